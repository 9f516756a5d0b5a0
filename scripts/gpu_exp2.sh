for cfg in "grid139 0 0" "grid139 2 0" "grid139 1 0" "grid139 0 2" "grid55 0 0"; do timeout 300 python scripts/phase_profile.py $cfg 2>&1 | tail -9; done
