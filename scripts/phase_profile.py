"""Per-phase clock profile of the local kernel (warp 0 of every CTA), to see where a tile's latency goes.
  python scripts/phase_profile.py [workload] [ctas_per_sm] [rot_mode]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
workload = sys.argv[1] if len(sys.argv) > 1 else "grid139"
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rot = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pd = importlib.import_module("soft-body-simulation-cuda_b200")
sc, p = bench.make_scene(pd, workload)
eng = pd.PdSolver(sc, use_graph=0, ctas_per_sm=ctas, rot_mode=rot)
eng.upload(V=bench.initial_velocity(sc.arrays()["X"]))
eng.Update(1)
a = eng.profile_local().astype(np.float64)
tiles = a[:, 7].sum()
names = ["phase B (positions, math, H stores)", "record LDG issue + gather wait", "barrier", "producer + vstage load issue", "wait part C (TMA)", "phase C", "gather issue (LDGSTS)"]
tot = a[:, :7].sum()
print(f"{workload} ctas/SM {ctas} rot {rot}: grid {a.shape[0]}, tiles {int(tiles)}, cycles per tile per CTA {tot / tiles:.0f}")
for i, n in enumerate(names):
    print(f"  {n:28s} {a[:, i].sum() / tiles:8.0f} cycles/tile  {100 * a[:, i].sum() / tot:5.1f}%")
