"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + per-phase SASS execution counts.
  python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--sass] [--traffic WORKLOAD KERNEL]
--traffic merges the capture's DRAM bytes per launch into profiles/ncu_traffic.json (bench.py's roofline.traffic)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__grid_size',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput', 'sm__throughput.avg.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']
for i, h in enumerate(hdr):
    if any(w in h for w in want) and 'Not Issued' not in h and '.max.' not in h and '.min.' not in h and 'per_second' not in h.replace('dram__bytes', ''):
        vals = [r[i] for r in data]
        if 'stalled' in h and all(float(v or 0) < 0.3 for v in vals):
            continue
        print(f"{h} [{units[i]}] {vals}")
if '--sass' in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    iS = h.index("Source"); iI = h.index("Instructions Executed"); iSm = h.index("# Samples"); iW = h.index("L1 Wavefronts Shared"); iWi = h.index("L1 Wavefronts Shared Ideal")
    out = []
    for r in rows[2:]:
        if len(r) < len(h): break
        out.append((r[iS].strip(), int(r[iI]), int(r[iSm]), int(r[iW]), int(r[iWi])))
    tot = sum(o[1] for o in out); samp = sum(o[2] for o in out)
    print("SASS instructions", len(out), "executed", tot, "samples", samp)
    prev = 0
    for i, o in enumerate(out):
        if 'BAR.SYNC' in o[0] or 'EXIT' in o[0] or 'TRYWAIT' in o[0] or i == len(out) - 1:
            seg = out[prev:i + 1]
            n = sum(x[1] for x in seg); s = sum(x[2] for x in seg); w = sum(x[3] for x in seg); wi = sum(x[4] for x in seg)
            if n: print(f"  [{prev:4d}:{i + 1:4d}] -> {o[0][:34]:34s} exec {n:10d} ({100 * n / tot:5.1f}%) samples {100 * s / samp:5.1f}% smemWF {w:9d} (ideal {wi:9d})")
            prev = i + 1

if '--traffic' in sys.argv:
    import json, os
    i = sys.argv.index('--traffic'); wl, kn = sys.argv[i + 1], sys.argv[i + 2]
    def col(name):
        j = next(k for k, h in enumerate(hdr) if h == name)
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[units[j]]
        return sum(float(r[j]) for r in data) / len(data) * scale
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    t.setdefault(wl, {})[kn] = {"dram_read_bytes": col('dram__bytes_read.sum'), "dram_write_bytes": col('dram__bytes_write.sum'),
                                "launches_averaged": len(data), "source": os.path.basename(rep)}
    json.dump(t, open(path, "w"), indent=1)
    print("wrote", path, t[wl][kn])
