"""Shared-memory / L1 data-pipe wavefronts of the local kernel per tile, MODELLED on the host from the layout the library
builds (no GPU): every warp-wide access of k_local is replayed with its real addresses (staging slots, H-scratch entries,
incidence rows) under the usual bank model -- 32 banks of 4 bytes, a 128-bit access in four quarter-warp passes, a pass or a
32-bit access needs as many wavefronts as the largest number of DISTINCT addresses falling into one bank (group).  The
default layout's total can be compared with ncu (l1tex__data_pipe_lsu_wavefronts_mem_shared per launch / tiles; DESIGN.md
section 7 quotes 707 per tile for grid139, which also contains the asynchronous copies' landing writes); the variants' totals
are the prediction the experiments of DESIGN.md section 9 are to be measured against.

    python scripts/wavefront_model.py [cells=40]                         # default layout, also evaluated with pads predicated off
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pd = importlib.import_module("soft-body-simulation-cuda_b200")
PLANES = False      # (the x|y|z-plane H scratch was measured on a B200 in round 2 and removed: no gain, profiles/r2_ab_planes_pred.txt)
ZERO_OFF = 4 * 256 * (4 if PLANES else 16)


def qw(n):          # quarter-warp passes of a 128-bit access with n leading active lanes in a warp
    return (n + 7) // 8


def model(L, sample=1):
    rec, T = L.records, L.tile_table
    tot = dict(records_ldg=0, positions_lds=0, positions_ideal=0, h_sts=0, rows_lds=0, gather_lds=0, gather_pred=0, slots_stg=0, stage_ldgsts_ideal=0,
               partc_landing=0, misc_ldg=0)
    ntiles = 0
    for ti in range(0, L.num_tiles, sample):
        ntiles += 1
        off = int(T[ti, 0]) * 16; ab = int(T[ti, 1] & 0xffff); cb = int(T[ti, 1] >> 16)
        nT = int(T[ti, 2] & 0xffff)
        tr = rec[off + 96:off + 96 + 48 * nT].view(np.uint32).reshape(3, nT, 4).transpose(1, 0, 2).reshape(nT, 12)
        halves = np.stack([tr[:, 10] & 0xffff, tr[:, 10] >> 16, tr[:, 11] & 0xffff, tr[:, 11] >> 16], 1).astype(np.int64)
        stage = (halves >> 4) & 0xff
        for w0 in range(0, nT, 32):
            act = min(32, nT - w0)
            tot["records_ldg"] += 3 * qw(act)
            tot["h_sts"] += 12 if PLANES else 4 * qw(act)
            for k in range(4):
                for q0 in range(w0, w0 + act, 8):
                    sl = np.unique(stage[q0:min(q0 + 8, nT), k])
                    tot["positions_lds"] += int(np.bincount(sl % 8, minlength=8).max())
                    tot["positions_ideal"] += 1
        nRows = cb // 128
        incT = rec[off + ab:off + ab + cb].view(np.uint16).reshape(nRows, 32, 2).astype(np.int64)
        real = incT < ZERO_OFF
        for g in range(8):
            w = int(T[ti, 4 + g]); rb = w & 63; nr = (w >> 6) & 63; nvalid = (w >> 12) & 63
            if nr == 0:
                continue
            trips = (nr + 1) // 2                                    # the default-mode loop takes two rows per trip (an odd last row is paired with zeros)
            tot["rows_lds"] += nr
            tot["gather_lds"] += trips * 2 * 2 * (3 if PLANES else 4)
            r = real[rb:rb + nr]                                     # (row, lane, half)
            tot["gather_pred"] += int(r.reshape(nr, 4, 8, 2).any(axis=2).sum())     # quarter-warps with a real entry, per row and half
            tot["slots_stg"] += qw(nvalid)
        nstage = int((L.vstage[256 * ti:256 * ti + 256] != 0xffffffff).sum())
        tot["stage_ldgsts_ideal"] += qw(nstage) if nstage <= 32 else sum(qw(min(32, nstage - s)) for s in range(0, nstage, 32))
        tot["partc_landing"] += cb // 128
        tot["misc_ldg"] += 8 + 8                                      # vstage word per thread, tile-table words per warp
    return {k: v / ntiles for k, v in tot.items()}, ntiles


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    sc = pd.Scene.kuhn_grid(n, n, n, 1.0, 0.05, 12345, (0, 10, 0), 1.0, 2e5)
    L = sc.layout()
    m, nt = model(L, sample=max(1, L.num_tiles // 400))
    shared = m["positions_lds"] + m["h_sts"] + m["rows_lds"] + m["gather_lds"] + m["stage_ldgsts_ideal"] + m["partc_landing"]
    glob = m["records_ldg"] + m["slots_stg"] + m["misc_ldg"]
    name = "default layout"
    print(f"{name}, grid {n}^3, {nt} tiles sampled: wavefronts per tile")
    for k, v in m.items():
        print(f"  {k:22s} {v:8.1f}")
    print(f"  shared-memory total     {shared:8.1f}   (LDGSTS landing at its ideal; ncu shows it 2.5x that)")
    print(f"  + global (LDG/STG)      {glob:8.1f}   = {shared + glob:.1f} through the L1 data pipe")
    if not PLANES:
        pred = shared - m["gather_lds"] + m["gather_pred"]
        print(f"  with -DPD_PHASEC_PRED=1 {pred:8.1f}   shared ({m['gather_lds'] - m['gather_pred']:.1f} fewer: {100 * (m['gather_lds'] - m['gather_pred']) / (shared + glob):.1f} % of the pipe's load)")
