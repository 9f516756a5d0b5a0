"""Host-only check of the PD_H_PLANES experiment (H scratch as x | y | z planes with a 32-colouring, layout.hpp): emulates the
local kernel's phase B stores and phase C gathers of every tile in numpy from the records the VARIANT library builds, with
the index arithmetic written exactly as csrc/pd_kernels.cuh does it (h_store / h_load / local_phase_c), and checks
  * every (tet, corner) contribution lands in its own scratch entry, every warp-wide STS.32 and every gathered LDS.32 of a
    row is free of bank conflicts (distinct entries fall into distinct banks; bank = index mod 32 in every plane),
  * every tile-local vertex's lane sums exactly the contributions of its incident (tet, corner) pairs, pads read zeros.
No GPU needed.  Build the variant first:
    PD_OUT=$PWD/soft-body-simulation-cuda_b200/variants/libpd_planes.so PD_DEFS="-DPD_H_PLANES=1" python soft-body-simulation-cuda_b200/build.py
    PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_planes.so python scripts/check_h_planes_layout.py"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
pd = importlib.import_module("soft-body-simulation-cuda_b200")

TILE_T, HSTRIDE, NZ = 256, 1024, 32                 # entries per corner block * 4 bytes; zero entries
PLANE_ENTRIES = 4 * TILE_T + NZ
ZERO_OFF = 4 * HSTRIDE


def check(sc, name):
    L = sc.layout()
    rec = L.records
    T = L.tile_table
    rng = np.random.default_rng(1)
    worst_rows = 0
    for ti in range(L.num_tiles):
        off = int(T[ti, 0]) * 16; ab = int(T[ti, 1] & 0xffff); cb = int(T[ti, 1] >> 16)
        nT = int(T[ti, 2] & 0xffff); nLocal = int(T[ti, 2] >> 16)
        tr = rec[off + 96:off + 96 + 48 * nT].view(np.uint32).reshape(3, nT, 4).transpose(1, 0, 2).reshape(nT, 12)
        halves = np.stack([tr[:, 10] & 0xffff, tr[:, 10] >> 16, tr[:, 11] & 0xffff, tr[:, 11] >> 16], 1).astype(np.int64)
        colour = ((halves >> 12) & 15) | ((halves & 1) << 4)                       # h_store
        stage = (halves >> 4) & 0xff
        assert ((halves & 0x000e) == 0).all()
        tl = np.arange(nT)[:, None]
        k = np.arange(4)[None, :]
        idx = k * TILE_T + (tl & ~31) + colour                                      # entry index inside a plane
        assert len(np.unique(idx)) == 4 * nT, (name, ti, "two contributions share a scratch entry")
        for w in range(0, nT, 32):                                                  # one warp-wide STS.32 per corner and plane
            for kk in range(4):
                c = colour[w:w + 32, kk]
                assert len(set(c.tolist())) == len(c), (name, ti, "store bank conflict")
        H = np.zeros(PLANE_ENTRIES)
        val = rng.integers(1, 1 << 20, size=(nT, 4)).astype(np.float64)
        H[idx.reshape(-1)] = val.reshape(-1)
        # local vertex of every corner: staging slot -> vstage -> position in the tile's vlist
        vl = L.vlist[256 * ti:256 * ti + 256]; vs = L.vstage[256 * ti:256 * ti + 256]
        rank_of = {int(v): i for i, v in enumerate(vl[:nLocal])}
        loc = np.vectorize(lambda s: rank_of[int(vs[s])])(stage)
        want = np.zeros(nLocal)
        np.add.at(want, loc.reshape(-1), val.reshape(-1))
        nRows = cb // 128
        worst_rows = max(worst_rows, nRows)
        incT = rec[off + ab:off + ab + cb].view(np.uint16).reshape(nRows, 32, 2).astype(np.int64)
        assert (incT % 4 == 0).all() and (incT < 4 * PLANE_ENTRIES).all()
        e = incT // 4
        for r in range(nRows):                                                      # one gathered LDS.32 per (row, half) and plane
            for h in range(2):
                a = e[r, :, h]
                ua = np.unique(a)
                assert len(set((ua % 32).tolist())) == len(ua), (name, ti, r, h, "load bank conflict")
        got = np.zeros(nLocal)
        for g in range(8):
            w = int(T[ti, 4 + g]); rb = w & 63; nr = (w >> 6) & 63; nvalid = (w >> 12) & 63
            if nr == 0:
                assert nvalid == 0 or all(want[32 * g + l] == 0 for l in range(nvalid))
                continue
            lanes = H[e[rb:rb + nr]].sum(axis=(0, 2))                               # local_phase_c: every row, both halves
            got[32 * g:32 * g + nvalid] = lanes[:nvalid]
            assert (e[rb:rb + nr, nvalid:, :] >= ZERO_OFF // 4).all()              # lanes without a vertex read zeros only
        assert np.array_equal(got, want), (name, ti, "a vertex sums the wrong contributions")
    print(f"{name}: {L.num_tiles} tiles ok (max rows {worst_rows})")


if __name__ == "__main__":
    if "planes" not in os.path.basename(pd.LIB_PATH):
        print("note: PD_B200_LIB does not point at the planes variant; this check is written for -DPD_H_PLANES=1", file=sys.stderr)
    import meshes
    with tempfile.TemporaryDirectory() as tmp:
        assets = meshes.write_assets(tmp)
        for ctx in ("C1 cube", "C5 house&sphere", "C2 armadillo&bunny"):
            check(pd.Scene.from_json(assets["json"], ctx), ctx)
    check(pd.Scene.kuhn_grid(14, 14, 14, 1.0, 0.05, 12345, (0, 10, 0), 1.0, 2e5), "grid14")
