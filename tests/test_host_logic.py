"""CPU tests (no GPU): C-ABI library loads and exports every declared symbol, the scene loader
matches the oracle's loader, and the device layout is bit-exact against the numpy restatement."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import layout_oracle as LO
import meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_exports_every_declared_symbol(pd):
    hdr = open(os.path.join(ROOT, "include", "pd_b200.h")).read()
    declared = set(re.findall(r"\b(pd_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pd_status"}
    L = ctypes.CDLL(pd.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    # the python binding covers exactly the header
    assert declared == set(pd.SYMBOLS), declared ^ set(pd.SYMBOLS)
    assert b"sm_100a" in pd.lib().pd_version()


def test_error_convention(pd):
    assert pd.lib().pd_scene_load_json(b"/nonexistent/context.json", None, None) is None
    assert "open" in pd.lib().pd_last_error().decode().lower()
    assert pd.lib().pd_step(None, 1) < 0
    with pytest.raises(pd.PdError):
        pd.Scene.from_arrays(np.zeros((4, 3), np.float32), np.array([[0, 1, 2, 9]], np.uint32), 1.0, 1.0)


def test_json_loader_matches_oracle_loader(pd, O, assets):
    for ctx in ["C1 cube", "C2 armadillo&bunny", "Armadillo&house", "C5 house&sphere"]:
        sc = pd.Scene.from_json(assets["json"], ctx)
        a = sc.arrays()
        osc, op = meshes.oracle_scene(O, assets, ctx)
        assert a["X"].shape == osc.X0.shape
        assert np.array_equal(a["X"].view(np.uint32), osc.X0.view(np.uint32)), ctx      # bit-exact transform
        assert np.array_equal(a["Tet"], osc.Tet)
        p = sc.params
        assert p["num_iterations"] == 100 and abs(p["dt"] - op["dt"]) < 1e-9 and p["gravity"] == op["gravity"]
    # context selection rules (context.cpp:366-369): first loadable when unnamed, explicit name otherwise
    assert pd.Scene.from_json(assets["json"]).counts()[:2] == (8, 6)
    assert pd.Scene.from_json(assets["json"], "not loaded").counts()[:2] == (4, 1)
    with pytest.raises(pd.PdError):
        pd.Scene.from_json(assets["json"], "no such context")


def test_fixed_bodies_match_oracle(pd, O, assets):
    sc = pd.Scene.from_json(assets["json"], "Armadillo&house")
    fixed = sc.arrays()["fixed"]
    kinds = [f.type for f in fixed]
    assert kinds.count(pd.PD_CYLINDER) == 5 and kinds.count(pd.PD_PLANE) == 6
    cfg = json.load(open(assets["json"]))
    fdefs = {d["name"]: d for d in cfg["fixedBodies"]}
    ctx = [c for c in cfg["contexts"] if c["name"] == "Armadillo&house"][0]
    for f, fb in zip(fixed, ctx["fixedBodies"]):
        d = fdefs[fb["name"]]
        pos = fb.get("pos", d.get("pos", [0, 0, 0])); rot = fb.get("rot", d.get("rot", [0, 0, 0])); sc3 = fb.get("scale", d.get("scale", [1, 1, 1]))
        if d["type"] == "cylinder":
            sc3 = [sc3[0], sc3[1], sc3[0]]
            assert f.radius == sc3[0]
        M = O.model_matrix(pos, rot, sc3, 0)
        assert np.array_equal(np.array(f.model[:], np.float32).view(np.uint32), M.view(np.uint32))
        if d["type"] == "plane":
            up = np.zeros(3, np.float32)
            pd.lib().pd_plane_up(np.array(f.model[:], np.float32).ctypes.data, up.ctypes.data)
            assert np.array_equal(up.view(np.uint32), O.plane_up(M).view(np.uint32))


def test_centralize_swaps_y_z_and_fixes_orientation(pd, assets):
    X, E, _ = meshes.raw_mesh("house2")
    path = os.path.join(assets["assets"], "house2", "house2.node")
    p = ctypes.c_void_p(); n = ctypes.c_int()
    assert pd.lib().pd_load_node(path.encode(), 1, ctypes.byref(p), ctypes.byref(n)) == 0
    Xc = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n.value, 3)).copy()
    pd.lib().pd_free(p)
    T = E[:, 1:5] - 1
    vol = lambda P: np.einsum("ij,ij->i", np.cross(P[T[:, 1]] - P[T[:, 0]], P[T[:, 2]] - P[T[:, 0]]), P[T[:, 3]] - P[T[:, 0]])
    assert (vol(X.astype(np.float64)) < 0).all() and (vol(Xc.astype(np.float64)) > 0).all()
    assert abs(Xc.mean(0)).max() < 1e-3


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_fixture_equals_reference_assets(pd, O):
    """Pins tests/golden/meshes.npz: the product's loader on the reference's own files gives the fixture."""
    for name in ["cube", "house2", "sphere", "bunny", "armadillo0"]:
        X, E, idx0 = meshes.raw_mesh(name)
        node, ele = [str(s) for s in meshes.npz()[name + "_files"]]
        Xo = O.load_node(os.path.join(REF, "assets", node), False)
        assert np.array_equal(Xo.view(np.uint32), X.view(np.uint32)), name
        start = 0 if name == "armadillo0" else 1
        To = O.load_ele(os.path.join(REF, "assets", ele), start)
        assert np.array_equal(To, (E[:, 1:5] - start).astype(np.uint32))
        p = ctypes.c_void_p(); n = ctypes.c_int()
        assert pd.lib().pd_load_node(os.path.join(REF, "assets", node).encode(), 0, ctypes.byref(p), ctypes.byref(n)) == 0
        Xp = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n.value, 3)).copy()
        pd.lib().pd_free(p)
        assert np.array_equal(Xp.view(np.uint32), X.view(np.uint32))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_loads_the_reference_context_json(pd):
    sc = pd.Scene.from_json(os.path.join(REF, "context.json"), "Armadillo&house")
    nV, nT, nF, nB = sc.counts()
    assert (nV, nT, nB) == (13054 + 400 + 482 + 8, 41960 + 1389 + 1217 + 6, 4) and nF == 11
    p = sc.params
    assert abs(p["dt"] - 0.01) < 1e-9 and p["gravity"] == 98 and p["num_iterations"] == 100


def _check_layout(pd, sc, reorder=True):
    a = sc.arrays()
    Lc = sc.layout(reorder)
    Lo = LO.build(a["X"], a["Tet"], a["mu"], reorder)
    for k in ["tet_order", "vert_order", "tet_new", "tile_tet_start", "tile_rec_off", "vslot_ptr", "vslot", "vlist"]:
        assert np.array_equal(getattr(Lc, k), Lo[k]), k
    assert (Lc.num_tiles, Lc.num_slots, Lc.max_local) == (Lo["num_tiles"], Lo["num_slots"], Lo["max_local"])
    assert Lc.record_bytes == int(Lc.tile_rec_off[-1])
    wf, ideal = LO.decode_and_check_records(Lc.records, Lc.tile_rec_off, Lo["tiles"], Lc.vlist, Lc.vstage)    # incl. DmInv/w bits and the incidence CSR
    assert np.array_equal(Lc.tile_table, LO.tile_table(Lo["tiles"]))        # what the local kernel reads per tile
    assert wf <= 1.35 * ideal, (wf, ideal)      # staging-slot colouring: few bank conflicts left on the position loads (identity: ~1.8x)
    return Lc


def test_layout_bit_exact_small_meshes(pd, assets):
    for ctx in ["C1 cube", "C5 house&sphere", "not loaded"]:
        _check_layout(pd, pd.Scene.from_json(assets["json"], ctx))
    _check_layout(pd, pd.Scene.from_json(assets["json"], "C5 house&sphere"), reorder=False)


def test_layout_bit_exact_bunny_and_grid(pd, assets):
    X, E, _ = meshes.raw_mesh("bunny")
    sc = pd.Scene.from_arrays(X * np.float32(35), (E[:, 1:5] - 1).astype(np.uint32), 10.0, 2e6)
    L = _check_layout(pd, sc)
    assert L.num_tiles == -(-8417 // 256) and L.max_local <= 256
    g = pd.Scene.kuhn_grid(7, 6, 5, 1.0, 0.05, 12345, (0, 10, 0), 1.0, 2e5)
    assert g.counts()[:2] == (8 * 7 * 6, 6 * 7 * 6 * 5)
    _check_layout(pd, g)


def test_layout_invariants(pd, assets):
    sc = pd.Scene.from_json(assets["json"], "C2 armadillo&bunny")
    nV, nT = sc.counts()[:2]
    L = sc.layout()
    assert sorted(L.tet_order.tolist()) == list(range(nT)) and sorted(L.vert_order.tolist()) == list(range(nV))
    # every (tile, local vertex) slot appears exactly once in the vertex->slot CSR, ascending per vertex
    used = np.flatnonzero(L.vlist != 0xffffffff)          # slots are padded per tile: tile * 256 + local vertex
    assert sorted(L.vslot.tolist()) == used.tolist() and len(used) == L.num_slots
    assert np.array_equal((L.vlist[L.vslot] & 0x7fffffff), np.repeat(np.arange(nV), np.diff(L.vslot_ptr.astype(np.int64))))
    for v in range(0, nV, 997):
        s = L.vslot[L.vslot_ptr[v]:L.vslot_ptr[v + 1]]
        assert (np.diff(s.astype(np.int64)) > 0).all() and len(s) >= 1
    # tiles hold <= 256 tets and <= 256 vertices; Morton order keeps them compact
    assert np.diff(L.tile_tet_start.astype(np.int64)).max() <= 256 and L.max_local <= 256
    assert L.num_slots < 1.2 * nT      # ~0.9 slots per tet on armadillo+bunny (compactness regression guard)


def test_kuhn_grid_generator(pd):
    g = pd.Scene.kuhn_grid(3, 4, 2, 1.5, 0.0, 1, (1, 2, 3), 2.0, 1e5)
    a = g.arrays()
    X, T = a["X"].astype(np.float64), a["Tet"].astype(np.int64)
    vol = np.einsum("ij,ij->i", np.cross(X[T[:, 1]] - X[T[:, 0]], X[T[:, 2]] - X[T[:, 0]]), X[T[:, 3]] - X[T[:, 0]]) / 6
    assert (vol > 0).all() and np.isclose(vol.sum(), 3 * 4 * 2 * 1.5 ** 3)
    assert np.allclose(X.min(0), [1, 2, 3]) and np.allclose(X.max(0), [1 + 4.5, 2 + 6, 3 + 3])
    # jitter stream = mt19937(seed) 24-bit uniforms, reproducible with numpy's legacy seeding
    g2 = pd.Scene.kuhn_grid(2, 2, 2, 1.0, 0.05, 12345, (0, 0, 0), 1.0, 1.0).arrays()["X"]
    rs = np.random.RandomState(12345)
    raw = np.array([rs.randint(0, 2 ** 32, dtype=np.uint64) for _ in range(81)], np.uint64)
    u = ((raw >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(27, 3)
    base = np.stack(np.meshgrid(np.arange(3), np.arange(3), np.arange(3), indexing="ij"), -1).transpose(2, 1, 0, 3).reshape(27, 3).astype(np.float32)
    exp = base + np.float32(0.05) * (np.float32(2.0) * u - np.float32(1.0))
    assert np.array_equal(g2.view(np.uint32), exp.astype(np.float32).view(np.uint32))


def test_partition_vertices(pd):
    for nV, w in [(10, 3), (2744000, 8), (7, 8), (1, 1)]:
        vb = np.zeros(w + 1, np.int32)
        assert pd.lib().pd_partition_vertices(nV, w, vb.ctypes.data) == 0
        assert np.array_equal(vb, LO.partition_vertices(nV, w))
        assert vb[0] == 0 and vb[-1] == nV and (np.diff(vb) >= 0).all()


def test_sparse_cholesky_prefactor_host(pd, O, assets):
    """cholesky_factor (host side of the small-mesh direct path) on the oracle's system matrix: L L^T == A^."""
    import scipy.sparse as sp
    osc, _ = meshes.oracle_scene(O, assets, "C5 house&sphere")
    rp, col, val = osc.system_matrix(O.make_params(dt=1 / 60, gravity=9.8, num_iterations=1))
    n = rp.shape[0] - 1
    A = sp.csr_matrix((val.astype(np.float64), col, rp), shape=(n, n))
    assert abs(A - A.T).max() <= 1e-6 * abs(A).max()          # symmetric up to summation-order rounding
    lp, lc, lv = pd.cholesky_factor(rp, col, val)
    L = sp.csr_matrix((lv.astype(np.float64), lc, lp), shape=(n, n))
    assert (lc[lp[1:] - 1] == np.arange(n)).all() and (np.diff(lp) >= 1).all()      # diagonal last in every row
    assert sp.triu(L, 1).nnz == 0 and (L.diagonal() > 0).all()
    R = (L @ L.T - A)
    assert abs(R).max() <= 2e-6 * abs(A).max()
    with pytest.raises(pd.PdError):      # not positive definite -> error, nothing returned
        pd.cholesky_factor(np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32), np.array([1.0, -1.0], np.float32))


def test_nested_dissection_order_host(pd, O):
    """The fill-reducing order of the Cholesky path (layout.cpp:nested_dissection_order): a permutation, deterministic, and
    it cuts nnz(L) of a grid's system matrix several-fold against the given (here: lexicographic) order."""
    sc = pd.Scene.kuhn_grid(14, 14, 14, 1.0, 0.05, 3, (0, 0, 0), 1.0, 2e5)
    a = sc.arrays()
    osc = O.Scene(a["X"], a["Tet"], a["mass"], a["mu"])
    rp, col, val = osc.system_matrix(O.make_params(dt=1 / 60, gravity=9.8, num_iterations=1))
    n = rp.shape[0] - 1
    perm, nat, ordd = pd.nested_dissection(rp, col, a["X"])
    perm2, _, _ = pd.nested_dissection(rp, col, a["X"], count_fill=False)
    assert np.array_equal(np.sort(perm), np.arange(n)) and np.array_equal(perm, perm2)
    print(f"nested dissection, 14^3-cell grid ({n} rows, nnz(A) {col.shape[0]}): nnz(L) {nat} -> {ordd}")
    assert ordd < 0.7 * nat                  # (banded order: n x bandwidth; the gap widens with n: 3x at 24^3 cells)
    # the factor of the permuted matrix solves the original system
    import scipy.sparse as sp
    A = sp.csr_matrix((val.astype(np.float64), col, rp), shape=(n, n))
    B = A[perm][:, perm].tocsr(); B.sort_indices()
    lp, lc, lv = pd.cholesky_factor(B.indptr, B.indices, B.data.astype(np.float32))
    L = sp.csr_matrix((lv.astype(np.float64), lc, lp), shape=(n, n))
    assert lc.shape[0] == ordd and abs(L @ L.T - B).max() <= 2e-6 * abs(B).max()


def _fan(n_tets, hub_first=True):
    """n_tets tets that all share vertex 0 (a hub of valence n_tets) and, pairwise, an edge: a closed fan of thin
    wedges around the z axis.  The hub's incidence list is far longer than one tile's row budget allows."""
    n = n_tets
    ang = 2 * np.pi * np.arange(n) / n
    ring_lo = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], 1)
    ring_hi = ring_lo + np.array([0, 0, 1.0])
    X = np.concatenate([[[0, 0, 0.5]], ring_lo, ring_hi]).astype(np.float32)
    lo = 1 + np.arange(n); hi = 1 + n + np.arange(n)
    T = np.stack([np.zeros(n, np.int64), lo, np.roll(lo, -1), hi], 1).astype(np.uint32)
    return X, T


def test_layout_edge_cases(pd):
    """Ragged and extreme inputs of the tile builder, each bit for bit against the numpy restatement: one tet;
    vertices no tet touches (renumbered last, no slots, nothing staged); 256 / 257 tets (tile boundary); a hub vertex of
    valence 700 (its list alone exceeds a tile's 56 incidence rows, so tiles must be halved until it fits)."""
    one = pd.Scene.from_arrays(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32), np.array([[0, 1, 2, 3]], np.uint32), 1.0, 1e5)
    L = _check_layout(pd, one)
    assert (L.num_tiles, L.num_slots) == (1, 4)
    # isolated vertices: ids 1 and 5 are not referenced
    X = np.random.default_rng(5).normal(size=(7, 3)).astype(np.float32)
    iso = pd.Scene.from_arrays(X, np.array([[0, 2, 3, 4], [2, 3, 4, 6]], np.uint32), 1.0, 1e5)
    L = _check_layout(pd, iso)
    assert sorted(L.vert_order[-2:].tolist()) == [1, 5]
    assert np.diff(L.vslot_ptr.astype(np.int64))[-2:].tolist() == [0, 0] and L.num_slots == 5
    assert int((L.vstage != 0xffffffff).sum()) == 5
    for n in (256, 257):
        g = pd.Scene.kuhn_grid(n, 1, 1, 1.0, 0.0, 1, (0, 0, 0), 1.0, 1e5)       # 6 n tets in a row of cells
        Xg, Tg = g.arrays()["X"], g.arrays()["Tet"][:n]
        sc = pd.Scene.from_arrays(Xg, Tg, 1.0, 1e5)
        L = _check_layout(pd, sc)
        assert L.num_tiles == (1 if n == 256 else 2)
    X, T = _fan(700)
    fan = pd.Scene.from_arrays(X, T, 1.0, 1e5)
    L = _check_layout(pd, fan)
    tt = np.diff(L.tile_tet_start.astype(np.int64))
    assert tt.max() <= 112 and tt.sum() == 700          # 2 list entries per row, 56 rows: at most 112 tets around the hub per tile
    hub = int(np.flatnonzero(L.vert_order == 0)[0])
    assert int(L.vslot_ptr[hub + 1] - L.vslot_ptr[hub]) == L.num_tiles      # the hub has a slot in every tile


def test_bad_meshes_are_rejected(pd):
    X = np.zeros((4, 3), np.float32)
    with pytest.raises(pd.PdError):
        pd.Scene.from_arrays(X, np.array([[0, 1, 2, 9]], np.uint32), 1.0, 1e5).layout()      # vertex index out of range
    for Xe, Te in [(np.zeros((3, 3), np.float32), np.zeros((0, 4), np.uint32)), (np.zeros((0, 3), np.float32), np.zeros((0, 4), np.uint32))]:
        with pytest.raises(pd.PdError):                                                          # empty meshes: an error code, not a crash
            pd.Scene.from_arrays(Xe, Te, 1.0, 1e5)


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path, pd):
    """The drop-in boundary is a C ABI: include/pd_b200.h must compile as C99 (no C++, no torch, no CUDA types), and a C
    program must link against the library and reach the host-only entry points (no GPU needed)."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "abi.c"
    src.write_text('#include "pd_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '  pd_params p; pd_default_params(&p);\n'
                   '  const float o[3] = {0.f, 1.f, 0.f};\n'
                   '  pd_scene* s = pd_scene_kuhn_grid(2, 2, 2, 1.0f, 0.0f, 1u, o, 1.0f, 1000.0f);\n'
                   '  int nv = 0, nt = 0; if (!s || pd_scene_counts(s, &nv, &nt, 0, 0) != PD_OK) return 1;\n'
                   '  pd_engine* e = pd_create(s, 0);            /* no GPU here: must fail loudly, not fall back */\n'
                   '  printf("%d %d %d %s|%s\\n", nv, nt, p.num_iterations, pd_version(), e ? "engine" : pd_last_error());\n'
                   '  if (e) pd_destroy(e);\n'
                   '  pd_scene_free(s); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(pd.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-l:" + os.path.basename(pd.LIB_PATH), "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    nv, nt, its = out.split()[:3]
    assert (int(nv), int(nt)) == (27, 48) and int(its) >= 1
    assert out.strip().endswith("|engine") or "CUDA" in out          # on a box without a GPU: "no CUDA device: ... no CPU fallback"


def test_gpu_scripts_parse(tmp_path):
    """scripts/*.py and scripts/*.sh only ever run on the GPU box: a syntax error there costs a gpurun call."""
    import glob
    import py_compile
    import subprocess
    for f in glob.glob(os.path.join(ROOT, "scripts", "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        py_compile.compile(f, doraise=True, cfile=str(tmp_path / (os.path.basename(f) + "c")))
    for f in glob.glob(os.path.join(ROOT, "scripts", "*.sh")):
        subprocess.check_call(["bash", "-n", f])


def test_no_ldgsts_with_uniform_register_offset():
    """ptxas 12.9 folds a uniform-register offset into an LDGSTS that also carries an L2 cache-hint descriptor and emits an
    encoding the B200 rejects (cudaErrorIllegalInstruction at run time; seen twice: DESIGN.md section 9, measured facts).  The
    local kernel keeps the destination opaque (pd_kernels.cuh:gather); this checks the built library's SASS."""
    import shutil
    import subprocess
    so = os.path.join(ROOT, "soft-body-simulation-cuda_b200", "libpd_b200.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(so) or not os.path.exists(cuobjdump):
        pytest.skip("library or cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, timeout=600).stdout
    lines = [l for l in sass.splitlines() if "LDGSTS" in l]
    assert lines, "no LDGSTS at all: the position gather is not asynchronous any more?"
    bad = [l.strip() for l in lines if "+UR" in l]
    assert not bad, bad[:3]


def test_fixed_body_helper_and_bench_scenes(pd):
    """pd.fixed_body() once handed ctypes the addresses of temporaries that were freed before the call (the benchmark's floor
    plane ended up at y = 450, upside down); and the scene builders of bench.py, no GPU needed."""
    import importlib
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    for _ in range(20):          # (the freed memory was reused by the very next temporary: every call was wrong, not one in many)
        fb = pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450))
        M = np.array(fb.model[:], np.float32)
        up = np.zeros(3, np.float32)
        pd.lib().pd_plane_up(M.ctypes.data, up.ctypes.data)
        assert np.allclose(M[12:15], 0.0) and np.allclose(up, [0.0, 1.0, 0.0]), (M[12:15], up)
    sc, p = bench.make_scene(pd, "grid24")
    planes, spheres, cyls = bench.fixed_arrays(pd, sc.arrays()["fixed"])
    assert len(planes) == 1 and not spheres and not cyls and np.allclose(planes[0][0], 0) and np.allclose(planes[0][1], [0, 1, 0])
    sc, p = bench.make_scene(pd, "armadillo")
    assert sc.counts()[:2] == (13054, 41960) and p["num_iterations"] == 100 and abs(p["dt"] - 0.01) < 1e-6 and p["gravity"] == 98.0
    assert len(bench.fixed_arrays(pd, sc.arrays()["fixed"])[0]) == 6          # floor + five walls
    cfg = bench.workload_config("armadillo", 13054, 41960, 100, p, False, 1)
    assert cfg["workload"] == "armadillo" and cfg["initial_velocity"] == "0"
