"""Mesh-mesh collision pass on the GPU (PdSolver::Update with SolverParams::handleCollision, pdSolver.cu:218-231):
the engine's pass against the reference's OWN continuous-collision arithmetic (collision/intersections.cu and
simulation/collisionUtil.cu compiled verbatim, oracle/ref_collision.cu) and against the CPU oracle's restatement of the
whole DetectCollision + CCDKernel sequence (oracle/pd_oracle.c:mesh_collision, all triangle pairs instead of the LBVH)."""
import os

import numpy as np
import pytest

import meshes
from test_collision_host import oracle_of, two_blocks

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
HAVE_REFC = os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libpd_ref_collision.so"))


def _random_queries(rng, n):
    """vertex-face and edge-edge queries around a contact within the step: a moving primitive crossing (or just missing) a
    second one, plus degenerate ones (parallel edges, coplanar starts, no motion)"""
    X = np.zeros((4 * n, 3), np.float32); XT = np.zeros_like(X)
    types = np.zeros(n, np.int32); verts = np.arange(4 * n, dtype=np.uint32).reshape(n, 4)
    for i in range(n):
        ee = i % 2 == 1
        types[i] = 2 if ee else 1
        if not ee:
            tri = rng.normal(size=(3, 3)); tri[:, 1] *= 0.1
            w = rng.dirichlet(np.ones(3)) if i % 7 else rng.normal(size=3)              # sometimes outside the triangle
            w = w / w.sum() if abs(w.sum()) > 1e-3 else np.ones(3) / 3
            hitp = w @ tri
            nrm = np.cross(tri[1] - tri[0], tri[2] - tri[0]); nrm /= np.linalg.norm(nrm) + 1e-12
            t = rng.uniform(-0.2, 1.3)
            d = rng.uniform(0.05, 1.0) * (1 if rng.random() < 0.5 else -1)
            p0 = hitp + nrm * d * t + 0.0 * rng.normal(size=3)
            p1 = hitp - nrm * d * (1 - t)
            x = np.vstack([p0, tri]); xt = np.vstack([p1, tri + 0.05 * rng.normal(size=(3, 3))])
        else:
            a = rng.normal(size=(2, 3)); b = rng.normal(size=(2, 3))
            mid = 0.5 * (a[0] + a[1])
            b += mid - 0.5 * (b[0] + b[1])                                               # edges cross near their middles
            nrm = np.cross(a[1] - a[0], b[1] - b[0]); nrm /= np.linalg.norm(nrm) + 1e-12
            t = rng.uniform(-0.2, 1.3)
            d = rng.uniform(0.05, 1.0)
            x = np.vstack([a + nrm * d * t, b]); xt = np.vstack([a - nrm * d * (1 - t), b + 0.05 * rng.normal(size=(2, 3))])
            if i % 11 == 1:
                x[3] = x[2] + (x[1] - x[0]); xt[3] = xt[2] + (xt[1] - xt[0])             # parallel edges
        if i % 13 == 0:
            xt = x.copy()                                                                 # nothing moves
        X[4 * i:4 * i + 4] = x; XT[4 * i:4 * i + 4] = xt
    return types, verts, X, XT


@pytest.mark.skipif(not HAVE_REFC, reason="reference collision harness (oracle/_ref/libpd_ref_collision.so) not built")
def test_ccd_arithmetic_vs_reference_and_oracle(pd, O):
    """ccd::collision_test (csrc/pd_collision.cuh) == the reference's ccdCollisionTest<float> on 20,000 queries: the same hit /
    miss decision and time of impact; the oracle's C restatement likewise (it rounds every operation separately, the two nvcc
    builds contract the same expressions into FMAs: decisions may differ on a handful of borderline roots)."""
    import ref
    rng = np.random.default_rng(5)
    types, verts, X, XT = _random_queries(rng, 20000)
    toi_r, nor_r = ref.ccd_queries(types, verts, X, XT)
    toi_e, nor_e = pd.ccd_batch(types, verts, X, XT)
    hit_r, hit_e = toi_r < 1.0, toi_e < 1.0
    print(f"ccd queries: {hit_r.sum()} of {types.shape[0]} hit in the reference ({(hit_r & (types == 1)).sum()} VF, {(hit_r & (types == 2)).sum()} EE); "
          f"engine decides differently on {(hit_r != hit_e).sum()}")
    assert hit_r.sum() > 2000 and (~hit_r).sum() > 2000
    assert (hit_r != hit_e).sum() <= 2
    both = hit_r & hit_e
    assert np.abs(toi_r[both] - toi_e[both]).max() <= 1e-5
    assert np.abs(nor_r[both] - nor_e[both]).max() <= 1e-4
    bit = np.array_equal(toi_r.view(np.uint32), toi_e.view(np.uint32))
    print(f"   toi bit-identical to the reference: {bit}; max |toi diff| {np.abs(toi_r[both] - toi_e[both]).max():.2e}; max normal diff {np.abs(nor_r[both] - nor_e[both]).max():.2e}")
    toi_o = np.array([O.ccd_test(types[i] == 2, verts[i], X, XT)[0] for i in range(0, types.shape[0], 4)], np.float32)
    hit_o = toi_o < 1.0
    print(f"   oracle vs reference on every 4th query: {(hit_o != hit_r[::4]).sum()} decisions differ")
    assert (hit_o != hit_r[::4]).sum() <= 5
    m = hit_o & hit_r[::4]
    assert np.abs(toi_o[m] - toi_r[::4][m]).max() <= 1e-4


@pytest.mark.skipif(not HAVE_REFC, reason="reference collision harness not built")
def test_oracle_ccd_kernel_vs_reference(O):
    """CCDKernel<float> (collisionUtil.cu:49-70) verbatim vs the oracle's restatement inside mesh_collision: the same X and V."""
    import ref
    rng = np.random.default_rng(9)
    nV = 5000
    X = rng.normal(size=(nV, 3)).astype(np.float32); XT = (X + 0.1 * rng.normal(size=(nV, 3))).astype(np.float32)
    V = rng.normal(size=(nV, 3)).astype(np.float32)
    tI = np.where(rng.random(nV) < 0.3, np.float32(0.5), np.float32(1.0)).astype(np.float32)
    n = rng.normal(size=(nV, 3)); n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    Xr, Vr = ref.ccd_kernel(X, XT, V, tI, n)
    hit = tI < 1
    dx = (XT - X)[hit]
    vn = (dx * n[hit]).sum(1, keepdims=True) * n[hit]
    assert np.array_equal(Xr[hit], X[hit]) and np.array_equal(Xr[~hit], XT[~hit]) and np.array_equal(Vr[~hit], V[~hit])
    assert np.abs(Vr[hit] + vn).max() <= 1e-6


def test_collision_pass_vs_oracle(pd, O):
    """Two blocks, the upper one driven into the lower one: tI, the contact normals, X, V and XTilde after every step against the
    oracle's all-pairs restatement.  Same overlapping pairs (the engine's refitted trees vs none at all), same vertices in
    contact; positions to the BASELINE tolerance."""
    sc, X0, V0 = two_blocks(pd, cells=3)
    p = sc.params
    eng = pd.PdSolver(sc)
    eng.upload(V=V0)
    osc = oracle_of(O, sc, V0)
    op = O.make_params(dt=p["dt"], gravity=p["gravity"], num_iterations=p["num_iterations"], threads=8)
    scale = float(np.linalg.norm(X0.max(0) - X0.min(0)))
    seen = 0
    for step in range(6):
        eng.Update(1); osc.step(op, 1)
        tI, nor, pairs = eng.collision()
        tIo, noro, pairso = osc.collision()
        X, V, XT = eng.download(); Xo, Vo, XTo = osc.get()
        hit, hito = tI < 1, tIo < 1
        seen += int(hito.sum())
        e = max(meshes.rel_err(X, Xo, scale), meshes.rel_err(XT, XTo, scale))
        print(f"collision step {step + 1}: pairs {pairs} (oracle {pairso}), vertices in contact {hit.sum()} (oracle {hito.sum()}), "
              f"differ {int((hit != hito).sum())}; rel err X/XTilde {e:.2e}; max |dV| {np.abs(V - Vo).max():.2e}")
        assert pairs == pairso
        assert np.array_equal(hit, hito)
        both = hit & hito
        if both.any():
            assert np.abs(nor[both] - noro[both]).max() <= 1e-4
        assert e <= 1e-4 and np.abs(V - Vo).max() <= 1e-3 * max(1.0, np.abs(Vo).max())
    assert seen > 0
    names, raw = eng.GetPerformanceData()
    eng.SetPerf(True); eng.Update(1)
    names, raw = eng.GetPerformanceData()
    assert names[3][0] == "collision handling(mesh)" and names[3][1] > 0 and names[2][1] > 0


def test_collision_flag_changes_nothing_without_contact(pd):
    """handleCollision = true (the reference's default) on bodies that never come near each other: bit-identical to the pass
    being off, through the plain-launch path the flag selects (no CUDA graph, tile kernels)."""
    sc, X0, V0 = two_blocks(pd, gap=3.0, speed=0.0)
    on = pd.PdSolver(sc, body_kernel=0)
    q = sc.params; q["handle_collision"] = 0
    sc.params = q
    off = pd.PdSolver(sc, body_kernel=0)
    on.Update(3); off.Update(3)
    for u, w in zip(on.download(), off.download()):
        assert np.array_equal(u.view(np.uint32), w.view(np.uint32))
    tI, nor, pairs = on.collision()
    assert (tI == 1).all() and pairs == 0
    with pytest.raises(pd.PdError):
        pd.PdSolver(sc_with_collision(pd), rank=0, world=2)


def sc_with_collision(pd):
    sc, _, _ = two_blocks(pd)
    return sc


def test_shipped_two_body_context_with_collision(pd, O, assets):
    """The shipped house + sphere context with the flag on, as the GUI runs it by default (handleCollision = true, context.h:44).
    The context is chaotic long before the bodies meet (the reference's own two runs are 5e-3 = half a length unit apart after
    40 steps, profiles/r1_noise_floor.txt), so the bar is: the sphere reaches the house within two steps of when it does in the
    oracle, contacts are found, and the state stays finite."""
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    p = sc.params
    p["handle_collision"] = 1; p["num_iterations"] = 30
    sc.params = p
    eng = pd.PdSolver(sc, rot_mode=1, reorder=0)
    osc, _ = meshes.oracle_scene(O, assets, "C5 house&sphere")
    tri, fa = sc.surface()
    osc.set_collision(True, tri, fa)
    op = O.make_params(dt=p["dt"], gravity=p["gravity"], muN=p["muN"], muT=p["muT"], rho=p["rho"], num_iterations=30, threads=8)
    first = first_o = None
    hits = 0
    for step in range(70):
        eng.Update(1); osc.step(op, 1)
        tI, nor, pairs = eng.collision()
        tIo, noro, pairso = osc.collision()
        hits += int((tI < 1).sum())
        if first is None and pairs > 0:
            first = step + 1
        if first_o is None and pairso > 0:
            first_o = step + 1
        if first is None and first_o is None:
            assert pairs == pairso == 0
    X = eng.download()[0]
    print(f"house + sphere with mesh collision: swept boxes first overlap at step {first} (oracle {first_o}); {hits} vertex contacts over 70 steps; "
          f"last step {pairs} overlapping triangle pairs (oracle {pairso})")
    assert first is not None and first_o is not None and abs(first - first_o) <= 2
    assert hits > 0 and np.isfinite(X).all()
