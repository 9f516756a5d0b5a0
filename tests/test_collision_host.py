"""Host side of the mesh-mesh collision pass, no GPU: the scene's surface triangles (dataLoader.cu:68-127,
343-369) and the CPU oracle's restatement of DetectCollision + CCDKernel (pdSolver.cu:218-225)."""
import numpy as np
import pytest

import meshes


def _boundary_faces_numpy(Tet):
    """independent restatement: faces that belong to exactly one tet, keyed by their sorted triple, in key order, wound like
    the tet face they come from"""
    F = [(0, 1, 2), (0, 2, 3), (0, 3, 1), (1, 3, 2)]
    faces = np.concatenate([Tet[:, f] for f in F], axis=0).astype(np.int64)
    key = np.sort(faces, axis=1)
    uniq, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    out = []
    for k in np.nonzero(cnt == 1)[0]:
        out.append(faces[np.nonzero(inv.reshape(-1) == k)[0][0]])
    return np.array(out, np.uint32)


def test_surface_of_a_scene(pd, assets):
    sc = pd.Scene.from_json(assets["json"], "C1 cube")
    tri, fa = sc.surface()
    assert tri.shape == (12, 3) and not fa.any()
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    a = sc.arrays()
    tri, fa = sc.surface()
    starts = list(a["body_vert_start"]) + [a["X"].shape[0]]
    tets_of = [np.all((a["Tet"] >= starts[b]) & (a["Tet"] < starts[b + 1]), axis=1) for b in range(2)]
    want = [_boundary_faces_numpy(a["Tet"][m]) for m in tets_of]
    assert np.array_equal(fa, np.repeat(np.arange(2, dtype=np.uint32), [w.shape[0] for w in want]))
    assert np.array_equal(tri, np.concatenate(want))
    # closed surfaces: every edge is shared by exactly two triangles of its body
    for b in range(2):
        t = tri[fa == b].astype(np.int64)
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [0, 2]]]), axis=1)
        assert (np.unique(e, axis=0, return_counts=True)[1] == 2).all()
    # caller-given triangles are kept as they are; a merge shifts them with the vertices
    g = pd.Scene.kuhn_grid(2, 2, 2, 1.0, 0.0, 1, (0, 0, 0), 1.0, 2e5)
    ga = g.arrays()
    gt, _ = g.surface()
    assert gt.shape[0] == 6 * 2 * 2 * 2
    given = pd.Scene.from_arrays(ga["X"], ga["Tet"], ga["mass"], ga["mu"], Tri=gt[::-1], TriFathers=np.zeros(gt.shape[0], np.uint32), params=g.params)
    assert np.array_equal(given.surface()[0], gt[::-1])
    m = pd.Scene.merge([g, given])
    mt, mf = m.surface()
    assert np.array_equal(mt, np.concatenate([gt, gt[::-1] + ga["X"].shape[0]])) and np.array_equal(mf, np.repeat([0, 1], gt.shape[0]).astype(np.uint32))


def two_blocks(pd, gap=0.03, speed=4.0, cells=2, iters=20):
    """two jittered Kuhn blocks, the upper one moving down onto the lower one: (merged scene, X0, V0)"""
    lo = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, 0.04, 5, (0.0, 1.0, 0.0), 1.0, 2e5)
    up = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, 0.04, 6, (0.3, 1.0 + cells + 0.08 + gap, 0.2), 1.0, 2e5)
    p = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=iters, handle_collision=1)
    lo.params = p; up.params = p
    sc = pd.Scene.merge([lo, up])
    sc.params = p
    X0 = sc.arrays()["X"]
    V0 = np.zeros_like(X0)
    V0[lo.counts()[0]:, 1] = -speed
    return sc, X0, V0


def oracle_of(O, sc, V0, collide=True):
    a = sc.arrays()
    osc = O.Scene(a["X"], a["Tet"], a["mass"], a["mu"])
    osc.set(V=V0)
    tri, fa = sc.surface()
    osc.set_collision(collide, tri, fa)
    return osc


def test_oracle_collision_pass(pd, O):
    sc, X0, V0 = two_blocks(pd)
    p = sc.params
    op = O.make_params(dt=p["dt"], gravity=p["gravity"], num_iterations=p["num_iterations"], threads=4)
    osc = oracle_of(O, sc, V0)
    free = oracle_of(O, sc, V0, collide=False)
    hit_any = False
    for step in range(4):
        Xb = osc.get()[0].copy()
        osc.step(op, 1); free.step(op, 1)
        X, V, XT = osc.get()
        tI, nor, pairs = osc.collision()
        hit = tI < 1.0
        if not hit_any and not hit.any():
            # until the first contact the pass changes nothing
            for u, w in zip(osc.get(), free.get()):
                assert np.array_equal(u, w)
        if hit.any():
            hit_any = True
            assert pairs > 0 and set(np.unique(tI)) <= {0.5, 1.0}
            assert np.array_equal(X[hit], Xb[hit])                          # CCDKernel: a vertex in contact keeps its position ...
            assert np.array_equal(X[~hit], XT[~hit])                        # ... every other one X <- XTilde
            n = nor[hit]
            assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
            dx = (XT - Xb)[hit]
            assert np.allclose(V[hit], -(dx * n).sum(1, keepdims=True) * n, atol=1e-6)      # V = -(n . dx) n
    assert hit_any
