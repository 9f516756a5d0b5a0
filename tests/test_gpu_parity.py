"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle (oracle/pd_oracle.c), against the reference's own CUDA kernels (oracle/_ref, when
the prebuilt harness travelled with the snapshot) and against the committed golden fixtures.

Tolerance (BASELINE.json north_star): max vertex-position relative error <= 1e-4 after 100 steps,
relative error = max_v |x_v - ref_v| / max(|ref_v|, bounding-box diagonal of the rest shape)."""
import os

import numpy as np
import pytest

import meshes

pytestmark = pytest.mark.gpu
TOL = 1e-4
HERE = os.path.dirname(os.path.abspath(__file__))


def _params(pd, scene, **kw):
    p = scene.params
    for k, v in kw.items():
        p[k] = v
    scene.params = p
    return p


def _oracle_params(O, p, **kw):
    d = dict(dt=p["dt"], gravity=p["gravity"], muN=p["muN"], muT=p["muT"], rho=p["rho"], tol=p["tol"],
             num_iterations=p["num_iterations"], threads=8)
    d.update(kw)
    return O.make_params(**d)


def _rest_scale(X):
    return float(np.linalg.norm(X.max(0) - X.min(0)))


@pytest.mark.parametrize("rot_mode", [0, 1])
def test_c1_cube_100_steps_vs_oracle(pd, O, assets, rot_mode):
    sc = pd.Scene.from_json(assets["json"], "C1 cube")
    p = _params(pd, sc, dt=1 / 60)
    osc, _ = meshes.oracle_scene(O, assets, "C1 cube")
    op = _oracle_params(O, p)
    eng = pd.PdSolver(sc, rot_mode=rot_mode)
    scale = _rest_scale(osc.X0)
    worst = 0.0
    for n in range(10):
        eng.Update(10)
        osc.step(op, 10)
        X, V, XT = eng.download()
        Xo, Vo, XTo = osc.get()
        worst = max(worst, meshes.rel_err(X, Xo, scale), meshes.rel_err(XT, XTo, scale))
    assert worst <= TOL, worst
    assert XT[:, 1].min() > -1e-3           # resting on the floor plane after the impact at step ~42
    assert np.abs(V).max() < 5.0


def test_c5_house_sphere_vs_oracle(pd, O, assets):
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    p = sc.params
    osc, _ = meshes.oracle_scene(O, assets, "C5 house&sphere")
    op = _oracle_params(O, p)
    eng = pd.PdSolver(sc)
    eng.Update(30)
    osc.step(op, 30)
    X, V, XT = eng.download()
    Xo, Vo, XTo = osc.get()
    scale = _rest_scale(osc.X0)
    assert meshes.rel_err(X, Xo, scale) <= TOL and meshes.rel_err(XT, XTo, scale) <= TOL


def test_setup_products_vs_oracle(pd, O, assets):
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    osc, _ = meshes.oracle_scene(O, assets, "C5 house&sphere")
    eng = pd.PdSolver(sc)
    md, c, B, V0 = eng.setup()
    mdo, co, Bo, V0o = osc.setup(_oracle_params(O, sc.params))
    assert np.array_equal(B.view(np.uint32), Bo.view(np.uint32))      # same arithmetic, bit-exact
    assert np.array_equal(V0.view(np.uint32), V0o.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), co.view(np.uint32))
    assert np.allclose(md, mdo, rtol=2e-6, atol=0)                    # summation order differs (tet reordering)


def test_rotation_paths(pd, O):
    rng = np.random.default_rng(7)
    n = 20000
    F = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        k = i % 5
        eps = [0.01, 0.1, 0.4, 1.0, 0.1][k]
        A = q @ (np.eye(3) + eps * rng.normal(size=(3, 3)))
        if k == 4:
            A[:, 2] *= -1          # inverted tets
        F[i] = A
    F[0] = 0; F[1] = np.eye(3); F[2] = np.diag([1, 1, -1]); F[3] = np.diag([2, 2, 1e-6]); F[4] = np.diag([1, 0, 0])
    Rsvd, used1 = pd.rotation_batch(F, rot_mode=1)
    assert used1.sum() == 0
    Ror = np.stack([O.rotation(f) for f in F])
    assert np.array_equal(Rsvd.view(np.uint32), Ror.view(np.uint32))  # device SVD == oracle SVD, bit for bit
    Rf, used = pd.rotation_batch(F, rot_mode=0)
    det = np.linalg.det(F.astype(np.float64))
    assert used[det <= 0.02].sum() == 0                               # inverted / flat tets take the SVD path
    assert used[(det > 0.5)].mean() > 0.95
    # where the fast path ran, it agrees with the exact polar factor to float accuracy
    idx = np.nonzero(used)[0]
    u, s, vt = np.linalg.svd(F[idx].astype(np.float64))
    Rex = u @ vt
    assert np.abs(Rf[idx] - Rex).max() < 2e-6
    assert np.abs(np.einsum("nij,nkj->nik", Rf[idx], Rf[idx]) - np.eye(3)).max() < 2e-6
    # and is at least as close to it as the reference's 4-sweep SVD
    assert np.abs(Rf[idx] - Rex).max() <= np.abs(Ror[idx] - Rex).max() + 1e-6


def test_determinism_and_reset(pd, assets):
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    eng = pd.PdSolver(sc)
    eng.Update(5)
    a = eng.download()
    eng.Reset()
    x0 = eng.download()
    assert np.array_equal(x0[0], sc.arrays()["X"]) and not x0[1].any()
    eng.Update(5)
    b = eng.download()
    for u, v in zip(a, b):
        assert np.array_equal(u.view(np.uint32), v.view(np.uint32))   # gather assembly: run-to-run bit-identical
    eng2 = pd.PdSolver(sc, use_graph=0)
    eng2.Update(5)
    for u, v in zip(a, eng2.download()):
        assert np.array_equal(u.view(np.uint32), v.view(np.uint32))   # graph replay == plain launches


def test_reordering_does_not_change_results_beyond_rounding(pd, assets):
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    a = pd.PdSolver(sc, reorder=1); b = pd.PdSolver(sc, reorder=0)
    a.Update(10); b.Update(10)
    Xa, Xb = a.download()[0], b.download()[0]
    assert meshes.rel_err(Xa, Xb, _rest_scale(sc.arrays()["X"])) < 1e-5


def test_host_and_device_entry_points(pd, assets):
    import ctypes as C
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    nV = sc.counts()[0]
    ref = pd.PdSolver(sc); ref.Update(3)
    Xr, Vr, XTr = ref.download()
    # e2e on host buffers (pinned)
    eng = pd.PdSolver(sc)
    X0 = sc.arrays()["X"]
    nbytes = X0.nbytes
    bufs = [pd.lib().pd_alloc_pinned(nbytes) for _ in range(6)]
    try:
        arr = [np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=(nV, 3)) for b in bufs]
        arr[0][:] = X0; arr[1][:] = 0; arr[2][:] = X0
        for _ in range(3):
            eng.step_host_ptr(1, bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5])
            arr[0][:] = arr[3]; arr[1][:] = arr[4]; arr[2][:] = arr[5]
        assert np.array_equal(arr[3], Xr) and np.array_equal(arr[4], Vr) and np.array_equal(arr[5], XTr)
    finally:
        for b in bufs:
            pd.lib().pd_free_pinned(b)
    # adopt the reference's device arrays (torch used only to own device memory)
    import torch
    dX = torch.from_numpy(X0.copy()).cuda(); dV = torch.zeros_like(dX); dXT = dX.clone()
    eng3 = pd.PdSolver(sc)
    torch.cuda.synchronize()
    eng3.update_device_ptr(3, dX.data_ptr(), dV.data_ptr(), dXT.data_ptr())
    assert np.array_equal(dX.cpu().numpy(), Xr) and np.array_equal(dV.cpu().numpy(), Vr) and np.array_equal(dXT.cpu().numpy(), XTr)


def test_params_and_perf_interface(pd, assets):
    sc = pd.Scene.from_json(assets["json"], "C1 cube")
    eng = pd.PdSolver(sc)
    p = eng.get_params()
    assert p["num_iterations"] == 100 and p["global_solver"] == pd.PD_JACOBI
    p["num_iterations"] = 7
    eng.SetPerf(True)
    eng.Update(2, p)
    names, raw = eng.GetPerformanceData()
    assert [n for n, _ in names] == ["local step", "global step", "collision handling(fixed)", "collision handling(mesh)"]
    assert names[0][1] > 0 and names[1][1] > 0 and names[2][1] > 0 and names[3][1] == 0
    assert raw.steps == 2 and raw.pd_iterations == 14 and raw.kernel_launches == 2 * (2 + 14)
    # handleCollision=true (mesh-mesh BVH/CCD) is outside the hot path and must be refused loudly
    p["handle_collision"] = 1
    with pytest.raises(pd.PdError):
        eng.set_params(p)


def test_dbc_pinned_vertices_vs_oracle(pd, O, assets):
    X, E, _ = meshes.raw_mesh("sphere")
    T = (E[:, 1:5] - 1).astype(np.uint32)
    X = (X * np.float32(20)) + np.float32([0, 50, 0])
    dbc = np.zeros(X.shape[0], np.float32)
    dbc[np.argsort(X[:, 1])[-12:]] = 1.0           # pin the top cap
    p = pd.SolverParams(dt=0.01, gravity=98.0, num_iterations=100)
    sc = pd.Scene.from_arrays(X, T, 10.0, 2e5, DBC=dbc, params=p)
    eng = pd.PdSolver(sc)
    osc = O.Scene(X, T, 10.0, 2e5, DBC=dbc)
    eng.Update(20); osc.step(_oracle_params(O, p), 20)
    Xg, Xo = eng.download()[0], osc.get()[0]
    assert meshes.rel_err(Xg, Xo, _rest_scale(X)) <= TOL
    assert np.abs(Xg[dbc > 0] - X[dbc > 0]).max() < 2e-2            # pinned vertices stay put (soft 1e6 weight)
    assert (Xg[:, 1].min() < X[:, 1].min() - 0.5)                    # the rest sags under gravity


def test_large_grid_properties(pd):
    """Size-independent properties at a size the oracle cannot reach in seconds (1M tets):
    rigid free fall is an exact fixed point of the local/global iteration (R = F = I)."""
    sc = pd.Scene.kuhn_grid(55, 55, 55, 1.0, 0.05, 12345, (0, 1000, 0), 1.0, 2e5)
    nV, nT = sc.counts()[:2]
    assert (nV, nT) == (175616, 998250)
    p = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=20)
    sc.params = p
    eng = pd.PdSolver(sc)
    X0 = sc.arrays()["X"]
    n = 10
    eng.Update(n)
    X, V, XT = eng.download()
    h = float(np.float32(1 / 60))
    drop = 0.5 * 9.8 * h * h * n * (n + 1)
    d = X - X0
    assert np.abs(d[:, 0]).max() < 2e-3 and np.abs(d[:, 2]).max() < 2e-3
    assert np.abs(d[:, 1] + drop).max() < 5e-3 * max(1.0, drop) + 2e-3
    assert np.abs(V[:, 1] + 9.8 * h * n).max() < 2e-2


@pytest.mark.skipif(not os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libpd_ref.so")), reason="reference harness not built")
@pytest.mark.parametrize("ctx,steps", [("C1 cube", 100), ("C5 house&sphere", 100), ("C2 armadillo&bunny", 100), ("Armadillo&house", 60)])
def test_vs_reference_cuda_kernels(pd, O, assets, ctx, steps):
    """The pin: the reference's own kernels (compiled verbatim) on the same GPU, same scene, same step count."""
    import ref
    sc = pd.Scene.from_json(assets["json"], ctx)
    if ctx == "C1 cube":
        _params(pd, sc, dt=1 / 60)
    p = sc.params
    a = sc.arrays()
    osc, _ = meshes.oracle_scene(O, assets, ctx)      # only used to get the fixed bodies in array form
    planes, spheres, cyls = _fixed_arrays(pd, a["fixed"])
    rs = ref.RefScene(a["X"], a["Tet"], a["mass"], a["mu"], planes=planes, spheres=spheres, cylinders=cyls)
    eng = pd.PdSolver(sc)
    eng.Update(steps)
    rs.step(steps, dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=p["num_iterations"])
    X, V, XT = eng.download()
    Xr, Vr, XTr = rs.get()
    scale = _rest_scale(a["X"])
    e1, e2 = meshes.rel_err(X, Xr, scale), meshes.rel_err(XT, XTr, scale)
    print(f"{ctx}: rel err X {e1:.3e} XTilde {e2:.3e}")
    assert e1 <= TOL and e2 <= TOL


def _fixed_arrays(pd, fixed):
    planes, spheres, cyls = [], [], []
    for f in fixed:
        M = np.array(f.model[:], np.float32)
        if f.type == pd.PD_PLANE:
            up = np.zeros(3, np.float32)
            pd.lib().pd_plane_up(M.ctypes.data, up.ctypes.data)
            planes.append((M[12:15].copy(), up))
        elif f.type == pd.PD_SPHERE:
            spheres.append((M[12:15].copy(), f.radius))
        else:
            ax = M[4:7] / np.float32(np.linalg.norm(M[4:8]))
            cyls.append((M[12:15].copy(), ax.astype(np.float32), f.radius))
    return planes, spheres, cyls


def test_golden_fixtures(pd, assets):
    """Committed outputs of the reference's CUDA kernels on a B200 (tests/golden/README.md)."""
    path = os.path.join(HERE, "golden", "reference_b200.npz")
    if not os.path.exists(path):
        pytest.skip("golden reference outputs not generated yet")
    z = np.load(path)
    for ctx in ["C1 cube", "C5 house&sphere", "C2 armadillo&bunny"]:
        key = ctx.split()[0]
        sc = pd.Scene.from_json(assets["json"], ctx)
        if key == "C1":
            _params(pd, sc, dt=1 / 60)
        eng = pd.PdSolver(sc)
        eng.Update(int(z[key + "_steps"]))
        X, V, XT = eng.download()
        scale = _rest_scale(sc.arrays()["X"])
        assert meshes.rel_err(X, z[key + "_X"], scale) <= TOL, ctx
        assert meshes.rel_err(XT, z[key + "_XTilde"], scale) <= TOL, ctx
