"""Analytic invariants of the PD step ("analytic invariants the builder should add"), checked on the CPU
oracle -- the reference's tests pin nothing on this path, so besides the replayed reference kernels (tests/golden) the oracle
is held to what the algorithm guarantees by construction.  CPU only, seconds."""
import numpy as np
import pytest

import meshes


def _grid(n=4, h=1.0, seed=3, jitter=0.05):
    """Kuhn 6-tet grid, n^3 cells, jittered vertices (so that DmInv varies)."""
    rng = np.random.default_rng(seed)
    g = np.arange(n + 1)
    X = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * np.float32(h)
    X += rng.uniform(-jitter, jitter, X.shape).astype(np.float32)
    X[:, 1] += 10.0
    vid = lambda i, j, k: (i * (n + 1) + j) * (n + 1) + k
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    T = []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                for p in perms:
                    c = [i, j, k]; path = [vid(*c)]
                    for ax in p:
                        c[ax] += 1; path.append(vid(*c))
                    T.append(path)
    T = np.array(T, np.uint32)
    # positive orientation
    d = np.linalg.det((X[T[:, 1:]] - X[T[:, :1]]).astype(np.float64))
    T[d < 0] = T[d < 0][:, [0, 2, 1, 3]]
    return X, T


def test_rest_pose_and_rigid_motion_carry_no_elastic_force(O):
    X, T = _grid()
    p = O.make_params(dt=1 / 60, gravity=0.0, num_iterations=20)
    sc = O.Scene(X, T, 1.0, 2e5)
    sc.step(p, 3)
    assert np.abs(sc.get()[0] - X).max() <= 2e-5                       # F = I, R = I: nothing moves without gravity
    # rigid rotation + translation of the rest shape: R = F, the elastic right-hand side vanishes
    a = 0.7
    Q = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    Y = ((X.astype(np.float64) - X.mean(0)) @ Q.T + X.mean(0) + [3.0, 1.0, -2.0]).astype(np.float32)
    sc = O.Scene(X, T, 1.0, 2e5)
    sc.set(X=Y, XTilde=Y)
    sc.step(p, 3)
    assert np.abs(sc.get()[0] - Y).max() <= 1e-4 * np.abs(Y).max()


def test_free_fall_is_an_exact_fixed_point_of_every_sweep(O):
    """Without contact q = s solves the global step exactly: y_n = y_0 - g h^2 n (n + 1) / 2, x and z stay."""
    X, T = _grid()
    g, h, n = 9.8, 1 / 60, 12
    sc = O.Scene(X, T, 1.0, 2e5)
    sc.step(O.make_params(dt=h, gravity=g, num_iterations=30), n)
    Y, V, _ = sc.get()
    drop = 0.5 * g * np.float32(h) ** 2 * n * (n + 1)
    assert np.abs(Y[:, [0, 2]] - X[:, [0, 2]]).max() <= 2e-5
    assert np.abs((X[:, 1] - Y[:, 1]) - drop).max() <= 5e-5
    assert np.abs(V[:, 1] + g * np.float32(h) * n).max() <= 2e-3


def test_matrix_diag_and_system_matrix(O):
    import scipy.sparse as sp
    X, T = _grid()
    p = O.make_params(dt=1 / 60, gravity=9.8, num_iterations=5)
    mass = np.linspace(1.0, 2.0, X.shape[0]).astype(np.float32)
    sc = O.Scene(X, T, mass, 2e5)
    md, c, B, V0 = sc.setup(p)
    rp, col, val = sc.system_matrix(p)
    n = X.shape[0]
    A = sp.csr_matrix((val.astype(np.float64), col, rp), shape=(n, n))
    assert np.allclose(c, mass / np.float32(1 / 60) ** 2, rtol=1e-6)                                  # setMDt_2, pdUtil.cu:48
    assert np.allclose(A.diagonal(), md.astype(np.float64) + c, rtol=2e-6)                            # diag(A^) = matrix_diag + M/h^2
    assert abs(A - A.T).max() <= 1e-6 * abs(A).max()                                                 # symmetric
    lap = A - sp.diags(c.astype(np.float64))
    assert abs(np.asarray(lap.sum(1)).ravel()).max() <= 1e-5 * abs(A).max()                          # Laplacian rows sum to zero
    # rest-shape products: V0 = |det Dm| / 6, DmInv Dm = I
    Dm = np.transpose(X[T[:, 1:]] - X[T[:, :1]], (0, 2, 1)).astype(np.float64)                       # columns = edges
    assert np.allclose(V0, np.abs(np.linalg.det(Dm)) / 6, rtol=1e-5)
    assert np.abs(np.einsum("tij,tjk->tik", Dm, B.astype(np.float64)) - np.eye(3)).max() <= 1e-4
    assert np.all(np.linalg.eigvalsh(A.toarray()) > 0)                                               # SPD: Cholesky / CG are applicable


def test_corotational_projection(O):
    """svd3_cuda.h is an approximate (4-sweep, rsqrt-based) SVD: R = U V^T is a proper rotation close to the polar factor."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for i in range(200):
        F = (np.eye(3) + 0.4 * rng.standard_normal((3, 3))).astype(np.float32)
        if i % 10 == 0:
            F[:, 2] *= -1                                            # inverted element: U, V stay proper, sigma_3 carries the sign
        R = O.rotation(F).astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() <= 2e-3 and np.linalg.det(R) > 0.99
        U, S, Vt = np.linalg.svd(F.astype(np.float64))
        if np.linalg.det(U @ Vt) < 0:
            U[:, 2] *= -1
        if S[1] - S[2] > 0.05 and S[2] > 0.05 and np.linalg.det(F) > 0:  # well separated: the polar rotation is unique and stable
            worst = max(worst, np.abs(R - U @ Vt).max())
        Us, Ss, Vs = O.svd3(F)
        assert np.allclose(np.sort(np.abs(Ss))[::-1], S, rtol=5e-3, atol=5e-3)
        assert np.abs((Us * Ss) @ Vs.T - F).max() <= 5e-3            # A = U S V^T
    assert worst <= 5e-3, worst


def test_pinned_vertices_stay(O):
    X, T = _grid()
    dbc = np.zeros(X.shape[0], np.float32)
    top = X[:, 1] > X[:, 1].max() - 0.3
    dbc[top] = 1.0
    sc = O.Scene(X, T, 1.0, 2e5, DBC=dbc)
    sc.step(O.make_params(dt=1 / 60, gravity=9.8, num_iterations=60), 20)
    Y = sc.get()[0]
    assert np.abs(Y[top] - X[top]).max() <= 1e-3                       # pulled to DBCX = X0 with weight 1e6 / h^2
    assert (X[~top, 1] - Y[~top, 1]).max() > 1e-3 and np.isfinite(Y).all()   # the rest sags under gravity
