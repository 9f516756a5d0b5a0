"""ctypes loader for oracle/_ref/libpd_ref.so: the REFERENCE's own CUDA kernels (compiled verbatim
from /root/reference by oracle/Makefile) behind the replay harness oracle/ref_harness.cu.
TEST INFRASTRUCTURE: parity pin + the reference arm of bench.py.  Needs a GPU to run."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libpd_ref.so")
SVD_CPU_PATH = os.path.join(_HERE, "_ref", "libref_svd_cpu.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_step.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_reset.argtypes = [C.c_void_p]
        L.ref_get.argtypes = [C.c_void_p] * 4
        L.ref_set.argtypes = [C.c_void_p] * 4
        L.ref_get_setup.argtypes = [C.c_void_p, C.c_float] + [C.c_void_p] * 4
        L.ref_get_perf.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_rotation.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_set_drag.argtypes = [C.c_void_p] * 4
        L.ref_drag_select.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.ref_get_drag.argtypes = [C.c_void_p] * 4
        _lib = L
    return _lib


class RefScene:
    def __init__(self, X, Tet, mass, mu, DBC=None, planes=(), spheres=(), cylinders=(), threads_per_block=128):
        """planes [(p0, up)], spheres [(c, r)], cylinders [(c, axis, r)] as for oracle.Scene"""
        X = np.ascontiguousarray(X, np.float32); Tet = np.ascontiguousarray(Tet, np.uint32)
        self.nV, self.nT = X.shape[0], Tet.shape[0]
        mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (self.nV,)))
        mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, np.float32), (self.nT,)))
        dbc = None if DBC is None else np.ascontiguousarray(DBC, np.float32)
        pl = np.ascontiguousarray(np.array([np.concatenate([p[0], p[1]]) for p in planes], np.float32).reshape(-1))
        sp = np.ascontiguousarray(np.array([np.concatenate([s[0], [s[1]]]) for s in spheres], np.float32).reshape(-1))
        cy = np.ascontiguousarray(np.array([np.concatenate([c[0], c[1], [c[2]]]) for c in cylinders], np.float32).reshape(-1))
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        self._h = lib().ref_create(self.nV, self.nT, p(X), p(Tet), p(mass), p(mu), p(dbc), len(planes), p(pl), len(spheres), p(sp),
                                   len(cylinders), p(cy), threads_per_block)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_destroy(self._h)
            self._h = None

    def step(self, n=1, dt=1 / 60, gravity=9.8, rho=0.9992, muN=0.5, muT=0.5, num_iterations=100, perf=False):
        rc = lib().ref_step(self._h, np.float32(dt), gravity, rho, muN, muT, num_iterations, n, int(perf))
        if rc:
            raise RuntimeError(f"reference harness CUDA error {rc}")

    def sync(self):
        return lib().ref_sync()

    def reset(self):
        lib().ref_reset(self._h)

    def get(self):
        X = np.zeros((self.nV, 3), np.float32); V = np.zeros_like(X); XT = np.zeros_like(X)
        lib().ref_get(self._h, X.ctypes.data, V.ctypes.data, XT.ctypes.data)
        return X, V, XT

    def set(self, X=None, V=None, XTilde=None):
        a = [None if t is None else np.ascontiguousarray(t, np.float32) for t in (X, V, XTilde)]
        lib().ref_set(self._h, *[None if t is None else t.ctypes.data for t in a])

    def set_drag(self, more_dbc=None, offset_x=None, target=(0.0, 0.0, 0.0)):
        if more_dbc is None:
            lib().ref_set_drag(self._h, None, None, None)
            return
        m = np.ascontiguousarray(more_dbc, np.float32).reshape(self.nV)
        o = np.ascontiguousarray(offset_x, np.float32).reshape(self.nV, 3)
        t = np.ascontiguousarray(target, np.float32).reshape(3)
        lib().ref_set_drag(self._h, m.ctypes.data, o.ctypes.data, t.ctypes.data)

    def drag_select(self, select_v, target, control_mag=10.0):
        t = np.ascontiguousarray(target, np.float32).reshape(3)
        lib().ref_drag_select(self._h, int(select_v), float(control_mag), t.ctypes.data)

    def get_drag(self):
        m = np.zeros(self.nV, np.float32); o = np.zeros((self.nV, 3), np.float32); d = np.zeros((self.nV, 3), np.float32)
        lib().ref_get_drag(self._h, m.ctypes.data, o.ctypes.data, d.ctypes.data)
        return m, o, d

    def setup(self, dt):
        md = np.zeros(self.nV, np.float32); c = np.zeros(self.nV, np.float32)
        B = np.zeros((self.nT, 3, 3), np.float32); V0 = np.zeros(self.nT, np.float32)
        lib().ref_get_setup(self._h, np.float32(dt), md.ctypes.data, c.ctypes.data, B.ctypes.data, V0.ctypes.data)
        return md, c, B.transpose(0, 2, 1).copy(), V0     # glm [col][row] -> row-major

    def perf(self):
        out = np.zeros(4, np.float32)
        lib().ref_get_perf(self._h, out.ctypes.data)
        return out


def rotation(F):
    F = np.ascontiguousarray(F, np.float32).reshape(-1, 9)
    R = np.zeros_like(F)
    rc = lib().ref_rotation(F.shape[0], F.ctypes.data, R.ctypes.data)
    if rc:
        raise RuntimeError(f"reference harness CUDA error {rc}")
    return R.reshape(-1, 3, 3)


# ---------------------------------------------------------------------------------------------------------------
# oracle/_ref/libpd_ref_solvers.so (oracle/ref_solvers.cu): the reference's direct / CG global-step back-ends
SOLVERS_PATH = os.path.join(_HERE, "_ref", "libpd_ref_solvers.so")
_slib = None


def solvers_available():
    return os.path.exists(SOLVERS_PATH)


def solvers_lib():
    global _slib
    if _slib is None:
        L = C.CDLL(SOLVERS_PATH)
        L.refs_create.restype = C.c_void_p
        L.refs_create.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
        L.refs_destroy.argtypes = [C.c_void_p]
        L.refs_step.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        L.refs_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refs_get.argtypes = [C.c_void_p] * 4
        L.refs_set.argtypes = [C.c_void_p] * 4
        L.refs_linear_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        _slib = L
    return _slib


class RefSolverScene:
    """PdSolver in a direct mode as the reference runs it: solver 1 = CuSolverCholesky (CholeskySpLinearSolver<float>),
    solver 2 = the same branch with PCGJacobiSolver<float> (warm start).  No fixed bodies."""

    def __init__(self, X, Tet, mass, mu, solver, DBC=None, threads_per_block=128):
        X = np.ascontiguousarray(X, np.float32); Tet = np.ascontiguousarray(Tet, np.uint32)
        self.nV, self.nT = X.shape[0], Tet.shape[0]
        mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (self.nV,)))
        mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, np.float32), (self.nT,)))
        dbc = None if DBC is None else np.ascontiguousarray(DBC, np.float32)
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        self._h = solvers_lib().refs_create(self.nV, self.nT, p(X), p(Tet), p(mass), p(mu), p(dbc), int(solver), threads_per_block)

    def __del__(self):
        if getattr(self, "_h", None):
            solvers_lib().refs_destroy(self._h)
            self._h = None

    def step(self, n=1, dt=1 / 60, gravity=9.8, tol=1e-4, num_iterations=10):
        rc = solvers_lib().refs_step(self._h, np.float32(dt), gravity, tol, num_iterations, n)
        if rc:
            raise RuntimeError(f"reference solver harness CUDA error {rc}")

    def stats(self):
        it, err = C.c_int(), C.c_float()
        solvers_lib().refs_stats(self._h, C.byref(it), C.byref(err))
        return it.value, err.value

    def get(self):
        X = np.zeros((self.nV, 3), np.float32); V = np.zeros_like(X); XT = np.zeros_like(X)
        solvers_lib().refs_get(self._h, X.ctypes.data, V.ctypes.data, XT.ctypes.data)
        return X, V, XT

    def set(self, X=None, V=None, XTilde=None):
        c = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        X, V, XTilde = c(X), c(V), c(XTilde)
        p = lambda a: None if a is None else a.ctypes.data
        solvers_lib().refs_set(self._h, p(X), p(V), p(XTilde))
        self._keep = (X, V, XTilde)


# ---------------------------------------------------------------------------------------------------------------
# oracle/_ref/libpd_ref_collision.so (oracle/ref_collision.cu): the reference's CCD arithmetic, verbatim
COLLISION_PATH = os.path.join(_HERE, "_ref", "libpd_ref_collision.so")
_clib = None


def collision_available():
    return os.path.exists(COLLISION_PATH)


def collision_lib():
    global _clib
    if _clib is None:
        L = C.CDLL(COLLISION_PATH)
        L.refc_ccd_queries.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refc_ccd_kernel.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int]
        _clib = L
    return _clib


def ccd_queries(types, verts, X, XTilde):
    """ccdCollisionTest<float> of the reference (intersections.cu:312-355) on n queries: types [n] (1 = VF, 2 = EE), verts [n, 4]
    vertex ids, X / XTilde [nV, 3] -> (toi [n], normals [n, 3])."""
    t = np.ascontiguousarray(types, np.int32); v = np.ascontiguousarray(verts, np.uint32).reshape(-1, 4)
    X = np.ascontiguousarray(X, np.float32); XT = np.ascontiguousarray(XTilde, np.float32)
    toi = np.zeros(t.shape[0], np.float32); nor = np.zeros((t.shape[0], 3), np.float32)
    rc = collision_lib().refc_ccd_queries(t.shape[0], t.ctypes.data, v.ctypes.data, X.shape[0], X.ctypes.data, XT.ctypes.data, toi.ctypes.data, nor.ctypes.data)
    if rc:
        raise RuntimeError(f"reference collision harness CUDA error {rc}")
    return toi, nor


def ccd_kernel(X, XTilde, V, tI, normals, dt=1 / 60, muT=0.5, muN=0.5, threads_per_block=128):
    """CCDKernel<float> of the reference (collisionUtil.cu:49-70) -> (X, V) after the kernel."""
    X = np.ascontiguousarray(X, np.float32).copy(); V = np.ascontiguousarray(V, np.float32).copy()
    XT = np.ascontiguousarray(XTilde, np.float32); tI = np.ascontiguousarray(tI, np.float32); n = np.ascontiguousarray(normals, np.float32)
    rc = collision_lib().refc_ccd_kernel(X.shape[0], X.ctypes.data, XT.ctypes.data, V.ctypes.data, tI.ctypes.data, n.ctypes.data, muT, muN, np.float32(dt), threads_per_block)
    if rc:
        raise RuntimeError(f"reference collision harness CUDA error {rc}")
    return X, V


def linear_solve(kind, A, row, col, b, guess=None, max_iter=0, tol=0.0):
    """LinearSolver<double>::Solve of the reference (kind 1 = CGSolver, IC(0); 2 = PCGJacobiSolver) on a host COO with duplicates -> x."""
    A = np.ascontiguousarray(A, np.float64); row = np.ascontiguousarray(row, np.int32); col = np.ascontiguousarray(col, np.int32)
    b = np.ascontiguousarray(b, np.float64); g = None if guess is None else np.ascontiguousarray(guess, np.float64)
    x = np.zeros(b.shape[0], np.float64)
    rc = solvers_lib().refs_linear_solve(kind, b.shape[0], A.shape[0], row.ctypes.data, col.ctypes.data, A.ctypes.data, b.ctypes.data,
                                         None if g is None else g.ctypes.data, x.ctypes.data, max_iter, tol)
    if rc:
        raise RuntimeError(f"reference solver harness CUDA error {rc}")
    return x
