"""ctypes loader for the CPU oracle (oracle/libpd_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. The product never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class FixedBodies(C.Structure):
    _fields_ = [("n_planes", C.c_int), ("plane_p0", C.c_void_p), ("plane_up", C.c_void_p),
                ("n_spheres", C.c_int), ("sphere_c", C.c_void_p), ("sphere_r", C.c_void_p),
                ("n_cyls", C.c_int), ("cyl_c", C.c_void_p), ("cyl_axis", C.c_void_p),
                ("cyl_r", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("dt", C.c_float), ("gravity", C.c_float), ("muN", C.c_float), ("muT", C.c_float),
                ("rho", C.c_float), ("tol", C.c_float), ("num_iterations", C.c_int),
                ("global_solver", C.c_int), ("pcg_max_iter", C.c_int), ("pcg_tol", C.c_float),
                ("threads", C.c_int)]


def make_params(dt=1.0 / 60.0, gravity=9.8, muN=0.5, muT=0.5, rho=0.9992, tol=1e-2,
                num_iterations=100, global_solver=0, pcg_max_iter=2000, pcg_tol=1e-5, threads=1):
    return Params(np.float32(dt), gravity, muN, muT, rho, tol, num_iterations, global_solver,
                  pcg_max_iter, pcg_tol, threads)


def build(force=False):
    so = os.path.join(_HERE, "libpd_oracle.so")
    src = os.path.join(_HERE, "pd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libpd_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.o_svd3.argtypes = [f32p, f32p, f32p, f32p]
        L.o_rotation.argtypes = [f32p, f32p]
        L.o_rest_shape.argtypes = [f32p, u32p, C.c_int, f32p, f32p]
        L.o_model_matrix.argtypes = [f32p, f32p, f32p, C.c_int, f32p]
        L.o_transform_vertices.argtypes = [f32p, C.c_int, f32p]
        L.o_plane_up.argtypes = [f32p, f32p]
        L.o_cylinder_axis.argtypes = [f32p, f32p]
        L.o_load_node.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.o_load_ele.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.o_free.argtypes = [C.c_void_p]
        L.o_scene_create.restype = C.c_void_p
        L.o_scene_create.argtypes = [C.c_int, C.c_int, f32p, u32p, f32p, f32p, f32p, C.POINTER(FixedBodies)]
        L.o_scene_destroy.argtypes = [C.c_void_p]
        L.o_scene_step.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
        L.o_scene_step_f64.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
        L.o_scene_reset.argtypes = [C.c_void_p]
        L.o_scene_get.argtypes = [C.c_void_p, f32p, f32p, f32p]
        L.o_scene_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.o_scene_get_setup.argtypes = [C.c_void_p, C.POINTER(Params), f32p, f32p, f32p, f32p]
        L.o_scene_system_matrix.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.o_scene_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.o_scene_set_drag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.o_scene_drag_select.argtypes = [C.c_void_p, C.c_int, C.c_float, f32p]
        L.o_scene_get_drag.argtypes = [C.c_void_p, f32p, f32p, f32p]
        L.o_scene_set_mu.argtypes = [C.c_void_p, f32p]
        L.o_scene_set_collision.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.o_scene_get_collision.argtypes = [C.c_void_p, f32p, f32p, C.POINTER(C.c_longlong)]
        L.o_ccd_test.restype = C.c_float
        L.o_ccd_test.argtypes = [C.c_int, u32p, f32p, f32p, f32p]
        _LIB = L
    return _LIB


def svd3(A):
    A = np.ascontiguousarray(A, np.float32).reshape(9)
    U = np.zeros(9, np.float32); S = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
    lib().o_svd3(A, U, S, V)
    return U.reshape(3, 3), S, V.reshape(3, 3)


def ccd_test(ee, q, X, XTilde):
    """ccdCollisionTest<float> (intersections.cu:312-355) on one query -> (toi, normal)."""
    q = np.ascontiguousarray(q, np.uint32).reshape(4)
    X = np.ascontiguousarray(X, np.float32).reshape(-1); XT = np.ascontiguousarray(XTilde, np.float32).reshape(-1)
    n = np.zeros(3, np.float32)
    t = lib().o_ccd_test(int(bool(ee)), q, X, XT, n)
    return float(t), n


def rotation(F):
    F = np.ascontiguousarray(F, np.float32).reshape(9)
    R = np.zeros(9, np.float32)
    lib().o_rotation(F, R)
    return R.reshape(3, 3)


def model_matrix(pos, rot, scale, soft_body_order):
    M = np.zeros(16, np.float32)
    lib().o_model_matrix(np.asarray(pos, np.float32), np.asarray(rot, np.float32),
                         np.asarray(scale, np.float32), int(soft_body_order), M)
    return M


def transform_vertices(X, M):
    X = np.ascontiguousarray(X, np.float32).copy()
    lib().o_transform_vertices(X.reshape(-1), X.shape[0], np.ascontiguousarray(M, np.float32))
    return X


def plane_up(M):
    up = np.zeros(3, np.float32)
    lib().o_plane_up(np.ascontiguousarray(M, np.float32), up)
    return up


def cylinder_axis(M):
    ax = np.zeros(3, np.float32)
    lib().o_cylinder_axis(np.ascontiguousarray(M, np.float32), ax)
    return ax


def load_node(path, centralize):
    p = C.c_void_p()
    n = lib().o_load_node(path.encode(), int(centralize), C.byref(p))
    if n < 0:
        raise IOError(f"o_load_node({path}) -> {n}")
    X = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, 3)).copy()
    lib().o_free(p)
    return X


def load_ele(path, start_index):
    p = C.c_void_p()
    n = lib().o_load_ele(path.encode(), int(start_index), C.byref(p))
    if n < 0:
        raise IOError(f"o_load_ele({path}) -> {n}")
    T = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n, 4)).copy()
    lib().o_free(p)
    return T


class Scene:
    """Merged soft bodies + fixed bodies + PD solver state (CPU oracle)."""

    def __init__(self, X, Tet, mass, mu, DBC=None, planes=(), spheres=(), cylinders=()):
        """planes: [(p0[3], up[3])], spheres: [(c[3], r)], cylinders: [(c[3], axis[3], r)]"""
        self.X0 = np.ascontiguousarray(X, np.float32)
        self.Tet = np.ascontiguousarray(Tet, np.uint32)
        self.nV, self.nT = self.X0.shape[0], self.Tet.shape[0]
        mass = np.broadcast_to(np.asarray(mass, np.float32), (self.nV,)).copy()
        mu = np.broadcast_to(np.asarray(mu, np.float32), (self.nT,)).copy()
        DBC = np.zeros(self.nV, np.float32) if DBC is None else np.ascontiguousarray(DBC, np.float32)
        self._keep = []

        def arr(rows, w):
            a = np.ascontiguousarray(np.asarray(rows, np.float32).reshape(-1, w)) if len(rows) else np.zeros((0, w), np.float32)
            self._keep.append(a)
            return a.ctypes.data

        fb = FixedBodies()
        fb.n_planes = len(planes)
        fb.plane_p0 = arr([p[0] for p in planes], 3); fb.plane_up = arr([p[1] for p in planes], 3)
        fb.n_spheres = len(spheres)
        fb.sphere_c = arr([s[0] for s in spheres], 3); fb.sphere_r = arr([[s[1]] for s in spheres], 1)
        fb.n_cyls = len(cylinders)
        fb.cyl_c = arr([c[0] for c in cylinders], 3); fb.cyl_axis = arr([c[1] for c in cylinders], 3)
        fb.cyl_r = arr([[c[2]] for c in cylinders], 1)
        self._h = lib().o_scene_create(self.nV, self.nT, self.X0.reshape(-1), self.Tet.reshape(-1), mass, mu, DBC, C.byref(fb))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().o_scene_destroy(self._h)
            self._h = None

    def step(self, params, n=1, f64=False):
        fn = lib().o_scene_step_f64 if f64 else lib().o_scene_step
        rc = fn(self._h, C.byref(params), n)
        if rc:
            raise RuntimeError(f"oracle step failed rc={rc}")

    def reset(self):
        lib().o_scene_reset(self._h)

    def get(self):
        X = np.zeros((self.nV, 3), np.float32); V = np.zeros_like(X); XT = np.zeros_like(X)
        lib().o_scene_get(self._h, X.reshape(-1), V.reshape(-1), XT.reshape(-1))
        return X, V, XT

    def set(self, X=None, V=None, XTilde=None):
        a = [np.ascontiguousarray(t, np.float32) if t is not None else None for t in (X, V, XTilde)]
        lib().o_scene_set(self._h, *[t.ctypes.data if t is not None else None for t in a])

    def setup(self, params):
        md = np.zeros(self.nV, np.float32); c = np.zeros(self.nV, np.float32)
        B = np.zeros((self.nT, 9), np.float32); V0 = np.zeros(self.nT, np.float32)
        lib().o_scene_get_setup(self._h, C.byref(params), md, c, B.reshape(-1), V0)
        return md, c, B.reshape(self.nT, 3, 3), V0

    def system_matrix(self, params):
        nnz = lib().o_scene_system_matrix(self._h, C.byref(params), None, None, None)
        rp = np.zeros(self.nV + 1, np.int32); col = np.zeros(nnz, np.int32); val = np.zeros(nnz, np.float32)
        lib().o_scene_system_matrix(self._h, C.byref(params), rp.ctypes.data, col.ctypes.data, val.ctypes.data)
        return rp, col, val

    def set_drag(self, more_dbc=None, offset_x=None, target=(0.0, 0.0, 0.0)):
        """SolverData::moreDBC / OffsetX / mouseSelection.target; more_dbc None clears (ResetMoreDBC(true))."""
        if more_dbc is None:
            lib().o_scene_set_drag(self._h, None, None, None)
            return
        m = np.ascontiguousarray(more_dbc, np.float32).reshape(self.nV)
        o = np.ascontiguousarray(offset_x, np.float32).reshape(self.nV, 3)
        t = np.ascontiguousarray(target, np.float32).reshape(3)
        lib().o_scene_set_drag(self._h, m.ctypes.data, o.ctypes.data, t.ctypes.data)

    def drag_select(self, select_v, target, control_mag=10.0):
        """Control_Kernel (simulationContext.cu:202-218) on the current X."""
        lib().o_scene_drag_select(self._h, int(select_v), float(control_mag), np.ascontiguousarray(target, np.float32).reshape(3))

    def set_mu(self, mu):
        lib().o_scene_set_mu(self._h, np.ascontiguousarray(mu, np.float32).reshape(self.nT))

    def set_collision(self, enable, Tri=None, TriFathers=None):
        """SolverParams::handleCollision + SolverData::Tri / dev_TriFathers (pdSolver.cu:218-225)."""
        if Tri is None:
            lib().o_scene_set_collision(self._h, int(bool(enable)), 0, None, None)
            return
        t = np.ascontiguousarray(Tri, np.uint32).reshape(-1, 3)
        f = None if TriFathers is None else np.ascontiguousarray(TriFathers, np.uint32).reshape(t.shape[0])
        lib().o_scene_set_collision(self._h, int(bool(enable)), t.shape[0], t.ctypes.data, None if f is None else f.ctypes.data)

    def collision(self):
        """(tI, normals, ordered overlapping triangle pairs) of the last collision pass."""
        t = np.ones(self.nV, np.float32); n = np.zeros((self.nV, 3), np.float32); p = C.c_longlong()
        lib().o_scene_get_collision(self._h, t, n.reshape(-1), C.byref(p))
        return t, n, p.value

    def get_drag(self):
        m = np.zeros(self.nV, np.float32); o = np.zeros((self.nV, 3), np.float32); d = np.zeros((self.nV, 3), np.float32)
        lib().o_scene_get_drag(self._h, m, o.reshape(-1), d.reshape(-1))
        return m, o, d

    def stats(self):
        a = C.c_int(); b = C.c_int()
        lib().o_scene_stats(self._h, C.byref(a), C.byref(b))
        return a.value, b.value
