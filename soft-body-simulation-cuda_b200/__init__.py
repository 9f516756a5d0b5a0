"""B200-native projective-dynamics step engine -- Python host bindings over the C ABI.

The product is the shared library `libpd_b200.so` (CUDA kernels for sm_100a + C++ host code,
sources under csrc/, C ABI in include/pd_b200.h).  This module only binds it with ctypes so
that tests and bench.py can drive it; it contains no numerics and has NO CPU fallback: the
engine constructors raise if the library or a sm_100 GPU is missing.

The class/method names mirror the reference's interface for this path
(src/simulation/solver/solver.h:11-28, projective/pdSolver.h:12-44,
src/simulation/simulationContext.h:17-84):  SolverParams, PdSolver.Update/Reset/SetPerf/
GetPerformanceData/SetGlobalSolver, SimulationContext.Update/Reset.

Import with  importlib.import_module("soft-body-simulation-cuda_b200")  (the hyphens rule out
a plain `import` statement).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PD_B200_LIB") or os.path.join(_HERE, "libpd_b200.so")     # PD_B200_LIB: kernel-variant experiments only

PD_JACOBI, PD_CHOLESKY, PD_PCG_JACOBI = 0, 1, 2
PD_PLANE, PD_SPHERE, PD_CYLINDER = 0, 1, 2


class PdError(RuntimeError):
    pass


class pd_fixed_body(C.Structure):
    _fields_ = [("type", C.c_int), ("model", C.c_float * 16), ("radius", C.c_float)]


class pd_params(C.Structure):
    _fields_ = [("dt", C.c_float), ("gravity", C.c_float), ("muN", C.c_float), ("muT", C.c_float),
                ("rho", C.c_float), ("tol", C.c_float), ("damp", C.c_float),
                ("num_iterations", C.c_int), ("global_solver", C.c_int), ("pcg_max_iter", C.c_int),
                ("pcg_tol", C.c_float), ("handle_collision", C.c_int), ("threads_per_block", C.c_int)]


class pd_scene_desc(C.Structure):
    _fields_ = [("num_verts", C.c_int), ("num_tets", C.c_int), ("X", C.c_void_p), ("Tet", C.c_void_p),
                ("mass", C.c_void_p), ("mu", C.c_void_p), ("DBC", C.c_void_p),
                ("num_fixed", C.c_int), ("fixed", C.POINTER(pd_fixed_body)),
                ("num_tris", C.c_int), ("Tri", C.c_void_p), ("TriFathers", C.c_void_p)]


class pd_engine_options(C.Structure):
    _fields_ = [("device", C.c_int), ("rot_mode", C.c_int), ("reorder", C.c_int), ("use_graph", C.c_int),
                ("ctas_per_sm", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("body_kernel", C.c_int)]


class pd_perf(C.Structure):
    _fields_ = [("local_step_ms", C.c_float), ("global_step_ms", C.c_float), ("collision_fixed_ms", C.c_float),
                ("collision_mesh_ms", C.c_float), ("step_ms_total", C.c_double), ("steps", C.c_longlong),
                ("pd_iterations", C.c_longlong), ("inner_iterations", C.c_longlong), ("kernel_launches", C.c_longlong)]


# every symbol include/pd_b200.h declares: name -> (restype, argtypes)
_VP, _I, _F, _CP = C.c_void_p, C.c_int, C.c_float, C.c_char_p
_PI, _PF = C.POINTER(C.c_int), C.POINTER(C.c_float)
SYMBOLS = {
    "pd_last_error": (_CP, []),
    "pd_version": (_CP, []),
    "pd_default_params": (None, [C.POINTER(pd_params)]),
    "pd_default_options": (None, [C.POINTER(pd_engine_options)]),
    "pd_scene_load_json": (_VP, [_CP, _CP, _CP]),
    "pd_scene_from_desc": (_VP, [C.POINTER(pd_scene_desc), C.POINTER(pd_params)]),
    "pd_scene_kuhn_grid": (_VP, [_I, _I, _I, _F, _F, C.c_uint32, _VP, _F, _F]),
    "pd_scene_merge": (_VP, [_VP, _I]),
    "pd_scene_free": (None, [_VP]),
    "pd_scene_counts": (_I, [_VP, _PI, _PI, _PI, _PI]),
    "pd_scene_get": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pd_scene_get_params": (_I, [_VP, C.POINTER(pd_params)]),
    "pd_scene_set_params": (_I, [_VP, C.POINTER(pd_params)]),
    "pd_scene_add_fixed": (_I, [_VP, C.POINTER(pd_fixed_body)]),
    "pd_scene_write_tetgen": (_I, [_VP, _CP, _CP]),
    "pd_load_node": (_I, [_CP, _I, C.POINTER(_VP), _PI]),
    "pd_load_ele": (_I, [_CP, _I, C.POINTER(_VP), _PI]),
    "pd_free": (None, [_VP]),
    "pd_model_matrix": (None, [_VP, _VP, _VP, _I, _VP]),
    "pd_transform_vertices": (None, [_VP, _I, _VP]),
    "pd_plane_up": (None, [_VP, _VP]),
    "pd_layout_build": (_VP, [_VP, _I]),
    "pd_layout_free": (None, [_VP]),
    "pd_layout_counts": (_I, [_VP, _PI, C.POINTER(C.c_uint32), C.POINTER(C.c_size_t), _PI]),
    "pd_layout_get": (_I, [_VP] * 10),
    "pd_layout_get_vstage": (_I, [_VP, _VP]),
    "pd_layout_tile_table": (_I, [_VP, _VP]),
    "pd_layout_matrix_diag": (_I, [_VP, _VP]),
    "pd_morton_keys": (_I, [_VP, _VP, _I, _VP]),
    "pd_partition_vertices": (_I, [_I, _I, _VP]),
    "pd_nested_dissection": (_I, [_I, _VP, _VP, _VP, _VP, _PI, _PI]),
    "pd_scene_get_surface": (_I, [_VP, _PI, _VP, _VP]),
    "pd_cholesky_factor": (_I, [_I, _VP, _VP, _VP, _PI, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    "pd_rank_plan_build": (_VP, [_VP, _I, _I]),
    "pd_rank_plan_free": (None, [_VP]),
    "pd_rank_plan_counts": (_I, [_VP, _VP]),
    "pd_rank_plan_get": (_I, [_VP] * 8),
    "pd_rank_layout": (_VP, [_VP, _VP]),
    "pd_dist_window_handle": (_I, [_VP, _VP]),
    "pd_dist_connect": (_I, [_VP, _VP]),
    "pd_dist_connect_local": (_I, [_VP, _I]),
    "pd_dist_step_lockstep": (_I, [_VP, _I, _I]),
    "pd_dist_status": (_I, [_VP, _VP]),
    "pd_dist_info": (_I, [_VP, _VP]),
    "pd_create": (_VP, [_VP, C.POINTER(pd_engine_options)]),
    "pd_create_from_json": (_VP, [_CP, _CP, _CP, C.POINTER(pd_engine_options)]),
    "pd_destroy": (None, [_VP]),
    "pd_step": (_I, [_VP, _I]),
    "pd_synchronize": (_I, [_VP]),
    "pd_step_timed": (_I, [_VP, _I, _PF]),
    "pd_set_params": (_I, [_VP, C.POINTER(pd_params)]),
    "pd_get_params": (_I, [_VP, C.POINTER(pd_params)]),
    "pd_set_global_solver": (_I, [_VP, _I]),
    "pd_reset": (_I, [_VP]),
    "pd_set_perf": (_I, [_VP, _I]),
    "pd_get_perf": (_I, [_VP, C.POINTER(pd_perf)]),
    "pd_download": (_I, [_VP, _VP, _VP, _VP]),
    "pd_upload_state": (_I, [_VP, _VP, _VP, _VP]),
    "pd_step_host": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pd_step_host_owned": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pd_dist_owned_ids": (_I, [_VP, _VP]),
    "pd_update_device": (_I, [_VP, _I, _VP, _VP, _VP]),
    "pd_set_drag": (_I, [_VP, _VP, _VP, _VP]),
    "pd_set_drag_device": (_I, [_VP, _VP, _VP, _VP]),
    "pd_drag_select": (_I, [_VP, _I, _F, _VP]),
    "pd_get_drag": (_I, [_VP, _VP, _VP, _VP, _PI]),
    "pd_update_mu": (_I, [_VP, _VP]),
    "pd_update_mu_device": (_I, [_VP, _VP]),
    "pd_get_setup": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "pd_get_system_matrix": (_I, [_VP, _PI, _VP, _VP, _VP]),
    "pd_get_solve_stats": (_I, [_VP, _PF, _PI]),
    "pd_linsolver_create": (_VP, [_I, _I, _I, C.c_double, _I]),
    "pd_linsolver_destroy": (None, [_VP]),
    "pd_linsolver_solve_device": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "pd_linsolver_solve_host": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "pd_linsolver_stats": (_I, [_VP, _PI, C.POINTER(C.c_double), _PI]),
    "pd_ccd_batch": (_I, [_I, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP]),
    "pd_get_collision": (_I, [_VP, _VP, _VP, C.POINTER(C.c_longlong)]),
    "pd_get_solver_sizes": (_I, [_VP, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "pd_time_kernels": (_I, [_VP, _I, _PF, _PF]),
    "pd_profile_local": (_I, [_VP, _VP]),
    "pd_engine_rot_mode": (_I, [_VP]),
    "pd_engine_info": (_I, [_VP, _PI, _PI, _PI, C.POINTER(C.c_uint32), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), _PI]),
    "pd_rotation_batch": (_I, [_I, _I, _I, _VP, _VP, _VP]),
    "pd_alloc_pinned": (_VP, [C.c_size_t]),
    "pd_free_pinned": (None, [_VP]),
}

_lib = None


def lib():
    """Load libpd_b200.so (built by build.py / __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdError(f"{LIB_PATH} is missing: build it with `python {os.path.join(_HERE, 'build.py')}`; "
                          "there is no CPU fallback for the PD engine")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _err():
    return lib().pd_last_error().decode()


def _check(rc):
    if rc != 0:
        raise PdError(f"pd_b200 error {rc}: {_err()}")


def _p(a):
    return None if a is None else a.ctypes.data


class SolverParams:
    """SolverParams<float> (src/def.h:82-100)."""

    def __init__(self, **kw):
        self.c = pd_params()
        lib().pd_default_params(C.byref(self.c))
        for k, v in kw.items():
            self[k] = v

    _alias = {"numIterations": "num_iterations", "globalSolver": "global_solver", "handleCollision": "handle_collision"}

    def __setitem__(self, k, v):
        k = self._alias.get(k, k)
        if not hasattr(self.c, k):
            raise KeyError(k)
        if k == "dt":
            v = float(np.float32(v))
        setattr(self.c, k, v)

    def __getitem__(self, k):
        return getattr(self.c, self._alias.get(k, k))


def fixed_body(kind, pos=(0, 0, 0), rot=(0, 0, 0), scale=(1, 1, 1), radius=1.0):
    """Build a pd_fixed_body the way Context::ReadFixedBodies does (context.cpp:222-317)."""
    fb = pd_fixed_body()
    fb.type = kind
    M = np.zeros(16, np.float32)
    if kind == PD_SPHERE:
        scale = (radius, radius, radius)
    elif kind == PD_CYLINDER:
        radius = scale[0]
        scale = (scale[0], scale[1], scale[0])
    # (named arrays: a temporary handed to _p() is freed before the call and its memory reused by the next temporary)
    pos_a, rot_a, scale_a = (np.ascontiguousarray(t, np.float32) for t in (pos, rot, scale))
    lib().pd_model_matrix(_p(pos_a), _p(rot_a), _p(scale_a), 0, _p(M))
    fb.model[:] = M.tolist()
    fb.radius = float(radius)
    return fb


class Scene:
    """Host-side merged scene: what DataLoader::AllocData + Impl::Init produce (no GPU needed)."""

    def __init__(self, handle):
        if not handle:
            raise PdError(_err())
        self._h = handle

    @classmethod
    def from_json(cls, json_path, context_name=None, asset_root=None):
        enc = lambda s: None if s is None else str(s).encode()
        return cls(lib().pd_scene_load_json(enc(json_path), enc(context_name), enc(asset_root)))

    @classmethod
    def from_arrays(cls, X, Tet, mass, mu, DBC=None, fixed=(), params=None, Tri=None, TriFathers=None):
        X = np.ascontiguousarray(X, np.float32); Tet = np.ascontiguousarray(Tet, np.uint32)
        nV, nT = X.shape[0], Tet.shape[0]
        mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (nV,)))
        mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, np.float32), (nT,)))
        dbc = None if DBC is None else np.ascontiguousarray(DBC, np.float32)
        arr = (pd_fixed_body * max(len(fixed), 1))(*fixed)
        tri = None if Tri is None else np.ascontiguousarray(Tri, np.uint32).reshape(-1, 3)
        fa = None if (Tri is None or TriFathers is None) else np.ascontiguousarray(TriFathers, np.uint32)
        d = pd_scene_desc(nV, nT, _p(X), _p(Tet), _p(mass), _p(mu), _p(dbc), len(fixed), arr, 0 if tri is None else tri.shape[0], _p(tri), _p(fa))
        return cls(lib().pd_scene_from_desc(C.byref(d), C.byref(params.c) if params else None))

    @classmethod
    def kuhn_grid(cls, nx, ny, nz, h=1.0, jitter=0.05, seed=12345, origin=(0, 0, 0), mass=1.0, mu=2e5):
        o = np.asarray(origin, np.float32)
        return cls(lib().pd_scene_kuhn_grid(nx, ny, nz, h, jitter, seed, _p(o), mass, mu))

    @classmethod
    def merge(cls, scenes):
        """Batch of independent contexts -> one scene (same params / fixed bodies); see pd_scene_merge."""
        arr = (C.c_void_p * len(scenes))(*[s._h for s in scenes])
        return cls(lib().pd_scene_merge(arr, len(scenes)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.pd_scene_free(self._h)
            self._h = None

    def counts(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().pd_scene_counts(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def arrays(self):
        nV, nT, nF, nB = self.counts()
        X = np.zeros((nV, 3), np.float32); T = np.zeros((nT, 4), np.uint32)
        mass = np.zeros(nV, np.float32); mu = np.zeros(nT, np.float32); dbc = np.zeros(nV, np.float32)
        fixed = (pd_fixed_body * max(nF, 1))()
        bvs = np.zeros(max(nB, 1), np.int32)
        _check(lib().pd_scene_get(self._h, _p(X), _p(T), _p(mass), _p(mu), _p(dbc), C.cast(fixed, C.c_void_p), _p(bvs)))
        return dict(X=X, Tet=T, mass=mass, mu=mu, DBC=dbc, fixed=[fixed[i] for i in range(nF)], body_vert_start=bvs[:nB])

    @property
    def params(self):
        p = SolverParams()
        _check(lib().pd_scene_get_params(self._h, C.byref(p.c)))
        return p

    @params.setter
    def params(self, p):
        _check(lib().pd_scene_set_params(self._h, C.byref(p.c)))

    def surface(self):
        """(Tri [n, 3], TriFathers [n]): the surface triangles the mesh-mesh collision pass uses (explicit .face / caller-given
        ones, else the boundary faces of every body's tets)."""
        n = C.c_int()
        _check(lib().pd_scene_get_surface(self._h, C.byref(n), None, None))
        tri = np.zeros((n.value, 3), np.uint32); fa = np.zeros(n.value, np.uint32)
        _check(lib().pd_scene_get_surface(self._h, C.byref(n), _p(tri), _p(fa)))
        return tri, fa

    def add_fixed(self, fb):
        _check(lib().pd_scene_add_fixed(self._h, C.byref(fb)))

    def write_tetgen(self, node_path, ele_path):
        _check(lib().pd_scene_write_tetgen(self._h, str(node_path).encode(), str(ele_path).encode()))

    def layout(self, reorder=True):
        return Layout(self, reorder)


class Layout:
    """Device layout (tet order, vertex renumbering, tile records, slot CSR), host only."""

    def __init__(self, scene, reorder=True, handle=None, counts=None):
        self._h = handle if handle is not None else lib().pd_layout_build(scene._h, int(reorder))
        if not self._h:
            raise PdError(_err())
        nV, nT = counts if counts is not None else scene.counts()[:2]
        nt, ns, rb, ml = C.c_int(), C.c_uint32(), C.c_size_t(), C.c_int()
        _check(lib().pd_layout_counts(self._h, C.byref(nt), C.byref(ns), C.byref(rb), C.byref(ml)))
        self.num_tiles, self.num_slots, self.record_bytes, self.max_local = nt.value, ns.value, rb.value, ml.value
        self.tet_order = np.zeros(nT, np.uint32); self.vert_order = np.zeros(nV, np.uint32)
        self.tet_new = np.zeros((nT, 4), np.uint32)
        self.tile_tet_start = np.zeros(self.num_tiles + 1, np.uint32)
        self.tile_rec_off = np.zeros(self.num_tiles + 1, np.uint64)
        self.records = np.zeros(self.record_bytes, np.uint8)
        self.vslot_ptr = np.zeros(nV + 1, np.uint32); self.vslot = np.zeros(self.num_slots, np.uint32)
        self.vlist = np.zeros(self.num_tiles * 256, np.uint32)      # padded: tile * TILE_NLMAX + local vertex
        _check(lib().pd_layout_get(self._h, _p(self.tet_order), _p(self.vert_order), _p(self.tet_new), _p(self.tile_tet_start),
                                   _p(self.tile_rec_off), _p(self.records), _p(self.vslot_ptr), _p(self.vslot), _p(self.vlist)))
        self.vstage = np.zeros(self.num_tiles * 256, np.uint32)     # staging slot -> vlist entry of the vertex staged there
        _check(lib().pd_layout_get_vstage(self._h, _p(self.vstage)))
        self.tile_table = np.zeros((self.num_tiles, 12), np.uint32)  # what the local kernel reads per tile (layout.hpp TILE_META_WORDS)
        _check(lib().pd_layout_tile_table(self._h, _p(self.tile_table)))
        self.num_tets = int(self.tile_tet_start[-1])                 # (a trimmed rank layout holds fewer tets than its plan's tiles)
        self.tet_order, self.tet_new = self.tet_order[:self.num_tets], self.tet_new[:self.num_tets]

    def matrix_diag(self):
        """SolverPrepare's matrix_diag per layout vertex, as Engine::prepare computes it (host only)."""
        md = np.zeros(len(self.vert_order), np.float32)
        _check(lib().pd_layout_matrix_diag(self._h, _p(md)))
        return md

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.pd_layout_free(self._h)
            self._h = None


class RankPlan:
    """Multi-GPU plan of one rank (host only): tiles evaluated, ghosts, neighbours, push lists."""

    def __init__(self, layout, world, rank):
        self._h = lib().pd_rank_plan_build(layout._h, world, rank)
        if not self._h:
            raise PdError(_err())
        self.world, self.rank = world, rank
        c = np.zeros(7, np.int32)
        _check(lib().pd_rank_plan_counts(self._h, _p(c)))
        self.num_owned, self.num_ghosts, self.num_tiles, self.num_neighbours, self.num_push, self.first_owned, self.num_interior_tiles = [int(x) for x in c]
        self.tiles = np.zeros(self.num_tiles, np.uint32); self.ghosts = np.zeros(self.num_ghosts, np.uint32)
        self.neighbours = np.zeros(self.num_neighbours, np.int32); self.n_loc_of = np.zeros(world, np.int32)
        self.push_src = np.zeros(self.num_push, np.uint32); self.push_dst = np.zeros(self.num_push, np.uint32)
        self.push_rank = np.zeros(self.num_push, np.int32)
        _check(lib().pd_rank_plan_get(self._h, _p(self.tiles), _p(self.ghosts), _p(self.neighbours), _p(self.n_loc_of),
                                      _p(self.push_src), _p(self.push_dst), _p(self.push_rank)))

    def local_layout(self, global_layout):
        h = lib().pd_rank_layout(global_layout._h, self._h)
        if not h:
            raise PdError(_err())
        nT = int(np.diff(global_layout.tile_tet_start.astype(np.int64))[self.tiles].sum())
        return Layout(None, handle=h, counts=(self.num_owned + self.num_ghosts, nT))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.pd_rank_plan_free(self._h)
            self._h = None


def dist_connect_local(engines):
    """All ranks of a multi-GPU engine inside ONE process (tests): wire the exchange windows directly."""
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().pd_dist_connect_local(arr, len(engines)))


def dist_step_lockstep(engines, n_steps=1):
    """One host thread drives all ranks phase by phase (see pd_dist_step_lockstep)."""
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().pd_dist_step_lockstep(arr, len(engines), n_steps))


class PdSolver:
    """PdSolver behind SimulationCUDAContext, on one B200 (pdSolver.h:12-44, simulationContext.h:17-84)."""

    SolverType = {"Jacobi": PD_JACOBI, "CuSolverCholesky": PD_CHOLESKY, "EigenCholesky": PD_CHOLESKY, "PCGJacobi": PD_PCG_JACOBI}

    def __init__(self, scene, device=0, rot_mode=-1, reorder=1, use_graph=1, ctas_per_sm=0, rank=0, world=1, body_kernel=-1):
        o = pd_engine_options(device, rot_mode, reorder, use_graph, ctas_per_sm, rank, world, body_kernel)
        self._h = lib().pd_create(scene._h, C.byref(o))
        if not self._h:
            raise PdError(_err())
        self.num_verts, self.num_tets = scene.counts()[:2]      # GLOBAL counts: host arrays always cover the whole scene
        self.rank, self.world = rank, world

    # --- multi-GPU plumbing (world > 1)
    def window_handle(self):
        h = np.zeros(64, np.uint8)
        _check(lib().pd_dist_window_handle(self._h, _p(h)))
        return h

    def connect(self, handles):
        """handles: (world, 64) uint8, rank order -- e.g. torch.distributed.all_gather of window_handle()"""
        h = np.ascontiguousarray(handles, np.uint8).reshape(self.world, 64)
        _check(lib().pd_dist_connect(self._h, _p(h)))

    def dist_status(self):
        s = C.c_uint()
        _check(lib().pd_dist_status(self._h, C.byref(s)))
        return s.value

    def dist_info(self):
        a = np.zeros(6, np.int32)
        _check(lib().pd_dist_info(self._h, _p(a)))
        return dict(zip(["num_owned", "num_ghosts", "num_neighbours", "num_push", "num_tets_local", "num_tiles_local"], [int(x) for x in a]))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.pd_destroy(self._h)
            self._h = None

    __del__ = close

    # --- Solver<float> interface
    def Update(self, n_steps=1, params=None):
        if params is not None:
            _check(lib().pd_set_params(self._h, C.byref(params.c)))
        _check(lib().pd_step(self._h, n_steps))

    def step_timed(self, n_steps):
        """n x Update bracketed by CUDA events on the engine's stream -> device milliseconds."""
        ms = C.c_float()
        _check(lib().pd_step_timed(self._h, n_steps, C.byref(ms)))
        return ms.value

    def Reset(self):
        _check(lib().pd_reset(self._h))

    def SetPerf(self, on):
        _check(lib().pd_set_perf(self._h, int(on)))

    def SetGlobalSolver(self, solver):
        _check(lib().pd_set_global_solver(self._h, self.SolverType.get(solver, solver)))

    def GetPerformanceData(self):
        p = pd_perf()
        _check(lib().pd_get_perf(self._h, C.byref(p)))
        return [("local step", p.local_step_ms), ("global step", p.global_step_ms),
                ("collision handling(fixed)", p.collision_fixed_ms), ("collision handling(mesh)", p.collision_mesh_ms)], p

    # --- state
    def synchronize(self):
        _check(lib().pd_synchronize(self._h))

    def set_params(self, params):
        _check(lib().pd_set_params(self._h, C.byref(params.c)))

    def get_params(self):
        p = SolverParams()
        _check(lib().pd_get_params(self._h, C.byref(p.c)))
        return p

    def download(self):
        X = np.zeros((self.num_verts, 3), np.float32); V = np.zeros_like(X); XT = np.zeros_like(X)
        _check(lib().pd_download(self._h, _p(X), _p(V), _p(XT)))
        return X, V, XT

    def upload(self, X=None, V=None, XTilde=None):
        a = [None if t is None else np.ascontiguousarray(t, np.float32) for t in (X, V, XTilde)]
        _check(lib().pd_upload_state(self._h, *[_p(t) for t in a]))

    def step_host_ptr(self, n, xin, vin, xtin, xout, vout, xtout):
        """e2e step on raw host pointers (ints), e.g. pinned buffers."""
        _check(lib().pd_step_host(self._h, n, xin, vin, xtin, xout, vout, xtout))

    def step_host_owned_ptr(self, n, xin, vin, xtin, xout, vout, xtout):
        """e2e step on this rank's SHARD (3 * num_owned floats per array, local owned order; see owned_ids)."""
        _check(lib().pd_step_host_owned(self._h, n, xin, vin, xtin, xout, vout, xtout))

    def owned_ids(self):
        """original vertex id of each owned vertex of this rank, in local order"""
        n = self.dist_info()["num_owned"]
        a = np.zeros(n, np.uint32)
        _check(lib().pd_dist_owned_ids(self._h, _p(a)))
        return a

    def update_device_ptr(self, n, dX, dV, dXT):
        _check(lib().pd_update_device(self._h, n, dX, dV, dXT))

    # --- mouse-drag soft constraints (SolverData::moreDBC / OffsetX / mouseSelection.target, def.h:14-18,31-32)
    def set_drag(self, more_dbc=None, offset_x=None, target=(0.0, 0.0, 0.0)):
        """more_dbc None ends the drag (SimulationCUDAContext::ResetMoreDBC(true), simulationContext.cu:220-226)."""
        if more_dbc is None:
            _check(lib().pd_set_drag(self._h, None, None, None))
            return
        m = np.ascontiguousarray(more_dbc, np.float32).reshape(self.num_verts)
        o = np.ascontiguousarray(offset_x, np.float32).reshape(self.num_verts, 3)
        t = np.ascontiguousarray(target, np.float32).reshape(3)
        _check(lib().pd_set_drag(self._h, _p(m), _p(o), _p(t)))

    def set_drag_device_ptr(self, d_more_dbc, d_offset_x, target):
        t = np.ascontiguousarray(target, np.float32).reshape(3)
        _check(lib().pd_set_drag_device(self._h, d_more_dbc, d_offset_x, _p(t)))

    def drag_select(self, select_v, target, control_mag=10.0):
        """Control_Kernel (simulationContext.cu:202-218) on the engine's current X; select_v = -1 clears."""
        t = np.ascontiguousarray(target, np.float32).reshape(3)
        _check(lib().pd_drag_select(self._h, int(select_v), float(control_mag), _p(t)))

    def get_drag(self):
        m = np.zeros(self.num_verts, np.float32); o = np.zeros((self.num_verts, 3), np.float32); d = np.zeros((self.num_verts, 3), np.float32)
        a = C.c_int()
        _check(lib().pd_get_drag(self._h, _p(m), _p(o), _p(d), C.byref(a)))
        return m, o, d, bool(a.value)

    def update_mu(self, mu):
        """SimulationCUDAContext::UpdateSoftBodyAttr (simulationContext.cu:165-176): new per-tet mu, original tet order."""
        m = np.ascontiguousarray(mu, np.float32).reshape(self.num_tets)
        _check(lib().pd_update_mu(self._h, _p(m)))

    def setup(self):
        md = np.zeros(self.num_verts, np.float32); c = np.zeros(self.num_verts, np.float32)
        B = np.zeros((self.num_tets, 9), np.float32); V0 = np.zeros(self.num_tets, np.float32)
        _check(lib().pd_get_setup(self._h, _p(md), _p(c), _p(B), _p(V0)))
        return md, c, B.reshape(-1, 3, 3), V0

    def system_matrix(self):
        """A^ as (rowptr, col, val) in the engine's renumbered vertex ids (scipy.sparse.csr_matrix-ready)."""
        nnz = C.c_int()
        _check(lib().pd_get_system_matrix(self._h, C.byref(nnz), None, None, None))
        n = self.dist_info()["num_owned"] + self.dist_info()["num_ghosts"]
        rp = np.zeros(n + 1, np.int32); col = np.zeros(nnz.value, np.int32); val = np.zeros(nnz.value, np.float32)
        _check(lib().pd_get_system_matrix(self._h, C.byref(nnz), _p(rp), _p(col), _p(val)))
        return rp, col, val

    def collision(self):
        """(tI [nV], normals [nV, 3], overlapping triangle pairs) of the last mesh-mesh collision pass."""
        nV = self.num_verts
        tI = np.ones(nV, np.float32); nor = np.zeros((nV, 3), np.float32); n = C.c_longlong()
        _check(lib().pd_get_collision(self._h, _p(tI), _p(nor), C.byref(n)))
        return tI, nor, n.value

    def solver_sizes(self):
        a, l = C.c_longlong(), C.c_longlong()
        _check(lib().pd_get_solver_sizes(self._h, C.byref(a), C.byref(l)))
        return dict(nnz_A=a.value, nnz_L=l.value)

    def solve_stats(self):
        err, it = C.c_float(), C.c_int()
        _check(lib().pd_get_solve_stats(self._h, C.byref(err), C.byref(it)))
        return err.value, it.value

    def time_kernels(self, reps=20):
        a, b = C.c_float(), C.c_float()
        _check(lib().pd_time_kernels(self._h, reps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile_local(self):
        """clock64 totals per phase of the local kernel: array (local_grid, 8), see include/pd_b200.h"""
        out = np.zeros((self.info()["local_grid"], 8), np.uint64)
        _check(lib().pd_profile_local(self._h, _p(out)))
        return out

    def info(self):
        nv, nt, ntl, lg = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ns = C.c_uint32(); sb, db = C.c_size_t(), C.c_size_t()
        _check(lib().pd_engine_info(self._h, C.byref(nv), C.byref(nt), C.byref(ntl), C.byref(ns), C.byref(sb), C.byref(db), C.byref(lg)))
        return dict(num_verts=nv.value, num_tets=nt.value, num_tiles=ntl.value, num_slots=ns.value,
                    tile_stream_bytes=sb.value, device_bytes=db.value, local_grid=lg.value, rot_mode=lib().pd_engine_rot_mode(self._h))


def nested_dissection(rowptr, col, xyz, count_fill=True):
    """Host-side fill-reducing order of the Cholesky path: (perm[new] = old, nnz(L) in the given order, nnz(L) reordered)."""
    rowptr = np.ascontiguousarray(rowptr, np.int32); col = np.ascontiguousarray(col, np.int32); xyz = np.ascontiguousarray(xyz, np.float32)
    n = rowptr.shape[0] - 1
    perm = np.zeros(n, np.int32); a, b = C.c_int(-1), C.c_int(-1)
    _check(lib().pd_nested_dissection(n, _p(rowptr), _p(col), _p(xyz), _p(perm), C.byref(a) if count_fill else None, C.byref(b) if count_fill else None))
    return perm, a.value, b.value


def cholesky_factor(rowptr, col, val):
    """Host-side sparse Cholesky (no GPU): returns L by rows (rowptr, col, val), diagonal last in every row."""
    rowptr = np.ascontiguousarray(rowptr, np.int32); col = np.ascontiguousarray(col, np.int32); val = np.ascontiguousarray(val, np.float32)
    n = rowptr.shape[0] - 1
    nnz = C.c_int(); lp, lc, lv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _check(lib().pd_cholesky_factor(n, _p(rowptr), _p(col), _p(val), C.byref(nnz), C.byref(lp), C.byref(lc), C.byref(lv)))
    out = (np.ctypeslib.as_array(C.cast(lp, C.POINTER(C.c_int)), shape=(n + 1,)).copy(),
           np.ctypeslib.as_array(C.cast(lc, C.POINTER(C.c_int)), shape=(nnz.value,)).copy(),
           np.ctypeslib.as_array(C.cast(lv, C.POINTER(C.c_float)), shape=(nnz.value,)).copy())
    for q in (lp, lc, lv):
        lib().pd_free(q)
    return out


PD_LS_CG_IC0, PD_LS_PCG_JACOBI = 1, 2


class LinearSolver:
    """LinearSolver<double> of the reference's IPC solver (linear.h:55-71): kind PD_LS_PCG_JACOBI or PD_LS_CG_IC0."""

    def __init__(self, kind, n, max_iter=0, tolerance=0.0, device=0):
        self.n = n
        self._h = lib().pd_linsolver_create(kind, n, max_iter, tolerance, device)
        if not self._h:
            raise PdError(_err())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pd_linsolver_destroy(self._h)
            self._h = None

    def solve(self, A, row, col, b, guess=None):
        """host COO (duplicates allowed) -> x"""
        A = np.ascontiguousarray(A, np.float64); row = np.ascontiguousarray(row, np.int32); col = np.ascontiguousarray(col, np.int32)
        b = np.ascontiguousarray(b, np.float64); g = None if guess is None else np.ascontiguousarray(guess, np.float64)
        x = np.zeros(self.n, np.float64)
        _check(lib().pd_linsolver_solve_host(self._h, self.n, _p(b), _p(x), _p(A), A.shape[0], _p(row), _p(col), _p(g)))
        return x

    def solve_device_ptr(self, d_b, d_x, d_A, nz, d_row, d_col, d_guess=None):
        _check(lib().pd_linsolver_solve_device(self._h, self.n, d_b, d_x, d_A, nz, d_row, d_col, d_guess))

    def stats(self):
        it, nnz, res = C.c_int(), C.c_int(), C.c_double()
        _check(lib().pd_linsolver_stats(self._h, C.byref(it), C.byref(res), C.byref(nnz)))
        return dict(iterations=it.value, residual=res.value, nnz=nnz.value)


def ccd_batch(types, verts, X, XTilde, device=0):
    """The collision pass's continuous-collision test on n queries -> (toi [n], normals [n, 3])."""
    t = np.ascontiguousarray(types, np.int32); v = np.ascontiguousarray(verts, np.uint32).reshape(-1, 4)
    X = np.ascontiguousarray(X, np.float32); XT = np.ascontiguousarray(XTilde, np.float32)
    toi = np.zeros(t.shape[0], np.float32); nor = np.zeros((t.shape[0], 3), np.float32)
    _check(lib().pd_ccd_batch(device, t.shape[0], _p(t), _p(v), X.shape[0], _p(X), _p(XT), _p(toi), _p(nor)))
    return toi, nor


def rotation_batch(F, rot_mode=0, device=0):
    """Corotational projection of a batch of 3x3 matrices on the GPU (test hook)."""
    F = np.ascontiguousarray(F, np.float32).reshape(-1, 9)
    R = np.zeros_like(F); used = np.zeros(F.shape[0], np.int32)
    _check(lib().pd_rotation_batch(device, rot_mode, F.shape[0], _p(F), _p(R), _p(used)))
    return R.reshape(-1, 3, 3), used
