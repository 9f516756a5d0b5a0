"""The SOURCE of the product kernels (csrc/pd_kernels.cuh, rotation.cuh: predictor, local step with its staging / H scratch /
ordered partial sums, Jacobi-Chebyshev vertex kernel, finish + fixed bodies, the mouse-drag instantiations) compiled for the
HOST (tests/emu: -DPD_HOST_EMU, one OS thread per CUDA thread, real barriers) and run on the device layout the product's
layout.cpp builds, against the CPU oracle.  This is not the GPU parity suite (`-m gpu` runs the real thing through the C ABI);
it checks the kernels' logic -- indexing, staging, summation order, arithmetic forms -- on every CPU run, and it is how the
experiment PD_DIST_TRIM (run time) was checked in a session without GPU time.
TEST INFRASTRUCTURE: nothing of tests/emu is linked into the product library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import meshes

pytestmark = pytest.mark.timeout(1200)        # 256 OS threads per emulated CTA: a scheduling problem must not hang the suite
HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
_libs = {}


def emu_lib(variant="default"):
    if variant not in _libs:
        name = "libpd_emu.so" if variant == "default" else f"libpd_emu_{variant}.so"
        subprocess.check_call(["make", "-C", EMU, name], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(EMU, name))
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_step.argtypes = [C.c_void_p] + [C.c_float] * 5 + [C.c_int, C.c_int]
        L.emu_step_concurrent.argtypes = [C.c_void_p] + [C.c_float] * 5 + [C.c_int, C.c_int]
        L.emu_get.argtypes = [C.c_void_p] * 4
        L.emu_set.argtypes = [C.c_void_p] * 4
        L.emu_set_drag.argtypes = [C.c_void_p] * 4
        L.emu_info.argtypes = [C.c_void_p] * 4
        L.emu_drag_select.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.emu_get_drag.argtypes = [C.c_void_p] * 3
        L.emu_enable_body_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.emu_variant.restype = C.c_char_p
        assert L.emu_variant().decode() == variant
        _libs[variant] = L
    return _libs[variant]


class Emu:
    def __init__(self, pd, sc, rot_mode=0, reorder=1, world=1, trim=0, grid=3, variant="default"):
        from test_gpu_parity import _fixed_arrays
        self.lib = emu_lib(variant)
        a = sc.arrays()
        planes, spheres, cyls = _fixed_arrays(pd, a["fixed"])
        f = lambda rows: np.ascontiguousarray(np.array([np.concatenate([np.ravel(x) for x in r]) for r in rows], np.float32).reshape(-1))
        pl, sp, cy = f(planes), f(spheres), f(cyls)
        self.nV = a["X"].shape[0]
        self.p = sc.params
        self._keep = (a, pl, sp, cy)
        p = lambda x: x.ctypes.data if x.size else None
        self.h = self.lib.emu_create(self.nV, a["Tet"].shape[0], p(a["X"]), p(a["Tet"]), p(a["mass"]), p(a["mu"]), p(a["DBC"]),
                                     len(planes), p(pl), len(spheres), p(sp), len(cyls), p(cy), rot_mode, reorder, world, trim, grid)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.emu_destroy(self.h)
            self.h = None

    def enable_body_kernel(self):
        """PD_BODY_KERNEL experiment: one emulated CTA per soft body, the whole step in one launch."""
        a = self._keep[0]
        bvs = np.ascontiguousarray(a["body_vert_start"], np.int32)
        assert self.lib.emu_enable_body_kernel(self.h, self.nV, a["Tet"].shape[0], a["X"].ctypes.data, a["Tet"].ctypes.data, a["mu"].ctypes.data,
                                               len(bvs), bvs.ctypes.data) == 0

    def step(self, n=1):
        p = self.p
        rc = self.lib.emu_step(self.h, p["dt"], p["gravity"], p["rho"], p["muN"], p["muT"], p["num_iterations"], n)
        if rc == 77:
            pytest.skip("this machine cannot create the OS threads the emulation needs (one per CUDA thread of a CTA)")
        assert rc == 0

    def step_concurrent(self, n=1):
        """world > 1: every rank on its own thread, the halo exchange inside the local kernel (flags, epochs, tickets)."""
        p = self.p
        rc = self.lib.emu_step_concurrent(self.h, p["dt"], p["gravity"], p["rho"], p["muN"], p["muT"], p["num_iterations"], n)
        if rc == 77:
            pytest.skip("this machine cannot create the OS threads the emulation needs")
        assert rc == 0, "a halo wait gave up" if rc == 2 else rc

    def get(self):
        X = np.zeros((self.nV, 3), np.float32); V = np.zeros_like(X); XT = np.zeros_like(X)
        self.lib.emu_get(self.h, X.ctypes.data, V.ctypes.data, XT.ctypes.data)
        return X, V, XT

    def set(self, X=None, V=None, XTilde=None):
        a = [None if t is None else np.ascontiguousarray(t, np.float32) for t in (X, V, XTilde)]
        self.lib.emu_set(self.h, *[None if t is None else t.ctypes.data for t in a])

    def set_drag(self, more=None, off=None, target=(0, 0, 0)):
        if more is None:
            assert self.lib.emu_set_drag(self.h, None, None, None) == 0
            return
        m = np.ascontiguousarray(more, np.float32); o = np.ascontiguousarray(off, np.float32); t = np.ascontiguousarray(target, np.float32)
        assert self.lib.emu_set_drag(self.h, m.ctypes.data, o.ctypes.data, t.ctypes.data) == 0

    def drag_select(self, v, target, mag=10.0):
        t = np.ascontiguousarray(target, np.float32)
        assert self.lib.emu_drag_select(self.h, int(v), float(mag), t.ctypes.data) == 0

    def get_drag(self):
        m = np.zeros(self.nV, np.float32); o = np.zeros((self.nV, 3), np.float32)
        self.lib.emu_get_drag(self.h, m.ctypes.data, o.ctypes.data)
        return m, o

    def info(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self.lib.emu_info(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(tets=a.value, tiles=b.value, ghosts=c.value)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _oparams(O, p, **kw):
    d = dict(dt=p["dt"], gravity=p["gravity"], muN=p["muN"], muT=p["muT"], rho=p["rho"], tol=p["tol"], num_iterations=p["num_iterations"],
             threads=min(8, os.cpu_count() or 1))
    d.update(kw)
    return O.make_params(**d)


def _scene(pd, assets, name, iters, **kw):
    sc = pd.Scene.from_json(assets["json"], name)
    p = sc.params
    p["num_iterations"] = iters
    for k, v in kw.items():
        p[k] = v
    sc.params = p
    return sc, p


@pytest.mark.parametrize("variant", ["default"])
def test_c1_cube_faithful_kernels_are_bit_exact_vs_oracle(pd, O, assets, variant):
    """Free fall, impact on the floor plane (step ~42 at 100 iterations per step) and rest, like the GPU suite's first test."""
    sc, p = _scene(pd, assets, "C1 cube", 100, dt=1 / 60)
    osc, _ = meshes.oracle_scene(O, assets, "C1 cube")
    op = _oparams(O, p)
    emu = Emu(pd, sc, rot_mode=1, reorder=0, variant=variant)
    chunks = 9 if variant == "default" else 2                       # the variant only changes the H scratch: no need to wait for the impact again
    for n in range(chunks):
        emu.step(5); osc.step(op, 5)
        for a, b in zip(emu.get(), osc.get()):
            assert np.array_equal(_bits(a), _bits(b)), f"step {5 * (n + 1)}"
    if variant == "default":
        assert np.abs(emu.get()[1][:, 1]).max() < 40.0 and emu.get()[2][:, 1].min() > -1e-3     # hit the floor plane (free fall would be at 73)


@pytest.mark.parametrize("variant,rot_mode,tol", [("default", 1, 1e-4), ("default", 0, 1e-4)])
def test_house_and_sphere_kernels_vs_oracle(pd, O, assets, variant, rot_mode, tol):
    """11 tiles over 3 emulated CTAs (several tiles per CTA: prologue, steady state and tail of the software pipeline),
    two bodies, fixed sphere + planes.  Faithful mode differs from the oracle only by the order of the per-tile partial
    sums; the default mode also by the Newton polar rotation."""
    sc, p = _scene(pd, assets, "C5 house&sphere", 30)
    osc, _ = meshes.oracle_scene(O, assets, "C5 house&sphere")
    op = _oparams(O, p)
    emu = Emu(pd, sc, rot_mode=rot_mode, variant=variant)
    scale = float(np.linalg.norm(osc.X0.max(0) - osc.X0.min(0)))
    emu.step(2); osc.step(op, 2)
    err = max(meshes.rel_err(a, b, scale) for a, b in zip(emu.get()[::2], osc.get()[::2]))
    print(f"emulated kernels ({variant}, rot_mode {rot_mode}) vs oracle, house+sphere, 2 steps from rest: {err:.2e}")
    assert err <= tol, err
    # (this context amplifies last-bit differences a thousandfold per step at 30 sweeps per step -- 1e-7, 1e-6, 1e-3 after
    #  steps 1, 2, 3, in every mode alike: the reference's Chebyshev-Jacobi iteration on the soft sphere, not the kernels)


@pytest.mark.parametrize("rot_mode", [1, 0])
def test_drag_kernels_vs_oracle(pd, O, rot_mode):
    """The DRAG instantiations of k_predict / k_vertex_jacobi / k_finish on the 6^3-cell grid of the GPU drag tests
    (tests/test_gpu_zz_drag.py): held vertices bit-exact, the rest within 2e-5 of the oracle."""
    from test_gpu_solvers import _grid, _oracle_of
    sc = _grid(pd)
    sc.params = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=50)
    p = sc.params
    osc = _oracle_of(O, sc)
    op = _oparams(O, p)
    emu = Emu(pd, sc, rot_mode=rot_mode)
    emu.step(2); osc.step(op, 2)
    X = osc.get()[0]
    pick = X.shape[0] // 2
    off = (X - X[pick]).astype(np.float32)
    more = np.where((off.astype(np.float64) ** 2).sum(1) < 1.44, np.float32(10), np.float32(0)).astype(np.float32)
    held = more > 0
    X0 = sc.arrays()["X"]
    scale = float(np.linalg.norm(X0.max(0) - X0.min(0)))
    worst = 0.0
    for k in range(4):
        target = (X[pick] + np.float32([0.15 * (k + 1), 0.1 * (k + 1), 0.0])).astype(np.float32)
        emu.set_drag(more, off, target); osc.set_drag(more, off, target)
        emu.step(1); osc.step(op, 1)
        Xe, Ve, XTe = emu.get()
        want = (target[None, :] + off[held]).astype(np.float32)
        assert np.array_equal(_bits(XTe[held]), _bits(want)) and not Ve[held].any()
        worst = max(worst, meshes.rel_err(Xe, osc.get()[0], scale))
    emu.set_drag(None); osc.set_drag(None)
    emu.step(2); osc.step(op, 2)
    worst = max(worst, meshes.rel_err(emu.get()[0], osc.get()[0], scale))
    print(f"emulated drag kernels, rot_mode {rot_mode}: worst rel err vs oracle {worst:.2e}")
    assert worst <= 2e-5 and np.abs(emu.get()[1][held, 1]).min() > 0


def test_cube_corner_dragged_faithful_kernels_bit_exact_vs_oracle(pd, O, assets):
    """What tests/test_gpu_zz_drag.py::test_control_kernel_and_reset_vs_oracle expects of the GPU: with a corner of the cube
    held by Control_Kernel's arrays, the faithful kernels still match the oracle bit for bit."""
    sc, p = _scene(pd, assets, "C1 cube", 100, dt=1 / 60)
    osc, _ = meshes.oracle_scene(O, assets, "C1 cube")
    op = _oparams(O, p)
    emu = Emu(pd, sc, rot_mode=1, reorder=0)
    emu.step(2); osc.step(op, 2)
    target = np.float32([0.5, 31.0, 0.25])
    osc.drag_select(3, target)
    more, off, _ = osc.get_drag()
    emu.drag_select(3, target)                                       # k_drag_select = Control_Kernel on the kernels' own X
    me, oe = emu.get_drag()
    assert np.array_equal(me, more) and np.array_equal(_bits(oe), _bits(off)) and me[3] == np.float32(10)
    emu.step(2); osc.step(op, 2)
    for a, b in zip(emu.get(), osc.get()):
        assert np.array_equal(_bits(a), _bits(b))
    assert np.array_equal(emu.get()[2][3], target)
    emu.set_drag(None); osc.set_drag(None)
    emu.step(2); osc.step(op, 2)
    for a, b in zip(emu.get(), osc.get()):
        assert np.array_equal(_bits(a), _bits(b))


@pytest.mark.parametrize("world,trim", [(2, 0), (3, 0), (3, 1), (8, 1)])
def test_partitioned_mesh_is_bit_identical_to_one_rank(pd, world, trim):
    """Multi-GPU logic on the host: `world` ranks in lock step (the halo push is a copy along the plans' push lists).  Every
    rank evaluates unchanged global tiles (trim 0) or its trimmed, re-packed boundary tiles (PD_DIST_TRIM experiment): the
    combined state must equal the single-rank run bit for bit."""
    sc = pd.Scene.kuhn_grid(7, 6, 5, 1.0, 0.05, 5, (0, 3, 0), 1.0, 2e5)
    sc.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
    sc.params = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=15)
    X0 = sc.arrays()["X"]
    V0 = np.zeros_like(X0); V0[:, 1] = 0.3 * np.sin(X0[:, 0])
    one = Emu(pd, sc)
    many = Emu(pd, sc, world=world, trim=trim)
    one.set(V=V0); many.set(V=V0)
    one.step(3); many.step(3)
    for a, b in zip(one.get(), many.get()):
        assert np.array_equal(_bits(a), _bits(b))
    print(f"world {world} trim {trim}: {many.info()} (one rank: {one.info()})")
    assert np.abs(one.get()[0] - X0).max() > 1e-3


def test_per_body_kernel_faithful_is_bit_exact_vs_oracle(pd, O, assets):
    """PD_BODY_KERNEL experiment (csrc/pd_body_kernel.cuh): one CTA per body, predictor + all iterations + end of step in one
    launch.  Its sums run in the reference's sequential (tet, corner) order, so with the SVD rotation (rot_mode 1) it matches
    the oracle BIT FOR BIT on a two-body scene with several hundred vertices each -- which the tiled path only does on
    single-tile meshes."""
    sc, p = _scene(pd, assets, "C5 house&sphere", 25)
    osc, _ = meshes.oracle_scene(O, assets, "C5 house&sphere")
    op = _oparams(O, p)
    emu = Emu(pd, sc, rot_mode=1)
    emu.enable_body_kernel()
    for n in range(3):
        emu.step(1); osc.step(op, 1)
        for a, b in zip(emu.get(), osc.get()):
            assert np.array_equal(_bits(a), _bits(b)), f"step {n + 1}"


def test_per_body_kernel_default_mode_vs_tile_kernels(pd, assets):
    """... and in the default mode (Newton polar rotation) it differs from the tile kernels only by the order of the sums."""
    grids = []
    for i in range(3):
        g = pd.Scene.kuhn_grid(4 + i, 4, 3, 1.0, 0.05, 11 + i, (6.0 * i, 0.4, 0.0), 1.0, 2e5)
        g.params = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=30)
        g.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
        grids.append(g)
    sc = pd.Scene.merge(grids)
    X0 = sc.arrays()["X"]
    V0 = np.zeros_like(X0); V0[:, 1] = 0.3 * np.sin(X0[:, 0])
    a, b = Emu(pd, sc), Emu(pd, sc)
    b.enable_body_kernel()
    a.set(V=V0); b.set(V=V0)
    a.step(6); b.step(6)                                           # the grids reach the floor within these steps
    scale = float(np.linalg.norm(X0.max(0) - X0.min(0)))
    err = max(meshes.rel_err(x, y, scale) for x, y in zip(a.get()[::2], b.get()[::2]))
    print(f"per-body kernel vs tile kernels, three grids, 6 steps: {err:.2e}")
    assert err <= 2e-5 and np.abs(a.get()[0] - X0).max() > 1e-3


@pytest.mark.parametrize("world,trim", [(2, 0), (3, 1)])
def test_in_kernel_halo_exchange_is_bit_identical_to_one_rank(pd, world, trim):
    """The exchange as the multi-GPU engine runs it: every rank on its own thread, the boundary
    positions pushed into the neighbours' buffers by the CTAs of the local kernel itself, ticket, epoch, flags, and the wait
    only before the first boundary tile -- with the triple-buffered positions rotating across steps.  All CTAs of a launch
    are resident at once here, as on the GPU.  (Logic only: the host's memory model is stronger than the GPU's.)"""
    sc = pd.Scene.kuhn_grid(7, 6, 5, 1.0, 0.05, 5, (0, 3, 0), 1.0, 2e5)
    sc.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
    sc.params = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=10)
    X0 = sc.arrays()["X"]
    V0 = np.zeros_like(X0); V0[:, 1] = 0.3 * np.sin(X0[:, 0])
    one = Emu(pd, sc)
    many = Emu(pd, sc, world=world, trim=trim, grid=2)
    one.set(V=V0); many.set(V=V0)
    one.step(4); many.step_concurrent(4)            # 4 steps x 11 phases: the buffer rotation base visits 0, 2, 1, 0
    for a, b in zip(one.get(), many.get()):
        assert np.array_equal(_bits(a), _bits(b))
