"""Mouse-drag soft constraints in the CPU oracle: the branches of
pdUtil.cu:56-69 (setMDt_2MoreDBC), :80-87 (computeSn), :159-164 (computeDBCLocal), :187-188 (updateVelPos) and
:201-206 (getErrorKern) plus Control_Kernel (simulationContext.cu:202-218).  The reference's tests hold no vectors for
these either; the properties below are what the reference's code guarantees by construction.  CPU only."""
import numpy as np

import meshes


def _scene(O, assets, name="C5 house&sphere"):
    sc, p = meshes.oracle_scene(O, assets, name)
    return sc, O.make_params(dt=p["dt"], gravity=p["gravity"], muN=p["muN"], muT=p["muT"], num_iterations=30, threads=4)


def _ball(X, v, r):
    off = (X - X[v]).astype(np.float32)
    more = np.where((off.astype(np.float64) ** 2).sum(1) < r * r, np.float32(10.0), np.float32(0.0)).astype(np.float32)
    return more, off


def test_dragged_vertices_are_held_at_target_plus_offset_with_zero_velocity(O, assets):
    sc, op = _scene(O, assets)
    sc.step(op, 2)
    X = sc.get()[0]
    v = 17
    more, off = _ball(X, v, 8.0)
    assert 1 < (more > 0).sum() < sc.nV // 2
    for k in range(3):
        target = (X[v] + np.float32([0.5 * (k + 1), 0.25 * (k + 1), 0.0])).astype(np.float32)
        sc.set_drag(more, off, target)
        sc.step(op, 1)
        Xn, Vn, XTn = sc.get()
        held = more > 0
        want = (target[None, :] + off[held]).astype(np.float32)             # one float add per component (pdUtil.cu:81)
        assert np.array_equal(XTn[held].view(np.uint32), want.view(np.uint32))
        assert np.array_equal(Xn[held].view(np.uint32), want.view(np.uint32))
        assert not Vn[held].any()                                            # updateVelPos: vel = 0
        assert np.array_equal(sc.get_drag()[2][held].view(np.uint32), want.view(np.uint32))   # computeSn also overwrites DBCX
        assert np.isfinite(Xn).all() and np.abs(Vn[~held]).max() > 0
    # release: ResetMoreDBC(true); the vertices fall again
    sc.set_drag(None)
    sc.step(op, 1)
    assert np.abs(sc.get()[1][more > 0, 1]).min() > 0


def test_zero_more_dbc_is_the_plain_step_bit_for_bit(O, assets):
    a, op = _scene(O, assets)
    b, _ = _scene(O, assets)
    b.set_drag(np.zeros(b.nV, np.float32), np.ones((b.nV, 3), np.float32), (1.0, 2.0, 3.0))
    a.step(op, 3); b.step(op, 3)
    for x, y in zip(a.get(), b.get()):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_control_kernel_selects_the_ball_around_the_picked_vertex(O, assets):
    sc, op = _scene(O, assets, "C1 cube")
    sc.step(op, 1)
    X = sc.get()[0]
    sc.drag_select(3, (0.0, 31.0, 0.0))
    more, off, _ = sc.get_drag()
    assert np.array_equal(off, (X - X[3]).astype(np.float32))
    d2 = (off.astype(np.float64) ** 2).sum(1)
    assert np.array_equal(more > 0, d2 < 0.002) and more[3] == np.float32(10.0)     # RADIUS_SQUARED, control_mag (simulationContext.cu:18,229)
    sc.step(op, 1)
    assert np.array_equal(sc.get()[2][3], np.float32([0.0, 31.0, 0.0]))
    sc.drag_select(-1, (0.0, 0.0, 0.0))                                                # select_v == -1: nothing is held
    assert not sc.get_drag()[0].any()
    sc.reset()                                                                          # Reset zeroes moreDBC (simulationContext.cu:240)
    sc.drag_select(3, (0.0, 31.0, 0.0)); sc.reset()
    assert not sc.get_drag()[0].any()


def test_dragged_mass_term_enters_the_direct_modes_right_hand_side_only(O, assets):
    """PCG / Cholesky branch: the matrix was assembled at prepare time (setMDt_2), so a drag changes b0 = (m + w)/h^2 s and
    the velocity of the held vertices, but the held vertices are NOT hard constraints there (the reference's behaviour)."""
    sc, p = meshes.oracle_scene(O, assets, "C1 cube")
    op = O.make_params(dt=1 / 60, gravity=p["gravity"], num_iterations=10, global_solver=2, tol=1e-6, pcg_tol=1e-6, threads=1)
    sc.step(op, 1)
    X = sc.get()[0]
    more = np.zeros(sc.nV, np.float32); more[5] = 10.0
    target = (X[5] + np.float32([0.2, 0.1, 0.0])).astype(np.float32)
    sc.set_drag(more, (X - X[5]).astype(np.float32), target)
    sc.step(op, 1)
    Xn, Vn, _ = sc.get()
    assert not Vn[5].any() and np.isfinite(Xn).all()
    assert np.abs(Xn[5] - target).max() > 1e-4          # pulled towards the target, not pinned to it


def test_live_mu_edit_acts_at_once_but_setup_products_wait_for_reset(O, assets):
    """UpdateSoftBodyAttr (simulationContext.cu:165-176) overwrites SolverData::mu; computeLocal reads it in every iteration
    while matrix_diag was computed by SolverPrepare: stale until Reset, exactly like the reference."""
    a, op = _scene(O, assets)
    b, _ = _scene(O, assets)
    a.step(op, 2); b.step(op, 2)
    md0 = b.setup(op)[0].copy()
    # a SOFTER material: with the stale (larger) diagonal the Jacobi sweeps stay stable.  Raising mu instead makes them
    # diverge within a step -- the reference's latent bug, restated as it is.
    mu2 = np.full(b.nT, 1.0e5, np.float32)
    b.set_mu(mu2)
    assert np.array_equal(b.setup(op)[0], md0)                       # stale on purpose
    a.step(op, 2); b.step(op, 2)
    assert np.isfinite(b.get()[0]).all() and np.abs(a.get()[0] - b.get()[0]).max() > 1e-6     # acts at once
    b.reset()
    assert not np.array_equal(b.setup(op)[0], md0)                   # re-prepared with the new mu
    c = O.Scene(b.X0, b.Tet, 10.0, mu2, planes=[])                    # same mesh built with the new mu from the start
    assert np.allclose(b.setup(op)[0], c.setup(op)[0], rtol=0, atol=0)
