"""The CPU oracle (oracle/pd_oracle.c) against the committed golden outputs of the REFERENCE's own CUDA kernels on a B200
(tests/golden/reference_b200.npz, generator tests/golden/make_reference_golden.py).  The reference's tests hold no vectors
for the PD path, so these fixtures -- outputs of pdUtil.cu / svd3_cuda.h / computeInvDmV0 compiled
verbatim and replayed launch for launch -- are the pin.  CPU only.

Bars: the 6-tet cube is one warp in the reference (deterministic): the oracle matches it bit for bit through the free fall
and stays within 2e-5 after the impact (where the reference's float atomics meet in a different order than a sequential
scatter); the armadillo is compared within max(1e-4, 10 x the reference's own run-to-run spread)."""
import os

import numpy as np
import pytest

import meshes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "reference_b200.npz"))


def _oracle(O, assets, name, **kw):
    sc, p = meshes.oracle_scene(O, assets, name)
    p.update(kw)
    return sc, O.make_params(dt=p["dt"], gravity=p["gravity"], muN=p["muN"], muT=p["muT"], num_iterations=p["num_iterations"],
                             threads=min(8, os.cpu_count() or 1))


def test_c1_cube_100_steps_vs_reference_golden(O, assets, golden):
    sc, op = _oracle(O, assets, "C1 cube", dt=1 / 60)
    sc.step(op, int(golden["C1_steps"]))
    X, V, XT = sc.get()
    scale = float(np.linalg.norm(sc.X0.max(0) - sc.X0.min(0)))
    ex, ext = meshes.rel_err(X, golden["C1_X"], scale), meshes.rel_err(XT, golden["C1_XTilde"], scale)
    ev = float(np.abs(V - golden["C1_V"]).max())
    print(f"oracle vs reference golden, C1 cube after {int(golden['C1_steps'])} steps: X {ex:.2e} XTilde {ext:.2e} |dV| {ev:.2e}")
    assert ex <= 2e-5 and ext <= 2e-5
    assert XT[:, 1].min() > -1e-3                      # at rest on the floor, like the reference


def test_armadillo_10_steps_vs_reference_golden(O, assets, golden):
    sc, op = _oracle(O, assets, "C2 armadillo")
    sc.step(op, 10)
    scale = float(np.linalg.norm(sc.X0.max(0) - sc.X0.min(0)))
    e10 = meshes.rel_err(sc.get()[2], golden["C2a_XTilde_10"], scale)
    print(f"oracle vs reference golden, armadillo after 10 steps: {e10:.2e} (reference run-to-run spread {float(golden['C2a_spread_10']):.2e})")
    assert e10 <= max(1e-4, 10 * float(golden["C2a_spread_10"]))
