"""Generates the golden outputs of the REFERENCE's own CUDA kernels on a B200.

Run on the GPU box (needs oracle/_ref/libpd_ref.so, built here from /root/reference by oracle/Makefile):
    gpurun -- 'python tests/golden/make_reference_golden.py gpurun_out/reference_b200.npz'
then copy gpurun_out/reference_b200.npz to tests/golden/.  Scenes = C1 cube (100 steps) and the armadillo of C2 (10 and 100 steps), launch-for-launch replay of
PdSolver::Update (oracle/ref_harness.cu).  Also records the reference's run-to-run spread (float
atomics reorder), which is the noise floor of any parity claim.
"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import meshes  # noqa: E402
import ref  # noqa: E402

if __name__ == "__main__":
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_b200.npz")
    pd = importlib.import_module("soft-body-simulation-cuda_b200")     # host-side scene loader only
    from test_gpu_parity import _fixed_arrays
    out = {}

    def ref_scene(a):
        planes, spheres, cyls = _fixed_arrays(pd, a["fixed"])
        return ref.RefScene(a["X"], a["Tet"], a["mass"], a["mu"], planes=planes, spheres=spheres, cylinders=cyls)

    with tempfile.TemporaryDirectory() as tmp:
        assets = meshes.write_assets(tmp)
        # C1: deterministic in the reference (6 tets, one warp)
        sc = pd.Scene.from_json(assets["json"], "C1 cube")
        p = sc.params; p["dt"] = 1 / 60
        a = sc.arrays()
        kw = dict(dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=p["num_iterations"])
        rs = ref_scene(a); rs.step(100, **kw)
        X, V, XT = rs.get()
        out.update(C1_steps=np.int32(100), C1_X=X, C1_V=V, C1_XTilde=XT)
        # armadillo: two runs, 10 and 100 steps, with the run-to-run spread (float atomics reorder)
        sc = pd.Scene.from_json(assets["json"], "C2 armadillo")
        p = sc.params
        a = sc.arrays()
        scale = float(np.linalg.norm(a["X"].max(0) - a["X"].min(0)))
        kw = dict(dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=p["num_iterations"])
        rsA, rsB = ref_scene(a), ref_scene(a)
        for n, tag in ((10, "10"), (90, "100")):
            rsA.step(n, **kw); rsB.step(n, **kw)
            XA, XB = rsA.get()[2], rsB.get()[2]
            out["C2a_XTilde_" + tag] = XA
            out["C2a_spread_" + tag] = np.float64(meshes.rel_err(XB, XA, scale))
            print("armadillo", tag, "steps: reference run-to-run spread", out["C2a_spread_" + tag])
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)
