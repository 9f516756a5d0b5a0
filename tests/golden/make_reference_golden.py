"""Generates the golden outputs of the REFERENCE's own CUDA kernels on a B200.

Run on the GPU box (needs oracle/_ref/libpd_ref.so, built here from /root/reference by oracle/Makefile):
    gpurun -- 'python tests/golden/make_reference_golden.py gpurun_out/reference_b200.npz'
then copy gpurun_out/reference_b200.npz to tests/golden/.  Scenes = the float contexts of
tests/meshes.py, 100 steps each, launch-for-launch replay of
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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import _fixed_arrays
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        assets = meshes.write_assets(tmp)
        for ctx, steps in [("C1 cube", 100), ("C5 house&sphere", 100), ("C2 armadillo&bunny", 100)]:
            key = ctx.split()[0]
            sc = pd.Scene.from_json(assets["json"], ctx)
            p = sc.params
            if key == "C1":
                p["dt"] = 1 / 60
            a = sc.arrays()
            planes, spheres, cyls = _fixed_arrays(pd, a["fixed"])
            runs = []
            for rep in range(2):
                rs = ref.RefScene(a["X"], a["Tet"], a["mass"], a["mu"], planes=planes, spheres=spheres, cylinders=cyls)
                traj = []
                for s in range(steps // 10):
                    rs.step(10, dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=p["num_iterations"])
                    traj.append(rs.get()[0].copy())
                runs.append((rs.get(), traj))
            (X, V, XT), traj = runs[0]
            spread = meshes.rel_err(runs[1][0][0], X, float(np.linalg.norm(a["X"].max(0) - a["X"].min(0))))
            out[key + "_steps"] = np.int32(steps)
            out[key + "_X"] = X; out[key + "_V"] = V; out[key + "_XTilde"] = XT
            out[key + "_X_every10"] = np.stack(traj)[:, :: max(1, X.shape[0] // 512)]   # subsampled trajectory
            out[key + "_run_to_run_rel"] = np.float64(spread)
            print(ctx, "run-to-run spread of the reference:", spread, "min y", XT[:, 1].min())
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)
