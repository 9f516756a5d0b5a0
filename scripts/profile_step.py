"""Runs a few PD steps of a bench workload with plain launches (no CUDA graph) so that ncu sees
every kernel: used for the launch list and the `--set full` captures under profiles/.
  python scripts/profile_step.py [workload] [steps] [iters]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    workload = sys.argv[1] if len(sys.argv) > 1 else "grid139"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    pd = importlib.import_module("soft-body-simulation-cuda_b200")
    sc, p = bench.make_scene(pd, workload)
    p["num_iterations"] = iters
    sc.params = p
    eng = pd.PdSolver(sc, use_graph=0)
    eng.upload(V=bench.initial_velocity(sc.arrays()["X"]))
    eng.Update(steps)
    eng.synchronize()
    print("done", sc.counts())
