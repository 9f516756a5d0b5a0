"""debug helper: a few PD steps of a small Kuhn grid (multi-tile) for compute-sanitizer runs"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pd = importlib.import_module("soft-body-simulation-cuda_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rot = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = pd.Scene.kuhn_grid(n, n, n, 1.0, 0.05, 1, (0, 5, 0), 1.0, 2e5)
p = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=4)
sc.params = p
eng = pd.PdSolver(sc, use_graph=0, rot_mode=rot)
eng.Update(2)
X, V, XT = eng.download()
print("ok", X[:2], eng.info())
