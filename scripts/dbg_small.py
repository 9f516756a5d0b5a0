"""debug helper: a few PD steps of a small Kuhn grid (multi-tile) for compute-sanitizer runs, including the paths written
without GPU time: a mouse drag (DRAG kernel instantiations, pd_set_drag / pd_drag_select), a live mu edit and -- with
PD_BODY_KERNEL=1 in the environment -- the per-body kernel.
    compute-sanitizer --tool memcheck python scripts/dbg_small.py 6
    PD_BODY_KERNEL=1 compute-sanitizer --tool memcheck python scripts/dbg_small.py 6"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pd = importlib.import_module("soft-body-simulation-cuda_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rot = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = pd.Scene.kuhn_grid(n, n, n, 1.0, 0.05, 1, (0, 5, 0), 1.0, 2e5)
sc.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
p = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=4)
sc.params = p
eng = pd.PdSolver(sc, use_graph=0, rot_mode=rot)
eng.Update(2)
X, V, XT = eng.download()
pick = X.shape[0] // 2
eng.drag_select(pick, X[pick] + np.float32([0.2, 0.1, 0.0]))
eng.Update(2)
off = (X - X[pick]).astype(np.float32)
more = np.where((off.astype(np.float64) ** 2).sum(1) < 1.5, np.float32(10), np.float32(0)).astype(np.float32)
eng.set_drag(more, off, X[pick] + np.float32([0.3, 0.2, 0.0]))
eng.Update(2)
print("drag", eng.get_drag()[3], int((more > 0).sum()))
eng.set_drag(None)
eng.update_mu(sc.arrays()["mu"] * np.float32(0.5))
eng.Update(2)
eng.Reset()
eng.Update(1)
X, V, XT = eng.download()
print("ok", np.isfinite(X).all(), X[:2], eng.info(), "launches", eng.GetPerformanceData()[1].kernel_launches)
