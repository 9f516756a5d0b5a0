"""Diagnostic (GPU box): PD + PCG on the grid family -- engine vs the reference's PCGJacobiSolver vs exact solves (oracle,
double Cholesky), per 5 steps, for two CG tolerances."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import meshes, oracle as O, ref
pd = importlib.import_module("soft-body-simulation-cuda_b200")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sc = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, 0.05, 12345, (0.0, 40.0, 0.0), 1.0, 2e5)
kw = dict(dt=1 / 60, gravity=9.8, num_iterations=10, tol=1e-6)
a = sc.arrays()
V0 = np.zeros_like(a["X"]); V0[:, 1] = 0.5 * np.sin(a["X"][:, 0] / 7.0)
scale = float(np.linalg.norm(a["X"].max(0) - a["X"].min(0)))
sc.params = pd.SolverParams(global_solver=2, pcg_max_iter=2000, pcg_tol=1e-5, **kw)
eng = pd.PdSolver(sc); eng.upload(V=V0)
sc.params = pd.SolverParams(global_solver=1, **kw)
ech = pd.PdSolver(sc); ech.upload(V=V0)
rs = ref.RefSolverScene(a["X"], a["Tet"], a["mass"], a["mu"], 2); rs.set(V=V0)
rs2 = ref.RefSolverScene(a["X"], a["Tet"], a["mass"], a["mu"], 2); rs2.set(V=V0)
rc = ref.RefSolverScene(a["X"], a["Tet"], a["mass"], a["mu"], 1); rc.set(V=V0)
osc = O.Scene(a["X"], a["Tet"], a["mass"], a["mu"]); osc.set(V=V0)
op = O.make_params(global_solver=1, threads=os.cpu_count() or 1, **kw)
for s in range(4):
    eng.Update(5); ech.Update(5); rs.step(5, **kw); rs2.step(5, **kw); rc.step(5, **kw); osc.step(op, 5)
    Xe, Xc, Xr, Xr2, Xrc, Xo = eng.download()[0], ech.download()[0], rs.get()[0], rs2.get()[0], rc.get()[0], osc.get()[0]
    e = lambda x, y: meshes.rel_err(x, y, scale)
    perf = eng.GetPerformanceData()[1]
    print(f"step {5*(s+1)}: engPCG-refPCG {e(Xe,Xr):.2e} refPCG-refPCG {e(Xr2,Xr):.2e} | vs exact (oracle f64 Cholesky): engPCG {e(Xe,Xo):.2e} engChol {e(Xc,Xo):.2e} refPCG {e(Xr,Xo):.2e} refCuSolverChol {e(Xrc,Xo):.2e} | eng stats {eng.solve_stats()} inner total {perf.inner_iterations} ref stats {rs.stats()}")
