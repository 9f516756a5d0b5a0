"""GPU measurement of the non-Jacobi global steps on BASELINE config 3 (55^3-cell Kuhn grid, 998,250 tets,
10 outer PD iterations per step): PCG-Jacobi with a fixed inner iteration count (throughput variant) and with the
reference's tolerance, and the prefactored sparse Cholesky solve on a small grid.  Prints one JSON line per case with the
roofline accounting of SURVEY.md 8d: B_cg = 8 nnz(A^) + 4 (nV + 1) + 144 nV bytes per CG iteration.

    python scripts/solver_bench.py [--cells 55] [--steps 5] [--inner 50]

Written in a session without GPU time: first run is round 2's.  Times are CUDA events on the engine's stream
(pd_step_timed); the PD / CG iteration counts come from the device-side counters (pd_get_perf)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pd = importlib.import_module("soft-body-simulation-cuda_b200")


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def case(cells, solver, steps, inner, pcg_tol, outer=10):
    sc = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, 0.05, 12345, (0.0, 10.0, 0.0), 1.0, 2e5)
    sc.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
    sc.params = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=outer, global_solver=solver, tol=1e-6, pcg_max_iter=inner, pcg_tol=pcg_tol)
    nV, nT = sc.counts()[:2]
    eng = pd.PdSolver(sc)
    X0 = sc.arrays()["X"]
    V0 = np.zeros_like(X0); V0[:, 1] = 0.5 * np.sin(X0[:, 0] / 7.0)
    eng.upload(V=V0)
    rp, col, val = eng.system_matrix()
    nnz = int(col.shape[0])
    eng.Update(2)                                   # warm-up (builds the matrix / factor)
    eng.synchronize()
    p0 = eng.GetPerformanceData()[1]
    pd0, in0 = p0.pd_iterations, p0.inner_iterations
    ms = eng.step_timed(steps)
    p1 = eng.GetPerformanceData()[1]
    pdi, inn = p1.pd_iterations - pd0, p1.inner_iterations - in0
    X = eng.download()[0]
    out = {"workload": f"grid{cells}", "num_verts": nV, "num_tets": nT, "nnz_A": nnz, "nnz_L": eng.solver_sizes()["nnz_L"], "solver": {1: "sparse Cholesky", 2: "PCG-Jacobi"}[solver],
           "steps": steps, "ms_per_step": ms / steps, "pd_iterations": int(pdi), "inner_iterations": int(inn),
           "pd_iters_per_s": pdi / (ms * 1e-3), "mtet_updates_per_s": nT * pdi / (ms * 1e-3) / 1e6, "finite": bool(np.isfinite(X).all())}
    if solver == 2 and inn > 0:
        b_cg = 8.0 * nnz + 4.0 * (nV + 1) + 144.0 * nV
        # upper bound on the CG share of the step: everything that is not the local step + RHS (timed separately below)
        t_local, _ = eng.time_kernels(reps=20)
        cg_ms = ms - pdi * t_local
        out.update({"inner_iters_per_s": inn / (ms * 1e-3), "bytes_per_cg_iteration": b_cg, "local_kernel_ms": t_local,
                    "cg_ms_upper_bound": cg_ms, "cg_achieved_gbs_lower_bound": b_cg * inn / (cg_ms * 1e-3) / 1e9,
                    "cg_frac_of_measured_peak_lower_bound": b_cg * inn / (cg_ms * 1e-3) / 1e9 / peaks()})
    return out


def reference_case(cells, solver, steps, outer=10):
    """The reference's own back-end on one B200 (oracle/_ref/libpd_ref_solvers.so: PdSolver's non-Jacobi branch replayed around
    CholeskySpLinearSolver<float> / PCGJacobiSolver<float>, default max_iter 2000, tolerance 1e-5), wall clock around synchronised
    steps.  None when the prebuilt harness did not travel with the snapshot."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import time
    import ref
    if not ref.solvers_available():
        return None
    sc = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, 0.05, 12345, (0.0, 10.0, 0.0), 1.0, 2e5)
    a = sc.arrays()
    rs = ref.RefSolverScene(a["X"], a["Tet"], a["mass"], a["mu"], solver)
    V0 = np.zeros_like(a["X"]); V0[:, 1] = 0.5 * np.sin(a["X"][:, 0] / 7.0)
    rs.set(V=V0)
    kw = dict(dt=1 / 60, gravity=9.8, tol=1e-6, num_iterations=outer)
    rs.step(2, **kw); rs.get()
    t0 = time.perf_counter()
    rs.step(steps, **kw); X = rs.get()[0]                    # the D2H read synchronises
    ms = (time.perf_counter() - t0) * 1e3
    return {"workload": f"grid{cells}", "impl": "reference", "solver": {1: "CholeskySpLinearSolver<float> (cuSOLVER)", 2: "PCGJacobiSolver<float> (cuSPARSE/cuBLAS)"}[solver],
            "steps": steps, "ms_per_step": ms / steps, "pd_iterations_last_step": rs.stats()[0], "finite": bool(np.isfinite(X).all())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=55)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--inner", type=int, default=50)
    a = ap.parse_args()
    print(json.dumps(case(a.cells, 2, a.steps, a.inner, 0.0)), flush=True)          # fixed inner count (pcg_tol 0 never triggers)
    print(json.dumps(case(a.cells, 2, a.steps, 2000, 1e-5)), flush=True)            # the reference's stopping rule
    print(json.dumps(case(min(a.cells, 24), 1, a.steps, 0, 0.0)), flush=True)       # small-mesh path (<= 262,144 vertices)
    for cells, solver in ((a.cells, 2), (min(a.cells, 24), 1)):                     # the reference's own back-ends beside them
        try:
            print(json.dumps(reference_case(cells, solver, a.steps)), flush=True)
        except Exception as ex:                                                     # e.g. cuSOLVER out of memory on the large grid
            print(json.dumps({"impl": "reference", "solver": solver, "error": str(ex)}), flush=True)
