"""torchrun worker: N ranks, one GPU each, step a grid scene through the multi-GPU engine and compare the
combined state with a single-GPU run of the same scene (rank 0), bit for bit."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pd = importlib.import_module("soft-body-simulation-cuda_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    sc = pd.Scene.kuhn_grid(n, n, n, 1.0, 0.05, 5, (0, 3, 0), 1.0, 2e5)
    sc.params = pd.SolverParams(dt=1 / 60, gravity=9.8, num_iterations=30)
    sc.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
    X0 = sc.arrays()["X"]
    V0 = np.zeros_like(X0); V0[:, 1] = 0.3 * np.sin(X0[:, 0])
    eng = pd.PdSolver(sc, device=local, rank=rank, world=world)
    h = torch.from_numpy(eng.window_handle()).cuda()
    allh = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    eng.connect(torch.stack(allh).cpu().numpy())
    eng.upload(V=V0)
    dist.barrier()
    eng.Update(steps)
    X, V, XT = eng.download()
    assert eng.dist_status() == 0, "halo wait timed out"
    t = torch.from_numpy(np.stack([X, V, XT])).cuda()
    dist.all_reduce(t)                      # the ranks' arrays are disjoint: the sum combines them
    dist.barrier()
    if rank == 0:
        ref = pd.PdSolver(sc, device=local)
        ref.upload(V=V0)
        ref.Update(steps)
        Xr, Vr, XTr = ref.download()
        got = t.cpu().numpy()
        assert np.abs(Xr - X0).max() > 1e-3
        ok = np.array_equal(got[0].view(np.uint32), Xr.view(np.uint32)) and np.array_equal(got[1].view(np.uint32), Vr.view(np.uint32)) \
            and np.array_equal(got[2].view(np.uint32), XTr.view(np.uint32))
        print("max |dX|", np.abs(got[0] - Xr).max(), eng.dist_info(), flush=True)
        assert ok, "multi-GPU result differs from the single-GPU run"
        print("DIST_GPU_OK world", world, flush=True)
    dist.barrier()

    # ---- distributed Jacobi-PCG global step: rows of A^ partitioned like the vertices, halo of p and the
    # dot-product all-reduce over peer memory inside the cooperative solve kernel.  The ranks add the partial dots in a
    # different grouping than one GPU does, so the bar is the BASELINE tolerance (1e-4), not bit equality.
    kw = dict(dt=1 / 60, gravity=9.8, num_iterations=6, tol=1e-7, global_solver=2, pcg_max_iter=40, pcg_tol=1e-4)
    sc2 = pd.Scene.kuhn_grid(n, n, n, 1.0, 0.05, 5, (0, 3, 0), 1.0, 2e5)
    sc2.params = pd.SolverParams(**kw)
    sc2.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
    eng2 = pd.PdSolver(sc2, device=local, rank=rank, world=world)
    h = torch.from_numpy(eng2.window_handle()).cuda()
    allh = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    eng2.connect(torch.stack(allh).cpu().numpy())
    eng2.upload(V=V0)
    dist.barrier()
    eng2.Update(steps)
    X, V, XT = eng2.download()
    st = eng2.GetPerformanceData()[1]
    assert eng2.dist_status() == 0, "a PCG halo / all-reduce wait timed out"
    t = torch.from_numpy(np.stack([X, V, XT])).cuda()
    dist.all_reduce(t)
    it = torch.tensor([st.inner_iterations, st.pd_iterations], dtype=torch.int64, device="cuda")
    itmax = it.clone(); dist.all_reduce(itmax, op=dist.ReduceOp.MAX)
    itmin = it.clone(); dist.all_reduce(itmin, op=dist.ReduceOp.MIN)
    assert torch.equal(itmax, itmin), "the ranks disagree on the CG / PD iteration counts"
    if rank == 0:
        ref = pd.PdSolver(sc2, device=local)
        ref.upload(V=V0)
        ref.Update(steps)
        Xr, Vr, XTr = ref.download()
        sr = ref.GetPerformanceData()[1]
        got = t.cpu().numpy()
        scale = np.abs(Xr).max()
        err = np.abs(got[0] - Xr).max() / scale
        print(f"dist PCG: rel err vs one GPU {err:.2e}; inner iterations {st.inner_iterations} vs {sr.inner_iterations}; PD iterations {st.pd_iterations} vs {sr.pd_iterations}", flush=True)
        assert np.abs(Xr - X0).max() > 1e-3 and err <= 1e-4, err
        assert st.inner_iterations > 0 and abs(st.inner_iterations - sr.inner_iterations) <= max(3, sr.inner_iterations // 20)
        print("DIST_PCG_OK world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
