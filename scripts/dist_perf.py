"""Multi-GPU diagnosis (torchrun): per-rank time of the local and the vertex kernel IN SITU (perf mode: events around
every kernel, plain launches, ranks in lock step through the halo flags) next to the same kernels timed alone.
  python -m torch.distributed.run --nproc-per-node N scripts/dist_perf.py [workload] [steps]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch, torch.distributed as dist

workload = sys.argv[1] if len(sys.argv) > 1 else "grid139"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pd = importlib.import_module("soft-body-simulation-cuda_b200")
rank, world, local = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
sc, p = bench.make_scene(pd, workload, rank)
eng = pd.PdSolver(sc, device=local, rank=rank, world=world)
if world > 1:
    h = torch.from_numpy(eng.window_handle()).cuda()
    allh = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    eng.connect(torch.stack(allh).cpu().numpy())
eng.upload(V=bench.initial_velocity(sc.arrays()["X"]))
eng.Update(2); eng.synchronize(); bench.barrier(world)
ms_graph = eng.step_timed(steps) / steps
eng.synchronize(); bench.barrier(world)
eng.SetPerf(True)
eng.Update(1); eng.synchronize(); bench.barrier(world)
names0, p0 = eng.GetPerformanceData()
eng.Update(steps); eng.synchronize(); bench.barrier(world)
names1, p1 = eng.GetPerformanceData()
eng.SetPerf(False)
bench.barrier(world)
tl, tv = eng.time_kernels(reps=20)
it = p["num_iterations"]
n_it = steps * it
row = torch.tensor([ms_graph, (dict(names1)["local step"] - dict(names0)["local step"]) / n_it, (dict(names1)["global step"] - dict(names0)["global step"]) / n_it, tl, tv], dtype=torch.float64, device="cuda")
rows = [torch.zeros_like(row) for _ in range(world)]
if world > 1:
    dist.all_gather(rows, row)
else:
    rows = [row]
if rank == 0:
    print(f"{workload} world {world}: per PD iteration, us")
    print("rank  graph step/it   in situ local   in situ vertex   alone local   alone vertex")
    for r, x in enumerate(rows):
        x = x.cpu().numpy()
        print(f"{r:4d}  {x[0] / it * 1e3:12.1f}  {x[1] * 1e3:14.1f}  {x[2] * 1e3:15.1f}  {x[3] * 1e3:12.1f}  {x[4] * 1e3:13.1f}")
if world > 1:
    dist.destroy_process_group()
