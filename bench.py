#!/usr/bin/env python
"""bench.py -- PD step throughput on B200 (see DESIGN.md section 7 for the measurement contract).

  python bench.py --gpus N --steps K --warmup W [--workload grid139] [--impl reference]

One "step" = one PdSolver::Update (predictor, `num of iterations` x (local step + global sweep),
velocity update, fixed-body response) of the synthetic Kuhn-grid workload; metric =
Mtet-updates/s = numTets x PD iterations / second, whole job.  Rank 0 prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (cells per axis, PD iterations per step, description)   SURVEY.md section 8d
    "grid139": (139, 100, "Kuhn 6-tet grid 139^3 cells (2,744,000 verts, 16,113,714 tets), float PD, Chebyshev-Jacobi 100 it/step"),
    "grid70": (70, 100, "Kuhn 6-tet grid 70^3 cells (357,911 verts, 2,058,000 tets) -- the per-GPU share of grid139 on 8 GPUs"),
    "grid55": (55, 100, "Kuhn 6-tet grid 55^3 cells (175,616 verts, 998,250 tets), float PD, Chebyshev-Jacobi 100 it/step"),
    "grid24": (24, 100, "Kuhn 6-tet grid 24^3 cells (82,944 tets) -- CPU-sized sample"),
    # independent (house2 + sphere) contexts, 64 per GPU, no inter-GPU communication (weak scaling)
    "batch64": (0, 100, "64 independent contexts per GPU, each house2 (1,389 tets) + sphere (1,217 tets) over a floor and a fixed sphere, float PD, Chebyshev-Jacobi 100 it/step"),
    # assets/armadillo0 with the parameters of the shipped "Armadillo&house" context (context.json:121-129).  The bunny of
    # config 2 is left out of the TIMED scene: its shipped definition diverges under Jacobi PD in the reference itself (NaN within
    # 5 steps, tests/test_gpu_parity.py::test_vs_reference_cuda_kernels checks exactly that on both sides)
    "armadillo": (0, 100, "assets/armadillo0 (13,054 verts, 41,960 tets) over the floor and five walls, dt 0.01, gravity 98, float PD, Chebyshev-Jacobi 100 it/step"),
}
ASSET_WORKLOADS = {"armadillo": "C2 armadillo"}      # context of the fixture (tests/meshes.py, regenerated from tests/golden/meshes.npz)
# SURVEY C3 as specified: PD + Jacobi-PCG global step, 10 outer PD iterations per step; (inner max, ||r|| tolerance)
SOLVER_WORKLOADS = {
    "grid55-pcg": (55, 10, 50, 0.0, "Kuhn 6-tet grid 55^3 cells (175,616 verts, 998,250 tets), float PD, 10 outer iterations x Jacobi-PCG with a FIXED 50 inner iterations (throughput variant)"),
    "grid55-pcg-tol": (55, 10, 2000, 1e-5, "Kuhn 6-tet grid 55^3 cells (175,616 verts, 998,250 tets), float PD, 10 outer iterations x Jacobi-PCG to the reference's stopping rule (||r|| < 1e-5, at most 2000)"),
}
for _k, _v in SOLVER_WORKLOADS.items():
    WORKLOADS[_k] = (_v[0], _v[1], _v[4])
BATCH_PER_GPU = 64
DT, GRAVITY, MU, MASS, JITTER, SEED = 1.0 / 60.0, 9.8, 2e5, 1.0, 0.05, 12345


def ncu_traffic(workload, kernel="k_local"):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json, written from the .ncu-rep by scripts/ncu_summary.py --traffic); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[workload][kernel]
        return float(t["dram_read_bytes"] + t["dram_write_bytes"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML in-process
    every 2 ms (an `nvidia-smi -lms` child needs >100 ms to produce its first row, longer than a short multi-GPU run);
    falls back to the nvidia-smi query loop if NVML cannot be loaded."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.sm, self.power, self.reasons, self.mx = [], [], set(), None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Start of the timed region: drop what was sampled before it.  (start() runs BEFORE the warm-up: loading NVML takes
        tens of milliseconds that differ from rank to rank, and a rank that enters the timed region late makes its neighbours
        wait for its halo inside their kernels -- seen once as a 7x outlier at N=2.)"""
        self.sm, self.power, self.reasons, self.rows = [], [], set(), []

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 2 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def make_batch_scene(pd, first, count):
    """Contexts first .. first+count-1 of the 512-context batch: the "C5 house&sphere" context of the
    fixture (tests/meshes.py, regenerated from tests/golden/meshes.npz -- nothing is read from /root/reference), pose i =
    base pose rotated about y by 2*pi*i/512 and lifted by (i mod 8)*5, merged into ONE scene (pd_scene_merge)."""
    import tempfile
    import meshes
    with tempfile.TemporaryDirectory() as tmp:
        assets = meshes.write_assets(tmp)
        base = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    a = base.arrays(); p = base.params
    p["num_iterations"] = WORKLOADS["batch64"][1]
    scenes = []
    for i in range(first, first + count):
        th = np.float32(2.0 * np.pi * i / 512.0)
        c, s_ = np.cos(th, dtype=np.float32), np.sin(th, dtype=np.float32)
        X = a["X"].copy()
        X[:, 0] = c * a["X"][:, 0] + s_ * a["X"][:, 2]
        X[:, 2] = -s_ * a["X"][:, 0] + c * a["X"][:, 2]
        X[:, 1] += np.float32((i % 8) * 5)
        scenes.append(pd.Scene.from_arrays(X, a["Tet"], a["mass"], a["mu"], fixed=a["fixed"], params=p))
    sc = pd.Scene.merge(scenes)
    sc.params = p
    return sc, p


def fixed_arrays(pd, fixed):
    """pd_fixed_body structs -> (planes [(p0, up)], spheres [(c, r)], cylinders [(c, axis, r)]) as oracle/ref.py takes them"""
    planes, spheres, cyls = [], [], []
    for f in fixed:
        M = np.array(f.model[:], np.float32)
        if f.type == pd.PD_PLANE:
            up = np.zeros(3, np.float32)
            pd.lib().pd_plane_up(M.ctypes.data, up.ctypes.data)
            planes.append((M[12:15].copy(), up))
        elif f.type == pd.PD_SPHERE:
            spheres.append((M[12:15].copy(), f.radius))
        else:
            ax = M[4:7] / np.float32(np.linalg.norm(M[4:8]))
            cyls.append((M[12:15].copy(), ax.astype(np.float32), f.radius))
    return planes, spheres, cyls


def make_scene(pd, workload, rank=0):
    if workload == "batch64":
        return make_batch_scene(pd, rank * BATCH_PER_GPU, BATCH_PER_GPU)
    if workload in ASSET_WORKLOADS:
        import tempfile
        import meshes
        with tempfile.TemporaryDirectory() as tmp:
            assets = meshes.write_assets(tmp)
            sc = pd.Scene.from_json(assets["json"], ASSET_WORKLOADS[workload])
        p = sc.params
        p["num_iterations"] = WORKLOADS[workload][1]
        sc.params = p
        return sc, p
    cells, iters, _ = WORKLOADS[workload]
    sc = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, JITTER, SEED, (0.0, 10.0, 0.0), MASS, MU)
    if workload in SOLVER_WORKLOADS:
        _, _, inner, cg_tol, _ = SOLVER_WORKLOADS[workload]
        p = pd.SolverParams(dt=DT, gravity=GRAVITY, num_iterations=iters, global_solver=2, tol=1e-6, pcg_max_iter=inner, pcg_tol=cg_tol)
    else:
        p = pd.SolverParams(dt=DT, gravity=GRAVITY, num_iterations=iters)
    sc.params = p
    sc.add_fixed(pd.fixed_body(pd.PD_PLANE, pos=(0, 0, 0), scale=(450, 450, 450)))
    return sc, p


def meshes_rel_err(x, r):
    import meshes
    return meshes.rel_err(x, r)


def initial_velocity(X):
    """v = 0.5 sin(x/7) y^   so that F is non-trivial from the first iteration."""
    V = np.zeros_like(X)
    V[:, 1] = 0.5 * np.sin(X[:, 0] / 7.0)
    return V


def cpu_baseline(workload_iters, threads):
    """The oracle port (oracle/pd_oracle.c, OpenMP) on a bounded sample of the same workload family: the same
    jittered Kuhn grid at 64^3 cells (1.57 M tets, far larger than the host caches, like grid139 is for the GPU),
    one warm-up step + two timed steps of 100 iterations = about 10-20 s on 16 host threads."""
    import oracle as O
    pd = importlib.import_module("soft-body-simulation-cuda_b200")
    cells = int(os.environ.get("PD_CPU_BASELINE_CELLS", "64"))
    sc = pd.Scene.kuhn_grid(cells, cells, cells, 1.0, JITTER, SEED, (0.0, 10.0, 0.0), MASS, MU)
    a = sc.arrays()
    osc = O.Scene(a["X"], a["Tet"], a["mass"], a["mu"], planes=[(np.zeros(3, np.float32), np.array([0, 1, 0], np.float32))])
    osc.set(V=initial_velocity(a["X"]))
    p = O.make_params(dt=DT, gravity=GRAVITY, num_iterations=workload_iters, threads=threads)
    osc.step(p, 1)                      # warm-up (also builds the setup products)
    t0 = time.perf_counter()
    nsteps = 2
    osc.step(p, nsteps)
    dt = time.perf_counter() - t0
    nT = a["Tet"].shape[0]
    val = nT * workload_iters * nsteps / dt / 1e6
    return {"value": val, "unit": "Mtet-updates/s", "cores": threads, "kind": "port",
            "sample": f"oracle/pd_oracle.c (OpenMP, {threads} threads): {nsteps} steps x {workload_iters} it of the {cells}^3-cell Kuhn grid ({nT} tets), {dt:.1f} s"}


def workload_config(workload, nV, nT, iters, p, batch, world):
    """The `config` object of the JSON line: identical in both arms (--impl b200 / reference) for the same workload."""
    stream_mb = 60.0 * nT / 1e6         # ~60 B/tet of tile stream read once per PD iteration
    return {"workload": workload, "description": WORKLOADS[workload][2], "num_verts": nV, "num_tets": nT,
            "pd_iterations_per_step": iters,
            "global_solver": (f"pcg-jacobi (max {SOLVER_WORKLOADS[workload][2]} inner iterations, ||r|| < {SOLVER_WORKLOADS[workload][3]:g}; PD stops at sqrt(err) < 1e-6)"
                              if workload in SOLVER_WORKLOADS else "chebyshev-jacobi"), "dt": p["dt"], "gravity": p["gravity"], "mu": 2e6 if workload in ASSET_WORKLOADS else MU,
            "rho": p["rho"], "muN": p["muN"], "muT": p["muT"],
            "initial_velocity": "0" if (batch or workload in ASSET_WORKLOADS) else "0.5*sin(x/7) y^",
            "l2": (f"inputs larger than L2: ~{stream_mb:.0f} MB of per-tet data are streamed per PD iteration, no flush needed" if stream_mb > 2 * 126
                   else f"working set smaller than L2 (~{stream_mb:.0f} MB per PD iteration): not flushed -- 100 iterations per step re-read the same data, L2-resident is this workload's steady state"),
            "parallelism": "single GPU" if world == 1 else ("contexts sharded 64 per GPU, no communication" if batch else f"vertex partition over {world} GPUs, tile-replicated boundary, NVLink peer-memory halo push (phase tag in every position) per PD iteration")}


PARITY_STEPS = 10


def ref_scene(pd, a):
    import ref
    planes, spheres, cyls = fixed_arrays(pd, a["fixed"])
    return ref.RefScene(a["X"], a["Tet"], a["mass"], a["mu"], planes=planes, spheres=spheres, cylinders=cyls)


def reference_positions(pd, a, p, iters, steps, V0, runs=2):
    """X after `steps` steps of the reference's own CUDA kernels (oracle/_ref, the CHECKER -- never timed here) on the scene
    arrays `a`, `runs` times from the same start (the reference sums with float atomics: the runs differ)."""
    rs = ref_scene(pd, a)
    kw = dict(dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=iters)
    out = []
    for _ in range(runs):
        rs.reset(); rs.set(V=V0)
        rs.step(steps, **kw); rs.sync()
        out.append(rs.get()[0].copy())
    del rs
    return out


def engine_positions(eng, V0, steps):
    eng.Reset(); eng.upload(V=V0)
    eng.Update(steps)
    return eng.download()[0]


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # (NCCL_DEBUG is left as the driver set it: its init lines carry the rank census the driver checks; the JSON line is
        # the LAST line rank 0 prints)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # rank census on stderr (one line per rank): the data plane is CUDA-IPC peer memory, not NCCL, so NCCL's own
        # log shows only the setup collectives
        print(f"[bench] rank {rank}/{world} local_rank {local} device cuda:{local} {torch.cuda.get_device_name(local)} "
              f"uuid {torch.cuda.get_device_properties(local).uuid}", file=sys.stderr, flush=True)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, local):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_b200(args):
    import torch
    pd = importlib.import_module("soft-body-simulation-cuda_b200")
    rank, world, local = dist_setup(args.gpus)
    _t0 = time.perf_counter()

    def trace(what):        # PD_BENCH_TRACE=1: wall-clock milestones per rank on stderr (diagnosis of multi-GPU stalls)
        if os.environ.get("PD_BENCH_TRACE"):
            st = eng.dist_status() if (world > 1 and not batch) else 0
            print(f"[trace] rank {rank} +{time.perf_counter() - _t0:8.3f} s {what} (halo status {st})", file=sys.stderr, flush=True)
    torch.cuda.set_device(local)
    batch = args.workload == "batch64"
    sc, p = make_scene(pd, args.workload, rank)
    nV, nT = sc.counts()[:2]
    iters = p["num_iterations"]
    # N > 1: the mesh is vertex-partitioned, one rank per GPU; every rank builds the same global layout and keeps
    # the tiles that touch its vertices.  torch.distributed (NCCL) carries only the setup
    # (window handles) and the timing reductions; the per-iteration halo goes over NVLink peer memory.
    # (the batch workload shards whole contexts instead: every rank steps its own 64, no communication at all)
    eng = pd.PdSolver(sc, device=local, rot_mode=args.rot_mode, ctas_per_sm=args.ctas_per_sm,
                      rank=0 if batch else rank, world=1 if batch else world)
    dist_info = None
    if world > 1 and not batch:
        import torch.distributed as dist
        h = torch.from_numpy(eng.window_handle()).cuda()
        allh = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(allh, h)
        eng.connect(torch.stack(allh).cpu().numpy())
        di = eng.dist_info()
        t = torch.tensor([di["num_owned"], di["num_ghosts"], di["num_tets_local"], di["num_push"]], dtype=torch.int64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        allt = torch.stack(allt).cpu().numpy()
        print(f"[bench] rank {rank}: owns {di['num_owned']} vertices, {di['num_ghosts']} ghosts, pushes {di['num_push']} positions per PD iteration "
              f"to peer windows over NVLink (CUDA IPC)", file=sys.stderr, flush=True)
        dist_info = {"owned_verts_per_rank": allt[:, 0].tolist(), "ghost_verts_per_rank": allt[:, 1].tolist(),
                     "tets_evaluated_per_rank": allt[:, 2].tolist(), "pushed_verts_per_rank": allt[:, 3].tolist(),
                     "redundant_tet_fraction": float(allt[:, 2].sum() / nT - 1.0)}
    X0 = sc.arrays()["X"]
    asset = args.workload in ASSET_WORKLOADS
    V0 = np.zeros_like(X0) if (batch or asset) else initial_velocity(X0)
    eng.upload(V=V0)
    info = eng.info()
    trace("engine ready")

    # ---- device-resident timing: inputs already in HBM, K steps bracketed by sync + events
    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup):
        eng.Update(1)
    eng.synchronize(); torch.cuda.synchronize(); trace("warm-up done"); barrier(world)
    sampler.mark()
    perf0 = eng.GetPerformanceData()[1].kernel_launches
    # the engine launches on its own stream, so the CUDA events are recorded there (pd_step_timed)
    dev_ms = eng.step_timed(args.steps)
    clocks = sampler.stop()
    barrier(world)
    ms_step = max_over_ranks(dev_ms / args.steps, world, local)
    launches = eng.GetPerformanceData()[1].kernel_launches - perf0
    trace("timed steps done")

    # ---- kernel-level roofline, measured live with CUDA events on the engine's stream
    barrier(world)
    t_local_ms, t_vertex_ms = eng.time_kernels(reps=20)
    trace("time_kernels done")
    peak, peak_src = peaks()
    # algorithmic bytes of what THIS rank's launch processes (56 B/tet + 24 B/vertex local, 68 B/vertex global)
    nT_launch = nT if (world == 1 or batch) else eng.dist_info()["num_tets_local"]
    nV_launch = nV if (world == 1 or batch) else eng.dist_info()["num_owned"]
    bytes_local = 56.0 * nT_launch + 24.0 * nV_launch
    bytes_iter = 56.0 * nT_launch + 68.0 * nV_launch
    ach_local = bytes_local / (t_local_ms * 1e-3) / 1e9
    ach_iter = bytes_iter / ((t_local_ms + t_vertex_ms) * 1e-3) / 1e9

    # ---- end to end: HOST (pinned) state in -> H2D, step, D2H -> host state out, every step.
    # N = 1 (and the batch): pd_step_host, whole arrays in the caller's numbering.  N > 1: every rank owns and moves
    # only its shard of X / V / XTilde (pd_step_host_owned, 3 * num_owned floats per array).
    import ctypes as C
    sharded = world > 1 and not batch
    X, V, XT = eng.download()
    if sharded:
        ids = eng.owned_ids()
        X, V, XT = X[ids], V[ids], XT[ids]
    nbytes = X.nbytes
    bufs = [pd.lib().pd_alloc_pinned(nbytes) for _ in range(6)]
    arr = [np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=X.shape) for b in bufs]
    arr[0][:] = X; arr[1][:] = V; arr[2][:] = XT
    e2e_steps = max(1, min(args.steps, 5))
    host_step = eng.step_host_owned_ptr if sharded else eng.step_host_ptr
    host_step(1, bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5])       # warm-up
    eng.synchronize(); barrier(world)
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        i, o = (3, 0) if s % 2 == 0 else (0, 3)
        host_step(1, bufs[i], bufs[i + 1], bufs[i + 2], bufs[o], bufs[o + 1], bufs[o + 2])
    t1 = time.perf_counter()
    e2e_ms = max_over_ranks((t1 - t0) * 1e3 / e2e_steps, world, local)
    final = arr[0] if e2e_steps % 2 == 1 else arr[3]
    finite = bool(np.isfinite(final).all())
    for b in bufs:
        pd.lib().pd_free_pinned(b)
    if world > 1:       # bytes moved by the whole job per step
        import torch.distributed as dist
        tb = torch.tensor([3 * nbytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(tb)
        job_bytes = int(tb.item())
    else:
        job_bytes = 3 * nbytes
    halo_ok = eng.dist_status() == 0 if (world > 1 and not batch) else True
    trace("e2e done")
    barrier(world)

    # ---- parity of what was just timed (outside every timed region; oracle/_ref is the CHECKER here, never the product)
    parity = None
    Xa = None
    if not args.no_parity and not batch:
        a = sc.arrays()
        if world == 1:
            import ref
            Xe = engine_positions(eng, V0, PARITY_STEPS)
            parity = {"workload": args.workload, "steps": PARITY_STEPS, "rot_mode": args.rot_mode,
                      "definition": "max_v |x_v - ref_v| / max(|ref_v|, bounding-box diagonal) after `steps` steps from the benchmark's initial state"}
            if ref.available():
                Xa, Xb = reference_positions(pd, a, p, iters, PARITY_STEPS, V0)
                parity.update({"reference": "oracle/_ref: the reference's CUDA kernels compiled verbatim (pdUtil.cu, svd3_cuda.h) replayed by oracle/ref_harness.cu",
                               "rel_err": meshes_rel_err(Xe, Xa), "reference_vs_reference": meshes_rel_err(Xb, Xa)})
            else:
                parity.update({"reference": "unavailable (oracle/_ref/libpd_ref.so not built)", "rel_err": None})
        else:
            # N > 1: the partitioned run must be BIT-identical to one GPU.  Every rank contributes its owned
            # vertices after 3 steps from the benchmark's initial state; rank 0 steps a single-GPU engine on the same mesh.
            import torch.distributed as dist
            steps_bi = 3
            Xd = engine_positions(eng, V0, steps_bi)            # owned rows filled, others 0
            trace("parity steps done")
            own = np.zeros(nV, np.uint8); own[eng.owned_ids()] = 1
            t = torch.from_numpy(np.ascontiguousarray(Xd * own[:, None])).cuda()
            dist.all_reduce(t)                                  # disjoint owners: the sum is the assembled state, bit for bit
            cnt = torch.from_numpy(own.astype(np.int32)).cuda(); dist.all_reduce(cnt)
            Xall = t.cpu().numpy()
            if rank == 0:
                one = pd.PdSolver(sc, device=local, rot_mode=args.rot_mode)
                X1 = engine_positions(one, V0, steps_bi)
                one.close()
                same = bool(np.array_equal(Xall, X1)) and bool(np.isfinite(X1).all()) and bool((cnt.cpu().numpy() == 1).all())
                import hashlib
                parity = {"workload": args.workload, "steps": steps_bi, "bit_identical_to_n1": same,
                          "max_abs_diff_to_n1": float(np.abs(Xall.astype(np.float64) - X1).max()),
                          "sha256_n": hashlib.sha256(Xall.tobytes()).hexdigest()[:16], "sha256_1": hashlib.sha256(X1.tobytes()).hexdigest()[:16],
                          "every_vertex_owned_once": bool((cnt.cpu().numpy() == 1).all())}
            barrier(world)

    # ---- the bit-faithful mode (rot_mode 1: the reference's SVD operation for operation, reference summation order),
    # second driver-visible value: its own timing and its own parity record (N = 1 only)
    faithful = None
    if world == 1 and not batch and not args.no_faithful and args.rot_mode == 0:
        import ref
        fe = pd.PdSolver(sc, device=local, rot_mode=1)
        fe.upload(V=V0)
        fe.Update(3)
        fe.synchronize()
        f_steps = max(1, min(args.steps, 2))
        f_ms = fe.step_timed(f_steps) / f_steps
        faithful = {"rot_mode": 1, "steps": f_steps, "ms_per_step": f_ms, "value": nT * iters / (f_ms * 1e-3) / 1e6, "unit": "Mtet-updates/s"}
        if Xa is not None:
            Xf = engine_positions(fe, V0, PARITY_STEPS)
            faithful["parity_rel_err"] = meshes_rel_err(Xf, Xa)
        fe.close()

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    nT_job = nT * world if batch else nT          # batch: every rank steps its own contexts (weak scaling)
    value = nT_job * iters / (ms_step * 1e-3) / 1e6
    line = {
        "metric": "Mtet-updates/s & PD iters/s at 1/2/4/8 B200; % of HBM roofline",
        "value": value, "unit": "Mtet-updates/s", "pd_iters_per_s": iters / (ms_step * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak" if batch else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, nV * (world if batch else 1), nT_job, iters, p, batch, world),
        "run": {"finite": finite, "multi_gpu": dist_info, "halo_ok": halo_ok, "tile_stream_bytes": info["tile_stream_bytes"], "rot_mode": args.rot_mode,
                "experiments": {k: os.environ[k] for k in ("PD_DIST_TRIM", "PD_B200_LIB", "PD_PDL", "PD_NO_STAGE_COLOR") if os.environ.get(k)} or None},
        "parity": parity,
        "faithful": faithful,
        "clocks": clocks,
        "e2e": {"value": nT_job * iters / (e2e_ms * 1e-3) / 1e6, "unit": "Mtet-updates/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": job_bytes, "d2h_bytes_per_step": job_bytes, "steps": e2e_steps,
                "api": ("pd_step_host_owned (include/pd_b200.h): every rank moves its own shard of X,V,XTilde, pinned host in and out every step" if sharded
                        else "pd_step_host (include/pd_b200.h): pinned host X,V,XTilde in and out every step")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_local (local step: F, rotation, ordered RHS partials)" + ("" if world == 1 else " -- rank 0's launch"),
                     "achieved": ach_local, "peak": peak, "unit": "GB/s", "frac": ach_local / peak,
                     "traffic": ncu_traffic(args.workload) if world == 1 else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_local, "launch_ms": t_local_ms,
                     "fused_iteration": {"achieved": ach_iter, "frac": ach_iter / peak, "algorithmic_bytes": bytes_iter,
                                         "ms": t_local_ms + t_vertex_ms, "vertex_kernel_ms": t_vertex_ms},
                     "frac_of_8TBs_nominal": ach_local / 8000.0},
        "cpu_baseline": cpu_baseline(iters, os.cpu_count() or 1) if (not args.no_cpu_baseline and world == 1) else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_b200_solver(args):
    """BASELINE config 3 as specified: PD with the Jacobi-PCG global step (pcgJacobi.cu:88-172 inside PdSolver's direct branch,
    pdSolver.cu:174-192) on ONE GPU.  value = tets x PD iterations actually executed / s (the PD loop may stop on its
    tolerance); the roofline is k_pcg_solve's: B_cg = 8 nnz(A^) + 4 (nV + 1) + 144 nV bytes per CG iteration."""
    import torch
    import ctypes as C
    pd = importlib.import_module("soft-body-simulation-cuda_b200")
    torch.cuda.set_device(0)
    sc, p = make_scene(pd, args.workload)
    nV, nT = sc.counts()[:2]
    iters = p["num_iterations"]
    X0 = sc.arrays()["X"]
    V0 = initial_velocity(X0)
    eng = pd.PdSolver(sc, device=0, rot_mode=args.rot_mode)
    eng.upload(V=V0)
    for _ in range(args.warmup):
        eng.Update(1)
    eng.synchronize(); torch.cuda.synchronize()
    c0 = eng.GetPerformanceData()[1]
    sampler = ClockSampler(0); sampler.start()
    dev_ms = eng.step_timed(args.steps)
    clocks = sampler.stop()
    c1 = eng.GetPerformanceData()[1]
    pd_it, inner = c1.pd_iterations - c0.pd_iterations, c1.inner_iterations - c0.inner_iterations
    ms_step = dev_ms / args.steps
    # the global-solve share, from events around every launch (perf mode: one host sync per step), on further steps
    eng.SetPerf(True)
    n0, r0 = eng.GetPerformanceData()
    eng.Update(args.steps)
    n1, r1 = eng.GetPerformanceData()
    eng.SetPerf(False)
    local_ms, solve_ms = n1[0][1] - n0[0][1], n1[1][1] - n0[1][1]
    inner_perf = r1.inner_iterations - r0.inner_iterations
    sizes = eng.solver_sizes()
    b_cg = 8.0 * sizes["nnz_A"] + 4.0 * (nV + 1) + 144.0 * nV
    peak, peak_src = peaks()
    ach = b_cg * inner_perf / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
    # end to end through pd_step_host
    X, V, XT = eng.download()
    nbytes = X.nbytes
    bufs = [pd.lib().pd_alloc_pinned(nbytes) for _ in range(6)]
    arr = [np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=X.shape) for b in bufs]
    arr[0][:] = X; arr[1][:] = V; arr[2][:] = XT
    e2e_steps = max(1, min(args.steps, 5))
    eng.step_host_ptr(1, bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5])
    eng.synchronize()
    e0 = eng.GetPerformanceData()[1].pd_iterations
    t0 = time.perf_counter()
    for s_ in range(e2e_steps):
        i, o = (3, 0) if s_ % 2 == 0 else (0, 3)
        eng.step_host_ptr(1, bufs[i], bufs[i + 1], bufs[i + 2], bufs[o], bufs[o + 1], bufs[o + 2])
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_it = eng.GetPerformanceData()[1].pd_iterations - e0
    finite = bool(np.isfinite(arr[0] if e2e_steps % 2 == 1 else arr[3]).all())
    for b in bufs:
        pd.lib().pd_free_pinned(b)
    # parity: the reference's own PCGJacobiSolver<float> in PdSolver's direct branch (oracle/ref_solvers.cu; that harness
    # has no fixed bodies: 5 steps of free flight from the benchmark's initial state, its default max 2000 / 1e-5)
    parity = None
    if not args.no_parity:
        import ref
        parity = {"workload": args.workload, "steps": 5, "definition": "max_v |x_v - ref_v| / max(|ref_v|, bounding-box diagonal)"}
        if ref.solvers_available():
            a = sc.arrays()
            kw = dict(dt=DT, gravity=GRAVITY, tol=1e-6, num_iterations=iters)
            outs = []
            for _ in range(2):
                rs = ref.RefSolverScene(a["X"], a["Tet"], a["mass"], a["mu"], 2); rs.set(V=V0)
                rs.step(5, **kw); outs.append(rs.get()[0].copy()); del rs
            q = sc.params; q["pcg_max_iter"] = 2000; q["pcg_tol"] = 1e-5; sc.params = q
            pe = pd.PdSolver(sc, device=0, rot_mode=args.rot_mode); pe.upload(V=V0); pe.Update(5)
            parity.update({"reference": "oracle/_ref: PdSolver's direct branch around the reference's PCGJacobiSolver<float> (cuSPARSE / cuBLAS)",
                           "rel_err": meshes_rel_err(pe.download()[0], outs[0]), "reference_vs_reference": meshes_rel_err(outs[1], outs[0])})
            pe.close()
        else:
            parity.update({"reference": "unavailable (oracle/_ref/libpd_ref_solvers.so not built)", "rel_err": None})
    value = nT * (pd_it / args.steps) / (ms_step * 1e-3) / 1e6
    line = {
        "metric": "Mtet-updates/s & PD iters/s at 1/2/4/8 B200; % of HBM roofline",
        "value": value, "unit": "Mtet-updates/s", "pd_iters_per_s": (pd_it / args.steps) / (ms_step * 1e-3), "inner_iters_per_s": (inner / args.steps) / (ms_step * 1e-3),
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, nV, nT, iters, p, False, 1),
        "run": {"finite": finite, "pd_iterations_per_step_executed": pd_it / args.steps, "inner_iterations_per_step": inner / args.steps,
                "nnz_A": sizes["nnz_A"], "rot_mode": args.rot_mode},
        "parity": parity, "clocks": clocks,
        "e2e": {"value": nT * (e2e_it / e2e_steps) / (e2e_ms * 1e-3) / 1e6, "unit": "Mtet-updates/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 3 * nbytes,
                "d2h_bytes_per_step": 3 * nbytes, "steps": e2e_steps, "api": "pd_step_host (include/pd_b200.h): pinned host X,V,XTilde in and out every step"},
        "gpu_launches": int(c1.kernel_launches - c0.kernel_launches),
        "roofline": {"bound": "hbm", "kernel": "k_pcg_solve (one cooperative launch per PD iteration = one whole PCG solve: fused SpMV + direction update + dots)",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic(args.workload, "k_pcg_solve"), "peak_source": peak_src,
                     "algorithmic_bytes_per_cg_iteration": b_cg, "cg_iterations_timed": int(inner_perf), "solve_kernels_ms": solve_ms, "local_and_rhs_kernels_ms": local_ms,
                     "us_per_cg_iteration": solve_ms * 1e3 / max(inner_perf, 1),
                     "note": "the working set (A^ and six vectors, ~46 MB) is L2 resident: the HBM roofline is the contract's yardstick, the kernel's own bound is the grid-wide synchronisations (two per CG iteration) and L2 latency"},
        "cpu_baseline": cpu_baseline(100, os.cpu_count() or 1) if not args.no_cpu_baseline else None,
    }
    print(json.dumps(line), flush=True)


def run_reference_solver(args):
    """Reference arm of the PCG workloads: the reference's own PCGJacobiSolver<float> inside PdSolver's direct branch (oracle/_ref,
    compiled from its sources: cuSPARSE SpMV + cuBLAS dots, host loop with three scalar read-backs per CG iteration) on one B200."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import ref
    import torch
    pd = importlib.import_module("soft-body-simulation-cuda_b200")      # host-side scene generator only
    sc, p = make_scene(pd, args.workload)
    nV, nT = sc.counts()[:2]
    iters = p["num_iterations"]
    a = sc.arrays()
    base = {"metric": "Mtet-updates/s & PD iters/s at 1/2/4/8 B200; % of HBM roofline", "unit": "Mtet-updates/s", "impl": "reference", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, nV, nT, iters, p, False, 1)}
    if not ref.solvers_available():
        print(json.dumps({**base, "unavailable": "oracle/_ref/libpd_ref_solvers.so not built"}))
        return
    rs = ref.RefSolverScene(a["X"], a["Tet"], a["mass"], a["mu"], 2)
    rs.set(V=initial_velocity(a["X"]))
    kw = dict(dt=DT, gravity=GRAVITY, tol=1e-6, num_iterations=iters)
    rs.step(args.warmup, **kw); rs.get()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    e0.record()
    rs.step(args.steps, **kw)
    e1.record(); e1.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    it_last = rs.stats()[0]
    val = nT * it_last / (ms * 1e-3) / 1e6
    base.update({"value": val, "ms_per_step": ms, "pd_iters_per_s": it_last / (ms * 1e-3), "clocks": clocks,
                 "reference_kind": "PdSolver's direct branch around the reference's PCGJacobiSolver<float> (its defaults: at most 2000 inner iterations, ||r|| < 1e-5), oracle/ref_solvers.cu on one B200",
                 "run": {"pd_iterations_last_step": it_last, "finite": bool(np.isfinite(rs.get()[0]).all())},
                 "cpu_baseline": None,
                 "e2e": {"value": val, "unit": "Mtet-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def run_reference(args):
    """Reference arm: the reference's own CUDA kernels (oracle/_ref, compiled verbatim from
    /root/reference) replayed launch for launch on ONE B200, same workload/metric.  If the prebuilt
    harness is absent, the oracle port on the host cores stands in (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ref
    if args.workload == "batch64":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm replays the grid workloads only (its harness takes one floor plane)"}))
        return
    pd = importlib.import_module("soft-body-simulation-cuda_b200")      # host-side scene generator only
    sc, p = make_scene(pd, args.workload)
    nV, nT = sc.counts()[:2]
    iters = p["num_iterations"]
    a = sc.arrays()
    cores = os.cpu_count() or 1
    base = {"metric": "Mtet-updates/s & PD iters/s at 1/2/4/8 B200; % of HBM roofline", "unit": "Mtet-updates/s",
            "impl": "reference", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, nV, nT, iters, p, False, 1)}
    cpu = cpu_baseline(iters, cores)
    if ref.available():
        rs = ref_scene(pd, a)       # the scene's own fixed bodies (the grid workloads: one floor plane)
        rs.set(V=np.zeros_like(a["X"]) if args.workload in ASSET_WORKLOADS else initial_velocity(a["X"]))
        kw = dict(dt=p["dt"], gravity=p["gravity"], rho=p["rho"], muN=p["muN"], muT=p["muT"], num_iterations=iters)
        rs.step(args.warmup, **kw); rs.sync()
        # same clock as the other arm: CUDA events on the stream the kernels run on (the harness launches on the legacy
        # default stream, which is torch's current stream); the host clock around the same region is kept as a cross-check
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        sampler = ClockSampler(0); sampler.start()
        t0 = time.perf_counter()
        e0.record()
        rs.step(args.steps, perf=False, **kw)
        e1.record(); e1.synchronize(); rs.sync()
        t1 = time.perf_counter()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1) / args.steps
        ms_host_clock = (t1 - t0) * 1e3 / args.steps
        t0 = time.perf_counter()
        rs.step(max(1, args.steps // 2), perf=True, **kw); rs.sync()     # as shipped: SetPerf(true), 2 event syncs per iteration
        ms_perf = (time.perf_counter() - t0) * 1e3 / max(1, args.steps // 2)
        val = nT * iters / (ms * 1e-3) / 1e6
        base.update({"value": val, "ms_per_step": ms, "ms_per_step_host_clock": ms_host_clock, "pd_iters_per_s": iters / (ms * 1e-3), "clocks": clocks,
                     "reference_kind": "the reference's CUDA kernels (pdUtil.cu, svd3_cuda.h, computeInvDmV0) compiled verbatim for sm_100a, replayed by oracle/ref_harness.cu on one B200",
                     "value_perf_true": nT * iters / (ms_perf * 1e-3) / 1e6, "ms_per_step_perf_true": ms_perf,
                     "gpu_launches": int(args.steps * (5 + 4 * iters + 1)),
                     "cpu_baseline": cpu,
                     "e2e": {"value": val, "unit": "Mtet-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:
        base.update({"value": cpu["value"], "ms_per_step": None, "cpu_baseline": cpu,
                     "reference_kind": "oracle port on host cores (oracle/_ref/libpd_ref.so not present)",
                     "e2e": {"value": cpu["value"], "unit": "Mtet-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="grid139", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity record (experiments: saves the reference replay)")
    ap.add_argument("--no-faithful", action="store_true", help="skip the rot_mode=1 record")
    ap.add_argument("--rot-mode", type=int, default=0, help="experiments only: 0 product default, 1 faithful SVD, 2 no projection (timing probe)")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="experiments only: cap the local kernel's resident CTAs per SM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload in SOLVER_WORKLOADS:
        if args.gpus > 1:
            sys.exit("the PCG workloads are single-GPU bench lines (the partitioned PCG path is covered by tests/dist_gpu_worker.py)")
        (run_reference_solver if args.impl == "reference" else run_b200_solver)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
