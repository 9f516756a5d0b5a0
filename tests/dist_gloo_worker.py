"""world_size-2 gloo worker for tests/test_dist_host.py::test_halo_exchange_world2_gloo (CPU only)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pd = importlib.import_module("soft-body-simulation-cuda_b200")


def value_of(gid):
    """a position that encodes the global vertex id exactly"""
    g = gid.astype(np.float32)
    return np.stack([g, g * np.float32(0.5) + np.float32(1.0), -g], 1)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sc = pd.Scene.kuhn_grid(10, 9, 8, 1.0, 0.05, 11, (0, 5, 0), 1.0, 2e5)
    G = sc.layout()
    P = pd.RankPlan(G, world, rank)
    n_loc = P.num_owned + P.num_ghosts
    # every rank must agree on everybody's window size
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_loc], dtype=torch.int64))
    assert [int(s) for s in sizes] == P.n_loc_of.tolist()
    q = np.full((n_loc, 3), np.nan, np.float32)
    q[:P.num_owned] = value_of(P.first_owned + np.arange(P.num_owned))
    # the push, as k_halo_push performs it: q_peer[dst] = q[src]  (here over gloo send/recv)
    reqs, recv = [], {}
    for n in P.neighbours.tolist():
        m = P.push_rank == n
        payload = torch.from_numpy(np.concatenate([P.push_dst[m].astype(np.float64)[:, None], q[P.push_src[m]].astype(np.float64)], 1).copy())
        cnt = torch.tensor([payload.shape[0]], dtype=torch.int64)
        reqs.append(dist.isend(cnt, n, tag=1))
        reqs.append(dist.isend(payload, n, tag=2))
    for n in P.neighbours.tolist():
        cnt = torch.zeros(1, dtype=torch.int64)
        dist.recv(cnt, n, tag=1)
        buf = torch.zeros(int(cnt), 4, dtype=torch.float64)
        dist.recv(buf, n, tag=2)
        recv[n] = buf.numpy()
    for r in reqs:
        r.wait()
    for n, buf in recv.items():
        q[buf[:, 0].astype(np.int64)] = buf[:, 1:].astype(np.float32)
    want = value_of(P.ghosts.astype(np.int64))
    assert np.array_equal(q[P.num_owned:].view(np.uint32), want.view(np.uint32)), "ghost entries differ from the owners' values"
    assert not np.isnan(q).any()
    dist.barrier()
    print(f"HALO_OK rank {rank}: {P.num_owned} owned, {P.num_ghosts} ghosts, {P.num_push} pushed, neighbours {P.neighbours.tolist()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
