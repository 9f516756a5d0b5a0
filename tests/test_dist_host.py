"""Host-side logic of the multi-GPU path (new work -- the reference is single-GPU):
vertex partition, tile selection, ghost and push lists, per-rank layouts.  No GPU needed; the world-size-2
test runs two processes over the gloo backend and performs one halo exchange with the real push lists."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

import layout_oracle as LO
import meshes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(autouse=True)
def _whole_boundary_tiles(monkeypatch):
    """The plan tests below state the SPECIFICATION: a rank evaluates every global tile that touches one of its vertices,
    unchanged.  The engine's default since round 2 trims and re-packs the boundary tiles (PD_DIST_TRIM, on unless set to 0);
    tests/test_dist_trim.py checks the trimmed plans against these."""
    monkeypatch.setenv("PD_DIST_TRIM", "0")


def _plans(pd, scene, world):
    G = scene.layout()
    return G, [pd.RankPlan(G, world, r) for r in range(world)]


def _tile_vertices(G, t):
    vl = G.vlist[256 * t:256 * t + 256]
    return (vl[vl != 0xffffffff] & 0x7fffffff).astype(np.int64)


def check_plans(pd, scene, world):
    nV = scene.counts()[0]
    G, plans = _plans(pd, scene, world)
    vbeg = LO.partition_vertices(nV, world)
    owner = np.searchsorted(vbeg, np.arange(nV), side="right") - 1
    need = [[[], []] for _ in range(world)]         # tiles each rank must evaluate: [interior, boundary]
    ghosts = [set() for _ in range(world)]
    for t in range(G.num_tiles):
        v = _tile_vertices(G, t)
        rs = set(owner[v].tolist())
        for r in rs:
            need[r][len(rs) > 1].append(t)
            ghosts[r].update(v[owner[v] != r].tolist())
    for r, P in enumerate(plans):
        assert (P.first_owned, P.num_owned) == (vbeg[r], vbeg[r + 1] - vbeg[r])
        assert P.tiles.tolist() == need[r][0] + need[r][1] and P.num_interior_tiles == len(need[r][0])     # interior first
        assert P.ghosts.tolist() == sorted(ghosts[r])
        assert P.n_loc_of.tolist() == [int(vbeg[q + 1] - vbeg[q]) + len(ghosts[q]) for q in range(world)]
        assert sorted(P.neighbours.tolist()) == sorted(set(owner[list(ghosts[r])].tolist())) if ghosts[r] else P.num_neighbours == 0
    # push lists: rank r's entry (src, dst, n) names the same vertex on both sides, every ghost is fed exactly once
    for n, Pn in enumerate(plans):
        fed = np.zeros(Pn.num_ghosts, np.int32)
        for r, Pr in enumerate(plans):
            m = Pr.push_rank == n
            gid = Pr.first_owned + Pr.push_src[m].astype(np.int64)
            slot = Pr.push_dst[m].astype(np.int64) - Pn.num_owned
            assert (slot >= 0).all() and np.array_equal(Pn.ghosts[slot].astype(np.int64), gid)
            assert (Pr.push_src[m] < Pr.num_owned).all()
            np.add.at(fed, slot, 1)
            assert (n in Pr.neighbours.tolist()) == bool(m.any()) or r == n
        assert (fed == 1).all()
    # neighbour relation is symmetric
    for r, Pr in enumerate(plans):
        for n in Pr.neighbours.tolist():
            assert r in plans[n].neighbours.tolist()
    return G, plans


def check_rank_layout(G, P):
    """The rank's layout = the selected global tiles, byte for byte except the re-based header fields; local
    ids = [own range | ghosts]; slot lists of owned vertices = the global ones, same order."""
    L = P.local_layout(G)
    assert L.num_tiles == P.num_tiles
    local_of = {int(P.first_owned + i): i for i in range(P.num_owned)}
    local_of.update({int(g): P.num_owned + i for i, g in enumerate(P.ghosts.tolist())})
    assert np.array_equal(L.vert_order, G.vert_order[np.array(sorted(local_of, key=local_of.get), np.int64)])
    grec, lrec = G.records, L.records
    for lt, gt in enumerate(P.tiles.tolist()):
        g0, g1 = int(G.tile_rec_off[gt]), int(G.tile_rec_off[gt + 1]); l0, l1 = int(L.tile_rec_off[lt]), int(L.tile_rec_off[lt + 1])
        assert g1 - g0 == l1 - l0
        gh = grec[g0:g0 + 32].view(np.uint32).copy(); lh = lrec[l0:l0 + 32].view(np.uint32).copy()
        assert lh[2] == lt * 256 and (int(lh[6]) | int(lh[7]) << 32) == l0        # slotBase, record offset re-based
        gh[[2, 6, 7]] = 0; lh[[2, 6, 7]] = 0
        assert np.array_equal(gh, lh) and np.array_equal(grec[g0 + 32:g1], lrec[l0 + 32:l1])
        gv = G.vlist[256 * gt:256 * gt + 256]; lv = L.vlist[256 * lt:256 * lt + 256]
        valid = gv != 0xffffffff
        assert np.array_equal(valid, lv != 0xffffffff) and np.array_equal(gv[valid] & 0x80000000, lv[valid] & 0x80000000)
        assert [local_of[int(x)] for x in (gv[valid] & 0x7fffffff)] == (lv[valid] & 0x7fffffff).tolist()
    ltile = {gt: lt for lt, gt in enumerate(P.tiles.tolist())}
    for i in range(0, P.num_owned, max(1, P.num_owned // 200)):
        g = P.first_owned + i
        gs = G.vslot[G.vslot_ptr[g]:G.vslot_ptr[g + 1]].astype(np.int64)
        ls = L.vslot[L.vslot_ptr[i]:L.vslot_ptr[i + 1]].astype(np.int64)
        assert [ltile[int(s) // 256] * 256 + int(s) % 256 for s in gs] == ls.tolist()
    return L


def test_rank_plans_grid(pd):
    sc = pd.Scene.kuhn_grid(12, 10, 9, 1.0, 0.05, 3, (0, 5, 0), 1.0, 2e5)
    for world in (2, 3, 8):
        G, plans = check_plans(pd, sc, world)
        for P in plans:
            check_rank_layout(G, P)
        # redundancy stays modest: boundary tiles are evaluated by more than one rank
        assert sum(P.num_tiles for P in plans) <= G.num_tiles * (1 + 0.6 * (world - 1))


def test_rank_plans_multibody_scene(pd, assets):
    sc = pd.Scene.from_json(assets["json"], "C2 armadillo&bunny")
    G, plans = check_plans(pd, sc, 4)
    for P in plans:
        check_rank_layout(G, P)
    assert check_plans(pd, sc, 1)[1][0].num_ghosts == 0


def test_world_one_is_the_single_gpu_layout(pd):
    sc = pd.Scene.kuhn_grid(5, 5, 5, 1.0, 0.05, 3, (0, 5, 0), 1.0, 2e5)
    G = sc.layout()
    P = pd.RankPlan(G, 1, 0)
    L = P.local_layout(G)
    assert (P.num_ghosts, P.num_neighbours, P.num_push, P.num_tiles) == (0, 0, 0, G.num_tiles)
    for k in ["vert_order", "tet_new", "tile_tet_start", "tile_rec_off", "vslot_ptr", "vslot", "vlist", "records"]:
        assert np.array_equal(getattr(L, k), getattr(G, k)), k


@pytest.mark.parametrize("trim", ["0", "1"])
def test_halo_exchange_world2_gloo(tmp_path, trim):
    """Two processes (gloo): each builds its own plan, they exchange one halo with the real push lists and
    check every ghost entry against the owner's value bit for bit.  trim = 1: the PD_DIST_TRIM experiment's shorter
    ghost / push lists (csrc/layout.hpp)."""
    port = 29500 + (os.getpid() + 7 * int(trim)) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", PD_DIST_TRIM=trim)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("HALO_OK") == 2, out.stdout[-2000:]
