"""PD_DIST_TRIM=1 (experiment, csrc/layout.hpp): a rank cuts its boundary tiles down to the tets that touch a vertex it owns and
packs them into fewer physical tiles.  The claim to check on the host: for every OWNED vertex the slots, their order, and the
ordered (tet, corner) incidence list behind every slot are exactly those of the single-GPU layout -- which is what makes the
partial sums, and the vertex sums over them, bit-identical -- and the local kernel's gathers still find the right positions.
Emulates the kernel's index arithmetic (phase B store offsets, phase C rows) in numpy.  CPU only."""
import os

import numpy as np
import pytest

OWNER = 0x80000000
ZERO_OFF = 4 * 256 * 16


def _slot_lists(L):
    """slot -> ordered [(original tet id, corner)] as phase C of the local kernel would sum them; also checks that every
    corner's staging slot holds that corner's vertex."""
    out = {}
    rec, T = L.records, L.tile_table
    for ti in range(L.num_tiles):
        off = int(T[ti, 0]) * 16; ab = int(T[ti, 1] & 0xffff); cb = int(T[ti, 1] >> 16)
        nT = int(T[ti, 2] & 0xffff); nLocal = int(T[ti, 2] >> 16)
        t0 = int(L.tile_tet_start[ti])
        assert int(L.tile_tet_start[ti + 1]) - t0 == nT
        tr = rec[off + 96:off + 96 + 48 * nT].view(np.uint32).reshape(3, nT, 4).transpose(1, 0, 2).reshape(nT, 12)
        halves = np.stack([tr[:, 10] & 0xffff, tr[:, 10] >> 16, tr[:, 11] & 0xffff, tr[:, 11] >> 16], 1).astype(np.int64)
        col = (halves >> 12) & 7; stage = (halves >> 4) & 0xff
        vs = L.vstage[256 * ti:256 * ti + 256]
        assert np.array_equal(vs[stage] & ~np.uint32(OWNER), L.tet_new[t0:t0 + nT]), ti       # the gather stages the right vertices
        where = {}
        for tl in range(nT):
            for k in range(4):
                o = k * 4096 + ((tl & ~7) | int(col[tl, k])) * 16                                # h_store
                assert o not in where
                where[o] = (int(L.tet_order[t0 + tl]), k)
        nRows = cb // 128
        incT = rec[off + ab:off + ab + cb].view(np.uint16).reshape(nRows, 32, 2).astype(np.int64)
        seen = 0
        for g in range(8):
            w = int(T[ti, 4 + g]); rb = w & 63; nr = (w >> 6) & 63; nvalid = (w >> 12) & 63
            for lane in range(nvalid):
                ent = incT[rb:rb + nr, lane, :].reshape(-1)
                real = ent[ent < ZERO_OFF]
                assert (ent[len(real):] >= ZERO_OFF).all()                                       # pads only behind the list
                out[256 * ti + 32 * g + lane] = [where[int(o)] for o in real]
                seen += len(real)
        assert seen == 4 * nT, ti                                                                # every contribution is summed exactly once
        vl = L.vlist[256 * ti:256 * ti + 256]
        for l in range(nLocal):                                                                  # ... by the slot of its own vertex
            v = int(vl[l] & ~np.uint32(OWNER))
            for (t_orig, k) in out[256 * ti + l]:
                pass
        assert (vl[nLocal:] == 0xffffffff).all()
    return out


@pytest.mark.parametrize("scene_name,world", [("grid", 2), ("grid", 3), ("grid", 8), ("multibody", 3)])
def test_trimmed_rank_layouts_keep_every_owned_slot(pd, assets, scene_name, world, monkeypatch):
    sc = pd.Scene.kuhn_grid(9, 10, 11, 1.0, 0.05, 5, (0, 3, 0), 1.0, 2e5) if scene_name == "grid" else pd.Scene.from_json(assets["json"], "C2 armadillo&bunny")
    nV, nT = sc.counts()[:2]
    tets = sc.arrays()["Tet"]
    G = sc.layout()
    glists = _slot_lists(G)
    full = trimmed = 0
    monkeypatch.setenv("PD_DIST_TRIM", "1")
    plans = [pd.RankPlan(G, world, r) for r in range(world)]              # the plan carries the switch to its layout
    monkeypatch.setenv("PD_DIST_TRIM", "0")
    plans0 = [pd.RankPlan(G, world, r) for r in range(world)]
    # push lists of the trimmed plans: the same vertex on both sides, every ghost fed exactly once by its owner
    for n, Pn in enumerate(plans):
        fed = np.zeros(Pn.num_ghosts, np.int32)
        for r, Pr in enumerate(plans):
            m = Pr.push_rank == n
            gid = Pr.first_owned + Pr.push_src[m].astype(np.int64)
            slot = Pr.push_dst[m].astype(np.int64) - Pn.num_owned
            assert (slot >= 0).all() and np.array_equal(Pn.ghosts[slot].astype(np.int64), gid) and (Pr.push_src[m] < Pr.num_owned).all()
            np.add.at(fed, slot, 1)
        assert (fed == 1).all()
        assert Pn.n_loc_of.tolist() == [q.num_owned + q.num_ghosts for q in plans]
        assert set(Pn.ghosts.tolist()) <= set(plans0[n].ghosts.tolist()) and Pn.tiles.tolist() == plans0[n].tiles.tolist()
        # the neighbour relation stays symmetric (a rank waits for the flag of every neighbour: an asymmetry would be a deadlock)
        for r in Pn.neighbours.tolist():
            assert n in plans[r].neighbours.tolist()
        assert sorted(Pn.neighbours.tolist()) == sorted(set(np.unique(Pn.push_rank).tolist()))
    print(f"{scene_name} world {world}: ghosts {sum(q.num_ghosts for q in plans0)} -> {sum(q.num_ghosts for q in plans)}")
    for rank in range(world):
        P = plans[rank]
        L0 = plans0[rank].local_layout(G)
        L = P.local_layout(G)
        full += L0.num_tets; trimmed += L.num_tets
        assert L.num_tiles <= L0.num_tiles and L.num_tets <= L0.num_tets
        # interior tiles are untouched and come first
        n_int = P.num_interior_tiles
        assert np.array_equal(L.tile_table[:n_int, 1:], L0.tile_table[:n_int, 1:])
        # the ghosts are exactly the foreign vertices of the kept tets
        used = np.unique(L.tet_new)
        assert used.max() < P.num_owned + P.num_ghosts and set(range(P.num_owned, P.num_owned + P.num_ghosts)) <= set(used.tolist())
        # exactly the tets that touch an owned vertex, once each
        first, n_own = P.first_owned, P.num_owned
        own_orig = set(G.vert_order[first:first + n_own].tolist())
        kept = L.tet_order.tolist()
        assert len(set(kept)) == len(kept)
        touching = {t for t in L0.tet_order.tolist() if own_orig & set(tets[t].tolist())}
        assert touching <= set(kept)
        assert set(kept[int(L.tile_tet_start[n_int]):]) <= touching                    # nothing else in the re-packed part
        # every owned vertex: same slots in the same order, each with the same ordered incidence list and owner flag
        lists = _slot_lists(L)
        for l in range(n_own):
            g = first + l
            sl = L.vslot[L.vslot_ptr[l]:L.vslot_ptr[l + 1]]
            sg = G.vslot[G.vslot_ptr[g]:G.vslot_ptr[g + 1]]
            assert len(sl) == len(sg)
            for a, b in zip(sl.tolist(), sg.tolist()):
                assert lists[a] == glists[b], (rank, l)
                assert int(L.vlist[a]) & ~OWNER == l and int(G.vlist[b]) & ~OWNER == g
                assert (int(L.vlist[a]) & OWNER) == (int(G.vlist[b]) & OWNER)
    print(f"{scene_name} world {world}: tets evaluated over all ranks {full} -> {trimmed} (mesh {nT}): redundant {full / nT - 1:.3f} -> {trimmed / nT - 1:.3f}")
    assert trimmed < full


def test_matrix_diag_host_vs_oracle_and_across_world_sizes(pd, O, assets, monkeypatch):
    """SolverPrepare's matrix_diag (computeSiTSi, pdUtil.cu:16-24) as Engine::prepare computes it from the tile records:
    equal to the oracle's up to the summation order (the reordered tets), and BIT-identical for every owned vertex of every
    rank layout, trimmed or not -- the sums run in ascending global tet order on every world size."""
    import meshes
    sc = pd.Scene.from_json(assets["json"], "C5 house&sphere")
    osc, p = meshes.oracle_scene(O, assets, "C5 house&sphere")
    mdo = osc.setup(O.make_params(dt=p["dt"], gravity=p["gravity"], num_iterations=5))[0]
    G = sc.layout()
    md = G.matrix_diag()
    assert np.allclose(md, mdo[G.vert_order], rtol=2e-6, atol=0)
    for trim in ("0", "1"):
        for world in (2, 3):
            for rank in range(world):
                monkeypatch.setenv("PD_DIST_TRIM", trim)
                P = pd.RankPlan(G, world, rank)
                L = P.local_layout(G)
                monkeypatch.setenv("PD_DIST_TRIM", "0")
                mdl = L.matrix_diag()
                own = md[P.first_owned:P.first_owned + P.num_owned]
                assert np.array_equal(mdl[:P.num_owned].view(np.uint32), own.view(np.uint32)), (trim, world, rank)
