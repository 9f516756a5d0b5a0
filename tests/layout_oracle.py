"""numpy restatement of the device-layout specification -- the oracle for
the bit-exact reorder / renumber / tile / incidence / slot checks.  Independent of the C++ builder
(soft-body-simulation-cuda_b200/csrc/layout.cpp): integer work in numpy, float32 ops one rounding each."""
import numpy as np

TILE_T = 256
TILE_NLMAX = 256
TILE_ROWSMAX = 56
TILE_HSTRIDE = TILE_T * 16
TILE_ZERO_OFF = 4 * TILE_HSTRIDE
TILE_OWNER_BIT = 0x80000000
TILE_OFF_TETS = 96
TILE_GROUP = 32      # vertices per phase-C group (layout.hpp PD_TILE_GROUP): 32 = one lane per vertex, 16 = two
TILE_LPV = 32 // TILE_GROUP


def spread3(x):
    x = x.astype(np.uint32) & np.uint32(0x3ff)
    x = (x | (x << np.uint32(16))) & np.uint32(0x030000ff)
    x = (x | (x << np.uint32(8))) & np.uint32(0x0300f00f)
    x = (x | (x << np.uint32(4))) & np.uint32(0x030c30c3)
    x = (x | (x << np.uint32(2))) & np.uint32(0x09249249)
    return x


def morton_keys(X, Tet):
    X = X.astype(np.float32); T = Tet.astype(np.int64)
    c = ((X[T[:, 0]] + X[T[:, 1]]) + (X[T[:, 2]] + X[T[:, 3]])) * np.float32(0.25)
    lo = c.min(0); hi = c.max(0)
    ext = np.float32((hi - lo).max())
    if not ext > 0:
        ext = np.float32(1.0)
    scale = np.float32(1023.0) / ext
    q = ((c - lo) * scale).astype(np.int32)          # truncation
    q = np.clip(q, 0, 1023).astype(np.uint32)
    return spread3(q[:, 0]) | (spread3(q[:, 1]) << np.uint32(1)) | (spread3(q[:, 2]) << np.uint32(2))


def rest_shape(X, Tet):
    """solverUtil.cuh:98-116 (glm::inverse, |det|/6) with the fused operations of the reference's nvcc
    build -- fma() cannot be written in numpy, so the values come from the C oracle's o_rest_shape
    (oracle/pd_oracle.c:rest_one), which tests pin bit-for-bit against the reference kernel on the GPU."""
    import oracle as O
    X = np.ascontiguousarray(X, np.float32); T = np.ascontiguousarray(Tet, np.uint32)
    B = np.zeros((T.shape[0], 9), np.float32); V0 = np.zeros(T.shape[0], np.float32)
    O.lib().o_rest_shape(X.reshape(-1), T.reshape(-1), T.shape[0], B.reshape(-1), V0)
    return B.reshape(-1, 3, 3), V0


def rup(x, a):
    return (x + a - 1) // a * a


def build(X, Tet, mu, reorder=True):
    nV, nT = X.shape[0], Tet.shape[0]
    if reorder and nT > 1:
        tet_order = np.argsort(morton_keys(X, Tet), kind="stable").astype(np.uint32)
    else:
        tet_order = np.arange(nT, dtype=np.uint32)
    T = Tet[tet_order].astype(np.int64)
    # first-touch renumbering
    flat = T.reshape(-1)
    _, first = np.unique(flat, return_index=True)
    touched = flat[np.sort(first)]
    untouched = np.setdiff1d(np.arange(nV), touched)
    vert_order = np.concatenate([touched, untouched]).astype(np.uint32)
    new_of_old = np.zeros(nV, np.int64); new_of_old[vert_order] = np.arange(nV)
    tet_new = new_of_old[T].astype(np.uint32)
    B, V0 = rest_shape(X, Tet)
    Br = B[tet_order].reshape(nT, 9); wr = (np.abs(V0) * mu.astype(np.float32))[tet_order]
    # tiles
    tile_tet_start = [0]; tile_rec_off = [0]; tiles = []; slot = 0; vlists = []
    vslots = [[] for _ in range(nV)]
    t0 = 0; max_local = 0
    seen_before = np.zeros(nV, bool)            # vertex already has a slot in an earlier tile
    while t0 < nT:
        seen = set(); t1 = t0
        while t1 < nT and t1 - t0 < TILE_T:
            new = [v for v in dict.fromkeys(tet_new[t1].tolist()) if v not in seen]
            if len(seen) + len(new) > TILE_NLMAX:
                break
            seen.update(new); t1 += 1
        while True:                              # shrink until the incidence rows fit
            nTets = t1 - t0
            tl_tets = tet_new[t0:t1].astype(np.int64)
            ids, counts = np.unique(tl_tets.reshape(-1), return_counts=True)
            order = np.lexsort((ids, -counts))   # in-tile incidence count descending, id ascending
            vl = ids[order]; cnt = counts[order]
            nLocal = len(vl); nGroups = (nLocal + TILE_GROUP - 1) // TILE_GROUP
            g_rows = [(int(cnt[TILE_GROUP * g]) + 2 * TILE_LPV - 1) // (2 * TILE_LPV) for g in range(nGroups)]
            g_base = np.concatenate([[0], np.cumsum(g_rows)]).astype(np.int64)
            nRows = int(g_base[-1])
            if nRows <= TILE_ROWSMAX or nTets == 1:
                break
            t1 = t0 + nTets // 2
        max_local = max(max_local, nLocal)
        lidx = np.zeros(nV, np.int64); lidx[vl] = np.arange(nLocal)
        cidx = lidx[tl_tets]                     # (nTets, 4) tile-local corner indices
        trec = np.zeros((nTets, 12), np.uint32)
        trec[:, :9] = Br[t0:t1].view(np.uint32)
        trec[:, 9] = wr[t0:t1].view(np.uint32)
        # incidence list of every tile-local vertex: ascending (tet, corner)
        owner = cidx.reshape(-1)
        o = np.argsort(owner, kind="stable")
        start = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=nLocal))])
        inc = [[(int(x) // 4, int(x) % 4) for x in o[start[l]:start[l + 1]]] for l in range(nLocal)]
        vlist = vl.astype(np.uint32)
        first = ~seen_before[vl]
        vlist = np.where(first, vlist | np.uint32(TILE_OWNER_BIT), vlist).astype(np.uint32)
        seen_before[vl] = True
        ab_bytes = TILE_OFF_TETS + 48 * nTets; c_bytes = 128 * nRows
        base = tile_rec_off[-1]
        gtab = np.zeros(16, np.uint32)
        for g in range(nGroups):
            gtab[g] = int(g_base[g]) | (g_rows[g] << 16)
        slot_base = len(tiles) * TILE_NLMAX       # padded slots: tile * TILE_NLMAX + tile-local vertex
        head = np.array([nTets, nLocal, slot_base, ab_bytes, c_bytes, nGroups, base & 0xffffffff, base >> 32], np.uint32).tobytes() + gtab.tobytes()
        # the H-scratch columns (bits 12..14 of the corner words, and the incidence entries built from them) are
        # the builder's free choice, constrained only by the conflict-freeness the tests check -> canonical form
        tiles.append(dict(head=head, tet40=trec[:, :10].copy(), corners=cidx.astype(np.int64), inc=inc,
                          g_base=g_base, g_rows=g_rows, ab_bytes=ab_bytes, c_bytes=c_bytes, base=base))
        vlists.append(np.concatenate([vlist, np.full(TILE_NLMAX - nLocal, 0xffffffff, np.uint32)]))   # slot-indexed, padded
        for l, v in enumerate(vl):
            vslots[int(v)].append(slot_base + l)
        slot += nLocal
        t0 = t1
        tile_tet_start.append(t0); tile_rec_off.append(base + ab_bytes + c_bytes)
    vslot_ptr = np.zeros(nV + 1, np.uint32)
    vslot_ptr[1:] = np.cumsum([len(s) for s in vslots])
    vslot = np.array([s for ss in vslots for s in ss], np.uint32)
    return dict(tet_order=tet_order, vert_order=vert_order, tet_new=tet_new,
                tile_tet_start=np.array(tile_tet_start, np.uint32), tile_rec_off=np.array(tile_rec_off, np.uint64),
                tiles=tiles, vslot_ptr=vslot_ptr, vslot=vslot, vlist=np.concatenate(vlists).astype(np.uint32),
                num_tiles=len(tiles), num_slots=slot, max_local=max_local, DmInv=B, V0=V0)


def partition_vertices(nV, world):
    return np.array([(nV * r) // world for r in range(world + 1)], np.int32)


def decode_and_check_records(records, tile_rec_off, tiles, vlist=None, vstage=None):
    """Decode the builder's packed tile records and compare them with the canonical tiles of build():
    header, group table, DmInv/w bits and corner indices bit for bit; the incidence rows entry by entry
    after mapping every H-scratch offset back to its (tet, corner); plus the bank-conflict freedom of the
    8-colouring (every quarter-warp STS.128 of phase B and LDS.128 of phase C hits 8 distinct 16-byte columns).
    The corner words hold STAGING slots, the builder's free choice like the H columns: with vlist/vstage given, the
    slot -> vertex map must be a bijection onto the tile's vertices and send every corner back to its vertex.
    Returns (position-load wavefronts, ideal) of phase B over all tiles (bank conflicts of the staging choice)."""
    rec = np.asarray(records, np.uint8)
    wf = ideal = 0
    for ti, T in enumerate(tiles):
        base = int(tile_rec_off[ti])
        assert base == T["base"] and int(tile_rec_off[ti + 1]) == base + T["ab_bytes"] + T["c_bytes"]
        assert rec[base:base + TILE_OFF_TETS].tobytes() == T["head"], ti
        nT = T["tet40"].shape[0]
        # three planes of 16 bytes per tet: plane p holds words 4p..4p+3 of every record
        tr = rec[base + TILE_OFF_TETS:base + TILE_OFF_TETS + 48 * nT].view(np.uint32).reshape(3, nT, 4).transpose(1, 0, 2).reshape(nT, 12)
        assert np.array_equal(tr[:, :10], T["tet40"]), ti
        halves = np.stack([tr[:, 10] & 0xffff, tr[:, 10] >> 16, tr[:, 11] & 0xffff, tr[:, 11] >> 16], 1).astype(np.int64)
        stage = (halves >> 4) & 0xff
        if vstage is None:
            assert np.array_equal(stage, T["corners"]), ti
        else:
            vl = np.asarray(vlist[256 * ti:256 * ti + 256]); vs = np.asarray(vstage[256 * ti:256 * ti + 256])
            nLocal = len(T["inc"])
            assert (vl[:nLocal] != 0xffffffff).all() and (vl[nLocal:] == 0xffffffff).all()
            assert sorted(vs[vs != 0xffffffff].tolist()) == sorted(vl[:nLocal].tolist()), ti         # bijection, owner bits kept
            assert np.array_equal(vs[stage], vl[T["corners"]]), ti                                  # corner -> slot -> its vertex
            for q in range(0, nT, 8):           # conflicts of the staged position loads: distinct slots per bank group
                for k in range(4):
                    sl = np.unique(stage[q:q + 8, k])
                    wf += int(np.bincount(sl % 8, minlength=8).max()); ideal += 1
        assert ((halves & 0x800f) == 0).all()
        col = (halves >> 12) & 7
        # stores: the 8 tets of a quarter-warp use 8 distinct columns per corner
        for q in range(0, nT, 8):
            for k in range(4):
                c = col[q:q + 8, k]
                assert len(set(c.tolist())) == len(c), (ti, q, k)
        nRows = T["c_bytes"] // 128
        incT = rec[base + T["ab_bytes"]:base + T["ab_bytes"] + T["c_bytes"]].view(np.uint16).reshape(nRows, 32, 2).astype(np.int64)
        # loads: per (row, half, quarter-warp) distinct addresses fall into distinct columns
        colsT = (incT // 16) % 8
        for r in range(nRows):
            for h in range(2):
                for qw in range(4):
                    a = incT[r, 8 * qw:8 * qw + 8, h]; c = colsT[r, 8 * qw:8 * qw + 8, h]
                    assert len(set(zip(a.tolist(), c.tolist()))) == len(set(c.tolist())), (ti, r, h, qw)
        assert ((incT % 16) == 0).all() and (incT < TILE_ZERO_OFF + 128).all()
        for l, lst in enumerate(T["inc"]):
            # vertex l: group l // TILE_GROUP; row r holds entries 2r, 2r+1 of its list in lane l % 32 (one lane per
            # vertex), or entries 4r, 4r+1 in lane l % 16 and 4r+2, 4r+3 in lane l % 16 + 16 (two lanes per vertex)
            g, lane = l // TILE_GROUP, l % TILE_GROUP
            rows = T["g_rows"][g]
            blk = incT[T["g_base"][g]:T["g_base"][g] + rows]
            ent = np.concatenate([blk[:, lane + TILE_GROUP * j, :] for j in range(TILE_LPV)], axis=1).reshape(-1)
            assert len(lst) <= len(ent)
            assert (ent[len(lst):] >= TILE_ZERO_OFF).all(), (ti, l)
            for e, (t, k) in enumerate(lst):
                assert ent[e] == k * TILE_HSTRIDE + ((t & ~7) | int(col[t, k])) * 16, (ti, l, e)
        # lanes beyond the tile's vertices hold only pads
        nLocal = len(T["inc"])
        for l in range(nLocal, TILE_GROUP * len(T["g_rows"])):
            g, lane = l // TILE_GROUP, l % TILE_GROUP
            blk = incT[T["g_base"][g]:T["g_base"][g] + T["g_rows"][g]]
            assert all((blk[:, lane + TILE_GROUP * j, :] >= TILE_ZERO_OFF).all() for j in range(TILE_LPV))
    return wf, ideal


def tile_table(tiles):
    """layout.cpp:build_tile_table restated from the canonical tiles of build(): per tile 12 words."""
    out = np.zeros((len(tiles), 4 + TILE_NLMAX // TILE_GROUP), np.uint32)
    for ti, T in enumerate(tiles):
        nTets = T["tet40"].shape[0]; nLocal = len(T["inc"])
        assert T["base"] % 16 == 0
        out[ti, 0] = T["base"] // 16
        out[ti, 1] = T["ab_bytes"] | (T["c_bytes"] << 16)
        out[ti, 2] = nTets | (nLocal << 16)
        for g in range(len(T["g_rows"])):
            n_valid = min(TILE_GROUP, nLocal - g * TILE_GROUP)
            out[ti, 4 + g] = int(T["g_base"][g]) | (int(T["g_rows"][g]) << 6) | (n_valid << 12)
        out[ti, 4:12] |= np.uint32(nTets << 18)
    return out
