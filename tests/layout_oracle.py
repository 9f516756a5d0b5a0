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
TILE_OFF_TETS = 80


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
    tile_tet_start = [0]; tile_rec_off = [0]; recs = []; slot = 0
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
            nLocal = len(vl); nGroups = (nLocal + 31) // 32
            g_rows = [(int(cnt[32 * g]) + 1) // 2 for g in range(nGroups)]
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
        trec[:, 10] = (cidx[:, 0] * 16) | ((cidx[:, 1] * 16) << 16)
        trec[:, 11] = (cidx[:, 2] * 16) | ((cidx[:, 3] * 16) << 16)
        # transposed incidence rows: entry e of local vertex l -> row g_base[l//32] + e//2, lane l%32, half e%2
        incT = np.full((nRows, 32, 2), TILE_ZERO_OFF, np.uint16)
        owner = cidx.reshape(-1)
        tl = np.repeat(np.arange(nTets), 4); k = np.tile(np.arange(4), nTets)
        swz = tl ^ ((tl >> 3) & 7)
        ent = (k * TILE_HSTRIDE + swz * 16).astype(np.uint16)
        o = np.argsort(owner, kind="stable")     # per vertex, ascending (tet, corner)
        ow = owner[o]
        start = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=nLocal))])
        e = np.arange(len(o)) - start[ow]
        incT[g_base[ow // 32] + e // 2, ow % 32, e % 2] = ent[o]
        vlist = vl.astype(np.uint32)
        first = ~seen_before[vl]
        vlist = np.where(first, vlist | np.uint32(TILE_OWNER_BIT), vlist).astype(np.uint32)
        seen_before[vl] = True
        ab_bytes = TILE_OFF_TETS + 48 * nTets + rup(4 * nLocal, 16); c_bytes = 128 * nRows
        base = tile_rec_off[-1]
        gtab = np.zeros(12, np.uint32)
        for g in range(nGroups):
            gtab[g] = int(g_base[g]) | (g_rows[g] << 16)
        rec = bytearray()
        rec += np.array([nTets, nLocal, slot, ab_bytes, c_bytes, nGroups, base & 0xffffffff, base >> 32], np.uint32).tobytes()
        rec += gtab.tobytes()
        rec += trec.tobytes()
        rec += vlist.tobytes() + b"\0" * (rup(4 * nLocal, 16) - 4 * nLocal)
        assert len(rec) == ab_bytes
        rec += incT.tobytes()
        assert len(rec) == ab_bytes + c_bytes
        recs.append(bytes(rec))
        for l, v in enumerate(vl):
            vslots[int(v)].append(slot + l)
        slot += nLocal
        t0 = t1
        tile_tet_start.append(t0); tile_rec_off.append(base + ab_bytes + c_bytes)
    vslot_ptr = np.zeros(nV + 1, np.uint32)
    vslot_ptr[1:] = np.cumsum([len(s) for s in vslots])
    vslot = np.array([s for ss in vslots for s in ss], np.uint32)
    return dict(tet_order=tet_order, vert_order=vert_order, tet_new=tet_new,
                tile_tet_start=np.array(tile_tet_start, np.uint32), tile_rec_off=np.array(tile_rec_off, np.uint64),
                records=np.frombuffer(b"".join(recs), np.uint8), vslot_ptr=vslot_ptr, vslot=vslot,
                num_tiles=len(recs), num_slots=slot, max_local=max_local, DmInv=B, V0=V0)


def partition_vertices(nV, world):
    return np.array([(nV * r) // world for r in range(world + 1)], np.int32)
