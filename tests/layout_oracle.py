"""numpy restatement of the device-layout specification -- the oracle for
the bit-exact reorder / renumber / tile / incidence / slot checks.  Independent of the C++ builder
(soft-body-simulation-cuda_b200/csrc/layout.cpp): integer work in numpy, float32 ops one rounding each."""
import numpy as np

TILE_T = 256
TILE_NLMAX = 384
TILE_HSTRIDE = TILE_T * 16


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
    """solverUtil.cuh:98-116 with glm's cofactor inverse, float32, one rounding per op."""
    X = X.astype(np.float32); T = Tet.astype(np.int64)
    x0 = X[T[:, 0]]
    m = [X[T[:, 1]] - x0, X[T[:, 2]] - x0, X[T[:, 3]] - x0]   # m[c][:, r]
    M = lambda c, r: m[c][:, r]
    det = (M(0, 0) * (M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) - M(1, 0) * (M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2))
           + M(2, 0) * (M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2)))
    ood = np.float32(1.0) / det
    inv = {}
    inv[0, 0] = +(M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) * ood
    inv[1, 0] = -(M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2)) * ood
    inv[2, 0] = +(M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)) * ood
    inv[0, 1] = -(M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2)) * ood
    inv[1, 1] = +(M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2)) * ood
    inv[2, 1] = -(M(0, 0) * M(2, 1) - M(2, 0) * M(0, 1)) * ood
    inv[0, 2] = +(M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2)) * ood
    inv[1, 2] = -(M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2)) * ood
    inv[2, 2] = +(M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1)) * ood
    B = np.zeros((T.shape[0], 3, 3), np.float32)      # row-major B[r][c] = inv[c][r]
    for r in range(3):
        for c in range(3):
            B[:, r, c] = inv[c, r]
    V0 = np.abs(det) / np.float32(6.0)
    return B, V0.astype(np.float32)


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
    while t0 < nT:
        seen = set(); t1 = t0
        while t1 < nT and t1 - t0 < TILE_T:
            new = [v for v in dict.fromkeys(tet_new[t1].tolist()) if v not in seen]
            if len(seen) + len(new) > TILE_NLMAX:
                break
            seen.update(new); t1 += 1
        tl_tets = tet_new[t0:t1].astype(np.int64)
        nTets = t1 - t0
        ids, counts = np.unique(tl_tets.reshape(-1), return_counts=True)
        order = np.lexsort((ids, -counts))                 # in-tile incidence count descending, id ascending
        vl = ids[order].astype(np.uint32)
        nLocal = len(vl)
        max_local = max(max_local, nLocal)
        lidx = np.zeros(nV, np.int64); lidx[vl] = np.arange(nLocal)
        cidx = lidx[tl_tets]                               # (nTets, 4) tile-local corner indices
        # 48-byte tet records: B[9], w, c01, c23 (corner index * 16, two u16 per word)
        trec = np.zeros((nTets, 12), np.uint32)
        trec[:, :9] = Br[t0:t1].view(np.uint32)
        trec[:, 9] = wr[t0:t1].view(np.uint32)
        trec[:, 10] = (cidx[:, 0] * 16) | ((cidx[:, 1] * 16) << 16)
        trec[:, 11] = (cidx[:, 2] * 16) | ((cidx[:, 3] * 16) << 16)
        # incidence lists: per tile-local vertex, ascending (tet, corner); entry = corner*HSTRIDE + tet*16
        owner = cidx.reshape(-1)
        tl = np.repeat(np.arange(nTets), 4); k = np.tile(np.arange(4), nTets)
        ent = (k * TILE_HSTRIDE + tl * 16).astype(np.uint16)
        inc = ent[np.argsort(owner, kind="stable")]
        inc_off = np.zeros(nLocal + 1, np.uint16)
        inc_off[1:] = np.cumsum(np.bincount(owner, minlength=nLocal)).astype(np.uint16)
        i_bytes = rup(8 * nTets, 16); io_bytes = rup(2 * (nLocal + 1), 16); v_bytes = rup(4 * nLocal, 16)
        rec_bytes = 16 + 48 * nTets + i_bytes + io_bytes + v_bytes
        rec = bytearray()
        rec += np.array([nTets, nLocal, slot, rec_bytes], np.uint32).tobytes()
        rec += trec.tobytes()
        rec += inc.tobytes() + b"\0" * (i_bytes - 8 * nTets)
        rec += inc_off.tobytes() + b"\0" * (io_bytes - 2 * (nLocal + 1))
        rec += vl.tobytes() + b"\0" * (v_bytes - 4 * nLocal)
        assert len(rec) == rec_bytes
        recs.append(bytes(rec))
        for l, v in enumerate(vl):
            vslots[int(v)].append(slot + l)
        slot += nLocal
        t0 = t1
        tile_tet_start.append(t0); tile_rec_off.append(tile_rec_off[-1] + rec_bytes)
    vslot_ptr = np.zeros(nV + 1, np.uint32)
    vslot_ptr[1:] = np.cumsum([len(s) for s in vslots])
    vslot = np.array([s for ss in vslots for s in ss], np.uint32)
    return dict(tet_order=tet_order, vert_order=vert_order, tet_new=tet_new,
                tile_tet_start=np.array(tile_tet_start, np.uint32), tile_rec_off=np.array(tile_rec_off, np.uint64),
                records=np.frombuffer(b"".join(recs), np.uint8), vslot_ptr=vslot_ptr, vslot=vslot,
                num_tiles=len(recs), num_slots=slot, max_local=max_local, DmInv=B, V0=V0)


def partition_vertices(nV, world):
    return np.array([(nV * r) // world for r in range(world + 1)], np.int32)
