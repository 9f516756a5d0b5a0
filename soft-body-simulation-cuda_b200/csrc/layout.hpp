// Device data layout for the PD step: tet reordering, vertex renumbering, tile-packed tet
// stream, tile-local incidence CSR, partial-sum slots, scalar system matrix, vertex partition.
// All of it is integer/index work that tests check bit-for-bit against an independent numpy
// restatement (tests/layout_oracle.py).  The specification lives in DESIGN.md section 3.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace pdb200 {

constexpr int TILE_T = 256;       // tets per tile (= threads per CTA of the local kernel)
constexpr int TILE_NLMAX = 256;   // max distinct vertices per tile (a tile is closed early beyond); = TILE_T: one vertex per thread
#ifndef PD_TILE_GROUP
#define PD_TILE_GROUP 32
#endif
constexpr int TILE_GROUP = PD_TILE_GROUP;             // tile-local vertices are summed in groups of 32 (one warp, one lane per vertex; measured
                                                     // faster on grid139) or 16 (two lanes per vertex: half the longest list, more instructions)
constexpr int TILE_LPV = 32 / TILE_GROUP;             // lanes per vertex in phase C
constexpr int TILE_NGROUPS = TILE_NLMAX / TILE_GROUP;   // warp w of the local kernel sums groups w (and w + 8)
static_assert(TILE_GROUP == 32 || TILE_GROUP == 16, "phase C group size");
constexpr int TILE_ROWSMAX = 56;  // max incidence rows per tile (a tile is shrunk beyond)
constexpr uint32_t TILE_HCOLOURS = 8;
constexpr uint32_t TILE_HSTRIDE = TILE_T * 16;           // bytes between corner planes of the per-tile H scratch
constexpr uint32_t TILE_ZERO_OFF = 4u * TILE_HSTRIDE;    // byte offset of the all-zero float4 padding entries point at
constexpr uint32_t TILE_HS_BYTES = TILE_ZERO_OFF + 128u; // four corner planes + eight zero slots, one per column
constexpr uint32_t TILE_OWNER_BIT = 0x80000000u;         // vlist flag: this slot is the vertex's first (owner) slot

// One packed tile record = two parts, each moved to shared memory by ONE
// Part AB (what phase B of the local kernel reads: header and group table on the host side only, the tet records
// straight from global memory):
//   +0    TileHeader                                                          32 B
//   +32   group table u32[16]: rowBase | nRows << 16 of each TILE_GROUP-vertex group     64 B
//   +96   tet records, 48 B (12 words) each: f32 B[9] (DmInv, row-major), f32 w = |V0|*mu, u32 c01, u32 c23 --
//         four 16-bit corner words: bits 4..11 = STAGING slot of the corner's vertex (so `word & 0xff0` is the byte
//         offset into the staged vertex array; the slot -> vertex map is Layout::vstage), bits 12..14 = H-scratch
//         column of this corner's contribution.  Staging slots are chosen per tile (layout.cpp:stage_slots) so that
//         the 8 tets of a quarter-warp read their corner-k positions from 8 different 16-byte bank groups wherever
//         the mesh allows (a heuristic 8-colouring; unlike the H columns it is not always conflict free).
//         Stored as THREE PLANES of 16 bytes per tet (plane p holds words 4p..4p+3 of every tet of the tile,
//         tile_tet_word()), so that thread t of the local kernel reads its record with three fully coalesced
//         16-byte loads straight from global memory (plane stride = 16 * nTets).
// The tile's vertex list lives OUTSIDE the record, in the global array Layout::vlist indexed by slot.
// Slots are PADDED per tile: slot = tile * TILE_NLMAX + tile-local vertex, so that the local kernel needs no
// per-tile offset to find them; entry = u32 global (renumbered) vertex id | TILE_OWNER_BIT, or 0xffffffff for
// the unused tail; tile-local vertices ordered by (in-tile incidence count descending, id ascending).  The
// position gather of the local kernel reads the same entries indexed by STAGING slot (Layout::vstage) with plain
// coalesced loads three tiles ahead of use.  Unused slots of the partial-sum array are never read or written.
// Part C (phase C, double buffered): the tile-local incidence lists, transposed per group of TILE_GROUP
// vertices.  TILE_GROUP 32: row r of group g holds, for each of the 32 lanes (vertices), entries 2r and 2r+1 of
// that vertex's list packed as two u16 in one u32.  TILE_GROUP 16: a vertex is summed by TWO lanes of a warp
// (combined by one shuffle): row r holds, for vertex l, entries 4r and 4r+1 of its list in lane l % 16 and
// entries 4r+2 and 4r+3 in lane l % 16 + 16.  An entry is the byte offset of one tet-corner
// contribution in the H scratch, tile_h_offset(tet, corner, column): the 8 tets of a quarter-warp share one
// 128-byte line per corner and the column (0..7) inside it comes from an 8-colouring (layout.cpp:color_tile)
// that makes every quarter-warp STS.128 of phase B and every quarter-warp LDS.128 of phase C conflict free.
// Lists are ascending in (tet, corner) and padded with TILE_ZERO_OFF + 16*c (eight zero slots, one per
// column, c = a column the pad's quarter-warp does not use).  128 B per row.
struct TileHeader {
    uint32_t nTets, nLocal, slotBase, abBytes, cBytes, nGroups, offLo, offHi;   // off = this record's byte offset in the stream
};
struct TileEntry {   // per-tile entry of the device tile table (one uint4)
    uint64_t off;    // byte offset of part AB in the record stream (part C follows at off + abBytes)
    uint32_t abBytes, cBytes;
};
inline uint32_t rup16(uint32_t x) { return (x + 15u) & ~15u; }
constexpr uint32_t TILE_OFF_TETS = 96u;
inline uint32_t tile_ab_bytes(uint32_t nTets) { return TILE_OFF_TETS + 48u * nTets; }
// byte offset (inside part AB) of word j (0..11) of the record of tile-local tet tl in a tile of nTets tets
inline size_t tile_tet_word(uint32_t nTets, uint32_t tl, uint32_t j) { return TILE_OFF_TETS + 16u * ((size_t)tl + (size_t)nTets * (j / 4u)) + 4u * (j % 4u); }
inline uint32_t tile_h_offset(uint32_t tet, uint32_t corner, uint32_t column) { return corner * TILE_HSTRIDE + ((tet & ~7u) | column) * 16u; }
inline uint32_t tile_corner_half(uint32_t localVertex, uint32_t column) { return (localVertex << 4) | (column << 12); }
inline uint32_t tile_corner_colour(uint32_t half) { return (half >> 12) & 7u; }
// TMA landing buffers (multiples of 128 B)
constexpr uint32_t TILE_ABMAX = ((TILE_OFF_TETS + 48u * TILE_T) + 127u) & ~127u;
constexpr uint32_t TILE_CMAX = 128u * TILE_ROWSMAX;

struct Layout {
    int nV = 0, nT = 0;
    // permutations: new index -> old index
    std::vector<uint32_t> tetOrder;     // nT
    std::vector<uint32_t> vertOrder;    // nV
    std::vector<uint32_t> vertNewOfOld; // nV (inverse of vertOrder)
    std::vector<uint32_t> tetNew;       // 4*nT, reordered tets in renumbered vertex ids
    // tiles
    int nTiles = 0;
    std::vector<uint32_t> tileTetStart; // nTiles+1, into the reordered tet list
    std::vector<uint64_t> tileRecOff;   // nTiles+1, byte offsets into `records` (16-B aligned)
    std::vector<TileEntry> tileTab;     // nTiles, what the local kernel's producer thread reads
    std::vector<uint8_t> records;       // packed tile records
    // partial-sum slots: slot = tile * TILE_NLMAX + localVertex
    uint32_t nSlots = 0;                // used slots (sum of the tiles' distinct-vertex counts)
    std::vector<uint32_t> vslotPtr;     // nV+1   vertex -> slots CSR (ascending slots)
    std::vector<uint32_t> vslot;        // nSlots
    std::vector<uint32_t> vstage;       // nTiles * TILE_NLMAX   STAGING slot -> the vlist entry of the vertex staged there (0xffffffff: none);
                                        // what the local kernel's position gather reads (stage_slots() in layout.cpp)
    std::vector<uint32_t> vlist;        // nTiles * TILE_NLMAX   slot -> vertex id | TILE_OWNER_BIT (first slot of the vertex), 0xffffffff unused
    int maxLocal = 0;
    // a rank's layout only (extract_rank_layout): position of every local tet in the GLOBAL reordered tet list, so that
    // float sums over a vertex's tets (matrix_diag) can be taken in the same order on every world size
    std::vector<uint32_t> tetGlobal;    // nT, empty for a single-GPU layout
};

// Per-tile entry of the DEVICE tile table the local kernel reads (build_tile_table): record offset / 16, part AB bytes |
// part C bytes << 16, tet count | vertex count << 16, spare, then one word per vertex group (warp w sums group w -- word
// 4 + w -- and, with 16-vertex groups, w + 8): rowBase | nRows << 6 | vertices in the group << 12, and in the first
// eight also the tile's tet count << 18
constexpr int TILE_META_WORDS = 4 + TILE_NGROUPS;
void build_tile_table(const Layout& L, std::vector<uint32_t>& meta);

// Reordering key of a tet: 30-bit Morton code of its quantised centroid
void morton_keys(const float* X, const uint32_t* Tet, int nT, std::vector<uint32_t>& keys);

// rest-shape products in the reference's arithmetic (solverUtil.cuh:98-116): DmInv row-major
void rest_shape(const float* X, const uint32_t* Tet, int nT, float* DmInv, float* V0);

// Build the whole layout.  X: 3*nV rest positions (original numbering); mu per tet.
// reorder=false keeps the input tet order (tiles are then consecutive input tets).
void build_layout(int nV, int nT, const float* X, const uint32_t* Tet, const float* mu, bool reorder, Layout& out);

// scalar system matrix A^ = diag(c) + sum_t w_t P^T (B^T G)^T (B^T G) P  (pdUtil.cu:9-54) in the
// renumbered vertex ids, CSR with ascending columns; duplicates summed in ascending reordered-tet order.
struct CsrMatrix {
    int n = 0;
    std::vector<int> rowPtr, col;
    std::vector<float> val;
};
// matrix_diag of PdSolver::SolverPrepare (pdUtil.cu:16-24) from the tile records, per layout-local vertex
void matrix_diag_host(const Layout& L, std::vector<float>& md);
void build_system_matrix(const Layout& L, const float* Xnew, const float* DmInv /*reordered*/, const float* w /*reordered*/,
                         const float* c /*nV renumbered*/, CsrMatrix& A, std::vector<float>& matrixDiag);

// Sparse Cholesky A^ = L L^T of the scalar system matrix in the given (Morton first-touch) vertex order, no further
// permutation: up-looking factorisation over the elimination tree, accumulated in double, stored as float.
// lPtr/lCol/lVal: L by rows (ascending columns, diagonal last); uPtr/uCol/uVal: L^T by rows (diagonal first).
// This is the host-side "prefactor once" of CholeskySpLinearSolver (cholesky.cu:133-158) / SimplicialCholesky
// (pdSolver.cu:103); throws if the matrix is not positive definite.
struct CholFactor {
    int n = 0;
    std::vector<int> lPtr, lCol, uPtr, uCol;
    std::vector<float> lVal, uVal;
};
void cholesky_factor(const CsrMatrix& A, CholFactor& F);
// Fill-reducing ordering for the factorisation (the reference orders with AMD inside cusolverSpXcsrcholAnalysis /
// Eigen::SimplicialCholesky, cholesky.cu:72-131, pdSolver.cu:103): GEOMETRIC nested dissection -- the vertex set is cut at
// the median of its longest coordinate axis, the vertices of the smaller boundary layer between the halves become the
// separator and are numbered LAST, recursively (leaves of <= 48 vertices keep their order).  xyz: n x 3 rest positions of
// the matrix's rows.  perm[new] = old.  Deterministic; a mesh graph needs nothing cleverer (nnz(L) on a 24^3-cell grid:
// 11.9 M in Morton order, 2.4 M here).
void nested_dissection_order(const CsrMatrix& A, const float* xyz, std::vector<int>& perm);
// B = P A P^T for perm[new] = old (rows of B sorted by column)
void permute_symmetric(const CsrMatrix& A, const std::vector<int>& perm, CsrMatrix& B);

// Per-body data of a scene made of many small soft bodies, one CTA per body (pd_body_kernel.cuh).
struct BodyDesc {           // device-visible POD, one per body
    uint32_t v0, nV;        // its vertices: entries [v0, v0 + nV) of BodyBatch::verts / md; body-local id = position
    uint32_t ptr0;          // its incidence pointers: incPtr[ptr0 .. ptr0 + nV], relative to inc0
    uint32_t nT;            // its tets
    uint32_t inc0;          // its incidence entries start at inc[inc0]; entry = body-local tet * 4 + corner
    uint32_t recOff16;      // byte offset / 16 of its record planes in BodyBatch::rec (plane p of tet t: 16 * (p * nT + t))
    uint32_t pad0, pad1;
};
struct BodyBatch {
    std::vector<BodyDesc> bodies;
    std::vector<uint32_t> verts;        // engine (renumbered) vertex id of every body-local vertex
    std::vector<uint8_t> rec;           // per tet 48 B in three planes: DmInv[9] (row-major), w = |V0| mu, four u16 body-local vertex ids
    std::vector<uint32_t> incPtr;
    std::vector<uint16_t> inc;          // ascending (tet, corner) in the ORIGINAL tet order: the reference's sequential scatter order
    std::vector<float> md;              // matrix_diag per body-local vertex, summed in that same order
    uint32_t nVmax = 0, nTmax = 0;      // capacities the kernel's shared-memory carve-up is sized for (multiples of 32)
};
// The soft bodies of a scene as its CONNECTED COMPONENTS (tets sharing a vertex): without mesh-mesh collision the PD system
// is block diagonal per component whatever the scene file calls a body.  Returns false when a component is not a contiguous
// vertex range (then the per-body kernel does not apply); starts = first vertex id of every component, ascending.
bool connected_body_ranges(int nV, int nT, const uint32_t* Tet, std::vector<int>& starts);
// bodyVertStart: first ORIGINAL vertex id of every body (ascending; bodies own contiguous vertex and tet ranges, as
// DataLoader::AllocData lays them out).  Throws if a tet spans two bodies or a body exceeds the u16 index ranges.
void build_body_batch(int nV, int nT, const float* X, const uint32_t* Tet, const float* mu, const std::vector<int>& bodyVertStart,
                      const uint32_t* vertNewOfOld, BodyBatch& out);

// Contiguous vertex partition of the renumbered ids into `world` ranks:
// rank r owns [vbeg[r], vbeg[r+1]).
void partition_vertices(int nV, int world, std::vector<int>& vbeg);

// Multi-GPU plan of one rank.  The partition is by VERTEX (contiguous ranges of the
// renumbered ids, which follow the Morton order of the tets and are therefore spatially compact); a rank
// evaluates every global TILE that touches one of its vertices, unchanged -- same record bytes, same order --
// so the partial sums and the per-vertex slot sums of its own vertices are bit-identical to the single-GPU
// run.  Vertices of those tiles that belong to other ranks are ghosts: their positions are pushed by their
// owners once per PD iteration.  Local vertex numbering: [own range in global order | ghosts ascending].
struct RankPlan {
    int rank = 0, world = 1;
    bool trim = false;                 // PD_DIST_TRIM experiment: trimmed ghosts here, trimmed + re-packed boundary tiles in extract_rank_layout
    std::vector<int> vbeg;             // world + 1
    int nOwn = 0, nGhost = 0;
    std::vector<uint32_t> tiles;       // global tile ids evaluated by this rank: first the interior tiles (all vertices owned),
                                       // then the boundary tiles (some ghost vertex), each group ascending
    int nInteriorTiles = 0;            // the local kernel needs the neighbours' halo only from tile nInteriorTiles on
    std::vector<uint32_t> ghosts;      // global (renumbered) ids of the ghost vertices, ascending
    std::vector<int> nLocOf;           // world: owned + ghost vertex count of every rank (sizes of the peers' windows)
    std::vector<int> neighbours;       // ranks exchanged with (symmetric), ascending
    // push list, sorted by (rank, dst): owned local vertex `src` here is ghost local vertex `dst` on `rank`
    std::vector<uint32_t> pushSrc, pushDst;
    std::vector<int> pushRank;
};
// trim (the engine's default, dist_trim_from_env): the ghosts of a rank are only the vertices of the tets it keeps (the tets that touch a
// vertex it owns), not every vertex of its boundary tiles -- the halo pushes shrink accordingly.  Same tiles either way.
void build_rank_plan(const Layout& G, int world, int rank, RankPlan& plan, bool trim = false);
// the rank's own Layout: selected tiles (headers re-based), local vertex ids, local slots; vertOrder maps local ->
// ORIGINAL vertex ids so that everything downstream of a Layout works unchanged
// plan.trim (the engine and pd_rank_plan_build take it from dist_trim_from_env: ON unless PD_DIST_TRIM=0): a boundary
// tile is cut down to the tets that touch a vertex this rank OWNS (the others only feed ghost vertices, whose sums are never
// read), and consecutive trimmed tiles are packed into physical tiles of up to TILE_T tets.  Every (global tile, vertex)
// pair keeps a slot of its own with its incidence list in the original order, so the partial sums -- and the vertex sums
// over the slots in ascending global tile order -- stay bit-identical to the single-GPU run.  Redundant tets on the 139^3
// grid at N = 8: 13 % -> 2.7 % (measured).
void extract_rank_layout(const Layout& G, const RankPlan& plan, Layout& out);      // trims iff plan.trim
bool dist_trim_from_env();

}  // namespace pdb200
