// Device data layout for the PD step: tet reordering, vertex renumbering, tile-packed tet
// stream, tile-local incidence CSR, partial-sum slots, scalar system matrix, vertex partition.
// All of it is integer/index work that tests check bit-for-bit against an independent numpy
// restatement (tests/layout_oracle.py).  The specification lives in DESIGN.md section 3.
#pragma once
#include <cstdint>
#include <vector>

namespace pdb200 {

constexpr int TILE_T = 256;       // tets per tile (= threads per CTA of the local kernel)
constexpr int TILE_NLMAX = 256;   // max distinct vertices per tile (a tile is closed early beyond); = TILE_T: one vertex per thread
constexpr int TILE_NGROUPS = TILE_NLMAX / 32;   // tile-local vertices are handled in groups of 32 (one warp)
constexpr int TILE_ROWSMAX = 56;  // max incidence rows per tile (a tile is shrunk beyond)
constexpr uint32_t TILE_HSTRIDE = TILE_T * 16;           // bytes between corner planes of the per-tile H scratch
constexpr uint32_t TILE_ZERO_OFF = 4u * TILE_HSTRIDE;    // byte offset of the all-zero float4 padding entries point at
constexpr uint32_t TILE_OWNER_BIT = 0x80000000u;         // vlist flag: this slot is the vertex's first (owner) slot

// One packed tile record = two parts, each moved to shared memory by ONE
// bulk copy (TMA).  Part AB (phases A and B of the local kernel, double buffered):
//   +0    TileHeader                                                          32 B
//   +32   group table u32[12]: rowBase | nRows << 16 of each 32-vertex group (8 used)  48 B
//   +80   tet records, 48 B each: f32 B[9] (DmInv, row-major), f32 w = |V0|*mu, u32 c01, u32 c23 --
//         the four tile-local corner indices * 16 (byte offsets into the staged vertex array)
//   offV  vlist u32[nLocal]: global (renumbered) vertex id of each tile-local vertex | TILE_OWNER_BIT,
//         ordered by (in-tile incidence count descending, id ascending)     (pad 16)
// Part C (phase C, single buffered): the tile-local incidence lists, transposed per group of 32
// vertices: row r of group g holds, for each of the 32 lanes (vertices), entries 2r and 2r+1 of that
// vertex's list packed as two u16 in one u32.  An entry is the byte offset of one tet-corner
// contribution in the H scratch, corner*TILE_HSTRIDE + swz(tet)*16 with swz(t) = t ^ ((t>>3)&7);
// lists are ascending in (tet, corner) and padded with TILE_ZERO_OFF.  128 B per row.
struct TileHeader {
    uint32_t nTets, nLocal, slotBase, abBytes, cBytes, nGroups, offLo, offHi;   // off = this record's byte offset in the stream
};
struct TileEntry {   // per-tile entry of the device tile table
    uint64_t off;    // byte offset of part AB in the record stream (part C follows at off + abBytes)
    uint32_t abBytes, cBytes;
};
inline uint32_t rup16(uint32_t x) { return (x + 15u) & ~15u; }
constexpr uint32_t TILE_OFF_TETS = 80u;
inline uint32_t tile_off_vlist(uint32_t nTets) { return TILE_OFF_TETS + 48u * nTets; }
inline uint32_t tile_ab_bytes(uint32_t nTets, uint32_t nLocal) { return tile_off_vlist(nTets) + rup16(4u * nLocal); }
inline uint32_t tile_swz(uint32_t t) { return t ^ ((t >> 3) & 7u); }
// TMA landing buffers (multiples of 128 B)
constexpr uint32_t TILE_ABMAX = ((TILE_OFF_TETS + 48u * TILE_T + 4u * TILE_NLMAX) + 127u) & ~127u;
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
    // partial-sum slots: slot = slotBase[tile] + localVertex
    uint32_t nSlots = 0;
    std::vector<uint32_t> vslotPtr;     // nV+1   vertex -> slots CSR (ascending slots)
    std::vector<uint32_t> vslot;        // nSlots
    int maxLocal = 0;
};

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
void build_system_matrix(const Layout& L, const float* Xnew, const float* DmInv /*reordered*/, const float* w /*reordered*/,
                         const float* c /*nV renumbered*/, CsrMatrix& A, std::vector<float>& matrixDiag);

// Contiguous vertex partition of the renumbered ids into `world` ranks:
// rank r owns [vbeg[r], vbeg[r+1]).
void partition_vertices(int nV, int world, std::vector<int>& vbeg);

}  // namespace pdb200
