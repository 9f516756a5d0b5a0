// Device data layout for the PD step: tet reordering, vertex renumbering, tile-packed tet
// stream, tile-local incidence CSR, partial-sum slots, scalar system matrix, vertex partition.
// All of it is integer/index work that tests check bit-for-bit against an independent numpy
// restatement (tests/layout_oracle.py).  The specification lives in DESIGN.md section 3.
#pragma once
#include <cstdint>
#include <vector>

namespace pdb200 {

constexpr int TILE_T = 256;      // tets per tile (= threads per CTA of the local kernel)
constexpr int TILE_NLMAX = 384;  // max distinct vertices per tile (a tile is closed early beyond)
constexpr int TILE_HSTRIDE = TILE_T * 16;   // bytes between corner planes of the per-tile H scratch

// One packed tile record, moved to shared memory by ONE bulk copy:
//   +0      TileHeader                                                   16 B
//   +16     tet records, 48 B each: f32 B[9] (DmInv, row-major), f32 w = |V0|*mu,
//           u32 c01, u32 c23 -- the four tile-local corner indices * 16 (byte offsets into the
//           staged vertex array), two u16 per word                       48*nTets
//   offI    inc  u16[4*nTets]: per tile-local vertex, ascending (tet, corner), each entry the byte
//           offset corner*TILE_HSTRIDE + tet*16 of that contribution in the H scratch   (pad 16)
//   offIO   incOff u16[nLocal+1]: tile-local incidence CSR offsets        (pad 16)
//   offV    vlist u32[nLocal]: global (renumbered) id of each tile-local vertex, ordered by
//           (in-tile incidence count descending, id ascending) so that the lanes of a warp
//           walk incidence lists of similar length                        (pad 16)
struct TileHeader {
    uint32_t nTets, nLocal, slotBase, recBytes;
};
inline uint32_t rup16(uint32_t x) { return (x + 15u) & ~15u; }
inline uint32_t tile_off_inc(uint32_t nTets) { return 16u + 48u * nTets; }
inline uint32_t tile_off_incoff(uint32_t nTets) { return tile_off_inc(nTets) + rup16(8u * nTets); }
inline uint32_t tile_off_vlist(uint32_t nTets, uint32_t nLocal) { return tile_off_incoff(nTets) + rup16(2u * (nLocal + 1u)); }
inline uint32_t tile_rec_bytes(uint32_t nTets, uint32_t nLocal) { return tile_off_vlist(nTets, nLocal) + rup16(4u * nLocal); }
// largest possible record, rounded to 128 B: the size of one TMA landing buffer
constexpr uint32_t TILE_RECMAX = ((16u + 48u * TILE_T + 8u * TILE_T + ((2u * (TILE_NLMAX + 1u) + 15u) & ~15u) + 4u * TILE_NLMAX) + 127u) & ~127u;

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
