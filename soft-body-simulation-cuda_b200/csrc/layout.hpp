// Device data layout for the PD step: tet reordering, vertex renumbering, tile-packed tet
// stream, tile-local incidence CSR, partial-sum slots, scalar system matrix, vertex partition.
// All of it is integer/index work that tests check bit-for-bit against an independent numpy
// restatement (tests/layout_oracle.py).  The specification lives in DESIGN.md section 3.
#pragma once
#include <cstdint>
#include <vector>

namespace pdb200 {

constexpr int TILE_T = 256;     // tets per tile (= threads per CTA of the local kernel)
constexpr int TILE_NLMAX = 512; // max distinct vertices per tile (tile is closed early beyond)

// Header of one packed tile record (16 bytes), followed by the sections described in DESIGN.md:
//   vlist   u32[nLp]            global (renumbered) vertex id of each tile-local vertex, ascending
//   cidx    u16[4][nTp] as uint2[nTp]   tile-local corner indices of each tet
//   Bm      f32[9][nTp]         DmInv, row-major entries, SoA over tets
//   w       f32[nTp]            |V0| * mu
//   incOff  u16[nLocal+1]       tile-local incidence CSR offsets   (padded to 16 B)
//   inc     u16[4*nTets]        entries tetLocal*4 + corner, ascending (padded to 16 B)
// nLp = roundup(nLocal,4), nTp = roundup(nTets,4).
struct TileHeader {
    uint32_t nTets, nLocal, slotBase, recBytes;
};

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
