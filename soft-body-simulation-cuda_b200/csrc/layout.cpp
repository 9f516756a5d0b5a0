// Layout builder (see layout.hpp and DESIGN.md section 3).  Compiled with -ffp-contract=off so
// the few float operations here (centroids, rest shape) round exactly once per operation and
// can be reproduced by numpy float32 arithmetic in the tests.
#include "layout.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>

namespace pdb200 {

static inline uint32_t spread3(uint32_t x)
{   // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

void morton_keys(const float* X, const uint32_t* Tet, int nT, std::vector<uint32_t>& keys)
{
    keys.assign((size_t)nT, 0u);
    if (nT == 0) return;
    std::vector<float> cen((size_t)nT * 3);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int t = 0; t < nT; ++t) {
        const uint32_t* tv = Tet + 4 * (size_t)t;
        for (int k = 0; k < 3; ++k) {
            const float c = ((X[3 * (size_t)tv[0] + k] + X[3 * (size_t)tv[1] + k]) +
                             (X[3 * (size_t)tv[2] + k] + X[3 * (size_t)tv[3] + k])) * 0.25f;
            cen[3 * (size_t)t + k] = c;
            lo[k] = std::min(lo[k], c);
            hi[k] = std::max(hi[k], c);
        }
    }
    float ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    if (!(ext > 0.f)) ext = 1.f;
    const float scale = 1023.0f / ext;
    for (int t = 0; t < nT; ++t) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) {
            const float f = (cen[3 * (size_t)t + k] - lo[k]) * scale;
            int qi = (int)f;                       // truncation, f >= 0
            q[k] = (uint32_t)std::min(std::max(qi, 0), 1023);
        }
        keys[t] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
    }
}

void rest_shape(const float* X, const uint32_t* Tet, int nT, float* DmInv, float* V0)
{
    for (int t = 0; t < nT; ++t) {
        const uint32_t* tv = Tet + 4 * (size_t)t;
        const float *x0 = X + 3 * (size_t)tv[0], *x1 = X + 3 * (size_t)tv[1], *x2 = X + 3 * (size_t)tv[2], *x3 = X + 3 * (size_t)tv[3];
        float m[3][3];   // m[c][r], column c = x_{c+1} - x_0  (glm layout)
        for (int r = 0; r < 3; ++r) { m[0][r] = x1[r] - x0[r]; m[1][r] = x2[r] - x0[r]; m[2][r] = x3[r] - x0[r]; }
        // glm::inverse / glm::determinant with the fused operations nvcc gives the reference's
        // computeInvDmV0 on sm_100a: minors as fma(a, b, -(c*d)), det = fma(m20, C2, fma(m00, C0, -(m10*C1)))
        auto minor = [](float a, float b, float c, float d) { return std::fma(a, b, -(c * d)); };
        const float c0 = minor(m[1][1], m[2][2], m[2][1], m[1][2]);
        const float c1 = minor(m[0][1], m[2][2], m[2][1], m[0][2]);
        const float c2 = minor(m[0][1], m[1][2], m[1][1], m[0][2]);
        const float det = std::fma(m[2][0], c2, std::fma(m[0][0], c0, -(m[1][0] * c1)));
        const float ood = 1.0f / det;
        float inv[3][3];
        inv[0][0] = +c0 * ood;
        inv[1][0] = -(minor(m[1][0], m[2][2], m[2][0], m[1][2]) * ood);
        inv[2][0] = +minor(m[1][0], m[2][1], m[2][0], m[1][1]) * ood;
        inv[0][1] = -(c1 * ood);
        inv[1][1] = +minor(m[0][0], m[2][2], m[2][0], m[0][2]) * ood;
        inv[2][1] = -(minor(m[0][0], m[2][1], m[2][0], m[0][1]) * ood);
        inv[0][2] = +c2 * ood;
        inv[1][2] = -(minor(m[0][0], m[1][2], m[1][0], m[0][2]) * ood);
        inv[2][2] = +minor(m[0][0], m[1][1], m[1][0], m[0][1]) * ood;
        float* B = DmInv + 9 * (size_t)t;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) B[r * 3 + c] = inv[c][r];
        V0[t] = std::fabs(det) / 6.0f;
    }
}

static inline size_t rup(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Proper edge colouring with 8 colours (Koenig: a bipartite multigraph of maximum degree 8 has one).
// Edge (tl, k) joins store group k*32 + tl/8 (the 8 tets whose STS.128 of corner k issue together) and the
// load group of its incidence entry, (row, half, lane/8) (the 8 lanes whose LDS.128 issue together); the
// colour is the 16-byte column inside the 128-byte H-scratch line of the store group, so that both the
// stores of phase B and the gathered loads of phase C are free of shared-memory bank conflicts.
// Deterministic: edges in (tet, corner) order, smallest free colours, alternating-path recolouring.
static void color_tile(uint32_t nTets, uint32_t nRows, const uint32_t* epos, uint8_t (*col)[4])
{
    // NC = 8 colours; store group of edge (tl, k) and load group of its incidence entry at position p = (row * 32 + lane) * 2 + half:
    //   8 consecutive tets x corner (one quarter-warp STS.128) / (row, half, lane / 8)
    constexpr uint32_t NC = TILE_HCOLOURS, NL = 4u * (uint32_t)TILE_T / NC, LG = 32u / NC;      // left nodes; load groups per (row, half)
    const uint32_t nE = 4u * nTets, nR = nRows * 2u * LG;
    std::vector<int16_t> atL((size_t)NL * NC, -1), atR((size_t)nR * NC, -1);
    std::vector<uint16_t> eu(nE), ev(nE);
    std::vector<int8_t> ec(nE, -1);
    std::vector<uint16_t> path;
    for (uint32_t e = 0; e < nE; ++e) {
        const uint32_t tl = e / 4u, k = e % 4u, p = epos[e];
        const uint32_t u = k * ((uint32_t)TILE_T / NC) + tl / NC, v = ((p / 64u) * 2u + (p & 1u)) * LG + ((p / 2u) % 32u) / NC;
        eu[e] = (uint16_t)u; ev[e] = (uint16_t)v;
        uint32_t a = 0, b = 0;
        while (a < NC && atL[(size_t)u * NC + a] >= 0) ++a;
        while (b < NC && atR[(size_t)v * NC + b] >= 0) ++b;
        if (a >= NC || b >= NC) throw std::runtime_error("tile colouring: degree above the colour count");
        if (atR[(size_t)v * NC + a] >= 0) {
            // free colour a at v: swap a <-> b along the alternating path that leaves v by colour a
            path.clear();
            uint32_t node = v; bool right = true; uint32_t c = a;
            for (;;) {
                const int16_t f = right ? atR[(size_t)node * NC + c] : atL[(size_t)node * NC + c];
                if (f < 0) break;
                path.push_back((uint16_t)f);
                node = right ? eu[f] : ev[f];
                right = !right;
                c = (c == a) ? b : a;
            }
            for (uint16_t f : path) { atL[(size_t)eu[f] * NC + ec[f]] = -1; atR[(size_t)ev[f] * NC + ec[f]] = -1; }
            for (uint16_t f : path) {
                ec[f] = (int8_t)((uint32_t)ec[f] == a ? b : a);
                atL[(size_t)eu[f] * NC + ec[f]] = (int16_t)f; atR[(size_t)ev[f] * NC + ec[f]] = (int16_t)f;
            }
        }
        ec[e] = (int8_t)a;
        atL[(size_t)u * NC + a] = (int16_t)e; atR[(size_t)v * NC + a] = (int16_t)e;
    }
    for (uint32_t e = 0; e < nE; ++e) col[e / 4u][e % 4u] = (uint8_t)ec[e];
}

// Staging slot of every tile-local vertex (the position gather of the local kernel writes vertex l to slot sigma[l];
// phase B reads it back with one LDS.128 per corner).  A quarter-warp (8 consecutive tets) reading corner k is conflict
// free when its distinct vertices sit in 8 different 16-byte bank groups = slots that differ mod 8.  Greedy colouring
// of the vertices (most constrained first) with the bank group as colour, then two improvement sweeps; at most 32
// vertices per colour (256 slots).  Integer work only, deterministic.  PD_NO_STAGE_COLOR=1 keeps sigma = identity.
static void stage_slots(uint32_t nTets, uint32_t nLocal, const uint32_t (*cl)[4], std::vector<uint8_t>& sigma)
{
    sigma.resize(nLocal);
    static const bool off = [] { const char* e = std::getenv("PD_NO_STAGE_COLOR"); return e && e[0] == '1'; }();
    if (off) { for (uint32_t l = 0; l < nLocal; ++l) sigma[l] = (uint8_t)l; return; }
    const uint32_t nQ = (nTets + 7u) / 8u, nSets = nQ * 4u;
    // sets = (quarter-warp, corner): distinct vertices; per vertex the sets it is in
    std::vector<std::vector<uint16_t>> setsOf(nLocal);
    for (uint32_t q = 0; q < nQ; ++q)
        for (uint32_t k = 0; k < 4; ++k) {
            uint32_t seen[8]; int ns = 0;
            for (uint32_t t = 8 * q; t < std::min(nTets, 8 * q + 8); ++t) {
                const uint32_t l = cl[t][k];
                bool dup = false;
                for (int i = 0; i < ns; ++i) dup |= seen[i] == l;
                if (!dup) { seen[ns++] = l; setsOf[l].push_back((uint16_t)(q * 4u + k)); }
            }
        }
    std::vector<uint8_t> cnt((size_t)nSets * 8, 0);      // vertices of colour c in set s
    std::vector<int> colour(nLocal, -1);
    uint32_t used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> order(nLocal);
    for (uint32_t l = 0; l < nLocal; ++l) order[l] = l;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return setsOf[a].size() > setsOf[b].size(); });
    auto best_colour = [&](uint32_t l) {
        int best = -1; uint32_t bestCost = 0xffffffffu;
        for (int c = 0; c < 8; ++c) {
            if (used[c] >= 32u) continue;
            uint32_t cost = 0;
            for (uint16_t s_ : setsOf[l]) cost += cnt[(size_t)s_ * 8 + c];
            cost = cost * 64u + used[c];                  // ties: the emptiest colour
            if (cost < bestCost) { bestCost = cost; best = c; }
        }
        return best;
    };
    for (uint32_t l : order) {
        const int c = best_colour(l);
        colour[l] = c; used[c]++;
        for (uint16_t s_ : setsOf[l]) cnt[(size_t)s_ * 8 + c]++;
    }
    for (int sweep = 0; sweep < 2; ++sweep)
        for (uint32_t l : order) {
            int c = colour[l];
            used[c]--;
            for (uint16_t s_ : setsOf[l]) cnt[(size_t)s_ * 8 + c]--;
            c = best_colour(l);
            colour[l] = c; used[c]++;
            for (uint16_t s_ : setsOf[l]) cnt[(size_t)s_ * 8 + c]++;
        }
    uint32_t next[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t l = 0; l < nLocal; ++l) { const int c = colour[l]; sigma[l] = (uint8_t)(c + 8 * (int)next[c]++); }
}

// tiles + records + slots for a mesh whose vertex ids are final
static void build_tiles(int nV, int nT, const uint32_t* tet, const float* DmInv, const float* w, Layout& L)
{
    L.tileTetStart.clear(); L.tileRecOff.clear(); L.records.clear(); L.tileTab.clear(); L.vlist.clear();
    std::vector<int> mark((size_t)nV, -1);
    std::vector<uint32_t> lidx((size_t)nV, 0);
    std::vector<uint32_t> vcount((size_t)nV + 1, 0);
    L.maxLocal = 0;
    uint32_t slot = 0;
    int t0 = 0, tile = 0;
    std::vector<uint32_t> vl;
    L.tileTetStart.push_back(0);
    L.tileRecOff.push_back(0);
    std::vector<uint32_t> cnt;
    std::vector<uint8_t> tileSigma;      // staging slot of every (tile, tile-local vertex)
    while (t0 < nT) {
        // greedy tile: up to TILE_T tets and TILE_NLMAX distinct vertices ...
        vl.clear();
        int t1 = t0;
        while (t1 < nT && t1 - t0 < TILE_T) {
            int added = 0;
            for (int k = 0; k < 4; ++k) {
                const uint32_t v = tet[4 * (size_t)t1 + k];
                if (mark[v] != tile) { mark[v] = tile; vl.push_back(v); ++added; }
            }
            if ((int)vl.size() > TILE_NLMAX) {       // undo this tet and close the tile
                for (int a = 0; a < added; ++a) { mark[vl.back()] = -1; vl.pop_back(); }
                break;
            }
            ++t1;
        }
        if (t1 == t0) throw std::runtime_error("tile builder: empty tile");
        // ... and at most TILE_ROWSMAX incidence rows: halve the tile until it fits
        uint32_t nLocal = 0, nTets = 0, nGroups = 0, nRows = 0;
        uint32_t gRowBase[TILE_NGROUPS], gRows[TILE_NGROUPS];
        for (;;) {
            nTets = (uint32_t)(t1 - t0);
            vl.clear();
            for (int t = t0; t < t1; ++t) for (int k = 0; k < 4; ++k) {
                const uint32_t v = tet[4 * (size_t)t + k];
                if (mark[v] != -2 - tile) { mark[v] = -2 - tile; vl.push_back(v); lidx[v] = 0; }
                lidx[v]++;
            }
            for (uint32_t v : vl) mark[v] = -1;
            // order: in-tile incidence count descending, id ascending
            std::sort(vl.begin(), vl.end(), [&](uint32_t a, uint32_t b) { return lidx[a] != lidx[b] ? lidx[a] > lidx[b] : a < b; });
            nLocal = (uint32_t)vl.size();
            nGroups = (nLocal + (uint32_t)TILE_GROUP - 1u) / (uint32_t)TILE_GROUP;
            nRows = 0;
            for (uint32_t g = 0; g < nGroups; ++g) {
                gRowBase[g] = nRows;
                gRows[g] = (lidx[vl[TILE_GROUP * g]] + 2u * TILE_LPV - 1u) / (2u * TILE_LPV);        // the group's first vertex has its largest count
                nRows += gRows[g];
            }
            if (nRows <= (uint32_t)TILE_ROWSMAX || nTets == 1) break;
            t1 = t0 + (int)(nTets / 2);
        }
        if (nRows > (uint32_t)TILE_ROWSMAX) throw std::runtime_error("tile builder: a single tet exceeds the incidence row budget");
        cnt.assign(nLocal, 0);
        for (uint32_t l = 0; l < nLocal; ++l) { cnt[l] = lidx[vl[l]]; }
        for (uint32_t l = 0; l < nLocal; ++l) { lidx[vl[l]] = l; vcount[vl[l] + 1]++; }
        L.maxLocal = std::max(L.maxLocal, (int)nLocal);
        const size_t abBytes = tile_ab_bytes(nTets), cBytes = 128 * (size_t)nRows;
        const size_t base = L.records.size();
        L.records.resize(base + abBytes + cBytes, 0);
        uint8_t* rec = L.records.data() + base;
        if ((uint64_t)(tile + 1) * TILE_NLMAX > 0xffffffffull) throw std::runtime_error("tile builder: slot index overflow");
        const uint32_t slotBase = (uint32_t)tile * (uint32_t)TILE_NLMAX;
        TileHeader h{nTets, nLocal, slotBase, (uint32_t)abBytes, (uint32_t)cBytes, nGroups, (uint32_t)(base & 0xffffffffu), (uint32_t)((uint64_t)base >> 32)};
        std::memcpy(rec, &h, 32);
        uint32_t* gtab = reinterpret_cast<uint32_t*>(rec + 32);
        for (uint32_t g = 0; g < nGroups; ++g) gtab[g] = gRowBase[g] | (gRows[g] << 16);
        L.vlist.insert(L.vlist.end(), vl.begin(), vl.end());          // owner bits are set once all tiles are known
        L.vlist.resize((size_t)(tile + 1) * TILE_NLMAX, 0xffffffffu);
        uint16_t* incT = reinterpret_cast<uint16_t*>(rec + abBytes);
        // incidence entry of (tet tl, corner k) -> position (row, lane, half) in the transposed rows
        std::vector<uint32_t> fill(nLocal, 0);
        std::vector<uint32_t> epos((size_t)nTets * 4);      // index into incT
        uint32_t cl[TILE_T][4];
        for (uint32_t tl = 0; tl < nTets; ++tl)
            for (int k = 0; k < 4; ++k) {
                const uint32_t l = lidx[tet[4 * ((size_t)t0 + tl) + k]];
                cl[tl][k] = l;
                const uint32_t e = fill[l]++, g = l / (uint32_t)TILE_GROUP, lane = l % (uint32_t)TILE_GROUP + (uint32_t)TILE_GROUP * ((e >> 1) % (uint32_t)TILE_LPV);
                epos[4 * tl + k] = ((gRowBase[g] + e / (2u * TILE_LPV)) * 32u + lane) * 2u + (e & 1u);
            }
        std::vector<uint8_t> sigma;
        stage_slots(nTets, nLocal, cl, sigma);
        // H-scratch column of every (tet, corner): proper 8-colouring of the bipartite multigraph
        // store groups {8 consecutive tets, corner k}  x  load groups {row, half, 8 consecutive lanes}
        uint8_t col[TILE_T][4];
        color_tile(nTets, nRows, epos.data(), col);
        // pads point at the zero slot of a column no real entry of their load group uses
        {
            constexpr uint32_t NC = TILE_HCOLOURS, LG = 32u / NC, ESZ = 16u;      // load groups per (row, half); entry bytes
            const size_t nLg = (size_t)nRows * 2 * LG;     // load group = (row, half, lane / NC)
            std::vector<uint32_t> used(nLg, 0);
            for (uint32_t tl = 0; tl < nTets; ++tl)
                for (int k = 0; k < 4; ++k) {
                    const uint32_t p = epos[4 * tl + k], row = p / 64u, lane = (p / 2u) % 32u, half = p & 1u;
                    used[(row * 2u + half) * LG + lane / NC] |= 1u << col[tl][k];
                }
            for (uint32_t row = 0; row < nRows; ++row)
                for (uint32_t lane = 0; lane < 32; ++lane)
                    for (uint32_t half = 0; half < 2; ++half) {
                        const uint32_t u = used[(row * 2u + half) * LG + lane / NC];
                        uint32_t c = 0;
                        while (c < NC - 1u && (u >> c & 1u)) ++c;    // a full group (every colour used) has no pad
                        incT[(row * 32u + lane) * 2u + half] = (uint16_t)(TILE_ZERO_OFF + ESZ * c);
                    }
        }
        for (uint32_t tl = 0; tl < nTets; ++tl) {
            const size_t t = (size_t)t0 + tl;
            float tr[12];
            for (int e = 0; e < 9; ++e) tr[e] = DmInv[9 * t + e];
            tr[9] = w[t];
            uint32_t h[4];
            for (int k = 0; k < 4; ++k) {
                h[k] = tile_corner_half(sigma[cl[tl][k]], col[tl][k]);
                incT[epos[4 * tl + k]] = (uint16_t)tile_h_offset(tl, (uint32_t)k, col[tl][k]);
            }
            const uint32_t c01 = h[0] | (h[1] << 16), c23 = h[2] | (h[3] << 16);
            std::memcpy(&tr[10], &c01, 4);
            std::memcpy(&tr[11], &c23, 4);
            for (uint32_t j = 0; j < 12; ++j) std::memcpy(rec + tile_tet_word(nTets, tl, j), &tr[j], 4);      // three 16-byte planes
        }
        L.tileTab.push_back(TileEntry{(uint64_t)base, (uint32_t)abBytes, (uint32_t)cBytes});
        tileSigma.insert(tileSigma.end(), sigma.begin(), sigma.end());
        tileSigma.resize((size_t)(tile + 1) * TILE_NLMAX, 0);
        slot += nLocal;
        t0 = t1;
        ++tile;
        L.tileTetStart.push_back((uint32_t)t0);
        L.tileRecOff.push_back((uint64_t)L.records.size());
    }
    L.nTiles = tile;
    L.nSlots = slot;
    // vertex -> slots CSR, ascending (tiles are visited in ascending order)
    for (int v = 0; v < nV; ++v) vcount[v + 1] += vcount[v];
    L.vslotPtr.assign(vcount.begin(), vcount.end());
    L.vslot.assign((size_t)slot, 0u);
    std::vector<uint32_t> fillv(vcount.begin(), vcount.end() - 1);
    for (size_t s = 0; s < L.vlist.size(); ++s) {
        const uint32_t v = L.vlist[s];
        if (v == 0xffffffffu) continue;
        if (fillv[v] == L.vslotPtr[v]) L.vlist[s] = v | TILE_OWNER_BIT;       // first slot of the vertex = owner
        L.vslot[fillv[v]++] = (uint32_t)s;
    }
    L.vstage.assign(L.vlist.size(), 0xffffffffu);
    for (size_t s = 0; s < L.vlist.size(); ++s)
        if (L.vlist[s] != 0xffffffffu) L.vstage[(s / TILE_NLMAX) * TILE_NLMAX + tileSigma[s]] = L.vlist[s];
}

// The device tile table the local kernel reads (TILE_META_WORDS words per tile, layout.hpp).
void build_tile_table(const Layout& L, std::vector<uint32_t>& meta)
{
    meta.assign(L.tileTab.size() * TILE_META_WORDS, 0u);
    for (size_t ti = 0; ti < L.tileTab.size(); ++ti) {
        const uint8_t* rec = L.records.data() + L.tileRecOff[ti];
        TileHeader h; std::memcpy(&h, rec, sizeof(h));
        uint32_t* m = &meta[ti * TILE_META_WORDS];
        if ((L.tileTab[ti].off & 15u) || (L.tileTab[ti].off >> 4) > 0xffffffffull || L.tileTab[ti].abBytes > 0xffffu || L.tileTab[ti].cBytes > 0xffffu)
            throw std::runtime_error("tile table: record offset / size out of range");
        m[0] = (uint32_t)(L.tileTab[ti].off >> 4);
        m[1] = L.tileTab[ti].abBytes | (L.tileTab[ti].cBytes << 16);
        m[2] = h.nTets | (h.nLocal << 16);
        for (uint32_t g = 0; g < (uint32_t)TILE_NGROUPS; ++g) {
            uint32_t gt = 0;
            if (g < h.nGroups) std::memcpy(&gt, rec + 32 + 4 * g, 4);          // rowBase | nRows << 16
            const uint32_t nValid = (g < h.nGroups) ? std::min<uint32_t>((uint32_t)TILE_GROUP, h.nLocal - g * (uint32_t)TILE_GROUP) : 0u;
            m[4 + g] = (gt & 63u) | ((gt >> 16) << 6) | (nValid << 12) | (g < 8u ? (h.nTets << 18) : 0u);
        }
    }
}

void build_layout(int nV, int nT, const float* X, const uint32_t* Tet, const float* mu, bool reorder, Layout& L)
{
    L = Layout();
    L.nV = nV; L.nT = nT;
    // 1. tet order: stable sort by Morton key of the centroid
    L.tetOrder.resize((size_t)nT);
    std::iota(L.tetOrder.begin(), L.tetOrder.end(), 0u);
    if (reorder && nT > 1) {
        std::vector<uint32_t> keys;
        morton_keys(X, Tet, nT, keys);
        std::stable_sort(L.tetOrder.begin(), L.tetOrder.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    }
    // 2. vertex renumbering: first touch over the reordered tets; untouched vertices last
    L.vertNewOfOld.assign((size_t)nV, 0xffffffffu);
    L.vertOrder.clear(); L.vertOrder.reserve((size_t)nV);
    for (int t = 0; t < nT; ++t)
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = Tet[4 * (size_t)L.tetOrder[t] + k];
            if (v >= (uint32_t)nV) throw std::runtime_error("tet references vertex out of range");
            if (L.vertNewOfOld[v] == 0xffffffffu) { L.vertNewOfOld[v] = (uint32_t)L.vertOrder.size(); L.vertOrder.push_back(v); }
        }
    for (int v = 0; v < nV; ++v)
        if (L.vertNewOfOld[v] == 0xffffffffu) { L.vertNewOfOld[v] = (uint32_t)L.vertOrder.size(); L.vertOrder.push_back((uint32_t)v); }
    L.tetNew.resize((size_t)nT * 4);
    for (int t = 0; t < nT; ++t)
        for (int k = 0; k < 4; ++k) L.tetNew[4 * (size_t)t + k] = L.vertNewOfOld[Tet[4 * (size_t)L.tetOrder[t] + k]];
    // 3. rest shape in the ORIGINAL numbering/order (same arithmetic as the reference), then permuted
    std::vector<float> B((size_t)nT * 9), V0((size_t)nT), Br((size_t)nT * 9), wr((size_t)nT);
    rest_shape(X, Tet, nT, B.data(), V0.data());
    for (int t = 0; t < nT; ++t) {
        const size_t o = L.tetOrder[t];
        std::memcpy(&Br[9 * (size_t)t], &B[9 * o], 36);
        wr[t] = std::fabs(V0[o]) * mu[o];
    }
    build_tiles(nV, nT, L.tetNew.data(), Br.data(), wr.data(), L);
}

// matrix_diag[v] = sum over the tets incident to v of w_t |col_i(B^T G)|^2 (PdUtil::computeSiTSi, pdUtil.cu:16-24), read back
// from the tile records, in ascending GLOBAL tet order: a rank keeps its interior tiles first and may have re-packed its
// boundary tiles (Layout::tetGlobal), and the float sums must not depend on the world size.
void matrix_diag_host(const Layout& L, std::vector<float>& md)
{
    md.assign((size_t)L.nV, 0.f);
    struct TetRef { uint32_t global; uint32_t tile, tl; };
    std::vector<TetRef> order; order.reserve((size_t)L.nT);
    size_t t = 0;
    for (int ti = 0; ti < L.nTiles; ++ti) {
        const uint32_t n = L.tileTetStart[(size_t)ti + 1] - L.tileTetStart[(size_t)ti];
        for (uint32_t tl = 0; tl < n; ++tl, ++t) order.push_back(TetRef{L.tetGlobal.empty() ? (uint32_t)t : L.tetGlobal[t], (uint32_t)ti, tl});
    }
    if (!L.tetGlobal.empty()) std::sort(order.begin(), order.end(), [](const TetRef& a, const TetRef& b) { return a.global < b.global; });
    int curTile = -1;
    const uint8_t* rec = nullptr; TileHeader h{}; const uint32_t* vstage = nullptr;
    for (const TetRef& r : order) {
        if ((int)r.tile != curTile) {
            curTile = (int)r.tile;
            rec = L.records.data() + L.tileRecOff[r.tile];
            std::memcpy(&h, rec, sizeof(h));
            vstage = L.vstage.data() + h.slotBase;      // the corner words hold staging slots
        }
        float B[12];
        for (uint32_t j = 0; j < 12; ++j) std::memcpy(&B[j], rec + tile_tet_word(h.nTets, r.tl, j), 4);
        const float w = B[9];
        uint32_t cw[2]; std::memcpy(cw, B + 10, 8);
        const uint32_t loc[4] = {(cw[0] >> 4) & 0xffu, (cw[0] >> 20) & 0xffu, (cw[1] >> 4) & 0xffu, (cw[1] >> 20) & 0xffu};
        for (int i = 0; i < 4; ++i) {
            float col[3];
            for (int c = 0; c < 3; ++c)
                col[c] = (i == 0) ? ((-B[0 * 3 + c] - B[1 * 3 + c]) - B[2 * 3 + c]) : B[(i - 1) * 3 + c];
            // computeSiTSi as nvcc fuses it: fma(c2,c2, fma(c0,c0, c1*c1)), then * (V0*mu)
            const float kii = std::fma(col[2], col[2], std::fma(col[0], col[0], col[1] * col[1]));
            md[vstage[loc[i]] & ~TILE_OWNER_BIT] += kii * w;
        }
    }
}

void build_system_matrix(const Layout& L, const float*, const float* DmInv, const float* w, const float* c,
                         CsrMatrix& A, std::vector<float>& matrixDiag)
{
    const int nV = L.nV, nT = L.nT;
    // incidence (ascending reordered tet order)
    std::vector<int> ptr((size_t)nV + 1, 0);
    for (size_t i = 0; i < (size_t)nT * 4; ++i) ptr[L.tetNew[i] + 1]++;
    for (int v = 0; v < nV; ++v) ptr[v + 1] += ptr[v];
    std::vector<int> inc((size_t)nT * 4), fill(ptr.begin(), ptr.end() - 1);
    for (int t = 0; t < nT; ++t) for (int k = 0; k < 4; ++k) inc[fill[L.tetNew[4 * (size_t)t + k]]++] = 4 * t + k;
    matrixDiag.assign((size_t)nV, 0.f);
    A.n = nV; A.rowPtr.assign((size_t)nV + 1, 0); A.col.clear(); A.val.clear();
    struct Ent { int c; float v; int seq; };
    std::vector<Ent> e;
    for (int v = 0; v < nV; ++v) {
        e.clear();
        int seq = 0;
        float md = 0.f;
        for (int q = ptr[v]; q < ptr[v + 1]; ++q) {
            const int t = inc[q] >> 2, i = inc[q] & 3;
            const float* B = DmInv + 9 * (size_t)t;
            float cols[4][3];
            for (int r = 0; r < 3; ++r) {
                const float m0 = B[0 * 3 + r], m1 = B[1 * 3 + r], m2 = B[2 * 3 + r];
                cols[0][r] = m0 * -1.0f + m1 * -1.0f + m2 * -1.0f;
                cols[1][r] = m0; cols[2][r] = m1; cols[3][r] = m2;
            }
            for (int j = 0; j < 4; ++j) {
                const float kji = cols[i][0] * cols[j][0] + cols[i][1] * cols[j][1] + cols[i][2] * cols[j][2];
                e.push_back({(int)L.tetNew[4 * (size_t)t + j], kji * w[t], seq++});
                if (j == i) md += kji * w[t];
            }
        }
        matrixDiag[v] = md;
        e.push_back({v, c[v], seq++});
        std::sort(e.begin(), e.end(), [](const Ent& a, const Ent& b) { return a.c != b.c ? a.c < b.c : a.seq < b.seq; });
        A.rowPtr[v] = (int)A.col.size();
        for (size_t a = 0; a < e.size();) {
            const int cc = e[a].c; float acc = 0.f;
            while (a < e.size() && e[a].c == cc) { acc += e[a].v; ++a; }
            A.col.push_back(cc); A.val.push_back(acc);
        }
    }
    A.rowPtr[nV] = (int)A.col.size();
}

void nested_dissection_order(const CsrMatrix& A, const float* xyz, std::vector<int>& perm)
{
    const int n = A.n;
    perm.clear(); perm.reserve((size_t)n);
    if (!xyz) { for (int i = 0; i < n; ++i) perm.push_back(i); return; }
    // side[v]: id of the sub-domain v currently belongs to (-1 once numbered / moved to a separator)
    std::vector<int> side((size_t)n, 0), work((size_t)n);
    for (int i = 0; i < n; ++i) work[(size_t)i] = i;
    struct Job { int lo, hi, id; bool emit; };          // work[lo, hi) is a sub-domain, or (emit) a separator to be numbered now
    std::vector<Job> stack{{0, n, 0, false}};
    int nextId = 1;
    std::vector<int> sepL, sepR;
    while (!stack.empty()) {
        const Job j = stack.back(); stack.pop_back();
        const int m = j.hi - j.lo;
        if (j.emit || m <= 48) {
            if (!j.emit) std::sort(work.begin() + j.lo, work.begin() + j.hi);      // leaves keep the matrix order
            for (int k = j.lo; k < j.hi; ++k) { perm.push_back(work[(size_t)k]); side[(size_t)work[(size_t)k]] = -1; }
            continue;
        }
        float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
        for (int k = j.lo; k < j.hi; ++k)
            for (int c = 0; c < 3; ++c) { const float x = xyz[3 * (size_t)work[(size_t)k] + c]; mn[c] = std::min(mn[c], x); mx[c] = std::max(mx[c], x); }
        int ax = 0;
        for (int c = 1; c < 3; ++c) if (mx[c] - mn[c] > mx[ax] - mn[ax]) ax = c;
        const int mid = j.lo + m / 2;
        std::nth_element(work.begin() + j.lo, work.begin() + mid, work.begin() + j.hi, [&](int a, int b) {
            const float xa = xyz[3 * (size_t)a + ax], xb = xyz[3 * (size_t)b + ax];
            return xa < xb || (xa == xb && a < b);
        });
        const int idL = nextId++, idR = nextId++;
        for (int k = j.lo; k < mid; ++k) side[(size_t)work[(size_t)k]] = idL;
        for (int k = mid; k < j.hi; ++k) side[(size_t)work[(size_t)k]] = idR;
        // the two boundary layers; the smaller one is the separator
        sepL.clear(); sepR.clear();
        for (int k = j.lo; k < j.hi; ++k) {
            const int v = work[(size_t)k], other = (k < mid) ? idR : idL;
            for (int e = A.rowPtr[v]; e < A.rowPtr[v + 1]; ++e)
                if (side[(size_t)A.col[e]] == other) { (k < mid ? sepL : sepR).push_back(v); break; }
        }
        const bool takeL = sepL.size() <= sepR.size();
        const std::vector<int>& sep = takeL ? sepL : sepR;
        if (sep.empty()) {          // disconnected halves: no separator at all
            stack.push_back({mid, j.hi, idR, false}); stack.push_back({j.lo, mid, idL, false});
            continue;
        }
        const int idS = nextId++;
        for (int v : sep) side[(size_t)v] = idS;
        // regroup work[lo, hi) as [left | right | separator]
        std::vector<int> tmp(work.begin() + j.lo, work.begin() + j.hi);
        int o = j.lo;
        for (int v : tmp) if (side[(size_t)v] == idL) work[(size_t)o++] = v;
        const int endL = o;
        for (int v : tmp) if (side[(size_t)v] == idR) work[(size_t)o++] = v;
        const int endR = o;
        for (int v : tmp) if (side[(size_t)v] == idS) work[(size_t)o++] = v;
        // numbered in the order left, right, separator: the stack pops in reverse
        stack.push_back({endR, j.hi, idS, true});
        stack.push_back({endL, endR, idR, false});
        stack.push_back({j.lo, endL, idL, false});
    }
    if ((int)perm.size() != n) throw std::runtime_error("nested dissection: internal error (not a permutation)");
}

void permute_symmetric(const CsrMatrix& A, const std::vector<int>& perm, CsrMatrix& B)
{
    const int n = A.n;
    std::vector<int> inv((size_t)n);
    for (int i = 0; i < n; ++i) inv[(size_t)perm[(size_t)i]] = i;
    B = CsrMatrix(); B.n = n;
    B.rowPtr.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) B.rowPtr[(size_t)i + 1] = B.rowPtr[(size_t)i] + (A.rowPtr[perm[(size_t)i] + 1] - A.rowPtr[perm[(size_t)i]]);
    B.col.resize(A.col.size()); B.val.resize(A.val.size());
    std::vector<std::pair<int, float>> row;
    for (int i = 0; i < n; ++i) {
        const int o = perm[(size_t)i];
        row.clear();
        for (int e = A.rowPtr[o]; e < A.rowPtr[o + 1]; ++e) row.emplace_back(inv[(size_t)A.col[e]], A.val[e]);
        std::sort(row.begin(), row.end(), [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.first < b.first; });
        int p = B.rowPtr[(size_t)i];
        for (const auto& cv : row) { B.col[(size_t)p] = cv.first; B.val[(size_t)p] = cv.second; ++p; }
    }
}

void cholesky_factor(const CsrMatrix& A, CholFactor& F)
{
    const int n = A.n;
    F = CholFactor();
    F.n = n;
    // elimination tree (Liu): A is symmetric, row k's entries with column < k are column k's upper part
    std::vector<int> parent((size_t)n, -1), anc((size_t)n, -1);
    for (int k = 0; k < n; ++k)
        for (int e = A.rowPtr[k]; e < A.rowPtr[k + 1]; ++e) {
            int i = A.col[e];
            while (i != -1 && i < k) {
                const int nx = anc[(size_t)i];
                anc[(size_t)i] = k;
                if (nx == -1) parent[(size_t)i] = k;
                i = nx;
            }
        }
    // columns of L grow row by row (entries arrive in ascending row order)
    std::vector<std::vector<int>> ci((size_t)n);
    std::vector<std::vector<double>> cx((size_t)n);
    std::vector<double> diag((size_t)n, 0.0), x((size_t)n, 0.0);
    std::vector<int> mark((size_t)n, -1), stack((size_t)n), path((size_t)n);
    for (int k = 0; k < n; ++k) {
        // pattern of row k of L = nodes reached from the entries of A(k, 0:k) up the elimination tree, topological order
        int top = n;
        mark[(size_t)k] = k;
        double d = 0.0;
        for (int e = A.rowPtr[k]; e < A.rowPtr[k + 1]; ++e) {
            const int c = A.col[e];
            if (c > k) continue;
            if (c == k) { d = (double)A.val[e]; continue; }
            x[(size_t)c] = (double)A.val[e];
            int len = 0;
            for (int i = c; mark[(size_t)i] != k; i = parent[(size_t)i]) { path[(size_t)len++] = i; mark[(size_t)i] = k; }
            while (len > 0) stack[(size_t)--top] = path[(size_t)--len];
        }
        for (int t = top; t < n; ++t) {
            const int j = stack[(size_t)t];
            const double lkj = x[(size_t)j] / diag[(size_t)j];
            x[(size_t)j] = 0.0;
            const std::vector<int>& rj = ci[(size_t)j];
            const std::vector<double>& vj = cx[(size_t)j];
            for (size_t q = 0; q < rj.size(); ++q) x[(size_t)rj[q]] -= vj[q] * lkj;
            d -= lkj * lkj;
            ci[(size_t)j].push_back(k);
            cx[(size_t)j].push_back(lkj);
        }
        if (!(d > 0.0)) throw std::runtime_error("Cholesky: system matrix is not positive definite at row " + std::to_string(k));
        diag[(size_t)k] = std::sqrt(d);
    }
    // L^T by rows = L by columns: diagonal first, then ascending rows
    F.uPtr.assign((size_t)n + 1, 0);
    for (int j = 0; j < n; ++j) F.uPtr[(size_t)j + 1] = F.uPtr[(size_t)j] + 1 + (int)ci[(size_t)j].size();
    F.uCol.resize((size_t)F.uPtr[(size_t)n]); F.uVal.resize((size_t)F.uPtr[(size_t)n]);
    std::vector<int> rowCount((size_t)n, 1);
    for (int j = 0; j < n; ++j) {
        int o = F.uPtr[(size_t)j];
        F.uCol[(size_t)o] = j; F.uVal[(size_t)o] = (float)diag[(size_t)j]; ++o;
        for (size_t q = 0; q < ci[(size_t)j].size(); ++q, ++o) {
            F.uCol[(size_t)o] = ci[(size_t)j][q]; F.uVal[(size_t)o] = (float)cx[(size_t)j][q];
            rowCount[(size_t)ci[(size_t)j][q]]++;
        }
    }
    // L by rows: ascending columns (columns are visited in ascending order), diagonal last
    F.lPtr.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) F.lPtr[(size_t)i + 1] = F.lPtr[(size_t)i] + rowCount[(size_t)i];
    F.lCol.resize((size_t)F.lPtr[(size_t)n]); F.lVal.resize((size_t)F.lPtr[(size_t)n]);
    std::vector<int> fill(F.lPtr.begin(), F.lPtr.end() - 1);
    for (int j = 0; j < n; ++j)
        for (size_t q = 0; q < ci[(size_t)j].size(); ++q) {
            const int i = ci[(size_t)j][q];
            F.lCol[(size_t)fill[(size_t)i]] = j; F.lVal[(size_t)fill[(size_t)i]] = (float)cx[(size_t)j][q];
            ++fill[(size_t)i];
        }
    for (int i = 0; i < n; ++i) { F.lCol[(size_t)fill[(size_t)i]] = i; F.lVal[(size_t)fill[(size_t)i]] = (float)diag[(size_t)i]; }
}

bool connected_body_ranges(int nV, int nT, const uint32_t* Tet, std::vector<int>& starts)
{
    starts.clear();
    std::vector<int> parent((size_t)nV);
    for (int v = 0; v < nV; ++v) parent[(size_t)v] = v;
    auto find = [&](int v) { while (parent[(size_t)v] != v) { parent[(size_t)v] = parent[(size_t)parent[(size_t)v]]; v = parent[(size_t)v]; } return v; };
    for (int t = 0; t < nT; ++t) {
        const int a = find((int)Tet[4 * (size_t)t]);
        for (int k = 1; k < 4; ++k) {
            const int b = find((int)Tet[4 * (size_t)t + k]);
            if (b != a) parent[(size_t)std::max(a, b)] = std::min(a, b);      // the root of a component is its smallest vertex
        }
    }
    // contiguous ranges <=> walking the vertices in order, the root changes only to the current vertex itself
    int cur = -1;
    for (int v = 0; v < nV; ++v) {
        const int r = find(v);
        if (r == v) { starts.push_back(v); cur = v; }
        else if (r != cur) { starts.clear(); return false; }
    }
    return !starts.empty();
}

void build_body_batch(int nV, int nT, const float* X, const uint32_t* Tet, const float* mu, const std::vector<int>& bodyVertStart,
                      const uint32_t* vertNewOfOld, BodyBatch& out)
{
    out = BodyBatch();
    const size_t nB = bodyVertStart.size();
    if (nB == 0) throw std::runtime_error("body batch: the scene names no bodies");
    std::vector<float> B((size_t)nT * 9), V0((size_t)nT);
    rest_shape(X, Tet, nT, B.data(), V0.data());
    int t = 0;
    for (size_t b = 0; b < nB; ++b) {
        const int vb = bodyVertStart[b], ve = (b + 1 < nB) ? bodyVertStart[b + 1] : nV;
        if (vb < 0 || ve <= vb || ve > nV) throw std::runtime_error("body batch: bad vertex range");
        const int tb = t;
        while (t < nT && (int)Tet[4 * (size_t)t] >= vb && (int)Tet[4 * (size_t)t] < ve) {
            for (int k = 1; k < 4; ++k)
                if ((int)Tet[4 * (size_t)t + k] < vb || (int)Tet[4 * (size_t)t + k] >= ve) throw std::runtime_error("body batch: a tet spans two bodies");
            ++t;
        }
        const uint32_t nVb = (uint32_t)(ve - vb), nTb = (uint32_t)(t - tb);
        if (nTb == 0) throw std::runtime_error("body batch: a body without tets");
        if (nVb > 65535u || nTb > 16383u) throw std::runtime_error("body batch: body too large for the per-body kernel");
        BodyDesc d{};
        d.v0 = (uint32_t)out.verts.size(); d.nV = nVb; d.ptr0 = (uint32_t)out.incPtr.size(); d.nT = nTb;
        d.inc0 = (uint32_t)out.inc.size(); d.recOff16 = (uint32_t)(out.rec.size() / 16);
        for (int v = vb; v < ve; ++v) out.verts.push_back(vertNewOfOld ? vertNewOfOld[v] : (uint32_t)v);
        // records: three planes of 16 bytes per tet
        const size_t base = out.rec.size();
        out.rec.resize(base + 48 * (size_t)nTb, 0);
        std::vector<uint32_t> cnt((size_t)nVb + 1, 0);
        std::vector<float> md((size_t)nVb, 0.f);
        for (uint32_t tl = 0; tl < nTb; ++tl) {
            const size_t tt = (size_t)tb + tl;
            float tr[12];
            for (int e = 0; e < 9; ++e) tr[e] = B[9 * tt + e];
            const float w = std::fabs(V0[tt]) * mu[tt];
            tr[9] = w;
            uint32_t loc[4];
            for (int k = 0; k < 4; ++k) { loc[k] = Tet[4 * tt + k] - (uint32_t)vb; cnt[loc[k] + 1]++; }
            const uint32_t c01 = loc[0] | (loc[1] << 16), c23 = loc[2] | (loc[3] << 16);
            std::memcpy(&tr[10], &c01, 4); std::memcpy(&tr[11], &c23, 4);
            for (uint32_t j = 0; j < 12; ++j) std::memcpy(out.rec.data() + base + 16 * ((size_t)(j / 4) * nTb + tl) + 4 * (j % 4), &tr[j], 4);
            for (int i = 0; i < 4; ++i) {       // computeSiTSi, as matrix_diag_host does it
                float col[3];
                for (int c = 0; c < 3; ++c) col[c] = (i == 0) ? ((-tr[0 * 3 + c] - tr[1 * 3 + c]) - tr[2 * 3 + c]) : tr[(i - 1) * 3 + c];
                const float kii = std::fma(col[2], col[2], std::fma(col[0], col[0], col[1] * col[1]));
                md[loc[i]] += kii * w;
            }
        }
        for (uint32_t l = 0; l < nVb; ++l) cnt[l + 1] += cnt[l];
        out.incPtr.insert(out.incPtr.end(), cnt.begin(), cnt.end());
        std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
        out.inc.resize(d.inc0 + 4 * (size_t)nTb);
        for (uint32_t tl = 0; tl < nTb; ++tl)
            for (uint32_t k = 0; k < 4; ++k) out.inc[d.inc0 + fill[Tet[4 * ((size_t)tb + tl) + k] - (uint32_t)vb]++] = (uint16_t)(tl * 4 + k);
        out.md.insert(out.md.end(), md.begin(), md.end());
        out.bodies.push_back(d);
        out.nVmax = std::max(out.nVmax, (nVb + 31u) & ~31u);
        out.nTmax = std::max(out.nTmax, (nTb + 31u) & ~31u);
    }
    if (t != nT) throw std::runtime_error("body batch: tets outside every body (bodies must own contiguous, ascending tet ranges)");
}

void partition_vertices(int nV, int world, std::vector<int>& vbeg)
{
    vbeg.assign((size_t)world + 1, 0);
    for (int r = 0; r <= world; ++r) vbeg[r] = (int)(((long long)nV * r) / world);
}

static inline int owner_of(const std::vector<int>& vbeg, uint32_t v)
{
    return (int)(std::upper_bound(vbeg.begin(), vbeg.end(), (int)v) - vbeg.begin()) - 1;
}

void build_rank_plan(const Layout& G, int world, int rank, RankPlan& P, bool trim)
{
    if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("rank plan: bad rank/world");
    P = RankPlan();
    P.rank = rank; P.world = world; P.trim = trim;
    partition_vertices(G.nV, world, P.vbeg);
    P.nOwn = P.vbeg[rank + 1] - P.vbeg[rank];
    // one pass over the tiles: which ranks evaluate it, and whose ghosts its vertices become
    std::vector<std::vector<uint32_t>> ghostsOf((size_t)world);
    std::vector<int> ranksOfTile;
    std::vector<uint32_t> boundaryTiles;
    for (int t = 0; t < G.nTiles; ++t) {
        const uint32_t* vl = G.vlist.data() + (size_t)t * TILE_NLMAX;
        ranksOfTile.clear();
        for (int l = 0; l < TILE_NLMAX && vl[l] != 0xffffffffu; ++l) {
            const int o = owner_of(P.vbeg, vl[l] & ~TILE_OWNER_BIT);
            if (std::find(ranksOfTile.begin(), ranksOfTile.end(), o) == ranksOfTile.end()) ranksOfTile.push_back(o);
        }
        if (std::find(ranksOfTile.begin(), ranksOfTile.end(), rank) != ranksOfTile.end())
            (ranksOfTile.size() > 1 ? boundaryTiles : P.tiles).push_back((uint32_t)t);
        if (ranksOfTile.size() > 1 && !trim)
            for (int l = 0; l < TILE_NLMAX && vl[l] != 0xffffffffu; ++l) {
                const uint32_t v = vl[l] & ~TILE_OWNER_BIT;
                const int o = owner_of(P.vbeg, v);
                for (int r : ranksOfTile) if (r != o) ghostsOf[(size_t)r].push_back(v);
            }
        if (ranksOfTile.size() > 1 && trim)          // a rank keeps a tet iff it owns one of its vertices; the others are its ghosts
            for (uint32_t tt = G.tileTetStart[(size_t)t]; tt < G.tileTetStart[(size_t)t + 1]; ++tt) {
                int o[4];
                for (int k = 0; k < 4; ++k) o[k] = owner_of(P.vbeg, G.tetNew[4 * (size_t)tt + k]);
                for (int k = 0; k < 4; ++k) {
                    bool first = true;
                    for (int j = 0; j < k; ++j) first = first && o[j] != o[k];
                    if (!first) continue;                                    // rank o[k] handled already for this tet
                    for (int j = 0; j < 4; ++j) if (o[j] != o[k]) ghostsOf[(size_t)o[k]].push_back(G.tetNew[4 * (size_t)tt + j]);
                }
            }
    }
    P.nInteriorTiles = (int)P.tiles.size();
    P.tiles.insert(P.tiles.end(), boundaryTiles.begin(), boundaryTiles.end());
    P.nLocOf.assign((size_t)world, 0);
    for (int r = 0; r < world; ++r) {
        std::vector<uint32_t>& g = ghostsOf[(size_t)r];
        std::sort(g.begin(), g.end());
        g.erase(std::unique(g.begin(), g.end()), g.end());
        P.nLocOf[(size_t)r] = (P.vbeg[r + 1] - P.vbeg[r]) + (int)g.size();
    }
    P.ghosts = ghostsOf[(size_t)rank];
    P.nGhost = (int)P.ghosts.size();
    // what this rank sends: its own vertices in the other ranks' ghost lists
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        const std::vector<uint32_t>& g = ghostsOf[(size_t)r];
        const int nOwnR = P.vbeg[r + 1] - P.vbeg[r];
        bool any = false;
        for (size_t i = 0; i < g.size(); ++i)
            if ((int)g[i] >= P.vbeg[rank] && (int)g[i] < P.vbeg[rank + 1]) {
                P.pushSrc.push_back(g[i] - (uint32_t)P.vbeg[rank]);
                P.pushDst.push_back((uint32_t)nOwnR + (uint32_t)i);
                P.pushRank.push_back(r);
                any = true;
            }
        if (any) P.neighbours.push_back(r);
    }
    // the relation is symmetric (a shared tile makes ghosts on both sides); keep the union anyway
    for (uint32_t g : P.ghosts) {
        const int o = owner_of(P.vbeg, g);
        if (std::find(P.neighbours.begin(), P.neighbours.end(), o) == P.neighbours.end()) P.neighbours.push_back(o);
    }
    std::sort(P.neighbours.begin(), P.neighbours.end());
}

bool dist_trim_from_env()
{
    // default ON since round 2 (same-box A/B on 2 GPUs: 3.21 -> 3.09-3.15 ms/step on grid70, redundant tets 6.0 % -> 1.5 %;
    // profiles/r2_trim_ab_n2_grid70.txt); PD_DIST_TRIM=0 keeps whole boundary tiles (A/B runs, the plan tests' specification)
    const char* e = std::getenv("PD_DIST_TRIM");
    return !(e && e[0] == '0');
}

void extract_rank_layout(const Layout& G, const RankPlan& P, Layout& L)
{
    const bool trim = P.trim;
    L = Layout();
    const int v0 = P.vbeg[P.rank];
    L.nV = P.nOwn + P.nGhost;
    std::vector<uint32_t> localOf((size_t)G.nV, 0xffffffffu);          // global renumbered id -> local id
    for (int i = 0; i < P.nOwn; ++i) localOf[(size_t)(v0 + i)] = (uint32_t)i;
    for (int i = 0; i < P.nGhost; ++i) localOf[P.ghosts[(size_t)i]] = (uint32_t)(P.nOwn + i);
    L.vertOrder.resize((size_t)L.nV);
    for (int g = 0; g < G.nV; ++g) if (localOf[(size_t)g] != 0xffffffffu) L.vertOrder[localOf[(size_t)g]] = G.vertOrder[(size_t)g];
    L.vertNewOfOld.assign((size_t)G.nV, 0xffffffffu);                   // ORIGINAL id -> local id (0xffffffff: not on this rank)
    for (int l = 0; l < L.nV; ++l) L.vertNewOfOld[L.vertOrder[(size_t)l]] = (uint32_t)l;
    std::vector<uint32_t> localTile((size_t)G.nTiles, 0xffffffffu);     // global tile -> local tile, for the tiles copied unchanged
    L.tileTetStart.push_back(0); L.tileRecOff.push_back(0);
    L.maxLocal = 0;
    L.nTiles = 0;
    auto open_tile = [&]() {
        L.vlist.resize((size_t)(L.nTiles + 1) * TILE_NLMAX, 0xffffffffu);
        L.vstage.resize((size_t)(L.nTiles + 1) * TILE_NLMAX, 0xffffffffu);
    };
    auto close_tile = [&]() {
        L.tileTetStart.push_back((uint32_t)L.tetOrder.size());
        L.tileRecOff.push_back((uint64_t)L.records.size());
        ++L.nTiles;
    };
    // a global tile exactly as it is: same record bytes, same slots
    auto copy_tile = [&](uint32_t gt) {
        const int lt = L.nTiles;
        open_tile();
        localTile[gt] = (uint32_t)lt;
        const uint64_t gb = G.tileRecOff[gt], ge = G.tileRecOff[gt + 1];
        const size_t base = L.records.size();
        L.records.insert(L.records.end(), G.records.begin() + (ptrdiff_t)gb, G.records.begin() + (ptrdiff_t)ge);
        TileHeader h; std::memcpy(&h, L.records.data() + base, sizeof(h));
        h.slotBase = (uint32_t)lt * (uint32_t)TILE_NLMAX;
        h.offLo = (uint32_t)(base & 0xffffffffu); h.offHi = (uint32_t)((uint64_t)base >> 32);
        std::memcpy(L.records.data() + base, &h, sizeof(h));
        L.tileTab.push_back(TileEntry{(uint64_t)base, h.abBytes, h.cBytes});
        L.maxLocal = std::max(L.maxLocal, (int)h.nLocal);
        const uint32_t t0 = G.tileTetStart[gt], t1 = G.tileTetStart[gt + 1];
        for (uint32_t t = t0; t < t1; ++t) {
            L.tetOrder.push_back(G.tetOrder[t]);
            L.tetGlobal.push_back(t);
            for (int k = 0; k < 4; ++k) L.tetNew.push_back(localOf[G.tetNew[4 * (size_t)t + k]]);
        }
        for (uint32_t l = 0; l < h.nLocal; ++l) {
            const uint32_t e = G.vlist[(size_t)gt * TILE_NLMAX + l];
            L.vlist[(size_t)lt * TILE_NLMAX + l] = localOf[e & ~TILE_OWNER_BIT] | (e & TILE_OWNER_BIT);
        }
        for (uint32_t sl = 0; sl < (uint32_t)TILE_NLMAX; ++sl) {
            const uint32_t e = G.vstage[(size_t)gt * TILE_NLMAX + sl];
            if (e != 0xffffffffu) L.vstage[(size_t)lt * TILE_NLMAX + sl] = localOf[e & ~TILE_OWNER_BIT] | (e & TILE_OWNER_BIT);
        }
        close_tile();
    };
    // ---- trimmed boundary tiles (experiment, see layout.hpp): slots of the (global tile, vertex) pairs they keep
    struct Part {               // one boundary tile cut down to the tets that touch an owned vertex
        uint32_t gt;
        std::vector<uint32_t> tets;                 // global reordered tet indices, ascending
        std::vector<uint32_t> verts;                // their distinct vertices: global id | owner bit of this tile's slot
    };
    std::vector<std::pair<uint64_t, uint32_t>> packedSlot;      // ((gt << 32) | global vertex, local slot), sorted below
    auto make_part = [&](uint32_t gt, Part& part) {
        part.gt = gt; part.tets.clear(); part.verts.clear();
        for (uint32_t t = G.tileTetStart[gt]; t < G.tileTetStart[gt + 1]; ++t) {
            bool owned = false;
            for (int k = 0; k < 4; ++k) owned = owned || localOf[G.tetNew[4 * (size_t)t + k]] < (uint32_t)P.nOwn;
            if (owned) part.tets.push_back(t);
        }
        // distinct vertices in the tile's own slot order, with the owner flag of that slot
        const uint32_t* vl = G.vlist.data() + (size_t)gt * TILE_NLMAX;
        std::vector<uint8_t> seen((size_t)TILE_NLMAX, 0);
        std::vector<uint32_t> ids;
        for (int l = 0; l < TILE_NLMAX && vl[l] != 0xffffffffu; ++l) ids.push_back(vl[l] & ~TILE_OWNER_BIT);
        for (uint32_t t : part.tets)
            for (int k = 0; k < 4; ++k) {
                const uint32_t v = G.tetNew[4 * (size_t)t + k];
                const size_t l = (size_t)(std::find(ids.begin(), ids.end(), v) - ids.begin());
                if (l >= ids.size()) throw std::runtime_error("rank layout: a tet's vertex is missing from its tile's list");
                seen[l] = 1;
            }
        for (size_t l = 0; l < ids.size(); ++l) if (seen[l]) part.verts.push_back(vl[l]);
    };
    // several consecutive parts -> ONE physical tile through the ordinary tile builder, on a small mesh whose vertex ids
    // are the (part, vertex) pairs; false if the builder needed more than one tile (the incidence row budget)
    auto emit_parts = [&](const Part* parts, size_t nParts) -> bool {
        std::vector<uint32_t> keyVert;              // key -> global id | owner bit
        std::vector<uint32_t> keyPart;
        std::vector<uint32_t> tet4;
        std::vector<float> B, w;
        std::vector<uint32_t> tetIdx;
        for (size_t pi = 0; pi < nParts; ++pi) {
            const Part& part = parts[pi];
            const uint32_t key0 = (uint32_t)keyVert.size();
            for (uint32_t e : part.verts) { keyVert.push_back(e); keyPart.push_back((uint32_t)pi); }
            const uint8_t* rec = G.records.data() + G.tileRecOff[part.gt];
            TileHeader h; std::memcpy(&h, rec, sizeof(h));
            for (uint32_t t : part.tets) {
                const uint32_t tl = t - G.tileTetStart[part.gt];
                for (int k = 0; k < 4; ++k) {
                    const uint32_t v = G.tetNew[4 * (size_t)t + k];
                    uint32_t key = key0;
                    while ((keyVert[key] & ~TILE_OWNER_BIT) != v) ++key;        // present by construction (make_part)
                    tet4.push_back(key);
                }
                for (uint32_t j = 0; j < 9; ++j) { float f; std::memcpy(&f, rec + tile_tet_word(h.nTets, tl, j), 4); B.push_back(f); }
                float wt; std::memcpy(&wt, rec + tile_tet_word(h.nTets, tl, 9), 4); w.push_back(wt);
                tetIdx.push_back(t);
            }
        }
        Layout T;
        T.nV = (int)keyVert.size(); T.nT = (int)tetIdx.size();
        build_tiles(T.nV, T.nT, tet4.data(), B.data(), w.data(), T);
        if (T.nTiles != 1) return false;
        const int lt = L.nTiles;
        open_tile();
        const size_t base = L.records.size();
        L.records.insert(L.records.end(), T.records.begin(), T.records.end());
        TileHeader h; std::memcpy(&h, L.records.data() + base, sizeof(h));
        h.slotBase = (uint32_t)lt * (uint32_t)TILE_NLMAX;
        h.offLo = (uint32_t)(base & 0xffffffffu); h.offHi = (uint32_t)((uint64_t)base >> 32);
        std::memcpy(L.records.data() + base, &h, sizeof(h));
        L.tileTab.push_back(TileEntry{(uint64_t)base, h.abBytes, h.cBytes});
        L.maxLocal = std::max(L.maxLocal, (int)h.nLocal);
        for (uint32_t t : tetIdx) {
            L.tetOrder.push_back(G.tetOrder[t]);
            L.tetGlobal.push_back(t);
            for (int k = 0; k < 4; ++k) L.tetNew.push_back(localOf[G.tetNew[4 * (size_t)t + k]]);
        }
        auto translate = [&](uint32_t e) -> uint32_t {      // the builder's entry (key | its own owner bit) -> local vertex | the global slot's owner bit
            const uint32_t kv = keyVert[e & ~TILE_OWNER_BIT];
            return localOf[kv & ~TILE_OWNER_BIT] | (kv & TILE_OWNER_BIT);
        };
        for (uint32_t l = 0; l < h.nLocal; ++l) {
            const uint32_t e = T.vlist[l];
            L.vlist[(size_t)lt * TILE_NLMAX + l] = translate(e);
            const uint32_t key = e & ~TILE_OWNER_BIT;
            packedSlot.push_back({((uint64_t)parts[keyPart[key]].gt << 32) | (keyVert[key] & ~TILE_OWNER_BIT), (uint32_t)lt * (uint32_t)TILE_NLMAX + l});
        }
        for (uint32_t sl = 0; sl < (uint32_t)TILE_NLMAX; ++sl)
            if (T.vstage[sl] != 0xffffffffu) L.vstage[(size_t)lt * TILE_NLMAX + sl] = translate(T.vstage[sl]);
        close_tile();
        return true;
    };

    for (int i = 0; i < P.nInteriorTiles; ++i) copy_tile(P.tiles[(size_t)i]);
    if (!trim) {
        for (size_t i = (size_t)P.nInteriorTiles; i < P.tiles.size(); ++i) copy_tile(P.tiles[i]);
    } else {
        std::vector<Part> group;
        size_t gTets = 0, gVerts = 0;
        auto flush = [&]() {
            if (group.empty()) return;
            if (!emit_parts(group.data(), group.size())) {
                if (group.size() == 1) throw std::runtime_error("rank layout: a trimmed tile does not fit one tile");
                for (const Part& part : group)
                    if (!emit_parts(&part, 1)) throw std::runtime_error("rank layout: a trimmed tile does not fit one tile");
            }
            group.clear(); gTets = gVerts = 0;
        };
        for (size_t i = (size_t)P.nInteriorTiles; i < P.tiles.size(); ++i) {
            Part part;
            make_part(P.tiles[i], part);
            if (part.tets.empty()) continue;            // (cannot happen: the tile is here because it holds an owned vertex)
            if (gTets + part.tets.size() > (size_t)TILE_T || gVerts + part.verts.size() > (size_t)TILE_NLMAX) flush();
            gTets += part.tets.size(); gVerts += part.verts.size();
            group.push_back(std::move(part));
        }
        flush();
        std::sort(packedSlot.begin(), packedSlot.end());
    }
    L.nT = (int)L.tetOrder.size();
    // vertex -> slots: the global lists restricted to this rank's tiles (complete for owned vertices), same order
    L.vslotPtr.assign((size_t)L.nV + 1, 0u);
    std::vector<uint32_t> globalOf((size_t)L.nV);
    for (int g = 0; g < G.nV; ++g) if (localOf[(size_t)g] != 0xffffffffu) globalOf[localOf[(size_t)g]] = (uint32_t)g;
    for (int l = 0; l < L.nV; ++l) {
        const uint32_t g = globalOf[(size_t)l];
        for (uint32_t e = G.vslotPtr[g]; e < G.vslotPtr[g + 1]; ++e) {
            const uint32_t slot = G.vslot[e], gt = slot / TILE_NLMAX, lt = localTile[gt];
            if (lt != 0xffffffffu) { L.vslot.push_back(lt * (uint32_t)TILE_NLMAX + slot % TILE_NLMAX); continue; }
            const uint64_t key = ((uint64_t)gt << 32) | g;
            const auto it = std::lower_bound(packedSlot.begin(), packedSlot.end(), std::make_pair(key, 0u));
            if (it != packedSlot.end() && it->first == key) L.vslot.push_back(it->second);
        }
        L.vslotPtr[(size_t)l + 1] = (uint32_t)L.vslot.size();
        if (l < P.nOwn && L.vslotPtr[(size_t)l + 1] - L.vslotPtr[(size_t)l] != G.vslotPtr[g + 1] - G.vslotPtr[g])
            throw std::runtime_error("rank layout: an owned vertex lost a tile");
    }
    L.nSlots = (uint32_t)L.vslot.size();
}

}  // namespace pdb200
