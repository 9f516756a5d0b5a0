// Host side of the mesh-mesh collision pass (see collision.hpp).
#include "collision.hpp"

#include <algorithm>
#include <array>
#include <numeric>
#include <stdexcept>

namespace pdb200 {

void boundary_faces(const uint32_t* Tet, int t0, int t1, std::vector<uint32_t>& tri)
{
    // the four faces of a tet in the reference's winding (dataLoader.cu:101-104)
    static const int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    struct Face { std::array<uint32_t, 3> key, wound; };
    std::vector<Face> faces;
    faces.reserve(4 * (size_t)(t1 - t0));
    for (int t = t0; t < t1; ++t)
        for (int f = 0; f < 4; ++f) {
            Face fc;
            for (int k = 0; k < 3; ++k) fc.wound[(size_t)k] = Tet[4 * (size_t)t + F[f][k]];
            fc.key = fc.wound;
            std::sort(fc.key.begin(), fc.key.end());
            faces.push_back(fc);
        }
    // the reference toggles a std::set per occurrence and keeps the LAST winding seen (std::map overwrite): a face survives when
    // it occurs an odd number of times; stable sort keeps the occurrences of a key in tet order
    std::stable_sort(faces.begin(), faces.end(), [](const Face& a, const Face& b) { return a.key < b.key; });
    for (size_t i = 0; i < faces.size();) {
        size_t j = i;
        while (j < faces.size() && faces[j].key == faces[i].key) ++j;
        if ((j - i) & 1u) for (int k = 0; k < 3; ++k) tri.push_back(faces[j - 1].wound[(size_t)k]);
        i = j;
    }
}

void scene_surface(const Scene& s, std::vector<uint32_t>& tri, std::vector<uint32_t>& father)
{
    tri.clear(); father.clear();
    if (s.triWholeScene) { tri = s.Tri; father = s.triFather; return; }
    const size_t nB = s.bodyTetStart.size();
    for (size_t b = 0; b < nB; ++b) {
        const size_t before = tri.size();
        if (b < s.bodyHasTri.size() && s.bodyHasTri[b]) {
            for (size_t t = 0; t < s.triFather.size(); ++t)
                if (s.triFather[t] == (uint32_t)b) for (int k = 0; k < 3; ++k) tri.push_back(s.Tri[3 * t + k]);
        } else {
            const int t0 = s.bodyTetStart[b], t1 = (b + 1 < nB) ? s.bodyTetStart[b + 1] : s.numTets;
            boundary_faces(s.Tet.data(), t0, t1, tri);
        }
        father.insert(father.end(), (tri.size() - before) / 3, (uint32_t)b);
    }
}

namespace {
struct Builder {
    const float* cen;                 // 3 per triangle
    std::vector<int>& order;          // triangle ids, permuted in place into leaf order
    std::vector<int>&left, &right;
    // returns the node id of the subtree over order[lo, hi): >= 0 internal, encoded leaves as -(position + 1)
    int build(int lo, int hi)
    {
        if (hi - lo == 1) return -(lo + 1);
        float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
        for (int k = lo; k < hi; ++k)
            for (int c = 0; c < 3; ++c) { const float x = cen[3 * (size_t)order[(size_t)k] + c]; mn[c] = std::min(mn[c], x); mx[c] = std::max(mx[c], x); }
        int ax = 0;
        for (int c = 1; c < 3; ++c) if (mx[c] - mn[c] > mx[ax] - mn[ax]) ax = c;
        const int mid = lo + (hi - lo) / 2;
        std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](int a, int b) {
            const float xa = cen[3 * (size_t)a + ax], xb = cen[3 * (size_t)b + ax];
            return xa < xb || (xa == xb && a < b);
        });
        const int me = (int)left.size();
        left.push_back(0); right.push_back(0);
        const int l = build(lo, mid), r = build(mid, hi);
        left[(size_t)me] = l; right[(size_t)me] = r;
        return me;
    }
};
}  // namespace

void build_collision_mesh(const Scene& s, CollisionMesh& M)
{
    M = CollisionMesh();
    std::vector<uint32_t> tri, father;
    scene_surface(s, tri, father);
    const int nT = (int)father.size();
    M.nTris = nT;
    M.nBodies = (int)s.bodyTetStart.size();
    if (s.triWholeScene) for (uint32_t f : father) M.nBodies = std::max(M.nBodies, (int)f + 1);
    for (size_t k = 0; k < tri.size(); ++k)
        if (tri[k] >= (uint32_t)s.numVerts) throw std::runtime_error("surface triangle " + std::to_string(k / 3) + " names a vertex outside the scene");
    if (nT == 0) { M.bodyRoot.assign((size_t)M.nBodies, -1); return; }
    // centroids at rest
    std::vector<float> cen(3 * (size_t)nT);
    for (int t = 0; t < nT; ++t)
        for (int c = 0; c < 3; ++c)
            cen[3 * (size_t)t + c] = (s.X[3 * (size_t)tri[3 * (size_t)t] + c] + s.X[3 * (size_t)tri[3 * (size_t)t + 1] + c] + s.X[3 * (size_t)tri[3 * (size_t)t + 2] + c]) * (1.0f / 3.0f);
    // triangles grouped by body (stable), one tree per body
    std::vector<int> order((size_t)nT);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return father[(size_t)a] < father[(size_t)b]; });
    std::vector<int> rootEnc((size_t)M.nBodies, 0);       // 0 = none
    Builder B{cen.data(), order, M.left, M.right};
    for (int lo = 0; lo < nT;) {
        int hi = lo;
        while (hi < nT && father[(size_t)order[(size_t)hi]] == father[(size_t)order[(size_t)lo]]) ++hi;
        const int enc = B.build(lo, hi);
        rootEnc[(size_t)father[(size_t)order[(size_t)lo]]] = (enc >= 0) ? enc + 1 : enc;       // internal ids shifted by one so that 0 stays "none"
        lo = hi;
    }
    M.nInternal = (int)M.left.size();
    auto node_id = [&](int enc) { return enc >= 0 ? enc : M.nInternal + (-enc - 1); };
    for (int i = 0; i < M.nInternal; ++i) { M.left[(size_t)i] = node_id(M.left[(size_t)i]); M.right[(size_t)i] = node_id(M.right[(size_t)i]); }
    M.parent.assign((size_t)M.nInternal + nT, -1);
    for (int i = 0; i < M.nInternal; ++i) { M.parent[(size_t)M.left[(size_t)i]] = i; M.parent[(size_t)M.right[(size_t)i]] = i; }
    M.bodyRoot.assign((size_t)M.nBodies, -1);
    for (int b = 0; b < M.nBodies; ++b)
        if (rootEnc[(size_t)b] != 0) M.bodyRoot[(size_t)b] = rootEnc[(size_t)b] > 0 ? rootEnc[(size_t)b] - 1 : M.nInternal + (-rootEnc[(size_t)b] - 1);
    // triangles in leaf order
    M.tri.resize(3 * (size_t)nT); M.father.resize((size_t)nT);
    for (int k = 0; k < nT; ++k) {
        const int t = order[(size_t)k];
        for (int c = 0; c < 3; ++c) M.tri[3 * (size_t)k + c] = tri[3 * (size_t)t + c];
        M.father[(size_t)k] = father[(size_t)t];
    }
    // unique edges, ascending (v0, v1); edge ids of every triangle's local edges (0,1), (0,2), (1,2)
    static const int E[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    std::vector<std::pair<uint32_t, uint32_t>> ed;
    ed.reserve(3 * (size_t)nT);
    for (int k = 0; k < nT; ++k)
        for (int e = 0; e < 3; ++e) {
            const uint32_t a = M.tri[3 * (size_t)k + E[e][0]], b = M.tri[3 * (size_t)k + E[e][1]];
            ed.emplace_back(std::min(a, b), std::max(a, b));
        }
    std::vector<std::pair<uint32_t, uint32_t>> uniq(ed);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    M.nEdges = (int)uniq.size();
    M.edge.resize(2 * uniq.size());
    for (size_t i = 0; i < uniq.size(); ++i) { M.edge[2 * i] = uniq[i].first; M.edge[2 * i + 1] = uniq[i].second; }
    M.triEdge.resize(3 * (size_t)nT);
    for (size_t i = 0; i < ed.size(); ++i) M.triEdge[i] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), ed[i]) - uniq.begin());
}

}  // namespace pdb200
