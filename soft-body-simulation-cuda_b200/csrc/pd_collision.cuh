// Mesh-mesh collision pass of PdSolver::Update (pdSolver.cu:218-231) -- the step between SolverStep and the fixed-body
// response when SolverParams::handleCollision is set -- headless, on the engine's own arrays.  Reference:
//   CollisionDetection::DetectCollision (bvh.cu:170-180): tI <- 1; BroadPhaseCCD (broadphase.cu:539-554): swept boxes of the
//   surface triangles (ccd.cu:53-66), LBVH, traverseTree (broadphase.cu:327-400) -> 12 queries per overlapping pair,
//   sortEachQuery + removeDuplicates (:453-491); NarrowPhase (narrowphase.cu:122-134): ccdCollisionTest per query
//   (intersections.cu:312-355), sort + unique keeps the earliest hit per vertex (VF) / per first edge (EE), storeTi (:74-119);
//   then CCDKernel (collisionUtil.cu:49-70): a vertex with tI < 1 keeps its OLD position and gets V = -(n . dx) n, every other
//   vertex X <- XTilde.
// What is built differently (collision.hpp): one tree per body from the rest shape, refitted every step (the set of overlapping
// leaf pairs does not depend on the tree); no Query structs, no sorts -- a pair's 12 tests fold straight into one 64-bit
// atomicMin per vertex / per edge (earliest hit first, smallest partner id on ties where the reference's unstable sort is
// arbitrary); the reference's racing storeTi writes become "the LAST group in the reference's sorted order wins", which is
// what its kernel gives when its threads run in index order.  Every run is bit-identical.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace pdb200 {

struct ColMeshDev {
    int nTris, nEdges, nInternal, nBodies, nV;
    const uint32_t* tri;        // 3 * nTris, ORIGINAL vertex ids (the reference's numbering: its sorts and tie-breaks use these), leaf order
    const uint32_t* father;     // nTris
    const uint32_t* edge;       // 2 * nEdges, sorted endpoints (original ids), ascending
    const uint32_t* triEdge;    // 3 * nTris
    const int *left, *right;    // nInternal
    const int* parent;          // nInternal + nTris
    const int* bodyRoot;        // nBodies
    const uint32_t* newOfOld;   // nV: engine (renumbered) id of an original vertex id
    float4 *boxMin, *boxMax;    // nInternal + nTris
    int* visit;                 // nInternal arrival counters of the refit (zeroed every step)
};
constexpr unsigned long long COL_EMPTY = 0xffffffffffffffffull;

// ------------------------------------------------------------------ float vector helpers in glm's operation order
namespace ccd {
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld(const float4* a, uint32_t i) { const float4 v = a[i]; return mk(v.x, v.y, v.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                 // func_geometric.inl:65-72
__device__ __forceinline__ V3 cross(V3 x, V3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ float length2(V3 a) { return dot(a, a); }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }                          // x * inversesqrt(dot(x, x))
__device__ __forceinline__ float stp(V3 u, V3 v, V3 w) { return dot(u, cross(v, w)); }                           // intersections.cu:27

// intersections.cu:95-114
__device__ __forceinline__ float newtons_method(float a, float b, float c, float d, float x0, int init_dir)
{
    if (init_dir != 0) {
        const float y0 = d + x0 * (c + x0 * (b + x0 * a)), ddy0 = 2 * b + x0 * (6 * a);
        if (ddy0 != 0) x0 += init_dir * sqrtf(fabsf(2 * y0 / ddy0));
    }
    for (int iter = 0; iter < 100; iter++) {
        const float y = d + x0 * (c + x0 * (b + x0 * a));
        const float dy = c + x0 * (2 * b + x0 * 3 * a);
        if (dy == 0) return x0;
        const float x1 = x0 - y / dy;
        if ((double)fabsf(x0 - x1) < 1e-6) return x0;
        x0 = x1;
    }
    return x0;
}
// intersections.cu:74-92
__device__ __forceinline__ int solve_quadratic(float a, float b, float c, float* x)
{
    const float d = b * b - 4 * a * c;
    if (d < 0) { x[0] = -b / (2 * a); return 0; }
    const float sgn = (float)(0.f < b) - (float)(b < 0.f);             // glm::sign
    const float q = -(b + sgn * sqrtf(d)) / 2;
    int i = 0;
    if ((double)fabsf(a) > 1e-12 * (double)fabsf(q)) x[i++] = q / a;
    if ((double)fabsf(q) > 1e-12 * (double)fabsf(c)) x[i++] = c / q;
    if (i == 2 && x[0] > x[1]) { const float t = x[0]; x[0] = x[1]; x[1] = t; }
    return i;
}
// intersections.cu:46-72
__device__ __forceinline__ int solve_cubic(float a, float b, float c, float d, float* x)
{
    float xc[2];
    const int ncrit = solve_quadratic(3 * a, 2 * b, c, xc);
    if (ncrit == 0) { x[0] = newtons_method(a, b, c, d, xc[0], 0); return 1; }
    if (ncrit == 1) return solve_quadratic(b, c, d, x);
    const float yc[2] = {d + xc[0] * (c + xc[0] * (b + xc[0] * a)), d + xc[1] * (c + xc[1] * (b + xc[1] * a))};
    int i = 0;
    if (yc[0] * a >= 0) x[i++] = newtons_method(a, b, c, d, xc[0], -1);
    if (yc[0] * yc[1] <= 0) {
        const int closer = fabsf(yc[0]) < fabsf(yc[1]) ? 0 : 1;
        x[i++] = newtons_method(a, b, c, d, xc[closer], closer == 0 ? 1 : -1);
    }
    if (yc[1] * a <= 0) x[i++] = newtons_method(a, b, c, d, xc[1], 1);
    return i;
}
// intersections.cu:115-134 / :157-174; w = the four weights
__device__ __forceinline__ float signed_vf_distance(V3 x, V3 y0, V3 y1, V3 y2, V3& n, float w[4])
{
    n = cross(normalize(y1 - y0), normalize(y2 - y0));
    if ((double)length2(n) < 1e-6) return FLT_MAX;
    n = normalize(n);
    const float h = dot(x - y0, n);
    const float b0 = stp(y1 - x, y2 - x, n), b1 = stp(y2 - x, y0 - x, n), b2 = stp(y0 - x, y1 - x, n);
    w[0] = 1; w[1] = -b0 / (b0 + b1 + b2); w[2] = -b1 / (b0 + b1 + b2); w[3] = -b2 / (b0 + b1 + b2);
    return h;
}
__device__ __forceinline__ float signed_ee_distance(V3 x0, V3 x1, V3 y0, V3 y1, V3& n, float w[4])
{
    n = cross(normalize(x1 - x0), normalize(y1 - y0));
    if ((double)length2(n) < 1e-6) return FLT_MAX;
    n = normalize(n);
    const float h = dot(x0 - y0, n);
    const float a0 = stp(y1 - x1, y0 - x1, n), a1 = stp(y0 - x0, y1 - x0, n), b0 = stp(x0 - y1, x1 - y1, n), b1 = stp(x1 - y0, x0 - y0, n);
    w[0] = a0 / (a0 + a1); w[1] = a1 / (a0 + a1); w[2] = -b0 / (b0 + b1); w[3] = -b1 / (b0 + b1);
    return h;
}
// ccdCollisionTest<float> (intersections.cu:312-355): ee = edge-edge query, else vertex-face; ids are ENGINE vertex ids.
// Returns the time of impact in [-1e-12, 1] or 1.0 (no hit); n = the contact normal of the accepted root.  (The reference
// leaves w, d and inside uninitialised when signed_*_distance bails out early; the FLT_MAX distance then fails |d| < 1e-6.)
__device__ __forceinline__ float collision_test(bool ee, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, const float4* __restrict__ X,
                                                const float4* __restrict__ XT, V3& n)
{
    const V3 x0 = ld(X, i0), x1 = ld(X, i1), x2 = ld(X, i2), x3 = ld(X, i3);
    const V3 v0 = ld(XT, i0) - x0, v1 = ld(XT, i1) - x1, v2 = ld(XT, i2) - x2, v3 = ld(XT, i3) - x3;
    const V3 x01 = x1 - x0, x02 = x2 - x0, x03 = x3 - x0, v01 = v1 - v0, v02 = v2 - v0, v03 = v3 - v0;
    const float a0 = stp(x01, x02, x03);
    const float a1 = stp(v01, x02, x03) + stp(x01, v02, x03) + stp(x01, x02, v03);
    const float a2 = stp(x01, v02, v03) + stp(v01, x02, v03) + stp(v01, v02, x03);
    const float a3 = stp(v01, v02, v03);
    n = mk(0.f, 0.f, 0.f);
    if ((double)fabsf(a0) < 1e-12 * (double)length(x01) * (double)length(x02) * (double)length(x03)) return 1.0f;      // initially coplanar
    float t[3];
    const int nsol = solve_cubic(a3, a2, a1, a0, t);
    for (int i = 0; i < nsol; i++) {
        if ((double)t[i] < -1e-12 || t[i] > 1) continue;
        const V3 xt0 = x0 + t[i] * v0, xt1 = x1 + t[i] * v1, xt2 = x2 + t[i] * v2, xt3 = x3 + t[i] * v3;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        float d;
        bool inside;
        if (!ee) {
            d = signed_vf_distance(xt0, xt1, xt2, xt3, n, w);
            inside = (double)fminf(-w[1], fminf(-w[2], -w[3])) >= -1e-3;
        } else {
            d = signed_ee_distance(xt0, xt1, xt2, xt3, n, w);
            inside = (double)fminf(w[0], fminf(w[1], fminf(-w[2], -w[3]))) >= -1e-3;
        }
        if (dot(n, w[1] * v1 + w[2] * v2 + w[3] * v3) > 0) n = -n;
        if ((double)fabsf(d) < 1e-6 && inside) return t[i];
    }
    return 1.0f;
}
}  // namespace ccd

// ------------------------------------------------------------------ 1. swept leaf boxes + bottom-up refit
// computeTriTrajBBoxCCD (ccd.cu:53-66): min / max over the triangle at X and at XTilde, grown by AABBThreshold = 0.01
__global__ void k_col_refit(ColMeshDev M, const float4* __restrict__ X, const float4* __restrict__ XT)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M.nTris) return;
    const uint32_t a = M.newOfOld[M.tri[3 * k]], b = M.newOfOld[M.tri[3 * k + 1]], c = M.newOfOld[M.tri[3 * k + 2]];
    const float4 p0 = X[a], p1 = X[b], p2 = X[c], p3 = XT[a], p4 = XT[b], p5 = XT[c];
    float4 mn, mx;
    mn.x = fminf(fminf(fminf(fminf(fminf(p0.x, p1.x), p2.x), p3.x), p4.x), p5.x) - 0.01f;
    mn.y = fminf(fminf(fminf(fminf(fminf(p0.y, p1.y), p2.y), p3.y), p4.y), p5.y) - 0.01f;
    mn.z = fminf(fminf(fminf(fminf(fminf(p0.z, p1.z), p2.z), p3.z), p4.z), p5.z) - 0.01f;
    mx.x = fmaxf(fmaxf(fmaxf(fmaxf(fmaxf(p0.x, p1.x), p2.x), p3.x), p4.x), p5.x) + 0.01f;
    mx.y = fmaxf(fmaxf(fmaxf(fmaxf(fmaxf(p0.y, p1.y), p2.y), p3.y), p4.y), p5.y) + 0.01f;
    mx.z = fmaxf(fmaxf(fmaxf(fmaxf(fmaxf(p0.z, p1.z), p2.z), p3.z), p4.z), p5.z) + 0.01f;
    mn.w = mx.w = 0.f;
    int node = M.nInternal + k;
    M.boxMin[node] = mn; M.boxMax[node] = mx;
    // climb: the SECOND thread to arrive at a node finds both children final (its own by program order, the other's through the
    // fence-before-counter of the first arrival) and carries the union further up
    node = M.parent[node];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&M.visit[node], 1) == 0) return;
        __threadfence();
        const int l = M.left[node], r = M.right[node];
        const float4 a0 = __ldcg(&M.boxMin[l]), a1 = __ldcg(&M.boxMax[l]), b0 = __ldcg(&M.boxMin[r]), b1 = __ldcg(&M.boxMax[r]);
        mn = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.f);
        mx = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.f);
        M.boxMin[node] = mn; M.boxMax[node] = mx;
        node = M.parent[node];
    }
}

// ------------------------------------------------------------------ 2. overlapping leaf pairs
__device__ __forceinline__ bool col_overlap(const float4 amin, const float4 amax, const float4 bmin, const float4 bmax)
{   // bboxIntersectionTest (intersections.cu:222-228)
    if (amax.x < bmin.x || amin.x > bmax.x) return false;
    if (amax.y < bmin.y || amin.y > bmax.y) return false;
    if (amax.z < bmin.z || amin.z > bmax.z) return false;
    return true;
}
// traverseTree (broadphase.cu:327-400): every triangle against the other triangles' boxes; (i, j) AND (j, i) are both found,
// as in the reference (each order carries its own vertex-face and first-edge tests)
__global__ void k_col_traverse(ColMeshDev M, int ignoreSelf, uint2* __restrict__ pairs, unsigned int* __restrict__ pairCount, unsigned int maxPairs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M.nTris) return;
    const float4 mn = M.boxMin[M.nInternal + i], mx = M.boxMax[M.nInternal + i];
    const uint32_t fi = M.father[i];
    const uint32_t a0 = M.tri[3 * i], a1 = M.tri[3 * i + 1], a2 = M.tri[3 * i + 2];
    int stack[64];
    for (int b = 0; b < M.nBodies; ++b) {
        const int root = M.bodyRoot[b];
        if (root < 0 || (ignoreSelf && (uint32_t)b == fi)) continue;          // a tree holds ONE father's triangles
        int sp = 0;
        stack[sp++] = root;
        while (sp) {
            const int node = stack[--sp];
            if (!col_overlap(mn, mx, M.boxMin[node], M.boxMax[node])) continue;
            if (node < M.nInternal) { stack[sp++] = M.left[node]; stack[sp++] = M.right[node]; continue; }
            const int j = node - M.nInternal;
            if (j == i) continue;
            const uint32_t b0 = M.tri[3 * j], b1 = M.tri[3 * j + 1], b2 = M.tri[3 * j + 2];
            if (a0 == b0 || a0 == b1 || a0 == b2 || a1 == b0 || a1 == b1 || a1 == b2 || a2 == b0 || a2 == b1 || a2 == b2) continue;    // isAdjacentTriangle
            const unsigned int at = atomicAdd(pairCount, 1u);
            if (at < maxPairs) pairs[at] = make_uint2((unsigned)i, (unsigned)j);
        }
    }
}

// ------------------------------------------------------------------ 3. narrow phase, folded into per-group minima
__device__ __forceinline__ void col_sort2(uint32_t& a, uint32_t& b) { if (a > b) { const uint32_t t = a; a = b; b = t; } }
__device__ __forceinline__ void col_sort3(uint32_t& a, uint32_t& b, uint32_t& c) { col_sort2(a, b); col_sort2(b, c); col_sort2(a, b); }      // sortThree
// earliest hit first (non-negative floats order like their bit patterns; the [-1e-12, 0) sliver counts as 0), then the
// smallest partner id
__device__ __forceinline__ unsigned long long col_key(float toi, uint32_t partner)
{
    return ((unsigned long long)__float_as_uint(fmaxf(toi, 0.f)) << 32) | partner;
}
// the vertices of test `test` (0..2 vertex-face, 3..11 edge-edge) of the ordered pair (i, j), in the reference's sorted form
// (sortEachQuery, broadphase.cu:453-478); false = the query degenerates (UNKNOWN)
__device__ __forceinline__ bool col_query(const ColMeshDev& M, uint32_t i, uint32_t j, int test, uint32_t q[4], uint32_t& group, uint32_t& partner)
{
    if (test < 3) {
        q[0] = M.tri[3 * i + test]; q[1] = M.tri[3 * j]; q[2] = M.tri[3 * j + 1]; q[3] = M.tri[3 * j + 2];
        if (q[0] == q[1] || q[0] == q[2] || q[0] == q[3]) return false;
        col_sort3(q[1], q[2], q[3]);
        group = q[0]; partner = j;
        return true;
    }
    const uint32_t e1 = M.triEdge[3 * i + (test - 3) / 3], e2 = M.triEdge[3 * j + (test - 3) % 3];
    if (e1 == e2) return false;
    q[0] = M.edge[2 * e1]; q[1] = M.edge[2 * e1 + 1]; q[2] = M.edge[2 * e2]; q[3] = M.edge[2 * e2 + 1];
    group = e1; partner = e2;
    return true;
}
__global__ void k_col_narrow(ColMeshDev M, const float4* __restrict__ X, const float4* __restrict__ XT, const uint2* __restrict__ pairs,
                             const unsigned int* __restrict__ pairCount, unsigned int maxPairs, unsigned long long* __restrict__ vfBest,
                             unsigned long long* __restrict__ eeBest)
{
    const unsigned int nP = min(*pairCount, maxPairs);
    for (unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 12u * nP; idx += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[idx / 12u];
        const int test = (int)(idx % 12u);
        uint32_t q[4], group, partner;
        if (!col_query(M, pr.x, pr.y, test, q, group, partner)) continue;
        ccd::V3 n;
        const float toi = ccd::collision_test(test >= 3, M.newOfOld[q[0]], M.newOfOld[q[1]], M.newOfOld[q[2]], M.newOfOld[q[3]], X, XT, n);
        if (toi < 1.0f) atomicMin(test < 3 ? &vfBest[group] : &eeBest[group], col_key(toi, partner));
    }
}

// ------------------------------------------------------------------ 4. storeTi (narrowphase.cu:74-119), deterministic
// Group g = vertex g (g < nV: its earliest vertex-face hit) or edge g - nV (its earliest edge-edge hit as FIRST edge).  The
// reference launches one thread per surviving query, vertex-face queries first (ascending vertex), then edge-edge (ascending
// edge), and lets their writes race; here the group with the HIGHEST index that touches a vertex wins it.
__device__ __forceinline__ bool col_group(const ColMeshDev& M, int g, const unsigned long long* vfBest, const unsigned long long* eeBest, bool& ee,
                                          uint32_t q[4])
{
    ee = g >= M.nV;
    const unsigned long long best = ee ? eeBest[g - M.nV] : vfBest[g];
    if (best == COL_EMPTY) return false;
    const uint32_t partner = (uint32_t)(best & 0xffffffffull);
    if (!ee) {
        q[0] = (uint32_t)g; q[1] = M.tri[3 * partner]; q[2] = M.tri[3 * partner + 1]; q[3] = M.tri[3 * partner + 2];
        col_sort3(q[1], q[2], q[3]);
    } else {
        const uint32_t e1 = (uint32_t)(g - M.nV);
        q[0] = M.edge[2 * e1]; q[1] = M.edge[2 * e1 + 1]; q[2] = M.edge[2 * partner]; q[3] = M.edge[2 * partner + 1];
    }
    return true;
}
__global__ void k_col_rank(ColMeshDev M, const unsigned long long* __restrict__ vfBest, const unsigned long long* __restrict__ eeBest, int* __restrict__ writer)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M.nV + M.nEdges) return;
    bool ee; uint32_t q[4];
    if (!col_group(M, g, vfBest, eeBest, ee, q)) return;
    const int nw = ee ? 2 : 4;                       // an edge-edge hit marks its FIRST edge only, a vertex-face hit all four vertices
    for (int k = 0; k < nw; ++k) atomicMax(&writer[q[k]], g + 1);
}
__global__ void k_col_store(ColMeshDev M, const float4* __restrict__ X, const float4* __restrict__ XT, const unsigned long long* __restrict__ vfBest,
                            const unsigned long long* __restrict__ eeBest, const int* __restrict__ writer, float* __restrict__ tI, float4* __restrict__ nors)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M.nV + M.nEdges) return;
    bool ee; uint32_t q[4];
    if (!col_group(M, g, vfBest, eeBest, ee, q)) return;
    ccd::V3 n;
    ccd::collision_test(ee, M.newOfOld[q[0]], M.newOfOld[q[1]], M.newOfOld[q[2]], M.newOfOld[q[3]], X, XT, n);      // the winning query again: its normal
    const int nw = ee ? 2 : 4;
    for (int k = 0; k < nw; ++k) {
        if (writer[q[k]] != g + 1) continue;
        const uint32_t v = M.newOfOld[q[k]];
        const float s = (!ee && k > 0) ? -1.f : 1.f;        // vertex-face: the vertex gets n, the face's vertices -n
        tI[v] = 0.5f;
        nors[v] = make_float4(s * n.x, s * n.y, s * n.z, 0.f);
    }
}

// ------------------------------------------------------------------ 5. CCDKernel (collisionUtil.cu:49-70)
__global__ void k_col_apply(int nV, float4* __restrict__ X, const float4* __restrict__ XT, float4* __restrict__ V, const float* __restrict__ tI,
                            const float4* __restrict__ nors)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 xt = XT[v];
    if (tI[v] < 1.0f) {
        const float4 x = X[v], n = nors[v];
        const ccd::V3 vel = ccd::mk(xt.x - x.x, xt.y - x.y, xt.z - x.z), nn = ccd::mk(n.x, n.y, n.z);
        const ccd::V3 vn = ccd::dot(vel, nn) * nn;
        V[v] = make_float4(-vn.x, -vn.y, -vn.z, 0.f);
    } else {
        X[v] = xt;
    }
}

}  // namespace pdb200
