// Hand-written sm_100a kernels for the projective-dynamics step (Jacobi/Chebyshev path).
// Reference behaviour restated per kernel; see DESIGN.md section 4 for the data layout and
// the roofline of each.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "layout.hpp"
#include "rotation.cuh"

namespace pdb200 {

// ------------------------------------------------------------------ device-side fixed bodies
struct DevFixedBodies {
    int nPlanes, nSpheres, nCyls;
    const float* planes;    // 6 per plane: p0[3], up[3]
    const float* spheres;   // 4 per sphere: c[3], r
    const float* cyls;      // 7 per cylinder: c[3], axis[3], r
};

// ------------------------------------------------------------------ mbarrier / bulk-copy (TMA) PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one bulk asynchronous copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ predictor
// gravity transform + setMDt_2MoreDBC + computeSn + the two D2D copies
// (pdSolver.cu:154-160, pdUtil.cu:56-95).  moreDBC == 0 (no mouse drag on the headless path).
// Writes q0 = prev = s, so4 = (s | DBCX, dbcFlag), cc = (c, c + matrix_diag).
__global__ void k_predict(int nV, const float4* __restrict__ X, const float4* __restrict__ V,
                          const float* __restrict__ mass, const float* __restrict__ dbc,
                          const float* __restrict__ md, const float4* __restrict__ X0,
                          float dt, float dt2Prepared, float gravity,
                          float4* __restrict__ q0, float4* __restrict__ qprev, float4* __restrict__ so4,
                          float2* __restrict__ cc)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float m = mass[v], isDbc = dbc[v];
    const float dt2 = dt * dt;
    // DBC vertices keep the massDt_2s computed by setMDt_2 at prepare time (pdUtil.cu:48,59)
    const float c = (isDbc == 0.f) ? m / dt2 : (m + isDbc * 1e6f) / dt2Prepared;
    const float4 x = X[v], vel = V[v];
    const float fy = -gravity * m;
    const float dt2_m_1 = 1.0f / c;
    float4 s;
    s.x = x.x + dt * vel.x + dt2_m_1 * 0.0f;
    s.y = x.y + dt * vel.y + dt2_m_1 * fy;
    s.z = x.z + dt * vel.z + dt2_m_1 * 0.0f;
    s.w = 0.f;
    q0[v] = s;
    qprev[v] = s;
    float4 so = s;
    if (isDbc > 0.f) { const float4 d = X0[v]; so = make_float4(d.x, d.y, d.z, 1.0f); }   // DBCX = X0 (pdSolver.cu:134)
    so4[v] = so;
    cc[v] = make_float2(c, c + md[v]);
}

// ------------------------------------------------------------------ local step (the hot kernel)
// PdUtil::computeLocal (pdUtil.cu:97-145) for one tile of <= TILE_T tets per CTA iteration:
//   0. ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) lands the packed tile record in
//      shared memory, double buffered: tile i+1 streams in while tile i is computed;
//   A. the tile's distinct vertex positions are gathered once into shared memory;
//   B. one tet per thread: 3 x LDS.128 (48-byte record: DmInv, w, corner offsets; the 48-byte
//      stride is bank-conflict free), 4 x LDS.128 positions, F = Ds*DmInv, rotation,
//      H = w (R - F) DmInv^T G (or w R DmInv^T G), 4 x STS.128 into the H scratch;
//   C. one tile-local vertex per thread: ordered sum over the vertex's incidence list (entries are
//      ready-made byte offsets into the H scratch) -> ONE partial sum per (tile, vertex) slot.
// No atomics anywhere: the reference's 12 float atomicAdds per tet become ordered sums, so results
// are run-to-run bit-identical.  Two __syncthreads per tile.
// F and H are written with the fused/rounded operation pattern nvcc gives the reference's glm
// expressions (see oracle/pd_oracle.c header), so that with ROT_MODE 1 every tet contribution is
// bit-identical to the reference kernel's.
constexpr uint32_t LOCAL_OFF_QS = 2u * TILE_RECMAX;
constexpr uint32_t LOCAL_OFF_HS = LOCAL_OFF_QS + 16u * TILE_NLMAX;
constexpr uint32_t LOCAL_OFF_BAR = LOCAL_OFF_HS + 4u * TILE_HSTRIDE;
constexpr uint32_t LOCAL_SMEM_BYTES = LOCAL_OFF_BAR + 16u;

__device__ __forceinline__ float dot3_nv(float a0, float b0, float a1, float b1, float a2, float b2)
{   // a0*b0 + a1*b1 + a2*b2 as nvcc contracts the reference's glm products
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

template <int ROT_MODE, bool JACOBI>
__global__ void __launch_bounds__(TILE_T, 4)
k_local(const uint8_t* __restrict__ records, const unsigned long long* __restrict__ recOff, int nTiles,
        const float4* __restrict__ q, float4* __restrict__ P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + LOCAL_OFF_BAR);
    const int tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    int tile = blockIdx.x;
    if (tid == 0 && tile < nTiles) {
        const unsigned long long o = recOff[tile];
        const uint32_t bytes = (uint32_t)(recOff[tile + 1] - o);
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(smem, records + o, bytes, &bar[0]);
    }
    for (int it = 0; tile < nTiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));

        const uint8_t* rec = smem + b * TILE_RECMAX;
        const uint4 hdr = *reinterpret_cast<const uint4*>(rec);      // nTets, nLocal, slotBase, recBytes
        const uint32_t nTets = hdr.x, nLocal = hdr.y;
        const uint32_t offI = 16u + 48u * nTets;
        const uint32_t offIO = offI + ((8u * nTets + 15u) & ~15u);
        const uint32_t offV = offIO + ((2u * (nLocal + 1u) + 15u) & ~15u);

        // phase A: stage the tile's vertex positions
        {
            const uint32_t* vlist = reinterpret_cast<const uint32_t*>(rec + offV);
            float4* qs = reinterpret_cast<float4*>(smem + LOCAL_OFF_QS);
            for (uint32_t l = tid; l < nLocal; l += TILE_T) qs[l] = __ldg(&q[vlist[l]]);
        }
        __syncthreads();   // positions staged; every warp is past phase C of the previous tile

        // the other buffer and the H scratch are free now: stream the next record in
        const int nextTile = tile + gridDim.x;
        if (tid == 0 && nextTile < nTiles) {
            const unsigned long long o = recOff[nextTile];
            const uint32_t bytes = (uint32_t)(recOff[nextTile + 1] - o);
            mbar_expect_tx(&bar[b ^ 1], bytes);
            bulk_g2s(smem + (b ^ 1) * TILE_RECMAX, records + o, bytes, &bar[b ^ 1]);
        }

        // phase B: one tet per thread
        if ((uint32_t)tid < nTets) {
            const float4* tr = reinterpret_cast<const float4*>(rec + 16 + 48 * tid);
            const float4 r0 = tr[0], r1 = tr[1], r2 = tr[2];
            const float B0 = r0.x, B1 = r0.y, B2 = r0.z, B3 = r0.w, B4 = r1.x, B5 = r1.y, B6 = r1.z, B7 = r1.w, B8 = r2.x;
            const float w = r2.y;
            const uint32_t c01 = __float_as_uint(r2.z), c23 = __float_as_uint(r2.w);
            const uint8_t* qsb = smem + LOCAL_OFF_QS;
            const float4 p0 = *reinterpret_cast<const float4*>(qsb + (c01 & 0xffffu));
            const float4 p1 = *reinterpret_cast<const float4*>(qsb + (c01 >> 16));
            const float4 p2 = *reinterpret_cast<const float4*>(qsb + (c23 & 0xffffu));
            const float4 p3 = *reinterpret_cast<const float4*>(qsb + (c23 >> 16));
            // Ds columns = edges; F = Ds * DmInv  (row-major F[r][c] = sum_k Ds[r][k] B[k][c])
            const float d00 = p1.x - p0.x, d01 = p2.x - p0.x, d02 = p3.x - p0.x;
            const float d10 = p1.y - p0.y, d11 = p2.y - p0.y, d12 = p3.y - p0.y;
            const float d20 = p1.z - p0.z, d21 = p2.z - p0.z, d22 = p3.z - p0.z;
            Mat3 F, R;
            F.m[0] = dot3_nv(d00, B0, d01, B3, d02, B6); F.m[1] = dot3_nv(d00, B1, d01, B4, d02, B7); F.m[2] = dot3_nv(d00, B2, d01, B5, d02, B8);
            F.m[3] = dot3_nv(d10, B0, d11, B3, d12, B6); F.m[4] = dot3_nv(d10, B1, d11, B4, d12, B7); F.m[5] = dot3_nv(d10, B2, d11, B5, d12, B8);
            F.m[6] = dot3_nv(d20, B0, d21, B3, d22, B6); F.m[7] = dot3_nv(d20, B1, d21, B4, d22, B7); F.m[8] = dot3_nv(d20, B2, d21, B5, d22, B8);
            corotation<ROT_MODE>(F, R);
            float M[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) M[e] = __fmul_rn(w, JACOBI ? __fsub_rn(R.m[e], F.m[e]) : R.m[e]);
            // H = M * DmInv^T ; column j (vertex j+1): H[r][j] = sum_k M[r][k] B[j][k] ; vertex 0: -(sum of columns)
            float4 h0, h1, h2, h3;
            h1.x = dot3_nv(M[0], B0, M[1], B1, M[2], B2); h2.x = dot3_nv(M[0], B3, M[1], B4, M[2], B5); h3.x = dot3_nv(M[0], B6, M[1], B7, M[2], B8);
            h1.y = dot3_nv(M[3], B0, M[4], B1, M[5], B2); h2.y = dot3_nv(M[3], B3, M[4], B4, M[5], B5); h3.y = dot3_nv(M[3], B6, M[4], B7, M[5], B8);
            h1.z = dot3_nv(M[6], B0, M[7], B1, M[8], B2); h2.z = dot3_nv(M[6], B3, M[7], B4, M[8], B5); h3.z = dot3_nv(M[6], B6, M[7], B7, M[8], B8);
            h0.x = __fsub_rn(__fsub_rn(-h1.x, h2.x), h3.x);
            h0.y = __fsub_rn(__fsub_rn(-h1.y, h2.y), h3.y);
            h0.z = __fsub_rn(__fsub_rn(-h1.z, h2.z), h3.z);
            h0.w = h1.w = h2.w = h3.w = 0.f;
            float4* Hs = reinterpret_cast<float4*>(smem + LOCAL_OFF_HS) + tid;
            Hs[0] = h0; Hs[TILE_T] = h1; Hs[2 * TILE_T] = h2; Hs[3 * TILE_T] = h3;
        }
        __syncthreads();   // H scratch complete

        // phase C: ordered gather per tile-local vertex -> partial-sum slot
        {
            const uint16_t* incOff = reinterpret_cast<const uint16_t*>(rec + offIO);
            const uint16_t* inc = reinterpret_cast<const uint16_t*>(rec + offI);
            const uint8_t* Hb = smem + LOCAL_OFF_HS;
            for (uint32_t l = tid; l < nLocal; l += TILE_T) {
                const uint32_t e0 = incOff[l], e1 = incOff[l + 1];
                float sx = 0.f, sy = 0.f, sz = 0.f;
                for (uint32_t e = e0; e < e1; ++e) {
                    const float4 h = *reinterpret_cast<const float4*>(Hb + inc[e]);
                    sx += h.x; sy += h.y; sz += h.z;
                }
                P[hdr.z + l] = make_float4(sx, sy, sz, 0.f);
            }
        }
        // no barrier here: the next iteration's first barrier (after its phase A, which touches
        // neither the H scratch nor this record buffer) orders phase C against every later write
    }
}

// ------------------------------------------------------------------ global step (Jacobi + Chebyshev)
// addM_h2Sn + computeDBCLocal + getErrorKern + chebyshevKern fused per vertex
// (pdUtil.cu:147-179,195-226).  Reads the ordered partial sums of the local step.
__global__ void k_vertex_jacobi(int nV, const float4* __restrict__ qcur, const float4* __restrict__ qprev,
                                float4* __restrict__ qnext, const float4* __restrict__ so4,
                                const float2* __restrict__ cc, const uint32_t* __restrict__ vslotPtr,
                                const uint32_t* __restrict__ vslot, const float4* __restrict__ P,
                                float omega, float wdbc)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 so = so4[v];
    const float2 c2 = cc[v];
    const float c = c2.x;
    float bx, by, bz;
    if (so.w > 0.f) {                 // computeDBCLocal overwrites b for pinned vertices
        bx = so.x * wdbc; by = so.y * wdbc; bz = so.z * wdbc;
    } else {
        bx = c * so.x; by = c * so.y; bz = c * so.z;
        const uint32_t e0 = vslotPtr[v], e1 = vslotPtr[v + 1];
        for (uint32_t e = e0; e < e1; ++e) {
            const float4 p = __ldg(&P[vslot[e]]);
            bx += p.x; by += p.y; bz += p.z;
        }
    }
    const float4 q = qcur[v], pr = qprev[v];
    const float den = c2.y;
    float nx = (bx - c * q.x) / den + q.x;
    float ny = (by - c * q.y) / den + q.y;
    float nz = (bz - c * q.z) / den + q.z;
    // under-relaxation in double, as the reference's `0.9 *` literal (pdUtil.cu:221)
    nx = (float)(0.9 * (double)(nx - q.x) + (double)q.x);
    ny = (float)(0.9 * (double)(ny - q.y) + (double)q.y);
    nz = (float)(0.9 * (double)(nz - q.z) + (double)q.z);
    nx = (nx - pr.x) * omega + pr.x;
    ny = (ny - pr.y) * omega + pr.y;
    nz = (nz - pr.z) * omega + pr.z;
    qnext[v] = make_float4(nx, ny, nz, 0.f);
}

// right-hand side for the direct / CG global solves: b = c*s_old + sum of partials (R, not R-F)
__global__ void k_vertex_rhs(int nV, const float4* __restrict__ so4, const float2* __restrict__ cc,
                             const uint32_t* __restrict__ vslotPtr, const uint32_t* __restrict__ vslot,
                             const float4* __restrict__ P, float wdbc, float4* __restrict__ rhs)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 so = so4[v];
    const float c = cc[v].x;
    float bx, by, bz;
    if (so.w > 0.f) {
        bx = so.x * wdbc; by = so.y * wdbc; bz = so.z * wdbc;
    } else {
        bx = c * so.x; by = c * so.y; bz = c * so.z;
        const uint32_t e0 = vslotPtr[v], e1 = vslotPtr[v + 1];
        for (uint32_t e = e0; e < e1; ++e) {
            const float4 p = __ldg(&P[vslot[e]]);
            bx += p.x; by += p.y; bz += p.z;
        }
    }
    rhs[v] = make_float4(bx, by, bz, 0.f);
}

// ------------------------------------------------------------------ end of step
// updateVelPos (pdUtil.cu:180-193), X <- XTilde (pdSolver.cu:227), then the three fixed-body
// kernels of FixedBodyData::HandleCollisions (fixedBodyData.cu:67-148) in their launch order.
__device__ __forceinline__ void fb_respond(float3& vel, const float3 n, float muT, float muN)
{
    const float vn = vel.x * n.x + vel.y * n.y + vel.z * n.z;
    const float3 vN = make_float3(vn * n.x, vn * n.y, vn * n.z);
    const float3 vT = make_float3(vel.x - vN.x, vel.y - vN.y, vel.z - vN.z);
    const float magT = sqrtf(vT.x * vT.x + vT.y * vT.y + vT.z * vT.z);
    const float magN = sqrtf(vN.x * vN.x + vN.y * vN.y + vN.z * vN.z);
    const float a = magT == 0.f ? 0.f : fmaxf(1.f - muT * (1.f + muN) * magN / magT, 0.f);
    vel.x = -muN * vN.x + a * vT.x;
    vel.y = -muN * vN.y + a * vT.y;
    vel.z = -muN * vN.z + a * vT.z;
}

__global__ void k_finish(int nV, const float4* __restrict__ qfinal, float dtInv, float4* __restrict__ X,
                         float4* __restrict__ XTilde, float4* __restrict__ V, DevFixedBodies fb, float muT, float muN)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 q = qfinal[v], xt = XTilde[v];
    float3 vel = make_float3((q.x - xt.x) * dtInv, (q.y - xt.y) * dtInv, (q.z - xt.z) * dtInv);
    float3 x = make_float3(q.x, q.y, q.z);
    X[v] = make_float4(x.x, x.y, x.z, 0.f);      // X keeps the un-projected position
    for (int j = 0; j < fb.nSpheres; ++j) {
        const float* s = fb.spheres + 4 * j;
        const float3 tc = make_float3(x.x - s[0], x.y - s[1], x.z - s[2]);
        const float d2 = tc.x * tc.x + tc.y * tc.y + tc.z * tc.z;
        const float d = sqrtf(d2);
        if (d < s[3]) {
            const float inv = 1.0f / sqrtf(d2);
            const float3 n = make_float3(tc.x * inv, tc.y * inv, tc.z * inv);
            const float push = s[3] - d;
            x.x += push * n.x; x.y += push * n.y; x.z += push * n.z;
            fb_respond(vel, n, muT, muN);
        }
    }
    for (int j = 0; j < fb.nPlanes; ++j) {
        const float* p = fb.planes + 6 * j;
        const float3 up = make_float3(p[3], p[4], p[5]);
        const float sd = (x.x - p[0]) * up.x + (x.y - p[1]) * up.y + (x.z - p[2]) * up.z;
        if (sd < 0.f && (vel.x * up.x + vel.y * up.y + vel.z * up.z) < 0.f) {
            x.x -= sd * up.x; x.y -= sd * up.y; x.z -= sd * up.z;
            fb_respond(vel, up, muT, muN);
        }
    }
    for (int j = 0; j < fb.nCyls; ++j) {
        const float* c = fb.cyls + 7 * j;
        const float3 ax = make_float3(c[3], c[4], c[5]);
        const float3 rel = make_float3(x.x - c[0], x.y - c[1], x.z - c[2]);
        // n = (I - a a^T) rel as the reference's matrix-vector product
        float3 nn;
        nn.x = (1.f - ax.x * ax.x) * rel.x + (0.f - ax.y * ax.x) * rel.y + (0.f - ax.z * ax.x) * rel.z;
        nn.y = (0.f - ax.x * ax.y) * rel.x + (1.f - ax.y * ax.y) * rel.y + (0.f - ax.z * ax.y) * rel.z;
        nn.z = (0.f - ax.x * ax.z) * rel.x + (0.f - ax.y * ax.z) * rel.y + (1.f - ax.z * ax.z) * rel.z;
        const float d2 = nn.x * nn.x + nn.y * nn.y + nn.z * nn.z;
        const float d = sqrtf(d2);
        if (d < c[6]) {
            const float inv = 1.0f / sqrtf(d2);
            const float3 n = make_float3(nn.x * inv, nn.y * inv, nn.z * inv);
            const float push = c[6] - d;
            x.x += push * n.x; x.y += push * n.y; x.z += push * n.z;
            fb_respond(vel, n, muT, muN);
        }
    }
    XTilde[v] = make_float4(x.x, x.y, x.z, 0.f);
    V[v] = make_float4(vel.x, vel.y, vel.z, 0.f);
}

// ------------------------------------------------------------------ layout conversion
// AoS float3 (reference numbering) <-> padded float4 (renumbered): newOfOld / oldOfNew permutations
__global__ void k_import3(int nV, const float* __restrict__ src3, const uint32_t* __restrict__ oldOfNew, float4* __restrict__ dst)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float* s = src3 + 3ull * oldOfNew[v];
    dst[v] = make_float4(s[0], s[1], s[2], 0.f);
}
__global__ void k_export3(int nV, const float4* __restrict__ src, const uint32_t* __restrict__ oldOfNew, float* __restrict__ dst3)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 s = src[v];
    float* d = dst3 + 3ull * oldOfNew[v];
    d[0] = s.x; d[1] = s.y; d[2] = s.z;
}

// ------------------------------------------------------------------ test hook
// corotation() on a batch of row-major 3x3 matrices (used by the parity tests of the rotation paths)
template <int ROT_MODE>
__global__ void k_rotation_batch(int n, const float* __restrict__ F, float* __restrict__ R, int* __restrict__ usedFast)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Mat3 f, r;
#pragma unroll
    for (int e = 0; e < 9; ++e) f.m[e] = F[9ull * i + e];
    bool fast = false;
    if (ROT_MODE == 0) fast = rotation_newton(f, r);
    if (!fast) rotation_svd(f, r);
#pragma unroll
    for (int e = 0; e < 9; ++e) R[9ull * i + e] = r.m[e];
    if (usedFast) usedFast[i] = fast ? 1 : 0;
}

}  // namespace pdb200
