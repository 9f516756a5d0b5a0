// One whole PdSolver::Update per launch for scenes made of many SMALL bodies -- BASELINE config 5, batches of independent
// house + sphere contexts.  Selected automatically (pd_engine_options::body_kernel, Engine ctor) when every connected body fits
// one CTA's shared memory; PD_BODY_KERNEL=0 keeps the tile kernels (A/B runs).  DESIGN.md section 4.
//
// Without mesh-mesh collision the soft bodies of a scene share no tet, so the PD system is block diagonal per body
//  and a body of a few thousand tets fits one SM: ONE CTA per body keeps the iterate, its
// predecessor, the right-hand-side base term and the per-corner contributions H in shared memory and runs predictor,
// `iterations` x (local step, Jacobi-Chebyshev sweep) and the end of step with two __syncthreads per iteration and no
// global traffic except the (L2-resident) tet records and incidence lists.  The tile path needs 2 + 2 * iterations
// launches per step, each far too small to fill the GPU (batch64: 166 k tets, 2.31 ms per step); this is one launch per step
// (1.43 ms, same-box A/B, profiles/r2_body_kernel_ab_batch64.txt).
//
// Arithmetic: the same device functions as the tile path (tet_contrib, finish_vertex) and the same forms as k_predict /
// k_vertex_jacobi.  The sum over a vertex's contributions runs sequentially over its incidence list in ascending
// (tet, corner) order of the ORIGINAL tet numbering, starting from (M/h^2) s_old -- the reference's order for a
// sequential scatter -- so with ROT_MODE 1 and matrix_diag summed in the same order the kernel matches the oracle bit for bit.
#pragma once
#include "pd_kernels.cuh"

namespace pdb200 {

// BodyDesc / BodyBatch: layout.hpp (built on the host by layout.cpp:build_body_batch)

// shared-memory carve-up for capacities (nVmax, nTmax, multiples of 32): q[3] | b0 | cc | H (four corner planes) | incidence
// pointers (nVmax + 32 words) | incidence entries (4 nTmax u16)
// pointers (nVmax + 32 words) | incidence entries (4 nTmax u16) | the tets the fast rotation path declined (nTmax u16) | their count
__host__ __device__ inline size_t body_smem_bytes(uint32_t nVmax, uint32_t nTmax) { return 72ull * nVmax + 64ull * nTmax + 4ull * (nVmax + 32u) + 8ull * nTmax + 2ull * nTmax + 16ull; }

template <int ROT_MODE>
__global__ void __launch_bounds__(512, 1)
k_body_step(const BodyDesc* __restrict__ bodies, const uint32_t* __restrict__ bverts, const uint8_t* __restrict__ brec,
            const uint32_t* __restrict__ bincPtr, const uint16_t* __restrict__ binc, const float* __restrict__ bmd,
            uint32_t nVmax, uint32_t nTmax, float4* __restrict__ X, float4* __restrict__ V, float4* __restrict__ XT,
            const float* __restrict__ mass, const float* __restrict__ dbc, const float4* __restrict__ dbcx,
            float dt, float dt2Prepared, float gravity, int iterations, float rho, float wdbc, DevFixedBodies fb, float muT, float muN,
            unsigned long long* __restrict__ perfNs)
{
    // perfNs (perf mode, else null): nanoseconds body 0 spent in {local step, global step, end of step}, added up over launches
    // -- the split PdSolver reports through GetPerformanceData (pdSolver.cu:26,165-203,228-231)
    const bool timed = perfNs != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long tLocal = 0, tGlobal = 0, tMark = 0;
    auto now = [&]() -> unsigned long long {
#ifdef PD_HOST_EMU
        return 0ull;
#else
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
#endif
    };
    PD_DYN_SMEM(smem);
    float4* const qb = reinterpret_cast<float4*>(smem);          // iterate buffer k at qb + k * nVmax
    float4* b0s = reinterpret_cast<float4*>(smem) + 3 * nVmax;
    float2* ccs = reinterpret_cast<float2*>(smem + 64ull * nVmax);
    float4* Hs = reinterpret_cast<float4*>(smem + 72ull * nVmax);
    uint32_t* ptrS = reinterpret_cast<uint32_t*>(smem + 72ull * nVmax + 64ull * nTmax);
    uint16_t* incS = reinterpret_cast<uint16_t*>(smem + 72ull * nVmax + 64ull * nTmax + 4ull * (nVmax + 32u));
    uint16_t* slowS = incS + 4ull * nTmax;                                   // tets whose contribution needs the slow rotation path
    unsigned int* nSlow = reinterpret_cast<unsigned int*>(slowS + nTmax);     // (nTmax is a multiple of 32: 4-byte aligned)
    const BodyDesc bd = bodies[blockIdx.x];
    const uint32_t tid = threadIdx.x, nth = blockDim.x;

    // the body's incidence lists stay in shared memory for the whole step (the gather below walks them every iteration)
    for (uint32_t l = tid; l <= bd.nV; l += nth) ptrS[l] = bincPtr[bd.ptr0 + l];
    for (uint32_t e = tid; e < 4u * bd.nT; e += nth) incS[e] = binc[bd.inc0 + e];
    // ---- predictor: gravity, setMDt_2MoreDBC (moreDBC = 0), computeSn, addM_h2Sn (k_predict<false>)
    const float dt2 = __fmul_rn(dt, dt);
    for (uint32_t l = tid; l < bd.nV; l += nth) {
        const uint32_t v = bverts[bd.v0 + l];
        const float m = mass[v], isDbc = dbc[v];
        const float c = (isDbc == 0.f) ? __fdiv_rn(m, dt2) : __fdiv_rn(__fadd_rn(m, __fmul_rn(isDbc, 1e6f)), dt2Prepared);
        const float4 x = X[v], vel = V[v];
        const float fy = __fmul_rn(-gravity, m);
        const float dt2_m_1 = __fdiv_rn(1.0f, c);
        float4 s;
        s.x = __fmaf_rn(0.0f, dt2_m_1, __fmaf_rn(vel.x, dt, x.x));
        s.y = __fmaf_rn(fy, dt2_m_1, __fmaf_rn(vel.y, dt, x.y));
        s.z = __fmaf_rn(0.0f, dt2_m_1, __fmaf_rn(vel.z, dt, x.z));
        s.w = 0.f;
        qb[l] = s;
        qb[2 * nVmax + l] = s;
        b0s[l] = make_float4(__fmul_rn(c, s.x), __fmul_rn(c, s.y), __fmul_rn(c, s.z), 0.f);
        ccs[l] = make_float2(isDbc > 0.f ? -c : c, __fadd_rn(c, bmd[bd.v0 + l]));
    }
    __syncthreads();

    const uint8_t* rec = brec + 16ull * bd.recOff16;
    if (tid == 0) *nSlow = 0u;
    __syncthreads();
    float omega = 1.0f;
    if (timed) tMark = now();
    for (int i = 0; i < iterations; ++i) {
        const float4* cur = qb + (uint32_t)(i % 3) * nVmax;
        const float4* prev = qb + (uint32_t)((i + 2) % 3) * nVmax;
        float4* next = qb + (uint32_t)((i + 1) % 3) * nVmax;
        // ---- local step: one tet per thread and trip; record = three coalesced 16-byte planes (DmInv, w, four u16 local ids)
        for (uint32_t t = tid; t < bd.nT; t += nth) {
            // (plain read-only loads: the records are re-read every iteration and should stay in L1 / L2)
            const float4* rp = reinterpret_cast<const float4*>(rec);
            const float4 r0 = __ldg(rp + t), r1 = __ldg(rp + bd.nT + t), r2 = __ldg(rp + 2ull * bd.nT + t);
            const uint32_t c01 = __float_as_uint(r2.z), c23 = __float_as_uint(r2.w);
            float4 h0, h1, h2, h3;
            if (ROT_MODE == 0) {
                // the fast rotation path only; a tet it declines (inverted / flat / strongly deformed: a few per cent of a body
                // that lies on the floor) is put on a list and taken by the slow path BELOW, compacted -- taken in place it drags
                // its whole warp through the ~900-instruction slow path (36 % of this kernel's instructions on batch64)
                if (!tet_contrib_fast<true>(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, cur[c01 & 0xffffu], cur[c01 >> 16],
                                            cur[c23 & 0xffffu], cur[c23 >> 16], h0, h1, h2, h3)) {
                    slowS[atomicAdd(nSlow, 1u)] = (uint16_t)t;
                    continue;
                }
            } else {
                tet_contrib_slow<ROT_MODE, true>(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, cur[c01 & 0xffffu], cur[c01 >> 16],
                                                 cur[c23 & 0xffffu], cur[c23 >> 16], h0, h1, h2, h3);
            }
            Hs[t] = h0; Hs[nTmax + t] = h1; Hs[2 * nTmax + t] = h2; Hs[3 * nTmax + t] = h3;
        }
        __syncthreads();
        if (ROT_MODE == 0) {
            // ... the declined tets, consecutive threads (the list's order varies from run to run, the results do not: every
            // tet writes its own four entries)
            const unsigned int ns = *nSlow;
            for (uint32_t i = tid; i < ns; i += nth) {
                const uint32_t t = slowS[i];
                const float4* rp = reinterpret_cast<const float4*>(rec);
                const float4 r0 = __ldg(rp + t), r1 = __ldg(rp + bd.nT + t), r2 = __ldg(rp + 2ull * bd.nT + t);
                const uint32_t c01 = __float_as_uint(r2.z), c23 = __float_as_uint(r2.w);
                float4 h0, h1, h2, h3;
                tet_contrib_slow<0, true>(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, cur[c01 & 0xffffu], cur[c01 >> 16],
                                          cur[c23 & 0xffffu], cur[c23 >> 16], h0, h1, h2, h3);
                Hs[t] = h0; Hs[nTmax + t] = h1; Hs[2 * nTmax + t] = h2; Hs[3 * nTmax + t] = h3;
            }
            __syncthreads();
            if (tid == 0) *nSlow = 0u;       // (the next iteration's first atomicAdd comes after the barrier that ends the sweep below)
        }
        if (timed) { const unsigned long long t = now(); tLocal += t - tMark; tMark = t; }
        // omega recurrence in float, pdSolver.cu:196-198 (the same operations Engine::enqueueIteration performs on the host)
        if (i <= 10) omega = 1.0f;
        else if (i == 11) omega = __fdiv_rn(2.0f, __fsub_rn(2.0f, __fmul_rn(rho, rho)));
        else omega = __fdiv_rn(4.0f, __fsub_rn(4.0f, __fmul_rn(__fmul_rn(rho, rho), omega)));
        // ---- right-hand side + Jacobi sweep + Chebyshev (addM_h2Sn, the scatter of computeLocal as an ordered gather,
        //      computeDBCLocal, getErrorKern, chebyshevKern: k_vertex_jacobi<false>)
        for (uint32_t l = tid; l < bd.nV; l += nth) {
            const float2 c2 = ccs[l];
            const float c = fabsf(c2.x);
            float bx, by, bz;
            if (c2.x < 0.f) {
                const float4 d = dbcx[bverts[bd.v0 + l]];
                bx = __fmul_rn(d.x, wdbc); by = __fmul_rn(d.y, wdbc); bz = __fmul_rn(d.z, wdbc);
            } else {
                const float4 bb = b0s[l];
                bx = bb.x; by = bb.y; bz = bb.z;
                uint32_t e = ptrS[l];
                const uint32_t e1 = ptrS[l + 1];
                auto hs = [&](uint32_t code) -> float4 { return Hs[(code & 3u) * nTmax + (code >> 2)]; };       // code = tet * 4 + corner
                for (; e + 4u <= e1; e += 4u) {      // four gathers in flight, added strictly in list order
                    const float4 h0 = hs(incS[e]), h1 = hs(incS[e + 1u]), h2 = hs(incS[e + 2u]), h3 = hs(incS[e + 3u]);
                    bx = __fadd_rn(bx, h0.x); by = __fadd_rn(by, h0.y); bz = __fadd_rn(bz, h0.z);
                    bx = __fadd_rn(bx, h1.x); by = __fadd_rn(by, h1.y); bz = __fadd_rn(bz, h1.z);
                    bx = __fadd_rn(bx, h2.x); by = __fadd_rn(by, h2.y); bz = __fadd_rn(bz, h2.z);
                    bx = __fadd_rn(bx, h3.x); by = __fadd_rn(by, h3.y); bz = __fadd_rn(bz, h3.z);
                }
                for (; e < e1; ++e) {
                    const float4 h = hs(incS[e]);
                    bx = __fadd_rn(bx, h.x); by = __fadd_rn(by, h.y); bz = __fadd_rn(bz, h.z);
                }
            }
            const float4 qq = cur[l], pr = prev[l];
            const float den = c2.y;
            float nx = __fadd_rn(__fdiv_rn(__fmaf_rn(-c, qq.x, bx), den), qq.x);
            float ny = __fadd_rn(__fdiv_rn(__fmaf_rn(-c, qq.y, by), den), qq.y);
            float nz = __fadd_rn(__fdiv_rn(__fmaf_rn(-c, qq.z, bz), den), qq.z);
            nx = (float)__fma_rn((double)__fsub_rn(nx, qq.x), 0.9, (double)qq.x);
            ny = (float)__fma_rn((double)__fsub_rn(ny, qq.y), 0.9, (double)qq.y);
            nz = (float)__fma_rn((double)__fsub_rn(nz, qq.z), 0.9, (double)qq.z);
            nx = __fmaf_rn(__fsub_rn(nx, pr.x), omega, pr.x);
            ny = __fmaf_rn(__fsub_rn(ny, pr.y), omega, pr.y);
            nz = __fmaf_rn(__fsub_rn(nz, pr.z), omega, pr.z);
            next[l] = make_float4(nx, ny, nz, 0.f);
        }
        __syncthreads();
        if (timed) { const unsigned long long t = now(); tGlobal += t - tMark; tMark = t; }
    }
    // ---- end of step: updateVelPos, X <- XTilde, fixed bodies (k_finish<false>)
    const float dtInv = __fdiv_rn(1.0f, dt);
    const float4* qf = qb + (uint32_t)(iterations % 3) * nVmax;
    for (uint32_t l = tid; l < bd.nV; l += nth) {
        const uint32_t v = bverts[bd.v0 + l];
        finish_vertex(qf[l], XT[v], dtInv, false, fb, muT, muN, &X[v], &XT[v], &V[v]);
    }
    if (timed) {
        __syncthreads();
        perfNs[0] += tLocal; perfNs[1] += tGlobal; perfNs[2] += now() - tMark;
    }
}

}  // namespace pdb200
