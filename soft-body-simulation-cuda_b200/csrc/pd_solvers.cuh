// Global-step back-ends other than Chebyshev-Jacobi: Jacobi-preconditioned CG and the prefactored sparse
// Cholesky solve, both on the SCALAR system matrix A^ (nV x nV) with the three coordinate right-hand sides
// interleaved as float4 -- the reference's 3nV x 3nV matrix is A^ (x) I3 (pdUtil.cu:26-37 writes the same K_ji to
// the x, y and z rows), so one pass over A^ serves all three.
//   PCG      : PCGJacobiSolver<float>::Solve, src/simulation/solver/linear/pcgJacobi.cu:88-172
//   Cholesky : CholeskySpLinearSolver<float>, src/simulation/solver/linear/cholesky.cu:133-192 (cuSOLVER) and
//              Eigen::SimplicialCholesky, src/simulation/solver/projective/pdSolver.cu:103,181-183
// Both run as ONE persistent cooperative kernel per solve: no host round trips for the scalars (the reference
// reads three of them back per CG iteration), no library calls.
#pragma once
#include <cstdint>
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace pdb200 {
namespace cg = cooperative_groups;

// device-resident bookkeeping of the non-Jacobi modes (one per engine)
struct SolveState {
    float err;               // computeError (pdSolver.cu:243-253): mean squared change of the iterate
    int done;                // sqrt(err) < tol reached: the remaining PD iterations of this step are skipped
    int pdIters;             // PD iterations executed (this step)
    long long innerIters;    // CG iterations executed (accumulated)
    long long pdItersTotal;
};

struct CsrDev {
    int n;
    const int* rowPtr;
    const int* col;
    const float* val;
    const float* invDiag;    // ExtractInverseDiagonalKernel, pcgJacobi.cu:6-19
};

#ifndef PD_SOLVE_THREADS
#define PD_SOLVE_THREADS 256
#endif
constexpr int SOLVE_THREADS = PD_SOLVE_THREADS;
constexpr int SOLVE_WARPS = SOLVE_THREADS / 32;       // shared scratch of a reduction: 3 * SOLVE_WARPS + 3 doubles
constexpr int SOLVE_SH = 3 * SOLVE_WARPS + 3;
constexpr int SOLVE_MAX_PARTIALS = 2048;       // >= grid size of the cooperative kernels

// Deterministic grid-wide sum of up to three doubles per thread: warp shuffle -> block -> one slot per block ->
// grid.sync -> every block adds the slots in the same fixed order.  Returns the sums in out[0..2].
// ONE grid-wide synchronisation per reduction: consecutive reductions alternate between two sets of slots (`flip`, a
// per-thread counter every thread of the grid advances alike), and a block can only reach the synchronisation of reduction
// k + 1 after it has read the slots of reduction k, so set k & 1 is free again when reduction k + 2 writes it.
__device__ __forceinline__ void grid_sum3(cg::grid_group& grid, double a, double b, double c, double* partials /* 2 x 3 * SOLVE_MAX_PARTIALS */,
                                          double* sh /* 3 * 8 + 3 doubles of shared memory */, double out[3], unsigned& flip)
{
    partials += (flip++ & 1u) * 3u * (unsigned)SOLVE_MAX_PARTIALS;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[warp] = a; sh[SOLVE_WARPS + warp] = b; sh[2 * SOLVE_WARPS + warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0, sc = 0;
        for (int w = 0; w < SOLVE_WARPS; ++w) { sa += sh[w]; sb += sh[SOLVE_WARPS + w]; sc += sh[2 * SOLVE_WARPS + w]; }
        partials[3 * blockIdx.x] = sa; partials[3 * blockIdx.x + 1] = sb; partials[3 * blockIdx.x + 2] = sc;
    }
    grid.sync();
    if (warp == 0) {
        double sa = 0, sb = 0, sc = 0;
        for (int i = lane; i < (int)gridDim.x; i += 32) { sa += partials[3 * i]; sb += partials[3 * i + 1]; sc += partials[3 * i + 2]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
            sc += __shfl_xor_sync(0xffffffffu, sc, o);
        }
        if (lane == 0) { sh[3 * SOLVE_WARPS] = sa; sh[3 * SOLVE_WARPS + 1] = sb; sh[3 * SOLVE_WARPS + 2] = sc; }
    }
    __syncthreads();
    out[0] = sh[3 * SOLVE_WARPS]; out[1] = sh[3 * SOLVE_WARPS + 1]; out[2] = sh[3 * SOLVE_WARPS + 2];
    __syncthreads();      // sh is reused by the next reduction
}

// ------------------------------------------------------------------ multi-GPU PCG
// The rows of A^ are partitioned like the vertices; every rank runs the same cooperative kernel on its rows.
// Two exchanges per CG iteration go over NVLink peer memory from INSIDE the kernel (no NCCL call, no host):
//  * the halo of p: boundary entries are stored straight into the neighbours' ghost entries of p, then a flag per
//    neighbour (sequence number) is raised; the SpMV starts once every neighbour's flag has arrived;
//  * the dot products (p.Ap | r.r, r.z | the PD error): an all-gather of the ranks' partial sums (three doubles +
//    a sequence number per rank, double-buffered by sequence parity) that every rank adds up in RANK ORDER, so all
//    ranks hold bit-identical scalars and take identical branches.
// Slot reuse is safe with two parities: a rank can only publish reduction k+2 after it has seen every rank's
// reduction k+1, which those ranks publish after all their blocks have consumed reduction k (grid.sync in grid_sum3).
struct DistSolve {
    int world, rank, nNbr, nPush, nGlobal;
    const int* nbr;                               // neighbour ranks
    const uint32_t *pushSrc, *pushDst, *pushNbr;  // the engine's push list (owned vertex -> ghost entry at neighbour slot)
    float4* const* peerP;                         // [nNbr]: the neighbours' p vectors
    unsigned long long* const* peerPFlag;         // [nNbr]: this rank's entry in the neighbours' p-flag arrays
    const unsigned long long* pflags;             // this rank's p-flag array, written by the peers (indexed by rank)
    double* const* peerRed;                       // [world]: every rank's reduction slots (own included)
    const double* red;                            // this rank's reduction slots: [parity][rank][a, b, c, seq]
    unsigned long long* seq;                      // device-resident counters: [0] p pushes, [1] reductions
    unsigned int* status;                         // set to 1 when a wait gave up
};
constexpr long long SOLVE_WAIT_LIMIT_CYCLES = 20000000000ll;

__device__ __forceinline__ unsigned long long solve_ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void solve_st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double solve_ld_relaxed_sys(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// in/out: the rank-local sums (identical in every thread of the grid, from grid_sum3) -> the sums over all ranks
__device__ __forceinline__ void dist_allreduce3(const DistSolve& D, unsigned long long seq, double* sh, double v[3])
{
    if (blockIdx.x == 0 && (int)threadIdx.x < D.world) {
        double* slot = D.peerRed[threadIdx.x] + ((seq & 1ull) * (unsigned)D.world + (unsigned)D.rank) * 4u;
        slot[0] = v[0]; slot[1] = v[1]; slot[2] = v[2];
        __threadfence_system();
        solve_st_release_sys(reinterpret_cast<unsigned long long*>(slot + 3), seq);
    }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double a = 0, b = 0, c = 0;
        if (lane < D.world) {
            const double* slot = D.red + ((seq & 1ull) * (unsigned)D.world + (unsigned)lane) * 4u;
            const long long t0 = clock64();
            while (solve_ld_acquire_sys(reinterpret_cast<const unsigned long long*>(slot + 3)) != seq) {
                if (clock64() - t0 > SOLVE_WAIT_LIMIT_CYCLES) { atomicExch(D.status, 1u); break; }
            }
            a = solve_ld_relaxed_sys(slot); b = solve_ld_relaxed_sys(slot + 1); c = solve_ld_relaxed_sys(slot + 2);
        }
        double sa = 0, sb = 0, sc = 0;
        for (int r = 0; r < D.world; ++r) {          // rank order: every rank gets the same bits
            sa += __shfl_sync(0xffffffffu, a, r); sb += __shfl_sync(0xffffffffu, b, r); sc += __shfl_sync(0xffffffffu, c, r);
        }
        if (lane == 0) { sh[3 * SOLVE_WARPS] = sa; sh[3 * SOLVE_WARPS + 1] = sb; sh[3 * SOLVE_WARPS + 2] = sc; }
    }
    __syncthreads();
    v[0] = sh[3 * SOLVE_WARPS]; v[1] = sh[3 * SOLVE_WARPS + 1]; v[2] = sh[3 * SOLVE_WARPS + 2];
    __syncthreads();
}

template <bool DIST>
__device__ __forceinline__ void solve_sum3(cg::grid_group& grid, const DistSolve& D, unsigned long long& rSeq, double a, double b, double c,
                                           double* partials, double* sh, double out[3], unsigned& flip)
{
    grid_sum3(grid, a, b, c, partials, sh, out, flip);
    if (DIST) { ++rSeq; dist_allreduce3(D, rSeq, sh, out); }
}

// boundary entries of p -> the neighbours' ghost entries, flags up, wait for the neighbours' (call after a grid.sync
// that completed p); returns with every ghost entry of p current
__device__ __forceinline__ void dist_push_p(cg::grid_group& grid, const DistSolve& D, unsigned long long& pSeq, const float4* p)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    for (int i = gtid; i < D.nPush; i += gstride) D.peerP[D.pushNbr[i]][D.pushDst[i]] = p[D.pushSrc[i]];
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();      // one per block, cumulative over the block's remote stores
    grid.sync();
    ++pSeq;
    if (blockIdx.x == 0 && (int)threadIdx.x < D.nNbr) solve_st_release_sys(D.peerPFlag[threadIdx.x], pSeq);
    if ((int)threadIdx.x < D.nNbr) {
        const unsigned long long* f = D.pflags + D.nbr[threadIdx.x];
        const long long t0 = clock64();
        while (solve_ld_acquire_sys(f) < pSeq) {
            if (clock64() - t0 > SOLVE_WAIT_LIMIT_CYCLES) { atomicExch(D.status, 1u); break; }
        }
    }
    __syncthreads();
}

// row v of y = A^ x for the three interleaved right-hand sides (CSR, ascending columns, one thread per row)
__device__ __forceinline__ float4 spmv_row(const CsrDev& A, int v, const float4* x)
{
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    const int e1 = A.rowPtr[v + 1];
    int e = A.rowPtr[v];
    for (; e + 4 <= e1; e += 4) {       // four gathers in flight; the sum keeps its ascending-column order
        const int c0 = A.col[e], c1 = A.col[e + 1], c2 = A.col[e + 2], c3 = A.col[e + 3];
        const float4 x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
        const float v0 = A.val[e], v1 = A.val[e + 1], v2 = A.val[e + 2], v3 = A.val[e + 3];
        a0 = __fadd_rn(a0, __fmul_rn(v0, x0.x)); a1 = __fadd_rn(a1, __fmul_rn(v0, x0.y)); a2 = __fadd_rn(a2, __fmul_rn(v0, x0.z));
        a0 = __fadd_rn(a0, __fmul_rn(v1, x1.x)); a1 = __fadd_rn(a1, __fmul_rn(v1, x1.y)); a2 = __fadd_rn(a2, __fmul_rn(v1, x1.z));
        a0 = __fadd_rn(a0, __fmul_rn(v2, x2.x)); a1 = __fadd_rn(a1, __fmul_rn(v2, x2.y)); a2 = __fadd_rn(a2, __fmul_rn(v2, x2.z));
        a0 = __fadd_rn(a0, __fmul_rn(v3, x3.x)); a1 = __fadd_rn(a1, __fmul_rn(v3, x3.y)); a2 = __fadd_rn(a2, __fmul_rn(v3, x3.z));
    }
    for (; e < e1; ++e) {
        const float a = A.val[e];
        const float4 xc = x[A.col[e]];
        a0 = __fadd_rn(a0, __fmul_rn(a, xc.x)); a1 = __fadd_rn(a1, __fmul_rn(a, xc.y)); a2 = __fadd_rn(a2, __fmul_rn(a, xc.z));
    }
    return make_float4(a0, a1, a2, 0.f);
}

// computeError + prev = x, fused into the tail of both solve kernels (pdSolver.cu:186-192,243-253)
template <bool DIST>
__device__ __forceinline__ void finish_pd_iteration(cg::grid_group& grid, int n, const float4* x, float4* xprev, float tol, SolveState* st,
                                                    int innerIters, double* partials, double* sh, const DistSolve& D, unsigned long long& rSeq, unsigned& flip)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    double acc = 0;
    for (int v = gtid; v < n; v += gstride) {
        const float4 a = xprev[v], b = x[v];
        const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y, dz = (double)a.z - (double)b.z;
        acc += dx * dx + dy * dy + dz * dz;
        xprev[v] = b;
    }
    double s[3];
    solve_sum3<DIST>(grid, D, rSeq, acc, 0.0, 0.0, partials, sh, s, flip);
    if (gtid == 0) {
        const float err = (float)(s[0] / (3.0 * (double)(DIST ? D.nGlobal : n)));
        st->err = err;
        st->pdIters += 1; st->pdItersTotal += 1; st->innerIters += innerIters;
        if (!(sqrtf(err) >= tol)) st->done = 1;
    }
}

// ------------------------------------------------------------------ Jacobi-preconditioned CG
// One launch = one PCGJacobiSolver::Solve with d_guess = x (warm start), pcgJacobi.cu:88-172:
//   r = b - A x;  loop k < max_iter: ||r||_2 < tol -> stop; z = D^-1 r; rho = r.z (|rho| < 1e-15 -> stop);
//   p = z + (rho/rho_prev) p; q = A p; alpha = rho / (p.q); x += alpha p; r -= alpha q.
// Dots accumulate in double and are rounded to float where the reference holds a float (cublasSdot results);
// vector updates use the same unfused multiply-then-add as the oracle (oracle/pd_oracle.c:pcg_solve).
// Three grid-wide synchronisations per CG iteration (p update | SpMV + p.q | x, r update + r.r, r.z): one per reduction
// (alternating partial slots) and one between the direction update and the product.  Measured and rejected on a B200:
// recomputing p = z + beta p_old inside the product (one synchronisation fewer, three gathers per non-zero instead of one:
// 30.0 vs 30.4 us per CG iteration on the 55^3 grid -- the gathers' L2 traffic is what bounds the product).
// DIST: A holds this rank's rows (n = owned vertices; columns index the rank's local vertex array, ghosts included);
// x and p carry ghost entries -- x's are current on entry (the engine's position halo), p's are exchanged here.
template <bool DIST>
__global__ void __launch_bounds__(SOLVE_THREADS)
k_pcg_solve(CsrDev A, const float4* __restrict__ b, float4* x, float4* __restrict__ r, float4* p, float4* __restrict__ q,
            float4* __restrict__ xprev, int maxIter, float cgTol, float pdTol, SolveState* st, double* partials, DistSolve D)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[SOLVE_SH];
    if (st->done) return;                       // uniform over the grid (and over the ranks): this PD iteration is skipped
    const int n = A.n, gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    unsigned long long pSeq = 0, rSeq = 0;
    if (DIST) { pSeq = D.seq[0]; rSeq = D.seq[1]; }
    unsigned flip = 0;
    double s[3];
    // r = b - A x ; rr = r.r ; rz = r.(D^-1 r)
    double arr = 0, arz = 0;
    for (int v = gtid; v < n; v += gstride) {
        const float4 ax = spmv_row(A, v, x), bb = b[v];
        const float4 rv = make_float4(__fsub_rn(bb.x, ax.x), __fsub_rn(bb.y, ax.y), __fsub_rn(bb.z, ax.z), 0.f);
        r[v] = rv;
        const float id = A.invDiag[v];
        arr += (double)rv.x * rv.x + (double)rv.y * rv.y + (double)rv.z * rv.z;
        arz += (double)rv.x * (double)__fmul_rn(rv.x, id) + (double)rv.y * (double)__fmul_rn(rv.y, id) + (double)rv.z * (double)__fmul_rn(rv.z, id);
    }
    solve_sum3<DIST>(grid, D, rSeq, arr, arz, 0.0, partials, sh, s, flip);
    float rho = (float)s[1], rhoPrev = 0.f;
    float rn = sqrtf((float)s[0]);
    int k = 0;
    for (; k < maxIter; ++k) {
        if (rn < cgTol) break;
        if (fabsf(rho) < 1e-15f) break;
        const float beta = (k == 0) ? 0.f : __fdiv_rn(rho, rhoPrev);
        double apq = 0;
        {
            for (int v = gtid; v < n; v += gstride) {       // p = z + beta p   (k == 0: p = z)
                const float4 rv = r[v];
                const float id = A.invDiag[v];
                const float zx = __fmul_rn(rv.x, id), zy = __fmul_rn(rv.y, id), zz = __fmul_rn(rv.z, id);
                float4 pv = make_float4(zx, zy, zz, 0.f);
                if (k > 0) {
                    const float4 po = p[v];
                    pv.x = __fadd_rn(__fmul_rn(beta, po.x), zx); pv.y = __fadd_rn(__fmul_rn(beta, po.y), zy); pv.z = __fadd_rn(__fmul_rn(beta, po.z), zz);
                }
                p[v] = pv;
            }
            grid.sync();
            if (DIST) dist_push_p(grid, D, pSeq, p);
            for (int v = gtid; v < n; v += gstride) {       // q = A p ; p.q
                const float4 qv = spmv_row(A, v, p), pv = p[v];
                q[v] = qv;
                apq += (double)pv.x * qv.x + (double)pv.y * qv.y + (double)pv.z * qv.z;
            }
        }
        solve_sum3<DIST>(grid, D, rSeq, apq, 0.0, 0.0, partials, sh, s, flip);
        const float alpha = __fdiv_rn(rho, (float)s[0]);
        arr = 0; arz = 0;
        const float4* pCur = p;
        for (int v = gtid; v < n; v += gstride) {       // x += alpha p ; r -= alpha q ; r.r ; r.z
            const float4 pv = pCur[v], qv = q[v];
            float4 xv = x[v], rv = r[v];
            xv.x = __fadd_rn(xv.x, __fmul_rn(alpha, pv.x)); xv.y = __fadd_rn(xv.y, __fmul_rn(alpha, pv.y)); xv.z = __fadd_rn(xv.z, __fmul_rn(alpha, pv.z));
            rv.x = __fsub_rn(rv.x, __fmul_rn(alpha, qv.x)); rv.y = __fsub_rn(rv.y, __fmul_rn(alpha, qv.y)); rv.z = __fsub_rn(rv.z, __fmul_rn(alpha, qv.z));
            x[v] = xv; r[v] = rv;
            const float id = A.invDiag[v];
            arr += (double)rv.x * rv.x + (double)rv.y * rv.y + (double)rv.z * rv.z;
            arz += (double)rv.x * (double)__fmul_rn(rv.x, id) + (double)rv.y * (double)__fmul_rn(rv.y, id) + (double)rv.z * (double)__fmul_rn(rv.z, id);
        }
        solve_sum3<DIST>(grid, D, rSeq, arr, arz, 0.0, partials, sh, s, flip);
        rhoPrev = rho;
        rho = (float)s[1];
        rn = sqrtf((float)s[0]);
    }
    finish_pd_iteration<DIST>(grid, n, x, xprev, pdTol, st, k, partials, sh, D, rSeq, flip);
    if (DIST && gtid == 0) { D.seq[0] = pSeq; D.seq[1] = rSeq; }
}

// ------------------------------------------------------------------ prefactored sparse Cholesky solve
// P A^ P^T = L L^T was factored once on the host (pd_engine.cu:prepareSolver: geometric nested-dissection ordering, then
// layout.cpp:cholesky_factor; the reference does the same inside SolverPrepare with cusolverSpXcsrcholAnalysis / Factor or
// Eigen::SimplicialCholesky, cholesky.cu:72-158, pdSolver.cu:103).  Per PD iteration: L y = P b, L^T z = y, x = P^T z for the
// three right-hand sides, as a "synchronisation-free" sparse triangular solve: ONE WARP PER ROW -- the lanes split the row's
// off-diagonal entries (the separator rows of the dissection hold hundreds), each lane waits for ITS dependencies (a tag in the
// fourth component of every solved entry, read and written with the value as one 128-bit word: no fences, one L2 round trip
// per hand-over), the partial sums are combined by a fixed shuffle tree.  Warp w takes rows w, w + W, ... in
// ascending (descending for L^T) order and the kernel is launched cooperatively (every warp resident), so the smallest
// unfinished row always has all its dependencies done and its warp working on it: progress is guaranteed.
// L is stored by rows (CSR, diagonal last) and L^T by rows as well (CSR, diagonal first).
struct CholDev {
    int n;
    const int *lPtr, *lCol; const float* lVal;        // L   by rows, ascending columns, diagonal LAST
    const int *uPtr, *uCol; const float* uVal;        // L^T by rows, ascending columns, diagonal FIRST
    const int* perm;                                  // perm[row of the factor] = vertex (engine numbering)
};

// A solved entry and its "final" tag travel in ONE 128-bit word (x, y, z, tag): b128 accesses are single-copy atomic, so
// value and flag need no fence between them and a consumer gets both with one L2 round trip.
__device__ __forceinline__ float4 ld_relaxed128(const float4* p)
{
    float4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed128(float4* p, float4 v)
{
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum over the off-diagonal entries [e0, e1) of one factor row of -l_e * v[col_e], lanes striding the entries; v[j] is final
// once its fourth component carries `tag`
__device__ __forceinline__ void chol_row_dot(const int* __restrict__ col, const float* __restrict__ val, int e0, int e1, int lane,
                                             const float4* v, int tag, float& a0, float& a1, float& a2)
{
    a0 = a1 = a2 = 0.f;
    // eight gathers in flight per lane (a separator row of the dissection holds thousands of entries; one L2 round trip per
    // entry and lane, taken one after the other, was 6 us per row on the critical path); only an entry that is not final
    // yet is polled again.  The sum keeps its order: ascending entries per lane.
    constexpr int D = 8;
    for (int e = e0 + lane; e < e1; e += 32 * D) {
        int j[D]; float l[D]; float4 vj[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int ek = e + 32 * k;
            j[k] = (ek < e1) ? col[ek] : -1;
            l[k] = (ek < e1) ? val[ek] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) if (j[k] >= 0) vj[k] = ld_relaxed128(v + j[k]);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (j[k] < 0) continue;
            for (unsigned spins = 0; __float_as_int(vj[k].w) != tag; ++spins) {
                if (spins > 4) __nanosleep(40);          // thousands of lanes polling back to back would congest L2 for the one load that matters
                vj[k] = ld_relaxed128(v + j[k]);
            }
            a0 = __fmaf_rn(-l[k], vj[k].x, a0); a1 = __fmaf_rn(-l[k], vj[k].y, a1); a2 = __fmaf_rn(-l[k], vj[k].z, a2);
        }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
}

__global__ void __launch_bounds__(SOLVE_THREADS)
k_chol_solve(CholDev C, const float4* __restrict__ b, float4* x, float4* y, float4* z, float4* __restrict__ xprev,
             int epochTag, float pdTol, SolveState* st, double* partials)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[SOLVE_SH];
    if (st->done) return;
    const int n = C.n, lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    const int tag = epochTag + 1;          // never 0 (the buffers start zeroed), one value per launch
    // forward: row i needs y[j] for every off-diagonal L_ij
    for (int i = warp; i < n; i += nWarps) {
        const int e0 = C.lPtr[i], e1 = C.lPtr[i + 1] - 1;
        // (everything the row's last instructions need is fetched BEFORE the dependencies are waited for: the hand-over from
        // row to row is the critical path of the whole solve)
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        float d = 1.f;
        if (lane == 0) { bb = b[C.perm[i]]; d = C.lVal[e1]; }
        float a0, a1, a2;
        chol_row_dot(C.lCol, C.lVal, e0, e1, lane, y, tag, a0, a1, a2);
        if (lane == 0) {
            st_relaxed128(y + i, make_float4(__fdiv_rn(__fadd_rn(bb.x, a0), d), __fdiv_rn(__fadd_rn(bb.y, a1), d), __fdiv_rn(__fadd_rn(bb.z, a2), d), __int_as_float(tag)));
        }
    }
    // (no grid-wide synchronisation here: the backward rows wait for y[i] through its tag like for any other dependency)
    // backward: row i of L^T needs z[j], j > i; rows are taken in descending order
    for (int t = warp; t < n; t += nWarps) {
        const int i = n - 1 - t;
        const int e0 = C.uPtr[i], e1 = C.uPtr[i + 1];
        float4 yy = make_float4(0.f, 0.f, 0.f, 0.f);
        float d = 1.f; int pi = 0;
        if (lane == 0) {
            d = C.uVal[e0]; pi = C.perm[i];
            do { yy = ld_relaxed128(y + i); } while (__float_as_int(yy.w) != tag);
        }
        float a0, a1, a2;
        chol_row_dot(C.uCol, C.uVal, e0 + 1, e1, lane, z, tag, a0, a1, a2);
        if (lane == 0) {
            const float4 r = make_float4(__fdiv_rn(__fadd_rn(yy.x, a0), d), __fdiv_rn(__fadd_rn(yy.y, a1), d), __fdiv_rn(__fadd_rn(yy.z, a2), d), __int_as_float(tag));
            st_relaxed128(z + i, r);
            x[pi] = make_float4(r.x, r.y, r.z, 0.f);
        }
    }
    grid.sync();
    unsigned long long noSeq = 0;
    unsigned flip = 0;
    finish_pd_iteration<false>(grid, n, x, xprev, pdTol, st, 0, partials, sh, DistSolve{}, noSeq, flip);
}

// start of a step in the non-Jacobi modes: err = 1, nothing skipped (pdSolver.cu:163)
__global__ void k_solve_begin(SolveState* st)
{
    st->err = 1.0f; st->done = 0; st->pdIters = 0;
}

}  // namespace pdb200
