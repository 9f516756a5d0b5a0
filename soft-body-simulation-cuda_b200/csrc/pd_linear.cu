// Linear back-ends of the reference's IPC (double) solver on the engine's fused CSR / CG kernels:
//   LinearSolver<double>::Solve(N, b, x, A, nz, rowIdx, colIdx, guess)            src/simulation/solver/linear/linear.h:55-71
//   PCGJacobiSolver<double> (Jacobi-preconditioned CG, max 2000, ||r|| < 1e-5)    linear/pcgJacobi.cu:88-172
//   CGSolver<double> (IC(0)-preconditioned CG, max 100, ||r|| < 1e-6)             linear/cg.cu:81-247
// as IPCSolver::SearchDirection calls them on the 3 nV x 3 nV Hessian in COO form with duplicates (IPC/ipc.cu:233-241).
// Everything runs on the device: COO -> CSR (sorted, duplicates summed in input order: the reference's sort_coo + coo2csr,
// linear.cu:7-39, without thrust), then ONE cooperative kernel per solve -- SpMV + direction update + dot products with the
// scalars on the device (the reference: cuSPARSE SpMV, cuBLAS dots, three host read-backs per iteration).  IC(0): a warp per
// row factorisation and triangular solves, rows handing over through (value, tag) pairs in one 128-bit word, as in the PD
// engine's Cholesky solve (pd_solvers.cuh).  No cuSPARSE / cuBLAS / thrust.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "pd_linear.hpp"

namespace cg = cooperative_groups;

namespace pdb200 {

#define LIN_CHECK(call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorName(e__) + " (" +          \
                                     cudaGetErrorString(e__) + ") at " + __FILE__ + ":" +                 \
                                     std::to_string(__LINE__) + ": " #call);                              \
    } while (0)

constexpr int LIN_THREADS = 256;
constexpr int LIN_MAX_BLOCKS = 2048;

// ------------------------------------------------------------------ COO -> CSR
__global__ void k_lin_count(int nz, const int* __restrict__ row, int N, int* __restrict__ cnt, int* __restrict__ bad)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nz) return;
    const int r = row[e];
    if (r < 0 || r >= N) { atomicOr(bad, 1); return; }
    atomicAdd(&cnt[r], 1);
}
// exclusive scan of cnt[0..N) into ptr[0..N] by ONE block (N is a few 10^5: tens of microseconds)
__global__ void k_lin_scan(int N, const int* __restrict__ cnt, int* __restrict__ ptr)
{
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < N ? cnt[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < N) ptr[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) ptr[N] = carry;
}
__global__ void k_lin_scatter(int nz, const int* __restrict__ row, int N, const int* __restrict__ ptr, int* __restrict__ cursor, int* __restrict__ slotEntry)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nz) return;
    const int r = row[e];
    if (r < 0 || r >= N) return;
    slotEntry[ptr[r] + atomicAdd(&cursor[r], 1)] = e;          // (any order: the row is sorted next)
}
// one thread per row: order the row's entries by (column, input index) -- the input index makes the duplicates' sum, and with
// it every bit of the solve, independent of the scatter's atomic order -- and merge duplicates; writes the compacted row
// in place and its length to len[r]
__global__ void k_lin_sort_rows(int N, const int* __restrict__ ptr, int* __restrict__ slotEntry, const int* __restrict__ col, const double* __restrict__ val,
                                int* __restrict__ ccol, double* __restrict__ cval, int* __restrict__ len, int* __restrict__ bad)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int b = ptr[r], e = ptr[r + 1];
    for (int i = b + 1; i < e; ++i) {              // insertion sort on (col, entry)
        const int x = slotEntry[i], cx = col[x];
        int j = i - 1;
        while (j >= b) {
            const int y = slotEntry[j], cy = col[y];
            if (cy < cx || (cy == cx && y < x)) break;
            slotEntry[j + 1] = y; --j;
        }
        slotEntry[j + 1] = x;
    }
    int o = b;
    for (int i = b; i < e;) {
        const int c = col[slotEntry[i]];
        if (c < 0 || c >= N) atomicOr(bad, 1);
        double s = 0.0;
        while (i < e && col[slotEntry[i]] == c) { s += val[slotEntry[i]]; ++i; }
        ccol[o] = c; cval[o] = s; ++o;
    }
    len[r] = o - b;
}
__global__ void k_lin_compact(int N, const int* __restrict__ ptr, const int* __restrict__ newPtr, const int* __restrict__ ccol, const double* __restrict__ cval,
                              int* __restrict__ ocol, double* __restrict__ oval)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int n = newPtr[r + 1] - newPtr[r];
    for (int k = 0; k < n; ++k) { ocol[newPtr[r] + k] = ccol[ptr[r] + k]; oval[newPtr[r] + k] = cval[ptr[r] + k]; }
}

// ------------------------------------------------------------------ reductions (double, deterministic)
struct LinState { int iterations; int status; double residual; };

__device__ __forceinline__ void lin_sum2(cg::grid_group& grid, double a, double b, double* partials /* 2 sets x 2 * LIN_MAX_BLOCKS */, double* sh, double out[2],
                                         unsigned& flip)
{
    partials += (flip++ & 1u) * 2u * (unsigned)LIN_MAX_BLOCKS;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[warp] = a; sh[8 + warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0;
        for (int w = 0; w < LIN_THREADS / 32; ++w) { sa += sh[w]; sb += sh[8 + w]; }
        partials[2 * blockIdx.x] = sa; partials[2 * blockIdx.x + 1] = sb;
    }
    grid.sync();
    if (warp == 0) {
        double sa = 0, sb = 0;
        for (int i = lane; i < (int)gridDim.x; i += 32) { sa += partials[2 * i]; sb += partials[2 * i + 1]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
        if (lane == 0) { sh[16] = sa; sh[17] = sb; }
    }
    __syncthreads();
    out[0] = sh[16]; out[1] = sh[17];
    __syncthreads();
}

struct LinCsr { int n; const int* ptr; const int* col; const double* val; };

__device__ __forceinline__ double lin_spmv_row(const LinCsr& A, int r, const double* x)
{
    double s = 0.0;
    const int e1 = A.ptr[r + 1];
    for (int e = A.ptr[r]; e < e1; ++e) s += A.val[e] * x[A.col[e]];
    return s;
}

// ------------------------------------------------------------------ IC(0) and its triangular solves (CGSolver, cg.cu:104-206)
// (value, tag) pairs in one 128-bit word; tag = launch-unique id: the row's entry is final
__device__ __forceinline__ double2 lin_ld128(const double2* p)
{
    unsigned long long a, b;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    return make_double2(__longlong_as_double((long long)a), __longlong_as_double((long long)b));
}
__device__ __forceinline__ void lin_st128(double2* p, double v, long long tag)
{
    const unsigned long long a = (unsigned long long)__double_as_longlong(v), b = (unsigned long long)tag;
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ double lin_wait(const double2* p, long long tag)
{
    double2 v = lin_ld128(p);
    for (unsigned spins = 0; __double_as_longlong(v.y) != tag; ++spins) { if (spins > 4) __nanosleep(40); v = lin_ld128(p); }
    return v.x;
}
// the same wait with ACQUIRE semantics: the caller goes on to read OTHER words the producer wrote before the tag
__device__ __forceinline__ double lin_wait_acquire(const double2* p, long long tag)
{
    for (unsigned spins = 0;; ++spins) {
        unsigned long long a, b;
        asm volatile("{\n\t.reg .b128 t;\n\tld.acquire.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}" : "=l"(a), "=l"(b) : "l"(p) : "memory");
        if ((long long)b == tag) return __longlong_as_double((long long)a);
        if (spins > 4) __nanosleep(40);
    }
}
__device__ __forceinline__ double lin_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Incomplete Cholesky on A's pattern (csric02): L_ik = (A_ik - sum_{j<k} L_ij L_kj) / L_kk, L_ii = sqrt(A_ii - sum_j L_ij^2).
// One warp per row, rows w, w + W, ... ascending; ic[e] = (L value, tag) for the LOWER entries (col <= row) of the CSR; `done[i]`
// = (L_ii, tag) announces row i.  mirror[e] = position of the transposed entry (col, row) in the CSR (the upper entries get
// L^T there for the backward solve).  A non-positive pivot sets status 2 (the reference's csric02 reports a zero pivot likewise).
__global__ void __launch_bounds__(LIN_THREADS) k_lin_ic0(LinCsr A, const int* __restrict__ diagPos, const int* __restrict__ mirror, double* __restrict__ ic,
                                                         double2* done, long long tag, LinState* st)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < A.n; i += nWarps) {
        const int b = A.ptr[i], dpos = diagPos[i];
        double diag = A.val[dpos];
        for (int t = b; t < dpos; ++t) {                 // the row's strictly lower entries, ascending column k
            const int k = A.col[t];
            const double lkk = lin_wait_acquire(done + k, tag);  // row k complete, its L entries visible
            // sum over j < k in BOTH patterns: lanes stride row k's lower entries, look each column up in row i's first t - b entries
            double s = 0.0;
            const int kb = A.ptr[k], kd = diagPos[k];
            for (int e = kb + lane; e < kd; e += 32) {
                const int j = A.col[e];
                int lo = b, hi = t;                        // binary search j in A.col[b, t)
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.col[mid] < j) lo = mid + 1; else hi = mid; }
                if (lo < t && A.col[lo] == j) s += ic[lo] * __ldcg(&ic[e]);      // (row k's entries: another SM wrote them -- not through L1)
            }
            s = lin_warp_sum(s);
            const double lik = (A.val[t] - s) / lkk;
            if (lane == 0) { ic[t] = lik; ic[mirror[t]] = lik; }
            __syncwarp();
            diag -= lik * lik;
        }
        double lii = 0.0;
        if (diag > 0.0) lii = sqrt(diag); else { if (lane == 0) atomicExch(&st->status, 2); lii = 1.0; }
        if (lane == 0) { ic[dpos] = lii; __threadfence(); lin_st128(done + i, lii, tag); }
        __syncwarp();
    }
}
// z = (L L^T)^-1 r inside the CG kernel: forward rows ascending, backward rows descending, one warp per row, hand-over by tags
__device__ __forceinline__ void lin_ic0_apply(const LinCsr& A, const int* diagPos, const double* ic, const double* r, double2* y, double2* z, double* zout,
                                              long long tag, int lane, int warp, int nWarps)
{
    for (int i = warp; i < A.n; i += nWarps) {
        const int b = A.ptr[i], d = diagPos[i];
        double s = 0.0;
        for (int e = b + lane; e < d; e += 32) s += ic[e] * lin_wait(y + A.col[e], tag);
        s = lin_warp_sum(s);
        if (lane == 0) lin_st128(y + i, (r[i] - s) / ic[d], tag);
    }
    for (int t = warp; t < A.n; t += nWarps) {
        const int i = A.n - 1 - t;
        const int d = diagPos[i], e1 = A.ptr[i + 1];
        double s = 0.0;
        for (int e = d + 1 + lane; e < e1; e += 32) s += ic[e] * lin_wait(z + A.col[e], tag);      // upper entries hold L^T
        s = lin_warp_sum(s);
        if (lane == 0) {
            const double v = (lin_wait(y + i, tag) - s) / ic[d];
            lin_st128(z + i, v, tag);
            zout[i] = v;
        }
    }
}

// ------------------------------------------------------------------ the solve: PCGJacobiSolver / CGSolver::Solve's loop
// IC0 = false: z = D^-1 r (ExtractInverseDiagonalKernel: |a_ii| < 1e-9 -> 1, pcgJacobi.cu:6-19).  Same sequence of operations as
// the reference's loop: ||r|| < tol -> stop; z; rho = r.z (Jacobi only: |rho| < 1e-15 -> stop); p = z + (rho / rho_prev) p;
// q = A p; alpha = rho / (p.q); x += alpha p; r -= alpha q.
template <bool IC0>
__global__ void __launch_bounds__(LIN_THREADS)
k_lin_cg(LinCsr A, const int* __restrict__ diagPos, const double* __restrict__ ic, const double* __restrict__ b, double* x, const double* guess,
         double* r, double* z, double* p, double* q, double2* ty, double2* tz, long long tagBase, int maxIter, double tol, double* partials, LinState* st)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[18];
    const int n = A.n, gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = gtid >> 5, nWarps = gstride >> 5;
    unsigned flip = 0;
    double s[2];
    // x = guess (or 0); r = b - A x
    if (guess) for (int v = gtid; v < n; v += gstride) x[v] = guess[v];
    else for (int v = gtid; v < n; v += gstride) x[v] = 0.0;
    grid.sync();
    double arr = 0;
    for (int v = gtid; v < n; v += gstride) {
        const double rv = guess ? b[v] - lin_spmv_row(A, v, x) : b[v];
        r[v] = rv;
        arr += rv * rv;
    }
    lin_sum2(grid, arr, 0.0, partials, sh, s, flip);
    double rn = sqrt(s[0]), rho = 0.0, rhoPrev = 0.0;
    int k = 0;
    for (; k < maxIter; ++k) {
        if (rn < tol) break;
        double arz = 0;
        if (IC0) {
            lin_ic0_apply(A, diagPos, ic, r, ty, tz, z, tagBase + k, lane, warp, nWarps);
            grid.sync();
            for (int v = gtid; v < n; v += gstride) arz += r[v] * z[v];
        } else {
            for (int v = gtid; v < n; v += gstride) {
                double dg = A.val[diagPos[v]];
                if (fabs(dg) < 1e-9) dg = 1.0;
                const double zv = r[v] * (1.0 / dg);
                z[v] = zv;
                arz += r[v] * zv;
            }
        }
        lin_sum2(grid, arz, 0.0, partials, sh, s, flip);
        rhoPrev = rho; rho = s[0];
        if (!IC0 && fabs(rho) < 1e-15) break;
        const double beta = (k == 0) ? 0.0 : rho / rhoPrev;
        for (int v = gtid; v < n; v += gstride) p[v] = (k == 0) ? z[v] : beta * p[v] + z[v];
        grid.sync();
        double apq = 0;
        for (int v = gtid; v < n; v += gstride) { const double qv = lin_spmv_row(A, v, p); q[v] = qv; apq += p[v] * qv; }
        lin_sum2(grid, apq, 0.0, partials, sh, s, flip);
        const double alpha = rho / s[0];
        arr = 0;
        for (int v = gtid; v < n; v += gstride) {
            x[v] += alpha * p[v];
            const double rv = r[v] - alpha * q[v];
            r[v] = rv;
            arr += rv * rv;
        }
        lin_sum2(grid, arr, 0.0, partials, sh, s, flip);
        rn = sqrt(s[0]);
    }
    if (gtid == 0) { st->iterations = k; st->residual = rn; }
}

__global__ void k_lin_diag_mirror(int N, const int* __restrict__ ptr, const int* __restrict__ col, int* __restrict__ diagPos, int* __restrict__ mirror, int* __restrict__ bad)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    int dp = -1;
    for (int e = ptr[r]; e < ptr[r + 1]; ++e) {
        const int c = col[e];
        if (c == r) dp = e;
        if (mirror) {           // position of (c, r) in row c (the Hessian's pattern is symmetric)
            int lo = ptr[c], hi = ptr[c + 1];
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < r) lo = mid + 1; else hi = mid; }
            if (lo < ptr[c + 1] && col[lo] == r) mirror[e] = lo; else { mirror[e] = e; atomicOr(bad, 2); }
        }
    }
    if (dp < 0) { atomicOr(bad, 4); dp = ptr[r]; }
    diagPos[r] = dp;
}

// ------------------------------------------------------------------ host side
struct LinearSolver::Impl {
    int device = 0, kind = 0, N = 0, maxIter = 0;
    double tol = 0;
    cudaStream_t stream = nullptr;
    // workspaces, grown on demand
    size_t capNz = 0;
    int *cnt = nullptr, *ptr = nullptr, *cursor = nullptr, *len = nullptr, *newPtr = nullptr, *slotEntry = nullptr, *ccol = nullptr, *col = nullptr, *diagPos = nullptr, *mirror = nullptr, *bad = nullptr;
    double *cval = nullptr, *val = nullptr, *ic = nullptr, *r = nullptr, *z = nullptr, *p = nullptr, *q = nullptr, *partials = nullptr;
    double2 *ty = nullptr, *tz = nullptr, *done = nullptr;
    LinState* st = nullptr;
    long long tag = 1;
    int grid = 0, numSms = 0;
    int lastIterations = 0, lastNnz = 0; double lastResidual = 0;
    std::vector<void*> allocs;
    template <typename T> T* alloc(size_t n) { void* p = nullptr; LIN_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); allocs.push_back(p); return static_cast<T*>(p); }
};

LinearSolver::LinearSolver(int kind, int N, int maxIter, double tol, int device) : d_(new Impl)
{
    if (kind != LIN_PCG_JACOBI && kind != LIN_CG_IC0) throw std::runtime_error("linear solver: kind must be 1 (CG + IC(0)) or 2 (PCG-Jacobi)");
    if (N <= 0) throw std::runtime_error("linear solver: N must be positive");
    Impl& d = *d_;
    d.kind = kind; d.N = N; d.device = device;
    d.maxIter = maxIter > 0 ? maxIter : (kind == LIN_PCG_JACOBI ? 2000 : 100);         // pcgJacobi.h:10, cg.h:10
    d.tol = tol > 0 ? tol : (kind == LIN_PCG_JACOBI ? 1e-5 : 1e-6);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device: the linear back-ends have no CPU fallback");
    LIN_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    LIN_CHECK(cudaGetDeviceProperties(&prop, device));
    d.numSms = prop.multiProcessorCount;
    LIN_CHECK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    const size_t n = (size_t)N;
    d.cnt = d.alloc<int>(n); d.ptr = d.alloc<int>(n + 1); d.cursor = d.alloc<int>(n); d.len = d.alloc<int>(n); d.newPtr = d.alloc<int>(n + 1);
    d.diagPos = d.alloc<int>(n); d.bad = d.alloc<int>(1);
    d.r = d.alloc<double>(n); d.z = d.alloc<double>(n); d.p = d.alloc<double>(n); d.q = d.alloc<double>(n);
    d.partials = d.alloc<double>(2 * 2 * (size_t)LIN_MAX_BLOCKS);
    d.ty = d.alloc<double2>(n); d.tz = d.alloc<double2>(n); d.done = d.alloc<double2>(n);
    LIN_CHECK(cudaMemset(d.ty, 0, n * 16)); LIN_CHECK(cudaMemset(d.tz, 0, n * 16)); LIN_CHECK(cudaMemset(d.done, 0, n * 16));
    d.st = d.alloc<LinState>(1);
    int perSm = 0, perSm2 = 0;
    LIN_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_lin_cg<false>, LIN_THREADS, 0));
    LIN_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm2, k_lin_cg<true>, LIN_THREADS, 0));
    perSm = std::max(1, std::min(std::min(perSm, perSm2), 4));
    d.grid = std::max(1, std::min(std::min(d.numSms * perSm, LIN_MAX_BLOCKS), (N + LIN_THREADS - 1) / LIN_THREADS));
}

LinearSolver::~LinearSolver()
{
    if (!d_) return;
    cudaSetDevice(d_->device);
    if (d_->stream) { cudaStreamSynchronize(d_->stream); cudaStreamDestroy(d_->stream); }
    for (void* p : d_->allocs) cudaFree(p);
}

void LinearSolver::stats(int* iterations, double* residual, int* nnz) const
{
    if (iterations) *iterations = d_->lastIterations;
    if (residual) *residual = d_->lastResidual;
    if (nnz) *nnz = d_->lastNnz;
}

void LinearSolver::solveDevice(int N, const double* b, double* x, const double* A, int nz, const int* rowIdx, const int* colIdx, const double* guess)
{
    Impl& d = *d_;
    if (N != d.N) throw std::runtime_error("linear solver: N differs from the size it was created for");
    if (!b || !x || !A || !rowIdx || !colIdx || nz <= 0) throw std::runtime_error("linear solver: NULL argument or empty matrix");
    LIN_CHECK(cudaSetDevice(d.device));
    // the caller's arrays were written on ITS stream (the reference: the legacy default stream)
    cudaEvent_t ev;
    LIN_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    LIN_CHECK(cudaEventRecord(ev, cudaStreamLegacy));
    LIN_CHECK(cudaStreamWaitEvent(d.stream, ev, 0));
    cudaEventDestroy(ev);
    if ((size_t)nz > d.capNz) {
        d.capNz = (size_t)nz + (size_t)nz / 4;
        d.slotEntry = d.alloc<int>(d.capNz); d.ccol = d.alloc<int>(d.capNz); d.col = d.alloc<int>(d.capNz); d.mirror = d.alloc<int>(d.capNz);
        d.cval = d.alloc<double>(d.capNz); d.val = d.alloc<double>(d.capNz); d.ic = d.alloc<double>(d.capNz);
    }
    cudaStream_t s = d.stream;
    const int tb = 256, gE = (nz + tb - 1) / tb, gN = (N + tb - 1) / tb;
    LIN_CHECK(cudaMemsetAsync(d.cnt, 0, (size_t)N * 4, s)); LIN_CHECK(cudaMemsetAsync(d.cursor, 0, (size_t)N * 4, s)); LIN_CHECK(cudaMemsetAsync(d.bad, 0, 4, s));
    k_lin_count<<<gE, tb, 0, s>>>(nz, rowIdx, N, d.cnt, d.bad);
    k_lin_scan<<<1, 1024, 0, s>>>(N, d.cnt, d.ptr);
    k_lin_scatter<<<gE, tb, 0, s>>>(nz, rowIdx, N, d.ptr, d.cursor, d.slotEntry);
    k_lin_sort_rows<<<gN, tb, 0, s>>>(N, d.ptr, d.slotEntry, colIdx, A, d.ccol, d.cval, d.len, d.bad);
    k_lin_scan<<<1, 1024, 0, s>>>(N, d.len, d.newPtr);
    k_lin_compact<<<gN, tb, 0, s>>>(N, d.ptr, d.newPtr, d.ccol, d.cval, d.col, d.val);
    const bool ic0 = d.kind == LIN_CG_IC0;
    k_lin_diag_mirror<<<gN, tb, 0, s>>>(N, d.newPtr, d.col, d.diagPos, ic0 ? d.mirror : nullptr, d.bad);
    int bad = 0, nnz = 0;
    LIN_CHECK(cudaMemcpyAsync(&bad, d.bad, 4, cudaMemcpyDeviceToHost, s));
    LIN_CHECK(cudaMemcpyAsync(&nnz, d.newPtr + N, 4, cudaMemcpyDeviceToHost, s));
    LIN_CHECK(cudaStreamSynchronize(s));
    if (bad & 1) throw std::runtime_error("linear solver: a COO index lies outside [0, N)");
    if (bad & 4) throw std::runtime_error("linear solver: a row has no diagonal entry");
    if (ic0 && (bad & 2)) throw std::runtime_error("linear solver: IC(0) needs a structurally symmetric matrix");
    d.lastNnz = nnz;
    LinCsr M{N, d.newPtr, d.col, d.val};
    LIN_CHECK(cudaMemsetAsync(d.st, 0, sizeof(LinState), s));
    long long tagBase = d.tag;
    if (ic0) {
        const long long ftag = d.tag++;
        k_lin_ic0<<<d.grid, LIN_THREADS, 0, s>>>(M, d.diagPos, d.mirror, d.ic, d.done, ftag, d.st);        // (not cooperative: grid <= resident capacity)
        tagBase = d.tag;
        d.tag += d.maxIter + 1;
    }
    const int* diagPos = d.diagPos; const double* ic = d.ic; double *r = d.r, *z = d.z, *p = d.p, *q = d.q, *partials = d.partials;
    double2 *ty = d.ty, *tz = d.tz; int maxIter = d.maxIter; double tol = d.tol; LinState* st = d.st;
    void* args[] = {&M, &diagPos, &ic, &b, &x, &guess, &r, &z, &p, &q, &ty, &tz, &tagBase, &maxIter, &tol, &partials, &st};
    LIN_CHECK(cudaLaunchCooperativeKernel(ic0 ? (const void*)k_lin_cg<true> : (const void*)k_lin_cg<false>, dim3(d.grid), dim3(LIN_THREADS), args, 0, s));
    LinState hs{};
    LIN_CHECK(cudaMemcpyAsync(&hs, d.st, sizeof(hs), cudaMemcpyDeviceToHost, s));
    LIN_CHECK(cudaStreamSynchronize(s));          // the caller reads x next, on its own stream
    d.lastIterations = hs.iterations; d.lastResidual = hs.residual;
    if (hs.status == 2) throw std::runtime_error("linear solver: IC(0) met a non-positive pivot (the matrix is not positive definite on its pattern)");
}

void LinearSolver::solveHost(int N, const double* b, double* x, const double* A, int nz, const int* rowIdx, const int* colIdx, const double* guess)
{
    Impl& d = *d_;
    LIN_CHECK(cudaSetDevice(d.device));
    double *db, *dx, *dA, *dg = nullptr; int *dr, *dc;
    LIN_CHECK(cudaMalloc(&db, 8ull * N)); LIN_CHECK(cudaMalloc(&dx, 8ull * N)); LIN_CHECK(cudaMalloc(&dA, 8ull * nz)); LIN_CHECK(cudaMalloc(&dr, 4ull * nz)); LIN_CHECK(cudaMalloc(&dc, 4ull * nz));
    LIN_CHECK(cudaMemcpy(db, b, 8ull * N, cudaMemcpyHostToDevice)); LIN_CHECK(cudaMemcpy(dA, A, 8ull * nz, cudaMemcpyHostToDevice));
    LIN_CHECK(cudaMemcpy(dr, rowIdx, 4ull * nz, cudaMemcpyHostToDevice)); LIN_CHECK(cudaMemcpy(dc, colIdx, 4ull * nz, cudaMemcpyHostToDevice));
    if (guess) { LIN_CHECK(cudaMalloc(&dg, 8ull * N)); LIN_CHECK(cudaMemcpy(dg, guess, 8ull * N, cudaMemcpyHostToDevice)); }
    try { solveDevice(N, db, dx, dA, nz, dr, dc, dg); }
    catch (...) { cudaFree(db); cudaFree(dx); cudaFree(dA); cudaFree(dr); cudaFree(dc); cudaFree(dg); throw; }
    LIN_CHECK(cudaMemcpy(x, dx, 8ull * N, cudaMemcpyDeviceToHost));
    cudaFree(db); cudaFree(dx); cudaFree(dA); cudaFree(dr); cudaFree(dc); cudaFree(dg);
}

}  // namespace pdb200
