// TEST INFRASTRUCTURE ONLY -- never part of libpd_b200.so.
//
// A minimal CUDA execution model for g++, so that the SOURCE of the product kernels (csrc/pd_kernels.cuh, rotation.cuh)
// can be compiled with -DPD_HOST_EMU and run on the host by the CPU test suite (tests/test_kernel_emulation.py): one OS
// thread per CUDA thread of a CTA, a real barrier for __syncthreads, CTAs one after the other, shared memory as a per-CTA
// buffer, cp.async / bulk copies executed at once by the issuing thread, mbarriers as phase counters.  It checks the
// kernels' LOGIC (indexing, staging, ordered sums, arithmetic forms) against the oracle without a GPU; it says nothing
// about timing, memory-model races between CTAs or instruction encodings -- those are the GPU suite's.
// Every PTX helper of the kernels has a host branch under PD_HOST_EMU next to its asm (same file, same function).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ inline __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3_emu { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace pd_emu {

struct Barrier {           // reusable barrier for the threads of one CTA
    std::mutex m; std::condition_variable cv; unsigned n = 0, waiting = 0, generation = 0;
    void arrive_and_wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = generation;
        if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != g; });
    }
};
struct Cta { uint8_t* smem = nullptr; Barrier bar; };
inline thread_local uint3_emu tIdx, bIdx;
inline thread_local dim3 bDim, gDim;
inline thread_local Cta* cta = nullptr;

// persistent worker threads (creating 256 OS threads per CTA launch would dominate the run time)
struct Pool {
    std::vector<std::thread> th;
    std::mutex m; std::condition_variable cv, doneCv;
    std::function<void(unsigned)> job;
    unsigned gen = 0, active = 0, remaining = 0;
    bool stop = false;
    void worker(unsigned i)
    {
        unsigned seen = 0;
        for (;;) {
            std::function<void(unsigned)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                if (i >= active) continue;
                f = job;
            }
            f(i);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--remaining == 0) doneCv.notify_all();
            }
        }
    }
    void run(unsigned n, std::function<void(unsigned)> f)
    {
        std::unique_lock<std::mutex> lk(m);
        while (th.size() < n) {
            const unsigned i = (unsigned)th.size();
            try { th.emplace_back([this, i] { worker(i); }); }
            catch (const std::system_error&) { throw std::runtime_error("host emulation: cannot create " + std::to_string(n) + " threads on this machine"); }
        }
        job = std::move(f); active = n; remaining = n; ++gen;
        cv.notify_all();
        doneCv.wait(lk, [&] { return remaining == 0; });
    }
    ~Pool()
    {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
};
inline Pool& pool() { static thread_local Pool p; return p; }      // one per driving thread (ranks may be driven concurrently)

// kernel<<<grid, block, smemBytes>>>(args...): CTAs sequentially, the threads of a CTA concurrently
template <typename K, typename... Args>
void launch(unsigned grid, unsigned block, size_t smemBytes, K kernel, Args... args)
{
    std::vector<uint8_t> raw(smemBytes + 256);
    for (unsigned b = 0; b < grid; ++b) {
        Cta c;
        c.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw.data()) + 127) & ~uintptr_t(127));
        std::memset(c.smem, 0xcd, smemBytes);        // uninitialised shared memory must not look like zeros
        c.bar.n = block;
        pool().run(block, [&](unsigned t) {
            tIdx = uint3_emu{t, 0, 0}; bIdx = uint3_emu{b, 0, 0}; bDim = dim3(block); gDim = dim3(grid); cta = &c;
            kernel(args...);
        });
    }
}
// the same with ALL CTAs of the grid resident at once (a persistent grid whose CTAs wait for each other or for other ranks)
template <typename K, typename... Args>
void launch_resident(unsigned grid, unsigned block, size_t smemBytes, K kernel, Args... args)
{
    std::vector<std::vector<uint8_t>> raw(grid, std::vector<uint8_t>(smemBytes + 256));
    std::vector<Cta> ctas(grid);
    for (unsigned b = 0; b < grid; ++b) {
        ctas[b].smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw[b].data()) + 127) & ~uintptr_t(127));
        std::memset(ctas[b].smem, 0xcd, smemBytes);
        ctas[b].bar.n = block;
    }
    pool().run(grid * block, [&](unsigned i) {
        tIdx = uint3_emu{i % block, 0, 0}; bIdx = uint3_emu{i / block, 0, 0}; bDim = dim3(block); gDim = dim3(grid); cta = &ctas[i / block];
        kernel(args...);
    });
}
// kernels without __syncthreads / shared memory (one thread per vertex): plain loops
template <typename K, typename... Args>
void launch_flat(unsigned grid, unsigned block, K kernel, Args... args)
{
    Cta c;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            tIdx = uint3_emu{t, 0, 0}; bIdx = uint3_emu{b, 0, 0}; bDim = dim3(block); gDim = dim3(grid); cta = &c;
            kernel(args...);
        }
}

}  // namespace pd_emu

#define threadIdx pd_emu::tIdx
#define blockIdx pd_emu::bIdx
#define blockDim pd_emu::bDim
#define gridDim pd_emu::gDim

inline void __syncthreads() { pd_emu::cta->bar.arrive_and_wait(); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline long long clock64()        // nanoseconds: the kernels' bounded waits (DIST_WAIT_LIMIT_CYCLES ~ 2e10) give up after 20 s instead of hanging a test
{
    return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline void __trap() { std::abort(); }

// single-rounding arithmetic (compile with -ffp-contract=off -mfma)
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __frsqrt_rn(float a) { return (float)(1.0 / std::sqrt((double)a)); }      // as oracle/pd_oracle.c:rsqrt_rn models it
inline float __fdividef(float a, float b) { return a / b; }                           // (approximate on the GPU: default rotation path only)
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
// fabsf, fmaf, fmaxf, fminf, sqrtf: the C library's (global namespace, <cmath>)
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int) { return v; }       // only in code paths the default build never takes (TILE_LPV == 2)
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// dynamic shared memory of the running CTA
#define PD_EMU_DYN_SMEM(name) uint8_t* name = pd_emu::cta->smem

// mbarrier as a 64-bit word: low half = completed phases, high half = bytes still expected in the current phase
namespace pd_emu {
inline void mbar_init(uint64_t* bar) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbar_expect(uint64_t* bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t)bytes << 32, __ATOMIC_SEQ_CST); }
inline void mbar_complete(uint64_t* bar, uint32_t bytes)
{
    const uint64_t v = __atomic_sub_fetch(bar, (uint64_t)bytes << 32, __ATOMIC_SEQ_CST);
    if ((v >> 32) == 0) __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);        // (one producer thread per barrier: no race between the two steps)
}
inline bool mbar_test(uint64_t* bar, uint32_t parity) { return ((uint32_t)__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1u) != parity; }
}  // namespace pd_emu
