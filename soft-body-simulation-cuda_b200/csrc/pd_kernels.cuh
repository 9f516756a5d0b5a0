// Hand-written sm_100a kernels for the projective-dynamics step (Jacobi/Chebyshev path).
// Reference behaviour restated per kernel; see DESIGN.md section 4 for the data layout and
// the roofline of each.
#pragma once
#include <cstdint>
#ifdef PD_HOST_EMU
#include "host_emu.hpp"      // tests/emu: the kernels' source compiled for the host by the CPU test suite (TEST ONLY, see there);
                             // every PTX helper below has a host branch next to its asm
#define PD_DYN_SMEM(name) PD_EMU_DYN_SMEM(name)
#else
#include <cuda_runtime.h>
#define PD_DYN_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#endif

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
#ifdef PD_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const uint8_t*>(p) - pd_emu::cta->smem); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { pd_emu::mbar_init(bar); }
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { pd_emu::mbar_expect(bar, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!pd_emu::mbar_test(bar, parity)) std::this_thread::yield(); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    std::memcpy(dst, src, bytes);
    pd_emu::mbar_complete(bar, bytes);
}
#else
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
{   // bounded: a bulk copy that never arrives (a bug, not a load condition) traps instead of hanging the GPU
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
// one bulk asynchronous copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

// L2 cache policies (the encodings CUTLASS uses for createpolicy.fractional.L2::evict_*): the tile
// stream is read once per iteration and is far larger than L2 -> evict_first; the per-vertex arrays the
// gathers hit (positions, b0) are re-read every iteration and fit in L2 -> evict_last.
constexpr unsigned long long L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
#ifdef PD_HOST_EMU
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, unsigned long long) { bulk_g2s(dst, src, bytes, bar); }
__device__ __forceinline__ void cp_async16_hint(uint32_t dstShared, const void* src, unsigned long long)
{   // (a ghost position may be arriving from another rank's thread right now: fourth lane -- the tag -- first, with acquire, so
    // that a fresh tag is never paired with a stale payload; see st_tagged_sys)
    uint32_t w = __atomic_load_n(static_cast<const uint32_t*>(src) + 3, __ATOMIC_ACQUIRE);
    std::memcpy(pd_emu::cta->smem + dstShared, src, 12);
    std::memcpy(pd_emu::cta->smem + dstShared + 12, &w, 4);
}
__device__ __forceinline__ void cp_async_commit() {}
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ void cp_async_wait_1() {}
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ float4 ldg_hint(const float4* p, unsigned long long) { return *p; }
#else
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, unsigned long long policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// 16-byte asynchronous gather global -> shared (LDGSTS), no register staging
__device__ __forceinline__ void cp_async16_hint(uint32_t dstShared, const void* src, unsigned long long policy)
{
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dstShared), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still draining; pdl_wait() blocks until the predecessor grid has
// completed and its writes are visible (a no-op for a normal launch), pdl_launch_dependents() lets the successor's
// CTAs be scheduled as soon as every CTA of this grid has got this far.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldg_hint(const float4* p, unsigned long long policy)
{
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}
#endif

// ------------------------------------------------------------------ multi-GPU halo exchange
// Every rank owns an exchange window [q0 | q1 | q2 | ...] that its neighbours map (CUDA IPC, or plain pointers inside one
// process).  Every phase that produced new positions (predictor, every PD iteration) is followed by a PUSH of the owner's
// boundary vertices straight into the neighbours' ghost entries over NVLink.  The push rides at the start of the local kernel
// that consumes those positions (below; k_halo_push is the stand-alone form for the PCG path and the lock-step test driver).
// FLAG IN THE DATA: a pushed position travels as ONE 16-byte store (x, y, z, tag), tag = the rank's phase count (`epoch`, the
// same number on every rank for the same phase), and a 16-byte aligned store is delivered whole.  The producer therefore
// needs NO fence, ticket or flag -- it fires its stores and goes on (a system-scope fence after remote stores costs an NVLink
// round trip, ~5 us per CTA and iteration when every CTA of the persistent grid had to publish its slice: measured,
// profiles/r2_dist_probes_n2_grid70.txt) -- and the consumer is the same launch on the other side: a thread that staged a ghost
// position for a boundary tile looks at the tag before the block barrier that hands the staging buffer to phase B and, only if
// it is stale, polls the entry itself.  A ghost entry of buffer k is rewritten every third phase (the buffers rotate), so a stale
// tag is the phase count minus three; the skew bound that makes the rotation safe is in pd_engine.cu:enqueueIteration.
struct DistWait {
    const unsigned long long* epoch;   // this rank's phase count = the tag of the positions this launch reads (bumped by k_predict /
                                       // k_vertex_jacobi / k_halo_push, never inside the local kernel)
    int nNbr;                          // 0 on a single GPU: no push, no check
    int nOwn;                          // local vertex ids >= nOwn are ghosts
    int firstTile;                     // first tile (in this rank's processing order) that reads ghost positions
    unsigned int* status;              // set to 1 when a wait gave up (peer died); results are then invalid
    // push at the START of the local kernel (nPush == 0: the pushes were separate launches): the boundary entries of
    // the position buffer this launch reads -> the neighbours' ghost entries, in push-list order (consecutive
    // destinations: full 128-byte NVLink writes)
    int nPush;
    const uint32_t* pushSrc;           // owned local vertex
    const uint32_t* pushDst;           // ghost index at the neighbour
    const uint32_t* pushNbr;           // neighbour slot
    float4* const* peerQ;              // [nNbr]: the neighbours' copies of the buffer this launch reads
    // programmatic dependent launch (PD_PDL): 0 / 1 = the successor may be scheduled as soon as every CTA has STARTED (measured
    // slower: its CTAs take SM slots from this kernel's), 2 = only once every CTA has finished its tiles (hides the launch latency)
    int pdlLate;
};
constexpr long long DIST_WAIT_LIMIT_CYCLES = 20000000000ll;    // ~10 s (ranks may enter their first step seconds apart): a hung peer must not hang this GPU

// one position + tag as a single 128-bit access at system scope (single-copy atomic: value and tag arrive together)
#ifdef PD_HOST_EMU
__device__ __forceinline__ void st_tagged_sys(float4* p, float4 v)
{   // the host has no 16-byte store that other threads see whole: payload first, then the tag with release
    p->x = v.x; p->y = v.y; p->z = v.z;
    __atomic_store_n(reinterpret_cast<uint32_t*>(&p->w), __float_as_uint(v.w), __ATOMIC_RELEASE);
}
__device__ __forceinline__ float4 ld_tagged_sys(const float4* p)
{   // ... and the tag first, with acquire
    float4 v;
    v.w = __uint_as_float(__atomic_load_n(reinterpret_cast<const uint32_t*>(&p->w), __ATOMIC_ACQUIRE));
    v.x = p->x; v.y = p->y; v.z = p->z;
    return v;
}
#else
__device__ __forceinline__ void st_tagged_sys(float4* p, float4 v)
{
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.sys.global.b128 [%0], t;\n\t}"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_tagged_sys(const float4* p)
{
    float4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.sys.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
#endif
// the tag of phase count e (never the 0 of a fresh buffer: the count starts at 1 before anything is consumed)
__device__ __forceinline__ float halo_tag(unsigned long long e) { return __uint_as_float((uint32_t)e); }

// ghost position `g` of the buffer this launch reads, staged at `slot` (shared memory) by the asynchronous gather: make sure it
// carries this phase's tag; poll the entry itself while it does not (bounded: a dead peer sets the status word instead of
// hanging this GPU)
__device__ __forceinline__ void halo_check_staged(float4* slot, const float4* g, uint32_t tag, unsigned int* status)
{
    if (__float_as_uint(slot->w) == tag) return;
    const long long t0 = clock64();
    for (;;) {
        const float4 v = ld_tagged_sys(g);
        if (__float_as_uint(v.w) == tag) { *slot = v; return; }
        if (clock64() - t0 > DIST_WAIT_LIMIT_CYCLES) { atomicExch(status, 1u); return; }
        __nanosleep(64);
    }
}

// stand-alone push (PCG path, lock-step test driver): positions of this rank's boundary vertices -> the neighbours' ghost
// entries, tagged with the NEW phase count, which the last block then makes current
__global__ void k_halo_push(int n, const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst, const uint32_t* __restrict__ nbrIdx,
                            const float4* __restrict__ q, float4* const* __restrict__ peerQ, unsigned long long* epoch, unsigned int* ticket)
{
    const unsigned long long e = *epoch + 1ull;            // (read before this block's ticket, hence before the last block bumps it)
    const float tag = halo_tag(e);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v = q[src[i]];
        v.w = tag;
        st_tagged_sys(&peerQ[nbrIdx[i]][dst[i]], v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) { *ticket = 0u; *epoch = e; }
    }
}

// ------------------------------------------------------------------ mouse-drag soft constraints
// SolverData<float>::moreDBC / OffsetX and MouseSelection::target (def.h:14-18,31-32), filled by Control_Kernel
// (simulationContext.cu:202-218) and consumed by five PdUtil kernels (pdUtil.cu:56-69,80-87,159-164,187-188,201-206).
// The per-vertex kernels below exist in two instantiations: DRAG = false is the headless path (moreDBC == 0
// everywhere, no extra loads), DRAG = true restates the reference's branches.  A dragged vertex is marked for the
// vertex kernel by a NEGATIVE cc.y (c + matrix_diag is positive otherwise).
struct DragArgs {
    const float* more;      // moreDBC per vertex (renumbered ids)
    const float4* offX;     // OffsetX
    float4* dbcx;           // DBCX: computeSn overwrites it with the drag target of every dragged vertex (pdUtil.cu:86)
    float tx, ty, tz;       // MouseSelection::target
    int hasDBC;             // SolverData::numDBC > 0: computeDBCLocal is launched at all (pdSolver.cu:170)
};

// ------------------------------------------------------------------ predictor
// gravity transform + setMDt_2MoreDBC + computeSn + the two D2D copies + addM_h2Sn
// (pdSolver.cu:154-160, pdUtil.cu:56-95,168-179).
// Writes q0 = prev = s, b0 = c * s_old (what addM_h2Sn recomputes every iteration; constant over
// a step), cc = (+-c, c + matrix_diag) with a NEGATIVE c marking a pinned (DBC) vertex.
// Arithmetic forms follow the reference's SASS: s = fma(f, 1/c, fma(v, dt, x)).
template <bool DRAG>
__global__ void k_predict(int nV, const float4* __restrict__ X, const float4* __restrict__ V,
                          const float* __restrict__ mass, const float* __restrict__ dbc,
                          const float* __restrict__ md, float dt, float dt2Prepared, float gravity,
                          float4* __restrict__ q0, float4* __restrict__ qprev, float4* __restrict__ b0,
                          float2* __restrict__ cc, DragArgs dr, unsigned long long* epochBump = nullptr)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    // multi-GPU: this kernel produced new positions -- one more phase (DistWait; the next local kernel pushes them with this tag)
    if (epochBump && v == 0) *epochBump += 1ull;
    if (v >= nV) return;
    const float m = mass[v], isDbc = dbc[v];
    const float dt2 = __fmul_rn(dt, dt);
    const float wi = DRAG ? dr.more[v] : 0.f;
    // DBC vertices keep the massDt_2s computed by setMDt_2 at prepare time (pdUtil.cu:48,59); the others get
    // (m + moreDBC) / dt^2 while they are dragged (pdUtil.cu:61-67)
    const float c = (isDbc == 0.f) ? __fdiv_rn((DRAG && wi > 0.f) ? __fadd_rn(m, wi) : m, dt2)
                                   : __fdiv_rn(__fadd_rn(m, __fmul_rn(isDbc, 1e6f)), dt2Prepared);
    float4 s;
    if (DRAG && wi > 0.f) {
        // computeSn, pdUtil.cu:80-87: sn = DBCX = target + OffsetX (whatever DBC says)
        const float4 o = dr.offX[v];
        s = make_float4(__fadd_rn(dr.tx, o.x), __fadd_rn(dr.ty, o.y), __fadd_rn(dr.tz, o.z), 0.f);
        dr.dbcx[v] = s;
    } else {
        const float4 x = X[v], vel = V[v];
        const float fy = __fmul_rn(-gravity, m);
        const float dt2_m_1 = __fdiv_rn(1.0f, c);
        s.x = __fmaf_rn(0.0f, dt2_m_1, __fmaf_rn(vel.x, dt, x.x));
        s.y = __fmaf_rn(fy, dt2_m_1, __fmaf_rn(vel.y, dt, x.y));
        s.z = __fmaf_rn(0.0f, dt2_m_1, __fmaf_rn(vel.z, dt, x.z));
        s.w = 0.f;
    }
    const float den = __fadd_rn(c, md[v]);
    const float2 c2 = make_float2(isDbc > 0.f ? -c : c, (DRAG && wi > 0.f) ? -den : den);
    cc[v] = c2;
    // the same two scalars ride in the unused w lanes of the arrays the Jacobi sweep reads anyway (b0.w = +-c; q.w = +-(c + md),
    // handed from iterate to iterate), so that sweep does not read cc at all: 8 bytes per vertex and iteration less
    s.w = c2.y;
    q0[v] = s;
    qprev[v] = s;
    b0[v] = make_float4(__fmul_rn(c, s.x), __fmul_rn(c, s.y), __fmul_rn(c, s.z), c2.x);
}

// ------------------------------------------------------------------ local step (the hot kernel)
// PdUtil::addM_h2Sn + computeLocal (pdUtil.cu:97-145,168-179), one tile of <= TILE_T tets per
// iteration of a persistent CTA (4 CTAs per SM).  v7: ONE __syncthreads per tile.  Per tile:
//   A. the tile's distinct vertex positions are gathered into shared memory by 16-byte cp.async
//      (LDGSTS) into a DOUBLE-BUFFERED staging area, issued more than a whole tile ahead (right after
//      the barrier of tile it-2); the vertex ids come from the global slot-indexed vlist, read
//      (coalesced) one more tile ahead into a register;
//   B. one tet per thread.  The 48-byte record (DmInv, w, corner words) is stored as three 16-byte
//      planes per tile and comes STRAIGHT FROM GLOBAL MEMORY into registers with three fully
//      coalesced LDG.128 (L2 evict_first), issued at the end of phase B of the previous tile so that
//      the latency hides behind that tile's phase C; the producer thread keeps the stream two tiles
//      ahead in L2 with one bulk prefetch per tile.  4 x LDS.128 positions, F = Ds*DmInv, rotation,
//      H = w (R - F) DmInv^T G (or w R DmInv^T G), 4 x STS.128 into this tile's H scratch (double
//      buffered: fast warps run ahead into the next tile while slow ones still sum the previous one);
//   -- barrier: H scratch complete, next tile's positions staged --
//   C. one tile-local vertex per thread, 32 vertices of similar incidence count per warp: ordered sum
//      over the vertex's incidence list (transposed rows, moved in by ONE bulk copy per tile (TMA,
//      cp.async.bulk + mbarrier, double buffered, requested a whole tile ahead); one conflict-free
//      LDS.32 yields two ready-made byte offsets into the H scratch) -> ONE partial sum per (tile,
//      vertex) slot.  Faithful mode: the vertex's first (owner) slot starts from b0 = (M/h^2) s_old, so
//      a vertex whose tets all sit in one tile gets exactly the reference's sequential
//      b = c*s; b += h1; b += h2; ...
// No atomics anywhere: the reference's 12 float atomicAdds per tet become ordered sums, so results
// are run-to-run bit-identical.
// F and H of the scalar path are written with the fused/rounded operation pattern nvcc gives the
// reference's glm expressions (see oracle/pd_oracle.c header), so that with ROT_MODE 1 every tet
// contribution is bit-identical to the reference kernel's.
constexpr uint32_t LOCAL_QS_BYTES = 16u * TILE_NLMAX;
constexpr uint32_t LOCAL_HS_BYTES = TILE_HS_BYTES;             // four corner planes of float4 + eight zero slots, one per column
constexpr uint32_t LOCAL_OFF_QS = 0u;
constexpr uint32_t LOCAL_OFF_HS = LOCAL_OFF_QS + 2u * LOCAL_QS_BYTES;
constexpr uint32_t LOCAL_OFF_C = LOCAL_OFF_HS + 2u * LOCAL_HS_BYTES;
constexpr uint32_t LOCAL_OFF_BAR = LOCAL_OFF_C + 2u * TILE_CMAX;
constexpr uint32_t LOCAL_SMEM_BYTES = LOCAL_OFF_BAR + 32u;      // two mbarriers
static_assert(4u * (LOCAL_SMEM_BYTES + 1024u) <= 233472u, "4 CTAs of the local kernel must fit one SM's shared memory");
static_assert(TILE_NLMAX == TILE_T, "one tile-local vertex per thread");
// (the per-tile entry of the device tile table, TILE_META_WORDS words, is described in layout.hpp)

__device__ __forceinline__ float dot3_nv(float a0, float b0, float a1, float b1, float a2, float b2)
{   // a0*b0 + a1*b1 + a2*b2 as nvcc contracts the reference's glm products
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}
#ifdef PD_HOST_EMU
__device__ __forceinline__ float4 ldg_stream(const void* p) { return *static_cast<const float4*>(p); }
__device__ __forceinline__ void bulk_prefetch_l2(const void*, uint32_t) {}
#else
__device__ __forceinline__ float4 ldg_stream(const void* p)
{   // read-once stream: no L1 allocation, L2 evict_first
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(L2_EVICT_FIRST));
    return v;
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#endif

// one H-scratch entry (x, y, z) at byte offset `off` of the tile's scratch Hb
__device__ __forceinline__ float4 h_load(const uint8_t* Hb, uint32_t off)
{
    return *reinterpret_cast<const float4*>(Hb + off);
}
// ... and the store of phase B: Hblk = the 128-byte line of the thread's quarter-warp in corner plane 0, half = the corner word
__device__ __forceinline__ void h_store(uint8_t* Hblk, uint32_t corner, uint32_t half, const float4 h)
{
    // conflict-free columns from the layout's 8-colouring (layout.cpp:color_tile): col * 16 = (half >> 8) & 0x70
    *reinterpret_cast<float4*>(Hblk + corner * TILE_HSTRIDE + ((half >> 8) & 0x70u)) = h;
}

// phase C of one vertex group of a tile, by one warp: Cb = the tile's incidence rows, Hb = its H scratch, gw = the
// group's tile-table word, slotG = the group's first slot.  32-vertex groups: one lane per vertex.  16-vertex groups:
// vertex i is summed by lanes i (entries 4r, 4r+1 of its list in row r) and i + 16 (entries 4r+2, 4r+3), combined by
// one shuffle.
// FAITHFUL: b = b0; b += h_1; b += h_2; ... strictly in (tet, corner) order (a second lane hands its entries over by
// shuffle), the vertex's first (owner) slot starting from b0 = (M/h^2) s_old -- the reference's sequential sum for a
// vertex whose tets share a tile.  Otherwise (product default) the slot holds the elastic terms only (the vertex
// kernel adds b0), summed in a different FIXED order without a long dependency chain: two rows per trip, their
// entries prefetched a trip ahead, four gathered LDS.128 in flight, the two entries of a lane accumulated separately
// ((x, y) as one packed FADD2).  A row past the group's count reads the zero slot.
template <bool FAITHFUL>
__device__ __forceinline__ void local_phase_c(const uint8_t* Cb, const uint8_t* Hb, uint32_t gw, uint32_t slotG, int lane,
                                              const uint32_t* __restrict__ vlist, const float4* __restrict__ b0, float4* __restrict__ P)
{
    const uint32_t nR = (gw >> 6) & 63u;
    if (nR == 0u) return;                                     // warp-uniform: no such group in this tile
    const uint32_t* row = reinterpret_cast<const uint32_t*>(Cb) + (gw & 63u) * 32u + lane;
    const bool writer = (uint32_t)lane < ((gw >> 12) & 63u);
    float sx, sy, sz;
    if (FAITHFUL) {
        sx = sy = sz = 0.f;
        if (writer) {
            const uint32_t ve = __ldg(&vlist[slotG + lane]);
            if (ve & TILE_OWNER_BIT) {
                const float4 bb = ldg_hint(&b0[ve & ~TILE_OWNER_BIT], L2_EVICT_LAST);
                sx = bb.x; sy = bb.y; sz = bb.z;
            }
        }
        for (uint32_t r = 0; r < nR; ++r) {
            const uint32_t e2 = row[32u * r];
            const float4 ha = h_load(Hb, e2 & 0xffffu);
            const float4 hb = h_load(Hb, e2 >> 16);
            sx = __fadd_rn(sx, ha.x); sy = __fadd_rn(sy, ha.y); sz = __fadd_rn(sz, ha.z);
            sx = __fadd_rn(sx, hb.x); sy = __fadd_rn(sy, hb.y); sz = __fadd_rn(sz, hb.z);
            if (TILE_LPV == 2) {
                sx = __fadd_rn(sx, __shfl_down_sync(0xffffffffu, ha.x, 16)); sy = __fadd_rn(sy, __shfl_down_sync(0xffffffffu, ha.y, 16)); sz = __fadd_rn(sz, __shfl_down_sync(0xffffffffu, ha.z, 16));
                sx = __fadd_rn(sx, __shfl_down_sync(0xffffffffu, hb.x, 16)); sy = __fadd_rn(sy, __shfl_down_sync(0xffffffffu, hb.y, 16)); sz = __fadd_rn(sz, __shfl_down_sync(0xffffffffu, hb.z, 16));
            }
        }
    } else {
        constexpr uint32_t ZZ = TILE_ZERO_OFF | (TILE_ZERO_OFF << 16);
        // (x, y) of every entry is added as one packed FADD2 (rotation.cuh), z as a scalar
        float2 axy = make_float2(0.f, 0.f), cxy = make_float2(0.f, 0.f);
        float az = 0.f, cz = 0.f;
        uint32_t e0 = row[0], e1 = (1u < nR) ? row[32] : ZZ;
#pragma unroll 1
        for (uint32_t r = 0; r < nR; r += 2) {
            const float4 h0 = h_load(Hb, e0 & 0xffffu);
            const float4 h1 = h_load(Hb, e0 >> 16);
            const float4 h2 = h_load(Hb, e1 & 0xffffu);
            const float4 h3 = h_load(Hb, e1 >> 16);
            const uint32_t* nx = row + 32u * (r + 2u);
            e0 = (r + 2u < nR) ? nx[0] : ZZ; e1 = (r + 3u < nR) ? nx[32] : ZZ;
            axy = f2add(axy, make_float2(h0.x, h0.y)); az = __fadd_rn(az, h0.z);
            cxy = f2add(cxy, make_float2(h1.x, h1.y)); cz = __fadd_rn(cz, h1.z);
            axy = f2add(axy, make_float2(h2.x, h2.y)); az = __fadd_rn(az, h2.z);
            cxy = f2add(cxy, make_float2(h3.x, h3.y)); cz = __fadd_rn(cz, h3.z);
        }
        const float2 sxy = f2add(axy, cxy);
        sx = sxy.x; sy = sxy.y; sz = __fadd_rn(az, cz);
        if (TILE_LPV == 2) {
            sx = __fadd_rn(sx, __shfl_down_sync(0xffffffffu, sx, 16));
            sy = __fadd_rn(sy, __shfl_down_sync(0xffffffffu, sy, 16));
            sz = __fadd_rn(sz, __shfl_down_sync(0xffffffffu, sz, 16));
        }
    }
    if (writer) P[slotG + lane] = make_float4(sx, sy, sz, 0.f);
}

// the contributions of ONE tet to the right-hand sides of its four vertices (PdUtil::computeLocal, pdUtil.cu:107-143):
// F = Ds DmInv, rotation, H = w (R - F) DmInv^T G (Jacobi mode) or w R DmInv^T G (direct / CG modes).  B0..B8 = DmInv
// (row-major), p0..p3 = the corner positions; shared by the tile kernel below and the per-body kernel (pd_body_kernel.cuh)
// (split in two so that a kernel can run the slow part on a COMPACTED list of the tets that need it: pd_body_kernel.cuh)
// ... the product default's FAST part: packed FP32, two unrolled polar Newton steps; false (nothing written) when the tet is
// inverted / flat / strongly deformed and needs tet_contrib_slow
template <bool JACOBI>
__device__ __forceinline__ bool tet_contrib_fast(const float B0, const float B1, const float B2, const float B3, const float B4, const float B5, const float B6,
                                            const float B7, const float B8, const float w, const float4 p0, const float4 p1, const float4 p2, const float4 p3,
                                            float4& h0, float4& h1, float4& h2, float4& h3)
{
        bool done = false;
        {
            // product default: packed FP32 (FFMA2) on matrix columns, rotation.cuh.  Edge k = p_{k+1} - p_0 is
            // column k of Ds; column c of F = sum_k edge_k B[k][c]
            const float2 p0xy = make_float2(p0.x, p0.y);
            Col3 e0, e1, e2;
            e0.xy = f2sub(make_float2(p1.x, p1.y), p0xy); e0.z = p1.z - p0.z;
            e1.xy = f2sub(make_float2(p2.x, p2.y), p0xy); e1.z = p2.z - p0.z;
            e2.xy = f2sub(make_float2(p3.x, p3.y), p0xy); e2.z = p3.z - p0.z;
            Col3 f0, f1, f2;
            f0.xy = f2fma(e2.xy, bc2(B6), f2fma(e0.xy, bc2(B0), f2mul(e1.xy, bc2(B3)))); f0.z = fmaf(e2.z, B6, fmaf(e0.z, B0, e1.z * B3));
            f1.xy = f2fma(e2.xy, bc2(B7), f2fma(e0.xy, bc2(B1), f2mul(e1.xy, bc2(B4)))); f1.z = fmaf(e2.z, B7, fmaf(e0.z, B1, e1.z * B4));
            f2.xy = f2fma(e2.xy, bc2(B8), f2fma(e0.xy, bc2(B2), f2mul(e1.xy, bc2(B5)))); f2.z = fmaf(e2.z, B8, fmaf(e0.z, B2, e1.z * B5));
            Col3 y0, y1, y2;
            if (rotation_newton2_packed(f0, f1, f2, y0, y1, y2)) {
                // M = w (R - F) = (w/4) Y - w F   (or w R), column k
                const float w4 = 0.25f * w;
                Col3 m0, m1, m2;
                if (JACOBI) {
                    m0.xy = f2fma(y0.xy, bc2(w4), f2mul(f0.xy, bc2(-w))); m0.z = fmaf(y0.z, w4, f0.z * -w);
                    m1.xy = f2fma(y1.xy, bc2(w4), f2mul(f1.xy, bc2(-w))); m1.z = fmaf(y1.z, w4, f1.z * -w);
                    m2.xy = f2fma(y2.xy, bc2(w4), f2mul(f2.xy, bc2(-w))); m2.z = fmaf(y2.z, w4, f2.z * -w);
                } else {
                    m0.xy = f2mul(y0.xy, bc2(w4)); m0.z = y0.z * w4;
                    m1.xy = f2mul(y1.xy, bc2(w4)); m1.z = y1.z * w4;
                    m2.xy = f2mul(y2.xy, bc2(w4)); m2.z = y2.z * w4;
                }
                // H = M DmInv^T: column j (vertex j+1) = sum_k M_k B[j][k]; vertex 0: -(sum of the columns)
                const float2 a1 = f2fma(m2.xy, bc2(B2), f2fma(m0.xy, bc2(B0), f2mul(m1.xy, bc2(B1))));
                const float2 a2 = f2fma(m2.xy, bc2(B5), f2fma(m0.xy, bc2(B3), f2mul(m1.xy, bc2(B4))));
                const float2 a3 = f2fma(m2.xy, bc2(B8), f2fma(m0.xy, bc2(B6), f2mul(m1.xy, bc2(B7))));
                const float z1 = fmaf(m2.z, B2, fmaf(m0.z, B0, m1.z * B1));
                const float z2 = fmaf(m2.z, B5, fmaf(m0.z, B3, m1.z * B4));
                const float z3 = fmaf(m2.z, B8, fmaf(m0.z, B6, m1.z * B7));
                const float2 a0 = f2sub(f2sub(make_float2(-a1.x, -a1.y), a2), a3);
                h0 = make_float4(a0.x, a0.y, (-z1 - z2) - z3, 0.f);
                h1 = make_float4(a1.x, a1.y, z1, 0.f);
                h2 = make_float4(a2.x, a2.y, z2, 0.f);
                h3 = make_float4(a3.x, a3.y, z3, 0.f);
                done = true;
            }
        }
        return done;
}
// ... and the scalar part in the reference's operation order (the faithful mode takes it for every tet)
template <int ROT_MODE, bool JACOBI>
__device__ __forceinline__ void tet_contrib_slow(const float B0, const float B1, const float B2, const float B3, const float B4, const float B5, const float B6,
                                            const float B7, const float B8, const float w, const float4 p0, const float4 p1, const float4 p2, const float4 p3,
                                            float4& h0, float4& h1, float4& h2, float4& h3)
{
        {
            // faithful mode, and the default mode's rare inverted / flat / strongly deformed tets: scalar path in
            // the reference's operation order.  Ds columns = edges; F = Ds * DmInv  (row-major F[r][c] = sum_k Ds[r][k] B[k][c])
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
            h1.x = dot3_nv(M[0], B0, M[1], B1, M[2], B2); h2.x = dot3_nv(M[0], B3, M[1], B4, M[2], B5); h3.x = dot3_nv(M[0], B6, M[1], B7, M[2], B8);
            h1.y = dot3_nv(M[3], B0, M[4], B1, M[5], B2); h2.y = dot3_nv(M[3], B3, M[4], B4, M[5], B5); h3.y = dot3_nv(M[3], B6, M[4], B7, M[5], B8);
            h1.z = dot3_nv(M[6], B0, M[7], B1, M[8], B2); h2.z = dot3_nv(M[6], B3, M[7], B4, M[8], B5); h3.z = dot3_nv(M[6], B6, M[7], B7, M[8], B8);
            h0.x = __fsub_rn(__fsub_rn(-h1.x, h2.x), h3.x);
            h0.y = __fsub_rn(__fsub_rn(-h1.y, h2.y), h3.y);
            h0.z = __fsub_rn(__fsub_rn(-h1.z, h2.z), h3.z);
            h0.w = h1.w = h2.w = h3.w = 0.f;
        }
}
template <int ROT_MODE, bool JACOBI>
__device__ __forceinline__ void tet_contrib(const float B0, const float B1, const float B2, const float B3, const float B4, const float B5, const float B6,
                                            const float B7, const float B8, const float w, const float4 p0, const float4 p1, const float4 p2, const float4 p3,
                                            float4& h0, float4& h1, float4& h2, float4& h3)
{
        if (ROT_MODE == 0 && tet_contrib_fast<JACOBI>(B0, B1, B2, B3, B4, B5, B6, B7, B8, w, p0, p1, p2, p3, h0, h1, h2, h3)) return;
        tet_contrib_slow<ROT_MODE, JACOBI>(B0, B1, B2, B3, B4, B5, B6, B7, B8, w, p0, p1, p2, p3, h0, h1, h2, h3);
}

// phase B of one tet: record (r0, r1, r2) in registers, positions from the staging buffer qsb, the four corner
// contributions into the quarter-warp's H scratch line
template <int ROT_MODE, bool JACOBI>
__device__ __forceinline__ void local_phase_b(const float4 r0, const float4 r1, const float4 r2, const uint8_t* qsb, uint8_t* Hline)
{
        const uint32_t c01 = __float_as_uint(r2.z), c23 = __float_as_uint(r2.w);
        const float4 p0 = *reinterpret_cast<const float4*>(qsb + (c01 & 0x0ff0u));
        const float4 p1 = *reinterpret_cast<const float4*>(qsb + ((c01 >> 16) & 0x0ff0u));
        const float4 p2 = *reinterpret_cast<const float4*>(qsb + (c23 & 0x0ff0u));
        const float4 p3 = *reinterpret_cast<const float4*>(qsb + ((c23 >> 16) & 0x0ff0u));
        float4 h0, h1, h2, h3;
        tet_contrib<ROT_MODE, JACOBI>(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, p0, p1, p2, p3, h0, h1, h2, h3);
        h_store(Hline, 0u, c01 & 0xffffu, h0);
        h_store(Hline, 1u, c01 >> 16, h1);
        h_store(Hline, 2u, c23 & 0xffffu, h2);
        h_store(Hline, 3u, c23 >> 16, h3);
}

#ifndef PD_VSTAGE_PREFETCH
#define PD_VSTAGE_PREFETCH 1
#endif
template <int ROT_MODE, bool JACOBI, bool PROF = false>
__global__ void __launch_bounds__(TILE_T, 4)      // (compiled for 3 CTAs per SM -- 70 registers -- the kernel is 4 % slower: profiles/r2_vslot4_mb3_ab_grid139.txt)
k_local(const uint8_t* __restrict__ records, const uint32_t* __restrict__ tileMeta, int nTiles,
        const uint32_t* __restrict__ vstage, const uint32_t* __restrict__ vlist, const float4* __restrict__ q, const float4* __restrict__ b0, float4* __restrict__ P,
        unsigned long long* __restrict__ prof, DistWait dw)
{
    // prof (may be null): per-phase clock64 totals of warp 0 of every CTA, 8 counters per CTA (scripts/phase_profile.py)
    PD_DYN_SMEM(smem);
    uint64_t* cbar = reinterpret_cast<uint64_t*>(smem + LOCAL_OFF_BAR);   // part C buffers 0, 1
    const int tid = threadIdx.x;
    const int nIt = ((int)blockIdx.x < nTiles) ? (nTiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (nIt == 0) return;

    if (tid == 0) {
        mbar_init(&cbar[0], 1);
        mbar_init(&cbar[1], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (tid < 16) *reinterpret_cast<float4*>(smem + LOCAL_OFF_HS + (tid >> 3) * LOCAL_HS_BYTES + TILE_ZERO_OFF + 16 * (tid & 7)) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    auto meta_of = [&](int k) -> const uint32_t* { return tileMeta + (size_t)(blockIdx.x + k * gridDim.x) * TILE_META_WORDS; };
    const int warp = tid >> 5, lane = tid & 31;
    // the vertex staged in this thread's staging slot for tile number k of this CTA (Layout::vstage)
    auto load_ve = [&](int k) -> uint32_t { return __ldg(&vstage[(blockIdx.x + k * gridDim.x) * (unsigned)TILE_NLMAX + tid]); };
    // multi-GPU: the rank's interior tiles come first; only its boundary tiles (>= firstTile) read ghost positions.  After this
    // thread's asynchronous copy for tile k has landed and BEFORE the barrier that hands the staging buffer to phase B, a thread
    // that staged a ghost checks its tag (DistWait) -- normally one shared-memory load and a compare: the neighbours pushed at the
    // start of their launch and the boundary tiles come last
    auto halo_check = [&](int k) {
        if (dw.nNbr > 0 && (int)(blockIdx.x + k * gridDim.x) >= dw.firstTile) {
            const uint32_t ve = load_ve(k) & ~TILE_OWNER_BIT;         // (0xffffffff = empty slot: not < anything below)
            if (ve != (0xffffffffu & ~TILE_OWNER_BIT) && (int)ve >= dw.nOwn)
                halo_check_staged(reinterpret_cast<float4*>(smem + LOCAL_OFF_QS + (k & 1) * LOCAL_QS_BYTES + 16 * tid), &q[ve],
                                  (uint32_t)*dw.epoch, dw.status);
        }
    };
    // this thread's record (three coalesced 16-byte planes) of a tile with nTets tets whose record starts at 16 * off16
    auto load_rec = [&](uint32_t off16, uint32_t nTets, float4& a0, float4& a1, float4& a2) {
        const uint8_t* base = records + 16ull * off16 + TILE_OFF_TETS + 16 * tid;
        if ((uint32_t)tid < nTets) { a0 = ldg_stream(base); a1 = ldg_stream(base + 16u * nTets); a2 = ldg_stream(base + 32u * nTets); }
    };
    // asynchronous gather of this thread's vertex into the staging buffers, alternating (tile k -> buffer k & 1).
    // The two destinations live in per-thread registers that swap every call: when ptxas 12.9 folds a
    // uniform-register buffer offset into an LDGSTS that also carries a cache-hint descriptor it emits an encoding
    // the B200 rejects ("illegal instruction")
    uint32_t gdst = smem_u32(smem + LOCAL_OFF_QS + 16 * tid);
    uint32_t gdstOther = gdst + LOCAL_QS_BYTES;
    auto gather = [&](uint32_t ve) {
#ifndef PD_HOST_EMU
        asm volatile("" : "+r"(gdst));      // (opaque to the optimiser: the destination stays a plain per-thread register, see above)
#endif
        if (ve != 0xffffffffu) {
            cp_async16_hint(gdst, &q[ve & ~TILE_OWNER_BIT], L2_EVICT_LAST);
            if (ROT_MODE == 1 && (ve & TILE_OWNER_BIT)) prefetch_l2(&b0[ve & ~TILE_OWNER_BIT]);
        }
        cp_async_commit();
        const uint32_t t = gdst; gdst = gdstOther; gdstOther = t;
    };
    // producer duties: lane 0 of the LAST warp (the tile-local vertices are sorted by incidence count, so warp 0
    // carries the longest phase C and the last warp the shortest, usually none at all).  Measured and rejected:
    // heaviest groups on the highest warp ids (+3 %), the record's table entry requested a phase earlier (+1 %),
    // 16-vertex groups with two lanes per vertex (-50 % longest list, +10 % instructions: +6 %); round 2: the position gather issued
    // by the upper half of the CTA only, two slots per thread (the heavy warps' stall after the barrier GREW: it is the scheduler's
    // memory-instruction queue behind the partner warp's scattered LDGSTS, +2 %, profiles/r2_half_cta_gather_ab_grid139.txt); two lanes per
    // vertex for the heaviest 16 vertices of a tile only (longest group 11.8 -> 6.0 rows, fewer rows in total: still +1.6 %,
    // profiles/r2_mixed_lpv_groups_ab_grid139.txt -- the tile is bound by the SM's throughput, not by its heaviest warp); warps 7 / 6 / 5
    // summing the odd rows of groups 0 / 1 / 2 and handing them over through shared memory and a named barrier
    // (bar.arrive / bar.sync on 64 threads): +19 % instructions, barrier stall 3.3 -> 4.0 per issue, +25 % time
    // (profiles/r2_k_local_split_c_ncu_summary.txt); x | y | z planes for the H scratch and predicated pad loads: no gain
    // (profiles/r2_ab_planes_pred.txt)
    const bool producer = tid == TILE_T - 32;
    auto fetch_c = [&](int k) {        // part C of tile k -> buffer k & 1
        const uint2 te = __ldg(reinterpret_cast<const uint2*>(meta_of(k)));       // off / 16, abBytes | cBytes << 16
        const uint32_t abBytes = te.y & 0xffffu, cBytes = te.y >> 16;
        mbar_expect_tx(&cbar[k & 1], cBytes);
        bulk_g2s_hint(smem + LOCAL_OFF_C + (k & 1) * TILE_CMAX, records + 16ull * te.x + abBytes, cBytes, &cbar[k & 1], L2_EVICT_FIRST);
    };
    auto prefetch_ab = [&](int k) {    // part AB of tile k -> L2
        const uint2 te = __ldg(reinterpret_cast<const uint2*>(meta_of(k)));
        bulk_prefetch_l2(records + 16ull * te.x, te.y & 0xffffu);
    };
    // ... and the tile's staging list (1 KB of the read-once, DRAM-resident vstage array) -> L2, two tiles before load_ve() reads
    // it: that load is consumed exactly one tile later by gather(), and a tile (~4,300 cycles per CTA) is about one LOADED HBM
    // round trip -- the phase profile showed the heaviest warp waiting ~1,000 cycles per tile for it (profiles/r1_phase_profile_*)
    auto prefetch_vs = [&](int k) {
#if PD_VSTAGE_PREFETCH
        bulk_prefetch_l2(vstage + (size_t)(blockIdx.x + k * gridDim.x) * TILE_NLMAX, 4u * TILE_NLMAX);
#endif
    };

    // ---- prologue: everything here touches only the immutable tile stream
    if (producer) {
        fetch_c(0);
        if (nIt > 1) { fetch_c(1); prefetch_ab(1); }
        if (nIt > 2) prefetch_ab(2);
        if (nIt > 3) prefetch_vs(3);
        if (nIt > 4) prefetch_vs(4);
    }
    // tile-table words of the current tile: group `warp` (with the tet count) and, with 16-vertex groups, `warp + 8`
    uint32_t mwA = __ldg(meta_of(0) + 4 + warp), mwB = (TILE_NGROUPS > 8) ? __ldg(meta_of(0) + 4 + (warp + 8) % TILE_NGROUPS) : 0u;
    uint32_t veN = load_ve(0);
    float4 r0, r1, r2;
    load_rec(__ldg(meta_of(0)), mwA >> 18, r0, r1, r2);
    // from here on the kernel reads positions written by the previous launch in the stream (and later overwrites
    // the slots that launch reads)
    if (dw.pdlLate == 0) pdl_launch_dependents();
    pdl_wait();
    if (dw.nPush > 0) {
        // multi-GPU halo push, spread over all CTAs, fire and forget: every position travels with this phase's tag in one
        // 16-byte store (DistWait); the exchange overlaps the interior tiles
        const float tag = halo_tag(*dw.epoch);
        for (int i = blockIdx.x * TILE_T + tid; i < dw.nPush; i += gridDim.x * TILE_T) {
            float4 v = q[dw.pushSrc[i]];
            v.w = tag;
            st_tagged_sys(&dw.peerQ[dw.pushNbr[i]][dw.pushDst[i]], v);
        }
    }
    gather(veN);
    veN = (nIt > 1) ? load_ve(1) : 0xffffffffu;
    gather(veN);
    veN = (nIt > 2) ? load_ve(2) : 0xffffffffu;                  // vlist entry of tile it+2
    cp_async_wait_1();       // this thread's part of tile 0 has landed
    halo_check(0);
    __syncthreads();         // tile 0 staged

    // the next tile's group words (phase C) ...
    uint32_t mwAN = 0u, mwBN = 0u;
    auto load_groups_next = [&](int k) {
        if (k < nIt) { const uint32_t* m = meta_of(k); mwAN = __ldg(m + 4 + warp); if (TILE_NGROUPS > 8) mwBN = __ldg(m + 4 + (warp + 8) % TILE_NGROUPS); }
    };
    load_groups_next(1);
    // ... and the record offset and tet count of the tile whose record is fetched next (raw words: any arithmetic on
    // a loaded value here would stall the warp at the load)
    uint32_t off16N = 0u, ntvN = 0u;
    auto load_rec_entry = [&](int k) {
        if (k < nIt) { const uint32_t* m = meta_of(k); off16N = __ldg(m); ntvN = __ldg(m + 2); }
    };
    load_rec_entry(1);

    long long tk = 0, acc[7] = {0, 0, 0, 0, 0, 0, 0};
#define PD_TICK(i) if (PROF) { const long long now = clock64(); acc[i] += now - tk; tk = now; }
    if (PROF) tk = clock64();
    for (int it = 0; it < nIt; ++it) {
        const int b = it & 1;
        // ---- phase B of tile it
        if ((uint32_t)tid < (mwA >> 18))
            local_phase_b<ROT_MODE, JACOBI>(r0, r1, r2, smem + LOCAL_OFF_QS + b * LOCAL_QS_BYTES,
                                            smem + LOCAL_OFF_HS + b * LOCAL_HS_BYTES + 16 * (tid & ~7));
        PD_TICK(0)
        // the record registers are free: fetch the next tile's record (consumed after this tile's phase C)
        if (it + 1 < nIt) load_rec(off16N, ntvN & 0xffffu, r0, r1, r2);
        cp_async_wait_all();   // this thread's part of tile it+1's positions has landed
        if (it + 1 < nIt) halo_check(it + 1);
        PD_TICK(1)
        __syncthreads();       // H scratch of tile it complete; tile it+1 staged; every warp is past phase C of tile it-1
        PD_TICK(2)
        // staging buffer b and part C buffer (it+1)&1 are free now
        gather(veN);           // tile it+2 into buffer b (an empty group when there is none)
        PD_TICK(6)
        if (producer) {
            if (it >= 1 && it + 1 < nIt) fetch_c(it + 1);
            if (it + 3 < nIt) prefetch_ab(it + 3);
            if (it + 5 < nIt) prefetch_vs(it + 5);
        }
        veN = (it + 3 < nIt) ? load_ve(it + 3) : 0xffffffffu;
        PD_TICK(3)
        // ---- phase C of tile it: this warp sums vertex group `warp` (and `warp + 8`)
        mbar_wait(&cbar[b], (uint32_t)((it >> 1) & 1));
        PD_TICK(4)
        {
            const uint8_t* Cb = smem + LOCAL_OFF_C + b * TILE_CMAX;
            const uint8_t* Hb = smem + LOCAL_OFF_HS + b * LOCAL_HS_BYTES;
            const uint32_t slot0 = (blockIdx.x + it * gridDim.x) * (unsigned)TILE_NLMAX;
            local_phase_c<ROT_MODE == 1>(Cb, Hb, mwA, slot0 + (uint32_t)(TILE_GROUP * warp), lane, vlist, b0, P);
            if (TILE_NGROUPS > 8) local_phase_c<ROT_MODE == 1>(Cb, Hb, mwB, slot0 + (uint32_t)(TILE_GROUP * (warp + 8)), lane, vlist, b0, P);
        }
        PD_TICK(5)
        mwA = mwAN; mwB = mwBN;
        load_groups_next(it + 2);
        load_rec_entry(it + 2);
    }
    if (dw.pdlLate) pdl_launch_dependents();
    if (PROF && tid == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) prof[8 * blockIdx.x + i] = (unsigned long long)acc[i];
        prof[8 * blockIdx.x + 7] = (unsigned long long)nIt;
    }
#undef PD_TICK
}

// ------------------------------------------------------------------ global step (Jacobi + Chebyshev)
// computeDBCLocal + getErrorKern + chebyshevKern fused per vertex (pdUtil.cu:147-166,195-226).
// b = ordered sum of the vertex's partial-sum slots (the first one already carries (M/h^2) s_old).
// Arithmetic forms follow the reference's SASS: next = fma(-c,q,b)/(c+md) + q (IEEE division);
// under-relaxation as one DFMA (the reference's `0.9 *` literal is a double); Chebyshev as one FFMA.
// Multi-GPU: the halo push of the new positions happens at the start of the NEXT local kernel (k_local, DistWait), in
// push-list order; this kernel is the same on one and on many GPUs.
#ifndef PD_FOLD_CC
#define PD_FOLD_CC 1
#endif
#ifndef PD_VERTEX_MINBLOCKS
#define PD_VERTEX_MINBLOCKS 8      // resident 256-thread blocks per SM (32 registers): measured 66 vs 76 us on grid139 against 6 (40 registers)
#endif
// b = (b0 or the owner slot) + the vertex's partial-sum slots in ascending slot order.  The first four slot indices
// and then their four partial sums are requested together (two dependent load levels instead of one pair per slot:
// this kernel is latency bound); the ordered sum itself is unchanged bit for bit.
template <bool BASE>
__device__ __forceinline__ void vertex_slot_sum(int v, const float4* __restrict__ b0, const uint32_t* __restrict__ vslotPtr,
                                                const uint32_t* __restrict__ vslot, const float4* __restrict__ P,
                                                float& bx, float& by, float& bz)
{
    uint32_t e0 = vslotPtr[v];
    const uint32_t e1 = vslotPtr[v + 1];
    float4 first;
    if (BASE || e0 == e1) first = b0[v];                        // a vertex without tets keeps b = (M/h^2) s_old
    uint32_t i0 = 0xffffffffu, i1 = 0xffffffffu, i2 = 0xffffffffu, i3 = 0xffffffffu;
    if (e0 < e1) i0 = __ldg(&vslot[e0]);
    if (e0 + 1u < e1) i1 = __ldg(&vslot[e0 + 1u]);
    if (e0 + 2u < e1) i2 = __ldg(&vslot[e0 + 2u]);
    if (e0 + 3u < e1) i3 = __ldg(&vslot[e0 + 3u]);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 p0 = (i0 != 0xffffffffu) ? __ldg(&P[i0]) : z;
    const float4 p1 = (i1 != 0xffffffffu) ? __ldg(&P[i1]) : z;
    const float4 p2 = (i2 != 0xffffffffu) ? __ldg(&P[i2]) : z;
    const float4 p3 = (i3 != 0xffffffffu) ? __ldg(&P[i3]) : z;
    if (BASE || e0 == e1) {
        bx = first.x; by = first.y; bz = first.z;
        if (i0 != 0xffffffffu) { bx = __fadd_rn(bx, p0.x); by = __fadd_rn(by, p0.y); bz = __fadd_rn(bz, p0.z); }
    } else {
        bx = p0.x; by = p0.y; bz = p0.z;                         // faithful mode: the owner slot already starts from b0
    }
    if (i1 != 0xffffffffu) { bx = __fadd_rn(bx, p1.x); by = __fadd_rn(by, p1.y); bz = __fadd_rn(bz, p1.z); }
    if (i2 != 0xffffffffu) { bx = __fadd_rn(bx, p2.x); by = __fadd_rn(by, p2.y); bz = __fadd_rn(bz, p2.z); }
    if (i3 != 0xffffffffu) { bx = __fadd_rn(bx, p3.x); by = __fadd_rn(by, p3.y); bz = __fadd_rn(bz, p3.z); }
    for (uint32_t e = e0 + 4u; e < e1; ++e) {
        const float4 p = __ldg(&P[vslot[e]]);
        bx = __fadd_rn(bx, p.x); by = __fadd_rn(by, p.y); bz = __fadd_rn(bz, p.z);
    }
}

template <bool BASE, bool DRAG = false>    // BASE: the slots hold the elastic terms only and b0 is added here (product default);
                           // otherwise the vertex's first slot already starts from b0 (faithful mode).
                           // DRAG: a vertex with cc.y < 0 is being dragged and keeps its position (getErrorKern, pdUtil.cu:201-206)
__global__ void __launch_bounds__(256, PD_VERTEX_MINBLOCKS) k_vertex_jacobi(int nV, const float4* __restrict__ qcur, const float4* __restrict__ qprev,
                                float4* __restrict__ qnext, const float4* __restrict__ X0 /* DBCX */, const float4* __restrict__ b0,
                                const float2* __restrict__ cc, const uint32_t* __restrict__ vslotPtr,
                                const uint32_t* __restrict__ vslot, const float4* __restrict__ P,
                                float omega, float wdbc, int flags = 0, unsigned long long* epochBump = nullptr)
{
    // flags: bit 0 = PD_PDL=2 (see DistWait::pdlLate); bit 1 = blocks walk the vertices from the END of the array (the engine's
    // default: the slots the local kernel wrote last -- still in L2 -- are read first, and the positions this kernel writes last
    // are the ones the next local kernel gathers first; -0.4 % per step on grid139)
    const int pdlLate = flags & 1;
    const int v = (int)((flags & 2) ? gridDim.x - 1u - blockIdx.x : blockIdx.x) * (int)blockDim.x + (int)threadIdx.x;
    if (pdlLate == 0) pdl_launch_dependents();
    pdl_wait();              // the slots come from the local kernel just before
    if (epochBump && blockIdx.x == 0 && threadIdx.x == 0) *epochBump += 1ull;      // multi-GPU: one more phase (see k_predict)
    if (v < nV) {
        const float4 q = qcur[v], pr = qprev[v];
#if PD_FOLD_CC
        // (+-c, +-(c + matrix_diag)) from the w lanes of b0 and of the current iterate (k_predict); the faithful mode, whose owner
        // slot already holds b0, keeps reading cc
        const float2 c2 = BASE ? make_float2(b0[v].w, q.w) : cc[v];
#else
        const float2 c2 = cc[v];
#endif
        const float c = fabsf(c2.x);
        float bx, by, bz;
        if (c2.x < 0.f) {                 // computeDBCLocal overwrites b for pinned vertices; DBCX = X0 (pdSolver.cu:134)
            const float4 d = X0[v];
            bx = __fmul_rn(d.x, wdbc); by = __fmul_rn(d.y, wdbc); bz = __fmul_rn(d.z, wdbc);
        } else {
            vertex_slot_sum<BASE>(v, b0, vslotPtr, vslot, P, bx, by, bz);
        }
        const float den = DRAG ? fabsf(c2.y) : c2.y;
        float nx = __fadd_rn(__fdiv_rn(__fmaf_rn(-c, q.x, bx), den), q.x);
        float ny = __fadd_rn(__fdiv_rn(__fmaf_rn(-c, q.y, by), den), q.y);
        float nz = __fadd_rn(__fdiv_rn(__fmaf_rn(-c, q.z, bz), den), q.z);
        if (DRAG && c2.y < 0.f) { nx = q.x; ny = q.y; nz = q.z; }      // next_x = sn for a dragged vertex, then chebyshevKern as usual
        // under-relaxation in double, as the reference's `0.9 *` literal (pdUtil.cu:221)
        nx = (float)__fma_rn((double)__fsub_rn(nx, q.x), 0.9, (double)q.x);
        ny = (float)__fma_rn((double)__fsub_rn(ny, q.y), 0.9, (double)q.y);
        nz = (float)__fma_rn((double)__fsub_rn(nz, q.z), 0.9, (double)q.z);
        nx = __fmaf_rn(__fsub_rn(nx, pr.x), omega, pr.x);
        ny = __fmaf_rn(__fsub_rn(ny, pr.y), omega, pr.y);
        nz = __fmaf_rn(__fsub_rn(nz, pr.z), omega, pr.z);
        const float4 out = make_float4(nx, ny, nz, q.w);
        qnext[v] = out;
    }
    if (pdlLate) pdl_launch_dependents();
}

// right-hand side for the direct / CG global solves: b = c*s_old + sum of partials (R, not R-F)
template <bool BASE, bool DRAG = false>
__global__ void k_vertex_rhs(int nV, const float4* __restrict__ X0 /* DBCX */, const float4* __restrict__ b0, const float2* __restrict__ cc,
                             const uint32_t* __restrict__ vslotPtr, const uint32_t* __restrict__ vslot,
                             const float4* __restrict__ P, float wdbc, float4* __restrict__ rhs, DragArgs dr)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    float bx, by, bz;
    const float2 c2 = cc[v];
    if (c2.x < 0.f) {
        const float4 d = X0[v];
        bx = __fmul_rn(d.x, wdbc); by = __fmul_rn(d.y, wdbc); bz = __fmul_rn(d.z, wdbc);
    } else if (DRAG && c2.y < 0.f && dr.hasDBC) {
        // computeDBCLocal's second branch (pdUtil.cu:159-164; the kernel runs only when numDBC > 0): b = DBCX * moreDBC
        const float4 d = X0[v];
        const float mw = dr.more[v];
        bx = __fmul_rn(d.x, mw); by = __fmul_rn(d.y, mw); bz = __fmul_rn(d.z, mw);
    } else {
        vertex_slot_sum<BASE>(v, b0, vslotPtr, vslot, P, bx, by, bz);
    }
    rhs[v] = make_float4(bx, by, bz, 0.f);
}

// ------------------------------------------------------------------ end of step
// updateVelPos (pdUtil.cu:180-193), X <- XTilde (pdSolver.cu:227), then the three fixed-body
// kernels of FixedBodyData::HandleCollisions (fixedBodyData.cu:67-148) in their launch order.
// Operation forms (which products are fused) follow the SASS nvcc produces for those kernels.
__device__ __forceinline__ void fb_respond(float3& vel, const float3 n, float kf, float muN)
{   // kf = (1 + muN) * muT
    const float vn = dot3_nv(vel.x, n.x, vel.y, n.y, vel.z, n.z);
    const float3 vN = make_float3(__fmul_rn(vn, n.x), __fmul_rn(vn, n.y), __fmul_rn(vn, n.z));
    const float3 vT = make_float3(__fsub_rn(vel.x, vN.x), __fsub_rn(vel.y, vN.y), __fsub_rn(vel.z, vN.z));
    const float magT = __fsqrt_rn(dot3_nv(vT.x, vT.x, vT.y, vT.y, vT.z, vT.z));
    float a = 0.f;
    if (magT != 0.f) {
        const float magN = __fsqrt_rn(dot3_nv(vN.x, vN.x, vN.y, vN.y, vN.z, vN.z));
        a = fmaxf(__fsub_rn(1.f, __fdiv_rn(__fmul_rn(kf, magN), magT)), 0.f);
    }
    vel.x = __fmaf_rn(vT.x, a, -__fmul_rn(vN.x, muN));
    vel.y = __fmaf_rn(vT.y, a, -__fmul_rn(vN.y, muN));
    vel.z = __fmaf_rn(vT.z, a, -__fmul_rn(vN.z, muN));
}

// updateVelPos (pdUtil.cu:180-193): q = the final iterate, xt = XTilde of the previous step
__device__ __forceinline__ float3 finish_velocity(const float4 q, const float4 xt, float dtInv, bool dragged)
{
    if (dragged) return make_float3(0.f, 0.f, 0.f);      // pdUtil.cu:187-188
    return make_float3(__fmul_rn(__fsub_rn(q.x, xt.x), dtInv), __fmul_rn(__fsub_rn(q.y, xt.y), dtInv), __fmul_rn(__fsub_rn(q.z, xt.z), dtInv));
}
// FixedBodyData::HandleCollisions on one vertex (fixedBodyData.cu:67-148): spheres, planes, cylinders in that order
__device__ __forceinline__ void fixed_body_response(float3& x, float3& vel, const DevFixedBodies& fb, float muT, float muN)
{
    const float kf = __fmul_rn(__fadd_rn(muN, 1.f), muT);
    for (int j = 0; j < fb.nSpheres; ++j) {
        const float* s = fb.spheres + 4 * j;
        const float3 tc = make_float3(__fsub_rn(x.x, s[0]), __fsub_rn(x.y, s[1]), __fsub_rn(x.z, s[2]));
        const float d = __fsqrt_rn(dot3_nv(tc.x, tc.x, tc.y, tc.y, tc.z, tc.z));
        if (d < s[3]) {
            const float inv = __fdiv_rn(1.0f, d);
            const float3 n = make_float3(__fmul_rn(tc.x, inv), __fmul_rn(tc.y, inv), __fmul_rn(tc.z, inv));
            const float push = __fsub_rn(s[3], d);
            x.x = __fmaf_rn(n.x, push, x.x); x.y = __fmaf_rn(n.y, push, x.y); x.z = __fmaf_rn(n.z, push, x.z);
            fb_respond(vel, n, kf, muN);
        }
    }
    for (int j = 0; j < fb.nPlanes; ++j) {
        const float* p = fb.planes + 6 * j;
        const float3 up = make_float3(p[3], p[4], p[5]);
        const float sd = dot3_nv(__fsub_rn(x.x, p[0]), up.x, __fsub_rn(x.y, p[1]), up.y, __fsub_rn(x.z, p[2]), up.z);
        if (sd < 0.f && dot3_nv(vel.x, up.x, vel.y, up.y, vel.z, up.z) < 0.f) {
            x.x = __fmaf_rn(-up.x, sd, x.x); x.y = __fmaf_rn(-up.y, sd, x.y); x.z = __fmaf_rn(-up.z, sd, x.z);
            fb_respond(vel, up, kf, muN);
        }
    }
    for (int j = 0; j < fb.nCyls; ++j) {
        const float* c = fb.cyls + 7 * j;
        const float3 ax = make_float3(c[3], c[4], c[5]);
        const float3 rel = make_float3(__fsub_rn(x.x, c[0]), __fsub_rn(x.y, c[1]), __fsub_rn(x.z, c[2]));
        // n = (I - a a^T) rel as the reference's matrix-vector product
        float3 nn;
        nn.x = dot3_nv(__fmaf_rn(-ax.x, ax.x, 1.f), rel.x, __fmaf_rn(-ax.y, ax.x, 0.f), rel.y, __fmaf_rn(-ax.z, ax.x, 0.f), rel.z);
        nn.y = dot3_nv(__fmaf_rn(-ax.x, ax.y, 0.f), rel.x, __fmaf_rn(-ax.y, ax.y, 1.f), rel.y, __fmaf_rn(-ax.z, ax.y, 0.f), rel.z);
        nn.z = dot3_nv(__fmaf_rn(-ax.x, ax.z, 0.f), rel.x, __fmaf_rn(-ax.y, ax.z, 0.f), rel.y, __fmaf_rn(-ax.z, ax.z, 1.f), rel.z);
        const float d = __fsqrt_rn(dot3_nv(nn.x, nn.x, nn.y, nn.y, nn.z, nn.z));
        if (d < c[6]) {
            const float inv = __fdiv_rn(1.0f, d);
            const float3 n = make_float3(__fmul_rn(nn.x, inv), __fmul_rn(nn.y, inv), __fmul_rn(nn.z, inv));
            const float push = __fsub_rn(c[6], d);
            x.x = __fmaf_rn(n.x, push, x.x); x.y = __fmaf_rn(n.y, push, x.y); x.z = __fmaf_rn(n.z, push, x.z);
            fb_respond(vel, n, kf, muN);
        }
    }
}
// one vertex of the end of step without mesh-mesh collision: velocity, X <- XTilde (pdSolver.cu:227; X keeps the un-projected
// position), fixed bodies on (XTilde, V)
__device__ __forceinline__ void finish_vertex(const float4 q, const float4 xt, float dtInv, bool dragged, const DevFixedBodies& fb, float muT, float muN,
                                              float4* __restrict__ Xout, float4* __restrict__ XTout, float4* __restrict__ Vout)
{
    float3 vel = finish_velocity(q, xt, dtInv, dragged);
    float3 x = make_float3(q.x, q.y, q.z);
    *Xout = make_float4(x.x, x.y, x.z, 0.f);
    fixed_body_response(x, vel, fb, muT, muN);
    *XTout = make_float4(x.x, x.y, x.z, 0.f);
    *Vout = make_float4(vel.x, vel.y, vel.z, 0.f);
}

template <bool DRAG>
__global__ void k_finish(int nV, const float4* __restrict__ qfinal, float dtInv, float4* __restrict__ X,
                         float4* __restrict__ XTilde, float4* __restrict__ V, DevFixedBodies fb, float muT, float muN,
                         const float* __restrict__ more)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    finish_vertex(qfinal[v], XTilde[v], dtInv, DRAG && more[v] > 0.f, fb, muT, muN, &X[v], &XTilde[v], &V[v]);
}

// The same end of step around the mesh-mesh collision pass (SolverParams::handleCollision, pdSolver.cu:218-231):
// k_finish_velocity (updateVelPos: V, XTilde <- q; X still holds the step's start), the pass of pd_collision.cuh on
// (X, XTilde, V), then k_fixed_bodies.
template <bool DRAG>
__global__ void k_finish_velocity(int nV, const float4* __restrict__ qfinal, float dtInv, float4* __restrict__ XTilde, float4* __restrict__ V,
                                  const float* __restrict__ more)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 q = qfinal[v];
    const float3 vel = finish_velocity(q, XTilde[v], dtInv, DRAG && more[v] > 0.f);
    XTilde[v] = make_float4(q.x, q.y, q.z, 0.f);
    V[v] = make_float4(vel.x, vel.y, vel.z, 0.f);
}
__global__ void k_fixed_bodies(int nV, float4* __restrict__ XTilde, float4* __restrict__ V, DevFixedBodies fb, float muT, float muN)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 xt = XTilde[v], vv = V[v];
    float3 x = make_float3(xt.x, xt.y, xt.z), vel = make_float3(vv.x, vv.y, vv.z);
    fixed_body_response(x, vel, fb, muT, muN);
    XTilde[v] = make_float4(x.x, x.y, x.z, 0.f);
    V[v] = make_float4(vel.x, vel.y, vel.z, 0.f);
}

// ------------------------------------------------------------------ layout conversion
// AoS float3 (reference numbering) <-> padded float4 (renumbered): newOfOld / oldOfNew permutations
__global__ void k_import3(int nV, const float* __restrict__ src3, const uint32_t* __restrict__ oldOfNew, float4* __restrict__ dst)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float* s = src3 + 3ull * (oldOfNew ? oldOfNew[v] : (uint32_t)v);      // no map: compact shard in local order
    dst[v] = make_float4(s[0], s[1], s[2], 0.f);
}
__global__ void k_export3(int nV, const float4* __restrict__ src, const uint32_t* __restrict__ oldOfNew, float* __restrict__ dst3)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float4 s = src[v];
    float* d = dst3 + 3ull * (oldOfNew ? oldOfNew[v] : (uint32_t)v);
    d[0] = s.x; d[1] = s.y; d[2] = s.z;
}

// scalar per-vertex array (reference numbering) -> renumbered; *anyPositive |= (value > 0)
__global__ void k_import1(int nV, const float* __restrict__ src, const uint32_t* __restrict__ oldOfNew, float* __restrict__ dst, int* __restrict__ anyPositive)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const float a = src[oldOfNew ? oldOfNew[v] : (uint32_t)v];
    dst[v] = a;
    if (a > 0.f) atomicOr(anyPositive, 1);
}
// Control_Kernel (simulationContext.cu:202-218; RADIUS_SQUARED 0.002 is a double literal, :18): vertices within the
// radius of the selected one (sel, renumbered id) get stiffness controlMag, and every free vertex its offset
__global__ void k_drag_select(int nV, const float4* __restrict__ X, const float* __restrict__ dbc, int sel, float controlMag,
                              float* __restrict__ more, float4* __restrict__ offX, int* __restrict__ anyPositive)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    float stiffness = 0.f;
    if (dbc[v] == 0.f && sel >= 0) {
        const float4 a = X[v], b = X[sel];
        const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
        offX[v] = make_float4(dx, dy, dz, 0.f);
        if ((double)dot3_nv(dx, dx, dy, dy, dz, dz) < 0.002) stiffness = controlMag;
    }
    more[v] = stiffness;
    if (stiffness > 0.f) atomicOr(anyPositive, 1);
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
