// Engine implementation (see pd_engine.hpp).
#include "pd_engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "collision.hpp"
#include "pd_kernels.cuh"
#include "pd_body_kernel.cuh"
#include "pd_collision.cuh"
#include "pd_solvers.cuh"

namespace pdb200 {

#define CUDA_CHECK(call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorName(e__) + " (" +          \
                                     cudaGetErrorString(e__) + ") at " + __FILE__ + ":" +                 \
                                     std::to_string(__LINE__) + ": " #call);                              \
    } while (0)

// kernel launch with (optionally) the programmatic-dependent-launch attribute: the kernel may be scheduled while
// the previous kernel of the stream drains; it calls pdl_wait() before touching anything that kernel wrote
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// Exchange window of one rank (nLoc = owned + ghost vertices), mapped by its peers:
//   [q0 | q1 | q2 | pcgP | (unused: the round-1 halo flags)[world] | epoch | ticket, status | pflags[world] | red[2][world][4 doubles]]
// (the Jacobi path's halo needs no flags: every pushed position carries its phase tag, DistWait in pd_kernels.cuh)
struct WindowLayout {
    size_t offP, offFlags, offEpoch, offTicket, offPFlags, offRed, bytes;
    WindowLayout(size_t nLoc, int world)
    {
        offP = 3 * nLoc * 16; offFlags = 4 * nLoc * 16; offEpoch = offFlags + 8 * (size_t)world; offTicket = offEpoch + 8;
        offPFlags = offEpoch + 16; offRed = offPFlags + 8 * (size_t)world; bytes = offRed + 2 * (size_t)world * 32;
    }
};

struct Engine::Impl {
    // tile stream
    uint8_t* records = nullptr;
    uint32_t* tileTab = nullptr;      // device tile table, TILE_META_WORDS words per tile (layout.hpp:build_tile_table)
    uint32_t *vslotPtr = nullptr, *vslot = nullptr, *vlist = nullptr, *vstage = nullptr;
    float4* P = nullptr;             // one partial RHS sum per (tile, tile-local vertex) slot
    // per-vertex state (renumbered, padded float4)
    float4* q[3] = {nullptr, nullptr, nullptr};
    float4 *b0 = nullptr, *X = nullptr, *V = nullptr, *XT = nullptr, *X0 = nullptr;
    float2* cc = nullptr;
    float *mass = nullptr, *dbc = nullptr, *md = nullptr;
    // mouse drag (allocated by the first setDrag*): moreDBC, OffsetX, a DBCX of its own (until then DBCX is X0) and the
    // "some moreDBC > 0" flag
    float* more = nullptr; float4* offX = nullptr; float4* dbcx = nullptr; int* dragFlag = nullptr;
    // EXPERIMENT PD_BODY_KERNEL=1 (pd_body_kernel.cuh): one CTA per small body, one launch per step
    BodyDesc* bBodies = nullptr; uint32_t* bVerts = nullptr; uint8_t* bRec = nullptr; uint32_t* bIncPtr = nullptr; uint16_t* bInc = nullptr; float* bMd = nullptr;
    int nBodies = 0; uint32_t bNVmax = 0, bNTmax = 0; size_t bSmem = 0;
    unsigned long long* bPerfNs = nullptr;
    DragArgs drag(const float t[3], int numDBC) const { return DragArgs{more, offX, dbcx, t[0], t[1], t[2], numDBC > 0 ? 1 : 0}; }
    uint32_t* oldOfNew = nullptr;
    float* stage3 = nullptr;          // 3 x (3 nV) floats, AoS staging for import/export
    float* fbData = nullptr;
    DevFixedBodies fb{};
    cudaGraphExec_t graphExec[3] = {nullptr, nullptr, nullptr};   // one per rotation base of the position buffers
    std::vector<cudaEvent_t> events;
    std::vector<void*> allocs;
    std::vector<float> hostMd;        // renumbered
    std::vector<float> hostWPrepared; // w = |V0| mu of every tet (tile order) as SolverPrepare saw it: the system matrix of the
                                      // direct / CG modes is assembled from THESE even if mu was edited since (pd_update_mu)
    // multi-GPU exchange window (this rank's) and the tables that point into the neighbours' windows
    uint8_t* window = nullptr;
    size_t windowBytes = 0;
    unsigned long long* epoch = nullptr;       // this rank's phase count = the tag of its halo pushes
    unsigned int *ticket = nullptr, *status = nullptr;
    uint32_t *pushSrc = nullptr, *pushDst = nullptr, *pushNbr = nullptr;
    int* nbrRanks = nullptr;
    float4** peerQ = nullptr;                  // [3 * nNbr]: buffer k of neighbour j
    int nPush = 0, nNbr = 0;
    // distributed PCG (pd_solvers.cuh: DistSolve): p lives in the window; flags / reduction slots / sequence counters
    unsigned long long *pflags = nullptr, *solveSeq = nullptr;
    double* red = nullptr;
    float4** peerP = nullptr; unsigned long long** peerPFlag = nullptr; double** peerRed = nullptr;
    std::vector<void*> ipcOpened;
    cudaEvent_t lockEvent = nullptr, callerEvent = nullptr;
    // non-Jacobi global solvers (PCG, sparse Cholesky)
    bool solverReady = false, cholReady = false;
    CsrDev A{0, nullptr, nullptr, nullptr, nullptr};
    CholDev C{0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float4 *rhs = nullptr, *cgR = nullptr, *cgP = nullptr, *cgQ = nullptr, *xprev = nullptr, *cholY = nullptr, *cholZ = nullptr;
    int cholEpoch = 0;
    SolveState* solveState = nullptr;
    double* partials = nullptr;
    int solveGrid = 0, cholGrid = 0;
    size_t nnzA = 0, nnzL = 0;
    DistWait wait{nullptr, 0, 0, 0x7fffffff, nullptr};
    // mesh-mesh collision pass (pd_collision.cuh), built by the first step that asks for it
    bool colReady = false;
    ColMeshDev cm{};
    uint2* colPairs = nullptr; unsigned int* colCount = nullptr; unsigned int colMaxPairs = 0;
    unsigned long long *colVf = nullptr, *colEe = nullptr;
    int* colWriter = nullptr; float* colTI = nullptr; float4* colNors = nullptr;
    cudaEvent_t colEv[2] = {nullptr, nullptr};
    long long colPairsLast = 0, colHitsLast = 0;
};

// give a dalloc'ed buffer back (re-prepared system matrix / factor after Reset)
void Engine::dfree(const void* p, size_t bytes)
{
    if (!p) return;
    auto& a = d_->allocs;
    auto it = std::find(a.begin(), a.end(), const_cast<void*>(p));
    if (it != a.end()) a.erase(it);
    cudaFree(const_cast<void*>(p));
    devBytes_ -= std::min(devBytes_, bytes);
}

template <typename T>
T* Engine::dalloc(size_t n)
{
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CUDA_CHECK(cudaMalloc(&p, bytes));
    d_->allocs.push_back(p);
    devBytes_ += bytes;
    return static_cast<T*>(p);
}

Engine::Engine(const Scene& scene, const EngineOptions& opt) : scene_(scene), params_(scene.params), opt_(opt), d_(new Impl)
{
    if (scene.numVerts <= 0 || scene.numTets <= 0) throw std::runtime_error("empty scene");
    if (opt.world < 1 || opt.rank < 0 || opt.rank >= opt.world) throw std::runtime_error("bad rank/world");
    if (params_.handleCollision && opt.world > 1)
        throw std::runtime_error("handleCollision=true (mesh-mesh collision) needs the whole surface on one GPU: single-GPU engines only");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("no CUDA device: the PD engine has no CPU fallback");
    CUDA_CHECK(cudaSetDevice(opt.device));
    cudaDeviceProp prop{};
    CUDA_CHECK(cudaGetDeviceProperties(&prop, opt.device));
    if (prop.major < 10)
        throw std::runtime_error("this build targets sm_100a (B200); found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
    numSms_ = prop.multiProcessorCount;
    if (const char* e = std::getenv("PD_PDL")) { usePdl_ = std::atoi(e) != 0; pdlLate_ = std::atoi(e) == 2 ? 1 : 0; }      // experiments: PD_PDL=1 / 2
    pdlActive_ = usePdl_;
    if (const char* e = std::getenv("PD_VERTEX_REVERSE")) vertexFlags_ = std::atoi(e) != 0 ? 2 : 0;      // (A/B runs; default on)
    CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));

    if (opt_.rotMode < 0) {
        // auto: a mesh that fits one tile is launch-latency bound whatever the rotation costs, and the reference's
        // float atomics are deterministic at that size -- reproduce it bit for bit (faithful SVD, input tet order)
        const bool oneTile = opt_.world == 1 && scene_.numTets <= TILE_T && scene_.numVerts <= TILE_NLMAX;
        opt_.rotMode = oneTile ? 1 : 0;
        if (oneTile) opt_.reorder = 0;
    }
    if (opt.world == 1) {
        build_layout(scene_.numVerts, scene_.numTets, scene_.X.data(), scene_.Tet.data(), scene_.mu.data(), opt_.reorder != 0, L_);
        nOwn_ = L_.nV;
    } else {
        // every rank builds the same global layout, then keeps the tiles that touch its vertex range
        Layout G;
        build_layout(scene_.numVerts, scene_.numTets, scene_.X.data(), scene_.Tet.data(), scene_.mu.data(), opt_.reorder != 0, G);
        build_rank_plan(G, opt.world, opt.rank, plan_, dist_trim_from_env());
        extract_rank_layout(G, plan_, L_);      // plan_.trim (PD_DIST_TRIM=1): trimmed + packed boundary tiles (experiment, layout.hpp)
        nOwn_ = plan_.nOwn;
    }
    nV_ = L_.nV;
    nT_ = L_.nT;
    if (nV_ <= 0 || nT_ <= 0) throw std::runtime_error("rank " + std::to_string(opt.rank) + " holds no tets: fewer ranks, or a larger mesh");

    // ---- device buffers
    Impl& d = *d_;
    d.records = dalloc<uint8_t>(L_.records.size());
    d.tileTab = dalloc<uint32_t>(L_.tileTab.size() * TILE_META_WORDS);
    d.vslotPtr = dalloc<uint32_t>(L_.vslotPtr.size());
    d.vslot = dalloc<uint32_t>(L_.vslot.size());
    d.vlist = dalloc<uint32_t>(L_.vlist.size());
    d.vstage = dalloc<uint32_t>(L_.vstage.size());
    d.P = dalloc<float4>((size_t)L_.nTiles * TILE_NLMAX);      // padded slots: tile * TILE_NLMAX + local vertex
    // the three position buffers (and the CG direction p) live in the exchange window, see WindowLayout
    const WindowLayout wl((size_t)nV_, opt.world);
    d.windowBytes = wl.bytes;
    d.window = dalloc<uint8_t>(d.windowBytes);
    CUDA_CHECK(cudaMemset(d.window, 0, d.windowBytes));
    for (int k = 0; k < 3; ++k) d.q[k] = reinterpret_cast<float4*>(d.window) + (size_t)k * nV_;
    d.cgP = reinterpret_cast<float4*>(d.window + wl.offP);
    d.epoch = reinterpret_cast<unsigned long long*>(d.window + wl.offEpoch);
    d.ticket = reinterpret_cast<unsigned int*>(d.window + wl.offTicket);
    d.status = d.ticket + 1;
    d.pflags = reinterpret_cast<unsigned long long*>(d.window + wl.offPFlags);
    d.red = reinterpret_cast<double*>(d.window + wl.offRed);
    d.solveSeq = dalloc<unsigned long long>(2);
    CUDA_CHECK(cudaMemset(d.solveSeq, 0, 16));
    d.b0 = dalloc<float4>(nV_); d.X = dalloc<float4>(nV_); d.V = dalloc<float4>(nV_);
    d.XT = dalloc<float4>(nV_); d.X0 = dalloc<float4>(nV_);
    d.cc = dalloc<float2>(nV_);
    d.mass = dalloc<float>(nV_); d.dbc = dalloc<float>(nV_); d.md = dalloc<float>(nV_);
    d.oldOfNew = dalloc<uint32_t>(nV_);
    d.stage3 = dalloc<float>(9 * (size_t)scene_.numVerts);
    d.dbcx = d.X0;                          // DBCX <- X0 (pdSolver.cu:134); a drag gives it storage of its own
    for (float f : scene_.DBC) if (f > 0.f) ++numDBC_;

    CUDA_CHECK(cudaMemcpy(d.records, L_.records.data(), L_.records.size(), cudaMemcpyHostToDevice));
    static_assert(sizeof(TileEntry) == sizeof(uint4), "tile table entry layout");
    {   // device tile table (layout.cpp:build_tile_table)
        std::vector<uint32_t> meta;
        build_tile_table(L_, meta);
        CUDA_CHECK(cudaMemcpy(d.tileTab, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
    }
    CUDA_CHECK(cudaMemcpy(d.vslotPtr, L_.vslotPtr.data(), L_.vslotPtr.size() * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d.vslot, L_.vslot.data(), L_.vslot.size() * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d.vlist, L_.vlist.data(), L_.vlist.size() * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d.vstage, L_.vstage.data(), L_.vstage.size() * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d.oldOfNew, L_.vertOrder.data(), (size_t)nV_ * 4, cudaMemcpyHostToDevice));
    {
        std::vector<float> m(nV_), b(nV_);
        std::vector<float4> x0(nV_);
        for (int v = 0; v < nV_; ++v) {
            const uint32_t o = L_.vertOrder[v];
            m[v] = scene_.mass[o]; b[v] = scene_.DBC[o];
            x0[v] = make_float4(scene_.X[3 * (size_t)o], scene_.X[3 * (size_t)o + 1], scene_.X[3 * (size_t)o + 2], 0.f);
        }
        CUDA_CHECK(cudaMemcpy(d.mass, m.data(), (size_t)nV_ * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.dbc, b.data(), (size_t)nV_ * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.X0, x0.data(), (size_t)nV_ * 16, cudaMemcpyHostToDevice));
    }
    // fixed bodies -> the three arrays the collision kernels consume
    {
        std::vector<float> planes, spheres, cyls;
        for (const FixedBody& f : scene_.fixed) {
            if (f.type == FB_PLANE) {
                float up[3]; plane_up(f.model, up);
                planes.insert(planes.end(), {f.model[12], f.model[13], f.model[14], up[0], up[1], up[2]});
            } else if (f.type == FB_SPHERE) {
                spheres.insert(spheres.end(), {f.model[12], f.model[13], f.model[14], f.radius});
            } else if (f.type == FB_CYLINDER) {
                float ax[3]; cylinder_axis(f.model, ax);
                cyls.insert(cyls.end(), {f.model[12], f.model[13], f.model[14], ax[0], ax[1], ax[2], f.radius});
            }
        }
        const size_t n = planes.size() + spheres.size() + cyls.size();
        d.fbData = dalloc<float>(n);
        std::vector<float> all;
        all.insert(all.end(), planes.begin(), planes.end());
        all.insert(all.end(), spheres.begin(), spheres.end());
        all.insert(all.end(), cyls.begin(), cyls.end());
        if (n) CUDA_CHECK(cudaMemcpy(d.fbData, all.data(), n * 4, cudaMemcpyHostToDevice));
        d.fb.nPlanes = (int)(planes.size() / 6); d.fb.nSpheres = (int)(spheres.size() / 4); d.fb.nCyls = (int)(cyls.size() / 7);
        d.fb.planes = d.fbData; d.fb.spheres = d.fbData + planes.size(); d.fb.cyls = d.fbData + planes.size() + spheres.size();
    }
    // local kernel geometry: fixed shared-memory carve-up (pd_kernels.cuh), persistent CTAs
    auto setAttr = [&](const void* fn) {
        CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOCAL_SMEM_BYTES));
    };
    setAttr((const void*)k_local<0, true>); setAttr((const void*)k_local<0, false>);
    setAttr((const void*)k_local<1, true>); setAttr((const void*)k_local<1, false>);
    setAttr((const void*)k_local<2, true>); setAttr((const void*)k_local<0, true, true>); setAttr((const void*)k_local<2, true, true>);
    int perSm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_local<0, true>, TILE_T, LOCAL_SMEM_BYTES));
    if (perSm < 1) throw std::runtime_error("local kernel does not fit on an SM");
    if (opt.ctasPerSm > 0) perSm = std::min(perSm, opt.ctasPerSm);
    localGrid_ = std::min(L_.nTiles, numSms_ * perSm);

    // Scenes made of SMALL bodies (every connected component fits one CTA's shared memory: BASELINE config 5, the cube, house +
    // sphere) step with one CTA per body and ONE launch per step (pd_body_kernel.cuh) instead of 2 + 2 * iterations launches
    // that cannot fill the GPU: batch64 2.31 -> 1.43 ms per step (profiles/r2_body_kernel_ab_batch64.txt).  PD_BODY_KERNEL=0 keeps the tile path (A/B runs).
    bodyKernel_ = opt.world == 1 && opt_.rotMode != 2 && opt_.bodyKernel != 0;
    if (const char* e = std::getenv("PD_BODY_KERNEL")) bodyKernel_ = bodyKernel_ && std::atoi(e) != 0;
    std::vector<int> bodyStarts;
    if (bodyKernel_) {
        int maxOptin = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, opt.device));
        bodyKernel_ = connected_body_ranges(scene_.numVerts, scene_.numTets, scene_.Tet.data(), bodyStarts);
        if (bodyKernel_) {      // cheap size check before anything is built: the largest component must fit
            std::vector<int> tetsOf(bodyStarts.size(), 0);
            for (int t = 0; t < scene_.numTets; ++t)
                ++tetsOf[(size_t)(std::upper_bound(bodyStarts.begin(), bodyStarts.end(), (int)scene_.Tet[4 * (size_t)t]) - bodyStarts.begin()) - 1];
            for (size_t b = 0; b < bodyStarts.size() && bodyKernel_; ++b) {
                const int nVb = ((b + 1 < bodyStarts.size()) ? bodyStarts[b + 1] : scene_.numVerts) - bodyStarts[b];
                bodyKernel_ = body_smem_bytes(((uint32_t)nVb + 31u) & ~31u, ((uint32_t)tetsOf[b] + 31u) & ~31u) <= (size_t)maxOptin;
            }
        }
    }
    if (bodyKernel_) {
        try {
            BodyBatch bb;
            build_body_batch(scene_.numVerts, scene_.numTets, scene_.X.data(), scene_.Tet.data(), scene_.mu.data(), bodyStarts, L_.vertNewOfOld.data(), bb);
            int maxOptin = 0;
            CUDA_CHECK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, opt.device));
            d.bSmem = body_smem_bytes(bb.nVmax, bb.nTmax);
            if (d.bSmem > (size_t)maxOptin) throw std::runtime_error("a body needs " + std::to_string(d.bSmem) + " bytes of shared memory");
            d.nBodies = (int)bb.bodies.size(); d.bNVmax = bb.nVmax; d.bNTmax = bb.nTmax;
            d.bBodies = dalloc<BodyDesc>(bb.bodies.size()); d.bVerts = dalloc<uint32_t>(bb.verts.size()); d.bRec = dalloc<uint8_t>(bb.rec.size());
            d.bIncPtr = dalloc<uint32_t>(bb.incPtr.size()); d.bInc = dalloc<uint16_t>(bb.inc.size()); d.bMd = dalloc<float>(bb.md.size());
            CUDA_CHECK(cudaMemcpy(d.bBodies, bb.bodies.data(), bb.bodies.size() * sizeof(BodyDesc), cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(d.bVerts, bb.verts.data(), bb.verts.size() * 4, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(d.bRec, bb.rec.data(), bb.rec.size(), cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(d.bIncPtr, bb.incPtr.data(), bb.incPtr.size() * 4, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(d.bInc, bb.inc.data(), bb.inc.size() * 2, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(d.bMd, bb.md.data(), bb.md.size() * 4, cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaFuncSetAttribute((const void*)k_body_step<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.bSmem));
            CUDA_CHECK(cudaFuncSetAttribute((const void*)k_body_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.bSmem));
        } catch (const std::exception& ex) {
            if (std::getenv("PD_BODY_KERNEL")) std::fprintf(stderr, "pd_b200: per-body kernel not used for this scene (%s); the tile path runs\n", ex.what());
            bodyKernel_ = false;
        }
    }

    reset();
}

Engine::~Engine()
{
    if (d_) {
        cudaSetDevice(opt_.device);
        if (stream_) cudaStreamSynchronize(stream_);
        for (cudaGraphExec_t g : d_->graphExec) if (g) cudaGraphExecDestroy(g);
        for (cudaEvent_t ev : d_->events) cudaEventDestroy(ev);
        if (d_->lockEvent) cudaEventDestroy(d_->lockEvent);
        for (cudaEvent_t ev : d_->colEv) if (ev) cudaEventDestroy(ev);
        if (d_->callerEvent) cudaEventDestroy(d_->callerEvent);
        for (void* p : d_->ipcOpened) cudaIpcCloseMemHandle(p);
        for (void* p : d_->allocs) cudaFree(p);
        if (stream_) cudaStreamDestroy(stream_);
    }
}

void Engine::synchronize()
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Engine::setParams(const SolverParams& p)
{
    if (p.handleCollision && opt_.world > 1) throw std::runtime_error("handleCollision=true (mesh-mesh collision) is for single-GPU engines");
    const SolverParams& o = params_;
    const bool same = o.handleCollision == p.handleCollision && o.dt == p.dt && o.gravity == p.gravity && o.muN == p.muN && o.muT == p.muT && o.rho == p.rho &&
                      o.tol == p.tol && o.numIterations == p.numIterations && o.globalSolver == p.globalSolver &&
                      o.pcgMaxIter == p.pcgMaxIter && o.pcgTol == p.pcgTol;
    params_ = p;
    if (!same) graphValid_ = false;
}

void Engine::reset()
{   // SimulationCUDAContext::Reset (simulationContext.cu:233-243): X = XTilde = X0, V = 0, re-prepare
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    CUDA_CHECK(cudaMemcpyAsync(d.X, d.X0, (size_t)nV_ * 16, cudaMemcpyDeviceToDevice, stream_));
    CUDA_CHECK(cudaMemcpyAsync(d.XT, d.X0, (size_t)nV_ * 16, cudaMemcpyDeviceToDevice, stream_));
    CUDA_CHECK(cudaMemsetAsync(d.V, 0, (size_t)nV_ * 16, stream_));
    if (d.more) {       // cudaMemset(moreDBC, 0) (simulationContext.cu:240); the next SolverPrepare restores DBCX <- X0 (pdSolver.cu:134)
        CUDA_CHECK(cudaMemsetAsync(d.more, 0, (size_t)nV_ * 4, stream_));
        CUDA_CHECK(cudaMemcpyAsync(d.dbcx, d.X0, (size_t)nV_ * 16, cudaMemcpyDeviceToDevice, stream_));
    }
    dragActive_ = false;
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    ready_ = false;
    perfc_ = PerfCounters();
}

void Engine::prepare()
{   // PdSolver::SolverPrepare (pdSolver.cu:40-139): matrix_diag and the dt baked into DBC rows.
    // matrix_diag[v] = sum over incident tets (ascending reordered order) of w_t |col_i(B^T G)|^2
    Impl& d = *d_;
    d.hostWPrepared.clear(); d.hostWPrepared.reserve((size_t)nT_);
    for (int ti = 0; ti < L_.nTiles; ++ti) {
        const uint8_t* rec = L_.records.data() + L_.tileRecOff[ti];
        TileHeader h; std::memcpy(&h, rec, sizeof(h));
        for (uint32_t t = 0; t < h.nTets; ++t) { float w; std::memcpy(&w, rec + tile_tet_word(h.nTets, t, 9), 4); d.hostWPrepared.push_back(w); }
    }
    matrix_diag_host(L_, d.hostMd);        // layout.cpp: ascending GLOBAL tet order, so the float sums do not depend on the world size
    CUDA_CHECK(cudaMemcpyAsync(d.md, d.hostMd.data(), (size_t)nV_ * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    dt2Prepared_ = params_.dt * params_.dt;
    ready_ = true;
    graphValid_ = false;
    d.solverReady = false; d.cholReady = false;      // the system matrix bakes in dt and mu
}

// pushBuf >= 0 (multi-GPU): q is position buffer number pushBuf and the kernel first pushes its boundary entries to
// the neighbours (DistWait in pd_kernels.cuh); -1: no push inside the kernel
void Engine::launchLocal(const float4* q, bool jacobi, unsigned long long* prof, int pushBuf, bool checkHalo)
{
    Impl& d = *d_;
    DistWait w = d.wait;
    if (!checkHalo) w.nNbr = 0;          // stand-alone timing launches on a buffer nobody pushed into: no tag to look for
    w.pdlLate = pdlLate_;
    if (pushBuf >= 0 && opt_.world > 1) { w.nPush = d.nPush; w.peerQ = d.peerQ + (size_t)pushBuf * d.nNbr; }
#define PD_LOCAL(RM, JAC) launch_pdl(k_local<RM, JAC>, dim3(localGrid_), dim3(TILE_T), LOCAL_SMEM_BYTES, stream_, pdlActive_, d.records, d.tileTab, L_.nTiles, d.vstage, d.vlist, q, d.b0, d.P, prof, w)
    if (prof) {
        if (opt_.rotMode == 2) k_local<2, true, true><<<localGrid_, TILE_T, LOCAL_SMEM_BYTES, stream_>>>(d.records, d.tileTab, L_.nTiles, d.vstage, d.vlist, q, d.b0, d.P, prof, w);
        else k_local<0, true, true><<<localGrid_, TILE_T, LOCAL_SMEM_BYTES, stream_>>>(d.records, d.tileTab, L_.nTiles, d.vstage, d.vlist, q, d.b0, d.P, prof, w);
    } else if (jacobi) {
        if (opt_.rotMode == 0) PD_LOCAL(0, true);
        else if (opt_.rotMode == 1) PD_LOCAL(1, true);
        else PD_LOCAL(2, true);
    } else {
        if (opt_.rotMode == 1) PD_LOCAL(1, false);
        else PD_LOCAL(0, false);
    }
#undef PD_LOCAL
}

void Engine::enqueuePredict()
{
    Impl& d = *d_;
    const SolverParams& p = params_;
    const int vb = 256, vg = (nOwn_ + vb - 1) / vb;
    base_ = (opt_.world > 1) ? (int)(phase_ % 3) : 0;
    omega_ = 1.0f;
    // multi-GPU: every kernel that produces new positions counts one phase (the tag of the halo push, DistWait); in the
    // lock-step test mode k_halo_push does that itself
    unsigned long long* bump = (opt_.world > 1 && !lockstep_) ? d.epoch : nullptr;
    if (dragActive_)
        k_predict<true><<<vg, vb, 0, stream_>>>(nOwn_, d.X, d.V, d.mass, d.dbc, d.md, p.dt, dt2Prepared_, p.gravity,
                                                d.q[base_], d.q[(base_ + 2) % 3], d.b0, d.cc, d.drag(dragTarget_, numDBC_), bump);
    else
        k_predict<false><<<vg, vb, 0, stream_>>>(nOwn_, d.X, d.V, d.mass, d.dbc, d.md, p.dt, dt2Prepared_, p.gravity,
                                                 d.q[base_], d.q[(base_ + 2) % 3], d.b0, d.cc, DragArgs{}, bump);
    if (lockstep_) enqueuePush(d.q[base_], base_);      // otherwise the first local kernel pushes its input itself
    ++phase_;
}

void Engine::enqueueIteration(int i, bool timed, size_t* ev)
{
    Impl& d = *d_;
    const SolverParams& p = params_;
    const int vb = 256, vg = (nOwn_ + vb - 1) / vb;
    const float dtInv = 1.0f / p.dt;
    const float wdbc = 1e6f * (dtInv * dtInv);
    auto rec = [&]() { if (timed) CUDA_CHECK(cudaEventRecord(d.events[(*ev)++], stream_)); };
    const int ic = (base_ + i) % 3, in = (base_ + i + 1) % 3, ip = (base_ + i + 2) % 3;
    const float4* cur = d.q[ic];
    const float4* prev = d.q[ip];
    float4* next = d.q[in];
    rec();
    // multi-GPU: the local kernel first pushes the boundary entries of the buffer it reads (`cur`) to the neighbours,
    // spread over its CTAs and in push-list order, every position tagged with this phase's count E (DistWait), and checks the
    // tags of the ghosts it stages for its boundary tiles.  Why no acknowledgement is needed (three rotating buffers): this
    // launch overwrites the neighbour's ghost entries of buffer E % 3, which that neighbour last READ in its local kernel of
    // phase E - 3.  This rank can only be here after its vertex kernel of phase E - 1, hence after its local kernel of phase
    // E - 1 consumed the neighbour's push of phase E - 1, which the neighbour issued at the start of ITS local kernel E - 1,
    // i.e. after it had finished phase E - 2 (stream order) -- so phase E - 3 is long over.  The same argument covers the step
    // boundary (finish + predictor instead of a vertex kernel).
    if (opt_.world > 1 && !connected_) throw std::runtime_error("multi-GPU engine stepped before pd_dist_connect");
    launchLocal(cur, true, nullptr, (opt_.world > 1 && !lockstep_) ? ic : -1);
    rec();
    // omega recurrence in float, pdSolver.cu:196-198
    if (i <= 10) omega_ = 1;
    else if (i == 11) omega_ = 2 / (2 - p.rho * p.rho);
    else omega_ = 4 / (4 - p.rho * p.rho * omega_);
    const float4* dbcx = d.dbcx;
    unsigned long long* bump = (opt_.world > 1 && !lockstep_) ? d.epoch : nullptr;      // one more phase (see enqueuePredict)
    if (dragActive_) {      // dragged vertices (cc.y < 0) keep their position (getErrorKern, pdUtil.cu:201-206)
        if (opt_.rotMode == 1) k_vertex_jacobi<false, true><<<vg, vb, 0, stream_>>>(nOwn_, cur, prev, next, dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, omega_, wdbc);
        else k_vertex_jacobi<true, true><<<vg, vb, 0, stream_>>>(nOwn_, cur, prev, next, dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, omega_, wdbc);
    }
    else if (opt_.rotMode == 1) launch_pdl(k_vertex_jacobi<false>, dim3(vg), dim3(vb), 0, stream_, pdlActive_, nOwn_, cur, prev, next, dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, omega_, wdbc, pdlLate_ | vertexFlags_, bump);
    else launch_pdl(k_vertex_jacobi<true>, dim3(vg), dim3(vb), 0, stream_, pdlActive_, nOwn_, cur, prev, next, dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, omega_, wdbc, pdlLate_ | vertexFlags_, bump);
    if (lockstep_) enqueuePush(next, in);
    ++phase_;
    rec();
}

void Engine::enqueueFinish() { enqueueEnd(d_->q[(base_ + params_.numIterations) % 3]); }

// updateVelPos, [mesh-mesh collision], X <- XTilde, fixed bodies (pdSolver.cu:206, 218-231)
void Engine::enqueueEnd(const float4* qfinal)
{
    Impl& d = *d_;
    const SolverParams& p = params_;
    const int vb = 256, vg = (nOwn_ + vb - 1) / vb;
    const float dtInv = 1.0f / p.dt;
    if (!p.handleCollision) {
        if (dragActive_) k_finish<true><<<vg, vb, 0, stream_>>>(nOwn_, qfinal, dtInv, d.X, d.XT, d.V, d.fb, p.muT, p.muN, d.more);
        else k_finish<false><<<vg, vb, 0, stream_>>>(nOwn_, qfinal, dtInv, d.X, d.XT, d.V, d.fb, p.muT, p.muN, nullptr);
        return;
    }
    if (dragActive_) k_finish_velocity<true><<<vg, vb, 0, stream_>>>(nOwn_, qfinal, dtInv, d.XT, d.V, d.more);
    else k_finish_velocity<false><<<vg, vb, 0, stream_>>>(nOwn_, qfinal, dtInv, d.XT, d.V, nullptr);
    collisionPass();
    k_fixed_bodies<<<vg, vb, 0, stream_>>>(nOwn_, d.XT, d.V, d.fb, p.muT, p.muN);
}

// ---------------------------------------------------------------- mesh-mesh collision (pd_collision.cuh)
__global__ void k_col_reset(int nV, int nEdges, int nInternal, float* tI, int* writer, unsigned long long* vf, unsigned long long* ee, int* visit, unsigned int* count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nV) { tI[i] = 1.0f; writer[i] = 0; vf[i] = COL_EMPTY; }       // thrust::fill(tI, 1.0f), bvh.cu:173-174
    if (i < nEdges) ee[i] = COL_EMPTY;
    if (i < nInternal) visit[i] = 0;
    if (i == 0) *count = 0u;
}

void Engine::prepareCollision()
{
    Impl& d = *d_;
    if (d.colReady) return;
    CollisionMesh M;
    build_collision_mesh(scene_, M);
    ColMeshDev& c = d.cm;
    c.nTris = M.nTris; c.nEdges = M.nEdges; c.nInternal = M.nInternal; c.nBodies = M.nBodies; c.nV = nV_;
    auto up = [&](const auto& v) {
        using T = typename std::decay<decltype(v)>::type::value_type;
        T* p = dalloc<T>(v.size());
        if (!v.empty()) CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    };
    c.tri = up(M.tri); c.father = up(M.father); c.edge = up(M.edge); c.triEdge = up(M.triEdge);
    c.left = up(M.left); c.right = up(M.right); c.parent = up(M.parent); c.bodyRoot = up(M.bodyRoot);
    c.newOfOld = up(L_.vertNewOfOld);
    c.boxMin = dalloc<float4>((size_t)M.nInternal + M.nTris); c.boxMax = dalloc<float4>((size_t)M.nInternal + M.nTris);
    c.visit = dalloc<int>((size_t)M.nInternal);
    d.colCount = dalloc<unsigned int>(1);
    d.colMaxPairs = (unsigned int)std::max<size_t>(1u << 12, 4 * (size_t)M.nTris);      // (the reference starts its query buffer at 2^15 and doubles it)
    d.colPairs = dalloc<uint2>(d.colMaxPairs);
    d.colVf = dalloc<unsigned long long>((size_t)nV_); d.colEe = dalloc<unsigned long long>((size_t)M.nEdges);
    d.colWriter = dalloc<int>((size_t)nV_); d.colTI = dalloc<float>((size_t)nV_); d.colNors = dalloc<float4>((size_t)nV_);
    CUDA_CHECK(cudaMemset(d.colNors, 0, (size_t)nV_ * 16));
    for (cudaEvent_t& ev : d.colEv) CUDA_CHECK(cudaEventCreate(&ev));
    d.colReady = true;
}

// DetectCollision + CCDKernel (pdSolver.cu:218-225) on (X = the step's start, XTilde = its end, V).  One host
// synchronisation per step (the pair count; the reference reads its query count back the same way, broadphase.cu:415-429).
void Engine::collisionPass()
{
    Impl& d = *d_;
    prepareCollision();
    const ColMeshDev& c = d.cm;
    const int tb = 128;
    if (perf_) CUDA_CHECK(cudaEventRecord(d.colEv[0], stream_));
    const int nReset = std::max(std::max(nV_, c.nEdges), std::max(c.nInternal, 1));
    k_col_reset<<<(nReset + 255) / 256, 256, 0, stream_>>>(nV_, c.nEdges, c.nInternal, d.colTI, d.colWriter, d.colVf, d.colEe, c.visit, d.colCount);
    unsigned int nPairs = 0;
    if (c.nTris > 0) {
        k_col_refit<<<(c.nTris + tb - 1) / tb, tb, 0, stream_>>>(c, d.X, d.XT);
        for (;;) {
            k_col_traverse<<<(c.nTris + tb - 1) / tb, tb, 0, stream_>>>(c, /*ignoreSelfCollision=*/1, d.colPairs, d.colCount, d.colMaxPairs);
            CUDA_CHECK(cudaMemcpyAsync(&nPairs, d.colCount, 4, cudaMemcpyDeviceToHost, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            if (nPairs <= d.colMaxPairs) break;
            // more overlapping pairs than the buffer holds: grow it and list them again (nothing has been modified yet)
            dfree(d.colPairs, (size_t)d.colMaxPairs * sizeof(uint2));
            while (d.colMaxPairs < nPairs) d.colMaxPairs *= 2;
            d.colPairs = dalloc<uint2>(d.colMaxPairs);
            CUDA_CHECK(cudaMemsetAsync(d.colCount, 0, 4, stream_));
        }
    }
    d.colPairsLast = nPairs;
    if (nPairs > 0) {       // BroadPhaseCCD returned queries: NarrowPhase + storeTi (bvh.cu:175-178)
        const int grid = (int)std::min<size_t>((12 * (size_t)nPairs + tb - 1) / tb, (size_t)numSms_ * 16);
        k_col_narrow<<<grid, tb, 0, stream_>>>(c, d.X, d.XT, d.colPairs, d.colCount, d.colMaxPairs, d.colVf, d.colEe);
        const int nG = nV_ + c.nEdges;
        k_col_rank<<<(nG + tb - 1) / tb, tb, 0, stream_>>>(c, d.colVf, d.colEe, d.colWriter);
        k_col_store<<<(nG + tb - 1) / tb, tb, 0, stream_>>>(c, d.X, d.XT, d.colVf, d.colEe, d.colWriter, d.colTI, d.colNors);
    }
    k_col_apply<<<(nV_ + 255) / 256, 256, 0, stream_>>>(nV_, d.X, d.XT, d.V, d.colTI, d.colNors);
    CUDA_CHECK(cudaGetLastError());
    if (perf_) {
        CUDA_CHECK(cudaEventRecord(d.colEv[1], stream_));
        CUDA_CHECK(cudaEventSynchronize(d.colEv[1]));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, d.colEv[0], d.colEv[1]));
        perfc_.collisionMesh += ms;
    }
}

// tI (1 = free, 0.5 = in a detected contact) and the contact normals of the last collision pass, caller's numbering
void Engine::getCollision(float* tI, float* normals, long long* numPairs)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    if (numPairs) *numPairs = d.colPairsLast;
    std::vector<float> t((size_t)nV_, 1.0f);
    std::vector<float4> n((size_t)nV_, make_float4(0.f, 0.f, 0.f, 0.f));
    if (d.colReady) {
        CUDA_CHECK(cudaMemcpy(t.data(), d.colTI, (size_t)nV_ * 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(n.data(), d.colNors, (size_t)nV_ * 16, cudaMemcpyDeviceToHost));
    }
    for (int v = 0; v < nV_; ++v) {
        const size_t g = L_.vertOrder[(size_t)v];
        if (tI) tI[g] = t[(size_t)v];
        if (normals) { normals[3 * g] = n[(size_t)v].x; normals[3 * g + 1] = n[(size_t)v].y; normals[3 * g + 2] = n[(size_t)v].z; }
    }
}

// multi-GPU: boundary positions of buffer `bufIndex` -> the neighbours' ghost entries, tagged with the new phase count
void Engine::enqueuePush(const float4* q, int bufIndex)
{
    Impl& d = *d_;
    if (opt_.world == 1) return;
    if (!connected_) throw std::runtime_error("multi-GPU engine stepped before pd_dist_connect");
    const int grid = std::max(1, std::min(numSms_, (d.nPush + 255) / 256));
    k_halo_push<<<grid, 256, 0, stream_>>>(d.nPush, d.pushSrc, d.pushDst, d.pushNbr, q, d.peerQ + (size_t)bufIndex * d.nNbr, d.epoch, d.ticket);
}

// The launch sequence of one PdSolver::Update in Jacobi mode.  `timed` brackets the local /
// global / collision parts with events (perf mode, no host sync inside the loop).
void Engine::enqueueStep(bool timed)
{
    Impl& d = *d_;
    size_t ev = 0;
    pdlActive_ = usePdl_ && !timed && !dragActive_;        // event records between the kernels (perf mode) want fully serialised launches
    enqueuePredict();
    for (int i = 0; i < params_.numIterations; ++i) enqueueIteration(i, timed, &ev);
    if (timed) CUDA_CHECK(cudaEventRecord(d.events[ev++], stream_));
    enqueueFinish();
    if (timed) CUDA_CHECK(cudaEventRecord(d.events[ev++], stream_));
    pdlActive_ = usePdl_;
}

// One CUDA graph per step.  Multi-GPU: the position buffers rotate across steps (base = phase_ % 3), so up to
// three graphs exist, one per base; the phase tags make the replayed kernels wait for the neighbours as usual.
void Engine::buildGraph()
{
    Impl& d = *d_;
    if (!graphValid_) {
        for (cudaGraphExec_t& g : d.graphExec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
        graphValid_ = true;
    }
    const int b = (opt_.world > 1) ? (int)(phase_ % 3) : 0;
    if (d.graphExec[b]) return;
    const long long phaseSaved = phase_;
    cudaGraph_t graph = nullptr;
    if (opt_.world > 1 && !connected_) throw std::runtime_error("multi-GPU engine stepped before pd_dist_connect");
    CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    try {
        enqueueStep(false);
    } catch (...) {                       // never leave the stream in capture mode
        cudaStreamEndCapture(stream_, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        phase_ = phaseSaved;
        throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(stream_, &graph));
    phase_ = phaseSaved;                  // capturing ran nothing
    CUDA_CHECK(cudaGraphInstantiate(&d.graphExec[b], graph, 0));
    cudaGraphDestroy(graph);
}

void Engine::step(int nSteps)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!ready_) prepare();
    Impl& d = *d_;
    if (params_.globalSolver != 0) {       // PCG-Jacobi / sparse Cholesky global step: plain launches, device-side early exit
        prepareSolver();
        if (perf_) {        // local step (+ right-hand side) | global solve, bracketed by events; one host sync per step
            const size_t need = 3 * (size_t)params_.numIterations + 2;
            while (d.events.size() < need) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); d.events.push_back(e); }
        }
        for (int s = 0; s < nSteps; ++s) {
            enqueueStepSolver(perf_);
            if (perf_) {
                CUDA_CHECK(cudaStreamSynchronize(stream_));
                for (int i = 0; i < params_.numIterations; ++i) {
                    float a = 0, b = 0;
                    CUDA_CHECK(cudaEventElapsedTime(&a, d.events[3 * (size_t)i], d.events[3 * (size_t)i + 1]));
                    CUDA_CHECK(cudaEventElapsedTime(&b, d.events[3 * (size_t)i + 1], d.events[3 * (size_t)i + 2]));
                    perfc_.localStep += a; perfc_.globalStep += b;
                }
                float c = 0;
                CUDA_CHECK(cudaEventElapsedTime(&c, d.events[3 * (size_t)params_.numIterations], d.events[3 * (size_t)params_.numIterations + 1]));
                perfc_.collisionFixed += c;
            }
        }
        CUDA_CHECK(cudaGetLastError());
        perfc_.steps += nSteps;
        perfc_.kernelLaunches += (long long)nSteps * (4 + 3 * params_.numIterations);
        return;
    }
    if (bodyKernel_ && !dragActive_ && !params_.handleCollision) {      // small bodies: one CTA per body, the whole step in one launch (pd_body_kernel.cuh)
        const SolverParams& p = params_;
        const float dtInv = 1.0f / p.dt, wdbc = 1e6f * (dtInv * dtInv);
        unsigned long long* perfNs = nullptr;
        if (perf_) {       // the four counters come from body 0's own clock (nothing else can split a single launch)
            if (!d.bPerfNs) { d.bPerfNs = dalloc<unsigned long long>(4); CUDA_CHECK(cudaMemsetAsync(d.bPerfNs, 0, 32, stream_)); }
            perfNs = d.bPerfNs;
        }
        for (int s = 0; s < nSteps; ++s) {
            if (opt_.rotMode == 1)
                k_body_step<1><<<d.nBodies, 512, d.bSmem, stream_>>>(d.bBodies, d.bVerts, d.bRec, d.bIncPtr, d.bInc, d.bMd, d.bNVmax, d.bNTmax, d.X, d.V, d.XT, d.mass,
                                                                     d.dbc, d.dbcx, p.dt, dt2Prepared_, p.gravity, p.numIterations, p.rho, wdbc, d.fb, p.muT, p.muN, perfNs);
            else
                k_body_step<0><<<d.nBodies, 512, d.bSmem, stream_>>>(d.bBodies, d.bVerts, d.bRec, d.bIncPtr, d.bInc, d.bMd, d.bNVmax, d.bNTmax, d.X, d.V, d.XT, d.mass,
                                                                     d.dbc, d.dbcx, p.dt, dt2Prepared_, p.gravity, p.numIterations, p.rho, wdbc, d.fb, p.muT, p.muN, perfNs);
        }
        CUDA_CHECK(cudaGetLastError());
        if (perf_) {
            unsigned long long ns[4];
            CUDA_CHECK(cudaMemcpyAsync(ns, d.bPerfNs, 32, cudaMemcpyDeviceToHost, stream_));
            CUDA_CHECK(cudaMemsetAsync(d.bPerfNs, 0, 32, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            perfc_.localStep += (float)(ns[0] * 1e-6); perfc_.globalStep += (float)(ns[1] * 1e-6); perfc_.collisionFixed += (float)(ns[2] * 1e-6);
        }
        perfc_.steps += nSteps;
        perfc_.pdIterations += (long long)nSteps * p.numIterations;
        perfc_.kernelLaunches += nSteps;
        return;
    }
    const int launchesPerStep = 2 + 2 * params_.numIterations;     // multi-GPU: the halo pushes ride inside the local kernels
    if (perf_) {
        const size_t need = 3 * (size_t)params_.numIterations + 2;
        while (d.events.size() < need) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); d.events.push_back(e); }
        for (int s = 0; s < nSteps; ++s) {
            const float meshBefore = perfc_.collisionMesh;
            enqueueStep(true);
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            size_t ev = 0;
            for (int i = 0; i < params_.numIterations; ++i) {
                float a = 0, b = 0;
                CUDA_CHECK(cudaEventElapsedTime(&a, d.events[ev], d.events[ev + 1]));
                CUDA_CHECK(cudaEventElapsedTime(&b, d.events[ev + 1], d.events[ev + 2]));
                perfc_.localStep += a; perfc_.globalStep += b;
                ev += 3;
            }
            float c = 0;
            CUDA_CHECK(cudaEventElapsedTime(&c, d.events[ev], d.events[ev + 1]));
            perfc_.collisionFixed += std::max(0.f, c - (perfc_.collisionMesh - meshBefore));      // (the bracket holds the mesh pass too)
        }
    } else if (opt_.useGraph && !dragActive_ && !params_.handleCollision) {     // (a drag moves its target every frame, the collision pass
                                                                                 // reads its pair count back: plain launches)
        for (int s = 0; s < nSteps; ++s) {
            buildGraph();
            CUDA_CHECK(cudaGraphLaunch(d.graphExec[(opt_.world > 1) ? (int)(phase_ % 3) : 0], stream_));
            phase_ += params_.numIterations + 1;
        }
    } else {
        for (int s = 0; s < nSteps; ++s) enqueueStep(false);
    }
    CUDA_CHECK(cudaGetLastError());
    perfc_.steps += nSteps;
    perfc_.pdIterations += (long long)nSteps * params_.numIterations;
    perfc_.kernelLaunches += (long long)nSteps * launchesPerStep;
}

float Engine::stepTimed(int nSteps)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!ready_) prepare();
    if (opt_.useGraph && !perf_ && !dragActive_ && !bodyKernel_ && !params_.handleCollision && params_.globalSolver == 0) buildGraph();
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaEventRecord(a, stream_));
    step(nSteps);
    CUDA_CHECK(cudaEventRecord(b, stream_));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    perfc_.stepMsTotal += ms;
    return ms;
}

// ---------------------------------------------------------------- state transfer
// The engine's stream is non-blocking, i.e. NOT ordered against the caller's legacy default stream, on which the
// reference writes the SolverData arrays asynchronously (Control_Kernel in ResetMoreDBC, the D2D copies of Reset, the
// BVH / CCD kernels): every entry point that reads CALLER DEVICE memory first makes the stream wait for whatever the
// legacy stream has been given so far.
void Engine::waitCallerStream()
{
    Impl& d = *d_;
    if (!d.callerEvent) CUDA_CHECK(cudaEventCreateWithFlags(&d.callerEvent, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(d.callerEvent, cudaStreamLegacy));
    CUDA_CHECK(cudaStreamWaitEvent(stream_, d.callerEvent, 0));
}

void Engine::importDevice(const float* dX, const float* dV, const float* dXTilde)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    waitCallerStream();
    Impl& d = *d_;
    const int vb = 256, vg = (nV_ + vb - 1) / vb;
    if (dX) k_import3<<<vg, vb, 0, stream_>>>(nV_, dX, d.oldOfNew, d.X);
    if (dV) k_import3<<<vg, vb, 0, stream_>>>(nV_, dV, d.oldOfNew, d.V);
    if (dXTilde) k_import3<<<vg, vb, 0, stream_>>>(nV_, dXTilde, d.oldOfNew, d.XT);
    CUDA_CHECK(cudaGetLastError());
}

void Engine::exportDevice(float* dX, float* dV, float* dXTilde)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    const int vb = 256, vg = (nV_ + vb - 1) / vb;
    // multi-GPU: owned vertices only (the caller combines the ranks' disjoint contributions)
    if (dX) k_export3<<<vg, vb, 0, stream_>>>(nOwn_, d.X, d.oldOfNew, dX);
    if (dV) k_export3<<<vg, vb, 0, stream_>>>(nOwn_, d.V, d.oldOfNew, dV);
    if (dXTilde) k_export3<<<vg, vb, 0, stream_>>>(nOwn_, d.XT, d.oldOfNew, dXTilde);
    CUDA_CHECK(cudaGetLastError());
}

void Engine::upload(const float* X, const float* V, const float* XTilde)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    const size_t nG = (size_t)scene_.numVerts, n = 3 * nG * sizeof(float);
    float *sx = d.stage3, *sv = d.stage3 + 3 * nG, *st = d.stage3 + 6 * nG;
    if (X) CUDA_CHECK(cudaMemcpyAsync(sx, X, n, cudaMemcpyHostToDevice, stream_));
    if (V) CUDA_CHECK(cudaMemcpyAsync(sv, V, n, cudaMemcpyHostToDevice, stream_));
    if (XTilde) CUDA_CHECK(cudaMemcpyAsync(st, XTilde, n, cudaMemcpyHostToDevice, stream_));
    importDevice(X ? sx : nullptr, V ? sv : nullptr, XTilde ? st : nullptr);
}

void Engine::download(float* X, float* V, float* XTilde)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    const size_t nG = (size_t)scene_.numVerts, n = 3 * nG * sizeof(float);
    float *sx = d.stage3, *sv = d.stage3 + 3 * nG, *st = d.stage3 + 6 * nG;
    if (opt_.world > 1) CUDA_CHECK(cudaMemsetAsync(d.stage3, 0, 3 * n, stream_));     // vertices of other ranks read as 0
    exportDevice(X ? sx : nullptr, V ? sv : nullptr, XTilde ? st : nullptr);
    if (X) CUDA_CHECK(cudaMemcpyAsync(X, sx, n, cudaMemcpyDeviceToHost, stream_));
    if (V) CUDA_CHECK(cudaMemcpyAsync(V, sv, n, cudaMemcpyDeviceToHost, stream_));
    if (XTilde) CUDA_CHECK(cudaMemcpyAsync(XTilde, st, n, cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Engine::stepHost(int nSteps, const float* Xin, const float* Vin, const float* XTin, float* Xout, float* Vout, float* XTout)
{
    upload(Xin, Vin, XTin);
    step(nSteps);
    download(Xout, Vout, XTout);
}

// Multi-GPU end to end: every rank moves only ITS shard -- 3 * numOwned() floats per array, in the rank's local
// owned-vertex order (ownedIds() gives the original vertex id of each entry).  On one GPU the shard is the whole
// state in the engine's renumbered order.
void Engine::stepHostOwned(int nSteps, const float* Xin, const float* Vin, const float* XTin, float* Xout, float* Vout, float* XTout)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    const size_t nO = (size_t)nOwn_, n = 3 * nO * sizeof(float);
    float *sx = d.stage3, *sv = d.stage3 + 3 * nO, *st = d.stage3 + 6 * nO;
    const int vb = 256, vg = (nOwn_ + vb - 1) / vb;
    if (Xin) { CUDA_CHECK(cudaMemcpyAsync(sx, Xin, n, cudaMemcpyHostToDevice, stream_)); k_import3<<<vg, vb, 0, stream_>>>(nOwn_, sx, nullptr, d.X); }
    if (Vin) { CUDA_CHECK(cudaMemcpyAsync(sv, Vin, n, cudaMemcpyHostToDevice, stream_)); k_import3<<<vg, vb, 0, stream_>>>(nOwn_, sv, nullptr, d.V); }
    if (XTin) { CUDA_CHECK(cudaMemcpyAsync(st, XTin, n, cudaMemcpyHostToDevice, stream_)); k_import3<<<vg, vb, 0, stream_>>>(nOwn_, st, nullptr, d.XT); }
    step(nSteps);
    if (Xout) { k_export3<<<vg, vb, 0, stream_>>>(nOwn_, d.X, nullptr, sx); CUDA_CHECK(cudaMemcpyAsync(Xout, sx, n, cudaMemcpyDeviceToHost, stream_)); }
    if (Vout) { k_export3<<<vg, vb, 0, stream_>>>(nOwn_, d.V, nullptr, sv); CUDA_CHECK(cudaMemcpyAsync(Vout, sv, n, cudaMemcpyDeviceToHost, stream_)); }
    if (XTout) { k_export3<<<vg, vb, 0, stream_>>>(nOwn_, d.XT, nullptr, st); CUDA_CHECK(cudaMemcpyAsync(XTout, st, n, cudaMemcpyDeviceToHost, stream_)); }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Engine::ownedIds(uint32_t* out) const
{
    std::memcpy(out, L_.vertOrder.data(), (size_t)nOwn_ * sizeof(uint32_t));
}

void Engine::getSetup(float* matrixDiag, float* massDt2, float* DmInv, float* V0)
{
    if (!ready_) prepare();
    const float dt2 = params_.dt * params_.dt;
    for (int v = 0; v < nV_; ++v) {
        const uint32_t o = L_.vertOrder[v];
        if (matrixDiag) matrixDiag[o] = d_->hostMd[v];
        if (massDt2) massDt2[o] = (scene_.mass[o] + scene_.DBC[o] * 1e6f) / dt2;
    }
    if (DmInv || V0) {
        std::vector<float> B((size_t)scene_.numTets * 9), v0((size_t)scene_.numTets);
        rest_shape(scene_.X.data(), scene_.Tet.data(), scene_.numTets, B.data(), v0.data());
        if (DmInv) std::memcpy(DmInv, B.data(), B.size() * 4);
        if (V0) std::memcpy(V0, v0.data(), v0.size() * 4);
    }
}

// ---------------------------------------------------------------- mouse-drag soft constraints
// SolverData<float>::moreDBC / OffsetX / mouseSelection.target (def.h:14-18,31-32).  The reference fills them with
// Control_Kernel (simulationContext.cu:202-218) between two Updates; PdSolver reads them at pdUtil.cu:56-69,80-87,
// 159-164,187-188,201-206.  While some moreDBC is positive the step runs the DRAG instantiations of the per-vertex
// kernels as plain launches; the local kernel does not change.
void Engine::ensureDragBuffers()
{
    Impl& d = *d_;
    if (opt_.world > 1) throw std::runtime_error("mouse-drag constraints on a partitioned mesh are outside the PD hot path (single-GPU engines only)");
    if (d.more) return;
    d.more = dalloc<float>(nV_); d.offX = dalloc<float4>(nV_); d.dbcx = dalloc<float4>(nV_); d.dragFlag = dalloc<int>(1);
    graphValid_ = false;                    // the captured launches hold the old DBCX pointer (X0)
    CUDA_CHECK(cudaMemsetAsync(d.more, 0, (size_t)nV_ * 4, stream_));
    CUDA_CHECK(cudaMemsetAsync(d.offX, 0, (size_t)nV_ * 16, stream_));
    CUDA_CHECK(cudaMemcpyAsync(d.dbcx, d.X0, (size_t)nV_ * 16, cudaMemcpyDeviceToDevice, stream_));
}

void Engine::finishDragUpdate(const float target[3])
{   // the stream has just written `more` and or-ed the flag
    int any = 0;
    CUDA_CHECK(cudaMemcpyAsync(&any, d_->dragFlag, 4, cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaGetLastError());
    dragActive_ = any != 0;
    if (target) std::memcpy(dragTarget_, target, 12);
}

void Engine::setDrag(const float* more, const float* offsetX, const float target[3])
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    if (!more) {                              // ResetMoreDBC(true)
        if (d.more) CUDA_CHECK(cudaMemsetAsync(d.more, 0, (size_t)nV_ * 4, stream_));
        dragActive_ = false;
        return;
    }
    ensureDragBuffers();
    const size_t nG = (size_t)scene_.numVerts;
    const int vb = 256, vg = (nV_ + vb - 1) / vb;
    float *so = d.stage3, *sm = d.stage3 + 3 * nG;
    CUDA_CHECK(cudaMemsetAsync(d.dragFlag, 0, 4, stream_));
    CUDA_CHECK(cudaMemcpyAsync(sm, more, nG * 4, cudaMemcpyHostToDevice, stream_));
    k_import1<<<vg, vb, 0, stream_>>>(nV_, sm, d.oldOfNew, d.more, d.dragFlag);
    if (offsetX) {
        CUDA_CHECK(cudaMemcpyAsync(so, offsetX, nG * 12, cudaMemcpyHostToDevice, stream_));
        k_import3<<<vg, vb, 0, stream_>>>(nV_, so, d.oldOfNew, d.offX);
    }
    finishDragUpdate(target);
}

void Engine::setDragDevice(const float* dMore, const float* dOffsetX, const float target[3])
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    if (!dMore) { setDrag(nullptr, nullptr, nullptr); return; }
    ensureDragBuffers();
    waitCallerStream();
    const int vb = 256, vg = (nV_ + vb - 1) / vb;
    CUDA_CHECK(cudaMemsetAsync(d.dragFlag, 0, 4, stream_));
    k_import1<<<vg, vb, 0, stream_>>>(nV_, dMore, d.oldOfNew, d.more, d.dragFlag);
    if (dOffsetX) k_import3<<<vg, vb, 0, stream_>>>(nV_, dOffsetX, d.oldOfNew, d.offX);
    finishDragUpdate(target);
}

void Engine::dragSelect(int selectV, float controlMag, const float target[3])
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    if (selectV < -1 || selectV >= scene_.numVerts) throw std::runtime_error("dragSelect: vertex index out of range");
    ensureDragBuffers();
    int sel = -1;
    if (selectV >= 0) sel = (int)L_.vertNewOfOld[(size_t)selectV];
    const int vb = 256, vg = (nV_ + vb - 1) / vb;
    CUDA_CHECK(cudaMemsetAsync(d.dragFlag, 0, 4, stream_));
    k_drag_select<<<vg, vb, 0, stream_>>>(nV_, d.X, d.dbc, sel, controlMag, d.more, d.offX, d.dragFlag);
    finishDragUpdate(target);
}

void Engine::getDrag(float* more, float* offsetX, float* dbcx)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    Impl& d = *d_;
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    std::vector<float> m((size_t)nV_, 0.f);
    std::vector<float4> o((size_t)nV_, make_float4(0.f, 0.f, 0.f, 0.f)), x((size_t)nV_);
    if (d.more) {
        CUDA_CHECK(cudaMemcpy(m.data(), d.more, (size_t)nV_ * 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(o.data(), d.offX, (size_t)nV_ * 16, cudaMemcpyDeviceToHost));
    }
    CUDA_CHECK(cudaMemcpy(x.data(), d.dbcx, (size_t)nV_ * 16, cudaMemcpyDeviceToHost));
    for (int v = 0; v < nV_; ++v) {
        const size_t g = L_.vertOrder[(size_t)v];
        if (more) more[g] = m[(size_t)v];
        if (offsetX) { offsetX[3 * g] = o[(size_t)v].x; offsetX[3 * g + 1] = o[(size_t)v].y; offsetX[3 * g + 2] = o[(size_t)v].z; }
        if (dbcx) { dbcx[3 * g] = x[(size_t)v].x; dbcx[3 * g + 1] = x[(size_t)v].y; dbcx[3 * g + 2] = x[(size_t)v].z; }
    }
}

// ---------------------------------------------------------------- live stiffness edit
// The tile stream carries w = |V0| * mu per tet (what computeLocal multiplies by, pdUtil.cu:124): rewrite that word of
// every record on the host copy and upload the stream again.  matrix_diag (hostMd) and the assembled system matrix are
// deliberately left alone -- the reference rebuilds them only in SolverPrepare, i.e. after Reset() (pdSolver.cu:212-216).
void Engine::updateMu(const float* mu)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!mu) throw std::runtime_error("updateMu: mu is NULL");
    Impl& d = *d_;
    std::vector<float> B((size_t)scene_.numTets * 9), v0((size_t)scene_.numTets);
    rest_shape(scene_.X.data(), scene_.Tet.data(), scene_.numTets, B.data(), v0.data());
    std::copy(mu, mu + scene_.numTets, scene_.mu.begin());
    bodyKernel_ = false;                               // (the per-body experiment keeps its own records: back to the tile path)
    CUDA_CHECK(cudaStreamSynchronize(stream_));        // no launch may still be reading the stream
    for (int ti = 0; ti < L_.nTiles; ++ti) {
        uint8_t* rec = L_.records.data() + L_.tileRecOff[(size_t)ti];
        TileHeader h; std::memcpy(&h, rec, sizeof(h));
        const uint32_t t0 = L_.tileTetStart[(size_t)ti];
        for (uint32_t tl = 0; tl < h.nTets; ++tl) {
            const size_t o = L_.tetOrder[(size_t)t0 + tl];                      // original tet id (also in a rank's layout)
            const float w = std::fabs(v0[o]) * scene_.mu[o];                    // as build_layout computes it
            std::memcpy(rec + tile_tet_word(h.nTets, tl, 9), &w, 4);
        }
    }
    CUDA_CHECK(cudaMemcpy(d.records, L_.records.data(), L_.records.size(), cudaMemcpyHostToDevice));
}

void Engine::updateMuDevice(const float* dMu)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!dMu) throw std::runtime_error("updateMuDevice: mu is NULL");
    std::vector<float> mu((size_t)scene_.numTets);
    CUDA_CHECK(cudaMemcpy(mu.data(), dMu, mu.size() * 4, cudaMemcpyDeviceToHost));
    updateMu(mu.data());
}

// ---------------------------------------------------------------- kernel timing helpers (bench)
float Engine::timeLocalKernelMs(int reps)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!ready_) prepare();
    Impl& d = *d_;
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaEventRecord(a, stream_));
    pdlActive_ = false;                    // isolated launches: no overlap with the neighbouring launch
    for (int r = 0; r < reps; ++r) {
        launchLocal(d.XT, true, nullptr, -1, false);
    }
    pdlActive_ = usePdl_;
    CUDA_CHECK(cudaEventRecord(b, stream_));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / (float)reps;
}

float Engine::timeVertexKernelMs(int reps)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!ready_) prepare();
    Impl& d = *d_;
    const int vb = 256, vg = (nOwn_ + vb - 1) / vb;
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    // a valid so4/cc is needed: run the predictor once
    k_predict<false><<<vg, vb, 0, stream_>>>(nOwn_, d.X, d.V, d.mass, d.dbc, d.md, params_.dt, dt2Prepared_, params_.gravity,
                                             d.q[0], d.q[2], d.b0, d.cc, DragArgs{});
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaEventRecord(a, stream_));
    for (int r = 0; r < reps; ++r)
        k_vertex_jacobi<true><<<vg, vb, 0, stream_>>>(nOwn_, d.q[0], d.q[2], d.q[1], d.dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, 1.0f, 1.0f);
    CUDA_CHECK(cudaEventRecord(b, stream_));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / (float)reps;
}

// ---------------------------------------------------------------- PCG / Cholesky global step
// A^ = diag(c) + sum_t w_t P^T (B^T G)^T (B^T G) P  (SolverPrepare's COO, pdSolver.cu:62-77, deduplicated on the host)
void Engine::prepareSolver()
{
    Impl& d = *d_;
    if (opt_.world > 1 && params_.globalSolver == 1)
        throw std::runtime_error("the sparse Cholesky global step does not shard (replicas only, one mesh per GPU); use Jacobi or PCG on a partitioned mesh");
    if (opt_.world > 1 && !connected_) throw std::runtime_error("multi-GPU engine stepped before pd_dist_connect");
    if (!d.solverReady) {
        std::vector<float> Br((size_t)nT_ * 9), wr((size_t)nT_), c((size_t)nV_);
        size_t t = 0;
        for (int ti = 0; ti < L_.nTiles; ++ti) {
            const uint8_t* rec = L_.records.data() + L_.tileRecOff[ti];
            TileHeader h; std::memcpy(&h, rec, sizeof(h));
            for (uint32_t k = 0; k < h.nTets; ++k, ++t) {
                float B[12];
                for (uint32_t j = 0; j < 12; ++j) std::memcpy(&B[j], rec + tile_tet_word(h.nTets, k, j), 4);
                std::memcpy(&Br[9 * t], B, 36);
                wr[t] = d.hostWPrepared[t];             // setup-time stiffness (computeSiTSi runs in SolverPrepare only, pdSolver.cu:62-70)
            }
        }
        const float dt2 = params_.dt * params_.dt;
        for (int v = 0; v < nV_; ++v) {
            const uint32_t o = L_.vertOrder[v];
            c[(size_t)v] = (scene_.mass[o] + scene_.DBC[o] * 1e6f) / dt2;       // setMDt_2, pdUtil.cu:48
        }
        CsrMatrix A; std::vector<float> md;
        build_system_matrix(L_, nullptr, Br.data(), wr.data(), c.data(), A, md);
        if (opt_.world > 1) {      // keep this rank's rows only (every tet of an owned vertex is here, so they are complete)
            A.n = nOwn_;
            A.rowPtr.resize((size_t)nOwn_ + 1);
            A.col.resize((size_t)A.rowPtr[(size_t)nOwn_]); A.val.resize((size_t)A.rowPtr[(size_t)nOwn_]);
        }
        hostA_ = A;
        const int nRows = A.n;
        std::vector<float> inv((size_t)nV_);
        for (int v = 0; v < nRows; ++v) {
            float dg = 1.0f;
            for (int e = A.rowPtr[v]; e < A.rowPtr[v + 1]; ++e) if (A.col[e] == v) { dg = A.val[e]; break; }
            if (std::fabs(dg) < 1e-9f) dg = 1.0f;                               // ExtractInverseDiagonalKernel, pcgJacobi.cu:6-19
            inv[(size_t)v] = 1.0f / dg;
        }
        if (d.A.rowPtr) {       // re-prepare after Reset: the previous matrix goes back first (same pattern, new values)
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            dfree(d.A.rowPtr, ((size_t)d.A.n + 1) * 4); dfree(d.A.col, d.nnzA * 4); dfree(d.A.val, d.nnzA * 4); dfree(d.A.invDiag, (size_t)nV_ * 4);
            d.A = CsrDev{0, nullptr, nullptr, nullptr, nullptr};
        }
        int* rp = dalloc<int>(A.rowPtr.size()); int* cl = dalloc<int>(A.col.size()); float* vl = dalloc<float>(A.val.size()); float* iv = dalloc<float>(nV_);
        CUDA_CHECK(cudaMemcpy(rp, A.rowPtr.data(), A.rowPtr.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(cl, A.col.data(), A.col.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(vl, A.val.data(), A.val.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(iv, inv.data(), (size_t)nV_ * 4, cudaMemcpyHostToDevice));
        d.A = CsrDev{nRows, rp, cl, vl, iv};
        d.nnzA = A.col.size();
        if (!d.rhs) {
            d.rhs = dalloc<float4>(nV_); d.cgR = dalloc<float4>(nV_); d.cgQ = dalloc<float4>(nV_);
            d.xprev = dalloc<float4>(nV_); d.cholY = dalloc<float4>(nV_); d.cholZ = dalloc<float4>(nV_);
            CUDA_CHECK(cudaMemset(d.cholY, 0, (size_t)nV_ * 16));        // tag 0 in every entry: nothing solved yet
            CUDA_CHECK(cudaMemset(d.cholZ, 0, (size_t)nV_ * 16));
            d.solveState = dalloc<SolveState>(1);
            CUDA_CHECK(cudaMemset(d.solveState, 0, sizeof(SolveState)));
            int perSm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg_solve<false>, SOLVE_THREADS, 0));
            int perSmD = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmD, k_pcg_solve<true>, SOLVE_THREADS, 0));
            perSm = std::min(perSm, perSmD);
            int perSm2 = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm2, k_chol_solve, SOLVE_THREADS, 0));
            perSm = std::max(1, perSm);
            // at most THREE blocks per SM: a CG iteration at config-3 size is grid barriers and reductions over one slot per block more
            // than bandwidth (same-box A/B on grid55-pcg: 12.6 ms per step with every resident block, 12.8 / 11.1 / 10.7-11.2 with
            // 1 / 2 / 3 per SM, profiles/r2_pcg_grid_size_ab.txt)
            d.solveGrid = std::min(numSms_ * std::min(perSm, 3), std::min(SOLVE_MAX_PARTIALS, (nOwn_ + SOLVE_THREADS - 1) / SOLVE_THREADS));
            d.solveGrid = std::max(d.solveGrid, 1);
            if (const char* e = std::getenv("PD_SOLVE_BLOCKS_PER_SM"))      // (experiments: fewer blocks = cheaper grid barriers, longer per-thread row loops)
                d.solveGrid = std::max(1, std::min(d.solveGrid, numSms_ * std::max(1, std::atoi(e))));
            // the triangular solves run one WARP per row: as many resident warps as rows, up to the whole GPU
            // (not more than 4 blocks per SM: warps beyond the rows that can make progress only poll)
            d.cholGrid = std::max(1, std::min(numSms_ * std::min(4, std::max(1, perSm2)), std::min(SOLVE_MAX_PARTIALS, (nOwn_ + SOLVE_THREADS / 32 - 1) / (SOLVE_THREADS / 32))));
            if (const char* e = std::getenv("PD_CHOL_BLOCKS_PER_SM"))       // (experiments)
                d.cholGrid = std::max(1, std::min(numSms_ * std::min(std::max(1, perSm2), std::max(1, std::atoi(e))), SOLVE_MAX_PARTIALS));
            d.partials = dalloc<double>(2 * 3 * (size_t)SOLVE_MAX_PARTIALS);      // two alternating sets of slots (grid_sum3)
        }
        d.solverReady = true;
    }
    if (params_.globalSolver == 1 && !d.cholReady) {
        if (nV_ > 262144) throw std::runtime_error("the sparse Cholesky global step is the small-mesh path (<= 262144 vertices); use Jacobi or PCG");
        // fill-reducing order first (geometric nested dissection on the rest positions, layout.cpp), then the factorisation
        CholFactor F;
        std::vector<int> perm;
        {
            std::vector<float> xyz(3 * (size_t)nV_);
            for (int v = 0; v < nV_; ++v)
                for (int c = 0; c < 3; ++c) xyz[3 * (size_t)v + c] = scene_.X[3 * (size_t)L_.vertOrder[(size_t)v] + c];
            nested_dissection_order(hostA_, xyz.data(), perm);
            CsrMatrix Ap;
            permute_symmetric(hostA_, perm, Ap);
            cholesky_factor(Ap, F);
        }
        if (d.C.lPtr) {
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            dfree(d.C.lPtr, ((size_t)d.C.n + 1) * 4); dfree(d.C.lCol, d.nnzL * 4); dfree(d.C.lVal, d.nnzL * 4);
            dfree(d.C.uPtr, ((size_t)d.C.n + 1) * 4); dfree(d.C.uCol, d.nnzL * 4); dfree(d.C.uVal, d.nnzL * 4); dfree(d.C.perm, (size_t)d.C.n * 4);
            d.C = CholDev{0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        }
        int* lp = dalloc<int>(F.lPtr.size()); int* lc = dalloc<int>(F.lCol.size()); float* lv = dalloc<float>(F.lVal.size());
        int* up = dalloc<int>(F.uPtr.size()); int* uc = dalloc<int>(F.uCol.size()); float* uv = dalloc<float>(F.uVal.size());
        CUDA_CHECK(cudaMemcpy(lp, F.lPtr.data(), F.lPtr.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(lc, F.lCol.data(), F.lCol.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(lv, F.lVal.data(), F.lVal.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(up, F.uPtr.data(), F.uPtr.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(uc, F.uCol.data(), F.uCol.size() * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(uv, F.uVal.data(), F.uVal.size() * 4, cudaMemcpyHostToDevice));
        int* pm = dalloc<int>(perm.size());
        CUDA_CHECK(cudaMemcpy(pm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice));
        d.C = CholDev{nV_, lp, lc, lv, up, uc, uv, pm};
        d.nnzL = F.lCol.size();
        d.cholReady = true;
    }
}

// One PdSolver::Update in the direct / CG modes (pdSolver.cu:141-208 with isJacobi == false): the iterate lives in
// q[0]; every PD iteration = local step (H = w R DmInv^T G) -> right-hand side -> ONE cooperative solve kernel that
// also evaluates computeError and raises the device-side `done` flag; later iterations of the step then return at once.
void Engine::enqueueStepSolver(bool timed)
{
    Impl& d = *d_;
    const SolverParams& p = params_;
    const int n = nOwn_;                    // multi-GPU: this rank's vertices (rows); ghosts are filled by the halo pushes
    const int vb = 256, vg = (n + vb - 1) / vb;
    const float dtInv = 1.0f / p.dt;
    const float wdbc = 1e6f * (dtInv * dtInv);
    const bool dist = opt_.world > 1;
    pdlActive_ = false;                     // cooperative solve kernels in between: plain, fully serialised launches
    const DragArgs dr = dragActive_ ? d.drag(dragTarget_, numDBC_) : DragArgs{};
    if (dragActive_) k_predict<true><<<vg, vb, 0, stream_>>>(n, d.X, d.V, d.mass, d.dbc, d.md, p.dt, dt2Prepared_, p.gravity, d.q[0], d.q[2], d.b0, d.cc, dr);
    else k_predict<false><<<vg, vb, 0, stream_>>>(n, d.X, d.V, d.mass, d.dbc, d.md, p.dt, dt2Prepared_, p.gravity, d.q[0], d.q[2], d.b0, d.cc, dr);
    enqueuePush(d.q[0], 0);
    CUDA_CHECK(cudaMemsetAsync(d.xprev, 0, (size_t)n * 16, stream_));        // cudaMemset(prev_x, 0, ...), pdSolver.cu:162
    k_solve_begin<<<1, 1, 0, stream_>>>(d.solveState);
    DistSolve ds{};
    if (dist) ds = DistSolve{opt_.world, opt_.rank, d.nNbr, d.nPush, scene_.numVerts, d.nbrRanks, d.pushSrc, d.pushDst, d.pushNbr,
                             d.peerP, d.peerPFlag, d.pflags, d.peerRed, d.red, d.solveSeq, d.status};
    auto rec = [&](size_t k) { if (timed) CUDA_CHECK(cudaEventRecord(d.events[k], stream_)); };
    for (int i = 0; i < p.numIterations; ++i) {
        rec(3 * (size_t)i);
        launchLocal(d.q[0], false);
        if (dragActive_) {
            if (opt_.rotMode == 1) k_vertex_rhs<false, true><<<vg, vb, 0, stream_>>>(n, d.dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, wdbc, d.rhs, dr);
            else k_vertex_rhs<true, true><<<vg, vb, 0, stream_>>>(n, d.dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, wdbc, d.rhs, dr);
        }
        else if (opt_.rotMode == 1) k_vertex_rhs<false><<<vg, vb, 0, stream_>>>(n, d.dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, wdbc, d.rhs, dr);
        else k_vertex_rhs<true><<<vg, vb, 0, stream_>>>(n, d.dbcx, d.b0, d.cc, d.vslotPtr, d.vslot, d.P, wdbc, d.rhs, dr);
        rec(3 * (size_t)i + 1);
        if (p.globalSolver == 2) {
            const float4* b = d.rhs; float4 *x = d.q[0], *r = d.cgR, *pp = d.cgP, *qq = d.cgQ, *xp = d.xprev;
            int maxIter = p.pcgMaxIter; float cgTol = p.pcgTol, pdTol = p.tol;
            SolveState* st = d.solveState; double* part = d.partials;
            void* args[] = {&d.A, &b, &x, &r, &pp, &qq, &xp, &maxIter, &cgTol, &pdTol, &st, &part, &ds};
            CUDA_CHECK(cudaLaunchCooperativeKernel(dist ? (const void*)k_pcg_solve<true> : (const void*)k_pcg_solve<false>,
                                                   dim3(d.solveGrid), dim3(SOLVE_THREADS), args, 0, stream_));
            enqueuePush(d.q[0], 0);         // multi-GPU: the new iterate's boundary entries -> the neighbours' ghosts
        } else {
            const float4* b = d.rhs; float4 *x = d.q[0], *y = d.cholY, *z = d.cholZ, *xp = d.xprev;
            int tag = d.cholEpoch; d.cholEpoch = (d.cholEpoch + 1) % 0x3fffffff; float pdTol = p.tol;
            SolveState* st = d.solveState; double* part = d.partials;
            void* args[] = {&d.C, &b, &x, &y, &z, &xp, &tag, &pdTol, &st, &part};
            CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)k_chol_solve, dim3(d.cholGrid), dim3(SOLVE_THREADS), args, 0, stream_));
        }
        rec(3 * (size_t)i + 2);
    }
    rec(3 * (size_t)p.numIterations);
    enqueueEnd(d.q[0]);
    rec(3 * (size_t)p.numIterations + 1);
    pdlActive_ = usePdl_;
}

void Engine::solverSizes(size_t& nnzA, size_t& nnzL) const { nnzA = d_->nnzA; nnzL = d_->nnzL; }

const CsrMatrix& Engine::systemMatrix()
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!ready_) prepare();
    const int saved = params_.globalSolver;
    if (saved == 0) params_.globalSolver = 2;      // build the matrix only, no factorisation
    try { prepareSolver(); } catch (...) { params_.globalSolver = saved; throw; }
    params_.globalSolver = saved;
    return hostA_;
}

// PD / inner iteration counts of the non-Jacobi modes live on the device; fold them into the host counters
void Engine::syncSolveStats()
{
    Impl& d = *d_;
    if (!d.solveState) return;
    CUDA_CHECK(cudaSetDevice(opt_.device));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    SolveState st;
    CUDA_CHECK(cudaMemcpy(&st, d.solveState, sizeof(st), cudaMemcpyDeviceToHost));
    perfc_.pdIterations += st.pdItersTotal;
    perfc_.innerIterations += st.innerIters;
    lastErr_ = st.err; lastPdIters_ = st.pdIters;
    st.pdItersTotal = 0; st.innerIters = 0;
    CUDA_CHECK(cudaMemcpy(d.solveState, &st, sizeof(st), cudaMemcpyHostToDevice));
}

// ---------------------------------------------------------------- multi-GPU plumbing
void Engine::windowHandle(void* out64)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, d_->window));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(out64, &h, 64);
}

// peerBase[r] = base of rank r's exchange window as seen from this device (nullptr for non-neighbours)
void Engine::setPeers(const std::vector<uint8_t*>& peerBase)
{
    Impl& d = *d_;
    const RankPlan& P = plan_;
    d.nNbr = (int)P.neighbours.size();
    d.nPush = (int)P.pushSrc.size();
    std::vector<float4*> pq((size_t)3 * std::max(d.nNbr, 1));
    std::vector<int> nbrIndexOfRank((size_t)opt_.world, -1);
    for (int j = 0; j < d.nNbr; ++j) {
        const int r = P.neighbours[(size_t)j];
        if (!peerBase[(size_t)r]) throw std::runtime_error("missing exchange window of neighbour rank " + std::to_string(r));
        nbrIndexOfRank[(size_t)r] = j;
        const size_t nLocR = (size_t)P.nLocOf[(size_t)r];
        const WindowLayout wl(nLocR, opt_.world);
        for (int k = 0; k < 3; ++k) pq[(size_t)k * d.nNbr + j] = reinterpret_cast<float4*>(peerBase[(size_t)r]) + (size_t)k * nLocR;
    }
    {   // distributed PCG: the neighbours' p vectors and p-flags, and EVERY rank's reduction slots (own included)
        std::vector<float4*> pp((size_t)std::max(d.nNbr, 1));
        std::vector<unsigned long long*> ppf((size_t)std::max(d.nNbr, 1));
        std::vector<double*> pr((size_t)opt_.world, nullptr);
        for (int j = 0; j < d.nNbr; ++j) {
            const int r = P.neighbours[(size_t)j];
            const WindowLayout wl((size_t)P.nLocOf[(size_t)r], opt_.world);
            pp[(size_t)j] = reinterpret_cast<float4*>(peerBase[(size_t)r] + wl.offP);
            ppf[(size_t)j] = reinterpret_cast<unsigned long long*>(peerBase[(size_t)r] + wl.offPFlags) + opt_.rank;
        }
        for (int r = 0; r < opt_.world; ++r) {
            if (r == opt_.rank) { pr[(size_t)r] = d.red; continue; }
            if (!peerBase[(size_t)r]) throw std::runtime_error("missing exchange window of rank " + std::to_string(r));
            const WindowLayout wl((size_t)P.nLocOf[(size_t)r], opt_.world);
            pr[(size_t)r] = reinterpret_cast<double*>(peerBase[(size_t)r] + wl.offRed);
        }
        d.peerP = dalloc<float4*>((size_t)d.nNbr); d.peerPFlag = dalloc<unsigned long long*>((size_t)d.nNbr); d.peerRed = dalloc<double*>((size_t)opt_.world);
        if (d.nNbr) {
            CUDA_CHECK(cudaMemcpy(d.peerP, pp.data(), (size_t)d.nNbr * sizeof(float4*), cudaMemcpyHostToDevice));
            CUDA_CHECK(cudaMemcpy(d.peerPFlag, ppf.data(), (size_t)d.nNbr * sizeof(unsigned long long*), cudaMemcpyHostToDevice));
        }
        CUDA_CHECK(cudaMemcpy(d.peerRed, pr.data(), (size_t)opt_.world * sizeof(double*), cudaMemcpyHostToDevice));
    }
    std::vector<uint32_t> nb((size_t)std::max(d.nPush, 1));
    for (int i = 0; i < d.nPush; ++i) nb[(size_t)i] = (uint32_t)nbrIndexOfRank[(size_t)P.pushRank[(size_t)i]];
    d.pushSrc = dalloc<uint32_t>(d.nPush); d.pushDst = dalloc<uint32_t>(d.nPush); d.pushNbr = dalloc<uint32_t>(d.nPush);
    d.nbrRanks = dalloc<int>(d.nNbr);
    d.peerQ = dalloc<float4*>(3 * (size_t)d.nNbr);
    if (d.nPush) {
        CUDA_CHECK(cudaMemcpy(d.pushSrc, P.pushSrc.data(), (size_t)d.nPush * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.pushDst, P.pushDst.data(), (size_t)d.nPush * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.pushNbr, nb.data(), (size_t)d.nPush * 4, cudaMemcpyHostToDevice));
    }
    if (d.nNbr) {
        CUDA_CHECK(cudaMemcpy(d.nbrRanks, P.neighbours.data(), (size_t)d.nNbr * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d.peerQ, pq.data(), (size_t)3 * d.nNbr * sizeof(float4*), cudaMemcpyHostToDevice));
    }
    d.wait = DistWait{d.epoch, d.nNbr, nOwn_, plan_.nInteriorTiles, d.status, 0, d.pushSrc, d.pushDst, d.pushNbr, d.peerQ};
    connected_ = true;
}

void Engine::connectIpc(const void* handles)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (opt_.world == 1) { connected_ = true; return; }
    std::vector<uint8_t*> base((size_t)opt_.world, nullptr);
    for (int r = 0; r < opt_.world; ++r) {      // every rank's window: the neighbours' for the halo, all of them for the CG all-reduce
        if (r == opt_.rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const uint8_t*>(handles) + 64 * (size_t)r, 64);
        void* p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        d_->ipcOpened.push_back(p);
        base[(size_t)r] = static_cast<uint8_t*>(p);
    }
    setPeers(base);
}

void Engine::connectLocal(Engine* const* engines, int n)
{
    for (int r = 0; r < n; ++r) {
        Engine* e = engines[r];
        if (e->opt_.world != n || e->opt_.rank != r) throw std::runtime_error("connectLocal: engines must be ranks 0..n-1 of world n, in order");
        CUDA_CHECK(cudaSetDevice(e->opt_.device));
        std::vector<uint8_t*> base((size_t)n, nullptr);
        for (int q = 0; q < n; ++q) {
            if (q == r) continue;
            if (engines[q]->opt_.device != e->opt_.device) {
                int can = 0;
                CUDA_CHECK(cudaDeviceCanAccessPeer(&can, e->opt_.device, engines[q]->opt_.device));
                if (!can) throw std::runtime_error("connectLocal: no peer access between the devices");
                cudaError_t pe = cudaDeviceEnablePeerAccess(engines[q]->opt_.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(pe);
                cudaGetLastError();
            }
            base[(size_t)q] = engines[q]->d_->window;
        }
        e->setPeers(base);
    }
}

// One host thread drives all ranks phase by phase (tests on one GPU, where a rank's local kernel could
// otherwise occupy every SM while it waits for a neighbour that cannot be scheduled): after each phase every
// stream waits for every other stream, so all tagged ghost positions have landed when a local kernel starts.
void Engine::stepLockstep(Engine* const* engines, int n, int nSteps)
{
    auto crossSync = [&]() {
        for (int r = 0; r < n; ++r) {
            Engine* e = engines[r];
            CUDA_CHECK(cudaSetDevice(e->opt_.device));
            if (!e->d_->lockEvent) CUDA_CHECK(cudaEventCreateWithFlags(&e->d_->lockEvent, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventRecord(e->d_->lockEvent, e->stream_));
        }
        for (int r = 0; r < n; ++r)
            for (int q = 0; q < n; ++q)
                if (q != r) { CUDA_CHECK(cudaSetDevice(engines[r]->opt_.device)); CUDA_CHECK(cudaStreamWaitEvent(engines[r]->stream_, engines[q]->d_->lockEvent, 0)); }
    };
    for (int r = 0; r < n; ++r) {
        CUDA_CHECK(cudaSetDevice(engines[r]->opt_.device));
        if (engines[r]->params_.globalSolver != 0) throw std::runtime_error("stepLockstep: Jacobi global solver only");
        if (!engines[r]->ready_) engines[r]->prepare();
    }
    const int iters = engines[0]->params_.numIterations;
    for (int r = 0; r < n; ++r) engines[r]->lockstep_ = true;
    for (int s = 0; s < nSteps; ++s) {
        for (int r = 0; r < n; ++r) { CUDA_CHECK(cudaSetDevice(engines[r]->opt_.device)); engines[r]->enqueuePredict(); }
        crossSync();
        for (int i = 0; i < iters; ++i) {
            for (int r = 0; r < n; ++r) { CUDA_CHECK(cudaSetDevice(engines[r]->opt_.device)); engines[r]->enqueueIteration(i, false, nullptr); }
            crossSync();
        }
        for (int r = 0; r < n; ++r) {
            CUDA_CHECK(cudaSetDevice(engines[r]->opt_.device));
            engines[r]->enqueueFinish();
            engines[r]->perfc_.steps += 1; engines[r]->perfc_.pdIterations += iters;
        }
    }
    for (int r = 0; r < n; ++r) { engines[r]->lockstep_ = false; CUDA_CHECK(cudaSetDevice(engines[r]->opt_.device)); CUDA_CHECK(cudaGetLastError()); }
}

unsigned int Engine::distStatus()
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    unsigned int s = 0;
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaMemcpy(&s, d_->status, 4, cudaMemcpyDeviceToHost));
    return s;
}

// per-phase clock totals of the local kernel (warp 0 of every CTA): out[8 * localGrid()]
void Engine::profileLocal(unsigned long long* out)
{
    CUDA_CHECK(cudaSetDevice(opt_.device));
    if (!ready_) prepare();
    unsigned long long* dprof = nullptr;
    CUDA_CHECK(cudaMalloc(&dprof, 64ull * localGrid_));
    CUDA_CHECK(cudaMemsetAsync(dprof, 0, 64ull * localGrid_, stream_));
    launchLocal(d_->XT, true, nullptr, -1, false);            // warm
    launchLocal(d_->XT, true, dprof, -1, false);
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaMemcpy(out, dprof, 64ull * localGrid_, cudaMemcpyDeviceToHost));
    cudaFree(dprof);
}

// test hook: the collision pass's continuous-collision test on a batch of queries (host pointers)
__global__ void k_ccd_batch(int n, const int* __restrict__ type, const uint32_t* __restrict__ v, const float4* __restrict__ X, const float4* __restrict__ XT,
                            float* __restrict__ toi, float* __restrict__ nor)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ccd::V3 nn;
    toi[i] = ccd::collision_test(type[i] == 2, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], X, XT, nn);
    nor[3 * i] = nn.x; nor[3 * i + 1] = nn.y; nor[3 * i + 2] = nn.z;
}
void ccd_batch(int device, int n, const int* type, const uint32_t* verts, int nV, const float* X, const float* XT, float* toi, float* normals)
{
    CUDA_CHECK(cudaSetDevice(device));
    std::vector<float4> x4((size_t)nV), t4((size_t)nV);
    for (int v = 0; v < nV; ++v) {
        x4[(size_t)v] = make_float4(X[3 * v], X[3 * v + 1], X[3 * v + 2], 0.f);
        t4[(size_t)v] = make_float4(XT[3 * v], XT[3 * v + 1], XT[3 * v + 2], 0.f);
    }
    int* dT; uint32_t* dV; float4 *dX, *dXT; float *dToi, *dN;
    CUDA_CHECK(cudaMalloc(&dT, 4ull * n)); CUDA_CHECK(cudaMalloc(&dV, 16ull * n)); CUDA_CHECK(cudaMalloc(&dX, 16ull * nV)); CUDA_CHECK(cudaMalloc(&dXT, 16ull * nV));
    CUDA_CHECK(cudaMalloc(&dToi, 4ull * n)); CUDA_CHECK(cudaMalloc(&dN, 12ull * n));
    CUDA_CHECK(cudaMemcpy(dT, type, 4ull * n, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(dV, verts, 16ull * n, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dX, x4.data(), 16ull * nV, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(dXT, t4.data(), 16ull * nV, cudaMemcpyHostToDevice));
    k_ccd_batch<<<(n + 127) / 128, 128>>>(n, dT, dV, dX, dXT, dToi, dN);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(toi, dToi, 4ull * n, cudaMemcpyDeviceToHost)); CUDA_CHECK(cudaMemcpy(normals, dN, 12ull * n, cudaMemcpyDeviceToHost));
    cudaFree(dT); cudaFree(dV); cudaFree(dX); cudaFree(dXT); cudaFree(dToi); cudaFree(dN);
}

void rotation_batch(int device, int rotMode, int n, const float* F, float* R, int* usedFast)
{
    CUDA_CHECK(cudaSetDevice(device));
    float *dF = nullptr, *dR = nullptr; int* dU = nullptr;
    CUDA_CHECK(cudaMalloc(&dF, 36ull * n)); CUDA_CHECK(cudaMalloc(&dR, 36ull * n)); CUDA_CHECK(cudaMalloc(&dU, 4ull * n));
    CUDA_CHECK(cudaMemcpy(dF, F, 36ull * n, cudaMemcpyHostToDevice));
    if (rotMode == 0) k_rotation_batch<0><<<(n + 127) / 128, 128>>>(n, dF, dR, dU);
    else k_rotation_batch<1><<<(n + 127) / 128, 128>>>(n, dF, dR, dU);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(R, dR, 36ull * n, cudaMemcpyDeviceToHost));
    if (usedFast) CUDA_CHECK(cudaMemcpy(usedFast, dU, 4ull * n, cudaMemcpyDeviceToHost));
    cudaFree(dF); cudaFree(dR); cudaFree(dU);
}

}  // namespace pdb200
