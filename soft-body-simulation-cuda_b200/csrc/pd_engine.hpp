// PD step engine: owns the device-resident state of one simulation context and replays
// PdSolver::Update (src/simulation/solver/projective/pdSolver.cu:210-232) on one B200.
#pragma once
#include <cstdint>
#include <memory>

#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "layout.hpp"
#include "scene.hpp"

namespace pdb200 {

struct PerfCounters {      // PdSolver::performanceData, pdSolver.cu:26 (milliseconds, accumulated)
    float localStep = 0, globalStep = 0, collisionFixed = 0, collisionMesh = 0;
    // extras
    double stepMsTotal = 0; long long steps = 0, pdIterations = 0, innerIterations = 0, kernelLaunches = 0;
};

struct EngineOptions {
    int device = 0;
    int rotMode = -1;       // -1 auto (default): 1 with reorder 0 on a mesh that fits ONE tile (where the reference is deterministic and the
                            // mode costs nothing: bit-exact reproduction), else 0;  0 Newton polar + SVD fallback;  1 always the Jacobi SVD
    int bodyKernel = -1;    // -1 auto: scenes whose connected components all fit one CTA's shared memory step with one CTA per body and one
                            // launch per step (pd_body_kernel.cuh); 0 never (the tile path)
    int reorder = 1;        // Morton reordering of tets / first-touch renumbering of vertices
    int useGraph = 1;       // replay each step as one CUDA graph
    int ctasPerSm = 0;      // 0 = occupancy query
    int rank = 0, world = 1; // multi-GPU: this engine is rank `rank` of `world` (one process per GPU)
};

class Engine {
public:
    Engine(const Scene& scene, const EngineOptions& opt);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    void setParams(const SolverParams& p);
    const SolverParams& params() const { return params_; }
    void step(int nSteps);                                  // PdSolver::Update x n
    float stepTimed(int nSteps);                            // same, returns device ms (events on the engine stream)
    void reset();                                           // SimulationCUDAContext::Reset
    void setPerf(bool on) { perf_ = on; }
    void synchronize();
    // host <-> device state in the reference's layout (AoS float3, original vertex numbering)
    void download(float* X, float* V, float* XTilde);
    void upload(const float* X, const float* V, const float* XTilde);
    // device <-> device against the reference's SolverData arrays (glm::vec3*, original numbering)
    void importDevice(const float* dX, const float* dV, const float* dXTilde);
    void exportDevice(float* dX, float* dV, float* dXTilde);
    // end-to-end step on host buffers (pinned or pageable): upload, n steps, download
    void stepHost(int nSteps, const float* Xin, const float* Vin, const float* XTin, float* Xout, float* Vout, float* XTout);

    // the same on this rank's SHARD only: 3 * numOwned() floats per array in local owned order (see ownedIds)
    void stepHostOwned(int nSteps, const float* Xin, const float* Vin, const float* XTin, float* Xout, float* Vout, float* XTout);
    void ownedIds(uint32_t* out) const;                     // original vertex id of each owned vertex, local order

    // ---- mouse-drag soft constraints: SolverData<float>::moreDBC / OffsetX / mouseSelection.target (def.h:14-18,31-32),
    // consumed at pdUtil.cu:56-69,80-87,159-164,187-188,201-206.  Arrays in the reference's layout and ORIGINAL vertex
    // numbering; more == nullptr (or no positive entry) ends the drag (ResetMoreDBC(true), simulationContext.cu:220-226).
    void setDrag(const float* more, const float* offsetX, const float target[3]);                 // host arrays
    void setDragDevice(const float* dMore, const float* dOffsetX, const float target[3]);         // the reference's device arrays
    void dragSelect(int selectV, float controlMag, const float target[3]);                        // Control_Kernel on the engine's X (simulationContext.cu:202-218)
    void getDrag(float* more, float* offsetX, float* dbcx);                                       // host, original numbering (parity checks)
    bool dragActive() const { return dragActive_; }
    // live stiffness edit: SolverData<float>::mu[numTets] changed (SimulationCUDAContext::UpdateSoftBodyAttr -> FillData,
    // simulationContext.cu:165-176).  computeLocal reads mu in every iteration (pdUtil.cu:97,124), so the new value acts at
    // once; matrix_diag and the assembled system matrix are SolverPrepare products and stay as they are until Reset().
    void updateMu(const float* mu);                    // host array, original tet order
    void updateMuDevice(const float* dMu);             // the reference's device array

    const PerfCounters& perf() const { return perfc_; }
    void syncSolveStats();                                  // PCG / Cholesky modes: fold the device-side iteration counters in
    float lastError() const { return lastErr_; }
    int lastPdIterations() const { return lastPdIters_; }
    const CsrMatrix& systemMatrix();                        // A^ in the renumbered vertex ids (builds it if needed)
    void resetPerf() { perfc_ = PerfCounters(); }
    const Layout& layout() const { return L_; }
    int numVerts() const { return nV_; }        // vertices held by this engine (owned + ghosts when world > 1)
    int numTets() const { return nT_; }         // tets evaluated by this engine
    int numOwned() const { return nOwn_; }
    int numVertsGlobal() const { return scene_.numVerts; }
    int numTetsGlobal() const { return scene_.numTets; }
    // ---- multi-GPU (world > 1)
    const RankPlan& plan() const { return plan_; }
    void windowHandle(void* out64);                                // cudaIpcMemHandle_t of this rank's exchange window
    void connectIpc(const void* handles);                          // world x 64 bytes in rank order (other processes)
    static void connectLocal(Engine* const* engines, int n);       // all ranks inside one process (tests)
    static void stepLockstep(Engine* const* engines, int n, int nSteps);   // one process driving all ranks phase by phase
    unsigned int distStatus();                                     // nonzero: a halo wait timed out
    // setup products in the ORIGINAL numbering/order (for parity checks against the oracle)
    void getSetup(float* matrixDiag, float* massDt2, float* DmInv, float* V0);
    // launch geometry, for the bench's launch count / roofline bookkeeping
    int localGrid() const { return localGrid_; }
    int rotMode() const { return opt_.rotMode; }
    void solverSizes(size_t& nnzA, size_t& nnzL) const;
    void getCollision(float* tI, float* normals, long long* numPairs);      // last mesh-mesh collision pass (host arrays, caller numbering)      // the effective mode (options.rotMode < 0 = auto is resolved in the ctor)
    size_t deviceBytes() const { return devBytes_; }
    size_t tileStreamBytes() const { return L_.records.size(); }
    cudaStream_t stream() const { return stream_; }
    // timing of one kernel family, measured with events over `reps` launches on the engine's stream
    float timeLocalKernelMs(int reps);
    float timeVertexKernelMs(int reps);
    void profileLocal(unsigned long long* out);     // 8 counters per CTA of the local kernel (clock64 per phase)

private:
    struct Impl;
    void prepare();
    void buildGraph();
    void enqueueStep(bool timed);
    void prepareSolver();
    void enqueueStepSolver(bool timed = false);
    void enqueuePredict();
    void enqueueIteration(int i, bool timed, size_t* ev);
    void enqueueFinish();
    void enqueueEnd(const float4* qfinal);
    void prepareCollision();
    void collisionPass();
    void enqueuePush(const float4* q, int bufIndex);
    void setPeers(const std::vector<uint8_t*>& peerBase);
    float* qbuf(int k) const;
    void ensureDragBuffers();
    void finishDragUpdate(const float target[3]);
    void launchLocal(const float4* q, bool jacobi, unsigned long long* prof = nullptr, int pushBuf = -1, bool checkHalo = true);
    template <typename T> T* dalloc(size_t n);
    void dfree(const void* p, size_t bytes);
    void waitCallerStream();

    int nV_ = 0, nT_ = 0, nOwn_ = 0;
    RankPlan plan_;
    long long phase_ = 0;     // multi-GPU: phases (predictor / iteration) completed so far; buffer rotation base = phase_ % 3
    int base_ = 0;            // position-buffer index of the current step's predictor output
    float omega_ = 1.f;
    CsrMatrix hostA_;         // scalar system matrix (renumbered ids), kept for the Cholesky factorisation and the parity tests
    float lastErr_ = 1.f; int lastPdIters_ = 0;
    bool connected_ = false;
    bool lockstep_ = false;   // driven phase by phase by stepLockstep: halo pushes are separate launches, not in the local kernel
    Scene scene_;             // host copy (original numbering)
    Layout L_;
    SolverParams params_;
    EngineOptions opt_;
    bool ready_ = false, perf_ = false, graphValid_ = false;
    bool pdlActive_ = false;
    int vertexFlags_ = 2;     // k_vertex_jacobi flag bit 1: its blocks walk the vertices from the END of the array (-0.4 % per step on grid139,
                              // profiles/r2_vertex_fold_reverse_ab_grid139.txt); PD_VERTEX_REVERSE=0 for A/B runs
    int pdlLate_ = 1;         // the dependents are released at the END of every CTA's work (PD_PDL=2, the default); 0: at its start (PD_PDL=1, measured slower)
    bool usePdl_ = true;      // programmatic dependent launch of the per-iteration kernels (pd_kernels.cuh: pdl_wait), late trigger: the next kernel's launch
                              // latency and prologue hide behind this one's tail: armadillo -15 %, grid55 -2.2 %, grid139 -1.3 %, N=2 grid70 -1.4 %
                              // (profiles/r2_pdl_late_ab.txt); PD_PDL=0 switches it off
                              // on B200 (grid139: 42.0 vs 36.9 ms/step, batch64: 2.68 vs 2.23), so it stays an opt-in experiment (PD_PDL=1)
    float dt2Prepared_ = 0.f;
    bool bodyKernel_ = false;             // EXPERIMENT PD_BODY_KERNEL=1: one CTA per small body, one launch per step (pd_body_kernel.cuh)
    bool dragActive_ = false;             // some vertex has moreDBC > 0: the DRAG kernel variants run, as plain launches (the target moves every frame)
    float dragTarget_[3] = {0.f, 0.f, 0.f};
    int numDBC_ = 0;                      // SolverData::numDBC
    PerfCounters perfc_;
    int localGrid_ = 0, numSms_ = 0;
    size_t devBytes_ = 0;
    cudaStream_t stream_ = nullptr;
    std::unique_ptr<Impl> d_;
};

// test hook: corotation() of n row-major 3x3 matrices on `device` (host pointers)
void rotation_batch(int device, int rotMode, int n, const float* F, float* R, int* usedFast);
// test hook: ccd::collision_test (pd_collision.cuh) on n queries; type 1 = vertex-face, 2 = edge-edge; verts 4 ids per query
void ccd_batch(int device, int n, const int* type, const uint32_t* verts, int nV, const float* X, const float* XT, float* toi, float* normals);

}  // namespace pdb200
