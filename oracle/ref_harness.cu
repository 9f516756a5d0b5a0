// Replay harness around the REFERENCE's own CUDA kernels (TEST INFRASTRUCTURE).
//
// oracle/Makefile compiles this file together with
//   /root/reference/src/simulation/solver/projective/pdUtil.cu   (all PdUtil::* kernels, verbatim)
// and the headers it pulls in (svd.cuh, external/svd3_cuda/svd3_cuda.h, solverUtil.cuh,
// vendored glm) from where they lie under /root/reference; nothing is copied into this repo.
// The result, oracle/_ref/libpd_ref.so, is "the reference's CUDA build on one B200": the
// parity pin for the PD path and the timed reference arm of bench.py.
//
// This file replays PdSolver::SolverPrepare (pdSolver.cu:40-77, without the Eigen/cuSOLVER
// factorisation of :79-133, which Jacobi mode never uses), PdSolver::SolverStep
// (pdSolver.cu:141-208) and PdSolver::Update (pdSolver.cu:210-232) launch for launch with the
// reference's launch shapes (threadsPerBlock from context.json:4).  pdSolver.cu itself cannot
// be compiled here (Eigen headers are a CPM download).  The three fixed-body kernels
// (fixedBodyData.cu:67-134) are restated below because their TU needs OpenGL headers.
#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/transform.h>

#include <cstdio>
#include <vector>

#include <def.h>
#include <simulation/solver/projective/pdUtil.cuh>
#include <simulation/solver/solverUtil.cuh>
#include <svd.cuh>

namespace {

struct gravity_force {   // pdSolver.cu:14-18
    const float g;
    gravity_force(float _g) : g(_g) {}
    __device__ glm::vec3 operator()(float mass) const { return glm::vec3{0.0f, -g * mass, 0.0f}; }
};

// ---- restated fixed-body kernels (fixedBodyData.cu:67-134); bodies as plain arrays
__global__ void refFloor(glm::vec3* X, glm::vec3* V, int numVerts, const float* planes, int numPlanes, float muT, float muN)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numVerts) return;
    for (int j = 0; j < numPlanes; j++) {
        glm::vec3 floorPos(planes[6 * j], planes[6 * j + 1], planes[6 * j + 2]);
        glm::vec3 floorUp(planes[6 * j + 3], planes[6 * j + 4], planes[6 * j + 5]);
        float signedDis = glm::dot(X[i] - floorPos, floorUp);
        if (signedDis < 0 && glm::dot(V[i], floorUp) < 0) {
            X[i] -= signedDis * floorUp;
            glm::vec3 vN = glm::dot(V[i], floorUp) * floorUp;
            glm::vec3 vT = V[i] - vN;
            float mag_vT = glm::length(vT);
            float a = mag_vT == 0 ? 0 : glm::max(1 - muT * (1 + muN) * glm::length(vN) / mag_vT, 0.0f);
            V[i] = -muN * vN + a * vT;
        }
    }
}
__global__ void refSphere(glm::vec3* X, glm::vec3* V, int numVerts, const float* spheres, int numSpheres, float muT, float muN)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numVerts) return;
    for (int j = 0; j < numSpheres; j++) {
        glm::vec3 c(spheres[4 * j], spheres[4 * j + 1], spheres[4 * j + 2]);
        float r = spheres[4 * j + 3];
        glm::vec3 toCenter = X[i] - c;
        float d = glm::length(toCenter);
        if (d < r) {
            glm::vec3 normal = glm::normalize(toCenter);
            X[i] += (r - d) * normal;
            glm::vec3 vN = glm::dot(V[i], normal) * normal;
            glm::vec3 vT = V[i] - vN;
            float mag_vT = glm::length(vT);
            float a = mag_vT == 0 ? 0 : glm::max(1 - muT * (1 + muN) * glm::length(vN) / mag_vT, 0.0f);
            V[i] = -muN * vN + a * vT;
        }
    }
}
__global__ void refCylinder(glm::vec3* X, glm::vec3* V, int numVerts, const float* cyls, int numCyls, float muT, float muN)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numVerts) return;
    for (int j = 0; j < numCyls; j++) {
        glm::vec3 c(cyls[7 * j], cyls[7 * j + 1], cyls[7 * j + 2]);
        glm::vec3 axis(cyls[7 * j + 3], cyls[7 * j + 4], cyls[7 * j + 5]);
        float r = cyls[7 * j + 6];
        glm::mat3 nnT = glm::mat3(1.f) - glm::outerProduct(axis, axis);
        glm::vec3 n = nnT * (X[i] - c);
        float d = glm::length(n);
        if (d < r) {
            glm::vec3 normal = glm::normalize(n);
            X[i] += (r - d) * normal;
            glm::vec3 vN = glm::dot(V[i], normal) * normal;
            glm::vec3 vT = V[i] - vN;
            float mag_vT = glm::length(vT);
            float a = mag_vT == 0 ? 0 : glm::max(1 - muT * (1 + muN) * glm::length(vN) / mag_vT, 0.0f);
            V[i] = -muN * vN + a * vT;
        }
    }
}

// Control_Kernel restated (simulationContext.cu:202-218; its TU needs OpenGL headers); RADIUS_SQUARED is the reference's
// double literal (simulationContext.cu:18)
#define RADIUS_SQUARED 0.002
__global__ void refControl(glm::vec3* X, float* fixed, float* more_fixed, glm::vec3* offset_X, const float control_mag, const int number, const int select_v)
{
    int i = blockDim.x * blockIdx.x + threadIdx.x;
    if (i >= number) return;
    float stiffness = 0;
    if (fixed[i] == 0 && select_v != -1) {
        glm::vec3 diff = X[i] - X[select_v];
        offset_X[i] = diff;
        float dist2 = glm::dot(diff, diff);
        if (dist2 < RADIUS_SQUARED) stiffness = control_mag;
    }
    more_fixed[i] = stiffness;
}

struct Ref {
    int nV = 0, nT = 0, numDBC = 0, tpb = 128;
    // SolverData<float>
    glm::vec3 *X = nullptr, *X0 = nullptr, *XTilde = nullptr, *V = nullptr, *DBCX = nullptr, *ExtForce = nullptr, *OffsetX = nullptr;
    indexType* Tet = nullptr;
    float *mass = nullptr, *mu = nullptr, *DBC = nullptr, *moreDBC = nullptr, *V0 = nullptr, *contact_area = nullptr, *degree = nullptr;
    glm::mat3* DmInv = nullptr;
    // PdSolver private
    float *massDt_2s = nullptr, *sn = nullptr, *sn_old = nullptr, *b = nullptr, *matrix_diag = nullptr, *next_x = nullptr, *prev_x = nullptr;
    float omega = 1.f;
    glm::vec3 target = glm::vec3(0.f);   // SolverData::mouseSelection.target (def.h:14-18)
    bool ready = false;
    // fixed bodies
    float *planes = nullptr, *spheres = nullptr, *cyls = nullptr;
    int nPlanes = 0, nSpheres = 0, nCyls = 0;
    float perf[4] = {0, 0, 0, 0};
    std::vector<void*> allocs;
    template <typename T> T* alloc(size_t n) { void* p = nullptr; cudaMalloc(&p, (n ? n : 1) * sizeof(T)); allocs.push_back(p); return (T*)p; }
};

void prepare(Ref& r, float dt)
{   // pdSolver.cu:40-77
    int vertBlocks = (r.nV + r.tpb - 1) / r.tpb, tetBlocks = (r.nT + r.tpb - 1) / r.tpb;
    size_t len = (size_t)r.nV * 3 + 48 * (size_t)r.nT;
    cudaMemset(r.matrix_diag, 0, sizeof(float) * r.nV);
    int *AColIdx, *ARowIdx; float* AVal;
    cudaMalloc((void**)&AColIdx, sizeof(int) * len);
    cudaMalloc((void**)&ARowIdx, sizeof(int) * len);
    cudaMalloc((void**)&AVal, sizeof(float) * len);
    PdUtil::computeSiTSi<<<tetBlocks, r.tpb>>>(ARowIdx, AColIdx, AVal, r.matrix_diag, r.V0, r.DmInv, r.Tet, r.mu, r.nT, r.nV);
    PdUtil::setMDt_2<<<vertBlocks, r.tpb>>>(r.nV, ARowIdx, AColIdx, AVal, 48 * r.nT, r.mass, dt * dt, r.massDt_2s, r.DBC, 1e6f);
    cudaMemcpy(r.DBCX, r.X0, sizeof(glm::vec3) * r.nV, cudaMemcpyDeviceToDevice);
    cudaDeviceSynchronize();
    cudaFree(ARowIdx); cudaFree(AColIdx); cudaFree(AVal);
    r.ready = true;
}

template <typename F>
float timed(const F& f, bool perf)
{   // measureExecutionTime, solverUtil.cuh:7-25
    if (!perf) { f(); return 0; }
    cudaEvent_t start, stop;
    cudaEventCreate(&start); cudaEventCreate(&stop);
    cudaEventRecord(start);
    f();
    cudaEventRecord(stop);
    cudaEventSynchronize(stop);
    float ms = 0;
    cudaEventElapsedTime(&ms, start, stop);
    cudaEventDestroy(start); cudaEventDestroy(stop);
    return ms;
}

void solver_step(Ref& r, float dt, float gravity, float rho, int numIterations, bool perf)
{   // pdSolver.cu:141-208, Jacobi branch
    const float dtInv = 1.0f / dt, dt2Inv = dtInv * dtInv;
    int vertBlocks = (r.nV + r.tpb - 1) / r.tpb, vert3Blocks = (r.nV * 3 + r.tpb - 1) / r.tpb, tetBlocks = (r.nT + r.tpb - 1) / r.tpb;
    thrust::transform(thrust::device_pointer_cast(r.mass), thrust::device_pointer_cast(r.mass) + r.nV,
                      thrust::device_pointer_cast(r.ExtForce), gravity_force(gravity));
    PdUtil::setMDt_2MoreDBC<<<vertBlocks, r.tpb>>>(r.nV, r.mass, dt * dt, r.massDt_2s, r.moreDBC, r.DBC);
    PdUtil::computeSn<<<vertBlocks, r.tpb>>>(r.nV, r.sn, dt, r.massDt_2s, r.X, r.V, r.ExtForce, r.moreDBC, r.OffsetX, r.DBCX, r.target);
    cudaMemcpy(r.sn_old, r.sn, sizeof(float) * (r.nV * 3), cudaMemcpyDeviceToDevice);
    cudaMemcpy(r.prev_x, r.sn, sizeof(float) * (r.nV * 3), cudaMemcpyDeviceToDevice);
    for (int i = 0; i < numIterations; i++) {
        r.perf[0] += timed([&]() {
            PdUtil::addM_h2Sn<<<vertBlocks, r.tpb>>>(r.b, r.sn_old, r.massDt_2s, r.nV);
            PdUtil::computeLocal<<<tetBlocks, r.tpb>>>(r.V0, r.mu, r.b, r.DmInv, r.sn, r.Tet, r.nT, true);
            if (r.numDBC > 0)
                PdUtil::computeDBCLocal<<<vertBlocks, r.tpb>>>(r.nV, r.DBC, r.moreDBC, r.DBCX, 1e6f * dt2Inv, r.b);
        }, perf);
        r.perf[1] += timed([&]() {
            PdUtil::getErrorKern<<<vertBlocks, r.tpb>>>(r.nV, r.next_x, r.b, r.massDt_2s, r.sn, r.matrix_diag, r.moreDBC);
            if (i <= 10) r.omega = 1;
            else if (i == 11) r.omega = 2 / (2 - rho * rho);
            else r.omega = 4 / (4 - rho * rho * r.omega);
            PdUtil::chebyshevKern<<<vert3Blocks, r.tpb>>>(r.nV * 3, r.next_x, r.prev_x, r.sn, r.omega);
        }, perf);
    }
    PdUtil::updateVelPos<<<vertBlocks, r.tpb>>>(r.sn, dtInv, r.XTilde, r.V, r.nV, r.moreDBC);
}

}  // namespace

extern "C" {

void* ref_create(int nV, int nT, const float* X, const unsigned* Tet, const float* mass, const float* mu, const float* DBC,
                 int nPlanes, const float* planes, int nSpheres, const float* spheres, int nCyls, const float* cyls, int threadsPerBlock)
{
    Ref* r = new Ref;
    r->nV = nV; r->nT = nT; r->tpb = threadsPerBlock > 0 ? threadsPerBlock : 128;
    size_t v3 = sizeof(glm::vec3) * nV;
    r->X = r->alloc<glm::vec3>(nV); r->X0 = r->alloc<glm::vec3>(nV); r->XTilde = r->alloc<glm::vec3>(nV); r->V = r->alloc<glm::vec3>(nV);
    r->DBCX = r->alloc<glm::vec3>(nV); r->ExtForce = r->alloc<glm::vec3>(nV); r->OffsetX = r->alloc<glm::vec3>(nV);
    r->Tet = r->alloc<indexType>(4 * (size_t)nT);
    r->mass = r->alloc<float>(nV); r->mu = r->alloc<float>(nT); r->DBC = r->alloc<float>(nV); r->moreDBC = r->alloc<float>(nV);
    r->V0 = r->alloc<float>(nT); r->DmInv = r->alloc<glm::mat3>(nT); r->contact_area = r->alloc<float>(nV); r->degree = r->alloc<float>(nV);
    r->massDt_2s = r->alloc<float>(nV); r->sn = r->alloc<float>(3 * (size_t)nV); r->sn_old = r->alloc<float>(3 * (size_t)nV);
    r->b = r->alloc<float>(3 * (size_t)nV); r->matrix_diag = r->alloc<float>(nV); r->next_x = r->alloc<float>(3 * (size_t)nV);
    r->prev_x = r->alloc<float>(3 * (size_t)nV);
    // DataLoader::AllocData, dataLoader.cu:291-378
    cudaMemcpy(r->X, X, v3, cudaMemcpyHostToDevice); cudaMemcpy(r->X0, X, v3, cudaMemcpyHostToDevice);
    cudaMemcpy(r->XTilde, X, v3, cudaMemcpyHostToDevice); cudaMemcpy(r->DBCX, X, v3, cudaMemcpyHostToDevice);
    cudaMemset(r->V, 0, v3); cudaMemset(r->ExtForce, 0, v3); cudaMemset(r->OffsetX, 0, v3);
    cudaMemset(r->moreDBC, 0, sizeof(float) * nV); cudaMemset(r->contact_area, 0, sizeof(float) * nV); cudaMemset(r->degree, 0, sizeof(float) * nV);
    cudaMemcpy(r->Tet, Tet, sizeof(indexType) * 4 * (size_t)nT, cudaMemcpyHostToDevice);
    cudaMemcpy(r->mass, mass, sizeof(float) * nV, cudaMemcpyHostToDevice);
    cudaMemcpy(r->mu, mu, sizeof(float) * nT, cudaMemcpyHostToDevice);
    if (DBC) { cudaMemcpy(r->DBC, DBC, sizeof(float) * nV, cudaMemcpyHostToDevice); for (int i = 0; i < nV; i++) if (DBC[i] > 0) r->numDBC++; }
    else cudaMemset(r->DBC, 0, sizeof(float) * nV);
    r->nPlanes = nPlanes; r->nSpheres = nSpheres; r->nCyls = nCyls;
    r->planes = r->alloc<float>(6 * (size_t)nPlanes); r->spheres = r->alloc<float>(4 * (size_t)nSpheres); r->cyls = r->alloc<float>(7 * (size_t)nCyls);
    if (nPlanes) cudaMemcpy(r->planes, planes, 24 * (size_t)nPlanes, cudaMemcpyHostToDevice);
    if (nSpheres) cudaMemcpy(r->spheres, spheres, 16 * (size_t)nSpheres, cudaMemcpyHostToDevice);
    if (nCyls) cudaMemcpy(r->cyls, cyls, 28 * (size_t)nCyls, cudaMemcpyHostToDevice);
    // FEMSolver ctor, femSolver.cu:6-17
    int blocks = (nT + r->tpb - 1) / r->tpb;
    computeInvDmV0<float><<<blocks, r->tpb>>>(r->V0, r->DmInv, nT, r->X, r->Tet, r->contact_area, r->degree);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "ref_create: %s\n", cudaGetErrorString(e)); }
    return r;
}

void ref_destroy(void* h)
{
    Ref* r = (Ref*)h;
    if (!r) return;
    for (void* p : r->allocs) cudaFree(p);
    delete r;
}

// PdSolver::Update (pdSolver.cu:210-232) with handleCollision == false, n times; returns 0 or a CUDA error code
int ref_step(void* h, float dt, float gravity, float rho, float muN, float muT, int numIterations, int nSteps, int perf)
{
    Ref& r = *(Ref*)h;
    for (int s = 0; s < nSteps; s++) {
        if (!r.ready) prepare(r, dt);
        solver_step(r, dt, gravity, rho, numIterations, perf != 0);
        cudaMemcpy(r.X, r.XTilde, sizeof(glm::vec3) * r.nV, cudaMemcpyDeviceToDevice);
        r.perf[2] += timed([&]() {   // FixedBodyData::HandleCollisions, fixedBodyData.cu:136-148
            int nb = (r.nV + r.tpb - 1) / r.tpb;
            if (r.nSpheres > 0) refSphere<<<nb, r.tpb>>>(r.XTilde, r.V, r.nV, r.spheres, r.nSpheres, muT, muN);
            if (r.nPlanes > 0) refFloor<<<nb, r.tpb>>>(r.XTilde, r.V, r.nV, r.planes, r.nPlanes, muT, muN);
            if (r.nCyls > 0) refCylinder<<<nb, r.tpb>>>(r.XTilde, r.V, r.nV, r.cyls, r.nCyls, muT, muN);
        }, perf != 0);
    }
    return (int)cudaGetLastError();
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

void ref_reset(void* h)
{   // simulationContext.cu:233-243
    Ref& r = *(Ref*)h;
    size_t v3 = sizeof(glm::vec3) * r.nV;
    cudaMemcpy(r.X, r.X0, v3, cudaMemcpyDeviceToDevice); cudaMemcpy(r.XTilde, r.X0, v3, cudaMemcpyDeviceToDevice);
    cudaMemset(r.V, 0, v3);
    cudaMemset(r.moreDBC, 0, sizeof(float) * r.nV);
    r.ready = false;
    for (float& p : r.perf) p = 0;
}

void ref_get(void* h, float* X, float* V, float* XTilde)
{
    Ref& r = *(Ref*)h;
    size_t v3 = sizeof(glm::vec3) * r.nV;
    if (X) cudaMemcpy(X, r.X, v3, cudaMemcpyDeviceToHost);
    if (V) cudaMemcpy(V, r.V, v3, cudaMemcpyDeviceToHost);
    if (XTilde) cudaMemcpy(XTilde, r.XTilde, v3, cudaMemcpyDeviceToHost);
}

void ref_set(void* h, const float* X, const float* V, const float* XTilde)
{
    Ref& r = *(Ref*)h;
    size_t v3 = sizeof(glm::vec3) * r.nV;
    if (X) cudaMemcpy(r.X, X, v3, cudaMemcpyHostToDevice);
    if (V) cudaMemcpy(r.V, V, v3, cudaMemcpyHostToDevice);
    if (XTilde) cudaMemcpy(r.XTilde, XTilde, v3, cudaMemcpyHostToDevice);
}

// mouse drag: SolverData::moreDBC / OffsetX / mouseSelection.target as the caller's Control_Kernel would leave them;
// moreDBC NULL = ResetMoreDBC(true) (simulationContext.cu:220-226)
void ref_set_drag(void* h, const float* moreDBC, const float* OffsetX, const float* target)
{
    Ref& r = *(Ref*)h;
    if (!moreDBC) { cudaMemset(r.moreDBC, 0, sizeof(float) * r.nV); return; }
    cudaMemcpy(r.moreDBC, moreDBC, sizeof(float) * r.nV, cudaMemcpyHostToDevice);
    if (OffsetX) cudaMemcpy(r.OffsetX, OffsetX, sizeof(glm::vec3) * r.nV, cudaMemcpyHostToDevice);
    if (target) r.target = glm::vec3(target[0], target[1], target[2]);
}
// ResetMoreDBC(false) while dragging (simulationContext.cu:227-230) with the reference's launch shape, then the target
void ref_drag_select(void* h, int select_v, float control_mag, const float* target)
{
    Ref& r = *(Ref*)h;
    refControl<<<r.nV / r.tpb + 1, r.tpb>>>(r.X, r.DBC, r.moreDBC, r.OffsetX, control_mag, r.nV, select_v);
    if (target) r.target = glm::vec3(target[0], target[1], target[2]);
}
void ref_get_drag(void* h, float* moreDBC, float* OffsetX, float* DBCX)
{
    Ref& r = *(Ref*)h;
    if (moreDBC) cudaMemcpy(moreDBC, r.moreDBC, sizeof(float) * r.nV, cudaMemcpyDeviceToHost);
    if (OffsetX) cudaMemcpy(OffsetX, r.OffsetX, sizeof(glm::vec3) * r.nV, cudaMemcpyDeviceToHost);
    if (DBCX) cudaMemcpy(DBCX, r.DBCX, sizeof(glm::vec3) * r.nV, cudaMemcpyDeviceToHost);
}

void ref_get_setup(void* h, float dt, float* matrix_diag, float* massDt_2s, float* DmInv /*9/tet glm column-major*/, float* V0)
{
    Ref& r = *(Ref*)h;
    if (!r.ready) prepare(r, dt);
    if (matrix_diag) cudaMemcpy(matrix_diag, r.matrix_diag, sizeof(float) * r.nV, cudaMemcpyDeviceToHost);
    if (massDt_2s) cudaMemcpy(massDt_2s, r.massDt_2s, sizeof(float) * r.nV, cudaMemcpyDeviceToHost);
    if (DmInv) cudaMemcpy(DmInv, r.DmInv, sizeof(glm::mat3) * r.nT, cudaMemcpyDeviceToHost);
    if (V0) cudaMemcpy(V0, r.V0, sizeof(float) * r.nT, cudaMemcpyDeviceToHost);
}

void ref_get_perf(void* h, float* out4) { Ref& r = *(Ref*)h; for (int i = 0; i < 4; i++) out4[i] = r.perf[i]; }

// svdGLM + R = U V^T exactly as PdUtil::computeLocal does it, for a batch of F (row-major 9 floats each)
__global__ void refRotationKernel(int n, const float* F, float* Rout)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    glm::mat3 Fm;   // glm is [col][row]
    for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) Fm[c][rr] = F[9 * i + rr * 3 + c];
    glm::mat3 U, S, R;
    svdGLM(Fm, U, S, R);
    R = U * glm::transpose(R);
    if (glm::determinant(R) < 0) R[2] = -R[2];
    for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) Rout[9 * i + rr * 3 + c] = R[c][rr];
}
int ref_rotation(int n, const float* F, float* R)
{
    float *dF, *dR;
    cudaMalloc(&dF, 36 * (size_t)n); cudaMalloc(&dR, 36 * (size_t)n);
    cudaMemcpy(dF, F, 36 * (size_t)n, cudaMemcpyHostToDevice);
    refRotationKernel<<<(n + 127) / 128, 128>>>(n, dF, dR);
    cudaMemcpy(R, dR, 36 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(dF); cudaFree(dR);
    return (int)cudaGetLastError();
}

}  // extern "C"
