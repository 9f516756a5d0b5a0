// Replay harness around the REFERENCE's own direct / CG global-step back-ends (TEST INFRASTRUCTURE).
//
// oracle/Makefile (target `refsolvers`) compiles this file together with, from where they lie under /root/reference,
//   src/simulation/solver/projective/pdUtil.cu                 (PdUtil::* kernels)
//   src/simulation/solver/linear/{linear,cholesky,pcgJacobi}.cu (sort_coo, CholeskySpLinearSolver, PCGJacobiSolver)
// and links cuSPARSE / cuSOLVER / cuBLAS; nothing is copied into this repo.  The result, oracle/_ref/libpd_ref_solvers.so,
// pins the engine's non-Jacobi global steps:
//   solver 1: PdSolver in SolverType::CuSolverCholesky mode -- SolverPrepare's COO (pdSolver.cu:40-77) handed to
//             CholeskySpLinearSolver<float> (pdSolver.cu:128; its constructor sorts and sums the duplicates itself,
//             cholesky.cu:133-158, so the Eigen detour of pdSolver.cu:88-127 is not needed) and the direct branch of
//             SolverStep (pdSolver.cu:141-208: computeLocal(isJacobi = false), ls->Solve, computeError, early exit);
//   solver 2: the same branch with PCGJacobiSolver<float>::Solve (pcgJacobi.cu:88-172, warm start from the current
//             iterate) in place of ls->Solve -- the combination BASELINE config 3 names; the reference itself only
//             reaches this solver from IPC.
// No fixed bodies here (their kernels live in ref_harness.cu): the tests use contact-free stretches.
#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/transform.h>
#include <thrust/transform_reduce.h>

#include <cstdio>
#include <memory>
#include <vector>

#include <def.h>
#include <simulation/solver/projective/pdUtil.cuh>
#include <simulation/solver/solverUtil.cuh>
#include <linear/cholesky.h>
#include <linear/pcgJacobi.h>
#include <linear/cg.h>

namespace {

struct gravity_force {   // pdSolver.cu:14-18
    const float g;
    gravity_force(float _g) : g(_g) {}
    __device__ glm::vec3 operator()(float mass) const { return glm::vec3{0.0f, -g * mass, 0.0f}; }
};

float compute_error(thrust::device_ptr<float> sn, thrust::device_ptr<float> sn_old, int size)
{   // computeError, pdSolver.cu:243-253
    return thrust::transform_reduce(
               thrust::counting_iterator<indexType>(0), thrust::counting_iterator<indexType>(size),
               [=] __host__ __device__(indexType i) { return (sn_old[i] - sn[i]) * (sn_old[i] - sn[i]); }, 0.0, thrust::plus<float>()) / size;
}

struct RefS {
    int nV = 0, nT = 0, tpb = 128, numDBC = 0, solver = 1;
    glm::vec3 *X = nullptr, *X0 = nullptr, *XTilde = nullptr, *V = nullptr, *DBCX = nullptr, *ExtForce = nullptr, *OffsetX = nullptr;
    indexType* Tet = nullptr;
    float *mass = nullptr, *mu = nullptr, *DBC = nullptr, *moreDBC = nullptr, *V0 = nullptr, *contact_area = nullptr, *degree = nullptr;
    glm::mat3* DmInv = nullptr;
    float *massDt_2s = nullptr, *sn = nullptr, *sn_old = nullptr, *b = nullptr, *matrix_diag = nullptr, *prev_x = nullptr;
    int *ARow = nullptr, *ACol = nullptr; float* AVal = nullptr; int len = 0;
    std::unique_ptr<LinearSolver<float>> ls;
    bool ready = false;
    int itersLast = 0; float errLast = 1.f;
    std::vector<void*> allocs;
    template <typename T> T* alloc(size_t n) { void* p = nullptr; cudaMalloc(&p, (n ? n : 1) * sizeof(T)); allocs.push_back(p); return (T*)p; }
};

void prepare(RefS& r, float dt)
{   // pdSolver.cu:40-77 and :128
    int vertBlocks = (r.nV + r.tpb - 1) / r.tpb, tetBlocks = (r.nT + r.tpb - 1) / r.tpb;
    r.len = r.nV * 3 + 48 * r.nT;
    const int ASize = 3 * r.nV;
    cudaMemset(r.matrix_diag, 0, sizeof(float) * r.nV);
    if (!r.ARow) { r.ARow = r.alloc<int>(r.len); r.ACol = r.alloc<int>(r.len); r.AVal = r.alloc<float>(r.len); }
    cudaMemset(r.ARow, 0, sizeof(int) * r.len); cudaMemset(r.ACol, 0, sizeof(int) * r.len); cudaMemset(r.AVal, 0, sizeof(int) * r.len);
    PdUtil::computeSiTSi<<<tetBlocks, r.tpb>>>(r.ARow, r.ACol, r.AVal, r.matrix_diag, r.V0, r.DmInv, r.Tet, r.mu, r.nT, r.nV);
    PdUtil::setMDt_2<<<vertBlocks, r.tpb>>>(r.nV, r.ARow, r.ACol, r.AVal, 48 * r.nT, r.mass, dt * dt, r.massDt_2s, r.DBC, 1e6f);
    cudaDeviceSynchronize();
    if (r.solver == 1) r.ls = std::make_unique<CholeskySpLinearSolver<float>>(r.tpb, r.ARow, r.ACol, r.AVal, ASize, r.len);
    else r.ls = std::make_unique<PCGJacobiSolver<float>>(ASize);
    cudaMemcpy(r.DBCX, r.X0, sizeof(glm::vec3) * r.nV, cudaMemcpyDeviceToDevice);
    cudaDeviceSynchronize();
    r.ready = true;
}

void solver_step(RefS& r, float dt, float gravity, float tol, int numIterations)
{   // pdSolver.cu:141-208, solverType != Jacobi
    const float dtInv = 1.0f / dt, dt2Inv = dtInv * dtInv;
    const int N = r.nV * 3;
    int vertBlocks = (r.nV + r.tpb - 1) / r.tpb, tetBlocks = (r.nT + r.tpb - 1) / r.tpb;
    thrust::device_ptr<float> x_prime_ptr(r.prev_x), x_ptr(r.sn);
    thrust::transform(thrust::device_pointer_cast(r.mass), thrust::device_pointer_cast(r.mass) + r.nV,
                      thrust::device_pointer_cast(r.ExtForce), gravity_force(gravity));
    PdUtil::setMDt_2MoreDBC<<<vertBlocks, r.tpb>>>(r.nV, r.mass, dt * dt, r.massDt_2s, r.moreDBC, r.DBC);
    PdUtil::computeSn<<<vertBlocks, r.tpb>>>(r.nV, r.sn, dt, r.massDt_2s, r.X, r.V, r.ExtForce, r.moreDBC, r.OffsetX, r.DBCX, glm::vec3(0.f));
    cudaMemcpy(r.sn_old, r.sn, sizeof(float) * N, cudaMemcpyDeviceToDevice);
    cudaMemset(r.prev_x, 0, N);            // the reference's byte count (pdSolver.cu:162): only the first N BYTES are cleared
    float err = 1;
    int i = 0;
    for (; i < numIterations && sqrt(err) >= tol; i++) {
        PdUtil::addM_h2Sn<<<vertBlocks, r.tpb>>>(r.b, r.sn_old, r.massDt_2s, r.nV);
        PdUtil::computeLocal<<<tetBlocks, r.tpb>>>(r.V0, r.mu, r.b, r.DmInv, r.sn, r.Tet, r.nT, false);
        if (r.numDBC > 0)
            PdUtil::computeDBCLocal<<<vertBlocks, r.tpb>>>(r.nV, r.DBC, r.moreDBC, r.DBCX, 1e6f * dt2Inv, r.b);
        if (r.solver == 1) r.ls->Solve(N, r.b, r.sn);
        else r.ls->Solve(N, r.b, r.sn, r.AVal, r.len, r.ARow, r.ACol, r.sn);
        err = compute_error(x_ptr, x_prime_ptr, N);
        cudaMemcpy(r.prev_x, r.sn, sizeof(float) * N, cudaMemcpyDeviceToDevice);
    }
    r.itersLast = i; r.errLast = err;
    PdUtil::updateVelPos<<<vertBlocks, r.tpb>>>(r.sn, dtInv, r.XTilde, r.V, r.nV, r.moreDBC);
}

}  // namespace

extern "C" {

void* refs_create(int nV, int nT, const float* X, const unsigned* Tet, const float* mass, const float* mu, const float* DBC,
                  int solver, int threadsPerBlock)
{
    RefS* r = new RefS;
    r->nV = nV; r->nT = nT; r->tpb = threadsPerBlock > 0 ? threadsPerBlock : 128; r->solver = solver;
    size_t v3 = sizeof(glm::vec3) * nV;
    r->X = r->alloc<glm::vec3>(nV); r->X0 = r->alloc<glm::vec3>(nV); r->XTilde = r->alloc<glm::vec3>(nV); r->V = r->alloc<glm::vec3>(nV);
    r->DBCX = r->alloc<glm::vec3>(nV); r->ExtForce = r->alloc<glm::vec3>(nV); r->OffsetX = r->alloc<glm::vec3>(nV);
    r->Tet = r->alloc<indexType>(4 * (size_t)nT);
    r->mass = r->alloc<float>(nV); r->mu = r->alloc<float>(nT); r->DBC = r->alloc<float>(nV); r->moreDBC = r->alloc<float>(nV);
    r->V0 = r->alloc<float>(nT); r->DmInv = r->alloc<glm::mat3>(nT); r->contact_area = r->alloc<float>(nV); r->degree = r->alloc<float>(nV);
    r->massDt_2s = r->alloc<float>(nV); r->sn = r->alloc<float>(3 * (size_t)nV); r->sn_old = r->alloc<float>(3 * (size_t)nV);
    r->b = r->alloc<float>(3 * (size_t)nV); r->matrix_diag = r->alloc<float>(nV); r->prev_x = r->alloc<float>(3 * (size_t)nV);
    cudaMemcpy(r->X, X, v3, cudaMemcpyHostToDevice); cudaMemcpy(r->X0, X, v3, cudaMemcpyHostToDevice);
    cudaMemcpy(r->XTilde, X, v3, cudaMemcpyHostToDevice); cudaMemcpy(r->DBCX, X, v3, cudaMemcpyHostToDevice);
    cudaMemset(r->V, 0, v3); cudaMemset(r->ExtForce, 0, v3); cudaMemset(r->OffsetX, 0, v3);
    cudaMemset(r->moreDBC, 0, sizeof(float) * nV); cudaMemset(r->contact_area, 0, sizeof(float) * nV); cudaMemset(r->degree, 0, sizeof(float) * nV);
    cudaMemcpy(r->Tet, Tet, sizeof(indexType) * 4 * (size_t)nT, cudaMemcpyHostToDevice);
    cudaMemcpy(r->mass, mass, sizeof(float) * nV, cudaMemcpyHostToDevice);
    cudaMemcpy(r->mu, mu, sizeof(float) * nT, cudaMemcpyHostToDevice);
    if (DBC) { cudaMemcpy(r->DBC, DBC, sizeof(float) * nV, cudaMemcpyHostToDevice); for (int i = 0; i < nV; i++) if (DBC[i] > 0) r->numDBC++; }
    else cudaMemset(r->DBC, 0, sizeof(float) * nV);
    int blocks = (nT + r->tpb - 1) / r->tpb;
    computeInvDmV0<float><<<blocks, r->tpb>>>(r->V0, r->DmInv, nT, r->X, r->Tet, r->contact_area, r->degree);   // femSolver.cu:6-17
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) fprintf(stderr, "refs_create: %s\n", cudaGetErrorString(e));
    return r;
}

void refs_destroy(void* h)
{
    RefS* r = (RefS*)h;
    if (!r) return;
    r->ls.reset();
    for (void* p : r->allocs) cudaFree(p);
    delete r;
}

// PdSolver::Update (pdSolver.cu:210-232) in a direct mode, handleCollision == false, no fixed bodies; n times
int refs_step(void* h, float dt, float gravity, float tol, int numIterations, int nSteps)
{
    RefS& r = *(RefS*)h;
    for (int s = 0; s < nSteps; s++) {
        if (!r.ready) prepare(r, dt);
        solver_step(r, dt, gravity, tol, numIterations);
        cudaMemcpy(r.X, r.XTilde, sizeof(glm::vec3) * r.nV, cudaMemcpyDeviceToDevice);
    }
    cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}

void refs_stats(void* h, int* pdIterationsLastStep, float* errLastStep)
{
    RefS& r = *(RefS*)h;
    if (pdIterationsLastStep) *pdIterationsLastStep = r.itersLast;
    if (errLastStep) *errLastStep = r.errLast;
}

void refs_get(void* h, float* X, float* V, float* XTilde)
{
    RefS& r = *(RefS*)h;
    size_t v3 = sizeof(glm::vec3) * r.nV;
    if (X) cudaMemcpy(X, r.X, v3, cudaMemcpyDeviceToHost);
    if (V) cudaMemcpy(V, r.V, v3, cudaMemcpyDeviceToHost);
    if (XTilde) cudaMemcpy(XTilde, r.XTilde, v3, cudaMemcpyDeviceToHost);
}

void refs_set(void* h, const float* X, const float* V, const float* XTilde)
{
    RefS& r = *(RefS*)h;
    size_t v3 = sizeof(glm::vec3) * r.nV;
    if (X) cudaMemcpy(r.X, X, v3, cudaMemcpyHostToDevice);
    if (V) cudaMemcpy(r.V, V, v3, cudaMemcpyHostToDevice);
    if (XTilde) cudaMemcpy(r.XTilde, XTilde, v3, cudaMemcpyHostToDevice);
}

// The IPC solver's linear back-ends on their own: LinearSolver<double>::Solve as IPCSolver::SearchDirection calls it
// (IPC/ipc.cu:233-241), kind 1 = CGSolver<double> (linear/cg.cu), kind 2 = PCGJacobiSolver<double> (linear/pcgJacobi.cu); host arrays
// in and out, the COO may hold duplicates.  maxIter / tol <= 0: the classes' defaults.
int refs_linear_solve(int kind, int N, int nz, const int* row, const int* col, const double* val, const double* b, const double* guess, double* x,
                      int maxIter, double tol)
{
    double *dA, *db, *dx, *dg = nullptr; int *dr, *dc;
    cudaMalloc(&dA, 8ull * nz); cudaMalloc(&dr, 4ull * nz); cudaMalloc(&dc, 4ull * nz); cudaMalloc(&db, 8ull * N); cudaMalloc(&dx, 8ull * N);
    cudaMemcpy(dA, val, 8ull * nz, cudaMemcpyHostToDevice); cudaMemcpy(dr, row, 4ull * nz, cudaMemcpyHostToDevice); cudaMemcpy(dc, col, 4ull * nz, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b, 8ull * N, cudaMemcpyHostToDevice); cudaMemset(dx, 0, 8ull * N);
    if (guess) { cudaMalloc(&dg, 8ull * N); cudaMemcpy(dg, guess, 8ull * N, cudaMemcpyHostToDevice); }
    {
        std::unique_ptr<LinearSolver<double>> ls;
        if (kind == 1) ls = (maxIter > 0) ? std::make_unique<CGSolver<double>>(N, maxIter, tol) : std::make_unique<CGSolver<double>>(N);
        else ls = (maxIter > 0) ? std::make_unique<PCGJacobiSolver<double>>(N, maxIter, tol) : std::make_unique<PCGJacobiSolver<double>>(N);
        ls->Solve(N, db, dx, dA, nz, dr, dc, dg);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(x, dx, 8ull * N, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dr); cudaFree(dc); cudaFree(db); cudaFree(dx); if (dg) cudaFree(dg);
    return (int)cudaGetLastError();
}

}  // extern "C"
