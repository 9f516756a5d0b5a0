// TEST INFRASTRUCTURE (oracle/_ref): the reference's own continuous-collision arithmetic -- collision/intersections.cu
// (ccdCollisionTest and everything under it) and simulation/collisionUtil.cu (CCDKernel) -- compiled VERBATIM from
// /root/reference and wrapped for the parity tests.  The reference's broad phase (bvh.cu / broadphase.cu / narrowphase.cu) is
// welded to its OpenGL classes (bvh.h -> openglcontext/wireframe.h) and cannot be compiled here; what it computes with
// integers and box comparisons is restated in oracle/pd_oracle.c (mesh_collision) instead.
#include <cstdint>
#include <cuda_runtime.h>

#include <collision/intersections.cu>
#include <simulation/collisionUtil.cu>

// one thread per query: type 1 = vertex-face, 2 = edge-edge (QueryType, aabb.h:54-58); exactly detectCollisionNarrow's body
// (narrowphase.cu:61-72)
__global__ void refc_narrow(int n, const int* type, const uint32_t* v, const glm::vec3* X, const glm::vec3* XT, float* toi, glm::vec3* nor)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Query q;
    q.type = type[i] == 2 ? QueryType::EE : QueryType::VF;
    q.v0 = v[4 * i]; q.v1 = v[4 * i + 1]; q.v2 = v[4 * i + 2]; q.v3 = v[4 * i + 3];
    glm::vec3 normal(0.f);
    q.toi = ccdCollisionTest<float>(q, X, XT, normal);
    toi[i] = (float)q.toi;
    nor[i] = normal;
}

extern "C" {
// host arrays in, host arrays out; returns a CUDA error code
int refc_ccd_queries(int n, const int* type, const uint32_t* v, int nV, const float* X, const float* XT, float* toi, float* normals)
{
    int* dT; uint32_t* dV; glm::vec3 *dX, *dXT, *dN; float* dToi;
    cudaMalloc(&dT, 4 * (size_t)n); cudaMalloc(&dV, 16 * (size_t)n); cudaMalloc(&dX, 12 * (size_t)nV); cudaMalloc(&dXT, 12 * (size_t)nV);
    cudaMalloc(&dN, 12 * (size_t)n); cudaMalloc(&dToi, 4 * (size_t)n);
    cudaMemcpy(dT, type, 4 * (size_t)n, cudaMemcpyHostToDevice); cudaMemcpy(dV, v, 16 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X, 12 * (size_t)nV, cudaMemcpyHostToDevice); cudaMemcpy(dXT, XT, 12 * (size_t)nV, cudaMemcpyHostToDevice);
    refc_narrow<<<(n + 127) / 128, 128>>>(n, dT, dV, dX, dXT, dToi, dN);
    cudaMemcpy(toi, dToi, 4 * (size_t)n, cudaMemcpyDeviceToHost); cudaMemcpy(normals, dN, 12 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(dT); cudaFree(dV); cudaFree(dX); cudaFree(dXT); cudaFree(dN); cudaFree(dToi);
    return (int)cudaGetLastError();
}
// CCDKernel<float> (collisionUtil.cu:49-70) as PdSolver::Update launches it (pdSolver.cu:222-223); X, V updated in place
int refc_ccd_kernel(int nV, float* X, const float* XT, float* V, const float* tI, const float* normals, float muT, float muN, float dt, int threadsPerBlock)
{
    glm::vec3 *dX, *dXT, *dV, *dN; float* dTI;
    cudaMalloc(&dX, 12 * (size_t)nV); cudaMalloc(&dXT, 12 * (size_t)nV); cudaMalloc(&dV, 12 * (size_t)nV); cudaMalloc(&dN, 12 * (size_t)nV); cudaMalloc(&dTI, 4 * (size_t)nV);
    cudaMemcpy(dX, X, 12 * (size_t)nV, cudaMemcpyHostToDevice); cudaMemcpy(dXT, XT, 12 * (size_t)nV, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, V, 12 * (size_t)nV, cudaMemcpyHostToDevice); cudaMemcpy(dN, normals, 12 * (size_t)nV, cudaMemcpyHostToDevice);
    cudaMemcpy(dTI, tI, 4 * (size_t)nV, cudaMemcpyHostToDevice);
    CCDKernel<float><<<(nV + threadsPerBlock - 1) / threadsPerBlock, threadsPerBlock>>>(dX, dXT, dV, dTI, dN, muT, muN, nV, dt);
    cudaMemcpy(X, dX, 12 * (size_t)nV, cudaMemcpyDeviceToHost); cudaMemcpy(V, dV, 12 * (size_t)nV, cudaMemcpyDeviceToHost);
    cudaFree(dX); cudaFree(dXT); cudaFree(dV); cudaFree(dN); cudaFree(dTI);
    return (int)cudaGetLastError();
}
}
