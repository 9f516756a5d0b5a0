// Microbenchmark (B200): issue cost of packed FP32 (FFMA2/FMUL2/FADD2) against scalar FFMA, alone and mixed with
// integer ALU work, to decide whether packing the local kernel's phase B relieves its issue-slot limit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
constexpr int NACC = 8, INNER = 64;
// MODE 0: scalar FFMA (3-reg), 1: FFMA2, 2: scalar FFMA + 1 LOP3 per FFMA, 3: FFMA2 + 1 LOP3 per FFMA2,
// 4: FFMA2 with scalar-broadcast operand, 5: FFMA2 + 2 LOP3 per FFMA2
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s, unsigned m)
{
    float2 acc[NACC];
    unsigned u[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) { acc[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f); u[j] = threadIdx.x + j; }
    float2 b = make_float2(s, s * 0.999f), c = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < INNER / NACC; ++r) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                if (MODE == 0 || MODE == 2) acc[j].x = fmaf(acc[j].x, b.x, c.x);
                if (MODE == 1 || MODE == 3 || MODE == 5) acc[j] = ffma2(acc[j], b, c);
                if (MODE == 4) acc[j] = ffma2(acc[j], make_float2(s, s), c);
                if (MODE == 2 || MODE == 3 || MODE == 5) u[j] = (u[j] ^ m) & (u[(j + 1) % NACC] | m);
                if (MODE == 5) u[j] = (u[j] | (m >> 1)) ^ u[(j + 3) % NACC];
            }
        }
    }
    float r = 0.f; unsigned q = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) { r += acc[j].x + acc[j].y; q ^= u[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)q;
}
template <int MODE>
void run(const char* name, float* d, int fpPerInst, int aluPerInst)
{
    const int iters = 2000, grid = 148 * 8;
    k<MODE><<<grid, 256>>>(d, 10, 0.999f, 0x5555u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, 0.999f, 0x5555u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warpInst = (double)grid * 8 * iters * INNER;           // FP instructions (warp level)
    const double perSmspPerNs = warpInst / (148.0 * 4) / (ms * 1e6);
    printf("%-44s %8.3f ms  FP warp-inst/SMSP/ns %.3f  (x%d lanes-ops)  fp32 FMA TFLOP/s %.1f  +alu/inst %d\n", name, ms, perSmspPerNs, fpPerInst,
           warpInst * 32 * fpPerInst * 2 / (ms * 1e-3) / 1e12, aluPerInst);
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA", d, 1, 0);
    run<1>("FFMA2", d, 2, 0);
    run<4>("FFMA2 scalar-broadcast operand", d, 2, 0);
    run<2>("scalar FFMA + 1 LOP3", d, 1, 1);
    run<3>("FFMA2 + 1 LOP3", d, 2, 1);
    run<5>("FFMA2 + 2 LOP3", d, 2, 2);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock attr %d kHz\n", clk);
    return 0;
}
