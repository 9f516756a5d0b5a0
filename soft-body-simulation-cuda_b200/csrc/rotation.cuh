// Corotational projection for one tet: R = U V^T of F = U S V^T with U, V proper rotations and the
// sign of det(F) carried by the smallest singular value.  This is what the reference obtains from
// svdGLM + `R = U * transpose(V)` (src/simulation/solver/projective/pdUtil.cu:112-122,
// src/simulation/solver/svd.cuh:7-14, external/svd3_cuda/svd3_cuda.h:34-1041).
//
// Two paths:
//  * fast path (det F comfortably positive): Newton iteration X <- (X + X^-T)/2 on F, which
//    converges quadratically to the orthogonal polar factor = U V^T.  ~41 FP32 instructions per
//    iteration, 2-4 iterations for the near-rotations a PD solve sees; the determinant that the
//    inverse needs anyway doubles as the convergence test (after one step every singular value
//    is >= 1, so det - 1 bounds the largest deviation).
//  * slow path (inverted, nearly flat or slowly converging tets): a 4-sweep approximate-Givens
//    Jacobi SVD with sorted singular values and Givens QR -- the McAdams et al. TR1690 scheme the
//    reference uses -- written with explicit single-rounding intrinsics so that it returns the
//    same bits as the CPU oracle (oracle/pd_oracle.c:o_svd3).
#pragma once
#ifdef PD_HOST_EMU
#include "host_emu.hpp"      // tests/emu (TEST ONLY): the packed operations below have host branches, one rounding per half
#else
#include <cuda_runtime.h>
#endif

namespace pdb200 {

struct Mat3 {          // row-major 3x3 in registers
    float m[9];
};

// ------------------------------------------------------------------ slow path: Jacobi SVD
namespace svd_detail {
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float rsq_refined(float x)
{
    float r = __frsqrt_rn(x);
    float h = mul(r, 0.5f);
    float t = mul(r, h);
    t = mul(r, t);
    t = mul(x, t);
    r = add(r, h);
    return sub(r, t);
}

// one Jacobi conjugation; roles as in oracle/pd_oracle.c:jacobi_pair
__device__ __forceinline__ void jacobi_pair(float& s11, float& s21, float& s22, float& s31, float& s32, float& s33,
                                            float& qw, float& qa, float& qb, float& qc)
{
    const float kTiny = 1.e-20f, kFourGammaSq = 5.8284273147583007813f;
    const float kSinPi8 = __uint_as_float(1053028117u), kCosPi8 = __uint_as_float(1064076127u);
    float sh = mul(s21, 0.5f);
    float d = sub(s11, s22);
    float t2 = mul(sh, sh);
    const bool big = t2 >= kTiny;
    sh = big ? sh : 0.0f;
    float ch = big ? d : 1.0f;
    float t1 = mul(sh, sh);
    t2 = mul(ch, ch);
    float t3 = add(t1, t2);
    float t4 = __frsqrt_rn(t3);
    sh = mul(t4, sh);
    ch = mul(t4, ch);
    t1 = mul(kFourGammaSq, t1);
    const bool usePi8 = t2 <= t1;
    sh = usePi8 ? kSinPi8 : sh;
    ch = usePi8 ? kCosPi8 : ch;
    t1 = mul(sh, sh);
    t2 = mul(ch, ch);
    const float c = sub(t2, t1);
    float s = mul(ch, sh);
    s = add(s, s);
    t3 = add(t1, t2);
    s33 = mul(s33, t3); s31 = mul(s31, t3); s32 = mul(s32, t3); s33 = mul(s33, t3);
    t1 = mul(s, s31); t2 = mul(s, s32);
    s31 = mul(c, s31); s32 = mul(c, s32);
    s31 = add(t2, s31); s32 = sub(s32, t1);
    t2 = mul(s, s);
    t1 = mul(s22, t2); t3 = mul(s11, t2);
    t4 = mul(c, c);
    s11 = mul(s11, t4); s22 = mul(s22, t4);
    s11 = add(s11, t1); s22 = add(s22, t3);
    t4 = sub(t4, t2);
    t2 = add(s21, s21);
    s21 = mul(s21, t4);
    t4 = mul(c, s);
    t2 = mul(t2, t4);
    d = mul(d, t4);
    s11 = add(s11, t2); s21 = sub(s21, d); s22 = sub(s22, t2);
    t1 = mul(sh, qa); t2 = mul(sh, qb); t3 = mul(sh, qc);
    sh = mul(sh, qw);
    qw = mul(ch, qw); qa = mul(ch, qa); qb = mul(ch, qb); qc = mul(ch, qc);
    qc = add(qc, sh); qw = sub(qw, t3); qa = add(qa, t2); qb = sub(qb, t1);
}

__device__ __forceinline__ void cond_swap(bool sw, float* B, float* V, int ca, int cb, int cneg, float& na, float& nb)
{
    if (sw) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float t = B[r * 3 + ca]; B[r * 3 + ca] = B[r * 3 + cb]; B[r * 3 + cb] = t;
            t = V[r * 3 + ca]; V[r * 3 + ca] = V[r * 3 + cb]; V[r * 3 + cb] = t;
        }
        float t = na; na = nb; nb = t;
    }
    const float f = sw ? -1.0f : 1.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { B[r * 3 + cneg] = mul(B[r * 3 + cneg], f); V[r * 3 + cneg] = mul(V[r * 3 + cneg], f); }
}

__device__ __forceinline__ void qr_givens(float* B, float* U, int rp, int rq, int cp)
{
    const float kSmall = 1.e-12f;
    const float apiv = B[rp * 3 + cp], aq = B[rq * 3 + cp];
    float sh = mul(aq, aq);
    sh = (sh >= kSmall) ? aq : 0.0f;
    float ch = sub(0.0f, apiv);
    ch = fmaxf(ch, apiv);
    ch = fmaxf(ch, kSmall);
    const bool pos = apiv >= 0.0f;
    float t1 = mul(ch, ch), t2 = mul(sh, sh);
    t2 = add(t1, t2);
    t1 = rsq_refined(t2);
    t1 = mul(t1, t2);
    ch = add(ch, t1);
    if (!pos) { const float t = ch; ch = sh; sh = t; }
    t1 = mul(ch, ch); t2 = mul(sh, sh);
    t2 = add(t1, t2);
    t1 = rsq_refined(t2);
    ch = mul(ch, t1); sh = mul(sh, t1);
    float c = mul(ch, ch), s = mul(sh, sh);
    c = sub(c, s);
    s = mul(sh, ch);
    s = add(s, s);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float x = B[rp * 3 + j], y = B[rq * 3 + j];
        const float u1 = mul(s, x), u2 = mul(s, y);
        x = mul(c, x); y = mul(c, y);
        B[rp * 3 + j] = add(x, u2); B[rq * 3 + j] = sub(y, u1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float x = U[i * 3 + rp], y = U[i * 3 + rq];
        const float u1 = mul(s, x), u2 = mul(s, y);
        x = mul(c, x); y = mul(c, y);
        U[i * 3 + rp] = add(x, u2); U[i * 3 + rq] = sub(y, u1);
    }
}
}  // namespace svd_detail

// Full SVD-based rotation (reference semantics, bit-identical to oracle o_rotation).  Out of line
// (it is the rare path); operands travel by value so that the caller keeps F and R in registers.
__device__ __noinline__ Mat3 rotation_svd_call(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8)
{
    using namespace svd_detail;
    const float A[9] = {a0, a1, a2, a3, a4, a5, a6, a7, a8};
    Mat3 Rm;
    float s11 = mul(A[0], A[0]); s11 = add(mul(A[3], A[3]), s11); s11 = add(mul(A[6], A[6]), s11);
    float s21 = mul(A[1], A[0]); s21 = add(mul(A[4], A[3]), s21); s21 = add(mul(A[7], A[6]), s21);
    float s31 = mul(A[2], A[0]); s31 = add(mul(A[5], A[3]), s31); s31 = add(mul(A[8], A[6]), s31);
    float s22 = mul(A[1], A[1]); s22 = add(mul(A[4], A[4]), s22); s22 = add(mul(A[7], A[7]), s22);
    float s32 = mul(A[2], A[1]); s32 = add(mul(A[5], A[4]), s32); s32 = add(mul(A[8], A[7]), s32);
    float s33 = mul(A[2], A[2]); s33 = add(mul(A[5], A[5]), s33); s33 = add(mul(A[8], A[8]), s33);
    float qw = 1.0f, qx = 0.0f, qy = 0.0f, qz = 0.0f;
#pragma unroll 1
    for (int sweep = 0; sweep < 4; ++sweep) {
        jacobi_pair(s11, s21, s22, s31, s32, s33, qw, qx, qy, qz);
        jacobi_pair(s22, s32, s33, s21, s31, s11, qw, qy, qz, qx);
        jacobi_pair(s33, s31, s11, s32, s21, s22, qw, qz, qx, qy);
    }
    float n2 = mul(qw, qw);
    n2 = add(mul(qx, qx), n2); n2 = add(mul(qy, qy), n2); n2 = add(mul(qz, qz), n2);
    const float rn = rsq_refined(n2);
    qw = mul(qw, rn); qx = mul(qx, rn); qy = mul(qy, rn); qz = mul(qz, rn);
    float t1 = mul(qx, qx), t2 = mul(qy, qy), t3 = mul(qz, qz);
    float v11 = mul(qw, qw);
    float v22 = sub(v11, t1);
    float v33 = sub(v22, t2);
    v33 = add(v33, t3);
    v22 = add(v22, t2); v22 = sub(v22, t3);
    v11 = add(v11, t1); v11 = sub(v11, t2); v11 = sub(v11, t3);
    t1 = add(qx, qx); t2 = add(qy, qy); t3 = add(qz, qz);
    float v32 = mul(qw, t1), v13 = mul(qw, t2), v21 = mul(qw, t3);
    t1 = mul(qy, t1); t2 = mul(qz, t2); t3 = mul(qx, t3);
    const float v12 = sub(t1, v21), v23 = sub(t2, v32), v31 = sub(t3, v13);
    v21 = add(t1, v21); v32 = add(t2, v32); v13 = add(t3, v13);
    float V[9] = {v11, v12, v13, v21, v22, v23, v31, v32, v33};
    float B[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float x = A[r * 3 + 0], y = A[r * 3 + 1], z = A[r * 3 + 2];
        float b1 = mul(v11, x); b1 = add(b1, mul(v21, y)); b1 = add(b1, mul(v31, z));
        float b2 = mul(v12, x); b2 = add(b2, mul(v22, y)); b2 = add(b2, mul(v32, z));
        float b3 = mul(v13, x); b3 = add(b3, mul(v23, y)); b3 = add(b3, mul(v33, z));
        B[r * 3 + 0] = b1; B[r * 3 + 1] = b2; B[r * 3 + 2] = b3;
    }
    float n1 = mul(B[0], B[0]); n1 = add(n1, mul(B[3], B[3])); n1 = add(n1, mul(B[6], B[6]));
    float nb = mul(B[1], B[1]); nb = add(nb, mul(B[4], B[4])); nb = add(nb, mul(B[7], B[7]));
    float n3 = mul(B[2], B[2]); n3 = add(n3, mul(B[5], B[5])); n3 = add(n3, mul(B[8], B[8]));
    cond_swap(n1 < nb, B, V, 0, 1, 1, n1, nb);
    cond_swap(n1 < n3, B, V, 0, 2, 0, n1, n3);
    cond_swap(nb < n3, B, V, 1, 2, 2, nb, n3);
    float U[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    qr_givens(B, U, 0, 1, 0);
    qr_givens(B, U, 0, 2, 0);
    qr_givens(B, U, 1, 2, 1);
    // R = U * V^T exactly as nvcc contracts the reference's glm product (computeLocal SASS: the
    // k=1 product is rounded, k=0 and k=2 are fused), then the det<0 column flip (pdUtil.cu:119-122)
    float* R = Rm.m;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[r * 3 + c] = __fmaf_rn(U[r * 3 + 2], V[c * 3 + 2], __fmaf_rn(U[r * 3 + 0], V[c * 3 + 0], mul(U[r * 3 + 1], V[c * 3 + 1])));
    const float det = add(sub(mul(R[0], sub(mul(R[4], R[8]), mul(R[5], R[7]))),
                              mul(R[3], sub(mul(R[1], R[8]), mul(R[7], R[2])))),
                          mul(R[6], sub(mul(R[1], R[5]), mul(R[4], R[2]))));
    if (det < 0.f) { R[2] = -R[2]; R[5] = -R[5]; R[8] = -R[8]; }
    return Rm;
}
__device__ __forceinline__ void rotation_svd(const Mat3& F, Mat3& R)
{
    R = rotation_svd_call(F.m[0], F.m[1], F.m[2], F.m[3], F.m[4], F.m[5], F.m[6], F.m[7], F.m[8]);
}

// ------------------------------------------------------------------ fast path: Newton polar
constexpr int   kNewtonMaxIter = 8;
constexpr float kNewtonDetMin = 0.02f;     // below this volume ratio the tet takes the slow path
constexpr float kNewtonTol = 4e-4f;        // det(X_k) - 1 < tol  =>  error after the update < tol^2/2

// returns true when converged; Rm holds the rotation
__device__ __forceinline__ bool rotation_newton(const Mat3& Fm, Mat3& Rm)
{
    float x0 = Fm.m[0], x1 = Fm.m[1], x2 = Fm.m[2], x3 = Fm.m[3], x4 = Fm.m[4], x5 = Fm.m[5], x6 = Fm.m[6],
          x7 = Fm.m[7], x8 = Fm.m[8];
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < kNewtonMaxIter; ++it) {
        // cofactor matrix: rows are cross products of the other two rows
        const float c0 = x4 * x8 - x5 * x7, c1 = x5 * x6 - x3 * x8, c2 = x3 * x7 - x4 * x6;
        const float c3 = x7 * x2 - x8 * x1, c4 = x8 * x0 - x6 * x2, c5 = x6 * x1 - x7 * x0;
        const float c6 = x1 * x5 - x2 * x4, c7 = x2 * x3 - x0 * x5, c8 = x0 * x4 - x1 * x3;
        const float det = x0 * c0 + x1 * c1 + x2 * c2;
        if (it == 0 && !(det > kNewtonDetMin)) break;          // inverted / flat: slow path
        const float h = __fdividef(0.5f, det);
        x0 = 0.5f * x0 + h * c0; x1 = 0.5f * x1 + h * c1; x2 = 0.5f * x2 + h * c2;
        x3 = 0.5f * x3 + h * c3; x4 = 0.5f * x4 + h * c4; x5 = 0.5f * x5 + h * c5;
        x6 = 0.5f * x6 + h * c6; x7 = 0.5f * x7 + h * c7; x8 = 0.5f * x8 + h * c8;
        if (it > 0 && det - 1.0f < kNewtonTol) { ok = (det > 0.5f); break; }
    }
    Rm.m[0] = x0; Rm.m[1] = x1; Rm.m[2] = x2; Rm.m[3] = x3; Rm.m[4] = x4; Rm.m[5] = x5;
    Rm.m[6] = x6; Rm.m[7] = x7; Rm.m[8] = x8;
    return ok;
}

// ------------------------------------------------------------------ packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2)
// sm_100 executes add/mul/fma on PAIRS of floats held in an aligned 64-bit register pair as ONE instruction
// (`fma.rn.f32x2` -> FFMA2): one issue slot for two FMAs (measured on B200, scripts/micro/ffma2_bench.cu: the FMA
// pipe is busy two cycles, the scheduler one), each half rounded exactly like the scalar operation.  An operand
// may be a scalar broadcast, have its halves swapped and either half negated at no cost (ptxas folds
// make_float2(s, s), make_float2(a.y, a.x), make_float2(-a.x, a.y) into operand modifiers).  The local kernel is
// bound by instruction issue, so its phase B works on matrix COLUMNS stored as (rows 0 and 1 packed, row 2 scalar).
#ifdef PD_HOST_EMU
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float rcp_approx(float x) { return 1.0f / x; }
#else
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2add(float2 a, float2 b)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2sub(float2 a, float2 b)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
        "sub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

struct Col3 {          // one matrix column: rows 0 and 1 packed, row 2 scalar
    float2 xy;
    float z;
};
// u x v: 2 packed + 2 scalar instructions
__device__ __forceinline__ Col3 cross3(const Col3& u, const Col3& v)
{
    Col3 c;
    const float2 t = f2mul(make_float2(-v.xy.y, v.xy.x), bc2(u.z));            // (-v.y u.z,  v.x u.z)
    c.xy = f2fma(make_float2(u.xy.y, -u.xy.x), bc2(v.z), t);                   // ( u.y v.z - v.y u.z, -u.x v.z + v.x u.z)
    c.z = fmaf(u.xy.x, v.xy.y, -(u.xy.y * v.xy.x));
    return c;
}
__device__ __forceinline__ float dot3c(const Col3& a, const Col3& b)
{
    const float2 p = f2mul(a.xy, b.xy);
    return fmaf(a.z, b.z, p.x) + p.y;
}

// Packed fast path of the corotational projection, the polar Newton iteration X <- (X + X^-T)/2 of
// rotation_newton() unrolled for the two steps a PD solve needs and carried in SCALED form (Y_k = 2^k X_k, so that
// an update is one fused multiply-add per entry:  Y1 = F + cof(F)/det(F),  Y2 = Y1 + 4 cof(Y1)/det(Y1) = 4 X2).
// Works on the COLUMNS of F (the polar factor of F^T is R^T, and cof(X) has the cross products of X's columns as
// its columns).  Same acceptance tests as rotation_newton(): det F > kNewtonDetMin, det X1 - 1 < kNewtonTol,
// det X1 > 0.5.  Returns false (Y untouched or partial) when the tet needs the general path.
__device__ __forceinline__ bool rotation_newton2_packed(const Col3& f0, const Col3& f1, const Col3& f2, Col3& y0, Col3& y1, Col3& y2)
{
    Col3 c0 = cross3(f1, f2), c1 = cross3(f2, f0), c2 = cross3(f0, f1);
    const float det0 = dot3c(f0, c0);
    if (!(det0 > kNewtonDetMin)) return false;
    const float g0 = rcp_approx(det0);
    y0.xy = f2fma(c0.xy, bc2(g0), f0.xy); y0.z = fmaf(c0.z, g0, f0.z);
    y1.xy = f2fma(c1.xy, bc2(g0), f1.xy); y1.z = fmaf(c1.z, g0, f1.z);
    y2.xy = f2fma(c2.xy, bc2(g0), f2.xy); y2.z = fmaf(c2.z, g0, f2.z);
    c0 = cross3(y1, y2); c1 = cross3(y2, y0); c2 = cross3(y0, y1);         // = 4 cof(X1)
    const float det1 = dot3c(y0, c0);                                        // = 8 det(X1)
    if (!(det1 < 8.0f * (1.0f + kNewtonTol) && det1 > 4.0f)) return false;
    const float g1 = 4.0f * rcp_approx(det1);
    y0.xy = f2fma(c0.xy, bc2(g1), y0.xy); y0.z = fmaf(c0.z, g1, y0.z);
    y1.xy = f2fma(c1.xy, bc2(g1), y1.xy); y1.z = fmaf(c1.z, g1, y1.z);
    y2.xy = f2fma(c2.xy, bc2(g1), y2.xy); y2.z = fmaf(c2.z, g1, y2.z);     // = 4 X2
    return true;
}

// ROT_MODE 0: Newton fast path with SVD fallback (product default); 1: always the SVD (faithful)
template <int ROT_MODE>
__device__ __forceinline__ void corotation(const Mat3& F, Mat3& R)
{
    if (ROT_MODE == 2) { R = F; return; }     // measurement only (scripts/experiments): no projection at all
    if (ROT_MODE == 0) {
        if (rotation_newton(F, R)) return;
    }
    rotation_svd(F, R);
}

}  // namespace pdb200
