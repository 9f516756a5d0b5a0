/*
 * pd_oracle.c -- CPU restatement of the reference's float PD step (see pd_oracle.h).
 * TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off keeps one rounding per float operation, which is what the
 * reference's SVD does explicitly (__fadd_rn/__fsub_rn, svd3_cuda.h:65-97).  Everywhere
 * else the reference is plain C++/glm that nvcc compiles with FMA contraction; the fused
 * operations are written out here with fmaf()/fma() exactly where nvcc 12.9 places them in
 * the sm_100a build of the reference kernels (read off `cuobjdump -sass oracle/_ref/libpd_ref.so`):
 *   a0*b0 + a1*b1 + a2*b2  (glm mat3*mat3, glm::dot)   ->  fma(a2,b2, fma(a0,b0, a1*b1))
 *   a*b - c*d                                          ->  fma(a,b, -(c*d))
 * and the per-kernel forms quoted at each use below.
 */
#include "pd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * 3x3 SVD, McAdams/Selle/Tamstorf/Teran/Sifakis TR1690 as shipped in
 * external/svd3_cuda/svd3_cuda.h:34-1041 (Kui Wu's CUDA port; vendored, no version tag).
 * Restated with helper routines; operation order and rounding follow the reference.
 * ---------------------------------------------------------------------------------------- */

/* svd3_cuda.h:25-30 */
/* the reference stores sin/cos(pi/8) as integer bit patterns (svd3_cuda.h:26-27); the cosine
 * is 1 ulp above the correctly rounded value, so the patterns are reproduced exactly */
/* a0*b0 + a1*b1 + a2*b2 as contracted by nvcc: the k=1 product is rounded, k=0 and k=2 are fused */
#define DOT3_NV(a0, b0, a1, b1, a2, b2) fmaf((a2), (b2), fmaf((a0), (b0), (a1) * (b1)))

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define O_SIN_PI8   bits2f(1053028117u)
#define O_COS_PI8   bits2f(1064076127u)
#define O_SMALL     1.e-12f
#define O_TINY      1.e-20f
#define O_4GAMMA2   5.8284273147583007813f

/* __frsqrt_rn: IEEE-rounded 1/sqrt(x) (svd3_cuda.h:119) */
static inline float rsqrt_rn(float x) { return (float)(1.0 / sqrt((double)x)); }

/* rsqrt followed by one Newton step, svd3_cuda.h:424-431 */
static inline float rsqrt_refined(float x)
{
    float r = rsqrt_rn(x);
    float h = r * 0.5f;
    float t = r * h;
    t = r * t;
    t = x * t;
    r = r + h;
    r = r - t;
    return r;
}

/* One Jacobi conjugation on the symmetric matrix S for the pair whose entries are
 * (*p11,*p21,*p22); (*p31,*p32,*p33) are the remaining off-diagonals / diagonal in the
 * cyclic role order the reference uses; (qa,qb,qc) are the quaternion vector components
 * in the same cyclic order.  svd3_cuda.h:107-196 (pair 1-2), :198-283, :285-370. */
static void jacobi_pair(float *p11, float *p21, float *p22, float *p31, float *p32, float *p33,
                        float *qw, float *qa, float *qb, float *qc)
{
    float s11 = *p11, s21 = *p21, s22 = *p22, s31 = *p31, s32 = *p32, s33 = *p33;
    /* approximate Givens quaternion, svd3_cuda.h:107-137 */
    float sh = s21 * 0.5f;
    float d = s11 - s22;
    float t2 = sh * sh;
    int big = (t2 >= O_TINY);
    sh = big ? sh : 0.0f;
    float ch = big ? d : 1.0f;
    float t1 = sh * sh;
    t2 = ch * ch;
    float t3 = t1 + t2;
    float t4 = rsqrt_rn(t3);
    sh = t4 * sh;
    ch = t4 * ch;
    t1 = O_4GAMMA2 * t1;
    int use_pi8 = (t2 <= t1);
    sh = use_pi8 ? O_SIN_PI8 : sh;
    ch = use_pi8 ? O_COS_PI8 : ch;
    t1 = sh * sh;
    t2 = ch * ch;
    float c = t2 - t1;
    float s = ch * sh;
    s = s + s;
    /* conjugation, svd3_cuda.h:146-176 */
    t3 = t1 + t2;
    s33 = s33 * t3;
    s31 = s31 * t3;
    s32 = s32 * t3;
    s33 = s33 * t3;
    t1 = s * s31;
    t2 = s * s32;
    s31 = c * s31;
    s32 = c * s32;
    s31 = t2 + s31;
    s32 = s32 - t1;
    t2 = s * s;
    t1 = s22 * t2;
    t3 = s11 * t2;
    t4 = c * c;
    s11 = s11 * t4;
    s22 = s22 * t4;
    s11 = s11 + t1;
    s22 = s22 + t3;
    t4 = t4 - t2;
    t2 = s21 + s21;
    s21 = s21 * t4;
    t4 = c * s;
    t2 = t2 * t4;
    d = d * t4;
    s11 = s11 + t2;
    s21 = s21 - d;
    s22 = s22 - t2;
    /* cumulative rotation, svd3_cuda.h:182-196 */
    float w = *qw, a = *qa, b = *qb, cc = *qc;
    t1 = sh * a;
    t2 = sh * b;
    t3 = sh * cc;
    sh = sh * w;
    w = ch * w;
    a = ch * a;
    b = ch * b;
    cc = ch * cc;
    cc = cc + sh;
    w = w - t3;
    a = a + t2;
    b = b - t1;
    *p11 = s11; *p21 = s21; *p22 = s22; *p31 = s31; *p32 = s32; *p33 = s33;
    *qw = w; *qa = a; *qb = b; *qc = cc;
}

/* conditional column swap of B and V with the sign flip that keeps V a rotation,
 * svd3_cuda.h:563-627 (and the two that follow) */
static void cond_swap_cols(int do_swap, float B[9], float V[9], int ca, int cb, int cneg,
                           float *na, float *nb)
{
    if (do_swap) {
        for (int r = 0; r < 3; r++) {
            float t = B[r * 3 + ca]; B[r * 3 + ca] = B[r * 3 + cb]; B[r * 3 + cb] = t;
            t = V[r * 3 + ca]; V[r * 3 + ca] = V[r * 3 + cb]; V[r * 3 + cb] = t;
        }
        float t = *na; *na = *nb; *nb = t;
    }
    float f = 1.0f + (do_swap ? -2.0f : 0.0f);
    for (int r = 0; r < 3; r++) {
        B[r * 3 + cneg] = B[r * 3 + cneg] * f;
        V[r * 3 + cneg] = V[r * 3 + cneg] * f;
    }
}

/* one Givens rotation of the QR step: zero B[rq][cp] against pivot B[rp][cp];
 * rotates rows rp,rq of B and columns rp,rq of U.  svd3_cuda.h:707-812 etc. */
static void qr_givens(float B[9], float U[9], int rp, int rq, int cp)
{
    float apiv = B[rp * 3 + cp], aq = B[rq * 3 + cp];
    float sh = aq * aq;
    sh = (sh >= O_SMALL) ? aq : 0.0f;
    float ch = 0.0f - apiv;
    ch = fmaxf(ch, apiv);
    ch = fmaxf(ch, O_SMALL);
    int pos = (apiv >= 0.0f);
    float t1 = ch * ch;
    float t2 = sh * sh;
    t2 = t1 + t2;
    t1 = rsqrt_refined(t2);
    t1 = t1 * t2;
    ch = ch + t1;
    if (!pos) { float t = ch; ch = sh; sh = t; }
    t1 = ch * ch;
    t2 = sh * sh;
    t2 = t1 + t2;
    t1 = rsqrt_refined(t2);
    ch = ch * t1;
    sh = sh * t1;
    float c = ch * ch;
    float s = sh * sh;
    c = c - s;
    s = sh * ch;
    s = s + s;
    for (int j = 0; j < 3; j++) {
        float x = B[rp * 3 + j], y = B[rq * 3 + j];
        float u1 = s * x, u2 = s * y;
        x = c * x; y = c * y;
        B[rp * 3 + j] = x + u2;
        B[rq * 3 + j] = y - u1;
    }
    for (int i = 0; i < 3; i++) {
        float x = U[i * 3 + rp], y = U[i * 3 + rq];
        float u1 = s * x, u2 = s * y;
        x = c * x; y = c * y;
        U[i * 3 + rp] = x + u2;
        U[i * 3 + rq] = y - u1;
    }
}

void o_svd3(const float A[9], float U[9], float S[3], float V[9])
{
    const float a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[3], a22 = A[4], a23 = A[5],
                a31 = A[6], a32 = A[7], a33 = A[8];
    /* normal equations S = A^T A, svd3_cuda.h:62-96 */
    float s11 = a11 * a11; s11 = a21 * a21 + s11; s11 = a31 * a31 + s11;
    float s21 = a12 * a11; s21 = a22 * a21 + s21; s21 = a32 * a31 + s21;
    float s31 = a13 * a11; s31 = a23 * a21 + s31; s31 = a33 * a31 + s31;
    float s22 = a12 * a12; s22 = a22 * a22 + s22; s22 = a32 * a32 + s22;
    float s32 = a13 * a12; s32 = a23 * a22 + s32; s32 = a33 * a32 + s32;
    float s33 = a13 * a13; s33 = a23 * a23 + s33; s33 = a33 * a33 + s33;
    float qw = 1.0f, qx = 0.0f, qy = 0.0f, qz = 0.0f;
    /* 4 cyclic Jacobi sweeps, svd3_cuda.h:103-371 */
    for (int sweep = 0; sweep < 4; sweep++) {
        jacobi_pair(&s11, &s21, &s22, &s31, &s32, &s33, &qw, &qx, &qy, &qz);
        jacobi_pair(&s22, &s32, &s33, &s21, &s31, &s11, &qw, &qy, &qz, &qx);
        jacobi_pair(&s33, &s31, &s11, &s32, &s21, &s22, &qw, &qz, &qx, &qy);
    }
    /* normalise quaternion, svd3_cuda.h:417-436 */
    float n2 = qw * qw;
    n2 = qx * qx + n2;
    n2 = qy * qy + n2;
    n2 = qz * qz + n2;
    float rn = rsqrt_refined(n2);
    qw = qw * rn; qx = qx * rn; qy = qy * rn; qz = qz * rn;
    /* quaternion -> V, svd3_cuda.h:442-470 */
    float t1 = qx * qx, t2 = qy * qy, t3 = qz * qz;
    float v11 = qw * qw;
    float v22 = v11 - t1;
    float v33 = v22 - t2;
    v33 = v33 + t3;
    v22 = v22 + t2;
    v22 = v22 - t3;
    v11 = v11 + t1;
    v11 = v11 - t2;
    v11 = v11 - t3;
    t1 = qx + qx; t2 = qy + qy; t3 = qz + qz;
    float v32 = qw * t1, v13 = qw * t2, v21 = qw * t3;
    t1 = qy * t1; t2 = qz * t2; t3 = qx * t3;
    float v12 = t1 - v21, v23 = t2 - v32, v31 = t3 - v13;
    v21 = t1 + v21; v32 = t2 + v32; v13 = t3 + v13;
    V[0] = v11; V[1] = v12; V[2] = v13; V[3] = v21; V[4] = v22; V[5] = v23;
    V[6] = v31; V[7] = v32; V[8] = v33;
    /* B = A V, svd3_cuda.h:476-536 */
    float B[9];
    for (int r = 0; r < 3; r++) {
        float x = A[r * 3 + 0], y = A[r * 3 + 1], z = A[r * 3 + 2];
        float b1 = v11 * x; b1 = b1 + v21 * y; b1 = b1 + v31 * z;
        float b2 = v12 * x; b2 = b2 + v22 * y; b2 = b2 + v32 * z;
        float b3 = v13 * x; b3 = b3 + v23 * y; b3 = b3 + v33 * z;
        B[r * 3 + 0] = b1; B[r * 3 + 1] = b2; B[r * 3 + 2] = b3;
    }
    /* sort columns by norm, svd3_cuda.h:542-693 */
    float n1 = B[0] * B[0]; n1 = n1 + B[3] * B[3]; n1 = n1 + B[6] * B[6];
    float nn2 = B[1] * B[1]; nn2 = nn2 + B[4] * B[4]; nn2 = nn2 + B[7] * B[7];
    float n3 = B[2] * B[2]; n3 = n3 + B[5] * B[5]; n3 = n3 + B[8] * B[8];
    cond_swap_cols(n1 < nn2, B, V, 0, 1, 1, &n1, &nn2);
    cond_swap_cols(n1 < n3, B, V, 0, 2, 0, &n1, &n3);
    cond_swap_cols(nn2 < n3, B, V, 1, 2, 2, &nn2, &n3);
    /* QR by Givens, svd3_cuda.h:699-1025 */
    U[0] = 1; U[1] = 0; U[2] = 0; U[3] = 0; U[4] = 1; U[5] = 0; U[6] = 0; U[7] = 0; U[8] = 1;
    qr_givens(B, U, 0, 1, 0);
    qr_givens(B, U, 0, 2, 0);
    qr_givens(B, U, 1, 2, 1);
    S[0] = B[0]; S[1] = B[4]; S[2] = B[8];
}

/* R = U * transpose(V); flip column 2 if det<0 (pdUtil.cu:115-122).  Row-major. */
void o_rotation(const float F[9], float R[9])
{
    float U[9], S[3], V[9];
    o_svd3(F, U, S, V);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            /* glm mat3*mat3 as nvcc contracts it (computeLocal SASS): fma(a2,b2, fma(a0,b0, a1*b1)) */
            R[r * 3 + c] = DOT3_NV(U[r * 3 + 0], V[c * 3 + 0], U[r * 3 + 1], V[c * 3 + 1], U[r * 3 + 2], V[c * 3 + 2]);
        }
    float det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[3] * (R[1] * R[8] - R[7] * R[2]) +
                R[6] * (R[1] * R[5] - R[4] * R[2]);
    if (det < 0) { R[2] = -R[2]; R[5] = -R[5]; R[8] = -R[8]; }
}

/* ------------------------------------------------------------------------------------------
 * rest shape: solverUtil.cuh:98-116 with glm::inverse (glm/detail/type_mat3x3.inl:37-57)
 * and glm::determinant (glm/detail/func_matrix.inl:231-240).  glm is column-major m[c][r];
 * here Dm, DmInv are row-major d[r*3+c].
 * ---------------------------------------------------------------------------------------- */
static void rest_one(const float *X, const uint32_t *t, float *Bi, float *V0)
{
    const float *x0 = X + 3 * t[0], *x1 = X + 3 * t[1], *x2 = X + 3 * t[2], *x3 = X + 3 * t[3];
    /* m[c][r]: column c = x_{c+1} - x0 */
    float m[3][3];
    for (int r = 0; r < 3; r++) { m[0][r] = x1[r] - x0[r]; m[1][r] = x2[r] - x0[r]; m[2][r] = x3[r] - x0[r]; }
    /* glm::inverse / glm::determinant as nvcc contracts them in computeInvDmV0 (SASS of the sm_100a
     * build): every 2x2 minor a*b - c*d is fma(a, b, -(c*d)); det = fma(m20, C2, fma(m00, C0, -(m10*C1))) */
#define MINOR(a, b, c, d) fmaf((a), (b), -((c) * (d)))
    float c0 = MINOR(m[1][1], m[2][2], m[2][1], m[1][2]);
    float c1 = MINOR(m[0][1], m[2][2], m[2][1], m[0][2]);
    float c2 = MINOR(m[0][1], m[1][2], m[1][1], m[0][2]);
    float det = fmaf(m[2][0], c2, fmaf(m[0][0], c0, -(m[1][0] * c1)));
    float ood = 1.0f / det;
    float inv[3][3]; /* inv[c][r] */
    inv[0][0] = +c0 * ood;
    inv[1][0] = -(MINOR(m[1][0], m[2][2], m[2][0], m[1][2]) * ood);
    inv[2][0] = +MINOR(m[1][0], m[2][1], m[2][0], m[1][1]) * ood;
    inv[0][1] = -(c1 * ood);
    inv[1][1] = +MINOR(m[0][0], m[2][2], m[2][0], m[0][2]) * ood;
    inv[2][1] = -(MINOR(m[0][0], m[2][1], m[2][0], m[0][1]) * ood);
    inv[0][2] = +c2 * ood;
    inv[1][2] = -(MINOR(m[0][0], m[1][2], m[1][0], m[0][2]) * ood);
    inv[2][2] = +MINOR(m[0][0], m[1][1], m[1][0], m[0][1]) * ood;
#undef MINOR
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Bi[r * 3 + c] = inv[c][r];
    *V0 = fabsf(det) / 6.0f;
}

void o_rest_shape(const float *X, const uint32_t *Tet, int nT, float *DmInv, float *V0)
{
    for (int t = 0; t < nT; t++) rest_one(X, Tet + 4 * t, DmInv + 9 * t, V0 + t);
}

/* ------------------------------------------------------------------------------------------
 * transforms (glm::translate/rotate/scale, gtc/matrix_transform.inl:40-86,114-125)
 * ---------------------------------------------------------------------------------------- */
static void mat4_identity(float M[16]) { memset(M, 0, 64); M[0] = M[5] = M[10] = M[15] = 1.0f; }

static void mat4_translate(float M[16], const float v[3])
{   /* Result[3] = m[0]*v0 + m[1]*v1 + m[2]*v2 + m[3] */
    for (int r = 0; r < 4; r++)
        M[12 + r] = M[0 + r] * v[0] + M[4 + r] * v[1] + M[8 + r] * v[2] + M[12 + r];
}

static void mat4_scale(float M[16], const float v[3])
{
    for (int r = 0; r < 4; r++) { M[0 + r] *= v[0]; M[4 + r] *= v[1]; M[8 + r] *= v[2]; }
}

static void mat4_rotate(float M[16], float angle, int axis_id)
{   /* gtc/matrix_transform.inl:52-86 with a unit coordinate axis */
    float axis[3] = {0, 0, 0};
    axis[axis_id] = 1.0f;
    float c = cosf(angle), s = sinf(angle);
    float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
    R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
    R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
    R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
    R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    float out[12];
    for (int k = 0; k < 3; k++)
        for (int r = 0; r < 4; r++)
            out[k * 4 + r] = M[0 + r] * R[k][0] + M[4 + r] * R[k][1] + M[8 + r] * R[k][2];
    memcpy(M, out, sizeof(out));
}

static float deg2rad(float d) { return d * 0.01745329251994329576923690768489f; }

void o_model_matrix(const float pos[3], const float rot[3], const float scale[3],
                    int soft_body_order, float M[16])
{
    mat4_identity(M);
    mat4_translate(M, pos);
    if (soft_body_order) mat4_scale(M, scale); /* dataLoader.cu:215-220: T*S*Rx*Ry*Rz */
    mat4_rotate(M, deg2rad(rot[0]), 0);
    mat4_rotate(M, deg2rad(rot[1]), 1);
    mat4_rotate(M, deg2rad(rot[2]), 2);
    if (!soft_body_order) mat4_scale(M, scale); /* utilities.cpp:141-150: T*Rx*Ry*Rz*S */
}

void o_transform_vertices(float *X, int nV, const float M[16])
{   /* glm mat4*vec4: m[0]*x + m[1]*y + m[2]*z + m[3]*w (type_mat4x4.inl operator*) */
    for (int i = 0; i < nV; i++) {
        float x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
        for (int r = 0; r < 3; r++) {
            float m0 = M[0 + r] * x, m1 = M[4 + r] * y, m2 = M[8 + r] * z, m3 = M[12 + r] * 1.0f;
            X[3 * i + r] = (m0 + m1) + (m2 + m3);
        }
    }
}

void o_plane_up(const float M[16], float up[3])
{   /* normalize(vec3(transpose(inverse(model)) * (0,1,0,0))) = normalize(row 1 of inverse(A)),
     * A = upper-left 3x3 (the model is affine, last row 0 0 0 1). */
    double a[3][3];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) a[r][c] = M[c * 4 + r];
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                 a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    float v[3];
    v[0] = (float)(-(a[1][0] * a[2][2] - a[1][2] * a[2][0]) / det);
    v[1] = (float)((a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det);
    v[2] = (float)(-(a[0][0] * a[1][2] - a[0][2] * a[1][0]) / det);
    float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    up[0] = v[0] / len; up[1] = v[1] / len; up[2] = v[2] / len;
}

void o_cylinder_axis(const float M[16], float axis[3])
{   /* normalize(model * (0,1,0,0)) = normalize(column 1) ; fixedBodyData.cu:116 */
    float v[3] = {M[4], M[5], M[6]};
    float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + M[7] * M[7]);
    axis[0] = v[0] / len; axis[1] = v[1] / len; axis[2] = v[2] / len;
}

/* ------------------------------------------------------------------------------------------
 * TetGen readers: dataLoader.cu:131-173 (.node) and :38-66 (.ele)
 * ---------------------------------------------------------------------------------------- */
int o_load_node(const char *path, int centralize, float **Xout)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    int n = 0;
    if (!fgets(line, sizeof line, f) || sscanf(line, "%d", &n) != 1 || n <= 0) { fclose(f); return -2; }
    float *X = (float *)calloc((size_t)n * 3, sizeof(float));
    float cx = 0, cy = 0, cz = 0;
    for (int i = 0; i < n && fgets(line, sizeof line, f); i++) {
        int idx; float x = 0, y = 0, z = 0;
        sscanf(line, "%d %f %f %f", &idx, &x, &y, &z);
        X[3 * i] = x; X[3 * i + 1] = y; X[3 * i + 2] = z;
        cx += x; cy += y; cz += z;
    }
    fclose(f);
    if (centralize) {   /* subtract centroid, then swap y and z (dataLoader.cu:162-169) */
        cx /= (float)n; cy /= (float)n; cz /= (float)n;
        for (int i = 0; i < n; i++) {
            X[3 * i] -= cx; X[3 * i + 1] -= cy; X[3 * i + 2] -= cz;
            float t = X[3 * i + 1]; X[3 * i + 1] = X[3 * i + 2]; X[3 * i + 2] = t;
        }
    }
    *Xout = X;
    return n;
}

int o_load_ele(const char *path, int start_index, uint32_t **Tout)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    int n = 0;
    if (!fgets(line, sizeof line, f) || sscanf(line, "%d", &n) != 1 || n <= 0) { fclose(f); return -2; }
    uint32_t *T = (uint32_t *)calloc((size_t)n * 4, sizeof(uint32_t));
    for (int t = 0; t < n && fgets(line, sizeof line, f); t++) {
        int a, b, c, d, e;   /* row is "a b c d e", tet = (b,c,d,e) - startIndex */
        sscanf(line, "%d %d %d %d %d", &a, &b, &c, &d, &e);
        T[4 * t] = (uint32_t)(b - start_index); T[4 * t + 1] = (uint32_t)(c - start_index);
        T[4 * t + 2] = (uint32_t)(d - start_index); T[4 * t + 3] = (uint32_t)(e - start_index);
    }
    fclose(f);
    *Tout = T;
    return n;
}

void o_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * scene + solver state
 * ---------------------------------------------------------------------------------------- */
struct o_scene {
    int nV, nT, numDBC;
    float *X, *X0, *XTilde, *V, *DBCX, *ExtForce;    /* 3 nV */
    float *mass, *DBC;                                 /* nV   */
    float *moreDBC, *OffsetX; float drag_target[3];    /* mouse drag: def.h:31-32, MouseSelection::target def.h:14-18 */
    float *mu;                                         /* nT   */
    uint32_t *Tet;
    float *DmInv, *V0;                                 /* FEMSolver ctor products */
    /* PdSolver private buffers (pdSolver.h:31-43) */
    float *massDt_2s, *matrix_diag, *sn, *sn_old, *b, *next_x, *prev_x;
    float *H;                /* per-tet 3x4 contribution buffer (parallel local step) */
    int *inc_ptr, *inc;      /* vertex -> (tet*4+corner), ascending = sequential scatter order */
    int ready;               /* Solver::solverReady */
    float prepared_dt;
    /* direct / CG back-ends */
    int *rowptr, *col; float *val; int nnz;
    double *chol;            /* dense lower Cholesky factor of A^ (double), nV*nV */
    float *cg_r, *cg_z, *cg_p, *cg_q;
    /* fixed bodies (owned copies) */
    o_fixed_bodies fb;
    float *fbbuf;
    int last_pd_iters, last_inner_iters;
    /* mesh-mesh collision (SolverData::Tri / dev_TriFathers / dev_tIs / dev_Normals, def.h:22,40-41,44) */
    int col_on, nTris; uint32_t *Tri, *TriFathers; float *tI, *Normals; long long col_pairs;
    /* f64 twin state */
    double *Xd, *Vd, *XTd;
};

static float *fdup(const float *src, size_t n)
{
    float *p = (float *)malloc(n * sizeof(float) + 4);
    if (src) memcpy(p, src, n * sizeof(float)); else memset(p, 0, n * sizeof(float));
    return p;
}

o_scene *o_scene_create(int nV, int nT, const float *X, const uint32_t *Tet, const float *mass,
                        const float *mu, const float *DBC, const o_fixed_bodies *fb)
{
    o_scene *s = (o_scene *)calloc(1, sizeof(o_scene));
    s->nV = nV; s->nT = nT;
    /* DataLoader::AllocData dataLoader.cu:291-378: X0 = DBCX = XTilde = X, V = 0 */
    s->X = fdup(X, 3 * (size_t)nV); s->X0 = fdup(X, 3 * (size_t)nV); s->XTilde = fdup(X, 3 * (size_t)nV);
    s->DBCX = fdup(X, 3 * (size_t)nV);
    s->V = fdup(NULL, 3 * (size_t)nV); s->ExtForce = fdup(NULL, 3 * (size_t)nV);
    s->mass = fdup(mass, nV); s->DBC = fdup(DBC, nV); s->mu = fdup(mu, nT);
    s->moreDBC = fdup(NULL, nV); s->OffsetX = fdup(NULL, 3 * (size_t)nV);   /* dataLoader.cu:305,316 (zero-filled) */
    for (int i = 0; i < nV; i++) if (s->DBC[i] > 0) s->numDBC++;
    s->Tet = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)nT);
    memcpy(s->Tet, Tet, sizeof(uint32_t) * 4 * (size_t)nT);
    /* FEMSolver ctor femSolver.cu:6-17 */
    s->DmInv = fdup(NULL, 9 * (size_t)nT); s->V0 = fdup(NULL, nT);
    o_rest_shape(s->X, s->Tet, nT, s->DmInv, s->V0);
    s->massDt_2s = fdup(NULL, nV); s->matrix_diag = fdup(NULL, nV);
    s->sn = fdup(NULL, 3 * (size_t)nV); s->sn_old = fdup(NULL, 3 * (size_t)nV); s->b = fdup(NULL, 3 * (size_t)nV);
    s->next_x = fdup(NULL, 3 * (size_t)nV); s->prev_x = fdup(NULL, 3 * (size_t)nV);
    s->H = fdup(NULL, 12 * (size_t)nT);
    /* incidence, ascending tet order */
    s->inc_ptr = (int *)calloc((size_t)nV + 1, sizeof(int));
    s->inc = (int *)malloc(sizeof(int) * 4 * (size_t)nT);
    for (int t = 0; t < nT; t++) for (int k = 0; k < 4; k++) s->inc_ptr[Tet[4 * t + k] + 1]++;
    for (int v = 0; v < nV; v++) s->inc_ptr[v + 1] += s->inc_ptr[v];
    int *fill = (int *)malloc(sizeof(int) * (size_t)nV);
    memcpy(fill, s->inc_ptr, sizeof(int) * (size_t)nV);
    for (int t = 0; t < nT; t++) for (int k = 0; k < 4; k++) s->inc[fill[Tet[4 * t + k]]++] = 4 * t + k;
    free(fill);
    /* fixed bodies */
    if (fb) {
        size_t tot = 6 * (size_t)fb->n_planes + 4 * (size_t)fb->n_spheres + 7 * (size_t)fb->n_cyls;
        s->fbbuf = fdup(NULL, tot + 1);
        float *p = s->fbbuf;
        s->fb = *fb;
        memcpy(p, fb->plane_p0, 12 * (size_t)fb->n_planes); s->fb.plane_p0 = p; p += 3 * fb->n_planes;
        memcpy(p, fb->plane_up, 12 * (size_t)fb->n_planes); s->fb.plane_up = p; p += 3 * fb->n_planes;
        memcpy(p, fb->sphere_c, 12 * (size_t)fb->n_spheres); s->fb.sphere_c = p; p += 3 * fb->n_spheres;
        memcpy(p, fb->sphere_r, 4 * (size_t)fb->n_spheres); s->fb.sphere_r = p; p += fb->n_spheres;
        memcpy(p, fb->cyl_c, 12 * (size_t)fb->n_cyls); s->fb.cyl_c = p; p += 3 * fb->n_cyls;
        memcpy(p, fb->cyl_axis, 12 * (size_t)fb->n_cyls); s->fb.cyl_axis = p; p += 3 * fb->n_cyls;
        memcpy(p, fb->cyl_r, 4 * (size_t)fb->n_cyls); s->fb.cyl_r = p;
    }
    s->Xd = (double *)malloc(sizeof(double) * 3 * (size_t)nV);
    s->Vd = (double *)calloc(3 * (size_t)nV, sizeof(double));
    s->XTd = (double *)malloc(sizeof(double) * 3 * (size_t)nV);
    for (size_t i = 0; i < 3 * (size_t)nV; i++) s->Xd[i] = s->XTd[i] = X[i];
    return s;
}

void o_scene_destroy(o_scene *s)
{
    if (s) { free(s->Tri); free(s->TriFathers); free(s->tI); free(s->Normals); }
    if (!s) return;
    free(s->X); free(s->X0); free(s->XTilde); free(s->V); free(s->DBCX); free(s->ExtForce);
    free(s->moreDBC); free(s->OffsetX);
    free(s->mass); free(s->DBC); free(s->mu); free(s->Tet); free(s->DmInv); free(s->V0);
    free(s->massDt_2s); free(s->matrix_diag); free(s->sn); free(s->sn_old); free(s->b);
    free(s->next_x); free(s->prev_x); free(s->H); free(s->inc_ptr); free(s->inc);
    free(s->rowptr); free(s->col); free(s->val); free(s->chol);
    free(s->cg_r); free(s->cg_z); free(s->cg_p); free(s->cg_q); free(s->fbbuf);
    free(s->Xd); free(s->Vd); free(s->XTd);
    free(s);
}

void o_scene_reset(o_scene *s)
{   /* simulationContext.cu:233-243 + solver.h:39-43 */
    size_t n = 3 * (size_t)s->nV;
    memcpy(s->X, s->X0, n * sizeof(float));
    memcpy(s->XTilde, s->X0, n * sizeof(float));
    memset(s->V, 0, n * sizeof(float));
    memset(s->moreDBC, 0, (size_t)s->nV * sizeof(float));   /* simulationContext.cu:240 */
    for (size_t i = 0; i < n; i++) { s->Xd[i] = s->XTd[i] = s->X0[i]; s->Vd[i] = 0; }
    s->ready = 0;
}

void o_scene_get(const o_scene *s, float *X, float *V, float *XTilde)
{
    size_t n = 3 * (size_t)s->nV * sizeof(float);
    if (X) memcpy(X, s->X, n);
    if (V) memcpy(V, s->V, n);
    if (XTilde) memcpy(XTilde, s->XTilde, n);
}

void o_scene_set(o_scene *s, const float *X, const float *V, const float *XTilde)
{
    size_t n = 3 * (size_t)s->nV * sizeof(float);
    if (X) memcpy(s->X, X, n);
    if (V) memcpy(s->V, V, n);
    if (XTilde) memcpy(s->XTilde, XTilde, n);
}

/* Mouse-drag soft constraints.  The reference fills SolverData::moreDBC / OffsetX with Control_Kernel
 * (simulationContext.cu:202-218) and MouseSelection::target in RayIntersect (:196); PdSolver consumes them at
 * pdUtil.cu:56-69,80-87,159-164,187-188,201-206.  moreDBC == NULL clears the drag (ResetMoreDBC(true), :220-226). */
void o_scene_set_drag(o_scene *s, const float *moreDBC, const float *OffsetX, const float target[3])
{
    if (!moreDBC) { memset(s->moreDBC, 0, (size_t)s->nV * sizeof(float)); return; }
    memcpy(s->moreDBC, moreDBC, (size_t)s->nV * sizeof(float));
    if (OffsetX) memcpy(s->OffsetX, OffsetX, 3 * (size_t)s->nV * sizeof(float));
    if (target) memcpy(s->drag_target, target, 12);
}

/* Control_Kernel, simulationContext.cu:202-218 (RADIUS_SQUARED 0.002, :18; control_mag 10, :229), on the
 * current X; then MouseSelection::target = `target`.  The SASS of the kernel fuses dot(diff, diff) like every
 * other glm::dot here (FMUL y, FFMA x, FFMA z). */
void o_scene_drag_select(o_scene *s, int select_v, float control_mag, const float target[3])
{
    for (int i = 0; i < s->nV; i++) {
        float stiffness = 0.0f;
        if (s->DBC[i] == 0 && select_v != -1) {
            float d[3];
            for (int k = 0; k < 3; k++) { d[k] = s->X[3 * i + k] - s->X[3 * select_v + k]; s->OffsetX[3 * i + k] = d[k]; }
            float dist2 = DOT3_NV(d[0], d[0], d[1], d[1], d[2], d[2]);
            if ((double)dist2 < 0.002) stiffness = control_mag;     /* RADIUS_SQUARED is a double literal: the comparison promotes dist2 */
        }
        s->moreDBC[i] = stiffness;
    }
    if (target) memcpy(s->drag_target, target, 12);
}

/* SimulationCUDAContext::UpdateSoftBodyAttr -> DataLoader::FillData (simulationContext.cu:165-176, dataLoader.cu:381-384):
 * SolverData::mu is overwritten; computeLocal reads it live, matrix_diag / the assembled matrix wait for the next
 * SolverPrepare (after Reset). */
void o_scene_set_mu(o_scene *s, const float *mu) { memcpy(s->mu, mu, (size_t)s->nT * sizeof(float)); }

void o_scene_get_drag(const o_scene *s, float *moreDBC, float *OffsetX, float *DBCX)
{
    if (moreDBC) memcpy(moreDBC, s->moreDBC, (size_t)s->nV * sizeof(float));
    if (OffsetX) memcpy(OffsetX, s->OffsetX, 3 * (size_t)s->nV * sizeof(float));
    if (DBCX) memcpy(DBCX, s->DBCX, 3 * (size_t)s->nV * sizeof(float));
}

/* AiSi = transpose(DmInv) * G ; columns (pdUtil.cu:16).  col0 = -(r0+r1+r2), col k = row k-1 */
static void aisi_cols(const float *Bi, float cols[4][3])
{
    for (int r = 0; r < 3; r++) {
        float m0 = Bi[0 * 3 + r], m1 = Bi[1 * 3 + r], m2 = Bi[2 * 3 + r]; /* M[k][r] = DmInv row k, comp r */
        cols[0][r] = (-m0 - m1) - m2;   /* FADD(-a,-b); FADD(-c, .) in the computeSiTSi SASS */
        cols[1][r] = m0; cols[2][r] = m1; cols[3][r] = m2;
    }
}

/* PdSolver::SolverPrepare pdSolver.cu:40-139 (the parts every mode needs) */
static void prepare(o_scene *s, const o_params *p)
{
    int nV = s->nV, nT = s->nT;
    float dt2 = p->dt * p->dt;
    memset(s->matrix_diag, 0, sizeof(float) * (size_t)nV);
    for (int t = 0; t < nT; t++) {   /* computeSiTSi pdUtil.cu:9-24, sequential atomics */
        float cols[4][3];
        aisi_cols(s->DmInv + 9 * t, cols);
        float coef = s->V0[t] * s->mu[t];
        for (int i = 0; i < 4; i++) {
            float kii = DOT3_NV(cols[i][0], cols[i][0], cols[i][1], cols[i][1], cols[i][2], cols[i][2]);
            s->matrix_diag[s->Tet[4 * t + i]] += kii * coef;
        }
    }
    for (int v = 0; v < nV; v++)     /* setMDt_2 pdUtil.cu:42-54 ; positional_weight = 1e6 */
        s->massDt_2s[v] = (s->mass[v] + s->DBC[v] * 1e6f) / dt2;
    memcpy(s->DBCX, s->X0, sizeof(float) * 3 * (size_t)nV); /* pdSolver.cu:134 */
    free(s->rowptr); free(s->col); free(s->val); free(s->chol);
    s->rowptr = s->col = NULL; s->val = NULL; s->chol = NULL; s->nnz = 0;
    s->prepared_dt = p->dt;
    s->ready = 1;
}

/* scalar system matrix A^ (CSR, sorted columns, duplicates summed in tet order) */
typedef struct { int c; float v; int seq; } ent_t;
static int ent_cmp(const void *a, const void *b)
{
    const ent_t *x = (const ent_t *)a, *y = (const ent_t *)b;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

static void build_matrix(o_scene *s, const o_params *p)
{
    if (!s->ready) prepare(s, p);
    if (s->rowptr) return;
    int nV = s->nV, nT = s->nT;
    /* row r gets, per incident tet, 4 entries (r, v_j) = K[j][i]*coef ; + diagonal c_r */
    int *cnt = (int *)calloc((size_t)nV + 1, sizeof(int));
    for (int v = 0; v < nV; v++) cnt[v + 1] = 4 * (s->inc_ptr[v + 1] - s->inc_ptr[v]) + 1;
    for (int v = 0; v < nV; v++) cnt[v + 1] += cnt[v];
    ent_t *e = (ent_t *)malloc(sizeof(ent_t) * (size_t)cnt[nV]);
    int *rowptr = (int *)calloc((size_t)nV + 1, sizeof(int));
    int *col = (int *)malloc(sizeof(int) * (size_t)cnt[nV]);
    float *val = (float *)malloc(sizeof(float) * (size_t)cnt[nV]);
    int nnz = 0;
    for (int v = 0; v < nV; v++) {
        int base = cnt[v], k = 0;
        for (int q = s->inc_ptr[v]; q < s->inc_ptr[v + 1]; q++) {
            int t = s->inc[q] >> 2, i = s->inc[q] & 3;
            float cols[4][3];
            aisi_cols(s->DmInv + 9 * t, cols);
            float coef = s->V0[t] * s->mu[t];
            for (int j = 0; j < 4; j++) {
                float kji = cols[i][0] * cols[j][0] + cols[i][1] * cols[j][1] + cols[i][2] * cols[j][2];
                e[base + k].c = (int)s->Tet[4 * t + j]; e[base + k].v = kji * coef; e[base + k].seq = k; k++;
            }
        }
        /* setMDt_2 at prepare time (pdUtil.cu:42-54): the assembled matrix never sees moreDBC or a later dt */
        e[base + k].c = v; e[base + k].v = (s->mass[v] + s->DBC[v] * 1e6f) / (s->prepared_dt * s->prepared_dt); e[base + k].seq = k; k++;
        qsort(e + base, (size_t)k, sizeof(ent_t), ent_cmp);
        rowptr[v] = nnz;
        for (int a = 0; a < k;) {
            int c = e[base + a].c; float acc = 0;
            while (a < k && e[base + a].c == c) { acc += e[base + a].v; a++; }
            col[nnz] = c; val[nnz] = acc; nnz++;
        }
    }
    rowptr[nV] = nnz;
    free(e); free(cnt);
    (void)nT;
    s->rowptr = rowptr; s->col = col; s->val = val; s->nnz = nnz;
}

int o_scene_system_matrix(o_scene *s, const o_params *p, int *rowptr, int *col, float *val)
{
    build_matrix(s, p);
    if (rowptr) {
        memcpy(rowptr, s->rowptr, sizeof(int) * ((size_t)s->nV + 1));
        memcpy(col, s->col, sizeof(int) * (size_t)s->nnz);
        memcpy(val, s->val, sizeof(float) * (size_t)s->nnz);
    }
    return s->nnz;
}

void o_scene_get_setup(o_scene *s, const o_params *p, float *matrix_diag, float *massDt_2s,
                       float *DmInv, float *V0)
{
    if (!s->ready) prepare(s, p);
    if (matrix_diag) memcpy(matrix_diag, s->matrix_diag, sizeof(float) * (size_t)s->nV);
    if (massDt_2s) memcpy(massDt_2s, s->massDt_2s, sizeof(float) * (size_t)s->nV);
    if (DmInv) memcpy(DmInv, s->DmInv, sizeof(float) * 9 * (size_t)s->nT);
    if (V0) memcpy(V0, s->V0, sizeof(float) * (size_t)s->nT);
}

void o_scene_stats(const o_scene *s, int *pd_iters, int *inner_iters)
{
    if (pd_iters) *pd_iters = s->last_pd_iters;
    if (inner_iters) *inner_iters = s->last_inner_iters;
}

/* PdUtil::computeLocal pdUtil.cu:97-145, one tet -> H[4][3] (column per corner).
 * glm products are evaluated left to right: ((|V0|*w) * R) * transpose(DmInv) * G */
static void local_one(const float *q, const uint32_t *tv, const float *Bi, float V0, float wi,
                      int is_jacobi, float *H)
{
    const float *v0 = q + 3 * tv[0], *v1 = q + 3 * tv[1], *v2 = q + 3 * tv[2], *v3 = q + 3 * tv[3];
    float Ds[9], F[9], R[9];
    for (int r = 0; r < 3; r++) { Ds[r * 3 + 0] = v1[r] - v0[r]; Ds[r * 3 + 1] = v2[r] - v0[r]; Ds[r * 3 + 2] = v3[r] - v0[r]; }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            F[r * 3 + c] = DOT3_NV(Ds[r * 3 + 0], Bi[0 * 3 + c], Ds[r * 3 + 1], Bi[1 * 3 + c], Ds[r * 3 + 2], Bi[2 * 3 + c]);
        }
    o_rotation(F, R);
    if (is_jacobi) for (int k = 0; k < 9; k++) R[k] = R[k] - F[k];
    float sc = fabsf(V0) * wi;
    float M1[9], M2[9];
    for (int k = 0; k < 9; k++) M1[k] = sc * R[k];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {   /* M1 * transpose(DmInv): (r,c) = sum_k M1[r][k]*Bi[c][k] */
            M2[r * 3 + c] = DOT3_NV(M1[r * 3 + 0], Bi[c * 3 + 0], M1[r * 3 + 1], Bi[c * 3 + 1], M1[r * 3 + 2], Bi[c * 3 + 2]);
        }
    for (int r = 0; r < 3; r++) {
        H[0 * 3 + r] = (-M2[r * 3 + 0] - M2[r * 3 + 1]) - M2[r * 3 + 2];   /* FADD(-a,-b); FADD(-c, .) in the SASS */
        H[1 * 3 + r] = M2[r * 3 + 0];
        H[2 * 3 + r] = M2[r * 3 + 1];
        H[3 * 3 + r] = M2[r * 3 + 2];
    }
}

/* addM_h2Sn + computeLocal + computeDBCLocal (pdSolver.cu:168-171).  The gather over the
 * ascending incidence list reproduces a sequential tet-order scatter bit for bit. */
static void local_step(o_scene *s, const o_params *p, int is_jacobi)
{
    int nV = s->nV, nT = s->nT;
    int nth = p->threads > 0 ? p->threads : 1;
    (void)nth;
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int t = 0; t < nT; t++)
        local_one(s->sn, s->Tet + 4 * t, s->DmInv + 9 * t, s->V0[t], s->mu[t], is_jacobi, s->H + 12 * (size_t)t);
    float dtInv = 1.0f / p->dt;
    float wdbc = 1e6f * (dtInv * dtInv);   /* positional_weight * dt2Inv, pdSolver.cu:144-146,171 */
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int v = 0; v < nV; v++) {
        float c = s->massDt_2s[v];
        float b0 = c * s->sn_old[3 * v], b1 = c * s->sn_old[3 * v + 1], b2 = c * s->sn_old[3 * v + 2];
        for (int k = s->inc_ptr[v]; k < s->inc_ptr[v + 1]; k++) {
            const float *h = s->H + 3 * (size_t)s->inc[k];
            b0 += h[0]; b1 += h[1]; b2 += h[2];
        }
        if (s->numDBC > 0 && s->DBC[v] > 0) {   /* computeDBCLocal pdUtil.cu:147-166 (launched only when numDBC > 0, pdSolver.cu:170) */
            b0 = s->DBCX[3 * v] * wdbc; b1 = s->DBCX[3 * v + 1] * wdbc; b2 = s->DBCX[3 * v + 2] * wdbc;
        } else if (s->numDBC > 0 && s->moreDBC[v] > 0) {   /* pdUtil.cu:159-164: the weight is moreDBC itself, not moreDBC / h^2 */
            float mw = s->moreDBC[v];
            b0 = s->DBCX[3 * v] * mw; b1 = s->DBCX[3 * v + 1] * mw; b2 = s->DBCX[3 * v + 2] * mw;
        }
        s->b[3 * v] = b0; s->b[3 * v + 1] = b1; s->b[3 * v + 2] = b2;
    }
}

/* fixed bodies, fixedBodyData.cu:67-148, order spheres -> planes -> cylinders */
/* glm::dot / glm::length as nvcc contracts them (refFloor SASS: FMUL y, FFMA x, FFMA z; IEEE sqrt) */
static inline float dot3(const float *a, const float *b) { return DOT3_NV(a[0], b[0], a[1], b[1], a[2], b[2]); }
static inline float len3(const float *v) { return sqrtf(dot3(v, v)); }

static void respond(float *V, const float *n, float muT, float muN)
{
    float vn = dot3(V, n);
    float vN[3] = {vn * n[0], vn * n[1], vn * n[2]};
    float vT[3] = {V[0] - vN[0], V[1] - vN[1], V[2] - vN[2]};
    float mag = len3(vT);
    float a = mag == 0 ? 0 : fmaxf(1 - muT * (1 + muN) * len3(vN) / mag, 0.0f);
    for (int k = 0; k < 3; k++) V[k] = fmaf(vT[k], a, -(vN[k] * muN));   /* FMUL vN*muN; FFMA(vT, a, -.) */
}

static void fixed_bodies(o_scene *s, float muT, float muN)
{
    const o_fixed_bodies *fb = &s->fb;
    for (int i = 0; i < s->nV; i++) {
        float *x = s->XTilde + 3 * i, *v = s->V + 3 * i;
        for (int j = 0; j < fb->n_spheres; j++) {
            const float *c = fb->sphere_c + 3 * j; float r = fb->sphere_r[j];
            float tc[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
            float d = len3(tc);
            if (d < r) {
                float inv = 1.0f / sqrtf(dot3(tc, tc));   /* glm::normalize = v * inversesqrt(dot) */
                float n[3] = {tc[0] * inv, tc[1] * inv, tc[2] * inv};
                for (int k = 0; k < 3; k++) x[k] = fmaf(r - d, n[k], x[k]);
                respond(v, n, muT, muN);
            }
        }
    }
    for (int i = 0; i < s->nV; i++) {
        float *x = s->XTilde + 3 * i, *v = s->V + 3 * i;
        for (int j = 0; j < fb->n_planes; j++) {
            const float *p0 = fb->plane_p0 + 3 * j, *up = fb->plane_up + 3 * j;
            float rel[3] = {x[0] - p0[0], x[1] - p0[1], x[2] - p0[2]};
            float sd = dot3(rel, up);
            if (sd < 0 && dot3(v, up) < 0) {
                for (int k = 0; k < 3; k++) x[k] = fmaf(-up[k], sd, x[k]);   /* FFMA(-up, sd, x) */
                respond(v, up, muT, muN);
            }
        }
    }
    for (int i = 0; i < s->nV; i++) {
        float *x = s->XTilde + 3 * i, *v = s->V + 3 * i;
        for (int j = 0; j < fb->n_cyls; j++) {
            const float *c = fb->cyl_c + 3 * j, *ax = fb->cyl_axis + 3 * j; float r = fb->cyl_r[j];
            float rel[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
            /* n = (I - a a^T) rel, as a matrix-vector product (fixedBodyData.cu:117-119) */
            float nn[3];
            for (int k = 0; k < 3; k++) {
                float m0 = fmaf(-ax[0], ax[k], k == 0 ? 1.0f : 0.0f);   /* FFMA(-a, a, 1|0) in the SASS */
                float m1 = fmaf(-ax[1], ax[k], k == 1 ? 1.0f : 0.0f);
                float m2 = fmaf(-ax[2], ax[k], k == 2 ? 1.0f : 0.0f);
                nn[k] = DOT3_NV(m0, rel[0], m1, rel[1], m2, rel[2]);
            }
            float d = len3(nn);
            if (d < r) {
                float inv = 1.0f / sqrtf(dot3(nn, nn));
                float n[3] = {nn[0] * inv, nn[1] * inv, nn[2] * inv};
                for (int k = 0; k < 3; k++) x[k] = fmaf(r - d, n[k], x[k]);
                respond(v, n, muT, muN);
            }
        }
    }
}

/* y = A^ (x) I3 * x on interleaved xyz */
static void spmv3(const o_scene *s, const float *x, float *y)
{
    for (int v = 0; v < s->nV; v++) {
        float a0 = 0, a1 = 0, a2 = 0;
        for (int k = s->rowptr[v]; k < s->rowptr[v + 1]; k++) {
            float a = s->val[k]; const float *xc = x + 3 * s->col[k];
            a0 += a * xc[0]; a1 += a * xc[1]; a2 += a * xc[2];
        }
        y[3 * v] = a0; y[3 * v + 1] = a1; y[3 * v + 2] = a2;
    }
}

static float dotn(const float *a, const float *b, size_t n)
{
    double acc = 0;
    for (size_t i = 0; i < n; i++) acc += (double)a[i] * (double)b[i];
    return (float)acc;
}

/* PCGJacobiSolver<float>::Solve pcgJacobi.cu:88-172 with d_guess = x (warm start) */
static int pcg_solve(o_scene *s, const o_params *p, const float *b, float *x)
{
    size_t N = 3 * (size_t)s->nV;
    if (!s->cg_r) { s->cg_r = fdup(NULL, N); s->cg_z = fdup(NULL, N); s->cg_p = fdup(NULL, N); s->cg_q = fdup(NULL, N); }
    float *r = s->cg_r, *z = s->cg_z, *pp = s->cg_p, *q = s->cg_q;
    spmv3(s, x, q);
    for (size_t i = 0; i < N; i++) r[i] = b[i] - q[i];
    float rho = 0, rho_t;
    int k;
    for (k = 0; k < p->pcg_max_iter; k++) {
        float rn = sqrtf(dotn(r, r, N));
        if (rn < p->pcg_tol) break;
        for (int v = 0; v < s->nV; v++) {   /* ExtractInverseDiagonal + ApplyJacobiPreconditioner */
            float d = 1.0f;
            for (int e = s->rowptr[v]; e < s->rowptr[v + 1]; e++) if (s->col[e] == v) { d = s->val[e]; break; }
            if (fabsf(d) < 1e-9f) d = 1.0f;
            float inv = 1.0f / d;
            z[3 * v] = r[3 * v] * inv; z[3 * v + 1] = r[3 * v + 1] * inv; z[3 * v + 2] = r[3 * v + 2] * inv;
        }
        rho_t = rho;
        rho = dotn(r, z, N);
        if (fabsf(rho) < 1e-15f) break;
        if (k == 0) memcpy(pp, z, N * sizeof(float));
        else { float beta = rho / rho_t; for (size_t i = 0; i < N; i++) pp[i] = beta * pp[i] + z[i]; }
        spmv3(s, pp, q);
        float pTq = dotn(pp, q, N);
        float alpha = rho / pTq;
        for (size_t i = 0; i < N; i++) { x[i] += alpha * pp[i]; r[i] -= alpha * q[i]; }
    }
    return k;
}

/* "exact" direct solve: dense Cholesky of A^ in double (small meshes), the stand-in for
 * Eigen::SimplicialCholesky (pdSolver.cu:103,181-183) and cusolverSp Cholesky
 * (cholesky.cu:133-192), whose results the reference's tests do not pin. */
static int direct_solve(o_scene *s, const float *b, float *x)
{
    int n = s->nV;
    if (!s->chol) {
        if (n > 8192) return -1;
        double *L = (double *)calloc((size_t)n * n, sizeof(double));
        for (int v = 0; v < n; v++)
            for (int e = s->rowptr[v]; e < s->rowptr[v + 1]; e++) L[(size_t)v * n + s->col[e]] = s->val[e];
        for (int j = 0; j < n; j++) {
            double d = L[(size_t)j * n + j];
            for (int k = 0; k < j; k++) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
            if (d <= 0) { free(L); return -2; }
            d = sqrt(d);
            L[(size_t)j * n + j] = d;
            for (int i = j + 1; i < n; i++) {
                double a = L[(size_t)i * n + j];
                for (int k = 0; k < j; k++) a -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
                L[(size_t)i * n + j] = a / d;
            }
        }
        s->chol = L;
    }
    const double *L = s->chol;
    double *y = (double *)malloc(sizeof(double) * (size_t)n);
    for (int d3 = 0; d3 < 3; d3++) {
        for (int i = 0; i < n; i++) {
            double a = b[3 * i + d3];
            for (int k = 0; k < i; k++) a -= L[(size_t)i * n + k] * y[k];
            y[i] = a / L[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double a = y[i];
            for (int k = i + 1; k < n; k++) a -= L[(size_t)k * n + i] * y[k];
            y[i] = a / L[(size_t)i * n + i];
        }
        for (int i = 0; i < n; i++) x[3 * i + d3] = (float)y[i];
    }
    free(y);
    return 0;
}

/* PdSolver::SolverStep pdSolver.cu:141-208 */
static int solver_step(o_scene *s, const o_params *p)
{
    int nV = s->nV;
    size_t N = 3 * (size_t)nV;
    float dt = p->dt;
    float dtInv = 1.0f / dt;
    float dt2 = dt * dt;
    for (int v = 0; v < nV; v++) {   /* gravity_force pdSolver.cu:14-18,154 */
        s->ExtForce[3 * v] = 0.0f; s->ExtForce[3 * v + 1] = -p->gravity * s->mass[v]; s->ExtForce[3 * v + 2] = 0.0f;
    }
    for (int v = 0; v < nV; v++)     /* setMDt_2MoreDBC pdUtil.cu:56-69 */
        if (s->DBC[v] == 0) {
            float wi = s->moreDBC[v];
            s->massDt_2s[v] = (wi > 0) ? (s->mass[v] + wi) / dt2 : s->mass[v] / dt2;
        }
    for (int v = 0; v < nV; v++) {   /* computeSn pdUtil.cu:73-95 */
        if (s->moreDBC[v] > 0) {     /* dragged vertex: sn = DBCX = target + OffsetX (pdUtil.cu:80-87) */
            for (int k = 0; k < 3; k++) s->sn[3 * v + k] = s->DBCX[3 * v + k] = s->drag_target[k] + s->OffsetX[3 * v + k];
            continue;
        }
        float dt2_m_1 = 1.0f / s->massDt_2s[v];
        for (int k = 0; k < 3; k++)
            s->sn[3 * v + k] = fmaf(s->ExtForce[3 * v + k], dt2_m_1, fmaf(s->V[3 * v + k], dt, s->X[3 * v + k]));   /* computeSn SASS: 2 FFMA */
    }
    memcpy(s->sn_old, s->sn, N * sizeof(float));
    int jacobi = (p->global_solver == 0);
    if (jacobi) memcpy(s->prev_x, s->sn, N * sizeof(float));
    else memset(s->prev_x, 0, N * sizeof(float));
    if (!jacobi) build_matrix(s, p);
    float err = 1.0f, omega = 1.0f;
    int it = 0, inner = 0;
    for (int i = 0; i < p->num_iterations && sqrtf(err) >= p->tol; i++, it++) {
        local_step(s, p, jacobi);
        if (jacobi) {
            for (int v = 0; v < nV; v++) {   /* getErrorKern pdUtil.cu:195-214 */
                float c = s->massDt_2s[v], md = s->matrix_diag[v];
                if (s->moreDBC[v] > 0) {     /* pdUtil.cu:201-206: a dragged vertex keeps sn */
                    for (int k = 0; k < 3; k++) s->next_x[3 * v + k] = s->sn[3 * v + k];
                    continue;
                }
                for (int k = 0; k < 3; k++)
                    s->next_x[3 * v + k] = fmaf(-c, s->sn[3 * v + k], s->b[3 * v + k]) / (c + md) + s->sn[3 * v + k];   /* FFMA(-c,q,b); IEEE div; FADD */
            }
            if (i <= 10) omega = 1;                                   /* pdSolver.cu:196-198 */
            else if (i == 11) omega = 2 / (2 - p->rho * p->rho);
            else omega = 4 / (4 - p->rho * p->rho * omega);
            for (size_t k = 0; k < N; k++) {  /* chebyshevKern pdUtil.cu:216-226 (0.9 is a double) */
                float nx = (float)fma((double)(s->next_x[k] - s->sn[k]), 0.9, (double)s->sn[k]);   /* DFMA */
                nx = fmaf(nx - s->prev_x[k], omega, s->prev_x[k]);                                /* FFMA */
                s->next_x[k] = nx;
                s->prev_x[k] = s->sn[k];
                s->sn[k] = nx;
            }
        } else {
            if (p->global_solver == 1) {
                if (direct_solve(s, s->b, s->sn) != 0) return -1;
            } else {
                inner += pcg_solve(s, p, s->b, s->sn);
            }
            double acc = 0;                    /* computeError pdSolver.cu:243-253 */
            for (size_t k = 0; k < N; k++) { double d = (double)s->prev_x[k] - (double)s->sn[k]; acc += d * d; }
            err = (float)(acc / (double)N);
            memcpy(s->prev_x, s->sn, N * sizeof(float));
        }
    }
    s->last_pd_iters = it; s->last_inner_iters = inner;
    for (int v = 0; v < nV; v++)   /* updateVelPos pdUtil.cu:180-193 */
        for (int k = 0; k < 3; k++) {
            float np = s->sn[3 * v + k];
            s->V[3 * v + k] = (s->moreDBC[v] > 0) ? 0.0f : (np - s->XTilde[3 * v + k]) * dtInv;   /* pdUtil.cu:187-190 */
            s->XTilde[3 * v + k] = np;
        }
    return 0;
}

static void mesh_collision(o_scene *s);

int o_scene_step(o_scene *s, const o_params *p, int n_steps)
{
    for (int n = 0; n < n_steps; n++) {   /* PdSolver::Update pdSolver.cu:210-232 */
        if (!s->ready) prepare(s, p);
        int rc = solver_step(s, p);
        if (rc) return rc;
        if (s->col_on) mesh_collision(s);                                  /* handleCollision: DetectCollision + CCDKernel, :218-225 */
        else memcpy(s->X, s->XTilde, sizeof(float) * 3 * (size_t)s->nV);   /* handleCollision == false, :227 */
        fixed_bodies(s, p->muT, p->muN);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * double-precision twin (Jacobi mode): same algorithm, exact polar rotation with the
 * reference's inversion convention (U,V proper, sigma_3 carries the sign).
 * ---------------------------------------------------------------------------------------- */
static void jacobi_eig3_d(double S[3][3], double Q[3][3])
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (fabs(S[p][q]) < 1e-300) continue;
                double th = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; k++) {
                    double a = S[k][p], b = S[k][q];
                    S[k][p] = c * a - sn * b; S[k][q] = sn * a + c * b;
                }
                for (int k = 0; k < 3; k++) {
                    double a = S[p][k], b = S[q][k];
                    S[p][k] = c * a - sn * b; S[q][k] = sn * a + c * b;
                }
                for (int k = 0; k < 3; k++) {
                    double a = Q[k][p], b = Q[k][q];
                    Q[k][p] = c * a - sn * b; Q[k][q] = sn * a + c * b;
                }
            }
    }
}

static void rotation_d(const double F[3][3], double R[3][3])
{
    double S[3][3], Q[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        S[i][j] = 0; for (int k = 0; k < 3; k++) S[i][j] += F[k][i] * F[k][j];
    }
    jacobi_eig3_d(S, Q);
    double lam[3] = {S[0][0], S[1][1], S[2][2]};
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; a++) for (int b = a + 1; b < 3; b++) if (lam[ord[b]] > lam[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Vm[3][3];
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) Vm[i][j] = Q[i][ord[j]];
    double detV = Vm[0][0] * (Vm[1][1] * Vm[2][2] - Vm[1][2] * Vm[2][1]) - Vm[0][1] * (Vm[1][0] * Vm[2][2] - Vm[1][2] * Vm[2][0]) +
                  Vm[0][2] * (Vm[1][0] * Vm[2][1] - Vm[1][1] * Vm[2][0]);
    if (detV < 0) for (int i = 0; i < 3; i++) Vm[i][2] = -Vm[i][2];
    /* U columns: u_j = F v_j / sigma_j for j=0,1 ; u_2 = u_0 x u_1 (proper) */
    double Um[3][3];
    for (int j = 0; j < 2; j++) {
        double u[3] = {0, 0, 0};
        for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) u[i] += F[i][k] * Vm[k][j];
        if (j == 1) { double d = u[0] * Um[0][0] + u[1] * Um[1][0] + u[2] * Um[2][0]; for (int i = 0; i < 3; i++) u[i] -= d * Um[i][0]; }
        double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (n < 1e-150) { u[0] = (j == 0); u[1] = (j == 1); u[2] = 0; n = 1; }
        for (int i = 0; i < 3; i++) Um[i][j] = u[i] / n;
    }
    Um[0][2] = Um[1][0] * Um[2][1] - Um[2][0] * Um[1][1];
    Um[1][2] = Um[2][0] * Um[0][1] - Um[0][0] * Um[2][1];
    Um[2][2] = Um[0][0] * Um[1][1] - Um[1][0] * Um[0][1];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        R[i][j] = 0; for (int k = 0; k < 3; k++) R[i][j] += Um[i][k] * Vm[j][k];
    }
}

int o_scene_step_f64(o_scene *s, const o_params *p, int n_steps)
{
    int nV = s->nV, nT = s->nT;
    size_t N = 3 * (size_t)nV;
    double dt = (double)p->dt, dt2 = dt * dt;
    double *c = (double *)malloc(sizeof(double) * (size_t)nV), *md = (double *)calloc((size_t)nV, sizeof(double));
    double *Bd = (double *)malloc(sizeof(double) * 9 * (size_t)nT), *w = (double *)malloc(sizeof(double) * (size_t)nT);
    double *sn = (double *)malloc(sizeof(double) * N), *so = (double *)malloc(sizeof(double) * N);
    double *prev = (double *)malloc(sizeof(double) * N), *b = (double *)malloc(sizeof(double) * N);
    for (int t = 0; t < nT; t++) {
        const uint32_t *tv = s->Tet + 4 * t;
        double m[3][3];
        for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) m[r][k] = (double)s->X0[3 * tv[k + 1] + r] - (double)s->X0[3 * tv[0] + r];
        double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                     m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
        double *Bi = Bd + 9 * (size_t)t;
        Bi[0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det; Bi[1] = -(m[0][1] * m[2][2] - m[0][2] * m[2][1]) / det; Bi[2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det;
        Bi[3] = -(m[1][0] * m[2][2] - m[1][2] * m[2][0]) / det; Bi[4] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det; Bi[5] = -(m[0][0] * m[1][2] - m[0][2] * m[1][0]) / det;
        Bi[6] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det; Bi[7] = -(m[0][0] * m[2][1] - m[0][1] * m[2][0]) / det; Bi[8] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
        w[t] = fabs(det) / 6.0 * (double)s->mu[t];
        for (int i = 0; i < 4; i++) {
            double col[3];
            for (int r = 0; r < 3; r++) col[r] = (i == 0) ? -(Bi[0 + r] + Bi[3 + r] + Bi[6 + r]) : Bi[3 * (i - 1) + r];
            md[tv[i]] += w[t] * (col[0] * col[0] + col[1] * col[1] + col[2] * col[2]);
        }
    }
    for (int v = 0; v < nV; v++) c[v] = ((double)s->mass[v] + (double)s->DBC[v] * 1e6) / dt2;
    for (int step = 0; step < n_steps; step++) {
        for (int v = 0; v < nV; v++)
            for (int k = 0; k < 3; k++) {
                double f = (k == 1) ? -(double)p->gravity * (double)s->mass[v] : 0.0;
                sn[3 * v + k] = s->Xd[3 * v + k] + dt * s->Vd[3 * v + k] + f / c[v];
            }
        memcpy(so, sn, sizeof(double) * N); memcpy(prev, sn, sizeof(double) * N);
        double omega = 1, rho = (double)p->rho;
        for (int it = 0; it < p->num_iterations; it++) {
            for (int v = 0; v < nV; v++) for (int k = 0; k < 3; k++) b[3 * v + k] = c[v] * so[3 * v + k];
            for (int t = 0; t < nT; t++) {
                const uint32_t *tv = s->Tet + 4 * t; const double *Bi = Bd + 9 * (size_t)t;
                double Ds[3][3], F[3][3], R[3][3], M[3][3];
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) Ds[r][k] = sn[3 * tv[k + 1] + r] - sn[3 * tv[0] + r];
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) { F[r][k] = 0; for (int j = 0; j < 3; j++) F[r][k] += Ds[r][j] * Bi[3 * j + k]; }
                rotation_d(F, R);
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) R[r][k] -= F[r][k];
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) { M[r][k] = 0; for (int j = 0; j < 3; j++) M[r][k] += R[r][j] * Bi[3 * k + j]; M[r][k] *= w[t]; }
                for (int r = 0; r < 3; r++) {
                    b[3 * tv[0] + r] -= M[r][0] + M[r][1] + M[r][2];
                    b[3 * tv[1] + r] += M[r][0]; b[3 * tv[2] + r] += M[r][1]; b[3 * tv[3] + r] += M[r][2];
                }
            }
            if (s->numDBC > 0)
                for (int v = 0; v < nV; v++) if (s->DBC[v] > 0) for (int k = 0; k < 3; k++) b[3 * v + k] = (double)s->X0[3 * v + k] * (1e6 / dt2);
            if (it <= 10) omega = 1; else if (it == 11) omega = 2 / (2 - rho * rho); else omega = 4 / (4 - rho * rho * omega);
            for (int v = 0; v < nV; v++) for (int k = 0; k < 3; k++) {
                size_t i = 3 * (size_t)v + k;
                double nx = (b[i] - c[v] * sn[i]) / (c[v] + md[v]) + sn[i];
                nx = 0.9 * (nx - sn[i]) + sn[i];
                nx = (nx - prev[i]) * omega + prev[i];
                prev[i] = sn[i]; sn[i] = nx;
            }
        }
        for (size_t i = 0; i < N; i++) { s->Vd[i] = (sn[i] - s->XTd[i]) / dt; s->XTd[i] = sn[i]; s->Xd[i] = sn[i]; }
        /* fixed bodies in double (planes/spheres/cylinders), same rules */
        const o_fixed_bodies *fb = &s->fb;
        for (int pass = 0; pass < 3; pass++)
            for (int i = 0; i < nV; i++) {
                double *x = s->XTd + 3 * i, *v = s->Vd + 3 * i;
                int nb = pass == 0 ? fb->n_spheres : pass == 1 ? fb->n_planes : fb->n_cyls;
                for (int j = 0; j < nb; j++) {
                    double n[3], push = 0; int hit = 0;
                    if (pass == 1) {
                        const float *p0 = fb->plane_p0 + 3 * j, *up = fb->plane_up + 3 * j;
                        double sd = (x[0] - p0[0]) * up[0] + (x[1] - p0[1]) * up[1] + (x[2] - p0[2]) * up[2];
                        double vn = v[0] * up[0] + v[1] * up[1] + v[2] * up[2];
                        if (sd < 0 && vn < 0) { hit = 1; push = -sd; n[0] = up[0]; n[1] = up[1]; n[2] = up[2]; }
                    } else {
                        const float *cc = pass == 0 ? fb->sphere_c + 3 * j : fb->cyl_c + 3 * j;
                        double r = pass == 0 ? fb->sphere_r[j] : fb->cyl_r[j];
                        double rel[3] = {x[0] - cc[0], x[1] - cc[1], x[2] - cc[2]};
                        if (pass == 2) {
                            const float *ax = fb->cyl_axis + 3 * j;
                            double d = rel[0] * ax[0] + rel[1] * ax[1] + rel[2] * ax[2];
                            for (int k = 0; k < 3; k++) rel[k] -= d * ax[k];
                        }
                        double d = sqrt(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
                        if (d < r) { hit = 1; push = r - d; for (int k = 0; k < 3; k++) n[k] = rel[k] / d; }
                    }
                    if (hit) {
                        for (int k = 0; k < 3; k++) x[k] += push * n[k];
                        double vn = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
                        double vN[3] = {vn * n[0], vn * n[1], vn * n[2]}, vT[3] = {v[0] - vN[0], v[1] - vN[1], v[2] - vN[2]};
                        double mag = sqrt(vT[0] * vT[0] + vT[1] * vT[1] + vT[2] * vT[2]);
                        double a = mag == 0 ? 0 : fmax(1 - (double)p->muT * (1 + (double)p->muN) * fabs(vn) / mag, 0.0);
                        for (int k = 0; k < 3; k++) v[k] = -(double)p->muN * vN[k] + a * vT[k];
                    }
                }
            }
    }
    for (size_t i = 0; i < N; i++) { s->X[i] = (float)s->Xd[i]; s->V[i] = (float)s->Vd[i]; s->XTilde[i] = (float)s->XTd[i]; }
    free(c); free(md); free(Bd); free(w); free(sn); free(so); free(prev); free(b);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * mesh-mesh collision of the PD path: CollisionDetection::DetectCollision + CCDKernel
 * (pdSolver.cu:218-225).  Restated step by step, WITHOUT the reference's LBVH: the set of
 * triangle pairs whose swept boxes overlap does not depend on the tree, so all pairs are tried
 * (the oracle is for small cases).  Where the reference is racy / unstable the oracle takes
 * "threads run in index order, sorts are stable over the previous order".
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } cv3;
static cv3 c_mk(float x, float y, float z) { cv3 r = {x, y, z}; return r; }
static cv3 c_ld(const float *a, uint32_t i) { return c_mk(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
static cv3 c_sub(cv3 a, cv3 b) { return c_mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static cv3 c_add(cv3 a, cv3 b) { return c_mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static cv3 c_mul(cv3 a, float s) { return c_mk(a.x * s, a.y * s, a.z * s); }
static cv3 c_smul(float s, cv3 a) { return c_mk(s * a.x, s * a.y, s * a.z); }
static cv3 c_neg(cv3 a) { return c_mk(-a.x, -a.y, -a.z); }
static float c_dot(cv3 a, cv3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                 /* glm func_geometric.inl:65-72 */
static cv3 c_cross(cv3 x, cv3 y) { return c_mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
static float c_len(cv3 a) { return sqrtf(c_dot(a, a)); }
static cv3 c_norm(cv3 a) { return c_mul(a, 1.0f / sqrtf(c_dot(a, a))); }                        /* x * inversesqrt(dot(x, x)) */
static float c_stp(cv3 u, cv3 v, cv3 w) { return c_dot(u, c_cross(v, w)); }                     /* intersections.cu:27 */

static float c_newton(float a, float b, float c, float d, float x0, int init_dir)
{   /* intersections.cu:95-114 */
    if (init_dir != 0) {
        float y0 = d + x0 * (c + x0 * (b + x0 * a)), ddy0 = 2 * b + x0 * (6 * a);
        if (ddy0 != 0) x0 += init_dir * sqrtf(fabsf(2 * y0 / ddy0));
    }
    for (int iter = 0; iter < 100; iter++) {
        float y = d + x0 * (c + x0 * (b + x0 * a));
        float dy = c + x0 * (2 * b + x0 * 3 * a);
        if (dy == 0) return x0;
        float x1 = x0 - y / dy;
        if ((double)fabsf(x0 - x1) < 1e-6) return x0;
        x0 = x1;
    }
    return x0;
}
static int c_quadratic(float a, float b, float c, float *x)
{   /* intersections.cu:74-92 */
    float d = b * b - 4 * a * c;
    if (d < 0) { x[0] = -b / (2 * a); return 0; }
    float sgn = (float)(0.0f < b) - (float)(b < 0.0f);
    float q = -(b + sgn * sqrtf(d)) / 2;
    int i = 0;
    if ((double)fabsf(a) > 1e-12 * (double)fabsf(q)) x[i++] = q / a;
    if ((double)fabsf(q) > 1e-12 * (double)fabsf(c)) x[i++] = c / q;
    if (i == 2 && x[0] > x[1]) { float t = x[0]; x[0] = x[1]; x[1] = t; }
    return i;
}
static int c_cubic(float a, float b, float c, float d, float *x)
{   /* intersections.cu:46-72 */
    float xc[2];
    int ncrit = c_quadratic(3 * a, 2 * b, c, xc);
    if (ncrit == 0) { x[0] = c_newton(a, b, c, d, xc[0], 0); return 1; }
    if (ncrit == 1) return c_quadratic(b, c, d, x);
    float yc[2] = {d + xc[0] * (c + xc[0] * (b + xc[0] * a)), d + xc[1] * (c + xc[1] * (b + xc[1] * a))};
    int i = 0;
    if (yc[0] * a >= 0) x[i++] = c_newton(a, b, c, d, xc[0], -1);
    if (yc[0] * yc[1] <= 0) {
        int closer = fabsf(yc[0]) < fabsf(yc[1]) ? 0 : 1;
        x[i++] = c_newton(a, b, c, d, xc[closer], closer == 0 ? 1 : -1);
    }
    if (yc[1] * a <= 0) x[i++] = c_newton(a, b, c, d, xc[1], 1);
    return i;
}
static float c_vf_distance(cv3 x, cv3 y0, cv3 y1, cv3 y2, cv3 *n, float w[4])
{   /* intersections.cu:115-134 */
    *n = c_cross(c_norm(c_sub(y1, y0)), c_norm(c_sub(y2, y0)));
    if ((double)c_dot(*n, *n) < 1e-6) return FLT_MAX;
    *n = c_norm(*n);
    float h = c_dot(c_sub(x, y0), *n);
    float b0 = c_stp(c_sub(y1, x), c_sub(y2, x), *n), b1 = c_stp(c_sub(y2, x), c_sub(y0, x), *n), b2 = c_stp(c_sub(y0, x), c_sub(y1, x), *n);
    w[0] = 1; w[1] = -b0 / (b0 + b1 + b2); w[2] = -b1 / (b0 + b1 + b2); w[3] = -b2 / (b0 + b1 + b2);
    return h;
}
static float c_ee_distance(cv3 x0, cv3 x1, cv3 y0, cv3 y1, cv3 *n, float w[4])
{   /* intersections.cu:157-174 */
    *n = c_cross(c_norm(c_sub(x1, x0)), c_norm(c_sub(y1, y0)));
    if ((double)c_dot(*n, *n) < 1e-6) return FLT_MAX;
    *n = c_norm(*n);
    float h = c_dot(c_sub(x0, y0), *n);
    float a0 = c_stp(c_sub(y1, x1), c_sub(y0, x1), *n), a1 = c_stp(c_sub(y0, x0), c_sub(y1, x0), *n);
    float b0 = c_stp(c_sub(x0, y1), c_sub(x1, y1), *n), b1 = c_stp(c_sub(x1, y0), c_sub(x0, y0), *n);
    w[0] = a0 / (a0 + a1); w[1] = a1 / (a0 + a1); w[2] = -b0 / (b0 + b1); w[3] = -b1 / (b0 + b1);
    return h;
}
/* ccdCollisionTest<float>, intersections.cu:312-355; ee: edge-edge query */
float o_ccd_test(int ee, const uint32_t q[4], const float *X, const float *XT, float nout[3])
{
    cv3 x0 = c_ld(X, q[0]), x1 = c_ld(X, q[1]), x2 = c_ld(X, q[2]), x3 = c_ld(X, q[3]);
    cv3 v0 = c_sub(c_ld(XT, q[0]), x0), v1 = c_sub(c_ld(XT, q[1]), x1), v2 = c_sub(c_ld(XT, q[2]), x2), v3 = c_sub(c_ld(XT, q[3]), x3);
    cv3 x01 = c_sub(x1, x0), x02 = c_sub(x2, x0), x03 = c_sub(x3, x0), v01 = c_sub(v1, v0), v02 = c_sub(v2, v0), v03 = c_sub(v3, v0);
    float a0 = c_stp(x01, x02, x03);
    float a1 = c_stp(v01, x02, x03) + c_stp(x01, v02, x03) + c_stp(x01, x02, v03);
    float a2 = c_stp(x01, v02, v03) + c_stp(v01, x02, v03) + c_stp(v01, v02, x03);
    float a3 = c_stp(v01, v02, v03);
    cv3 n = c_mk(0, 0, 0);
    float res = 1.0f;
    if (!((double)fabsf(a0) < 1e-12 * (double)c_len(x01) * (double)c_len(x02) * (double)c_len(x03))) {
        float t[3];
        int nsol = c_cubic(a3, a2, a1, a0, t);
        for (int i = 0; i < nsol; i++) {
            if ((double)t[i] < -1e-12 || t[i] > 1) continue;
            cv3 xt0 = c_add(x0, c_smul(t[i], v0)), xt1 = c_add(x1, c_smul(t[i], v1)), xt2 = c_add(x2, c_smul(t[i], v2)), xt3 = c_add(x3, c_smul(t[i], v3));
            float w[4] = {0, 0, 0, 0}, d;
            int inside;
            if (!ee) {
                d = c_vf_distance(xt0, xt1, xt2, xt3, &n, w);
                inside = (double)fminf(-w[1], fminf(-w[2], -w[3])) >= -1e-3;
            } else {
                d = c_ee_distance(xt0, xt1, xt2, xt3, &n, w);
                inside = (double)fminf(w[0], fminf(w[1], fminf(-w[2], -w[3]))) >= -1e-3;
            }
            if (c_dot(n, c_add(c_add(c_smul(w[1], v1), c_smul(w[2], v2)), c_smul(w[3], v3))) > 0) n = c_neg(n);
            if ((double)fabsf(d) < 1e-6 && inside) { res = t[i]; break; }
        }
    }
    nout[0] = n.x; nout[1] = n.y; nout[2] = n.z;
    return res;
}

typedef struct { int type; uint32_t v[4]; float toi; float n[3]; } o_query;   /* type 1 = VF, 2 = EE (QueryType, aabb.h:54-58) */
static void swap_u(uint32_t *a, uint32_t *b) { uint32_t t = *a; *a = *b; *b = t; }
static int q_cmp_ids(const void *pa, const void *pb)
{   /* QueryComparator, broadphase.cu:213-223 */
    const o_query *a = (const o_query *)pa, *b = (const o_query *)pb;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;
    for (int k = 0; k < 4; k++) if (a->v[k] != b->v[k]) return a->v[k] < b->v[k] ? -1 : 1;
    return 0;
}
static int q_cmp_toi(const void *pa, const void *pb)
{   /* CompareQuery, narrowphase.cu:9-41; ties fall back to the id order (the reference's sort is unstable there) */
    const o_query *a = (const o_query *)pa, *b = (const o_query *)pb;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;
    if (a->v[0] != b->v[0]) return a->v[0] < b->v[0] ? -1 : 1;
    if (a->type == 2 && a->v[1] != b->v[1]) return a->v[1] < b->v[1] ? -1 : 1;
    if (a->toi != b->toi) return a->toi < b->toi ? -1 : 1;
    return q_cmp_ids(pa, pb);
}
static int q_same_group(const o_query *a, const o_query *b)
{   /* EqualQuery, narrowphase.cu:43-61 */
    if (a->type != b->type) return 0;
    return a->type == 1 ? a->v[0] == b->v[0] : (a->v[0] == b->v[0] && a->v[1] == b->v[1]);
}

static void mesh_collision(o_scene *s)
{
    const int nV = s->nV, nT = s->nTris;
    const uint32_t *tri = s->Tri, *fa = s->TriFathers;
    for (int v = 0; v < nV; v++) s->tI[v] = 1.0f;                        /* thrust::fill(tI, 1.0f), bvh.cu:173-174 */
    /* swept boxes: computeTriTrajBBoxCCD, ccd.cu:53-66 */
    float *bmin = (float *)malloc(sizeof(float) * 3 * (size_t)(nT > 0 ? nT : 1)), *bmax = (float *)malloc(sizeof(float) * 3 * (size_t)(nT > 0 ? nT : 1));
    for (int t = 0; t < nT; t++)
        for (int c = 0; c < 3; c++) {
            float mn = FLT_MAX, mx = -FLT_MAX;
            const float *src[2] = {s->X, s->XTilde};
            float p[6];
            for (int h = 0; h < 2; h++) for (int k = 0; k < 3; k++) p[3 * h + k] = src[h][3 * (size_t)tri[3 * t + k] + c];
            mn = fminf(fminf(fminf(fminf(fminf(p[0], p[1]), p[2]), p[3]), p[4]), p[5]);
            mx = fmaxf(fmaxf(fmaxf(fmaxf(fmaxf(p[0], p[1]), p[2]), p[3]), p[4]), p[5]);
            bmin[3 * t + c] = mn - 0.01f; bmax[3 * t + c] = mx + 0.01f;
        }
    /* traverseTree + fillQuery (broadphase.cu:267-400): every ORDERED pair of overlapping, non-adjacent triangles of different fathers */
    size_t cap = 1024, nq = 0;
    o_query *Q = (o_query *)malloc(cap * sizeof(o_query));
    long long pairs = 0;
    static const int ET[6] = {0, 1, 0, 2, 1, 2};                          /* edgeIndicesTable */
    for (int i = 0; i < nT; i++)
        for (int j = 0; j < nT; j++) {
            if (i == j || fa[i] == fa[j]) continue;                       /* PdSolver passes ignoreSelfCollision = true */
            int adj = 0;
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) adj |= tri[3 * i + a] == tri[3 * j + b];
            if (adj) continue;
            int hit = 1;
            for (int c = 0; c < 3; c++) if (bmax[3 * i + c] < bmin[3 * j + c] || bmin[3 * i + c] > bmax[3 * j + c]) hit = 0;
            if (!hit) continue;
            pairs++;
            if (nq + 12 > cap) { cap *= 2; Q = (o_query *)realloc(Q, cap * sizeof(o_query)); }
            for (int k = 0; k < 3; k++) {
                o_query q = {1, {tri[3 * i + k], tri[3 * j], tri[3 * j + 1], tri[3 * j + 2]}, 0.f, {0, 0, 0}};
                Q[nq++] = q;
            }
            for (int e = 0; e < 3; e++) {
                uint32_t v0 = tri[3 * i + ET[2 * e]], v1 = tri[3 * i + ET[2 * e + 1]];
                const uint32_t w[3][2] = {{tri[3 * j], tri[3 * j + 1]}, {tri[3 * j], tri[3 * j + 2]}, {tri[3 * j + 1], tri[3 * j + 2]}};
                for (int f = 0; f < 3; f++) { o_query q = {2, {v0, v1, w[f][0], w[f][1]}, 0.f, {0, 0, 0}}; Q[nq++] = q; }
            }
        }
    s->col_pairs = pairs;
    /* sortEachQuery (broadphase.cu:453-478), removeDuplicates (:488-498) */
    size_t m = 0;
    for (size_t k = 0; k < nq; k++) {
        o_query q = Q[k];
        if (q.type == 1) {
            if (q.v[0] == q.v[1] || q.v[0] == q.v[2] || q.v[0] == q.v[3]) continue;
            if (q.v[1] > q.v[2]) swap_u(&q.v[1], &q.v[2]);
            if (q.v[2] > q.v[3]) swap_u(&q.v[2], &q.v[3]);
            if (q.v[1] > q.v[2]) swap_u(&q.v[1], &q.v[2]);
        } else {
            if (q.v[0] > q.v[1]) swap_u(&q.v[0], &q.v[1]);
            if (q.v[2] > q.v[3]) swap_u(&q.v[2], &q.v[3]);
            if (q.v[0] == q.v[2] && q.v[1] == q.v[3]) continue;
        }
        Q[m++] = q;
    }
    nq = m;
    qsort(Q, nq, sizeof(o_query), q_cmp_ids);
    m = 0;
    for (size_t k = 0; k < nq; k++) if (m == 0 || q_cmp_ids(&Q[m - 1], &Q[k]) != 0) Q[m++] = Q[k];
    nq = m;
    if (nq > 0) {                                                         /* BroadPhaseCCD returned true: NarrowPhase, narrowphase.cu:122-134 */
        for (size_t k = 0; k < nq; k++) Q[k].toi = o_ccd_test(Q[k].type == 2, Q[k].v, s->X, s->XTilde, Q[k].n);
        qsort(Q, nq, sizeof(o_query), q_cmp_toi);
        m = 0;
        for (size_t k = 0; k < nq; k++) if (m == 0 || !q_same_group(&Q[m - 1], &Q[k])) Q[m++] = Q[k];
        nq = m;
        for (size_t k = 0; k < nq; k++) {                                 /* storeTi, narrowphase.cu:74-119, threads in index order */
            const o_query *q = &Q[k];
            if (!(q->toi < 1.0f)) continue;
            int nw = q->type == 2 ? 2 : 4;
            for (int u = 0; u < nw; u++) {
                float sgn = (q->type == 1 && u > 0) ? -1.0f : 1.0f;
                s->tI[q->v[u]] = 0.5f;
                for (int c = 0; c < 3; c++) s->Normals[3 * (size_t)q->v[u] + c] = sgn * q->n[c];
            }
        }
    }
    free(Q); free(bmin); free(bmax);
    /* CCDKernel, collisionUtil.cu:49-70 */
    for (int v = 0; v < nV; v++) {
        if (s->tI[v] < 1.0f) {
            cv3 n = c_ld(s->Normals, (uint32_t)v), vel = c_sub(c_ld(s->XTilde, (uint32_t)v), c_ld(s->X, (uint32_t)v));
            cv3 vn = c_smul(c_dot(vel, n), n);
            s->V[3 * v] = -vn.x; s->V[3 * v + 1] = -vn.y; s->V[3 * v + 2] = -vn.z;
        } else {
            for (int c = 0; c < 3; c++) s->X[3 * v + c] = s->XTilde[3 * v + c];
        }
    }
}

void o_scene_set_collision(o_scene *s, int enable, int nTris, const uint32_t *Tri, const uint32_t *TriFathers)
{
    s->col_on = enable;
    if (Tri) {
        free(s->Tri); free(s->TriFathers);
        s->nTris = nTris;
        s->Tri = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)(nTris > 0 ? nTris : 1));
        s->TriFathers = (uint32_t *)calloc((size_t)(nTris > 0 ? nTris : 1), sizeof(uint32_t));
        memcpy(s->Tri, Tri, sizeof(uint32_t) * 3 * (size_t)nTris);
        if (TriFathers) memcpy(s->TriFathers, TriFathers, sizeof(uint32_t) * (size_t)nTris);
    }
    if (!s->tI) {
        s->tI = (float *)malloc(sizeof(float) * (size_t)s->nV);
        s->Normals = (float *)calloc(3 * (size_t)s->nV, sizeof(float));
        for (int v = 0; v < s->nV; v++) s->tI[v] = 1.0f;
    }
}
void o_scene_get_collision(const o_scene *s, float *tI, float *normals, long long *pairs)
{
    if (tI) for (int v = 0; v < s->nV; v++) tI[v] = s->tI ? s->tI[v] : 1.0f;
    if (normals) { if (s->Normals) memcpy(normals, s->Normals, sizeof(float) * 3 * (size_t)s->nV); else memset(normals, 0, sizeof(float) * 3 * (size_t)s->nV); }
    if (pairs) *pairs = s->col_pairs;
}
