// Compiles the reference's own 3x3 SVD header (external/svd3_cuda/svd3_cuda.h, included
// from /root/reference where it lies -- nothing is copied) for the HOST, by mapping the
// CUDA intrinsics it uses onto their IEEE host equivalents.  Used only to pin
// oracle/pd_oracle.c:o_svd3 bit-for-bit (tests/test_oracle_svd.py).  TEST INFRASTRUCTURE.
#include <cmath>
#include <algorithm>
#define __device__
#define __forceinline__ inline
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __frsqrt_rn(float a) { return (float)(1.0 / std::sqrt((double)a)); }
using std::max;
#include <svd3_cuda.h>

extern "C" void ref_svd3(const float* A, float* U, float* S, float* V)
{
    svd<float>(A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], A[8],
               U[0], U[1], U[2], U[3], U[4], U[5], U[6], U[7], U[8],
               S[0], S[1], S[2],
               V[0], V[1], V[2], V[3], V[4], V[5], V[6], V[7], V[8]);
}
