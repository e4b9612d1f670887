// Device-side scalar math with pinned rounding.
//
// rl_sincosf restates, in double precision exactly as published, the single-precision sin/cos
// algorithm of glibc >= 2.28 (sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h}; ARM
// optimized-routines), which is the libm the reference calls at RangeLib.h:713-714 (BL) and
// :931-932 (RM).  The same restatement exists on the CPU as oracle/rangelib_oracle.c:orc_sinf
// (tests pin that against libm bit for bit); tests/test_gpu_parity.py pins THIS one against
// libm through rl_debug_sincosf.  |x| >= 120 uses the published large-argument reduction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rl {

static __device__ __constant__ uint32_t c_inv_pio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

// The double constants live in constant memory: a DMUL / DADD takes a constant-bank operand directly, where a 64-bit
// literal costs two moves per use (ncu source view of the lidar-fan kernel: 30 of the 109 instructions of a set-up
// batch's sincosf were such moves).
static __device__ __constant__ double c_sc[11] = {
    0x1.45F306DC9C883p+23,      // 0 HPI_INV
    0x1.921FB54442D18p0,        // 1 HPI
    0x1p0,                      // 2 C0
    -0x1.ffffffd0c621cp-2,      // 3 C1
    0x1.55553e1068f19p-5,       // 4 C2
    -0x1.6c087e89a359dp-10,     // 5 C3
    0x1.99343027bf8c3p-16,      // 6 C4
    -0x1.555545995a603p-3,      // 7 S1
    0x1.1107605230bc4p-7,       // 8 S2
    -0x1.994eb3774cf24p-13,     // 9 S3
    0x1.921FB54442D18p-62};     // 10 PI63
#ifndef RL_SC_LITERALS
#define RL_SC_HPI_INV c_sc[0]
#define RL_SC_HPI c_sc[1]
#define RL_SC_C0 c_sc[2]
#define RL_SC_C1 c_sc[3]
#define RL_SC_C2 c_sc[4]
#define RL_SC_C3 c_sc[5]
#define RL_SC_C4 c_sc[6]
#define RL_SC_S1 c_sc[7]
#define RL_SC_S2 c_sc[8]
#define RL_SC_S3 c_sc[9]
#define RL_SC_PI63 c_sc[10]
#else
#define RL_SC_HPI_INV 0x1.45F306DC9C883p+23
#define RL_SC_HPI 0x1.921FB54442D18p0
#define RL_SC_C0 0x1p0
#define RL_SC_C1 (-0x1.ffffffd0c621cp-2)
#define RL_SC_C2 0x1.55553e1068f19p-5
#define RL_SC_C3 (-0x1.6c087e89a359dp-10)
#define RL_SC_C4 0x1.99343027bf8c3p-16
#define RL_SC_S1 (-0x1.555545995a603p-3)
#define RL_SC_S2 0x1.1107605230bc4p-7
#define RL_SC_S3 (-0x1.994eb3774cf24p-13)
#define RL_SC_PI63 0x1.921FB54442D18p-62
#endif

// sine series on the reduced argument (sincosf.h: sinf_poly, even n)
__device__ __forceinline__ float sc_sin_poly(double x, double x2) {
  double x3 = __dmul_rn(x, x2);
  double s1 = __dadd_rn(RL_SC_S2, __dmul_rn(x2, RL_SC_S3));
  double x7 = __dmul_rn(x3, x2);
  double s = __dadd_rn(x, __dmul_rn(x3, RL_SC_S1));
  return __double2float_rn(__dadd_rn(s, __dmul_rn(x7, s1)));
}
// cosine series (odd n), positive-coefficient table
__device__ __forceinline__ float sc_cos_poly(double x2) {
  double x4 = __dmul_rn(x2, x2);
  double c2 = __dadd_rn(RL_SC_C3, __dmul_rn(x2, RL_SC_C4));
  double c1 = __dadd_rn(RL_SC_C0, __dmul_rn(x2, RL_SC_C1));
  double x6 = __dmul_rn(x4, x2);
  double c = __dadd_rn(c1, __dmul_rn(x4, RL_SC_C2));
  return __double2float_rn(__dadd_rn(c, __dmul_rn(x6, c2)));
}

__device__ __forceinline__ double sc_reduce_large(uint32_t xi, int* np) {
  const uint32_t* arr = &c_inv_pio4[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint64_t)(uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return __dmul_rn(x, RL_SC_PI63);
}

// sinf(y) and cosf(y) together: one argument reduction, two polynomials.
//
// Written without data-dependent branches for |y| < 120 (a warp of rays holds 32 unrelated headings):
//  * the published |y| < pi/4 shortcut (abstop12 < 0x3f4, i.e. |y| < 0.75) is the general path with n = 0 --
//    reduce_fast then returns x - 0 * hpi = x and the same two polynomials run on it;
//  * the published sign handling multiplies the argument of the odd polynomial by sign[n & 3] and switches to
//    a table of negated coefficients for n & 2; every operation in both polynomials is odd / linear in that
//    sign and round-to-nearest is symmetric, so negating the float results is bit-identical and saves the
//    multiplications.
__device__ __forceinline__ void rl_sincosf(float y, float* sp, float* cp) {
  const uint32_t bits = __float_as_uint(y);
  const uint32_t top = (bits >> 20) & 0x7ff;
  double x = (double)y;
  int n;
  int q;  // quadrant index that selects the signs
  if (top < 0x42fu) {  // |y| < 120
    const double r = __dmul_rn(x, RL_SC_HPI_INV);
    n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, RL_SC_HPI));
    q = n;
  } else if (top < 0x7f8u) {
    const int sign = bits >> 31;
    x = sc_reduce_large(bits, &n);
    q = n + sign;
  } else {  // inf / nan
    *sp = __fsub_rn(y, y);
    *cp = *sp;
    return;
  }
  const double x2 = __dmul_rn(x, x);
  float a = sc_sin_poly(x, x2);  // even-index polynomial, argument sign[q & 3] = {1,-1,-1,1}
  float b = sc_cos_poly(x2);     // odd-index polynomial, negated table for q & 2
  a = (((q + 1) & 2) != 0) ? -a : a;
  b = ((q & 2) != 0) ? -b : b;
  float sn = ((n & 1) == 0) ? a : b;  // sin uses polynomial n, cos uses n ^ 1
  float cs = ((n & 1) == 0) ? b : a;
  if (top < 0x398u) {  // |y| < 2^-12: sinf returns y, cosf returns 1
    sn = y;
    cs = 1.0f;
  }
  *sp = sn;
  *cp = cs;
}

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
// C (int)f on x86 (cvttss2si): truncation; out of range / NaN -> INT_MIN
__device__ __forceinline__ int f2i(float f) {
  if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
  return __float2int_rz(f);
}

}  // namespace rl
