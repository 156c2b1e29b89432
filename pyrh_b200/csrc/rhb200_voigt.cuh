// rhb200_voigt.cuh -- Humlicek (1982) W(z) = H + iF, the Voigt / Faraday-Voigt pair.
// Reference: VoigtHumlicek rh/voigt.c:381-419, Humlicek1..4 rh/humlicek.c:28-117 with the
// struct-by-value complex helpers of rh/complex.c:27-155.  The operation ORDER of those
// helpers is kept (cmplx_mult: ar*br - ai*bi, ar*bi + ai*br; cmplx_div: (ar*br+ai*bi)/d ...)
// so every intermediate rounds like the reference; region selection is the integer part
// that must be bit-exact.
#pragma once
#include "rhb200_math.cuh"

namespace rhv {

struct cplx { double r, i; };

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  cplx c; c.r = a.r*b.r - a.i*b.i; c.i = a.r*b.i + a.i*b.r; return c;
}
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  cplx c; double d = b.r*b.r + b.i*b.i;
  c.r = (a.r*b.r + a.i*b.i) / d; c.i = (a.i*b.r - a.r*b.i) / d; return c;
}

__device__ __forceinline__ int humlicek_region(double a, double v) {
  double s = fabs(v) + a;                       // voigt.c:404
  if (s >= 15.0) return 1;
  if (s >= 5.5) return 2;
  if (a >= 0.195*fabs(v) - 0.176) return 3;
  return 4;
}

// NEED_F = false drops everything that only feeds Im W (the Faraday-Voigt function is used by
// RLKProfile only when MAGNETO_OPTICAL is on, kurucz.c:815-822): the imaginary quotient of the
// final complex division and, in region IV, the whole sin() evaluation.  Re W is unchanged bit
// for bit.
template <bool NEED_F>
__device__ __forceinline__ double humlicek_t(double a, double v, double *F) {
  cplx z = {a, -v}, z1, z2;
  double Wr, Wi = 0.0;
  const int reg = humlicek_region(a, v);
  // complex quotient z1/z2 (complex.c:84-99); the imaginary part only when needed
#define RH_CDIV_OUT(NUM, DEN)                                                   \
  { const double d__ = (DEN).r*(DEN).r + (DEN).i*(DEN).i;                       \
    Wr = ((NUM).r*(DEN).r + (NUM).i*(DEN).i) / d__;                             \
    if (NEED_F) Wi = ((NUM).i*(DEN).r - (NUM).r*(DEN).i) / d__; }
  if (reg == 1) {                               // humlicek.c:28-38
    z1.r = 0.5641896*z.r; z1.i = 0.5641896*z.i;
    z2 = cmul(z, z); z2.r = z2.r + 0.5;
    RH_CDIV_OUT(z1, z2)
  } else if (reg == 2) {                        // humlicek.c:43-55
    cplx u = cmul(z, z), t;
    z1.r = 0.5641896*u.r; z1.i = 0.5641896*u.i;
    t.r = z1.r + 1.410474; t.i = z1.i;
    z1 = cmul(z, t);
    t.r = u.r + 3.0; t.i = u.i;
    z2 = cmul(u, t);
    z2.r = z2.r + 0.75;
    RH_CDIV_OUT(z1, z2)
  } else if (reg == 3) {                        // humlicek.c:62-83
    const double A[5] = {0.5642236, 3.778987, 11.96482, 20.20933, 16.4955};
    const double B[5] = {6.699398, 21.69274, 39.27121, 38.82363, 16.4955};
    z1.r = A[0]; z1.i = 0.0;
    z2.r = z.r + B[0]; z2.i = z.i;
#pragma unroll
    for (int n = 1; n < 5; n++) {
      z1 = cmul(z1, z); z1.r = z1.r + A[n];
      z2 = cmul(z2, z); z2.r = z2.r + B[n];
    }
    RH_CDIV_OUT(z1, z2)
  } else {                                      // humlicek.c:90-117
    const double A[7] = {0.56419, 1.320522, 35.7668, 219.031, 1540.787, 3321.99, 36183.31};
    const double B[7] = {1.841439, 61.57037, 364.2191, 2186.181, 9022.228, 24322.84, 32066.6};
    cplx mz = {-1.0*z.r, -1.0*z.i};
    cplx u = cmul(z, mz);
    z1.r = A[0]; z1.i = 0.0;
    z2.r = u.r + B[0]; z2.i = u.i;
#pragma unroll
    for (int n = 1; n < 7; n++) {
      z1 = cmul(u, z1); z1.r = z1.r + A[n];
      z2 = cmul(u, z2); z2.r = z2.r + B[n];
    }
    cplx mu = {-1.0*u.r, -1.0*u.i};
    const double ex = rhm::rh_exp(mu.r);         // complex.c:104-109
    const cplx zz = cmul(z, z1);
    RH_CDIV_OUT(zz, z2)
    Wr = ex*rhm::rh_cos(mu.i) - Wr;
    if (NEED_F) Wi = ex*rhm::rh_sin(mu.i) - Wi;
  }
#undef RH_CDIV_OUT
  if (NEED_F) *F = Wi;
  return Wr;
}

__device__ __noinline__ double humlicek(double a, double v, double *F) { return humlicek_t<true>(a, v, F); }
__device__ __forceinline__ double humlicek_H(double a, double v) { return humlicek_t<false>(a, v, nullptr); }

}  // namespace rhv
