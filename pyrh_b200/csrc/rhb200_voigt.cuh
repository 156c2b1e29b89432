// rhb200_voigt.cuh -- Humlicek (1982) W(z) = H + iF, the Voigt / Faraday-Voigt pair.
// Reference: VoigtHumlicek rh/voigt.c:381-419, Humlicek1..4 rh/humlicek.c:28-117 with the
// struct-by-value complex helpers of rh/complex.c:27-155.  The operation ORDER of those
// helpers is kept (cmplx_mult: ar*br - ai*bi, ar*bi + ai*br; cmplx_div: (ar*br+ai*bi)/d ...)
// so every intermediate rounds like the reference; region selection is the integer part
// that must be bit-exact.
#pragma once
#include "rhb200_math.cuh"
#ifndef RH_PI
#define RH_PI 3.14159265358979
#endif

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

static __device__ __noinline__ double humlicek(double a, double v, double *F) { return humlicek_t<true>(a, v, F); }
__device__ __forceinline__ double humlicek_H(double a, double v) { return humlicek_t<false>(a, v, nullptr); }

// ---- VoigtArmstrong (voigt.c:126-243): unpolarised Voigt function H(a, v) of Profile() for
//      NO_STOKES active lines.  K1 (Chebyshev series + recurrence with data-dependent exit) and K3
//      (10-point quadrature) are pure + - * / plus exp/cos; K2 (1 <= a < 2.5, v < 4) calls atan() and log(),
//      evaluated with the glibc-exact rh_atan / rh_log of rhb200_math.cuh: all three branches are
//      bit-identical to the reference.
__device__ __forceinline__ int armstrong_region(double a, double v) {
  v = fabs(v);
  if ((a < 1.0 && v < 4.0) || (a < 1.8/(v + 1.0))) return 1;
  if (a < 2.5 && v < 4.0) return 2;
  return 3;
}

static __device__ __noinline__ double voigt_armstrong(double a, double v) {
  const double T[10] = {0.2453407083, 0.7374737285, 1.2340762153, 1.7385377121, 2.2549740020,
                        2.7888060584, 3.3478545673, 3.9447640401, 4.6036824495, 5.3874808900};
  const double W[10] = {4.6224366960e-01, 2.8667550536e-01, 1.0901720602e-01, 2.4810520887e-02,
                        3.2437733422e-03, 2.2833863601e-04, 7.8025564785e-06, 1.0860693707e-07,
                        4.3993409922e-10, 2.2293936455e-13};
  if (v < 0.0) v = -v;
  const int reg = armstrong_region(a, v);
  if (reg == 1) {                                           // VoigtK1, voigt.c:146-211
    const double Cc[34] = { 0.1999999999972224, -0.1840000000029998, 0.1558399999965025, -0.1216640000043988,
      0.0877081599940391, -0.0585141248086907, 0.0362157301623914, -0.0208497654398036, 0.0111960116346270,
      -0.56231896167109e-02, 0.26487634172265e-02, -0.11732670757704e-02, 0.4899519978088e-03, -0.1933630801528e-03,
      0.722877446788e-04, -0.256555124979e-04, 0.86620736841e-05, -0.27876379719e-05, 0.8566873627e-06,
      -0.2518433784e-06, 0.709360221e-07, -0.191732257e-07, 0.49801256e-08, -0.12447734e-08, 0.2997777e-09,
      -0.696450e-10, 0.156262e-10, -0.33897e-11, 0.7116e-12, -0.1447e-12, 0.285e-13, -0.55e-14, 0.10e-14, -0.2e-15 };
    const double a2 = a*a, v2 = v*v;
    double u1, dn01, dn02;
    if ((v2 - a2) > 70.0) u1 = 0.0;
    else u1 = rhm::rh_exp(a2 - v2) * rhm::rh_cos(2.0*v*a);
    if (v > 5.0) {
      const double v2i = 1.0 / v2;
      dn01 = -v2i * (0.5 + v2i*(0.75 + v2i*(1.875 + v2i*(6.5625 +
             v2i*(29.53125 + v2i*(1162.4218 + v2i*1055.7421))))));
      dn02 = (1.0 - dn01) / (2.0 * v);
    } else {
      double bn01 = 0.0, bn02 = 0.0, bn = 0.0;
      const double v1 = v / 5.0, coef = 4.0 * v1*v1 - 2.0;
#pragma unroll 1
      for (int n = 33; n >= 0; n--) { bn = coef*bn01 - bn02 + Cc[n]; bn02 = bn01; bn01 = bn; }
      dn02 = v1*(bn - bn02);
      dn01 = 1.0 - 2.0*v*dn02;
    }
    double funct = a*dn01;
    if (a > 1.0E-08) {
      double q = 1.0, an = a;
#pragma unroll 1
      for (int n = 2; n <= 50; n++) {
        const double dn = (v*dn01 + dn02) * (-2.0/n);
        dn02 = dn01; dn01 = dn;
        if (n % 2) {
          q = -q; an *= a2;
          const double g = dn * an;
          funct += q*g;
          if (fabs(g/funct) <= 1.0E-08) return (u1 - 1.12837917*funct);
        }
      }
    }
    return (u1 - 1.12837917*funct);
  }
  if (reg == 2) {                                           // VoigtK2, voigt.c:215-229
    double g = 0.0;
    const double a2 = a*a;
#pragma unroll 1
    for (int n = 0; n < 10; n++) {
      const double r = T[n] - v, s = T[n] + v;
      g += (4.0*T[n]*T[n] - 2.0) * (r*rhm::rh_atan(r/a) + s*rhm::rh_atan(s/a) -
            0.5*a*(rhm::rh_log(a2 + r*r) + rhm::rh_log(a2 + s*s))) * W[n];
    }
    return g/RH_PI;
  }
  double g = 0.0;                                           // VoigtK3, voigt.c:233-243
  const double a2 = a*a;
#pragma unroll 1
  for (int n = 0; n < 10; n++)
    g += (1.0/((v - T[n])*(v - T[n]) + a2) + 1.0/((v + T[n])*(v + T[n]) + a2)) * W[n];
  return (a*g)/RH_PI;
}

}  // namespace rhv
