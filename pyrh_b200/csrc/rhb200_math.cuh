// rhb200_math.cuh -- exp / pow / sin / cos that reproduce glibc 2.39's results bit for bit.
//
// The reference links glibc's libm; 1-ulp differences in exp() are amplified by the
// cancellation in Bezier3_coeffs (bezier_aux.c:341-349) and can flip an fp32 rounding of the
// DELO matrix (hard part 2), so the device path evaluates the *same
// algorithms* glibc 2.39 selects on FMA-capable x86-64 (the ifunc'ed __exp_fma, __pow_fma,
// __sin_fma, __cos_fma variants):
//   exp, pow : table-driven algorithms of S. Nagy (ARM optimized-routines; glibc
//              sysdeps/ieee754/dbl-64/e_exp.c, e_pow.c), N = 128 tables
//   sin, cos : IBM Accurate Mathematical Library (glibc sysdeps/ieee754/dbl-64/s_sin.c):
//              1/128-spaced sin/cos table + short polynomials, 3-part pi/2 reduction
// The placement of fused multiply-adds follows the instruction sequence of those variants
// (every fused operation below is written as RH_FMA, everything else rounds separately:
// the file must be compiled with -fmad=false / -ffp-contract=off).  Tables come from
// tools/gen_math_tables.py (computed from first principles, word-identical to glibc's).
//
// Domains outside what the hot path can reach fall back to the toolchain's libm and are
// documented at each function.  The same source compiles for the host (plain C++), which is how
// tests/test_math_cpu.py checks it against this machine's glibc on 10^7 arguments.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include "rhb200_math_tables.inc"
#include "rhb200_log_tables.inc"
#include "rhb200_atan_table.inc"

#if defined(__CUDACC__)
#define RH_FN __device__ __forceinline__
#define RH_TABLE static __device__ const
#define RH_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RH_ASUINT(x) ((uint64_t) __double_as_longlong(x))
#define RH_ASDOUBLE(u) __longlong_as_double((long long) (u))
#else
#define RH_FN static inline
#define RH_TABLE static const
#define RH_FMA(a, b, c) __builtin_fma((a), (b), (c))
static inline uint64_t rh_asuint_(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static inline double rh_asdouble_(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
#define RH_ASUINT(x) rh_asuint_(x)
#define RH_ASDOUBLE(u) rh_asdouble_(u)
#endif

namespace rhm {

RH_TABLE unsigned long long exp_tab[256] = { RH_EXP_TABLE };
RH_TABLE double sincos_tab[440] = { RH_SINCOS_TABLE };
RH_TABLE double powlog_tab[512] = { RH_POWLOG_TABLE };

// ------------------------------------------------------------------------ exp
namespace detail {

constexpr double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p+52,
                 NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47,
                 C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3,
                 C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;

// results with |x| >= 512: scale may over/underflow (glibc e_exp.c specialcase())
RH_FN double exp_specialcase(double tmp, uint64_t sbits, uint64_t ki)
{
  if ((ki & 0x80000000ull) == 0) {
    sbits -= 1009ull << 52;
    const double scale = RH_ASDOUBLE(sbits);
    return 0x1p1009 * RH_FMA(scale, tmp, scale);
  }
  sbits += 1022ull << 52;
  const double scale = RH_ASDOUBLE(sbits);
  const double st = scale * tmp;
  double y = scale + st;
  if (fabs(y) < 1.0) {
    const double one = (y < 0.0) ? -1.0 : 1.0;
    double lo = scale - y + st;
    const double hi = one + y;
    lo = one - hi + y + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = RH_ASDOUBLE(sbits & 0x8000000000000000ull);
  }
  return 0x1p-1022 * y;
}

// exp(x + xtail) for the reduced argument handling shared by exp and pow
RH_FN double exp_core(double x, double xtail, bool has_tail, uint32_t abstop)
{
  const double z = RH_FMA(x, InvLn2N, Shift);
  const uint64_t ki = RH_ASUINT(z);
  const double kd = z - Shift;
  double r = RH_FMA(kd, NegLn2hiN, x);
  r = RH_FMA(kd, NegLn2loN, r);
  if (has_tail) r = xtail + r;
  const uint64_t idx = 2 * (ki & 127);
  const uint64_t top = ki << 45;
  const double tail = RH_ASDOUBLE(exp_tab[idx]);
  const uint64_t sbits = exp_tab[idx + 1] + top;
  const double r2 = r * r;
  const double p23 = RH_FMA(r, C3, C2);
  const double tr = r + tail;
  const double p45 = RH_FMA(r, C5, C4);
  const double t1 = RH_FMA(p23, r2, tr);
  const double tmp = RH_FMA(r2 * r2, p45, t1);
  if (abstop == 0) return exp_specialcase(tmp, sbits, ki);
  const double scale = RH_ASDOUBLE(sbits);
  return RH_FMA(scale, tmp, scale);
}

}  // namespace detail

// glibc 2.39 __exp (e_exp.c), FMA variant.  Full domain.
RH_FN double rh_exp(double x)
{
  const uint64_t ix = RH_ASUINT(x);
  uint32_t abstop = (uint32_t) (ix >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;          // |x| < 2^-54
    if (abstop >= 0x409u) {                                       // |x| >= 1024, inf, nan
      if (ix == 0xfff0000000000000ull) return 0.0;
      if (abstop >= 0x7ffu) return 1.0 + x;
      return (ix >> 63) ? 0.0 : HUGE_VAL;
    }
    abstop = 0;                                                   // 512 <= |x| < 1024
  }
  return detail::exp_core(x, 0.0, false, abstop);
}

// ------------------------------------------------------------------------ pow
// glibc 2.39 __pow (e_pow.c), FMA variant, for x positive, finite and normal and
// 2^-65 <= |y| < 2^63 -- the only domain the hot path reaches (T^0.3, T^((1-alpha)/2),
// (C1/T)^-1.5).  Anything else is forwarded to the toolchain's pow().
RH_FN double rh_pow(double x, double y)
{
  const uint64_t ix = RH_ASUINT(x), iy = RH_ASUINT(y);
  const uint32_t topx = (uint32_t) (ix >> 52), topy = (uint32_t) (iy >> 52);
  if (topx - 1u >= 0x7feu || (topy & 0x7ff) - 0x3beu >= 0x80u) return pow(x, y);

  // log_inline(): log(x) = k ln2 + log(c) + log1p(z/c - 1) in double-double (hi, lo)
  constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
  constexpr double A0 = -0x1p-1, A1 = -0x1.5555555555560p-1, A2 = 0x1.0000000000006p-1,
                   A3 = 0x1.999999959554ep-1, A4 = -0x1.555555529a47ap-1,
                   A5 = -0x1.2495b9b4845e9p+0, A6 = 0x1.0002b8b263fc3p+0;
  const uint64_t tmp = ix - 0x3fe6955500000000ull;
  const int i = (int) ((tmp >> 45) & 127);
  const int k = (int) ((int64_t) tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffull << 52));
  const double z = RH_ASDOUBLE(iz), kd = (double) k;
  const double invc = powlog_tab[4*i], logc = powlog_tab[4*i + 2], logctail = powlog_tab[4*i + 3];
  const double r = RH_FMA(z, invc, -1.0);
  const double t1 = RH_FMA(kd, Ln2hi, logc);
  const double t2 = t1 + r;
  const double lo1 = RH_FMA(kd, Ln2lo, logctail);
  const double lo2 = t1 - t2 + r;
  const double ar = A0 * r, ar2 = r * ar, ar3 = r * ar2;
  const double hi = t2 + ar2;
  const double lo3 = RH_FMA(ar, r, -ar2);
  const double lo4 = t2 - hi + ar2;
  const double q56 = RH_FMA(r, A6, A5), q34 = RH_FMA(r, A4, A3), q12 = RH_FMA(r, A2, A1);
  const double q = RH_FMA(ar2, RH_FMA(q56, ar2, q34), q12);
  const double lo = RH_FMA(ar3, q, lo1 + lo2 + lo3 + lo4);
  const double lhi = hi + lo;
  const double llo = hi - lhi + lo;

  const double ehi = y * lhi;
  const double elo = RH_FMA(y, llo, RH_FMA(lhi, y, -ehi));

  // exp_inline(ehi, elo, sign_bias = 0)
  uint32_t abstop = (uint32_t) (RH_ASUINT(ehi) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + ehi;
    if (abstop >= 0x409u) return (RH_ASUINT(ehi) >> 63) ? 0.0 : HUGE_VAL;
    abstop = 0;
  }
  return detail::exp_core(ehi, elo, true, abstop);
}

// -------------------------------------------------------------------- sin, cos
namespace detail {

constexpr double big = 0x1.8p+45, toint = 0x1.8p+52, hpinv = 0x1.45f306dc9c883p-1,
                 hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54,
                 mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde973c000000p-27,
                 pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83,
                 s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13,
                 s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26,
                 sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7,
                 cs2 = 0x1p-1, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;

// s_sin.c do_cos(): cos(x + dx) for |x| < 0.86
RH_FN double do_cos(double x, double dx)
{
  if (x < 0) dx = -dx;
  const double u = big + fabs(x);
  const int k = ((int) (uint32_t) RH_ASUINT(u)) * 4;
  const double xr = (fabs(x) - (u - big)) + dx;
  const double xx = xr * xr;
  const double s = RH_FMA(xr * xx, RH_FMA(sn5, xx, sn3), xr);
  const double c = xx * RH_FMA(RH_FMA(cs6, xx, cs4), xx, cs2);
  const double sn = sincos_tab[k], ssn = sincos_tab[k+1], cs = sincos_tab[k+2], ccs = sincos_tab[k+3];
  const double cor = RH_FMA(-s, sn, RH_FMA(-c, cs, RH_FMA(-s, ssn, ccs)));
  return cs + cor;
}

// s_sin.c do_sin(): sin(x + dx) for |x| < 0.86
RH_FN double do_sin(double x, double dx)
{
  const double xold = x;
  if (fabs(x) < 0.126) {                               // TAYLOR_SIN
    const double xx = x * x;
    const double p = RH_FMA(RH_FMA(RH_FMA(RH_FMA(s5, xx, s4), xx, s3), xx, s2), xx, s1);
    const double t = RH_FMA(xx, RH_FMA(p, x, -(0.5 * dx)), dx);
    return x + t;
  }
  if (x <= 0) dx = -dx;
  const double u = big + fabs(x);
  const int k = ((int) (uint32_t) RH_ASUINT(u)) * 4;
  const double xr = fabs(x) - (u - big);
  const double xx = xr * xr;
  const double s = xr + RH_FMA(xr * xx, RH_FMA(sn5, xx, sn3), dx);
  const double c = RH_FMA(xr, dx, xx * RH_FMA(RH_FMA(cs6, xx, cs4), xx, cs2));
  const double sn = sincos_tab[k], ssn = sincos_tab[k+1], cs = sincos_tab[k+2], ccs = sincos_tab[k+3];
  const double cor = RH_FMA(s, cs, RH_FMA(-c, sn, RH_FMA(s, ccs, ssn)));
  return copysign(sn + cor, xold);
}

// s_sin.c reduce_sincos(): x = n pi/2 + (a + da), 2.426 < |x| < 105414350
RH_FN int reduce_sincos(double x, double &a, double &da)
{
  const double t = RH_FMA(x, hpinv, toint);
  const double xn = t - toint;
  const int n = (int) ((uint32_t) RH_ASUINT(t) & 3u);
  double y = RH_FMA(-xn, mp1, x);
  y = RH_FMA(-xn, mp2, y);
  const double t2 = RH_FMA(-xn, pp3, y);
  double db = RH_FMA(-pp3, xn, y - t2);
  const double b = RH_FMA(-xn, pp4, t2);
  db = db + RH_FMA(-xn, pp4, t2 - b);
  a = b; da = db;
  return n;
}

RH_FN double do_sincos(double a, double da, int n)
{
  const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
  return (n & 2) ? -r : r;
}

}  // namespace detail

// glibc 2.39 __sin (s_sin.c), FMA variant, for |x| < 105414350; larger arguments (glibc's
// __branred path), inf and nan are forwarded to the toolchain's sin().
RH_FN double rh_sin(double x)
{
  using namespace detail;
  const uint32_t k = (uint32_t) (RH_ASUINT(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;
  if (k < 0x3feb6000u) return do_sin(x, 0.0);
  if (k < 0x400368fdu) return copysign(do_cos(hp0 - fabs(x), hp1), x);
  if (k < 0x419921fbu) { double a, da; const int n = reduce_sincos(x, a, da); return do_sincos(a, da, n); }
  return sin(x);
}

RH_FN double rh_cos(double x)
{
  using namespace detail;
  const uint32_t k = (uint32_t) (RH_ASUINT(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;
  if (k < 0x3feb6000u) return do_cos(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = hp0 - fabs(x);
    const double a = y + hp1;
    const double da = (y - a) + hp1;
    return do_sin(a, da);
  }
  if (k < 0x419921fbu) { double a, da; const int n = reduce_sincos(x, a, da); return do_sincos(a, da, n + 1); }
  return cos(x);
}

// ---------------------------------------------------------------------------------------------
// log: glibc 2.39 sysdeps/ieee754/dbl-64/e_log.c (S. Nagy), the __log_fma ifunc variant.  Fused operations are
// placed exactly as in that variant's instruction sequence (libm.so.6 of this image, function at the
// IRELATIVE target of log@GLIBC_2.29); tables: rhb200_log_tables.inc.
RH_TABLE double rh_log_A[5] = {RH_LOG_POLY_A};
RH_TABLE double rh_log_B[11] = {RH_LOG_POLY_B};
RH_TABLE double rh_log_T[256] = {RH_LOG_TAB};      // {invc, logc} x 128

RH_FN double rh_log(double x)
{
  uint64_t ix = RH_ASUINT(x);
  const uint64_t LO = 0x3FEE000000000000ULL, HI = 0x3FF1090000000000ULL;     // 1 - 2^-4, 1 + 0x1.09p-4
  if (ix - LO < HI - LO) {
    if (ix == 0x3FF0000000000000ULL) return 0.0;
    const double *B = rh_log_B;
    const double r = x - 1.0, r2 = r * r, r3 = r * r2;
    double p2 = RH_FMA(r, B[2], B[1]);   p2 = RH_FMA(r2, B[3], p2);           // B1 + r B2 + r2 B3
    double p5 = RH_FMA(r, B[5], B[4]);   p5 = RH_FMA(r2, B[6], p5);           // B4 + r B5 + r2 B6
    double p8 = RH_FMA(r, B[8], B[7]);   p8 = RH_FMA(r2, B[9], p8);  p8 = RH_FMA(r3, B[10], p8);
    double P = RH_FMA(p8, r3, p5);
    P = RH_FMA(P, r3, p2);
    const double w = RH_FMA(r, 0x1p27, r);                                   // r + r*2^27
    const double rhi = RH_FMA(-0x1p27, r, w);
    const double rlo = r - rhi;
    const double rhi2 = rhi * rhi;
    const double hi = RH_FMA(rhi2, B[0], r);
    double lo = RH_FMA(rhi2, B[0], r - hi);
    lo = RH_FMA(B[0] * rlo, rhi + r, lo);
    const double y = RH_FMA(P, r3, lo);
    return hi + y;
  }
  const uint32_t top = (uint32_t) (ix >> 48);
  if (top - 0x0010 >= 0x7ff0 - 0x0010) {
    if (ix * 2 == 0) return -1.0 / 0.0;                                      // log(+-0) = -inf
    if (ix == 0x7FF0000000000000ULL) return x;                               // log(inf) = inf
    if ((top & 0x8000) || (top & 0x7ff0) == 0x7ff0) return (x - x) / (x - x);
    ix = RH_ASUINT(x * 0x1p52);                                              // subnormal: normalise
    ix -= 52ULL << 52;
  }
  const uint64_t tmp = ix - 0x3FE6000000000000ULL;
  const int i = (int) ((tmp >> 45) & 127);
  const int64_t k = (int64_t) tmp >> 52;
  const uint64_t iz = ix - (tmp & (0xFFFULL << 52));
  const double invc = rh_log_T[2*i], logc = rh_log_T[2*i + 1];
  const double z = RH_ASDOUBLE(iz);
  const double r = RH_FMA(z, invc, -1.0);
  const double kd = (double) k;
  const double w = RH_FMA(kd, RH_LOG_LN2HI, logc);
  const double hi = w + r;
  const double lo = RH_FMA(kd, RH_LOG_LN2LO, (w - hi) + r);
  const double r2 = r * r;
  const double *A = rh_log_A;
  const double q = RH_FMA(RH_FMA(r, A[4], A[3]), r2, RH_FMA(r, A[2], A[1]));
  const double y = RH_FMA(r * r2, q, RH_FMA(r2, A[0], lo));
  return y + hi;
}

// log10: glibc 2.39 sysdeps/ieee754/dbl-64/e_log10.c (fdlibm scaling around the log above; no fused operations)
RH_FN double rh_log10(double x)
{
  const double two54 = 1.80143985094819840000e+16, ivln10 = 4.34294481903251816668e-01,
               log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
  uint64_t u = RH_ASUINT(x);
  int32_t hx = (int32_t) (u >> 32);
  const uint32_t lx = (uint32_t) u;
  int32_t k = 0;
  if (hx < 0x00100000) {
    if (((hx & 0x7fffffff) | lx) == 0) return -two54 / fabs(x);
    if (hx < 0) return (x - x) / (x - x);
    k -= 54; x *= two54;
    u = RH_ASUINT(x); hx = (int32_t) (u >> 32);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  const int32_t i = (int32_t) (((uint32_t) k & 0x80000000u) >> 31);
  hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
  const double y = (double) (k + i);
  x = RH_ASDOUBLE(((uint64_t) (uint32_t) hx << 32) | (RH_ASUINT(x) & 0xffffffffULL));
  const double z = y * log10_2lo + ivln10 * rh_log(x);
  return z + y * log10_2hi;
}


// ------------------------------------------------------------------------ atan
// glibc 2.39 sysdeps/ieee754/dbl-64/s_atan.c (IBM Accurate Mathematical Library, slow paths removed in 2.28), as
// compiled for FMA hosts (__atan_fma): five ranges of u = |x| with the fused operations where that variant has them.
//   u < A = 0x1.bb67ap-27: x;  A <= u < 1/16: odd polynomial d3..d13;  1/16 <= u < 1: table cij (241 rows: x_i,
//   atan x_i, c1..c5) + degree-5 polynomial in u - x_i;  1 <= u < 16: pi/2 - atan(1/u) through the same table with the
//   rounding error of 1/u carried along;  16 <= u < 0x1.49ff2p+52: pi/2 - (1/u) polynomial;  beyond: +-pi/2.
RH_TABLE double atan_cij[241 * 7] = { RH_ATAN_TABLE };

RH_FN double rh_atan(double x)
{
  const double A = 0x1.bb67a00000000p-27, B = 0x1.0p-4, C = 1.0, D = 16.0, E = 0x1.49ff200000000p+52;
  const double d3 = -0x1.5555555555555p-2, d5 = 0x1.99999999997fdp-3, d7 = -0x1.24924923f7603p-3,
               d9 = 0x1.c71c6e5129a3bp-4, d11 = -0x1.7458022b13c25p-4, d13 = 0x1.375f08b31cbcep-4;
  const double HPI = 0x1.921fb54442d18p+0, HPI1 = 0x1.1a62633145c07p-54, TWO52 = 0x1.0p+52, TWO8 = 256.0;
  if (x != x) return x + x;
  const double u = (x < 0) ? -x : x;
  const uint64_t sign = RH_ASUINT(x) & 0x8000000000000000ULL;
  double y;
  if (u < C) {
    if (u < B) {
      if (u < A) return x;
      const double v = x * x;
      double yy = RH_FMA(d13, v, d11);
      yy = RH_FMA(yy, v, d9);
      yy = RH_FMA(yy, v, d7);
      yy = RH_FMA(yy, v, d5);
      yy = RH_FMA(yy, v, d3);
      return RH_FMA(x * v, yy, x);
    }
    const int i = (int) (RH_FMA(u, TWO8, TWO52) - TWO52) - 16;
    const double *c = atan_cij + 7 * i;
    const double z = u - c[0];
    double yy = RH_FMA(c[6], z, c[5]);
    yy = RH_FMA(yy, z, c[4]);
    yy = RH_FMA(yy, z, c[3]);
    yy = RH_FMA(yy, z, c[2]);
    y = RH_FMA(z, yy, c[1]);
  } else if (u < D) {
    const double w = 1.0 / u;
    const double t1 = w * u;
    const double t2 = RH_FMA(u, w, -t1);
    const double a = (1.0 - t1) - t2;                       // ww = w * a, fused below
    const int i = (int) (RH_FMA(w, TWO8, TWO52) - TWO52) - 16;
    const double *c = atan_cij + 7 * i;
    const double z = RH_FMA(a, w, w - c[0]);
    double yy = RH_FMA(c[6], z, c[5]);
    yy = RH_FMA(yy, z, c[4]);
    yy = RH_FMA(yy, z, c[3]);
    yy = RH_FMA(yy, z, c[2]);
    yy = RH_FMA(-z, yy, HPI1);
    y = (HPI - c[1]) + yy;
  } else if (u < E) {
    const double w = 1.0 / u;
    const double t1 = w * u;
    const double t3 = HPI - w;
    const double v = w * w;
    double yy = RH_FMA(d13, v, d11);
    yy = RH_FMA(yy, v, d9);
    yy = RH_FMA(yy, v, d7);
    yy = RH_FMA(yy, v, d5);
    yy = RH_FMA(yy, v, d3);
    const double t2 = RH_FMA(u, w, -t1);
    const double a = (1.0 - t1) - t2;
    const double cor = ((HPI - t3) - w) + HPI1;
    const double r = RH_FMA(-a, w, cor);
    const double r2 = RH_FMA(-(w * v), yy, r);
    y = t3 + r2;
  } else y = HPI;
  return RH_ASDOUBLE((RH_ASUINT(y) & 0x7fffffffffffffffULL) | sign);
}

}  // namespace rhm
