/* oracle/port/rhport_lines.c -- TEST INFRASTRUCTURE ONLY (see rhport.h).
 *
 * CPU restatement of the reference's LTE Kurucz-line opacity:
 *   Voigt/Faraday-Voigt (Humlicek 1982)   rh/voigt.c:381-419, rh/humlicek.c:28-117,
 *                                         rh/complex.c:27-155
 *   Linear / Hunt                          rh/linear.c:22-58, rh/hunt.c:17-78
 *   LTEpops_elem (Saha)                    rh/ltepops.c:116-159
 *   RLKProfile                             rh/kurucz.c:729-828
 *   rlk_opacity                            rh/kurucz.c:511-725
 * Arithmetic order follows the reference expression by expression; compile
 * with -O2 -ffp-contract=off (the reference x86-64 build has no FMA).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "rhport.h"

/* ---- complex helpers, same operation order as complex.c ---------------- */
typedef struct { double r, i; } cplx;

static inline cplx c_mul(cplx a, cplx b)        /* complex.c:53-64 */
{ cplx c; c.r = a.r*b.r - a.i*b.i; c.i = a.r*b.i + a.i*b.r; return c; }
static inline cplx c_scl(double a, cplx b)      /* complex.c:69-79 */
{ cplx c; c.r = a*b.r; c.i = a*b.i; return c; }
static inline cplx c_div(cplx a, cplx b)        /* complex.c:84-99 */
{ cplx c; double d = b.r*b.r + b.i*b.i;
  c.r = (a.r*b.r + a.i*b.i) / d; c.i = (a.i*b.r - a.r*b.i) / d; return c; }
static inline cplx c_addr(cplx a, double b)     /* complex.c:128-138 */
{ cplx c; c.r = a.r + b; c.i = a.i; return c; }

int rp_humlicek_region(double a, double v)      /* voigt.c:404-413 */
{
  double s = fabs(v) + a;
  if (s >= 15.0) return 1;
  if (s >= 5.5)  return 2;
  if (a >= 0.195*fabs(v) - 0.176) return 3;
  return 4;
}

double rp_voigt_humlicek(double a, double v, double *F)
{
  cplx z = { a, -v }, W, z1, z2, u;
  int n;
  switch (rp_humlicek_region(a, v)) {
  case 1:                                       /* humlicek.c:28-38 */
    z1 = c_scl(0.5641896, z);
    z2 = c_addr(c_mul(z, z), 0.5);
    W  = c_div(z1, z2);
    break;
  case 2:                                       /* humlicek.c:43-55 */
    u  = c_mul(z, z);
    z1 = c_scl(0.5641896, u);
    z1 = c_mul(z, c_addr(z1, 1.410474));
    z2 = c_mul(u, c_addr(u, 3.0));
    z2 = c_addr(z2, 0.75);
    W  = c_div(z1, z2);
    break;
  case 3: {                                     /* humlicek.c:62-83 */
    static const double A[5] = {0.5642236, 3.778987, 11.96482, 20.20933, 16.4955};
    static const double B[5] = {6.699398, 21.69274,  39.27121, 38.82363, 16.4955};
    z1.r = A[0]; z1.i = 0.0;
    z2 = c_addr(z, B[0]);
    for (n = 1; n < 5; n++) {
      z1 = c_addr(c_mul(z1, z), A[n]);
      z2 = c_addr(c_mul(z2, z), B[n]);
    }
    W = c_div(z1, z2);
    break;
  }
  default: {                                    /* humlicek.c:90-117 */
    static const double A[7] =
      {0.56419, 1.320522, 35.7668, 219.031, 1540.787, 3321.99, 36183.31};
    static const double B[7] =
      {1.841439, 61.57037, 364.2191, 2186.181, 9022.228, 24322.84, 32066.6};
    cplx mu_, e, q;
    u = c_mul(z, c_scl(-1.0, z));
    z1.r = A[0]; z1.i = 0.0;
    z2 = c_addr(u, B[0]);
    for (n = 1; n < 7; n++) {
      z1 = c_addr(c_mul(u, z1), A[n]);
      z2 = c_addr(c_mul(u, z2), B[n]);
    }
    mu_ = c_scl(-1.0, u);
    { cplx cs; cs.r = cos(mu_.i); cs.i = sin(mu_.i);       /* complex.c:104-109 */
      e = c_scl(exp(mu_.r), cs); }
    q = c_div(c_mul(z, z1), z2);
    W.r = e.r - q.r; W.i = e.i - q.i;
  }
  }
  if (F) *F = W.i;
  return W.r;
}

/* ---- Linear() with hunt=TRUE; for xmin < x < xmax Hunt() returns the unique
        bracket j with xt[j] <= x < xt[j+1], found here by bisection
        (linear.c:22-58, hunt.c:59-68; ascending tables only: Tpf is) ------- */
void rp_linear(int Ntable, const double *xt, const double *yt, int N,
               const double *x, double *y)
{
  double xmin = xt[0], xmax = xt[Ntable-1];
  for (int n = 0; n < N; n++) {
    if (x[n] <= xmin) y[n] = yt[0];
    else if (x[n] >= xmax) y[n] = yt[Ntable-1];
    else {
      int lo = 0, hi = Ntable;
      while (hi - lo > 1) {
        int mid = (hi + lo) >> 1;
        if (x[n] >= xt[mid]) lo = mid; else hi = mid;
      }
      double fx = (xt[lo+1] - x[n]) / (xt[lo+1] - xt[lo]);
      y[n] = fx*yt[lo] + (1 - fx)*yt[lo+1];
    }
  }
}

/* ---- LTEpops_elem, ltepops.c:116-159 ------------------------------------ */
void rp_ltepops_elem(const rp_linetable *lt, int ielem, const rp_column *col, double *n)
{
  const double *e = lt->elems + (long) ielem * RE_NFIELD;
  int Nst = (int) e[RE_NSTAGE], pfrow = (int) e[RE_PFROW], N = col->Ndep, k, i;
  double C1 = (RP_HPLANCK/(2.0*RP_PI*RP_M_ELECTRON)) * (RP_HPLANCK/RP_KBOLTZMANN);
  double *sum = malloc(N*sizeof(double)), *CT_ne = malloc(N*sizeof(double)),
         *Uk = malloc(N*sizeof(double)), *Ukp1 = malloc(N*sizeof(double)), *tmp;

  for (k = 0; k < N; k++) {
    CT_ne[k] = 2.0 * pow(C1/col->T[k], -1.5) / col->ne[k];
    sum[k] = 1.0;
    n[k] = 1.0;
  }
  rp_linear(lt->npf, lt->Tpf, lt->pf + (long) pfrow*lt->npf, N, col->T, Uk);
  for (i = 1; i < Nst; i++) {
    rp_linear(lt->npf, lt->Tpf, lt->pf + (long)(pfrow+i)*lt->npf, N, col->T, Ukp1);
    for (k = 0; k < N; k++) {
      n[i*N+k] = n[(i-1)*N+k] * CT_ne[k] *
        exp(Ukp1[k] - Uk[k] - e[RE_IONPOT0 + i-1]/(RP_KBOLTZMANN*col->T[k]));
      sum[k] += n[i*N+k];
    }
    tmp = Uk; Uk = Ukp1; Ukp1 = tmp;
  }
  for (k = 0; k < N; k++) n[k] = e[RE_ABUND] * col->nHtot[k] / sum[k];
  for (i = 1; i < Nst; i++)
    for (k = 0; k < N; k++) n[i*N+k] *= n[k];
  free(sum); free(CT_ne); free(Uk); free(Ukp1);
}

/* ---- RLKProfile, kurucz.c:729-828 (magneto_optical = FALSE) ------------- */
static double rlk_profile(const rp_linetable *lt, const double *L, const rp_column *c,
                          int k, int to_obs, double lambda,
                          double *phi_Q, double *phi_U, double *phi_V)
{
  const double *el = lt->elems + (long)((int) L[RL_ELEM]) * RE_NFIELD;
  double vtherm = 2.0*RP_KBOLTZMANN/(RP_AMU * el[RE_WEIGHT]);
  double vbroad = sqrt(vtherm*c->T[k] + c->vturb[k]*c->vturb[k]);
  double v, sv, GvdW, adamp, phi;

  v = (lambda/L[RL_LAMBDA0] - 1.0) * RP_CLIGHT/vbroad;
  if (c->moving) {
    if (to_obs) v += (c->muz * c->vel[k]) / vbroad;     /* vproject(), project.c:28-36 */
    else        v -= (c->muz * c->vel[k]) / vbroad;
  }
  sv = 1.0 / (RP_SQRTPI * vbroad);

  if (L[RL_GRAD]) {
    switch ((int) L[RL_VDWAALS]) {
    case RP_UNSOLD:  GvdW = L[RL_CROSS] * pow(c->T[k], 0.3); break;
    case RP_BARKLEM: GvdW = L[RL_CROSS] * pow(c->T[k], (1.0 - L[RL_ALPHA])/2.0); break;
    default:         GvdW = L[RL_GVDW]; break;
    }
    adamp = (L[RL_GRAD] + L[RL_GSTARK] * c->ne[k] +
             GvdW * (c->nHtot[k] - c->np[k])) *
      (L[RL_LAMBDA0] * RP_NM_TO_M) / (4.0*RP_PI * vbroad);
  } else {
    phi = (fabs(v) <= RP_MAX_GAUSS_DOPPLER) ? exp(-v*v) : 0.0;
    return phi * sv;
  }

  if (L[RL_POLARIZABLE]) {
    double sin2_gamma = 1.0 - c->cos_gamma[k]*c->cos_gamma[k];
    double vB = (RP_LARMOR * L[RL_LAMBDA0]) * c->B[k] / vbroad;
    double sign = to_obs ? 1.0 : -1.0;
    double phi_sm = 0.0, phi_pi = 0.0, phi_sp = 0.0, H, F, phi_sigma, phi_delta;
    int zoff = (int) L[RL_ZOFF], nc = (int) L[RL_NCOMP], nz;

    for (nz = 0; nz < nc; nz++) {
      H = rp_voigt_humlicek(adamp, v - lt->zshift[zoff+nz]*vB, &F);
      switch (lt->zq[zoff+nz]) {
      case -1: phi_sm += lt->zstrength[zoff+nz] * H; break;
      case  0: phi_pi += lt->zstrength[zoff+nz] * H; break;
      case  1: phi_sp += lt->zstrength[zoff+nz] * H; break;
      }
    }
    phi_sigma = phi_sp + phi_sm;
    phi_delta = 0.5*phi_pi - 0.25*phi_sigma;
    phi = (phi_delta*sin2_gamma + 0.5*phi_sigma) * sv;
    *phi_Q = sign * phi_delta * sin2_gamma * c->cos_2chi[k] * sv;
    *phi_U = phi_delta * sin2_gamma * c->sin_2chi[k] * sv;
    *phi_V = sign * 0.5*(phi_sp - phi_sm) * c->cos_gamma[k] * sv;
    return phi;
  }
  /* non-polarizable lines use VoigtArmstrong in the reference (kurucz.c:824);
     not restated here: every line of the pinned configurations is polarizable */
  return NAN;
}

/* ---- rlk_opacity, kurucz.c:511-725 (rlkscatter = FALSE, no model-atom veto:
        callers pass only lines whose element has no overlapping explicit
        AtomicLine, kurucz.c:616-631) ------------------------------------- */
int rp_rlk_opacity(const rp_linetable *lt, const rp_column *col, const double *elem_n,
                   double lambda, int to_obs, double *chi, double *eta)
{
  int N = col->Ndep, nl = lt->nline, flags = 0, n, k, Nwhite, Nblue, Nred;
  const double *LT = lt->lines;
  double dlamb_char = lambda * RP_Q_WING * (lt->vmicro_char / RP_CLIGHT);
  double hc = RP_HPLANCK * RP_CLIGHT, fourPI = 4.0 * RP_PI, hc_4PI = hc / fourPI;
  double *pf;

  if (nl == 0) return 0;
  if (lambda < LT[RL_LAMBDA0] - dlamb_char ||
      lambda > LT[(long)(nl-1)*RL_NFIELD + RL_LAMBDA0] + dlamb_char) return 0;

  /* rlk_locate (kurucz.c:453-507) with *low = 0 is a plain bisection */
  { int lo = 0, hi = nl;
    while (hi - lo > 1) {
      int mid = (hi + lo) >> 1;
      if (lambda >= LT[(long)mid*RL_NFIELD + RL_LAMBDA0]) lo = mid; else hi = mid;
    }
    Nwhite = lo; }
  Nblue = Nwhite;
  while (LT[(long)Nblue*RL_NFIELD + RL_LAMBDA0] + dlamb_char > lambda && Nblue > 0) Nblue--;
  Nred = Nwhite;
  while (LT[(long)Nred*RL_NFIELD + RL_LAMBDA0] - dlamb_char < lambda && Nred < nl-1) Nred++;

  if (Nred >= Nblue)
    for (k = 0; k < 4*N; k++) { chi[k] = 0.0; eta[k] = 0.0; }

  pf = malloc(N * sizeof(double));
  for (n = Nblue; n <= Nred; n++) {
    const double *L = LT + (long) n*RL_NFIELD;
    if (fabs(L[RL_LAMBDA0] - lambda) <= dlamb_char) {
      int ie = (int) L[RL_ELEM], st = (int) L[RL_STAGE];
      const double *el = lt->elems + (long) ie*RE_NFIELD;
      if (!(st < (int) el[RE_NSTAGE] - 1)) continue;
      {
        double hc_la      = (RP_HPLANCK * RP_CLIGHT) / (L[RL_LAMBDA0] * RP_NM_TO_M);
        double Bijhc_4PI  = hc_4PI * L[RL_BIJ] * L[RL_ISO_FRAC] * L[RL_HFS_FRAC] * L[RL_GI];
        double twohnu3_c2 = L[RL_AJI] / L[RL_BJI];
        const double *nst = elem_n + ((long) ie*RE_MAXSTAGE + st) * N;

        flags |= 1;
        if (L[RL_POLARIZABLE]) flags |= 2;
        rp_linear(lt->npf, lt->Tpf, lt->pf + (long)((int) el[RE_PFROW] + st)*lt->npf,
                  N, col->T, pf);
        for (k = 0; k < N; k++) {
          double phi_Q = 0, phi_U = 0, phi_V = 0;
          double phi = rlk_profile(lt, L, col, k, to_obs, lambda, &phi_Q, &phi_U, &phi_V);
          if (phi) {
            double kT    = 1.0 / (RP_KBOLTZMANN * col->T[k]);
            double ni_gi = nst[k] * exp(-L[RL_EI]*kT - pf[k]);
            double nj_gj = ni_gi * exp(-hc_la * kT);
            double chi_l = Bijhc_4PI * (ni_gi - nj_gj);
            double eta_l = Bijhc_4PI * twohnu3_c2 * nj_gj;
            chi[k] += chi_l * phi;
            eta[k] += eta_l * phi;
            if (L[RL_POLARIZABLE] && L[RL_GRAD]) {
              chi[N+k]   += chi_l * phi_Q;
              chi[2*N+k] += chi_l * phi_U;
              chi[3*N+k] += chi_l * phi_V;
              eta[N+k]   += eta_l * phi_Q;
              eta[2*N+k] += eta_l * phi_U;
              eta[3*N+k] += eta_l * phi_V;
            }
          }
        }
      }
    }
  }
  free(pf);
  return flags;
}

/* ---- bookkeeping helper for DESIGN.md / bench.py: histogram of the Humlicek region taken
        by every Voigt evaluation of rlk_opacity over a wavelength grid (hist[1..4]) and the
        number of (line, depth, wavelength) profile evaluations (hist[0]) ---------------- */
void rp_region_hist(const rp_linetable *lt, const rp_column *c, int Nlambda,
                    const double *lambda, int to_obs, long *hist)
{
  for (int i = 0; i < 5; i++) hist[i] = 0;
  for (int nl = 0; nl < Nlambda; nl++) {
    double dlamb_char = lambda[nl] * RP_Q_WING * (lt->vmicro_char / RP_CLIGHT);
    for (int n = 0; n < lt->nline; n++) {
      const double *L = lt->lines + (long) n*RL_NFIELD;
      if (fabs(L[RL_LAMBDA0] - lambda[nl]) > dlamb_char) continue;
      const double *el = lt->elems + (long)((int) L[RL_ELEM]) * RE_NFIELD;
      double vtherm = 2.0*RP_KBOLTZMANN/(RP_AMU * el[RE_WEIGHT]);
      for (int k = 0; k < c->Ndep; k++) {
        double vbroad = sqrt(vtherm*c->T[k] + c->vturb[k]*c->vturb[k]);
        double v = (lambda[nl]/L[RL_LAMBDA0] - 1.0) * RP_CLIGHT/vbroad, GvdW, adamp, vB;
        if (c->moving) v += (to_obs ? 1.0 : -1.0) * (c->muz * c->vel[k]) / vbroad;
        switch ((int) L[RL_VDWAALS]) {
        case RP_UNSOLD:  GvdW = L[RL_CROSS] * pow(c->T[k], 0.3); break;
        case RP_BARKLEM: GvdW = L[RL_CROSS] * pow(c->T[k], (1.0 - L[RL_ALPHA])/2.0); break;
        default:         GvdW = L[RL_GVDW]; break;
        }
        adamp = (L[RL_GRAD] + L[RL_GSTARK] * c->ne[k] + GvdW * (c->nHtot[k] - c->np[k])) *
          (L[RL_LAMBDA0] * RP_NM_TO_M) / (4.0*RP_PI * vbroad);
        vB = (RP_LARMOR * L[RL_LAMBDA0]) * c->B[k] / vbroad;
        hist[0]++;
        for (int nz = 0; nz < (int) L[RL_NCOMP]; nz++)
          hist[rp_humlicek_region(adamp, v - lt->zshift[(int) L[RL_ZOFF]+nz]*vB)]++;
      }
    }
  }
}

/* ---- VoigtArmstrong (K1/K2/K3), rh/voigt.c:126-243: unpolarised Voigt function used by
        Profile() for NO_STOKES active lines, by passive_bb and by non-polarizable Kurucz lines */
static const double arm_t[10] = {0.2453407083, 0.7374737285, 1.2340762153, 1.7385377121,
                                 2.2549740020, 2.7888060584, 3.3478545673, 3.9447640401,
                                 4.6036824495, 5.3874808900};
static const double arm_w[10] = {4.6224366960e-01, 2.8667550536e-01, 1.0901720602e-01,
                                 2.4810520887e-02, 3.2437733422e-03, 2.2833863601e-04,
                                 7.8025564785e-06, 1.0860693707e-07, 4.3993409922e-10,
                                 2.2293936455e-13};
static const double arm_c[34] = { 0.1999999999972224, -0.1840000000029998, 0.1558399999965025, -0.1216640000043988,
  0.0877081599940391, -0.0585141248086907, 0.0362157301623914, -0.0208497654398036, 0.0111960116346270,
  -0.56231896167109e-02, 0.26487634172265e-02, -0.11732670757704e-02, 0.4899519978088e-03, -0.1933630801528e-03,
  0.722877446788e-04, -0.256555124979e-04, 0.86620736841e-05, -0.27876379719e-05, 0.8566873627e-06,
  -0.2518433784e-06, 0.709360221e-07, -0.191732257e-07, 0.49801256e-08, -0.12447734e-08, 0.2997777e-09,
  -0.696450e-10, 0.156262e-10, -0.33897e-11, 0.7116e-12, -0.1447e-12, 0.285e-13, -0.55e-14, 0.10e-14, -0.2e-15 };

static double voigt_k1(double a, double v)              /* voigt.c:146-211 */
{
  int n;
  double a2 = a*a, v2 = v*v, u1, dn01, dn02, dn, v2i, funct, an, q, g, coef, bn01, bn02, bn = 0.0, v1;
  if ((v2 - a2) > 70.0) u1 = 0.0;
  else u1 = exp(a2 - v2) * cos(2.0*v*a);
  if (v > 5.0) {
    v2i = 1.0 / v2;
    dn01 = -v2i * (0.5 + v2i*(0.75 + v2i*(1.875 + v2i*(6.5625 +
           v2i*(29.53125 + v2i*(1162.4218 + v2i*1055.7421))))));
    dn02 = (1.0 - dn01) / (2.0 * v);
  } else {
    bn01 = bn02 = 0.0;
    v1   = v / 5.0;
    coef = 4.0 * v1*v1 - 2.0;
    for (n = 33; n >= 0; n--) { bn = coef*bn01 - bn02 + arm_c[n]; bn02 = bn01; bn01 = bn; }
    dn02 = (double) (v1*(bn - bn02));
    dn01 = 1.0 - 2.0*v*dn02;
  }
  funct = a*dn01;
  if (a > 1.0E-08) {
    q = 1.0; an = a;
    for (n = 2; n <= 50; n++) {
      dn = (v*dn01 + dn02) * (-2.0/n);
      dn02 = dn01; dn01 = dn;
      if (n % 2) {
        q = -q; an *= a2; g = dn * an; funct += q*g;
        if (fabs(g/funct) <= 1.0E-08) return (u1 - 1.12837917*funct);
      }
    }
  }
  return (u1 - 1.12837917*funct);
}
static double voigt_k2(double a, double v)              /* voigt.c:215-229 */
{
  double g = 0.0, r, s, a2 = a*a;
  for (int n = 0; n < 10; n++) {
    r = arm_t[n] - v; s = arm_t[n] + v;
    g += (4.0*arm_t[n]*arm_t[n] - 2.0) * (r*atan(r/a) + s*atan(s/a) -
          0.5*a*(log(a2 + r*r) + log(a2 + s*s))) * arm_w[n];
  }
  return g/RP_PI;
}
static double voigt_k3(double a, double v)              /* voigt.c:233-243 */
{
  double g = 0.0, a2 = a*a;
  for (int n = 0; n < 10; n++)
    g += (1.0/((v - arm_t[n])*(v - arm_t[n]) + a2) + 1.0/((v + arm_t[n])*(v + arm_t[n]) + a2)) * arm_w[n];
  return (a*g)/RP_PI;
}
int rp_armstrong_region(double a, double v)             /* voigt.c:126-137 */
{
  if (v < 0.0) v = -v;
  if ((a < 1.0 && v < 4.0) || (a < 1.8/(v + 1.0))) return 1;
  if (a < 2.5 && v < 4.0) return 2;
  return 3;
}
double rp_voigt_armstrong(double a, double v)
{
  if (v < 0.0) v = -v;
  switch (rp_armstrong_region(a, v)) {
  case 1: return voigt_k1(a, v);
  case 2: return voigt_k2(a, v);
  default: return voigt_k3(a, v);
  }
}
