/* oracle/port/rhport_solvers.c -- TEST INFRASTRUCTURE ONLY (see rhport.h).
 *
 * CPU restatement of the reference's 1-D formal solvers:
 *   Piece_Stokes_Bezier3_1D   rh/rhf1d/bezier_1D.c:52-300
 *   Piecewise_Bezier3_1D      rh/rhf1d/bezier_1D.c:306-541 (without the log gf RF)
 *   StokesK                   rh/stokesopac.c:28-87 (magneto_optical = FALSE)
 *   cent_deriv, m4m, m4v, MatInv (scalar + SSE), Bezier3_coeffs
 *                             rh/bezier_aux.c:34-360
 *   w3                        rh/w3.c:43-63
 *   Planck                    rh/planck.c:38-66
 * Compile with -O2 -ffp-contract=off: the reference x86-64 build has no FMA.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#if defined(__x86_64__)
#include <x86intrin.h>
#endif
#include "rhport.h"

#define RP_MAX(x, y) (((x) > (y) ? (x) : (y)))

double rp_planck(double T, double lambda)              /* planck.c:38-66 */
{
  double hc_kla     = (RP_HPLANCK * RP_CLIGHT) / (RP_KBOLTZMANN * RP_NM_TO_M * lambda);
  double l = RP_NM_TO_M * lambda;
  double twohnu3_c2 = (2.0*RP_HPLANCK*RP_CLIGHT) / (l*l*l);
  double hc_Tkla = hc_kla / T;
  if (hc_Tkla <= 150.0) return twohnu3_c2 / (exp(hc_Tkla) - 1.0);
  return 0.0;
}

void rp_w3(double dtau, double *w)                     /* w3.c:43-63 */
{
  double expdt, delta;
  if (dtau < 5.0E-4) {
    w[0]   = dtau*(1.0 - 0.5*dtau);
    delta  = dtau*dtau;
    w[1]   = delta*(0.5 - dtau/3.0);
    delta *= dtau;
    w[2]   = delta*(1.0/3.0 - 0.25*dtau);
  } else if (dtau > 50.0) {
    w[1] = w[0] = 1.0;
    w[2] = 2.0;
  } else {
    expdt = exp(-dtau);
    w[0]  = 1.0 - expdt;
    w[1]  = w[0] - dtau*expdt;
    w[2]  = 2.0*w[1] - dtau*dtau * expdt;
  }
}

double rp_cent_deriv(double dsup, double dsdn, double chiup, double chic, double chidn)
{                                                       /* bezier_aux.c:34-51 */
  double fim1 = (chic - chiup) / dsup, fi = (chidn - chic) / dsdn, alpha;
  if (fim1*fi > 0) {
    alpha = 0.333333333333333333333333 * (1.0 + dsdn / (dsdn+dsup));
    return (fim1 * fi) / ((1.0 - alpha) * fim1 + alpha*fi);
  }
  return 0.0;
}

void rp_bezier3_coeffs(double dt, double *alpha, double *beta, double *gamma,
                       double *theta, double *eps)      /* bezier_aux.c:333-359 */
{
  double dt2 = dt*dt, dt3 = dt2 * dt, dt4;
  if (dt >= 5.e-2) {
    *eps = exp(-dt);
    *alpha = (-6.0 + 6.0 * dt - 3.0 * dt2 + dt3 + 6.0 * eps[0]) / dt3;
    dt3 = 1.0/dt3;
    *beta  = (6.0 + (-6.0 - dt * (6.0 + dt * (3.0 + dt))) * eps[0]) * dt3;
    *gamma = 3.0 * (6.0 + (-4.0 + dt)*dt - 2.0 * (3.0 + dt) * eps[0]) * dt3;
    *theta = 3.0 * (eps[0] * (6.0 + dt2 + 4.0 * dt) + 2.0 * dt - 6.0) * dt3;
  } else {
    dt4 = dt2*dt2;
    *eps = 1.0 - dt + 0.5 * dt2 - dt3 / 6.0 + dt4 / 24.0;
    *alpha = 0.25 * dt - 0.05 * dt2 + dt3 / 120.0 - dt4 / 840.0;
    *beta  = 0.25 * dt - 0.20 * dt2 + dt3 / 12.0  - dt4 / 42.0;
    *gamma = 0.25 * dt - 0.10 * dt2 + dt3 * 0.025 - dt4 / 210.0;
    *theta = 0.25 * dt - 0.15 * dt2 + dt3 * 0.05  - dt4 / 84.0;
  }
}

/* scalar Cramer inverse in float, bezier_aux.c:235-328 (cofactor expansion on
   the transposed matrix; every operation rounds to float) */
void rp_matinv_scalar(float *mat)
{
  float t[12], s[16], d[16], det;
  int i, j;
  for (i = 0; i < 4; i++) {
    s[i] = mat[i*4]; s[i+4] = mat[i*4+1]; s[i+8] = mat[i*4+2]; s[i+12] = mat[i*4+3];
  }
  t[0] = s[10]*s[15]; t[1] = s[11]*s[14]; t[2] = s[9]*s[15];  t[3] = s[11]*s[13];
  t[4] = s[9]*s[14];  t[5] = s[10]*s[13]; t[6] = s[8]*s[15];  t[7] = s[11]*s[12];
  t[8] = s[8]*s[14];  t[9] = s[10]*s[12]; t[10] = s[8]*s[13]; t[11] = s[9]*s[12];

  d[0]  = t[0]*s[5] + t[3]*s[6] + t[4]*s[7];   d[0] -= t[1]*s[5] + t[2]*s[6] + t[5]*s[7];
  d[1]  = t[1]*s[4] + t[6]*s[6] + t[9]*s[7];   d[1] -= t[0]*s[4] + t[7]*s[6] + t[8]*s[7];
  d[2]  = t[2]*s[4] + t[7]*s[5] + t[10]*s[7];  d[2] -= t[3]*s[4] + t[6]*s[5] + t[11]*s[7];
  d[3]  = t[5]*s[4] + t[8]*s[5] + t[11]*s[6];  d[3] -= t[4]*s[4] + t[9]*s[5] + t[10]*s[6];
  d[4]  = t[1]*s[1] + t[2]*s[2] + t[5]*s[3];   d[4] -= t[0]*s[1] + t[3]*s[2] + t[4]*s[3];
  d[5]  = t[0]*s[0] + t[7]*s[2] + t[8]*s[3];   d[5] -= t[1]*s[0] + t[6]*s[2] + t[9]*s[3];
  d[6]  = t[3]*s[0] + t[6]*s[1] + t[11]*s[3];  d[6] -= t[2]*s[0] + t[7]*s[1] + t[10]*s[3];
  d[7]  = t[4]*s[0] + t[9]*s[1] + t[10]*s[2];  d[7] -= t[5]*s[0] + t[8]*s[1] + t[11]*s[2];

  t[0] = s[2]*s[7];  t[1] = s[3]*s[6];  t[2] = s[1]*s[7];  t[3] = s[3]*s[5];
  t[4] = s[1]*s[6];  t[5] = s[2]*s[5];  t[6] = s[0]*s[7];  t[7] = s[3]*s[4];
  t[8] = s[0]*s[6];  t[9] = s[2]*s[4];  t[10] = s[0]*s[5]; t[11] = s[1]*s[4];

  d[8]  = t[0]*s[13] + t[3]*s[14] + t[4]*s[15];   d[8]  -= t[1]*s[13] + t[2]*s[14] + t[5]*s[15];
  d[9]  = t[1]*s[12] + t[6]*s[14] + t[9]*s[15];   d[9]  -= t[0]*s[12] + t[7]*s[14] + t[8]*s[15];
  d[10] = t[2]*s[12] + t[7]*s[13] + t[10]*s[15];  d[10] -= t[3]*s[12] + t[6]*s[13] + t[11]*s[15];
  d[11] = t[5]*s[12] + t[8]*s[13] + t[11]*s[14];  d[11] -= t[4]*s[12] + t[9]*s[13] + t[10]*s[14];
  d[12] = t[2]*s[10] + t[5]*s[11] + t[1]*s[9];    d[12] -= t[4]*s[11] + t[0]*s[9] + t[3]*s[10];
  d[13] = t[8]*s[11] + t[0]*s[8] + t[7]*s[10];    d[13] -= t[6]*s[10] + t[9]*s[11] + t[1]*s[8];
  d[14] = t[6]*s[9] + t[11]*s[11] + t[3]*s[8];    d[14] -= t[10]*s[11] + t[2]*s[8] + t[7]*s[9];
  d[15] = t[10]*s[10] + t[4]*s[8] + t[9]*s[9];    d[15] -= t[8]*s[9] + t[11]*s[10] + t[5]*s[8];

  det = s[0]*d[0] + s[1]*d[1] + s[2]*d[2] + s[3]*d[3];
  det = 1.0/det;                    /* double division, rounded to float */
  for (j = 0; j < 16; j++) d[j] *= det;
  memcpy(mat, d, 16*sizeof(float));
}

/* SSE Cramer inverse, bezier_aux.c:133-232 (Intel AP-928 sequence incl. rcpss +
   one Newton step).  Same intrinsic sequence is required for bit parity with
   the -DSIMDON build on the *same* CPU (rcpss is micro-architecture defined). */
void rp_matinv_simd(float *mat)
{
#if defined(__x86_64__)
  __m128 m0, m1, m2, m3, r0, r1, r2, r3, det, t;
  t  = _mm_setzero_ps(); r1 = _mm_setzero_ps(); r3 = _mm_setzero_ps();
  t  = _mm_loadh_pi(_mm_loadl_pi(t, (__m64*)(mat)), (__m64*)(mat+4));
  r1 = _mm_loadh_pi(_mm_loadl_pi(r1, (__m64*)(mat+8)), (__m64*)(mat+12));
  r0 = _mm_shuffle_ps(t, r1, 0x88);
  r1 = _mm_shuffle_ps(r1, t, 0xDD);
  t  = _mm_loadh_pi(_mm_loadl_pi(t, (__m64*)(mat+2)), (__m64*)(mat+6));
  r3 = _mm_loadh_pi(_mm_loadl_pi(r3, (__m64*)(mat+10)), (__m64*)(mat+14));
  r2 = _mm_shuffle_ps(t, r3, 0x88);
  r3 = _mm_shuffle_ps(r3, t, 0xDD);

  t  = _mm_mul_ps(r2, r3);           t = _mm_shuffle_ps(t, t, 0xB1);
  m0 = _mm_mul_ps(r1, t);            m1 = _mm_mul_ps(r0, t);
  t  = _mm_shuffle_ps(t, t, 0x4E);
  m0 = _mm_sub_ps(_mm_mul_ps(r1, t), m0);
  m1 = _mm_sub_ps(_mm_mul_ps(r0, t), m1);
  m1 = _mm_shuffle_ps(m1, m1, 0x4E);

  t  = _mm_mul_ps(r1, r2);           t = _mm_shuffle_ps(t, t, 0xB1);
  m0 = _mm_add_ps(_mm_mul_ps(r3, t), m0);
  m3 = _mm_mul_ps(r0, t);
  t  = _mm_shuffle_ps(t, t, 0x4E);
  m0 = _mm_sub_ps(m0, _mm_mul_ps(r3, t));
  m3 = _mm_sub_ps(_mm_mul_ps(r0, t), m3);
  m3 = _mm_shuffle_ps(m3, m3, 0x4E);

  t  = _mm_mul_ps(_mm_shuffle_ps(r1, r1, 0x4E), r3);
  t  = _mm_shuffle_ps(t, t, 0xB1);
  r2 = _mm_shuffle_ps(r2, r2, 0x4E);
  m0 = _mm_add_ps(_mm_mul_ps(r2, t), m0);
  m2 = _mm_mul_ps(r0, t);
  t  = _mm_shuffle_ps(t, t, 0x4E);
  m0 = _mm_sub_ps(m0, _mm_mul_ps(r2, t));
  m2 = _mm_sub_ps(_mm_mul_ps(r0, t), m2);
  m2 = _mm_shuffle_ps(m2, m2, 0x4E);

  t  = _mm_mul_ps(r0, r1);           t = _mm_shuffle_ps(t, t, 0xB1);
  m2 = _mm_add_ps(_mm_mul_ps(r3, t), m2);
  m3 = _mm_sub_ps(_mm_mul_ps(r2, t), m3);
  t  = _mm_shuffle_ps(t, t, 0x4E);
  m2 = _mm_sub_ps(_mm_mul_ps(r3, t), m2);
  m3 = _mm_sub_ps(m3, _mm_mul_ps(r2, t));

  t  = _mm_mul_ps(r0, r3);           t = _mm_shuffle_ps(t, t, 0xB1);
  m1 = _mm_sub_ps(m1, _mm_mul_ps(r2, t));
  m2 = _mm_add_ps(_mm_mul_ps(r1, t), m2);
  t  = _mm_shuffle_ps(t, t, 0x4E);
  m1 = _mm_add_ps(_mm_mul_ps(r2, t), m1);
  m2 = _mm_sub_ps(m2, _mm_mul_ps(r1, t));

  t  = _mm_mul_ps(r0, r2);           t = _mm_shuffle_ps(t, t, 0xB1);
  m1 = _mm_add_ps(_mm_mul_ps(r3, t), m1);
  m3 = _mm_sub_ps(m3, _mm_mul_ps(r1, t));
  t  = _mm_shuffle_ps(t, t, 0x4E);
  m1 = _mm_sub_ps(m1, _mm_mul_ps(r3, t));
  m3 = _mm_add_ps(_mm_mul_ps(r1, t), m3);

  det = _mm_mul_ps(r0, m0);
  det = _mm_add_ps(_mm_shuffle_ps(det, det, 0x4E), det);
  det = _mm_add_ss(_mm_shuffle_ps(det, det, 0xB1), det);
  t   = _mm_rcp_ss(det);
  det = _mm_sub_ss(_mm_add_ss(t, t), _mm_mul_ss(det, _mm_mul_ss(t, t)));
  det = _mm_shuffle_ps(det, det, 0x00);
  m0 = _mm_mul_ps(det, m0);
  _mm_storel_pi((__m64*)(mat), m0);     _mm_storeh_pi((__m64*)(mat+2), m0);
  m1 = _mm_mul_ps(det, m1);
  _mm_storel_pi((__m64*)(mat+4), m1);   _mm_storeh_pi((__m64*)(mat+6), m1);
  m2 = _mm_mul_ps(det, m2);
  _mm_storel_pi((__m64*)(mat+8), m2);   _mm_storeh_pi((__m64*)(mat+10), m2);
  m3 = _mm_mul_ps(det, m3);
  _mm_storel_pi((__m64*)(mat+12), m3);  _mm_storeh_pi((__m64*)(mat+14), m3);
#else
  rp_matinv_scalar(mat);
#endif
}

static void m4m(double a[4][4], double b[4][4], double c[4][4])   /* bezier_aux.c:88-99 */
{
  int i, j, k;
  memset(&c[0][0], 0, sizeof(double)*16);
  for (j = 0; j < 4; j++)
    for (i = 0; i < 4; i++)
      for (k = 0; k < 4; k++)
        c[j][i] += a[k][i]*b[j][k];
}

static void m4v(float a[4][4], double b[4], double c[4])          /* bezier_aux.c:109-116 */
{
  int k, i;
  memset(&c[0], 0, sizeof(double)*4);
  for (i = 0; i < 4; i++)
    for (k = 0; k < 4; k++)
      c[i] += ((double) a[i][k]) * b[k];
}

/* StokesK, stokesopac.c:28-87, with chiQUV = as->chi[QUV] + as->chi_c[QUV]
   already summed by the caller and magneto_optical = FALSE */
static void stokesK(const double *chiQUV, int N, int k, double chi_I, double K[4][4])
{
  int i, j;
  for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) K[j][i] = 0.0;
  K[0][1] = chiQUV[k]; K[0][2] = chiQUV[N+k]; K[0][3] = chiQUV[2*N+k];
  for (j = 0; j < 3; j++)
    for (i = j+1; i < 4; i++) { K[j][i] /= chi_I; K[i][j] = K[j][i]; }
}

static const double ident[4][4] = {{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,1}};

void rp_stokes_bezier3(int Ndep, const double *z, double muz, int to_obs,
                       const double *chi, const double *S, const double *chiQUV,
                       const double *T, double lambda, int bc_top, int bc_bottom,
                       int matinv_simd, double *I, double *Psi)
{
  int k, n, m, i, j, k_start, k_end, dk, N = Ndep;
  double dtau_uw, dtau_dw = 0.0, c1, c2, w[3], dsdn2, dchi_dn, I_upw[4];
  double dchi_up, dchi_c, dt03, dsup, dsdn, dt, eps = 0, alpha = 0, beta = 0, gamma = 0, theta = 0;
  double Ku[4][4], K0[4][4], Kd[4][4], dKu[4][4], dK0[4][4];
  double Su[4], S0[4], Sd[4], dSu[4], dS0[4];
  double A[4][4], Ma[4][4], Mb[4][4], Mc[4][4], V0[4], V1[4];
  double imu = 1.0 / muz;
  float Md[4][4];
  void (*matinv)(float *) = matinv_simd ? rp_matinv_simd : rp_matinv_scalar;

  if (to_obs) { dk = -1; k_start = Ndep-1; k_end = 0; }
  else        { dk =  1; k_start = 0;      k_end = Ndep-1; }
  dtau_uw = 0.5 * imu * (chi[k_start] + chi[k_start+dk]) * fabs(z[k_start] - z[k_start+dk]);

  for (n = 0; n < 4; n++) I_upw[n] = 0.0;            /* bezier_1D.c:93-126 */
  if (to_obs) {
    if (bc_bottom == RP_THERMALIZED) {
      double B0 = rp_planck(T[Ndep-2], lambda), B1 = rp_planck(T[Ndep-1], lambda);
      I_upw[0] = B1 - (B0 - B1) / dtau_uw;
    }
  }
  (void) bc_top;                                      /* ZERO */
  for (n = 0; n < 4; n++) I[n*N + k_start] = I_upw[n];
  if (Psi) Psi[k_start] = 0.0;

  k = k_start+dk;
  dsup = fabs(z[k] - z[k-dk]) * imu;
  dsdn = fabs(z[k+dk] - z[k]) * imu;
  dchi_up = (chi[k] - chi[k-dk])/dsup;
  dchi_c = rp_cent_deriv(dsup, dsdn, chi[k-dk], chi[k], chi[k+dk]);

  c2 = RP_MAX(chi[k]    - (dsup/3.0) * dchi_c,  0.0);
  c1 = RP_MAX(chi[k-dk] + (dsup/3.0) * dchi_up, 0.0);
  dtau_uw = 0.25 * dsup * (chi[k] + chi[k-dk] + c1 + c2);

  stokesK(chiQUV, N, k_start,    chi[k_start],    Ku);
  stokesK(chiQUV, N, k_start+dk, chi[k_start+dk], K0);
  for (n = 0; n < 4; n++) { Su[n] = S[n*N + k_start]; S0[n] = S[n*N + k_start+dk]; }

  for (n = 0; n < 4; n++) {
    dSu[n] = (S0[n] - Su[n]) / dtau_uw;
    for (m = 0; m < 4; m++) dKu[n][m] = (K0[n][m] - Ku[n][m]) / dtau_uw;
  }

  for (k = k_start+dk; k != k_end; k += dk) {
    dsdn = fabs(z[k+dk] - z[k]) * imu;
    if (abs(k - k_end) > 1) {
      dsdn2 = fabs(z[k+2*dk] - z[k+dk]) * imu;
      dchi_dn = rp_cent_deriv(dsdn, dsdn2, chi[k], chi[k+dk], chi[k+2*dk]);
    } else
      dchi_dn = (chi[k+dk] - chi[k])/dsdn;

    c2 = RP_MAX(chi[k]    + (dsdn/3.0) * dchi_c , 0.0);
    c1 = RP_MAX(chi[k+dk] - (dsdn/3.0) * dchi_dn, 0.0);
    dtau_dw = 0.25 * dsdn * (chi[k] + chi[k+dk] + c1 + c2);
    dt = dtau_uw; dt03 = dt / 3.0;

    rp_bezier3_coeffs(dt, &alpha, &beta, &gamma, &theta, &eps);
    if (Psi) Psi[k] = alpha + gamma;

    stokesK(chiQUV, N, k+dk, chi[k+dk], Kd);
    for (n = 0; n < 4; n++) Sd[n] = S[n*N + k+dk];

    for (j = 0; j < 4; j++)
      for (i = 0; i < 4; i++)
        dK0[j][i] = rp_cent_deriv(dtau_uw, dtau_dw, Ku[j][i], K0[j][i], Kd[j][i]);
    for (i = 0; i < 4; i++)
      dS0[i] = rp_cent_deriv(dtau_uw, dtau_dw, Su[i], S0[i], Sd[i]);

    m4m(Ku, Ku, Ma);
    m4m(K0, K0, A);

    for (j = 0; j < 4; j++) {
      for (i = 0; i < 4; i++) {
        Md[j][i] = ident[j][i] + alpha * K0[j][i] - gamma *
          -(dt03 * (A[j][i] + dK0[j][i] + K0[j][i]) + K0[j][i]);
        Ma[j][i] = eps * ident[j][i] - beta * Ku[j][i] + theta *
          (dt03 * (Ma[j][i] + dKu[j][i] + Ku[j][i]) - Ku[j][i]);
        Mb[j][i] = beta * ident[j][i] + theta * (ident[j][i] - dt03 * Ku[j][i]);
        Mc[j][i] = alpha* ident[j][i] + gamma * (ident[j][i] + dt03 * K0[j][i]);
      }
    }
    memset(V0, 0, 4*sizeof(double));
    for (i = 0; i < 4; i++) {
      for (j = 0; j < 4; j++)
        V0[i] += Ma[i][j] * I[j*N + k-dk] + Mb[i][j] * Su[j] + Mc[i][j] * S0[j];
      V0[i] += dt03 * (gamma * dS0[i] - theta * dSu[i]);
    }
    matinv(Md[0]);
    m4v(Md, V0, V1);
    for (i = 0; i < 4; i++) I[i*N + k] = V1[i];

    memcpy(Su, S0, sizeof(Su)); memcpy(S0, Sd, sizeof(S0)); memcpy(dSu, dS0, sizeof(dSu));
    memcpy(Ku, K0, sizeof(Ku)); memcpy(K0, Kd, sizeof(K0)); memcpy(dKu, dK0, sizeof(dKu));
    dtau_uw = dtau_dw; dsup = dsdn; dchi_up = dchi_c; dchi_c = dchi_dn;
  }

  k = k_end;                                          /* bezier_1D.c:268-299 */
  dtau_uw = 0.5*imu * (chi[k] + chi[k-dk]) * fabs(z[k] - z[k-dk]);
  rp_w3(dtau_uw, w);
  for (n = 0; n < 4; n++) V0[n] = w[0]*S[n*N + k] + w[1] * -dSu[n];
  if (Psi) Psi[k] = w[0] - w[1] / dtau_uw;
  for (n = 0; n < 4; n++) {
    for (m = 0; m < 4; m++) {
      A[n][m]  = -w[1]/dtau_uw * Ku[n][m];
      Md[n][m] = (w[0] - w[1]/dtau_uw) * K0[n][m];
    }
    A[n][n]  = 1.0 - w[0];
    Md[n][n] = 1.0;
  }
  for (n = 0; n < 4; n++)
    for (m = 0; m < 4; m++)
      V0[n] += A[n][m] * I[m*N + k-dk];
  matinv(Md[0]);
  m4v(Md, V0, V1);
  for (n = 0; n < 4; n++) I[n*N + k] = V1[n];
}

void rp_bezier3_scalar(int Ndep, const double *z, double muz, int to_obs,
                       const double *chi, const double *S, const double *T, double lambda,
                       int bc_top, int bc_bottom, double *I, double *Psi)
{
  rp_bezier3_scalar_rf(Ndep, z, muz, to_obs, chi, S, T, lambda, bc_top, bc_bottom, I, Psi, 0, NULL, NULL, NULL);
}

/* With npar > 0 (get_atomic_rfs, up-ray only): dchi, deta [Ndep][npar] = spectrum.dchi_c_lam / deta_c_lam
   [nspect]; dI [Ndep][npar] out.  I must enter holding the preceding down-ray solution: the reference
   reads the not yet overwritten I[k], I[k+dk] (bezier_1D.c:424, 483). */
void rp_bezier3_scalar_rf(int Ndep, const double *z, double muz, int to_obs,
                          const double *chi, const double *S, const double *T, double lambda,
                          int bc_top, int bc_bottom, double *I, double *Psi,
                          int npar, const double *dchi, const double *deta, double *dI)
{                                                     /* bezier_1D.c:306-541 */
  int k, k_start, k_end, dk, idp;
  double Zk, Zkm1, Zkp1, dZk[RP_MAXPAR], dZup[RP_MAXPAR], dI_upw[RP_MAXPAR];
  double dtau_uw, dtau_dw = 0.0, dS_uw, I_upw = 0.0, c1, c2, w[3], zmu = 1.0 / muz;
  double dsup, dsdn, dt03, eps = 0, alpha = 0, beta = 0, gamma = 0, theta = 0;
  double dS_up, dS_c = 0.0, dchi_up, dchi_c, dchi_dn = 0.0, dsdn2;

  if (to_obs) { dk = -1; k_start = Ndep-1; k_end = 0; }
  else        { dk =  1; k_start = 0;      k_end = Ndep-1; }
  dtau_uw = 0.5 * zmu * (chi[k_start] + chi[k_start+dk]) * fabs(z[k_start] - z[k_start+dk]);

  if (to_obs) {
    if (bc_bottom == RP_THERMALIZED) {
      double B0 = rp_planck(T[Ndep-2], lambda), B1 = rp_planck(T[Ndep-1], lambda);
      I_upw = B1 - (B0 - B1) / dtau_uw;
    }
  } else {
    if (bc_top == RP_THERMALIZED) {
      double B0 = rp_planck(T[0], lambda), B1 = rp_planck(T[1], lambda);
      I_upw = B0 - (B1 - B0) / dtau_uw;
    }
  }
  I[k_start] = I_upw;
  if (Psi) Psi[k_start] = 0.0;

  k = k_start+dk;
  dsup = fabs(z[k] - z[k-dk]) * zmu;
  dsdn = fabs(z[k+dk] - z[k]) * zmu;
  dchi_up = (chi[k] - chi[k-dk]) / dsup;
  dchi_c = rp_cent_deriv(dsup, dsdn, chi[k-dk], chi[k], chi[k+dk]);
  c1 = RP_MAX(chi[k]    - (dsup/3.0) * dchi_c,  0.0);
  c2 = RP_MAX(chi[k-dk] + (dsup/3.0) * dchi_up, 0.0);
  dtau_uw = dsup * (chi[k] + chi[k-dk] + c1 + c2) * 0.25;
  dS_up = (S[k] - S[k-dk]) / dtau_uw;

  if (!to_obs || npar > RP_MAXPAR) npar = 0;
  for (idp = 0; idp < npar; idp++) {                  /* bezier_1D.c:416-428 */
    dI[k_start*npar + idp] = 0.0;
    dI_upw[idp] = 0.0;
    Zk = -dchi[k*npar + idp]/chi[k] * I[k] + deta[k*npar + idp]/chi[k];
    Zkm1 = -dchi[(k-dk)*npar + idp]/chi[k-dk] * I[k-dk] + deta[(k-dk)*npar + idp]/chi[k-dk];
    dZup[idp] = (Zk - Zkm1) / dtau_uw;
  }

  for (k = k_start+dk; k != k_end+dk; k += dk) {
    if (k != k_end) {
      dsdn = fabs(z[k+dk] - z[k]) * zmu;
      if (abs(k - k_end) > 1) {
        dsdn2 = fabs(z[k+2*dk] - z[k+dk]) * zmu;
        dchi_dn = rp_cent_deriv(dsdn, dsdn2, chi[k], chi[k+dk], chi[k+2*dk]);
      } else
        dchi_dn = (chi[k+dk]-chi[k])/dsdn;
      c1 = RP_MAX(chi[k]    + (dsdn/3.0) * dchi_c,  0.0);
      c2 = RP_MAX(chi[k+dk] - (dsdn/3.0) * dchi_dn, 0.0);
      dtau_dw = dsdn * (chi[k] + chi[k+dk] + c1 + c2) * 0.25;
      dt03    = dtau_uw / 3.0;
      rp_bezier3_coeffs(dtau_uw, &alpha, &beta, &gamma, &theta, &eps);
      dS_c = rp_cent_deriv(dtau_uw, dtau_dw, S[k-dk], S[k], S[k+dk]);
      c1 = RP_MAX(S[k]    - dt03 * dS_c , 0.0);
      c2 = RP_MAX(S[k-dk] + dt03 * dS_up, 0.0);
      I[k] = I_upw*eps + alpha*S[k] + beta*S[k-dk] + gamma * c1 + theta * c2;
      for (idp = 0; idp < npar; idp++) {              /* bezier_1D.c:477-490 */
        Zk = -dchi[k*npar + idp] * I[k] + deta[k*npar + idp];
        Zk /= chi[k];
        Zkm1 = -dchi[(k-dk)*npar + idp] * I[k-dk] + deta[(k-dk)*npar + idp];
        Zkm1 /= chi[k-dk];
        Zkp1 = -dchi[(k+dk)*npar + idp] * I[k+dk] + deta[(k+dk)*npar + idp];
        Zkp1 /= chi[k+dk];
        dZk[idp] = rp_cent_deriv(dtau_uw, dtau_dw, Zkm1, Zk, Zkp1);
        c1 = RP_MAX(Zk - dt03 * dZk[idp], 0.0);
        c2 = RP_MAX(Zkm1 + dt03 * dZup[idp], 0.0);
        dI[k*npar + idp] = dI_upw[idp]*eps + alpha*Zk + beta*Zkm1 + gamma*c1 + theta*c2;
      }
      if (Psi) Psi[k] = alpha + gamma;
    } else {
      dtau_uw = 0.5 * zmu * (chi[k] + chi[k-dk]) * fabs(z[k] - z[k-dk]);
      dS_uw = -(S[k] - S[k-dk]) / dtau_uw;
      rp_w3(dtau_uw, w);
      I[k] = (1.0 - w[0])*I_upw + w[0]*S[k] + w[1]*dS_uw;
      for (idp = 0; idp < npar; idp++) {              /* bezier_1D.c:509-516 */
        Zk = dchi[k*npar + idp]/chi[k] * I[k] - deta[k*npar + idp]/chi[k];
        Zkm1 = dchi[(k-dk)*npar + idp]/chi[k-dk] * I[k-dk] - deta[(k-dk)*npar + idp]/chi[k-dk];
        dZk[idp] = -(Zk - Zkm1) / dtau_uw;
        dI[k*npar + idp] = (1.0 - w[0])*dI_upw[idp] + w[0]*Zk + w[1]*dZk[idp];
      }
      if (Psi) Psi[k] = w[0] - w[1] / dtau_uw;
    }
    I_upw = I[k];
    dsup = dsdn; dchi_up = dchi_c; dchi_c = dchi_dn; dtau_uw = dtau_dw; dS_up = dS_c;
    for (idp = 0; idp < npar; idp++) { dI_upw[idp] = dI[k*npar + idp]; dZup[idp] = dZk[idp]; }
  }
}

/* Feautrier, rh/rhf1d/feautrier.c:56-202, F_order = STANDARD (the only order Formal() requests,
   formal.c:299).  Returns the emergent intensity; fills P and (if non-NULL) Psi. */
double rp_feautrier(int Ndep, const double *z, double muz, const double *chi, const double *S,
                    const double *T, double lambda, int bc_top, int bc_bottom, double *P, double *Psi)
{
  int k;
  double r0 = 0.0, h0 = 0.0, rN = 0.0, hN = 0.0, f0, fN, zmu = 0.5 / muz, dtau_mid, Iplus;
  double *dtau = malloc(Ndep*sizeof(double)), *abc = malloc(Ndep*sizeof(double)),
         *A1 = malloc(Ndep*sizeof(double)), *C1 = malloc(Ndep*sizeof(double)),
         *F = malloc(Ndep*sizeof(double)), *G = malloc(Ndep*sizeof(double)),
         *ztmp = malloc(Ndep*sizeof(double)), *Stmp = malloc(Ndep*sizeof(double));

  for (k = 0; k < Ndep-1; k++) dtau[k] = zmu * (chi[k] + chi[k+1]) * (z[k] - z[k+1]);

  if (bc_top == RP_THERMALIZED) {
    double B0 = rp_planck(T[0], lambda), B1 = rp_planck(T[1], lambda);
    h0 = B0 - (B1 - B0) / dtau[0];
  }
  f0      = (1.0 - r0) / (1.0 + r0);
  abc[0]  = 1.0 + 2.0*f0 / dtau[0];
  C1[0]   = 2.0 / (dtau[0]*dtau[0]);
  Stmp[0] = S[0] + 2.0*h0 / ((1.0 + r0)*dtau[0]);

  if (bc_bottom == RP_THERMALIZED) {
    double B0 = rp_planck(T[Ndep-2], lambda), B1 = rp_planck(T[Ndep-1], lambda);
    hN = B1 - (B0 - B1) / dtau[Ndep-2];
  }
  fN           = (1.0 - rN) / (1.0 + rN);
  abc[Ndep-1]  = 1.0 + 2.0*fN / dtau[Ndep-2];
  A1[Ndep-1]   = 2.0 / (dtau[Ndep-2]*dtau[Ndep-2]);
  Stmp[Ndep-1] = S[Ndep-1] + 2.0*hN / ((1.0 + rN)*dtau[Ndep-2]);

  for (k = 1; k < Ndep-1; k++) {
    dtau_mid = 0.5*(dtau[k] + dtau[k-1]);
    A1[k]   = 1.0 / (dtau_mid * dtau[k-1]);
    C1[k]   = 1.0 / (dtau_mid * dtau[k]);
    abc[k]  = 1.0;
    Stmp[k] = S[k];
  }
  F[0]    = abc[0] / C1[0];
  ztmp[0] = Stmp[0] / (abc[0] + C1[0]);
  for (k = 1; k < Ndep-1; k++) {
    F[k]    = (abc[k] + A1[k]*F[k-1]/(1.0 + F[k-1])) / C1[k];
    ztmp[k] = (Stmp[k] + A1[k]*ztmp[k-1]) / (C1[k] * (1.0 + F[k]));
  }
  P[Ndep-1] = (Stmp[Ndep-1]+ A1[Ndep-1]*ztmp[Ndep-2]) /
    (abc[Ndep-1] + A1[Ndep-1]*(F[Ndep-2] / (1.0 + F[Ndep-2])));
  for (k = Ndep-2; k >= 0; k--) P[k] = P[k+1] / (1.0 + F[k]) + ztmp[k];

  if (Psi) {
    G[Ndep-1] = abc[Ndep-1] / A1[Ndep-1];
    for (k = Ndep-2; k >= 1; k--) G[k] = (abc[k] + C1[k]*G[k+1]/(1.0 + G[k+1])) / A1[k];
    Psi[0] = 1.0 / (abc[0] + C1[0]*G[1]/(1.0 + G[1]));
    for (k = 1; k < Ndep-1; k++)
      Psi[k] = 1.0 / (abc[k] + A1[k]*F[k-1]/(1.0 + F[k-1]) + C1[k]*G[k+1]/(1.0 + G[k+1]));
    Psi[Ndep-1] = 1.0 / (abc[Ndep-1] + A1[Ndep-1]*F[Ndep-2]/(1.0 + F[Ndep-2]));
  }
  Iplus = (1.0 + f0)*P[0] - h0/(1.0 + r0);
  free(dtau); free(abc); free(A1); free(C1); free(F); free(G); free(ztmp); free(Stmp);
  return Iplus;
}

/* Formal() for the LTE FULL_STOKES case, formal.c:157-275 with solve_NLTE =
   FALSE (no sca_c*J term), Nrays = 1, and only the emergent (to_obs) ray:
   chi_c = chi_ai + chi_lines (background.c:476-537), S = eta/chi (formal.c:178-208) */
void rp_lte_stokes_column(const rp_linetable *lt, const rp_column *col,
                          int Nlambda, const double *lambda,
                          const double *chi_ai, const double *eta_ai,
                          int bc_top, int bc_bottom, double *stokes)
{
  int N = col->Ndep, nl, k, n, ie;
  double *elem_n = calloc((size_t) lt->nelem * RE_MAXSTAGE * N, sizeof(double));
  double *chl = malloc(4*N*sizeof(double)), *etl = malloc(4*N*sizeof(double));
  double *chi = malloc(N*sizeof(double)), *S = malloc(4*N*sizeof(double));
  double *I = malloc(4*N*sizeof(double)), *chiQUV = malloc(3*N*sizeof(double));

  for (ie = 0; ie < lt->nelem; ie++)
    rp_ltepops_elem(lt, ie, col, elem_n + (long) ie*RE_MAXSTAGE*N);

  for (nl = 0; nl < Nlambda; nl++) {
    int fl = rp_rlk_opacity(lt, col, elem_n, lambda[nl], 1, chl, etl);
    const double *ca = chi_ai + (long) nl*N, *ea = eta_ai + (long) nl*N;
    if (!(fl & 1)) memset(chl, 0, 4*N*sizeof(double)), memset(etl, 0, 4*N*sizeof(double));
    for (k = 0; k < N; k++) {
      chi[k] = ca[k] + chl[k];
      S[k]   = ea[k] + etl[k];
    }
    for (k = N; k < 4*N; k++) { chiQUV[k-N] = 0.0 + chl[k]; S[k] = 0.0 + (0.0 + etl[k]); }
    for (n = 0; n < 4; n++)
      for (k = 0; k < N; k++) S[n*N+k] /= chi[k];
    if (fl & 1) {
      rp_stokes_bezier3(N, col->height, col->muz, 1, chi, S, chiQUV, col->T, lambda[nl],
                        bc_top, bc_bottom, lt->matinv_simd, I, NULL);
      for (n = 0; n < 4; n++) stokes[(long) n*Nlambda + nl] = I[n*N];
    } else {
      /* no line in the window: not angle dependent -> Feautrier (formal.c:100-103, 289-309);
         J = 0 on the single LTE pass so the sca_c*Jdag term vanishes */
      stokes[nl] = rp_feautrier(N, col->height, col->muz, chi, S, col->T, lambda[nl],
                                bc_top, bc_bottom, I, NULL);
      for (n = 1; n < 4; n++) stokes[(long) n*Nlambda + nl] = 0.0;
    }
  }
  free(elem_n); free(chl); free(etl); free(chi); free(S); free(I); free(chiQUV);
}
