/* rhport_piecewise.c -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Plain-C restatement of the reference's non-Bezier short-characteristics solvers:
 *   Piecewise_Linear_1D   rh/rhf1d/piecewise_1D.c:44-127     (S_INTERPOLATION = S_LINEAR)
 *   Piecewise_1D          rh/rhf1d/piecewise_1D.c:134-253    (S_INTERPOLATION = S_PARABOLIC)
 *   Piece_Stokes_1D       rh/rhf1d/piecestokes_1D.c:49-174   (S_INTERPOLATION_STOKES = DELO_PARABOLIC)
 *   w2                    rh/w3.c:25-39
 * Pinned bit-exact against calls recorded from the compiled reference (tests/golden/falc_solvers.npz).
 * Only the boundary conditions the pyrh path produces are restated (ZERO / THERMALIZED).
 */
#include <math.h>
#include "rhport.h"

void rp_w2(double dtau, double *w)                      /* w3.c:25-39 */
{
  double expdt;
  if (dtau < 5.0E-4) {
    w[0] = dtau*(1.0 - 0.5*dtau);
    w[1] = (dtau*dtau) * (0.5 - dtau/3.0);
  } else if (dtau > 50.0) {
    w[1] = w[0] = 1.0;
  } else {
    expdt = exp(-dtau);
    w[0]  = 1.0 - expdt;
    w[1]  = w[0] - dtau*expdt;
  }
}

/* upwind boundary intensity shared by the three solvers (piecewise_1D.c:70-107) */
static double boundary_I(int Ndep, int to_obs, int bc_top, int bc_bottom, const double *T,
                         double lambda, double dtau_uw)
{
  if (to_obs) {
    if (bc_bottom == RP_THERMALIZED) {
      double B0 = rp_planck(T[Ndep-2], lambda), B1 = rp_planck(T[Ndep-1], lambda);
      return B1 - (B0 - B1) / dtau_uw;
    }
  } else if (bc_top == RP_THERMALIZED) {
    double B0 = rp_planck(T[0], lambda), B1 = rp_planck(T[1], lambda);
    return B0 - (B1 - B0) / dtau_uw;
  }
  return 0.0;
}

void rp_piecewise_linear(int Ndep, const double *z, double muz, int to_obs,
                         const double *chi, const double *S, const double *T, double lambda,
                         int bc_top, int bc_bottom, double *I, double *Psi)
{
  int k, k_start, k_end, dk;
  double dtau_uw, dS_uw, I_uw, w[2] = {0.0, 0.0}, zmu = 0.5 / muz;

  if (to_obs) { dk = -1; k_start = Ndep-1; k_end = 0; }
  else        { dk =  1; k_start = 0;      k_end = Ndep-1; }
  dtau_uw = zmu * (chi[k_start] + chi[k_start+dk]) * fabs(z[k_start] - z[k_start+dk]);
  dS_uw = (S[k_start] - S[k_start+dk]) / dtau_uw;
  I_uw = boundary_I(Ndep, to_obs, bc_top, bc_bottom, T, lambda, dtau_uw);
  I[k_start] = I_uw;
  if (Psi) Psi[k_start] = 0.0;

  for (k = k_start+dk; k != k_end; k += dk) {
    rp_w2(dtau_uw, w);
    I[k] = (1.0 - w[0])*I_uw + w[0]*S[k] + w[1]*dS_uw;
    if (Psi) Psi[k] = w[0] - w[1] / dtau_uw;
    dtau_uw = zmu * (chi[k] + chi[k+dk]) * fabs(z[k] - z[k+dk]);
    dS_uw   = (S[k] - S[k+dk]) / dtau_uw;
    I_uw = I[k];
  }
  /* the last point re-uses the weights of the previous interval (piecewise_1D.c:125-126) */
  I[k_end] = (1.0 - w[0])*I_uw + w[0]*S[k_end] + w[1]*dS_uw;
  if (Psi) Psi[k_end] = w[0] - w[1] / dtau_uw;
}

void rp_piecewise_parabolic(int Ndep, const double *z, double muz, int to_obs,
                            const double *chi, const double *S, const double *T, double lambda,
                            int bc_top, int bc_bottom, double *I, double *Psi)
{
  int k, k_start, k_end, dk;
  double dtau_uw, dtau_dw = 0.0, dS_uw, I_uw, dS_dw = 0.0, c1, c2, w[3], zmu = 0.5 / muz;

  if (to_obs) { dk = -1; k_start = Ndep-1; k_end = 0; }
  else        { dk =  1; k_start = 0;      k_end = Ndep-1; }
  dtau_uw = zmu * (chi[k_start] + chi[k_start+dk]) * fabs(z[k_start] - z[k_start+dk]);
  I_uw = boundary_I(Ndep, to_obs, bc_top, bc_bottom, T, lambda, dtau_uw);
  I[k_start] = I_uw;
  if (Psi) Psi[k_start] = 0.0;
  dS_uw = (S[k_start] - S[k_start+dk]) / dtau_uw;

  for (k = k_start+dk; k != k_end+dk; k += dk) {
    rp_w3(dtau_uw, w);
    if (k != k_end) {
      dtau_dw = zmu * (chi[k] + chi[k+dk]) * fabs(z[k] - z[k+dk]);
      dS_dw   = (S[k] - S[k+dk]) / dtau_dw;
      c1 = (dS_uw*dtau_dw + dS_dw*dtau_uw);
      c2 = (dS_uw - dS_dw);
      I[k] = (1.0 - w[0])*I_uw + w[0]*S[k] + (w[1]*c1 + w[2]*c2) / (dtau_uw + dtau_dw);
      if (I[k] < 0.0) {                                  /* fall back to linear, :223-228 */
        c1   = dS_uw;
        I[k] = (1.0 - w[0])*I_uw + w[0]*S[k] + w[1]*c1;
        if (Psi) Psi[k] = w[0] - w[1]/dtau_uw;
      } else if (Psi) {
        c1 = dtau_uw - dtau_dw;
        Psi[k] = w[0] + (w[1]*c1 - w[2]) / (dtau_uw * dtau_dw);
      }
    } else {
      I[k] = (1.0 - w[0])*I_uw + w[0]*S[k] + w[1]*dS_uw;
      if (Psi) Psi[k] = w[0] - w[1] / dtau_uw;
    }
    I_uw = I[k];
    dS_uw   = dS_dw;
    dtau_uw = dtau_dw;
  }
}

/* StokesK, stokesopac.c:28-87 with magneto_optical = FALSE and chiQUV = as->chi[QUV] + as->chi_c[QUV] */
static void stokesK(const double *chiQUV, int N, int k, double chi_I, double K[4][4])
{
  int i, j;
  for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) K[j][i] = 0.0;
  K[0][1] = chiQUV[k]; K[0][2] = chiQUV[N+k]; K[0][3] = chiQUV[2*N+k];
  for (j = 0; j < 3; j++)
    for (i = j+1; i < 4; i++) { K[j][i] /= chi_I; K[i][j] = K[j][i]; }
}

/* S, I: [4][Ndep]; chiQUV: [3][Ndep] */
void rp_stokes_parabolic(int Ndep, const double *z, double muz, int to_obs,
                         const double *chi_I, const double *S, const double *chiQUV,
                         const double *T, double lambda, int bc_top, int bc_bottom,
                         double *I, double *Psi)
{
  int k, n, m, k_start, k_end, dk, N = Ndep;
  double dtau_uw, dtau_dw = 0.0, dS_uw[4], dS_dw[4] = {0, 0, 0, 0}, c1, c2, w[3], I_upw[4], zmu = 0.5 / muz,
         P[4], Q[4][4], R[16], K[4][4], K_upw[4][4];

  if (to_obs) { dk = -1; k_start = Ndep-1; k_end = 0; }
  else        { dk =  1; k_start = 0;      k_end = Ndep-1; }
  dtau_uw = zmu * (chi_I[k_start] + chi_I[k_start+dk]) * fabs(z[k_start] - z[k_start+dk]);
  stokesK(chiQUV, N, k_start, chi_I[k_start], K_upw);

  I_upw[0] = boundary_I(Ndep, to_obs, bc_top, bc_bottom, T, lambda, dtau_uw);
  for (n = 1; n < 4; n++) I_upw[n] = 0.0;
  for (n = 0; n < 4; n++) dS_uw[n] = (S[n*N + k_start] - S[n*N + k_start+dk]) / dtau_uw;
  for (n = 0; n < 4; n++) I[n*N + k_start] = I_upw[n];
  if (Psi) Psi[k_start] = 0.0;

  for (k = k_start+dk; k != k_end+dk; k += dk) {
    rp_w3(dtau_uw, w);
    stokesK(chiQUV, N, k, chi_I[k], K);
    if (k != k_end) {
      dtau_dw = zmu * (chi_I[k] + chi_I[k+dk]) * fabs(z[k] - z[k+dk]);
      for (n = 0; n < 4; n++) {
        dS_dw[n] = (S[n*N + k] - S[n*N + k+dk]) / dtau_dw;
        c1 = dS_uw[n]*dtau_dw + dS_dw[n]*dtau_uw;
        c2 = dS_uw[n] - dS_dw[n];
        P[n] = w[0]*S[n*N + k] + (w[1]*c1 + w[2]*c2) / (dtau_uw + dtau_dw);
      }
      if (Psi) {
        c1 = dtau_uw - dtau_dw;
        Psi[k] = w[0] + (w[1]*c1 - w[2]) / (dtau_uw * dtau_dw);
      }
    } else {
      for (n = 0; n < 4; n++) P[n] = w[0]*S[n*N + k] + w[1]*dS_uw[n];
      if (Psi) Psi[k] = w[0] - w[1] / dtau_uw;
    }
    for (n = 0; n < 4; n++) {
      for (m = 0; m < 4; m++) {
        Q[n][m] = -w[1]/dtau_uw * K_upw[n][m];
        R[n*4 + m] = (w[0] - w[1]/dtau_uw) * K[n][m];
      }
      Q[n][n] = 1.0 - w[0];
      R[n*4 + n] = 1.0;
    }
    for (n = 0; n < 4; n++)
      for (m = 0; m < 4; m++) P[n] += Q[n][m] * I_upw[m];

    rp_solve_linear_eq(4, R, P, 1);                      /* piecestokes_1D.c:156 */

    for (n = 0; n < 4; n++) I[n*N + k] = P[n];
    dtau_uw = dtau_dw;
    for (n = 0; n < 4; n++) {
      I_upw[n] = I[n*N + k];
      dS_uw[n] = dS_dw[n];
      for (m = 0; m < 4; m++) K_upw[n][m] = K[n][m];
    }
  }
}
