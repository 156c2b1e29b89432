// rhb200_piecewise.cuh -- the reference's linear / parabolic short-characteristics rays and the
// quasi-parabolic DELO polarised ray, one thread per ray (depth sequential).
//
// Reference: Piecewise_Linear_1D rh/rhf1d/piecewise_1D.c:44-127   (S_INTERPOLATION = S_LINEAR)
//            Piecewise_1D        rh/rhf1d/piecewise_1D.c:134-253  (S_INTERPOLATION = S_PARABOLIC)
//            Piece_Stokes_1D     rh/rhf1d/piecestokes_1D.c:49-174 (S_INTERPOLATION_STOKES = DELO_PARABOLIC)
//            w2, w3              rh/w3.c:25-63;  StokesK rh/stokesopac.c:28-87;  SolveLinearEq rh/ludcmp.c:36
// Operation order follows the reference statement by statement (bit-exact, -fmad=false).
#pragma once
#include "rhb200_delo.cuh"
#include "rhb200_lu.cuh"

namespace rhp {

__device__ __forceinline__ void w2(double dtau, double &w0, double &w1) {      // w3.c:25-39
  if (dtau < 5.0E-4) {
    w0 = dtau*(1.0 - 0.5*dtau);
    w1 = (dtau*dtau) * (0.5 - dtau/3.0);
  } else if (dtau > 50.0) {
    w1 = w0 = 1.0;
  } else {
    const double e = rhm::rh_exp(-dtau);
    w0 = 1.0 - e;
    w1 = w0 - dtau*e;
  }
}

__device__ __forceinline__ void w3(double dtau, double &w0, double &w1, double &w2_) {   // w3.c:43-63
  if (dtau < 5.0E-4) {
    w0 = dtau*(1.0 - 0.5*dtau);
    double delta = dtau*dtau;
    w1 = delta*(0.5 - dtau/3.0);
    delta *= dtau;
    w2_ = delta*(1.0/3.0 - 0.25*dtau);
  } else if (dtau > 50.0) {
    w1 = w0 = 1.0;
    w2_ = 2.0;
  } else {
    const double e = rhm::rh_exp(-dtau);
    w0 = 1.0 - e;
    w1 = w0 - dtau*e;
    w2_ = 2.0*w1 - (dtau*dtau) * e;
  }
}

// upwind boundary intensity, piecewise_1D.c:70-107 (ZERO / THERMALIZED)
__device__ __forceinline__ double boundary_I(int ndep, int to_obs, int bc_top, int bc_bottom,
                                             const double *__restrict__ T, double lambda, double dtau_uw) {
  if (to_obs) {
    if (bc_bottom == RHB200_BC_THERMALIZED) {
      const double B0 = rhd::planck(T[ndep-2], lambda), B1 = rhd::planck(T[ndep-1], lambda);
      return B1 - (B0 - B1) / dtau_uw;
    }
  } else if (bc_top == RHB200_BC_THERMALIZED) {
    const double B0 = rhd::planck(T[0], lambda), B1 = rhd::planck(T[1], lambda);
    return B0 - (B1 - B0) / dtau_uw;
  }
  return 0.0;
}

__device__ __forceinline__ void linear_ray(const int ndep, const double *__restrict__ z, const double muz,
                                           const int to_obs, const int bc_top, const int bc_bottom,
                                           const double *__restrict__ T, const double lambda,
                                           const double *__restrict__ chi, const double *__restrict__ S,
                                           double *__restrict__ I, double *__restrict__ Psi)
{
  const double zmu = 0.5 / muz;
  const int dk = to_obs ? -1 : 1;
  const int ks = to_obs ? ndep-1 : 0, ke = to_obs ? 0 : ndep-1;
  double dtau_uw = zmu * (chi[ks] + chi[ks+dk]) * fabs(z[ks] - z[ks+dk]);
  double dS_uw = (S[ks] - S[ks+dk]) / dtau_uw;
  double I_uw = boundary_I(ndep, to_obs, bc_top, bc_bottom, T, lambda, dtau_uw);
  double w0 = 0.0, w1 = 0.0;
  I[ks] = I_uw;
  if (Psi) Psi[ks] = 0.0;
  for (int k = ks+dk; k != ke; k += dk) {
    w2(dtau_uw, w0, w1);
    I[k] = (1.0 - w0)*I_uw + w0*S[k] + w1*dS_uw;
    if (Psi) Psi[k] = w0 - w1 / dtau_uw;
    dtau_uw = zmu * (chi[k] + chi[k+dk]) * fabs(z[k] - z[k+dk]);
    dS_uw   = (S[k] - S[k+dk]) / dtau_uw;
    I_uw = I[k];
  }
  // the end point re-uses the previous interval's weights (piecewise_1D.c:125-126)
  I[ke] = (1.0 - w0)*I_uw + w0*S[ke] + w1*dS_uw;
  if (Psi) Psi[ke] = w0 - w1 / dtau_uw;
}

__device__ __forceinline__ void parabolic_ray(const int ndep, const double *__restrict__ z, const double muz,
                                              const int to_obs, const int bc_top, const int bc_bottom,
                                              const double *__restrict__ T, const double lambda,
                                              const double *__restrict__ chi, const double *__restrict__ S,
                                              double *__restrict__ I, double *__restrict__ Psi)
{
  const double zmu = 0.5 / muz;
  const int dk = to_obs ? -1 : 1;
  const int ks = to_obs ? ndep-1 : 0, ke = to_obs ? 0 : ndep-1;
  double dtau_uw = zmu * (chi[ks] + chi[ks+dk]) * fabs(z[ks] - z[ks+dk]);
  double I_uw = boundary_I(ndep, to_obs, bc_top, bc_bottom, T, lambda, dtau_uw);
  I[ks] = I_uw;
  if (Psi) Psi[ks] = 0.0;
  double dS_uw = (S[ks] - S[ks+dk]) / dtau_uw, dtau_dw = 0.0, dS_dw = 0.0;
  for (int k = ks+dk; k != ke+dk; k += dk) {
    double w0, w1, w2_;
    w3(dtau_uw, w0, w1, w2_);
    double Ik;
    if (k != ke) {
      dtau_dw = zmu * (chi[k] + chi[k+dk]) * fabs(z[k] - z[k+dk]);
      dS_dw   = (S[k] - S[k+dk]) / dtau_dw;
      double c1 = (dS_uw*dtau_dw + dS_dw*dtau_uw);
      const double c2 = (dS_uw - dS_dw);
      Ik = (1.0 - w0)*I_uw + w0*S[k] + (w1*c1 + w2_*c2) / (dtau_uw + dtau_dw);
      if (Ik < 0.0) {                            // linear fallback, piecewise_1D.c:223-228
        c1 = dS_uw;
        Ik = (1.0 - w0)*I_uw + w0*S[k] + w1*c1;
        if (Psi) Psi[k] = w0 - w1/dtau_uw;
      } else if (Psi) {
        c1 = dtau_uw - dtau_dw;
        Psi[k] = w0 + (w1*c1 - w2_) / (dtau_uw * dtau_dw);
      }
    } else {
      Ik = (1.0 - w0)*I_uw + w0*S[k] + w1*dS_uw;
      if (Psi) Psi[k] = w0 - w1 / dtau_uw;
    }
    I[k] = Ik;
    I_uw = Ik;
    dS_uw = dS_dw;
    dtau_uw = dtau_dw;
  }
}

// Piece_Stokes_1D with the IO policies of rhb200_delo.cu (chi, K'[0][1..3], S[4], storeI, storePsi).
// K' = [[0,q,u,v],[q,0,0,0],[u,0,0,0],[v,0,0,0]] (MAGNETO_OPTICAL = FALSE): Q and R are built as the
// full 4x4 matrices the reference hands to SolveLinearEq (zeros multiply to exact zeros, and the LU
// pivot search sees the same entries).
template <class IO>
__device__ __forceinline__ void stokes_parabolic_ray(IO &io, const int ndep, const double *__restrict__ z,
                                                     const double muz, const int to_obs,
                                                     const int bc_top, const int bc_bottom,
                                                     const double *__restrict__ T, const double lambda)
{
  const double zmu = 0.5 / muz;
  const int dk = to_obs ? -1 : 1;
  const int ks = to_obs ? ndep-1 : 0, ke = to_obs ? 0 : ndep-1;
  double chi_k = io.chi(ks), chi_n = io.chi(ks+dk);
  double dtau_uw = zmu * (chi_k + chi_n) * fabs(z[ks] - z[ks+dk]), dtau_dw = 0.0;
  double Ku[3], Kc[3], Sk[4], Sn[4], dS_uw[4], dS_dw[4] = {0.0, 0.0, 0.0, 0.0}, I_upw[4], P[4];
  io.K(ks, Ku);
  I_upw[0] = boundary_I(ndep, to_obs, bc_top, bc_bottom, T, lambda, dtau_uw);
  I_upw[1] = I_upw[2] = I_upw[3] = 0.0;
  io.S(ks, Sk); io.S(ks+dk, Sn);
#pragma unroll
  for (int n = 0; n < 4; n++) dS_uw[n] = (Sk[n] - Sn[n]) / dtau_uw;
  io.storeI(ks, I_upw);
  io.storePsi(ks, 0.0);

  for (int k = ks+dk; k != ke+dk; k += dk) {
    double w0, w1, w2_;
    w3(dtau_uw, w0, w1, w2_);
    io.K(k, Kc);
#pragma unroll
    for (int n = 0; n < 4; n++) Sk[n] = Sn[n];
    chi_k = chi_n;
    if (k != ke) {
      chi_n = io.chi(k+dk);
      io.S(k+dk, Sn);
      dtau_dw = zmu * (chi_k + chi_n) * fabs(z[k] - z[k+dk]);
#pragma unroll
      for (int n = 0; n < 4; n++) {
        dS_dw[n] = (Sk[n] - Sn[n]) / dtau_dw;
        const double c1 = dS_uw[n]*dtau_dw + dS_dw[n]*dtau_uw;
        const double c2 = dS_uw[n] - dS_dw[n];
        P[n] = w0*Sk[n] + (w1*c1 + w2_*c2) / (dtau_uw + dtau_dw);
      }
      const double c1 = dtau_uw - dtau_dw;
      io.storePsi(k, w0 + (w1*c1 - w2_) / (dtau_uw * dtau_dw));
    } else {
#pragma unroll
      for (int n = 0; n < 4; n++) P[n] = w0*Sk[n] + w1*dS_uw[n];
      io.storePsi(k, w0 - w1 / dtau_uw);
    }
    // Q = -w1/dtau K_upw (diag 1-w0), R = (w0 - w1/dtau) K (diag 1): piecestokes_1D.c:139-151
    const double qf = -w1/dtau_uw, rf = (w0 - w1/dtau_uw);
    double R[16];
#pragma unroll
    for (int i = 0; i < 16; i++) R[i] = rf * 0.0;
#pragma unroll
    for (int m = 1; m < 4; m++) { R[m] = rf * Kc[m-1]; R[4*m] = rf * Kc[m-1]; }
    R[0] = R[5] = R[10] = R[15] = 1.0;
    {
      const double q00 = 1.0 - w0;
      const double q1 = qf * Ku[0], q2 = qf * Ku[1], q3 = qf * Ku[2], qz = qf * 0.0;
      // P[n] += sum_m Q[n][m] I_upw[m], m ascending
      P[0] += q00 * I_upw[0]; P[0] += q1 * I_upw[1]; P[0] += q2 * I_upw[2]; P[0] += q3 * I_upw[3];
      P[1] += q1 * I_upw[0];  P[1] += q00 * I_upw[1]; P[1] += qz * I_upw[2]; P[1] += qz * I_upw[3];
      P[2] += q2 * I_upw[0];  P[2] += qz * I_upw[1];  P[2] += q00 * I_upw[2]; P[2] += qz * I_upw[3];
      P[3] += q3 * I_upw[0];  P[3] += qz * I_upw[1];  P[3] += qz * I_upw[2]; P[3] += q00 * I_upw[3];
    }
    rhlu::solve_linear_eq<4>(4, R, P, true);
    io.storeI(k, P);
    dtau_uw = dtau_dw;
#pragma unroll
    for (int n = 0; n < 4; n++) { I_upw[n] = P[n]; dS_uw[n] = dS_dw[n]; }
    Ku[0] = Kc[0]; Ku[1] = Kc[1]; Ku[2] = Kc[2];
  }
}

}  // namespace rhp
