// rhb200_bezier.cuh -- scalar cubic-Bezier short-characteristics ray (unpolarised).
// Reference: Piecewise_Bezier3_1D rh/rhf1d/bezier_1D.c:306-541.  RF = true adds the log gf
// response-function branch (:416-428, :477-490, :509-516): dchi/deta [ndep][npar] are
// spectrum.dchi_c_lam/deta_c_lam of the wavelength, dI [ndep][npar] the output; I must then enter
// holding the preceding down-ray solution, which the reference reads at not yet updated depths.
#pragma once
#include "rhb200_delo.cuh"

namespace rhz {

#define RHB200_MAXPAR 16

template <bool RF>
__device__ __forceinline__ void bezier3_ray_t(const int ndep, const double *__restrict__ z, const double muz,
                                              const int to_obs, const int bc_top, const int bc_bottom,
                                              const double *__restrict__ T, const double lambda,
                                              const double *__restrict__ chi, const double *__restrict__ S,
                                              double *I, double *__restrict__ Psi,
                                              const int npar, const double *__restrict__ dchi,
                                              const double *__restrict__ deta, double *__restrict__ dI,
                                              const bool psi_over_chi = false /* store Psi[k] / chi[k] (formal.c:247-248) */)
{
  using namespace rhd;
  const double zmu = 1.0 / muz;
  const rhdiv::Recip third(3.0);
  const int dk = to_obs ? -1 : 1;
  const int ks = to_obs ? ndep-1 : 0, ke = to_obs ? 0 : ndep-1;

  double dtau_uw = 0.5 * zmu * (chi[ks] + chi[ks+dk]) * fabs(z[ks] - z[ks+dk]);
  double I_upw = 0.0;                                   // bezier_1D.c:352-386
  if (to_obs) {
    if (bc_bottom == RHB200_BC_THERMALIZED) {
      const double B0 = planck(T[ndep-2], lambda), B1 = planck(T[ndep-1], lambda);
      I_upw = B1 - (B0 - B1) / dtau_uw;
    }
  } else if (bc_top == RHB200_BC_THERMALIZED) {
    const double B0 = planck(T[0], lambda), B1 = planck(T[1], lambda);
    I_upw = B0 - (B1 - B0) / dtau_uw;
  }
  I[ks] = I_upw;
  if (Psi) Psi[ks] = psi_over_chi ? 0.0 / chi[ks] : 0.0;

  int k = ks + dk;
  double dsup = fabs(z[k] - z[k-dk]) * zmu;
  double dsdn = fabs(z[k+dk] - z[k]) * zmu;
  double dchi_up = (chi[k] - chi[k-dk]) / dsup;
  double fchi = (chi[k+dk] - chi[k]) / dsdn;
  double dchi_c = fb_deriv(dchi_up, fchi, fb_alpha(dsup, dsdn));
  {
    const double dsup3 = third.div(dsup);
    const double c1 = RH_MAX0(chi[k]    - dsup3 * dchi_c);
    const double c2 = RH_MAX0(chi[k-dk] + dsup3 * dchi_up);
    dtau_uw = dsup * (chi[k] + chi[k-dk] + c1 + c2) * 0.25;
  }
  double dS_up = (S[k] - S[k-dk]) / dtau_uw;
  double fS = dS_up, dtau_dw = 0.0, dchi_dn = 0.0, dS_c = 0.0;
  double dZk[RF ? RHB200_MAXPAR : 1], dZup[RF ? RHB200_MAXPAR : 1], dI_upw[RF ? RHB200_MAXPAR : 1];
  if (RF) {
    for (int p = 0; p < npar; p++) {                     // bezier_1D.c:416-428
      dI[ks*npar + p] = 0.0;
      dI_upw[p] = 0.0;
      const double Zk = -dchi[k*npar + p]/chi[k] * I[k] + deta[k*npar + p]/chi[k];
      const double Zkm1 = -dchi[(k-dk)*npar + p]/chi[k-dk] * I[k-dk] + deta[(k-dk)*npar + p]/chi[k-dk];
      dZup[p] = (Zk - Zkm1) / dtau_uw;
    }
  }

  for (; k != ke + dk; k += dk) {
    if (k != ke) {
      dsdn = fabs(z[k+dk] - z[k]) * zmu;
      double fnext = fchi;
      if (abs(k - ke) > 1) {
        const double dsdn2 = fabs(z[k+2*dk] - z[k+dk]) * zmu;
        fnext = (chi[k+2*dk] - chi[k+dk]) / dsdn2;
        dchi_dn = fb_deriv(fchi, fnext, fb_alpha(dsdn, dsdn2));
      } else
        dchi_dn = fchi;
      const double dsdn3 = third.div(dsdn);
      double c1 = RH_MAX0(chi[k]    + dsdn3 * dchi_c);
      double c2 = RH_MAX0(chi[k+dk] - dsdn3 * dchi_dn);
      dtau_dw = dsdn * (chi[k] + chi[k+dk] + c1 + c2) * 0.25;
      const double dt03 = third.div(dtau_uw);
      double alpha, beta, gamma, theta, eps;
      bezier3_coeffs(dtau_uw, alpha, beta, gamma, theta, eps);
      const double fi = (S[k+dk] - S[k]) / dtau_dw;
      dS_c = fb_deriv(fS, fi, fb_alpha(dtau_uw, dtau_dw));
      fS = fi;
      c1 = RH_MAX0(S[k]    - dt03 * dS_c);
      c2 = RH_MAX0(S[k-dk] + dt03 * dS_up);
      I[k] = I_upw*eps + alpha*S[k] + beta*S[k-dk] + gamma * c1 + theta * c2;
      if (RF) {
        const double ca = fb_alpha(dtau_uw, dtau_dw);
        for (int p = 0; p < npar; p++) {                 // bezier_1D.c:477-490
          double Zk = -dchi[k*npar + p] * I[k] + deta[k*npar + p];
          Zk /= chi[k];
          double Zkm1 = -dchi[(k-dk)*npar + p] * I[k-dk] + deta[(k-dk)*npar + p];
          Zkm1 /= chi[k-dk];
          double Zkp1 = -dchi[(k+dk)*npar + p] * I[k+dk] + deta[(k+dk)*npar + p];   // I[k+dk]: down-ray value
          Zkp1 /= chi[k+dk];
          dZk[p] = fb_deriv((Zk - Zkm1) / dtau_uw, (Zkp1 - Zk) / dtau_dw, ca);
          const double z1 = RH_MAX0(Zk - dt03 * dZk[p]);
          const double z2 = RH_MAX0(Zkm1 + dt03 * dZup[p]);
          dI[k*npar + p] = dI_upw[p]*eps + alpha*Zk + beta*Zkm1 + gamma*z1 + theta*z2;
        }
      }
      if (Psi) Psi[k] = psi_over_chi ? (alpha + gamma) / chi[k] : alpha + gamma;
      fchi = fnext;
    } else {
      dtau_uw = 0.5 * zmu * (chi[k] + chi[k-dk]) * fabs(z[k] - z[k-dk]);
      const double dS_uw = -(S[k] - S[k-dk]) / dtau_uw;
      double w0, w1;
      w3(dtau_uw, w0, w1);
      I[k] = (1.0 - w0)*I_upw + w0*S[k] + w1*dS_uw;
      if (RF) {
        for (int p = 0; p < npar; p++) {                 // bezier_1D.c:509-516
          const double Zk = dchi[k*npar + p]/chi[k] * I[k] - deta[k*npar + p]/chi[k];
          const double Zkm1 = dchi[(k-dk)*npar + p]/chi[k-dk] * I[k-dk] - deta[(k-dk)*npar + p]/chi[k-dk];
          dZk[p] = -(Zk - Zkm1) / dtau_uw;
          dI[k*npar + p] = (1.0 - w0)*dI_upw[p] + w0*Zk + w1*dZk[p];
        }
      }
      if (Psi) Psi[k] = psi_over_chi ? (w0 - w1 / dtau_uw) / chi[k] : w0 - w1 / dtau_uw;
    }
    I_upw = I[k];
    dsup = dsdn; dchi_up = dchi_c; dchi_c = dchi_dn; dtau_uw = dtau_dw; dS_up = dS_c;
    if (RF) for (int p = 0; p < npar; p++) { dI_upw[p] = dI[k*npar + p]; dZup[p] = dZk[p]; }
  }
}

__device__ __forceinline__ void bezier3_ray(const int ndep, const double *__restrict__ z, const double muz,
                                            const int to_obs, const int bc_top, const int bc_bottom,
                                            const double *__restrict__ T, const double lambda,
                                            const double *__restrict__ chi, const double *__restrict__ S,
                                            double *__restrict__ I, double *__restrict__ Psi, const bool psi_over_chi = false)
{
  bezier3_ray_t<false>(ndep, z, muz, to_obs, bc_top, bc_bottom, T, lambda, chi, S, I, Psi, 0, nullptr, nullptr, nullptr, psi_over_chi);
}

// The same ray (no response function) with chi, S and z held in a sliding register window and the next step's three new
// values loaded one step ahead: three loads per step instead of a dozen re-reads, issued a whole step before their
// first use.  Same values, same arithmetic, same order as bezier3_ray_t<false>.
__device__ __forceinline__ void bezier3_ray_w(const int ndep, const double *__restrict__ z, const double muz,
                                              const int to_obs, const int bc_top, const int bc_bottom,
                                              const double *__restrict__ T, const double lambda,
                                              const double *__restrict__ chi, const double *__restrict__ S,
                                              double *__restrict__ I, double *__restrict__ Psi, const bool psi_over_chi)
{
  using namespace rhd;
  const double zmu = 1.0 / muz;
  const rhdiv::Recip third(3.0);
  const int dk = to_obs ? -1 : 1;
  const int ks = to_obs ? ndep-1 : 0, ke = to_obs ? 0 : ndep-1;
  // window around k = ks + dk: m = k - dk, 0 = k, p = k + dk, pp = k + 2 dk
  double chi_m = chi[ks], chi_0 = chi[ks+dk], S_m = S[ks], S_0 = S[ks+dk], z_m = z[ks], z_0 = z[ks+dk];
  double chi_p = chi[ks+2*dk], S_p = S[ks+2*dk], z_p = z[ks+2*dk];
  double chi_pp = (ndep > 3) ? chi[ks+3*dk] : 0.0, z_pp = (ndep > 3) ? z[ks+3*dk] : 0.0;

  double dtau_uw = 0.5 * zmu * (chi_m + chi_0) * fabs(z_m - z_0);
  double I_upw = 0.0;                                   // bezier_1D.c:352-386
  if (to_obs) {
    if (bc_bottom == RHB200_BC_THERMALIZED) {
      const double B0 = planck(T[ndep-2], lambda), B1 = planck(T[ndep-1], lambda);
      I_upw = B1 - (B0 - B1) / dtau_uw;
    }
  } else if (bc_top == RHB200_BC_THERMALIZED) {
    const double B0 = planck(T[0], lambda), B1 = planck(T[1], lambda);
    I_upw = B0 - (B1 - B0) / dtau_uw;
  }
  I[ks] = I_upw;
  if (Psi) Psi[ks] = psi_over_chi ? 0.0 / chi_m : 0.0;

  int k = ks + dk;
  double dsup = fabs(z_0 - z_m) * zmu;
  double dsdn = fabs(z_p - z_0) * zmu;
  double dchi_up = (chi_0 - chi_m) / dsup;
  double fchi = (chi_p - chi_0) / dsdn;
  double dchi_c = fb_deriv(dchi_up, fchi, fb_alpha(dsup, dsdn));
  {
    const double dsup3 = third.div(dsup);
    const double c1 = RH_MAX0(chi_0 - dsup3 * dchi_c);
    const double c2 = RH_MAX0(chi_m + dsup3 * dchi_up);
    dtau_uw = dsup * (chi_0 + chi_m + c1 + c2) * 0.25;
  }
  double dS_up = (S_0 - S_m) / dtau_uw;
  double fS = dS_up, dtau_dw = 0.0, dchi_dn = 0.0, dS_c = 0.0;

  for (; k != ke + dk; k += dk) {
    // the values the NEXT step adds to the window: chi and z at k + 3 dk, S at k + 2 dk
    const int k3 = k + 3*dk, k2 = k + 2*dk;
    const bool in3 = k3 >= 0 && k3 < ndep, in2 = k2 >= 0 && k2 < ndep;
    const double chi_n = in3 ? chi[k3] : 0.0, z_n = in3 ? z[k3] : 0.0, S_n = in2 ? S[k2] : 0.0;
    double Ik;
    if (k != ke) {
      dsdn = fabs(z_p - z_0) * zmu;
      double fnext = fchi;
      if (abs(k - ke) > 1) {
        const double dsdn2 = fabs(z_pp - z_p) * zmu;
        fnext = (chi_pp - chi_p) / dsdn2;
        dchi_dn = fb_deriv(fchi, fnext, fb_alpha(dsdn, dsdn2));
      } else
        dchi_dn = fchi;
      const double dsdn3 = third.div(dsdn);
      double c1 = RH_MAX0(chi_0 + dsdn3 * dchi_c);
      double c2 = RH_MAX0(chi_p - dsdn3 * dchi_dn);
      dtau_dw = dsdn * (chi_0 + chi_p + c1 + c2) * 0.25;
      const double dt03 = third.div(dtau_uw);
      double alpha, beta, gamma, theta, eps;
      bezier3_coeffs(dtau_uw, alpha, beta, gamma, theta, eps);
      const double fi = (S_p - S_0) / dtau_dw;
      dS_c = fb_deriv(fS, fi, fb_alpha(dtau_uw, dtau_dw));
      fS = fi;
      c1 = RH_MAX0(S_0 - dt03 * dS_c);
      c2 = RH_MAX0(S_m + dt03 * dS_up);
      Ik = I_upw*eps + alpha*S_0 + beta*S_m + gamma * c1 + theta * c2;
      I[k] = Ik;
      if (Psi) Psi[k] = psi_over_chi ? (alpha + gamma) / chi_0 : alpha + gamma;
      fchi = fnext;
    } else {
      dtau_uw = 0.5 * zmu * (chi_0 + chi_m) * fabs(z_0 - z_m);
      const double dS_uw = -(S_0 - S_m) / dtau_uw;
      double w0, w1;
      w3(dtau_uw, w0, w1);
      Ik = (1.0 - w0)*I_upw + w0*S_0 + w1*dS_uw;
      I[k] = Ik;
      if (Psi) Psi[k] = psi_over_chi ? (w0 - w1 / dtau_uw) / chi_0 : w0 - w1 / dtau_uw;
    }
    I_upw = Ik;
    dsup = dsdn; dchi_up = dchi_c; dchi_c = dchi_dn; dtau_uw = dtau_dw; dS_up = dS_c;
    chi_m = chi_0; chi_0 = chi_p; chi_p = chi_pp; chi_pp = chi_n;
    z_m = z_0; z_0 = z_p; z_p = z_pp; z_pp = z_n;
    S_m = S_0; S_0 = S_p; S_p = S_n;
  }
}

}  // namespace rhz
