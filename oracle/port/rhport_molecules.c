/* rhport_molecules.c -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Plain-C restatement of the reference's LTE molecular background lines:
 *   MolecularOpacity   rh/opacity.c:711-839
 *   MolProfile         rh/opacity.c:844-916   (MAGNETO_OPTICAL = FALSE)
 * for one column, one wavelength, one direction.  Pinned bit-exact against calls recorded from the
 * compiled reference (tests/golden/falc_molecules.npz).
 */
#include <math.h>
#include <stddef.h>
#include "rhport.h"

/* mlines [nmline][16]: lambda0, Ei, gi, Bij, Aji, Bji, isotope_frac, qwing, polarizable, molecule, zoff, ncomp
   mol [nmol][3][N]: n, pf, vbroad.  chi, eta [4][N].  Returns flags: bit0 hasline, bit1 ispolarized. */
int rp_molecular_opacity(int N, int nmol, int nmline, const double *mlines,
                         const int *zq, const double *zshift, const double *zstrength,
                         double vmicro_char, double lambda, double muz, int moving, int to_obs,
                         const double *T, const double *vel, const double *B,
                         const double *cos_gamma, const double *cos_2chi, const double *sin_2chi,
                         const double *mol, double *chi, double *eta)
{
  const double hc_4PI = (RP_HPLANCK * RP_CLIGHT) / (4.0 * RP_PI);
  int flags = 0, m, n, k, nz;
  for (k = 0; k < 4*N; k++) { chi[k] = 0.0; eta[k] = 0.0; }
  for (m = 0; m < nmol; m++) {
    int first = -1, last = -1;
    for (n = 0; n < nmline; n++)
      if ((int) mlines[n*16 + 9] == m) { if (first < 0) first = n; last = n; }
    if (first < 0) continue;
    double dl0 = lambda * mlines[first*16 + 7] * (vmicro_char / RP_CLIGHT);
    double dlN = lambda * mlines[last*16 + 7] * (vmicro_char / RP_CLIGHT);
    if (!(lambda >= mlines[first*16] - dl0 && lambda <= mlines[last*16] + dlN)) continue;
    const double *Mn = mol + (size_t) m*3*N, *Mpf = Mn + N, *Mvb = Mn + 2*N;
    for (n = first; n <= last; n++) {
      const double *L = mlines + n*16;
      double lambda0 = L[0], dl = lambda * L[7] * (vmicro_char / RP_CLIGHT);
      if (!(fabs(lambda0 - lambda) <= dl)) continue;
      double hc_la = (RP_HPLANCK * RP_CLIGHT) / (lambda0 * RP_NM_TO_M);
      double Bijhc_4PI = hc_4PI * L[3] * L[6] * L[2];
      double twohnu3_c2 = L[4] / L[5];
      int pol = L[8] != 0.0, zoff = (int) L[10], nc = (int) L[11];
      flags |= 1;
      if (pol) flags |= 2;
      for (k = 0; k < N; k++) {
        if (!(Mn[k] > 0.0)) continue;
        double adamp = L[4] * (lambda0 * RP_NM_TO_M) / (4.0*RP_PI * Mvb[k]);
        double v = (lambda/lambda0 - 1.0) * RP_CLIGHT/Mvb[k];
        double sv, phi, phi_Q = 0, phi_U = 0, phi_V = 0;
        if (moving) { if (to_obs) v += (muz * vel[k]) / Mvb[k]; else v -= (muz * vel[k]) / Mvb[k]; }
        sv = 1.0 / (RP_SQRTPI * Mvb[k]);
        if (pol) {
          double sin2_gamma = 1.0 - cos_gamma[k]*cos_gamma[k];
          double vB = (RP_LARMOR * lambda0) * B[k] / Mvb[k];
          double sign = to_obs ? 1.0 : -1.0, phi_sm = 0, phi_pi = 0, phi_sp = 0, F;
          for (nz = 0; nz < nc; nz++) {
            double H = rp_voigt_humlicek(adamp, v - zshift[zoff+nz]*vB, &F);
            switch (zq[zoff+nz]) {
            case -1: phi_sm += zstrength[zoff+nz] * H; break;
            case  0: phi_pi += zstrength[zoff+nz] * H; break;
            case  1: phi_sp += zstrength[zoff+nz] * H;
            }
          }
          double phi_sigma = phi_sp + phi_sm, phi_delta = 0.5*phi_pi - 0.25*phi_sigma;
          phi   = (phi_delta*sin2_gamma + 0.5*phi_sigma) * sv;
          phi_Q = sign * phi_delta * sin2_gamma * cos_2chi[k] * sv;
          phi_U = phi_delta * sin2_gamma * sin_2chi[k] * sv;
          phi_V = sign * 0.5*(phi_sp - phi_sm) * cos_gamma[k] * sv;
        } else
          phi = rp_voigt_armstrong(adamp, v) * sv;
        double kT = 1.0 / (RP_KBOLTZMANN * T[k]);
        double ni_gi = Mn[k] * exp(-L[1] * kT) / Mpf[k];
        double nj_gj = ni_gi * exp(-hc_la * kT);
        double chi_l = Bijhc_4PI * (ni_gi - nj_gj), eta_l = Bijhc_4PI * twohnu3_c2 * nj_gj;
        chi[k] += chi_l * phi;
        eta[k] += eta_l * phi;
        if (pol) {
          chi[N+k] += chi_l * phi_Q; chi[2*N+k] += chi_l * phi_U; chi[3*N+k] += chi_l * phi_V;
          eta[N+k] += eta_l * phi_Q; eta[2*N+k] += eta_l * phi_U; eta[3*N+k] += eta_l * phi_V;
        }
      }
    }
  }
  return flags;
}

/* passive_bb, rh/metal.c:174-344: bound-bound lines of PASSIVE model atoms in the background (unpolarised,
   VoigtArmstrong with the Damping() output supplied per line, Gaussian when line->Voigt is off).
   plines [nline][8]: lambda0, qwing, Bij, Bji, Aji, Voigt, Ncomponent, comp_off;  comp: c_shift / c_fraction;
   pcol [nline][4][N]: n_i, n_j, vbroad, adamp.  chi, eta [N].  Returns hasline. */
int rp_passive_bb(int N, int nline, const double *plines, const double *c_shift, const double *c_fraction,
                  double vmicro_char, double lambda, double muz, int moving, int to_obs,
                  const double *vel, const double *pcol, double *chi, double *eta)
{
  const double hc_4PI = (RP_HPLANCK * RP_CLIGHT) / (4.0 * RP_PI);
  int hasline = 0, n, nc, k;
  for (k = 0; k < N; k++) { chi[k] = 0.0; eta[k] = 0.0; }
  for (n = 0; n < nline; n++) {
    const double *L = plines + n*8;
    const double *ni = pcol + (size_t) n*4*N, *nj = ni + N, *vbroad = ni + 2*N, *adamp = ni + 3*N;
    double lambda0 = L[0], dlambda = lambda0 * L[1] * (vmicro_char / RP_CLIGHT);
    if (!(fabs(lambda - lambda0) <= dlambda)) continue;
    hasline = 1;
    double gij = L[3] / L[2], twohnu3_c2 = L[4] / L[3];
    int ncomp = (int) L[6], off = (int) L[7];
    for (nc = 0; nc < ncomp; nc++) {
      for (k = 0; k < N; k++) {
        double phi, Vij;
        double v = (lambda - lambda0 - c_shift[off+nc]) * RP_CLIGHT / (lambda0 * vbroad[k]);
        if (moving) { if (to_obs) v += (muz * vel[k]) / vbroad[k]; else v -= (muz * vel[k]) / vbroad[k]; }
        if (L[5] != 0.0) phi = rp_voigt_armstrong(adamp[k], v) * c_fraction[off+nc];
        else             phi = exp(-(v*v));
        Vij = hc_4PI * L[2] * phi / (RP_SQRTPI*vbroad[k]);
        chi[k] += Vij * (ni[k] - gij * nj[k]);
        eta[k] += twohnu3_c2 * gij * Vij * nj[k];
      }
    }
  }
  return hasline;
}
