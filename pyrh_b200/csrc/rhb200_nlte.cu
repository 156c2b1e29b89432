// rhb200_nlte.cu -- NLTE (MALI) machinery on the device: atoms, CRD, unpolarised radiation.
//
// Reference (all file:line relative to the reference root):
//   Iterate / solveSpectrum / updatePopulations   rh/iterate.c:48-253, rh/statequil.c:177-216
//   Formal (scalar branches)                      rh/rhf1d/formal.c:44-346
//   Opacity                                       rh/opacity.c:64-390
//   addtoCoupling / addtoGamma / addtoRates       rh/fillgamma.c:251-332, :82-246, :375-461
//   statEquil, SolveLinearEq/LUdecomp/LUbacksubst rh/statequil.c:40-103, rh/ludcmp.c:36-177
//   NgInit / Accelerate / MaxChange               rh/accelerate.c:34-147, rh/maxchange.c:32-50
//
// B200 design.  The reference walks wavelengths one at a time and adds
// every ray's contribution into Gamma and the rates as it goes.  Here one MALI iteration is
//   (1) opacity   : one thread per (column, ray, depth)      -> chi, S
//   (2) rays      : one thread per (column, ray), sequential in depth (Bezier3 / Feautrier) -> I, Psi
//   (3) J         : one thread per (column, lambda, depth), rays summed in the reference's order
//   (4) Gamma     : one thread per (column, transition, depth): walks the transition's own
//                   wavelengths and rays IN THE REFERENCE'S ORDER and keeps the two Gamma entries
//                   and the two rates of that transition in registers.  A Gamma pair {ij, ji} is
//                   only ever touched while its own transition is processed (fillgamma.c:139-187),
//                   so the sequence of floating-point additions into every element is exactly the
//                   reference's: results are bit-identical and run-to-run deterministic, with no
//                   atomics and no cross-thread reduction.
//   (5) statEquil : one thread per (column, atom, depth): Crout LU with implicit-scaling partial
//                   pivoting + one refinement step, same pivot choices as ludcmp.c
//   (6) Ng        : one block per (column, atom); every normal-equation entry is summed by one
//                   thread in the reference's k-order
// Columns that reached ITER_LIMIT are frozen (per-column `active` flag), like separate rhf1d calls.
#include <algorithm>
#include <chrono>
#include <vector>
#include "rhb200_common.cuh"
#include "rhb200_bezier.cuh"
#include "rhb200_feautrier.cuh"
#include "rhb200_lu.cuh"
#include "rhb200_piecewise.cuh"
#include "rhb200_voigt.cuh"

namespace {

enum { TR_ATOM = RHB200_TR_ATOM, TR_TYPE = RHB200_TR_TYPE, TR_I = RHB200_TR_I, TR_J = RHB200_TR_J,
       TR_NBLUE = RHB200_TR_NBLUE, TR_NLAMBDA = RHB200_TR_NLAMBDA, TR_AJI = RHB200_TR_AJI,
       TR_BJI = RHB200_TR_BJI, TR_BIJ = RHB200_TR_BIJ, TR_ISOFRAC = RHB200_TR_ISOFRAC,
       TR_WOFF = RHB200_TR_WOFF, TR_PHIROW = RHB200_TR_PHIROW, TR_LINEIDX = RHB200_TR_LINEIDX,
       TR_LAMBDA0 = RHB200_TR_LAMBDA0,
       TR_NFIELD = RHB200_TR_NFIELD };

struct Plan {            // device copy of the shared problem structure
  int Nspect, Nrays, Ndep, Natom, Ntrans, nas, nray, nlev, ngam, nphirow, nline, bc_top, bc_bottom, solver;
  int ns_lo, ns_hi, add_C;   // wavelength shard [ns_lo, ns_hi) of this rank; add_C: this rank adds the collisional part
  const double *lambda, *muz, *wmu, *trans, *tr_lambda, *tr_wlambda, *tr_alpha;
  const int *atom_nlevel, *lev_off, *gam_off, *as_first, *as_trans, *angle_dep, *ray_off,
            *ray_ns, *ray_mu, *ray_dir, *prow_tr, *line_tr;
  // fixed partition of every transition's wavelengths into segments (deterministic two-stage rate accumulation)
  int nseg; const int *seg_tr, *seg_lo, *seg_hi, *tr_seg0;   // [nseg] transition, [lo, hi) wavelengths; [Ntrans+1] first segment
  // atom-major rate accumulation (default mode): chunks of 16 wavelengths per atom; every transition the atom has in a
  // chunk owns a slot of partial sums; as_slot maps an active-set entry to its slot
  int naseg, nslot, aseg_maxslot; const int *aseg_atom, *aseg_lo, *aseg_hi, *aseg_slot0, *as_slot, *tr_slot0, *tr_slots;
  // FULL_STOKES formal solution with polarizable ACTIVE lines (after adjustStokesMode(), zeeman.c:303-345)
  int stokes;                                   // input.StokesMode == FULL_STOKES for the passes run now
  int stokes_prof;                              // input.StokesMode > FIELD_FREE when Profile() ran: Zeeman profiles (profile.c:112)
  int stokes_solver;                            // S_INTERPOLATION_STOKES
  const int *line_pol, *line_zoff, *zq;         // [nline] line->polarizable, [nline+1] slices of the Zeeman pattern tables
  const double *zshift, *zstrength;
  const int *pol_as, *pol_c;                    // [Nspect] containsPolarized(as) (with StokesMode FULL_STOKES), backgrflags.ispolarized
  // angle-averaged PRD (scatter.c:51-290, redistribute.c:38-106): the PRD lines, their rows in rho, containsPRDline(as)
  int nprd, nrho;
  const int *prd_tr, *prd_roff, *tr_prd, *prd_ns;
  const int *ns_mask;                           // solveSpectrum(.., redistribute = TRUE): only these wavelengths (or NULL: all)
};

struct Cols {            // device per-column arrays
  const double *T, *height, *nstar, *ntotal, *C, *chi_c, *eta_c, *sca_c, *adamp, *vbroad, *vel;
  double *phi, *wphi;
  double *n, *J, *Gamma, *Rij, *Rji, *gw, *chi, *S, *I, *Psi, *scr, *dJ, *Iem;
  double *part;            // [ncol][nseg][4][Ndep] partial {Gij, Gji, Rij, Rji} of the segments
  // FULL_STOKES passes: atmos.B [ncol][Ndep]; cos_gamma, cos_2chi, sin_2chi [ncol][Nrays][3][Ndep] (Bproject, project.c);
  // phi_Q/U/V [3][ncol][nphirow][Ndep]; background chi_c / eta_c Q,U,V [ncol][Nspect][3][Ndep];
  // per ray chi_Q,U,V and S_Q,U,V [ncol][nray][3][Ndep]; emergent Q,U,V [ncol][nray][3]
  const double *B, *bproj, *chi_cQ, *eta_cQ;
  double *phiQ, *chiQ, *SQ, *IemQ;
  double *rho;             // line->rho_prd [ncol][nrho][Ndep] of the PRD lines (profile ratio of emission to absorption)
  const double *Qelast;    // line->Qelast [ncol][nline][Ndep] (Damping(), broad.c:305-308)
  double *IQ;              // Stokes Q, U, V along every FULL_STOKES ray [ncol][nray][3][Ndep] (Stokes I_eff, fillgamma.c:106-129)
  const int *active;
};

// ---- (0) wavelength-integration weights and g_ij per active-set entry (opacity.c:214-217, 233-243):
//      they depend only on nstar, T, wphi, so they are evaluated once per column
__global__ void __launch_bounds__(128)
nlte_setup_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nas * N) return;
  const int k = (int) (t % N);
  const size_t ce = t / N;
  const int e = (int) (ce % P.nas), col = (int) (ce / P.nas);
  // wavelength of this entry: find ns by scanning is avoided: entry -> ns table is not needed,
  // la follows from the transition's position; we store ns in as_ns (packed after as_trans)
  const int tr_id = P.as_trans[e], ns = P.as_trans[P.nas + e];
  const double *tr = P.trans + (size_t) tr_id * TR_NFIELD;
  const int la = ns - (int) tr[TR_NBLUE], a = (int) tr[TR_ATOM];
  const double hc = RH_HPLANCK * RH_CLIGHT, fourPI = 4.0 * RH_PI, hc_4PI = hc / fourPI;
  const double hc_k = hc / (RH_KBOLTZMANN * RH_NM_TO_M);
  double g, w;
  if (tr[TR_TYPE] == 0.0) {
    g = tr[TR_BJI] / tr[TR_BIJ];
    if (P.nprd > 0 && P.tr_prd[tr_id] >= 0)      // PRD correction to the emission profile, opacity.c:207-217
      g *= C.rho[((size_t) col * P.nrho + P.prd_roff[P.tr_prd[tr_id]] + la) * N + k];
    const double wlambda = P.tr_wlambda[(int) tr[TR_WOFF] + la];
    w = wlambda * C.wphi[((size_t) col * P.nline + (int) tr[TR_LINEIDX]) * N + k] / hc_4PI;
  } else {
    const double lc = P.tr_lambda[(int) tr[TR_WOFF] + la], wlambda = P.tr_wlambda[(int) tr[TR_WOFF] + la];
    const double *ns_i = C.nstar + ((size_t) col * P.nlev + P.lev_off[a] + (int) tr[TR_I]) * N;
    const double *ns_j = C.nstar + ((size_t) col * P.nlev + P.lev_off[a] + (int) tr[TR_J]) * N;
    g = ns_i[k] / ns_j[k] * rhm::rh_exp(-hc_k / (lc * C.T[(size_t) col * N + k]));
    w = fourPI/RH_HPLANCK * (wlambda/lc);
  }
  double *o = C.gw + (((size_t) col * P.nas + e) * 2) * N + k;
  o[0] = g; o[N] = w;
}

// ---- Profile() of un-polarised single-component lines in a moving atmosphere (profile.c:311-323):
//      one thread per (column, profile row, depth); row = (line, la, mu, to_obs)
__global__ void __launch_bounds__(128)
nlte_profile_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nphirow * N) return;
  const int k = (int) (t % N);
  const size_t cr = t / N;
  const int row = (int) (cr % P.nphirow), col = (int) (cr / P.nphirow);
  const double *tr = P.trans + (size_t) P.prow_tr[row] * TR_NFIELD;
  const int lamu = row - (int) tr[TR_PHIROW];
  const int to_obs = lamu & 1, mu = (lamu >> 1) % P.Nrays, la = (lamu >> 1) / P.Nrays;
  const int a = (int) tr[TR_ATOM], li = (int) tr[TR_LINEIDX];
  const double lambda0 = tr[TR_LAMBDA0], lam = P.tr_lambda[(int) tr[TR_WOFF] + la];
  const double vbroad = C.vbroad[((size_t) col * P.Natom + a) * N + k];
  const double v = (lam - lambda0 - 0.0) * RH_CLIGHT / (vbroad * lambda0);            // profile.c:196-197
  const double v_los = (P.muz[mu] * C.vel[(size_t) col * N + k]) / vbroad;               // :188-190
  const double sign = to_obs ? 1.0 : -1.0;
  const double vk = v + sign * v_los;
  const double adamp = C.adamp[((size_t) col * P.nline + li) * N + k];
  if (P.stokes_prof && P.line_pol[li]) {
    // profile.c:112-116, 174-184, 239-305: Zeeman components through Voigt(.., HUMLICEK), one isotope component
    const double Larmor = (RH_Q_ELECTRON / (4.0*RH_PI*RH_M_ELECTRON)) * (lambda0*RH_NM_TO_M);
    const double vB = Larmor * C.B[(size_t) col * N + k] / vbroad;
    const double sv = 1.0 / (RH_SQRTPI * vbroad);
    const double *bp = C.bproj + (((size_t) col * P.Nrays + mu) * 3) * N + k;
    const double cos_gamma = bp[0], cos_2chi = bp[N], sin_2chi = bp[2*(size_t) N];
    const double sin2_gamma = 1.0 - cos_gamma*cos_gamma;
    double phi_sm = 0.0, phi_pi = 0.0, phi_sp = 0.0;
    for (int nz = P.line_zoff[li]; nz < P.line_zoff[li+1]; nz++) {
      const double H = rhv::humlicek_H(adamp, vk - P.zshift[nz]*vB);
      const int q = P.zq[nz];
      if (q == -1)     phi_sm += P.zstrength[nz] * H;
      else if (q == 0) phi_pi += P.zstrength[nz] * H;
      else if (q == 1) phi_sp += P.zstrength[nz] * H;
    }
    const double phi_sigma = (phi_sp + phi_sm) * 1.0;
    const double phi_delta = 0.5*phi_pi * 1.0 - 0.25*phi_sigma;
    const size_t plane = (size_t) ncol * P.nphirow * N;
    C.phi[t]            = 0.0 + (phi_delta*sin2_gamma + 0.5*phi_sigma) * sv;
    C.phiQ[t]           = 0.0 + sign * phi_delta * sin2_gamma * cos_2chi * sv;
    C.phiQ[plane + t]   = 0.0 + phi_delta * sin2_gamma * sin_2chi * sv;
    C.phiQ[2*plane + t] = 0.0 + sign * 0.5*(phi_sp - phi_sm) * cos_gamma * sv;
    return;
  }
  const double H = rhv::voigt_armstrong(adamp, vk);
  C.phi[t] = 0.0 + H * 1.0 / (RH_SQRTPI * vbroad);                                      // :317-318
}

// Bproject() for every ray of the plan (rhf1d/project.c:38-80, geometry of rhf1d/anglequad.c: mux = sqrt(1 - muz^2),
// muy = 0) from the pyrh rows B [G], gamma, chi: one thread per (column, depth)
__global__ void __launch_bounds__(128)
nlte_bproject_kernel(Plan P, int ncol, int nrow, const double *__restrict__ in, double *__restrict__ B, double *__restrict__ bproj)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * N) return;
  const int col = (int) (t / N), k = (int) (t % N);
  const double *a = in + (size_t) col * nrow * N + k;
  B[t] = a[(size_t) 5 * N] / 1e4;                                                         // pyrh_compute1dray.c:266
  const double gamma_B = a[(size_t) 6 * N], chi_B = a[(size_t) 7 * N];
  for (int mu = 0; mu < P.Nrays; mu++) {
    const double muz = P.muz[mu];
    double cg, c2, s2;
    if (muz == 1.0) {                                                                     // project.c:52-58
      cg = rhm::rh_cos(gamma_B);
      c2 = rhm::rh_cos(2.0 * chi_B);
      s2 = rhm::rh_sin(2.0 * chi_B);
    } else {                                                                              // project.c:60-77
      const double mux = sqrt(1.0 - muz*muz), muy = 0.0;
      const double csc_theta = 1.0 / sqrt(1.0 - muz*muz);
      const double sin_gamma = rhm::rh_sin(gamma_B);
      const double bx = sin_gamma * rhm::rh_cos(chi_B);
      const double by = sin_gamma * rhm::rh_sin(chi_B);
      const double bz = rhm::rh_cos(gamma_B);
      const double b3 = mux*bx + muy*by + muz*bz;
      const double b1 = csc_theta * (bz - muz*b3);
      const double b2 = csc_theta * (muy*bx - mux*by);
      cg = b3;
      c2 = (b1*b1 - b2*b2) / (1.0 - b3*b3);
      s2 = 2.0 * b1*b2 / (1.0 - b3*b3);
    }
    double *o = bproj + (((size_t) col * P.Nrays + mu) * 3) * N + k;
    o[0] = cg; o[N] = c2; o[2*(size_t) N] = s2;
  }
}

// wphi[k] = 1 / sum_{la,mu,dir} phi wlambda 0.5 wmu, summed in the reference's order (profile.c:320,358)
__global__ void __launch_bounds__(64)
nlte_wphi_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nline * N) return;
  const int k = (int) (t % N);
  const size_t cl = t / N;
  const int li = (int) (cl % P.nline), col = (int) (cl / P.nline);
  const double *tr = P.trans + (size_t) P.line_tr[li] * TR_NFIELD;
  const int Nla = (int) tr[TR_NLAMBDA], row0 = (int) tr[TR_PHIROW];
  double w = 0.0;
  for (int la = 0; la < Nla; la++)
    for (int mu = 0; mu < P.Nrays; mu++) {
      const double wlamu = P.tr_wlambda[(int) tr[TR_WOFF] + la] * 0.5*P.wmu[mu];
      for (int to_obs = 0; to_obs <= 1; to_obs++)
        w += C.phi[((size_t) col * P.nphirow + row0 + 2*(P.Nrays*la + mu) + to_obs) * N + k] * wlamu;
    }
  C.wphi[t] = 1.0 / w;
}

// V_ij of active-set entry e at (ray, depth): Bij hc/4pi phi for lines (opacity.c:188-193),
// alpha(lambda) for continua (:236)
__device__ __forceinline__ double vij_of(const Plan &P, const Cols &C, int col, const double *tr, int ns,
                                         int mu, int dir, int k)
{
  const int la = ns - (int) tr[TR_NBLUE];
  if (tr[TR_TYPE] == 0.0) {
    const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
    const double Bijxhc_4PI = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC];
    const int lamu = 2*(P.Nrays*la + mu) + dir;
    return Bijxhc_4PI * __ldg(C.phi + ((size_t) col * P.nphirow + (int) tr[TR_PHIROW] + lamu) * P.Ndep + k);
  }
  return P.tr_alpha[(int) tr[TR_WOFF] + la];
}
__device__ __forceinline__ double twohnu3_of(const Plan &P, const double *tr, int ns)
{
  if (tr[TR_TYPE] == 0.0) return tr[TR_AJI] / tr[TR_BJI];
  const double nm3 = RH_NM_TO_M*RH_NM_TO_M*RH_NM_TO_M;
  const double twohc = 2.0*(RH_HPLANCK * RH_CLIGHT) / nm3;
  const double lc = P.tr_lambda[(int) tr[TR_WOFF] + ns - (int) tr[TR_NBLUE]];
  return twohc / (lc*lc*lc);
}

// ---- (1) Opacity() + the chi/S assembly of Formal(): one wavelength per block, threads over (column,
// depth); one thread evaluates all rays of its wavelength (batches of NLTE_RB) so that the
// ray-independent factors (n_i - g n_j, thn g, n_j, background, Jdag) are loaded once, and the walk over
// the active set is uniform across the block.  Products keep the reference's association:
// V*(n_i - g n_j), ((thn*g)*V)*n_j (opacity.c:252-260).
// blocks of 64 threads per SM the rate kernels are compiled for.  Measured per launch, 256 columns (configs[4] sample /
// configs[3]): 12 -> 5.14 / 12.0 ms (80 registers, 850 B of spills), 10 -> 3.88 / 8.97, 8 -> 3.42 / 7.76 (128 registers, still
// spilling), 6 and 5 -> 2.79 / 6.93 (149 / 126 registers, no spills), 4 -> 5.78 / 12.7
#ifndef NLTE_GAMMA_MINB
#define NLTE_GAMMA_MINB 6
#endif
#ifndef NLTE_RB
#define NLTE_RB 6
#endif
// blocks/SM of the opacity kernel, measured per launch at 256 columns (configs[4] sample / configs[3]): 3 -> 1.51 / 1.40 ms,
// 4 -> 1.24 / 1.10, 5 -> 1.13 / 0.97 (96 registers), 6 -> 1.41 / 1.03
#ifndef NLTE_OPAC_MINB
#define NLTE_OPAC_MINB 5
#endif
__global__ void __launch_bounds__(128, NLTE_OPAC_MINB)
nlte_opacity_kernel(Plan P, Cols C, int ncol)
{
  // consecutive blocks take consecutive wavelengths of the same (column, depth) chunk: their rays are
  // adjacent in memory, which keeps the DRAM pages of chi/S/phi open
  const int N = P.Ndep, ns = (int) (blockIdx.x % P.Nspect);
  const size_t t = (size_t) (blockIdx.x / P.Nspect) * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * N) return;
  const int k = (int) (t % N), col = (int) (t / N);
  if (!C.active[col]) return;
  if (ns < P.ns_lo || ns >= P.ns_hi) return;           // another rank's wavelength
  if (P.ns_mask && !P.ns_mask[ns]) return;
  const int first = P.as_first[ns], nact = P.as_first[ns+1] - first;
  const double *ncol_ = C.n + (size_t) col * P.nlev * N;
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  const size_t lk = ((size_t) col * P.Nspect + ns) * N + k;
  const double chi_c = __ldg(C.chi_c + lk), eta_c = __ldg(C.eta_c + lk);
  const double scaJ = __ldg(C.sca_c + lk) * C.J[lk];                         // J still holds Jdag here
  const int r_end = P.ray_off[ns+1];
  for (int r0 = P.ray_off[ns]; r0 < r_end; r0 += NLTE_RB) {
    const int nb = r_end - r0 < NLTE_RB ? r_end - r0 : NLTE_RB;
    double as_chi[NLTE_RB], as_eta[NLTE_RB], eta_atom[NLTE_RB];
    int lamu[NLTE_RB];
#pragma unroll
    for (int q = 0; q < NLTE_RB; q++) {
      as_chi[q] = as_eta[q] = eta_atom[q] = 0.0;
      lamu[q] = (q < nb) ? 2*P.ray_mu[r0 + q] + P.ray_dir[r0 + q] : 0;
    }
    int cur_atom = -1;
    for (int n = 0; n < nact; n++) {
      const double *tr = P.trans + (size_t) P.as_trans[first+n] * TR_NFIELD;
      const int a = (int) tr[TR_ATOM], la = ns - (int) tr[TR_NBLUE];
      if (a != cur_atom) {                      // as->eta += atom->rhth.eta, opacity.c:375-380
        if (cur_atom >= 0) {
#pragma unroll
          for (int q = 0; q < NLTE_RB; q++) as_eta[q] += eta_atom[q];
        }
#pragma unroll
        for (int q = 0; q < NLTE_RB; q++) eta_atom[q] = 0.0;
        cur_atom = a;
      }
      const double thn = twohnu3_of(P, tr, ns);
      if (thn == 0.0) continue;                 // opacity.c:252
      const double g = C.gw[(((size_t) col * P.nas + first + n) * 2) * N + k];
      const double n_i = ncol_[(size_t)(P.lev_off[a] + (int) tr[TR_I]) * N + k];
      const double n_j = ncol_[(size_t)(P.lev_off[a] + (int) tr[TR_J]) * N + k];
      const double diff = n_i - g*n_j, tg = thn * g;
      if (tr[TR_TYPE] == 0.0) {                 // line: V = Bij hc/4pi phi(ray), opacity.c:188-193
        const double c = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC];
        const double *ph = C.phi + ((size_t) col * P.nphirow + (int) tr[TR_PHIROW] + 2*P.Nrays*la) * N + k;
        double V[NLTE_RB];
#pragma unroll
        for (int q = 0; q < NLTE_RB; q++) V[q] = (q < nb) ? c * __ldg(ph + (size_t) lamu[q] * N) : 0.0;
#pragma unroll
        for (int q = 0; q < NLTE_RB; q++) { as_chi[q] += V[q] * diff; eta_atom[q] += tg * V[q] * n_j; }
      } else {                                  // continuum: V = alpha(lambda), opacity.c:236
        const double V = P.tr_alpha[(int) tr[TR_WOFF] + la];
#pragma unroll
        for (int q = 0; q < NLTE_RB; q++) { as_chi[q] += V * diff; eta_atom[q] += tg * V * n_j; }
      }
    }
#pragma unroll
    for (int q = 0; q < NLTE_RB; q++) {
      if (q < nb) {
        if (cur_atom >= 0) as_eta[q] += eta_atom[q];
        const double chi = as_chi[q] + chi_c;                                // formal.c:178-182, 293-296
        const double S = (as_eta[q] + eta_c + scaJ) / chi;
        const size_t rk = ((size_t) col * P.nray + r0 + q) * N + k;
        C.chi[rk] = chi;
        C.S[rk] = S;
      }
    }
  }
}

// ---- (1b) FULL_STOKES passes: Q, U, V of Opacity() (opacity.c:262-296, 375-380) and of Formal()'s source vector and
//      StokesK numerators (formal.c:184-206, stokesopac.c:49-60) at the wavelengths that hold a polarizable ACTIVE
//      line or a polarised background line.  One thread per (column, wavelength, depth), after nlte_opacity_kernel
//      (reads its chi_I).  chiQ [ray][3][N] = K'[0][1..3] numerators, SQ [ray][3][N] = S_Q,U,V.
__global__ void __launch_bounds__(128)
nlte_opacity_quv_kernel(Plan P, Cols C, int ncol)
{
  const int N = P.Ndep, ns = (int) (blockIdx.x % P.Nspect);
  const size_t t = (size_t) (blockIdx.x / P.Nspect) * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * N) return;
  const int k = (int) (t % N), col = (int) (t / N);
  if (!C.active[col]) return;
  if (ns < P.ns_lo || ns >= P.ns_hi) return;
  if (P.ns_mask && !P.ns_mask[ns]) return;
  const int pol_as = P.pol_as[ns], pol_c = P.pol_c[ns];
  if (!pol_as && !pol_c) return;
  const int first = P.as_first[ns], nact = P.as_first[ns+1] - first;
  const double *ncol_ = C.n + (size_t) col * P.nlev * N;
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  const size_t plane = (size_t) ncol * P.nphirow * N;
  double cq[3] = {0.0, 0.0, 0.0}, eq[3] = {0.0, 0.0, 0.0};
  if (pol_c) {
    const double *bc = C.chi_cQ + (((size_t) col * P.Nspect + ns) * 3) * N + k, *be = C.eta_cQ + (((size_t) col * P.Nspect + ns) * 3) * N + k;
    for (int s = 0; s < 3; s++) { cq[s] = bc[(size_t) s * N]; eq[s] = be[(size_t) s * N]; }
  }
  for (int r = P.ray_off[ns]; r < P.ray_off[ns+1]; r++) {
    const int lamu = 2*P.ray_mu[r] + P.ray_dir[r];
    double chi_q[3] = {0.0, 0.0, 0.0}, eta_q[3] = {0.0, 0.0, 0.0}, eta_atom[3] = {0.0, 0.0, 0.0};
    int cur_atom = -1;
    if (pol_as)
      for (int n = 0; n < nact; n++) {
        const double *tr = P.trans + (size_t) P.as_trans[first+n] * TR_NFIELD;
        const int a = (int) tr[TR_ATOM];
        if (a != cur_atom) {                      // as->eta += atom->rhth.eta for all four records, opacity.c:375-380
          if (cur_atom >= 0) for (int s = 0; s < 3; s++) eta_q[s] += eta_atom[s];
          for (int s = 0; s < 3; s++) eta_atom[s] = 0.0;
          cur_atom = a;
        }
        if (tr[TR_TYPE] != 0.0 || !P.line_pol[(int) tr[TR_LINEIDX]]) continue;
        const int la = ns - (int) tr[TR_NBLUE];
        const double twohnu3_c2 = tr[TR_AJI] / tr[TR_BJI];
        if (twohnu3_c2 == 0.0) continue;          // opacity.c:252
        const double g = C.gw[(((size_t) col * P.nas + first + n) * 2) * N + k];
        const double n_i = ncol_[(size_t)(P.lev_off[a] + (int) tr[TR_I]) * N + k];
        const double n_j = ncol_[(size_t)(P.lev_off[a] + (int) tr[TR_J]) * N + k];
        const double Bijxhc_4PI = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC];
        const double chi_l = Bijxhc_4PI * (n_i - g*n_j);                       // opacity.c:271-272
        const double eta_l = Bijxhc_4PI * twohnu3_c2 * g * n_j;                // :283-284
        const double *ph = C.phiQ + ((size_t) col * P.nphirow + (int) tr[TR_PHIROW] + 2*P.Nrays*la + lamu) * N + k;
        for (int s = 0; s < 3; s++) {
          const double f = ph[(size_t) s * plane];
          chi_q[s] += chi_l * f;
          eta_atom[s] += eta_l * f;
        }
      }
    if (cur_atom >= 0) for (int s = 0; s < 3; s++) eta_q[s] += eta_atom[s];
    const size_t rk = ((size_t) col * P.nray + r) * N + k;
    const double chi = C.chi[rk];
    double *oc = C.chiQ + ((size_t) col * P.nray + r) * 3 * N + k, *os = C.SQ + ((size_t) col * P.nray + r) * 3 * N + k;
    for (int s = 0; s < 3; s++) {
      double K = 0.0, S = 0.0;
      if (pol_as) { K = chi_q[s]; S += eta_q[s]; }          // stokesopac.c:49-52, formal.c:187-189
      if (pol_c)  { K += cq[s];   S += eq[s]; }             // :60-63, :192-194
      oc[(size_t) s * N] = K;
      os[(size_t) s * N] = S / chi;                         // formal.c:204-207
    }
  }
}

// ---- (2) formal solution of every ray: Piecewise_Bezier3_1D for angle-dependent wavelengths,
//      Feautrier otherwise (formal.c:157-309)
struct NlteFeauIO {
  const double *__restrict__ chi_, *__restrict__ S_, *__restrict__ h;
  double *P_, *Psi_, *scr; int ndep;
  __device__ __forceinline__ double chi(int k) const { return chi_[k]; }
  __device__ __forceinline__ double S(int k) const { return S_[k]; }
  __device__ __forceinline__ double z(int k) const { return h[k]; }
  __device__ __forceinline__ void putF(int k, double v) { scr[k] = v; }
  __device__ __forceinline__ void putZ(int k, double v) { scr[ndep + k] = v; }
  __device__ __forceinline__ double getF(int k) const { return scr[k]; }
  __device__ __forceinline__ double getZ(int k) const { return scr[ndep + k]; }
  __device__ __forceinline__ void storeP(int k, double v) { P_[k] = v; }
  __device__ __forceinline__ void storePsi(int k, double v) { Psi_[k] = v / chi_[k]; }     // Psi[k] /= chi[k], formal.c:300-301
  __device__ __forceinline__ bool wantPsi() const { return Psi_ != nullptr; }
};

// FULL_STOKES rays (formal.c:184-217): Piece_Stokes_Bezier3_1D / Piece_Stokes_1D on chi_I, S[4], K' numerators
struct NlteStokesIO {
  const double *__restrict__ chi_, *__restrict__ S_, *__restrict__ SQ_, *__restrict__ q_;
  double *I_, *Psi_, *IemQ_, *IQ_;
  int ndep, kem;
  __device__ __forceinline__ double chi(int k) const { return chi_[k]; }
  __device__ __forceinline__ void K(int k, double x[3]) const {   // StokesK, stokesopac.c:72-77
    const double c = chi_[k];
    x[0] = q_[k] / c; x[1] = q_[ndep + k] / c; x[2] = q_[2*ndep + k] / c;
  }
  __device__ __forceinline__ void S(int k, double s[4]) const {
    s[0] = S_[k]; s[1] = SQ_[k]; s[2] = SQ_[ndep+k]; s[3] = SQ_[2*ndep+k];
  }
  __device__ __forceinline__ void storeI(int k, const double I[4]) {
    I_[k] = I[0];
    if (IQ_) { IQ_[k] = I[1]; IQ_[ndep + k] = I[2]; IQ_[2*ndep + k] = I[3]; }
    if (k == kem) { IemQ_[0] = I[1]; IemQ_[1] = I[2]; IemQ_[2] = I[3]; }
  }
  __device__ __forceinline__ void storePsi(int k, double p) { if (Psi_) Psi_[k] = p / chi_[k]; }     // formal.c:247-248
  __device__ __forceinline__ void prefetch(int, int) const {}
};
__global__ void __launch_bounds__(128, 4)
nlte_ray_stokes_kernel(Plan P, Cols C, int ncol, int eval_operator)
{
  const size_t cr = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (cr >= (size_t) ncol * P.nray) return;
  const int r = (int) (cr % P.nray), col = (int) (cr / P.nray), N = P.Ndep;
  if (!C.active[col]) return;
  const int ns = P.ray_ns[r], mu = P.ray_mu[r], dir = P.ray_dir[r];
  if (ns < P.ns_lo || ns >= P.ns_hi) return;
  if (P.ns_mask && !P.ns_mask[ns]) return;
  if (!(P.pol_as[ns] || P.pol_c[ns])) return;              // solveStokes, formal.c:94-95
  const double *h = C.height + (size_t) col * N, *T = C.T + (size_t) col * N;
  NlteStokesIO io{C.chi + cr * N, C.S + cr * N, C.SQ + cr * 3 * N, C.chiQ + cr * 3 * N, C.I + cr * N,
                  eval_operator ? C.Psi + cr * N : nullptr, C.IemQ + cr * 3, eval_operator ? C.IQ + cr * 3 * N : nullptr, N, 0};
  if (P.stokes_solver == RHB200_DELO_PARABOLIC)
    rhp::stokes_parabolic_ray(io, N, h, P.muz[mu], dir, P.bc_top, P.bc_bottom, T, P.lambda[ns]);
  else
    rhd::delo_bezier3_ray(io, N, h, P.muz[mu], dir, P.bc_top, P.bc_bottom, T, P.lambda[ns]);
  C.Iem[cr] = C.I[cr * N];
}

// SOLVER is a template parameter so that each instantiation carries one solver's registers only
template <int SOLVER, int MINB>
__global__ void __launch_bounds__(128, MINB)
nlte_ray_kernel(Plan P, Cols C, int ncol, int eval_operator)
{
  const size_t cr = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (cr >= (size_t) ncol * P.nray) return;
  const int r = (int) (cr % P.nray), col = (int) (cr / P.nray), N = P.Ndep;
  if (!C.active[col]) return;
  const int ns = P.ray_ns[r], mu = P.ray_mu[r], dir = P.ray_dir[r];
  if (ns < P.ns_lo || ns >= P.ns_hi) return;
  if (P.ns_mask && !P.ns_mask[ns]) return;
  if (P.stokes && (P.pol_as[ns] || P.pol_c[ns])) return;   // nlte_ray_stokes_kernel
  const double *h = C.height + (size_t) col * N, *T = C.T + (size_t) col * N;
  double *Psi = eval_operator ? C.Psi + cr * N : nullptr;
  // every solver hands back Psi / chi (formal.c:247-248, 300-301): the rate kernels would otherwise divide once per
  // (transition, ray-point) instead of once per ray-point
  if (P.angle_dep[ns]) {
    if (SOLVER == RHB200_S_LINEAR || SOLVER == RHB200_S_PARABOLIC) {        // S_INTERPOLATION, formal.c:229-235
      if (SOLVER == RHB200_S_LINEAR)
        rhp::linear_ray(N, h, P.muz[mu], dir, P.bc_top, P.bc_bottom, T, P.lambda[ns], C.chi + cr * N,
                        C.S + cr * N, C.I + cr * N, Psi);
      else
        rhp::parabolic_ray(N, h, P.muz[mu], dir, P.bc_top, P.bc_bottom, T, P.lambda[ns], C.chi + cr * N,
                           C.S + cr * N, C.I + cr * N, Psi);
      if (Psi) { const double *chi = C.chi + cr * N; for (int k = 0; k < N; k++) Psi[k] = Psi[k] / chi[k]; }
    } else
      rhz::bezier3_ray_w(N, h, P.muz[mu], dir, P.bc_top, P.bc_bottom, T, P.lambda[ns], C.chi + cr * N,
                         C.S + cr * N, C.I + cr * N, Psi, true);
    C.Iem[cr] = C.I[cr * N];                         // spectrum.I[nspect][mu] = I[0] (formal.c:270)
  } else {
    NlteFeauIO io{C.chi + cr * N, C.S + cr * N, h, C.I + cr * N, Psi, C.scr + cr * 2 * N, N};
    C.Iem[cr] = rhf::feautrier_ray(io, N, P.muz[mu], P.bc_top, P.bc_bottom, T, P.lambda[ns]);   // formal.c:299
  }
}

// ---- (3) J = sum_rays wmu I in the reference's ray order; dJ = |1 - Jdag/J| (formal.c:252-256, 313-318)
__global__ void __launch_bounds__(128)
nlte_J_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;     // linear in memory: (column, wavelength, depth)
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.Nspect * N) return;
  const int k = (int) (t % N);
  const size_t cl = t / N;
  const int ns = (int) (cl % P.Nspect), col = (int) (cl / P.Nspect);
  if (!C.active[col]) return;
  if (ns < P.ns_lo || ns >= P.ns_hi) return;
  if (P.ns_mask && !P.ns_mask[ns]) return;
  const int ad = P.angle_dep[ns];
  double J = 0.0;
  const int r_end = P.ray_off[ns+1];
  for (int r0 = P.ray_off[ns]; r0 < r_end; r0 += NLTE_RB) {            // loads of a batch first, ordered sum after
    double Iv[NLTE_RB], wv[NLTE_RB];
#pragma unroll
    for (int q = 0; q < NLTE_RB; q++) {
      const bool on = r0 + q < r_end;
      Iv[q] = on ? __ldg(C.I + ((size_t) col * P.nray + r0 + q) * N + k) : 0.0;
      wv[q] = on ? P.wmu[P.ray_mu[r0 + q]] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < NLTE_RB; q++)
      if (r0 + q < r_end) J += (ad ? 0.5 * wv[q] : wv[q]) * Iv[q];
  }
  const double Jdag = C.J[t];
  C.J[t] = J;
  C.dJ[t] = fabs(1.0 - Jdag / J);
}

// ---- angle-averaged PRD.  GII: Gouttebroze's approximation with Uitenbroek's cross-redistribution form (giigen.c:87-147)
#define PRD_QCORE   2.0
#define PRD_QWING   4.0
#define PRD_QSPREAD 5.0
#define PRD_DQ      0.25
__device__ __forceinline__ double prd_gzero(double x) { return 1.0 / (fabs(x) + sqrt(x*x + 1.273239545)); }
__device__ __forceinline__ double prd_gii(double adamp, double waveratio, double q_emit, double q_abs)
{
  if (q_emit < 0.0) { q_emit = -q_emit; q_abs = -q_abs; }
  double pcore = 0.0, gii = 0.0;
  if (q_emit < PRD_QWING) {
    if (q_abs < -PRD_QWING || q_abs > q_emit + waveratio*PRD_QSPREAD) return gii;
    if (fabs(q_abs) <= q_emit) gii = prd_gzero(q_emit);
    else gii = rhm::rh_exp(q_emit*q_emit - q_abs*q_abs) * prd_gzero(q_abs);
    if (q_emit >= PRD_QCORE) {
      const double phicore = rhm::rh_exp(-(q_emit*q_emit));
      const double phiwing = adamp / (RH_SQRTPI * (adamp*adamp + q_emit*q_emit));
      pcore = phicore / (phicore + phiwing);
    }
  }
  if (q_emit >= PRD_QCORE) {
    const double aq_emit = waveratio * q_emit;
    if (q_emit >= PRD_QWING) {
      if (fabs(q_abs - aq_emit) > waveratio*PRD_QSPREAD) return gii;
      pcore = 0.0;
    }
    const double umin = fabs((q_abs - aq_emit) / (1.0 + waveratio));
    double giiwing = (1.0 + waveratio) * (1.0 - 2.0*umin*prd_gzero(umin)) * rhm::rh_exp(-(umin*umin));
    if (waveratio == 1.0) {
      const double epsilon = q_abs / aq_emit;
      giiwing *= (2.75 - (2.5 - 0.75*epsilon) * epsilon);
    } else {
      const double u1 = fabs((q_abs - aq_emit) / (waveratio - 1.0));
      giiwing -= fabs(1.0 - waveratio) * (1.0 - 2.0*u1*prd_gzero(u1)) * rhm::rh_exp(-(u1*u1));
    }
    giiwing = giiwing / (2.0 * waveratio * RH_SQRTPI);
    gii = pcore*gii + (1.0 - pcore)*giiwing;
  }
  return gii;
}

// PRDScatter() (scatter.c:51-290, LINEAR representation, no cross redistribution): one thread per (column, row of rho,
// depth).  drho [ncol]: MaxChange of the Ng structure of order 0 (maxchange.c:32-50), as the bit pattern of a double >= 0
__global__ void __launch_bounds__(128)
nlte_prd_scatter_kernel(Plan P, Cols C, int ncol, unsigned long long *__restrict__ drho)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nrho * N) return;
  const int k = (int) (t % N), row = (int) ((t / N) % P.nrho), col = (int) (t / ((size_t) N * P.nrho));
  if (!C.active[col]) return;
  int p = 0;
  while (row >= P.prd_roff[p+1]) p++;
  const int la = row - P.prd_roff[p], tid = P.prd_tr[p];
  const double *tr = P.trans + (size_t) tid * TR_NFIELD;
  const int a = (int) tr[TR_ATOM], i = (int) tr[TR_I], j = (int) tr[TR_J], Nl = P.atom_nlevel[a], li = (int) tr[TR_LINEIDX];
  const int Nblue = (int) tr[TR_NBLUE], Nla = (int) tr[TR_NLAMBDA];
  const double lambda0 = tr[TR_LAMBDA0], Bij = tr[TR_BIJ];
  const double *lam = P.tr_lambda + (int) tr[TR_WOFF];
  const double vbroad = C.vbroad[((size_t) col * P.Natom + a) * N + k];
  const double adamp = C.adamp[((size_t) col * P.nline + li) * N + k];      // = (Grad + Qelast) cDop / vbroad, scatter.c:118
  // total rate out of the upper level, scatter.c:122-137
  double Pj = C.Qelast[((size_t) col * P.nline + li) * N + k];
  const size_t gbase = ((size_t) col * P.ngam + P.gam_off[a]) * N;
  for (int ip = 0; ip < Nl; ip++) Pj += C.C[gbase + (size_t) (ip*Nl + j) * N + k];
  for (int t2 = 0; t2 < P.Ntrans; t2++) {        // the atom's lines (kr order), then its continua (kr order)
    const double *q = P.trans + (size_t) t2 * TR_NFIELD;
    if ((int) q[TR_ATOM] != a) continue;
    const size_t rk = ((size_t) col * P.Ntrans + t2) * N + k;
    if ((int) q[TR_J] == j) Pj += C.Rji[rk];
    if ((int) q[TR_I] == j) Pj += C.Rij[rk];
  }
  const double *n = C.n + ((size_t) col * P.nlev + P.lev_off[a]) * N + k;
  const double gamma = n[(size_t) i * N] / n[(size_t) j * N] * Bij / Pj;
  const double Jbar = C.Rij[((size_t) col * P.Ntrans + tid) * N + k] / Bij;
  const double *Jk = C.J + ((size_t) col * P.Nspect + Nblue) * N + k;        // J_k[la'] = Jk[la' * N]
#define QABS(l) ((lam[l] - lambda0) * RH_CLIGHT / (lambda0 * vbroad))
  const double q_emit = QABS(la);
  double q0, qN;                                                                // scatter.c:176-193 with waveratio = 1
  if (fabs(q_emit) < PRD_QCORE) { q0 = -PRD_QWING; qN = PRD_QWING; }
  else if (fabs(q_emit) < PRD_QWING) {
    if (q_emit > 0.0) { q0 = -PRD_QWING; qN = 1.0 * (q_emit + PRD_QSPREAD); }
    else              { q0 = 1.0 * (q_emit - PRD_QSPREAD); qN = PRD_QWING; }
  } else { q0 = 1.0 * (q_emit - PRD_QSPREAD); qN = 1.0 * (q_emit + PRD_QSPREAD); }
  const int Np = (int) ((qN - q0) / PRD_DQ) + 1;
  const double xmin = QABS(0), xmax = QABS(Nla - 1);
  int jt = 0;
  double qa0 = xmin, qa1 = QABS(1);
  double qp = q0, gnorm = 0.0, scatInt = 0.0;
  for (int lap = 0; lap < Np; lap++) {
    if (lap > 0) qp = qp + PRD_DQ;                                              // :195-196
    double Jv;                                                                  // Linear(.., hunt), linear.c:22-51
    if (qp <= xmin) Jv = Jk[0];
    else if (qp >= xmax) Jv = Jk[(size_t) (Nla - 1) * N];
    else {
      while (jt < Nla - 2 && qa1 <= qp) { jt++; qa0 = qa1; qa1 = QABS(jt + 1); }
      const double fx = (qa1 - qp) / (qa1 - qa0);
      Jv = fx*Jk[(size_t) jt * N] + (1 - fx)*Jk[(size_t) (jt + 1) * N];
    }
    double wq = PRD_DQ;                                                         // :231-236 (later assignments win)
    if (lap == 0) wq = 5.0/12.0 * PRD_DQ;
    if (lap == 1) wq = 13.0/12.0 * PRD_DQ;
    if (lap == Np-1) wq = 5.0/12.0 * PRD_DQ;
    if (lap == Np-2) wq = 13.0/12.0 * PRD_DQ;
    const double gii = prd_gii(adamp, 1.0, q_emit, qp) * wq;
    gnorm += gii;
    scatInt += Jv * gii;
  }
#undef QABS
  const double rho_new = 1.0 + gamma*(scatInt/gnorm - Jbar);                    // :109-113, 281
  const double rho_old = C.rho[t];
  C.rho[t] = rho_new;
  if (rho_new != 0.0) {
    const double d = fabs((rho_new - rho_old) / rho_new);
    atomicMax(drho + col, (unsigned long long) __double_as_longlong(d));
  }
}

// addtoRates(.., redistribute = TRUE) after zeroRates(TRUE) (fillgamma.c:337-461): the rates of the PRD lines alone,
// rays in the reference's order.  One thread per (column, PRD line, depth)
__global__ void __launch_bounds__(64)
nlte_prd_rates_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nprd * N) return;
  const int k = (int) (t % N), p = (int) ((t / N) % P.nprd), col = (int) (t / ((size_t) N * P.nprd));
  if (!C.active[col]) return;
  const int tid = P.prd_tr[p];
  const double *tr = P.trans + (size_t) tid * TR_NFIELD;
  const int Nblue = (int) tr[TR_NBLUE], Nla = (int) tr[TR_NLAMBDA];
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  const double c = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC], thn = tr[TR_AJI] / tr[TR_BJI];
  double Rij = 0.0, Rji = 0.0;
  for (int ns = Nblue; ns < Nblue + Nla; ns++) {
    const int first = P.as_first[ns], nact = P.as_first[ns+1] - first;
    int e = -1;
    for (int n = 0; n < nact; n++) if (P.as_trans[first+n] == tid) { e = first + n; break; }
    if (e < 0) continue;
    const double *gw = C.gw + (((size_t) col * P.nas + e) * 2) * N + k;
    const double g = gw[0], w = gw[N];
    const int ad = P.angle_dep[ns], la = ns - Nblue;
    const double *ph = C.phi + ((size_t) col * P.nphirow + (int) tr[TR_PHIROW] + 2*P.Nrays*la) * N + k;
    for (int r = P.ray_off[ns]; r < P.ray_off[ns+1]; r++) {
      const int mu = P.ray_mu[r];
      const double wmu = ad ? 0.5 * P.wmu[mu] : P.wmu[mu];
      const double I = C.I[((size_t) col * P.nray + r) * N + k];
      const double V = c * ph[(size_t) (2*mu + P.ray_dir[r]) * N];
      const double wlamu = V * w * wmu;
      Rij += I * wlamu;
      Rji += g * (thn + I) * wlamu;
    }
  }
  C.Rij[((size_t) col * P.Ntrans + tid) * N + k] = Rij;
  C.Rji[((size_t) col * P.Ntrans + tid) * N + k] = Rji;
}

__global__ void nlte_fill_kernel(double *__restrict__ a, size_t n, double v)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) a[t] = v;
}

// dJmax per column (solveSpectrum's return value, iterate.c:236-244): max is order independent
__global__ void __launch_bounds__(256)
nlte_dJmax_kernel(Plan P, Cols C, int ncol, double *dJmax)
{
  const int col = blockIdx.x;
  if (col >= ncol || !C.active[col]) return;
  __shared__ double sm[256];
  const size_t per = (size_t) P.Nspect * P.Ndep;
  double m = 0.0;
  for (size_t i = threadIdx.x; i < per; i += blockDim.x) {
    const double d = C.dJ[(size_t) col * per + i];
    if (d > m) m = d;                                 // NaN (J = 0 and Jdag = 0) never wins, like the reference's MAX
  }
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) { if (threadIdx.x < s2) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s2]); __syncthreads(); }
  if (threadIdx.x == 0) dJmax[col] = sm[0];
}

// ---- (4) addtoCoupling + addtoGamma + addtoRates for one transition at one depth.
// One transition per block, threads over (column, depth): the walk over this transition's wavelengths,
// rays and active-set entries is uniform across the block (no divergent trip counts, broadcast loads of
// the plan tables).  Everything that does not depend on the ray is gathered once per wavelength into a
// small per-thread cache: the products are formed in the reference's association order
// (V*w)*(n_i - g n_j), (thn*g)*V, ((thn*g)*V)*n_j, so hoisting the right-hand factors changes no rounding.
#define NLTE_MAXACT 12
// SEG: the thread covers one SEGMENT of its transition's wavelengths and stores partial sums (stage 1 of the
// fixed-partition reduction, nlte_gamma_sum_kernel adds them in segment order); !SEG: the whole transition in the
// reference's order, bit-identical to fillgamma.c.
template <bool SEG, bool STK>
__global__ void __launch_bounds__(64, NLTE_GAMMA_MINB)
nlte_gamma_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * N) return;
  const int k = (int) (t % N), col = (int) (t / N);
  const int seg = SEG ? (int) blockIdx.y : 0;
  const int tid = SEG ? P.seg_tr[seg] : (int) blockIdx.y;   // heavy transitions (lines) start first
  if (!C.active[col]) return;
  const double *trs = P.trans + (size_t) tid * TR_NFIELD;
  const int a = (int) trs[TR_ATOM], i = (int) trs[TR_I], j = (int) trs[TR_J], Nl = P.atom_nlevel[a];
  const int Nblue = (int) trs[TR_NBLUE], Nla = (int) trs[TR_NLAMBDA];
  const double *ncol_ = C.n + (size_t) col * P.nlev * N;
  const size_t gbase = ((size_t) col * P.ngam + P.gam_off[a]) * N;
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  double Gij = 0.0, Gji = 0.0;
  if (!SEG && P.add_C) { Gij = C.C[gbase + (size_t)(i*Nl + j) * N + k]; Gji = C.C[gbase + (size_t)(j*Nl + i) * N + k]; }   // initGammaAtom
  double Rij = 0.0, Rji = 0.0;                                                                          // zeroRates

  const int w_lo = SEG ? P.seg_lo[seg] : Nblue, w_hi = SEG ? P.seg_hi[seg] : Nblue + Nla;
  const int ns_first = w_lo > P.ns_lo ? w_lo : P.ns_lo, ns_last = w_hi < P.ns_hi ? w_hi : P.ns_hi;
  for (int ns = ns_first; ns < ns_last; ns++) {
    const int first = P.as_first[ns], nact = P.as_first[ns+1] - first;
    const int ad = P.angle_dep[ns];
    const bool pol = STK && P.pol_as[ns];        // FULL_STOKES && containsPolarized(as), fillgamma.c:106,122
    const size_t plane = (size_t) ncol * P.nphirow * N;
    double e_el[STK ? NLTE_MAXACT : 1];          // Bij hc/4pi (2h nu^3/c^2) g_ij n_j of polarizable lines (opacity.c:283-284)
    // ---- entries of this atom at this wavelength (ray independent part)
    int    e_flag[NLTE_MAXACT];                 // bit0 im==i, bit1 jm==j, bit2 jm==i, bit3 self, bit4 thn != 0, bit5 line
    double e_w[NLTE_MAXACT], e_diff[NLTE_MAXACT], e_tg[NLTE_MAXACT], e_nj[NLTE_MAXACT], e_c[NLTE_MAXACT];
    const double *e_phi[NLTE_MAXACT];
    double gs = 0.0, thns = 0.0;
    int m = 0, n_jp_eq_i = 0;
    for (int n = 0; n < nact; n++) {
      const int tm = P.as_trans[first+n];
      const double *tr = P.trans + (size_t) tm * TR_NFIELD;
      if ((int) tr[TR_ATOM] != a) continue;
      if (m == NLTE_MAXACT) { m++; break; }
      const int im = (int) tr[TR_I], jm = (int) tr[TR_J], la = ns - (int) tr[TR_NBLUE];
      const double *gw = C.gw + (((size_t) col * P.nas + first + n) * 2) * N + k;
      const double g = gw[0], w = gw[N];
      const double thn = twohnu3_of(P, tr, ns);
      const double n_i = ncol_[(size_t)(P.lev_off[a] + im) * N + k], n_j = ncol_[(size_t)(P.lev_off[a] + jm) * N + k];
      const bool line = tr[TR_TYPE] == 0.0;
      e_flag[m] = (im == i ? 1 : 0) | (jm == j ? 2 : 0) | (jm == i ? 4 : 0) | (tm == tid ? 8 : 0) |
                  (thn != 0.0 ? 16 : 0) | (line ? 32 : 0);
      if (STK) {
        const bool lp = pol && line && P.line_pol[(int) tr[TR_LINEIDX]] != 0;
        if (lp) { e_flag[m] |= 64; e_el[m] = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC] * thn * g * n_j; }
      }
      e_w[m] = w; e_diff[m] = n_i - g*n_j; e_tg[m] = thn * g; e_nj[m] = n_j;
      if (line) {                                // V = Bij hc/4pi phi(ray), opacity.c:188-193
        e_c[m] = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC];
        e_phi[m] = C.phi + ((size_t) col * P.nphirow + (int) tr[TR_PHIROW] + 2*P.Nrays*la) * N + k;
      } else {                                   // V = alpha(lambda), opacity.c:236
        e_c[m] = P.tr_alpha[(int) tr[TR_WOFF] + la];
        e_phi[m] = nullptr;
      }
      if (jm == i) n_jp_eq_i++;
      if (tm == tid) { gs = g; thns = thn; }
      m++;
    }
    if (m > NLTE_MAXACT) { Gij = Gji = Rij = Rji = __longlong_as_double(0x7ff8000000000000LL); break; }   // refused on the host
    // ---- rays of this wavelength in batches of NLTE_RB: all loads of a batch are issued before the
    //      (ordered) accumulation, which is what hides the memory latency of this otherwise serial walk
    const int r_end = P.ray_off[ns+1];
    for (int r0 = P.ray_off[ns]; r0 < r_end; r0 += NLTE_RB) {
      const int nb = r_end - r0 < NLTE_RB ? r_end - r0 : NLTE_RB;
      double Iv[NLTE_RB], Pv[NLTE_RB], wmuv[NLTE_RB];
      int lamu[NLTE_RB];
      double eta_atom[NLTE_RB], chi_up_i[NLTE_RB], Uji_down_j[NLTE_RB], chi_down_j[NLTE_RB], Uji_down_i[NLTE_RB], Vs[NLTE_RB];
      double eQ[STK ? NLTE_RB : 1], eU[STK ? NLTE_RB : 1], eV[STK ? NLTE_RB : 1], Iq[STK ? NLTE_RB : 1], Iu[STK ? NLTE_RB : 1], Iv4[STK ? NLTE_RB : 1];
      double ws = 0.0;
#pragma unroll
      for (int q = 0; q < NLTE_RB; q++) {
        eta_atom[q] = chi_up_i[q] = Uji_down_j[q] = chi_down_j[q] = Uji_down_i[q] = Vs[q] = 0.0;
        Iv[q] = Pv[q] = wmuv[q] = 0.0; lamu[q] = 0;
        if (q < nb) {
          const int r = r0 + q, mu = P.ray_mu[r];
          const size_t rk = ((size_t) col * P.nray + r) * N + k;
          Iv[q] = __ldg(C.I + rk);
          Pv[q] = __ldg(C.Psi + rk);                               // Psi / chi (formal.c:248 / :301), divided by the ray kernel
          wmuv[q] = ad ? 0.5 * P.wmu[mu] : P.wmu[mu];
          lamu[q] = 2*mu + P.ray_dir[r];
        }
        if (STK) {
          eQ[q] = eU[q] = eV[q] = Iq[q] = Iu[q] = Iv4[q] = 0.0;
          if (pol && q < nb) {
            const double *iq = C.IQ + (((size_t) col * P.nray + r0 + q) * 3) * N + k;
            Iq[q] = __ldg(iq); Iu[q] = __ldg(iq + N); Iv4[q] = __ldg(iq + 2*(size_t) N);
          }
        }
      }
      for (int e = 0; e < m; e++) {               // entries in active-set order: per-ray sums keep their order
        const int f = e_flag[e];
        const double c = e_c[e], tg = e_tg[e], w = e_w[e], diff = e_diff[e], nj = e_nj[e];
        const double *ph = e_phi[e];
        double V[NLTE_RB];
#pragma unroll
        for (int q = 0; q < NLTE_RB; q++)
          V[q] = (q < nb) ? ((f & 32) ? c * __ldg(ph + (size_t) lamu[q] * N) : c) : 0.0;
        if (f & 8) ws = w;
#pragma unroll
        for (int q = 0; q < NLTE_RB; q++) {
          if (f & 16) {
            const double tgV = tg * V[q];
            eta_atom[q] += tgV * nj;                              // opacity.c:257-258
            const double chicc = V[q] * w * diff;                 // fillgamma.c:321-327
            if (f & 1) chi_up_i[q] += chicc;
            if (f & 2) { chi_down_j[q] += chicc; Uji_down_j[q] += tgV; }
            if (f & 4) Uji_down_i[q] += tgV;
          }
          if (f & 8) Vs[q] = V[q];
        }
        if (STK && (f & 64)) {                      // eta_Q,U,V of this atom (opacity.c:283-291)
          const double el = e_el[e];
#pragma unroll
          for (int q = 0; q < NLTE_RB; q++)
            if (q < nb) {
              const double *pq = C.phiQ + (size_t) (ph - C.phi) + (size_t) lamu[q] * N;
              eQ[q] += el * __ldg(pq); eU[q] += el * __ldg(pq + plane); eV[q] += el * __ldg(pq + 2*plane);
            }
        }
      }
#pragma unroll
      for (int q = 0; q < NLTE_RB; q++) {
        if (q < nb) {
          const double I = Iv[q], Psi = Pv[q], wmu = wmuv[q];
          const double Ieff = (STK && pol) ? I + Iq[q] + Iu[q] + Iv4[q] - Psi * (eta_atom[q] + eQ[q] + eU[q] + eV[q])   // fillgamma.c:122-128
                                           : I - Psi * eta_atom[q];               // fillgamma.c:130-133
          const double wlamu = Vs[q] * ws * wmu;
          Gji += Ieff * wlamu;                                     // fillgamma.c:163-168
          Gij += (thns + Ieff) * gs * wlamu;
          Gij -= chi_up_i[q] * Psi * Uji_down_j[q] * wmu;          // fillgamma.c:172-175
          for (int z = 0; z < n_jp_eq_i; z++)                      // fillgamma.c:180-196
            Gji += chi_down_j[q] * Psi * Uji_down_i[q] * wmu;
          Rij += I * wlamu;                                        // fillgamma.c:448-452
          Rji += gs * (thns + I) * wlamu;
        }
      }
    }
  }
  if (SEG) {
    double *o = C.part + (((size_t) col * P.nseg + seg) * 4) * N + k;
    o[0] = Gij; o[N] = Gji; o[2*(size_t) N] = Rij; o[3*(size_t) N] = Rji;
    return;
  }
  C.Gamma[gbase + (size_t)(i*Nl + j) * N + k] = Gij;
  C.Gamma[gbase + (size_t)(j*Nl + i) * N + k] = Gji;
  C.Rij[((size_t) col * P.Ntrans + tid) * N + k] = Rij;
  C.Rji[((size_t) col * P.Ntrans + tid) * N + k] = Rji;
}

// ---- (4b) the default rate accumulation, ATOM-major: one thread per (column, depth) x (atom, chunk of wavelengths).
// The transition-major kernels above load I, Psi, chi once per (transition, wavelength, ray) -- for H + Ca II nine
// times per ray-point, because the continua overlap every line; here they are loaded once per (atom, wavelength, ray)
// and addtoCoupling's per-level sums (chi_up[i], chi_down[j], Uji_down[j], fillgamma.c:251-332) are formed once and
// shared by all transitions of the atom, as in the reference.  Partial {Gij, Gji, Rij, Rji} per (chunk, transition
// slot) are added by nlte_gamma_slot_sum_kernel in chunk order: a fixed partition, deterministic and independent of
// batch size and chunking.  Within a chunk every sum keeps the reference's order.
#define NLTE_MAXSLOT 24
#define NLTE_MAXLEV 8
__global__ void __launch_bounds__(64, NLTE_GAMMA_MINB)
nlte_gamma_atom_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * N) return;
  const int k = (int) (t % N), col = (int) (t / N);
  if (!C.active[col]) return;
  const int sg = (int) blockIdx.y, a = P.aseg_atom[sg], slot0 = P.aseg_slot0[sg], nslot = P.aseg_slot0[sg+1] - slot0;
  const double *ncol_ = C.n + (size_t) col * P.nlev * N;
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  // per-thread accumulators and level sums live in shared memory, [index][thread]: dynamically indexed, conflict free
  extern __shared__ double sh_gamma[];
  double *acc = sh_gamma + threadIdx.x;                       // acc[(slot*4 + c)*64]
  double *lev = sh_gamma + (size_t) P.aseg_maxslot * 4 * 64 + threadIdx.x;     // lev[(x*NLTE_MAXLEV + l)*64], x = 0 chi_up, 1 chi_down, 2 Uji_down
#define ACC(q, c) acc[((q)*4 + (c))*64]
#define LEV(x, l) lev[((x)*NLTE_MAXLEV + (l))*64]
  for (int q = 0; q < nslot; q++) ACC(q, 0) = ACC(q, 1) = ACC(q, 2) = ACC(q, 3) = 0.0;
  const int ns_first = P.aseg_lo[sg] > P.ns_lo ? P.aseg_lo[sg] : P.ns_lo, ns_last = P.aseg_hi[sg] < P.ns_hi ? P.aseg_hi[sg] : P.ns_hi;
  for (int ns = ns_first; ns < ns_last; ns++) {
    const int first = P.as_first[ns], nact = P.as_first[ns+1] - first;
    const int ad = P.angle_dep[ns];
    int    e_im[NLTE_MAXACT], e_jm[NLTE_MAXACT], e_slot[NLTE_MAXACT], e_on[NLTE_MAXACT];
    double e_w[NLTE_MAXACT], e_diff[NLTE_MAXACT], e_tg[NLTE_MAXACT], e_nj[NLTE_MAXACT], e_c[NLTE_MAXACT], e_g[NLTE_MAXACT], e_thn[NLTE_MAXACT];
    const double *e_phi[NLTE_MAXACT];
    int cntj[NLTE_MAXLEV];                          // number of entries whose upper level is this level (fillgamma.c:180-196)
    for (int l = 0; l < NLTE_MAXLEV; l++) cntj[l] = 0;
    int m = 0;
    for (int n = 0; n < nact && m < NLTE_MAXACT; n++) {
      const int tm = P.as_trans[first+n];
      const double *tr = P.trans + (size_t) tm * TR_NFIELD;
      if ((int) tr[TR_ATOM] != a) continue;
      const int im = (int) tr[TR_I], jm = (int) tr[TR_J], la = ns - (int) tr[TR_NBLUE];
      const double *gw = C.gw + (((size_t) col * P.nas + first + n) * 2) * N + k;
      const double g = gw[0], w = gw[N];
      const double thn = twohnu3_of(P, tr, ns);
      const double n_i = ncol_[(size_t)(P.lev_off[a] + im) * N + k], n_j = ncol_[(size_t)(P.lev_off[a] + jm) * N + k];
      e_im[m] = im; e_jm[m] = jm; e_slot[m] = P.as_slot[first+n]; e_on[m] = thn != 0.0;
      e_w[m] = w; e_diff[m] = n_i - g*n_j; e_tg[m] = thn * g; e_nj[m] = n_j; e_g[m] = g; e_thn[m] = thn;
      if (tr[TR_TYPE] == 0.0) {
        e_c[m] = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC];
        e_phi[m] = C.phi + ((size_t) col * P.nphirow + (int) tr[TR_PHIROW] + 2*P.Nrays*la) * N + k;
      } else {
        e_c[m] = P.tr_alpha[(int) tr[TR_WOFF] + la];
        e_phi[m] = nullptr;
      }
      cntj[jm]++;
      m++;
    }
    const int r_end = P.ray_off[ns+1];
    for (int r0 = P.ray_off[ns]; r0 < r_end; r0 += NLTE_RB) {
      const int nb = r_end - r0 < NLTE_RB ? r_end - r0 : NLTE_RB;
      double Iv[NLTE_RB], Pv[NLTE_RB], wmuv[NLTE_RB];
      int lamu[NLTE_RB];
#pragma unroll
      for (int q = 0; q < NLTE_RB; q++) {           // the loads of a batch of rays are issued together
        Iv[q] = Pv[q] = wmuv[q] = 0.0; lamu[q] = 0;
        if (q < nb) {
          const int r = r0 + q, mu = P.ray_mu[r];
          const size_t rk = ((size_t) col * P.nray + r) * N + k;
          Iv[q] = __ldg(C.I + rk);
          Pv[q] = __ldg(C.Psi + rk);
          wmuv[q] = ad ? 0.5 * P.wmu[mu] : P.wmu[mu];
          lamu[q] = 2*mu + P.ray_dir[r];
        }
      }
      for (int q = 0; q < nb; q++) {
        double V[NLTE_MAXACT];
        for (int l = 0; l < NLTE_MAXLEV; l++) LEV(0, l) = LEV(1, l) = LEV(2, l) = 0.0;
        double eta_atom = 0.0;
        for (int e = 0; e < m; e++) {               // Opacity() + addtoCoupling(), active-set order
          const double v = e_phi[e] ? e_c[e] * __ldg(e_phi[e] + (size_t) lamu[q] * N) : e_c[e];
          V[e] = v;
          if (e_on[e]) {
            const double tgV = e_tg[e] * v;
            eta_atom += tgV * e_nj[e];
            const double chicc = v * e_w[e] * e_diff[e];
            LEV(0, e_im[e]) += chicc;
            LEV(1, e_jm[e]) += chicc;
            LEV(2, e_jm[e]) += tgV;
          }
        }
        const double I = Iv[q], Psi = Pv[q], wmu = wmuv[q];
        const double Ieff = I - Psi * eta_atom;
        for (int e = 0; e < m; e++) {               // addtoGamma() + addtoRates() of every transition of the atom
          const int i = e_im[e], j = e_jm[e];
          const double wlamu = V[e] * e_w[e] * wmu;
          const int sl = e_slot[e];
          double Gij = ACC(sl, 0), Gji = ACC(sl, 1);
          Gji += Ieff * wlamu;
          Gij += (e_thn[e] + Ieff) * e_g[e] * wlamu;
          Gij -= LEV(0, i) * Psi * LEV(2, j) * wmu;
          for (int z = 0; z < cntj[i]; z++) Gji += LEV(1, j) * Psi * LEV(2, i) * wmu;
          ACC(sl, 0) = Gij; ACC(sl, 1) = Gji;
          ACC(sl, 2) += I * wlamu;
          ACC(sl, 3) += e_g[e] * (e_thn[e] + I) * wlamu;
        }
      }
    }
  }
  for (int q = 0; q < nslot; q++) {
    double *o = C.part + (((size_t) col * P.nslot + slot0 + q) * 4) * N + k;
    o[0] = ACC(q, 0); o[N] = ACC(q, 1); o[2*(size_t) N] = ACC(q, 2); o[3*(size_t) N] = ACC(q, 3);
  }
#undef ACC
#undef LEV
}

__global__ void __launch_bounds__(128)
nlte_gamma_slot_sum_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.Ntrans * N) return;
  const int k = (int) (t % N), tid = (int) ((t / N) % P.Ntrans), col = (int) (t / ((size_t) N * P.Ntrans));
  if (!C.active[col]) return;
  const double *trs = P.trans + (size_t) tid * TR_NFIELD;
  const int a = (int) trs[TR_ATOM], i = (int) trs[TR_I], j = (int) trs[TR_J], Nl = P.atom_nlevel[a];
  const size_t gbase = ((size_t) col * P.ngam + P.gam_off[a]) * N;
  double Gij = 0.0, Gji = 0.0, Rij = 0.0, Rji = 0.0;
  if (P.add_C) { Gij = C.C[gbase + (size_t)(i*Nl + j) * N + k]; Gji = C.C[gbase + (size_t)(j*Nl + i) * N + k]; }
  for (int q = P.tr_slot0[tid]; q < P.tr_slot0[tid+1]; q++) {
    const double *p = C.part + (((size_t) col * P.nslot + P.tr_slots[q]) * 4) * N + k;
    Gij += p[0]; Gji += p[N]; Rij += p[2*(size_t) N]; Rji += p[3*(size_t) N];
  }
  C.Gamma[gbase + (size_t)(i*Nl + j) * N + k] = Gij;
  C.Gamma[gbase + (size_t)(j*Nl + i) * N + k] = Gji;
  C.Rij[t] = Rij;
  C.Rji[t] = Rji;
}

// stage 2: collisional part + the segments' partial sums in segment (= wavelength) order: the same partition for every
// batch size, chunking and rank count, so results are run-to-run and layout independent (not the reference's add order:
// populations agree to ~1e-13 instead of to the bit; north_star asks 1e-6)
__global__ void __launch_bounds__(128)
nlte_gamma_sum_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.Ntrans * N) return;
  const int k = (int) (t % N), tid = (int) ((t / N) % P.Ntrans), col = (int) (t / ((size_t) N * P.Ntrans));
  if (!C.active[col]) return;
  const double *trs = P.trans + (size_t) tid * TR_NFIELD;
  const int a = (int) trs[TR_ATOM], i = (int) trs[TR_I], j = (int) trs[TR_J], Nl = P.atom_nlevel[a];
  const size_t gbase = ((size_t) col * P.ngam + P.gam_off[a]) * N;
  double Gij = 0.0, Gji = 0.0, Rij = 0.0, Rji = 0.0;
  if (P.add_C) { Gij = C.C[gbase + (size_t)(i*Nl + j) * N + k]; Gji = C.C[gbase + (size_t)(j*Nl + i) * N + k]; }
  for (int sg = P.tr_seg0[tid]; sg < P.tr_seg0[tid+1]; sg++) {
    const double *p = C.part + (((size_t) col * P.nseg + sg) * 4) * N + k;
    Gij += p[0]; Gji += p[N]; Rij += p[2*(size_t) N]; Rji += p[3*(size_t) N];
  }
  C.Gamma[gbase + (size_t)(i*Nl + j) * N + k] = Gij;
  C.Gamma[gbase + (size_t)(j*Nl + i) * N + k] = Gji;
  C.Rij[t] = Rij;
  C.Rji[t] = Rji;
}

// Gamma entries that belong to no radiative transition keep the collisional value (initGammaAtom)
__global__ void nlte_gamma_init_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t) P.ngam * P.Ndep;
  if (t >= (size_t) ncol * per) return;
  if (!C.active[t / per]) return;
  C.Gamma[t] = P.add_C ? C.C[t] : 0.0;
}

// wavelength shard: entries of J owned by other ranks are zeroed before the closing sum-allreduce
__global__ void nlte_zero_foreign_J_kernel(Plan P, Cols C, int ncol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.Nspect * N) return;
  const int ns = (int) ((t / N) % P.Nspect);
  if (ns < P.ns_lo || ns >= P.ns_hi) C.J[t] = 0.0;
}

// ---- (5) SolveLinearEq: rhb200_lu.cuh
using rhlu::solve_linear_eq;

// statEquil, statequil.c:40-103: one thread per (column, atom, depth)
template <int MAXN>
__global__ void __launch_bounds__(64)
nlte_statequil_kernel(Plan P, Cols C, int ncol, int isum)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.Natom * N) return;
  const int k = (int) (t % N);
  const size_t ca = t / N;
  const int a = (int) (ca % P.Natom), col = (int) (ca / P.Natom);
  if (!C.active[col]) return;
  const int Nl = P.atom_nlevel[a];
  double G[MAXN*MAXN], nk[MAXN];
  double *n = C.n + ((size_t) col * P.nlev + P.lev_off[a]) * N;
  const double *Gam = C.Gamma + ((size_t) col * P.ngam + P.gam_off[a]) * N;
  for (int i = 0; i < Nl; i++) {
    nk[i] = n[(size_t) i*N + k];
    for (int j = 0; j < Nl; j++) G[i*Nl+j] = Gam[(size_t)(i*Nl+j)*N + k];
  }
  int ie = isum;
  if (isum == -1) {
    ie = 0; double nmax = 0.0;
    for (int i = 0; i < Nl; i++) if (nk[i] > nmax) { nmax = nk[i]; ie = i; }
  }
  for (int i = 0; i < Nl; i++) {
    double GamDiag = 0.0;
    G[i*Nl+i] = 0.0; nk[i] = 0.0;
    for (int j = 0; j < Nl; j++) GamDiag += G[j*Nl+i];
    G[i*Nl+i] = -GamDiag;
  }
  nk[ie] = C.ntotal[((size_t) col * P.Natom + a) * N + k];
  for (int j = 0; j < Nl; j++) G[ie*Nl+j] = 1.0;
  solve_linear_eq<MAXN>(Nl, G, nk, true);
  for (int i = 0; i < Nl; i++) n[(size_t) i*N + k] = nk[i];
}

// batched SolveLinearEq entry (test hook): systems [nsys][N*N] + [nsys][N]
template <int MAXN>
__global__ void solve_batch_kernel(int nsys, int N, double *A, double *b, int improve)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsys) return;
  double G[MAXN*MAXN], x[MAXN];
  for (int i = 0; i < N*N; i++) G[i] = A[(size_t) s*N*N + i];
  for (int i = 0; i < N; i++) x[i] = b[(size_t) s*N + i];
  solve_linear_eq<MAXN>(N, G, x, improve != 0);
  for (int i = 0; i < N; i++) b[(size_t) s*N + i] = x[i];
}

// ---- (6) Accelerate + MaxChange (accelerate.c:68-147, maxchange.c:32-50): one block per (column, atom)
__global__ void __launch_bounds__(128)
nlte_ng_kernel(Plan P, Cols C, int ncol, double *previous /*[col][atom-offset][(Norder+2)][Nl*N]*/,
               const size_t *prev_off, int Norder, int Ndelay, int Nperiod, int count, double *dpops)
{
  const int ca = blockIdx.x, a = ca % P.Natom, col = ca / P.Natom;
  if (col >= ncol || !C.active[col]) return;
  const int Nn = P.atom_nlevel[a] * P.Ndep, tid = threadIdx.x, nt = blockDim.x;
  double *sol = C.n + ((size_t) col * P.nlev + P.lev_off[a]) * P.Ndep;
  double *prev = previous + ((size_t) col * prev_off[P.Natom] + prev_off[a]);
  __shared__ double sA[16], sb[4], smax[128];
  // store the current solution (count is the value BEFORE the increment)
  const int slot = count % (Norder + 2);
  for (int k = tid; k < Nn; k += nt) prev[(size_t) slot*Nn + k] = sol[k];
  __syncthreads();
  const int cnt = count + 1;
  if ((Norder > 0) && (cnt >= Ndelay) && !((cnt - Ndelay) % Nperiod)) {
    auto delta = [&](int i, int k) {
      const int ip = (cnt - 1 - i) % (Norder + 2), ipp = (cnt - 2 - i) % (Norder + 2);
      return prev[(size_t) ip*Nn + k] - prev[(size_t) ipp*Nn + k];
    };
    // entries: b[j] (j < Norder) then A[i][j]; each summed sequentially over k by one thread
    if (tid < Norder + Norder*Norder) {
      double s = 0.0;
      if (tid < Norder) {
        const int j = tid;
        for (int k = 0; k < Nn; k++) {
          const double w = 1.0 / fabs(sol[k]), d0 = delta(0, k);
          s += w * d0*(d0 - delta(j+1, k));
        }
        sb[j] = s;
      } else {
        const int e = tid - Norder, i = e / Norder, j = e % Norder;
        for (int k = 0; k < Nn; k++) {
          const double w = 1.0 / fabs(sol[k]), d0 = delta(0, k);
          s += w * (delta(j+1, k) - d0) * (delta(i+1, k) - d0);
        }
        sA[i*Norder + j] = s;
      }
    }
    __syncthreads();
    if (tid == 0) {
      double A[16], b[4];
      for (int i = 0; i < Norder*Norder; i++) A[i] = sA[i];
      for (int i = 0; i < Norder; i++) b[i] = sb[i];
      solve_linear_eq<4>(Norder, A, b, true);
      for (int i = 0; i < Norder; i++) sb[i] = b[i];
    }
    __syncthreads();
    const int i0 = (cnt - 1) % (Norder + 2);
    for (int k = tid; k < Nn; k += nt) {
      double s = sol[k];
      for (int i = 0; i < Norder; i++) {
        const int ip = (cnt - 2 - i) % (Norder + 2);
        s += sb[i] * (prev[(size_t) ip*Nn + k] - prev[(size_t) i0*Nn + k]);
      }
      sol[k] = s;
    }
    __syncthreads();
    for (int k = tid; k < Nn; k += nt) prev[(size_t) i0*Nn + k] = sol[k];
    __syncthreads();
  }
  // MaxChange
  double dmax = 0.0;
  if (cnt >= 2) {
    const double *old = prev + (size_t)((cnt - 2) % (Norder + 2))*Nn, *nw = prev + (size_t)((cnt - 1) % (Norder + 2))*Nn;
    for (int k = tid; k < Nn; k += nt)
      if (nw[k] != 0.0) dmax = fmax(dmax, fabs((nw[k] - old[k]) / nw[k]));
  }
  smax[tid] = dmax;
  __syncthreads();
  for (int s = nt/2; s > 0; s >>= 1) { if (tid < s) smax[tid] = fmax(smax[tid], smax[tid+s]); __syncthreads(); }
  if (tid == 0) dpops[(size_t) col * P.Natom + a] = smax[0];
}

// spectrum.I[nspect][0] of every wavelength -> spec[col][Nspect]: the up-ray of ray mu = 0 (formal.c:270), Feautrier's
// emergent intensity where the wavelength is angle independent (formal.c:299-305)
__global__ void __launch_bounds__(128)
nlte_pack_spectrum_kernel(Plan P, Cols C, int ncol, double *__restrict__ spec)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * P.Nspect) return;
  const int ns = (int) (t % P.Nspect), col = (int) (t / P.Nspect);
  const int r = P.ray_off[ns] + (P.angle_dep[ns] ? 1 : 0);      // rays of a wavelength: (mu 0, down), (mu 0, up), ...
  spec[t] = C.Iem[(size_t) col * P.nray + r];
}
// spectrum.Stokes_Q/U/V[nspect][0] (formal.c:273-277): zero where Formal() did not solve for them (initSolution's calloc)
__global__ void __launch_bounds__(128)
nlte_pack_quv_kernel(Plan P, Cols C, int ncol, double *__restrict__ quv /* [ncol][3][Nspect] */)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * P.Nspect) return;
  const int ns = (int) (t % P.Nspect), col = (int) (t / P.Nspect);
  const bool solved = P.stokes && (P.pol_as[ns] || P.pol_c[ns]);
  const int r = P.ray_off[ns] + (P.angle_dep[ns] ? 1 : 0);
  for (int s = 0; s < 3; s++)
    quv[((size_t) col * 3 + s) * P.Nspect + ns] = solved ? C.IemQ[((size_t) col * P.nray + r) * 3 + s] : 0.0;
}

template <class T> int up(T **d, const T *h, size_t n)
{
  *d = nullptr;
  RH_CUDA(cudaMalloc((void **) d, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) RH_CUDA(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return RHB200_OK;
}

struct DevArena {           // everything allocated for one call, freed on scope exit
  std::vector<void *> ptrs;
  ~DevArena() { for (void *p : ptrs) cudaFree(p); }
  template <class T> int alloc(T **d, size_t n, bool zero = false) {
    *d = nullptr;
    cudaError_t e = cudaMalloc((void **) d, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) { cudaGetLastError(); rhb200_set_error("cudaMalloc(%zu): %s", n * sizeof(T), cudaGetErrorString(e)); return RHB200_ENOMEM; }
    ptrs.push_back(*d);
    if (zero) RH_CUDA(cudaMemset(*d, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return RHB200_OK;
  }
  template <class T> int upload(T **d, const T *h, size_t n) {
    int rc = alloc(d, n);
    if (rc != RHB200_OK) return rc;
    if (n) RH_CUDA(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return RHB200_OK;
  }
};

#define RH_CHECK(expr) do { int rc__ = (expr); if (rc__ != RHB200_OK) return rc__; } while (0)
#define RH_GRID(n, b) (unsigned) (((size_t) (n) + (b) - 1) / (b)), (b)

}  // namespace

extern "C" int rhb200_solve_linear_eq_batch(rhb200_ctx *c, int nsys, int N, double *A, double *b, int improve)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  if (nsys <= 0 || N <= 0 || N > 32 || !A || !b) { rhb200_set_error("bad arguments (N <= 32)"); return RHB200_EINVAL; }
  DevArena ar;
  double *dA, *db;
  RH_CHECK(ar.upload(&dA, A, (size_t) nsys*N*N));
  RH_CHECK(ar.upload(&db, b, (size_t) nsys*N));
  if (N <= 8) solve_batch_kernel<8><<<RH_GRID(nsys, 64), 0, c->stream>>>(nsys, N, dA, db, improve);
  else if (N <= 16) solve_batch_kernel<16><<<RH_GRID(nsys, 64), 0, c->stream>>>(nsys, N, dA, db, improve);
  else solve_batch_kernel<32><<<RH_GRID(nsys, 32), 0, c->stream>>>(nsys, N, dA, db, improve);
  RH_CUDA(cudaGetLastError());
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CUDA(cudaMemcpy(b, db, (size_t) nsys*N*sizeof(double), cudaMemcpyDeviceToHost));
  return RHB200_OK;
}

void rh_nlte_shard_range(int Nspect, const int *ray_off, int rank, int nrank, int *lo, int *hi);

// ---- the solver as an engine: plan-level tables built once (build), per-batch work arrays (alloc), inputs either
//      uploaded from the host (the function-level entry points) or bound as device arrays the front end filled
//      (rhb200_nlte_compute1d_batch); then prepare -> scatter / iterate in any order, all on ctx->stream.
struct NlteEngine {
  rhb200_ctx *c = nullptr;
  DevArena plan_ar, work_ar;
  Plan P{};
  Cols C{};
  int Ns = 0, N = 0, Na = 0, Nt = 0, Nr = 0, nlev = 0, ngam = 0, nas = 0, nray = 0, maxnl = 0, nphirow = 0, nline = 0;
  int isum = -1, Norder = 0, Ndelay = 0, Nperiod = 1, ncol = 0, nrank = 1, rank = 0;
  bool device_profiles = false;
  std::vector<int> lev_off, gam_off, angle_dep, ray_off, ray_ns, ray_mu, ray_dir, nlevel;
  std::vector<size_t> prev_off;
  int *d_active = nullptr; double *d_prev = nullptr; size_t *d_prev_off = nullptr; double *d_dpops = nullptr, *d_dJmax = nullptr;
  std::vector<int> active;

  int allreduce(double *buf, size_t count, int op) {
    if (nrank == 1) return RHB200_OK;
    if (c->nccl_comm) return rh_nccl_allreduce(c, buf, count, op);        // enqueued on the compute stream, no host sync
    RH_CUDA(cudaStreamSynchronize(c->stream));
    if (c->shard_fn(c->shard_user, buf, count, op) != 0) { rhb200_set_error("allreduce callback failed"); return RHB200_ECUDA; }
    return RHB200_OK;
  }

  int build(rhb200_ctx *ctx, const rhb200_nlte_plan *pl) {
    c = ctx;
    Ns = pl->Nspect; N = pl->Ndep; Na = pl->Natom; Nt = pl->Ntrans; Nr = pl->Nrays;
    if (Ns <= 0 || N < 3 || Na <= 0 || Na > 15 || Nt <= 0 || Nr <= 0) { rhb200_set_error("rhb200_nlte: bad plan sizes"); return RHB200_EINVAL; }
    if (!pl->moving) { rhb200_set_error("static atmospheres (angle-independent line profiles) are not implemented"); return RHB200_EUNSUPPORTED; }
    if (pl->Ngorder > 4 || pl->Ngorder < 0) { rhb200_set_error("NG_ORDER > 4 is not implemented"); return RHB200_EUNSUPPORTED; }
    lev_off.assign(Na+1, 0); gam_off.assign(Na+1, 0); nlevel.assign(pl->atom_nlevel, pl->atom_nlevel + Na);
    for (int a = 0; a < Na; a++) {
      maxnl = std::max(maxnl, pl->atom_nlevel[a]);
      lev_off[a+1] = lev_off[a] + pl->atom_nlevel[a];
      gam_off[a+1] = gam_off[a] + pl->atom_nlevel[a]*pl->atom_nlevel[a];
    }
    if (maxnl > 32) { rhb200_set_error("atoms with more than 32 levels are not implemented"); return RHB200_EUNSUPPORTED; }
    nlev = lev_off[Na]; ngam = gam_off[Na]; nas = pl->as_first[Ns]; nphirow = pl->nphirow; nline = pl->nline;
    isum = pl->isum;
    // derived host tables: angle dependence (formal.c:100-103), ray list in the reference's order
    angle_dep.assign(Ns, 0); ray_off.assign(Ns+1, 0);
    std::vector<int> as_pack(2*(size_t) nas);
    for (int ns = 0; ns < Ns; ns++) {
      bool bb = false;
      for (int e = pl->as_first[ns]; e < pl->as_first[ns+1]; e++) {
        as_pack[e] = pl->as_trans[e]; as_pack[nas + e] = ns;
        const double ty = pl->trans[(size_t) pl->as_trans[e]*RHB200_TR_NFIELD + RHB200_TR_TYPE];
        if (ty == 0.0) bb = true;
        if (ty != 0.0 && ty != 1.0) { rhb200_set_error("transition type must be 0 (line) or 1 (continuum)"); return RHB200_EINVAL; }
      }
      {                                          // the rate kernel caches one atom's entries of a wavelength
        std::vector<int> per_atom(Na, 0);
        for (int e = pl->as_first[ns]; e < pl->as_first[ns+1]; e++) {
          const int a = (int) pl->trans[(size_t) pl->as_trans[e]*RHB200_TR_NFIELD + RHB200_TR_ATOM];
          if (a < 0 || a >= Na) { rhb200_set_error("transition atom index out of range"); return RHB200_EINVAL; }
          if (++per_atom[a] > NLTE_MAXACT) {
            rhb200_set_error("more than %d active transitions of one atom at one wavelength", NLTE_MAXACT);
            return RHB200_EUNSUPPORTED;
          }
        }
      }
      angle_dep[ns] = pl->moving && (bb || pl->bg_hasline[ns]);
      for (int mu = 0; mu < Nr; mu++)
        for (int dir = 0; dir <= (angle_dep[ns] ? 1 : 0); dir++) { ray_ns.push_back(ns); ray_mu.push_back(mu); ray_dir.push_back(dir); }
      ray_off[ns+1] = (int) ray_ns.size();
    }
    nray = (int) ray_ns.size();
    P.Nspect = Ns; P.Nrays = Nr; P.Ndep = N; P.Natom = Na; P.Ntrans = Nt; P.nas = nas; P.nray = nray;
    P.nlev = nlev; P.ngam = ngam; P.nphirow = nphirow; P.nline = nline; P.bc_top = pl->bc_top; P.bc_bottom = pl->bc_bottom;
    P.solver = c->s_interpolation;
    // wavelength shard of one atmosphere (rhb200_nlte_set_shard): contiguous chunk balanced by ray count
    nrank = c->shard_nrank > 1 ? c->shard_nrank : 1; rank = nrank > 1 ? c->shard_rank : 0;
    rh_nlte_shard_range(Ns, ray_off.data(), rank, nrank, &P.ns_lo, &P.ns_hi);
    P.add_C = (rank == 0);
    DevArena &ar = plan_ar;
    double *dd; int *di;
#define UPI2(field, src, n) RH_CHECK(ar.upload(&di, src, (size_t) (n))); P.field = di
#define UPD(field, src, n) RH_CHECK(ar.upload(&dd, src, (size_t) (n))); P.field = dd
#define UPI(field, src, n) RH_CHECK(ar.upload(&di, src, (size_t) (n))); P.field = di
    UPD(lambda, pl->lambda, Ns); UPD(muz, pl->muz, Nr); UPD(wmu, pl->wmu, Nr);
    UPD(trans, pl->trans, (size_t) Nt*RHB200_TR_NFIELD);
    UPD(tr_lambda, pl->tr_lambda, pl->ntrl); UPD(tr_wlambda, pl->tr_wlambda, pl->ntrl); UPD(tr_alpha, pl->tr_alpha, pl->ntrl);
    UPI(atom_nlevel, pl->atom_nlevel, Na); UPI(lev_off, lev_off.data(), Na+1); UPI(gam_off, gam_off.data(), Na+1);
    UPI(as_first, pl->as_first, Ns+1); UPI(as_trans, as_pack.data(), 2*(size_t) nas);
    UPI(angle_dep, angle_dep.data(), Ns); UPI(ray_off, ray_off.data(), Ns+1);
    UPI(ray_ns, ray_ns.data(), nray); UPI(ray_mu, ray_mu.data(), nray); UPI(ray_dir, ray_dir.data(), nray);
    {                                            // profile-row and line-index maps (device-side Profile())
      std::vector<int> prow_tr(std::max(1, nphirow), -1), line_tr(std::max(1, nline), -1);
      bool ok = true;
      for (int t = 0; t < Nt; t++) {
        const double *tr = pl->trans + (size_t) t*RHB200_TR_NFIELD;
        if (tr[RHB200_TR_TYPE] != 0.0) continue;
        const int row0 = (int) tr[RHB200_TR_PHIROW], nrow = 2*Nr*(int) tr[RHB200_TR_NLAMBDA], li = (int) tr[RHB200_TR_LINEIDX];
        if (row0 < 0 || row0 + nrow > nphirow || li < 0 || li >= nline) { rhb200_set_error("transition %d: bad profile rows / line index", t); return RHB200_EINVAL; }
        if (!(tr[RHB200_TR_LAMBDA0] > 0.0)) ok = false;
        for (int r = 0; r < nrow; r++) prow_tr[row0 + r] = t;
        line_tr[li] = t;
      }
      for (int r = 0; r < nphirow; r++) if (prow_tr[r] < 0) ok = false;
      for (int l = 0; l < nline; l++) if (line_tr[l] < 0) ok = false;
      profile_maps_ok = ok;
      UPI(prow_tr, prow_tr.data(), prow_tr.size()); UPI(line_tr, line_tr.data(), line_tr.size());
    }
#undef UPD
#undef UPI
    {                                            // segments of the rate accumulation (RHB200_NLTE_GAMMA_SEG wavelengths each)
      int seglen = 16;
      if (const char *e = getenv("RHB200_NLTE_GAMMA_SEG")) { const int v = atoi(e); if (v > 0) seglen = v; }
      exact_rates = c->nlte_exact_rates != 0;
      if (const char *e = getenv("RHB200_NLTE_EXACT")) exact_rates = atoi(e) != 0;
      std::vector<int> seg_tr, seg_lo, seg_hi, tr_seg0(Nt + 1, 0);
      for (int t = 0; t < Nt; t++) {
        const double *tr = pl->trans + (size_t) t*RHB200_TR_NFIELD;
        const int b0 = (int) tr[RHB200_TR_NBLUE], b1 = b0 + (int) tr[RHB200_TR_NLAMBDA];
        tr_seg0[t] = (int) seg_tr.size();
        for (int lo = b0; lo < b1; lo += seglen) { seg_tr.push_back(t); seg_lo.push_back(lo); seg_hi.push_back(std::min(b1, lo + seglen)); }
      }
      tr_seg0[Nt] = (int) seg_tr.size();
      P.nseg = nseg = (int) seg_tr.size();
      UPI2(seg_tr, seg_tr.data(), std::max(1, nseg)); UPI2(seg_lo, seg_lo.data(), std::max(1, nseg));
      UPI2(seg_hi, seg_hi.data(), std::max(1, nseg)); UPI2(tr_seg0, tr_seg0.data(), Nt + 1);
      // atom-major chunks (nlte_gamma_atom_kernel): per atom, global wavelength chunks of `seglen`
      std::vector<int> aseg_atom, aseg_lo, aseg_hi, aseg_slot0(1, 0), as_slot(std::max(1, nas), 0), slot_tr;
      bool ok = true;
      int maxslot = 1;
      for (int a = 0; a < Na; a++) if (pl->atom_nlevel[a] > NLTE_MAXLEV) ok = false;
      for (int a = 0; a < Na && ok; a++)
        for (int lo = 0; lo < Ns; lo += seglen) {
          const int hi = std::min(Ns, lo + seglen);
          std::vector<int> trs;
          for (int ns = lo; ns < hi; ns++)
            for (int e = pl->as_first[ns]; e < pl->as_first[ns+1]; e++) {
              const int t = pl->as_trans[e];
              if ((int) pl->trans[(size_t) t*RHB200_TR_NFIELD + RHB200_TR_ATOM] != a) continue;
              size_t q = 0;
              while (q < trs.size() && trs[q] != t) q++;
              if (q == trs.size()) trs.push_back(t);
              as_slot[e] = (int) q;
            }
          if (trs.empty()) continue;
          aseg_atom.push_back(a); aseg_lo.push_back(lo); aseg_hi.push_back(hi);
          slot_tr.insert(slot_tr.end(), trs.begin(), trs.end());
          aseg_slot0.push_back((int) slot_tr.size());
          maxslot = std::max(maxslot, (int) trs.size());
        }
      if (maxslot > NLTE_MAXSLOT) ok = false;
      // measured (profiles/r2_bench_n1.json): the atom-major kernel wins where several transitions of an atom share the
      // wavelengths (H + Ca II: 313 -> 279 ms per 256 columns) and loses where mostly one does (Ca II alone: 251 -> 406)
      size_t pairs = 0;
      for (int ns = 0; ns < Ns; ns++) {
        std::vector<char> seen(Na, 0);
        for (int e = pl->as_first[ns]; e < pl->as_first[ns+1]; e++) {
          const int a = (int) pl->trans[(size_t) pl->as_trans[e]*RHB200_TR_NFIELD + RHB200_TR_ATOM];
          if (!seen[a]) { seen[a] = 1; pairs++; }
        }
      }
      use_atom_rates = ok && !exact_rates && pairs > 0 && (double) nas / (double) pairs >= 3.0;
      if (const char *e = getenv("RHB200_NLTE_GAMMA_ATOM")) use_atom_rates = ok && !exact_rates && atoi(e) != 0;
      if (use_atom_rates) {
        std::vector<int> tr_slot0(Nt + 1, 0), tr_slots;
        for (int t = 0; t < Nt; t++) {
          tr_slot0[t] = (int) tr_slots.size();
          for (size_t q = 0; q < slot_tr.size(); q++) if (slot_tr[q] == t) tr_slots.push_back((int) q);
        }
        tr_slot0[Nt] = (int) tr_slots.size();
        naseg = (int) aseg_atom.size(); nslot_total = (int) slot_tr.size();
        P.naseg = naseg; P.nslot = nslot_total; P.aseg_maxslot = maxslot;
        UPI2(aseg_atom, aseg_atom.data(), std::max(1, naseg)); UPI2(aseg_lo, aseg_lo.data(), std::max(1, naseg));
        UPI2(aseg_hi, aseg_hi.data(), std::max(1, naseg)); UPI2(aseg_slot0, aseg_slot0.data(), naseg + 1);
        UPI2(as_slot, as_slot.data(), std::max(1, nas)); UPI2(tr_slot0, tr_slot0.data(), Nt + 1);
        UPI2(tr_slots, tr_slots.data(), std::max<size_t>(1, tr_slots.size()));
        gamma_smem = ((size_t) maxslot * 4 + 3 * NLTE_MAXLEV) * 64 * sizeof(double);
        RH_CUDA(cudaFuncSetAttribute(nlte_gamma_atom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gamma_smem));
      }
    }
    Norder = pl->Ngorder; Ndelay = std::max(pl->Ngdelay, Norder + 2); Nperiod = std::max(1, pl->Ngperiod);
    prev_off.assign(Na+1, 0);
    for (int a = 0; a < Na; a++) prev_off[a+1] = prev_off[a] + (size_t)(Norder+2) * pl->atom_nlevel[a] * N;
    RH_CHECK(ar.upload(&d_prev_off, prev_off.data(), Na+1));
    return RHB200_OK;
  }
  bool profile_maps_ok = false, exact_rates = false;
  bool has_zeeman = false;                       // set_zeeman() was called: the FULL_STOKES passes are available
  int nprd = 0, nrho = 0, prd_nmax = 0; double prd_limit = 0.0;   // set_prd(): PRD lines, PRD_N_MAX_ITER, PRD_ITER_LIMIT
  int *d_prd_ns = nullptr; unsigned long long *d_drho = nullptr;
  double *d_chi_cQ = nullptr, *d_eta_cQ = nullptr;   // background Q, U, V records of this engine [ncol][Ns][3][N]
  int nseg = 0, naseg = 0, nslot_total = 0;
  bool use_atom_rates = false;
  size_t gamma_smem = 0;

  // doubles of device memory per column that alloc() takes (chunk sizing of the front end)
  // Zeeman patterns of the polarizable ACTIVE lines (Zeeman(), zeeman.c:186-281; line->polarizable, readatom.c:352-368)
  // and the flags Formal() derives from them with StokesMode FULL_STOKES (formal.c:86-95)
  int set_zeeman(const rhb200_nlte_plan *pl, const int *line_pol, const int *line_zoff, const int *zq, const double *zshift,
                 const double *zstrength) {
    if (!line_pol || !line_zoff) { rhb200_set_error("rhb200_nlte: Zeeman tables missing"); return RHB200_EINVAL; }
    DevArena &ar = plan_ar;
    int *di; double *dd;
    const int nz = line_zoff[nline];
    if (nz > 0 && (!zq || !zshift || !zstrength)) { rhb200_set_error("rhb200_nlte: Zeeman tables missing"); return RHB200_EINVAL; }
    RH_CHECK(ar.upload(&di, line_pol, (size_t) std::max(1, nline))); P.line_pol = di;
    RH_CHECK(ar.upload(&di, line_zoff, (size_t) nline + 1)); P.line_zoff = di;
    RH_CHECK(ar.upload(&di, zq, (size_t) nz)); P.zq = di;
    RH_CHECK(ar.upload(&dd, zshift, (size_t) nz)); P.zshift = dd;
    RH_CHECK(ar.upload(&dd, zstrength, (size_t) nz)); P.zstrength = dd;
    std::vector<int> pol_as(Ns, 0), pol_c(Ns, 0), wflags(Ns, 0);
    for (int ns = 0; ns < Ns; ns++)
      for (int e = pl->as_first[ns]; e < pl->as_first[ns+1]; e++) {
        const double *tr = pl->trans + (size_t) pl->as_trans[e]*RHB200_TR_NFIELD;
        if (tr[RHB200_TR_TYPE] == 0.0 && line_pol[(int) tr[RHB200_TR_LINEIDX]]) pol_as[ns] = 1;
      }
    if (c->wav.nlambda != Ns) { rhb200_set_error("rhb200_set_wavelengths() must hold plan->lambda"); return RHB200_ESTATE; }
    RH_CUDA(cudaMemcpy(wflags.data(), c->wav.flags, (size_t) Ns * sizeof(int), cudaMemcpyDeviceToHost));
    for (int ns = 0; ns < Ns; ns++) pol_c[ns] = (wflags[ns] & 2) ? 1 : 0;       // backgrflags.ispolarized
    RH_CHECK(ar.upload(&di, pol_as.data(), (size_t) Ns)); P.pol_as = di;
    RH_CHECK(ar.upload(&di, pol_c.data(), (size_t) Ns)); P.pol_c = di;
    P.stokes = 0; P.stokes_prof = 0; P.stokes_solver = c->s_interpolation_stokes;
    has_zeeman = true;
    return RHB200_OK;
  }
  void set_stokes(bool on) { P.stokes = P.stokes_prof = (on && has_zeeman) ? 1 : 0; }
  // POLARIZATION_FREE: Zeeman profiles, scalar transfer
  void set_stokes_profiles_only(bool on) { P.stokes = 0; P.stokes_prof = (on && has_zeeman) ? 1 : 0; }

  // line->PRD of the ACTIVE lines (readatom.c:255-258) and the keywords of Redistribute() (iterate.c:98-108)
  int set_prd(const rhb200_nlte_plan *pl, const int *line_prd, int PRD_NmaxIter, double PRDiterLimit) {
    std::vector<int> prd_tr, roff(1, 0), tr_prd(Nt, -1), prd_ns(Ns, 0);
    for (int t = 0; t < Nt; t++) {
      const double *tr = pl->trans + (size_t) t*RHB200_TR_NFIELD;
      if (tr[RHB200_TR_TYPE] != 0.0 || !line_prd[(int) tr[RHB200_TR_LINEIDX]]) continue;
      tr_prd[t] = (int) prd_tr.size();
      prd_tr.push_back(t);
      roff.push_back(roff.back() + (int) tr[RHB200_TR_NLAMBDA]);
      for (int ns = (int) tr[RHB200_TR_NBLUE]; ns < (int) tr[RHB200_TR_NBLUE] + (int) tr[RHB200_TR_NLAMBDA]; ns++) prd_ns[ns] = 1;
    }
    nprd = (int) prd_tr.size(); nrho = roff.back(); prd_nmax = PRD_NmaxIter; prd_limit = PRDiterLimit;
    P.nprd = nprd; P.nrho = nrho; P.ns_mask = nullptr;
    if (nprd == 0) return RHB200_OK;
    if (nrank > 1) { rhb200_set_error("PRD lines in a wavelength-sharded solve are not implemented"); return RHB200_EUNSUPPORTED; }
    DevArena &ar = plan_ar;
    int *di;
    RH_CHECK(ar.upload(&di, prd_tr.data(), prd_tr.size())); P.prd_tr = di;
    RH_CHECK(ar.upload(&di, roff.data(), roff.size())); P.prd_roff = di;
    RH_CHECK(ar.upload(&di, tr_prd.data(), tr_prd.size())); P.tr_prd = di;
    RH_CHECK(ar.upload(&di, prd_ns.data(), prd_ns.size())); P.prd_ns = di; d_prd_ns = di;
    return RHB200_OK;
  }

  // Redistribute(PRD_NmaxIter, PRDiterlimit) (redistribute.c:38-106) with Ng order 0: PRDScatter of every PRD line,
  // then solveSpectrum(FALSE, TRUE) over the wavelengths that hold one; per column until its rho changes by < limit.
  // dpops [ncol]: this iteration's dpopsmax (a negative PRD_ITER_LIMIT follows it, iterate.c:103-106)
  int redistribute(const std::vector<double> &dpops) {
    if (nprd == 0 || prd_nmax <= 0) return RHB200_OK;
    cudaStream_t st = c->stream;
    const size_t cN = (size_t) ncol * N;
    std::vector<int> act(active);                   // columns still iterating in Iterate()
    std::vector<unsigned long long> h_drho(ncol);
    int left = 0;
    for (int col = 0; col < ncol; col++) left += act[col];
    bool masked = false;
    for (int it = 1; it <= prd_nmax && left > 0; it++) {
      RH_CUDA(cudaMemsetAsync(d_drho, 0, (size_t) ncol * sizeof(unsigned long long), st));
      { ScopedKernelTimer t(c, RHB200_K_OTHER);
        nlte_prd_scatter_kernel<<<RH_GRID(cN*nrho, 128), 0, st>>>(P, C, ncol, d_drho);
        nlte_setup_kernel<<<RH_GRID(cN*nas, 128), 0, st>>>(P, C, ncol); }
      P.ns_mask = d_prd_ns;
      { ScopedKernelTimer t(c, RHB200_K_OPACITY);
        nlte_opacity_kernel<<<(unsigned) (((cN + 127) / 128) * Ns), 128, 0, st>>>(P, C, ncol);
        if (P.stokes) nlte_opacity_quv_kernel<<<(unsigned) (((cN + 127) / 128) * Ns), 128, 0, st>>>(P, C, ncol); }
      { ScopedKernelTimer t(c, RHB200_K_BEZIER);
        launch_rays(0); }
      { ScopedKernelTimer t(c, RHB200_K_J);
        nlte_J_kernel<<<RH_GRID(cN*Ns, 128), 0, st>>>(P, C, ncol); }
      { ScopedKernelTimer t(c, RHB200_K_GAMMA);
        nlte_prd_rates_kernel<<<RH_GRID(cN*nprd, 64), 0, st>>>(P, C, ncol); }
      P.ns_mask = nullptr;
      RH_CUDA(cudaGetLastError());
      RH_CUDA(cudaMemcpyAsync(h_drho.data(), d_drho, (size_t) ncol * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      RH_CUDA(cudaStreamSynchronize(st));
      bool changed = false;
      for (int col = 0; col < ncol; col++) {
        if (!act[col]) continue;
        double d; memcpy(&d, &h_drho[col], sizeof d);
        const double limit = prd_limit < 0.0 ? std::max(dpops[col], -prd_limit) : prd_limit;
        if (d < limit) { act[col] = 0; left--; changed = true; }
      }
      if (changed && left > 0 && it < prd_nmax) { RH_CUDA(cudaMemcpyAsync(d_active, act.data(), ncol*sizeof(int), cudaMemcpyHostToDevice, st)); masked = true; }
    }
    if (masked) RH_CUDA(cudaMemcpyAsync(d_active, active.data(), ncol*sizeof(int), cudaMemcpyHostToDevice, st));
    return RHB200_OK;
  }

  // B, Bproject() of every ray of this engine from the pyrh rows (d_in [ncol][nrow][N])
  int bproject(const double *d_in, int nrow) {
    if (!has_zeeman) return RHB200_OK;
    nlte_bproject_kernel<<<RH_GRID((size_t) ncol * N, 128), 0, c->stream>>>(P, ncol, nrow, d_in, (double *) C.B, (double *) C.bproj);
    RH_CUDA(cudaGetLastError());
    return RHB200_OK;
  }

  size_t doubles_per_column(bool own_inputs) const {
    size_t d = (size_t) N * ((size_t) ngam + 2*(size_t) Nt + 2*(size_t) nas + 6*(size_t) nray + (size_t) Ns + nphirow + nline +
                             (exact_rates ? 0 : 4*(size_t) std::max(nseg, nslot_total))) + nray + prev_off[Na] + Na;
    if (own_inputs) d += (size_t) N * (2 + 2*(size_t) nlev + Na + ngam + nline + Na + 1 + 4*(size_t) Ns);
    d += (size_t) N * ((size_t) nrho + nline) + 1;
    if (has_zeeman) d += (size_t) N * (1 + 3*(size_t) Nr + 3*(size_t) nphirow + 6*(size_t) Ns + 9*(size_t) nray) + 3*(size_t) nray;
    return d;
  }

  // work arrays of one batch.  own_inputs: also the input arrays (the front end's kernels fill them on the device)
  int alloc(int ncol_, bool device_profiles_, bool own_inputs) {
    ncol = ncol_; device_profiles = device_profiles_;
    if (ncol <= 0) { rhb200_set_error("rhb200_nlte: ncol must be positive"); return RHB200_EINVAL; }
    if (device_profiles && !profile_maps_ok) { rhb200_set_error("profile rows / lambda0 of the line transitions are incomplete"); return RHB200_EINVAL; }
    DevArena &ar = work_ar;
    const size_t cN = (size_t) ncol * N;
    RH_CHECK(ar.alloc(&C.phi, cN*nphirow)); RH_CHECK(ar.alloc(&C.wphi, cN*nline));
    RH_CHECK(ar.alloc(&C.Gamma, cN*ngam)); RH_CHECK(ar.alloc(&C.Rij, cN*Nt, true)); RH_CHECK(ar.alloc(&C.Rji, cN*Nt, true));
    RH_CHECK(ar.alloc(&C.gw, cN*nas*2));
    if (!exact_rates) RH_CHECK(ar.alloc(&C.part, cN*std::max(nseg, nslot_total)*4));
    RH_CHECK(ar.alloc(&C.chi, cN*nray)); RH_CHECK(ar.alloc(&C.S, cN*nray)); RH_CHECK(ar.alloc(&C.I, cN*nray));
    RH_CHECK(ar.alloc(&C.Psi, cN*nray)); RH_CHECK(ar.alloc(&C.scr, cN*nray*2)); RH_CHECK(ar.alloc(&C.dJ, cN*Ns, true));
    RH_CHECK(ar.alloc(&C.Iem, (size_t) ncol*nray, true));
    { double *dd; RH_CHECK(ar.alloc(&dd, cN*std::max(1, nline), true)); C.Qelast = dd; }
    RH_CHECK(ar.alloc(&C.rho, cN*std::max(1, nrho))); RH_CHECK(ar.alloc(&d_drho, (size_t) ncol, true));
    if (has_zeeman) {
      double *dd;
      RH_CHECK(ar.alloc(&dd, cN)); C.B = dd;
      RH_CHECK(ar.alloc(&dd, cN*Nr*3)); C.bproj = dd;
      RH_CHECK(ar.alloc(&C.phiQ, cN*nphirow*3, true));
      RH_CHECK(ar.alloc(&d_chi_cQ, cN*Ns*3, true)); RH_CHECK(ar.alloc(&d_eta_cQ, cN*Ns*3, true));
      C.chi_cQ = d_chi_cQ; C.eta_cQ = d_eta_cQ;
      RH_CHECK(ar.alloc(&C.chiQ, cN*nray*3)); RH_CHECK(ar.alloc(&C.SQ, cN*nray*3)); RH_CHECK(ar.alloc(&C.IQ, cN*nray*3));
      RH_CHECK(ar.alloc(&C.IemQ, (size_t) ncol*nray*3, true));
    }
    active.assign(ncol, 1);
    RH_CHECK(ar.upload(&d_active, active.data(), ncol));
    C.active = d_active;
    RH_CHECK(ar.alloc(&d_prev, (size_t) ncol * prev_off[Na], true));
    RH_CHECK(ar.alloc(&d_dpops, (size_t) ncol * Na, true));
    RH_CHECK(ar.alloc(&d_dJmax, ncol, true));
    if (own_inputs) {
      double *dd;
#define OWN(field, n) RH_CHECK(ar.alloc(&dd, (size_t) (n))); C.field = dd
      OWN(T, cN); OWN(height, cN); OWN(nstar, cN*nlev); OWN(ntotal, cN*Na); OWN(C, cN*ngam);
      OWN(adamp, cN*nline); OWN(vbroad, cN*Na); OWN(vel, cN);
      OWN(chi_c, cN*Ns); OWN(eta_c, cN*Ns); OWN(sca_c, cN*Ns);
#undef OWN
      RH_CHECK(ar.alloc(&C.n, cN*nlev)); RH_CHECK(ar.alloc(&C.J, cN*Ns, true));
    }
    return RHB200_OK;
  }

  int bind_host(const rhb200_nlte_columns *cols) {
    DevArena &ar = work_ar;
    const size_t cN = (size_t) ncol * N;
    double *dd;
#define UPC(field, src, n) RH_CHECK(ar.upload(&dd, src, (size_t) (n))); C.field = dd
    UPC(T, cols->T, cN); UPC(height, cols->height, cN);
    UPC(nstar, cols->nstar, cN*nlev); UPC(ntotal, cols->ntotal, cN*Na); UPC(C, cols->C, cN*ngam);
    if (device_profiles) {
      if (!cols->adamp || !cols->vbroad || !cols->vel) { rhb200_set_error("phi == NULL needs adamp, vbroad and vel"); return RHB200_EINVAL; }
      UPC(adamp, cols->adamp, cN*nline); UPC(vbroad, cols->vbroad, cN*Na); UPC(vel, cols->vel, cN);
    } else {
      if (!cols->wphi) { rhb200_set_error("wphi missing"); return RHB200_EINVAL; }
      RH_CUDA(cudaMemcpy(C.phi, cols->phi, cN*nphirow*sizeof(double), cudaMemcpyHostToDevice));
      RH_CUDA(cudaMemcpy(C.wphi, cols->wphi, cN*nline*sizeof(double), cudaMemcpyHostToDevice));
    }
    UPC(chi_c, cols->chi_c, cN*Ns); UPC(eta_c, cols->eta_c, cN*Ns); UPC(sca_c, cols->sca_c, cN*Ns);
#undef UPC
    RH_CHECK(ar.upload(&C.n, cols->n, cN*nlev));
    RH_CHECK(ar.upload(&C.J, cols->J, cN*Ns));
    return RHB200_OK;
  }

  void launch_rays(int eval_operator) {
    static int ray_minb = -1;
    if (ray_minb < 0) { const char *e = getenv("RHB200_NLTE_RAY_MINB"); ray_minb = e ? atoi(e) : 6; }   // with the register-window solver: 8 -> 5.05 / 2.77 ms, 6 -> 4.30 / 2.45, 4 -> 4.49 / 2.69 (configs[4] sample / configs[3], 256 columns)
    const unsigned blocks = (unsigned) (((size_t) ncol*nray + 127) / 128);
#define RH_RAYS(S, M) nlte_ray_kernel<S, M><<<blocks, 128, 0, c->stream>>>(P, C, ncol, eval_operator)
#define RH_RAYS_M(S) do { if (ray_minb >= 8) RH_RAYS(S, 8); else if (ray_minb >= 6) RH_RAYS(S, 6); else if (ray_minb == 5) RH_RAYS(S, 5); else RH_RAYS(S, 4); } while (0)
    if (P.solver == RHB200_S_LINEAR) RH_RAYS_M(RHB200_S_LINEAR);
    else if (P.solver == RHB200_S_PARABOLIC) RH_RAYS_M(RHB200_S_PARABOLIC);
    else RH_RAYS_M(RHB200_S_BEZIER3);
#undef RH_RAYS_M
#undef RH_RAYS
    if (P.stokes) { ScopedKernelTimer t(c, RHB200_K_DELO); nlte_ray_stokes_kernel<<<blocks, 128, 0, c->stream>>>(P, C, ncol, eval_operator); }
  }

  // Profile() + wphi (when evaluated here), the per-entry weights, NgInit (accelerate.c:57-59)
  int prepare(double *phi_out, double *wphi_out, bool ng_init) {
    cudaStream_t st = c->stream;
    const size_t cN = (size_t) ncol * N;
    if (device_profiles) {
      { ScopedKernelTimer t(c, RHB200_K_PREP);
        nlte_profile_kernel<<<RH_GRID(cN*nphirow, 128), 0, st>>>(P, C, ncol); }
      { ScopedKernelTimer t(c, RHB200_K_PREP);
        nlte_wphi_kernel<<<RH_GRID(cN*nline, 64), 0, st>>>(P, C, ncol); }
      RH_CUDA(cudaGetLastError());
      if (phi_out) {
        RH_CUDA(cudaStreamSynchronize(st));
        RH_CUDA(cudaMemcpy(phi_out, C.phi, cN*nphirow*sizeof(double), cudaMemcpyDeviceToHost));
      }
      if (wphi_out) {
        RH_CUDA(cudaStreamSynchronize(st));
        RH_CUDA(cudaMemcpy(wphi_out, C.wphi, cN*nline*sizeof(double), cudaMemcpyDeviceToHost));
      }
    }
    if (nprd > 0 && ng_init) {                       // Profile() of a PRD line starts from rho = 1 (profile.c:83-103)
      nlte_fill_kernel<<<RH_GRID(cN*nrho, 256), 0, st>>>(C.rho, cN*nrho, 1.0);
      RH_CUDA(cudaGetLastError());
    }
    { ScopedKernelTimer t(c, RHB200_K_OTHER);
      nlte_setup_kernel<<<RH_GRID(cN*nas, 128), 0, st>>>(P, C, ncol); }
    RH_CUDA(cudaGetLastError());
    if (ng_init)
      for (int col = 0; col < ncol; col++)
        for (int a = 0; a < Na; a++)
          RH_CUDA(cudaMemcpyAsync(d_prev + (size_t) col*prev_off[Na] + prev_off[a],
                                  C.n + ((size_t) col*nlev + lev_off[a])*N, (size_t) nlevel[a]*N*sizeof(double),
                                  cudaMemcpyDeviceToDevice, st));
    return RHB200_OK;
  }

  // solveSpectrum(FALSE, FALSE) repeated: initScatter (update_J 1), the passes after Iterate() (2), the final formal
  // solution (0).  Iem_host [ncol][Ns][Nr] or NULL; d_Iem_spec: device [ncol][Ns] emergent intensity of ray mu = 0 or NULL
  int scatter(int NmaxScatter, int update_J, double limit, int *nscat_out, double *Iem_host, double *d_Iem_spec,
              double *d_quv_spec = nullptr /* device [ncol][3][Ns] emergent Q, U, V of ray mu = 0 */) {
    cudaStream_t st = c->stream;
    const size_t cN = (size_t) ncol * N;
    std::vector<int> nscat(ncol, 0);
    std::vector<double> h_dJ(ncol, 0.0);
    std::vector<int> act(ncol, 1);
    int nact_s = ncol;
    for (int it = 0; it < NmaxScatter && nact_s > 0; it++) {
      { ScopedKernelTimer t(c, RHB200_K_OPACITY);
        nlte_opacity_kernel<<<(unsigned) (((cN + 127) / 128) * Ns), 128, 0, st>>>(P, C, ncol);
        if (P.stokes) nlte_opacity_quv_kernel<<<(unsigned) (((cN + 127) / 128) * Ns), 128, 0, st>>>(P, C, ncol); }
      { ScopedKernelTimer t(c, RHB200_K_BEZIER);
        launch_rays(0); }
      if (update_J) {
        { ScopedKernelTimer t(c, RHB200_K_J);
          nlte_J_kernel<<<RH_GRID(cN*Ns, 128), 0, st>>>(P, C, ncol); }
        nlte_dJmax_kernel<<<ncol, 256, 0, st>>>(P, C, ncol, d_dJmax);
        RH_CUDA(cudaGetLastError());
        RH_CHECK(allreduce(d_dJmax, ncol, RHB200_REDUCE_MAX));
        RH_CUDA(cudaMemcpyAsync(h_dJ.data(), d_dJmax, ncol*sizeof(double), cudaMemcpyDeviceToHost, st));
        RH_CUDA(cudaStreamSynchronize(st));
      }
      bool changed = false;
      for (int col = 0; col < ncol; col++) {
        if (!act[col]) continue;
        nscat[col] = it + 1;
        // initscatter.c:65 stops on dJmax < limit; the post-Iterate loop of rhf1d() on <= (pyrh_compute1dray.c:335)
        if (!update_J || (update_J == 2 ? h_dJ[col] <= limit : h_dJ[col] < limit)) { act[col] = 0; nact_s--; changed = true; }
      }
      if (changed && nact_s > 0) RH_CUDA(cudaMemcpyAsync(d_active, act.data(), ncol*sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (NmaxScatter > 0) {
      RH_CUDA(cudaStreamSynchronize(st));
      std::vector<int> all(ncol, 1);
      RH_CUDA(cudaMemcpy(d_active, all.data(), ncol*sizeof(int), cudaMemcpyHostToDevice));   // all columns active again
      if (Iem_host) {
        // emergent intensity per (column, wavelength, mu): the up-ray of angle-dependent wavelengths
        std::vector<double> h_Iem((size_t) ncol*nray);
        RH_CHECK(allreduce(C.Iem, (size_t) ncol*nray, RHB200_REDUCE_SUM));     // foreign rays hold 0
        RH_CUDA(cudaStreamSynchronize(st));
        RH_CUDA(cudaMemcpy(h_Iem.data(), C.Iem, h_Iem.size()*sizeof(double), cudaMemcpyDeviceToHost));
        for (int col = 0; col < ncol; col++)
          for (int r = 0; r < nray; r++)
            if (!angle_dep[ray_ns[r]] || ray_dir[r] == 1)
              Iem_host[((size_t) col*Ns + ray_ns[r])*Nr + ray_mu[r]] = h_Iem[(size_t) col*nray + r];
      }
      if (d_Iem_spec) {
        nlte_pack_spectrum_kernel<<<RH_GRID((size_t) ncol*Ns, 128), 0, st>>>(P, C, ncol, d_Iem_spec);
        if (d_quv_spec && has_zeeman) nlte_pack_quv_kernel<<<RH_GRID((size_t) ncol*Ns, 128), 0, st>>>(P, C, ncol, d_quv_spec);
        else if (d_quv_spec) RH_CUDA(cudaMemsetAsync(d_quv_spec, 0, (size_t) ncol * 3 * Ns * sizeof(double), st));
        RH_CUDA(cudaGetLastError());
      }
    }
    if (nscat_out) memcpy(nscat_out, nscat.data(), ncol*sizeof(int));
    return RHB200_OK;
  }

  // Iterate(): the MALI loop (iterate.c:48-143)
  int iterate(int NmaxIter, double iterLimit, int *niter_out, double *dpops_hist, int dump_iter,
              double *gamma_dump, double *rates_dump) {
    cudaStream_t st = c->stream;
    const size_t cN = (size_t) ncol * N;
    std::vector<double> h_dpops((size_t) ncol * Na);
    std::vector<int> niter(ncol, 0);
    active.assign(ncol, 1);
    int nactive = ncol;
    for (int it = 1; it <= NmaxIter && nactive > 0; it++) {
      { ScopedKernelTimer t(c, RHB200_K_OTHER);
        nlte_gamma_init_kernel<<<RH_GRID(cN*ngam, 256), 0, st>>>(P, C, ncol); }
      { ScopedKernelTimer t(c, RHB200_K_OPACITY);
        nlte_opacity_kernel<<<(unsigned) (((cN + 127) / 128) * Ns), 128, 0, st>>>(P, C, ncol);
        if (P.stokes) nlte_opacity_quv_kernel<<<(unsigned) (((cN + 127) / 128) * Ns), 128, 0, st>>>(P, C, ncol); }
      { ScopedKernelTimer t(c, RHB200_K_BEZIER);
        launch_rays(1); }
      { ScopedKernelTimer t(c, RHB200_K_GAMMA);
        if (exact_rates) {
          if (P.stokes) nlte_gamma_kernel<false, true><<<dim3((unsigned) ((cN + 63) / 64), (unsigned) Nt), 64, 0, st>>>(P, C, ncol);
          else nlte_gamma_kernel<false, false><<<dim3((unsigned) ((cN + 63) / 64), (unsigned) Nt), 64, 0, st>>>(P, C, ncol);
        } else if (use_atom_rates && !P.stokes) {
          nlte_gamma_atom_kernel<<<dim3((unsigned) ((cN + 63) / 64), (unsigned) naseg), 64, gamma_smem, st>>>(P, C, ncol);
          nlte_gamma_slot_sum_kernel<<<RH_GRID(cN*Nt, 128), 0, st>>>(P, C, ncol);
        } else {
          if (P.stokes) nlte_gamma_kernel<true, true><<<dim3((unsigned) ((cN + 63) / 64), (unsigned) nseg), 64, 0, st>>>(P, C, ncol);
          else nlte_gamma_kernel<true, false><<<dim3((unsigned) ((cN + 63) / 64), (unsigned) nseg), 64, 0, st>>>(P, C, ncol);
          nlte_gamma_sum_kernel<<<RH_GRID(cN*Nt, 128), 0, st>>>(P, C, ncol);
        } }
      { ScopedKernelTimer t(c, RHB200_K_J);
        nlte_J_kernel<<<RH_GRID(cN*Ns, 128), 0, st>>>(P, C, ncol); }
      // the exchange step of a wavelength-sharded atmosphere: radiative rates add up over ranks
      if (c->nccl_comm && nrank > 1) RH_CHECK(rh_nccl_group(0));           // one fused launch for the three buffers
      RH_CHECK(allreduce(C.Gamma, cN*ngam, RHB200_REDUCE_SUM));
      RH_CHECK(allreduce(C.Rij, cN*Nt, RHB200_REDUCE_SUM));
      RH_CHECK(allreduce(C.Rji, cN*Nt, RHB200_REDUCE_SUM));
      if (c->nccl_comm && nrank > 1) RH_CHECK(rh_nccl_group(1));
      if (it == dump_iter) {
        RH_CUDA(cudaStreamSynchronize(st));
        if (gamma_dump) RH_CUDA(cudaMemcpy(gamma_dump, C.Gamma, cN*ngam*sizeof(double), cudaMemcpyDeviceToHost));
        if (rates_dump) {
          RH_CUDA(cudaMemcpy(rates_dump, C.Rij, cN*Nt*sizeof(double), cudaMemcpyDeviceToHost));
          RH_CUDA(cudaMemcpy(rates_dump + cN*Nt, C.Rji, cN*Nt*sizeof(double), cudaMemcpyDeviceToHost));
        }
      }
      { ScopedKernelTimer t(c, RHB200_K_STATEQ);
        if (maxnl <= 8) nlte_statequil_kernel<8><<<RH_GRID(cN*Na, 64), 0, st>>>(P, C, ncol, isum);
        else if (maxnl <= 16) nlte_statequil_kernel<16><<<RH_GRID(cN*Na, 64), 0, st>>>(P, C, ncol, isum);
        else nlte_statequil_kernel<32><<<RH_GRID(cN*Na, 64), 0, st>>>(P, C, ncol, isum); }
      { ScopedKernelTimer t(c, RHB200_K_NG);
        nlte_ng_kernel<<<ncol*Na, 128, 0, st>>>(P, C, ncol, d_prev, d_prev_off, Norder, Ndelay, Nperiod, it, d_dpops); }
      RH_CUDA(cudaGetLastError());
      RH_CUDA(cudaMemcpyAsync(h_dpops.data(), d_dpops, h_dpops.size()*sizeof(double), cudaMemcpyDeviceToHost, st));
      RH_CUDA(cudaStreamSynchronize(st));
      if (nprd > 0) {                               // iterate.c:98-108: before the convergence test of this iteration
        std::vector<double> dcol(ncol, 0.0);
        for (int col = 0; col < ncol; col++)
          for (int a = 0; a < Na; a++) dcol[col] = std::max(dcol[col], h_dpops[(size_t) col*Na + a]);
        RH_CHECK(redistribute(dcol));
      }
      bool changed = false;
      for (int col = 0; col < ncol; col++) {
        if (!active[col]) continue;
        double d = 0.0;
        for (int a = 0; a < Na; a++) d = std::max(d, h_dpops[(size_t) col*Na + a]);
        niter[col] = it;
        if (dpops_hist) dpops_hist[(size_t) col*NmaxIter + it-1] = d;
        if (d < iterLimit) { active[col] = 0; nactive--; changed = true; }     // iterate.c:114
      }
      if (changed) RH_CUDA(cudaMemcpyAsync(d_active, active.data(), ncol*sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (nrank > 1) {                                   // every rank returns the full J
      nlte_zero_foreign_J_kernel<<<RH_GRID(cN*Ns, 256), 0, st>>>(P, C, ncol);
      RH_CUDA(cudaGetLastError());
      RH_CHECK(allreduce(C.J, cN*Ns, RHB200_REDUCE_SUM));
    }
    RH_CUDA(cudaStreamSynchronize(st));
    {                                                  // frozen columns thaw: later passes treat every column
      std::vector<int> all(ncol, 1);
      RH_CUDA(cudaMemcpy(d_active, all.data(), ncol*sizeof(int), cudaMemcpyHostToDevice));
    }
    if (niter_out) memcpy(niter_out, niter.data(), ncol*sizeof(int));
    return RHB200_OK;
  }
};

static int nlte_run(rhb200_ctx *c, const rhb200_nlte_plan *pl, int ncol,
                    const rhb200_nlte_columns *cols, int NmaxScatter, int update_J, int NmaxIter, double iterLimit,
                    int *niter_out, double *dpops_hist, int dump_iter,
                    double *gamma_dump, double *rates_dump, double *phi_out, double *wphi_out,
                    double *Iem_out, int *nscatter_out)
{
  if (!c || !pl || !cols) { rhb200_set_error("null argument"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  if (NmaxIter < 0) { rhb200_set_error("rhb200_nlte_iterate: bad sizes"); return RHB200_EINVAL; }
  NlteEngine E;
  RH_CHECK(E.build(c, pl));
  RH_CHECK(E.alloc(ncol, cols->phi == nullptr, false));
  RH_CHECK(E.bind_host(cols));
  RH_CHECK(E.prepare(phi_out, wphi_out, true));
  RH_CHECK(E.scatter(NmaxScatter, update_J, iterLimit, nscatter_out, Iem_out, nullptr));
  RH_CHECK(E.iterate(NmaxIter, iterLimit, niter_out, dpops_hist, dump_iter, gamma_dump, rates_dump));
  const size_t cN = (size_t) ncol * E.N;
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CUDA(cudaMemcpy(cols->n, E.C.n, cN*E.nlev*sizeof(double), cudaMemcpyDeviceToHost));
  RH_CUDA(cudaMemcpy(cols->J, E.C.J, cN*E.Ns*sizeof(double), cudaMemcpyDeviceToHost));
  return RHB200_OK;
}

// contiguous wavelength chunk of rank `rank`: boundaries where the cumulative ray count crosses rank*nray/nrank
void rh_nlte_shard_range(int Nspect, const int *ray_off, int rank, int nrank, int *lo, int *hi)
{
  const long nray = ray_off[Nspect];
  auto bound = [&](int r) {
    if (r <= 0) return 0;
    if (r >= nrank) return Nspect;
    const long target = nray * r / nrank;
    int ns = 0;
    while (ns < Nspect && ray_off[ns] < target) ns++;
    return ns;
  };
  *lo = bound(rank); *hi = bound(rank + 1);
}

extern "C" int rhb200_nlte_set_exact_rates(rhb200_ctx *c, int exact)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  c->nlte_exact_rates = exact != 0;
  return RHB200_OK;
}

// ---- native NCCL for the wavelength shard: the library loads libnccl at run time (dlopen), nothing is linked
#include <dlfcn.h>
namespace {
struct NcclApi {
  void *h = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  struct Id { char b[128]; };
  int (*CommInitRank)(void **, int, Id, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int nccl_load()
{
  if (g_nccl.h) return RHB200_OK;
  const char *names[] = {getenv("RHB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) { if (n && (g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
  if (!g_nccl.h) { rhb200_set_error("cannot load libnccl (set RHB200_NCCL_LIB): %s", dlerror()); return RHB200_ESTATE; }
#define SYM(f, name) *(void **) &g_nccl.f = dlsym(g_nccl.h, name); if (!g_nccl.f) { rhb200_set_error("libnccl lacks %s", name); g_nccl.h = nullptr; return RHB200_ESTATE; }
  SYM(AllReduce, "ncclAllReduce") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return RHB200_OK;
}
}  // namespace
int rh_nccl_allreduce(rhb200_ctx *c, double *buf, size_t count, int op)
{
  const int rc = g_nccl.AllReduce(buf, buf, count, /* ncclDouble */ 8, op == RHB200_REDUCE_MAX ? /* ncclMax */ 2 : /* ncclSum */ 0,
                                  c->nccl_comm, c->stream);
  if (rc != 0) { rhb200_set_error("ncclAllReduce: %s", g_nccl.GetErrorString(rc)); return RHB200_ECUDA; }
  return RHB200_OK;
}
int rh_nccl_group(int end)
{
  const int rc = end ? g_nccl.GroupEnd() : g_nccl.GroupStart();
  if (rc != 0) { rhb200_set_error("ncclGroup%s: %s", end ? "End" : "Start", g_nccl.GetErrorString(rc)); return RHB200_ECUDA; }
  return RHB200_OK;
}
extern "C" int rhb200_nccl_unique_id(char id[128])
{
  RH_CHECK(nccl_load());
  const int rc = g_nccl.GetUniqueId(id);
  if (rc != 0) { rhb200_set_error("ncclGetUniqueId: %s", g_nccl.GetErrorString(rc)); return RHB200_ECUDA; }
  return RHB200_OK;
}
extern "C" int rhb200_nlte_set_shard_nccl_id(rhb200_ctx *c, int rank, int nrank, const char id[128])
{
  if (!c || !id || nrank < 1 || rank < 0 || rank >= nrank) { rhb200_set_error("bad shard arguments"); return RHB200_EINVAL; }
  RH_CHECK(nccl_load());
  RH_CUDA(cudaSetDevice(c->device));
  NcclApi::Id uid;
  memcpy(uid.b, id, sizeof uid.b);
  void *comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, nrank, uid, rank);
  if (rc != 0) { rhb200_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(rc)); return RHB200_ECUDA; }
  c->nccl_comm = comm; c->nccl_owned = true; c->shard_rank = rank; c->shard_nrank = nrank; c->shard_fn = nullptr; c->shard_user = nullptr;
  return RHB200_OK;
}
extern "C" int rhb200_nlte_set_shard_nccl(rhb200_ctx *c, int rank, int nrank, void *nccl_comm)
{
  if (!c || nrank < 1 || rank < 0 || rank >= nrank || (nrank > 1 && !nccl_comm)) { rhb200_set_error("bad shard arguments"); return RHB200_EINVAL; }
  if (nrank > 1) RH_CHECK(nccl_load());
  c->nccl_comm = nrank > 1 ? nccl_comm : nullptr; c->nccl_owned = false;
  c->shard_rank = rank; c->shard_nrank = nrank; c->shard_fn = nullptr; c->shard_user = nullptr;
  return RHB200_OK;
}
void rh_nccl_release(rhb200_ctx *c)
{
  if (c->nccl_comm && c->nccl_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
  c->nccl_comm = nullptr; c->nccl_owned = false;
}

extern "C" int rhb200_nlte_set_shard(rhb200_ctx *c, int rank, int nrank, rhb200_allreduce_fn fn, void *user)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  if (nrank < 1 || rank < 0 || rank >= nrank || (nrank > 1 && !fn)) { rhb200_set_error("bad shard arguments"); return RHB200_EINVAL; }
  rh_nccl_release(c);
  c->shard_rank = rank; c->shard_nrank = nrank; c->shard_fn = fn; c->shard_user = user;
  return RHB200_OK;
}

extern "C" int rhb200_nlte_shard_range(const rhb200_nlte_plan *pl, int rank, int nrank, int *ns_lo, int *ns_hi)
{
  if (!pl || !ns_lo || !ns_hi || nrank < 1 || rank < 0 || rank >= nrank) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  std::vector<int> ray_off(pl->Nspect + 1, 0);
  for (int ns = 0; ns < pl->Nspect; ns++) {
    bool bb = false;
    for (int e = pl->as_first[ns]; e < pl->as_first[ns+1]; e++)
      if (pl->trans[(size_t) pl->as_trans[e]*RHB200_TR_NFIELD + RHB200_TR_TYPE] == 0.0) bb = true;
    const bool ad = pl->moving && (bb || pl->bg_hasline[ns]);
    ray_off[ns+1] = ray_off[ns] + pl->Nrays * (ad ? 2 : 1);
  }
  rh_nlte_shard_range(pl->Nspect, ray_off.data(), rank, nrank, ns_lo, ns_hi);
  return RHB200_OK;
}

extern "C" int rhb200_nlte_iterate(rhb200_ctx *c, const rhb200_nlte_plan *pl, int ncol,
                                   const rhb200_nlte_columns *cols, int NmaxScatter, int NmaxIter, double iterLimit,
                                   int *niter_out, double *dpops_hist, int dump_iter,
                                   double *gamma_dump, double *rates_dump, double *phi_out, double *wphi_out)
{
  return nlte_run(c, pl, ncol, cols, NmaxScatter, 1, NmaxIter, iterLimit, niter_out, dpops_hist, dump_iter,
                  gamma_dump, rates_dump, phi_out, wphi_out, nullptr, nullptr);
}

extern "C" int rhb200_nlte_formal(rhb200_ctx *c, const rhb200_nlte_plan *pl, int ncol,
                                  const rhb200_nlte_columns *cols, int npass, int update_J, double dJlimit,
                                  double *Iem, int *npass_done)
{
  if (npass < 1) { rhb200_set_error("npass must be >= 1"); return RHB200_EINVAL; }
  return nlte_run(c, pl, ncol, cols, npass, update_J, 0, dJlimit, nullptr, nullptr, 0, nullptr, nullptr,
                  nullptr, nullptr, Iem, npass_done);
}

#include "rhb200_nlte_front.cuh"
