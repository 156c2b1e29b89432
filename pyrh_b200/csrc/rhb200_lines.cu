// rhb200_lines.cu -- LTE Kurucz-line opacity on the device.
//
// Reference:
//   LTEpops_elem   rh/ltepops.c:116-159     Saha chain per element
//   Linear / Hunt  rh/linear.c:22-58, rh/hunt.c:17-78
//   RLKProfile     rh/kurucz.c:729-828      Voigt-Faraday x Zeeman pattern
//   rlk_opacity    rh/kurucz.c:511-725      chi/eta (I,Q,U,V) of all lines in the window
//   Background     rh/background.c:476-537  chi_c = chi_ai + lines
//   Formal         rh/rhf1d/formal.c:178-208  S = eta/chi ; StokesK stokesopac.c:72-77
//
// B200 design.  The reference evaluates rlk_opacity once per
// (wavelength, mu, direction) and inside it recomputes, for every depth, quantities that do
// not depend on wavelength at all (Doppler width, damping, Saha-Boltzmann populations: two
// pow(), two exp(), one sqrt per line and depth).  Here a tiny "prep" kernel evaluates them once
// per (column, line, depth) with the *same expressions*, and the opacity kernel -- one thread
// per ray-point, consecutive DEPTHS of one wavelength along the warp -- only does the
// wavelength-dependent part: the Zeeman-component loop of Humlicek W(z) evaluations.
#include "rhb200_common.cuh"
#include "rhb200_voigt.cuh"
#include "rhb200_div.cuh"

namespace {

// Linear(), linear.c:22-58 with hunt=TRUE: for xmin < x < xmax the bracket returned by Hunt
// (hunt.c) is the unique j with xt[j] <= x < xt[j+1]; found by the same bisection (hunt.c:59-68)
__device__ __forceinline__ double linear_interp(int nt, const double *__restrict__ xt,
                                                const double *__restrict__ yt, double x)
{
  if (x <= xt[0]) return yt[0];
  if (x >= xt[nt-1]) return yt[nt-1];
  int lo = 0, hi = nt;
  while (hi - lo > 1) {
    const int mid = (hi + lo) >> 1;
    if (x >= xt[mid]) lo = mid; else hi = mid;
  }
  const double fx = (xt[lo+1] - x) / (xt[lo+1] - xt[lo]);
  return fx*yt[lo] + (1 - fx)*yt[lo+1];
}

// one thread per (column, depth)
__global__ void __launch_bounds__(128)
prep_kernel(int ncol, int ndep, double muz, int moving, int rlkscatter,
            int nline, int nelem, int npf,
            const double *__restrict__ lines, const double *__restrict__ elems,
            const double *__restrict__ pf, const double *__restrict__ Tpf,
            const double *__restrict__ atmos,
            double *__restrict__ elem_n,      // [ncol][nelem][MAXSTAGE][ndep]
            double *__restrict__ lineprep)    // [ncol][nline][LP_NFIELD][ndep]
{
  const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t) ncol * ndep) return;
  const int col = (int) (idx / ndep), k = (int) (idx - (size_t) col * ndep);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double T = at[RHB200_AT_T*ndep + k], ne = at[RHB200_AT_NE*ndep + k];
  const double vturb = at[RHB200_AT_VTURB*ndep + k], vel = at[RHB200_AT_VEL*ndep + k];
  const double B = at[RHB200_AT_B*ndep + k], nHtot = at[RHB200_AT_NHTOT*ndep + k];
  const double np = at[RHB200_AT_NP*ndep + k];

  // ---- LTEpops_elem, ltepops.c:116-159
  const double C1 = (RH_HPLANCK/(2.0*RH_PI*RH_M_ELECTRON)) * (RH_HPLANCK/RH_KBOLTZMANN);
  const double CT_ne = 2.0 * rhm::rh_pow(C1/T, -1.5) / ne;
  for (int ie = 0; ie < nelem; ie++) {
    const double *e = elems + (size_t) ie * RHB200_RE_NFIELD;
    const int nst = (int) e[RHB200_RE_NSTAGE], pfrow = (int) e[RHB200_RE_PFROW];
    double *n = elem_n + (((size_t) col*nelem + ie) * RHB200_RE_MAXSTAGE) * ndep + k;
    double sum = 1.0, nprev = 1.0;
    double Uk = linear_interp(npf, Tpf, pf + (size_t) pfrow*npf, T);
    n[0] = 1.0;
    for (int i = 1; i < nst; i++) {
      const double Ukp1 = linear_interp(npf, Tpf, pf + (size_t)(pfrow+i)*npf, T);
      const double ni = nprev * CT_ne *
        rhm::rh_exp(Ukp1 - Uk - e[RHB200_RE_IONPOT0 + i-1]/(RH_KBOLTZMANN*T));
      n[(size_t) i*ndep] = ni;
      sum += ni;
      nprev = ni;
      Uk = Ukp1;
    }
    const double n0 = e[RHB200_RE_ABUND] * nHtot / sum;
    n[0] = n0;
    for (int i = 1; i < nst; i++) n[(size_t) i*ndep] *= n0;
  }

  // ---- wavelength-independent part of RLKProfile / rlk_opacity per line
  const double hc = RH_HPLANCK * RH_CLIGHT, fourPI = 4.0 * RH_PI, hc_4PI = hc / fourPI;
  for (int nl = 0; nl < nline; nl++) {
    const double *L = lines + (size_t) nl * RHB200_RL_NFIELD;
    const int ie = (int) L[RHB200_RL_ELEM], st = (int) L[RHB200_RL_STAGE];
    const double *e = elems + (size_t) ie * RHB200_RE_NFIELD;
    const double lambda0 = L[RHB200_RL_LAMBDA0];
    double *out = lineprep + ((size_t) col*nline + nl) * LP_NFIELD * ndep + k;

    const double vtherm = 2.0*RH_KBOLTZMANN/(RH_AMU * e[RHB200_RE_WEIGHT]);     // kurucz.c:745-746
    const double vbroad = sqrt(vtherm*T + vturb*vturb);
    const double w = moving ? (muz * vel) / vbroad : 0.0;                        // kurucz.c:749-754
    const double sv = 1.0 / (RH_SQRTPI * vbroad);
    double adamp = 0.0;
    if (L[RHB200_RL_GRAD] != 0.0) {                                              // kurucz.c:758-775
      double GvdW;
      const int vdw = (int) L[RHB200_RL_VDWAALS];
      if (vdw == 0)      GvdW = L[RHB200_RL_CROSS] * rhm::rh_pow(T, 0.3);
      else if (vdw == 2) GvdW = L[RHB200_RL_CROSS] * rhm::rh_pow(T, (1.0 - L[RHB200_RL_ALPHA])/2.0);
      else               GvdW = L[RHB200_RL_GVDW];
      adamp = (L[RHB200_RL_GRAD] + L[RHB200_RL_GSTARK] * ne + GvdW * (nHtot - np)) *
        (lambda0 * RH_NM_TO_M) / (4.0*RH_PI * vbroad);
    }
    const double vB = (RH_LARMOR * lambda0) * B / vbroad;                        // kurucz.c:783

    // Boltzmann factors, kurucz.c:638-640,666,675-680
    const double hc_la      = (RH_HPLANCK * RH_CLIGHT) / (lambda0 * RH_NM_TO_M);
    const double Bijhc_4PI  = hc_4PI * L[RHB200_RL_BIJ] * L[RHB200_RL_ISO_FRAC] *
                              L[RHB200_RL_HFS_FRAC] * L[RHB200_RL_GI];
    const double twohnu3_c2 = L[RHB200_RL_AJI] / L[RHB200_RL_BJI];
    const double pfk = linear_interp(npf, Tpf, pf + (size_t)((int) e[RHB200_RE_PFROW] + st)*npf, T);
    const double kT = 1.0 / (RH_KBOLTZMANN * T);
    const double nstage = elem_n[(((size_t) col*nelem + ie) * RHB200_RE_MAXSTAGE + st) * ndep + k];
    const double ni_gi = nstage * rhm::rh_exp(-L[RHB200_RL_EI]*kT - pfk);
    const double nj_gj = ni_gi * rhm::rh_exp(-hc_la * kT);
    out[(size_t) LP_VBROAD*ndep] = vbroad;
    out[(size_t) LP_ADAMP*ndep]  = adamp;
    out[(size_t) LP_VB*ndep]     = vB;
    out[(size_t) LP_W*ndep]      = w;
    out[(size_t) LP_SV*ndep]     = sv;
    out[(size_t) LP_CHIL*ndep]   = Bijhc_4PI * (ni_gi - nj_gj);
    out[(size_t) LP_ETAL*ndep]   = Bijhc_4PI * twohnu3_c2 * nj_gj;
    double epsilon = 1.0;
    if (rlkscatter) {                                                            // kurucz.c:641-652, 682-685
      const double Cc = 2 * RH_PI * (RH_Q_ELECTRON/8.854187817E-12) * (RH_Q_ELECTRON/RH_M_ELECTRON) / RH_CLIGHT;
      const double l0m = lambda0 * RH_NM_TO_M;
      const double x = (st == 0) ? 0.68 : 0.0;
      const double C3 = Cc / (((st == 0) ? 2.15E-6 : 3.96E-6) * (l0m * l0m));
      const double dE = L[RHB200_RL_EJ] - L[RHB200_RL_EI];
      epsilon = 1.0 / (1.0 + C3 * rhm::rh_pow(T, 1.5) / (ne * rhm::rh_pow(RH_KBOLTZMANN * T / dE, 1 + x)));
    }
    out[(size_t) LP_EPS*ndep]    = epsilon;
  }
}

struct LineSums { double chi[4], eta[4], scatt; };

// Zeeman components are read with a warp-uniform index.  Small line lists pass them BY VALUE in
// the kernel parameter block (constant bank: one broadcast LDC per read, no L1 traffic);
// larger ones fall back to global memory.
#define RH_ZPARAM_MAX 160
struct ZeemanParam {
  double shift[RH_ZPARAM_MAX], strength[RH_ZPARAM_MAX];
  signed char q[RH_ZPARAM_MAX];
};
struct ZeemanGlobal { const int *q; const double *shift, *strength; };
__device__ __forceinline__ int    zq_of(const ZeemanParam &z, int i) { return z.q[i]; }
__device__ __forceinline__ double zshift_of(const ZeemanParam &z, int i) { return z.shift[i]; }
__device__ __forceinline__ double zstrength_of(const ZeemanParam &z, int i) { return z.strength[i]; }
__device__ __forceinline__ int    zq_of(const ZeemanGlobal &z, int i) { return __ldg(z.q + i); }
__device__ __forceinline__ double zshift_of(const ZeemanGlobal &z, int i) { return __ldg(z.shift + i); }
__device__ __forceinline__ double zstrength_of(const ZeemanGlobal &z, int i) { return __ldg(z.strength + i); }

// Sum of all contributing lines at one (column, depth, wavelength): the body of the
// n-loop of rlk_opacity (kurucz.c:605-720) with RLKProfile (kurucz.c:729-828) inlined.
// lp_colk points at lineprep[col][0][0][k]; line nl / field f is at lp_colk[(nl*LP_NFIELD + f)*ndep].
// ARM: the table holds lines that are not polarizable (VoigtArmstrong branch); compiled out otherwise, the branch
// costs the hot kernel 2 % in registers and code even when never taken
// RLKS: keyword RLK_SCATTER -- each line's opacity is split into a thermal part (epsilon) and a scattering part that
// only enters the total opacity (kurucz.c:682-694, background.c:538-543)
template <bool ARM, bool RLKS, class ZT>
__device__ __forceinline__ void line_sums(LineSums &s, const double lambda, const int to_obs,
                                          const int first, const int count,
                                          const int *__restrict__ widx,
                                          const double *__restrict__ lines, const ZT &zee,
                                          const double *__restrict__ lp_colk, const int ndep,
                                          const double cos_gamma, const double cos_2chi,
                                          const double sin_2chi)
{
#pragma unroll
  for (int i = 0; i < 4; i++) { s.chi[i] = 0.0; s.eta[i] = 0.0; }
  s.scatt = 0.0;
  const double sign = to_obs ? 1.0 : -1.0;
  for (int j = 0; j < count; j++) {
    const int nl = __ldg(widx + first + j);
    const double *L = lines + (size_t) nl * RHB200_RL_NFIELD;
    const double *P = lp_colk + (size_t) nl * LP_NFIELD * ndep;
    const double vbroad = __ldg(P + (size_t) LP_VBROAD*ndep), sv = __ldg(P + (size_t) LP_SV*ndep);
    double v = (lambda/__ldg(L + RHB200_RL_LAMBDA0) - 1.0) * RH_CLIGHT/vbroad;
    if (to_obs) v += __ldg(P + (size_t) LP_W*ndep); else v -= __ldg(P + (size_t) LP_W*ndep);

    double phi, phi_Q = 0.0, phi_U = 0.0, phi_V = 0.0;
    const bool has_grad = (__ldg(L + RHB200_RL_GRAD) != 0.0);
    const bool polarizable = (__ldg(L + RHB200_RL_POLARIZABLE) != 0.0);
    if (!has_grad) {
      phi = ((fabs(v) <= RH_MAX_GAUSS_DOPPLER) ? rhm::rh_exp(-v*v) : 0.0) * sv;   // kurucz.c:777-778
    } else if (ARM && !polarizable) {                   // kurucz.c:823-824: Voigt(adamp, v, NULL, ARMSTRONG) * sv
      phi = rhv::voigt_armstrong(__ldg(P + (size_t) LP_ADAMP*ndep), v) * sv;
    } else {
      const double adamp = __ldg(P + (size_t) LP_ADAMP*ndep), vB = __ldg(P + (size_t) LP_VB*ndep);
      const double sin2_gamma = 1.0 - cos_gamma*cos_gamma;
      double phi_sm = 0.0, phi_pi = 0.0, phi_sp = 0.0;
      const int zoff = (int) __ldg(L + RHB200_RL_ZOFF), nc = (int) __ldg(L + RHB200_RL_NCOMP);
      for (int nz = 0; nz < nc; nz++) {
        const double H = rhv::humlicek_H(adamp, v - zshift_of(zee, zoff + nz)*vB);
        const int q = zq_of(zee, zoff + nz);
        const double st = zstrength_of(zee, zoff + nz);
        if (q == -1)     phi_sm += st * H;
        else if (q == 0) phi_pi += st * H;
        else if (q == 1) phi_sp += st * H;
      }
      const double phi_sigma = phi_sp + phi_sm;
      const double phi_delta = 0.5*phi_pi - 0.25*phi_sigma;
      phi   = (phi_delta*sin2_gamma + 0.5*phi_sigma) * sv;
      phi_Q = sign * phi_delta * sin2_gamma * cos_2chi * sv;
      phi_U = phi_delta * sin2_gamma * sin_2chi * sv;
      phi_V = sign * 0.5*(phi_sp - phi_sm) * cos_gamma * sv;
    }
    if (phi != 0.0) {                                   // kurucz.c:674
      double chi_l = __ldg(P + (size_t) LP_CHIL*ndep), eta_l = __ldg(P + (size_t) LP_ETAL*ndep);
      if (RLKS) {
        const double epsilon = __ldg(P + (size_t) LP_EPS*ndep);
        s.scatt += (1.0 - epsilon) * chi_l * phi;
        chi_l *= epsilon;
        eta_l *= epsilon;
      }
      s.chi[0] += chi_l * phi;
      s.eta[0] += eta_l * phi;
      if (polarizable && has_grad) {                    // kurucz.c:701-708
        s.chi[1] += chi_l * phi_Q;  s.chi[2] += chi_l * phi_U;  s.chi[3] += chi_l * phi_V;
        s.eta[1] += eta_l * phi_Q;  s.eta[2] += eta_l * phi_U;  s.eta[3] += eta_l * phi_V;
      }
    }
  }
}

// Thread mapping of both opacity kernels: one thread per ray-point, flattened with DEPTH
// fastest (t = ray*ndep + k).  A warp therefore holds 32 consecutive depths of one wavelength
// (occasionally the tail of one ray and the head of the next): the Doppler coordinate
// v = (lambda/lambda0 - 1) c / vbroad(k) varies slowly along the warp, so all lanes take the same
// Humlicek region except near region boundaries (ncu: 21 -> ~30 active threads / instruction
// against the wavelength-along-warp mapping), and every global access is unit stride:
// chi_ai/eta_ai[ray][k] (the reference's own layout), lineprep[col][line][field][k], atmos[col][f][k].

// FUSED: total opacity + source vector + reduced propagation matrix per ray-point, written as
// one 64-byte record {chi_I, K'_Q, K'_U, K'_V, S_I, S_Q, S_U, S_V} at raypts[ray][k].
template <int MINB, class ZT, bool ARM, bool RLKS>
__global__ void __launch_bounds__(128, MINB)
opacity_fused_kernel(int ncol, int nlambda, int ndep, int to_obs, int nline,
                     const double *__restrict__ lambda, const int *__restrict__ wfirst,
                     const int *__restrict__ wcount, const int *__restrict__ widx,
                     const double *__restrict__ lines, const __grid_constant__ ZT zee,
                     const double *__restrict__ atmos, const double *__restrict__ lineprep,
                     const double *__restrict__ chi_ai, const double *__restrict__ eta_ai,
                     double *__restrict__ raypts,
                     const double *__restrict__ mol_chi, const double *__restrict__ mol_eta,
                     const double *__restrict__ sca_ai, const int *__restrict__ wflags, int all_scalar, int mol_pol)
{
  const size_t npts = (size_t) ncol * nlambda * ndep;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;

  LineSums s;
  line_sums<ARM, RLKS>(s, __ldg(lambda + l), to_obs, __ldg(wfirst + l), __ldg(wcount + l), widx, lines,
            zee, lineprep + (size_t) col*nline*LP_NFIELD*ndep + k, ndep,
            __ldg(at + RHB200_AT_COS_GAMMA*ndep + k), __ldg(at + RHB200_AT_COS_2CHI*ndep + k),
            __ldg(at + RHB200_AT_SIN_2CHI*ndep + k));

  // background.c:476-537: chi_c = chi_ai + chi_lines (I), 0 + lines (Q,U,V);
  // formal.c:178-208: chi = 0 + chi_c, S = (0 + eta_c)/chi; stokesopac.c:72-77: K' = chi_QUV/chi_I
  // molecular lines come last in Background() (background.c:548-566): chi_c = (chi_ai + Kurucz) + molecules
  double chi = __ldg(chi_ai + t) + s.chi[0], eta = __ldg(eta_ai + t) + s.eta[0];
  if (RLKS) chi += s.scatt;                        // background.c:538-543: sca_c += scatt; chi_c += scatt
  if (mol_chi) {
    chi += __ldg(mol_chi + t); eta += __ldg(mol_eta + t);
    if (mol_pol) {                                 // polarizable molecular lines: chi_c[QUV] = (0 + Kurucz) + molecules
#pragma unroll
      for (int i = 1; i < 4; i++) { s.chi[i] += __ldg(mol_chi + i*npts + t); s.eta[i] += __ldg(mol_eta + i*npts + t); }
    }
  }
  const rhdiv::Recip rchi(chi);                  // seven IEEE quotients, one reciprocal refinement
  double2 *o = reinterpret_cast<double2 *>(raypts + t * RP_NFIELD);
  o[0] = make_double2(chi, rchi.div(s.chi[1]));
  o[1] = make_double2(rchi.div(s.chi[2]), rchi.div(s.chi[3]));
  o[2] = make_double2(rchi.div(eta), rchi.div(s.eta[1]));
  o[3] = make_double2(rchi.div(s.eta[2]), rchi.div(s.eta[3]));
  // N_MAX_SCATTER > 0: wavelengths solved for I alone keep the emissivity and the scattering opacity for the Lambda
  // iteration of S = (eta + sca J)/chi (formal.c:289-309) in the slots their (zero) S_U, S_V would take
  if (sca_ai && (all_scalar || (__ldg(wflags + l) & 2) == 0)) o[3] = make_double2(eta, RLKS ? __ldg(sca_ai + t) + s.scatt : __ldg(sca_ai + t));
}

// RAW: exactly the output of rlk_opacity(), chi/eta [ncol][nlambda][4][ndep]
__global__ void __launch_bounds__(128)
opacity_raw_kernel(int ncol, int nlambda, int ndep, int to_obs, int nline,
                   const double *__restrict__ lambda, const int *__restrict__ wfirst,
                   const int *__restrict__ wcount, const int *__restrict__ widx,
                   const double *__restrict__ lines, const int *__restrict__ zq,
                   const double *__restrict__ zshift, const double *__restrict__ zstrength,
                   const double *__restrict__ atmos, const double *__restrict__ lineprep,
                   double *__restrict__ chi, double *__restrict__ eta)
{
  const size_t npts = (size_t) ncol * nlambda * ndep;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  LineSums s;
  const ZeemanGlobal zee{zq, zshift, zstrength};
  line_sums<true, false>(s, __ldg(lambda + l), to_obs, __ldg(wfirst + l), __ldg(wcount + l), widx, lines,
            zee, lineprep + (size_t) col*nline*LP_NFIELD*ndep + k, ndep,
            __ldg(at + RHB200_AT_COS_GAMMA*ndep + k), __ldg(at + RHB200_AT_COS_2CHI*ndep + k),
            __ldg(at + RHB200_AT_SIN_2CHI*ndep + k));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    chi[(r*4 + i)*ndep + k] = s.chi[i];
    eta[(r*4 + i)*ndep + k] = s.eta[i];
  }
}

// NLTE background (background.c:519-546 with atmos.Stokes set but only the intensity record used, NO_STOKES): the I
// component of rlk_opacity() is added to chi_c / eta_c [ncol][nlambda][ndep] in place, where a line is in the window
// (hasline: the reference adds nothing otherwise).  Same line_sums as the raw kernel, hence the same roundings.
__global__ void __launch_bounds__(128)
opacity_addI_kernel(int ncol, int nlambda, int ndep, int to_obs, int nline,
                    const double *__restrict__ lambda, const int *__restrict__ wfirst,
                    const int *__restrict__ wcount, const int *__restrict__ widx,
                    const double *__restrict__ lines, const int *__restrict__ zq,
                    const double *__restrict__ zshift, const double *__restrict__ zstrength,
                    const double *__restrict__ atmos, const double *__restrict__ lineprep,
                    double *__restrict__ chi_c, double *__restrict__ eta_c,
                    double *__restrict__ chi_quv, double *__restrict__ eta_quv /* [ncol][nlambda][3][ndep] or NULL */)
{
  const size_t npts = (size_t) ncol * nlambda * ndep;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  const int count = __ldg(wcount + l);
  if (count == 0) return;
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  LineSums s;
  const ZeemanGlobal zee{zq, zshift, zstrength};
  line_sums<true, false>(s, __ldg(lambda + l), to_obs, __ldg(wfirst + l), count, widx, lines,
            zee, lineprep + (size_t) col*nline*LP_NFIELD*ndep + k, ndep,
            __ldg(at + RHB200_AT_COS_GAMMA*ndep + k), __ldg(at + RHB200_AT_COS_2CHI*ndep + k),
            __ldg(at + RHB200_AT_SIN_2CHI*ndep + k));
  chi_c[t] += s.chi[0];
  eta_c[t] += s.eta[0];
  if (chi_quv) {                                   // the Q, U, V records of a polarised background (readj.c:303-337)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      chi_quv[(r*3 + i)*ndep + k] = s.chi[i + 1];
      eta_quv[(r*3 + i)*ndep + k] = s.eta[i + 1];
    }
  }
}

// chi_c += a, eta_c += b where the wavelength has entries in the window list (molecular lines of the NLTE background)
__global__ void __launch_bounds__(128)
add_where_kernel(size_t n, int nlambda, int ndep, const int *__restrict__ wcount, const double *__restrict__ a,
                 const double *__restrict__ b, double *__restrict__ chi_c, double *__restrict__ eta_c)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int l = (int) ((t / ndep) % nlambda);
  if (__ldg(wcount + l) == 0) return;
  chi_c[t] += a[t];
  eta_c[t] += b[t];
}

// d chi_c / d log gf and d eta_c / d log gf of the wavelengths solved with the scalar ray (kurucz.c:696-699:
// spectrum.dchi_c_lam[nspect][k][p] = chi_l * phi * LN10 of the line parameter p belongs to, in the up direction -- the
// last rlk_opacity() call before the up-ray of Formal() reads it).  One thread per (column, I-only wavelength, depth).
// The reference leaves the entry uninitialised (matrix3d_double mallocs) where the line is outside the wavelength's
// window; it is 0 here.  scratch [ncol][nunpol][fields][ndep]: dchi [ndep][npar] at field 3, deta at field 3 + npar.
template <bool ARM, bool RLKS>
__global__ void __launch_bounds__(128)
loggf_dopac_kernel(int ncol, int nlambda, int ndep, int nline, const int *__restrict__ nolines, int nnoline,
                   const int *__restrict__ unpol_rank, int nunpol, int fields, int npar, const int *__restrict__ par_line,
                   const double *__restrict__ lambda, const int *__restrict__ wfirst, const int *__restrict__ wcount,
                   const int *__restrict__ widx, const double *__restrict__ lines, const int *__restrict__ zq,
                   const double *__restrict__ zshift, const double *__restrict__ zstrength,
                   const double *__restrict__ atmos, const double *__restrict__ lineprep, double *__restrict__ scratch)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nnoline * ndep) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nnoline), l = __ldg(nolines + (int) (r - (size_t) col * nnoline));
  const int rank = __ldg(unpol_rank + l);
  if (rank < 0) return;
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const ZeemanGlobal zee{zq, zshift, zstrength};
  double *dchi = scratch + ((size_t) col * nunpol + rank) * fields * ndep + (size_t) 3 * ndep, *deta = dchi + (size_t) npar * ndep;
  const int first = __ldg(wfirst + l), count = __ldg(wcount + l);
  for (int p = 0; p < npar; p++) {
    const int nl = __ldg(par_line + p);
    double dc = 0.0, de = 0.0;
    for (int j = 0; j < count; j++) {
      if (__ldg(widx + first + j) != nl) continue;
      LineSums s;
      line_sums<ARM, RLKS>(s, __ldg(lambda + l), 1, first + j, 1, widx, lines, zee,
                           lineprep + (size_t) col*nline*LP_NFIELD*ndep + k, ndep,
                           __ldg(at + RHB200_AT_COS_GAMMA*ndep + k), __ldg(at + RHB200_AT_COS_2CHI*ndep + k),
                           __ldg(at + RHB200_AT_SIN_2CHI*ndep + k));
      dc = s.chi[0] * 2.30258509299404568402;          // LN10 = log(10), kurucz.c:94
      de = s.eta[0] * 2.30258509299404568402;
      break;
    }
    dchi[(size_t) k * npar + p] = dc;
    deta[(size_t) k * npar + p] = de;
  }
}

// MolecularOpacity + MolProfile (opacity.c:711-916): LTE lines of PASSIVE molecules in the background.
// One thread per (column, wavelength, depth); widx lists, per wavelength, the lines that pass the
// reference's window tests (opacity.c:774-787, evaluated on the host) in molecule / line order, so the
// accumulation order is the reference's.  mol: [ncol][nmol][3][ndep] = molecule->n, pf, vbroad.
__global__ void __launch_bounds__(128)
mol_opacity_raw_kernel(int intensity_only, int ncol, int nlambda, int ndep, int nmol, double muz, int moving, int to_obs,
                       const double *__restrict__ lambda, const int *__restrict__ wfirst,
                       const int *__restrict__ wcount, const int *__restrict__ widx,
                       const double *__restrict__ mlines, const int *__restrict__ zq,
                       const double *__restrict__ zshift, const double *__restrict__ zstrength,
                       const double *__restrict__ atmos, const double *__restrict__ mol,
                       double *__restrict__ chi, double *__restrict__ eta)
{
  const size_t npts = (size_t) ncol * nlambda * ndep;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double T = at[RHB200_AT_T*ndep + k], vel = at[RHB200_AT_VEL*ndep + k], B = at[RHB200_AT_B*ndep + k];
  const double cos_gamma = at[RHB200_AT_COS_GAMMA*ndep + k], cos_2chi = at[RHB200_AT_COS_2CHI*ndep + k],
               sin_2chi = at[RHB200_AT_SIN_2CHI*ndep + k];
  const double lam = __ldg(lambda + l);
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  double c4[4] = {0.0, 0.0, 0.0, 0.0}, e4[4] = {0.0, 0.0, 0.0, 0.0};
  const int first = __ldg(wfirst + l), count = __ldg(wcount + l);
  for (int j = 0; j < count; j++) {
    const double *L = mlines + (size_t) __ldg(widx + first + j) * RHB200_ML_NFIELD;
    const double *M = mol + (((size_t) col * nmol + (int) L[RHB200_ML_MOL]) * 3) * ndep + k;
    const double n = M[0], pf = M[ndep], vbroad = M[2*(size_t) ndep];
    if (!(n > 0.0)) continue;                                            // opacity.c:797
    const double lambda0 = L[RHB200_ML_LAMBDA0];
    // MolProfile, opacity.c:844-916
    const double adamp = L[RHB200_ML_AJI] * (lambda0 * RH_NM_TO_M) / (4.0*RH_PI * vbroad);
    double v = (lam/lambda0 - 1.0) * RH_CLIGHT/vbroad;
    if (moving) { if (to_obs) v += (muz * vel) / vbroad; else v -= (muz * vel) / vbroad; }
    const double sv = 1.0 / (RH_SQRTPI * vbroad);
    double phi, phi_Q = 0.0, phi_U = 0.0, phi_V = 0.0;
    const bool polarizable = L[RHB200_ML_POLARIZABLE] != 0.0;
    if (polarizable) {
      const double sin2_gamma = 1.0 - cos_gamma*cos_gamma;
      const double vB = (RH_LARMOR * lambda0) * B / vbroad;
      const double sign = to_obs ? 1.0 : -1.0;
      double phi_sm = 0.0, phi_pi = 0.0, phi_sp = 0.0;
      const int zoff = (int) L[RHB200_ML_ZOFF], nc = (int) L[RHB200_ML_NCOMP];
      for (int nz = 0; nz < nc; nz++) {
        const double H = rhv::humlicek_H(adamp, v - __ldg(zshift + zoff + nz)*vB);
        const int q = __ldg(zq + zoff + nz);
        const double st = __ldg(zstrength + zoff + nz);
        if (q == -1)     phi_sm += st * H;
        else if (q == 0) phi_pi += st * H;
        else if (q == 1) phi_sp += st * H;
      }
      const double phi_sigma = phi_sp + phi_sm;
      const double phi_delta = 0.5*phi_pi - 0.25*phi_sigma;
      phi   = (phi_delta*sin2_gamma + 0.5*phi_sigma) * sv;
      phi_Q = sign * phi_delta * sin2_gamma * cos_2chi * sv;
      phi_U = phi_delta * sin2_gamma * sin_2chi * sv;
      phi_V = sign * 0.5*(phi_sp - phi_sm) * cos_gamma * sv;
    } else
      phi = rhv::voigt_armstrong(adamp, v) * sv;
    // MolecularOpacity, opacity.c:786-788, 802-822
    const double hc_la     = (RH_HPLANCK * RH_CLIGHT) / (lambda0 * RH_NM_TO_M);
    const double Bijhc_4PI = hc_4PI * L[RHB200_ML_BIJ] * L[RHB200_ML_ISO_FRAC] * L[RHB200_ML_GI];
    const double twohnu3_c2 = L[RHB200_ML_AJI] / L[RHB200_ML_BJI];
    const double kT    = 1.0 / (RH_KBOLTZMANN * T);
    const double ni_gi = n * rhm::rh_exp(-L[RHB200_ML_EI] * kT) / pf;
    const double nj_gj = ni_gi * rhm::rh_exp(-hc_la * kT);
    const double chi_l = Bijhc_4PI * (ni_gi - nj_gj);
    const double eta_l = Bijhc_4PI * twohnu3_c2 * nj_gj;
    c4[0] += chi_l * phi;
    e4[0] += eta_l * phi;
    if (polarizable) {
      c4[1] += chi_l * phi_Q;  c4[2] += chi_l * phi_U;  c4[3] += chi_l * phi_V;
      e4[1] += eta_l * phi_Q;  e4[2] += eta_l * phi_U;  e4[3] += eta_l * phi_V;
    }
  }
  if (intensity_only) {          // fused path: [ncol][nlambda][ndep], zero where there is no line; 2 = I, Q, U, V planes
    chi[t] = c4[0];
    eta[t] = e4[0];
    if (intensity_only == 2) {
#pragma unroll
      for (int i = 1; i < 4; i++) { chi[i*npts + t] = c4[i]; eta[i*npts + t] = e4[i]; }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    chi[(r*4 + i)*ndep + k] = c4[i];
    eta[(r*4 + i)*ndep + k] = e4[i];
  }
}

// molecule->n (from the chemistry kernel), molecule->pf = partfunction(T) (chemequil.c:395-443) and the Doppler width
// (readmolecule.c:233-237) of the molecules that have line lists: mol [ncol][nsel][3][ndep].
// msel [nsel][16] = chem index, weight, fit, Tmin, Tmax, Npf, pf_coef[8]
__global__ void __launch_bounds__(128)
mol_prep_kernel(int ncol, int ndep, int nsel, const double *__restrict__ msel, const double *__restrict__ atmos,
                const double *__restrict__ molden, double *__restrict__ mol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nsel * ndep) return;
  const int k = (int) (t % ndep), q = (int) ((t / ndep) % nsel), col = (int) (t / ((size_t) ndep * nsel));
  const double *M = msel + (size_t) q * 16;
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double T = at[RHB200_AT_T*ndep + k], vturb = at[RHB200_AT_VTURB*ndep + k];
  double pf = 0.0;
  if (!(T < M[3] || T > M[4])) {
    const int fit = (int) M[2], npf = (int) M[5];
    const double *c = M + 6;
    if (fit == 0) {            // KURUCZ_70
      pf = c[0]; for (int i = 1; i < npf; i++) pf = pf*T + c[i];
      pf = rhm::rh_exp(pf);
    } else if (fit == 1) {     // KURUCZ_85
      const double x = T * 1.0E-4;
      pf = c[0]; for (int i = 1; i < npf; i++) pf = pf*x + c[i];
      pf = rhm::rh_exp(pf);
    } else if (fit == 2) {     // SAUVAL_TATUM_84
      const double x = rhm::rh_log10(5.03974756E+03 / T);
      pf = c[0]; for (int i = 1; i < npf; i++) pf = pf*x + c[i];
      pf = rhm::rh_exp(2.30258509299404568402 * (pf));
    } else if (fit == 3) {     // IRWIN_81
      const double x = rhm::rh_log(T);
      pf = c[0]; for (int i = 1; i < npf; i++) pf = pf*x + c[i];
      pf = rhm::rh_exp(pf);
    }
  }
  const double vtherm = 2.0*RH_KBOLTZMANN / (RH_AMU * M[1]);
  double *o = mol + (((size_t) col * nsel + q) * 3) * ndep + k;
  o[0] = molden[((size_t) col * nsel + q) * ndep + k];
  o[ndep] = pf;
  o[2*(size_t) ndep] = sqrt(vtherm*T + vturb*vturb);
}

// passive_bb (metal.c:174-344): bound-bound lines of PASSIVE model atoms, unpolarised.  One thread per
// (column, wavelength, depth); lines in the reference's order (atoms, then lines), components inside.
// pcol: [ncol][nline][4][ndep] = n_i, n_j, vbroad, Damping() output.
__global__ void __launch_bounds__(128)
passive_bb_kernel(int accumulate, int ncol, int nlambda, int ndep, int nline, double muz, int moving, int to_obs,
                  const double *__restrict__ lambda, const int *__restrict__ wfirst,
                  const int *__restrict__ wcount, const int *__restrict__ widx,
                  const double *__restrict__ plines, const double *__restrict__ c_shift,
                  const double *__restrict__ c_fraction, const double *__restrict__ atmos,
                  const double *__restrict__ pcol, double *__restrict__ chi, double *__restrict__ eta)
{
  const size_t npts = (size_t) ncol * nlambda * ndep;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  const double vel = atmos[((size_t) col * RHB200_AT_NFIELD + RHB200_AT_VEL) * ndep + k];
  const double lam = __ldg(lambda + l);
  const double hc_4PI = (RH_HPLANCK * RH_CLIGHT) / (4.0 * RH_PI);
  double c = 0.0, e = 0.0;
  const int first = __ldg(wfirst + l), count = __ldg(wcount + l);
  for (int jn = 0; jn < count; jn++) {
    const int n = __ldg(widx + first + jn);
    const double *L = plines + (size_t) n * RHB200_PB_NFIELD;
    const double *P = pcol + (((size_t) col * nline + n) * 4) * ndep + k;
    const double n_i = P[0], n_j = P[ndep], vbroad = P[2*(size_t) ndep], adamp = P[3*(size_t) ndep];
    const double lambda0 = L[RHB200_PB_LAMBDA0], Bij = L[RHB200_PB_BIJ];
    const double gij = L[RHB200_PB_BJI] / Bij, twohnu3_c2 = L[RHB200_PB_AJI] / L[RHB200_PB_BJI];
    const bool voigt = L[RHB200_PB_VOIGT] != 0.0;
    const int ncomp = (int) L[RHB200_PB_NCOMP], off = (int) L[RHB200_PB_COMPOFF];
    for (int nc = 0; nc < ncomp; nc++) {
      double v = (lam - lambda0 - __ldg(c_shift + off + nc)) * RH_CLIGHT / (lambda0 * vbroad);
      if (moving) { if (to_obs) v += (muz * vel) / vbroad; else v -= (muz * vel) / vbroad; }
      const double phi = voigt ? rhv::voigt_armstrong(adamp, v) * __ldg(c_fraction + off + nc)
                               : rhm::rh_exp(-(v*v));
      const double Vij = hc_4PI * Bij * phi / (RH_SQRTPI*vbroad);
      c += Vij * (n_i - gij * n_j);
      e += twohnu3_c2 * gij * Vij * n_j;
    }
  }
  if (accumulate) {                 // fused path: chi_c = chi_ai + passive lines, only where there is one (background.c:500-515)
    if (count > 0) { chi[t] += c; eta[t] += e; }
  } else {
    chi[t] = c;
    eta[t] = e;
  }
}

// Per (column, passive line, depth): populations of the two levels, the atom's Doppler width (readatom.c:199-201) and
// Damping() (broad.c:273-314: van der Waals + quadratic Stark + hydrogen's linear Stark, in that order) -> pcol.
__global__ void __launch_bounds__(128)
passive_prep_kernel(int ncol, int ndep, int npl, int nlev, const double *__restrict__ plrows,
                    const double *__restrict__ atmos, const double *__restrict__ pops, double *__restrict__ pcol,
                    double *__restrict__ qelast_out /* line->Qelast [ncol][npl][ndep] (broad.c:305-308) or NULL */)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * npl * ndep) return;
  const int k = (int) (t % ndep), n = (int) ((t / ndep) % npl), col = (int) (t / ((size_t) ndep * npl));
  const double *L = plrows + (size_t) n * RHB200_PL_NFIELD;
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double *P = pops + (size_t) col * nlev * ndep + k;
  const double T = at[RHB200_AT_T*ndep + k], ne = at[RHB200_AT_NE*ndep + k], vturb = at[RHB200_AT_VTURB*ndep + k];
  const double vtherm = 2.0*RH_KBOLTZMANN/(RH_AMU * L[RHB200_PL_WEIGHT]);
  const double vbroad = sqrt(vtherm*T + vturb*vturb);
  double adamp = 0.0, Qelast = 0.0;
  if (L[RHB200_PL_VOIGT] != 0.0) {
    const int vdw = (int) L[RHB200_PL_VDW_TYPE];
    if (vdw >= 0) {                                                            // VanderWaals, broad.c:60-140
      double GvdW;
      if (vdw == 0) GvdW = L[RHB200_PL_VDW_A] * rhm::rh_pow(T, 0.3);
      else if (vdw == 2) GvdW = L[RHB200_PL_VDW_A] * rhm::rh_pow(T, L[RHB200_PL_VDW_B]) +      // BARKLEM, broad.c:125-136
                                L[RHB200_PL_VDW_C] * rhm::rh_pow(T, 0.3);
      else GvdW = L[RHB200_PL_VDW_A] * rhm::rh_pow(T, L[RHB200_PL_VDW_B]) +
                  L[RHB200_PL_VDW_C] * rhm::rh_pow(T, L[RHB200_PL_VDW_D]) * L[RHB200_PL_HE_ABUND];
      GvdW *= P[0];                                                            // atmos.H->n[0][k]
      Qelast += GvdW;
    }
    const int stark = (int) L[RHB200_PL_STARK_TYPE];
    if (stark == 1) Qelast += L[RHB200_PL_STARK_A] * ne;                       // Stark, broad.c:147-215
    else if (stark == 2) {
      const double vrel = rhm::rh_pow(L[RHB200_PL_STARK_C] * T, 0.16666667) * L[RHB200_PL_STARK_CM];
      Qelast += L[RHB200_PL_STARK_A] * vrel * ne;
    }
    if (L[RHB200_PL_IS_H] != 0.0) Qelast += L[RHB200_PL_LINSTARK_C] * rhm::rh_pow(ne, 0.66666667);   // StarkLinear
    const double cDop = (RH_NM_TO_M * L[RHB200_PL_LAMBDA0]) / (4.0 * RH_PI);
    adamp = (L[RHB200_PL_GRAD] + Qelast) * cDop / vbroad;
  }
  double *o = pcol + (((size_t) col * npl + n) * 4) * ndep + k;
  o[0] = P[(size_t) ((int) L[RHB200_PL_LEVEL_I]) * ndep];
  o[ndep] = P[(size_t) ((int) L[RHB200_PL_LEVEL_J]) * ndep];
  o[2*(size_t) ndep] = vbroad;
  o[3*(size_t) ndep] = adamp;
  if (qelast_out) qelast_out[t] = Qelast;
}

__global__ void voigt_kernel(int n, const double *__restrict__ a, const double *__restrict__ v,
                             double *__restrict__ H, double *__restrict__ F, int *__restrict__ region)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double f;
  H[i] = rhv::humlicek(a[i], v[i], &f);
  F[i] = f;
  if (region) region[i] = rhv::humlicek_region(a[i], v[i]);
}

__global__ void voigt_armstrong_kernel(int n, const double *__restrict__ a, const double *__restrict__ v,
                                       double *__restrict__ H, int *__restrict__ region)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  H[i] = rhv::voigt_armstrong(a[i], v[i]);
  if (region) region[i] = rhv::armstrong_region(a[i], v[i]);
}

__global__ void math_probe_kernel(int n, int func, const double *__restrict__ x,
                                  const double *__restrict__ y, double *__restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  switch (func) {
  case 0: r = rhm::rh_exp(x[i]); break;
  case 1: r = rhm::rh_sin(x[i]); break;
  case 2: r = rhm::rh_cos(x[i]); break;
  case 4: { rhdiv::Recip rc(y[i]); r = rc.div(x[i]); break; }     // shared-reciprocal division
  case 5: r = x[i] / y[i]; break;                                  // compiler's IEEE division
  case 6: r = rhm::rh_atan(x[i]); break;
  case 7: r = rhm::rh_log(x[i]); break;
  case 8: r = rhm::rh_log10(x[i]); break;
  default: r = rhm::rh_pow(x[i], y[i]); break;
  }
  out[i] = r;
}

}  // namespace

int rh_launch_mol_opacity_raw(rhb200_ctx *ctx, int ncol, int nlambda, int ndep, int nmol, double muz, int moving,
                              int to_obs, const double *d_lambda, const int *d_first, const int *d_count,
                              const int *d_idx, const double *d_mlines, const int *d_zq, const double *d_zshift,
                              const double *d_zstrength, const double *d_atmos, const double *d_mol,
                              double *d_chi, double *d_eta)
{
  const size_t n = (size_t) ncol * nlambda * ndep;
  if (n == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    mol_opacity_raw_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(0, ncol, nlambda, ndep, nmol, muz,
        moving, to_obs, d_lambda, d_first, d_count, d_idx, d_mlines, d_zq, d_zshift, d_zstrength, d_atmos, d_mol,
        d_chi, d_eta);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

// MolecularOpacity of one chunk in the fused path: densities / partition functions / Doppler
// widths of the molecules with line lists, then chi and eta of their lines per ray-point, which the Kurucz-line
// kernel adds last (background.c:548-566)
int rh_molecular_chunk(rhb200_ctx *ctx, int ncol, int ndep, double muz, const double *d_atmos, const double *d_molden,
                       double *d_mol, double *d_molchi, double *d_moleta)
{
  const DevWave &w = ctx->wav;
  if (w.nmw == 0 || ncol == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_PREP);
    const size_t n = (size_t) ncol * w.nmsel * ndep;
    mol_prep_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(ncol, ndep, w.nmsel, w.ml_sel, d_atmos, d_molden, d_mol);
  }
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    const size_t n = (size_t) ncol * w.nlambda * ndep;
    mol_opacity_raw_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(w.mol_pol ? 2 : 1, ncol, w.nlambda, ndep, w.nmsel, muz, 1, 1,
        w.lambda, w.mw_first, w.mw_count, w.mw_idx, w.ml_rows, w.mz_q, w.mz_shift, w.mz_strength, d_atmos, d_mol, d_molchi, d_moleta);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_passive_bb(rhb200_ctx *ctx, int ncol, int nlambda, int ndep, int nline, double muz, int moving,
                         int to_obs, const double *d_lambda, const int *d_first, const int *d_count,
                         const int *d_idx, const double *d_plines, const double *d_cshift, const double *d_cfrac,
                         const double *d_atmos, const double *d_pcol, double *d_chi, double *d_eta)
{
  const size_t n = (size_t) ncol * nlambda * ndep;
  if (n == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    passive_bb_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(0, ncol, nlambda, ndep, nline, muz, moving,
        to_obs, d_lambda, d_first, d_count, d_idx, d_plines, d_cshift, d_cfrac, d_atmos, d_pcol, d_chi, d_eta);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

// passive_bb of one chunk in the fused path: damping parameters and populations, then the lines are added to the
// background the Kurucz-line kernel starts from
int rh_passive_chunk(rhb200_ctx *ctx, int ncol, int ndep, double muz, const double *d_atmos, const double *d_pops, int nlev,
                     double *d_pcol, double *d_chi_ai, double *d_eta_ai)
{
  const DevWave &w = ctx->wav;
  if (w.npl == 0 || ncol == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_PREP);
    const size_t n = (size_t) ncol * w.npl * ndep;
    passive_prep_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(ncol, ndep, w.npl, nlev, w.pl_rows, d_atmos, d_pops, d_pcol, nullptr);
  }
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    const size_t n = (size_t) ncol * w.nlambda * ndep;
    passive_bb_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(1, ncol, w.nlambda, ndep, w.npl, muz, 1, 1,
        w.lambda, w.pw_first, w.pw_count, w.pw_idx, w.pl_pb, w.pl_cshift, w.pl_cfrac, d_atmos, d_pcol, d_chi_ai, d_eta_ai);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_prep(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                   const double *d_atmos, double *d_elem_n, double *d_lineprep)
{
  const size_t n = (size_t) ncol * ndep;
  if (n == 0) return RHB200_OK;
  const int threads = 128;
  const unsigned blocks = (unsigned) ((n + threads - 1) / threads);
  {
    ScopedKernelTimer t(ctx, RHB200_K_PREP);
    prep_kernel<<<blocks, threads, 0, ctx->stream>>>(ncol, ndep, muz, moving, ctx->tab.rlkscatter, ctx->tab.nline,
        ctx->tab.nelem, ctx->tab.npf, ctx->tab.lines, ctx->tab.elems, ctx->tab.pf, ctx->tab.Tpf,
        d_atmos, d_elem_n, d_lineprep);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_opacity_fused(rhb200_ctx *ctx, int ncol, int ndep, int to_obs,
                            const double *d_atmos, const double *d_lineprep,
                            const double *d_chi_ai, const double *d_eta_ai, double *d_raypts,
                            const double *d_molchi, const double *d_moleta, const double *d_sca)
{
  const size_t nray = (size_t) ncol * ctx->wav.nlambda;
  if (nray == 0) return RHB200_OK;
  const size_t npts = nray * (size_t) ndep;
  const unsigned block = 128, grid = (unsigned) ((npts + block - 1) / block);
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    static int variant = -1;
    if (variant < 0) { const char *e = getenv("RHB200_OPACITY_MINB"); variant = e ? atoi(e) : 10; }      // measured (ms per 4096 columns): 4 -> 11.7, 5 -> 10.1, 6 -> 9.2, 8 -> 8.84, 10 -> 8.70, 12 -> 8.91; 9 is 0.06 ms per 2048-column e2e call behind 10
    bool arm = false;                                     // any line with damping wings that is not polarizable
    for (int n = 0; n < ctx->tab.nline; n++) {
      const double *L = ctx->h_lines.data() + (size_t) n * RHB200_RL_NFIELD;
      arm = arm || (L[RHB200_RL_GRAD] != 0.0 && L[RHB200_RL_POLARIZABLE] == 0.0);
    }
#define RH_OPF_ARGS(Z) (ncol, ctx->wav.nlambda, ndep, to_obs, ctx->tab.nline, ctx->wav.lambda, ctx->wav.first, \
        ctx->wav.count, ctx->wav.idx, ctx->tab.lines, Z, d_atmos, d_lineprep, d_chi_ai, d_eta_ai, d_raypts, d_molchi, d_moleta, d_sca, ctx->wav.flags, ctx->no_stokes, ctx->wav.mol_pol)
#define RH_LAUNCH_OPF(M, ZT, Z) do { if (ctx->tab.rlkscatter) opacity_fused_kernel<M, ZT, true, true><<<grid, block, 0, ctx->stream>>>RH_OPF_ARGS(Z); \
        else if (arm) opacity_fused_kernel<M, ZT, true, false><<<grid, block, 0, ctx->stream>>>RH_OPF_ARGS(Z); \
        else opacity_fused_kernel<M, ZT, false, false><<<grid, block, 0, ctx->stream>>>RH_OPF_ARGS(Z); } while (0)
#define RH_LAUNCH_OPF_V(ZT, Z) switch (variant) {             \
    case 4: RH_LAUNCH_OPF(4, ZT, Z); break;                   \
    case 5: RH_LAUNCH_OPF(5, ZT, Z); break;                   \
    case 6: RH_LAUNCH_OPF(6, ZT, Z); break;                   \
    case 9: RH_LAUNCH_OPF(9, ZT, Z); break;                   \
    case 10: RH_LAUNCH_OPF(10, ZT, Z); break;                 \
    case 12: RH_LAUNCH_OPF(12, ZT, Z); break;                 \
    default: RH_LAUNCH_OPF(8, ZT, Z); break; }
    if (ctx->tab.ncomp <= RH_ZPARAM_MAX && !getenv("RHB200_ZEEMAN_GLOBAL")) {
      ZeemanParam zp;
      memset(&zp, 0, sizeof(zp));
      for (int i = 0; i < ctx->tab.ncomp; i++) {
        zp.shift[i] = ctx->h_zshift[i]; zp.strength[i] = ctx->h_zstrength[i]; zp.q[i] = (signed char) ctx->h_zq[i];
      }
      RH_LAUNCH_OPF_V(ZeemanParam, zp)
    } else {
      const ZeemanGlobal zg{ctx->tab.zq, ctx->tab.zshift, ctx->tab.zstrength};
      RH_LAUNCH_OPF_V(ZeemanGlobal, zg)
    }
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_loggf_dopac(rhb200_ctx *ctx, int ncol, int ndep, const double *d_atmos, const double *d_lineprep, double *d_scratch)
{
  const int nn = ctx->wav.nnoline;
  if (nn == 0 || ncol == 0 || ctx->lrf_npar == 0) return RHB200_OK;
  const size_t n = (size_t) ncol * nn * ndep;
  const unsigned grid = (unsigned) ((n + 127) / 128);
  bool arm = false;
  for (int i = 0; i < ctx->tab.nline; i++) {
    const double *L = ctx->h_lines.data() + (size_t) i * RHB200_RL_NFIELD;
    arm = arm || (L[RHB200_RL_GRAD] != 0.0 && L[RHB200_RL_POLARIZABLE] == 0.0);
  }
#define RH_DOPAC_ARGS (ncol, ctx->wav.nlambda, ndep, ctx->tab.nline, ctx->wav.noline, nn, ctx->wav.unpol_rank, ctx->wav.nunpol, \
      ctx->scal_fields(), ctx->lrf_npar, ctx->d_lrf_lines, ctx->wav.lambda, ctx->wav.first, ctx->wav.count, ctx->wav.idx, ctx->tab.lines, \
      ctx->tab.zq, ctx->tab.zshift, ctx->tab.zstrength, d_atmos, d_lineprep, d_scratch)
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    if (ctx->tab.rlkscatter) loggf_dopac_kernel<true, true><<<grid, 128, 0, ctx->stream>>>RH_DOPAC_ARGS;
    else if (arm) loggf_dopac_kernel<true, false><<<grid, 128, 0, ctx->stream>>>RH_DOPAC_ARGS;
    else loggf_dopac_kernel<false, false><<<grid, 128, 0, ctx->stream>>>RH_DOPAC_ARGS;
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_opacity_addI(rhb200_ctx *ctx, int ncol, int ndep, int to_obs, const double *d_atmos, const double *d_lineprep,
                           double *d_chi_c, double *d_eta_c, double *d_chi_quv, double *d_eta_quv)
{
  const size_t npts = (size_t) ncol * ctx->wav.nlambda * ndep;
  if (npts == 0 || ctx->tab.nline == 0) return RHB200_OK;
  if (ctx->tab.rlkscatter) { rhb200_set_error("RLK_SCATTER in the NLTE background is not implemented"); return RHB200_EUNSUPPORTED; }
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    opacity_addI_kernel<<<(unsigned) ((npts + 127) / 128), 128, 0, ctx->stream>>>(ncol, ctx->wav.nlambda, ndep, to_obs,
        ctx->tab.nline, ctx->wav.lambda, ctx->wav.first, ctx->wav.count, ctx->wav.idx,
        ctx->tab.lines, ctx->tab.zq, ctx->tab.zshift, ctx->tab.zstrength, d_atmos, d_lineprep, d_chi_c, d_eta_c,
        d_chi_quv, d_eta_quv);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_add_molecular(rhb200_ctx *ctx, int ncol, int ndep, const double *d_molchi, const double *d_moleta,
                            double *d_chi_c, double *d_eta_c)
{
  const size_t n = (size_t) ncol * ctx->wav.nlambda * ndep;
  if (n == 0 || ctx->wav.nmw == 0) return RHB200_OK;
  add_where_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(n, ctx->wav.nlambda, ndep, ctx->wav.mw_count,
                                                                          d_molchi, d_moleta, d_chi_c, d_eta_c);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

// Damping() + Doppler width of an arbitrary line table (rows RHB200_PL_*): the ACTIVE lines of the NLTE front end.
// d_pcol [ncol][nline][4][ndep] = n_i, n_j, vbroad, adamp
int rh_launch_line_damping(rhb200_ctx *ctx, int ncol, int ndep, int nline, const double *d_plrows, const double *d_atmos,
                           const double *d_pops, int nlev, double *d_pcol, double *d_qelast)
{
  const size_t n = (size_t) ncol * nline * ndep;
  if (n == 0) return RHB200_OK;
  ScopedKernelTimer t(ctx, RHB200_K_PREP);
  passive_prep_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(ncol, ndep, nline, nlev, d_plrows, d_atmos, d_pops, d_pcol, d_qelast);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_opacity_raw(rhb200_ctx *ctx, int ncol, int ndep, int to_obs,
                          const double *d_atmos, const double *d_lineprep,
                          double *d_chi, double *d_eta)
{
  const size_t nray = (size_t) ncol * ctx->wav.nlambda;
  if (nray == 0) return RHB200_OK;
  const size_t npts = nray * (size_t) ndep;
  const unsigned block = 128, grid = (unsigned) ((npts + block - 1) / block);
  {
    ScopedKernelTimer t(ctx, RHB200_K_OPACITY);
    opacity_raw_kernel<<<grid, block, 0, ctx->stream>>>(ncol, ctx->wav.nlambda, ndep, to_obs,
        ctx->tab.nline, ctx->wav.lambda, ctx->wav.first, ctx->wav.count, ctx->wav.idx,
        ctx->tab.lines, ctx->tab.zq, ctx->tab.zshift, ctx->tab.zstrength, d_atmos, d_lineprep,
        d_chi, d_eta);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_voigt(rhb200_ctx *ctx, int n, const double *d_a, const double *d_v,
                    double *d_H, double *d_F, int *d_region)
{
  if (n == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_OTHER);
    voigt_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, d_a, d_v, d_H, d_F, d_region);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_voigt_armstrong(rhb200_ctx *ctx, int n, const double *d_a, const double *d_v,
                              double *d_H, int *d_region)
{
  if (n == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_OTHER);
    voigt_armstrong_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, d_a, d_v, d_H, d_region);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_math_probe(rhb200_ctx *ctx, int n, int func, const double *d_x, const double *d_y,
                         double *d_out)
{
  if (n == 0) return RHB200_OK;
  math_probe_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, func, d_x, d_y, d_out);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}
