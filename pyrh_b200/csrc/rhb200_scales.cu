// rhb200_scales.cu -- the pyrh boundary of one column on the device: what rhf1d() does to the nine
// atmosphere rows pyrh.compute1d hands it before Background() and between Background() and Iterate().
//
//   pyrh_rows_kernel   unit conversion in place of pyrh_compute1dray.c:230-270 (the reference's own
//                      expressions: *= KM_TO_M, /= CUBE(CM_TO_M), /= 1e4; POW10 = exp(LG10*x), rh.h:50)
//                      and Bproject(), rhf1d/project.c:38-80, both branches (mu == 1 and inclined rays)
//   static_cols_kernel atmos.moving per column (pyrh_compute1dray.c:272-278)
//   proton_kernel      np = atmos.H->n[Nlevel-1] after ChemicalEquilibrium (kurucz.c:772)
//   scales_kernel      convertScales(), rhf1d/multiatmos.c:100-177: tau_500 / column mass -> height with the
//                      reference-wavelength opacity, then the shift that puts tau_ref = 1 at height 0
//                      (Linear(), linear.c:22-51 with Locate(), hunt.c:92-117)
#include "rhb200_common.cuh"
#include "rhb200_math.cuh"

#define RH_KM_TO_M  1.0E+03
#define RH_CM_TO_M  1.0E-02
#define RH_G_TO_KG  1.0E-03
#define RH_LG10     2.30258509299404568402
#define SQ(x)   ((x)*(x))
#define CUBE(x) ((x)*(x)*(x))

// one thread per (column, depth).  in: [ncol][9][ndep] pyrh rows (pyrh.pyx:621-625); out: [ncol][RHB200_AT_NFIELD][ndep].
// The height row receives the scale the column came with, already in SI: tau_ref (TAU500), cmass (COLUMN_MASS)
// or height (GEOMETRIC); scales_kernel turns the first two into heights.
__global__ void __launch_bounds__(128)
pyrh_rows_kernel(int ncol, int ndep, int nrow_in, int atm_scale, double muz,
                 const double *__restrict__ in, double *__restrict__ atmos)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t % ndep);
  const double *a = in + (size_t) col * nrow_in * ndep + k;
  double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep + k;
  const double scale = a[0];
  double s;
  if (atm_scale == 0)      s = rhm::rh_exp(RH_LG10 * (scale));                                      // :232-233
  else if (atm_scale == 1) s = rhm::rh_exp(RH_LG10 * (scale)) * (RH_G_TO_KG / SQ(RH_CM_TO_M));      // :237-238
  else                     s = scale * RH_KM_TO_M;                                                  // :243-244
  at[(size_t) RHB200_AT_HEIGHT * ndep] = s;
  at[(size_t) RHB200_AT_T * ndep]     = a[(size_t) 1 * ndep];
  at[(size_t) RHB200_AT_NE * ndep]    = a[(size_t) 2 * ndep] / CUBE(RH_CM_TO_M);
  at[(size_t) RHB200_AT_VEL * ndep]   = a[(size_t) 3 * ndep] * RH_KM_TO_M;
  at[(size_t) RHB200_AT_VTURB * ndep] = a[(size_t) 4 * ndep] * RH_KM_TO_M;
  at[(size_t) RHB200_AT_B * ndep]     = a[(size_t) 5 * ndep] / 1e4;
  at[(size_t) RHB200_AT_NHTOT * ndep] = a[(size_t) 8 * ndep] / CUBE(RH_CM_TO_M);
  at[(size_t) RHB200_AT_NP * ndep]    = 0.0;
  const double gamma_B = a[(size_t) 6 * ndep], chi_B = a[(size_t) 7 * ndep];
  double cg, c2, s2;
  if (muz == 1.0) {                                                                                 // project.c:52-58
    cg = rhm::rh_cos(gamma_B);
    c2 = rhm::rh_cos(2.0 * chi_B);
    s2 = rhm::rh_sin(2.0 * chi_B);
  } else {                                                                                          // project.c:60-77
    const double mux = sqrt(1.0 - SQ(muz)), muy = 0.0;                                              // pyrh_compute1dray.c:294-296
    const double csc_theta = 1.0 / sqrt(1.0 - SQ(muz));
    const double sin_gamma = rhm::rh_sin(gamma_B);
    const double bx = sin_gamma * rhm::rh_cos(chi_B);
    const double by = sin_gamma * rhm::rh_sin(chi_B);
    const double bz = rhm::rh_cos(gamma_B);
    const double b3 = mux*bx + muy*by + muz*bz;
    const double b1 = csc_theta * (bz - muz*b3);
    const double b2 = csc_theta * (muy*bx - mux*by);
    cg = b3;
    c2 = (SQ(b1) - SQ(b2)) / (1.0 - SQ(b3));
    s2 = 2.0 * b1*b2 / (1.0 - SQ(b3));
  }
  at[(size_t) RHB200_AT_COS_GAMMA * ndep] = cg;
  at[(size_t) RHB200_AT_COS_2CHI * ndep]  = c2;
  at[(size_t) RHB200_AT_SIN_2CHI * ndep]  = s2;
}

// one thread per column: a column none of whose |v| reaches VMACRO_TRESH is static (atmos.moving = FALSE), and
// the only thing the LTE path does with atmos.moving is to drop the Doppler shift (kurucz.c:749-754); zeroing
// the device copy of the velocity row gives the same +0.0 shift
__global__ void static_cols_kernel(int ncol, int ndep, double vmacro_tresh, double *__restrict__ atmos,
                                   int *__restrict__ col_moving)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  double *v = atmos + ((size_t) col * RHB200_AT_NFIELD + RHB200_AT_VEL) * ndep;
  bool moving = false;
  for (int k = 0; k < ndep; k++) if (fabs(v[k]) >= vmacro_tresh) { moving = true; break; }
  if (!moving) for (int k = 0; k < ndep; k++) v[k] = 0.0;
  col_moving[col] = moving ? 1 : 0;      // also selects the solver of unpolarised-line wavelengths (formal.c:100-103)
}

__global__ void __launch_bounds__(128)
proton_kernel(int ncol, int ndep, int nlev, int proton_level, const double *__restrict__ pops, double *__restrict__ atmos)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t % ndep);
  atmos[((size_t) col * RHB200_AT_NFIELD + RHB200_AT_NP) * ndep + k] = pops[((size_t) col * nlev + proton_level) * ndep + k];
}

// convertScales() for one column, sequential in depth like the reference.  `in` hands over the column: chi(k) = total
// opacity at the reference wavelength (spectrum.chi_c_lam[ref_index], readj.c:319: the last record written for that
// wavelength, i.e. continuum + lines of the up-ray), nH(k), T(k), ne(k) and scale(k) = the scale the column came with
// (tau_ref, cmass or height, SI).  height / tau / cmass [ndep] receive the three scales.
template <class In>
__device__ __forceinline__ void convert_scales(const In &in, int ndep, int atm_scale, double wght_per_H, double total_abund,
                                               double gravity, double *height, double *__restrict__ tau,
                                               double *__restrict__ cmass)
{
#define CHI(k) in.chi(k)
#define RHO(k) ((RH_AMU * wght_per_H) * in.nH(k))
  if (atm_scale == 0) {                                  // TAU500, multiatmos.c:140-151
    double tprev = in.scale(0), hprev = 0.0;
    tau[0] = tprev;
    height[0] = 0.0;
    double cprev = (tprev / CHI(0)) * RHO(0);
    cmass[0] = cprev;
    for (int k = 1; k < ndep; k++) {
      const double tk = in.scale(k);
      tau[k] = tk;
      const double hk = hprev - 2.0 * (tk - tprev) / (CHI(k-1) + CHI(k));
      const double ck = cprev + 0.5*(RHO(k-1) + RHO(k)) * (hprev - hk);
      height[k] = hk; cmass[k] = ck;
      hprev = hk; tprev = tk; cprev = ck;
    }
  } else if (atm_scale == 1) {                           // COLUMN_MASS, :128-138
    double cprev = in.scale(0), hprev = 0.0;
    double tprev = CHI(0) / RHO(0) * cprev;
    tau[0] = tprev; cmass[0] = cprev;
    height[0] = 0.0;
    for (int k = 1; k < ndep; k++) {
      const double ck = in.scale(k);
      const double hk = hprev - 2.0*(ck - cprev) / (RHO(k-1) + RHO(k));
      const double tk = tprev + 0.5*(CHI(k-1) + CHI(k)) * (hprev - hk);
      height[k] = hk; tau[k] = tk; cmass[k] = ck;
      hprev = hk; cprev = ck; tprev = tk;
    }
  } else {                                               // GEOMETRIC, :153-163: heights stay as they came
    double cprev = (in.nH(0) * total_abund + in.ne(0)) * (RH_KBOLTZMANN * in.T(0) / gravity);
    double tprev = 0.5 * CHI(0) * (in.scale(0) - in.scale(1));
    if (tprev > 1.0) tprev = 0.0;
    cmass[0] = cprev; tau[0] = tprev;
    height[0] = in.scale(0);
    for (int k = 1; k < ndep; k++) {
      const double ck = cprev + 0.5*(RHO(k-1) + RHO(k)) * (in.scale(k-1) - in.scale(k));
      const double tk = tprev + 0.5*(CHI(k-1) + CHI(k)) * (in.scale(k-1) - in.scale(k));
      cmass[k] = ck; tau[k] = tk;
      height[k] = in.scale(k);
      cprev = ck; tprev = tk;
    }
  }
  if (atm_scale != 2) {
    // Linear(Ndep, tau_ref, height, 1, &unity, &h_zero, FALSE), multiatmos.c:166-173
    const double unity = 1.0;
    double h_zero;
    const bool ascend = tau[1] > tau[0];
    const double xmin = ascend ? tau[0] : tau[ndep-1], xmax = ascend ? tau[ndep-1] : tau[0];
    if (unity <= xmin)      h_zero = ascend ? height[0] : height[ndep-1];
    else if (unity >= xmax) h_zero = ascend ? height[ndep-1] : height[0];
    else {
      const bool asc2 = tau[ndep-1] > tau[0];              // Locate(), hunt.c:97
      int lo = 0, hi = ndep;
      while (hi - lo > 1) {
        const int mid = (hi + lo) >> 1;
        if (asc2 ? (unity >= tau[mid]) : (unity <= tau[mid])) lo = mid; else hi = mid;
      }
      const double fx = (tau[lo+1] - unity) / (tau[lo+1] - tau[lo]);
      h_zero = fx*height[lo] + (1 - fx)*height[lo+1];
    }
    for (int k = 0; k < ndep; k++) height[k] = height[k] - h_zero;
  }
#undef CHI
#undef RHO
}

// one thread per column.  The height row of `atmos` holds the incoming scale and receives the heights;
// scratch [ncol][2][ndep] keeps tau_ref and cmass; scales_out [ncol][3][ndep] = height, tau_ref, cmass.
struct ColumnScalesIn {
  const double *rp; size_t kstride; const double *at; int ndep;
  __device__ __forceinline__ double chi(int k) const { return rp[(size_t) k * kstride]; }
  __device__ __forceinline__ double nH(int k) const { return at[(size_t) RHB200_AT_NHTOT * ndep + k]; }
  __device__ __forceinline__ double T(int k) const { return at[(size_t) RHB200_AT_T * ndep + k]; }
  __device__ __forceinline__ double ne(int k) const { return at[(size_t) RHB200_AT_NE * ndep + k]; }
  __device__ __forceinline__ double scale(int k) const { return at[(size_t) RHB200_AT_HEIGHT * ndep + k]; }
};
__global__ void scales_kernel(int ncol, int ndep, int atm_scale, double wght_per_H,
                              double total_abund, double gravity,
                              const double *__restrict__ chi_ref, size_t chi_col_stride, size_t chi_k_stride,
                              double *atmos, double *__restrict__ scratch, double *__restrict__ scales_out)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  double *height = atmos + ((size_t) col * RHB200_AT_NFIELD + RHB200_AT_HEIGHT) * ndep;
  double *tau = scratch + (size_t) col * 2 * ndep, *cmass = tau + ndep;
  ColumnScalesIn in{chi_ref + (size_t) col * chi_col_stride, chi_k_stride, atmos + (size_t) col * RHB200_AT_NFIELD * ndep, ndep};
  // the walk reads scale(k) once, before it writes height[k] (same row): in-place like the reference
  convert_scales(in, ndep, atm_scale, wght_per_H, total_abund, gravity, height, tau, cmass);
  if (scales_out) {
    double *o = scales_out + (size_t) col * 3 * ndep;
    for (int k = 0; k < ndep; k++) { o[k] = height[k]; o[ndep + k] = tau[k]; o[2*ndep + k] = cmass[k]; }
  }
}

// finite-difference response functions, single-depth perturbations (rhb200_rf_fd_batch): virtual column
// v = ((b*npar + p)*ndep + kp)*2 + s is base column b with parameter p changed at depth kp alone.  Everything before
// convertScales() is local in depth, so its depth-kp values are those of the PSEUDO column (b, p, s) -- the base with
// the parameter changed at every depth (full column b*(1 + 2 npar) + 1 + 2p + s).  This kernel walks the scales of
// the virtual column from the base's rows with the depth-kp entries replaced.  vws [nv][4][ndep]: height, T, tau, cmass.
struct VirtualScalesIn {
  const double *rp_b, *rp_q; size_t kstride; const double *at_b, *at_q; int ndep, kp;
  __device__ __forceinline__ double chi(int k) const { return (k == kp ? rp_q : rp_b)[(size_t) k * kstride]; }
  __device__ __forceinline__ double nH(int k) const { return (k == kp ? at_q : at_b)[(size_t) RHB200_AT_NHTOT * ndep + k]; }
  __device__ __forceinline__ double T(int k) const { return (k == kp ? at_q : at_b)[(size_t) RHB200_AT_T * ndep + k]; }
  __device__ __forceinline__ double ne(int k) const { return (k == kp ? at_q : at_b)[(size_t) RHB200_AT_NE * ndep + k]; }
  __device__ __forceinline__ double scale(int k) const { return at_b[(size_t) RHB200_AT_HEIGHT * ndep + k]; }
};
__global__ void __launch_bounds__(64)
vscales_kernel(int nv, int npar, int ndep, int nsel, const int *__restrict__ sel, int atm_scale, double wght_per_H, double total_abund, double gravity,
               const double *__restrict__ chi_ref, size_t chi_col_stride, size_t chi_k_stride,
               const double *__restrict__ atmos, double *__restrict__ vws)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int s = v & 1, kq = (v >> 1) % nsel, kp = sel ? sel[kq] : kq, p = ((v >> 1) / nsel) % npar, b = ((v >> 1) / nsel) / npar;
  const size_t fb = (size_t) b * (1 + 2*npar), fq = fb + 1 + 2*p + s;
  VirtualScalesIn in{chi_ref + fb * chi_col_stride, chi_ref + fq * chi_col_stride, chi_k_stride,
                     atmos + fb * RHB200_AT_NFIELD * ndep, atmos + fq * RHB200_AT_NFIELD * ndep, ndep, kp};
  double *w = vws + (size_t) v * 4 * ndep;
  for (int k = 0; k < ndep; k++) w[ndep + k] = in.T(k);
  convert_scales(in, ndep, atm_scale, wght_per_H, total_abund, gravity, w, w + 2*(size_t) ndep, w + 3*(size_t) ndep);
}

// ---- finite-difference response functions: perturbed copies of the base columns, made where they are consumed.
// virtual column v = ((base*npar + p)*ndep + kp)*2 + s: row rows[p] of column `base` changed at depth kp by
// +delta[p] (s = 0) or -delta[p] (s = 1); everything else is copied.  One thread per (virtual column, row, depth).
__global__ void __launch_bounds__(128)
rf_expand_kernel(int v0, int n, int ndep, int nsel, const int *__restrict__ sel, int nrow, int npar, const int *__restrict__ rows, const double *__restrict__ delta,
                 const double *__restrict__ base, double *__restrict__ out)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) n * nrow * ndep) return;
  const int k = (int) (t % ndep), r = (int) ((t / ndep) % nrow);
  const int v = v0 + (int) (t / ((size_t) ndep * nrow));
  const int s = v & 1, kq = (v >> 1) % nsel, kp = sel ? sel[kq] : kq, p = ((v >> 1) / nsel) % npar, b = ((v >> 1) / nsel) / npar;
  double x = base[((size_t) b * nrow + r) * ndep + k];
  if (r == rows[p] && k == kp) x = s ? x - delta[p] : x + delta[p];
  out[t] = x;
}

// the full columns of a chunk of base columns b0 .. b0+nb-1: column (b - b0)*(1 + 2 npar) is the base itself,
// + 1 + 2p + s the pseudo column with row rows[p] changed by +delta[p] (s = 0) / -delta[p] (s = 1) at EVERY depth
__global__ void __launch_bounds__(128)
rf_expand_full_kernel(int b0, int nb, int ndep, int nrow, int npar, const int *__restrict__ rows, const double *__restrict__ delta,
                      const double *__restrict__ base, double *__restrict__ out)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int nfull = 1 + 2*npar;
  if (t >= (size_t) nb * nfull * nrow * ndep) return;
  const int k = (int) (t % ndep), r = (int) ((t / ndep) % nrow);
  const int f = (int) (t / ((size_t) ndep * nrow)), b = f / nfull, j = f % nfull;
  double x = base[((size_t) (b0 + b) * nrow + r) * ndep + k];
  if (j > 0 && r == rows[(j - 1) >> 1]) x = ((j - 1) & 1) ? x - delta[(j - 1) >> 1] : x + delta[(j - 1) >> 1];
  out[t] = x;
}

// rf[pair][4][nlambda] = (S(+delta) - S(-delta)) / (2 delta); one thread per element
__global__ void __launch_bounds__(128)
rf_diff_kernel(int v0, int npair, int nsel, int n4l, int npar, const double *__restrict__ delta,
               const double *__restrict__ stokes, double *__restrict__ rf)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) npair * n4l) return;
  const int pair = (int) (t / n4l), e = (int) (t % n4l);
  const int p = (((v0 >> 1) + pair) / nsel) % npar;
  const double a = stokes[(size_t) (2*pair) * n4l + e], b = stokes[(size_t) (2*pair + 1) * n4l + e];
  rf[t] = (a - b) / (2.0 * delta[p]);
}

int rh_launch_rf_expand(rhb200_ctx *c, int v0, int n, int ndep, int nrow, int npar, const int *d_rows,
                        const double *d_delta, const double *d_base, double *d_in, int nsel, const int *d_sel)
{
  const size_t tot = (size_t) n * nrow * ndep;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  rf_expand_kernel<<<(unsigned) ((tot + 127) / 128), 128, 0, c->stream>>>(v0, n, ndep, nsel, d_sel, nrow, npar, d_rows, d_delta, d_base, d_in);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_rf_expand_full(rhb200_ctx *c, int b0, int nb, int ndep, int nrow, int npar, const int *d_rows,
                             const double *d_delta, const double *d_base, double *d_in)
{
  const size_t tot = (size_t) nb * (1 + 2*npar) * nrow * ndep;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  rf_expand_full_kernel<<<(unsigned) ((tot + 127) / 128), 128, 0, c->stream>>>(b0, nb, ndep, nrow, npar, d_rows, d_delta, d_base, d_in);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_vscales(rhb200_ctx *c, int nb, int npar, int ndep, int iref, int atm_scale, double wght_per_H,
                      double total_abund, double gravity, const double *d_raypts, const double *d_atmos, double *d_vws,
                      int nsel, const int *d_sel)
{
  const int nv = nb * npar * nsel * 2;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  vscales_kernel<<<(unsigned) ((nv + 63) / 64), 64, 0, c->stream>>>(nv, npar, ndep, nsel, d_sel, atm_scale, wght_per_H, total_abund, gravity,
      d_raypts + (size_t) iref * ndep * RP_NFIELD + RP_CHI, (size_t) c->wav.nlambda * ndep * RP_NFIELD, (size_t) RP_NFIELD,
      d_atmos, d_vws);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_rf_diff(rhb200_ctx *c, int v0, int n, int nsel, int nlambda, int npar, const double *d_delta,
                      const double *d_stokes, double *d_rf)
{
  const size_t tot = (size_t) (n / 2) * 4 * nlambda;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  rf_diff_kernel<<<(unsigned) ((tot + 127) / 128), 128, 0, c->stream>>>(v0, n / 2, nsel, 4 * nlambda, npar, d_delta, d_stokes, d_rf);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_pyrh_rows(rhb200_ctx *c, int ncol, int ndep, int nrow_in, int atm_scale, double muz, double vmacro_tresh,
                        const double *d_in, double *d_atmos, int *d_col_moving)
{
  const size_t cn = (size_t) ncol * ndep;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  pyrh_rows_kernel<<<(unsigned) ((cn + 127) / 128), 128, 0, c->stream>>>(ncol, ndep, nrow_in, atm_scale, muz, d_in, d_atmos);
  if (vmacro_tresh > 0.0)
    static_cols_kernel<<<(unsigned) ((ncol + 63) / 64), 64, 0, c->stream>>>(ncol, ndep, vmacro_tresh, d_atmos, d_col_moving);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_proton(rhb200_ctx *c, int ncol, int ndep, int nlev, int proton_level, const double *d_pops, double *d_atmos)
{
  const size_t cn = (size_t) ncol * ndep;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  proton_kernel<<<(unsigned) ((cn + 127) / 128), 128, 0, c->stream>>>(ncol, ndep, nlev, proton_level, d_pops, d_atmos);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_scales(rhb200_ctx *c, int ncol, int ndep, int iref, int atm_scale, double wght_per_H,
                     double total_abund, double gravity,
                     const double *d_raypts, double *d_atmos, double *d_scratch, double *d_scales_out)
{
  if (atm_scale == 2 && !d_scales_out) return RHB200_OK; // GEOMETRIC: the heights are the input
  ScopedKernelTimer t(c, RHB200_K_PREP);
  scales_kernel<<<(unsigned) ((ncol + 31) / 32), 32, 0, c->stream>>>(ncol, ndep, atm_scale, wght_per_H, total_abund, gravity,
      d_raypts + (size_t) iref * ndep * RP_NFIELD + RP_CHI, (size_t) c->wav.nlambda * ndep * RP_NFIELD, (size_t) RP_NFIELD,
      d_atmos, d_scratch, d_scales_out);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

// the same with the reference-wavelength opacity read from a plain [ncol][nlambda][ndep] array (NLTE: spectrum.chi_c_lam)
int rh_launch_scales_chi(rhb200_ctx *c, int ncol, int ndep, int nlambda, int iref, int atm_scale, double wght_per_H,
                         double total_abund, double gravity,
                         const double *d_chi_c, double *d_atmos, double *d_scratch, double *d_scales_out)
{
  if (atm_scale == 2 && !d_scales_out) return RHB200_OK;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  scales_kernel<<<(unsigned) ((ncol + 31) / 32), 32, 0, c->stream>>>(ncol, ndep, atm_scale, wght_per_H, total_abund, gravity,
      d_chi_c + (size_t) iref * ndep, (size_t) nlambda * ndep, (size_t) 1, d_atmos, d_scratch, d_scales_out);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}
