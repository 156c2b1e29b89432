// rhb200_delo.cu -- kernels + launchers for the polarised DELO-Bezier3 solver and the
// scalar cubic-Bezier short-characteristics solver (one thread per ray, depth sequential).
#include "rhb200_delo.cuh"
#include "rhb200_bezier.cuh"
#include "rhb200_feautrier.cuh"
#include "rhb200_piecewise.cuh"

namespace {

// ---- IO policy: ray-point records written by the fused opacity kernel, raypts[ray][k] =
//      {chi_I, K'_Q, K'_U, K'_V, S_I, S_Q, S_U, S_V}: one thread owns one ray and walks its own
//      64-byte records with four 128-bit loads per depth (every fetched sector is fully used).
struct RayPtsIO {
  const double2 *__restrict__ rp;  // this ray's records, 4 x double2 per depth
  double *out;                     // stokes + col*4*nlambda + l, stride nlambda between I,Q,U,V
  int nlambda, kout;
  __device__ __forceinline__ double chi(int k) const { return __ldg(reinterpret_cast<const double *>(rp + 4*(size_t)k)); }
  __device__ __forceinline__ void K(int k, double x[3]) const {
    const double2 a = __ldg(rp + 4*(size_t)k), b = __ldg(rp + 4*(size_t)k + 1);
    x[0] = a.y; x[1] = b.x; x[2] = b.y;
  }
  __device__ __forceinline__ void S(int k, double s[4]) const {
    const double2 a = __ldg(rp + 4*(size_t)k + 2), b = __ldg(rp + 4*(size_t)k + 3);
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
  }
  __device__ __forceinline__ void storeI(int k, const double I[4]) {
    if (k == kout) {
      out[0] = I[0]; out[nlambda] = I[1]; out[2*(size_t)nlambda] = I[2]; out[3*(size_t)nlambda] = I[3];
    }
  }
  __device__ __forceinline__ void storePsi(int, double) {}
  __device__ __forceinline__ void prefetch(int k, int ndep) const {
    if (k >= 0 && k < ndep) asm volatile("prefetch.global.L1 [%0];" :: "l"(rp + 4*(size_t)k));
  }
};

// ---- IO policy of the single-depth finite-difference columns (rhb200_rf_fd_batch): the base column's records
//      everywhere except at depth kp, where the pseudo column's record (parameter changed at every depth) is read
struct RayPtsPatchIO {
  const double2 *__restrict__ rp, *__restrict__ rq;
  int kp;
  double *out;
  int nlambda, kout;
  __device__ __forceinline__ const double2 *rec(int k) const { return (k == kp ? rq : rp) + 4*(size_t)k; }
  __device__ __forceinline__ double chi(int k) const { return __ldg(reinterpret_cast<const double *>(rec(k))); }
  __device__ __forceinline__ void K(int k, double x[3]) const {
    const double2 *r = rec(k);
    const double2 a = __ldg(r), b = __ldg(r + 1);
    x[0] = a.y; x[1] = b.x; x[2] = b.y;
  }
  __device__ __forceinline__ void S(int k, double s[4]) const {
    const double2 *r = rec(k);
    const double2 a = __ldg(r + 2), b = __ldg(r + 3);
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
  }
  __device__ __forceinline__ void storeI(int k, const double I[4]) {
    if (k == kout) {
      out[0] = I[0]; out[nlambda] = I[1]; out[2*(size_t)nlambda] = I[2]; out[3*(size_t)nlambda] = I[3];
    }
  }
  __device__ __forceinline__ void storePsi(int, double) {}
  __device__ __forceinline__ void prefetch(int k, int ndep) const {
    if (k >= 0 && k < ndep) asm volatile("prefetch.global.L1 [%0];" :: "l"(rec(k)));
  }
};

// ---- IO policy: reference layouts chi[nray][ndep], S[nray][4][ndep], chiQUV[nray][3][ndep]
struct GenericIO {
  const double *__restrict__ chi_, *__restrict__ S_, *__restrict__ q_;
  double *I_, *Psi_;
  int ndep;
  __device__ __forceinline__ double chi(int k) const { return chi_[k]; }
  __device__ __forceinline__ void K(int k, double x[3]) const {   // StokesK, stokesopac.c:72-77
    const double c = chi_[k];
    x[0] = q_[k] / c; x[1] = q_[ndep + k] / c; x[2] = q_[2*ndep + k] / c;
  }
  __device__ __forceinline__ void S(int k, double s[4]) const {
    s[0] = S_[k]; s[1] = S_[ndep+k]; s[2] = S_[2*ndep+k]; s[3] = S_[3*ndep+k];
  }
  __device__ __forceinline__ void storeI(int k, const double I[4]) {
    I_[k] = I[0]; I_[ndep+k] = I[1]; I_[2*ndep+k] = I[2]; I_[3*ndep+k] = I[3];
  }
  __device__ __forceinline__ void storePsi(int k, double p) { if (Psi_) Psi_[k] = p; }
  __device__ __forceinline__ void prefetch(int, int) const {}
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
delo_raypts_kernel(int ncol, int nlambda, int ndep, double muz, int bc_top, int bc_bottom,
                   const double *__restrict__ atmos, const double *__restrict__ lambda,
                   const int *__restrict__ wflags,
                   const double *__restrict__ raypts, double *__restrict__ stokes)
{
  const size_t nray = (size_t) ncol * nlambda;
  const size_t r = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nray) return;
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  if ((__ldg(wflags + l) & 2) == 0) return;      // no polarised line at this wavelength: scalar kernel (formal.c:84-103)
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  RayPtsIO io{reinterpret_cast<const double2 *>(raypts + r * (size_t) ndep * RP_NFIELD),
              stokes + (size_t) col * 4 * nlambda + l, nlambda, 0};
  rhd::delo_bezier3_ray(io, ndep, at + RHB200_AT_HEIGHT * ndep, muz, 1, bc_top, bc_bottom,
                        at + RHB200_AT_T * ndep, __ldg(lambda + l));
}

// the BASE column's polarised rays once, saving the state of the sweep at every depth (delo_bezier3_ray MODE 1):
// perturbations of parameters that leave the depth scales alone (v_z, v_mic, B, gamma, chi) resume from it.
// Heights / T: those of virtual column (b, pn, 0, 0), pn such a parameter -- bit-identical to the base column's.
__global__ void __launch_bounds__(128, 4)
delo_base_state_kernel(int nb, int npar, int pn, int nsel, int nlambda, int ndep, double muz, int bc_top, int bc_bottom,
                       const double *__restrict__ vws, const double *__restrict__ lambda, const int *__restrict__ wflags,
                       const double *__restrict__ raypts, double *__restrict__ state)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) nb * nlambda) return;
  const int b = (int) (t / nlambda), l = (int) (t - (size_t) b * nlambda);
  if ((__ldg(wflags + l) & 2) == 0) return;
  const size_t fb = (size_t) b * (1 + 2*npar);
  double sink[4];
  RayPtsIO io{reinterpret_cast<const double2 *>(raypts + (fb * nlambda + l) * (size_t) ndep * RP_NFIELD), sink, 1, -1};
  const double *w = vws + ((((size_t) b * npar + pn) * nsel + 0) * 2 + 0) * 4 * ndep;
  rhd::delo_bezier3_ray<RayPtsIO, 1>(io, ndep, w, muz, 1, bc_top, bc_bottom, w + ndep, __ldg(lambda + l),
                                      state + t * (size_t) ndep * DELO_NSTATE);
}

// single-depth finite-difference columns: one thread per (virtual column, wavelength); vws [nv][4][ndep] = height, T, ..
// of the virtual column (vscales_kernel), raypts those of the chunk's full columns (base + pseudo).
// neutral [npar] (or NULL): the parameter leaves heights and T alone -> resume the base sweep at depth kp + 2
__global__ void __launch_bounds__(128, 4)
delo_vcols_kernel(int nv, int npar, int nlambda, int ndep, double muz, int bc_top, int bc_bottom, int parabolic,
                  const double *__restrict__ vws, const double *__restrict__ lambda, const int *__restrict__ wflags,
                  const double *__restrict__ raypts, double *__restrict__ stokes,
                  const int *__restrict__ neutral, double *__restrict__ state, int nsel, const int *__restrict__ sel)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) nv * nlambda) return;
  const int v = (int) (t / nlambda), l = (int) (t - (size_t) v * nlambda);
  if ((__ldg(wflags + l) & 2) == 0) return;
  const int s = v & 1, kq = (v >> 1) % nsel, kp = sel ? sel[kq] : kq, p = ((v >> 1) / nsel) % npar, b = ((v >> 1) / nsel) / npar;
  const size_t fb = (size_t) b * (1 + 2*npar), fq = fb + 1 + 2*p + s;
  RayPtsPatchIO io{reinterpret_cast<const double2 *>(raypts + (fb * nlambda + l) * (size_t) ndep * RP_NFIELD),
                   reinterpret_cast<const double2 *>(raypts + (fq * nlambda + l) * (size_t) ndep * RP_NFIELD), kp,
                   stokes + (size_t) v * 4 * nlambda + l, nlambda, 0};
  const double *w = vws + (size_t) v * 4 * ndep;
  if (parabolic) rhp::stokes_parabolic_ray(io, ndep, w, muz, 1, bc_top, bc_bottom, w + ndep, __ldg(lambda + l));
  else if (state && neutral[p] && kp + 2 <= ndep - 2)
    rhd::delo_bezier3_ray<RayPtsPatchIO, 2>(io, ndep, w, muz, 1, bc_top, bc_bottom, w + ndep, __ldg(lambda + l),
                                            state + ((size_t) b * nlambda + l) * (size_t) ndep * DELO_NSTATE, kp + 2);
  else           rhd::delo_bezier3_ray(io, ndep, w, muz, 1, bc_top, bc_bottom, w + ndep, __ldg(lambda + l));
}

// fused LTE path with S_INTERPOLATION_STOKES = DELO_PARABOLIC (formal.c:215-216)
__global__ void __launch_bounds__(128)
stokes_parabolic_raypts_kernel(int ncol, int nlambda, int ndep, double muz, int bc_top, int bc_bottom,
                               const double *__restrict__ atmos, const double *__restrict__ lambda,
                               const int *__restrict__ wflags,
                               const double *__restrict__ raypts, double *__restrict__ stokes)
{
  const size_t nray = (size_t) ncol * nlambda;
  const size_t r = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nray) return;
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  if ((__ldg(wflags + l) & 2) == 0) return;
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  RayPtsIO io{reinterpret_cast<const double2 *>(raypts + r * (size_t) ndep * RP_NFIELD),
              stokes + (size_t) col * 4 * nlambda + l, nlambda, 0};
  rhp::stokes_parabolic_ray(io, ndep, at + RHB200_AT_HEIGHT * ndep, muz, 1, bc_top, bc_bottom,
                            at + RHB200_AT_T * ndep, __ldg(lambda + l));
}

__global__ void __launch_bounds__(128)
delo_generic_kernel(int solver, int nray, int ndep, double muz, int to_obs, int bc_top, int bc_bottom,
                    const int *__restrict__ ray_col, const double *__restrict__ ray_lambda,
                    const double *__restrict__ height, const double *__restrict__ T,
                    const double *__restrict__ chi, const double *__restrict__ S,
                    const double *__restrict__ chiQUV, double *__restrict__ I, double *__restrict__ Psi)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nray) return;
  const int col = ray_col[r];
  GenericIO io{chi + (size_t) r*ndep, S + (size_t) r*4*ndep, chiQUV + (size_t) r*3*ndep,
               I + (size_t) r*4*ndep, Psi ? Psi + (size_t) r*ndep : nullptr, ndep};
  if (solver == RHB200_DELO_PARABOLIC)
    rhp::stokes_parabolic_ray(io, ndep, height + (size_t) col*ndep, muz, to_obs, bc_top, bc_bottom,
                              T + (size_t) col*ndep, ray_lambda[r]);
  else
    rhd::delo_bezier3_ray(io, ndep, height + (size_t) col*ndep, muz, to_obs, bc_top, bc_bottom,
                          T + (size_t) col*ndep, ray_lambda[r]);
}

__global__ void __launch_bounds__(128)
bezier3_kernel(int solver, int nray, int ndep, double muz, int to_obs, int bc_top, int bc_bottom,
               const int *__restrict__ ray_col, const double *__restrict__ ray_lambda,
               const double *__restrict__ height, const double *__restrict__ T,
               const double *__restrict__ chi, const double *__restrict__ S,
               double *__restrict__ I, double *__restrict__ Psi)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nray) return;
  const int col = ray_col[r];
  const double *z = height + (size_t) col*ndep, *Tc = T + (size_t) col*ndep;
  const double *c = chi + (size_t) r*ndep, *s = S + (size_t) r*ndep;
  double *Ir = I + (size_t) r*ndep, *Pr = Psi ? Psi + (size_t) r*ndep : nullptr;
  switch (solver) {                                           // formal.c:229-235
  case RHB200_S_LINEAR:    rhp::linear_ray(ndep, z, muz, to_obs, bc_top, bc_bottom, Tc, ray_lambda[r], c, s, Ir, Pr); break;
  case RHB200_S_PARABOLIC: rhp::parabolic_ray(ndep, z, muz, to_obs, bc_top, bc_bottom, Tc, ray_lambda[r], c, s, Ir, Pr); break;
  default:                 rhz::bezier3_ray(ndep, z, muz, to_obs, bc_top, bc_bottom, Tc, ray_lambda[r], c, s, Ir, Pr);
  }
}

// Formal()'s two passes at one (wavelength, mu) in NO_STOKES mode with get_atomic_rfs (formal.c:167-283):
// the down-ray fills I, the up-ray overwrites it while its response-function branch still reads the
// down-ray values at the depths it has not reached yet.
__global__ void __launch_bounds__(128)
bezier3_rf_kernel(int nray, int ndep, double muz, int bc_top, int bc_bottom,
                  const int *__restrict__ ray_col, const double *__restrict__ ray_lambda,
                  const double *__restrict__ height, const double *__restrict__ T,
                  const double *__restrict__ chi_dn, const double *__restrict__ S_dn,
                  const double *__restrict__ chi_up, const double *__restrict__ S_up,
                  int npar, const double *__restrict__ dchi, const double *__restrict__ deta,
                  double *I, double *__restrict__ dI)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nray) return;
  const int col = ray_col[r];
  const double *z = height + (size_t) col*ndep, *Tc = T + (size_t) col*ndep;
  double *Ir = I + (size_t) r*ndep;
  rhz::bezier3_ray_t<false>(ndep, z, muz, 0, bc_top, bc_bottom, Tc, ray_lambda[r], chi_dn + (size_t) r*ndep,
                            S_dn + (size_t) r*ndep, Ir, nullptr, 0, nullptr, nullptr, nullptr);
  rhz::bezier3_ray_t<true>(ndep, z, muz, 1, bc_top, bc_bottom, Tc, ray_lambda[r], chi_up + (size_t) r*ndep,
                           S_up + (size_t) r*ndep, Ir, nullptr, npar, dchi + (size_t) r*ndep*npar,
                           deta + (size_t) r*ndep*npar, dI + (size_t) r*ndep*npar);
}

// ---- Feautrier IO policies
struct FeauRayPtsIO {           // fused LTE path: line-free rays; F and z live in the unused K' slots
  double *rp;                   // this ray's records [k][RP_NFIELD]
  const double *__restrict__ h;
  __device__ __forceinline__ double chi(int k) const { return rp[(size_t) k*RP_NFIELD + RP_CHI]; }
  __device__ __forceinline__ double S(int k) const { return rp[(size_t) k*RP_NFIELD + RP_SI]; }
  __device__ __forceinline__ double z(int k) const { return __ldg(h + k); }
  __device__ __forceinline__ void putF(int k, double v) { rp[(size_t) k*RP_NFIELD + RP_KQ] = v; }
  __device__ __forceinline__ void putZ(int k, double v) { rp[(size_t) k*RP_NFIELD + RP_KU] = v; }
  __device__ __forceinline__ double getF(int k) const { return rp[(size_t) k*RP_NFIELD + RP_KQ]; }
  __device__ __forceinline__ double getZ(int k) const { return rp[(size_t) k*RP_NFIELD + RP_KU]; }
  bool keepP = false;           // N_MAX_SCATTER > 0: J = P (one ray, wmu = 1) goes into the S_Q slot
  double dJ = 0.0;              // max |1 - Jdag/J| over the depths (formal.c:312-316), Jdag = what the slot held
  __device__ __forceinline__ void storeP(int k, double v) {
    if (keepP) {
      double *j = rp + (size_t) k*RP_NFIELD + RP_SQ;
      const double d = fabs(1.0 - *j / v);
      dJ = (dJ > d) ? dJ : d;
      *j = v;
    }
  }
  __device__ __forceinline__ void storePsi(int, double) {}
  __device__ __forceinline__ bool wantPsi() const { return false; }
};
struct FeauGenericIO {          // reference layouts; P and Psi double as the F / z scratch
  const double *__restrict__ chi_, *__restrict__ S_, *__restrict__ h;
  double *P_, *Psi_, *scr;      // scr: [2][ndep] scratch when Psi is not requested
  int ndep;
  __device__ __forceinline__ double chi(int k) const { return chi_[k]; }
  __device__ __forceinline__ double S(int k) const { return S_[k]; }
  __device__ __forceinline__ double z(int k) const { return h[k]; }
  __device__ __forceinline__ void putF(int k, double v) { scr[k] = v; }
  __device__ __forceinline__ void putZ(int k, double v) { scr[ndep + k] = v; }
  __device__ __forceinline__ double getF(int k) const { return scr[k]; }
  __device__ __forceinline__ double getZ(int k) const { return scr[ndep + k]; }
  __device__ __forceinline__ void storeP(int k, double v) { P_[k] = v; }
  __device__ __forceinline__ void storePsi(int k, double v) { Psi_[k] = v; }
  __device__ __forceinline__ bool wantPsi() const { return Psi_ != nullptr; }
};

// rays of the fused LTE path without a polarised line, solved for I alone.  formal.c:84-103: angle_dep is false when
// the wavelength has no line, or has one but the column is static -> Feautrier (:289-309); a moving column with an
// unpolarised line takes the scalar S_INTERPOLATION ray (:223-236).  scratch [ncol][nunpol][fields][ndep]: chi, S, I, ...
__global__ void __launch_bounds__(128)
feautrier_raypts_kernel(int ncol, int nlambda, int ndep, double muz, int bc_top, int bc_bottom,
                        const int *__restrict__ nolines, int nnoline,
                        const double *__restrict__ atmos, const double *__restrict__ lambda,
                        double *__restrict__ raypts, double *__restrict__ stokes,
                        const int *__restrict__ wflags, const int *__restrict__ unpol_rank, int nunpol,
                        int solver, int moving, const int *__restrict__ col_moving, double *__restrict__ scratch, int fields, int keepP)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nnoline) return;
  const int col = (int) (t / nnoline), l = __ldg(nolines + (int) (t - (size_t) col * nnoline));
  const size_t r = (size_t) col * nlambda + l;
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  double *rp = raypts + r * (size_t) ndep * RP_NFIELD;
  double I0;
  const int fl = __ldg(wflags + l);                  // polarised line (NO_STOKES only): always angle dependent
  const bool scalar_ray = (fl & 2) || ((fl & 1) && (col_moving ? col_moving[col] != 0 : moving != 0));
  if (scalar_ray) {
    double *c = scratch + ((size_t) col * nunpol + __ldg(unpol_rank + l)) * fields * ndep, *s = c + ndep, *Ir = s + ndep;
    for (int k = 0; k < ndep; k++) { c[k] = rp[(size_t) k * RP_NFIELD + RP_CHI]; s[k] = rp[(size_t) k * RP_NFIELD + RP_SI]; }
    const double *z = at + RHB200_AT_HEIGHT * ndep, *Tc = at + RHB200_AT_T * ndep;
    const double lam = __ldg(lambda + l);
    switch (solver) {                                         // formal.c:229-235
    case RHB200_S_LINEAR:    rhp::linear_ray(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr); break;
    case RHB200_S_PARABOLIC: rhp::parabolic_ray(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr); break;
    default:                 rhz::bezier3_ray(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr);
    }
    I0 = Ir[0];
  } else {
    FeauRayPtsIO io{rp, at + RHB200_AT_HEIGHT * ndep};
    io.keepP = keepP != 0;                                   // S_Q slot holds 0 = the J Iterate() starts from
    I0 = rhf::feautrier_ray(io, ndep, muz, bc_top, bc_bottom, at + RHB200_AT_T * ndep, __ldg(lambda + l));
  }
  double *out = stokes + (size_t) col * 4 * nlambda + l;
  out[0] = I0; out[nlambda] = 0.0; out[2*(size_t) nlambda] = 0.0; out[3*(size_t) nlambda] = 0.0;
}

// the same rays for the single-depth finite-difference columns: chi and S of the virtual column are gathered into its
// own scratch rows (the base records are shared by 2 npar ndep virtual columns, so Feautrier's F / z cannot live in
// them).  scratch [nv][nnoline][5][ndep]: chi, S, I / P, F, z.  vmacro_tresh = 0 on this path: every column is moving.
__global__ void __launch_bounds__(128)
noline_vcols_kernel(int nv, int npar, int nlambda, int ndep, double muz, int bc_top, int bc_bottom,
                    const int *__restrict__ nolines, int nnoline, const double *__restrict__ vws,
                    const double *__restrict__ lambda, const double *__restrict__ raypts, double *__restrict__ stokes,
                    const int *__restrict__ wflags, int solver, double *__restrict__ scratch, int nsel, const int *__restrict__ sel)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) nv * nnoline) return;
  const int v = (int) (t / nnoline), l = __ldg(nolines + (int) (t - (size_t) v * nnoline));
  const int sg = v & 1, kq = (v >> 1) % nsel, kp = sel ? sel[kq] : kq, p = ((v >> 1) / nsel) % npar, b = ((v >> 1) / nsel) / npar;
  const size_t fb = (size_t) b * (1 + 2*npar), fq = fb + 1 + 2*p + sg;
  const double *rb = raypts + (fb * nlambda + l) * (size_t) ndep * RP_NFIELD, *rq = raypts + (fq * nlambda + l) * (size_t) ndep * RP_NFIELD;
  const double *z = vws + (size_t) v * 4 * ndep, *Tc = z + ndep;
  double *c = scratch + t * 5 * (size_t) ndep, *s = c + ndep, *Ir = s + ndep;
  for (int k = 0; k < ndep; k++) {
    const double *r = (k == kp ? rq : rb) + (size_t) k * RP_NFIELD;
    c[k] = r[RP_CHI]; s[k] = r[RP_SI];
  }
  const int fl = __ldg(wflags + l);
  const double lam = __ldg(lambda + l);
  double I0;
  if (fl & 3) {                                               // a line and a moving column: the scalar S_INTERPOLATION ray
    switch (solver) {                                         // formal.c:229-235
    case RHB200_S_LINEAR:    rhp::linear_ray(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr); break;
    case RHB200_S_PARABOLIC: rhp::parabolic_ray(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr); break;
    default:                 rhz::bezier3_ray(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr);
    }
    I0 = Ir[0];
  } else {
    FeauGenericIO io{c, s, z, Ir, nullptr, Ir + ndep, ndep};
    I0 = rhf::feautrier_ray(io, ndep, muz, bc_top, bc_bottom, Tc, lam);
  }
  double *out = stokes + (size_t) v * 4 * nlambda + l;
  out[0] = I0; out[nlambda] = 0.0; out[2*(size_t) nlambda] = 0.0; out[3*(size_t) nlambda] = 0.0;
}

// One pass of the LTE scattering iteration (pyrh_compute1dray.c:332-337 -> solveSpectrum -> Formal, angle-independent
// branch formal.c:289-309) over the Feautrier wavelengths of the columns that have not converged yet:
// S = (eta + sca Jdag)/chi, new J = P, the column's dJmax through an atomic max on the bit pattern (dJ >= 0).
__global__ void __launch_bounds__(128)
scatter_pass_kernel(int ncol, int nlambda, int ndep, double muz, int bc_top, int bc_bottom,
                    const int *__restrict__ nolines, int nnoline, const double *__restrict__ atmos,
                    const double *__restrict__ lambda, double *__restrict__ raypts, double *__restrict__ stokes,
                    const int *__restrict__ wflags, int moving, const int *__restrict__ col_moving,
                    const int *__restrict__ done, unsigned long long *__restrict__ colmax)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nnoline) return;
  const int col = (int) (t / nnoline), l = __ldg(nolines + (int) (t - (size_t) col * nnoline));
  if (done[col]) return;
  const int fl = __ldg(wflags + l);
  if ((fl & 2) || ((fl & 1) && (col_moving ? col_moving[col] != 0 : moving != 0))) return;   // angle dependent: no J feedback in LTE
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  double *rp = raypts + ((size_t) col * nlambda + l) * (size_t) ndep * RP_NFIELD;
  for (int k = 0; k < ndep; k++) {
    double *r = rp + (size_t) k * RP_NFIELD;
    r[RP_SI] = (r[RP_SU] + r[RP_SV] * r[RP_SQ]) / r[RP_CHI];
  }
  FeauRayPtsIO io{rp, at + RHB200_AT_HEIGHT * ndep};
  io.keepP = true;
  const double I0 = rhf::feautrier_ray(io, ndep, muz, bc_top, bc_bottom, at + RHB200_AT_T * ndep, __ldg(lambda + l));
  stokes[(size_t) col * 4 * nlambda + l] = I0;
  atomicMax(colmax + col, (unsigned long long) __double_as_longlong(io.dJ));
}

// get_atomic_rfs (formal.c:59-65, 124-127, 278-282): the log gf response function of the emergent intensity at the
// wavelengths Formal() solves with Piecewise_Bezier3_1D, i.e. the scalar rays of feautrier_raypts_kernel when
// S_INTERPOLATION = S_BEZIER3; every other wavelength keeps the zeros Formal() starts from.  The down-ray that
// precedes the up-ray in Formal() -- its intensities are what the response-function branch reads at the depths the
// up-ray has not reached (bezier_1D.c:483) -- sees the SAME opacity as the up-ray: pyrh keeps one background record
// per wavelength in memory (spectrum.chi_c_lam[nspect], no direction index), and the last Background() pass, the one
// that stays, is to_obs = 1.  The scratch row of the ray still holds chi, S of feautrier_raypts_kernel; dchi/deta
// were written by loggf_dopac_kernel.  One thread per (column, wavelength).
__global__ void __launch_bounds__(128)
loggf_rf_kernel(int ncol, int nlambda, int ndep, double muz, int bc_top, int bc_bottom, const double *__restrict__ atmos,
                const double *__restrict__ lambda, const int *__restrict__ wflags,
                const int *__restrict__ unpol_rank, int nunpol, int solver, int no_stokes, int moving,
                const int *__restrict__ col_moving, double *__restrict__ scratch, int fields, int npar, double *__restrict__ rf)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nlambda) return;
  const int col = (int) (t / nlambda), l = (int) (t - (size_t) col * nlambda);
  double *out = rf + t * npar;
  const int fl = __ldg(wflags + l), rank = __ldg(unpol_rank + l);
  const bool mov = col_moving ? col_moving[col] != 0 : moving != 0;
  const bool scalar_ray = rank >= 0 && (no_stokes ? ((fl & 2) || ((fl & 1) && mov)) : (fl == 1 && mov));
  if (!scalar_ray || solver != RHB200_S_BEZIER3) {
    for (int p = 0; p < npar; p++) out[p] = 0.0;
    return;
  }
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double *z = at + RHB200_AT_HEIGHT * ndep, *Tc = at + RHB200_AT_T * ndep;
  double *c = scratch + ((size_t) col * nunpol + rank) * fields * ndep, *s = c + ndep, *Ir = s + ndep,
         *dchi = Ir + ndep, *deta = dchi + (size_t) npar * ndep, *dI = deta + (size_t) npar * ndep;
  const double lam = __ldg(lambda + l);
  rhz::bezier3_ray_t<false>(ndep, z, muz, 0, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr, 0, nullptr, nullptr, nullptr);
  rhz::bezier3_ray_t<true>(ndep, z, muz, 1, bc_top, bc_bottom, Tc, lam, c, s, Ir, nullptr, npar, dchi, deta, dI);
  for (int p = 0; p < npar; p++) out[p] = dI[p];       // dI[0][p], formal.c:278-282
}

__global__ void scatter_update_kernel(int ncol, double limit, int *__restrict__ done, unsigned long long *__restrict__ colmax)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  if (!done[col] && __longlong_as_double((long long) colmax[col]) <= limit) done[col] = 1;   // solveSpectrum() <= iterLimit: break
  colmax[col] = 0ull;
}

__global__ void __launch_bounds__(128)
feautrier_generic_kernel(int nray, int ndep, double muz, int bc_top, int bc_bottom,
                         const int *__restrict__ ray_col, const double *__restrict__ ray_lambda,
                         const double *__restrict__ height, const double *__restrict__ T,
                         const double *__restrict__ chi, const double *__restrict__ S,
                         double *__restrict__ P, double *__restrict__ Psi, double *__restrict__ Iem,
                         double *__restrict__ scratch)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nray) return;
  const int col = ray_col[r];
  FeauGenericIO io{chi + (size_t) r*ndep, S + (size_t) r*ndep, height + (size_t) col*ndep,
                   P + (size_t) r*ndep, Psi ? Psi + (size_t) r*ndep : nullptr,
                   scratch + (size_t) r*2*ndep, ndep};
  Iem[r] = rhf::feautrier_ray(io, ndep, muz, bc_top, bc_bottom, T + (size_t) col*ndep, ray_lambda[r]);
}

}  // namespace

int rh_launch_delo_raypts(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom,
                          const double *d_atmos, const double *d_raypts, double *d_stokes)
{
  const size_t nray = (size_t) ncol * ctx->wav.nlambda;
  if (nray == 0) return RHB200_OK;
  const int threads = 128;
  const unsigned blocks = (unsigned) ((nray + threads - 1) / threads);
  {
    ScopedKernelTimer t(ctx, RHB200_K_DELO);
    static int variant = -1;
    if (variant < 0) { const char *e = getenv("RHB200_DELO_MINB"); variant = e ? atoi(e) : 4; }
#define RH_LAUNCH_DELO(M) delo_raypts_kernel<M><<<blocks, threads, 0, ctx->stream>>>(ncol, ctx->wav.nlambda, ndep, muz, \
        bc_top, bc_bottom, d_atmos, ctx->wav.lambda, ctx->wav.flags, d_raypts, d_stokes)
    if (ctx->s_interpolation_stokes == RHB200_DELO_PARABOLIC)
      stokes_parabolic_raypts_kernel<<<blocks, threads, 0, ctx->stream>>>(ncol, ctx->wav.nlambda, ndep, muz,
          bc_top, bc_bottom, d_atmos, ctx->wav.lambda, ctx->wav.flags, d_raypts, d_stokes);
    else switch (variant) {
    case 4: RH_LAUNCH_DELO(4); break;
    case 5: RH_LAUNCH_DELO(5); break;
    case 6: RH_LAUNCH_DELO(6); break;
    default: RH_LAUNCH_DELO(3); break;
    }
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_delo_vcols(rhb200_ctx *ctx, int nb, int npar, int ndep, double muz, int bc_top, int bc_bottom,
                         const double *d_vws, const double *d_raypts, double *d_stokes,
                         const int *d_neutral, int pn, double *d_state, int nsel, const int *d_sel)
{
  const int nv = nb * npar * nsel * 2;
  const size_t nray = (size_t) nv * ctx->wav.nlambda;
  if (nray == 0) return RHB200_OK;
  const bool parabolic = ctx->s_interpolation_stokes == RHB200_DELO_PARABOLIC;
  if (parabolic || pn < 0) d_state = nullptr;
  {
    ScopedKernelTimer t(ctx, RHB200_K_DELO);
    if (d_state)
      delo_base_state_kernel<<<(unsigned) (((size_t) nb * ctx->wav.nlambda + 127) / 128), 128, 0, ctx->stream>>>(nb, npar, pn, nsel,
          ctx->wav.nlambda, ndep, muz, bc_top, bc_bottom, d_vws, ctx->wav.lambda, ctx->wav.flags, d_raypts, d_state);
    delo_vcols_kernel<<<(unsigned) ((nray + 127) / 128), 128, 0, ctx->stream>>>(nv, npar, ctx->wav.nlambda, ndep, muz, bc_top, bc_bottom,
        parabolic, d_vws, ctx->wav.lambda, ctx->wav.flags, d_raypts, d_stokes, d_neutral, d_state, nsel, d_sel);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_noline_vcols(rhb200_ctx *ctx, int nb, int npar, int ndep, double muz, int bc_top, int bc_bottom,
                           const double *d_vws, const double *d_raypts, double *d_stokes, double *d_scratch,
                           int nsel, const int *d_sel)
{
  const int nv = nb * npar * nsel * 2, nn = ctx->wav.nnoline;
  const size_t n = (size_t) nv * nn;
  if (n == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    noline_vcols_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(nv, npar, ctx->wav.nlambda, ndep, muz, bc_top, bc_bottom,
        ctx->wav.noline, nn, d_vws, ctx->wav.lambda, d_raypts, d_stokes, ctx->wav.flags, ctx->s_interpolation, d_scratch, nsel, d_sel);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_delo_generic(rhb200_ctx *ctx, int solver, int nray, int ndep, double muz, int to_obs,
                           int bc_top, int bc_bottom, const int *d_ray_col,
                           const double *d_ray_lambda, const double *d_height, const double *d_T,
                           const double *d_chi, const double *d_S, const double *d_chiQUV,
                           double *d_I, double *d_Psi)
{
  if (nray == 0) return RHB200_OK;
  const int threads = 128, blocks = (nray + threads - 1) / threads;
  {
    ScopedKernelTimer t(ctx, RHB200_K_DELO);
    delo_generic_kernel<<<blocks, threads, 0, ctx->stream>>>(solver, nray, ndep, muz, to_obs, bc_top, bc_bottom,
                                                              d_ray_col, d_ray_lambda, d_height, d_T,
                                                              d_chi, d_S, d_chiQUV, d_I, d_Psi);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_bezier3(rhb200_ctx *ctx, int solver, int nray, int ndep, double muz, int to_obs,
                      int bc_top, int bc_bottom, const int *d_ray_col,
                      const double *d_ray_lambda, const double *d_height, const double *d_T,
                      const double *d_chi, const double *d_S, double *d_I, double *d_Psi)
{
  if (nray == 0) return RHB200_OK;
  const int threads = 128, blocks = (nray + threads - 1) / threads;
  {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    bezier3_kernel<<<blocks, threads, 0, ctx->stream>>>(solver, nray, ndep, muz, to_obs, bc_top, bc_bottom,
                                                         d_ray_col, d_ray_lambda, d_height, d_T,
                                                         d_chi, d_S, d_I, d_Psi);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_bezier3_rf(rhb200_ctx *ctx, int nray, int ndep, double muz, int bc_top, int bc_bottom,
                         const int *d_ray_col, const double *d_ray_lambda, const double *d_height, const double *d_T,
                         const double *d_chi_dn, const double *d_S_dn, const double *d_chi_up, const double *d_S_up,
                         int npar, const double *d_dchi, const double *d_deta, double *d_I, double *d_dI)
{
  if (nray == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    bezier3_rf_kernel<<<(nray + 127) / 128, 128, 0, ctx->stream>>>(nray, ndep, muz, bc_top, bc_bottom, d_ray_col,
        d_ray_lambda, d_height, d_T, d_chi_dn, d_S_dn, d_chi_up, d_S_up, npar, d_dchi, d_deta, d_I, d_dI);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_feautrier_raypts(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom,
                               const double *d_atmos, double *d_raypts, double *d_stokes,
                               int moving, const int *d_col_moving, double *d_scratch)
{
  const int nn = ctx->wav.nnoline;
  if (nn == 0 || ncol == 0) return RHB200_OK;
  const size_t n = (size_t) ncol * nn;
  {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    feautrier_raypts_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(
        ncol, ctx->wav.nlambda, ndep, muz, bc_top, bc_bottom, ctx->wav.noline, nn, d_atmos,
        ctx->wav.lambda, d_raypts, d_stokes, ctx->wav.flags, ctx->wav.unpol_rank, ctx->wav.nunpol,
        ctx->s_interpolation, moving, d_col_moving, d_scratch, ctx->scal_fields(), ctx->n_max_scatter > 0);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_loggf_rf(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom, const double *d_atmos,
                       int moving, const int *d_col_moving, double *d_scratch, double *d_rf)
{
  const size_t n = (size_t) ncol * ctx->wav.nlambda;
  if (n == 0 || ctx->lrf_npar == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    loggf_rf_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(
        ncol, ctx->wav.nlambda, ndep, muz, bc_top, bc_bottom, d_atmos, ctx->wav.lambda, ctx->wav.flags,
        ctx->wav.unpol_rank, ctx->wav.nunpol, ctx->s_interpolation, ctx->no_stokes, moving, d_col_moving, d_scratch,
        ctx->scal_fields(), ctx->lrf_npar, d_rf);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_feautrier(rhb200_ctx *ctx, int nray, int ndep, double muz, int bc_top, int bc_bottom,
                        const int *d_ray_col, const double *d_ray_lambda, const double *d_height,
                        const double *d_T, const double *d_chi, const double *d_S,
                        double *d_P, double *d_Psi, double *d_Iem, double *d_scratch)
{
  if (nray == 0) return RHB200_OK;
  {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    feautrier_generic_kernel<<<(nray + 127) / 128, 128, 0, ctx->stream>>>(
        nray, ndep, muz, bc_top, bc_bottom, d_ray_col, d_ray_lambda, d_height, d_T, d_chi, d_S,
        d_P, d_Psi, d_Iem, d_scratch);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_launch_scatter_passes(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom, const double *d_atmos,
                             double *d_raypts, double *d_stokes, int moving, const int *d_col_moving,
                             unsigned long long *d_colmax, int *d_done)
{
  const int nn = ctx->wav.nnoline;
  if (ctx->n_max_scatter <= 0 || nn == 0 || ncol == 0) return RHB200_OK;
  RH_CUDA(cudaMemsetAsync(d_colmax, 0, (size_t) ncol * sizeof(unsigned long long), ctx->stream));
  RH_CUDA(cudaMemsetAsync(d_done, 0, (size_t) ncol * sizeof(int), ctx->stream));
  const size_t n = (size_t) ncol * nn;
  for (int it = 0; it < ctx->n_max_scatter; it++) {
    ScopedKernelTimer t(ctx, RHB200_K_BEZIER);
    scatter_pass_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(ncol, ctx->wav.nlambda, ndep, muz, bc_top, bc_bottom,
        ctx->wav.noline, nn, d_atmos, ctx->wav.lambda, d_raypts, d_stokes, ctx->wav.flags, moving, d_col_moving, d_done, d_colmax);
    scatter_update_kernel<<<(unsigned) ((ncol + 127) / 128), 128, 0, ctx->stream>>>(ncol, ctx->scatter_limit, d_done, d_colmax);
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}
