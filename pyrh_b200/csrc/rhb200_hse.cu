// rhb200_hse.cu -- electron density from the LTE ionisation of all elements (Solve_ne, rh/solvene.c:55-140, with
// getfjk :145-200 and getKuruczpf :205-215) for pyrh.get_ne_from_nH (rhf1d/pyrh_hse.c:555-677), and the
// layer-by-layer variant get_ne() of the hydrostatic-equilibrium solver (rhf1d/pyrh_background.c).
#include "rhb200_common.cuh"
#include "rhb200_math.cuh"
#include <algorithm>
#include <cmath>

#define RH_EV 1.60217733E-19
#define N_MAX_ELECTRON_ITERATIONS 10      // solvene.c:38
#define MAX_ELECTRON_ERROR 1.0E-2         // solvene.c:37

struct ElementTable {
  int nelem = 0, npf = 0, npf_rows = 0;
  double *elems = nullptr, *pf = nullptr, *Tpf = nullptr;
  ~ElementTable() { cudaFree(elems); cudaFree(pf); cudaFree(Tpf); }
};

namespace {

// Linear() at one point inside the table, clamped outside (linear.c:22-51; Hunt and Locate agree on the bracket)
__device__ __forceinline__ double pf_interp(int nt, const double *__restrict__ xt, const double *__restrict__ yt, double x)
{
  if (x <= xt[0]) return yt[0];
  if (x >= xt[nt-1]) return yt[nt-1];
  int lo = 0, hi = nt;
  while (hi - lo > 1) { const int mid = (hi + lo) >> 1; if (x >= xt[mid]) lo = mid; else hi = mid; }
  const double fx = (xt[lo+1] - x) / (xt[lo+1] - xt[lo]);
  return fx*yt[lo] + (1 - fx)*yt[lo+1];
}

// one depth point: Newton iteration on the charge-conservation equation.  uk_zero: get_ne() of the HSE solver starts
// from PhiH with U = 0 (pyrh_background.c), Solve_ne with the interpolated ln U of H I (solvene.c:83-86).
__device__ double solve_ne_point(int nelem, int npf, const double *__restrict__ elems, const double *__restrict__ pf,
                                 const double *__restrict__ Tpf, double T, double nHtot, double ne_in,
                                 int fromscratch, int uk_zero)
{
  const double C1 = (RH_HPLANCK/(2.0*RH_PI*RH_M_ELECTRON)) * (RH_HPLANCK/RH_KBOLTZMANN);
  double ne_old, ne = ne_in;
  if (fromscratch) {
    const double Uk = uk_zero ? 0.0 : pf_interp(npf, Tpf, pf + (size_t) ((int) elems[RHB200_RE_PFROW]) * npf, T);
    const double PhiH = 0.5 * rhm::rh_pow(C1/T, 1.5) * rhm::rh_exp(Uk + elems[RHB200_RE_IONPOT0]/(RH_KBOLTZMANN*T));
    ne_old = (sqrt(1.0 + 4.0*nHtot*PhiH) - 1.0) / (2.0*PhiH);
    ne = ne_old;
  } else ne_old = ne_in;
  double fjk[RHB200_RE_MAXSTAGE], dfjk[RHB200_RE_MAXSTAGE];
  for (int niter = 0; niter < N_MAX_ELECTRON_ITERATIONS; niter++) {
    double error = ne_old / nHtot, sum = 0.0;
    for (int n = 0; n < nelem; n++) {
      const double *e = elems + (size_t) n * RHB200_RE_NFIELD;
      const int nst = (int) e[RHB200_RE_NSTAGE], row = (int) e[RHB200_RE_PFROW];
      // getfjk, LTE branch (solvene.c:172-198)
      const double CT_ne = 2.0 * rhm::rh_pow(C1/T, -1.5) / ne_old;
      double sum1 = 1.0, sum2 = 0.0;
      fjk[0] = 1.0; dfjk[0] = 0.0;
      double Uk = pf_interp(npf, Tpf, pf + (size_t) row * npf, T);
      for (int j = 1; j < nst; j++) {
        const double Ukp1 = pf_interp(npf, Tpf, pf + (size_t) (row + j) * npf, T);
        fjk[j]  = fjk[j-1] * CT_ne * rhm::rh_exp(Ukp1 - Uk - e[RHB200_RE_IONPOT0 + j-1]/(RH_KBOLTZMANN*T));
        dfjk[j] = -j * fjk[j] / ne_old;
        sum1 += fjk[j];
        sum2 += dfjk[j];
        Uk = Ukp1;
      }
      for (int j = 0; j < nst; j++) {
        fjk[j] /= sum1;
        dfjk[j] = (dfjk[j] - fjk[j] * sum2) / sum1;
      }
      if (n == 0) {                                          // H-minus, solvene.c:106-111
        const double PhiHmin = 0.25*rhm::rh_pow(C1/T, 1.5) * rhm::rh_exp(0.754 * RH_EV / (RH_KBOLTZMANN * T));
        error += ne_old * fjk[0] * PhiHmin;
        sum   -= (fjk[0] + ne_old * dfjk[0]) * PhiHmin;
      }
      for (int j = 1; j < nst; j++) {
        const double akj = e[RHB200_RE_ABUND] * j;
        error -= akj * fjk[j];
        sum   += akj * dfjk[j];
      }
    }
    ne = ne_old - nHtot * error / (1.0 - nHtot * sum);
    const double dne = fabs((ne - ne_old)/ne_old);
    ne_old = ne;
    if (dne <= MAX_ELECTRON_ERROR) break;
  }
  return ne;
}

// The same for small batches, one WARP per depth point: a GPU thread needs ~1 ms for the ~500 Saha factors of one
// Newton step, which is what a single-column pyrh.hse call waits for 125 times.  Lane l evaluates getfjk() of the
// elements l, l + 32, ... and leaves the terms akj fjk[j], akj dfjk[j] in shared memory; lane 0 then adds them to
// `error` and `sum` in the reference's (element, stage) order, so the result is the bit pattern of solve_ne_point().
// sh: 2 * nelem * RHB200_RE_MAXSTAGE doubles per warp.
__device__ double solve_ne_point_coop(int nelem, int npf, const double *__restrict__ elems, const double *__restrict__ pf,
                                      const double *__restrict__ Tpf, double T, double nHtot, double ne_in,
                                      int fromscratch, int uk_zero, double *sh)
{
  constexpr int MS = RHB200_RE_MAXSTAGE;
  const int lane = threadIdx.x & 31;
  double *sh_e = sh, *sh_s = sh + (size_t) nelem * MS;
  const double C1 = (RH_HPLANCK/(2.0*RH_PI*RH_M_ELECTRON)) * (RH_HPLANCK/RH_KBOLTZMANN);
  double ne_old, ne = ne_in;
  if (fromscratch) {
    const double Uk = uk_zero ? 0.0 : pf_interp(npf, Tpf, pf + (size_t) ((int) elems[RHB200_RE_PFROW]) * npf, T);
    const double PhiH = 0.5 * rhm::rh_pow(C1/T, 1.5) * rhm::rh_exp(Uk + elems[RHB200_RE_IONPOT0]/(RH_KBOLTZMANN*T));
    ne_old = (sqrt(1.0 + 4.0*nHtot*PhiH) - 1.0) / (2.0*PhiH);
    ne = ne_old;
  } else ne_old = ne_in;
  double fjk[MS], dfjk[MS];
  for (int niter = 0; niter < N_MAX_ELECTRON_ITERATIONS; niter++) {
    for (int n = lane; n < nelem; n += 32) {
      const double *e = elems + (size_t) n * RHB200_RE_NFIELD;
      const int nst = (int) e[RHB200_RE_NSTAGE], row = (int) e[RHB200_RE_PFROW];
      const double CT_ne = 2.0 * rhm::rh_pow(C1/T, -1.5) / ne_old;
      double sum1 = 1.0, sum2 = 0.0;
      fjk[0] = 1.0; dfjk[0] = 0.0;
      double Uk = pf_interp(npf, Tpf, pf + (size_t) row * npf, T);
      for (int j = 1; j < nst; j++) {
        const double Ukp1 = pf_interp(npf, Tpf, pf + (size_t) (row + j) * npf, T);
        fjk[j]  = fjk[j-1] * CT_ne * rhm::rh_exp(Ukp1 - Uk - e[RHB200_RE_IONPOT0 + j-1]/(RH_KBOLTZMANN*T));
        dfjk[j] = -j * fjk[j] / ne_old;
        sum1 += fjk[j];
        sum2 += dfjk[j];
        Uk = Ukp1;
      }
      for (int j = 0; j < nst; j++) {
        fjk[j] /= sum1;
        dfjk[j] = (dfjk[j] - fjk[j] * sum2) / sum1;
      }
      if (n == 0) {
        const double PhiHmin = 0.25*rhm::rh_pow(C1/T, 1.5) * rhm::rh_exp(0.754 * RH_EV / (RH_KBOLTZMANN * T));
        sh_e[0] = ne_old * fjk[0] * PhiHmin;
        sh_s[0] = (fjk[0] + ne_old * dfjk[0]) * PhiHmin;
      }
      for (int j = 1; j < nst; j++) {
        const double akj = e[RHB200_RE_ABUND] * j;
        sh_e[(size_t) n * MS + j] = akj * fjk[j];
        sh_s[(size_t) n * MS + j] = akj * dfjk[j];
      }
    }
    __syncwarp();
    double dne = 0.0;
    if (lane == 0) {
      double error = ne_old / nHtot, sum = 0.0;
      for (int n = 0; n < nelem; n++) {
        const int nst = (int) elems[(size_t) n * RHB200_RE_NFIELD + RHB200_RE_NSTAGE];
        if (n == 0) { error += sh_e[0]; sum -= sh_s[0]; }
        for (int j = 1; j < nst; j++) { error -= sh_e[(size_t) n * MS + j]; sum += sh_s[(size_t) n * MS + j]; }
      }
      ne = ne_old - nHtot * error / (1.0 - nHtot * sum);
      dne = fabs((ne - ne_old)/ne_old);
    }
    ne = __shfl_sync(0xffffffffu, ne, 0);
    dne = __shfl_sync(0xffffffffu, dne, 0);
    __syncwarp();
    ne_old = ne;
    if (dne <= MAX_ELECTRON_ERROR) break;
  }
  return ne;
}

__global__ void __launch_bounds__(128)
solve_ne_kernel(size_t n, int nelem, int npf, const double *__restrict__ elems, const double *__restrict__ pf,
                const double *__restrict__ Tpf, const double *__restrict__ T, const double *__restrict__ nHtot,
                double *__restrict__ ne, int fromscratch, int uk_zero)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ne[t] = solve_ne_point(nelem, npf, elems, pf, Tpf, T[t], nHtot[t], ne[t], fromscratch, uk_zero);
}

// ---- hydrostatic equilibrium, hse() of rhf1d/pyrh_hse.c:67-400, for a batch of columns.  The reference walks the
//      layers top-down and iterates each to convergence; here every step of that walk is a kernel over the columns.
struct HseCols {
  int ncol, ndep, atm_scale;
  double wght_per_H, total_abund, gravity, LOG10;
  double *scale /* tau_ref or height [m] */, *T, *ne, *nHtot, *rho, *pg, *opac;   // [ncol][ndep]
  int *done, *iter, *nactive;
};

// start of layer k (pyrh_hse.c:212-216, 281-296)
__global__ void hse_layer_init_kernel(HseCols H, int k)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= H.ncol) return;
  const size_t o = (size_t) col * H.ndep + k;
  if (k == 0) {
    H.ne[o] = 0;
    H.nHtot[o] = (H.pg[o]/RH_KBOLTZMANN/H.T[o] - H.ne[o])/H.total_abund;
  } else {
    double deltaP;
    if (H.atm_scale == 2) deltaP = H.gravity * H.rho[o-1] * (H.scale[o-1] - H.scale[o]);
    else                  deltaP = H.gravity * H.rho[o-1]/H.opac[o-1] * (H.scale[o] - H.scale[o-1]);
    H.pg[o] = H.pg[o-1] + deltaP;
    H.nHtot[o] = (H.pg[o]/RH_KBOLTZMANN/H.T[o] - H.ne[o-1])/H.total_abund;
  }
  H.done[col] = 0; H.iter[col] = 0;
}

// first half of one iteration of layer k: density, electron density (get_ne from scratch), the one-depth atmosphere
// rows the LTE-population / chemistry / continuum kernels read (pyrh_hse.c:227-234, 299-306)
__global__ void hse_pre_kernel(HseCols H, int k, int nelem, int npf, const double *__restrict__ elems,
                               const double *__restrict__ pf, const double *__restrict__ Tpf, double *__restrict__ atL)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= H.ncol || H.done[col]) return;
  const size_t o = (size_t) col * H.ndep + k;
  H.rho[o] = (RH_AMU * H.wght_per_H) * H.nHtot[o];
  const double ne = solve_ne_point(nelem, npf, elems, pf, Tpf, H.T[o], H.nHtot[o], 0.0, 1, 1);
  H.ne[o] = ne;
  double *a = atL + (size_t) col * RHB200_AT_NFIELD;
  for (int f = 0; f < RHB200_AT_NFIELD; f++) a[f] = 0.0;
  a[RHB200_AT_T] = H.T[o]; a[RHB200_AT_NE] = ne; a[RHB200_AT_NHTOT] = H.nHtot[o];
}

// the same with one warp per column (small batches: see solve_ne_point_coop)
__global__ void __launch_bounds__(32)
hse_pre_coop_kernel(HseCols H, int k, int nelem, int npf, const double *__restrict__ elems,
                    const double *__restrict__ pf, const double *__restrict__ Tpf, double *__restrict__ atL)
{
  extern __shared__ double hse_sh[];
  const int col = blockIdx.x;
  if (col >= H.ncol || H.done[col]) return;
  const size_t o = (size_t) col * H.ndep + k;
  const double ne = solve_ne_point_coop(nelem, npf, elems, pf, Tpf, H.T[o], H.nHtot[o], 0.0, 1, 1, hse_sh);
  if (threadIdx.x == 0) {
    H.rho[o] = (RH_AMU * H.wght_per_H) * H.nHtot[o];
    H.ne[o] = ne;
    double *a = atL + (size_t) col * RHB200_AT_NFIELD;
    for (int f = 0; f < RHB200_AT_NFIELD; f++) a[f] = 0.0;
    a[RHB200_AT_T] = H.T[o]; a[RHB200_AT_NE] = ne; a[RHB200_AT_NHTOT] = H.nHtot[o];
  }
}

// second half: pressure integration and the new total hydrogen density (pyrh_hse.c:236-254, 308-372)
__global__ void hse_post_kernel(HseCols H, int k, const double *__restrict__ chi_layer)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= H.ncol || H.done[col]) return;
  const size_t o = (size_t) col * H.ndep + k;
  H.opac[o] = chi_layer[col];
  if (k > 0) {
    if (H.atm_scale == 0) {
      const double dlogtau = rhm::rh_log10(H.scale[o]) - rhm::rh_log10(H.scale[o-1]);
      const double beta1 = H.rho[o-1]/H.opac[o-1] * H.scale[o-1];
      const double beta2 = H.rho[o]/H.opac[o] * H.scale[o];
      H.pg[o] = H.pg[o-1] + H.LOG10 * H.gravity * (beta2 + beta1)/2 * dlogtau;
    } else {
      const double deltaP = H.gravity * (H.scale[o-1] - H.scale[o]) * sqrt(H.rho[o]*H.rho[o-1]);
      H.pg[o] = H.pg[o-1] + deltaP;
    }
  }
  const double nHtot_old = H.nHtot[o];
  H.nHtot[o] = (H.pg[o]/RH_KBOLTZMANN/H.T[o] - H.ne[o]) / H.total_abund;
  H.iter[col] += 1;
  const double eta = fabs((H.nHtot[o] - nHtot_old)/H.nHtot[o]);
  if (eta <= 1e-2 || H.iter[col] >= 50) { H.done[col] = 1; atomicSub(H.nactive, 1); }       // NMAX_HSE_ITER, pyrh_hse.c:63
}

}  // namespace

// pyrh.hse (pyrh.pyx:427-489) for a batch of columns: gas pressure, electron and hydrogen densities of an atmosphere
// in hydrostatic equilibrium from its temperature run and the pressure at the top.
extern "C" int rhb200_hse_batch(rhb200_ctx *c, int ncol, int ndep, int atm_scale, const double *scale, const double *T,
                                const double *pg_top, double wght_per_H, double total_abund, double gravity,
                                double *ne, double *nHtot, double *rho, double *pg)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  ElementTable *t = (ElementTable *) c->elements;
  if (!t) { rhb200_set_error("rhb200_set_elements() has not been called"); return RHB200_ESTATE; }
  if (!c->cont || !rh_continuum_has_chemistry(c)) { rhb200_set_error("rhb200_set_continuum() / rhb200_set_chemistry() have not been called"); return RHB200_ESTATE; }
  if (rh_continuum_nlambda(c) != 1 || c->wav.nlambda != 1) { rhb200_set_error("the HSE solver works at the reference wavelength alone: set the one-wavelength grid {500 nm} (pyrh_hse.c:197-200)"); return RHB200_ESTATE; }
  if (atm_scale != 0 && atm_scale != 2) { rhb200_set_error("hse(): only the tau500 (0) and height (2) scales are integrated by the reference (pyrh_hse.c:284-289)"); return RHB200_EUNSUPPORTED; }
  if (ncol < 0 || ndep < 2 || !scale || !T || !pg_top || !ne || !nHtot || !rho || !pg) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  if (ncol == 0) return RHB200_OK;
  const size_t n = (size_t) ncol * ndep;
  const int natom = rh_continuum_natom(c), nlev = rh_continuum_nlev(c);
  double *d = nullptr; int *di = nullptr;
  const size_t layer_doubles = (size_t) ncol * (RHB200_AT_NFIELD + (natom + 4) + nlev + 8 + 2);
  RH_CUDA(cudaMalloc((void **) &d, (7 * n + layer_doubles) * sizeof(double)));
  if (cudaMalloc((void **) &di, (2 * (size_t) ncol + 1) * sizeof(int)) != cudaSuccess) { cudaFree(d); rhb200_set_error("cudaMalloc failed"); return RHB200_ENOMEM; }
  HseCols H{ncol, ndep, atm_scale, wght_per_H, total_abund, gravity, log(10.0),
            d, d + n, d + 2*n, d + 3*n, d + 4*n, d + 5*n, d + 6*n, di, di + ncol, di + 2*(size_t) ncol};
  double *atL = d + 7*n, *chemL = atL + (size_t) ncol * RHB200_AT_NFIELD, *popsL = chemL + (size_t) ncol * (natom + 4),
         *tprepL = popsL + (size_t) ncol * nlev, *chiL = tprepL + (size_t) ncol * 8, *etaL = chiL + ncol;
  int rc = RHB200_OK;
  cudaError_t e = cudaSuccess;
  auto fail = [&](const char *what) { rhb200_set_error("rhb200_hse_batch: %s: %s", what, cudaGetErrorString(e)); rc = RHB200_ECUDA; };
  std::vector<double> h(n);
  for (int col = 0; col < ncol && rc == RHB200_OK; col++)                 // pyrh_hse.c:150-165: tau = POW10(scale), height = km -> m
    for (int k = 0; k < ndep; k++) {
      const double s = scale[(size_t) col * ndep + k];
      h[(size_t) col * ndep + k] = atm_scale == 0 ? exp(2.30258509299404568402 * (s)) : s * 1.0E+03;
    }
  if ((e = cudaMemcpy(H.scale, h.data(), n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(H.T, T, n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemset(H.ne, 0, 5 * n * sizeof(double))) != cudaSuccess) fail("upload");
  if (rc == RHB200_OK) {
    std::fill(h.begin(), h.end(), 0.0);
    for (int col = 0; col < ncol; col++) h[(size_t) col * ndep] = pg_top[col];
    if ((e = cudaMemcpy(H.pg, h.data(), n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) fail("upload");
  }
  rh_continuum_set_hse_mode(c, 1);
  static int coop_max = -1;
  if (coop_max < 0) { const char *e = getenv("RHB200_HSE_COOP_MAX"); coop_max = e ? atoi(e) : 65536; }   // measured: 4096 columns 0.53 -> 0.17 s, 16 384 columns 0.48 -> 0.40 s against the thread-per-column kernel
  const unsigned gb = (unsigned) ((ncol + 63) / 64);
  for (int k = 0; k < ndep && rc == RHB200_OK; k++) {
    hse_layer_init_kernel<<<gb, 64, 0, c->stream>>>(H, k);
    int nactive = ncol;
    if ((e = cudaMemcpyAsync(H.nactive, &nactive, sizeof(int), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) { fail("flag"); break; }
    for (int it = 0; it < 50 && nactive > 0 && rc == RHB200_OK; it++) {
      if (ncol <= coop_max)                        // one warp per column (a single thread needs ~1 ms per Newton solve)
        hse_pre_coop_kernel<<<ncol, 32, 2 * (size_t) t->nelem * RHB200_RE_MAXSTAGE * sizeof(double), c->stream>>>(
            H, k, t->nelem, t->npf, t->elems, t->pf, t->Tpf, atL);
      else
        hse_pre_kernel<<<gb, 64, 0, c->stream>>>(H, k, t->nelem, t->npf, t->elems, t->pf, t->Tpf, atL);
      rc = rh_continuum_chunk(c, ncol, 1, atL, chemL, popsL, tprepL, chiL, etaL, 1);
      if (rc != RHB200_OK) break;
      hse_post_kernel<<<gb, 64, 0, c->stream>>>(H, k, chiL);
      if ((e = cudaMemcpyAsync(&nactive, H.nactive, sizeof(int), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
          (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) fail("iteration");
    }
  }
  rh_continuum_set_hse_mode(c, 0);
  if (rc == RHB200_OK && ((e = cudaMemcpy(ne, H.ne, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess ||
      (e = cudaMemcpy(nHtot, H.nHtot, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess ||
      (e = cudaMemcpy(rho, H.rho, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess ||
      (e = cudaMemcpy(pg, H.pg, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess)) fail("download");
  cudaFree(d); cudaFree(di);
  return rc;
}

void rh_elements_free(rhb200_ctx *c)
{
  if (c->elements) { delete (ElementTable *) c->elements; c->elements = nullptr; }
}

// All elements of the periodic table with their partition functions (atmos.elements[], abundance.c:85-215):
// rows as in rhb200_set_lines (RHB200_RE_*), hydrogen first.
extern "C" int rhb200_set_elements(rhb200_ctx *c, int nelem, const double *elems, int npf_rows, int npf,
                                   const double *pf, const double *Tpf)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  if (nelem < 1 || npf < 2 || npf_rows < 1 || !elems || !pf || !Tpf) { rhb200_set_error("rhb200_set_elements: bad arguments"); return RHB200_EINVAL; }
  for (int e = 0; e < nelem; e++) {
    const double *E = elems + (size_t) e * RHB200_RE_NFIELD;
    const int nst = (int) E[RHB200_RE_NSTAGE], row = (int) E[RHB200_RE_PFROW];
    if (nst < 1 || nst > RHB200_RE_MAXSTAGE || row < 0 || row + nst > npf_rows) { rhb200_set_error("element row %d: Nstage/pf rows out of range", e); return RHB200_EINVAL; }
  }
  rh_elements_free(c);
  ElementTable *t = new ElementTable();
  c->elements = t;
  t->nelem = nelem; t->npf = npf; t->npf_rows = npf_rows;
  RH_CUDA(cudaMalloc((void **) &t->elems, (size_t) nelem * RHB200_RE_NFIELD * sizeof(double)));
  RH_CUDA(cudaMalloc((void **) &t->pf, (size_t) npf_rows * npf * sizeof(double)));
  RH_CUDA(cudaMalloc((void **) &t->Tpf, (size_t) npf * sizeof(double)));
  RH_CUDA(cudaMemcpy(t->elems, elems, (size_t) nelem * RHB200_RE_NFIELD * sizeof(double), cudaMemcpyHostToDevice));
  RH_CUDA(cudaMemcpy(t->pf, pf, (size_t) npf_rows * npf * sizeof(double), cudaMemcpyHostToDevice));
  RH_CUDA(cudaMemcpy(t->Tpf, Tpf, (size_t) npf * sizeof(double), cudaMemcpyHostToDevice));
  return RHB200_OK;
}

// Solve_ne for n independent depth points (any batch of columns flattened): T [K], nHtot [m^-3]; ne [m^-3] is the
// starting guess when fromscratch == 0 and the result.  Hydrogen in LTE (atmos.H_LTE, the only mode pyrh uses).
extern "C" int rhb200_solve_ne_batch(rhb200_ctx *c, size_t n, const double *T, const double *nHtot, double *ne,
                                     int fromscratch)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  ElementTable *t = (ElementTable *) c->elements;
  if (!t) { rhb200_set_error("rhb200_set_elements() has not been called"); return RHB200_ESTATE; }
  if (!T || !nHtot || !ne) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  if (n == 0) return RHB200_OK;
  double *d = nullptr;
  RH_CUDA(cudaMalloc((void **) &d, 3 * n * sizeof(double)));
  int rc = RHB200_OK;
  cudaError_t e;
  if ((e = cudaMemcpy(d, T, n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d + n, nHtot, n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d + 2*n, ne, n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) rc = RHB200_ECUDA;
  if (rc == RHB200_OK) {
    ScopedKernelTimer tm(c, RHB200_K_PREP);
    solve_ne_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(n, t->nelem, t->npf, t->elems, t->pf, t->Tpf,
                                                                        d, d + n, d + 2*n, fromscratch, 0);
  }
  if (rc == RHB200_OK && ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(c->stream)) != cudaSuccess ||
      (e = cudaMemcpy(ne, d + 2*n, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess)) rc = RHB200_ECUDA;
  if (rc != RHB200_OK) rhb200_set_error("rhb200_solve_ne_batch: %s", cudaGetErrorString(e));
  cudaFree(d);
  return rc;
}
