// rhb200_hse.cu -- electron density from the LTE ionisation of all elements (Solve_ne, rh/solvene.c:55-140, with
// getfjk :145-200 and getKuruczpf :205-215) for pyrh.get_ne_from_nH (rhf1d/pyrh_hse.c:555-677), and the
// layer-by-layer variant get_ne() of the hydrostatic-equilibrium solver (rhf1d/pyrh_background.c).
#include "rhb200_common.cuh"
#include "rhb200_math.cuh"

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

__global__ void __launch_bounds__(128)
solve_ne_kernel(size_t n, int nelem, int npf, const double *__restrict__ elems, const double *__restrict__ pf,
                const double *__restrict__ Tpf, const double *__restrict__ T, const double *__restrict__ nHtot,
                double *__restrict__ ne, int fromscratch, int uk_zero)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ne[t] = solve_ne_point(nelem, npf, elems, pf, Tpf, T[t], nHtot[t], ne[t], fromscratch, uk_zero);
}

}  // namespace

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
