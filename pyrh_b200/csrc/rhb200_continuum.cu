// rhb200_continuum.cu -- the angle-independent background continuum of Background() on the device.
//
// Reference: Background            rh/background.c:329-466   (order of the sums, Planck, fudge = off)
//            Thomson               rh/thomson.c:31-44
//            Hminus_bf, Hminus_ff  rh/hydrogen.c:310-365, 367-549   (tables: Geltman 1962; Stilley & Callaway 1970)
//            Hydrogen_bf/_ff, Gaunt_bf/_ff   rh/hydrogen.c:151-305
//            H2plus_ff, H2minus_ff, Rayleigh_H2   rh/hydrogen.c:621-995 (Bates 1952; Bell 1980; Victor & Dalgarno 1969)
//            Rayleigh              rh/rayleigh.c:33-100
//            OH_bf_opac, CH_bf_opac   rh/ohchbf.c:53-691      (Kurucz, van Dishoeck & Tarafdar 1987)
//            Metal_bf              rh/metal.c:71-170
//            splineCoef/splineEval rh/spline.c:33-97; Linear rh/linear.c:22-58; Locate/Hunt rh/hunt.c; bilinear hydrogen.c:998
//
// Split: everything that depends on the wavelength only (spline and table look-ups in lambda, Gaunt
// factors, Rayleigh cross-sections, which bound-free edges are open) is evaluated ONCE per wavelength grid
// on the host with the reference's expressions and the same libm; the device evaluates the per-depth part
// (Boltzmann/stimulated-emission factors with the glibc-exact exp, populations, the T-interpolations) for
// every (column, wavelength, depth) and sums the contributions in Background()'s order.
// The published cross-section tables are inputs (the RH host owns them), see INTEGRATION.md.
#include <cmath>
#include <vector>
#include "rhb200_common.cuh"
#include "rhb200_math.cuh"
#include "rhb200_delo.cuh"
#include "rhb200_lu.cuh"

#ifndef RH_CHECK
#define RH_CHECK(expr) do { int rc__ = (expr); if (rc__ != RHB200_OK) return rc__; } while (0)
#endif

namespace {

#define RH_E_RYDBERG  2.1798741E-18
#define RH_EV         1.60217733E-19
#define RH_THETA0     5.03974756E+03
#define RH_MEGABARN_TO_M2 1.0E-22
#define RH_LG10       2.30258509299404568402
#define RH_CM_TO_M    1.0E-02
#define RH_Q_ELECTRON 1.60217733E-19
#define RH_EPSILON_0  8.854187817E-12
#define SQ(x)   ((x)*(x))
#define CUBE(x) ((x)*(x)*(x))

// ---- host: hunt.c Locate (= Hunt's result for in-range values of a strictly monotonic table)
int locate(int n, const double *a, double v)
{
  const bool ascend = a[n-1] > a[0];
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (hi + lo) >> 1;
    if (ascend ? (v >= a[mid]) : (v <= a[mid])) lo = mid; else hi = mid;
  }
  return lo;
}

// spline.c:33-97
struct Spline {
  std::vector<double> M;
  const double *x, *y; int N; bool ascend; double xmin, xmax;
  void coef(int n, const double *xt, const double *yt) {
    N = n; x = xt; y = yt;
    ascend = x[1] > x[0];
    xmin = ascend ? x[0] : x[N-1];
    xmax = ascend ? x[N-1] : x[0];
    std::vector<double> q(N), u(N);
    M.assign(N, 0.0);
    double hj = x[1] - x[0], D = (y[1] - y[0]) / hj;
    q[0] = u[0] = 0.0;
    for (int j = 1; j < N-1; j++) {
      const double hj1 = x[j+1] - x[j];
      const double mu = hj / (hj + hj1);
      const double D1 = (y[j+1] - y[j]) / hj1;
      const double p = mu*q[j-1] + 2;
      q[j] = (mu - 1) / p;
      u[j] = ((D1 - D) * 6/(hj + hj1) - mu*u[j-1]) / p;
      hj = hj1; D = D1;
    }
    M[N-1] = 0.0;
    for (int j = N-2; j >= 0; j--) M[j] = q[j]*M[j+1] + u[j];
  }
  double eval(double xv) const {
    if (xv <= xmin) return ascend ? y[0] : y[N-1];
    if (xv >= xmax) return ascend ? y[N-1] : y[0];
    const int j = locate(N, x, xv);
    const double hj = x[j+1] - x[j], fx = (xv - x[j]) / hj, fx1 = 1 - fx;
    return fx1*y[j] + fx*y[j+1] + (fx1*(SQ(fx1) - 1) * M[j] + fx*(SQ(fx) - 1) * M[j+1]) * SQ(hj)/6.0;
  }
};

double linear_host(int n, const double *xt, const double *yt, double x)      // linear.c:22-58
{
  const bool ascend = xt[1] > xt[0];
  const double xmin = ascend ? xt[0] : xt[n-1], xmax = ascend ? xt[n-1] : xt[0];
  if (x <= xmin) return ascend ? yt[0] : yt[n-1];
  if (x >= xmax) return ascend ? yt[n-1] : yt[0];
  const int j = locate(n, xt, x);
  const double fx = (xt[j+1] - x) / (xt[j+1] - xt[j]);
  return fx*yt[j] + (1 - fx)*yt[j+1];
}

double gaunt_bf(double lambda, double n_eff, int charge)                     // hydrogen.c:269-282
{
  const double x = ((RH_HPLANCK*RH_CLIGHT)/(lambda * RH_NM_TO_M)) / (RH_E_RYDBERG * SQ(charge));
  const double x3 = pow(x, 0.33333333);
  const double nsqx = 1.0 / (SQ(n_eff) * x);
  return 1.0 + 0.1728*x3 * (1.0 - 2.0*nsqx) - 0.0496*SQ(x3) * (1.0 - (1.0 - nsqx)*0.66666667*nsqx);
}

#ifndef CONT_TL        // -DCONT_TL=n through RHB200_NVCC_EXTRA; measured per 2048-column call: 2 -> 16.7 ms, 3 -> 15.9, 4 -> 15.5
#define CONT_TL 4      // wavelengths per thread of the tiled kernel
#endif
#define CONT_MAXBF 96  // open bound-free edges per family staged in shared memory
// per-wavelength coefficient record (doubles), then the lists of open bound-free edges
enum { WC_FLAGS = 0, WC_LAMBDA, WC_HCKLA_B, WC_TWOHNU3_B, WC_HCKLA_A, WC_TWOHNU3_A, WC_ALPHA_HMBF, WC_LI_HMFF,
       WC_E_OH, WC_E_CH, WC_NU3, WC_CY, WC_GA1, WC_GA2, WC_SIG_RAY_H, WC_SIG_RAY_HE, WC_LI_H2P, WC_SIG_RH2,
       WC_LI_H2M, WC_HBF_FIRST, WC_HBF_COUNT, WC_MBF_FIRST, WC_MBF_COUNT,
       WC_FUDGE_HMIN, WC_FUDGE_SCAT, WC_FUDGE_METAL /* 1.0 without fudge: exact */, WC_NFIELD = 28 };
enum { F_HMBF = 1, F_HMFF = 2, F_OH = 4, F_CH = 8, F_RAY_H = 16, F_RAY_HE = 32, F_H2P = 64, F_RH2 = 128, F_H2M = 256 };

struct DevModel {       // device copies
  double *wc = nullptr, *bfl = nullptr;        // [nlambda][WC_NFIELD]; bound-free list entries {i, j, alpha}
  int *tile_uniform = nullptr;                 // [ceil(nlambda/CONT_TL)]: same open edges for every wavelength of the tile
  double *hmff_kappa = nullptr, *h2m_kappa = nullptr, *h2p_kappa = nullptr, *oh_cross = nullptr, *ch_cross = nullptr;
  double *hmff_theta = nullptr, *h2m_theta = nullptr, *h2p_temp = nullptr, *oh_T = nullptr, *ch_T = nullptr;
  int n_hmff_lambda, n_hmff_theta, n_h2m_lambda, n_h2m_theta, n_h2p_lambda, n_h2p_temp, n_oh_T, n_oh_E, n_ch_T, n_ch_E;
  int nlev, nlev_H, lev0_He, H_active, solve_NLTE;
  int hse_mode = 0;      // pyrh_Background() of the HSE solver: no Metal_bf, scattering added (rhf1d/pyrh_background.c:144-430)
  double sigma_T, sigma_ff;
};

// ---- device
__device__ __forceinline__ double bilinear_d(int Ncol, int Nrow, const double *__restrict__ f, double x, double y)
{                                                                          // hydrogen.c:998-1020
  const int i = (int) x; const double fx = x - i;
  const int i1 = (i == Ncol-1) ? i : i + 1;
  const int j = (int) y; const double fy = y - j;
  const int j1 = (j == Nrow-1) ? j : j + 1;
  return (1.0 - fx)*(1.0 - fy) * f[j*Ncol+i] + fx*(1.0 - fy) * f[j*Ncol+i1] +
         (1.0 - fx)*fy * f[j1*Ncol+i] + fx*fy * f[j1*Ncol+i1];
}

__device__ __forceinline__ int locate_d(int n, const double *__restrict__ a, double v)   // ascending tables
{
  int lo = 0, hi = n;
  while (hi - lo > 1) { const int mid = (hi + lo) >> 1; if (v >= a[mid]) lo = mid; else hi = mid; }
  return lo;
}

// fractional index used by Hminus_ff / H2minus_ff (theta) and H2plus_ff (T): hydrogen.c:480-494
__device__ __forceinline__ double frac_index(int n, const double *__restrict__ tab, double v)
{
  if (v <= tab[0]) return 0;
  if (v >= tab[n-1]) return n - 1;
  const int idx = locate_d(n, tab, v);
  return (double) idx + (v - tab[idx]) / (tab[idx+1] - tab[idx]);
}

// T-only quantities per (column, depth): tprep[col][TP_NFIELD][ndep]
enum { TP_TH_HMFF = 0, TP_TH_H2M, TP_T_H2P, TP_T_OH, TP_T_CH, TP_NFIELD };
__global__ void __launch_bounds__(128)
cont_prep_kernel(int ncol, int ndep, DevModel M, const double *__restrict__ T, size_t Tstride, double *__restrict__ tprep)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t % ndep);
  const double Tk = T[(size_t) col * Tstride + k], theta = RH_THETA0 / Tk;
  double *o = tprep + (size_t) col * TP_NFIELD * ndep + k;
  o[(size_t) TP_TH_HMFF*ndep] = frac_index(M.n_hmff_theta, M.hmff_theta, theta);
  o[(size_t) TP_TH_H2M*ndep]  = M.h2m_theta ? frac_index(M.n_h2m_theta, M.h2m_theta, theta) : 0.0;
  o[(size_t) TP_T_H2P*ndep]   = frac_index(M.n_h2p_temp, M.h2p_temp, Tk);
  // OH / CH (ohchbf.c:360-366): outside the tabulated temperatures the contribution is zero -> index -1
  double ti = -1.0;
  if (M.oh_T && !(Tk < M.oh_T[0] || Tk > M.oh_T[M.n_oh_T-1])) {
    const int i2 = locate_d(M.n_oh_T, M.oh_T, Tk);
    ti = (double) i2 + (Tk - M.oh_T[i2]) / (M.oh_T[i2+1] - M.oh_T[i2]);
  }
  o[(size_t) TP_T_OH*ndep] = ti;
  ti = -1.0;
  if (M.ch_T && !(Tk < M.ch_T[0] || Tk > M.ch_T[M.n_ch_T-1])) {
    const int i2 = locate_d(M.n_ch_T, M.ch_T, Tk);
    ti = (double) i2 + (Tk - M.ch_T[i2]) / (M.ch_T[i2+1] - M.ch_T[i2]);
  }
  o[(size_t) TP_T_CH*ndep] = ti;
}

// one thread per (column, wavelength, depth); sums in Background()'s order (background.c:343-465)
__global__ void __launch_bounds__(128)
continuum_kernel(int ncol, int nlambda, int ndep, DevModel M,
                 const double *__restrict__ T, const double *__restrict__ ne, size_t astride,
                 const double *__restrict__ nHmin, const double *__restrict__ nH2, const double *__restrict__ nOH,
                 const double *__restrict__ nCH, size_t cstride,
                 const double *__restrict__ pn, const double *__restrict__ ps, const double *__restrict__ tprep,
                 double *__restrict__ chi_ai, double *__restrict__ eta_ai, double *__restrict__ sca_ai,
                 double *__restrict__ contrib /* optional [ray][13][2][ndep]: each contribution on its own */)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nlambda * ndep) return;
  const size_t r = t / ndep;
  const int k = (int) (t - r * ndep);
  const int col = (int) (r / nlambda), l = (int) (r - (size_t) col * nlambda);
  const size_t ak = (size_t) col * astride + k;      // T, ne: [col][astride]
  const size_t ck = (size_t) col * cstride + k;      // nHmin, nH2, nOH, nCH: [col][cstride]
#define RH_DBG(id, a, b) do { if (contrib) { contrib[((r*13 + (id))*2 + 0)*ndep + k] = (a); contrib[((r*13 + (id))*2 + 1)*ndep + k] = (b); } } while (0)
  const double *W = M.wc + (size_t) l * WC_NFIELD;
  const int flags = (int) W[WC_FLAGS];
  const double lambda = W[WC_LAMBDA];
  const double Tk = T[ak], nek = ne[ak];
  const double *n_ = pn + (size_t) col * M.nlev * ndep + k, *s_ = ps + (size_t) col * M.nlev * ndep + k;
  const double *tp = tprep + (size_t) col * TP_NFIELD * ndep + k;
  const double Bnu = rhd::planck(Tk, lambda);                               // background.c:334
  const double nH0 = n_[0], np = n_[(size_t) (M.nlev_H-1) * ndep];

  double chi_a = 0.0, eta_a = 0.0, sca_a = nek * M.sigma_T;                 // Thomson, thomson.c:42
  RH_DBG(0, sca_a, 0.0);
  const double hc_kla_B = W[WC_HCKLA_B], twohnu3_B = W[WC_TWOHNU3_B];
  const double stimB = rhm::rh_exp(-hc_kla_B/Tk);
  if (flags & F_HMBF) {                                                     // hydrogen.c:357-361
    const double alpha_bf = W[WC_ALPHA_HMBF];
    chi_a += nHmin[ck] * (1.0 - stimB) * alpha_bf;
    eta_a += nHmin[ck] * twohnu3_B * stimB * alpha_bf;
    RH_DBG(1, nHmin[ck] * (1.0 - stimB) * alpha_bf, nHmin[ck] * twohnu3_B * stimB * alpha_bf);
  }
  const double pe = nek * RH_KBOLTZMANN * Tk;
  if (flags & F_HMFF) {                                                     // hydrogen.c:541-546
    const double kappa = bilinear_d(M.n_hmff_theta, M.n_hmff_lambda, M.hmff_kappa, tp[(size_t) TP_TH_HMFF*ndep], W[WC_LI_HMFF]);
    const double chi = (nH0 * 1.0E-29) * pe * kappa;
    chi_a += chi; eta_a += chi * Bnu;
    RH_DBG(2, chi, 0.0);
  }
  chi_a *= W[WC_FUDGE_HMIN]; eta_a *= W[WC_FUDGE_HMIN];                      // background.c:364-371 (1.0 without fudge)
  if (flags & F_OH) {                                                       // ohchbf.c:372-388
    const double ti = tp[(size_t) TP_T_OH*ndep];
    double chi = 0.0, eta = 0.0;
    if (ti >= 0.0) {
      const double kappa = rhm::rh_exp(RH_LG10 * bilinear_d(M.n_oh_T, M.n_oh_E, M.oh_cross, ti, W[WC_E_OH])) * SQ(RH_CM_TO_M);
      chi = nOH[ck] * (1.0 - stimB) * kappa;
      eta = nOH[ck] * twohnu3_B * stimB * kappa;
    }
    chi_a += chi; eta_a += eta;
    RH_DBG(3, chi, eta);
  }
  if (flags & F_CH) {
    const double ti = tp[(size_t) TP_T_CH*ndep];
    double chi = 0.0, eta = 0.0;
    if (ti >= 0.0) {
      const double kappa = rhm::rh_exp(RH_LG10 * bilinear_d(M.n_ch_T, M.n_ch_E, M.ch_cross, ti, W[WC_E_CH])) * SQ(RH_CM_TO_M);
      chi = nCH[ck] * (1.0 - stimB) * kappa;
      eta = nCH[ck] * twohnu3_B * stimB * kappa;
    }
    chi_a += chi; eta_a += eta;
    RH_DBG(4, chi, eta);
  }
  const double hc_kla_A = W[WC_HCKLA_A], twohnu3_A = W[WC_TWOHNU3_A];
  const double explaA = rhm::rh_exp(-hc_kla_A/Tk);
  {                                                                         // Hydrogen_bf, hydrogen.c:189-224
    const int first = (int) W[WC_HBF_FIRST], cnt = (int) W[WC_HBF_COUNT];
    if (cnt > 0) {
      const double npstar = s_[(size_t) (M.nlev_H-1) * ndep];
      double chi = 0.0, eta = 0.0;
      for (int c = 0; c < cnt; c++) {
        const double *e = M.bfl + (size_t) (first + c) * 3;
        const int i = (int) e[0];
        const double sigma = e[2];
        const double gijk = s_[(size_t) i * ndep]/npstar * explaA;
        chi += sigma * (1.0 - explaA) * n_[(size_t) i * ndep];
        eta += twohnu3_A * gijk * sigma * np;
      }
      chi_a += chi; eta_a += eta;
      RH_DBG(5, chi, eta);
    }
  }
  {                                                                         // Hydrogen_ff, hydrogen.c:255-260
    const double stim = 1.0 - stimB;
    const double y = (W[WC_CY] * Tk) / (RH_HPLANCK*RH_CLIGHT);                // Gaunt_ff, hydrogen.c:296-303
    const double gIII = 1.0 + W[WC_GA1] * (1.0 + y) - W[WC_GA2] * (1.0 + (1.0 + y)*0.33333333*y);
    const double g_ff = (gIII > 1.0) ? gIII : 1.0;
    const double chi = M.sigma_ff / sqrt(Tk) * W[WC_NU3] * nek * np * stim * g_ff;
    chi_a += chi; eta_a += chi * Bnu;
    RH_DBG(6, chi, 0.0);
  }
  if (flags & F_RAY_H)  { sca_a += W[WC_SIG_RAY_H] * nH0; RH_DBG(7, W[WC_SIG_RAY_H] * nH0, 0.0); }   // rayleigh.c:92-93
  if (flags & F_RAY_HE) { sca_a += W[WC_SIG_RAY_HE] * n_[(size_t) M.lev0_He * ndep]; RH_DBG(8, W[WC_SIG_RAY_HE] * n_[(size_t) M.lev0_He * ndep], 0.0); }
  if (flags & F_H2P) {                                                      // hydrogen.c:929-933
    const double kappa = bilinear_d(M.n_h2p_temp, M.n_h2p_lambda, M.h2p_kappa, tp[(size_t) TP_T_H2P*ndep], W[WC_LI_H2P]);
    const double chi = (nH0 * 1.0E-29) * (np * 1.0E-20) * kappa;
    chi_a += chi; eta_a += chi * Bnu;
    RH_DBG(9, chi, 0.0);
  }
  if (flags & F_RH2) { sca_a += W[WC_SIG_RH2] * nH2[ck]; RH_DBG(10, W[WC_SIG_RH2] * nH2[ck], 0.0); }   // hydrogen.c:985-986
  if (flags & F_H2M) {                                                      // hydrogen.c:782-790
    double chi = 0.0;
    if (nH2[ck] > 0.0) {
      const double kappa = bilinear_d(M.n_h2m_theta, M.n_h2m_lambda, M.h2m_kappa, tp[(size_t) TP_TH_H2M*ndep], W[WC_LI_H2M]);
      chi = (nH2[ck] * 1.0E-29) * pe * kappa;
    }
    chi_a += chi; eta_a += chi * Bnu;
    RH_DBG(11, chi, 0.0);
  }
  {                                                                         // Metal_bf, metal.c:104-158 (fudge = 1)
    const int first = (int) W[WC_MBF_FIRST], cnt = (int) W[WC_MBF_COUNT];
    double chi = 0.0, eta = 0.0;
    for (int c = 0; c < cnt; c++) {
      const double *e = M.bfl + (size_t) (first + c) * 3;
      const int i = (int) e[0], j = (int) e[1];
      const double alpha_la = e[2];
      const double gijk = s_[(size_t) i * ndep]/s_[(size_t) j * ndep] * explaA;
      chi += alpha_la * (1.0 - explaA) * n_[(size_t) i * ndep];
      eta += twohnu3_A * gijk * alpha_la * n_[(size_t) j * ndep];
    }
    chi_a += chi * W[WC_FUDGE_METAL]; eta_a += eta * W[WC_FUDGE_METAL];       // background.c:447-451
    if (cnt > 0) RH_DBG(12, chi, eta);
  }
  sca_a *= W[WC_FUDGE_SCAT];                                                // background.c:456-464
  if (M.solve_NLTE) chi_a += sca_a;                                         // background.c:462
  chi_ai[t] = chi_a; eta_ai[t] = eta_a;
  if (sca_ai) sca_ai[t] = sca_a;
}

// Tiled variant used by the fused path: one thread per (column, tile of CONT_TL wavelengths, depth).
// Measured on B200 (2048 columns x 301 wavelengths x 70 depths): per-wavelength kernel 12.2 ms; tiles of
// 8 / 4 / 2 wavelengths at 8 / 32 / 48 warps per SM: 8.4 / 3.97 / 4.96 ms -- occupancy beats tile width.  The level
// populations of a bound-free edge are loaded (and their ratio formed) once per tile instead of once per
// wavelength; every wavelength keeps its own accumulators and receives its terms in the same order, so the
// sums are bit-identical to continuum_kernel's.  Tiles whose wavelengths do not share the same open edges fall
// back to the per-wavelength walk.
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
continuum_tile_kernel(int ncol, int nlambda, int ndep, DevModel M,
                      const double *__restrict__ T, const double *__restrict__ ne, size_t astride,
                      const double *__restrict__ nHmin, const double *__restrict__ nH2, const double *__restrict__ nOH,
                      const double *__restrict__ nCH, size_t cstride,
                      const double *__restrict__ pn, const double *__restrict__ ps, const double *__restrict__ tprep,
                      double *__restrict__ chi_ai, double *__restrict__ eta_ai, double *__restrict__ sca_ai /* or NULL */)
{
  // blockIdx.x = wavelength tile, blockIdx.y = chunk of 128 (column, depth) pairs: the tile's coefficient
  // records and cross-sections are staged in shared memory once per block
  const int tl = blockIdx.x;
  const int l0 = tl * CONT_TL, nl = (nlambda - l0 < CONT_TL) ? nlambda - l0 : CONT_TL;
  __shared__ double shW[CONT_TL][WC_NFIELD];
  __shared__ double sh_alpha[2][CONT_MAXBF][CONT_TL];
  __shared__ int sh_ij[2][CONT_MAXBF][2];
  for (int x = threadIdx.x; x < nl * WC_NFIELD; x += blockDim.x) shW[x / WC_NFIELD][x % WC_NFIELD] = M.wc[(size_t) l0 * WC_NFIELD + x];
  __syncthreads();
  const bool uniform = M.tile_uniform[tl] != 0 && (int) shW[0][WC_HBF_COUNT] <= CONT_MAXBF && (int) shW[0][WC_MBF_COUNT] <= CONT_MAXBF;
  if (uniform) {
    for (int fam = 0; fam < 2; fam++) {
      const int FI = fam ? WC_MBF_FIRST : WC_HBF_FIRST, cnt = (int) shW[0][fam ? WC_MBF_COUNT : WC_HBF_COUNT];
      for (int x = threadIdx.x; x < cnt * nl; x += blockDim.x) {
        const int c = x / nl, q = x % nl;
        const double *e = M.bfl + (size_t) ((int) shW[q][FI] + c) * 3;
        sh_alpha[fam][c][q] = e[2];
        if (q == 0) { sh_ij[fam][c][0] = (int) e[0]; sh_ij[fam][c][1] = (int) e[1]; }
      }
    }
  }
  __syncthreads();
  const size_t t = (size_t) blockIdx.y * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t - (size_t) col * ndep);
  const size_t ak = (size_t) col * astride + k, ck = (size_t) col * cstride + k;
  const double Tk = T[ak], nek = ne[ak];
  const double *n_ = pn + (size_t) col * M.nlev * ndep + k, *s_ = ps + (size_t) col * M.nlev * ndep + k;
  const double *tp = tprep + (size_t) col * TP_NFIELD * ndep + k;
  const double nH0 = n_[0], np = n_[(size_t) (M.nlev_H-1) * ndep], npstar = s_[(size_t) (M.nlev_H-1) * ndep];
  const double pe = nek * RH_KBOLTZMANN * Tk;
  const double nHm = nHmin[ck], nH2k = nH2 ? nH2[ck] : 0.0;
  const double th_hmff = tp[(size_t) TP_TH_HMFF*ndep], th_h2m = tp[(size_t) TP_TH_H2M*ndep], t_h2p = tp[(size_t) TP_T_H2P*ndep];
  const double ti_oh = tp[(size_t) TP_T_OH*ndep], ti_ch = tp[(size_t) TP_T_CH*ndep];

  // Per-wavelength state of the thread lives in shared memory ([q][thread]: conflict-free) -- the running sums chi_a /
  // eta_a, the Planck function and the two exponentials.  Only the accumulators of the bound-free loop stay in
  // registers, so nothing spills: the previous version kept 20 doubles per thread live across that loop and wrote a
  // 416-byte stack frame per thread through L2 to DRAM (5x the kernel's algorithmic traffic).
  __shared__ double sh_chi[CONT_TL][128], sh_eta[CONT_TL][128], sh_Bnu[CONT_TL][128], sh_stimB[CONT_TL][128], sh_expA[CONT_TL][128];
  const int tx = threadIdx.x;
#pragma unroll 1
  for (int q = 0; q < nl; q++) {
    const double *W = shW[q];
    const int flags = (int) W[WC_FLAGS];
    double chi_q = 0.0, eta_q = 0.0;
    const double Bnu_q = rhd::planck(Tk, W[WC_LAMBDA]);
    const double stimB_q = rhm::rh_exp(-W[WC_HCKLA_B]/Tk);
    sh_Bnu[q][tx] = Bnu_q; sh_stimB[q][tx] = stimB_q;
    sh_expA[q][tx] = rhm::rh_exp(-W[WC_HCKLA_A]/Tk);
    const double twohnu3_B = W[WC_TWOHNU3_B];
    if (flags & F_HMBF) {
      const double alpha_bf = W[WC_ALPHA_HMBF];
      chi_q += nHm * (1.0 - stimB_q) * alpha_bf;
      eta_q += nHm * twohnu3_B * stimB_q * alpha_bf;
    }
    if (flags & F_HMFF) {
      const double kappa = bilinear_d(M.n_hmff_theta, M.n_hmff_lambda, M.hmff_kappa, th_hmff, W[WC_LI_HMFF]);
      const double chi = (nH0 * 1.0E-29) * pe * kappa;
      chi_q += chi; eta_q += chi * Bnu_q;
    }
    chi_q *= W[WC_FUDGE_HMIN]; eta_q *= W[WC_FUDGE_HMIN];            // background.c:364-371 (1.0 without fudge)
    if (flags & F_OH) {
      double chi = 0.0, eta = 0.0;
      if (ti_oh >= 0.0) {
        const double kappa = rhm::rh_exp(RH_LG10 * bilinear_d(M.n_oh_T, M.n_oh_E, M.oh_cross, ti_oh, W[WC_E_OH])) * SQ(RH_CM_TO_M);
        chi = nOH[ck] * (1.0 - stimB_q) * kappa;
        eta = nOH[ck] * twohnu3_B * stimB_q * kappa;
      }
      chi_q += chi; eta_q += eta;
    }
    if (flags & F_CH) {
      double chi = 0.0, eta = 0.0;
      if (ti_ch >= 0.0) {
        const double kappa = rhm::rh_exp(RH_LG10 * bilinear_d(M.n_ch_T, M.n_ch_E, M.ch_cross, ti_ch, W[WC_E_CH])) * SQ(RH_CM_TO_M);
        chi = nCH[ck] * (1.0 - stimB_q) * kappa;
        eta = nCH[ck] * twohnu3_B * stimB_q * kappa;
      }
      chi_q += chi; eta_q += eta;
    }
    sh_chi[q][tx] = chi_q; sh_eta[q][tx] = eta_q;
  }
  // ---- bound-free families: 0 = Hydrogen_bf (ratio to the proton density), 1 = Metal_bf; between them
  //      Hydrogen_ff, H2plus_ff, H2minus_ff enter in Background()'s order
#pragma unroll 1
  for (int fam = 0; fam < 2; fam++) {
    const int FI = fam ? WC_MBF_FIRST : WC_HBF_FIRST, CI = fam ? WC_MBF_COUNT : WC_HBF_COUNT;
    double chi_f[CONT_TL], eta_f[CONT_TL], explaA[CONT_TL];
#pragma unroll
    for (int q = 0; q < CONT_TL; q++) { chi_f[q] = eta_f[q] = 0.0; explaA[q] = (q < nl) ? sh_expA[q][tx] : 0.0; }
    if (uniform) {
      const int cnt = (int) shW[0][CI];
#pragma unroll 2
      for (int c = 0; c < cnt; c++) {
        const int i = sh_ij[fam][c][0], j = sh_ij[fam][c][1];
        const double n_i = n_[(size_t) i * ndep];
        const double n_up = fam ? n_[(size_t) j * ndep] : np;
        const double ratio = fam ? s_[(size_t) i * ndep]/s_[(size_t) j * ndep] : s_[(size_t) i * ndep]/npstar;
#pragma unroll
        for (int q = 0; q < CONT_TL; q++) {
          if (q < nl) {
            const double a = sh_alpha[fam][c][q];
            const double gijk = ratio * explaA[q];
            chi_f[q] += a * (1.0 - explaA[q]) * n_i;
            eta_f[q] += shW[q][WC_TWOHNU3_A] * gijk * a * n_up;
          }
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < CONT_TL; q++) {                  // static indices: the per-wavelength arrays stay in registers
        if (q >= nl) continue;
        const double *W = shW[q];
        const int first = (int) W[FI], cnt = (int) W[CI];
        double chi = 0.0, eta = 0.0;
        for (int c = 0; c < cnt; c++) {
          const double *e = M.bfl + (size_t) (first + c) * 3;
          const int i = (int) e[0], j = (int) e[1];
          const double gijk = (fam ? s_[(size_t) i * ndep]/s_[(size_t) j * ndep] : s_[(size_t) i * ndep]/npstar) * explaA[q];
          chi += e[2] * (1.0 - explaA[q]) * n_[(size_t) i * ndep];
          eta += W[WC_TWOHNU3_A] * gijk * e[2] * (fam ? n_[(size_t) j * ndep] : np);
        }
        chi_f[q] = chi; eta_f[q] = eta;
      }
    }
#pragma unroll
    for (int q = 0; q < CONT_TL; q++) {
      if (q < nl) {
        const double *W = shW[q];
        const int flags = (int) W[WC_FLAGS];
        double chi_q = sh_chi[q][tx], eta_q = sh_eta[q][tx];
        if (fam == 0) {
          const double Bnu_q = sh_Bnu[q][tx];
          if ((int) W[WC_HBF_COUNT] > 0) { chi_q += chi_f[q]; eta_q += eta_f[q]; }
          {                                                                   // Hydrogen_ff
            const double stim = 1.0 - sh_stimB[q][tx];
            const double y = (W[WC_CY] * Tk) / (RH_HPLANCK*RH_CLIGHT);
            const double gIII = 1.0 + W[WC_GA1] * (1.0 + y) - W[WC_GA2] * (1.0 + (1.0 + y)*0.33333333*y);
            const double g_ff = (gIII > 1.0) ? gIII : 1.0;
            const double chi = M.sigma_ff / sqrt(Tk) * W[WC_NU3] * nek * np * stim * g_ff;
            chi_q += chi; eta_q += chi * Bnu_q;
          }
          if (flags & F_H2P) {
            const double kappa = bilinear_d(M.n_h2p_temp, M.n_h2p_lambda, M.h2p_kappa, t_h2p, W[WC_LI_H2P]);
            const double chi = (nH0 * 1.0E-29) * (np * 1.0E-20) * kappa;
            chi_q += chi; eta_q += chi * Bnu_q;
          }
          if (flags & F_H2M) {
            double chi = 0.0;
            if (nH2k > 0.0) {
              const double kappa = bilinear_d(M.n_h2m_theta, M.n_h2m_lambda, M.h2m_kappa, th_h2m, W[WC_LI_H2M]);
              chi = (nH2k * 1.0E-29) * pe * kappa;
            }
            chi_q += chi; eta_q += chi * Bnu_q;
          }
          sh_chi[q][tx] = chi_q; sh_eta[q][tx] = eta_q;
        } else {
          if (!M.hse_mode) { chi_q += chi_f[q] * W[WC_FUDGE_METAL]; eta_q += eta_f[q] * W[WC_FUDGE_METAL]; }
          double chi_out = chi_q;
          const size_t o = ((size_t) col * nlambda + l0 + q) * ndep + k;
          if (M.solve_NLTE || M.hse_mode || sca_ai) {                                       // background.c:456-464
            double sca = nek * M.sigma_T;
            if (flags & F_RAY_H)  sca += W[WC_SIG_RAY_H] * nH0;
            if (flags & F_RAY_HE) sca += W[WC_SIG_RAY_HE] * n_[(size_t) M.lev0_He * ndep];
            if (flags & F_RH2)    sca += W[WC_SIG_RH2] * nH2k;
            sca *= W[WC_FUDGE_SCAT];
            if (M.solve_NLTE || M.hse_mode) chi_out += sca;                                 // LTE: sca_c stays separate
            if (sca_ai) sca_ai[o] = sca;
          }
          chi_ai[o] = chi_out; eta_ai[o] = eta_q;
        }
      }
    }
  }
}

template <class T> int up(T **d, const T *h, size_t n)
{
  *d = nullptr;
  if (n == 0 || !h) return RHB200_OK;
  RH_CUDA(cudaMalloc((void **) d, n * sizeof(T)));
  RH_CUDA(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return RHB200_OK;
}

struct Holder {          // frees the device copies
  std::vector<void *> p;
  ~Holder() { for (void *q : p) if (q) cudaFree(q); }
  template <class T> int put(T **d, const T *h, size_t n) { int rc = up(d, h, n); if (*d) p.push_back(*d); return rc; }
};

}  // namespace

// builds the per-wavelength coefficients on the host (reference expressions, same libm) and uploads everything
static int build_model(const rhb200_continuum_model *m, int nlambda, const double *lambda, DevModel &D, Holder &H)
{
  const double *lev = m->lev;
  auto lvE = [&](int g) { return lev[5*(size_t) g + 1]; };
  auto lvStage = [&](int g) { return (int) lev[5*(size_t) g + 2]; };
  if (m->do_fudge && (m->n_fudge < 2 || !m->fudge_lambda || !m->fudge)) { rhb200_set_error("do_fudge without a fudge table"); return RHB200_EINVAL; }
  auto fudge_at = [&](int row, double lam) {                 // Linear(), linear.c:22-51 with Locate()
    if (!m->do_fudge) return 1.0;
    const int N = m->n_fudge; const double *x = m->fudge_lambda, *y = m->fudge + (size_t) row * N;
    const bool ascend = x[1] > x[0];
    const double xmin = ascend ? x[0] : x[N-1], xmax = ascend ? x[N-1] : x[0];
    if (lam <= xmin) return ascend ? y[0] : y[N-1];
    if (lam >= xmax) return ascend ? y[N-1] : y[0];
    int lo = 0, hi = N;
    const bool asc2 = x[N-1] > x[0];
    while (hi - lo > 1) { const int mid = (hi + lo) >> 1; if (asc2 ? (lam >= x[mid]) : (lam <= x[mid])) lo = mid; else hi = mid; }
    const double fx = (x[lo+1] - lam) / (x[lo+1] - x[lo]);
    return fx*y[lo] + (1 - fx)*y[lo+1];
  };
  // splines of the tabulated bound-free cross-sections (metal.c:137-140) and of H- bf (hydrogen.c:343)
  std::vector<Spline> sp(m->ncont);
  for (int c = 0; c < m->ncont; c++) {
    const double *b = m->bf + 10*(size_t) c;
    if (b[5] == 0.0 && (int) b[0] != 0)
      sp[c].coef((int) b[7], m->tab_lambda + (int) b[8], m->tab_alpha + (int) b[8]);
  }
  Spline hm;
  hm.coef(m->n_hmbf, m->hmbf_lambda, m->hmbf_alpha);

  const double twohc  = (2.0 * RH_HPLANCK * RH_CLIGHT) / CUBE(RH_NM_TO_M);
  const double hc_k   = (RH_HPLANCK * RH_CLIGHT) / (RH_KBOLTZMANN * RH_NM_TO_M);
  const double sigma0 = 32.0/(3.0*sqrt(3.0)) * SQ(RH_Q_ELECTRON)/(4.0*RH_PI*RH_EPSILON_0) /
                        (RH_M_ELECTRON * RH_CLIGHT) * RH_HPLANCK/(2.0*RH_E_RYDBERG);
  const double C_ray = 2*RH_PI * (RH_Q_ELECTRON/RH_EPSILON_0) * (RH_Q_ELECTRON/RH_M_ELECTRON) / RH_CLIGHT;
  const double sigma_e = 8.0*RH_PI/3.0 * pow(RH_Q_ELECTRON/(sqrt(4.0*RH_PI*RH_EPSILON_0) * (sqrt(RH_M_ELECTRON)*RH_CLIGHT)), 4);
  {
    const double C0 = SQ(RH_Q_ELECTRON)/(4.0*RH_PI*RH_EPSILON_0) / sqrt(RH_M_ELECTRON);
    D.sigma_ff = 4.0/3.0 * sqrt(2.0*RH_PI/(3.0 * RH_KBOLTZMANN)) * CUBE(C0) / (RH_HPLANCK * RH_CLIGHT);
    D.sigma_T = sigma_e;
  }
  std::vector<double> wc((size_t) nlambda * WC_NFIELD, 0.0), bfl;
  for (int l = 0; l < nlambda; l++) {
    const double lam = lambda[l];
    double *W = wc.data() + (size_t) l * WC_NFIELD;
    int flags = 0;
    if (!(lam > 0.0)) { rhb200_set_error("wavelength %d is not positive", l); return RHB200_EINVAL; }
    W[WC_LAMBDA] = lam;
    W[WC_FUDGE_HMIN] = fudge_at(0, lam); W[WC_FUDGE_SCAT] = fudge_at(1, lam); W[WC_FUDGE_METAL] = fudge_at(2, lam);
    W[WC_HCKLA_B]   = (RH_HPLANCK * RH_CLIGHT) / (RH_KBOLTZMANN * RH_NM_TO_M * lam);
    W[WC_TWOHNU3_B] = (2.0 * RH_HPLANCK * RH_CLIGHT) / CUBE(RH_NM_TO_M * lam);
    W[WC_HCKLA_A]   = hc_k / lam;
    W[WC_TWOHNU3_A] = twohc / CUBE(lam);
    // Hminus_bf, hydrogen.c:340-349
    if (!((lam <= m->hmbf_lambda[0]) || (lam >= m->hmbf_lambda[m->n_hmbf-1]))) {
      double a = hm.eval(lam);
      a *= 1.0E-21;
      W[WC_ALPHA_HMBF] = a; flags |= F_HMBF;
    }
    // Hminus_ff, hydrogen.c:474-476, 523-525
    if (lam >= m->hmff_lambda[m->n_hmff_lambda-1]) { rhb200_set_error("Hminus_ff_long (lambda >= %g nm) is not implemented", m->hmff_lambda[m->n_hmff_lambda-1]); return RHB200_EUNSUPPORTED; }
    {
      const int idx = locate(m->n_hmff_lambda, m->hmff_lambda, lam);
      W[WC_LI_HMFF] = (double) idx + (lam - m->hmff_lambda[idx]) / (m->hmff_lambda[idx+1] - m->hmff_lambda[idx]);
      flags |= F_HMFF;
    }
    // OH / CH, ohchbf.c:345-351
    const double Eev = (RH_HPLANCK * RH_CLIGHT) / (lam * RH_NM_TO_M) / RH_EV;
    if (m->has_OH && m->oh_E && !(Eev < m->oh_E[0] || Eev > m->oh_E[m->n_oh_E-1])) {
      const int idx = locate(m->n_oh_E, m->oh_E, Eev);
      W[WC_E_OH] = (double) idx + (Eev - m->oh_E[idx]) / (m->oh_E[idx+1] - m->oh_E[idx]); flags |= F_OH;
    }
    if (m->has_CH && m->ch_E && !(Eev < m->ch_E[0] || Eev > m->ch_E[m->n_ch_E-1])) {
      const int idx = locate(m->n_ch_E, m->ch_E, Eev);
      W[WC_E_CH] = (double) idx + (Eev - m->ch_E[idx]) / (m->ch_E[idx+1] - m->ch_E[idx]); flags |= F_CH;
    }
    // Hydrogen_bf, hydrogen.c:189-206 (H continua are the entries with atom == 0)
    W[WC_HBF_FIRST] = (double) (bfl.size() / 3);
    if (!m->H_active) {
      for (int c = 0; c < m->ncont; c++) {
        const double *b = m->bf + 10*(size_t) c;
        if ((int) b[0] != 0) continue;
        const double lambdaEdge = b[3];
        if (lam <= lambdaEdge && lam >= b[4]) {
          const int i = (int) b[1], j = (int) b[2];
          const double n_eff = sqrt(RH_E_RYDBERG / (lvE(j) - lvE(i)));
          const double g_bf = gaunt_bf(lam, n_eff, lvStage(i) + 1);
          const double sigma = sigma0 * n_eff * g_bf * CUBE(lam/lambdaEdge);
          bfl.push_back(i); bfl.push_back(j); bfl.push_back(sigma);
        }
      }
    }
    W[WC_HBF_COUNT] = (double) (bfl.size() / 3) - W[WC_HBF_FIRST];
    // Hydrogen_ff + Gaunt_ff, hydrogen.c:243-246, 296-300
    W[WC_NU3] = CUBE((lam * RH_NM_TO_M) / RH_CLIGHT);
    {
      const double x = ((RH_HPLANCK * RH_CLIGHT)/(lam * RH_NM_TO_M)) / (RH_E_RYDBERG * SQ(1));
      const double x3 = pow(x, 0.33333333);
      W[WC_CY] = 2.0 * lam * RH_NM_TO_M * RH_KBOLTZMANN;
      W[WC_GA1] = 0.1728*x3;
      W[WC_GA2] = 0.0496*SQ(x3);
    }
    // Rayleigh (H: atom 0, He: atom 1), rayleigh.c:55-93
    for (int atom = 0; atom < 2; atom++) {             // ray[][0]: 0 = hydrogen, 1 = helium
      if (atom == 1 && m->atom_He < 0) continue;
      double lambda_limit = 1.0E6; int nl = 0;
      for (int r = 0; r < m->nray; r++) {
        const double *R = m->ray + 8*(size_t) r;
        if ((int) R[0] != atom) continue;
        nl++;
        const double lambda_red = R[1] * (1.0 + R[2] * m->vmicro_char / RH_CLIGHT);
        lambda_limit = lambda_limit < lambda_red ? lambda_limit : lambda_red;
      }
      if (nl == 0) continue;                       // no line from the ground state: lambda_limit stays LONG_WAVELENGTH
      if (lam > lambda_limit) {
        double fomega = 0.0;
        for (int r = 0; r < m->nray; r++) {
          const double *R = m->ray + 8*(size_t) r;
          if ((int) R[0] != atom) continue;
          const double lambda_red = R[1] * (1.0 + R[2] * m->vmicro_char / RH_CLIGHT);
          if (lam > lambda_red) {
            const double lambda2 = 1.0 / (SQ(lam / R[1]) - 1.0);
            const double f = R[3] * (R[4] / R[5]) * SQ(R[1]*RH_NM_TO_M) / C_ray;
            fomega += f * SQ(lambda2);
          }
        }
        W[atom == 0 ? WC_SIG_RAY_H : WC_SIG_RAY_HE] = sigma_e * fomega;
        flags |= (atom == 0 ? F_RAY_H : F_RAY_HE);
      }
    }
    // H2plus_ff, hydrogen.c:883-884, 917-919
    if (lam < m->h2pff_lambda[m->n_h2pff_lambda-1]) {
      const int idx = locate(m->n_h2pff_lambda, m->h2pff_lambda, lam);
      W[WC_LI_H2P] = idx + (lam - m->h2pff_lambda[idx]) / (m->h2pff_lambda[idx+1] - m->h2pff_lambda[idx]); flags |= F_H2P;
    }
    // Rayleigh_H2, hydrogen.c:970-981
    if (m->has_H2 && lam >= 121.57) {
      double s;
      if (lam <= m->rh2_lambda[m->n_rh2-1]) s = linear_host(m->n_rh2, m->rh2_lambda, m->rh2_sigma, lam);
      else { const double lambda2 = 1.0 / SQ(lam); s = (m->rh2_a[0] + (m->rh2_a[1] + m->rh2_a[2]*lambda2) * lambda2) * SQ(lambda2); }
      s *= RH_MEGABARN_TO_M2;
      W[WC_SIG_RH2] = s; flags |= F_RH2;
    }
    // H2minus_ff, hydrogen.c:724-725, 762-764
    if (m->has_H2 && lam < m->h2mff_lambda[m->n_h2mff_lambda-1]) {
      const int idx = locate(m->n_h2mff_lambda, m->h2mff_lambda, lam);
      W[WC_LI_H2M] = idx + (lam - m->h2mff_lambda[idx]) / (m->h2mff_lambda[idx+1] - m->h2mff_lambda[idx]); flags |= F_H2M;
    }
    // Metal_bf, metal.c:104-141: PASSIVE atoms other than hydrogen
    W[WC_MBF_FIRST] = (double) (bfl.size() / 3);
    for (int c = 0; c < m->ncont; c++) {
      const double *b = m->bf + 10*(size_t) c;
      if ((int) b[0] == 0 || b[9] != 0.0) continue;
      if (lam <= b[3] && lam >= b[4]) {
        const int i = (int) b[1], j = (int) b[2];
        double alpha_la;
        if (b[5] != 0.0) {
          const int Z = lvStage(j);
          const double n_eff = Z*sqrt(RH_E_RYDBERG / (lvE(j) - lvE(i)));
          const double gbf_0 = gaunt_bf(b[3], n_eff, Z);
          alpha_la = b[6] * CUBE(lam/b[3]) * gaunt_bf(lam, n_eff, Z) / gbf_0;
        } else
          alpha_la = sp[c].eval(lam);
        bfl.push_back(i); bfl.push_back(j); bfl.push_back(alpha_la);
      }
    }
    W[WC_MBF_COUNT] = (double) (bfl.size() / 3) - W[WC_MBF_FIRST];
    W[WC_FLAGS] = (double) flags;
  }
  if (bfl.empty()) bfl.assign(3, 0.0);
  {
    const int ntile = (nlambda + CONT_TL - 1) / CONT_TL;
    std::vector<int> uni(ntile, 1);
    for (int tl = 0; tl < ntile; tl++) {
      const double *W0 = wc.data() + (size_t) tl * CONT_TL * WC_NFIELD;
      for (int l = tl * CONT_TL + 1; l < nlambda && l < (tl + 1) * CONT_TL && uni[tl]; l++) {
        const double *W = wc.data() + (size_t) l * WC_NFIELD;
        for (int fam = 0; fam < 2 && uni[tl]; fam++) {
          const int fi = fam ? WC_MBF_FIRST : WC_HBF_FIRST, ci = fam ? WC_MBF_COUNT : WC_HBF_COUNT;
          if (W[ci] != W0[ci]) { uni[tl] = 0; break; }
          for (int c = 0; c < (int) W[ci]; c++)
            if (bfl[3*((size_t) W[fi] + c)] != bfl[3*((size_t) W0[fi] + c)] ||
                bfl[3*((size_t) W[fi] + c) + 1] != bfl[3*((size_t) W0[fi] + c) + 1]) { uni[tl] = 0; break; }
        }
      }
    }
    RH_CHECK(H.put(&D.tile_uniform, uni.data(), uni.size()));
  }
  RH_CHECK(H.put(&D.wc, wc.data(), wc.size()));
  RH_CHECK(H.put(&D.bfl, bfl.data(), bfl.size()));
  RH_CHECK(H.put(&D.hmff_kappa, m->hmff_kappa, (size_t) m->n_hmff_lambda * m->n_hmff_theta));
  RH_CHECK(H.put(&D.hmff_theta, m->hmff_theta, (size_t) m->n_hmff_theta));
  RH_CHECK(H.put(&D.h2m_kappa, m->h2mff_kappa, (size_t) m->n_h2mff_lambda * m->n_h2mff_theta));
  RH_CHECK(H.put(&D.h2m_theta, m->h2mff_theta, (size_t) m->n_h2mff_theta));
  RH_CHECK(H.put(&D.h2p_kappa, m->h2pff_kappa, (size_t) m->n_h2pff_lambda * m->n_h2pff_temp));
  RH_CHECK(H.put(&D.h2p_temp, m->h2pff_temp, (size_t) m->n_h2pff_temp));
  if (m->has_OH) { RH_CHECK(H.put(&D.oh_cross, m->oh_cross, (size_t) m->n_oh_T * m->n_oh_E)); RH_CHECK(H.put(&D.oh_T, m->oh_T, (size_t) m->n_oh_T)); }
  if (m->has_CH) { RH_CHECK(H.put(&D.ch_cross, m->ch_cross, (size_t) m->n_ch_T * m->n_ch_E)); RH_CHECK(H.put(&D.ch_T, m->ch_T, (size_t) m->n_ch_T)); }
  D.n_hmff_lambda = m->n_hmff_lambda; D.n_hmff_theta = m->n_hmff_theta;
  D.n_h2m_lambda = m->n_h2mff_lambda; D.n_h2m_theta = m->n_h2mff_theta;
  D.n_h2p_lambda = m->n_h2pff_lambda; D.n_h2p_temp = m->n_h2pff_temp;
  D.n_oh_T = m->n_oh_T; D.n_oh_E = m->n_oh_E; D.n_ch_T = m->n_ch_T; D.n_ch_E = m->n_ch_E;
  D.nlev = m->nlev; D.nlev_H = m->nlev_H; D.H_active = m->H_active; D.solve_NLTE = m->solve_NLTE;
  D.lev0_He = 0;
  for (int g = 0; g < m->nlev; g++) if ((int) lev[5*(size_t) g] == m->atom_He) { D.lev0_He = g; break; }
  return RHB200_OK;
}

static int check_model(const rhb200_continuum_model *m)
{
  if (!m || m->natom < 1 || m->nlev < 2 || m->nlev_H < 2 || m->ncont < 0 || !m->lev || (m->ncont && !m->bf) ||
      !m->hmbf_lambda || !m->hmbf_alpha || m->n_hmbf < 3 || !m->hmff_lambda || !m->hmff_theta || !m->hmff_kappa ||
      !m->h2pff_lambda || !m->h2pff_temp || !m->h2pff_kappa ||
      (m->has_H2 && (!m->h2mff_lambda || !m->h2mff_theta || !m->h2mff_kappa || !m->rh2_lambda || !m->rh2_sigma || !m->rh2_a)) ||
      (m->has_OH && (!m->oh_T || !m->oh_E || !m->oh_cross)) || (m->has_CH && (!m->ch_T || !m->ch_E || !m->ch_cross))) {
    rhb200_set_error("incomplete continuum model"); return RHB200_EINVAL;
  }
  for (int c = 0; c < m->ncont; c++) {
    const double *b = m->bf + 10*(size_t) c;
    if ((int) b[1] < 0 || (int) b[1] >= m->nlev || (int) b[2] < 0 || (int) b[2] >= m->nlev ||
        (b[5] == 0.0 && ((int) b[7] < 3 || (int) b[8] < 0 || (int) b[8] + (int) b[7] > m->ntab))) {
      rhb200_set_error("continuum %d: level index or table slice out of range", c); return RHB200_EINVAL;
    }
  }
  return RHB200_OK;
}

// device-pointer core shared by the host entry point and the fused LTE path
int rh_continuum_dev(rhb200_ctx *c, const rhb200_continuum_model *m, int nlambda, const double *h_lambda,
                     int ncol, int ndep, const double *d_T, const double *d_ne, const double *d_nHmin,
                     const double *d_nH2, const double *d_nOH, const double *d_nCH, const double *d_n,
                     const double *d_nstar, double *d_chi, double *d_eta, double *d_sca, double *d_contrib)
{
  RH_CHECK(check_model(m));
  DevModel D; Holder H;
  RH_CHECK(build_model(m, nlambda, h_lambda, D, H));
  double *d_tprep = nullptr;
  RH_CUDA(cudaMalloc((void **) &d_tprep, (size_t) ncol * TP_NFIELD * ndep * sizeof(double)));
  H.p.push_back(d_tprep);
  {
    ScopedKernelTimer t(c, RHB200_K_PREP);
    cont_prep_kernel<<<(unsigned) (((size_t) ncol * ndep + 127) / 128), 128, 0, c->stream>>>(ncol, ndep, D, d_T, (size_t) ndep, d_tprep);
  }
  {
    ScopedKernelTimer t(c, RHB200_K_OTHER);
    const size_t n = (size_t) ncol * nlambda * ndep;
    continuum_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(ncol, nlambda, ndep, D, d_T, d_ne, (size_t) ndep,
        d_nHmin, d_nH2, d_nOH, d_nCH, (size_t) ndep, d_n, d_nstar, d_tprep, d_chi, d_eta, d_sca, d_contrib);
  }
  RH_CUDA(cudaGetLastError());
  RH_CUDA(cudaStreamSynchronize(c->stream));      // the Holder frees the tables on return
  return RHB200_OK;
}

// ---- LTEpops (ltepops.c:33-113, Debeye forced off as in the reference) + the per-atom rescaling of
//      ChemicalEquilibrium (chemequil.c:334-343).  One thread per (column, depth); chem [col][natom+4][ndep] =
//      fraction per atom (1 for atoms that are in no molecule), nHmin, nH2, nOH, nCH.
__global__ void __launch_bounds__(128)
ltepops_kernel(int ncol, int ndep, int natom, int nlev, const double *__restrict__ lev /*[nlev][5]*/,
               const int *__restrict__ atom_first /*[natom+1]*/, const double *__restrict__ abundance,
               const double *__restrict__ atmos, const double *__restrict__ chem, double *__restrict__ pops,
               const double *__restrict__ ntot_in /* [ncol][natom][ndep] or NULL: atom->ntotal as a previous
                                                     ChemicalEquilibrium() left it (second Background() of an NLTE run) */)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t % ndep);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double T = at[RHB200_AT_T*ndep + k], ne = at[RHB200_AT_NE*ndep + k], nHtot = at[RHB200_AT_NHTOT*ndep + k];
  const double c1 = (RH_HPLANCK/(2.0*RH_PI*RH_M_ELECTRON)) * (RH_HPLANCK/RH_KBOLTZMANN);
  const double cNe_T = 0.5*ne * rhm::rh_pow(c1/T, 1.5);
  double *P = pops + (size_t) col * nlev * ndep + k;
  for (int a = 0; a < natom; a++) {
    const int l0 = atom_first[a], l1 = atom_first[a+1];
    const double E0 = lev[5*(size_t) l0 + 1], g0 = lev[5*(size_t) l0 + 3];
    const int st0 = (int) lev[5*(size_t) l0 + 2];
    double sum = 1.0;
    for (int i = l0 + 1; i < l1; i++) {
      const double dE = lev[5*(size_t) i + 1] - E0;
      const double gi0 = lev[5*(size_t) i + 3] / g0;
      const int dZ = (int) lev[5*(size_t) i + 2] - st0;
      const double dE_kT = dE / (RH_KBOLTZMANN * T);
      double ns = gi0 * rhm::rh_exp(-dE_kT);
      for (int m = 1; m <= dZ; m++) ns /= cNe_T;
      P[(size_t) i * ndep] = ns;
      sum += ns;
    }
    const double ntotal = ntot_in ? ntot_in[((size_t) col * natom + a) * ndep + k] : abundance[a] * nHtot;   // readatom.c:190
    const double n0 = ntotal / sum;
    const double fraction = chem ? chem[((size_t) col * (natom + 4) + a) * ndep + k] : 1.0;   // else chemeq_kernel rescales
    P[(size_t) l0 * ndep] = n0 * fraction;                          // chemequil.c:339
    for (int i = l0 + 1; i < l1; i++) P[(size_t) i * ndep] = (P[(size_t) i * ndep] * n0) * fraction;
  }
}

// ---- ChemicalEquilibrium (chemequil.c:107-392) with equilconstant (:456-530): Newton-Raphson on the number
//      conservation + Saha equations of the nuclei bound in molecules, per (column, depth).  NG_CHEM_ORDER = 0,
//      so Accelerate() only stores iterates and MaxChange() decides convergence (N_MAX_CHEM_ITER 10,
//      CHEM_ITER_LIMIT 1e-3, background.h:18-19).  Writes the chem block the continuum consumes and rescales
//      the LTE populations of the nuclei's model atoms (:334-343).
#define CHEM_MAXEQ 24
#define CHEM_MAXNUC 8
enum { MC_FIT = 0, MC_CHARGE, MC_NNUCLEI, MC_NELEMENT, MC_NEQC, MC_TMIN, MC_TMAX, MC_EDISS, MC_EQC0 = 8, MC_NUC0 = 16,
       MC_CNT0 = 20, MC_NFIELD = 32 };

// pow(x, n) for the small integer exponents of the chemical network (constituent counts, molecular charge): pow() of
// this libm returns x itself for y = 1 and 1 for y = 0 (the exact result is representable and the routine's error is
// far below half an ulp; checked on 3 M arguments over the whole exponent range), so those two cases -- the molecular
// charge always, most constituent counts -- skip the log/exp evaluation; every other exponent takes rh_pow
__device__ __forceinline__ double pow_count(double x, int n)
{
  if (n == 1 && x > 1.0e-300 && x < 1.0e300) return x;
  if (n == 0) return 1.0;
  return rhm::rh_pow(x, (double) n);
}

__device__ __forceinline__ double equilconstant_d(const double *__restrict__ m, double T)
{
  if (T < m[MC_TMIN] || T > m[MC_TMAX]) return 0.0;
  const int fit = (int) m[MC_FIT], neqc = (int) m[MC_NEQC], nnuc = (int) m[MC_NNUCLEI], charge = (int) m[MC_CHARGE];
  const double *c = m + MC_EQC0;
  double eqc = c[0], cgs_to_SI = 1.0;
  if (fit == 0 || fit == 1) {                               // KURUCZ_70 / KURUCZ_85
    const double t = (fit == 0) ? T : T * 1.0E-4;
    const double kT = RH_KBOLTZMANN * T;
    const int mk = nnuc - 1 - charge;
    for (int i = 1; i < neqc; i++) eqc = eqc*t + c[i];
    eqc = rhm::rh_exp(m[MC_EDISS]/kT + eqc - 1.5*mk*rhm::rh_log(T));
    cgs_to_SI = rhm::rh_pow(CUBE(RH_CM_TO_M), (double) mk);
  } else if (fit == 2 || fit == 3) {                        // SAUVAL_TATUM_84 (IRWIN_81 falls through to it)
    const double theta = RH_THETA0 / T;
    const double t = rhm::rh_log10(theta);
    const double kT = RH_KBOLTZMANN * T;
    for (int i = 1; i < neqc; i++) eqc = eqc*t + c[i];
    eqc = rhm::rh_exp(RH_LG10 * ((m[MC_EDISS]/RH_EV) * theta - eqc)) * kT;
  } else {                                                  // TSUJI_73
    const double theta = RH_THETA0 / T;
    const double kT = RH_KBOLTZMANN * T;
    for (int i = 1; i < neqc; i++) eqc = eqc*theta + c[i];
    eqc = SQ(kT) * rhm::rh_exp(RH_LG10 * (-eqc));
    cgs_to_SI = rhm::rh_pow(CUBE(RH_CM_TO_M) / 1.0E-07, (double) (nnuc - 1));
  }
  return eqc * cgs_to_SI;
}

// SH: the Jacobian lives in shared memory, one element per thread interleaved (rhlu::Interleaved) -- the O(N^3)
// accesses of the LU then never leave the SM; with the matrix in thread-local memory the kernel is L2-bound.
template <int MAXEQ, bool SH>
__global__ void __launch_bounds__(64)
chemeq_kernel(int ncol, int ndep, int natom, int nlev, const double *__restrict__ lev, const int *__restrict__ atom_first,
              const double *__restrict__ abundance, const double *__restrict__ atmos,
              int nnuc, const int *__restrict__ nuc_atom, int nmol, const double *__restrict__ mol,
              int iH2, int iOH, int iCH, int NmaxIter, double iterLimit,
              double *__restrict__ pops, double *__restrict__ chem,
              int nsel, const int *__restrict__ molsel, double *__restrict__ molout /* [ncol][nsel][ndep] or NULL */,
              double *__restrict__ ntot_out /* [ncol][natom][ndep] or NULL: atom->ntotal after chemequil.c:342 */,
              int ntot_is_input /* the array holds atom->ntotal of a previous call: the fraction is relative to it (:336) */)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t % ndep);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double T = at[RHB200_AT_T*ndep + k], ne = at[RHB200_AT_NE*ndep + k], nHtot = at[RHB200_AT_NHTOT*ndep + k];
  double *P = pops + (size_t) col * nlev * ndep + k;
  const int Neq = nnuc + nmol;
  double n[MAXEQ], f[MAXEQ], a[MAXEQ], df_local[SH ? 1 : MAXEQ*MAXEQ], prev[2][MAXEQ];
  double fn0[CHEM_MAXNUC], Phi[MAXEQ];
  extern __shared__ double chem_smem[];
  auto DF = [&](int i, int j) -> double & {
    if constexpr (SH) return chem_smem[(size_t) (i*Neq + j) * blockDim.x + threadIdx.x];
    else return df_local[i*Neq + j];
  };
  for (int i = 0; i < Neq; i++) a[i] = 0.0;
  for (int i = 0; i < nnuc; i++) {                           // chemequil.c:233-245 (every nucleus has a model atom)
    const int am = nuc_atom[i], l0 = atom_first[am], l1 = atom_first[am+1];
    double s = 0.0;
    for (int j = l0; j < l1; j++) {
      if ((int) lev[5*(size_t) j + 2] > 0) break;
      s += P[(size_t) j * ndep];
    }
    a[i] = abundance[am] * nHtot;
    // atom->ntotal[k] == abundance * nHtot (readatom.c:190) unless a previous call reduced it (chemequil.c:342)
    fn0[i] = s / (ntot_is_input ? ntot_out[((size_t) col * natom + am) * ndep + k] : a[i]);
  }
  const double CI = (RH_HPLANCK/(2.0*RH_PI*RH_M_ELECTRON)) * (RH_HPLANCK/RH_KBOLTZMANN);
  const double PhiHmin = 0.25*rhm::rh_pow(CI/T, 1.5) * rhm::rh_exp(0.754 * RH_EV / (RH_KBOLTZMANN * T));
  const double fHmin = ne * fn0[0]*PhiHmin;
  for (int i = 0; i < nmol; i++) Phi[i] = equilconstant_d(mol + (size_t) i * MC_NFIELD, T);
  for (int i = 0; i < nnuc; i++) n[i] = a[i];
  for (int i = 0; i < nmol; i++) n[nnuc+i] = 0.0;
  int count = 1;
  for (int i = 0; i < Neq; i++) prev[0][i] = n[i];
  int niter = 1;
  while (niter <= NmaxIter) {
    for (int i = 0; i < Neq; i++) {
      f[i] = n[i] - a[i];
      for (int j = 0; j < Neq; j++) DF(i, j) = 0.0;
      DF(i, i) = 1.0;
    }
    f[0] += fHmin * n[0];
    DF(0, 0) += fHmin;
    for (int i = 0; i < nmol; i++) {
      const double *m = mol + (size_t) i * MC_NFIELD;
      const int nel = (int) m[MC_NELEMENT];
      double saha = Phi[i];
      for (int j = 0; j < nel; j++) {
        const int nu = (int) m[MC_NUC0 + j];
        const int cnt = (int) m[MC_CNT0 + j];
        saha *= pow_count(fn0[nu] * n[nu], cnt);
        f[nu] += cnt * n[nnuc + i];
      }
      saha /= pow_count(ne, (int) m[MC_CHARGE]);
      f[nnuc + i] -= saha;
      for (int j = 0; j < nel; j++) {
        const int nu = (int) m[MC_NUC0 + j];
        const int cnt = (int) m[MC_CNT0 + j];
        DF(nu, nnuc + i) += cnt;
        DF(nnuc + i, nu) = -saha * (cnt/n[nu]);
      }
    }
    if constexpr (SH) rhlu::solve_linear_eq_mat<MAXEQ>(Neq, rhlu::Interleaved{chem_smem + threadIdx.x, Neq, (int) blockDim.x}, f, true);
    else rhlu::solve_linear_eq<MAXEQ>(Neq, df_local, f, true);
    for (int i = 0; i < Neq; i++) n[i] -= f[i];
    {                                                        // Accelerate (store) + MaxChange, accelerate.c:75-79, maxchange.c:38-46
      const int slot = count % 2;
      for (int i = 0; i < Neq; i++) prev[slot][i] = n[i];
      count++;
    }
    double dmax = 0.0;
    {
      const double *old = prev[(count - 2) % 2], *nw = prev[(count - 1) % 2];
      for (int i = 0; i < Neq; i++)
        if (nw[i] != 0.0) { const double d = fabs((nw[i] - old[i]) / nw[i]); dmax = (dmax > d) ? dmax : d; }
    }
    if (dmax <= iterLimit) break;
    niter++;
  }
  double *ch = chem + (size_t) col * (natom + 4) * ndep + k;
  for (int am = 0; am < natom; am++) ch[(size_t) am * ndep] = 1.0;
  for (int i = 0; i < nnuc; i++) {                           // chemequil.c:334-343
    const int am = nuc_atom[i];
    const double fraction = n[i] / (ntot_is_input ? ntot_out[((size_t) col * natom + am) * ndep + k] : a[i]);
    ch[(size_t) am * ndep] = fraction;
    for (int j = atom_first[am]; j < atom_first[am+1]; j++) P[(size_t) j * ndep] *= fraction;
    if (ntot_out) ntot_out[((size_t) col * natom + am) * ndep + k] = n[i];
  }
  ch[(size_t) natom * ndep] = ne * (n[0] * PhiHmin);         // nHmin, chemequil.c:347
  ch[(size_t) (natom + 1) * ndep] = iH2 >= 0 ? n[nnuc + iH2] : 0.0;
  ch[(size_t) (natom + 2) * ndep] = iOH >= 0 ? n[nnuc + iOH] : 0.0;
  ch[(size_t) (natom + 3) * ndep] = iCH >= 0 ? n[nnuc + iCH] : 0.0;
  if (molout) for (int q = 0; q < nsel; q++) molout[((size_t) col * nsel + q) * ndep + k] = n[nnuc + molsel[q]];   // molecule->n of those with line lists
}

// ---- the same Newton-Raphson, LPS lanes per system (LPS = 16: two systems per warp; 32: one).  Thread-per-system
//      keeps 5.4 KB of matrices in local memory and is L2-bound; here lane l owns row l of the Jacobian, which
//      lives in shared memory (row stride LPS+1: conflict-free column walks).  Every floating-point result keeps
//      the reference's operation order:
//        * Crout's sums A[i][j] -= A[i][k]*A[k][j] run k = 0, 1, ... for every element (ludcmp.c:108-124); the
//          lanes do them right-looking -- as soon as A[k][j] is final it is broadcast and every lane i > k
//          subtracts its product -- which only changes WHEN each subtraction happens, not their order;
//        * the pivot is the LAST row attaining the maximum of vv[i]*|A[i][j]| (`>=`, :121): ballot + highest lane;
//        * forward substitution likewise right-looking with the reference's skip of leading zeros (`ii`, :163-168),
//          the row interchanges applied up front (equivalent: index[i] >= i);
//        * back substitution sums j = i+1 .. N-1 ascending while x[j] become known descending (:170-175): that
//          order is inherently serial, one lane does it;
//        * residual rows, the molecule rows (pow, the expensive part) and equilconstant run one per lane.
template <int LPS>
__global__ void __launch_bounds__(128, 7)      // 72 registers, 7 x 27 KB of shared memory per SM; measured e2e per 2048 columns: unbounded (80 regs) 15.38 ms, 7 -> 15.16, 8 -> 15.24
chemeq_coop_kernel(int ncol, int ndep, int natom, int nlev, const double *__restrict__ lev, const int *__restrict__ atom_first,
                   const double *__restrict__ abundance, const double *__restrict__ atmos,
                   int nnuc, const int *__restrict__ nuc_atom, int nmol, const double *__restrict__ mol,
                   int iH2, int iOH, int iCH, int NmaxIter, double iterLimit,
                   double *__restrict__ pops, double *__restrict__ chem,
              int nsel, const int *__restrict__ molsel, double *__restrict__ molout /* [ncol][nsel][ndep] or NULL */,
              double *__restrict__ ntot_out /* [ncol][natom][ndep] or NULL: atom->ntotal after chemequil.c:342 */,
              int ntot_is_input /* the array holds atom->ntotal of a previous call: the fraction is relative to it (:336) */)
{
  constexpr int LD = LPS + 1, SPB = 128 / LPS, PER = LPS*LD + 9*LPS;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double chem_smem[];
  const int l = threadIdx.x % LPS, sl = threadIdx.x / LPS;
  const size_t t = (size_t) blockIdx.x * SPB + sl;
  const bool live = t < (size_t) ncol * ndep;
  const size_t tt = live ? t : 0;
  const int col = (int) (tt / ndep), k = (int) (tt % ndep);
  double *A = chem_smem + (size_t) sl * PER, *x = A + LPS*LD, *bc = x + LPS, *r = bc + LPS,
         *nv = r + LPS, *av = nv + LPS, *vv = av + LPS, *fn0 = vv + LPS, *Phi = fn0 + LPS;
  int *idx = (int *) (Phi + LPS);
  const unsigned half_shift = (LPS == 32) ? 0u : (unsigned) (16 * ((threadIdx.x & 31) / 16));
  const unsigned half_mask = (LPS == 32) ? FULL : 0xffffu;

  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * ndep;
  const double T = at[RHB200_AT_T*ndep + k], ne = at[RHB200_AT_NE*ndep + k], nHtot = at[RHB200_AT_NHTOT*ndep + k];
  double *P = pops + (size_t) col * nlev * ndep + k;
  const int Neq = nnuc + nmol;
  const bool row = l < Neq;
  if (l < nnuc) {                                            // chemequil.c:233-245
    const int am = nuc_atom[l], l0 = atom_first[am], l1 = atom_first[am+1];
    double s = 0.0;
    for (int j = l0; j < l1; j++) {
      if ((int) lev[5*(size_t) j + 2] > 0) break;
      s += P[(size_t) j * ndep];
    }
    av[l] = abundance[am] * nHtot;
    fn0[l] = s / (ntot_is_input ? ntot_out[((size_t) col * natom + am) * ndep + k] : av[l]);
    nv[l] = av[l];
  } else if (row) { av[l] = 0.0; nv[l] = 0.0; }
  if (l < nmol) Phi[l] = equilconstant_d(mol + (size_t) l * MC_NFIELD, T);
  __syncwarp();
  const double CI = (RH_HPLANCK/(2.0*RH_PI*RH_M_ELECTRON)) * (RH_HPLANCK/RH_KBOLTZMANN);
  const double PhiHmin = 0.25*rhm::rh_pow(CI/T, 1.5) * rhm::rh_exp(0.754 * RH_EV / (RH_KBOLTZMANN * T));
  const double fHmin = ne * fn0[0]*PhiHmin;
  // The Jacobian is kept only as LU factors; SolveLinearEq's iterative improvement (ludcmp.c:60-73) needs the original
  // matrix again, and row l regenerates its entries instead of reading a copy (44 -> 27 KB of shared memory per block):
  // a nucleus row is 1 (+ fHmin) on the diagonal and the constituent counts -- 4 bits per molecule -- in the molecule
  // columns; a molecule row is 1 on the diagonal and the <= 4 derivatives it has just computed
  unsigned long long cpack[LPS / 16];
#pragma unroll
  for (int w = 0; w < LPS / 16; w++) cpack[w] = 0ull;
  if (l < nnuc)
    for (int i = 0; i < nmol; i++) {
      const double *m = mol + (size_t) i * MC_NFIELD;
      const int nel = (int) m[MC_NELEMENT];
      int cl = 0;
      for (int j = 0; j < nel; j++) if ((int) m[MC_NUC0 + j] == l) cl += (int) m[MC_CNT0 + j];
#pragma unroll
      for (int w = 0; w < LPS / 16; w++) if (i / 16 == w) cpack[w] |= (unsigned long long) (cl & 15) << (4 * (i % 16));
    }
  // a molecule row keeps its own record in registers (12 bits per constituent: nucleus, count; then the number of
  // constituents) so that the Newton loop reads nothing from global memory; the host routes networks that name a
  // nucleus twice in one molecule to the per-thread kernel
  unsigned long long mpack = 0ull;
  int mcharge = 0;
  if (row && l >= nnuc) {
    const double *m = mol + (size_t) (l - nnuc) * MC_NFIELD;
    const int nel = (int) m[MC_NELEMENT];
    for (int j = 0; j < nel; j++)
      mpack |= ((unsigned long long) ((int) m[MC_NUC0 + j] & 255) | ((unsigned long long) ((int) m[MC_CNT0 + j] & 15) << 8)) << (12 * j);
    mpack |= (unsigned long long) nel << 48;
    mcharge = (int) m[MC_CHARGE];
  }
  int jn[4] = {-1, -1, -1, -1};
  double jv[4] = {0.0, 0.0, 0.0, 0.0};
  double n_l = row ? nv[l] : 0.0, n_old = n_l;              // Accelerate()'s two stored iterates, element l
  bool done = !live;
  int imax = 0;

  auto backsubst = [&](double *v) {                          // LUbacksubst, ludcmp.c:156-177
    double s = row ? v[l] : 0.0;                              // the row exchanges of the decomposition, element l in lane l
    for (int i = 0; i < Neq; i++) {
      const int ip = idx[i];
      s = __shfl_sync(FULL, s, (l == i) ? ip : (l == ip) ? i : l, LPS);
    }
    int ii = -1;
    for (int j = 0; j < Neq; j++) {
      const double xj = __shfl_sync(FULL, s, j, LPS);
      if (ii < 0 && xj != 0.0) ii = j;
      if (row && l > j && ii >= 0) s -= A[l*LD + j] * xj;
    }
    // back substitution, ludcmp.c:170-175: x_i = (v_i - sum_{j>i} A_ij x_j) / A_ii with the subtractions in ascending j.
    // Lane j forms its product A_ij x_j as soon as x_j exists; the ordered sum then runs over register values fetched
    // with shuffles that do not depend on the running sum, instead of one lane walking shared memory serially
    double xv = row ? s : 0.0;
    for (int i = Neq - 1; i >= 0; i--) {
      const double p = (row && l > i) ? A[i*LD + l] * xv : 0.0;
      double sum = __shfl_sync(FULL, xv, i, LPS);
      for (int j = i + 1; j < Neq; j++) sum -= __shfl_sync(FULL, p, j, LPS);
      const double xi = sum / A[i*LD + i];
      if (l == i) xv = xi;
    }
    if (row) v[l] = xv;
    __syncwarp();
  };

  for (int niter = 1; niter <= NmaxIter; niter++) {
    // ---- f and the Jacobian, row l (chemequil.c:262-298)
    if (row) {
      double fl = nv[l] - av[l];
      for (int j = 0; j < Neq; j++) A[l*LD + j] = 0.0;
      A[l*LD + l] = 1.0;
      if (l == 0) { fl += fHmin * nv[0]; A[0] += fHmin; }
      if (l < nnuc) {
        for (int i = 0; i < nmol; i++) {                     // chemequil.c:268-281 through the packed counts
          unsigned long long wsel = cpack[0];
#pragma unroll
          for (int w = 1; w < LPS / 16; w++) if (i / 16 == w) wsel = cpack[w];
          const int cnt = (int) ((wsel >> (4 * (i % 16))) & 15ull);
          if (cnt) { fl += cnt * nv[nnuc + i]; A[l*LD + nnuc + i] += cnt; }
        }
      } else {
        const int nel = (int) (mpack >> 48);
        double saha = Phi[l - nnuc];
        for (int j = 0; j < nel; j++) {
          const int nu = (int) ((mpack >> (12 * j)) & 255ull);
          saha *= pow_count(fn0[nu] * nv[nu], (int) ((mpack >> (12 * j + 8)) & 15ull));
        }
        saha /= pow_count(ne, mcharge);
        fl -= saha;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (j < nel) {
            const int nu = (int) ((mpack >> (12 * j)) & 255ull);
            const int cnt = (int) ((mpack >> (12 * j + 8)) & 15ull);
            jn[j] = nu; jv[j] = -saha * (cnt/nv[nu]);
            A[l*LD + nu] = jv[j];
          } else jn[j] = -1;
        }
      }
      x[l] = fl; bc[l] = fl;
      double big = 0.0;                                      // LUdecomp, ludcmp.c:92-101
      for (int j = 0; j < Neq; j++) { const double temp = fabs(A[l*LD + j]); if (temp > big) big = temp; }
      vv[l] = 1.0 / big;
    }
    __syncwarp();
    imax = 0;
    for (int j = 0; j < Neq; j++) {                          // ludcmp.c:103-150
      double s = row ? A[l*LD + j] : 0.0;
      for (int kk = 0; kk < j; kk++) {
        const double v = __shfl_sync(FULL, s, kk, LPS);
        if (row && l > kk) s -= A[l*LD + kk] * v;
      }
      if (row) A[l*LD + j] = s;
      const bool cand = row && l >= j;
      const double dum = cand ? vv[l]*fabs(s) : -1.0;
      double mx = dum;
      for (int off = LPS/2; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, off, LPS));
      const unsigned bal = (__ballot_sync(FULL, cand && dum == mx) >> half_shift) & half_mask;
      if (bal) imax = 31 - __clz(bal);
      __syncwarp();
      if (j != imax) {
        if (row) { const double d2 = A[imax*LD + l]; A[imax*LD + l] = A[j*LD + l]; A[j*LD + l] = d2; }
        if (l == 0) vv[imax] = vv[j];
      }
      if (l == 0) idx[j] = imax;
      __syncwarp();
      double piv = A[j*LD + j];
      __syncwarp();
      if (piv == 0.0) { piv = 1.0e-20; if (l == 0) A[j*LD + j] = piv; }
      const double inv = 1.0 / piv;
      if (row && l > j) A[l*LD + j] *= inv;
      __syncwarp();
    }
    backsubst(x);
    if (row) {                                               // one step of iterative improvement, ludcmp.c:60-73
      double rr = bc[l];
      if (l < nnuc) {
        const double diag = (l == 0) ? 1.0 + fHmin : 1.0;
        for (int j = 0; j < nnuc; j++) rr -= ((j == l) ? diag : 0.0) * x[j];
        for (int i = 0; i < nmol; i++) {
          unsigned long long wsel = cpack[0];
#pragma unroll
          for (int w = 1; w < LPS / 16; w++) if (i / 16 == w) wsel = cpack[w];
          rr -= (double) (int) ((wsel >> (4 * (i % 16))) & 15ull) * x[nnuc + i];
        }
      } else {
        for (int j = 0; j < Neq; j++) {
          double a = (j == l) ? 1.0 : 0.0;
#pragma unroll
          for (int e = 0; e < 4; e++) if (jn[e] == j) a = jv[e];
          rr -= a * x[j];
        }
      }
      r[l] = rr;
    }
    __syncwarp();
    backsubst(r);
    double d = 0.0;
    if (row && !done) {
      const double nw = n_l - (x[l] + r[l]);
      n_old = n_l; n_l = nw;
      nv[l] = nw;
      if (nw != 0.0) d = fabs((nw - n_old) / nw);           // MaxChange, maxchange.c:38-46
    }
    for (int off = LPS/2; off > 0; off >>= 1) d = fmax(d, __shfl_xor_sync(FULL, d, off, LPS));
    if (d <= iterLimit) done = true;
    __syncwarp();
    if (__all_sync(FULL, done)) break;
  }
  if (!live) return;
  double *ch = chem + (size_t) col * (natom + 4) * ndep + k;
  for (int am = l; am < natom; am += LPS) ch[(size_t) am * ndep] = 1.0;
  __syncwarp(__activemask());
  if (l < nnuc) {                                            // chemequil.c:334-343
    const int am = nuc_atom[l];
    const double fraction = nv[l] / (ntot_is_input ? ntot_out[((size_t) col * natom + am) * ndep + k] : av[l]);
    ch[(size_t) am * ndep] = fraction;
    for (int j = atom_first[am]; j < atom_first[am+1]; j++) P[(size_t) j * ndep] *= fraction;
    if (ntot_out) ntot_out[((size_t) col * natom + am) * ndep + k] = nv[l];
  }
  if (l == 0) {
    ch[(size_t) natom * ndep] = ne * (nv[0] * PhiHmin);      // nHmin, chemequil.c:347
    ch[(size_t) (natom + 1) * ndep] = iH2 >= 0 ? nv[nnuc + iH2] : 0.0;
    ch[(size_t) (natom + 2) * ndep] = iOH >= 0 ? nv[nnuc + iOH] : 0.0;
    ch[(size_t) (natom + 3) * ndep] = iCH >= 0 ? nv[nnuc + iCH] : 0.0;
    if (molout) for (int q = 0; q < nsel; q++) molout[((size_t) col * nsel + q) * ndep + k] = nv[nnuc + molsel[q]];
  }
}

struct ContinuumState {
  DevModel D; Holder H;
  int natom = 0, nlev = 0, nlambda = 0, has_H2 = 0, has_OH = 0, has_CH = 0, proton_level = 0;
  double *d_lev = nullptr, *d_abund = nullptr; int *d_first = nullptr;
  // chemistry on the device (rhb200_set_chemistry)
  int nnuc = 0, nmol = 0, iH2 = -1, iOH = -1, iCH = -1;
  bool mol_repeats = false;      // some molecule names a nucleus twice, or a count above 15: per-thread kernel
  int *d_nuc_atom = nullptr; double *d_mol = nullptr;
  int nsel = 0; int *d_molsel = nullptr;      // molecules whose densities the molecular-line kernels need
  std::vector<int> h_molsel, h_first;
  std::vector<double> h_abund;
};

void rh_continuum_free(rhb200_ctx *c)
{
  if (c->cont) { delete (ContinuumState *) c->cont; c->cont = nullptr; }
}

int rh_continuum_nlev(const rhb200_ctx *c) { return c->cont ? ((ContinuumState *) c->cont)->nlev : 0; }
int rh_continuum_natom(const rhb200_ctx *c) { return c->cont ? ((ContinuumState *) c->cont)->natom : 0; }
// last level of the first model atom: hydrogen comes first in atoms.input (atmos.H = &atmos.atoms[0], readatom.c)
int rh_continuum_set_molsel(rhb200_ctx *c, int nsel, const int *chem_index)
{
  ContinuumState *S = (ContinuumState *) c->cont;
  if (!S) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  for (int q = 0; q < nsel; q++)
    if (chem_index[q] < 0 || chem_index[q] >= S->nmol) { rhb200_set_error("molecule index %d outside the chemical network (rhb200_set_chemistry first)", chem_index[q]); return RHB200_EINVAL; }
  if ((int) S->h_molsel.size() == nsel && std::equal(chem_index, chem_index + nsel, S->h_molsel.begin()) && S->nsel == nsel) return RHB200_OK;
  S->h_molsel.assign(chem_index, chem_index + nsel);
  S->nsel = nsel;
  return nsel ? S->H.put(&S->d_molsel, chem_index, (size_t) nsel) : RHB200_OK;
}
void rh_continuum_set_hse_mode(rhb200_ctx *c, int on) { if (c->cont) ((ContinuumState *) c->cont)->D.hse_mode = on; }
int rh_continuum_nlambda(const rhb200_ctx *c) { return c->cont ? ((ContinuumState *) c->cont)->nlambda : 0; }
int rh_continuum_has_chemistry(const rhb200_ctx *c) { return c->cont && ((ContinuumState *) c->cont)->nmol > 0; }
int rh_continuum_proton_level(const rhb200_ctx *c) { return c->cont ? ((ContinuumState *) c->cont)->proton_level : 0; }

static int launch_chemeq(rhb200_ctx *c, ContinuumState *S, int cc, int ndep, const double *d_atmos, double *d_pops, double *d_chem,
                         double *d_molout = nullptr, double *d_ntot = nullptr, int ntot_is_input = 0)
{
  const size_t cn = (size_t) cc * ndep;
  const int na = S->natom;
  const int Neq = S->nnuc + S->nmol;
  // RHB200_CHEM_KERNEL = coop (default) | local | shared: A/B switch between the three bit-identical variants.
  // Measured per 2048 columns x 70 depths (B200): coop 3.0 ms (issue-bound: 48 % issue slots, the serial back
  // substitution is 27 % of the instructions), local 3.3 ms (L2-bound), shared 4.5 ms (3 warps/SM, latency-bound)
  static const char *variant = getenv("RHB200_CHEM_KERNEL");
  const bool serial = (variant && !strcmp(variant, "local")) || S->mol_repeats, coop = !variant || !strcmp(variant, "coop");
  const int TPB = 32;
#define CHEM_ARGS cc, ndep, na, S->nlev, S->d_lev, S->d_first, S->d_abund, d_atmos, S->nnuc, S->d_nuc_atom, S->nmol, S->d_mol, \
                S->iH2, S->iOH, S->iCH, 10, 1.0E-3, d_pops, d_chem, S->nsel, S->d_molsel, d_molout, d_ntot, ntot_is_input
  if (serial) {
    if (Neq <= 16) chemeq_kernel<16, false><<<(unsigned) ((cn + 63) / 64), 64, 0, c->stream>>>(CHEM_ARGS);
    else chemeq_kernel<CHEM_MAXEQ, false><<<(unsigned) ((cn + 63) / 64), 64, 0, c->stream>>>(CHEM_ARGS);
  } else if (!coop) {
    const size_t sh = (size_t) Neq * Neq * TPB * sizeof(double);
    if (Neq <= 16) {
      RH_CUDA(cudaFuncSetAttribute(chemeq_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sh));
      chemeq_kernel<16, true><<<(unsigned) ((cn + TPB - 1) / TPB), TPB, sh, c->stream>>>(CHEM_ARGS);
    } else {
      RH_CUDA(cudaFuncSetAttribute(chemeq_kernel<CHEM_MAXEQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sh));
      chemeq_kernel<CHEM_MAXEQ, true><<<(unsigned) ((cn + TPB - 1) / TPB), TPB, sh, c->stream>>>(CHEM_ARGS);
    }
  } else if (Neq <= 16) {
    const size_t sh = (size_t) 8 * (16*17 + 9*16) * sizeof(double);
    RH_CUDA(cudaFuncSetAttribute(chemeq_coop_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sh));
    chemeq_coop_kernel<16><<<(unsigned) ((cn + 7) / 8), 128, sh, c->stream>>>(CHEM_ARGS);
  } else {
    const size_t sh = (size_t) 4 * (32*33 + 9*32) * sizeof(double);
    RH_CUDA(cudaFuncSetAttribute(chemeq_coop_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sh));
    chemeq_coop_kernel<32><<<(unsigned) ((cn + 3) / 4), 128, sh, c->stream>>>(CHEM_ARGS);
  }
#undef CHEM_ARGS
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

// LTE populations + continuum of one chunk of columns, all on ctx->stream, in three stages (the NLTE front end runs
// CollisionRate between the first two, like SetLTEQuantities does: ltepops.c:224-249).
// d_pops [cc][nlev][ndep] and d_tprep [cc][5][ndep] are workspace; out d_chi, d_eta [cc][nlambda][ndep].
int rh_continuum_ltepops(rhb200_ctx *c, int cc, int ndep, const double *d_atmos, const double *d_chem_host, double *d_pops,
                         const double *d_ntot_in)
{
  ContinuumState *S = (ContinuumState *) c->cont;
  if (!S) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  if (S->nlambda != c->wav.nlambda) { rhb200_set_error("wavelengths changed after rhb200_set_continuum()"); return RHB200_ESTATE; }
  const size_t cn = (size_t) cc * ndep;
  ScopedKernelTimer t(c, RHB200_K_PREP);
  ltepops_kernel<<<(unsigned) ((cn + 127) / 128), 128, 0, c->stream>>>(cc, ndep, S->natom, S->nlev, S->d_lev, S->d_first,
                                                                       S->d_abund, d_atmos, d_chem_host, d_pops, d_ntot_in);
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_continuum_chemeq(rhb200_ctx *c, int cc, int ndep, const double *d_atmos, double *d_pops, double *d_chem,
                        double *d_molout, double *d_ntot, int ntot_is_input)
{
  ContinuumState *S = (ContinuumState *) c->cont;
  if (!S) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  if (S->nmol == 0) { rhb200_set_error("rhb200_set_chemistry() has not been called"); return RHB200_ESTATE; }
  ScopedKernelTimer t(c, RHB200_K_PREP);
  return launch_chemeq(c, S, cc, ndep, d_atmos, d_pops, d_chem, d_molout, d_ntot, ntot_is_input);
}

// d_pops_n: the populations Background() reads through atom->n (differs from the LTE populations d_pops_star only for
// an ACTIVE hydrogen atom, whose n is the NLTE solution -- zero before initSolution(): hydrogen.c:100-141)
int rh_continuum_opac(rhb200_ctx *c, int cc, int ndep, const double *d_atmos, const double *d_chem, const double *d_pops_n,
                      const double *d_pops_star, double *d_tprep, double *d_chi, double *d_eta, double *d_sca)
{
  ContinuumState *S = (ContinuumState *) c->cont;
  if (!S) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  const size_t cn = (size_t) cc * ndep;
  const int na = S->natom;
  // T and ne are rows of the atmosphere block; the chem block carries nHmin, nH2, nOH, nCH after the fractions:
  // both blocks are [col][field][ndep], hence field pointers and per-column strides.
  {
    ScopedKernelTimer t(c, RHB200_K_PREP);
    cont_prep_kernel<<<(unsigned) ((cn + 127) / 128), 128, 0, c->stream>>>(cc, ndep, S->D, d_atmos + (size_t) RHB200_AT_T * ndep,
                                                                           (size_t) RHB200_AT_NFIELD * ndep, d_tprep);
  }
  {
    ScopedKernelTimer t(c, RHB200_K_OTHER);
    const int ntile = (S->nlambda + CONT_TL - 1) / CONT_TL;
    const size_t as = (size_t) RHB200_AT_NFIELD * ndep, cs = (size_t) (na + 4) * ndep;
    const size_t nck = (cn + 127) / 128;
    if (nck > 65535) { rhb200_set_error("chunk too large for the continuum kernel grid"); return RHB200_EINVAL; }
    static int minb = -1;                                   // RHB200_CONT_MINB: occupancy / spill trade-off
    if (minb < 0) { const char *e = getenv("RHB200_CONT_MINB"); minb = e ? atoi(e) : 8; }
#define RH_CONT_ARGS (cc, S->nlambda, ndep, S->D, d_atmos + (size_t) RHB200_AT_T * ndep, d_atmos + (size_t) RHB200_AT_NE * ndep, as, d_chem + (size_t) na * ndep, d_chem + (size_t) (na + 1) * ndep, d_chem + (size_t) (na + 2) * ndep, d_chem + (size_t) (na + 3) * ndep, cs, d_pops_n, d_pops_star, d_tprep, d_chi, d_eta, d_sca)
    const dim3 grid((unsigned) ntile, (unsigned) nck);
    switch (minb) {
    case 4: continuum_tile_kernel<4><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    case 5: continuum_tile_kernel<5><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    case 6: continuum_tile_kernel<6><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    case 7: continuum_tile_kernel<7><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    case 10: continuum_tile_kernel<10><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    case 12: continuum_tile_kernel<12><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    default: continuum_tile_kernel<8><<<grid, 128, 0, c->stream>>>RH_CONT_ARGS; break;
    }
#undef RH_CONT_ARGS
  }
  RH_CUDA(cudaGetLastError());
  return RHB200_OK;
}

int rh_continuum_chunk(rhb200_ctx *c, int cc, int ndep, const double *d_atmos, const double *d_chem,
                       double *d_pops, double *d_tprep, double *d_chi, double *d_eta, int chem_on_device, double *d_molout,
                       double *d_sca)
{
  RH_CHECK(rh_continuum_ltepops(c, cc, ndep, d_atmos, chem_on_device ? nullptr : d_chem, d_pops, nullptr));
  if (chem_on_device) RH_CHECK(rh_continuum_chemeq(c, cc, ndep, d_atmos, d_pops, (double *) d_chem, d_molout, nullptr, 0));
  return rh_continuum_opac(c, cc, ndep, d_atmos, d_chem, d_pops, d_pops, d_tprep, d_chi, d_eta, d_sca);
}

// model-atom bookkeeping the NLTE front end needs: first level row of each atom, abundance * nHtot is formed there
int rh_continuum_atom_first(const rhb200_ctx *c, int atom)
{
  const ContinuumState *S = (const ContinuumState *) c->cont;
  return (S && atom >= 0 && atom <= S->natom) ? S->h_first[atom] : -1;
}
double rh_continuum_abundance(const rhb200_ctx *c, int atom)
{
  const ContinuumState *S = (const ContinuumState *) c->cont;
  return (S && atom >= 0 && atom < S->natom) ? S->h_abund[atom] : 0.0;
}

extern "C" int rhb200_set_continuum(rhb200_ctx *c, const rhb200_continuum_model *m, const double *abundance)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  RH_CHECK(check_model(m));
  if (!abundance) { rhb200_set_error("abundance missing"); return RHB200_EINVAL; }
  if (c->wav.nlambda == 0) { rhb200_set_error("rhb200_set_wavelengths() has not been called"); return RHB200_ESTATE; }
  // lev[g][4] != 0 marks the levels of an ACTIVE atom: its bound-free continua must then carry bf[c][9] != 0 so that
  // Metal_bf skips them (metal.c:105), and hydrogen's go with H_active (hydrogen.c:179)
  for (int cb = 0; cb < m->ncont; cb++) {
    const double *b = m->bf + 10*(size_t) cb;
    const bool act = m->lev[5*(size_t) ((int) b[1]) + 4] != 0.0;
    if (act && (int) b[0] != 0 && b[9] == 0.0) { rhb200_set_error("continuum %d belongs to an ACTIVE atom but is not flagged (bf[9])", cb); return RHB200_EINVAL; }
    if (act && (int) b[0] == 0 && !m->H_active) { rhb200_set_error("hydrogen levels are flagged ACTIVE but H_active is 0"); return RHB200_EINVAL; }
  }
  rh_continuum_free(c);
  ContinuumState *S = new ContinuumState();
  c->cont = S;
  int rc = build_model(m, c->wav.nlambda, c->h_lambda.data(), S->D, S->H);
  if (rc != RHB200_OK) { rh_continuum_free(c); return rc; }
  S->natom = m->natom; S->nlev = m->nlev; S->nlambda = c->wav.nlambda;
  S->has_H2 = m->has_H2; S->has_OH = m->has_OH; S->has_CH = m->has_CH;
  std::vector<int> first(m->natom + 1, -1);
  for (int g = 0; g < m->nlev; g++) {
    const int a = (int) m->lev[5*(size_t) g];
    if (a < 0 || a >= m->natom || (g > 0 && a < (int) m->lev[5*(size_t) (g-1)])) { rh_continuum_free(c); rhb200_set_error("level table must be grouped by atom"); return RHB200_EINVAL; }
    if (first[a] < 0) first[a] = g;
  }
  first[m->natom] = m->nlev;
  for (int a = m->natom - 1; a >= 0; a--) if (first[a] < 0) first[a] = first[a+1];
  S->proton_level = first[1] - 1;
  S->h_first = first; S->h_abund.assign(abundance, abundance + m->natom);
  if ((rc = S->H.put(&S->d_lev, m->lev, (size_t) m->nlev * 5)) != RHB200_OK ||
      (rc = S->H.put(&S->d_abund, abundance, (size_t) m->natom)) != RHB200_OK ||
      (rc = S->H.put(&S->d_first, first.data(), first.size())) != RHB200_OK) { rh_continuum_free(c); return rc; }
  return RHB200_OK;
}

extern "C" int rhb200_set_chemistry(rhb200_ctx *c, int nnuclei, const int *nucleus_atom, int nmol, const double *mol)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  ContinuumState *S = (ContinuumState *) c->cont;
  if (!S) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  if (nnuclei < 1 || nnuclei > CHEM_MAXNUC || nmol < 1 || nnuclei + nmol > CHEM_MAXEQ || !nucleus_atom || !mol) {
    rhb200_set_error("chemical network: 1..%d nuclei and at most %d equations", CHEM_MAXNUC, CHEM_MAXEQ); return RHB200_EINVAL;
  }
  for (int i = 0; i < nnuclei; i++)
    if (nucleus_atom[i] < 0 || nucleus_atom[i] >= S->natom) {
      rhb200_set_error("nucleus %d has no model atom (getfjk path, solvene.c:143, is not implemented)", i); return RHB200_EUNSUPPORTED;
    }
  if (nucleus_atom[0] != 0) { rhb200_set_error("first nucleus must be hydrogen (chemequil.c:146)"); return RHB200_EINVAL; }
  S->iH2 = S->iOH = S->iCH = -1;
  S->mol_repeats = false;
  for (int i = 0; i < nmol; i++) {
    const double *m = mol + (size_t) i * MC_NFIELD;
    const int nel = (int) m[MC_NELEMENT], fit = (int) m[MC_FIT];
    if (nel < 1 || nel > 4 || (int) m[MC_NEQC] < 1 || (int) m[MC_NEQC] > 8 || fit < 0 || fit > 4) { rhb200_set_error("molecule %d: bad element / coefficient count or fit", i); return RHB200_EINVAL; }
    for (int j = 0; j < nel; j++) if ((int) m[MC_NUC0 + j] < 0 || (int) m[MC_NUC0 + j] >= nnuclei) { rhb200_set_error("molecule %d: nucleus index out of range", i); return RHB200_EINVAL; }
    for (int j = 0, tot = 0; j < nel; j++) {                 // the cooperative kernel packs a nucleus' count per molecule in 4 bits
      const int cnt = (int) m[MC_CNT0 + j];
      if (cnt < 1) { rhb200_set_error("molecule %d: constituent count %d", i, cnt); return RHB200_EINVAL; }
      tot += cnt;
      if (tot > 15) S->mol_repeats = true;
      for (int q = 0; q < j; q++) if ((int) m[MC_NUC0 + q] == (int) m[MC_NUC0 + j]) S->mol_repeats = true;
    }
    if (m[24] != 0.0) S->iH2 = i;
    if (m[25] != 0.0) S->iOH = i;
    if (m[26] != 0.0) S->iCH = i;
  }
  S->nnuc = nnuclei; S->nmol = nmol;
  RH_CHECK(S->H.put(&S->d_nuc_atom, nucleus_atom, (size_t) nnuclei));
  RH_CHECK(S->H.put(&S->d_mol, mol, (size_t) nmol * MC_NFIELD));
  return RHB200_OK;
}

// unit-level: LTE populations + chemical equilibrium of ncol columns; chem [ncol][natom+4][ndep], pops [ncol][nlev][ndep]
extern "C" int rhb200_chemistry_batch(rhb200_ctx *c, int ncol, int ndep, const double *atmos, double *chem, double *pops)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  ContinuumState *S = (ContinuumState *) c->cont;
  if (!S || S->nmol == 0) { rhb200_set_error("rhb200_set_continuum() / rhb200_set_chemistry() have not been called"); return RHB200_ESTATE; }
  if (ncol <= 0 || ndep <= 0 || !atmos || !chem) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  Holder H;
  const size_t cn = (size_t) ncol * ndep;
  double *d_at, *d_ch, *d_pp;
  RH_CHECK(H.put(&d_at, atmos, cn * RHB200_AT_NFIELD));
  RH_CUDA(cudaMalloc((void **) &d_ch, cn * (S->natom + 4) * sizeof(double))); H.p.push_back(d_ch);
  RH_CUDA(cudaMalloc((void **) &d_pp, cn * S->nlev * sizeof(double))); H.p.push_back(d_pp);
  ltepops_kernel<<<(unsigned) ((cn + 127) / 128), 128, 0, c->stream>>>(ncol, ndep, S->natom, S->nlev, S->d_lev, S->d_first,
                                                                       S->d_abund, d_at, nullptr, d_pp, nullptr);
  RH_CHECK(launch_chemeq(c, S, ncol, ndep, d_at, d_pp, d_ch));
  RH_CUDA(cudaGetLastError());
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CUDA(cudaMemcpy(chem, d_ch, cn * (S->natom + 4) * sizeof(double), cudaMemcpyDeviceToHost));
  if (pops) RH_CUDA(cudaMemcpy(pops, d_pp, cn * S->nlev * sizeof(double), cudaMemcpyDeviceToHost));
  return RHB200_OK;
}

extern "C" int rhb200_continuum_batch(rhb200_ctx *c, const rhb200_continuum_model *m, int nlambda, const double *lambda,
                                      int ncol, int ndep, const double *T, const double *ne, const double *nHmin,
                                      const double *nH2, const double *nOH, const double *nCH,
                                      const double *pops_n, const double *pops_nstar,
                                      double *chi_ai, double *eta_ai, double *sca_ai, double *contrib)
{
  if (!c) { rhb200_set_error("null context"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  RH_CHECK(check_model(m));
  if (nlambda <= 0 || ncol <= 0 || ndep <= 0 || !lambda || !T || !ne || !nHmin || !pops_n || !pops_nstar || !chi_ai || !eta_ai ||
      (m->has_H2 && !nH2) || (m->has_OH && !nOH) || (m->has_CH && !nCH)) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  Holder H;
  const size_t cn = (size_t) ncol * ndep, pn = (size_t) ncol * m->nlev * ndep, on = (size_t) ncol * nlambda * ndep;
  double *dT, *dne, *dHm, *dH2 = nullptr, *dOH = nullptr, *dCH = nullptr, *dn, *ds, *dchi, *deta, *dsca;
  RH_CHECK(H.put(&dT, T, cn)); RH_CHECK(H.put(&dne, ne, cn)); RH_CHECK(H.put(&dHm, nHmin, cn));
  if (m->has_H2) RH_CHECK(H.put(&dH2, nH2, cn));
  if (m->has_OH) RH_CHECK(H.put(&dOH, nOH, cn));
  if (m->has_CH) RH_CHECK(H.put(&dCH, nCH, cn));
  RH_CHECK(H.put(&dn, pops_n, pn));
  if (pops_nstar == pops_n) ds = dn; else RH_CHECK(H.put(&ds, pops_nstar, pn));
  RH_CUDA(cudaMalloc((void **) &dchi, on * sizeof(double))); H.p.push_back(dchi);
  RH_CUDA(cudaMalloc((void **) &deta, on * sizeof(double))); H.p.push_back(deta);
  RH_CUDA(cudaMalloc((void **) &dsca, on * sizeof(double))); H.p.push_back(dsca);
  double *dcon = nullptr;
  if (contrib) { RH_CUDA(cudaMalloc((void **) &dcon, on * 26 * sizeof(double))); H.p.push_back(dcon); RH_CUDA(cudaMemset(dcon, 0, on * 26 * sizeof(double))); }
  RH_CHECK(rh_continuum_dev(c, m, nlambda, lambda, ncol, ndep, dT, dne, dHm, dH2, dOH, dCH, dn, ds, dchi, deta, dsca, dcon));
  if (contrib) RH_CUDA(cudaMemcpy(contrib, dcon, on * 26 * sizeof(double), cudaMemcpyDeviceToHost));
  RH_CUDA(cudaMemcpy(chi_ai, dchi, on * sizeof(double), cudaMemcpyDeviceToHost));
  RH_CUDA(cudaMemcpy(eta_ai, deta, on * sizeof(double), cudaMemcpyDeviceToHost));
  if (sca_ai) RH_CUDA(cudaMemcpy(sca_ai, dsca, on * sizeof(double), cudaMemcpyDeviceToHost));
  return RHB200_OK;
}
