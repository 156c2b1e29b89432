// rhb200_common.cuh -- shared declarations of librhb200.so (sm_100a).
// All device arithmetic is compiled with -fmad=false: the reference x86-64
// build has no FMA contraction, and parity is bit-for-bit where the reference
// is IEEE-deterministic (see DESIGN.md "Arithmetic contract").
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/rhb200.h"

// physical constants: rh/constant.h:28-73 (same literals, same expressions)
#define RH_CLIGHT      2.99792458E+08
#define RH_HPLANCK     6.6260755E-34
#define RH_KBOLTZMANN  1.380658E-23
#define RH_AMU         1.6605402E-27
#define RH_M_ELECTRON  9.1093897E-31
#define RH_Q_ELECTRON  1.60217733E-19
#define RH_NM_TO_M     1.0E-09
#define RH_PI          3.14159265358979
#define RH_SQRTPI      1.77245385090551
// constant.h:68 -- the macro has no outer parentheses in the reference: keep it so
#define RH_LARMOR      (RH_Q_ELECTRON / (4.0*RH_PI*RH_M_ELECTRON)) * RH_NM_TO_M
#define RH_Q_WING            20.0   // kurucz.c:89
#define RH_MAX_GAUSS_DOPPLER 7.0    // kurucz.c:92

void rhb200_set_error(const char *fmt, ...);

#define RH_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      rhb200_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,            \
                       cudaGetErrorString(e__));                               \
      return RHB200_ECUDA;                                                     \
    }                                                                          \
  } while (0)

// per-(column, line, depth) quantities that do not depend on wavelength
// (hoisted out of RLKProfile / rlk_opacity, kurucz.c:669-691,744-783)
enum { LP_VBROAD = 0, LP_ADAMP, LP_VB, LP_W, LP_SV, LP_CHIL, LP_ETAL, LP_EPS /* RLK_SCATTER: destruction probability */, LP_NFIELD = 8 };
// ray-point record handed from the opacity kernel to the DELO kernel
enum { RP_CHI = 0, RP_KQ, RP_KU, RP_KV, RP_SI, RP_SQ, RP_SU, RP_SV, RP_NFIELD };

struct DevTables {
  int nline = 0, ncomp = 0, nelem = 0, npf_rows = 0, npf = 0;
  double *lines = nullptr, *zshift = nullptr, *zstrength = nullptr, *elems = nullptr,
         *pf = nullptr, *Tpf = nullptr;
  int *zq = nullptr;
  double vmicro_char = 0.0;
  int rlkscatter = 0;        // keyword RLK_SCATTER (kurucz.c:641-652, 682-694)
};

struct DevWave {
  int nlambda = 0, nidx = 0;
  double *lambda = nullptr;   // [nlambda]
  int *first = nullptr, *count = nullptr, *idx = nullptr, *flags = nullptr;
  // wavelengths without a POLARISED line (flags bit 1 clear): solved for I alone (formal.c:84-103, 223-236, 289-309)
  // -- Feautrier when there is no line at all or the column is static, else the scalar S_INTERPOLATION ray
  int *noline = nullptr, nnoline = 0;
  int *unpol_rank = nullptr, nunpol = 0;
  // passive_bb lines whose windows touch the grid (rhb200_set_passive_lines): compact tables of the ACTIVE subset
  int npl = 0, npw = 0;
  int *pw_first = nullptr, *pw_count = nullptr, *pw_idx = nullptr;
  // molecular lines of PASSIVE molecules (rhb200_set_molecular_lines): all lines on the device, windows per wavelength
  int nml = 0, nmsel = 0, nmw = 0;
  int *mw_first = nullptr, *mw_count = nullptr, *mw_idx = nullptr;
  double *ml_rows = nullptr /* [nml][RHB200_ML_NFIELD] */, *ml_sel = nullptr /* [nmsel][16] */;
  int mol_pol = 0;                                     // a polarizable molecular line (MolZeeman pattern) lies in some window
  int *mz_q = nullptr; double *mz_shift = nullptr, *mz_strength = nullptr;   // Zeeman components of the molecular lines
  double *pl_rows = nullptr /* [npl][RHB200_PL_NFIELD] */, *pl_pb = nullptr /* [npl][RHB200_PB_NFIELD] */, *pl_cshift = nullptr, *pl_cfrac = nullptr;   // [nlambda] rank among the wavelengths with flags == 1 (line, unpolarised), else -1
};

struct KTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

struct rhb200_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  DevTables tab;
  DevWave wav;
  std::vector<double> h_lines, h_elems, h_lambda, h_zshift, h_zstrength;
  std::vector<int> h_zq;
  std::vector<int> h_first, h_count, h_idx, h_flags, h_noline;
  std::vector<double> h_model_lines;
  std::vector<double> h_plines, h_pcshift, h_pcfrac;   // rhb200_set_passive_lines
  std::vector<int> h_mzq; std::vector<double> h_mzshift, h_mzstrength;   // MolZeeman components (rhb200_set_molecular_lines_zeeman)
  std::vector<double> h_mlines, h_msel;                // rhb200_set_molecular_lines     // rhb200_set_model_lines: [n][4] element row, stage, lambda0 [nm], qwing
  // formal solver selection (keyword.input S_INTERPOLATION / S_INTERPOLATION_STOKES, inputs.h:26-27)
  int s_interpolation = RHB200_S_BEZIER3, s_interpolation_stokes = RHB200_DELO_BEZIER3;
  int no_stokes = 0;         // STOKES_MODE = NO_STOKES (rhb200_set_stokes_mode)
  int n_max_scatter = 0; double scatter_limit = 1.0e-2;   // N_MAX_SCATTER / ITER_LIMIT in LTE (rhb200_set_scatter)
  // analytic log gf response functions (rhb200_set_loggf_rf): parameter p belongs to row lrf_lines[p] of the line table
  int lrf_npar = 0; int *d_lrf_lines = nullptr;
  // doubles per depth point of the scalar-ray scratch: chi, S, I (+ dchi, deta, dI[npar] in RF mode)
  int scal_fields() const { return 3 + 3*lrf_npar; }
  // NLTE rate accumulation: 0 = fixed-partition two-stage reduction (default), 1 = the reference's add order, bit-identical
  int nlte_exact_rates = 0;
  double total_abund = 0.0, gravity = 0.0;      // rhb200_set_gravity: the column mass of a height scale (multiatmos.c:153-155)
  // wavelength shard of the NLTE solve (rhb200_nlte_set_shard)
  int shard_rank = 0, shard_nrank = 1;
  rhb200_allreduce_fn shard_fn = nullptr; void *shard_user = nullptr;
  bool nccl_owned = false;
  void *nccl_comm = nullptr;                     // rhb200_nlte_set_shard_nccl: ncclAllReduce on the compute stream, no host sync
  void *elements = nullptr;  // all elements + partition functions (rhb200_set_elements), owned by rhb200_hse.cu
  void *cont = nullptr;      // background-continuum state (rhb200_set_continuum), owned by rhb200_continuum.cu
  void *nlte_front = nullptr;   // plans / engines / work arrays of rhb200_nlte_compute1d_batch, owned by rhb200_nlte.cu
  // workspace (grown on demand)
  void *ws = nullptr; size_t ws_bytes = 0;
  void *flush = nullptr; size_t flush_bytes = 0;
  // instrumentation
  bool timing = false;
  double k_ms[RHB200_K_COUNT] = {0};
  long k_launch[RHB200_K_COUNT] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

// timed launch helper: records CUDA events around the launch on ctx->stream when
// instrumentation is on (synchronising: only used for measurement runs)
// NVTX range names = the labels the reference passes to getCPU() for the same work (rh/getcpu.c is stubbed out in pyrh;
// examples/time.out shows them): a timeline in Nsight Systems reads like the reference's time.out
static const char *const RH_NVTX_LABEL[RHB200_K_COUNT] = {
  "Read Atmosphere / LTE populations / line prep", "Background Opacity", "Spectrum & Operator (Stokes ray)",
  "Spectrum & Operator (scalar ray)", "Background continuum / scales", "Spectrum & Operator (Gamma, rates)",
  "Spectrum & Operator (J)", "Populations (statEquil)", "Populations (Ng)"};
struct RhRange {                     // phase-level range (rhf1d()'s getCPU(1..2, ...) labels)
  explicit RhRange(const char *label) { nvtxRangePushA(label); }
  ~RhRange() { nvtxRangePop(); }
};

struct ScopedKernelTimer {
  rhb200_ctx *c; int which;
  ScopedKernelTimer(rhb200_ctx *ctx, int w) : c(ctx), which(w) {
    nvtxRangePushA(RH_NVTX_LABEL[w]);
    if (c->timing) cudaEventRecord(c->ev0, c->stream);
  }
  ~ScopedKernelTimer() {
    nvtxRangePop();
    c->k_launch[which] += 1;
    if (c->timing) {
      cudaEventRecord(c->ev1, c->stream);
      cudaEventSynchronize(c->ev1);
      float ms = 0.f; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
      c->k_ms[which] += ms;
    }
  }
};

int rh_ws_reserve(rhb200_ctx *ctx, size_t bytes);
void rh_continuum_free(rhb200_ctx *ctx);
void rh_nlte_front_free(rhb200_ctx *ctx);
void rh_elements_free(rhb200_ctx *ctx);
int rh_continuum_nlev(const rhb200_ctx *ctx);
int rh_continuum_natom(const rhb200_ctx *ctx);
int rh_continuum_proton_level(const rhb200_ctx *ctx);
void rh_continuum_set_hse_mode(rhb200_ctx *ctx, int on);
int rh_continuum_nlambda(const rhb200_ctx *ctx);
int rh_continuum_has_chemistry(const rhb200_ctx *ctx);
int rh_launch_pyrh_rows(rhb200_ctx *ctx, int ncol, int ndep, int nrow_in, int atm_scale, double muz, double vmacro_tresh,
                        const double *d_in, double *d_atmos, int *d_col_moving);
int rh_launch_rf_expand(rhb200_ctx *ctx, int v0, int n, int ndep, int nrow, int npar, const int *d_rows,
                        const double *d_delta, const double *d_base, double *d_in, int nsel, const int *d_sel);
int rh_launch_rf_diff(rhb200_ctx *ctx, int v0, int n, int nsel, int nlambda, int npar, const double *d_delta,
                      const double *d_stokes, double *d_rf);
int rh_launch_proton(rhb200_ctx *ctx, int ncol, int ndep, int nlev, int proton_level, const double *d_pops, double *d_atmos);
int rh_launch_scales(rhb200_ctx *ctx, int ncol, int ndep, int iref, int atm_scale, double wght_per_H,
                     double total_abund, double gravity, const double *d_raypts, double *d_atmos, double *d_scratch, double *d_scales_out);
int rh_continuum_chunk(rhb200_ctx *ctx, int cc, int ndep, const double *d_atmos, const double *d_chem,
                       double *d_pops, double *d_tprep, double *d_chi, double *d_eta, int chem_on_device,
                       double *d_molout = nullptr, double *d_sca = nullptr);
int rh_nccl_allreduce(rhb200_ctx *ctx, double *buf, size_t count, int op);
int rh_nccl_group(int end);
void rh_nccl_release(rhb200_ctx *ctx);
int rh_continuum_set_molsel(rhb200_ctx *ctx, int nsel, const int *chem_index);
int rh_continuum_ltepops(rhb200_ctx *ctx, int cc, int ndep, const double *d_atmos, const double *d_chem_host, double *d_pops,
                         const double *d_ntot_in);
int rh_continuum_chemeq(rhb200_ctx *ctx, int cc, int ndep, const double *d_atmos, double *d_pops, double *d_chem,
                        double *d_molout, double *d_ntot, int ntot_is_input);
int rh_continuum_opac(rhb200_ctx *ctx, int cc, int ndep, const double *d_atmos, const double *d_chem, const double *d_pops_n,
                      const double *d_pops_star, double *d_tprep, double *d_chi, double *d_eta, double *d_sca);
int rh_continuum_atom_first(const rhb200_ctx *ctx, int atom);
double rh_continuum_abundance(const rhb200_ctx *ctx, int atom);
int rh_launch_scales_chi(rhb200_ctx *ctx, int ncol, int ndep, int nlambda, int iref, int atm_scale, double wght_per_H,
                         double total_abund, double gravity,
                         const double *d_chi_c, double *d_atmos, double *d_scratch, double *d_scales_out);

// launchers implemented in the .cu files (device pointers)
int rh_launch_mol_opacity_raw(rhb200_ctx *ctx, int ncol, int nlambda, int ndep, int nmol, double muz, int moving,
                              int to_obs, const double *d_lambda, const int *d_first, const int *d_count,
                              const int *d_idx, const double *d_mlines, const int *d_zq, const double *d_zshift,
                              const double *d_zstrength, const double *d_atmos, const double *d_mol,
                              double *d_chi, double *d_eta);
int rh_launch_passive_bb(rhb200_ctx *ctx, int ncol, int nlambda, int ndep, int nline, double muz, int moving,
                         int to_obs, const double *d_lambda, const int *d_first, const int *d_count,
                         const int *d_idx, const double *d_plines, const double *d_cshift, const double *d_cfrac,
                         const double *d_atmos, const double *d_pcol, double *d_chi, double *d_eta);
int rh_passive_chunk(rhb200_ctx *ctx, int ncol, int ndep, double muz, const double *d_atmos, const double *d_pops, int nlev,
                      double *d_pcol /* [ncol][npl][4][ndep] */, double *d_chi_ai, double *d_eta_ai);
int rh_molecular_chunk(rhb200_ctx *ctx, int ncol, int ndep, double muz, const double *d_atmos, const double *d_molden,
                        double *d_mol /* [ncol][nmsel][3][ndep] */, double *d_molchi, double *d_moleta /* [ncol][nlambda][ndep], or [4] of those planes (I, Q, U, V) when wav.mol_pol */);
int rh_launch_prep(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                   const double *d_atmos, double *d_elem_n, double *d_lineprep);
int rh_launch_opacity_fused(rhb200_ctx *ctx, int ncol, int ndep, int to_obs,
                            const double *d_atmos, const double *d_lineprep,
                            const double *d_chi_ai, const double *d_eta_ai,
                            double *d_raypts /* [nray][ndep][RP_NFIELD] */,
                            const double *d_molchi = nullptr, const double *d_moleta = nullptr, const double *d_sca = nullptr);
int rh_launch_loggf_dopac(rhb200_ctx *ctx, int ncol, int ndep, const double *d_atmos, const double *d_lineprep,
                          double *d_scratch /* [ncol][nunpol][scal_fields()][ndep] */);
int rh_launch_loggf_rf(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom, const double *d_atmos,
                       int moving, const int *d_col_moving,
                       double *d_scratch, double *d_rf /* [ncol][nlambda][npar] */);
int rh_launch_opacity_addI(rhb200_ctx *ctx, int ncol, int ndep, int to_obs, const double *d_atmos, const double *d_lineprep,
                           double *d_chi_c, double *d_eta_c, double *d_chi_quv = nullptr, double *d_eta_quv = nullptr);
int rh_launch_add_molecular(rhb200_ctx *ctx, int ncol, int ndep, const double *d_molchi, const double *d_moleta,
                            double *d_chi_c, double *d_eta_c);
int rh_launch_line_damping(rhb200_ctx *ctx, int ncol, int ndep, int nline, const double *d_plrows, const double *d_atmos,
                           const double *d_pops, int nlev, double *d_pcol, double *d_qelast = nullptr);
int rh_launch_opacity_raw(rhb200_ctx *ctx, int ncol, int ndep, int to_obs,
                          const double *d_atmos, const double *d_lineprep,
                          double *d_chi, double *d_eta /* [ncol][nlambda][4][ndep] */);
int rh_launch_delo_raypts(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom,
                          const double *d_atmos, const double *d_raypts,
                          double *d_stokes /* [ncol][4][nlambda] */);
// single-depth finite-difference columns (rhb200_rf_fd_batch): see vscales_kernel in rhb200_scales.cu
int rh_launch_rf_expand_full(rhb200_ctx *ctx, int b0, int nb, int ndep, int nrow, int npar, const int *d_rows,
                             const double *d_delta, const double *d_base, double *d_in);
int rh_launch_vscales(rhb200_ctx *ctx, int nb, int npar, int ndep, int iref, int atm_scale, double wght_per_H,
                      double total_abund, double gravity, const double *d_raypts, const double *d_atmos,
                      double *d_vws /* [nv][4][ndep] */, int nsel, const int *d_sel);
int rh_launch_delo_vcols(rhb200_ctx *ctx, int nb, int npar, int ndep, double muz, int bc_top, int bc_bottom,
                         const double *d_vws, const double *d_raypts, double *d_stokes /* [nv][4][nlambda] */,
                         const int *d_neutral /* [npar] or NULL */, int pn /* a neutral parameter or -1 */,
                         double *d_state /* [nb][nlambda][ndep][DELO_NSTATE] or NULL */, int nsel, const int *d_sel);
int rh_launch_noline_vcols(rhb200_ctx *ctx, int nb, int npar, int ndep, double muz, int bc_top, int bc_bottom,
                           const double *d_vws, const double *d_raypts, double *d_stokes,
                           double *d_scratch /* [nv][nnoline][5][ndep] */, int nsel, const int *d_sel);
int rh_launch_delo_generic(rhb200_ctx *ctx, int solver /* RHB200_DELO_* */, int nray, int ndep, double muz, int to_obs,
                           int bc_top, int bc_bottom, const int *d_ray_col,
                           const double *d_ray_lambda, const double *d_height, const double *d_T,
                           const double *d_chi, const double *d_S, const double *d_chiQUV,
                           double *d_I, double *d_Psi);
int rh_launch_bezier3_rf(rhb200_ctx *ctx, int nray, int ndep, double muz, int bc_top, int bc_bottom,
                         const int *d_ray_col, const double *d_ray_lambda, const double *d_height, const double *d_T,
                         const double *d_chi_dn, const double *d_S_dn, const double *d_chi_up, const double *d_S_up,
                         int npar, const double *d_dchi, const double *d_deta, double *d_I, double *d_dI);
int rh_launch_bezier3(rhb200_ctx *ctx, int solver /* RHB200_S_* */, int nray, int ndep, double muz, int to_obs,
                      int bc_top, int bc_bottom, const int *d_ray_col,
                      const double *d_ray_lambda, const double *d_height, const double *d_T,
                      const double *d_chi, const double *d_S, double *d_I, double *d_Psi);
int rh_launch_feautrier_raypts(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom,
                               const double *d_atmos, double *d_raypts, double *d_stokes,
                               int moving, const int *d_col_moving, double *d_scratch /* [ncol][nunpol][3][ndep] */);
int rh_launch_scatter_passes(rhb200_ctx *ctx, int ncol, int ndep, double muz, int bc_top, int bc_bottom, const double *d_atmos,
                             double *d_raypts, double *d_stokes, int moving, const int *d_col_moving,
                             unsigned long long *d_colmax /* [ncol] */, int *d_done /* [ncol] */);
int rh_launch_feautrier(rhb200_ctx *ctx, int nray, int ndep, double muz, int bc_top, int bc_bottom,
                        const int *d_ray_col, const double *d_ray_lambda, const double *d_height,
                        const double *d_T, const double *d_chi, const double *d_S,
                        double *d_P, double *d_Psi, double *d_Iem, double *d_scratch);
int rh_launch_voigt(rhb200_ctx *ctx, int n, const double *d_a, const double *d_v,
                    double *d_H, double *d_F, int *d_region);
int rh_launch_voigt_armstrong(rhb200_ctx *ctx, int n, const double *d_a, const double *d_v,
                              double *d_H, int *d_region);
int rh_launch_math_probe(rhb200_ctx *ctx, int n, int func, const double *d_x, const double *d_y,
                         double *d_out);
int rh_fp64_peak(rhb200_ctx *ctx, double *tf_fma, double *tf_nofma);
