// rhb200_abi.cu -- the C ABI of librhb200.so (include/rhb200.h): context management,
// shared tables, the batched LTE Stokes pipeline and the function-level entry points.
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <mutex>
#include <thread>
#include "rhb200_common.cuh"

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";

void rhb200_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *rhb200_last_error(void) { return g_err; }
extern "C" int rhb200_version(void) { return RHB200_VERSION; }

extern "C" int rhb200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int rhb200_device_info(int device, char *name, int name_len, int *sm_count,
                                  size_t *mem_bytes, int *cc_major, int *cc_minor)
{
  cudaDeviceProp p;
  RH_CUDA(cudaGetDeviceProperties(&p, device));
  if (name && name_len > 0) { strncpy(name, p.name, name_len-1); name[name_len-1] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (mem_bytes) *mem_bytes = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return RHB200_OK;
}

// ----------------------------------------------------------------- context
extern "C" rhb200_ctx *rhb200_open(int device)
{
  int n = rhb200_device_count();
  if (n <= 0) { rhb200_set_error("no CUDA device visible (librhb200 has no CPU fallback)"); return nullptr; }
  if (device < 0 || device >= n) { rhb200_set_error("device %d out of range (%d visible)", device, n); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { rhb200_set_error("cudaSetDevice(%d) failed", device); return nullptr; }
  rhb200_ctx *c = new rhb200_ctx();
  c->device = device;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, device);
  c->sm_count = p.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    rhb200_set_error("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  return c;
}

static void free_tables(rhb200_ctx *c)
{
  DevTables &t = c->tab;
  cudaFree(t.lines); cudaFree(t.zshift); cudaFree(t.zstrength); cudaFree(t.elems);
  cudaFree(t.pf); cudaFree(t.Tpf); cudaFree(t.zq);
  t = DevTables();
  cudaFree(c->d_lrf_lines); c->d_lrf_lines = nullptr; c->lrf_npar = 0;      // rows of the table that is going away
}
static void free_wave(rhb200_ctx *c)
{
  DevWave &w = c->wav;
  cudaFree(w.lambda); cudaFree(w.first); cudaFree(w.count); cudaFree(w.idx); cudaFree(w.flags); cudaFree(w.noline); cudaFree(w.unpol_rank);
  cudaFree(w.mw_first); cudaFree(w.mw_count); cudaFree(w.mw_idx); cudaFree(w.ml_rows); cudaFree(w.ml_sel);
  cudaFree(w.mz_q); cudaFree(w.mz_shift); cudaFree(w.mz_strength);
  cudaFree(w.pw_first); cudaFree(w.pw_count); cudaFree(w.pw_idx); cudaFree(w.pl_rows); cudaFree(w.pl_pb); cudaFree(w.pl_cshift); cudaFree(w.pl_cfrac);
  w = DevWave();
}

extern "C" void rhb200_close(rhb200_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_tables(c); free_wave(c); rh_continuum_free(c); rh_elements_free(c); rh_nlte_front_free(c); rh_nccl_release(c);
  cudaFree(c->ws); cudaFree(c->flush);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c;
}

int rh_ws_reserve(rhb200_ctx *c, size_t bytes)
{
  if (bytes <= c->ws_bytes) return RHB200_OK;
  if (c->ws) { RH_CUDA(cudaFree(c->ws)); c->ws = nullptr; c->ws_bytes = 0; }
  cudaError_t e = cudaMalloc(&c->ws, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    rhb200_set_error("workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return RHB200_ENOMEM;
  }
  c->ws_bytes = bytes;
  return RHB200_OK;
}

template <class T>
static int upload(T **dptr, const T *h, size_t n)
{
  *dptr = nullptr;
  if (n == 0) return RHB200_OK;
  RH_CUDA(cudaMalloc((void **) dptr, n * sizeof(T)));
  RH_CUDA(cudaMemcpy(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return RHB200_OK;
}

#define RH_CHECK(expr) do { int rc__ = (expr); if (rc__ != RHB200_OK) return rc__; } while (0)
#define RH_NEED_CTX(c) do { if (!(c)) { rhb200_set_error("null context"); return RHB200_EINVAL; } \
                            RH_CUDA(cudaSetDevice((c)->device)); } while (0)

extern "C" int rhb200_set_lines(rhb200_ctx *c, int nline, const double *lines, int ncomp,
                                const int *zq, const double *zshift, const double *zstrength,
                                int nelem, const double *elems, int npf_rows, int npf,
                                const double *pf, const double *Tpf, double vmicro_char,
                                int magneto_optical, int rlkscatter)
{
  RH_NEED_CTX(c);
  if (magneto_optical) { rhb200_set_error("MAGNETO_OPTICAL = TRUE is not implemented (the reference overflows chip_c there, readj.c:328)"); return RHB200_EUNSUPPORTED; }

  if (nline < 0 || nelem < 0 || npf < 2 || (nline > 0 && (!lines || !elems || !pf || !Tpf))) {
    rhb200_set_error("rhb200_set_lines: bad arguments"); return RHB200_EINVAL;
  }
  for (int n = 0; n < nline; n++) {
    const double *L = lines + (size_t) n * RHB200_RL_NFIELD;
    if (n > 0 && L[RHB200_RL_LAMBDA0] < L[RHB200_RL_LAMBDA0 - RHB200_RL_NFIELD]) {
      rhb200_set_error("line table not sorted by lambda0 at row %d (background.c:292-294)", n); return RHB200_EINVAL;
    }
    const int ie = (int) L[RHB200_RL_ELEM];
    if (ie < 0 || ie >= nelem) { rhb200_set_error("line %d: element row %d out of range", n, ie); return RHB200_EINVAL; }
    const int zo = (int) L[RHB200_RL_ZOFF], nc = (int) L[RHB200_RL_NCOMP];
    if (zo < 0 || nc < 0 || zo + nc > ncomp) { rhb200_set_error("line %d: Zeeman slice out of range", n); return RHB200_EINVAL; }
  }
  for (int e = 0; e < nelem; e++) {
    const double *E = elems + (size_t) e * RHB200_RE_NFIELD;
    const int nst = (int) E[RHB200_RE_NSTAGE], row = (int) E[RHB200_RE_PFROW];
    if (nst < 1 || nst > RHB200_RE_MAXSTAGE || row < 0 || row + nst > npf_rows) {
      rhb200_set_error("element row %d: Nstage/pf rows out of range", e); return RHB200_EINVAL;
    }
  }
  if (Tpf && !(Tpf[1] > Tpf[0])) { rhb200_set_error("Tpf must ascend"); return RHB200_EINVAL; }
  free_tables(c);
  DevTables &t = c->tab;
  t.nline = nline; t.ncomp = ncomp; t.nelem = nelem; t.npf_rows = npf_rows; t.npf = npf;
  t.vmicro_char = vmicro_char;
  t.rlkscatter = rlkscatter ? 1 : 0;
  RH_CHECK(upload(&t.lines, lines, (size_t) nline * RHB200_RL_NFIELD));
  RH_CHECK(upload(&t.zq, zq, (size_t) ncomp));
  RH_CHECK(upload(&t.zshift, zshift, (size_t) ncomp));
  RH_CHECK(upload(&t.zstrength, zstrength, (size_t) ncomp));
  RH_CHECK(upload(&t.elems, elems, (size_t) nelem * RHB200_RE_NFIELD));
  RH_CHECK(upload(&t.pf, pf, (size_t) npf_rows * npf));
  RH_CHECK(upload(&t.Tpf, Tpf, (size_t) npf));
  c->h_lines.assign(lines, lines + (size_t) nline * RHB200_RL_NFIELD);
  c->h_elems.assign(elems, elems + (size_t) nelem * RHB200_RE_NFIELD);
  c->h_zq.assign(zq, zq + ncomp); c->h_zshift.assign(zshift, zshift + ncomp); c->h_zstrength.assign(zstrength, zstrength + ncomp);
  free_wave(c);     // windows depend on the line table
  c->h_model_lines.clear();
  c->h_plines.clear(); c->h_pcshift.clear(); c->h_pcfrac.clear();
  c->h_mlines.clear(); c->h_msel.clear(); c->h_mzq.clear(); c->h_mzshift.clear(); c->h_mzstrength.clear();
  return RHB200_OK;
}

// In-place update of line strengths (log gf overrides of pyrh.compute1d, kurucz.c:247-257): the windows, Zeeman patterns
// and every other table stay; only Aji / Bji / Bij of the named rows change, on the host copy and on the device.
extern "C" int rhb200_update_line_strengths(rhb200_ctx *c, int n, const int *rows, const double *Aji, const double *Bji,
                                            const double *Bij)
{
  RH_NEED_CTX(c);
  if (n < 0 || (n > 0 && (!rows || !Aji || !Bji || !Bij))) { rhb200_set_error("rhb200_update_line_strengths: bad arguments"); return RHB200_EINVAL; }
  for (int i = 0; i < n; i++)
    if (rows[i] < 0 || rows[i] >= c->tab.nline) { rhb200_set_error("line row %d out of range", rows[i]); return RHB200_EINVAL; }
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CUDA(cudaStreamSynchronize(c->copy_stream));
  for (int i = 0; i < n; i++) {
    double *L = c->h_lines.data() + (size_t) rows[i] * RHB200_RL_NFIELD;
    L[RHB200_RL_AJI] = Aji[i]; L[RHB200_RL_BJI] = Bji[i]; L[RHB200_RL_BIJ] = Bij[i];
    // BJI, AJI, BIJ are adjacent fields of the row
    RH_CUDA(cudaMemcpy(c->tab.lines + (size_t) rows[i] * RHB200_RL_NFIELD + RHB200_RL_BJI, L + RHB200_RL_BJI, 3 * sizeof(double),
                       cudaMemcpyHostToDevice));
  }
  return RHB200_OK;
}

// MolecularOpacity in the fused LTE path (opacity.c:711-839): LTE lines of PASSIVE molecules.
// mlines [nline][RHB200_ML_NFIELD] grouped by molecule, ascending in lambda0 inside each (RHB200_ML_MOL = row of
// `molecules`); molecules [nmol][16] = {index in the chemical network of rhb200_set_chemistry, molecular weight,
// enum fit_type, Tmin, Tmax, Npf, pf_coef[0..7]} (readmolecule.c:199-237).  Polarizable lines (Hund's-case data in the
// list, readmolecule.c:859-912) carry RHB200_ML_POLARIZABLE != 0 and the range [RHB200_ML_ZOFF, + RHB200_ML_NCOMP) of
// their MolZeeman components (molzeeman.c:196-319) in zq / zshift / zstrength.  Needs rhb200_set_continuum and
// rhb200_set_chemistry before the first batch call; call before rhb200_set_wavelengths; rhb200_set_lines clears the table.
extern "C" int rhb200_set_molecular_lines_zeeman(rhb200_ctx *c, int nline, const double *mlines, int nmol, const double *molecules,
                                                 int ncomp, const int *zq, const double *zshift, const double *zstrength)
{
  RH_NEED_CTX(c);
  if (nline < 0 || nmol < 0 || ncomp < 0 || (nline > 0 && (!mlines || !molecules || nmol == 0)) ||
      (ncomp > 0 && (!zq || !zshift || !zstrength))) { rhb200_set_error("rhb200_set_molecular_lines: bad arguments"); return RHB200_EINVAL; }
  std::vector<int> chem(nmol);
  for (int m = 0; m < nmol; m++) {
    chem[m] = (int) molecules[(size_t) m * 16];
    if ((int) molecules[(size_t) m * 16 + 5] > 8) { rhb200_set_error("molecule %d: more than 8 partition-function coefficients", m); return RHB200_EINVAL; }
  }
  for (int n = 0; n < nline; n++) {
    const double *L = mlines + (size_t) n * RHB200_ML_NFIELD;
    const int m = (int) L[RHB200_ML_MOL];
    if (m < 0 || m >= nmol) { rhb200_set_error("molecular line %d: molecule index out of range", n); return RHB200_EINVAL; }
    if (n > 0 && m < (int) L[RHB200_ML_MOL - RHB200_ML_NFIELD]) { rhb200_set_error("molecular lines must be grouped by molecule"); return RHB200_EINVAL; }
    if (L[RHB200_ML_POLARIZABLE] != 0.0) {
      const int zoff = (int) L[RHB200_ML_ZOFF], nc = (int) L[RHB200_ML_NCOMP];
      if (zoff < 0 || nc < 0 || zoff + nc > ncomp) { rhb200_set_error("molecular line %d is polarizable: its MolZeeman components [%d, %d) are outside the tables (%d)", n, zoff, zoff + nc, ncomp); return RHB200_EINVAL; }
    }
  }
  c->h_mlines.assign(mlines, mlines + (size_t) nline * RHB200_ML_NFIELD);
  c->h_msel.assign(molecules, molecules + (size_t) nmol * 16);
  c->h_mzq.assign(zq, zq + ncomp); c->h_mzshift.assign(zshift, zshift + ncomp); c->h_mzstrength.assign(zstrength, zstrength + ncomp);
  free_wave(c);
  return RHB200_OK;
}
extern "C" int rhb200_set_molecular_lines(rhb200_ctx *c, int nline, const double *mlines, int nmol, const double *molecules)
{
  return rhb200_set_molecular_lines_zeeman(c, nline, mlines, nmol, molecules, 0, nullptr, nullptr, nullptr);
}

// keywords N_MAX_SCATTER / ITER_LIMIT in LTE (pyrh_compute1dray.c:332-337): after the formal solution the reference
// Lambda-iterates the continuum-scattering term of the angle-independent wavelengths, S = (eta + sca J)/chi, until
// max |1 - Jdag/J| <= iter_limit or n_max_scatter passes.  0 (default) = the single pass of Iterate().  Needs the
// continuum on the device (the scattering opacity comes from there).
extern "C" int rhb200_set_scatter(rhb200_ctx *c, int n_max_scatter, double iter_limit)
{
  RH_NEED_CTX(c);
  if (n_max_scatter < 0 || !(iter_limit >= 0.0)) { rhb200_set_error("rhb200_set_scatter: bad arguments"); return RHB200_EINVAL; }
  c->n_max_scatter = n_max_scatter; c->scatter_limit = iter_limit;
  return RHB200_OK;
}

// keyword STOKES_MODE (inputs.h enum StokesMode): FULL_STOKES solves I, Q, U, V where a polarised line is present;
// NO_STOKES solves I alone everywhere (Q = U = V = 0) -- the line profiles are the same Zeeman-split ones either way,
// because pyrh always sets atmos.Stokes (pyrh_compute1dray.c:259).  Call before rhb200_set_wavelengths.
extern "C" int rhb200_set_stokes_mode(rhb200_ctx *c, int full_stokes)
{
  RH_NEED_CTX(c);
  c->no_stokes = full_stokes ? 0 : 1;
  free_wave(c);
  return RHB200_OK;
}

// get_atomic_rfs (inputs.h:92): the lines whose log gf the response function is taken for -- rows of the table passed
// to rhb200_set_lines (which is sorted by wavelength; RLK_Line.loggf_rf_ind = p for row line_rows[p], kurucz.c:254-257)
extern "C" int rhb200_set_loggf_rf(rhb200_ctx *c, int npar, const int *line_rows)
{
  RH_NEED_CTX(c);
  if (npar < 0 || npar > 16 || (npar > 0 && !line_rows)) { rhb200_set_error("rhb200_set_loggf_rf: 0 <= npar <= 16"); return RHB200_EINVAL; }
  for (int p = 0; p < npar; p++)
    if (line_rows[p] < -1 || line_rows[p] >= c->tab.nline) { rhb200_set_error("rhb200_set_loggf_rf: row %d outside the line table", line_rows[p]); return RHB200_EINVAL; }   // -1: a parameter no line carries (all zero)
  if (c->d_lrf_lines) { cudaFree(c->d_lrf_lines); c->d_lrf_lines = nullptr; }
  c->lrf_npar = npar;
  if (npar > 0) {
    RH_CUDA(cudaMalloc((void **) &c->d_lrf_lines, (size_t) npar * sizeof(int)));
    RH_CUDA(cudaMemcpy(c->d_lrf_lines, line_rows, (size_t) npar * sizeof(int), cudaMemcpyHostToDevice));
  }
  return RHB200_OK;
}

extern "C" int rhb200_set_passive_lines(rhb200_ctx *c, int nline, const double *plines, int ncomp,
                                        const double *c_shift, const double *c_fraction)
{
  RH_NEED_CTX(c);
  if (nline < 0 || ncomp < 0 || (nline > 0 && (!plines || !c_shift || !c_fraction))) { rhb200_set_error("rhb200_set_passive_lines: bad arguments"); return RHB200_EINVAL; }
  for (int n = 0; n < nline; n++) {
    const double *L = plines + (size_t) n * RHB200_PL_NFIELD;
    const int off = (int) L[RHB200_PL_COMPOFF], nc = (int) L[RHB200_PL_NCOMP];
    if (off < 0 || nc < 1 || off + nc > ncomp) { rhb200_set_error("passive line %d: component slice out of range", n); return RHB200_EINVAL; }
  }
  c->h_plines.assign(plines, plines + (size_t) nline * RHB200_PL_NFIELD);
  c->h_pcshift.assign(c_shift, c_shift + ncomp);
  c->h_pcfrac.assign(c_fraction, c_fraction + ncomp);
  free_wave(c);
  return RHB200_OK;
}

// Lines of the explicit model atoms (PASSIVE atoms of atoms.input): a Kurucz line of the same element and
// ionisation stage does not contribute at wavelengths inside such a line's wing window, where passive_bb already
// accounts for it (kurucz.c:617-633).  rows [n][4] = {row of the element in the table of rhb200_set_lines,
// stage of the model line's lower level, lambda0 [nm], qwing}.  Call between rhb200_set_lines and
// rhb200_set_wavelengths; rhb200_set_lines clears the table.
extern "C" int rhb200_set_model_lines(rhb200_ctx *c, int n, const double *rows)
{
  RH_NEED_CTX(c);
  if (n < 0 || (n > 0 && !rows)) { rhb200_set_error("rhb200_set_model_lines: bad arguments"); return RHB200_EINVAL; }
  c->h_model_lines.assign(rows, rows + (size_t) n * 4);
  free_wave(c);
  return RHB200_OK;
}

// Per-wavelength line window: the integer part of rlk_opacity (kurucz.c:538-566, 605-634),
// evaluated once on the host with the reference's own comparisons.
extern "C" int rhb200_set_wavelengths(rhb200_ctx *c, int nlambda, const double *lambda)
{
  RH_NEED_CTX(c);
  if (nlambda <= 0 || !lambda) { rhb200_set_error("rhb200_set_wavelengths: bad arguments"); return RHB200_EINVAL; }
  const int N = c->tab.nline;
  const double *LT = c->h_lines.data();
  auto lam0 = [&](int n) { return LT[(size_t) n * RHB200_RL_NFIELD + RHB200_RL_LAMBDA0]; };
  c->h_lambda.assign(lambda, lambda + nlambda);
  c->h_first.assign(nlambda, 0); c->h_count.assign(nlambda, 0); c->h_flags.assign(nlambda, 0);
  c->h_idx.clear();
  for (int l = 0; l < nlambda; l++) {
    const double lam = lambda[l];
    c->h_first[l] = (int) c->h_idx.size();
    if (N == 0) continue;
    const double dlamb_char = lam * RH_Q_WING * (c->tab.vmicro_char / RH_CLIGHT);
    if (lam < lam0(0) - dlamb_char || lam > lam0(N-1) + dlamb_char) continue;
    int lo = 0, hi = N;                              // rlk_locate with *low = 0: plain bisection
    while (hi - lo > 1) { const int mid = (hi + lo) >> 1; if (lam >= lam0(mid)) lo = mid; else hi = mid; }
    int Nblue = lo, Nred = lo;
    while (lam0(Nblue) + dlamb_char > lam && Nblue > 0) Nblue--;
    while (lam0(Nred) - dlamb_char < lam && Nred < N-1) Nred++;
    for (int n = Nblue; n <= Nred; n++) {
      if (std::fabs(lam0(n) - lam) <= dlamb_char) {
        const double *L = LT + (size_t) n * RHB200_RL_NFIELD;
        const double *E = c->h_elems.data() + (size_t) ((int) L[RHB200_RL_ELEM]) * RHB200_RE_NFIELD;
        bool contributes = (int) L[RHB200_RL_STAGE] < (int) E[RHB200_RE_NSTAGE] - 1;   // kurucz.c:614
        for (size_t kr = 0; contributes && kr < c->h_model_lines.size() / 4; kr++) {           // kurucz.c:617-633
          const double *M = c->h_model_lines.data() + 4 * kr;
          if ((int) M[0] != (int) L[RHB200_RL_ELEM]) continue;
          const double dlamb_wing = M[2] * M[3] * (c->tab.vmicro_char / RH_CLIGHT);
          if (std::fabs(lam - M[2]) <= dlamb_wing && (int) M[1] == (int) L[RHB200_RL_STAGE]) contributes = false;
        }
        if (contributes) {
          c->h_idx.push_back(n);
          c->h_flags[l] |= 1;
          if (L[RHB200_RL_POLARIZABLE] != 0.0) c->h_flags[l] |= 2;
        }
      }
    }
    c->h_count[l] = (int) c->h_idx.size() - c->h_first[l];
  }
  // passive_bb windows (metal.c:245-249) over the lines of rhb200_set_passive_lines, in table order; the lines that
  // are hit anywhere form the compact ACTIVE table the kernels index
  const int NPL = (int) (c->h_plines.size() / RHB200_PL_NFIELD);
  std::vector<int> prank(NPL, -1), pact, pw_first(nlambda, 0), pw_count(nlambda, 0), pw_idx;
  std::vector<std::vector<int>> hits(nlambda);
  for (int l = 0; l < nlambda; l++)
    for (int n = 0; n < NPL; n++) {
      const double *L = c->h_plines.data() + (size_t) n * RHB200_PL_NFIELD;
      const double dlambda = L[RHB200_PL_LAMBDA0] * L[RHB200_PL_QWING] * (c->tab.vmicro_char / RH_CLIGHT);
      if (std::fabs(lambda[l] - L[RHB200_PL_LAMBDA0]) <= dlambda) {
        if (prank[n] < 0) { prank[n] = 0; }
        hits[l].push_back(n);
        c->h_flags[l] |= 1;                                  // backgrflags.hasline
      }
    }
  for (int n = 0; n < NPL; n++) if (prank[n] == 0) { prank[n] = (int) pact.size(); pact.push_back(n); }
  for (int l = 0; l < nlambda; l++) {
    pw_first[l] = (int) pw_idx.size(); pw_count[l] = (int) hits[l].size();
    for (int n : hits[l]) pw_idx.push_back(prank[n]);
  }
  // MolecularOpacity windows (opacity.c:774-787), molecule by molecule, lines in table order
  const int NML = (int) (c->h_mlines.size() / RHB200_ML_NFIELD), NMS = (int) (c->h_msel.size() / 16);
  std::vector<int> mw_first(nlambda, 0), mw_count(nlambda, 0), mw_idx, mfirst(NMS, -1), mlast(NMS, -1);
  int mol_pol = 0;
  for (int n = 0; n < NML; n++) {
    const int m = (int) c->h_mlines[(size_t) n * RHB200_ML_NFIELD + RHB200_ML_MOL];
    if (mfirst[m] < 0) mfirst[m] = n;
    mlast[m] = n;
  }
  for (int l = 0; l < nlambda; l++) {
    const double lam = lambda[l], vc = c->tab.vmicro_char / RH_CLIGHT;
    mw_first[l] = (int) mw_idx.size();
    for (int m = 0; m < NMS; m++) {
      if (mfirst[m] < 0) continue;
      const double *L0 = c->h_mlines.data() + (size_t) mfirst[m] * RHB200_ML_NFIELD, *LN = c->h_mlines.data() + (size_t) mlast[m] * RHB200_ML_NFIELD;
      const double dl0 = lam * L0[RHB200_ML_QWING] * vc, dlN = lam * LN[RHB200_ML_QWING] * vc;
      if (!(lam >= L0[RHB200_ML_LAMBDA0] - dl0 && lam <= LN[RHB200_ML_LAMBDA0] + dlN)) continue;
      for (int n = mfirst[m]; n <= mlast[m]; n++) {
        const double *L = c->h_mlines.data() + (size_t) n * RHB200_ML_NFIELD;
        const double dl = lam * L[RHB200_ML_QWING] * vc;
        if (std::fabs(L[RHB200_ML_LAMBDA0] - lam) <= dl) {
          mw_idx.push_back(n); c->h_flags[l] |= 1;
          if (L[RHB200_ML_POLARIZABLE] != 0.0) { c->h_flags[l] |= 2; mol_pol = 1; }     // backgrflags.ispolarized, opacity.c:794-797
        }
      }
    }
    mw_count[l] = (int) mw_idx.size() - mw_first[l];
  }
  free_wave(c);
  DevWave &w = c->wav;
  w.nml = NML; w.nmsel = NMS; w.nmw = (int) mw_idx.size();
  if (w.nmw) {
    RH_CHECK(upload(&w.ml_rows, c->h_mlines.data(), c->h_mlines.size()));
    RH_CHECK(upload(&w.ml_sel, c->h_msel.data(), c->h_msel.size()));
    RH_CHECK(upload(&w.mw_first, mw_first.data(), (size_t) nlambda));
    RH_CHECK(upload(&w.mw_count, mw_count.data(), (size_t) nlambda));
    RH_CHECK(upload(&w.mw_idx, mw_idx.data(), mw_idx.size()));
    w.mol_pol = mol_pol;
    if (mol_pol) {
      RH_CHECK(upload(&w.mz_q, c->h_mzq.data(), c->h_mzq.size()));
      RH_CHECK(upload(&w.mz_shift, c->h_mzshift.data(), c->h_mzshift.size()));
      RH_CHECK(upload(&w.mz_strength, c->h_mzstrength.data(), c->h_mzstrength.size()));
    }
  }
  w.npl = (int) pact.size(); w.npw = (int) pw_idx.size();
  if (w.npl) {
    std::vector<double> rows((size_t) w.npl * RHB200_PL_NFIELD), pb((size_t) w.npl * RHB200_PB_NFIELD, 0.0);
    for (int a = 0; a < w.npl; a++) {
      const double *L = c->h_plines.data() + (size_t) pact[a] * RHB200_PL_NFIELD;
      std::copy(L, L + RHB200_PL_NFIELD, rows.begin() + (size_t) a * RHB200_PL_NFIELD);
      double *B = pb.data() + (size_t) a * RHB200_PB_NFIELD;
      B[RHB200_PB_LAMBDA0] = L[RHB200_PL_LAMBDA0]; B[RHB200_PB_QWING] = L[RHB200_PL_QWING]; B[RHB200_PB_BIJ] = L[RHB200_PL_BIJ];
      B[RHB200_PB_BJI] = L[RHB200_PL_BJI]; B[RHB200_PB_AJI] = L[RHB200_PL_AJI]; B[RHB200_PB_VOIGT] = L[RHB200_PL_VOIGT];
      B[RHB200_PB_NCOMP] = L[RHB200_PL_NCOMP]; B[RHB200_PB_COMPOFF] = L[RHB200_PL_COMPOFF];
    }
    RH_CHECK(upload(&w.pl_rows, rows.data(), rows.size()));
    RH_CHECK(upload(&w.pl_pb, pb.data(), pb.size()));
    RH_CHECK(upload(&w.pl_cshift, c->h_pcshift.data(), c->h_pcshift.size()));
    RH_CHECK(upload(&w.pl_cfrac, c->h_pcfrac.data(), c->h_pcfrac.size()));
    RH_CHECK(upload(&w.pw_first, pw_first.data(), (size_t) nlambda));
    RH_CHECK(upload(&w.pw_count, pw_count.data(), (size_t) nlambda));
    RH_CHECK(upload(&w.pw_idx, pw_idx.data(), pw_idx.size()));
  }
  w.nlambda = nlambda; w.nidx = (int) c->h_idx.size();
  RH_CHECK(upload(&w.lambda, lambda, (size_t) nlambda));
  RH_CHECK(upload(&w.first, c->h_first.data(), (size_t) nlambda));
  RH_CHECK(upload(&w.count, c->h_count.data(), (size_t) nlambda));
  RH_CHECK(upload(&w.flags, c->h_flags.data(), (size_t) nlambda));
  if (w.nidx) RH_CHECK(upload(&w.idx, c->h_idx.data(), (size_t) w.nidx));
  else RH_CUDA(cudaMalloc((void **) &w.idx, sizeof(int)));
  c->h_noline.clear();
  std::vector<int> rank(nlambda, -1);
  w.nunpol = 0;
  for (int l = 0; l < nlambda; l++) {
    // FULL_STOKES: wavelengths without a polarised line are solved for I alone -- with the scalar ray where a line is
    // present and the column moves; NO_STOKES (solveStokes false, formal.c:93-95): every wavelength is, and a polarised
    // background line still makes the wavelength angle dependent, i.e. takes the scalar ray (formal.c:100-103)
    if (c->no_stokes || (c->h_flags[l] & 2) == 0) c->h_noline.push_back(l);
    if (c->no_stokes ? (c->h_flags[l] & 1) : (c->h_flags[l] == 1)) rank[l] = w.nunpol++;
  }
  w.nnoline = (int) c->h_noline.size();
  if (w.nnoline) RH_CHECK(upload(&w.noline, c->h_noline.data(), (size_t) w.nnoline));
  RH_CHECK(upload(&w.unpol_rank, rank.data(), (size_t) nlambda));
  return RHB200_OK;
}

extern "C" int rhb200_get_line_windows(rhb200_ctx *c, int *first, int *count, int *idx, int cap, int *nidx)
{
  if (!c || c->wav.nlambda == 0) { rhb200_set_error("wavelengths not set"); return RHB200_ESTATE; }
  if (first) memcpy(first, c->h_first.data(), c->h_first.size()*sizeof(int));
  if (count) memcpy(count, c->h_count.data(), c->h_count.size()*sizeof(int));
  if (nidx) *nidx = (int) c->h_idx.size();
  if (idx) memcpy(idx, c->h_idx.data(), std::min((size_t) cap, c->h_idx.size())*sizeof(int));
  return RHB200_OK;
}

// atmos.backgrflags of the grid (background.c:335-338): bit 0 hasline (Kurucz, passive_bb or molecular line in the
// window), bit 1 ispolarized -- the integer part of Background(), fixed once the tables and the grid are set
extern "C" int rhb200_get_wavelength_flags(rhb200_ctx *c, int *flags)
{
  if (!c || c->wav.nlambda == 0 || !flags) { rhb200_set_error("wavelengths not set"); return RHB200_ESTATE; }
  memcpy(flags, c->h_flags.data(), c->h_flags.size()*sizeof(int));
  return RHB200_OK;
}

static int need_state(rhb200_ctx *c, bool wave)
{
  if (c->tab.nelem == 0 && c->tab.nline == 0 && c->tab.Tpf == nullptr) {
    rhb200_set_error("rhb200_set_lines() has not been called"); return RHB200_ESTATE;
  }
  if (wave && c->wav.nlambda == 0) { rhb200_set_error("rhb200_set_wavelengths() has not been called"); return RHB200_ESTATE; }
  return RHB200_OK;
}

// ------------------------------------------------------- batched LTE Stokes
static size_t align_up(size_t x) { return (x + 255) & ~(size_t) 255; }

struct ChunkLayout {
  size_t elem_n, lineprep, raypts, scal, colmov, total;
  ChunkLayout(const rhb200_ctx *c, int cc, int ndep) {
    elem_n   = align_up((size_t) cc * std::max(1, c->tab.nelem) * RHB200_RE_MAXSTAGE * ndep * sizeof(double));
    lineprep = align_up((size_t) cc * std::max(1, c->tab.nline) * ndep * LP_NFIELD * sizeof(double));
    raypts   = align_up((size_t) cc * c->wav.nlambda * ndep * RP_NFIELD * sizeof(double));
    scal     = align_up((size_t) cc * std::max(1, c->wav.nunpol) * c->scal_fields() * ndep * sizeof(double));
    colmov   = align_up((size_t) cc * (2 * sizeof(int) + sizeof(unsigned long long)));     // moving flags, scatter done flags, dJmax
    total = elem_n + lineprep + raypts + scal + colmov;
  }
};

static int chunk_columns(const rhb200_ctx *c, int ncol, int ndep, int nslots)
{
  size_t budget = (size_t) 8 << 30;
  if (const char *e = getenv("RHB200_WS_GB")) { double g = atof(e); if (g > 0.01) budget = (size_t) (g * (double) ((size_t) 1 << 30)); }
  ChunkLayout one(c, 1, ndep);
  size_t per = one.total + (size_t) (RHB200_AT_NFIELD * ndep + 2 * c->wav.nlambda * ndep + 4 * c->wav.nlambda) * sizeof(double);
  if (c->cont) per += (size_t) (rh_continuum_natom(c) + 4 + rh_continuum_nlev(c) + 8) * ndep * sizeof(double);
  size_t cc = budget / ((size_t) nslots * per);
  if (cc < 1) cc = 1;
  if (const char *e = getenv("RHB200_CHUNK_COLS")) { int v = atoi(e); if (v > 0) return std::min(ncol, v); }
  if (cc >= (size_t) ncol) return ncol;
  const size_t nchunk = ((size_t) ncol + cc - 1) / cc;      // equal chunks instead of full ones plus a small tail
  return (int) (((size_t) ncol + nchunk - 1) / nchunk);
}

// per-column atmos.moving flags of a chunk (written by the pyrh-rows step when VMACRO_TRESH > 0)
static int *chunk_col_moving(const rhb200_ctx *c, int cc, int ndep, char *ws)
{
  ChunkLayout L(c, cc, ndep);
  return (int *) (ws + L.elem_n + L.lineprep + L.raypts + L.scal + (size_t) cc * sizeof(unsigned long long));   // after dJmax[cc]
}

// pyrh boundary (rhb200_compute1d_batch): the columns arrive as pyrh.compute1d's nine rows
struct PyrhIn {
  const double *atmosphere; int nrow, atm_scale, iref; double wght_per_H, vmacro_tresh; double *scales;
  double total_abund = 0.0, gravity = 1.0; int scales_only = 0;      // rhb200_get_scales_batch
  // finite-difference response functions (rhb200_rf_fd_batch): the columns of the call are VIRTUAL -- column
  // v = ((base*npar + p)*ndep + k)*2 + s is base column `base` with row rf_rows[p] changed by +delta (s = 0) or
  // -delta (s = 1) at depth k; they are expanded on the device from d_base and only the differences travel back
  int rf_npar = 0; const int *d_rf_rows = nullptr; const double *d_rf_delta = nullptr; const double *d_base = nullptr;
  int rf_nsel = 0; const int *d_rf_sel = nullptr;         // perturbed depths: sel[0 .. nsel) (NULL: all ndep)
  double *rf_out = nullptr;
  double *lrf_out = nullptr;                     // analytic log gf response functions [ncol][nlambda][lrf_npar] (rhb200_compute1d_rf_batch)
};
struct ScalesStep { int iref, atm_scale; double wght_per_H; double *d_scratch, *d_scales_out;
                    double total_abund = 0.0, gravity = 1.0; int scales_only = 0; };

static int run_chunk_dev(rhb200_ctx *c, int cc, int ndep, double muz, int moving, int bc_top, int bc_bottom,
                         const double *d_atmos, const double *d_chi_ai, const double *d_eta_ai,
                         double *d_stokes, char *ws, const ScalesStep *sc = nullptr, bool per_column_moving = false,
                         const double *d_molchi = nullptr, const double *d_moleta = nullptr, const double *d_sca = nullptr,
                         double *d_lrf = nullptr)
{
  ChunkLayout L(c, cc, ndep);
  double *d_elem_n = (double *) ws, *d_lineprep = (double *) (ws + L.elem_n),
         *d_raypts = (double *) (ws + L.elem_n + L.lineprep), *d_scal = (double *) (ws + L.elem_n + L.lineprep + L.raypts);
  const int *d_colmov = per_column_moving ? chunk_col_moving(c, cc, ndep, ws) : nullptr;
  // NO_STOKES: only I is solved; Q, U, V are the zeros initSolution() callocs (spectrum.Stokes_Q/U/V, initial_xdr.c:96-100)
  if (c->no_stokes && d_stokes) RH_CUDA(cudaMemsetAsync(d_stokes, 0, (size_t) cc * 4 * c->wav.nlambda * sizeof(double), c->stream));
  RH_CHECK(rh_launch_prep(c, cc, ndep, muz, moving, d_atmos, d_elem_n, d_lineprep));
  RH_CHECK(rh_launch_opacity_fused(c, cc, ndep, 1, d_atmos, d_lineprep, d_chi_ai, d_eta_ai, d_raypts, d_molchi, d_moleta, d_sca));
  // convertScales() sits between Background() and Iterate() (pyrh_compute1dray.c:310-311): the height row is
  // only read by the formal solvers below
  if (sc) RH_CHECK(rh_launch_scales(c, cc, ndep, sc->iref, sc->atm_scale, sc->wght_per_H, sc->total_abund, sc->gravity,
                                    d_raypts, (double *) d_atmos, sc->d_scratch, sc->d_scales_out));
  if (sc && sc->scales_only) return RHB200_OK;
  if (!c->no_stokes) RH_CHECK(rh_launch_delo_raypts(c, cc, ndep, muz, bc_top, bc_bottom, d_atmos, d_raypts, d_stokes));
  RH_CHECK(rh_launch_feautrier_raypts(c, cc, ndep, muz, bc_top, bc_bottom, d_atmos, d_raypts, d_stokes,
                                      moving, d_colmov, d_scal));
  if (c->n_max_scatter > 0) {
    if (!d_sca) { rhb200_set_error("N_MAX_SCATTER > 0 needs the continuum on the device (the *_pops, *_atmos or compute1d entry points)"); return RHB200_ESTATE; }
    char *flagbase = ws + L.elem_n + L.lineprep + L.raypts + L.scal;
    unsigned long long *d_colmax = (unsigned long long *) flagbase;
    int *d_done = (int *) (flagbase + (size_t) cc * sizeof(unsigned long long)) + cc;       // after the moving flags
    RH_CHECK(rh_launch_scatter_passes(c, cc, ndep, muz, bc_top, bc_bottom, d_atmos, d_raypts, d_stokes, moving, d_colmov,
                                      d_colmax, d_done));
  }
  if (d_lrf) {                                   // get_atomic_rfs
    RH_CHECK(rh_launch_loggf_dopac(c, cc, ndep, d_atmos, d_lineprep, d_scal));
    RH_CHECK(rh_launch_loggf_rf(c, cc, ndep, muz, bc_top, bc_bottom, d_atmos, moving, d_colmov, d_scal, d_lrf));
  }
  return RHB200_OK;
}

static int check_batch_args(rhb200_ctx *c, int ncol, int ndep, double muz, int bc_top, int bc_bottom)
{
  RH_CHECK(need_state(c, true));
  if (ncol < 0 || ndep < 3 || !(muz > 0.0 && muz <= 1.0)) { rhb200_set_error("bad ncol/ndep/muz (%d, %d, %g)", ncol, ndep, muz); return RHB200_EINVAL; }
  if (bc_top != RHB200_BC_ZERO && bc_top != RHB200_BC_THERMALIZED) { rhb200_set_error("top boundary: only ZERO / THERMALIZED are implemented"); return RHB200_EUNSUPPORTED; }
  if (bc_bottom != RHB200_BC_THERMALIZED && bc_bottom != RHB200_BC_ZERO) { rhb200_set_error("bottom boundary: only THERMALIZED / ZERO are implemented"); return RHB200_EUNSUPPORTED; }
  return RHB200_OK;
}

extern "C" int rhb200_lte_stokes_batch_dev(rhb200_ctx *c, int ncol, int ndep, double muz, int moving,
                                           int bc_top, int bc_bottom, const double *d_atmos,
                                           const double *d_chi_ai, const double *d_eta_ai, double *d_stokes)
{
  RH_NEED_CTX(c);
  RH_CHECK(check_batch_args(c, ncol, ndep, muz, bc_top, bc_bottom));
  if (ncol == 0) return RHB200_OK;
  const int nl = c->wav.nlambda;
  const int cc = chunk_columns(c, ncol, ndep, 1);
  ChunkLayout L(c, cc, ndep);
  RH_CHECK(rh_ws_reserve(c, L.total));
  for (int c0 = 0; c0 < ncol; c0 += cc) {
    const int n = std::min(cc, ncol - c0);
    RH_CHECK(run_chunk_dev(c, n, ndep, muz, moving, bc_top, bc_bottom,
                           d_atmos + (size_t) c0 * RHB200_AT_NFIELD * ndep,
                           d_chi_ai + (size_t) c0 * nl * ndep, d_eta_ai + (size_t) c0 * nl * ndep,
                           d_stokes + (size_t) c0 * 4 * nl, (char *) c->ws));
  }
  RH_CUDA(cudaStreamSynchronize(c->stream));
  return RHB200_OK;
}

// HOST pointers.  Two slots (device input/output buffers + workspace), each driven by its own
// stream: the H2D copy of chunk i+1 overlaps the kernels of chunk i when the host buffers are
// pinned (rhb200_host_alloc_pinned); pageable memory still works, just without overlap.
// chem != NULL: the background continuum is evaluated on the device from LTE populations
// (rhb200_set_continuum) instead of being copied in as chi_ai / eta_ai
static int lte_batch_host(rhb200_ctx *c, int ncol, int ndep, double muz, int moving,
                          int bc_top, int bc_bottom, const double *atmos,
                          const double *chi_ai, const double *eta_ai, const double *chem, double *stokes,
                          int chem_on_device = 0, const PyrhIn *py = nullptr)
{
  RH_NEED_CTX(c);
  RhRange whole("rhf1d (LTE, batch)");
  RH_CHECK(check_batch_args(c, ncol, ndep, muz, bc_top, bc_bottom));
  if (ncol == 0) return RHB200_OK;
  if ((!atmos && !py) || (!stokes && !(py && (py->rf_out || py->scales_only || py->lrf_out))) || (!chem && !chem_on_device && (!chi_ai || !eta_ai))) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  const bool cont_dev = chem || chem_on_device;
  if (py && py->lrf_out) {
    if (c->lrf_npar <= 0) { rhb200_set_error("rhb200_set_loggf_rf() has not been called"); return RHB200_ESTATE; }
    if (c->lrf_npar > ndep) { rhb200_set_error("log gf response functions: npar (%d) must not exceed ndep (%d)", c->lrf_npar, ndep); return RHB200_EINVAL; }
  }
  if (cont_dev && !c->cont) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  const int nl = c->wav.nlambda;
  const int nslots = 2;
  int cc = chunk_columns(c, ncol, ndep, nslots);
  if (py && py->rf_npar && (cc & 1)) cc = std::max(2, cc - 1);      // +delta / -delta pairs stay in one chunk
  ChunkLayout L(c, cc, ndep);
  const size_t b_at = align_up((size_t) cc * RHB200_AT_NFIELD * ndep * sizeof(double));
  const size_t b_op = align_up((size_t) cc * nl * ndep * sizeof(double));
  const size_t b_st = align_up((size_t) cc * 4 * nl * sizeof(double));
  const int nchem = cont_dev ? rh_continuum_natom(c) + 4 : 0;
  const size_t b_ch = cont_dev ? align_up((size_t) cc * nchem * ndep * sizeof(double)) : 0;
  const size_t b_pp = cont_dev ? align_up((size_t) cc * rh_continuum_nlev(c) * ndep * sizeof(double)) : 0;
  const size_t b_tp = cont_dev ? align_up((size_t) cc * 8 * ndep * sizeof(double)) : 0;
  const size_t b_in = py ? align_up((size_t) cc * py->nrow * ndep * sizeof(double)) : 0;     // pyrh rows as they arrive
  const size_t b_sc = py ? align_up((size_t) cc * 5 * ndep * sizeof(double)) : 0;            // {tau, cmass} scratch + {height, tau_ref, cmass} out
  if (c->wav.npl > 0 && !cont_dev) { rhb200_set_error("passive_bb lines need the populations of the device continuum: use the *_pops, *_atmos or compute1d entry points"); return RHB200_ESTATE; }
  const size_t b_pc = align_up((size_t) cc * std::max(1, c->wav.npl) * 4 * ndep * sizeof(double));   // passive_bb: n_i, n_j, vbroad, adamp
  if (c->wav.nmw > 0 && !cont_dev) { rhb200_set_error("molecular lines need the chemistry of the device continuum: use the *_atmos or compute1d entry points"); return RHB200_ESTATE; }
  const bool mol_on = c->wav.nmw > 0;
  if (mol_on) {                                  // tell the chemistry kernel which densities to keep
    std::vector<int> chem(c->wav.nmsel);
    for (int m = 0; m < c->wav.nmsel; m++) chem[m] = (int) c->h_msel[(size_t) m * 16];
    RH_CHECK(rh_continuum_set_molsel(c, c->wav.nmsel, chem.data()));
  }
  const size_t b_md = mol_on ? align_up((size_t) cc * c->wav.nmsel * 4 * ndep * sizeof(double)) : 0;   // densities + {n, pf, vbroad}
  const size_t b_mo = mol_on ? (c->wav.mol_pol ? 4 : 1) * b_op : 0;                                     // chi, eta of the molecular lines (I, or I Q U V planes)
  const size_t b_sa = (c->n_max_scatter > 0 && cont_dev) ? b_op : 0;                                  // scattering opacity
  const size_t slot = L.total + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc + b_pc + b_md + 2*b_mo + b_sa;
  RH_CHECK(rh_ws_reserve(c, nslots * slot));
  cudaStream_t streams[2] = {c->stream, c->copy_stream};
  cudaStream_t saved = c->stream;
  int rc = RHB200_OK, i = 0;
  for (int c0 = 0; c0 < ncol && rc == RHB200_OK; c0 += cc, i++) {
    const int n = std::min(cc, ncol - c0);
    char *base = (char *) c->ws + (size_t) (i % nslots) * slot;
    double *d_at = (double *) base, *d_chi = (double *) (base + b_at), *d_eta = (double *) (base + b_at + b_op),
           *d_st = (double *) (base + b_at + 2*b_op);
    double *d_ch = (double *) (base + b_at + 2*b_op + b_st), *d_pp = (double *) (base + b_at + 2*b_op + b_st + b_ch),
           *d_tp = (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp);
    double *d_in = (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp);
    double *d_sc = (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in);
    double *d_pc = (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc);
    double *d_md = (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc + b_pc);
    double *d_mchi = mol_on ? (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc + b_pc + b_md) : nullptr;
    double *d_meta = mol_on ? (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc + b_pc + b_md + b_mo) : nullptr;
    double *d_sa = b_sa ? (double *) (base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc + b_pc + b_md + 2*b_mo) : nullptr;
    char *ws = base + b_at + 2*b_op + b_st + b_ch + b_pp + b_tp + b_in + b_sc + b_pc + b_md + 2*b_mo + b_sa;
    cudaStream_t st = streams[i % nslots];
    c->stream = st;
    cudaError_t e;
    if (py) {
      if (py->rf_npar) {
        rc = rh_launch_rf_expand(c, c0, n, ndep, py->nrow, py->rf_npar, py->d_rf_rows, py->d_rf_delta, py->d_base, d_in,
                                 py->rf_nsel, py->d_rf_sel);
        if (rc != RHB200_OK) break;
      } else
      if ((e = cudaMemcpyAsync(d_in, py->atmosphere + (size_t) c0 * py->nrow * ndep,
                               (size_t) n * py->nrow * ndep * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) {
        rhb200_set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
      }
      rc = rh_launch_pyrh_rows(c, n, ndep, py->nrow, py->atm_scale, muz, py->vmacro_tresh, d_in, d_at,
                               chunk_col_moving(c, n, ndep, ws));
      if (rc != RHB200_OK) break;
    } else
    if ((e = cudaMemcpyAsync(d_at, atmos + (size_t) c0 * RHB200_AT_NFIELD * ndep,
                             (size_t) n * RHB200_AT_NFIELD * ndep * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) {
      rhb200_set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
    if (cont_dev) {
      if (chem && (e = cudaMemcpyAsync(d_ch, chem + (size_t) c0 * nchem * ndep, (size_t) n * nchem * ndep * sizeof(double),
                               cudaMemcpyHostToDevice, st)) != cudaSuccess) {
        rhb200_set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
      }
      if (mol_on && chem) { rhb200_set_error("molecular lines need the chemistry on the device (not the *_pops entry point)"); rc = RHB200_ESTATE; break; }
      rc = rh_continuum_chunk(c, n, ndep, d_at, d_ch, d_pp, d_tp, d_chi, d_eta, chem ? 0 : 1, mol_on ? d_md : nullptr, d_sa);
      if (rc != RHB200_OK) break;
      if (mol_on) {                              // MolecularOpacity, background.c:548-566
        rc = rh_molecular_chunk(c, n, ndep, muz, d_at, d_md, d_md + (size_t) n * c->wav.nmsel * ndep, d_mchi, d_meta);
        if (rc != RHB200_OK) break;
      }
      rc = rh_passive_chunk(c, n, ndep, muz, d_at, d_pp, rh_continuum_nlev(c), d_pc, d_chi, d_eta);   // background.c:494-515
      if (rc != RHB200_OK) break;
      if (py) {                                  // np = atmos.H->n[Nlevel-1] (kurucz.c:772); H is the first model atom
        rc = rh_launch_proton(c, n, ndep, rh_continuum_nlev(c), rh_continuum_proton_level(c), d_pp, d_at);
        if (rc != RHB200_OK) break;
      }
    } else if ((e = cudaMemcpyAsync(d_chi, chi_ai + (size_t) c0 * nl * ndep, (size_t) n * nl * ndep * sizeof(double),
                             cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_eta, eta_ai + (size_t) c0 * nl * ndep, (size_t) n * nl * ndep * sizeof(double),
                             cudaMemcpyHostToDevice, st)) != cudaSuccess) {
      rhb200_set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
    ScalesStep sc{py ? py->iref : 0, py ? py->atm_scale : 0, py ? py->wght_per_H : 0.0, d_sc,
                  (py && py->scales) ? d_sc + (size_t) cc * 2 * ndep : nullptr};
    if (py) { sc.total_abund = py->total_abund; sc.gravity = py->gravity; sc.scales_only = py->scales_only; }
    rc = run_chunk_dev(c, n, ndep, muz, moving, bc_top, bc_bottom, d_at, d_chi, d_eta, d_st, ws, py ? &sc : nullptr,
                       py && py->vmacro_tresh > 0.0, d_mchi, d_meta, d_sa, (py && py->lrf_out) ? d_eta : nullptr);
    if (rc != RHB200_OK) break;
    // d_eta (eta_ai, [n][nl][ndep]) is free once the opacity kernel has run: the response functions live there
    if (py && py->lrf_out && (e = cudaMemcpyAsync(py->lrf_out + (size_t) c0 * nl * c->lrf_npar, d_eta, (size_t) n * nl * c->lrf_npar * sizeof(double),
                                                  cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
      rhb200_set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
    if (sc.d_scales_out && (e = cudaMemcpyAsync(py->scales + (size_t) c0 * 3 * ndep, sc.d_scales_out, (size_t) n * 3 * ndep * sizeof(double),
                                                cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
      rhb200_set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
    if (py && py->scales_only) continue;
    if (py && py->rf_npar) {                     // (S+ - S-) / (2 delta); d_chi is free once the opacity kernel has run
      rc = rh_launch_rf_diff(c, c0, n, py->rf_nsel, nl, py->rf_npar, py->d_rf_delta, d_st, d_chi);
      if (rc != RHB200_OK) break;
      if ((e = cudaMemcpyAsync(py->rf_out + (size_t) (c0 / 2) * 4 * nl, d_chi, (size_t) (n / 2) * 4 * nl * sizeof(double),
                               cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
        rhb200_set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
      }
    } else
    if (stokes && (e = cudaMemcpyAsync(stokes + (size_t) c0 * 4 * nl, d_st, (size_t) n * 4 * nl * sizeof(double),
                             cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
      rhb200_set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
  }
  c->stream = saved;
  cudaError_t e1 = cudaStreamSynchronize(streams[0]), e2 = cudaStreamSynchronize(streams[1]);
  if (rc == RHB200_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
    rhb200_set_error("kernel execution failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    rc = RHB200_ECUDA;
  }
  return rc;
}

extern "C" int rhb200_lte_stokes_batch(rhb200_ctx *c, int ncol, int ndep, double muz, int moving,
                                       int bc_top, int bc_bottom, const double *atmos,
                                       const double *chi_ai, const double *eta_ai, double *stokes)
{
  if (!chi_ai || !eta_ai) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  return lte_batch_host(c, ncol, ndep, muz, moving, bc_top, bc_bottom, atmos, chi_ai, eta_ai, nullptr, stokes);
}

extern "C" int rhb200_lte_stokes_batch_pops(rhb200_ctx *c, int ncol, int ndep, double muz, int moving,
                                            int bc_top, int bc_bottom, const double *atmos,
                                            const double *chem, double *stokes)
{
  if (!chem) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  return lte_batch_host(c, ncol, ndep, muz, moving, bc_top, bc_bottom, atmos, nullptr, nullptr, chem, stokes);
}

extern "C" int rhb200_lte_stokes_batch_atmos(rhb200_ctx *c, int ncol, int ndep, double muz, int moving,
                                             int bc_top, int bc_bottom, const double *atmos, double *stokes)
{
  return lte_batch_host(c, ncol, ndep, muz, moving, bc_top, bc_bottom, atmos, nullptr, nullptr, nullptr, stokes, 1);
}

// pyrh.compute1d() / rhf1d() in LTE for a batch of columns (pyrh.pyx:537-668, pyrh_compute1dray.c:112-357):
// the nine pyrh rows in, Stokes spectra on the context's wavelength grid (lambda_ref included, as in
// spectrum.lambda; _solveray drops it when packing, pyrh_solveray.c:130-150) out
extern "C" int rhb200_compute1d_batch(rhb200_ctx *c, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                      const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                      int bc_top, int bc_bottom, double *stokes, double *scales)
{
  RH_NEED_CTX(c);
  if (!atmosphere) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  if (nrow < 9) { rhb200_set_error("atmosphere needs the 9 rows of pyrh.compute1d (pyrh.pyx:621-625), got %d", nrow); return RHB200_EINVAL; }
  if (atm_scale < 0 || atm_scale > 2) { rhb200_set_error("atm_scale must be 0 (tau500), 1 (column mass) or 2 (height)"); return RHB200_EINVAL; }
  if (iref < 0 || iref >= c->wav.nlambda) { rhb200_set_error("iref outside the wavelength grid"); return RHB200_EINVAL; }
  if (!(wght_per_H > 0.0) && atm_scale == 1) { rhb200_set_error("wght_per_H (abundance.c:220) is needed for the column-mass scale"); return RHB200_EINVAL; }
  PyrhIn py{atmosphere, nrow, atm_scale, iref, wght_per_H, vmacro_tresh, scales};
  if (scales && atm_scale == 2) {
    if (!(c->gravity > 0.0)) { rhb200_set_error("scales on a height grid: the column-mass row needs rhb200_set_gravity() (multiatmos.c:153-155)"); return RHB200_EINVAL; }
    py.total_abund = c->total_abund; py.gravity = c->gravity;
  }
  return lte_batch_host(c, ncol, ndep, mu, 1, bc_top, bc_bottom, nullptr, nullptr, nullptr, nullptr, stokes, 1, &py);
}

extern "C" int rhb200_set_gravity(rhb200_ctx *c, double total_abund, double gravity)
{
  RH_NEED_CTX(c);
  if (!(total_abund > 0.0) || !(gravity > 0.0)) { rhb200_set_error("total_abund and gravity [m/s^2] must be positive"); return RHB200_EINVAL; }
  c->total_abund = total_abund; c->gravity = gravity;
  return RHB200_OK;
}

extern "C" int rhb200_shard_columns(int ncol, int rank, int nrank, int *first, int *count)
{
  if (ncol < 0 || nrank < 1 || rank < 0 || rank >= nrank || !first || !count) { rhb200_set_error("rhb200_shard_columns: bad arguments"); return RHB200_EINVAL; }
  const int base = ncol / nrank, rem = ncol % nrank;
  *first = rank * base + std::min(rank, rem);
  *count = base + (rank < rem ? 1 : 0);
  return RHB200_OK;
}

// One host batch fanned over the GPUs of the box from one process: a host thread per context, contiguous column blocks,
// no collective (columns are independent).  The error text of a failing block is copied into the caller's thread.
extern "C" int rhb200_compute1d_batch_multi(int nctx, rhb200_ctx *const *ctxs, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                            const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                            int bc_top, int bc_bottom, double *stokes, double *scales)
{
  if (nctx < 1 || !ctxs) { rhb200_set_error("rhb200_compute1d_batch_multi: no contexts"); return RHB200_EINVAL; }
  for (int i = 0; i < nctx; i++) {
    if (!ctxs[i]) { rhb200_set_error("context %d is NULL", i); return RHB200_EINVAL; }
    if (ctxs[i]->wav.nlambda != ctxs[0]->wav.nlambda) { rhb200_set_error("contexts differ in their wavelength grids"); return RHB200_EINVAL; }
    for (int j = 0; j < i; j++) if (ctxs[j]->device == ctxs[i]->device) { rhb200_set_error("contexts %d and %d share device %d", j, i, ctxs[i]->device); return RHB200_EINVAL; }
  }
  if (nctx == 1) return rhb200_compute1d_batch(ctxs[0], ncol, ndep, nrow, mu, atm_scale, atmosphere, iref, wght_per_H, vmacro_tresh,
                                               bc_top, bc_bottom, stokes, scales);
  const int nl = ctxs[0]->wav.nlambda;
  std::vector<int> rc(nctx, RHB200_OK);
  std::vector<std::string> err(nctx);
  std::vector<std::thread> th;
  for (int i = 0; i < nctx; i++)
    th.emplace_back([&, i]() {
      int first = 0, count = 0;
      rhb200_shard_columns(ncol, i, nctx, &first, &count);
      if (count == 0) return;
      rc[i] = rhb200_compute1d_batch(ctxs[i], count, ndep, nrow, mu, atm_scale, atmosphere + (size_t) first * nrow * ndep, iref,
                                     wght_per_H, vmacro_tresh, bc_top, bc_bottom,
                                     stokes ? stokes + (size_t) first * 4 * nl : nullptr,
                                     scales ? scales + (size_t) first * 3 * ndep : nullptr);
      if (rc[i] != RHB200_OK) err[i] = rhb200_last_error();
    });
  for (auto &t : th) t.join();
  for (int i = 0; i < nctx; i++)
    if (rc[i] != RHB200_OK) { rhb200_set_error("device %d: %s", ctxs[i]->device, err[i].c_str()); return rc[i]; }
  return RHB200_OK;
}

// get_atomic_rfs = 1 (pyrh.pyx:604-606, 658-660): rhb200_compute1d_batch plus mySpectrum.rfs for the log gf parameters
// registered with rhb200_set_loggf_rf
extern "C" int rhb200_compute1d_rf_batch(rhb200_ctx *c, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                         const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                         int bc_top, int bc_bottom, double *stokes, double *scales, double *rfs)
{
  RH_NEED_CTX(c);
  if (!atmosphere || !rfs) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  if (nrow < 9) { rhb200_set_error("atmosphere needs the 9 rows of pyrh.compute1d (pyrh.pyx:621-625), got %d", nrow); return RHB200_EINVAL; }
  if (atm_scale < 0 || atm_scale > 2) { rhb200_set_error("atm_scale must be 0 (tau500), 1 (column mass) or 2 (height)"); return RHB200_EINVAL; }
  if (iref < 0 || iref >= c->wav.nlambda) { rhb200_set_error("iref outside the wavelength grid"); return RHB200_EINVAL; }
  if (!(wght_per_H > 0.0) && atm_scale == 1) { rhb200_set_error("wght_per_H (abundance.c:220) is needed for the column-mass scale"); return RHB200_EINVAL; }
  PyrhIn py{atmosphere, nrow, atm_scale, iref, wght_per_H, vmacro_tresh, scales};
  if (scales && atm_scale == 2) {
    if (!(c->gravity > 0.0)) { rhb200_set_error("scales on a height grid: the column-mass row needs rhb200_set_gravity() (multiatmos.c:153-155)"); return RHB200_EINVAL; }
    py.total_abund = c->total_abund; py.gravity = c->gravity;
  }
  py.lrf_out = rfs;
  return lte_batch_host(c, ncol, ndep, mu, 1, bc_top, bc_bottom, nullptr, nullptr, nullptr, nullptr, stokes, 1, &py);
}

// pyrh.get_scales() (pyrh.pyx:491-534, rhf1d/pyrh_hse.c:402-553) for a batch: Background() at the reference
// wavelength + convertScales(); the context's line table is normally empty (the reference sets Nrlk = 0 there)
extern "C" int rhb200_get_scales_batch(rhb200_ctx *c, int ncol, int ndep, int nrow, int atm_scale,
                                       const double *atmosphere, int iref, double wght_per_H, double total_abund,
                                       double gravity, double vmacro_tresh, double *scales)
{
  RH_NEED_CTX(c);
  if (!atmosphere || !scales) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  if (nrow < 9 || atm_scale < 0 || atm_scale > 2 || iref < 0 || iref >= c->wav.nlambda) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  if (!(wght_per_H > 0.0) || !(total_abund > 0.0) || !(gravity > 0.0)) { rhb200_set_error("wght_per_H, total_abund (abundance.c:219-220) and gravity [m/s^2] must be positive"); return RHB200_EINVAL; }
  PyrhIn py{atmosphere, nrow, atm_scale, iref, wght_per_H, vmacro_tresh, scales};
  py.total_abund = total_abund; py.gravity = gravity; py.scales_only = 1;
  return lte_batch_host(c, ncol, ndep, 1.0, 1, RHB200_BC_ZERO, RHB200_BC_THERMALIZED, nullptr, nullptr, nullptr, nullptr,
                        nullptr, 1, &py);
}

// ------------------------------------------------ function-level entry points
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 8));
    if (e != cudaSuccess) { cudaGetLastError(); rhb200_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return RHB200_ENOMEM; }
    return RHB200_OK;
  }
  int from_host(const void *h, size_t bytes) {
    RH_CHECK(alloc(bytes));
    if (bytes) RH_CUDA(cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice));
    return RHB200_OK;
  }
  template <class T> T *as() { return (T *) p; }
};

// Single-depth perturbations without the 2 npar ndep full syntheses: everything before convertScales() is local in
// depth, so the depth-kp opacities of "parameter p changed at depth kp" are those of the pseudo column "parameter p
// changed at every depth".  Per base column 1 + 2 npar full columns go through Background() (instead of 2 npar ndep);
// each virtual column then gets its own convertScales() walk (vscales_kernel) and formal solution, reading the base
// column's ray-point records with the depth-kp record taken from the pseudo column.  Same arithmetic on the same
// numbers as the brute-force path: results are bit-identical (tests/test_gpu_parity.py).
// Two slots / streams: the D2H of chunk i overlaps the kernels of chunk i+1.
static int rf_fd_single_depth(rhb200_ctx *c, int ncol, int ndep, int nrow, double mu, int atm_scale, int iref, double wght_per_H,
                              int bc_top, int bc_bottom, int npar, const int *d_rows, const double *d_delta,
                              const double *d_base, double *rf, const int *h_rows, int nsel, const int *d_sel)
{
  RhRange whole("rhf1d (LTE, finite-difference response functions)");
  RH_CHECK(check_batch_args(c, ncol, ndep, mu, bc_top, bc_bottom));
  if (!c->cont) { rhb200_set_error("rhb200_set_continuum() has not been called"); return RHB200_ESTATE; }
  const int nl = c->wav.nlambda, nfull1 = 1 + 2*npar, nv1 = 2*npar*nsel, nslots = 2;
  const bool mol_on = c->wav.nmw > 0;
  if (mol_on) {
    std::vector<int> chem(c->wav.nmsel);
    for (int m = 0; m < c->wav.nmsel; m++) chem[m] = (int) c->h_msel[(size_t) m * 16];
    RH_CHECK(rh_continuum_set_molsel(c, c->wav.nmsel, chem.data()));
  }
  const int nchem = rh_continuum_natom(c) + 4, nlev = rh_continuum_nlev(c);
  // parameters that leave T, n_e, n_H and the opacity at the reference wavelength alone (v_z, v_mic, B, gamma, chi with a
  // line-free reference wavelength) leave the depth scales alone: below the perturbed depth their rays are the base
  // column's, bit for bit, and resume from its saved sweep (delo_base_state_kernel)
  std::vector<int> neutral(npar, 0);
  int pn = -1;
  {
    int fl = 1;
    RH_CUDA(cudaMemcpy(&fl, c->wav.flags + iref, sizeof(int), cudaMemcpyDeviceToHost));
    const bool resume = (fl & 3) == 0 && !c->no_stokes && !(getenv("RHB200_RF_FD_NO_RESUME") && atoi(getenv("RHB200_RF_FD_NO_RESUME")));
    for (int p = 0; p < npar && resume; p++)
      if (h_rows[p] >= 3 && h_rows[p] <= 7) { neutral[p] = 1; if (pn < 0) pn = p; }
  }
  struct Lay { size_t in, at, chi, eta, ch, pp, tp, pc, md, mchi, meta, ws, vws, vst, vsc, rfo, st, neu, total; };
  auto layout = [&](int nb) {
    Lay y; size_t o = 0;
    const size_t nf = (size_t) nb * nfull1, nv = (size_t) nb * nv1;
    auto take = [&](size_t bytes) { const size_t at = o; o += align_up(bytes); return at; };
    y.in  = take(nf * nrow * ndep * sizeof(double));
    y.at  = take(nf * RHB200_AT_NFIELD * ndep * sizeof(double));
    y.chi = take(nf * nl * ndep * sizeof(double));
    y.eta = take(nf * nl * ndep * sizeof(double));
    y.ch  = take(nf * nchem * ndep * sizeof(double));
    y.pp  = take(nf * nlev * ndep * sizeof(double));
    y.tp  = take(nf * 8 * ndep * sizeof(double));
    y.pc  = take(nf * std::max(1, c->wav.npl) * 4 * ndep * sizeof(double));
    y.md  = take(mol_on ? nf * c->wav.nmsel * 4 * ndep * sizeof(double) : 0);
    y.mchi = take(mol_on ? (c->wav.mol_pol ? 4 : 1) * nf * nl * ndep * sizeof(double) : 0);
    y.meta = take(mol_on ? (c->wav.mol_pol ? 4 : 1) * nf * nl * ndep * sizeof(double) : 0);
    y.ws  = take(ChunkLayout(c, (int) nf, ndep).total);
    y.vws = take(nv * 4 * ndep * sizeof(double));
    y.vst = take(nv * 4 * nl * sizeof(double));
    y.vsc = take(nv * std::max(1, c->wav.nnoline) * 5 * ndep * sizeof(double));
    y.rfo = take(nv / 2 * 4 * nl * sizeof(double));
    y.st  = take(pn >= 0 ? (size_t) nb * nl * ndep * 44 * sizeof(double) : 0);      // DELO_NSTATE doubles per (ray, depth)
    y.neu = take((size_t) npar * sizeof(int));
    y.total = o;
    return y;
  };
  size_t budget = (size_t) 8 << 30;
  if (const char *e = getenv("RHB200_WS_GB")) { double g = atof(e); if (g > 0.01) budget = (size_t) (g * (double) ((size_t) 1 << 30)); }
  const size_t per = layout(1).total;
  int nb = (int) std::max<size_t>(1, std::min<size_t>((size_t) ncol, budget / (nslots * per)));
  if (const char *e = getenv("RHB200_RF_CHUNK_COLS")) { int v = atoi(e); if (v > 0) nb = std::min(ncol, v); }
  if (nb < ncol) { const int nchunk = (ncol + nb - 1) / nb; nb = (ncol + nchunk - 1) / nchunk; }
  const Lay y = layout(nb);
  RH_CHECK(rh_ws_reserve(c, nslots * y.total));
  cudaStream_t streams[2] = {c->stream, c->copy_stream};
  cudaStream_t saved = c->stream;
  int rc = RHB200_OK, i = 0;
  for (int b0 = 0; b0 < ncol && rc == RHB200_OK; b0 += nb, i++) {
    const int n = std::min(nb, ncol - b0), nf = n * nfull1, nv = n * nv1;
    char *base = (char *) c->ws + (size_t) (i % nslots) * y.total;
    auto D = [&](size_t off) { return (double *) (base + off); };
    c->stream = streams[i % nslots];
    ChunkLayout L(c, nf, ndep);
    char *ws = base + y.ws;
    double *d_elem_n = (double *) ws, *d_lineprep = (double *) (ws + L.elem_n), *d_raypts = (double *) (ws + L.elem_n + L.lineprep);
#define RF_STEP(call) if ((rc = (call)) != RHB200_OK) break
    RF_STEP(rh_launch_rf_expand_full(c, b0, n, ndep, nrow, npar, d_rows, d_delta, d_base, D(y.in)));
    RF_STEP(rh_launch_pyrh_rows(c, nf, ndep, nrow, atm_scale, mu, 0.0, D(y.in), D(y.at), nullptr));
    RF_STEP(rh_continuum_chunk(c, nf, ndep, D(y.at), D(y.ch), D(y.pp), D(y.tp), D(y.chi), D(y.eta), 1, mol_on ? D(y.md) : nullptr, nullptr));
    if (mol_on) RF_STEP(rh_molecular_chunk(c, nf, ndep, mu, D(y.at), D(y.md), D(y.md) + (size_t) nf * c->wav.nmsel * ndep, D(y.mchi), D(y.meta)));
    RF_STEP(rh_passive_chunk(c, nf, ndep, mu, D(y.at), D(y.pp), nlev, D(y.pc), D(y.chi), D(y.eta)));
    RF_STEP(rh_launch_proton(c, nf, ndep, nlev, rh_continuum_proton_level(c), D(y.pp), D(y.at)));
    RF_STEP(rh_launch_prep(c, nf, ndep, mu, 1, D(y.at), d_elem_n, d_lineprep));
    RF_STEP(rh_launch_opacity_fused(c, nf, ndep, 1, D(y.at), d_lineprep, D(y.chi), D(y.eta), d_raypts,
                                    mol_on ? D(y.mchi) : nullptr, mol_on ? D(y.meta) : nullptr, nullptr));
    RF_STEP(rh_launch_vscales(c, n, npar, ndep, iref, atm_scale, wght_per_H, 0.0, 1.0, d_raypts, D(y.at), D(y.vws), nsel, d_sel));
    cudaError_t e;
    if (c->no_stokes && (e = cudaMemsetAsync(D(y.vst), 0, (size_t) nv * 4 * nl * sizeof(double), c->stream)) != cudaSuccess) {
      rhb200_set_error("memset failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
    if ((e = cudaMemcpyAsync(base + y.neu, neutral.data(), (size_t) npar * sizeof(int), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) {
      rhb200_set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
    if (!c->no_stokes) RF_STEP(rh_launch_delo_vcols(c, n, npar, ndep, mu, bc_top, bc_bottom, D(y.vws), d_raypts, D(y.vst),
                                                    (const int *) (base + y.neu), pn, pn >= 0 ? D(y.st) : nullptr, nsel, d_sel));
    RF_STEP(rh_launch_noline_vcols(c, n, npar, ndep, mu, bc_top, bc_bottom, D(y.vws), d_raypts, D(y.vst), D(y.vsc), nsel, d_sel));
    RF_STEP(rh_launch_rf_diff(c, b0 * nv1, nv, nsel, nl, npar, d_delta, D(y.vst), D(y.rfo)));
#undef RF_STEP
    if ((e = cudaMemcpyAsync(rf + (size_t) b0 * (nv1 / 2) * 4 * nl, D(y.rfo), (size_t) (nv / 2) * 4 * nl * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) {
      rhb200_set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = RHB200_ECUDA; break;
    }
  }
  c->stream = saved;
  cudaError_t e1 = cudaStreamSynchronize(streams[0]), e2 = cudaStreamSynchronize(streams[1]);
  if (rc == RHB200_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
    rhb200_set_error("kernel execution failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    rc = RHB200_ECUDA;
  }
  return rc;
}

// Finite-difference response functions of the LTE Stokes spectrum to the atmosphere rows, per depth point:
// what a pyrh caller (an inversion code) obtains from 2 x npar x ndep calls of pyrh.compute1d per column.
extern "C" int rhb200_rf_fd_depths_batch(rhb200_ctx *c, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                         const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                         int bc_top, int bc_bottom, int npar, const int *par_rows, const double *par_delta,
                                         int nsel, const int *depths, double *rf);
extern "C" int rhb200_rf_fd_batch(rhb200_ctx *c, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                  const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                  int bc_top, int bc_bottom, int npar, const int *par_rows, const double *par_delta,
                                  double *rf)
{
  return rhb200_rf_fd_depths_batch(c, ncol, ndep, nrow, mu, atm_scale, atmosphere, iref, wght_per_H, vmacro_tresh, bc_top, bc_bottom,
                                   npar, par_rows, par_delta, ndep, nullptr, rf);
}

// the same at SELECTED depths only (the nodes of an inversion): depths [nsel] ascending indices (NULL: all),
// rf [ncol][npar][nsel][4][nlambda]; work and D2H traffic scale with nsel / ndep
extern "C" int rhb200_rf_fd_depths_batch(rhb200_ctx *c, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                         const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                         int bc_top, int bc_bottom, int npar, const int *par_rows, const double *par_delta,
                                         int nsel, const int *depths, double *rf)
{
  RH_NEED_CTX(c);
  if (nsel < 1 || nsel > ndep) { rhb200_set_error("nsel must be 1 .. ndep"); return RHB200_EINVAL; }
  if (depths) for (int q = 0; q < nsel; q++) if (depths[q] < 0 || depths[q] >= ndep) { rhb200_set_error("depth index %d out of range", depths[q]); return RHB200_EINVAL; }
  if (!atmosphere || !rf || !par_rows || !par_delta) { rhb200_set_error("null buffer"); return RHB200_EINVAL; }
  if (nrow < 9 || atm_scale < 0 || atm_scale > 2 || iref < 0 || iref >= c->wav.nlambda || npar < 1 || ncol < 0) {
    rhb200_set_error("bad arguments"); return RHB200_EINVAL;
  }
  for (int p = 0; p < npar; p++)
    if (par_rows[p] < 0 || par_rows[p] >= nrow || !(par_delta[p] > 0.0)) { rhb200_set_error("parameter %d: row %d / delta %g", p, par_rows[p], par_delta[p]); return RHB200_EINVAL; }
  const long long nvirt = (long long) ncol * npar * nsel * 2;
  if (nvirt > 0x7fffffffLL) { rhb200_set_error("too many perturbed columns in one call (%lld)", nvirt); return RHB200_EINVAL; }
  if (ncol == 0) return RHB200_OK;
  RH_CUDA(cudaSetDevice(c->device));
  DevBuf base, rows, delta, sel;
  if (depths) RH_CHECK(sel.from_host(depths, (size_t) nsel * sizeof(int)));
  const int *d_sel = depths ? (const int *) sel.p : nullptr;
  RH_CHECK(base.alloc((size_t) ncol * nrow * ndep * sizeof(double)));
  RH_CHECK(rows.alloc((size_t) npar * sizeof(int)));
  RH_CHECK(delta.alloc((size_t) npar * sizeof(double)));
  RH_CUDA(cudaMemcpy(base.p, atmosphere, (size_t) ncol * nrow * ndep * sizeof(double), cudaMemcpyHostToDevice));
  RH_CUDA(cudaMemcpy(rows.p, par_rows, (size_t) npar * sizeof(int), cudaMemcpyHostToDevice));
  RH_CUDA(cudaMemcpy(delta.p, par_delta, (size_t) npar * sizeof(double), cudaMemcpyHostToDevice));
  // the depth-local route needs: no scale-row parameter (the scale couples the depths), every column moving
  // (VMACRO_TRESH = 0; atmos.moving is a property of the whole column) and no scattering iteration (J couples them)
  bool local = c->n_max_scatter <= 0 && !(vmacro_tresh > 0.0) && c->cont != nullptr;
  for (int p = 0; p < npar; p++) if (par_rows[p] == 0) local = false;
  if (const char *e = getenv("RHB200_RF_FD_BRUTE")) if (atoi(e) != 0) local = false;
  if (local)
    return rf_fd_single_depth(c, ncol, ndep, nrow, mu, atm_scale, iref, wght_per_H, bc_top, bc_bottom, npar,
                              (const int *) rows.p, (const double *) delta.p, (const double *) base.p, rf, par_rows, nsel, d_sel);
  PyrhIn py{nullptr, nrow, atm_scale, iref, wght_per_H, vmacro_tresh, nullptr};
  py.rf_npar = npar; py.d_rf_rows = (const int *) rows.p; py.d_rf_delta = (const double *) delta.p;
  py.d_base = (const double *) base.p; py.rf_out = rf; py.rf_nsel = nsel; py.d_rf_sel = d_sel;
  return lte_batch_host(c, (int) nvirt, ndep, mu, 1, bc_top, bc_bottom, nullptr, nullptr, nullptr, nullptr, nullptr, 1, &py);
}
static int to_host(void *h, const void *d, size_t bytes)
{
  if (bytes) RH_CUDA(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
  return RHB200_OK;
}

extern "C" int rhb200_ltepops_elem_batch(rhb200_ctx *c, int ncol, int ndep, const double *atmos, double *n)
{
  RH_NEED_CTX(c);
  RH_CHECK(need_state(c, false));
  if (ncol <= 0 || ndep <= 0 || !atmos || !n) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  DevBuf at, en, lp;
  const size_t nb = (size_t) ncol * c->tab.nelem * RHB200_RE_MAXSTAGE * ndep * sizeof(double);
  RH_CHECK(at.from_host(atmos, (size_t) ncol * RHB200_AT_NFIELD * ndep * sizeof(double)));
  RH_CHECK(en.alloc(nb));
  RH_CUDA(cudaMemset(en.p, 0, std::max<size_t>(nb, 8)));
  RH_CHECK(lp.alloc((size_t) ncol * std::max(1, c->tab.nline) * ndep * LP_NFIELD * sizeof(double)));
  RH_CHECK(rh_launch_prep(c, ncol, ndep, 1.0, 1, at.as<double>(), en.as<double>(), lp.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  return to_host(n, en.p, nb);
}

extern "C" int rhb200_rlk_opacity_batch(rhb200_ctx *c, int ncol, int ndep, double muz, int moving,
                                        int to_obs, const double *atmos, double *chi, double *eta, int *flags)
{
  RH_NEED_CTX(c);
  if (c->tab.rlkscatter) { rhb200_set_error("rhb200_rlk_opacity_batch returns chi / eta only: RLK_SCATTER = TRUE is part of the fused path"); return RHB200_EUNSUPPORTED; }
  RH_CHECK(need_state(c, true));
  if (ncol <= 0 || ndep <= 0 || !atmos || !chi || !eta) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  const int nl = c->wav.nlambda;
  DevBuf at, en, lp, dchi, deta;
  const size_t ob = (size_t) ncol * nl * 4 * ndep * sizeof(double);
  RH_CHECK(at.from_host(atmos, (size_t) ncol * RHB200_AT_NFIELD * ndep * sizeof(double)));
  RH_CHECK(en.alloc((size_t) ncol * std::max(1, c->tab.nelem) * RHB200_RE_MAXSTAGE * ndep * sizeof(double)));
  RH_CHECK(lp.alloc((size_t) ncol * std::max(1, c->tab.nline) * ndep * LP_NFIELD * sizeof(double)));
  RH_CHECK(dchi.alloc(ob)); RH_CHECK(deta.alloc(ob));
  RH_CHECK(rh_launch_prep(c, ncol, ndep, muz, moving, at.as<double>(), en.as<double>(), lp.as<double>()));
  RH_CHECK(rh_launch_opacity_raw(c, ncol, ndep, to_obs, at.as<double>(), lp.as<double>(),
                                 dchi.as<double>(), deta.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(chi, dchi.p, ob));
  RH_CHECK(to_host(eta, deta.p, ob));
  if (flags) memcpy(flags, c->h_flags.data(), (size_t) nl * sizeof(int));
  return RHB200_OK;
}

extern "C" int rhb200_molecular_opacity_batch(rhb200_ctx *c, int ncol, int ndep, double muz, int moving, int to_obs,
                                              int nmol, int nmline, const double *mlines,
                                              int ncomp, const int *zq, const double *zshift, const double *zstrength,
                                              double vmicro_char, int nlambda, const double *lambda,
                                              const double *atmos, const double *mol,
                                              double *chi, double *eta, int *flags)
{
  RH_NEED_CTX(c);
  if (ncol <= 0 || ndep <= 0 || nmol <= 0 || nmline <= 0 || nlambda <= 0 || !mlines || !lambda || !atmos || !mol ||
      !chi || !eta || ncomp < 0 || (ncomp > 0 && (!zq || !zshift || !zstrength))) {
    rhb200_set_error("bad arguments"); return RHB200_EINVAL;
  }
  // first / last line of each molecule (the reference's outer window test uses mrt[0] and mrt[Nrt-1])
  std::vector<int> mfirst(nmol, -1), mlast(nmol, -1);
  for (int n = 0; n < nmline; n++) {
    const double *L = mlines + (size_t) n * RHB200_ML_NFIELD;
    const int m = (int) L[RHB200_ML_MOL];
    if (m < 0 || m >= nmol) { rhb200_set_error("molecular line %d: molecule index out of range", n); return RHB200_EINVAL; }
    if (n > 0 && m < (int) L[RHB200_ML_MOL - RHB200_ML_NFIELD]) { rhb200_set_error("molecular lines must be grouped by molecule"); return RHB200_EINVAL; }
    const int zo = (int) L[RHB200_ML_ZOFF], nc = (int) L[RHB200_ML_NCOMP];
    if (zo < 0 || nc < 0 || zo + nc > ncomp) { rhb200_set_error("molecular line %d: Zeeman slice out of range", n); return RHB200_EINVAL; }
    if (mfirst[m] < 0) mfirst[m] = n;
    mlast[m] = n;
  }
  std::vector<int> first(nlambda, 0), count(nlambda, 0), idx, fl(nlambda, 0);
  const double vc = vmicro_char / RH_CLIGHT;
  for (int l = 0; l < nlambda; l++) {
    const double lam = lambda[l];
    first[l] = (int) idx.size();
    for (int m = 0; m < nmol; m++) {
      if (mfirst[m] < 0) continue;
      const double *L0 = mlines + (size_t) mfirst[m] * RHB200_ML_NFIELD, *LN = mlines + (size_t) mlast[m] * RHB200_ML_NFIELD;
      const double dl0 = lam * L0[RHB200_ML_QWING] * vc, dlN = lam * LN[RHB200_ML_QWING] * vc;        // opacity.c:774-777
      if (!(lam >= L0[RHB200_ML_LAMBDA0] - dl0 && lam <= LN[RHB200_ML_LAMBDA0] + dlN)) continue;     // :779-780
      for (int n = mfirst[m]; n <= mlast[m]; n++) {
        const double *L = mlines + (size_t) n * RHB200_ML_NFIELD;
        const double dl = lam * L[RHB200_ML_QWING] * vc;
        if (std::fabs(L[RHB200_ML_LAMBDA0] - lam) <= dl) {                                           // :784-786
          idx.push_back(n);
          fl[l] |= 1;
          if (L[RHB200_ML_POLARIZABLE] != 0.0) fl[l] |= 2;
        }
      }
    }
    count[l] = (int) idx.size() - first[l];
  }
  DevBuf dl_, df, dc, di, dm, dq, dsh, dst, dat, dmol, dchi, deta;
  const size_t ob = (size_t) ncol * nlambda * 4 * ndep * sizeof(double);
  RH_CHECK(dl_.from_host(lambda, (size_t) nlambda * sizeof(double)));
  RH_CHECK(df.from_host(first.data(), (size_t) nlambda * sizeof(int)));
  RH_CHECK(dc.from_host(count.data(), (size_t) nlambda * sizeof(int)));
  RH_CHECK(di.from_host(idx.data(), idx.size() * sizeof(int)));
  RH_CHECK(dm.from_host(mlines, (size_t) nmline * RHB200_ML_NFIELD * sizeof(double)));
  RH_CHECK(dq.from_host(zq, (size_t) ncomp * sizeof(int)));
  RH_CHECK(dsh.from_host(zshift, (size_t) ncomp * sizeof(double)));
  RH_CHECK(dst.from_host(zstrength, (size_t) ncomp * sizeof(double)));
  RH_CHECK(dat.from_host(atmos, (size_t) ncol * RHB200_AT_NFIELD * ndep * sizeof(double)));
  RH_CHECK(dmol.from_host(mol, (size_t) ncol * nmol * 3 * ndep * sizeof(double)));
  RH_CHECK(dchi.alloc(ob)); RH_CHECK(deta.alloc(ob));
  RH_CHECK(rh_launch_mol_opacity_raw(c, ncol, nlambda, ndep, nmol, muz, moving, to_obs, dl_.as<double>(), df.as<int>(),
                                     dc.as<int>(), di.as<int>(), dm.as<double>(), dq.as<int>(), dsh.as<double>(),
                                     dst.as<double>(), dat.as<double>(), dmol.as<double>(), dchi.as<double>(),
                                     deta.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(chi, dchi.p, ob));
  RH_CHECK(to_host(eta, deta.p, ob));
  if (flags) memcpy(flags, fl.data(), (size_t) nlambda * sizeof(int));
  return RHB200_OK;
}

extern "C" int rhb200_passive_bb_batch(rhb200_ctx *c, int ncol, int ndep, double muz, int moving, int to_obs,
                                       int nline, const double *plines, int ncomp, const double *c_shift,
                                       const double *c_fraction, double vmicro_char, int nlambda, const double *lambda,
                                       const double *atmos, const double *pcol, double *chi, double *eta, int *flags)
{
  RH_NEED_CTX(c);
  if (ncol <= 0 || ndep <= 0 || nline <= 0 || nlambda <= 0 || ncomp <= 0 || !plines || !c_shift || !c_fraction ||
      !lambda || !atmos || !pcol || !chi || !eta) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  for (int n = 0; n < nline; n++) {
    const double *L = plines + (size_t) n * RHB200_PB_NFIELD;
    const int nc = (int) L[RHB200_PB_NCOMP], off = (int) L[RHB200_PB_COMPOFF];
    if (nc < 1 || off < 0 || off + nc > ncomp) { rhb200_set_error("passive line %d: component slice out of range", n); return RHB200_EINVAL; }
  }
  std::vector<int> first(nlambda, 0), count(nlambda, 0), idx, fl(nlambda, 0);
  for (int l = 0; l < nlambda; l++) {
    first[l] = (int) idx.size();
    for (int n = 0; n < nline; n++) {
      const double *L = plines + (size_t) n * RHB200_PB_NFIELD;
      const double dlambda = L[RHB200_PB_LAMBDA0] * L[RHB200_PB_QWING] * (vmicro_char / RH_CLIGHT);   // metal.c:245-246
      if (std::fabs(lambda[l] - L[RHB200_PB_LAMBDA0]) <= dlambda) { idx.push_back(n); fl[l] |= 1; }     // :248
    }
    count[l] = (int) idx.size() - first[l];
  }
  DevBuf dl_, df, dc, di, dp_, dcs, dcf, dat, dcol, dchi, deta;
  const size_t ob = (size_t) ncol * nlambda * ndep * sizeof(double);
  RH_CHECK(dl_.from_host(lambda, (size_t) nlambda * sizeof(double)));
  RH_CHECK(df.from_host(first.data(), (size_t) nlambda * sizeof(int)));
  RH_CHECK(dc.from_host(count.data(), (size_t) nlambda * sizeof(int)));
  RH_CHECK(di.from_host(idx.data(), idx.size() * sizeof(int)));
  RH_CHECK(dp_.from_host(plines, (size_t) nline * RHB200_PB_NFIELD * sizeof(double)));
  RH_CHECK(dcs.from_host(c_shift, (size_t) ncomp * sizeof(double)));
  RH_CHECK(dcf.from_host(c_fraction, (size_t) ncomp * sizeof(double)));
  RH_CHECK(dat.from_host(atmos, (size_t) ncol * RHB200_AT_NFIELD * ndep * sizeof(double)));
  RH_CHECK(dcol.from_host(pcol, (size_t) ncol * nline * 4 * ndep * sizeof(double)));
  RH_CHECK(dchi.alloc(ob)); RH_CHECK(deta.alloc(ob));
  RH_CHECK(rh_launch_passive_bb(c, ncol, nlambda, ndep, nline, muz, moving, to_obs, dl_.as<double>(), df.as<int>(),
                                dc.as<int>(), di.as<int>(), dp_.as<double>(), dcs.as<double>(), dcf.as<double>(),
                                dat.as<double>(), dcol.as<double>(), dchi.as<double>(), deta.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(chi, dchi.p, ob));
  RH_CHECK(to_host(eta, deta.p, ob));
  if (flags) memcpy(flags, fl.data(), (size_t) nlambda * sizeof(int));
  return RHB200_OK;
}

extern "C" int rhb200_set_solvers(rhb200_ctx *c, int s_interpolation, int s_interpolation_stokes)
{
  RH_NEED_CTX(c);
  if (s_interpolation < RHB200_S_LINEAR || s_interpolation > RHB200_S_BEZIER3) {
    rhb200_set_error("Unknown radiation solver: %d", s_interpolation); return RHB200_EINVAL;     /* formal.c:240 */
  }
  if (s_interpolation_stokes != RHB200_DELO_PARABOLIC && s_interpolation_stokes != RHB200_DELO_BEZIER3) {
    rhb200_set_error("Unknown polarization solver: %d", s_interpolation_stokes); return RHB200_EINVAL;   /* formal.c:218 */
  }
  c->s_interpolation = s_interpolation;
  c->s_interpolation_stokes = s_interpolation_stokes;
  return RHB200_OK;
}

extern "C" int rhb200_stokes_bezier3_batch(rhb200_ctx *c, int nray, int ncol, int ndep, double muz, int to_obs,
                                           int bc_top, int bc_bottom, const int *ray_col,
                                           const double *ray_lambda, const double *height, const double *T,
                                           const double *chi, const double *S, const double *chiQUV,
                                           double *I, double *Psi)
{
  return rhb200_stokes_ray_batch(c, RHB200_DELO_BEZIER3, nray, ncol, ndep, muz, to_obs, bc_top, bc_bottom, ray_col,
                                 ray_lambda, height, T, chi, S, chiQUV, I, Psi);
}

extern "C" int rhb200_stokes_ray_batch(rhb200_ctx *c, int solver, int nray, int ncol, int ndep, double muz, int to_obs,
                                       int bc_top, int bc_bottom, const int *ray_col,
                                       const double *ray_lambda, const double *height, const double *T,
                                       const double *chi, const double *S, const double *chiQUV,
                                       double *I, double *Psi)
{
  RH_NEED_CTX(c);
  if (solver != RHB200_DELO_PARABOLIC && solver != RHB200_DELO_BEZIER3) {
    rhb200_set_error("Unknown polarization solver: %d", solver); return RHB200_EINVAL;
  }
  if (nray <= 0 || ncol <= 0 || ndep < 3 || !ray_col || !ray_lambda || !height || !T || !chi || !S || !chiQUV || !I) {
    rhb200_set_error("bad arguments"); return RHB200_EINVAL;
  }
  for (int r = 0; r < nray; r++) if (ray_col[r] < 0 || ray_col[r] >= ncol) { rhb200_set_error("ray_col[%d] out of range", r); return RHB200_EINVAL; }
  DevBuf rc, rl, h, t, dchi, dS, dq, dI, dP;
  const size_t rb = (size_t) nray * ndep * sizeof(double);
  RH_CHECK(rc.from_host(ray_col, (size_t) nray * sizeof(int)));
  RH_CHECK(rl.from_host(ray_lambda, (size_t) nray * sizeof(double)));
  RH_CHECK(h.from_host(height, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(t.from_host(T, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(dchi.from_host(chi, rb)); RH_CHECK(dS.from_host(S, 4*rb)); RH_CHECK(dq.from_host(chiQUV, 3*rb));
  RH_CHECK(dI.alloc(4*rb));
  if (Psi) RH_CHECK(dP.alloc(rb));
  RH_CHECK(rh_launch_delo_generic(c, solver, nray, ndep, muz, to_obs, bc_top, bc_bottom, rc.as<int>(), rl.as<double>(),
                                  h.as<double>(), t.as<double>(), dchi.as<double>(), dS.as<double>(),
                                  dq.as<double>(), dI.as<double>(), Psi ? dP.as<double>() : nullptr));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(I, dI.p, 4*rb));
  if (Psi) RH_CHECK(to_host(Psi, dP.p, rb));
  return RHB200_OK;
}

extern "C" int rhb200_bezier3_batch(rhb200_ctx *c, int nray, int ncol, int ndep, double muz, int to_obs,
                                    int bc_top, int bc_bottom, const int *ray_col, const double *ray_lambda,
                                    const double *height, const double *T, const double *chi, const double *S,
                                    double *I, double *Psi)
{
  return rhb200_scalar_ray_batch(c, RHB200_S_BEZIER3, nray, ncol, ndep, muz, to_obs, bc_top, bc_bottom, ray_col,
                                 ray_lambda, height, T, chi, S, I, Psi);
}

extern "C" int rhb200_scalar_ray_batch(rhb200_ctx *c, int solver, int nray, int ncol, int ndep, double muz, int to_obs,
                                       int bc_top, int bc_bottom, const int *ray_col, const double *ray_lambda,
                                       const double *height, const double *T, const double *chi, const double *S,
                                       double *I, double *Psi)
{
  RH_NEED_CTX(c);
  if (solver < RHB200_S_LINEAR || solver > RHB200_S_BEZIER3) {
    rhb200_set_error("Unknown radiation solver: %d", solver); return RHB200_EINVAL;
  }
  if (nray <= 0 || ncol <= 0 || ndep < 3 || !ray_col || !ray_lambda || !height || !T || !chi || !S || !I) {
    rhb200_set_error("bad arguments"); return RHB200_EINVAL;
  }
  for (int r = 0; r < nray; r++) if (ray_col[r] < 0 || ray_col[r] >= ncol) { rhb200_set_error("ray_col[%d] out of range", r); return RHB200_EINVAL; }
  DevBuf rc, rl, h, t, dchi, dS, dI, dP;
  const size_t rb = (size_t) nray * ndep * sizeof(double);
  RH_CHECK(rc.from_host(ray_col, (size_t) nray * sizeof(int)));
  RH_CHECK(rl.from_host(ray_lambda, (size_t) nray * sizeof(double)));
  RH_CHECK(h.from_host(height, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(t.from_host(T, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(dchi.from_host(chi, rb)); RH_CHECK(dS.from_host(S, rb));
  RH_CHECK(dI.alloc(rb));
  if (Psi) RH_CHECK(dP.alloc(rb));
  RH_CHECK(rh_launch_bezier3(c, solver, nray, ndep, muz, to_obs, bc_top, bc_bottom, rc.as<int>(), rl.as<double>(),
                             h.as<double>(), t.as<double>(), dchi.as<double>(), dS.as<double>(),
                             dI.as<double>(), Psi ? dP.as<double>() : nullptr));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(I, dI.p, rb));
  if (Psi) RH_CHECK(to_host(Psi, dP.p, rb));
  return RHB200_OK;
}

extern "C" int rhb200_bezier3_rf_batch(rhb200_ctx *c, int nray, int ncol, int ndep, double muz,
                                       int bc_top, int bc_bottom, const int *ray_col, const double *ray_lambda,
                                       const double *height, const double *T,
                                       const double *chi_dn, const double *S_dn, const double *chi_up, const double *S_up,
                                       int npar, const double *dchi, const double *deta, double *I, double *dI)
{
  RH_NEED_CTX(c);
  if (nray <= 0 || ncol <= 0 || ndep < 3 || !ray_col || !ray_lambda || !height || !T || !chi_dn || !S_dn ||
      !chi_up || !S_up || !dchi || !deta || !I || !dI) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  if (npar < 1 || npar > 16) { rhb200_set_error("npar must be in 1..16"); return RHB200_EINVAL; }
  for (int r = 0; r < nray; r++) if (ray_col[r] < 0 || ray_col[r] >= ncol) { rhb200_set_error("ray_col[%d] out of range", r); return RHB200_EINVAL; }
  DevBuf rc, rl, h, t, cd, sd, cu, su, dc, de, dIb, ddI;
  const size_t rb = (size_t) nray * ndep * sizeof(double);
  RH_CHECK(rc.from_host(ray_col, (size_t) nray * sizeof(int)));
  RH_CHECK(rl.from_host(ray_lambda, (size_t) nray * sizeof(double)));
  RH_CHECK(h.from_host(height, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(t.from_host(T, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(cd.from_host(chi_dn, rb)); RH_CHECK(sd.from_host(S_dn, rb));
  RH_CHECK(cu.from_host(chi_up, rb)); RH_CHECK(su.from_host(S_up, rb));
  RH_CHECK(dc.from_host(dchi, rb * npar)); RH_CHECK(de.from_host(deta, rb * npar));
  RH_CHECK(dIb.alloc(rb)); RH_CHECK(ddI.alloc(rb * npar));
  RH_CHECK(rh_launch_bezier3_rf(c, nray, ndep, muz, bc_top, bc_bottom, rc.as<int>(), rl.as<double>(), h.as<double>(),
                                t.as<double>(), cd.as<double>(), sd.as<double>(), cu.as<double>(), su.as<double>(),
                                npar, dc.as<double>(), de.as<double>(), dIb.as<double>(), ddI.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(I, dIb.p, rb));
  RH_CHECK(to_host(dI, ddI.p, rb * npar));
  return RHB200_OK;
}

extern "C" int rhb200_feautrier_batch(rhb200_ctx *c, int nray, int ncol, int ndep, double muz,
                                      int bc_top, int bc_bottom, const int *ray_col, const double *ray_lambda,
                                      const double *height, const double *T, const double *chi, const double *S,
                                      double *P, double *Psi, double *Iem)
{
  RH_NEED_CTX(c);
  if (nray <= 0 || ncol <= 0 || ndep < 3 || !ray_col || !ray_lambda || !height || !T || !chi || !S || !P || !Iem) {
    rhb200_set_error("bad arguments"); return RHB200_EINVAL;
  }
  for (int r = 0; r < nray; r++) if (ray_col[r] < 0 || ray_col[r] >= ncol) { rhb200_set_error("ray_col[%d] out of range", r); return RHB200_EINVAL; }
  DevBuf rc, rl, h, t, dchi, dS, dP, dPsi, dI, scr;
  const size_t rb = (size_t) nray * ndep * sizeof(double);
  RH_CHECK(rc.from_host(ray_col, (size_t) nray * sizeof(int)));
  RH_CHECK(rl.from_host(ray_lambda, (size_t) nray * sizeof(double)));
  RH_CHECK(h.from_host(height, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(t.from_host(T, (size_t) ncol * ndep * sizeof(double)));
  RH_CHECK(dchi.from_host(chi, rb)); RH_CHECK(dS.from_host(S, rb));
  RH_CHECK(dP.alloc(rb)); RH_CHECK(dI.alloc((size_t) nray * sizeof(double))); RH_CHECK(scr.alloc(2*rb));
  if (Psi) RH_CHECK(dPsi.alloc(rb));
  RH_CHECK(rh_launch_feautrier(c, nray, ndep, muz, bc_top, bc_bottom, rc.as<int>(), rl.as<double>(),
                               h.as<double>(), t.as<double>(), dchi.as<double>(), dS.as<double>(),
                               dP.as<double>(), Psi ? dPsi.as<double>() : nullptr, dI.as<double>(), scr.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(P, dP.p, rb));
  RH_CHECK(to_host(Iem, dI.p, (size_t) nray * sizeof(double)));
  if (Psi) RH_CHECK(to_host(Psi, dPsi.p, rb));
  return RHB200_OK;
}

extern "C" int rhb200_voigt_humlicek(rhb200_ctx *c, int n, const double *a, const double *v,
                                     double *H, double *F, int *region)
{
  RH_NEED_CTX(c);
  if (n <= 0 || !a || !v || !H || !F) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  DevBuf da, dv, dH, dF, dR;
  const size_t b = (size_t) n * sizeof(double);
  RH_CHECK(da.from_host(a, b)); RH_CHECK(dv.from_host(v, b));
  RH_CHECK(dH.alloc(b)); RH_CHECK(dF.alloc(b));
  if (region) RH_CHECK(dR.alloc((size_t) n * sizeof(int)));
  RH_CHECK(rh_launch_voigt(c, n, da.as<double>(), dv.as<double>(), dH.as<double>(), dF.as<double>(),
                           region ? dR.as<int>() : nullptr));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(H, dH.p, b)); RH_CHECK(to_host(F, dF.p, b));
  if (region) RH_CHECK(to_host(region, dR.p, (size_t) n * sizeof(int)));
  return RHB200_OK;
}

extern "C" int rhb200_voigt_armstrong(rhb200_ctx *c, int n, const double *a, const double *v,
                                      double *H, int *region)
{
  RH_NEED_CTX(c);
  if (n <= 0 || !a || !v || !H) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  DevBuf da, dv, dH, dR;
  const size_t b = (size_t) n * sizeof(double);
  RH_CHECK(da.from_host(a, b)); RH_CHECK(dv.from_host(v, b)); RH_CHECK(dH.alloc(b));
  if (region) RH_CHECK(dR.alloc((size_t) n * sizeof(int)));
  RH_CHECK(rh_launch_voigt_armstrong(c, n, da.as<double>(), dv.as<double>(), dH.as<double>(),
                                     region ? dR.as<int>() : nullptr));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  RH_CHECK(to_host(H, dH.p, b));
  if (region) RH_CHECK(to_host(region, dR.p, (size_t) n * sizeof(int)));
  return RHB200_OK;
}

extern "C" int rhb200_math_probe(rhb200_ctx *c, int n, int func, const double *x, const double *y, double *out)
{
  RH_NEED_CTX(c);
  if (n <= 0 || !x || !out || func < 0 || func > 8 || (func >= 3 && func <= 5 && !y)) { rhb200_set_error("bad arguments"); return RHB200_EINVAL; }
  DevBuf dx, dy, dout;
  const size_t b = (size_t) n * sizeof(double);
  RH_CHECK(dx.from_host(x, b));
  if (y) RH_CHECK(dy.from_host(y, b)); else RH_CHECK(dy.alloc(8));
  RH_CHECK(dout.alloc(b));
  RH_CHECK(rh_launch_math_probe(c, n, func, dx.as<double>(), dy.as<double>(), dout.as<double>()));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  return to_host(out, dout.p, b);
}

// -------------------------------------------------------- memory helpers
extern "C" int rhb200_dev_alloc(rhb200_ctx *c, size_t bytes, void **dptr)
{
  RH_NEED_CTX(c);
  if (!dptr) { rhb200_set_error("null dptr"); return RHB200_EINVAL; }
  cudaError_t e = cudaMalloc(dptr, std::max<size_t>(bytes, 8));
  if (e != cudaSuccess) { cudaGetLastError(); rhb200_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return RHB200_ENOMEM; }
  return RHB200_OK;
}
extern "C" int rhb200_dev_free(rhb200_ctx *c, void *dptr) { RH_NEED_CTX(c); RH_CUDA(cudaFree(dptr)); return RHB200_OK; }
extern "C" int rhb200_host_alloc_pinned(size_t bytes, void **hptr)
{
  if (!hptr) { rhb200_set_error("null hptr"); return RHB200_EINVAL; }
  cudaError_t e = cudaHostAlloc(hptr, std::max<size_t>(bytes, 8), cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); rhb200_set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); return RHB200_ENOMEM; }
  return RHB200_OK;
}
extern "C" int rhb200_host_free_pinned(void *hptr) { RH_CUDA(cudaFreeHost(hptr)); return RHB200_OK; }
extern "C" int rhb200_memcpy_h2d(rhb200_ctx *c, void *dst, const void *src, size_t bytes)
{ RH_NEED_CTX(c); RH_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return RHB200_OK; }
extern "C" int rhb200_memcpy_d2h(rhb200_ctx *c, void *dst, const void *src, size_t bytes)
{ RH_NEED_CTX(c); RH_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return RHB200_OK; }
extern "C" int rhb200_synchronize(rhb200_ctx *c) { RH_NEED_CTX(c); RH_CUDA(cudaDeviceSynchronize()); return RHB200_OK; }

extern "C" int rhb200_flush_l2(rhb200_ctx *c)
{
  RH_NEED_CTX(c);
  const size_t bytes = (size_t) 256 << 20;      // 256 MiB > 126 MB L2
  if (!c->flush) {
    cudaError_t e = cudaMalloc(&c->flush, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); rhb200_set_error("flush buffer: %s", cudaGetErrorString(e)); return RHB200_ENOMEM; }
    c->flush_bytes = bytes;
  }
  static int v = 0;
  RH_CUDA(cudaMemsetAsync(c->flush, (++v) & 0xff, c->flush_bytes, c->stream));
  RH_CUDA(cudaStreamSynchronize(c->stream));
  return RHB200_OK;
}

// --------------------------------------------------------- instrumentation
extern "C" int rhb200_timing_enable(rhb200_ctx *c, int on) { if (!c) return RHB200_EINVAL; c->timing = on != 0; return RHB200_OK; }
extern "C" int rhb200_timing_reset(rhb200_ctx *c)
{
  if (!c) return RHB200_EINVAL;
  for (int i = 0; i < RHB200_K_COUNT; i++) { c->k_ms[i] = 0.0; c->k_launch[i] = 0; }
  return RHB200_OK;
}
extern "C" int rhb200_timing_get(rhb200_ctx *c, int which, double *ms, long *launches)
{
  if (!c || which < 0 || which >= RHB200_K_COUNT) return RHB200_EINVAL;
  if (ms) *ms = c->k_ms[which];
  if (launches) *launches = c->k_launch[which];
  return RHB200_OK;
}
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
extern "C" int rhb200_timer_begin(rhb200_ctx *c)
{
  RH_NEED_CTX(c);
  if (!g_t0) { RH_CUDA(cudaEventCreate(&g_t0)); RH_CUDA(cudaEventCreate(&g_t1)); }
  RH_CUDA(cudaDeviceSynchronize());
  RH_CUDA(cudaEventRecord(g_t0, c->stream));
  return RHB200_OK;
}
extern "C" int rhb200_timer_end(rhb200_ctx *c, double *ms)
{
  RH_NEED_CTX(c);
  if (!g_t0 || !ms) { rhb200_set_error("timer not started"); return RHB200_EINVAL; }
  RH_CUDA(cudaEventRecord(g_t1, c->stream));
  RH_CUDA(cudaEventSynchronize(g_t1));
  RH_CUDA(cudaDeviceSynchronize());
  float f = 0.f;
  RH_CUDA(cudaEventElapsedTime(&f, g_t0, g_t1));
  *ms = f;
  return RHB200_OK;
}
extern "C" int rhb200_fp64_peak(rhb200_ctx *c, double *tf_fma, double *tf_nofma)
{
  RH_NEED_CTX(c);
  return rh_fp64_peak(c, tf_fma, tf_nofma);
}
