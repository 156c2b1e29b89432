// rhb200_nlte_front.cuh -- NLTE through the drop-in call: rhb200_nlte_compute1d_batch.
// Included at the end of rhb200_nlte.cu (same translation unit as the MALI engine).
//
// Reference: rhf1d() with input.solve_NLTE (rh/rhf1d/pyrh_compute1dray.c:112-388), SetLTEQuantities / CollisionRate
// (rh/ltepops.c:224-249, rh/collision.c:450-946), Background (rh/background.c:139-700), getProfiles / Damping
// (rh/profile.c:470-500, rh/broad.c:273-314), initSolution (rh/initial_xdr.c:59-414), initScatter (rh/initscatter.c:32-76),
// _solveray (rh/rhf1d/pyrh_solveray.c:75-187).  Everything per column is evaluated on the device; the host part is
// the chunk loop and the two convergence read-backs per iteration the engine already does.

namespace {

#define NF_MAXATOM 16
struct FrontAtoms {
  int n;                          // ACTIVE atoms
  int model[NF_MAXATOM];          // index among the model atoms
  int first_lev[NF_MAXATOM];      // first row in the model-atom level table
  double abund[NF_MAXATOM];       // atom->abundance
};

// Linear() (linear.c:22-51) / splineEval() (spline.c:70-100) of one collisional coefficient table at temperature x
__device__ __forceinline__ double coll_interp(const double *__restrict__ xt, const double *__restrict__ yt,
                                              const double *__restrict__ M, int n, double x)
{
  const bool ascend = xt[1] > xt[0];
  const double xmin = ascend ? xt[0] : xt[n-1], xmax = ascend ? xt[n-1] : xt[0];
  if (x <= xmin) return ascend ? yt[0] : yt[n-1];
  if (x >= xmax) return ascend ? yt[n-1] : yt[0];
  int lo = 0, hi = n;
  const bool asc2 = xt[n-1] > xt[0];
  while (hi - lo > 1) { const int mid = (hi + lo) >> 1; if (asc2 ? (x >= xt[mid]) : (x <= xt[mid])) lo = mid; else hi = mid; }
  if (n > 2) {
    const double hj = xt[lo+1] - xt[lo];
    const double fx = (x - xt[lo]) / hj;
    const double fx1 = 1 - fx;
    return fx1*yt[lo] + fx*yt[lo+1] + (fx1*(fx1*fx1 - 1) * M[lo] + fx*(fx*fx - 1) * M[lo+1]) * (hj*hj)/6.0;
  }
  const double fx = (xt[lo+1] - x) / (xt[lo+1] - xt[lo]);
  return fx*yt[lo] + (1 - fx)*yt[lo+1];
}

// CollisionRate (collision.c:450-946) for the keywords of the shipped atoms: one thread per (column, depth) walks the
// records in file order, so every C[ij][k] receives its contributions in the reference's order.  pops: the LTE
// populations right after LTEpops(), i.e. BEFORE ChemicalEquilibrium rescales them (SetLTEQuantities runs first).
__global__ void __launch_bounds__(128)
nlte_collision_kernel(FrontAtoms A, Plan P, int ncol, int nlev_model, int H_nlevel, double C0,
                      int ncoll, const double *__restrict__ coll, const double *__restrict__ cT,
                      const double *__restrict__ cC, const double *__restrict__ cM,
                      const double *__restrict__ atmos, const double *__restrict__ pops, double *__restrict__ Cout)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * N) return;
  const int col = (int) (t / N), k = (int) (t % N);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * N;
  const double T = at[RHB200_AT_T*N + k], ne = at[RHB200_AT_NE*N + k];
  const double *pp = pops + (size_t) col * nlev_model * N + k;
  double *Cc = Cout + (size_t) col * P.ngam * N + k;
  for (int g = 0; g < P.ngam; g++) Cc[(size_t) g * N] = 0.0;
  for (int r = 0; r < ncoll; r++) {
    const double *R = coll + (size_t) r * RHB200_CO_NFIELD;
    const int a = (int) R[RHB200_CO_ATOM], type = (int) R[RHB200_CO_TYPE], i = (int) R[RHB200_CO_I], j = (int) R[RHB200_CO_J];
    const int nt = (int) R[RHB200_CO_NT], off = (int) R[RHB200_CO_TOFF], Nl = P.atom_nlevel[a];
    const double Ck = coll_interp(cT + off, cC + off, cM + off, nt, T);
    double *Cij = Cc + (size_t) (P.gam_off[a] + i*Nl + j) * N, *Cji = Cc + (size_t) (P.gam_off[a] + j*Nl + i) * N;
    const double ns_i = pp[(size_t) (A.first_lev[a] + i) * N], ns_j = pp[(size_t) (A.first_lev[a] + j) * N];
    switch (type) {
    case RHB200_CO_OMEGA: {                                    // collision.c:690-697
      // atom->g[j] travels in the record's spare field
      const double Cdown = C0 * ne * Ck / (R[7] * sqrt(T));
      *Cij += Cdown; *Cji += Cdown * ns_j/ns_i; break; }
    case RHB200_CO_CE: {                                       // :702-708 (gij = g[i]/g[j] in the spare field)
      const double Cdown = Ck * ne * R[7] * sqrt(T);
      *Cij += Cdown; *Cji += Cdown * ns_j/ns_i; break; }
    case RHB200_CO_CI: {                                       // :713-720
      const double Cup = Ck * ne * rhm::rh_exp(-R[RHB200_CO_DE]/(RH_KBOLTZMANN*T)) * sqrt(T);
      *Cji += Cup; *Cij += Cup * ns_i/ns_j; break; }
    case RHB200_CO_CP: {                                       // :725-731, np = atmos.H->n[Nlevel-1]
      const double Cdown = pp[(size_t) (H_nlevel-1) * N] * Ck;
      *Cij += Cdown; *Cji += Cdown * ns_j/ns_i; break; }
    case RHB200_CO_CH: {                                       // :736-741
      const double Cup = pp[0] * Ck;
      *Cji += Cup; *Cij += Cup * ns_i/ns_j; break; }
    case RHB200_CO_CH0: *Cij += pp[0] * Ck; break;             // :746-747
    default: *Cji += pp[(size_t) (H_nlevel-1) * N] * Ck; break;   // CH+, :753-755
    }
  }
}

// ntotal = abundance * nHtot of every model atom (readatom.c:188-190); ChemicalEquilibrium then overwrites the atoms
// that are bound in molecules (chemequil.c:342)
__global__ void __launch_bounds__(128)
nlte_ntot_init_kernel(int ncol, int ndep, int natom_model, const double *__restrict__ abund_model,
                      const double *__restrict__ atmos, double *__restrict__ ntot)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * natom_model * ndep) return;
  const int k = (int) (t % ndep), a = (int) ((t / ndep) % natom_model), col = (int) (t / ((size_t) ndep * natom_model));
  ntot[t] = abund_model[a] * atmos[((size_t) col * RHB200_AT_NFIELD + RHB200_AT_NHTOT) * ndep + k];
}

// dense per-column inputs of the engine from the atmosphere block and the model-atom populations: T, vel, nstar,
// ntotal of the ACTIVE atoms, and n = nstar (initSolution, LTE_POPULATIONS: initial_xdr.c:286-291)
__global__ void __launch_bounds__(128)
nlte_gather_kernel(FrontAtoms A, Plan P, Cols C, int ncol, int nlev_model, int natom_model, const double *__restrict__ atmos,
                   const double *__restrict__ pops, const double *__restrict__ ntot, double *__restrict__ n_out)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * N) return;
  const int col = (int) (t / N), k = (int) (t % N);
  const double *at = atmos + (size_t) col * RHB200_AT_NFIELD * N;
  ((double *) C.T)[t] = at[RHB200_AT_T*N + k];
  ((double *) C.vel)[t] = at[RHB200_AT_VEL*N + k];
  for (int a = 0; a < A.n; a++) {
    ((double *) C.ntotal)[((size_t) col * P.Natom + a) * N + k] = ntot[((size_t) col * natom_model + A.model[a]) * N + k];
    for (int i = 0; i < P.atom_nlevel[a]; i++) {
      const double v = pops[((size_t) col * nlev_model + A.first_lev[a] + i) * N + k];
      const size_t o = ((size_t) col * P.nlev + P.lev_off[a] + i) * N + k;
      ((double *) C.nstar)[o] = v;
      n_out[o] = v;
    }
  }
}

// second Background() (_solveray): nstar is re-derived from the ntotal the first ChemicalEquilibrium() left, rescaled by
// the second one, which also multiplies the NLTE populations of ACTIVE atoms bound in molecules by its fraction
// (chemequil.c:336-341: atom->n != atom->nstar there)
__global__ void __launch_bounds__(128)
nlte_regather_kernel(FrontAtoms A, Plan P, Cols C, int ncol, int nlev_model, int natom_model,
                     const double *__restrict__ pops, const double *__restrict__ chem, double *__restrict__ n_io)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * N) return;
  const int col = (int) (t / N), k = (int) (t % N);
  for (int a = 0; a < A.n; a++) {
    const double fraction = chem[((size_t) col * (natom_model + 4) + A.model[a]) * N + k];
    for (int i = 0; i < P.atom_nlevel[a]; i++) {
      const size_t o = ((size_t) col * P.nlev + P.lev_off[a] + i) * N + k;
      ((double *) C.nstar)[o] = pops[((size_t) col * nlev_model + A.first_lev[a] + i) * N + k];
      n_io[o] *= fraction;
    }
  }
}

__global__ void __launch_bounds__(128)
nlte_height_kernel(int ncol, int ndep, const double *__restrict__ atmos, double *__restrict__ height)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * ndep) return;
  const int col = (int) (t / ndep), k = (int) (t % ndep);
  height[t] = atmos[((size_t) col * RHB200_AT_NFIELD + RHB200_AT_HEIGHT) * ndep + k];
}

// populations Background() reads through atom->n: the LTE ones, except that an ACTIVE hydrogen atom shows its NLTE
// populations there -- zero until initSolution() (first Background() call), the converged ones in _solveray()
__global__ void __launch_bounds__(128)
nlte_popsn_kernel(int ncol, int ndep, int nlev_model, int H_nlevel, const double *__restrict__ pops,
                  const double *__restrict__ nH /* [ncol][nlevH...] engine layout or NULL = zeros */, int nlev_engine,
                  int H_lev_off, double *__restrict__ popsn)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) ncol * nlev_model * ndep) return;
  const int k = (int) (t % ndep), l = (int) ((t / ndep) % nlev_model), col = (int) (t / ((size_t) ndep * nlev_model));
  double v = pops[t];
  if (l < H_nlevel) v = nH ? nH[((size_t) col * nlev_engine + H_lev_off + l) * ndep + k] : 0.0;
  popsn[t] = v;
}

// adamp, vbroad of the ACTIVE lines from the damping kernel's [ncol][nline][4][ndep] block
__global__ void __launch_bounds__(128)
nlte_damping_gather_kernel(Plan P, Cols C, int ncol, const double *__restrict__ pcol)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nline * N) return;
  const int k = (int) (t % N), li = (int) ((t / N) % P.nline), col = (int) (t / ((size_t) N * P.nline));
  const double *p = pcol + (((size_t) col * P.nline + li) * 4) * N + k;
  ((double *) C.adamp)[t] = p[3*(size_t) N];
  const int a = (int) P.trans[(size_t) P.line_tr[li] * TR_NFIELD + TR_ATOM];
  ((double *) C.vbroad)[((size_t) col * P.Natom + a) * N + k] = p[2*(size_t) N];    // same value from every line of the atom
}

// adjustStokesMode() runs Profile() -- and with it Damping() -- for the POLARIZABLE lines only (zeeman.c:329-341)
__global__ void __launch_bounds__(128)
nlte_adamp_select_kernel(Plan P, int ncol, const double *__restrict__ src, double *__restrict__ dst)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.Ndep;
  if (t >= (size_t) ncol * P.nline * N) return;
  const int li = (int) ((t / N) % P.nline);
  if (P.line_pol[li]) dst[t] = src[t];
}

struct FrontState {                 // kept with the context between calls of rhb200_nlte_compute1d_batch
  uint64_t key = 0;
  int cc = 0;
  NlteEngine E, F;
  DevArena ar;
  double *d_in = nullptr, *d_at = nullptr, *d_pops = nullptr, *d_popsn = nullptr, *d_chem = nullptr, *d_ntot = nullptr, *d_tprep = nullptr,
         *d_sc = nullptr, *d_pc = nullptr, *d_apc = nullptr, *d_elem_n = nullptr, *d_lineprep = nullptr, *d_md = nullptr, *d_mol = nullptr,
         *d_mchi = nullptr, *d_meta = nullptr, *d_spec = nullptr, *d_abund = nullptr, *d_coll = nullptr, *d_cT = nullptr, *d_cC = nullptr,
         *d_cM = nullptr, *d_plrows = nullptr, *d_adamp2 = nullptr, *d_chi2 = nullptr, *d_eta2 = nullptr, *d_sca2 = nullptr,
         *d_quv = nullptr;
};

struct FrontDebug {                 // device copies kept for rhb200_nlte_front_debug (test hook)
  std::vector<std::vector<double>> arr;
};
FrontDebug g_front_debug;

}  // namespace

void rh_nlte_front_free(rhb200_ctx *c)
{
  if (c->nlte_front) { delete (FrontState *) c->nlte_front; c->nlte_front = nullptr; }
}

extern "C" int rhb200_nlte_front_debug(rhb200_ctx *c, int which, double *out, size_t count)
{
  if (!c || !out || which < 0 || which >= (int) g_front_debug.arr.size() || g_front_debug.arr[which].size() != count) {
    rhb200_set_error("rhb200_nlte_front_debug: nothing recorded for which = %d with %zu entries (have %zu)", which, count,
                     (which >= 0 && which < (int) g_front_debug.arr.size()) ? g_front_debug.arr[which].size() : (size_t) 0);
    return RHB200_EINVAL;
  }
  memcpy(out, g_front_debug.arr[which].data(), count * sizeof(double));
  return RHB200_OK;
}

extern "C" int rhb200_nlte_compute1d_stokes_batch(rhb200_ctx *c, const rhb200_nlte_plan *pl, const rhb200_nlte_front *fr,
                                                  int ncol, int ndep, int nrow, double mu, int atm_scale, const double *atmosphere,
                                                  int iref, double wght_per_H, double vmacro_tresh,
                                                  double *spectrum, double *quv, double *out_n, double *out_nstar, int *niter_out,
                                                  int *passes_out, double *scales);
extern "C" int rhb200_nlte_compute1d_batch(rhb200_ctx *c, const rhb200_nlte_plan *pl, const rhb200_nlte_front *fr,
                                           int ncol, int ndep, int nrow, double mu, int atm_scale, const double *atmosphere,
                                           int iref, double wght_per_H, double vmacro_tresh,
                                           double *spectrum, double *out_n, double *out_nstar, int *niter_out, int *passes_out,
                                           double *scales)
{
  return rhb200_nlte_compute1d_stokes_batch(c, pl, fr, ncol, ndep, nrow, mu, atm_scale, atmosphere, iref, wght_per_H, vmacro_tresh,
                                            spectrum, nullptr, out_n, out_nstar, niter_out, passes_out, scales);
}

extern "C" int rhb200_nlte_compute1d_stokes_batch(rhb200_ctx *c, const rhb200_nlte_plan *pl, const rhb200_nlte_front *fr,
                                                  int ncol, int ndep, int nrow, double mu, int atm_scale, const double *atmosphere,
                                                  int iref, double wght_per_H, double vmacro_tresh,
                                                  double *spectrum, double *quv, double *out_n, double *out_nstar, int *niter_out,
                                                  int *passes_out, double *scales)
{
  if (!c || !pl || !fr || !fr->plan1 || !atmosphere) { rhb200_set_error("null argument"); return RHB200_EINVAL; }
  RH_CUDA(cudaSetDevice(c->device));
  if (ncol <= 0 || nrow < 9 || ndep != pl->Ndep || atm_scale < 0 || atm_scale > 2 || !(mu > 0.0 && mu <= 1.0)) {
    rhb200_set_error("rhb200_nlte_compute1d_batch: bad ncol / nrow / ndep / atm_scale / mu"); return RHB200_EINVAL;
  }
  if (c->wav.nlambda != pl->Nspect) { rhb200_set_error("rhb200_set_wavelengths() must hold plan->lambda (%d vs %d wavelengths)", c->wav.nlambda, pl->Nspect); return RHB200_ESTATE; }
  if (iref < 0 || iref >= pl->Nspect) { rhb200_set_error("iref outside the wavelength grid"); return RHB200_EINVAL; }
  if (!c->cont || !rh_continuum_has_chemistry(c)) { rhb200_set_error("rhb200_set_continuum() / rhb200_set_chemistry() have not been called"); return RHB200_ESTATE; }
  if (fr->stokes < 0 || fr->stokes > 3) { rhb200_set_error("front->stokes must be 0 (NO_STOKES), 1 (FIELD_FREE), 2 (FULL_STOKES) or 3 (POLARIZATION_FREE)"); return RHB200_EUNSUPPORTED; }
  const bool stokes = fr->stokes != 0, full_stokes = fr->stokes == 2, pol_free = fr->stokes == 3;
  bool prd = false;
  if (fr->line_prd && fr->PRD_NmaxIter > 0) for (int l = 0; l < pl->nline; l++) prd = prd || fr->line_prd[l] != 0;
  if (!c->no_stokes) { rhb200_set_error("the background of the NLTE path is set up with rhb200_set_stokes_mode(ctx, 0): the FULL_STOKES passes of FIELD_FREE are selected by front->stokes"); return RHB200_EUNSUPPORTED; }
  if (stokes && (!fr->line_pol || !fr->line_zoff)) { rhb200_set_error("front->stokes needs line_pol / line_zoff and the Zeeman tables"); return RHB200_EINVAL; }
  if (pl->Natom > NF_MAXATOM) { rhb200_set_error("too many ACTIVE atoms"); return RHB200_EUNSUPPORTED; }
  if (vmacro_tresh > 0.0) { rhb200_set_error("VMACRO_TRESH > 0 (columns that may be static) is not implemented on the NLTE path"); return RHB200_EUNSUPPORTED; }
  if (scales && atm_scale == 2 && !(c->gravity > 0.0)) { rhb200_set_error("scales on a height grid: the column-mass row needs rhb200_set_gravity() (multiatmos.c:153-155)"); return RHB200_EINVAL; }
  const int N = ndep, Ns = pl->Nspect;
  const int nlev_model = rh_continuum_nlev(c), natom_model = rh_continuum_natom(c);
  const int H_nlevel = rh_continuum_proton_level(c) + 1;

  FrontAtoms A{};
  A.n = pl->Natom;
  int H_engine_atom = -1;
  for (int a = 0; a < pl->Natom; a++) {
    A.model[a] = fr->atom_model[a];
    if (A.model[a] < 0 || A.model[a] >= natom_model) { rhb200_set_error("ACTIVE atom %d: model index out of range", a); return RHB200_EINVAL; }
    A.first_lev[a] = rh_continuum_atom_first(c, A.model[a]);
    A.abund[a] = rh_continuum_abundance(c, A.model[a]);
    if (rh_continuum_atom_first(c, A.model[a] + 1) - A.first_lev[a] != pl->atom_nlevel[a]) { rhb200_set_error("ACTIVE atom %d: level count differs from the model atom's", a); return RHB200_EINVAL; }
    if (A.model[a] == 0) H_engine_atom = a;
  }
  const bool H_active = H_engine_atom >= 0;
  for (int r = 0; r < fr->ncoll; r++) {
    const double *R = fr->coll + (size_t) r * RHB200_CO_NFIELD;
    const int a = (int) R[RHB200_CO_ATOM], i = (int) R[RHB200_CO_I], j = (int) R[RHB200_CO_J], nt = (int) R[RHB200_CO_NT], off = (int) R[RHB200_CO_TOFF];
    if (a < 0 || a >= pl->Natom || i < 0 || j <= i || j >= pl->atom_nlevel[a] || nt < 2 || off < 0 || off + nt > fr->ncolltab ||
        (int) R[RHB200_CO_TYPE] < 0 || (int) R[RHB200_CO_TYPE] > RHB200_CO_CHPLUS) {
      rhb200_set_error("collision record %d is malformed", r); return RHB200_EINVAL;
    }
  }

  const bool trace = getenv("RHB200_NLTE_TRACE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_prev = now();
  RhRange whole("rhf1d (NLTE, batch)");
  auto mark = [&](const char *what) {
    nvtxMarkA(what);
    if (!trace) return;
    cudaStreamSynchronize(c->stream);
    const double t = now();
    fprintf(stderr, "[rhb200 nlte] %-28s %9.3f ms\n", what, t - t_prev);
    t_prev = t;
  };
  // ---- state kept with the context between calls (plans, engines, work arrays): rebuilt only when the problem's
  //      structure, the tables' sizes, the chunk size or mu change -- an inversion calls this thousands of times
  rhb200_nlte_plan p1 = *fr->plan1;
  const double mu1[1] = {mu}, w1[1] = {1.0};
  p1.muz = mu1; p1.wmu = w1; p1.Nrays = 1;
  const int nline_k = c->tab.nline, nelem = c->tab.nelem;
  const bool mol_on = c->wav.nmw > 0;
  if (c->wav.mol_pol) { rhb200_set_error("polarizable molecular lines with ACTIVE atoms are not implemented"); return RHB200_EUNSUPPORTED; }
  uint64_t key = 1469598103934665603ull;
  auto mix = [&](const void *ptr, size_t bytes) {
    const unsigned char *q = (const unsigned char *) ptr;
    for (size_t i = 0; i < bytes; i++) { key ^= q[i]; key *= 1099511628211ull; }
  };
  {
    const int ints[] = {pl->Nspect, pl->Nrays, pl->Ndep, pl->Natom, pl->Ntrans, pl->moving, pl->Ngorder, pl->Ngdelay, pl->Ngperiod,
                        pl->isum, pl->bc_top, pl->bc_bottom, pl->ntrl, pl->nphirow, pl->nline, fr->ncoll, fr->ncolltab, nrow, ncol,
                        nline_k, nelem, c->wav.npl, c->wav.nmsel, c->wav.nmw, nlev_model, natom_model, c->s_interpolation,
                        c->nlte_exact_rates, getenv("RHB200_NLTE_EXACT") ? atoi(getenv("RHB200_NLTE_EXACT")) + 2 : 0,
                        getenv("RHB200_NLTE_GAMMA_SEG") ? atoi(getenv("RHB200_NLTE_GAMMA_SEG")) : 0,
                        getenv("RHB200_NLTE_CHUNK_COLS") ? atoi(getenv("RHB200_NLTE_CHUNK_COLS")) : 0};
    mix(ints, sizeof ints); mix(&mu, sizeof mu);
    mix(pl->lambda, sizeof(double) * pl->Nspect); mix(pl->muz, sizeof(double) * pl->Nrays); mix(pl->wmu, sizeof(double) * pl->Nrays);
    mix(pl->atom_nlevel, sizeof(int) * pl->Natom); mix(pl->trans, sizeof(double) * pl->Ntrans * RHB200_TR_NFIELD);
    mix(fr->plan1->trans, sizeof(double) * pl->Ntrans * RHB200_TR_NFIELD);
    mix(pl->tr_lambda, sizeof(double) * pl->ntrl); mix(pl->tr_wlambda, sizeof(double) * pl->ntrl); mix(pl->tr_alpha, sizeof(double) * pl->ntrl);
    mix(pl->as_first, sizeof(int) * (pl->Nspect + 1)); mix(pl->as_trans, sizeof(int) * pl->as_first[pl->Nspect]);
    mix(pl->bg_hasline, sizeof(int) * pl->Nspect); mix(fr->atom_model, sizeof(int) * pl->Natom);
    mix(fr->coll, sizeof(double) * fr->ncoll * RHB200_CO_NFIELD); mix(fr->coll_T, sizeof(double) * fr->ncolltab);
    mix(fr->coll_coef, sizeof(double) * fr->ncolltab); mix(fr->line_rows, sizeof(double) * pl->nline * RHB200_PL_NFIELD);
    mix(&fr->stokes, sizeof(int)); mix(&c->s_interpolation_stokes, sizeof(int));
    if (prd) { mix(fr->line_prd, sizeof(int) * pl->nline); mix(&fr->PRD_NmaxIter, sizeof(int)); mix(&fr->PRDiterLimit, sizeof(double)); }
    if (stokes) {
      const int nz = fr->line_zoff[pl->nline];
      mix(fr->line_pol, sizeof(int) * pl->nline); mix(fr->line_zoff, sizeof(int) * (pl->nline + 1));
      if (nz > 0) { mix(fr->zq, sizeof(int) * nz); mix(fr->zshift, sizeof(double) * nz); mix(fr->zstrength, sizeof(double) * nz); }
    }
    for (int a = 0; a < natom_model; a++) { const double ab = rh_continuum_abundance(c, a); mix(&ab, sizeof ab); }
  }
  FrontState *S = (FrontState *) c->nlte_front;
  const bool reuse = S && S->key == key;
  mark(reuse ? "state reused (hash)" : "hash");
  if (!reuse) {
    delete S;
    c->nlte_front = nullptr;
    S = new FrontState();
    S->key = key;
    // the two engines: Nrays rays for initScatter / Iterate, one ray at `mu` for _solveray()'s pass
    int rc = S->E.build(c, pl);
    if (rc == RHB200_OK) rc = S->F.build(c, &p1);
    if (rc == RHB200_OK && stokes) rc = S->E.set_zeeman(pl, fr->line_pol, fr->line_zoff, fr->zq, fr->zshift, fr->zstrength);
    if (rc == RHB200_OK && stokes) rc = S->F.set_zeeman(&p1, fr->line_pol, fr->line_zoff, fr->zq, fr->zshift, fr->zstrength);
    if (rc == RHB200_OK && prd) rc = S->E.set_prd(pl, fr->line_prd, fr->PRD_NmaxIter, fr->PRDiterLimit);
    if (rc == RHB200_OK && prd) rc = S->F.set_prd(&p1, fr->line_prd, 0, 0.0);
    if (rc == RHB200_OK && S->E.nrank > 1) { rhb200_set_error("wavelength sharding is only available through rhb200_nlte_iterate"); rc = RHB200_EUNSUPPORTED; }
    if (rc != RHB200_OK) { delete S; return rc; }
    // ---- chunk size from the workspace budget
    size_t budget = (size_t) 48 << 30;
    if (const char *e = getenv("RHB200_NLTE_WS_GB")) { const double g = atof(e); if (g > 0.01) budget = (size_t) (g * (double) ((size_t) 1 << 30)); }
    const size_t per_col = sizeof(double) * (S->E.doubles_per_column(true) + S->F.doubles_per_column(false) +
        (size_t) N * ((size_t) nrow + RHB200_AT_NFIELD + 2*(size_t) nlev_model + 2*(size_t) (natom_model + 4) + 8 + 5 +
                      (size_t) std::max(1, c->wav.npl) * 4 + (size_t) std::max(1, pl->nline) * 4 +
                      (size_t) std::max(1, nelem) * RHB200_RE_MAXSTAGE + (size_t) std::max(1, nline_k) * LP_NFIELD +
                      (size_t) c->wav.nmsel * 4 + 2*(size_t) (c->wav.nmw > 0 ? Ns : 0)) + (size_t) Ns);
    int cc = (int) std::min<size_t>((size_t) ncol, std::max<size_t>(1, budget / per_col));
    if (const char *e = getenv("RHB200_NLTE_CHUNK_COLS")) { const int v = atoi(e); if (v > 0) cc = std::min(ncol, v); }
    { const size_t nchunk = ((size_t) ncol + cc - 1) / cc; cc = (int) (((size_t) ncol + nchunk - 1) / nchunk); }
    if ((size_t) cc * N > (size_t) 65535 * 128) cc = (int) ((size_t) 65535 * 128 / N);      // grid.y limit of the continuum kernel
    S->cc = cc;
    auto build_state = [&]() -> int {
      NlteEngine &E = S->E, &F = S->F;
      DevArena &ar = S->ar;
      RH_CHECK(E.alloc(cc, true, true));
      RH_CHECK(F.alloc(cc, true, false));
      // F shares every input with E; only the background of the final pass and the profiles are its own
      F.C.T = E.C.T; F.C.height = E.C.height; F.C.nstar = E.C.nstar; F.C.ntotal = E.C.ntotal; F.C.C = E.C.C;
      F.C.vbroad = E.C.vbroad; F.C.vel = E.C.vel; F.C.n = E.C.n; F.C.J = E.C.J;
      F.C.rho = E.C.rho;                               // the final pass keeps the profile ratio of the PRD lines (profile.c:83-90)
      const size_t cN = (size_t) cc * N;
      RH_CHECK(ar.alloc(&S->d_in, cN * nrow)); RH_CHECK(ar.alloc(&S->d_at, cN * RHB200_AT_NFIELD));
      RH_CHECK(ar.alloc(&S->d_pops, cN * nlev_model)); RH_CHECK(ar.alloc(&S->d_popsn, cN * nlev_model));
      RH_CHECK(ar.alloc(&S->d_chem, cN * (natom_model + 4))); RH_CHECK(ar.alloc(&S->d_ntot, cN * natom_model));
      RH_CHECK(ar.alloc(&S->d_tprep, cN * 8)); RH_CHECK(ar.alloc(&S->d_sc, cN * 5));
      RH_CHECK(ar.alloc(&S->d_pc, cN * std::max(1, c->wav.npl) * 4)); RH_CHECK(ar.alloc(&S->d_apc, cN * std::max(1, pl->nline) * 4));
      RH_CHECK(ar.alloc(&S->d_elem_n, cN * std::max(1, nelem) * RHB200_RE_MAXSTAGE));
      RH_CHECK(ar.alloc(&S->d_lineprep, cN * std::max(1, nline_k) * LP_NFIELD));
      if (mol_on) {
        RH_CHECK(ar.alloc(&S->d_md, cN * c->wav.nmsel)); RH_CHECK(ar.alloc(&S->d_mol, cN * c->wav.nmsel * 3));
        RH_CHECK(ar.alloc(&S->d_mchi, cN * Ns)); RH_CHECK(ar.alloc(&S->d_meta, cN * Ns));
      }
      RH_CHECK(ar.alloc(&S->d_spec, (size_t) cc * Ns)); RH_CHECK(ar.alloc(&S->d_quv, (size_t) cc * Ns * 3));
      RH_CHECK(ar.alloc(&S->d_adamp2, cN * std::max(1, pl->nline)));
      RH_CHECK(ar.alloc(&S->d_chi2, cN * Ns)); RH_CHECK(ar.alloc(&S->d_eta2, cN * Ns)); RH_CHECK(ar.alloc(&S->d_sca2, cN * Ns));
      std::vector<double> ab(natom_model);
      for (int a = 0; a < natom_model; a++) ab[a] = rh_continuum_abundance(c, a);
      RH_CHECK(ar.upload(&S->d_abund, ab.data(), ab.size()));
      // g[j] (OMEGA) / g[i]/g[j] (CE) travel in the spare field of the record: supplied by the host in field 7
      RH_CHECK(ar.upload(&S->d_coll, fr->coll, (size_t) std::max(1, fr->ncoll) * RHB200_CO_NFIELD));
      RH_CHECK(ar.upload(&S->d_cT, fr->coll_T, (size_t) std::max(1, fr->ncolltab)));
      RH_CHECK(ar.upload(&S->d_cC, fr->coll_coef, (size_t) std::max(1, fr->ncolltab)));
      RH_CHECK(ar.upload(&S->d_cM, fr->coll_M, (size_t) std::max(1, fr->ncolltab)));
      RH_CHECK(ar.upload(&S->d_plrows, fr->line_rows, (size_t) std::max(1, pl->nline) * RHB200_PL_NFIELD));
      return RHB200_OK;
    };
    rc = build_state();
    if (rc != RHB200_OK) { delete S; return rc; }
    c->nlte_front = S;
    mark("plans, engines, work arrays");
  }
  if (mol_on) {
    std::vector<int> chem(c->wav.nmsel);
    for (int m = 0; m < c->wav.nmsel; m++) chem[m] = (int) c->h_msel[(size_t) m * 16];
    RH_CHECK(rh_continuum_set_molsel(c, c->wav.nmsel, chem.data()));
  }
  NlteEngine &E = S->E, &F = S->F;
  const int cc = S->cc;
  const size_t cN = (size_t) cc * N;
  double *d_in = S->d_in, *d_at = S->d_at, *d_pops = S->d_pops, *d_popsn = S->d_popsn, *d_chem = S->d_chem, *d_ntot = S->d_ntot,
         *d_tprep = S->d_tprep, *d_sc = S->d_sc, *d_pc = S->d_pc, *d_apc = S->d_apc, *d_elem_n = S->d_elem_n, *d_lineprep = S->d_lineprep,
         *d_md = S->d_md, *d_mol = S->d_mol, *d_mchi = S->d_mchi, *d_meta = S->d_meta, *d_spec = S->d_spec, *d_abund = S->d_abund,
         *d_coll = S->d_coll, *d_cT = S->d_cT, *d_cC = S->d_cC, *d_cM = S->d_cM, *d_plrows = S->d_plrows, *d_adamp2 = S->d_adamp2,
         *d_chi2 = S->d_chi2, *d_eta2 = S->d_eta2, *d_sca2 = S->d_sca2;
  F.C.adamp = E.C.adamp;
  // collision.c:470-471
  const double C0 = ((2.1798741E-18/sqrt(RH_M_ELECTRON)) * RH_PI*(5.29177349E-11*5.29177349E-11)) * sqrt(8.0/(RH_PI*RH_KBOLTZMANN));
  const double mu_last = pl->muz[pl->Nrays - 1];
  cudaStream_t st = c->stream;
  g_front_debug.arr.clear();

  // background of one chunk for ray direction muz (up): continuum with the populations seen through atom->n, passive_bb,
  // Kurucz lines, molecular lines -> chi_c (scattering included: background.c:462), eta_c, sca_c
  auto background = [&](int n, double muz, const double *pops_n, double *chi_c, double *eta_c, double *sca_c,
                        double *chi_quv, double *eta_quv) -> int {
    RH_CHECK(rh_continuum_opac(c, n, N, d_at, d_chem, pops_n, d_pops, d_tprep, chi_c, eta_c, sca_c));
    if (mol_on) RH_CHECK(rh_molecular_chunk(c, n, N, muz, d_at, d_md, d_mol, d_mchi, d_meta));
    RH_CHECK(rh_passive_chunk(c, n, N, muz, d_at, pops_n, nlev_model, d_pc, chi_c, eta_c));
    RH_CHECK(rh_launch_proton(c, n, N, nlev_model, rh_continuum_proton_level(c), pops_n, d_at));
    RH_CHECK(rh_launch_prep(c, n, N, muz, 1, d_at, d_elem_n, d_lineprep));
    RH_CHECK(rh_launch_opacity_addI(c, n, N, 1, d_at, d_lineprep, chi_c, eta_c, chi_quv, eta_quv));
    if (mol_on) RH_CHECK(rh_launch_add_molecular(c, n, N, d_mchi, d_meta, chi_c, eta_c));
    return RHB200_OK;
  };
  auto keep = [&](int which, const double *d, size_t count) -> int {
    if ((int) g_front_debug.arr.size() <= which) g_front_debug.arr.resize(which + 1);
    g_front_debug.arr[which].resize(count);
    RH_CUDA(cudaStreamSynchronize(st));
    RH_CUDA(cudaMemcpy(g_front_debug.arr[which].data(), d, count * sizeof(double), cudaMemcpyDeviceToHost));
    return RHB200_OK;
  };
  const bool debug_keep = getenv("RHB200_NLTE_FRONT_DEBUG") != nullptr;     // test hook, read per call

  for (int c0 = 0; c0 < ncol; c0 += cc) {
    const int n = std::min(cc, ncol - c0);
    const size_t nN = (size_t) n * N;
    E.ncol = n; F.ncol = n;
    RH_CUDA(cudaMemcpyAsync(d_in, atmosphere + (size_t) c0 * nrow * N, nN * nrow * sizeof(double), cudaMemcpyHostToDevice, st));
    // ---- first Background(): rays of the angle quadrature, record of the last one (mu = Nrays-1, up)
    RH_CHECK(rh_launch_pyrh_rows(c, n, N, nrow, atm_scale, mu_last, 0.0, d_in, d_at, nullptr));
    RH_CHECK(rh_continuum_ltepops(c, n, N, d_at, nullptr, d_pops, nullptr));
    nlte_collision_kernel<<<RH_GRID(nN, 128), 0, st>>>(A, E.P, n, nlev_model, H_nlevel, C0, fr->ncoll, d_coll, d_cT, d_cC, d_cM,
                                                       d_at, d_pops, (double *) E.C.C);
    nlte_ntot_init_kernel<<<RH_GRID(nN * natom_model, 128), 0, st>>>(n, N, natom_model, d_abund, d_at, d_ntot);
    RH_CUDA(cudaGetLastError());
    RH_CHECK(rh_continuum_chemeq(c, n, N, d_at, d_pops, d_chem, mol_on ? d_md : nullptr, d_ntot, 0));
    nlte_gather_kernel<<<RH_GRID(nN, 128), 0, st>>>(A, E.P, E.C, n, nlev_model, natom_model, d_at, d_pops, d_ntot, E.C.n);
    RH_CUDA(cudaGetLastError());
    const double *pops_n = d_pops;
    if (H_active) {
      nlte_popsn_kernel<<<RH_GRID(nN * nlev_model, 128), 0, st>>>(n, N, nlev_model, H_nlevel, d_pops, nullptr, E.nlev,
                                                                  E.lev_off[H_engine_atom], d_popsn);
      RH_CUDA(cudaGetLastError());
      pops_n = d_popsn;
    }
    RH_CHECK(background(n, mu_last, pops_n, (double *) E.C.chi_c, (double *) E.C.eta_c, (double *) E.C.sca_c,
                        stokes ? E.d_chi_cQ : nullptr, stokes ? E.d_eta_cQ : nullptr));
    if (stokes) RH_CHECK(E.bproject(d_in, nrow));
    E.set_stokes(full_stokes);                       // FULL_STOKES: polarised profiles, rays and I_eff from the first pass on
    if (pol_free) E.set_stokes_profiles_only(true);  // POLARIZATION_FREE: Zeeman-broadened profiles, I alone (opacity.c:97, formal.c:94)
    RH_CHECK(rh_launch_scales_chi(c, n, N, Ns, iref, atm_scale, wght_per_H, c->total_abund, c->gravity > 0.0 ? c->gravity : 1.0, E.C.chi_c, d_at, d_sc,
                                  scales ? d_sc + 2*cN : nullptr));
    nlte_height_kernel<<<RH_GRID(nN, 128), 0, st>>>(n, N, d_at, (double *) E.C.height);
    // ---- getProfiles(): Damping() with the populations atom->n shows at this point (line->Qelast of PRD lines kept)
    RH_CHECK(rh_launch_line_damping(c, n, N, pl->nline, d_plrows, d_at, pops_n, nlev_model, d_apc, prd ? (double *) E.C.Qelast : nullptr));
    nlte_damping_gather_kernel<<<RH_GRID(nN * pl->nline, 128), 0, st>>>(E.P, E.C, n, d_apc);
    RH_CUDA(cudaGetLastError());
    RH_CUDA(cudaMemsetAsync(E.C.J, 0, nN * Ns * sizeof(double), st));                  // initSolution: J = 0
    F.C.adamp = E.C.adamp;
    if (debug_keep && c0 == 0) {
      RH_CHECK(keep(0, E.C.C, nN * E.ngam)); RH_CHECK(keep(1, E.C.nstar, nN * E.nlev)); RH_CHECK(keep(2, E.C.ntotal, nN * E.Na));
      RH_CHECK(keep(3, E.C.adamp, nN * E.nline)); RH_CHECK(keep(4, E.C.vbroad, nN * E.Na));
      RH_CHECK(keep(5, E.C.chi_c, nN * Ns)); RH_CHECK(keep(6, E.C.eta_c, nN * Ns)); RH_CHECK(keep(7, E.C.sca_c, nN * Ns));
      RH_CHECK(keep(8, E.C.height, nN));
    }
    mark("background + damping");
    // ---- initScatter, Iterate, the scattering passes after it
    RH_CHECK(E.prepare(nullptr, nullptr, true));
    mark("profiles + setup");
    std::vector<int> niter(n, 0), pass_a(n, 0), pass_b(n, 0);
    RH_CHECK(E.scatter(fr->NmaxIter ? fr->NmaxScatter : 0, 1, fr->iterLimit, pass_a.data(), nullptr, nullptr));
    mark("initScatter");
    RH_CHECK(E.iterate(fr->NmaxIter, fr->iterLimit, niter.data(), nullptr, 0, nullptr, nullptr));
    mark("Iterate");
    if (pol_free) E.set_stokes(true);                // adjustStokesMode(): FULL_STOKES from here on, the profiles are kept (zeeman.c:319-321)
    if (stokes && !full_stokes && !pol_free) {       // adjustStokesMode(), pyrh_compute1dray.c:332: Profile() of the polarizable lines again
      if (H_active) {                                // their Damping() now sees hydrogen's NLTE populations
        nlte_popsn_kernel<<<RH_GRID(nN * nlev_model, 128), 0, st>>>(n, N, nlev_model, H_nlevel, d_pops, E.C.n, E.nlev,
                                                                    E.lev_off[H_engine_atom], d_popsn);
        RH_CUDA(cudaGetLastError());
        RH_CHECK(rh_launch_line_damping(c, n, N, pl->nline, d_plrows, d_at, pops_n, nlev_model, d_apc));
        Cols Ctmp = E.C;
        Ctmp.adamp = d_adamp2;
        nlte_damping_gather_kernel<<<RH_GRID(nN * pl->nline, 128), 0, st>>>(E.P, Ctmp, n, d_apc);
        nlte_adamp_select_kernel<<<RH_GRID(nN * pl->nline, 128), 0, st>>>(E.P, n, d_adamp2, (double *) E.C.adamp);
        RH_CUDA(cudaGetLastError());
      }
      E.set_stokes(true);
      RH_CHECK(E.prepare(nullptr, nullptr, false));
      mark("adjustStokesMode");
    }
    RH_CHECK(E.scatter(fr->NmaxScatter, 2, fr->iterLimit, pass_b.data(), nullptr, nullptr));
    E.set_stokes(full_stokes);
    if (pol_free) E.set_stokes_profiles_only(true);
    mark("passes after Iterate");
    if (niter_out) memcpy(niter_out + c0, niter.data(), n * sizeof(int));
    if (passes_out) for (int q = 0; q < n; q++) { passes_out[2*(size_t) (c0 + q)] = pass_a[q]; passes_out[2*(size_t) (c0 + q) + 1] = pass_b[q]; }
    // ---- _solveray(): one ray at mu; Background() and getProfiles() again
    RH_CHECK(rh_launch_pyrh_rows(c, n, N, nrow, atm_scale, mu, 0.0, d_in, d_at, nullptr));
    // pyrh_rows rewrote the scale row: put the heights back
    RH_CUDA(cudaMemcpy2DAsync(d_at + (size_t) RHB200_AT_HEIGHT * N, (size_t) RHB200_AT_NFIELD * N * sizeof(double),
                              E.C.height, (size_t) N * sizeof(double), (size_t) N * sizeof(double), n, cudaMemcpyDeviceToDevice, st));
    // SetLTEQuantities + ChemicalEquilibrium again, from the ntotal the first pass left (chemequil.c:342)
    RH_CHECK(rh_continuum_ltepops(c, n, N, d_at, nullptr, d_pops, d_ntot));
    RH_CHECK(rh_continuum_chemeq(c, n, N, d_at, d_pops, d_chem, mol_on ? d_md : nullptr, d_ntot, 1));
    nlte_regather_kernel<<<RH_GRID(nN, 128), 0, st>>>(A, E.P, E.C, n, nlev_model, natom_model, d_pops, d_chem, E.C.n);
    RH_CUDA(cudaGetLastError());
    if (H_active) {
      nlte_popsn_kernel<<<RH_GRID(nN * nlev_model, 128), 0, st>>>(n, N, nlev_model, H_nlevel, d_pops, E.C.n, E.nlev,
                                                                  E.lev_off[H_engine_atom], d_popsn);
      RH_CUDA(cudaGetLastError());
    }
    RH_CHECK(background(n, mu, pops_n, d_chi2, d_eta2, d_sca2, stokes ? F.d_chi_cQ : nullptr, stokes ? F.d_eta_cQ : nullptr));
    if (stokes) { RH_CHECK(F.bproject(d_in, nrow)); F.set_stokes(true); }
    F.C.chi_c = d_chi2; F.C.eta_c = d_eta2; F.C.sca_c = d_sca2;
    {                                                // Damping() again: hydrogen's populations changed (NLTE solution if
      RH_CHECK(rh_launch_line_damping(c, n, N, pl->nline, d_plrows, d_at, pops_n, nlev_model, d_apc));   // ACTIVE, else the
      F.C.adamp = d_adamp2;                                                                              // re-derived LTE ones)
      Cols Ctmp = F.C;
      nlte_damping_gather_kernel<<<RH_GRID(nN * pl->nline, 128), 0, st>>>(F.P, Ctmp, n, d_apc);    // vbroad: unchanged values
      RH_CUDA(cudaGetLastError());
    }
    mark("second background + damping");
    RH_CHECK(F.prepare(nullptr, nullptr, false));
    RH_CHECK(F.scatter(1, 0, 0.0, nullptr, nullptr, d_spec, quv ? S->d_quv : nullptr));
    mark("final pass");
    if (debug_keep && c0 == 0) {
      RH_CHECK(keep(9, E.C.J, nN * Ns));
      RH_CHECK(keep(10, d_chi2, nN * Ns)); RH_CHECK(keep(11, d_eta2, nN * Ns)); RH_CHECK(keep(12, d_sca2, nN * Ns));
      RH_CHECK(keep(13, F.C.phi, nN * F.nphirow)); RH_CHECK(keep(14, F.C.wphi, nN * F.nline)); RH_CHECK(keep(15, F.C.adamp, nN * F.nline));
    }
    RH_CUDA(cudaStreamSynchronize(st));
    if (spectrum) RH_CUDA(cudaMemcpy(spectrum + (size_t) c0 * Ns, d_spec, (size_t) n * Ns * sizeof(double), cudaMemcpyDeviceToHost));
    if (quv) RH_CUDA(cudaMemcpy(quv + (size_t) c0 * 3 * Ns, S->d_quv, (size_t) n * 3 * Ns * sizeof(double), cudaMemcpyDeviceToHost));
    if (out_n) RH_CUDA(cudaMemcpy(out_n + (size_t) c0 * E.nlev * N, E.C.n, nN * E.nlev * sizeof(double), cudaMemcpyDeviceToHost));
    if (out_nstar) RH_CUDA(cudaMemcpy(out_nstar + (size_t) c0 * E.nlev * N, E.C.nstar, nN * E.nlev * sizeof(double), cudaMemcpyDeviceToHost));
    if (scales) RH_CUDA(cudaMemcpy(scales + (size_t) c0 * 3 * N, d_sc + 2*cN, nN * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return RHB200_OK;
}
