/* oracle/probe.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiled *into* oracle/_ref/liboracle_{scalar,simd}.so next to the unmodified
 * reference objects.  It contains no reference code: it only *observes* the
 * reference through `ld --wrap=<symbol>` (see oracle/build_ref.sh), copying the
 * arguments/results of hot-path functions and a snapshot of the reference's
 * global state (atmos, geometry, spectrum: pyrh_compute1dray.c:47-53) into an
 * in-memory record log that tests/golden generators read through ctypes.
 *
 * Wrapped reference functions (file:line of the real definition):
 *   rlk_opacity              rh/kurucz.c:511
 *   writeBackground          rh/readj.c:284
 *   Piece_Stokes_Bezier3_1D  rh/rhf1d/bezier_1D.c:52
 *   Piecewise_Bezier3_1D     rh/rhf1d/bezier_1D.c:306
 *   Feautrier                rh/rhf1d/feautrier.c:56
 *   Formal                   rh/rhf1d/formal.c:44
 *   Opacity                  rh/opacity.c:64
 *   addtoGamma / addtoRates  rh/fillgamma.c:82 / :375
 *   statEquil                rh/statequil.c:40
 *   Accelerate               rh/accelerate.c:68
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "rh.h"
#include "atom.h"
#include "atmos.h"
#include "rhf1d/geometry.h"
#include "spectrum.h"
#include "background.h"
#include "inputs.h"
#include "constant.h"
#include "accelerate.h"

extern Atmosphere atmos;
extern Geometry geometry;
extern Spectrum spectrum;
extern InputData input;

/* ------------------------------------------------------------------ log */

typedef struct {
  char    tag[24];
  int     meta[8];
  long    n;
  double *data;
} ProbeRec;

static ProbeRec *recs = NULL;
static long nrec = 0, cap = 0;
static unsigned probe_mask = 0;   /* bit field, see PROBE_* below */
static int snapshot_done = 0;
static int mol_snapshot_done = 0;
static int cont_snapshot_done = 0;
#define PBB_MAXSEEN 256
static AtomicLine *pbb_seen[PBB_MAXSEEN];
static int pbb_nseen = 0;

enum { PROBE_RLK = 1, PROBE_BG = 2, PROBE_DELO = 4, PROBE_SNAP = 8,
       PROBE_BEZ = 16, PROBE_FEAU = 32, PROBE_NLTE = 64, PROBE_FORMAL = 128 };

void probe_enable(unsigned mask) { probe_mask = mask; }

void probe_reset(void)
{
  for (long i = 0; i < nrec; i++) free(recs[i].data);
  nrec = 0;
  snapshot_done = 0;
  mol_snapshot_done = 0;
  pbb_nseen = 0;
  cont_snapshot_done = 0;
  /* rhf1d() only ever SETS atmos.Nloggf / Nlam (pyrh_compute1dray.c:199-211): after a call with log gf overrides a
     call without them reads the previous caller's (freed) arrays in readKuruczLines().  The driver calls this before
     every rhf1d(), so each call sees only its own overrides -- what pyrh users get from a fresh process. */
  atmos.Nloggf = 0;
  atmos.Nlam = 0;
}
long      probe_count(void)       { return nrec; }
ProbeRec *probe_get(long i)       { return (i >= 0 && i < nrec) ? &recs[i] : NULL; }

static double *rec_new(const char *tag, long n, int m0, int m1, int m2, int m3,
                       int m4, int m5)
{
  if (nrec == cap) {
    cap = cap ? 2*cap : 1024;
    recs = (ProbeRec *) realloc(recs, cap * sizeof(ProbeRec));
  }
  ProbeRec *r = &recs[nrec++];
  memset(r, 0, sizeof(*r));
  strncpy(r->tag, tag, sizeof(r->tag)-1);
  r->meta[0] = m0; r->meta[1] = m1; r->meta[2] = m2; r->meta[3] = m3;
  r->meta[4] = m4; r->meta[5] = m5;
  r->n = n;
  r->data = (double *) malloc((n > 0 ? n : 1) * sizeof(double));
  return r->data;
}

static void rec_copy(const char *tag, const double *src, long n,
                     int m0, int m1, int m2, int m3)
{
  double *d = rec_new(tag, n, m0, m1, m2, m3, 0, 0);
  if (src) memcpy(d, src, n*sizeof(double)); else memset(d, 0, n*sizeof(double));
}

/* ------------------------------------------------------------ snapshot */

static void snapshot(void)
{
  int N = atmos.Nspace, i, n, k;
  if (snapshot_done) return;
  snapshot_done = 1;

  rec_copy("T", atmos.T, N, 0,0,0,0);
  rec_copy("ne", atmos.ne, N, 0,0,0,0);
  rec_copy("vturb", atmos.vturb, N, 0,0,0,0);
  rec_copy("vel", geometry.vel, N, 0,0,0,0);
  rec_copy("nHtot", atmos.nHtot, N, 0,0,0,0);
  rec_copy("height", geometry.height, N, 0,0,0,0);
  rec_copy("tau_ref", geometry.tau_ref, N, 0,0,0,0);
  rec_copy("cmass", geometry.cmass, N, 0,0,0,0);
  {
    double *d = rec_new("abund_sums", 4, 0,0,0,0,0,0);     /* abundance.c:219-221 */
    d[0] = atmos.wght_per_H; d[1] = atmos.totalAbund; d[2] = atmos.avgMolWght; d[3] = atmos.gravity;
  }
  if (atmos.Stokes && atmos.B) {
    rec_copy("B", atmos.B, N, 0,0,0,0);
    rec_copy("gamma_B", atmos.gamma_B, N, 0,0,0,0);
    rec_copy("chi_B", atmos.chi_B, N, 0,0,0,0);
    for (i = 0; i < atmos.Nrays; i++) {
      rec_copy("cos_gamma", atmos.cos_gamma[i], N, i,0,0,0);
      rec_copy("cos_2chi",  atmos.cos_2chi[i],  N, i,0,0,0);
      rec_copy("sin_2chi",  atmos.sin_2chi[i],  N, i,0,0,0);
    }
  }
  rec_copy("np", atmos.H->n[atmos.H->Nlevel-1], N, 0,0,0,0);
  rec_copy("muz", geometry.muz, geometry.Nrays, 0,0,0,0);
  rec_copy("wmu", geometry.wmu, geometry.Nrays, 0,0,0,0);
  rec_copy("lambda", spectrum.lambda, spectrum.Nspect, 0,0,0,0);
  {
    double *d = rec_new("flags", 16, 0,0,0,0,0,0);
    d[0] = atmos.moving; d[1] = atmos.Stokes; d[2] = input.magneto_optical;
    d[3] = input.rlkscatter; d[4] = geometry.vboundary[TOP];
    d[5] = geometry.vboundary[BOTTOM]; d[6] = atmos.vmicro_char;
    d[7] = atmos.lambda_ref; d[8] = input.StokesMode; d[9] = input.solve_NLTE;
    d[10] = atmos.Nrays; d[11] = atmos.Nrlk; d[12] = input.S_interpolation;
    d[13] = input.S_interpolation_stokes; d[14] = atmos.H_LTE; d[15] = input.LS_Lande;
  }
  {
    double *d = rec_new("backgrflags", 2*spectrum.Nspect, 0,0,0,0,0,0);
    for (n = 0; n < spectrum.Nspect; n++) {
      d[2*n]   = atmos.backgrflags[n].hasline;
      d[2*n+1] = atmos.backgrflags[n].ispolarized;
    }
  }
  if (atmos.Npf > 0 && atmos.Tpf) rec_copy("Tpf", atmos.Tpf, atmos.Npf, 0,0,0,0);

  /* Kurucz lines (sorted, as used by rlk_opacity) + the element tables they need */
  for (n = 0; n < atmos.Nrlk; n++) {
    RLK_Line *r = &atmos.rlk_lines[n];
    int Nc = r->zm ? r->zm->Ncomponent : 0;
    double *d = rec_new("rlk_line", 32 + 3*Nc, n, Nc, 0,0,0,0);
    d[0] = r->lambda0; d[1] = r->gi; d[2] = r->gj; d[3] = r->Ei; d[4] = r->Ej;
    d[5] = r->Bji; d[6] = r->Aji; d[7] = r->Bij; d[8] = r->Si; d[9] = r->Sj;
    d[10] = r->Grad; d[11] = r->GStark; d[12] = r->GvdWaals;
    d[13] = r->hyperfine_frac; d[14] = r->isotope_frac; d[15] = r->gL_i;
    d[16] = r->gL_j; d[17] = r->cross; d[18] = r->alpha; d[19] = r->polarizable;
    d[20] = r->vdwaals; d[21] = r->pt_index; d[22] = r->stage; d[23] = r->Li;
    d[24] = r->Lj; d[25] = Nc; d[26] = r->li; d[27] = r->lj;
    d[28] = r->get_loggf_rf; d[29] = r->loggf_rf_ind; d[30] = 0; d[31] = 0;
    for (i = 0; i < Nc; i++) {
      d[32+3*i]   = r->zm->q[i];
      d[32+3*i+1] = r->zm->shift[i];
      d[32+3*i+2] = r->zm->strength[i];
    }
    Element *e = &atmos.elements[r->pt_index - 1];
    {
      double *h = rec_new("elem", 8 + e->Nstage, r->pt_index, e->Nstage, 0,0,0,0);
      h[0] = e->weight; h[1] = e->abund; h[2] = e->abundance_set;
      h[3] = (e->model != NULL); h[4] = e->Nstage; h[5] = h[6] = h[7] = 0;
      for (i = 0; i < e->Nstage; i++)
        h[8+i] = (i < e->Nstage-1 && e->ionpot) ? e->ionpot[i] : 0.0;
    }
    if (e->pf)
      for (i = 0; i < e->Nstage; i++)
        rec_copy("elem_pf", e->pf[i], atmos.Npf, r->pt_index, i, 0,0);
    if (e->n)
      for (i = 0; i < e->Nstage; i++)
        rec_copy("elem_n", e->n[i], N, r->pt_index, i, 0,0);
  }
  (void) k;
}

void probe_snapshot(void) { snapshot_done = 0; snapshot(); }

/* ------------------------------------------------------------ wrappers */

flags __real_rlk_opacity(double lambda, int nspect, int mu, bool_t to_obs,
                         double *chi, double *eta, double *scatt, double *chip);
flags __wrap_rlk_opacity(double lambda, int nspect, int mu, bool_t to_obs,
                         double *chi, double *eta, double *scatt, double *chip)
{
  flags f = __real_rlk_opacity(lambda, nspect, mu, to_obs, chi, eta, scatt, chip);
  if (probe_mask & PROBE_RLK) {
    int N = atmos.Nspace, ns = atmos.Stokes ? 4 : 1;
    double *d = rec_new("rlk", 2*ns*N, nspect, mu, to_obs, f.hasline,
                        f.ispolarized, ns);
    if (f.hasline) {
      memcpy(d, chi, ns*N*sizeof(double));
      memcpy(d + ns*N, eta, ns*N*sizeof(double));
    } else
      memset(d, 0, 2*ns*N*sizeof(double));
  }
  return f;
}

int __real_writeBackground(int nspect, int mu, bool_t to_obs, double *chi_c,
                           double *eta_c, double *sca_c, double *chip_c);
int __wrap_writeBackground(int nspect, int mu, bool_t to_obs, double *chi_c,
                           double *eta_c, double *sca_c, double *chip_c)
{
  if (probe_mask & PROBE_BG) {
    int N = atmos.Nspace;
    int ns = atmos.backgrflags[nspect].ispolarized ? 4 : 1;
    double *d = rec_new("bg", (2*ns+1)*N, nspect, mu, to_obs, ns, 0, 0);
    memcpy(d, chi_c, ns*N*sizeof(double));
    memcpy(d + ns*N, eta_c, ns*N*sizeof(double));
    memcpy(d + 2*ns*N, sca_c, N*sizeof(double));
  }
  return __real_writeBackground(nspect, mu, to_obs, chi_c, eta_c, sca_c, chip_c);
}

void __real_Piece_Stokes_Bezier3_1D(int nspect, int mu, bool_t to_obs,
                                    double *chi, double **S, double **I, double *Psi);
void __wrap_Piece_Stokes_Bezier3_1D(int nspect, int mu, bool_t to_obs,
                                    double *chi, double **S, double **I, double *Psi)
{
  __real_Piece_Stokes_Bezier3_1D(nspect, mu, to_obs, chi, S, I, Psi);
  if (probe_mask & PROBE_DELO) {
    int N = atmos.Nspace, n;
    ActiveSet *as = &spectrum.as[nspect];
    /* layout: chi[N], S[4][N], I[4][N], Psi[N], chiQUV_total[3][N] */
    double *d = rec_new("delo", 13*N, nspect, mu, to_obs, Psi != NULL, 0, 0);
    memcpy(d, chi, N*sizeof(double));
    for (n = 0; n < 4; n++) memcpy(d + (1+n)*N, S[n], N*sizeof(double));
    for (n = 0; n < 4; n++) memcpy(d + (5+n)*N, I[n], N*sizeof(double));
    if (Psi) memcpy(d + 9*N, Psi, N*sizeof(double));
    else memset(d + 9*N, 0, N*sizeof(double));
    for (n = 0; n < 3*N; n++) {
      double v = 0.0;
      if (containsPolarized(as)) v += as->chi[N + n];
      if (atmos.backgrflags[nspect].ispolarized) v += as->chi_c[N + n];
      d[10*N + n] = v;
    }
  }
}

void __real_Piecewise_Bezier3_1D(int nspect, int mu, bool_t to_obs, double *chi,
                                 double *S, double *I, double *Psi, double **dI);
void __wrap_Piecewise_Bezier3_1D(int nspect, int mu, bool_t to_obs, double *chi,
                                 double *S, double *I, double *Psi, double **dI)
{
  /* log gf response function (bezier_1D.c:416-428, 477-490, 509-516): the up-ray reads the not yet
     overwritten I[] of the preceding down-ray, so that array is part of the input */
  int rf = (probe_mask & PROBE_BEZ) && input.get_atomic_rfs && to_obs && dI != NULL;
  double *rfrec = NULL;
  int np_ = input.n_atomic_pars;
  if (rf) {
    int N = atmos.Nspace, k, p;
    rfrec = rec_new("bezrf", (long) N + 3L*N*np_, nspect, mu, np_, 0, 0, 0);
    memcpy(rfrec, I, N*sizeof(double));                                   /* I before the call */
    for (k = 0; k < N; k++)
      for (p = 0; p < np_; p++) {
        rfrec[N + (long) k*np_ + p] = spectrum.dchi_c_lam[nspect][k][p];
        rfrec[N + (long) N*np_ + (long) k*np_ + p] = spectrum.deta_c_lam[nspect][k][p];
      }
  }
  __real_Piecewise_Bezier3_1D(nspect, mu, to_obs, chi, S, I, Psi, dI);
  if (rf) {
    int N = atmos.Nspace, k, p;
    for (k = 0; k < N; k++)
      for (p = 0; p < np_; p++) rfrec[N + 2L*N*np_ + (long) k*np_ + p] = dI[k][p];
  }
  if (probe_mask & PROBE_BEZ) {
    int N = atmos.Nspace;
    double *d = rec_new("bez", 4*N, nspect, mu, to_obs, Psi != NULL, 0, 0);
    memcpy(d, chi, N*sizeof(double));
    memcpy(d + N, S, N*sizeof(double));
    memcpy(d + 2*N, I, N*sizeof(double));
    if (Psi) memcpy(d + 3*N, Psi, N*sizeof(double));
    else memset(d + 3*N, 0, N*sizeof(double));
  }
}

/* Piecewise_Linear_1D / Piecewise_1D (rh/rhf1d/piecewise_1D.c:44,134): S_INTERPOLATION = S_LINEAR | S_PARABOLIC */
static void rec_scalar_ray(const char *tag, int nspect, int mu, int to_obs, double *chi, double *S,
                           double *I, double *Psi)
{
  int N = atmos.Nspace;
  double *d = rec_new(tag, 4*N, nspect, mu, to_obs, Psi != NULL, 0, 0);
  memcpy(d, chi, N*sizeof(double));
  memcpy(d + N, S, N*sizeof(double));
  memcpy(d + 2*N, I, N*sizeof(double));
  if (Psi) memcpy(d + 3*N, Psi, N*sizeof(double));
  else memset(d + 3*N, 0, N*sizeof(double));
}
void __real_Piecewise_1D(int nspect, int mu, bool_t to_obs, double *chi, double *S, double *I, double *Psi);
void __wrap_Piecewise_1D(int nspect, int mu, bool_t to_obs, double *chi, double *S, double *I, double *Psi)
{
  __real_Piecewise_1D(nspect, mu, to_obs, chi, S, I, Psi);
  if (probe_mask & PROBE_BEZ) rec_scalar_ray("par", nspect, mu, to_obs, chi, S, I, Psi);
}
void __real_Piecewise_Linear_1D(int nspect, int mu, bool_t to_obs, double *chi, double *S, double *I, double *Psi);
void __wrap_Piecewise_Linear_1D(int nspect, int mu, bool_t to_obs, double *chi, double *S, double *I, double *Psi)
{
  __real_Piecewise_Linear_1D(nspect, mu, to_obs, chi, S, I, Psi);
  if (probe_mask & PROBE_BEZ) rec_scalar_ray("lin", nspect, mu, to_obs, chi, S, I, Psi);
}

/* Piece_Stokes_1D (rh/rhf1d/piecestokes_1D.c:49): S_INTERPOLATION_STOKES = DELO_PARABOLIC; same record
   layout as "delo" */
void __real_Piece_Stokes_1D(int nspect, int mu, bool_t to_obs, double *chi, double **S, double **I, double *Psi);
void __wrap_Piece_Stokes_1D(int nspect, int mu, bool_t to_obs, double *chi, double **S, double **I, double *Psi)
{
  __real_Piece_Stokes_1D(nspect, mu, to_obs, chi, S, I, Psi);
  if (probe_mask & PROBE_DELO) {
    int N = atmos.Nspace, n;
    ActiveSet *as = &spectrum.as[nspect];
    double *d = rec_new("pst", 13*N, nspect, mu, to_obs, Psi != NULL, 0, 0);
    memcpy(d, chi, N*sizeof(double));
    for (n = 0; n < 4; n++) memcpy(d + (1+n)*N, S[n], N*sizeof(double));
    for (n = 0; n < 4; n++) memcpy(d + (5+n)*N, I[n], N*sizeof(double));
    if (Psi) memcpy(d + 9*N, Psi, N*sizeof(double));
    else memset(d + 9*N, 0, N*sizeof(double));
    for (n = 0; n < 3*N; n++) {
      double v = 0.0;
      if (containsPolarized(as)) v += as->chi[N + n];
      if (atmos.backgrflags[nspect].ispolarized) v += as->chi_c[N + n];
      d[10*N + n] = v;
    }
  }
}

double __real_Feautrier(int nspect, int mu, double *chi, double *S,
                        enum FeautrierOrder order, double *P, double *Psi);
double __wrap_Feautrier(int nspect, int mu, double *chi, double *S,
                        enum FeautrierOrder order, double *P, double *Psi)
{
  double Iem = __real_Feautrier(nspect, mu, chi, S, order, P, Psi);
  if (probe_mask & PROBE_FEAU) {
    int N = atmos.Nspace;
    double *d = rec_new("feau", 4*N + 1, nspect, mu, order, Psi != NULL, 0, 0);
    memcpy(d, chi, N*sizeof(double));
    memcpy(d + N, S, N*sizeof(double));
    memcpy(d + 2*N, P, N*sizeof(double));
    if (Psi) memcpy(d + 3*N, Psi, N*sizeof(double));
    else memset(d + 3*N, 0, N*sizeof(double));
    d[4*N] = Iem;
  }
  return Iem;
}

double __real_Formal(int nspect, bool_t eval_operator, bool_t redistribute);
double __wrap_Formal(int nspect, bool_t eval_operator, bool_t redistribute)
{
  if (probe_mask & PROBE_SNAP) snapshot();
  double dJ = __real_Formal(nspect, eval_operator, redistribute);
  if (probe_mask & PROBE_FORMAL) {
    int N = atmos.Nspace;
    double *d = rec_new("formal", N + 1, nspect, eval_operator, redistribute, 0,0,0);
    memcpy(d, spectrum.J[nspect], N*sizeof(double));
    d[N] = dJ;
  }
  return dJ;
}

/* --- NLTE observers: record inputs/outputs of the rate machinery -------- */

void __real_Opacity(int nspect, int mu, bool_t to_obs, bool_t initialize);
void __wrap_Opacity(int nspect, int mu, bool_t to_obs, bool_t initialize)
{
  __real_Opacity(nspect, mu, to_obs, initialize);
  if (probe_mask & PROBE_NLTE) {
    int N = atmos.Nspace;
    ActiveSet *as = &spectrum.as[nspect];
    double *d = rec_new("opac", 2*N, nspect, mu, to_obs, initialize, 0, 0);
    memcpy(d, as->chi, N*sizeof(double));
    memcpy(d + N, as->eta, N*sizeof(double));
  }
}

void __real_addtoGamma(int nspect, double wmu, double *P, double *Psi);
void __wrap_addtoGamma(int nspect, double wmu, double *P, double *Psi)
{
  if (probe_mask & PROBE_NLTE) {
    int N = atmos.Nspace;
    double *d = rec_new("gam_in", 2*N + 1, nspect, 0,0,0,0,0);
    memcpy(d, P, N*sizeof(double));
    memcpy(d + N, Psi, N*sizeof(double));
    d[2*N] = wmu;
  }
  __real_addtoGamma(nspect, wmu, P, Psi);
}

void __real_addtoRates(int nspect, int mu, bool_t to_obs, double wmu, double *I,
                       bool_t redistribute);
void __wrap_addtoRates(int nspect, int mu, bool_t to_obs, double wmu, double *I,
                       bool_t redistribute)
{
  __real_addtoRates(nspect, mu, to_obs, wmu, I, redistribute);
}

void __real_statEquil(Atom *atom, int isum);
void __wrap_statEquil(Atom *atom, int isum)
{
  int N = atmos.Nspace, Nl = atom->Nlevel;
  if (probe_mask & PROBE_NLTE) {
    double *d = rec_new("se_in", (long)(Nl*Nl + 2*Nl)*N + N, Nl, isum, 0,0,0,0);
    memcpy(d, atom->Gamma[0], (long)Nl*Nl*N*sizeof(double));
    memcpy(d + (long)Nl*Nl*N, atom->n[0], (long)Nl*N*sizeof(double));
    memcpy(d + (long)(Nl*Nl+Nl)*N, atom->nstar[0], (long)Nl*N*sizeof(double));
    memcpy(d + (long)(Nl*Nl+2*Nl)*N, atom->ntotal, N*sizeof(double));
  }
  __real_statEquil(atom, isum);
  if (probe_mask & PROBE_NLTE) {
    double *d = rec_new("se_out", (long)Nl*N, Nl, isum, 0,0,0,0);
    memcpy(d, atom->n[0], (long)Nl*N*sizeof(double));
  }
}

bool_t __real_Accelerate(struct Ng *Ngs, double *solution);
bool_t __wrap_Accelerate(struct Ng *Ngs, double *solution)
{
  long N = Ngs->N;
  if (probe_mask & PROBE_NLTE) {
    double *d = rec_new("ng_in", N, Ngs->Norder, Ngs->Ndelay, Ngs->Nperiod,
                        Ngs->count, 0, 0);
    memcpy(d, solution, N*sizeof(double));
  }
  bool_t acc = __real_Accelerate(Ngs, solution);
  if (probe_mask & PROBE_NLTE) {
    double *d = rec_new("ng_out", N, Ngs->Norder, acc, 0,0,0,0);
    memcpy(d, solution, N*sizeof(double));
  }
  return acc;
}

/* ===================================================================== NLTE
   Full dump of the NLTE problem the reference is about to iterate (state after
   initSolution + initScatter), taken at the entry of Iterate() (rh/iterate.c:48),
   per-iteration Gamma / rates / populations from updatePopulations()
   (rh/statequil.c:177), and every SolveLinearEq() call (rh/ludcmp.c:36).       */

static void dump_nlte_problem(int NmaxIter, double iterLimit)
{
  int N = atmos.Nspace, a, kr, n, la, k, Ns = spectrum.Nspect;
  {
    double *d = rec_new("nl_hdr", 16, 0,0,0,0,0,0);
    d[0] = Ns; d[1] = atmos.Nrays; d[2] = atmos.Nactiveatom; d[3] = N; d[4] = atmos.moving;
    d[5] = input.Ngorder; d[6] = input.Ngdelay; d[7] = input.Ngperiod; d[8] = input.isum;
    d[9] = NmaxIter; d[10] = iterLimit; d[11] = input.NmaxScatter; d[12] = input.StokesMode;
    d[13] = geometry.vboundary[TOP]; d[14] = geometry.vboundary[BOTTOM]; d[15] = input.S_interpolation;
  }
  rec_copy("nl_lambda", spectrum.lambda, Ns, 0,0,0,0);
  rec_copy("nl_muz", geometry.muz, geometry.Nrays, 0,0,0,0);
  rec_copy("nl_wmu", geometry.wmu, geometry.Nrays, 0,0,0,0);
  rec_copy("nl_T", atmos.T, N, 0,0,0,0);
  rec_copy("nl_height", geometry.height, N, 0,0,0,0);
  rec_copy("nl_vel", geometry.vel, N, 0,0,0,0);
  {
    double *J = rec_new("nl_J", (long) Ns*N, 0,0,0,0,0,0);
    double *b = rec_new("nl_bg", (long) 3*Ns*N, 0,0,0,0,0,0);
    double *f = rec_new("nl_bgflags", 2*Ns, 0,0,0,0,0,0);
    for (n = 0; n < Ns; n++) {
      memcpy(J + (long) n*N, spectrum.J[n], N*sizeof(double));
      memcpy(b + (long) n*N, spectrum.chi_c_lam[n], N*sizeof(double));
      memcpy(b + (long) (Ns + n)*N, spectrum.eta_c_lam[n], N*sizeof(double));
      memcpy(b + (long) (2*Ns + n)*N, spectrum.sca_c_lam[n], N*sizeof(double));
      f[2*n] = atmos.backgrflags[n].hasline; f[2*n+1] = atmos.backgrflags[n].ispolarized;
    }
  }
  for (a = 0; a < atmos.Nactiveatom; a++) {
    Atom *atom = atmos.activeatoms[a];
    int Nl = atom->Nlevel;
    double *d = rec_new("nl_atom", 2*Nl, a, Nl, atom->Nline, atom->Ncont, 0, 0);
    for (n = 0; n < Nl; n++) { d[n] = atom->g[n]; d[Nl+n] = atom->E[n]; }
    rec_copy("nl_n", atom->n[0], (long) Nl*N, a,0,0,0);
    rec_copy("nl_nstar", atom->nstar[0], (long) Nl*N, a,0,0,0);
    rec_copy("nl_ntotal", atom->ntotal, N, a,0,0,0);
    rec_copy("nl_C", atom->C[0], (long) Nl*Nl*N, a,0,0,0);
    rec_copy("nl_vbroad", atom->vbroad, N, a,0,0,0);
    for (kr = 0; kr < atom->Nline; kr++) {
      AtomicLine *L = &atom->line[kr];
      int Nla = L->Nlambda;
      double *h = rec_new("nl_line", 8 + 2*Nla + N, a, kr, L->i, L->j, Nla, L->Nblue);
      h[0] = L->lambda0; h[1] = L->Aji; h[2] = L->Bji; h[3] = L->Bij; h[4] = L->isotope_frac;
      h[5] = L->symmetric; h[6] = L->PRD; h[7] = L->polarizable;
      {   /* inputs of Profile() (profile.c:67): damping parameter, components */
        double *ad = rec_new("nl_adamp", N, a, kr, L->Voigt, L->Ncomponent, 0, 0);
        memset(ad, 0, N*sizeof(double));
        if (L->Voigt) Damping(L, ad);
        double *cc = rec_new("nl_comp", 2*L->Ncomponent, a, kr, L->Ncomponent, 0,0,0);
        for (n = 0; n < L->Ncomponent; n++) { cc[n] = L->c_shift[n]; cc[L->Ncomponent+n] = L->c_fraction[n]; }
      }
      for (la = 0; la < Nla; la++) { h[8+la] = L->lambda[la]; h[8+Nla+la] = getwlambda_line(L, la); }
      memcpy(h + 8 + 2*Nla, L->wphi, N*sizeof(double));
      {
        int nrow = (atmos.moving) ? 2*atmos.Nrays*Nla : Nla;
        double *p = rec_new("nl_phi", (long) nrow*N, a, kr, nrow, 0,0,0);
        for (n = 0; n < nrow; n++) memcpy(p + (long) n*N, L->phi[n], N*sizeof(double));
      }
    }
    for (kr = 0; kr < atom->Ncont; kr++) {
      AtomicContinuum *Cn = &atom->continuum[kr];
      int Nla = Cn->Nlambda;
      double *h = rec_new("nl_cont", 4 + 3*Nla, a, kr, Cn->i, Cn->j, Nla, Cn->Nblue);
      h[0] = Cn->lambda0; h[1] = Cn->alpha0; h[2] = Cn->hydrogenic; h[3] = Cn->isotope_frac;
      for (la = 0; la < Nla; la++) {
        h[4+la] = Cn->lambda[la]; h[4+Nla+la] = Cn->alpha[la]; h[4+2*Nla+la] = getwlambda_cont(Cn, la);
      }
    }
  }
  for (n = 0; n < Ns; n++) {
    ActiveSet *as = &spectrum.as[n];
    int cnt = 0, m;
    for (a = 0; a < atmos.Nactiveatom; a++) cnt += as->Nactiveatomrt[a];
    double *d = rec_new("nl_as", 3*cnt + 1, n, cnt, 0,0,0,0);
    cnt = 0;
    for (a = 0; a < atmos.Nactiveatom; a++) {
      Atom *atom = atmos.activeatoms[a];
      for (m = 0; m < as->Nactiveatomrt[a]; m++) {
        d[3*cnt] = a;
        if (as->art[a][m].type == ATOMIC_LINE) {
          d[3*cnt+1] = 0; d[3*cnt+2] = (double) (as->art[a][m].ptype.line - atom->line);
        } else {
          d[3*cnt+1] = 1; d[3*cnt+2] = (double) (as->art[a][m].ptype.continuum - atom->continuum);
        }
        cnt++;
      }
    }
    d[3*cnt] = 0;
  }
  (void) k;
}

void __real_Iterate(int NmaxIter, double iterLimit);
void __wrap_Iterate(int NmaxIter, double iterLimit)
{
  int a, N = atmos.Nspace, n;
  if ((probe_mask & PROBE_NLTE) && atmos.Nactiveatom > 0) dump_nlte_problem(NmaxIter, iterLimit);
  /* Iterate() frees atom->Gamma at exit but keeps n and J */
  __real_Iterate(NmaxIter, iterLimit);
  if ((probe_mask & PROBE_NLTE) && atmos.Nactiveatom > 0) {
    for (a = 0; a < atmos.Nactiveatom; a++) {
      Atom *atom = atmos.activeatoms[a];
      rec_copy("nl_n_final", atom->n[0], (long) atom->Nlevel*N, a,0,0,0);
    }
    double *J = rec_new("nl_J_final", (long) spectrum.Nspect*N, 0,0,0,0,0,0);
    for (n = 0; n < spectrum.Nspect; n++) memcpy(J + (long) n*N, spectrum.J[n], N*sizeof(double));
  }
}

static int up_count = 0;
double __real_updatePopulations(int niter);
double __wrap_updatePopulations(int niter)
{
  int a, N = atmos.Nspace, kr;
  if (probe_mask & PROBE_NLTE) {
    for (a = 0; a < atmos.Nactiveatom; a++) {
      Atom *atom = atmos.activeatoms[a];
      int Nl = atom->Nlevel;
      rec_copy("up_gamma", atom->Gamma[0], (long) Nl*Nl*N, a, niter, 0,0);
      double *r = rec_new("up_rates", (long) 2*(atom->Nline + atom->Ncont)*N, a, niter, atom->Nline, atom->Ncont, 0,0);
      for (kr = 0; kr < atom->Nline; kr++) {
        memcpy(r + (long) (2*kr)*N, atom->line[kr].Rij, N*sizeof(double));
        memcpy(r + (long) (2*kr+1)*N, atom->line[kr].Rji, N*sizeof(double));
      }
      for (kr = 0; kr < atom->Ncont; kr++) {
        memcpy(r + (long) (2*(atom->Nline+kr))*N, atom->continuum[kr].Rij, N*sizeof(double));
        memcpy(r + (long) (2*(atom->Nline+kr)+1)*N, atom->continuum[kr].Rji, N*sizeof(double));
      }
    }
  }
  double dp = __real_updatePopulations(niter);
  if (probe_mask & PROBE_NLTE) {
    for (a = 0; a < atmos.Nactiveatom; a++) {
      Atom *atom = atmos.activeatoms[a];
      double *d = rec_new("up_n", (long) atom->Nlevel*N + 1, a, niter, 0,0,0,0);
      memcpy(d, atom->n[0], (long) atom->Nlevel*N*sizeof(double));
      d[(long) atom->Nlevel*N] = dp;
    }
  }
  up_count++;
  return dp;
}

void __real_SolveLinearEq(int N, double **A, double *b, bool_t improve);
void __wrap_SolveLinearEq(int N, double **A, double *b, bool_t improve)
{
  double *d = NULL;
  int i;
  if ((probe_mask & PROBE_NLTE) && N <= 32) {
    d = rec_new("lu", (long) N*N + 2*N, N, improve, 0,0,0,0);
    for (i = 0; i < N; i++) memcpy(d + (long) i*N, A[i], N*sizeof(double));
    memcpy(d + (long) N*N, b, N*sizeof(double));
  }
  __real_SolveLinearEq(N, A, b, improve);
  if (d) memcpy(d + (long) N*N + N, b, N*sizeof(double));
}

/* The final single-mu formal solution of _solveray() (rh/rhf1d/pyrh_solveray.c:75-106):
   background and line profiles were recomputed for the new angle set; record them + the result */
double __real_solveSpectrum(bool_t eval_operator, bool_t redistribute);
double __wrap_solveSpectrum(bool_t eval_operator, bool_t redistribute)
{
  int rec = (probe_mask & PROBE_NLTE) && atmos.Nactiveatom > 0 && atmos.Nrays == 1 && !spectrum.updateJ;
  int N = atmos.Nspace, Ns = spectrum.Nspect, n, a, kr;
  if (rec) {
    double *b = rec_new("fs_bg", (long) 3*Ns*N, 0,0,0,0,0,0);
    double *f = rec_new("fs_bgflags", 2*Ns, 0,0,0,0,0,0);
    for (n = 0; n < Ns; n++) {
      memcpy(b + (long) n*N, spectrum.chi_c_lam[n], N*sizeof(double));
      memcpy(b + (long) (Ns + n)*N, spectrum.eta_c_lam[n], N*sizeof(double));
      memcpy(b + (long) (2*Ns + n)*N, spectrum.sca_c_lam[n], N*sizeof(double));
      f[2*n] = atmos.backgrflags[n].hasline; f[2*n+1] = atmos.backgrflags[n].ispolarized;
    }
    rec_copy("fs_muz", geometry.muz, 1, 0,0,0,0);
    rec_copy("fs_wmu", geometry.wmu, 1, 0,0,0,0);
    for (a = 0; a < atmos.Nactiveatom; a++) {
      Atom *atom = atmos.activeatoms[a];
      for (kr = 0; kr < atom->Nline; kr++) {
        AtomicLine *L = &atom->line[kr];
        int nrow = 2*L->Nlambda;
        double *p = rec_new("fs_phi", (long) nrow*N, a, kr, nrow, 0,0,0);
        for (n = 0; n < nrow; n++) memcpy(p + (long) n*N, L->phi[n], N*sizeof(double));
        rec_copy("fs_wphi", L->wphi, N, a, kr, 0,0);
        {   /* Profile() inputs as they stand at this (second) getProfiles call */
          double *ad = rec_new("fs_adamp", N, a, kr, L->Voigt, 0,0,0);
          memset(ad, 0, N*sizeof(double));
          if (L->Voigt) Damping(L, ad);
        }
      }
      rec_copy("fs_vbroad", atom->vbroad, N, a,0,0,0);
    }
    rec_copy("fs_vel", geometry.vel, N, 0,0,0,0);
  }
  double dJ = __real_solveSpectrum(eval_operator, redistribute);
  if (rec) {
    double *I = rec_new("fs_I", Ns, 0,0,0,0,0,0);
    for (n = 0; n < Ns; n++) I[n] = spectrum.I[n][0];
  }
  return dJ;
}

/* ------------------------------------------------------------------ Zeeman helpers
   Direct calls into the reference's host-side Zeeman machinery (kurucz.c:832-969, zeeman.c:37-299)
   on caller-built structs: used by oracle/gen_golden_zeeman.py only. */
ZeemanMultiplet *RLKZeeman(RLK_Line *rlk);
bool_t RLKdeterminate(char *labeli, char *labelj, RLK_Line *rlk);

int probe_rlk_determinate(const char *labeli, const char *labelj, double gi, double gj, double *out /*Si,Li,Sj,Lj*/)
{
  RLK_Line r;
  char li[64], lj[64];
  memset(&r, 0, sizeof(r));
  r.gi = gi; r.gj = gj;
  strncpy(li, labeli, 63); li[63] = 0;
  strncpy(lj, labelj, 63); lj[63] = 0;
  int ok = RLKdeterminate(li, lj, &r);
  out[0] = r.Si; out[1] = r.Li; out[2] = r.Sj; out[3] = r.Lj;
  return ok;
}

int probe_rlk_zeeman(double gi, double gj, double Si, int Li, double Sj, int Lj, double gL_i, double gL_j,
                     int LS_Lande, int cap, int *q, double *shift, double *strength)
{
  RLK_Line r;
  int n, save = input.LS_Lande;
  memset(&r, 0, sizeof(r));
  r.gi = gi; r.gj = gj; r.Si = Si; r.Li = Li; r.Sj = Sj; r.Lj = Lj; r.gL_i = gL_i; r.gL_j = gL_j;
  input.LS_Lande = LS_Lande;
  ZeemanMultiplet *zm = RLKZeeman(&r);
  input.LS_Lande = save;
  int nc = zm->Ncomponent;
  for (n = 0; n < nc && n < cap; n++) { q[n] = zm->q[n]; shift[n] = zm->shift[n]; strength[n] = zm->strength[n]; }
  free(zm->q); free(zm->shift); free(zm->strength); free(zm);
  return nc;
}

/* Zeeman() of a model-atom line (zeeman.c:186-281): labels/statistical weights of the two levels */
int probe_zeeman_atom(const char *label_i, double g_i, const char *label_j, double g_j, double g_Lande_eff,
                      int cap, int *q, double *shift, double *strength)
{
  Atom atom;
  AtomicLine line;
  char *labels[2], li[ATOM_LABEL_WIDTH+1], lj[ATOM_LABEL_WIDTH+1];
  double g[2];
  int n;
  memset(&atom, 0, sizeof(atom)); memset(&line, 0, sizeof(line));
  strncpy(li, label_i, ATOM_LABEL_WIDTH); li[ATOM_LABEL_WIDTH] = 0;
  strncpy(lj, label_j, ATOM_LABEL_WIDTH); lj[ATOM_LABEL_WIDTH] = 0;
  labels[0] = li; labels[1] = lj; g[0] = g_i; g[1] = g_j;
  atom.label = labels; atom.g = g;
  line.atom = &atom; line.i = 0; line.j = 1; line.g_Lande_eff = g_Lande_eff;
  ZeemanMultiplet *zm = Zeeman(&line);
  int nc = zm->Ncomponent;
  for (n = 0; n < nc && n < cap; n++) { q[n] = zm->q[n]; shift[n] = zm->shift[n]; strength[n] = zm->strength[n]; }
  free(zm->q); free(zm->shift); free(zm->strength); free(zm);
  return nc;
}

int probe_determinate(const char *label, double g, double *out /* n, S, L, J */)
{
  char l[ATOM_LABEL_WIDTH+1];
  int n = 0, L = 0;
  double S = 0, J = 0;
  strncpy(l, label, ATOM_LABEL_WIDTH); l[ATOM_LABEL_WIDTH] = 0;
  int ok = determinate(l, g, &n, &S, &L, &J);
  out[0] = n; out[1] = S; out[2] = L; out[3] = J;
  return ok;
}

/* ------------------------------------------------------------------ molecular background lines
   MolecularOpacity (rh/opacity.c:711-839): outputs per (wavelength, mu, direction) and, once, the
   molecules that own a line list: n, pf, vbroad [Nspace] and the line table with Zeeman patterns. */
ZeemanMultiplet *MolZeeman(MolecularLine *mrt);
flags __real_MolecularOpacity(double lambda, int nspect, int mu, bool_t to_obs,
                              double *chi, double *eta, double *chip);
flags __wrap_MolecularOpacity(double lambda, int nspect, int mu, bool_t to_obs,
                              double *chi, double *eta, double *chip)
{
  flags f = __real_MolecularOpacity(lambda, nspect, mu, to_obs, chi, eta, chip);
  if (probe_mask & PROBE_RLK) {
    int N = atmos.Nspace, ns = atmos.Stokes ? 4 : 1, n, kr, c;
    if (f.hasline) {
      double *d = rec_new("mol", 2*ns*N, nspect, mu, to_obs, f.hasline, f.ispolarized, ns);
      memcpy(d, chi, ns*N*sizeof(double));
      memcpy(d + ns*N, eta, ns*N*sizeof(double));
    }
    if (!mol_snapshot_done && mu == atmos.Nrays-1 && to_obs) {
      for (n = 0; n < atmos.Nmolecule; n++) {
        Molecule *m = &atmos.molecules[n];
        if (m->Nrt <= 0 || m->active) continue;
        double *d = rec_new("mol_col", 3*N, n, m->Nrt, 0,0,0,0);
        memcpy(d, m->n, N*sizeof(double));
        memcpy(d + N, m->pf, N*sizeof(double));
        memcpy(d + 2*N, m->vbroad, N*sizeof(double));
        for (kr = 0; kr < m->Nrt; kr++) {
          MolecularLine *l = &m->mrt[kr];
          ZeemanMultiplet *zm = l->zm;
          int own = 0;
          if (l->polarizable && zm == NULL) {            /* lazily built in the reference too (opacity.c:796) */
            double keep = l->g_Lande_eff;                /* MolZeeman() stores g_eff in the line (molzeeman.c:308): left there, */
            zm = MolZeeman(l); own = 1;                  /* the reference's own later call would build a normal triplet instead */
            l->g_Lande_eff = keep;
          }
          int nc = zm ? zm->Ncomponent : 0;
          double *r = rec_new("mol_line", 10 + 3*nc, n, kr, nc, l->polarizable, 0, 0);
          r[0] = l->lambda0; r[1] = l->Ei; r[2] = l->gi; r[3] = l->Bij; r[4] = l->Aji; r[5] = l->Bji;
          r[6] = l->isotope_frac; r[7] = l->qwing; r[8] = l->polarizable; r[9] = l->gj;
          for (c = 0; c < nc; c++) { r[10+c] = zm->q[c]; r[10+nc+c] = zm->shift[c]; r[10+2*nc+c] = zm->strength[c]; }
          if (own) { free(zm->q); free(zm->shift); free(zm->strength); free(zm); }
        }
      }
      mol_snapshot_done = 1;
    }
  }
  return f;
}

/* ------------------------------------------------------------------ passive bound-bound lines
   passive_bb (rh/metal.c:174-344): outputs per (wavelength, mu, direction) that found a line, and once per
   contributing line its parameters + the per-depth inputs (n_i, n_j, vbroad, Damping()). */
flags __real_passive_bb(double lambda, int nspect, int mu, bool_t to_obs, double *chi, double *eta, double *chip);
flags __wrap_passive_bb(double lambda, int nspect, int mu, bool_t to_obs, double *chi, double *eta, double *chip)
{
  flags f = __real_passive_bb(lambda, nspect, mu, to_obs, chi, eta, chip);
  if ((probe_mask & PROBE_RLK) && f.hasline) {
    int N = atmos.Nspace, m, kr, l, c;
    double *d = rec_new("pbb", 2*N, nspect, mu, to_obs, 0, 0, 0);
    memcpy(d, chi, N*sizeof(double));
    memcpy(d + N, eta, N*sizeof(double));
    for (m = 0; m < atmos.Natom; m++) {
      Atom *atom = atmos.atoms + m;
      if (atom->active) continue;
      double **n = (atom->n != atom->nstar) ? atom->n : atom->nstar;
      for (kr = 0; kr < atom->Nline; kr++) {
        AtomicLine *line = atom->line + kr;
        double dlambda = line->lambda0 * line->qwing * (atmos.vmicro_char / CLIGHT);
        if (!(fabs(lambda - line->lambda0) <= dlambda)) continue;
        int seen = 0;
        for (l = 0; l < pbb_nseen; l++) if (pbb_seen[l] == line) seen = 1;
        if (seen || pbb_nseen == PBB_MAXSEEN) continue;
        pbb_seen[pbb_nseen++] = line;
        int nc = line->Ncomponent;
        double *r = rec_new("pbb_line", 8 + 2*nc + 4L*N, m, kr, nc, line->Voigt, line->i, line->j);
        r[0] = line->lambda0; r[1] = line->qwing; r[2] = line->Bij; r[3] = line->Bji; r[4] = line->Aji;
        r[5] = line->Voigt; r[6] = nc; r[7] = 0.0;
        for (c = 0; c < nc; c++) { r[8+c] = line->c_shift[c]; r[8+nc+c] = line->c_fraction[c]; }
        double *a = r + 8 + 2*nc;
        memcpy(a, n[line->i], N*sizeof(double));
        memcpy(a + N, n[line->j], N*sizeof(double));
        memcpy(a + 2*N, atom->vbroad, N*sizeof(double));
        if (line->Voigt) Damping(line, a + 3*N); else memset(a + 3*N, 0, N*sizeof(double));
      }
    }
  }
  return f;
}

/* ------------------------------------------------------------------ background continuum
   The angle-independent contributions Background() sums per wavelength (rh/background.c:343-465).
   PROBE_CONT records each contribution's output (tag "cont", meta[0] = function id below, data[0] = lambda)
   and, once, every input they read: level populations of all model atoms, bound-free continua with their
   cross-section tables, the ground-state lines Rayleigh() sums, molecular densities, nHmin. */
#define PROBE_CONT 256
enum { CF_THOMSON = 0, CF_HMINUS_BF, CF_HMINUS_FF, CF_OH_BF, CF_CH_BF, CF_H_BF, CF_H_FF, CF_RAYLEIGH_H,
       CF_RAYLEIGH_HE, CF_H2PLUS_FF, CF_RAYLEIGH_H2, CF_H2MINUS_FF, CF_METAL_BF };
static void cont_snapshot(void)
{
  int N = atmos.Nspace, m, i, kr, nlev = 0, ncont = 0, ntab = 0, nray = 0;
  if (cont_snapshot_done) return;
  cont_snapshot_done = 1;
  for (m = 0; m < atmos.Natom; m++) {
    Atom *a = &atmos.atoms[m];
    nlev += a->Nlevel; ncont += a->Ncont;
    for (kr = 0; kr < a->Ncont; kr++) ntab += a->continuum[kr].Nlambda;
    if (m < 2) for (kr = 0; kr < a->Nline; kr++) if (a->line[kr].i == 0) nray++;
  }
  double *h = rec_new("ct_hdr", 16, atmos.Natom, nlev, ncont, ntab, nray, 0);
  h[0] = atmos.Natom; h[1] = nlev; h[2] = ncont; h[3] = ntab; h[4] = nray; h[5] = atmos.H->active;
  h[6] = (atmos.elements[1].model != NULL); h[7] = (atmos.OH != NULL); h[8] = (atmos.CH != NULL);
  h[9] = (atmos.H2 != NULL); h[10] = input.solve_NLTE; h[11] = atmos.vmicro_char; h[12] = input.do_fudge;
  h[13] = atmos.H->Nlevel; h[14] = atmos.moving; h[15] = atmos.Stokes;
  double *lev = rec_new("ct_lev", 5L*nlev, 0,0,0,0,0,0);          /* atom, E, stage, g, active */
  double *pn = rec_new("ct_n", (long) nlev*N, 0,0,0,0,0,0);
  double *ps = rec_new("ct_nstar", (long) nlev*N, 0,0,0,0,0,0);
  double *bf = rec_new("ct_bf", 10L*ncont, 0,0,0,0,0,0);
  double *tl = rec_new("ct_tab_lambda", ntab > 0 ? ntab : 1, 0,0,0,0,0,0);
  double *ta = rec_new("ct_tab_alpha", ntab > 0 ? ntab : 1, 0,0,0,0,0,0);
  double *ry = rec_new("ct_ray", 8L*(nray > 0 ? nray : 1), 0,0,0,0,0,0);
  int l0 = 0, c0 = 0, t0 = 0, r0 = 0;
  for (m = 0; m < atmos.Natom; m++) {
    Atom *a = &atmos.atoms[m];
    for (i = 0; i < a->Nlevel; i++) {
      lev[5*(l0+i)] = m; lev[5*(l0+i)+1] = a->E[i]; lev[5*(l0+i)+2] = a->stage[i]; lev[5*(l0+i)+3] = a->g[i];
      lev[5*(l0+i)+4] = a->active;
      memcpy(pn + (long) (l0+i)*N, a->n[i], N*sizeof(double));
      memcpy(ps + (long) (l0+i)*N, a->nstar[i], N*sizeof(double));
    }
    for (kr = 0; kr < a->Ncont; kr++) {
      AtomicContinuum *c = &a->continuum[kr];
      double *b = bf + 10L*(c0+kr);
      b[0] = m; b[1] = l0 + c->i; b[2] = l0 + c->j; b[3] = c->lambda0; b[4] = c->lambda[0]; b[5] = c->hydrogenic;
      b[6] = c->alpha0; b[7] = c->Nlambda; b[8] = t0; b[9] = a->active;
      memcpy(tl + t0, c->lambda, c->Nlambda*sizeof(double));
      memcpy(ta + t0, c->alpha, c->Nlambda*sizeof(double));
      t0 += c->Nlambda;
    }
    if (m < 2) for (kr = 0; kr < a->Nline; kr++) if (a->line[kr].i == 0) {
      AtomicLine *L = &a->line[kr];
      double *r = ry + 8L*r0++;
      r[0] = m; r[1] = L->lambda0; r[2] = L->qwing; r[3] = L->Aji; r[4] = a->g[L->j]; r[5] = a->g[0]; r[6] = l0; r[7] = a->stage[0];
    }
    l0 += a->Nlevel; c0 += a->Ncont;
  }
  rec_copy("ct_T", atmos.T, N, 0,0,0,0);
  rec_copy("ct_ne", atmos.ne, N, 0,0,0,0);
  rec_copy("ct_nHmin", atmos.nHmin, N, 0,0,0,0);
  if (atmos.H2) rec_copy("ct_nH2", atmos.H2->n, N, 0,0,0,0);
  if (atmos.OH) rec_copy("ct_nOH", atmos.OH->n, N, 0,0,0,0);
  if (atmos.CH) rec_copy("ct_nCH", atmos.CH->n, N, 0,0,0,0);
}

static void cont_rec(int id, double lambda, int ok, const double *a, const double *b)
{
  int N = atmos.Nspace;
  double *d = rec_new("cont", 1 + 2L*N, id, ok, 0,0,0,0);
  d[0] = lambda;
  if (a && ok) memcpy(d + 1, a, N*sizeof(double)); else memset(d + 1, 0, N*sizeof(double));
  if (b && ok) memcpy(d + 1 + N, b, N*sizeof(double)); else memset(d + 1 + N, 0, N*sizeof(double));
}
#define CONT_ON ((probe_mask & PROBE_CONT) && atmos.active_layer == -1)

void __real_Thomson(double *chi);
void __wrap_Thomson(double *chi) { __real_Thomson(chi); if (CONT_ON) { cont_snapshot(); cont_rec(CF_THOMSON, 0.0, 1, chi, NULL); } }
#define WRAP2(NAME, ID) \
  bool_t __real_##NAME(double lambda, double *chi, double *eta); \
  bool_t __wrap_##NAME(double lambda, double *chi, double *eta) { \
    bool_t ok = __real_##NAME(lambda, chi, eta); if (CONT_ON) cont_rec(ID, lambda, ok, chi, eta); return ok; }
#define WRAP1(NAME, ID) \
  bool_t __real_##NAME(double lambda, double *chi); \
  bool_t __wrap_##NAME(double lambda, double *chi) { \
    bool_t ok = __real_##NAME(lambda, chi); if (CONT_ON && lambda != 0.0) cont_rec(ID, lambda, ok, chi, NULL); return ok; }
WRAP2(Hminus_bf, CF_HMINUS_BF)
WRAP1(Hminus_ff, CF_HMINUS_FF)
WRAP2(OH_bf_opac, CF_OH_BF)
WRAP2(CH_bf_opac, CF_CH_BF)
WRAP2(Hydrogen_bf, CF_H_BF)
WRAP1(H2plus_ff, CF_H2PLUS_FF)
WRAP1(Rayleigh_H2, CF_RAYLEIGH_H2)
WRAP1(H2minus_ff, CF_H2MINUS_FF)
void __real_Hydrogen_ff(double lambda, double *chi);
void __wrap_Hydrogen_ff(double lambda, double *chi) { __real_Hydrogen_ff(lambda, chi); if (CONT_ON) cont_rec(CF_H_FF, lambda, 1, chi, NULL); }
bool_t __real_Rayleigh(double lambda, Atom *atom, double *scatt);
bool_t __wrap_Rayleigh(double lambda, Atom *atom, double *scatt)
{
  bool_t ok = __real_Rayleigh(lambda, atom, scatt);
  if (CONT_ON) cont_rec(atom == atmos.H ? CF_RAYLEIGH_H : CF_RAYLEIGH_HE, lambda, ok, scatt, NULL);
  return ok;
}
bool_t __real_Metal_bf(double lambda, int Nmetal, struct Atom *metals, double *chi, double *eta);
bool_t __wrap_Metal_bf(double lambda, int Nmetal, struct Atom *metals, double *chi, double *eta)
{
  bool_t ok = __real_Metal_bf(lambda, Nmetal, metals, chi, eta);
  if (CONT_ON) cont_rec(CF_METAL_BF, lambda, ok, chi, eta);
  return ok;
}

/* ChemicalEquilibrium (rh/chemequil.c:107-392) rescales the populations of every model atom that is bound in
   molecules by n_after / ntotal_before (:336-342): record both ntotal arrays and the abundances. */
void __real_ChemicalEquilibrium(int NmaxIter, double iterLimit);
void __wrap_ChemicalEquilibrium(int NmaxIter, double iterLimit)
{
  int rec = (probe_mask & PROBE_CONT) && atmos.active_layer == -1, N = atmos.Nspace, m;
  double *pre = NULL;
  if (rec) {
    pre = rec_new("ce_ntotal_pre", (long) atmos.Natom*N, atmos.Natom, 0,0,0,0,0);
    for (m = 0; m < atmos.Natom; m++) memcpy(pre + (long) m*N, atmos.atoms[m].ntotal, N*sizeof(double));
    double *ab = rec_new("ce_abundance", atmos.Natom, 0,0,0,0,0,0);
    for (m = 0; m < atmos.Natom; m++) ab[m] = atmos.atoms[m].abundance;
    /* the chemical network: nuclei (elements bound in some molecule, atmos.elements order) with the index of
       their model atom, and per molecule the fit data equilconstant() reads (chemequil.c:456-530) */
    {
      int i, j, nn = 0, nu, idx[128];
      for (i = 0; i < atmos.Nelem; i++) if (atmos.elements[i].Nmolecule > 0) idx[nn++] = i;
      double *nuc = rec_new("ce_nuclei", 2L*nn, nn, 0,0,0,0,0);
      for (i = 0; i < nn; i++) {
        Element *e = &atmos.elements[idx[i]];
        nuc[2*i] = idx[i];
        nuc[2*i+1] = e->model ? (double) (e->model - atmos.atoms) : -1.0;
      }
      for (i = 0; i < atmos.Nmolecule; i++) {
        Molecule *mo = &atmos.molecules[i];
        double *r = rec_new("ce_mol", 32, i, mo->fit, mo->charge, mo->Nnuclei, mo->Nelement, mo->Neqc);
        memset(r, 0, 32*sizeof(double));
        r[0] = mo->fit; r[1] = mo->charge; r[2] = mo->Nnuclei; r[3] = mo->Nelement; r[4] = mo->Neqc;
        r[5] = mo->Tmin; r[6] = mo->Tmax; r[7] = mo->Ediss;
        for (j = 0; j < mo->Neqc && j < 8; j++) r[8+j] = mo->eqc_coef[j];
        for (j = 0; j < mo->Nelement && j < 4; j++) {
          for (nu = 0; nu < nn; nu++) if (idx[nu] == mo->pt_index[j]) r[16+j] = nu;
          r[20+j] = mo->pt_count[j];
        }
        r[24] = (mo == atmos.H2); r[25] = (mo == atmos.OH); r[26] = (mo == atmos.CH);
      }
    }
  }
  __real_ChemicalEquilibrium(NmaxIter, iterLimit);
  if (rec) {
    double *post = rec_new("ce_ntotal_post", (long) atmos.Natom*N, atmos.Natom, 0,0,0,0,0);
    for (m = 0; m < atmos.Natom; m++) memcpy(post + (long) m*N, atmos.atoms[m].ntotal, N*sizeof(double));
  }
}
