/* integration/pyrh_b200_bridge.c -- the RH host side of the B200 path.
 *
 * Compiled INTO the reference's pyrh C library (together with integration/pyrh_b200.patch, by
 * integration/build_bridged.sh) and linked against librhb200.so.  The reference's own host code keeps doing what it
 * does once per call -- readInput, readAbundance, readAtomicModels, readMolecularModels, readKuruczLines, Bproject,
 * SortLambda (rh/rhf1d/pyrh_compute1dray.c:112-307) -- and this file
 *   1. flattens that parsed state (atmos.rlk_lines, atmos.elements, atmos.atoms, atmos.molecules, spectrum.lambda,
 *      spectrum.as[], the collisional sections of ACTIVE atoms) into the tables of include/rhb200.h,
 *   2. runs everything per column on the device (rhb200_compute1d_batch / rhb200_compute1d_rf_batch in LTE,
 *      rhb200_nlte_compute1d_batch with ACTIVE atoms) instead of Background() ... Iterate() ... _solveray(),
 *   3. hands back the reference's own `mySpectrum` (rh/rhf1d/pyrh_compute1dray.h:12-19): malloc'd lam/sI/sQ/sU/sV,
 *      rfs, atom_pops pointing at the live atom->n / atom->nstar like _solveray() does (pyrh_solveray.c:119-186).
 * `rhf1d()` keeps its exact signature; `rhf1d_batch()` is the non-breaking addition for many columns.
 * The other entry points rh.pxd binds are hooked the same way in rh/rhf1d/pyrh_hse.c: hse() -> rhb200_hse_batch,
 * get_scales() -> rhb200_get_scales_batch, get_ne_from_nH() -> rhb200_solve_ne_batch (end of this file).
 *
 * Nothing here is a fallback: when the device library refuses a configuration the call aborts through the reference's
 * Error(ERROR_LEVEL_2) convention.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rh.h"
#include "atom.h"
#include "atmos.h"
#include "spectrum.h"
#include "geometry.h"
#include "inputs.h"
#include "constant.h"
#include "background.h"
#include "error.h"
#include "pyrh_compute1dray.h"
#include "rhb200.h"

extern Atmosphere atmos;
extern Geometry geometry;
extern Spectrum spectrum;
extern InputData input;
extern char messageStr[];

ZeemanMultiplet *RLKZeeman(RLK_Line *rlk);          /* kurucz.c:832 */

#define NPL RHB200_PL_NFIELD
/* the pyrh build silences Error()'s text (options.c:63-65) before it exit()s: say why on stderr first */
#define FAIL(msg) do { sprintf(messageStr, "pyrh_b200 bridge: %s", msg); fprintf(stderr, "%s\n", messageStr); \
                       Error(ERROR_LEVEL_2, "pyrh_b200", messageStr); } while (0)
#define CHECK(call) do { if ((call) != RHB200_OK) { snprintf(messageStr, 900, "pyrh_b200 bridge: %.200s -> %s", #call, rhb200_last_error()); \
                                                   fprintf(stderr, "%s\n", messageStr); \
                                                   Error(ERROR_LEVEL_2, "pyrh_b200", messageStr); } } while (0)

/* ---- state kept between calls: one device context, the pyrh-unit copy of the caller's rows, a pending batch */
static rhb200_ctx *g_ctx = NULL;
/* PYRH_B200_TRACE=1: FNV-1a hash of every table handed to the library (and the Kurucz line rows themselves), to stderr:
   two calls that should be identical and are not show which parsed table differs -- the reference's readers have
   state of their own (e.g. the stale-buffer read of kurucz.c:273) */
static void trace(const char *name, const void *p, size_t bytes)
{
  const unsigned char *q = (const unsigned char *) p;
  unsigned long long h = 1469598103934665603ull;
  size_t i;
  if (!getenv("PYRH_B200_TRACE")) return;
  for (i = 0; p && i < bytes; i++) { h ^= q[i]; h *= 1099511628211ull; }
  fprintf(stderr, "pyrh_b200 trace: %-12s %9lu bytes %016llx\n", name, (unsigned long) bytes, h);
}
static double *g_rows = NULL;       /* [9][Ndep] pyrh units, saved before rhf1d() converts the caller's arrays in place */
static int g_rows_ndep = 0, g_atm_scale = 0;
static struct { int ncol; const double *atm; double *stokes, *pops_n, *pops_nstar; int *niter; } g_batch = {0};

typedef struct { double *v; int n, cap; } dvec;
static void dv_push(dvec *a, double x)
{
  if (a->n == a->cap) { a->cap = a->cap ? 2*a->cap : 64; a->v = (double *) realloc(a->v, a->cap * sizeof(double)); }
  a->v[a->n++] = x;
}
typedef struct { int *v; int n, cap; } ivec;
static void iv_push(ivec *a, int x)
{
  if (a->n == a->cap) { a->cap = a->cap ? 2*a->cap : 64; a->v = (int *) realloc(a->v, a->cap * sizeof(int)); }
  a->v[a->n++] = x;
}

/* called first thing in rhf1d() (patch): the caller's rows in pyrh units, before the in-place conversion to SI
   (pyrh_compute1dray.c:263-270) -- the device does that conversion itself, with the reference's expressions */
void pyrh_b200_save_inputs(int Ndep, int atm_scale, double *scale, double *temp, double *ne, double *vz, double *vmic,
                           double *mag, double *gamma, double *chi, double *nH)
{
  double *src[9] = {scale, temp, ne, vz, vmic, mag, gamma, chi, nH};
  int r;
  if (Ndep != g_rows_ndep) { g_rows = (double *) realloc(g_rows, (size_t) 9 * Ndep * sizeof(double)); g_rows_ndep = Ndep; }
  for (r = 0; r < 9; r++) {
    if (src[r]) memcpy(g_rows + (size_t) r * Ndep, src[r], Ndep * sizeof(double));
    else memset(g_rows + (size_t) r * Ndep, 0, Ndep * sizeof(double));          /* get_scales() has no field rows */
  }
  g_atm_scale = atm_scale;
  /* rhf1d() only ever SETS atmos.Nloggf / Nlam (pyrh_compute1dray.c:199-211): a call without log gf / wavelength
     overrides after one with them would make readKuruczLines() read the previous caller's freed arrays */
  atmos.Nloggf = 0; atmos.Nlam = 0;
}

/* ---- published opacity tables (pyrh_b200/data/opacity_tables.txt; hydrogen.c / ohchbf.c keep them function-static) */
typedef struct { char name[24]; int n; double *v; } table_t;
static table_t g_tab[32];
static int g_ntab = 0;
static const double *tab(const char *name, int *n)
{
  int i;
  for (i = 0; i < g_ntab; i++) if (!strcmp(g_tab[i].name, name)) { if (n) *n = g_tab[i].n; return g_tab[i].v; }
  sprintf(messageStr, "opacity table %s missing", name);
  Error(ERROR_LEVEL_2, "pyrh_b200", messageStr);
  return NULL;
}
static void load_tables(void)
{
  const char *dir = getenv("RHB200_DATA");
  char path[1024];
  FILE *fp;
  if (g_ntab) return;
  if (!dir) FAIL("RHB200_DATA (directory of opacity_tables.txt) is not set");
  snprintf(path, sizeof path, "%s/opacity_tables.txt", dir);
  if (!(fp = fopen(path, "r"))) FAIL("cannot open opacity_tables.txt");
  while (g_ntab < 32 && fscanf(fp, "%23s %d", g_tab[g_ntab].name, &g_tab[g_ntab].n) == 2) {
    table_t *t = &g_tab[g_ntab++];
    int i;
    t->v = (double *) malloc(t->n * sizeof(double));
    for (i = 0; i < t->n; i++) if (fscanf(fp, "%lf", &t->v[i]) != 1) FAIL("opacity_tables.txt is truncated");
  }
  fclose(fp);
}

/* ---- Damping() constants of one model-atom line -> RHB200_PL_* row (broad.c:60-264; depth-independent factors) */
static void line_row(Atom *atom, int atom_index, int lev0, AtomicLine *line, int compoff, double *r)
{
  const double FOURPIEPS0 = 4.0 * PI * EPSILON_0;
  const double H_weight = atmos.elements[0].weight, He_weight = atmos.elements[1].weight, He_abund = atmos.elements[1].abund;
  const double weight = atom->weight;
  const int i = line->i, j = line->j;
  double *cv = line->cvdWaals;
  memset(r, 0, NPL * sizeof(double));
  r[RHB200_PL_ATOM] = atom_index; r[RHB200_PL_LEVEL_I] = lev0 + i; r[RHB200_PL_LEVEL_J] = lev0 + j;
  r[RHB200_PL_LAMBDA0] = line->lambda0; r[RHB200_PL_QWING] = line->qwing;
  r[RHB200_PL_BIJ] = line->Bij; r[RHB200_PL_BJI] = line->Bji; r[RHB200_PL_AJI] = line->Aji; r[RHB200_PL_VOIGT] = line->Voigt ? 1.0 : 0.0;
  r[RHB200_PL_NCOMP] = line->Ncomponent; r[RHB200_PL_COMPOFF] = compoff; r[RHB200_PL_GRAD] = line->Grad;
  r[RHB200_PL_WEIGHT] = weight; r[RHB200_PL_IS_H] = strstr(atom->ID, "H ") ? 1.0 : 0.0; r[RHB200_PL_HE_ABUND] = He_abund;
  r[RHB200_PL_VDW_TYPE] = -1;
  if (cv[0] > 0.0 || cv[2] > 0.0) {                                  /* VanderWaals, broad.c:60-140 */
    if (line->vdWaals == UNSOLD || line->vdWaals == BARKLEM) {
      const double vrel35_He = pow(8.0*KBOLTZMANN/(PI*AMU*weight) * (1.0 + weight/He_weight), 0.3);
      const int Z = atom->stage[j] + 1;
      int ic = j + 1;
      double d1, d2, deltaR, ZR, C625;
      while (atom->stage[ic] < atom->stage[j] + 1) ic++;
      d1 = E_RYDBERG/(atom->E[ic] - atom->E[j]); d2 = E_RYDBERG/(atom->E[ic] - atom->E[i]);
      deltaR = d1*d1 - d2*d2;
      ZR = Z * RBOHR;
      C625 = pow(2.5 * ((Q_ELECTRON*Q_ELECTRON)/FOURPIEPS0) * (ABARH/FOURPIEPS0) * 2*PI*(ZR*ZR)/HPLANCK * deltaR, 0.4);
      if (line->vdWaals == BARKLEM) {
        r[RHB200_PL_VDW_TYPE] = 2;
        r[RHB200_PL_VDW_A] = cv[0]; r[RHB200_PL_VDW_B] = (1.0 - cv[1])/2.0;
        r[RHB200_PL_VDW_C] = 8.08 * cv[2] * He_abund * vrel35_He * C625;
      } else {
        const double vrel35_H = pow(8.0*KBOLTZMANN/(PI*AMU*weight) * (1.0 + weight/H_weight), 0.3);
        r[RHB200_PL_VDW_TYPE] = 0;
        r[RHB200_PL_VDW_A] = 8.08 * (cv[0]*vrel35_H + cv[2]*He_abund*vrel35_He) * C625;
      }
    } else {                                                          /* RIDDER_RENSBERGEN */
      const double CUBE_CM = CM_TO_M*CM_TO_M*CM_TO_M;
      const double gH = 1.0E-8 * CUBE_CM * pow(1.0 + H_weight/weight, cv[1]);
      const double gHe = 1.0E-9 * CUBE_CM * pow(1.0 + He_weight/weight, cv[3]);
      r[RHB200_PL_VDW_TYPE] = 1;
      r[RHB200_PL_VDW_A] = gH*cv[0]; r[RHB200_PL_VDW_B] = cv[1]; r[RHB200_PL_VDW_C] = gHe*cv[2]; r[RHB200_PL_VDW_D] = cv[3];
    }
  }
  if (line->cStark < 0.0) {                                            /* Stark, broad.c:147-215 */
    r[RHB200_PL_STARK_TYPE] = 1; r[RHB200_PL_STARK_A] = fabs(line->cStark);
  } else if (line->cStark != 0.0) {
    const double m_electron = M_ELECTRON/AMU;
    const double Cc = 8.0*KBOLTZMANN/(PI*AMU*weight);
    const double Cm = pow(1.0 + weight/m_electron, 0.16666667) + pow(1.0 + weight/28.0, 0.16666667);
    const int Z = atom->stage[i] + 1;
    int ic = i + 1;
    double E_Ryd, neff_l, neff_u, Z2, tu, tl, C4;
    while (atom->stage[ic] < atom->stage[i] + 1 && ic < atom->Nlevel) ic++;
    E_Ryd = E_RYDBERG/(1.0 + M_ELECTRON/(weight*AMU));
    neff_l = Z*sqrt(E_Ryd/(atom->E[ic] - atom->E[i]));
    neff_u = Z*sqrt(E_Ryd/(atom->E[ic] - atom->E[j]));
    Z2 = Z*Z;
    tu = neff_u*(5.0*(neff_u*neff_u) + 1.0); tl = neff_l*(5.0*(neff_l*neff_l) + 1.0);
    C4 = ((Q_ELECTRON*Q_ELECTRON)/(4.0*PI*EPSILON_0)) * RBOHR * (2.0*PI*(RBOHR*RBOHR)/HPLANCK)/(18.0*Z2*Z2) * (tu*tu - tl*tl);
    r[RHB200_PL_STARK_TYPE] = 2; r[RHB200_PL_STARK_A] = 11.37*pow(line->cStark*C4, 0.66666667);
    r[RHB200_PL_STARK_C] = Cc; r[RHB200_PL_STARK_CM] = Cm;
  }
  if (strstr(atom->ID, "H ")) {                                        /* StarkLinear, broad.c:222-264 */
    int n_lower = 0, n_upper = 0;
    double a1;
    sscanf(atom->label[i], "H I %d", &n_lower);
    sscanf(atom->label[j], "H I %d", &n_upper);
    a1 = (n_upper - n_lower == 1) ? 0.642 : 1.0;
    r[RHB200_PL_LINSTARK_C] = a1 * 0.6 * (n_upper*n_upper - n_lower*n_lower) * (CM_TO_M*CM_TO_M);
  }
}

/* ---- everything that does not depend on the column: the reference's parsed state -> device tables */
typedef struct {
  int natom, nlev, *lev0, nmol;
  double *lam; int nlam, iref;
  int lrf_npar;
} tables_t;
static tables_t T;

static int g_no_kurucz = 0;          /* hse() / get_scales(): atmos.Nrlk = 0, no line list is read (pyrh_hse.c:145,441) */

static int stable_by_lambda0(RLK_Line *L, int n, int *order)
{
  int a, b;
  for (a = 0; a < n; a++) order[a] = a;
  for (a = 1; a < n; a++) {                        /* insertion sort: stable, like the library's Python host */
    const int x = order[a];
    for (b = a; b > 0 && L[order[b-1]].lambda0 > L[x].lambda0; b--) order[b] = order[b-1];
    order[b] = x;
  }
  return 0;
}

static void build_tables(int get_atomic_rfs, int fudge_num, double *fudge_lam, double *fudge)
{
  int n, m, i, kr, k;
  dvec lines = {0}, zs = {0}, zt = {0}, elems = {0}, pf = {0};
  ivec zq = {0};
  int *elem_row = (int *) malloc(atmos.Nelem * sizeof(int)), nelem = 0, *order;
  static double *keep[32];
  static int nkeep = 0;
  for (n = 0; n < nkeep; n++) free(keep[n]);
  nkeep = 0;

  if (input.magneto_optical && !g_no_kurucz) FAIL("MAGNETO_OPTICAL = TRUE is refused (the reference overflows chip_c there, readj.c:328)");
  if (!g_ctx && !(g_ctx = rhb200_open(getenv("RHB200_DEVICE") ? atoi(getenv("RHB200_DEVICE")) : 0))) FAIL(rhb200_last_error());

  /* -- Kurucz lines (Background() reads and sorts them, background.c:284-294) */
  if (atmos.Nrlk == 0 && !g_no_kurucz) readKuruczLines(input.KuruczData);
  for (n = 0; n < atmos.Nelem; n++) elem_row[n] = -1;
  for (n = 0; n < atmos.Nrlk; n++) {               /* element rows in order of first appearance in the files */
    const int e = atmos.rlk_lines[n].pt_index - 1;
    if (elem_row[e] < 0) elem_row[e] = nelem++;
  }
  order = (int *) malloc((atmos.Nrlk + 1) * sizeof(int));
  stable_by_lambda0(atmos.rlk_lines, atmos.Nrlk, order);
  for (n = 0; n < atmos.Nrlk; n++) {
    RLK_Line *rlk = &atmos.rlk_lines[order[n]];
    double r[RHB200_RL_NFIELD];
    memset(r, 0, sizeof r);
    r[RHB200_RL_LAMBDA0] = rlk->lambda0; r[RHB200_RL_GI] = rlk->gi; r[RHB200_RL_GJ] = rlk->gj;
    r[RHB200_RL_EI] = rlk->Ei; r[RHB200_RL_EJ] = rlk->Ej; r[RHB200_RL_BJI] = rlk->Bji; r[RHB200_RL_AJI] = rlk->Aji;
    r[RHB200_RL_BIJ] = rlk->Bij; r[RHB200_RL_GRAD] = rlk->Grad; r[RHB200_RL_GSTARK] = rlk->GStark; r[RHB200_RL_GVDW] = rlk->GvdWaals;
    r[RHB200_RL_HFS_FRAC] = rlk->hyperfine_frac; r[RHB200_RL_ISO_FRAC] = rlk->isotope_frac;
    r[RHB200_RL_CROSS] = rlk->cross; r[RHB200_RL_ALPHA] = (rlk->vdwaals == BARKLEM) ? rlk->alpha : 0.0;
    r[RHB200_RL_POLARIZABLE] = rlk->polarizable ? 1.0 : 0.0; r[RHB200_RL_VDWAALS] = (double) rlk->vdwaals;
    r[RHB200_RL_ELEM] = elem_row[rlk->pt_index - 1]; r[RHB200_RL_STAGE] = rlk->stage;
    r[RHB200_RL_ZOFF] = zq.n;
    if (rlk->polarizable) {
      if (rlk->zm == NULL) rlk->zm = RLKZeeman(rlk);            /* kurucz.c:659 */
      r[RHB200_RL_NCOMP] = rlk->zm->Ncomponent;
      for (i = 0; i < rlk->zm->Ncomponent; i++) { iv_push(&zq, rlk->zm->q[i]); dv_push(&zs, rlk->zm->shift[i]); dv_push(&zt, rlk->zm->strength[i]); }
    }
    for (i = 0; i < RHB200_RL_NFIELD; i++) dv_push(&lines, r[i]);
  }
  {
    int *row_elem = (int *) malloc((nelem + 1) * sizeof(int)), pfrow = 0;
    for (n = 0; n < atmos.Nelem; n++) if (elem_row[n] >= 0) row_elem[elem_row[n]] = n;
    for (m = 0; m < nelem; m++) {
      Element *el = &atmos.elements[row_elem[m]];
      double r[RHB200_RE_NFIELD];
      if (el->Nstage > RHB200_RE_MAXSTAGE) FAIL("element with more ionisation stages than RHB200_RE_MAXSTAGE");
      memset(r, 0, sizeof r);
      r[RHB200_RE_WEIGHT] = el->weight; r[RHB200_RE_ABUND] = el->abund; r[RHB200_RE_NSTAGE] = el->Nstage; r[RHB200_RE_PFROW] = pfrow;
      for (i = 0; i < el->Nstage; i++) r[RHB200_RE_IONPOT0 + i] = el->ionpot[i];
      for (i = 0; i < RHB200_RE_NFIELD; i++) dv_push(&elems, r[i]);
      for (i = 0; i < el->Nstage; i++) for (k = 0; k < atmos.Npf; k++) dv_push(&pf, el->pf[i][k]);
      pfrow += el->Nstage;
    }
    if (zq.n == 0) { iv_push(&zq, 0); dv_push(&zs, 0.0); dv_push(&zt, 0.0); zq.n = 0; }
    if (!pf.n) for (k = 0; k < atmos.Npf; k++) dv_push(&pf, 0.0);
    trace("lines", lines.v, lines.n * sizeof(double));
    if (getenv("PYRH_B200_TRACE")) for (n = 0; n < lines.n; n++) fprintf(stderr, "pyrh_b200 linefield %d %d %.17g\n", n / RHB200_RL_NFIELD, n % RHB200_RL_NFIELD, lines.v[n]);
 trace("zs", zs.v, zq.n * sizeof(double)); trace("zt", zt.v, zq.n * sizeof(double));
    trace("elems", elems.v, elems.n * sizeof(double)); trace("pf", pf.v, pf.n * sizeof(double)); trace("Tpf", atmos.Tpf, atmos.Npf * sizeof(double));
    CHECK(rhb200_set_lines(g_ctx, atmos.Nrlk, lines.v, zq.n, zq.v, zs.v, zt.v, nelem, elems.v, nelem ? pf.n / atmos.Npf : 0,
                           atmos.Npf, pf.v, atmos.Tpf, atmos.vmicro_char, 0, input.rlkscatter ? 1 : 0));
    free(row_elem);
  }
  /* get_atomic_rfs: parameter p <-> table row of the line that carries it (kurucz.c:250-259) */
  T.lrf_npar = 0;
  if (get_atomic_rfs && atmos.Nloggf > 0) {
    int *rows = (int *) malloc(atmos.Nloggf * sizeof(int)), p;
    for (p = 0; p < atmos.Nloggf; p++) rows[p] = -1;
    for (n = 0; n < atmos.Nrlk; n++) {
      RLK_Line *rlk = &atmos.rlk_lines[order[n]];
      if (rlk->get_loggf_rf && rlk->loggf_rf_ind >= 0 && rlk->loggf_rf_ind < atmos.Nloggf) rows[rlk->loggf_rf_ind] = n;
    }
    CHECK(rhb200_set_loggf_rf(g_ctx, atmos.Nloggf, rows));
    T.lrf_npar = atmos.Nloggf;
    free(rows);
  }
  free(order);

  /* -- model atoms: level table, bound-free continua, Rayleigh lines, passive / model / active line tables */
  {
    dvec lev = {0}, bf = {0}, tl = {0}, ta = {0}, ray = {0}, pl = {0}, cs = {0}, cf = {0}, ml = {0}, ab = {0};
    rhb200_continuum_model M;
    int any_active = 0, l0 = 0;
    T.natom = atmos.Natom;
    T.lev0 = (int *) realloc(T.lev0, (atmos.Natom + 1) * sizeof(int));
    for (m = 0; m < atmos.Natom; m++) {
      Atom *atom = &atmos.atoms[m];
      const double act = atom->active ? 1.0 : 0.0;
      T.lev0[m] = l0;
      if (atom->active) any_active = 1;
      dv_push(&ab, atom->abundance);
      for (i = 0; i < atom->Nlevel; i++) { dv_push(&lev, m); dv_push(&lev, atom->E[i]); dv_push(&lev, atom->stage[i]); dv_push(&lev, atom->g[i]); dv_push(&lev, act); }
      for (kr = 0; kr < atom->Ncont; kr++) {
        AtomicContinuum *c = &atom->continuum[kr];
        /* an ACTIVE atom's continua were remapped onto the merged grid by SortLambda; Background() never looks at them */
        const int nla = atom->active ? 3 : c->Nlambda;
        dv_push(&bf, m); dv_push(&bf, l0 + c->i); dv_push(&bf, l0 + c->j); dv_push(&bf, c->lambda0); dv_push(&bf, c->lambda[0]);
        dv_push(&bf, c->hydrogenic ? 1.0 : 0.0); dv_push(&bf, c->alpha0); dv_push(&bf, nla); dv_push(&bf, tl.n); dv_push(&bf, act);
        for (i = 0; i < nla; i++) { dv_push(&tl, atom->active ? c->lambda0 - (nla - 1 - i) : c->lambda[i]); dv_push(&ta, atom->active ? 0.0 : c->alpha[i]); }
      }
      if (m < 2)                                    /* Rayleigh(): lines from the ground level of H and He */
        for (kr = 0; kr < atom->Nline; kr++) {
          AtomicLine *line = &atom->line[kr];
          if (line->i == 0) {
            dv_push(&ray, m); dv_push(&ray, line->lambda0); dv_push(&ray, line->qwing); dv_push(&ray, line->Aji);
            dv_push(&ray, atom->g[line->j]); dv_push(&ray, atom->g[0]); dv_push(&ray, l0); dv_push(&ray, atom->stage[0]);
          }
        }
      for (kr = 0; kr < atom->Nline; kr++) {        /* rlk_opacity's duplicate check covers every model atom (kurucz.c:617-633) */
        AtomicLine *line = &atom->line[kr];
        if (elem_row[atom->periodic_table] >= 0) {
          dv_push(&ml, elem_row[atom->periodic_table]); dv_push(&ml, atom->stage[line->i]); dv_push(&ml, line->lambda0); dv_push(&ml, line->qwing);
        }
        if (!atom->active) {                        /* passive_bb skips ACTIVE atoms (metal.c:237) */
          double r[NPL];
          line_row(atom, m, l0, line, cs.n, r);
          for (i = 0; i < NPL; i++) dv_push(&pl, r[i]);
          for (i = 0; i < line->Ncomponent; i++) { dv_push(&cs, line->c_shift[i]); dv_push(&cf, line->c_fraction[i]); }
        }
      }
      l0 += atom->Nlevel;
    }
    T.lev0[atmos.Natom] = T.nlev = l0;
    trace("passive", pl.v, pl.n * sizeof(double)); trace("pcs", cs.v, cs.n * sizeof(double)); trace("pcf", cf.v, cs.n * sizeof(double)); trace("modellines", ml.v, ml.n * sizeof(double));
    if (input.allow_passive_bb) CHECK(rhb200_set_passive_lines(g_ctx, pl.n / NPL, pl.v, cs.n, cs.v, cf.v));
    else CHECK(rhb200_set_passive_lines(g_ctx, 0, NULL, 0, NULL, NULL));
    CHECK(rhb200_set_model_lines(g_ctx, ml.n / 4, ml.v));
    {                                               /* MolecularOpacity (opacity.c:711-839): line lists of PASSIVE molecules */
      dvec mrows = {0}, msel = {0}, mzs = {0}, mzt = {0};
      ivec mzq = {0};
      int nsel = 0;
      for (n = 0; n < atmos.Nmolecule; n++) {
        Molecule *mo = &atmos.molecules[n];
        if (mo->Nrt <= 0) continue;
        if (mo->Npf > 8) FAIL("molecule with more than 8 partition-function coefficients");
        for (kr = 0; kr < mo->Nrt; kr++) {            /* readMolecule() left them sorted by lambda0 (readmolecule.c:247) */
          MolecularLine *mrt = &mo->mrt[kr];
          double r[RHB200_ML_NFIELD];
          memset(r, 0, sizeof r);
          if (mrt->polarizable) {                     /* the reference's own MolZeeman() pattern (opacity.c:796 builds it lazily) */
            if (mrt->zm == NULL) mrt->zm = MolZeeman(mrt);
            r[RHB200_ML_POLARIZABLE] = 1.0; r[RHB200_ML_ZOFF] = mzq.n; r[RHB200_ML_NCOMP] = mrt->zm->Ncomponent;
            for (i = 0; i < mrt->zm->Ncomponent; i++) { iv_push(&mzq, mrt->zm->q[i]); dv_push(&mzs, mrt->zm->shift[i]); dv_push(&mzt, mrt->zm->strength[i]); }
          }
          r[RHB200_ML_LAMBDA0] = mrt->lambda0; r[RHB200_ML_EI] = mrt->Ei; r[RHB200_ML_GI] = mrt->gi; r[RHB200_ML_BIJ] = mrt->Bij;
          r[RHB200_ML_AJI] = mrt->Aji; r[RHB200_ML_BJI] = mrt->Bji; r[RHB200_ML_ISO_FRAC] = mrt->isotope_frac;
          r[RHB200_ML_QWING] = mrt->qwing; r[RHB200_ML_MOL] = nsel;
          for (i = 0; i < RHB200_ML_NFIELD; i++) dv_push(&mrows, r[i]);
        }
        {
          double r[16];
          memset(r, 0, sizeof r);
          r[0] = n; r[1] = mo->weight; r[2] = (double) mo->fit; r[3] = mo->Tmin; r[4] = mo->Tmax; r[5] = mo->Npf;
          for (i = 0; i < mo->Npf; i++) r[6 + i] = mo->pf_coef[i];
          for (i = 0; i < 16; i++) dv_push(&msel, r[i]);
        }
        nsel++;
      }
      trace("mollines", mrows.v, mrows.n * sizeof(double)); trace("molsel", msel.v, msel.n * sizeof(double)); trace("molzs", mzs.v, mzq.n * sizeof(double));
      CHECK(rhb200_set_molecular_lines_zeeman(g_ctx, mrows.n / RHB200_ML_NFIELD, mrows.v, nsel, msel.v, mzq.n, mzq.v, mzs.v, mzt.v));
      free(mrows.v); free(msel.v); free(mzq.v); free(mzs.v); free(mzt.v);
    }
    CHECK(rhb200_set_scatter(g_ctx, input.NmaxScatter, input.iterLimit));
    /* with ACTIVE atoms the polarised passes are selected through rhb200_nlte_front.stokes; the background is set up I-only */
    CHECK(rhb200_set_stokes_mode(g_ctx, input.StokesMode == FULL_STOKES && !any_active));
    /* -- wavelengths: spectrum.lambda as SortLambda left it */
    T.nlam = spectrum.Nspect;
    T.lam = (double *) realloc(T.lam, T.nlam * sizeof(double));
    memcpy(T.lam, spectrum.lambda, T.nlam * sizeof(double));
    T.iref = -1;
    for (n = 0; n < T.nlam; n++) if (T.lam[n] == atmos.lambda_ref) T.iref = n;
    if (T.iref < 0) FAIL("LAMBDA_REF = 0: convertScales needs the reference wavelength");
    trace("lambda", T.lam, T.nlam * sizeof(double));
    CHECK(rhb200_set_wavelengths(g_ctx, T.nlam, T.lam));
    CHECK(rhb200_set_solvers(g_ctx, (int) input.S_interpolation, (int) input.S_interpolation_stokes));
    /* -- continuum model */
    load_tables();
    memset(&M, 0, sizeof M);
    M.natom = atmos.Natom; M.nlev = T.nlev; M.ncont = bf.n / 10; M.ntab = tl.n; M.nray = ray.n / 8;
    M.lev = lev.v; M.bf = bf.v; M.tab_lambda = tl.v; M.tab_alpha = ta.v; M.ray = ray.v;
    M.nlev_H = atmos.atoms[0].Nlevel;
    M.atom_He = (atmos.Natom > 1 && atmos.elements[1].model == &atmos.atoms[1]) ? 1 : -1;
    M.H_active = atmos.atoms[0].active ? 1 : 0; M.solve_NLTE = any_active;
    M.vmicro_char = atmos.vmicro_char;
    for (n = 0; n < atmos.Nmolecule; n++) {
      if (!strcmp(atmos.molecules[n].ID, "OH")) M.has_OH = 1;
      if (!strcmp(atmos.molecules[n].ID, "CH")) M.has_CH = 1;
      if (!strcmp(atmos.molecules[n].ID, "H2")) M.has_H2 = 1;
    }
    M.hmbf_lambda = tab("hmbf_lambda", &M.n_hmbf); M.hmbf_alpha = tab("hmbf_alpha", NULL);
    M.hmff_lambda = tab("hmff_lambda", &M.n_hmff_lambda); M.hmff_theta = tab("hmff_theta", &M.n_hmff_theta); M.hmff_kappa = tab("hmff_kappa", NULL);
    M.h2mff_lambda = tab("h2mff_lambda", &M.n_h2mff_lambda); M.h2mff_theta = tab("h2mff_theta", &M.n_h2mff_theta); M.h2mff_kappa = tab("h2mff_kappa", NULL);
    M.h2pff_lambda = tab("h2pff_lambda", &M.n_h2pff_lambda); M.h2pff_temp = tab("h2pff_temp", &M.n_h2pff_temp); M.h2pff_kappa = tab("h2pff_kappa", NULL);
    M.rh2_a = tab("rh2_a", NULL); M.rh2_lambda = tab("rh2_lambda", &M.n_rh2); M.rh2_sigma = tab("rh2_sigma", NULL);
    M.oh_T = tab("oh_T", &M.n_oh_T); M.oh_E = tab("oh_E", &M.n_oh_E); M.oh_cross = tab("oh_cross", NULL);
    M.ch_T = tab("ch_T", &M.n_ch_T); M.ch_E = tab("ch_E", &M.n_ch_E); M.ch_cross = tab("ch_cross", NULL);
    if (fudge_lam != NULL) { M.do_fudge = 1; M.n_fudge = fudge_num; M.fudge_lambda = fudge_lam; M.fudge = fudge; }
    trace("lev", lev.v, lev.n * sizeof(double)); trace("bf", bf.v, bf.n * sizeof(double)); trace("tab_lambda", tl.v, tl.n * sizeof(double)); trace("tab_alpha", ta.v, ta.n * sizeof(double));
    trace("ray", ray.v, ray.n * sizeof(double)); trace("abund", ab.v, ab.n * sizeof(double)); trace("vmicro", &M.vmicro_char, sizeof(double));
    CHECK(rhb200_set_continuum(g_ctx, &M, ab.v));
    keep[nkeep++] = lev.v; keep[nkeep++] = bf.v; keep[nkeep++] = tl.v; keep[nkeep++] = ta.v; keep[nkeep++] = ray.v;
    free(pl.v); free(cs.v); free(cf.v); free(ml.v); free(ab.v);
  }
  /* -- chemical network (chemequil.c:130-170): nuclei = elements bound in molecules, in periodic-table order */
  {
    int *nuc_elem = (int *) malloc(atmos.Nelem * sizeof(int)), *nuc_atom, nnuc = 0;
    double *mol = (double *) calloc((size_t) atmos.Nmolecule * 32, sizeof(double));
    for (n = 0; n < atmos.Nelem; n++) if (atmos.elements[n].Nmolecule > 0) nuc_elem[nnuc++] = n;
    nuc_atom = (int *) malloc((nnuc + 1) * sizeof(int));
    for (n = 0; n < nnuc; n++) {
      Atom *model = atmos.elements[nuc_elem[n]].model;
      if (!model) FAIL("a nucleus bound in molecules has no model atom (getfjk path, chemequil.c:222-228)");
      nuc_atom[n] = (int) (model - atmos.atoms);
    }
    for (n = 0; n < atmos.Nmolecule; n++) {
      Molecule *mo = &atmos.molecules[n];
      double *r = mol + (size_t) n * 32;
      if (mo->active) FAIL("ACTIVE molecules are not implemented");
      if (mo->Neqc > 8 || mo->Nelement > 4) FAIL("molecule with more than 8 equilibrium coefficients or 4 constituents");
      r[0] = (double) mo->fit; r[1] = mo->charge; r[2] = mo->Nnuclei; r[3] = mo->Nelement; r[4] = mo->Neqc;
      r[5] = mo->Tmin; r[6] = mo->Tmax; r[7] = mo->Ediss;
      for (i = 0; i < mo->Neqc; i++) r[8 + i] = mo->eqc_coef[i];
      for (i = 0; i < mo->Nelement; i++) {
        for (k = 0; k < nnuc; k++) if (nuc_elem[k] == mo->pt_index[i]) r[16 + i] = k;
        r[20 + i] = mo->pt_count[i];
      }
      r[24] = !strcmp(mo->ID, "H2"); r[25] = !strcmp(mo->ID, "OH"); r[26] = !strcmp(mo->ID, "CH");
    }
    T.nmol = atmos.Nmolecule;
    trace("chem_mol", mol, (size_t) atmos.Nmolecule * 32 * sizeof(double)); trace("nuc_atom", nuc_atom, nnuc * sizeof(int));
    CHECK(rhb200_set_chemistry(g_ctx, nnuc, nuc_atom, atmos.Nmolecule, mol));
    free(nuc_elem); free(nuc_atom); free(mol);
  }
  free(lines.v); free(zs.v); free(zt.v); free(zq.v); free(elems.v); free(pf.v); free(elem_row);
}

/* ---- NLTE: spectrum.as[] / atoms -> rhb200_nlte_plan + rhb200_nlte_front, solve, fill atom->n / atom->nstar */
static void spline_coef(int N, const double *x, const double *y, double *M)      /* splineCoef, spline.c:31-66 */
{
  double *q = (double *) malloc(N * sizeof(double)), *u = (double *) malloc(N * sizeof(double));
  double hj = x[1] - x[0], D = (y[1] - y[0]) / hj, hj1, mu, D1, p;
  int j;
  q[0] = u[0] = 0.0;
  for (j = 1; j < N-1; j++) {
    hj1 = x[j+1] - x[j];
    mu = hj / (hj + hj1);
    D1 = (y[j+1] - y[j]) / hj1;
    p = mu*q[j-1] + 2;
    q[j] = (mu - 1) / p;
    u[j] = ((D1 - D) * 6/(hj + hj1) - mu*u[j-1]) / p;
    hj = hj1; D = D1;
  }
  M[N-1] = 0.0;
  for (j = N-2; j >= 0; j--) M[j] = q[j]*M[j+1] + u[j];
  free(q); free(u);
}

static void collisions(Atom *atom, int a, dvec *rows, dvec *tT, dvec *tC, dvec *tM)
{
  char line[MAX_LINE_SIZE], key[MAX_LINE_SIZE], *tok;
  double *Tg = NULL, coef[64];
  int nT = 0, n;
  fseek(atom->fp_input, atom->offset_coll, SEEK_SET);
  while (getLine(atom->fp_input, "#", line, FALSE) != EOF) {
    int type = -1, i1, i2, i, j;
    if (!(tok = strtok(line, " "))) continue;
    strcpy(key, tok);
    if (!strcmp(key, "TEMP")) {
      nT = atoi(strtok(NULL, " "));
      if (nT > 64) FAIL("collisional temperature grid with more than 64 points");
      Tg = (double *) realloc(Tg, nT * sizeof(double));
      for (n = 0; n < nT; n++) { if (!(tok = strtok(NULL, " "))) FAIL("short TEMP record"); sscanf(tok, "%lf", Tg + n); }
      continue;
    }
    if (strstr(key, "END")) break;
    if (!strcmp(key, "OMEGA")) type = RHB200_CO_OMEGA; else if (!strcmp(key, "CE")) type = RHB200_CO_CE;
    else if (!strcmp(key, "CI")) type = RHB200_CO_CI; else if (!strcmp(key, "CP")) type = RHB200_CO_CP;
    else if (!strcmp(key, "CH")) type = RHB200_CO_CH; else if (!strcmp(key, "CH0")) type = RHB200_CO_CH0;
    else if (!strcmp(key, "CH+")) type = RHB200_CO_CHPLUS;
    else { sprintf(messageStr, "collision keyword %s is not ported (collision.c:516-936)", key); Error(ERROR_LEVEL_2, "pyrh_b200", messageStr); }
    if (!Tg) FAIL("collision record before any TEMP record");
    i1 = atoi(strtok(NULL, " ")); i2 = atoi(strtok(NULL, " "));
    for (n = 0; n < nT; n++) { if (!(tok = strtok(NULL, " "))) FAIL("short collision record"); sscanf(tok, "%lf", coef + n); }
    i = MIN(i1, i2); j = MAX(i1, i2);
    {
      double r[RHB200_CO_NFIELD], M[64];
      memset(r, 0, sizeof r); memset(M, 0, sizeof M);
      r[RHB200_CO_ATOM] = a; r[RHB200_CO_TYPE] = type; r[RHB200_CO_I] = i; r[RHB200_CO_J] = j; r[RHB200_CO_NT] = nT;
      r[RHB200_CO_TOFF] = tT->n; r[RHB200_CO_DE] = atom->E[j] - atom->E[i];
      r[7] = (type == RHB200_CO_OMEGA) ? atom->g[j] : (type == RHB200_CO_CE) ? atom->g[i]/atom->g[j] : 0.0;
      if (nT > 2) spline_coef(nT, Tg, coef, M);
      for (n = 0; n < RHB200_CO_NFIELD; n++) dv_push(rows, r[n]);
      for (n = 0; n < nT; n++) { dv_push(tT, Tg[n]); dv_push(tC, coef[n]); dv_push(tM, M[n]); }
    }
  }
  fseek(atom->fp_input, atom->offset_coll, SEEK_SET);
  free(Tg);
}

static void solve_nlte(double mu, int ncol, const double *rows9, int nrow, double *spec_out, double *quv_out, double *n_out, double *ns_out, int *niter)
{
  const int Ns = spectrum.Nspect, Na = atmos.Nactiveatom, N = atmos.Nspace;
  rhb200_nlte_plan P, P1;
  rhb200_nlte_front F;
  dvec trans = {0}, trans1 = {0}, wl = {0}, ww = {0}, wa = {0}, co = {0}, tT = {0}, tC = {0}, tM = {0}, lr = {0};
  ivec asf = {0}, ast = {0}, lpol = {0}, lzoff = {0}, zq = {0}, lprd = {0};
  dvec zs = {0}, zt = {0};
  const int field_free = input.StokesMode != NO_STOKES;                   /* polarised passes needed */
  int *nlevel = (int *) malloc(Na * sizeof(int)), *model = (int *) malloc(Na * sizeof(int)), *hasline = (int *) malloc(Ns * sizeof(int));
  int **lidx = (int **) malloc(Na * sizeof(int *)), **cidx = (int **) malloc(Na * sizeof(int *));
  int a, kr, la, ns, n, phirow = 0, phirow1 = 0, nline = 0, ntr = 0;
  double mu1 = mu, w1 = 1.0;
  iv_push(&lzoff, 0);
  if (!atmos.moving) FAIL("static atmospheres with ACTIVE atoms are not implemented");
  for (a = 0; a < Na; a++) {
    Atom *atom = atmos.activeatoms[a];
    nlevel[a] = atom->Nlevel; model[a] = (int) (atom - atmos.atoms);
    if (atom->initial_solution != LTE_POPULATIONS) FAIL("initial solutions other than LTE_POPULATIONS are not implemented");
    if (atom->Nfixed > 0) FAIL("fixed transitions are not implemented");
    lidx[a] = (int *) malloc((atom->Nline + 1) * sizeof(int)); cidx[a] = (int *) malloc((atom->Ncont + 1) * sizeof(int));
    for (kr = 0; kr < atom->Nline; kr++) {                 /* device order: an atom's lines, then its continua */
      AtomicLine *L = &atom->line[kr];
      double r[RHB200_TR_NFIELD], pr[NPL];
      if (L->PRD && (input.PRD_angle_dep || input.XRD)) FAIL("PRD_ANGLE_DEP / XRD (angle-dependent and cross redistribution) are not implemented");
      if (L->PRD && input.PRD_Ngorder > 0) FAIL("PRD_NG_ORDER > 0 is not implemented");
      iv_push(&lprd, L->PRD ? 1 : 0);
      if (L->Ncomponent > 1) FAIL("multi-component ACTIVE lines are not implemented");
      memset(r, 0, sizeof r);
      r[RHB200_TR_ATOM] = a; r[RHB200_TR_TYPE] = 0; r[RHB200_TR_I] = L->i; r[RHB200_TR_J] = L->j; r[RHB200_TR_NBLUE] = L->Nblue;
      r[RHB200_TR_NLAMBDA] = L->Nlambda; r[RHB200_TR_AJI] = L->Aji; r[RHB200_TR_BJI] = L->Bji; r[RHB200_TR_BIJ] = L->Bij;
      r[RHB200_TR_ISOFRAC] = L->isotope_frac; r[RHB200_TR_WOFF] = wl.n; r[RHB200_TR_KR] = kr; r[RHB200_TR_LINEIDX] = nline;
      r[RHB200_TR_LAMBDA0] = L->lambda0;
      r[RHB200_TR_PHIROW] = phirow;
      for (n = 0; n < RHB200_TR_NFIELD; n++) dv_push(&trans, r[n]);
      r[RHB200_TR_PHIROW] = phirow1;
      for (n = 0; n < RHB200_TR_NFIELD; n++) dv_push(&trans1, r[n]);
      for (la = 0; la < L->Nlambda; la++) { dv_push(&wl, L->lambda[la]); dv_push(&ww, getwlambda_line(L, la)); dv_push(&wa, 0.0); }
      phirow += 2 * atmos.Nrays * L->Nlambda; phirow1 += 2 * L->Nlambda;
      line_row(atom, model[a], T.lev0[model[a]], L, 0, pr);
      for (n = 0; n < NPL; n++) dv_push(&lr, pr[n]);
      if (field_free && L->polarizable) {                  /* Zeeman(line), zeeman.c:186-281: what Profile() uses (profile.c:115) */
        ZeemanMultiplet *zm = Zeeman(L);
        for (n = 0; n < zm->Ncomponent; n++) { iv_push(&zq, zm->q[n]); dv_push(&zs, zm->shift[n]); dv_push(&zt, zm->strength[n]); }
        freeZeeman(zm); free(zm);
        iv_push(&lpol, 1);
      } else iv_push(&lpol, 0);
      iv_push(&lzoff, zq.n);
      lidx[a][kr] = ntr++; nline++;
    }
    for (kr = 0; kr < atom->Ncont; kr++) {
      AtomicContinuum *Cn = &atom->continuum[kr];
      double r[RHB200_TR_NFIELD];
      memset(r, 0, sizeof r);
      r[RHB200_TR_ATOM] = a; r[RHB200_TR_TYPE] = 1; r[RHB200_TR_I] = Cn->i; r[RHB200_TR_J] = Cn->j; r[RHB200_TR_NBLUE] = Cn->Nblue;
      r[RHB200_TR_NLAMBDA] = Cn->Nlambda; r[RHB200_TR_WOFF] = wl.n; r[RHB200_TR_PHIROW] = -1; r[RHB200_TR_KR] = kr; r[RHB200_TR_LINEIDX] = -1;
      for (n = 0; n < RHB200_TR_NFIELD; n++) { dv_push(&trans, r[n]); dv_push(&trans1, r[n]); }
      for (la = 0; la < Cn->Nlambda; la++) { dv_push(&wl, Cn->lambda[la]); dv_push(&ww, getwlambda_cont(Cn, la)); dv_push(&wa, Cn->alpha[la]); }
      cidx[a][kr] = ntr++;
    }
    collisions(atom, a, &co, &tT, &tC, &tM);
  }
  for (ns = 0; ns < Ns; ns++) {                             /* active sets in SortLambda's order */
    ActiveSet *as = &spectrum.as[ns];
    iv_push(&asf, ast.n);
    for (a = 0; a < Na; a++)
      for (n = 0; n < as->Nactiveatomrt[a]; n++) {
        AtomicTransition *t = &as->art[a][n];
        Atom *atom = atmos.activeatoms[a];
        iv_push(&ast, t->type == ATOMIC_LINE ? lidx[a][t->ptype.line - atom->line] : cidx[a][t->ptype.continuum - atom->continuum]);
      }
  }
  iv_push(&asf, ast.n);
  CHECK(rhb200_get_wavelength_flags(g_ctx, hasline));
  for (ns = 0; ns < Ns; ns++) hasline[ns] &= 1;
  memset(&P, 0, sizeof P);
  P.Nspect = Ns; P.Nrays = atmos.Nrays; P.Ndep = N; P.Natom = Na; P.Ntrans = ntr; P.moving = 1;
  P.Ngorder = input.Ngorder; P.Ngdelay = input.Ngdelay; P.Ngperiod = input.Ngperiod; P.isum = input.isum;
  P.bc_top = RHB200_BC_ZERO; P.bc_bottom = RHB200_BC_THERMALIZED;
  if (geometry.vboundary[TOP] != ZERO) FAIL("irradiated top boundary is not implemented");
  P.ntrl = wl.n; P.nphirow = phirow; P.nline = nline;
  P.lambda = spectrum.lambda; P.muz = geometry.muz; P.wmu = geometry.wmu; P.atom_nlevel = nlevel; P.trans = trans.v;
  P.tr_lambda = wl.v; P.tr_wlambda = ww.v; P.tr_alpha = wa.v; P.as_first = asf.v; P.as_trans = ast.v; P.bg_hasline = hasline;
  P1 = P; P1.Nrays = 1; P1.muz = &mu1; P1.wmu = &w1; P1.trans = trans1.v; P1.nphirow = phirow1;
  memset(&F, 0, sizeof F);
  F.atom_model = model; F.ncoll = co.n / RHB200_CO_NFIELD; F.ncolltab = tT.n; F.coll = co.v; F.coll_T = tT.v; F.coll_coef = tC.v;
  F.coll_M = tM.v; F.line_rows = lr.v; F.NmaxScatter = input.NmaxScatter; F.NmaxIter = input.NmaxIter; F.iterLimit = input.iterLimit;
  F.plan1 = &P1;
  if (zq.n == 0) { iv_push(&zq, 0); dv_push(&zs, 0.0); dv_push(&zt, 0.0); }
  F.line_prd = lprd.v; F.PRD_NmaxIter = input.PRD_NmaxIter; F.PRDiterLimit = input.PRDiterLimit;
  F.stokes = input.StokesMode == FULL_STOKES ? 2 : (input.StokesMode == FIELD_FREE ? 1 : (input.StokesMode == POLARIZATION_FREE ? 3 : 0)); F.line_pol = lpol.v; F.line_zoff = lzoff.v; F.zq = zq.v; F.zshift = zs.v; F.zstrength = zt.v;
  CHECK(rhb200_nlte_compute1d_stokes_batch(g_ctx, &P, &F, ncol, N, nrow, mu, g_atm_scale, rows9, T.iref, atmos.wght_per_H,
                                           atmos.vmacro_tresh, spec_out, quv_out, n_out, ns_out, niter, NULL, NULL));
  free(lprd.v); free(lpol.v); free(lzoff.v); free(zq.v); free(zs.v); free(zt.v);
  for (a = 0; a < Na; a++) { free(lidx[a]); free(cidx[a]); }
  free(lidx); free(cidx); free(nlevel); free(model); free(hasline);
  free(trans.v); free(trans1.v); free(wl.v); free(ww.v); free(wa.v); free(co.v); free(tT.v); free(tC.v); free(tM.v); free(lr.v);
  free(asf.v); free(ast.v);
}

/* ---- the hook the patched rhf1d() calls after getBoundary() (pyrh_compute1dray.c:308): returns 1 with `spec` filled */
int pyrh_b200_solve(double mu, int get_atomic_rfs, int get_populations, int fudge_num, double *fudge_lam, double *fudge,
                    mySpectrum *spec)
{
  const int N = atmos.Nspace, Ns = spectrum.Nspect, Nlw = Ns - 1;
  const int ncol = g_batch.ncol > 0 ? g_batch.ncol : 1;
  const double *rows = g_batch.ncol > 0 ? g_batch.atm : g_rows;
  int n, index, a;
  memset(spec, 0, sizeof *spec);
  if (atmos.hydrostatic) FAIL("HYDROSTATIC = TRUE (Hydrostatic() inside Iterate()) is not implemented");
  if (input.backgr_pol) FAIL("BACKGROUND_POLARIZATION = TRUE (scattering polarisation through J20) is not implemented");
  build_tables(get_atomic_rfs, fudge_num, fudge_lam, fudge);
  spec->nlw = Nlw; spec->Nrays = atmos.Nrays; spec->stokes = 1;
  spec->lam = (double *) malloc(Nlw * sizeof(double));
  spec->sI = (double *) calloc(Nlw, sizeof(double)); spec->sQ = (double *) calloc(Nlw, sizeof(double));
  spec->sU = (double *) calloc(Nlw, sizeof(double)); spec->sV = (double *) calloc(Nlw, sizeof(double));
  if (input.solve_NLTE) {
    int nlev = 0;
    double *I = (double *) malloc((size_t) ncol * Ns * sizeof(double)), *pn, *ps;
    double *quv = (double *) calloc((size_t) ncol * 3 * Ns, sizeof(double));
    for (a = 0; a < atmos.Nactiveatom; a++) nlev += atmos.activeatoms[a]->Nlevel;
    pn = (double *) malloc((size_t) ncol * nlev * N * sizeof(double)); ps = (double *) malloc((size_t) ncol * nlev * N * sizeof(double));
    solve_nlte(mu, ncol, rows, 9, I, quv, pn, ps, g_batch.niter);
    for (n = 0, index = 0; n < Ns; n++)
      if (spectrum.lambda[n] != atmos.lambda_ref) {
        spec->lam[index] = spectrum.lambda[n]; spec->sI[index] = I[n];
        spec->sQ[index] = quv[n]; spec->sU[index] = quv[Ns + n]; spec->sV[index] = quv[2*(size_t) Ns + n];
        index++;
      }
    if (g_batch.ncol > 0) {
      int c;
      for (c = 0; c < ncol; c++)
        for (n = 0, index = 0; n < Ns; n++)
          if (spectrum.lambda[n] != atmos.lambda_ref) {
            int q;
            g_batch.stokes[((size_t) c * 4) * Nlw + index] = I[(size_t) c * Ns + n];
            for (q = 0; q < 3; q++) g_batch.stokes[((size_t) c * 4 + q + 1) * Nlw + index] = quv[((size_t) c * 3 + q) * Ns + n];
            index++;
          }
      if (g_batch.pops_n) memcpy(g_batch.pops_n, pn, (size_t) ncol * nlev * N * sizeof(double));
      if (g_batch.pops_nstar) memcpy(g_batch.pops_nstar, ps, (size_t) ncol * nlev * N * sizeof(double));
    }
    {                                                /* column 0 -> the live atom->n / atom->nstar, like the reference */
      int l0 = 0;
      for (a = 0; a < atmos.Nactiveatom; a++) {
        Atom *atom = atmos.activeatoms[a];
        memcpy(atom->n[0], pn + (size_t) l0 * N, (size_t) atom->Nlevel * N * sizeof(double));
        memcpy(atom->nstar[0], ps + (size_t) l0 * N, (size_t) atom->Nlevel * N * sizeof(double));
        l0 += atom->Nlevel;
      }
    }
    free(I); free(quv); free(pn); free(ps);
  } else {
    const int nl = Ns, npar = T.lrf_npar;
    double *st = (double *) malloc((size_t) ncol * 4 * nl * sizeof(double)), *rf = NULL;
    const int bc_top = RHB200_BC_ZERO, bc_bot = RHB200_BC_THERMALIZED;
    if (geometry.vboundary[TOP] != ZERO) FAIL("irradiated top boundary is not implemented");
    if (input.get_atomic_rfs && npar > 0) {
      rf = (double *) calloc((size_t) ncol * nl * npar, sizeof(double));
      CHECK(rhb200_compute1d_rf_batch(g_ctx, ncol, N, 9, mu, g_atm_scale, rows, T.iref, atmos.wght_per_H, atmos.vmacro_tresh,
                                      bc_top, bc_bot, st, NULL, rf));
    } else
      trace("rows", rows, (size_t) ncol * 9 * N * sizeof(double)); trace("wght_per_H", &atmos.wght_per_H, sizeof(double)); trace("vmacro_tresh", &atmos.vmacro_tresh, sizeof(double));
      CHECK(rhb200_compute1d_batch(g_ctx, ncol, N, 9, mu, g_atm_scale, rows, T.iref, atmos.wght_per_H, atmos.vmacro_tresh,
                                   bc_top, bc_bot, st, NULL));
    if (input.get_atomic_rfs) spec->rfs = matrix_double(Nlw, input.n_atomic_pars);
    for (n = 0, index = 0; n < Ns; n++) {
      if (spectrum.lambda[n] == atmos.lambda_ref) continue;
      spec->lam[index] = spectrum.lambda[n];
      spec->sI[index] = st[n]; spec->sQ[index] = st[nl + n]; spec->sU[index] = st[2*nl + n]; spec->sV[index] = st[3*nl + n];
      if (rf) { int p; for (p = 0; p < npar; p++) spec->rfs[index][p] = rf[(size_t) n * npar + p]; }
      index++;
    }
    if (g_batch.ncol > 0) {
      int c, q;
      for (c = 0; c < ncol; c++)
        for (q = 0; q < 4; q++)
          for (n = 0, index = 0; n < Ns; n++)
            if (spectrum.lambda[n] != atmos.lambda_ref) g_batch.stokes[((size_t) c * 4 + q) * Nlw + index++] = st[((size_t) c * 4 + q) * nl + n];
    }
    free(st); free(rf);
  }
  if (get_populations) {                             /* pyrh_solveray.c:171-185 */
    spec->Nactive_atoms = atmos.Nactiveatom;
    spec->atom_pops = (AtomPops *) malloc((atmos.Nactiveatom + 1) * sizeof(AtomPops));
    for (a = 0; a < atmos.Nactiveatom; a++) {
      Atom *atom = atmos.activeatoms[a];
      strcpy(spec->atom_pops[a].ID, atom->ID);
      spec->atom_pops[a].Nlevel = atom->Nlevel; spec->atom_pops[a].Nz = N;
      spec->atom_pops[a].n = atom->n; spec->atom_pops[a].nstar = atom->nstar;
    }
  }
  /* what _solveray() releases on the reference's path (pyrh_solveray.c:153-166) */
  if (spectrum.lambda != NULL) { free(spectrum.lambda); spectrum.lambda = NULL; }
  return 1;
}

/* ---- non-breaking addition: many columns through one call.  atmosphere [ncol][9][Ndep] in pyrh units (the rows of
   pyrh.compute1d); stokes [ncol][4][nlw] with nlw = the number of wavelengths rhf1d() returns (spectrum.lambda minus
   lambda_ref); pops_n / pops_nstar [ncol][sum Nlevel][Ndep] and niter [ncol] may be NULL.  Returns the mySpectrum of
   column 0 (its lam member is the wavelength axis of `stokes`). */
mySpectrum rhf1d_batch(char *cwd, double mu, int Ndep, int ncol, double *atmosphere, int atm_scale, int Nwave, double *lam,
                       int fudge_num, double *fudge_lam, double *fudge, int Nloggf, int *loggf_ids, double *loggf_values,
                       int Nlam, int *lam_ids, double *lam_values, int Nabun, int *atomic_id, double *atomic_abundance,
                       double *stokes, double *pops_n, double *pops_nstar, int *niter)
{
  mySpectrum spec;
  double *col0 = (double *) malloc((size_t) 9 * Ndep * sizeof(double));
  memcpy(col0, atmosphere, (size_t) 9 * Ndep * sizeof(double));        /* rhf1d() converts its arguments in place */
  g_batch.ncol = ncol; g_batch.atm = atmosphere; g_batch.stokes = stokes; g_batch.pops_n = pops_n; g_batch.pops_nstar = pops_nstar;
  g_batch.niter = niter;
  spec = rhf1d(cwd, mu, Ndep, col0, col0 + Ndep, col0 + 2*Ndep, col0 + 3*Ndep, col0 + 4*Ndep, col0 + 5*Ndep, col0 + 6*Ndep,
               col0 + 7*Ndep, col0 + 8*Ndep, atm_scale, Nwave, lam, fudge_num, fudge_lam, fudge, Nloggf, loggf_ids, loggf_values,
               Nlam, lam_ids, lam_values, Nabun, atomic_id, atomic_abundance, 0, pops_n != NULL, 0, cwd);
  memset(&g_batch, 0, sizeof g_batch);
  free(col0);
  return spec;
}

/* ---- pyrh_hse.c: hse(), get_scales(), get_ne_from_nH() */
static void set_all_elements(void)               /* atmos.elements[] as Solve_ne sees it (solvene.c:55-140) */
{
  dvec elems = {0}, pf = {0};
  int n, i, k, pfrow = 0;
  if (!g_ctx && !(g_ctx = rhb200_open(getenv("RHB200_DEVICE") ? atoi(getenv("RHB200_DEVICE")) : 0))) FAIL(rhb200_last_error());
  for (n = 0; n < atmos.Nelem; n++) {
    Element *el = &atmos.elements[n];
    double r[RHB200_RE_NFIELD];
    if (!el->abundance_set) FAIL("element without an abundance (abundance.c:207-215 feeds raw partition functions into Solve_ne)");
    if (el->Nstage > RHB200_RE_MAXSTAGE) FAIL("element with more ionisation stages than RHB200_RE_MAXSTAGE");
    memset(r, 0, sizeof r);
    r[RHB200_RE_WEIGHT] = el->weight; r[RHB200_RE_ABUND] = el->abund; r[RHB200_RE_NSTAGE] = el->Nstage; r[RHB200_RE_PFROW] = pfrow;
    for (i = 0; i < el->Nstage; i++) r[RHB200_RE_IONPOT0 + i] = el->ionpot[i];
    for (i = 0; i < RHB200_RE_NFIELD; i++) dv_push(&elems, r[i]);
    for (i = 0; i < el->Nstage; i++) for (k = 0; k < atmos.Npf; k++) dv_push(&pf, el->pf[i][k]);
    pfrow += el->Nstage;
  }
  CHECK(rhb200_set_elements(g_ctx, atmos.Nelem, elems.v, pfrow, atmos.Npf, pf.v, atmos.Tpf));
  free(elems.v); free(pf.v);
}

/* hse() after SortLambda({500 nm}) (patch, pyrh_hse.c:200): the layer-by-layer walk of :212-367 for this column.
   atmos.T / ne / nHtot alias the caller's arrays (:168-171), pg[0] holds the top pressure [Pa]. */
int pyrh_b200_hse(int Ndep, int atm_scale, double *scale, double *rho, double *pg, int fudge_num, double *fudge_lam, double *fudge)
{
  double pg_top = pg[0];
  if (atm_scale == 1) FAIL("hse() on a column-mass scale: the reference integrates no pressure there (pyrh_hse.c:283-288)");
  g_no_kurucz = 1;
  build_tables(0, fudge_num, fudge_lam, fudge);
  g_no_kurucz = 0;
  set_all_elements();
  CHECK(rhb200_hse_batch(g_ctx, 1, Ndep, atm_scale, scale, atmos.T, &pg_top, atmos.wght_per_H, atmos.totalAbund, atmos.gravity,
                         atmos.ne, atmos.nHtot, rho, pg));
  return 1;
}

/* get_scales() after getBoundary() (patch, pyrh_hse.c:508): Background() at lam_ref + convertScales() (:513-514).
   The rows in pyrh units were saved by pyrh_b200_save_inputs() before the in-place conversion (:479-485); the
   results go where the reference's geometry pointers point (:445-462). */
int pyrh_b200_get_scales(int Ndep)
{
  double *sc = (double *) malloc((size_t) 3 * Ndep * sizeof(double));
  g_no_kurucz = 1;
  build_tables(0, 0, NULL, NULL);
  g_no_kurucz = 0;
  CHECK(rhb200_get_scales_batch(g_ctx, 1, Ndep, 9, g_atm_scale, g_rows, T.iref, atmos.wght_per_H, atmos.totalAbund, atmos.gravity,
                                atmos.vmacro_tresh, sc));
  memcpy(geometry.height, sc, Ndep * sizeof(double));
  memcpy(geometry.tau_ref, sc + Ndep, Ndep * sizeof(double));
  memcpy(geometry.cmass, sc + 2 * (size_t) Ndep, Ndep * sizeof(double));
  free(sc);
  return 1;
}

/* get_ne_from_nH() in place of Background(FALSE, TRUE) with SOLVE_NE = ONCE (patch, pyrh_hse.c:644-645):
   atmos.T [K], atmos.nHtot [m^-3] -> atmos.ne [m^-3] */
int pyrh_b200_solve_ne(int Ndep)
{
  set_all_elements();
  CHECK(rhb200_solve_ne_batch(g_ctx, (size_t) Ndep, atmos.T, atmos.nHtot, atmos.ne, 1));
  return 1;
}

void pyrh_b200_close(void)
{
  if (g_ctx) { rhb200_close(g_ctx); g_ctx = NULL; }
}
