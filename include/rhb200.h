/* include/rhb200.h -- C ABI of librhb200.so, the B200 (sm_100a) hot path of RH / pyrh.
 *
 * Plain C: pointers and sizes only.  This is the boundary the RH host library
 * (reference: /root/reference/rh) binds to; every entry point names the
 * reference interface it replaces (file:line relative to the reference root).
 * The reference-side stubs a maintainer would add are shown in INTEGRATION.md.
 *
 * Conventions
 *   - all floating point data is IEEE double, SI units, exactly the values the
 *     reference holds in `atmos`, `geometry`, `spectrum` when Formal() starts
 *     (i.e. after the in-place conversion in pyrh_compute1dray.c:263-270);
 *   - "ray" = one (column, wavelength) pair at the single LTE angle mu;
 *   - every function returns 0 on success, a negative RHB200_E* code otherwise,
 *     and rhb200_last_error() describes the failure.  The library never calls
 *     exit() (the reference's Error(ERROR_LEVEL_2) does: rh/error.c:45-60);
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with RHB200_ENODEV.
 */
#ifndef RHB200_H
#define RHB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RHB200_VERSION 100

enum {
  RHB200_OK = 0,
  RHB200_ENODEV = -1,   /* no CUDA device / driver */
  RHB200_ECUDA = -2,    /* CUDA runtime error (message in rhb200_last_error) */
  RHB200_EINVAL = -3,   /* bad argument */
  RHB200_ESTATE = -4,   /* tables / wavelengths not set */
  RHB200_ENOMEM = -5,
  RHB200_EUNSUPPORTED = -6   /* keyword combination outside the implemented path */
};

/* ---- line-table row layout: the numeric fields of RLK_Line, rh/atom.h:156-167,
        as filled by readKuruczLines (rh/kurucz.c:121-431) and sorted by
        qsort(rlk_ascend) (rh/background.c:292-294).  One row of doubles per line. */
enum {
  RHB200_RL_LAMBDA0 = 0, /* nm, vacuum */
  RHB200_RL_GI, RHB200_RL_GJ,
  RHB200_RL_EI, RHB200_RL_EJ,                 /* J */
  RHB200_RL_BJI, RHB200_RL_AJI, RHB200_RL_BIJ,
  RHB200_RL_GRAD, RHB200_RL_GSTARK, RHB200_RL_GVDW,
  RHB200_RL_HFS_FRAC, RHB200_RL_ISO_FRAC,
  RHB200_RL_CROSS, RHB200_RL_ALPHA,
  RHB200_RL_POLARIZABLE,                      /* 0/1 */
  RHB200_RL_VDWAALS,                          /* enum vdWaals, rh/atom.h:32 */
  RHB200_RL_ELEM,                             /* row index into the element table */
  RHB200_RL_STAGE,
  RHB200_RL_ZOFF, RHB200_RL_NCOMP,            /* slice of the Zeeman component arrays */
  RHB200_RL_NFIELD = 24
};
/* ---- element-table row layout: Element, rh/atom.h:139-145 */
enum {
  RHB200_RE_WEIGHT = 0, RHB200_RE_ABUND, RHB200_RE_NSTAGE,
  RHB200_RE_PFROW,                            /* first row of this element in pf[][] */
  RHB200_RE_IONPOT0,                          /* ionpot[0 .. NSTAGE-2], J */
  RHB200_RE_MAXSTAGE = 12,
  RHB200_RE_NFIELD = RHB200_RE_IONPOT0 + RHB200_RE_MAXSTAGE
};
/* ---- per-column atmosphere rows [RHB200_AT_NFIELD][Ndep] (atmos.h:55-79, geometry.h:19-25) */
enum {
  RHB200_AT_T = 0, RHB200_AT_NE, RHB200_AT_VTURB, RHB200_AT_VEL, RHB200_AT_B,
  RHB200_AT_COS_GAMMA, RHB200_AT_COS_2CHI, RHB200_AT_SIN_2CHI,   /* Bproject(), rhf1d/project.c:38 */
  RHB200_AT_NHTOT, RHB200_AT_NP,              /* np = atmos.H->n[Nlevel-1] (kurucz.c:772) */
  RHB200_AT_HEIGHT,
  RHB200_AT_NFIELD
};
enum { RHB200_BC_IRRADIATED = 0, RHB200_BC_ZERO = 1, RHB200_BC_THERMALIZED = 2 };  /* geometry.h:13 */

typedef struct rhb200_ctx rhb200_ctx;

/* ---- library / device ------------------------------------------------ */
int         rhb200_version(void);
const char *rhb200_last_error(void);
int         rhb200_device_count(void);
/* device name, SM count, memory; any pointer may be NULL */
int         rhb200_device_info(int device, char *name, int name_len, int *sm_count,
                               size_t *mem_bytes, int *cc_major, int *cc_minor);

/* ---- context: one per (process, GPU); owns a stream pair, the device copies
        of the shared tables and a reusable workspace.  Replaces the reference's
        process-global structs (pyrh_compute1dray.c:47-53). */
rhb200_ctx *rhb200_open(int device);
void        rhb200_close(rhb200_ctx *ctx);

/* Shared read-only tables.  Replaces atmos.rlk_lines / atmos.elements / atmos.Tpf
   as set up by readKuruczLines (kurucz.c:121), RLKZeeman (kurucz.c:832) and
   readAbundance + pf_Kurucz.input (abundance.c:72-203).
   vmicro_char: keyword VMICRO_CHAR in m/s.  magneto_optical, rlkscatter: keywords
   MAGNETO_OPTICAL, RLK_SCATTER (both must be 0: RHB200_EUNSUPPORTED otherwise). */
int rhb200_set_lines(rhb200_ctx *ctx,
                     int nline, const double *lines /* [nline][RHB200_RL_NFIELD] */,
                     int ncomp, const int *zq, const double *zshift, const double *zstrength,
                     int nelem, const double *elems /* [nelem][RHB200_RE_NFIELD] */,
                     int npf_rows, int npf, const double *pf /* [npf_rows][npf] */,
                     const double *Tpf /* [npf] */,
                     double vmicro_char, int magneto_optical, int rlkscatter);

/* Wavelengths [nm, vacuum] at which Stokes spectra are wanted (spectrum.lambda
   minus lambda_ref, pyrh_solveray.c:130-150).  Builds the per-wavelength line
   windows of rlk_opacity (kurucz.c:538-566,608) on the host: integer work,
   bit-exact. */
/* log gf overrides between calls (pyrh.compute1d's loggf_ids / loggf_values, kurucz.c:247-257) without rebuilding
   anything: rows[n] of the table passed to rhb200_set_lines get new Aji / Bji / Bij; windows, Zeeman patterns, element
   tables and the device context stay as they are. */
int rhb200_update_line_strengths(rhb200_ctx *ctx, int n, const int *rows, const double *Aji, const double *Bji, const double *Bij);

int rhb200_set_wavelengths(rhb200_ctx *ctx, int nlambda, const double *lambda);

/* MolecularOpacity in the fused LTE path (opacity.c:711-839, MolProfile :844-916): LTE lines of PASSIVE molecules,
   added to the background after the Kurucz lines like Background() does (background.c:548-566).  mlines [nline][RHB200_ML_NFIELD] grouped by molecule, ascending lambda0 inside each, RHB200_ML_MOL = row
   of `molecules`; molecules [nmol][16] = {index in the chemical network of rhb200_set_chemistry, molecular weight,
   enum fit_type, Tmin, Tmax, Npf, pf_coef[0..7]}.  Densities come from the chemistry kernel, partition functions
   (partfunction, chemequil.c:395-443) and Doppler widths are formed on the device.  Call after rhb200_set_continuum /
   rhb200_set_chemistry and before rhb200_set_wavelengths; rhb200_set_lines clears the table. */
int rhb200_set_molecular_lines(rhb200_ctx *ctx, int nline, const double *mlines, int nmol, const double *molecules);
/* The same with polarizable lines (line lists that carry Hund's-case data, readmolecule.c:859-912): rows with
   RHB200_ML_POLARIZABLE != 0 name their MolZeeman components (molzeeman.c:196-319: q, shift in Larmor units, strength
   normalised per q) as [RHB200_ML_ZOFF, + RHB200_ML_NCOMP) of zq / zshift / zstrength [ncomp]; their wavelengths are
   flagged polarised (backgrflags.ispolarized, opacity.c:794-797), the profile is MolProfile's Zeeman sum
   (opacity.c:866-899) and chi_c / eta_c get the Q, U, V parts after the Kurucz lines' (background.c:548-566).  Without
   MAGNETO_OPTICAL only.  rhb200_set_molecular_lines() is the ncomp = 0 case and refuses polarizable rows. */
int rhb200_set_molecular_lines_zeeman(rhb200_ctx *ctx, int nline, const double *mlines, int nmol, const double *molecules,
                                      int ncomp, const int *zq, const double *zshift, const double *zstrength);

/* get_atomic_rfs (rh/inputs.h:92, pyrh.pyx:604-606): the Kurucz lines whose log gf the analytic response function is
   taken for.  line_rows[p] = row of the table passed to rhb200_set_lines that carries parameter p (the reference's
   RLK_Line.loggf_rf_ind, kurucz.c:254-257), or -1 for a parameter no line carries (its column is 0); npar <= 16.  rhb200_set_lines() clears the registration. */
int rhb200_set_loggf_rf(rhb200_ctx *ctx, int npar, const int *line_rows);

/* keywords N_MAX_SCATTER and ITER_LIMIT in LTE (pyrh_compute1dray.c:332-337): n_max_scatter > 0 makes the batched LTE
   entry points Lambda-iterate the continuum-scattering term of the angle-independent (Feautrier) wavelengths,
   S = (eta + sca J)/chi (formal.c:289-309), per column until max |1 - Jdag/J| <= iter_limit or n_max_scatter passes.
   Default 0: the single pass of Iterate().  Only with the continuum on the device. */
int rhb200_set_scatter(rhb200_ctx *ctx, int n_max_scatter, double iter_limit);

/* keyword STOKES_MODE: 1 = FULL_STOKES (default), 0 = NO_STOKES -- I alone at every wavelength with the scalar
   S_INTERPOLATION ray (formal.c:93-103, 223-236), Q = U = V = 0; the reference's own test configuration
   (tests/keyword.input).  Call before rhb200_set_wavelengths. */
int rhb200_set_stokes_mode(rhb200_ctx *ctx, int full_stokes);

/* passive_bb in the fused LTE path (metal.c:174-344): bound-bound lines of the PASSIVE model atoms, hydrogen
   included, unpolarised, added to the background before the Kurucz lines exactly like Background() does
   (background.c:494-515); they also set hasline, i.e. select the scalar ray for their wavelengths.  plines
   [nline][RHB200_PL_NFIELD] in the reference's order (atoms, then lines) with the depth-independent factors of
   Damping() (broad.c:60-314) evaluated by the host; populations, Doppler widths and damping parameters are formed
   on the device from the LTE populations of the continuum model (entry points with the continuum on the device
   only).  Call after rhb200_set_lines and before rhb200_set_wavelengths; rhb200_set_lines clears the table. */
enum {
  RHB200_PL_ATOM = 0,        /* model-atom index (order of atoms.input) */
  RHB200_PL_LEVEL_I, RHB200_PL_LEVEL_J,          /* rows of the level table of rhb200_continuum_model */
  RHB200_PL_LAMBDA0, RHB200_PL_QWING, RHB200_PL_BIJ, RHB200_PL_BJI, RHB200_PL_AJI,
  RHB200_PL_VOIGT,           /* 0: Gaussian */
  RHB200_PL_NCOMP, RHB200_PL_COMPOFF,            /* slice of c_shift / c_fraction */
  RHB200_PL_GRAD,
  RHB200_PL_VDW_TYPE,        /* -1 none, 0 UNSOLD: A * T^0.3, 1 RIDDER_RENSBERGEN: A T^B + C T^D * He abundance,
                                2 BARKLEM (barklem.c:216-312, broad.c:125-136): A T^B + C T^0.3, B = (1 - alpha)/2, C = the Unsold helium term */
  RHB200_PL_VDW_A, RHB200_PL_VDW_B, RHB200_PL_VDW_C, RHB200_PL_VDW_D, RHB200_PL_HE_ABUND,
  RHB200_PL_STARK_TYPE,      /* 0 none, 1: A * ne (cStark < 0), 2: A * (C T)^(1/6) * Cm * ne */
  RHB200_PL_STARK_A, RHB200_PL_STARK_C, RHB200_PL_STARK_CM,
  RHB200_PL_LINSTARK_C,      /* hydrogen: C * ne^(2/3) (StarkLinear) */
  RHB200_PL_IS_H, RHB200_PL_WEIGHT,
  RHB200_PL_NFIELD = 28
};
int rhb200_set_passive_lines(rhb200_ctx *ctx, int nline, const double *plines, int ncomp,
                             const double *c_shift, const double *c_fraction);

/* Lines of the explicit (PASSIVE) model atoms: inside the wing window of such a line a Kurucz line of the same
   element and ionisation stage does not contribute (rlk_opacity, kurucz.c:617-633: passive_bb accounts for it).
   rows [n][4] = {element row of the table given to rhb200_set_lines, stage of the model line's lower level,
   lambda0 [nm, vacuum], qwing}.  Call after rhb200_set_lines (which clears the table) and before
   rhb200_set_wavelengths. */
int rhb200_set_model_lines(rhb200_ctx *ctx, int n, const double *rows);
/* read back the window table: for wavelength i the lines first[i] .. first[i]+count[i]-1
   of the *contributing* list idx[] (test hook for the integer part) */
int rhb200_get_line_windows(rhb200_ctx *ctx, int *first /*[nlambda]*/, int *count /*[nlambda]*/,
                            int *idx /*[cap]*/, int cap, int *nidx);
/* atmos.backgrflags of the current grid (background.c:335-338, 500-566): flags[nlambda], bit 0 = hasline (a Kurucz,
   passive_bb or molecular line has the wavelength in its window), bit 1 = ispolarized */
int rhb200_get_wavelength_flags(rhb200_ctx *ctx, int *flags);

/* ---- the hot path ------------------------------------------------------
   LTE FULL_STOKES synthesis of `ncol` independent columns.  Replaces, per column,
     Background(): rlk_opacity() for every (lambda, mu, to_obs)   background.c:519-546, kurucz.c:511-828
     Iterate() -> solveSpectrum() -> Formal() -> Piece_Stokes_Bezier3_1D
                                                   iterate.c:48, formal.c:157-275, bezier_1D.c:52-300
     _solveray() packing                                          pyrh_solveray.c:130-150
   of the reference's rhf1d() (pyrh_compute1dray.c:112-389) for the case
   solve_NLTE = FALSE (no ACTIVE atom), STOKES_MODE = FULL_STOKES,
   S_INTERPOLATION_STOKES = DELO_BEZIER3, Nrays = 1.

   atmos   [ncol][RHB200_AT_NFIELD][ndep]   SI
   chi_ai, eta_ai [ncol][nlambda][ndep]     angle-independent background (chi_ai/eta_ai of
                                            background.c:343-465 incl. passive_bb), reference layout
   stokes  [ncol][4][nlambda]               emergent I,Q,U,V at mu (W m^-2 Hz^-1 sr^-1)
   moving: atmos.moving (pyrh_compute1dray.c:261-278); bc_*: geometry.vboundary (getBoundary)
   All pointers are HOST pointers; H2D/D2H copies are part of the call. */
int rhb200_lte_stokes_batch(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                            int bc_top, int bc_bottom,
                            const double *atmos, const double *chi_ai, const double *eta_ai,
                            double *stokes);
/* same, with DEVICE pointers (inputs already resident in HBM; no copies) */
int rhb200_lte_stokes_batch_dev(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                                int bc_top, int bc_bottom,
                                const double *d_atmos, const double *d_chi_ai,
                                const double *d_eta_ai, double *d_stokes);

/* ---- function-level entry points (each mirrors one reference function; used by
        the parity tests and by hosts that keep the rest of Formal() on the CPU).
        HOST pointers. ---------------------------------------------------- */

/* LTEpops_elem (ltepops.c:116-159) for every element row: n [ncol][nelem][RHB200_RE_MAXSTAGE][ndep] */
int rhb200_ltepops_elem_batch(rhb200_ctx *ctx, int ncol, int ndep, const double *atmos, double *n);

/* rlk_opacity (kurucz.c:511-725) at every set wavelength, direction to_obs:
   chi, eta [ncol][nlambda][4][ndep] (line part only, accumulated from 0 in line order),
   flags [nlambda] bit0 hasline bit1 ispolarized (may be NULL) */
int rhb200_rlk_opacity_batch(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                             int to_obs, const double *atmos, double *chi, double *eta, int *flags);

/* MolecularOpacity + MolProfile + the window tests that mrt_locate serves (opacity.c:656-916): LTE lines of
   PASSIVE molecules at every given wavelength, direction to_obs.
   mlines [nmline][RHB200_ML_NFIELD], grouped by molecule in the reference's order (molecule index
   ascending, lines as read = ascending lambda0); Zeeman slices as for Kurucz lines (MolZeeman output;
   non-polarizable lines use VoigtArmstrong, opacity.c:911).
   mol [ncol][nmol][3][ndep] = molecule->n [m^-3], molecule->pf, molecule->vbroad [m/s] (host: chemical
   equilibrium).  Out: chi, eta [ncol][nlambda][4][ndep] accumulated from 0 in line order,
   flags [nlambda] bit0 hasline, bit1 ispolarized (may be NULL). */
enum {
  RHB200_ML_LAMBDA0 = 0, RHB200_ML_EI, RHB200_ML_GI, RHB200_ML_BIJ, RHB200_ML_AJI, RHB200_ML_BJI,
  RHB200_ML_ISO_FRAC, RHB200_ML_QWING, RHB200_ML_POLARIZABLE, RHB200_ML_MOL, RHB200_ML_ZOFF, RHB200_ML_NCOMP,
  RHB200_ML_NFIELD = 16
};
int rhb200_molecular_opacity_batch(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving, int to_obs,
                                   int nmol, int nmline, const double *mlines,
                                   int ncomp, const int *zq, const double *zshift, const double *zstrength,
                                   double vmicro_char, int nlambda, const double *lambda,
                                   const double *atmos, const double *mol,
                                   double *chi, double *eta, int *flags);

/* passive_bb (metal.c:174-344): bound-bound lines of PASSIVE model atoms (hydrogen included) in the
   background, unpolarised: VoigtArmstrong with the host's Damping() output, Gaussian when line->Voigt is off,
   line components (c_shift, c_fraction).  plines [nline][RHB200_PB_NFIELD] in the reference's order (atoms, then
   lines of each atom); pcol [ncol][nline][4][ndep] = n_i, n_j (LTE or NLTE populations the host holds), atom
   vbroad, adamp.  Out: chi, eta [ncol][nlambda][ndep]; flags [nlambda] bit0 hasline (may be NULL). */
enum {
  RHB200_PB_LAMBDA0 = 0, RHB200_PB_QWING, RHB200_PB_BIJ, RHB200_PB_BJI, RHB200_PB_AJI, RHB200_PB_VOIGT,
  RHB200_PB_NCOMP, RHB200_PB_COMPOFF, RHB200_PB_NFIELD = 8
};
int rhb200_passive_bb_batch(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving, int to_obs,
                            int nline, const double *plines, int ncomp, const double *c_shift,
                            const double *c_fraction, double vmicro_char, int nlambda, const double *lambda,
                            const double *atmos, const double *pcol, double *chi, double *eta, int *flags);

/* Piece_Stokes_Bezier3_1D (bezier_1D.c:52-300) + StokesK (stokesopac.c:28-87) for nray rays.
   ray_col[nray] selects the column (height, T) of each ray, ray_lambda[nray] its wavelength [nm].
   chi [nray][ndep], S [nray][4][ndep], chiQUV [nray][3][ndep] (numerators of K', un-divided),
   out I [nray][4][ndep], Psi [nray][ndep] or NULL.  height,T: [ncol][ndep]. */
int rhb200_stokes_bezier3_batch(rhb200_ctx *ctx, int nray, int ncol, int ndep, double muz, int to_obs,
                                int bc_top, int bc_bottom,
                                const int *ray_col, const double *ray_lambda,
                                const double *height, const double *T,
                                const double *chi, const double *S, const double *chiQUV,
                                double *I, double *Psi);

/* Piecewise_Bezier3_1D (bezier_1D.c:306-541, without the log gf response function):
   chi, S [nray][ndep]; out I, Psi [nray][ndep] (Psi may be NULL) */
int rhb200_bezier3_batch(rhb200_ctx *ctx, int nray, int ncol, int ndep, double muz, int to_obs,
                         int bc_top, int bc_bottom,
                         const int *ray_col, const double *ray_lambda,
                         const double *height, const double *T,
                         const double *chi, const double *S, double *I, double *Psi);

/* ---- Host-side Zeeman machinery (once per line list; pure host code, no device needed) ----
   These return a count / flag >= 0 on success and a negative RHB200_E* code on error.
   RLKdeterminate (kurucz.c:925-969): S, L of both levels from the 10-character Kurucz term labels;
   returns 1 if determined (the line is then polarizable in Stokes mode, kurucz.c:295), else 0. */
int rhb200_rlk_determinate(const char *labeli, const char *labelj, double *Si, int *Li, double *Sj, int *Lj);
/* Lande (zeeman.c:139-146): LS-coupling g factor */
double rhb200_lande(double S, int L, double J);
/* RLKZeeman (kurucz.c:832-921): Zeeman components of a Kurucz line, in the reference's (Ml outer,
   Mu inner) order, strengths normalised per q.  gL_i/gL_j: Lande factors from the line list
   (-99e-3 = not given), used unless LS_Lande (keyword LS_LANDE).  Returns Ncomponent; nothing is
   written when Ncomponent > cap. */
int rhb200_rlk_zeeman(double gi, double gj, double Si, int Li, double Sj, int Lj, double gL_i, double gL_j,
                      int LS_Lande, int cap, int *q, double *shift, double *strength);
/* determinate (zeeman.c:37-85): n, S, L, J of a model-atom level from its 20-character label; 1 = determined */
int rhb200_determinate(const char *label, double g, int *n, double *S, int *L, double *J);
/* Zeeman (zeeman.c:186-281): pattern of a model-atom line; g_Lande_eff != 0 selects the normal triplet */
int rhb200_zeeman(const char *label_i, double g_i, const char *label_j, double g_j, double g_Lande_eff,
                  int cap, int *q, double *shift, double *strength);

/* ---- Background continuum on the device ----
   The angle-independent part of Background() (background.c:343-465): Thomson, H- bf/ff, OH/CH bf, H bf/ff,
   Rayleigh (H, He, H2), H2+ ff, H2- ff and the bound-free continua of all PASSIVE model atoms, summed in the
   reference's order with its expressions.  The RH host keeps what does not depend on the column (the model
   below, filled once) and what it computes per column anyway (LTE populations + chemical equilibrium); the
   library evaluates everything that scales with columns x wavelengths x depths.
   lambda >= 9113 nm (Hminus_ff_long) returns RHB200_EUNSUPPORTED. */
typedef struct rhb200_continuum_model {
  int natom, nlev, ncont, ntab, nray;
  const double *lev;         /* [nlev][5]  atom index, E [J], stage, g, atom->active; atmos.atoms order, H first */
  const double *bf;          /* [ncont][10] atom, level i, level j (rows of lev), lambda0, lambda[0], hydrogenic,
                                alpha0, Nlambda, offset into tab_*, atom->active */
  const double *tab_lambda, *tab_alpha;      /* [ntab] continuum->lambda / ->alpha of the tabulated edges */
  const double *ray;         /* [nray][8] lines from the ground state for Rayleigh(): 0 = H / 1 = He, lambda0,
                                qwing, Aji, g_j, g_0, (unused), (unused) */
  int nlev_H;                /* atmos.H->Nlevel (levels 0 .. nlev_H-1 of lev; the last is the proton) */
  int atom_He;               /* atom index of the helium model, -1 if absent (background.c:285) */
  int H_active, has_OH, has_CH, has_H2, solve_NLTE, do_fudge;
  double vmicro_char;        /* [m/s] */
  /* published tables the reference holds as function-static arrays (hydrogen.c, ohchbf.c) */
  const double *hmbf_lambda, *hmbf_alpha;                     int n_hmbf;
  const double *hmff_lambda, *hmff_theta, *hmff_kappa;        int n_hmff_lambda, n_hmff_theta;
  const double *h2mff_lambda, *h2mff_theta, *h2mff_kappa;     int n_h2mff_lambda, n_h2mff_theta;
  const double *h2pff_lambda, *h2pff_temp, *h2pff_kappa;      int n_h2pff_lambda, n_h2pff_temp;
  const double *rh2_a, *rh2_lambda, *rh2_sigma;               int n_rh2;
  const double *oh_T, *oh_E, *oh_cross;                       int n_oh_T, n_oh_E;
  const double *ch_T, *ch_E, *ch_cross;                       int n_ch_T, n_ch_E;
  /* opacity fudge factors (do_fudge != 0; pyrh.compute1d's fudge_wave / fudge_value, pyrh_compute1dray.c:183-196,
     background.c:364-371, 438-451, 456-464): n_fudge wavelengths [nm] and three rows of factors -- H-, scattering,
     metal bound-free -- interpolated linearly in wavelength */
  int n_fudge;  const double *fudge_lambda, *fudge /* [3][n_fudge] */;
} rhb200_continuum_model;
/* T, ne, nHmin, nH2, nOH, nCH [ncol][ndep] (SI; the molecular ones may be NULL when has_* is 0);
   pops_n, pops_nstar [ncol][nlev][ndep] = atom->n / atom->nstar of every level (may be the same pointer);
   out chi_ai, eta_ai, sca_ai [ncol][nlambda][ndep] (sca_ai may be NULL). */
int rhb200_continuum_batch(rhb200_ctx *ctx, const rhb200_continuum_model *model, int nlambda, const double *lambda,
                           int ncol, int ndep, const double *T, const double *ne, const double *nHmin,
                           const double *nH2, const double *nOH, const double *nCH,
                           const double *pops_n, const double *pops_nstar,
                           double *chi_ai, double *eta_ai, double *sca_ai,
                           double *contrib /* NULL, or [ncol][nlambda][13][2][ndep]: every contribution on its own
                              (diagnostics; order Thomson, H- bf, H- ff, OH bf, CH bf, H bf, H ff, Rayleigh H,
                              Rayleigh He, H2+ ff, Rayleigh H2, H2- ff, metal bf; [..][0] = chi or scatt, [..][1] = eta) */);

/* Fused LTE path with the continuum on the device.  rhb200_set_continuum (after rhb200_set_wavelengths)
   builds the per-wavelength coefficients once; abundance [natom] = atom->abundance (readatom.c:190).
   rhb200_lte_stokes_batch_pops is rhb200_lte_stokes_batch without chi_ai / eta_ai: per column it takes
   chem [ncol][natom+4][ndep] = the factor ChemicalEquilibrium() applies to each model atom's populations
   (ntotal_after / ntotal_before, chemequil.c:336; 1 for atoms in no molecule), then nHmin, nH2, nOH, nCH;
   the library evaluates LTEpops (ltepops.c:33-113), the continuum, the line opacity and the formal solution.
   All model atoms must be PASSIVE with LTE populations. */
int rhb200_set_continuum(rhb200_ctx *ctx, const rhb200_continuum_model *model, const double *abundance);
int rhb200_lte_stokes_batch_pops(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                                 int bc_top, int bc_bottom, const double *atmos, const double *chem,
                                 double *stokes);

/* ChemicalEquilibrium (chemequil.c:107-392) on the device.  After rhb200_set_continuum: the nuclei that are bound
   in molecules (atmos.elements order, hydrogen first) with the index of their model atom, and per molecule a
   32-double record {fit (enum fit_type, atom.h:32), charge, Nnuclei, Nelement, Neqc, Tmin, Tmax, Ediss [J],
   eqc_coef[8], nucleus index of each element [4], pt_count[4], is H2, is OH, is CH, 0...}.
   rhb200_lte_stokes_batch_atmos then needs nothing per column but the atmosphere rows: LTEpops, chemical
   equilibrium, continuum, line opacity and formal solution all run on the device.
   rhb200_chemistry_batch exposes the first two steps: chem [ncol][natom+4][ndep] (layout of
   rhb200_lte_stokes_batch_pops), pops [ncol][nlev][ndep] (may be NULL). */
int rhb200_set_chemistry(rhb200_ctx *ctx, int nnuclei, const int *nucleus_atom, int nmol, const double *mol);
int rhb200_chemistry_batch(rhb200_ctx *ctx, int ncol, int ndep, const double *atmos, double *chem, double *pops);
int rhb200_lte_stokes_batch_atmos(rhb200_ctx *ctx, int ncol, int ndep, double muz, int moving,
                                  int bc_top, int bc_bottom, const double *atmos, double *stokes);

/* pyrh.compute1d() / rhf1d() for a batch of columns, LTE (pyrh.pyx:537-668; pyrh_compute1dray.h:26-36,
   pyrh_compute1dray.c:112-357).  `atmosphere` [ncol][nrow >= 9][ndep] holds the rows pyrh.compute1d takes, in its
   units (pyrh.pyx:621-625): 0 scale (log10 tau500 | log10 column mass [g cm^-2] | height [km] for atm_scale
   0 | 1 | 2, pyrh_compute1dray.c:230-246), 1 T [K], 2 ne [cm^-3], 3 v_z [km/s], 4 v_mic [km/s], 5 B [G],
   6 gamma [rad], 7 chi [rad], 8 nH_tot [cm^-3].  Everything the reference derives per column runs on the device:
   the unit conversion (:263-270), Bproject() for mu == 1 and inclined rays (rhf1d/project.c:38-80), atmos.moving
   against VMACRO_TRESH (`vmacro_tresh`, m/s; :272-278), LTEpops, ChemicalEquilibrium, the proton density
   (kurucz.c:772), the background continuum, the line opacity, convertScales() (rhf1d/multiatmos.c:100-177) and the
   formal solution.  Needs rhb200_set_lines, rhb200_set_wavelengths, rhb200_set_continuum, rhb200_set_chemistry.
   The wavelength grid is spectrum.lambda, i.e. it CONTAINS the reference wavelength (500 nm, atmos.lambda_ref) at
   index `iref` (sortlambda.c adds it; convertScales looks it up with Locate()); `stokes` [ncol][4][nlambda]
   therefore has one more column than the spectrum _solveray() packs (pyrh_solveray.c:130-150 drops it).
   `wght_per_H` = sum over elements of abundance x atomic weight (abundance.c:186-220).
   `scales` (may be NULL) [ncol][3][ndep]: height [m], tau_ref and column mass [kg m^-2] the reference would hold in
   geometry.height / tau_ref / cmass after convertScales() (for atm_scale 2 the column-mass row needs total_abund
   and gravity: rhb200_set_gravity first, else RHB200_EINVAL). */
int rhb200_compute1d_batch(rhb200_ctx *ctx, int ncol, int ndep, int nrow, double mu, int atm_scale,
                           const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                           int bc_top, int bc_bottom, double *stokes, double *scales);

/* One host batch over several GPUs of the box, from one process (BASELINE configs[1]: "column-sharded over 1/2/4/8
   GPUs"): ctxs[nctx] are contexts opened on different devices and set up with the same tables; the columns are cut into
   nctx contiguous blocks (rhb200_shard_columns), every block runs rhb200_compute1d_batch on its own context from its own
   host thread -- copies and kernels of the devices overlap -- and the call returns when all are done.  No data-path
   collective: columns are independent.  Arguments as rhb200_compute1d_batch; page-locked host buffers
   (rhb200_host_alloc_pinned) keep the copies of the devices concurrent.  Returns the first error of any block. */
int rhb200_compute1d_batch_multi(int nctx, rhb200_ctx *const *ctxs, int ncol, int ndep, int nrow, double mu, int atm_scale,
                                 const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                                 int bc_top, int bc_bottom, double *stokes, double *scales);
/* block [*first, *first + *count) of `ncol` columns that shard `rank` of `nrank` takes (contiguous, sizes differ by at
   most one) */
int rhb200_shard_columns(int ncol, int rank, int nrank, int *first, int *count);

/* rhb200_compute1d_batch with get_atomic_rfs = 1: additionally rfs [ncol][nlambda][npar] = atmos.atomic_rfs[nspect][0][p]
   (formal.c:278-282, pyrh_solveray.c:144-147; pyrh.compute1d returns its transpose without the lambda_ref entry), the
   response of the emergent intensity to log gf of the lines registered with rhb200_set_loggf_rf.  As in the reference
   it is non-zero only at wavelengths Formal() solves with Piecewise_Bezier3_1D (bezier_1D.c:416-516): an unpolarised
   line in a moving column, or any line in NO_STOKES mode; wavelengths solved by the polarised solver or by Feautrier
   return 0.  Where the reference reads uninitialised memory (a registered line outside the wavelength's window:
   sortlambda.c:608-611 mallocs dchi_c_lam, kurucz.c:696 only writes inside the window) the entry is 0 here.
   stokes, scales may be NULL. */
int rhb200_compute1d_rf_batch(rhb200_ctx *ctx, int ncol, int ndep, int nrow, double mu, int atm_scale,
                              const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                              int bc_top, int bc_bottom, double *stokes, double *scales, double *rfs);

/* atmos.totalAbund (abundance.c:219) and atmos.gravity [m s^-2] (multiatmos.c:69,82): the compute1d entry points need
   them for ONE thing, the column-mass row of `scales` when the column came on a height grid (atm_scale 2,
   multiatmos.c:153-155); asking for that row without this call is RHB200_EINVAL. */
int rhb200_set_gravity(rhb200_ctx *ctx, double total_abund, double gravity);

/* pyrh.get_scales() for a batch (pyrh.pyx:491-534, rhf1d/pyrh_hse.c:402-553): Background() at the reference
   wavelength and convertScales(), nothing else.  The reference sets atmos.Nrlk = 0 there, so the context normally
   holds an empty line table (rhb200_set_lines with nline = 0) and the one-wavelength grid {lam_ref}, iref = 0.
   total_abund = sum of abundances, gravity = 10^4.4 cm s^-2 in m s^-2 (multiatmos.c:69,82; abundance.c:219) enter
   the column mass of a height scale (multiatmos.c:153-155).  scales [ncol][3][ndep] = height [m], tau_ref,
   column mass [kg m^-2]. */
int rhb200_get_scales_batch(rhb200_ctx *ctx, int ncol, int ndep, int nrow, int atm_scale, const double *atmosphere,
                            int iref, double wght_per_H, double total_abund, double gravity, double vmacro_tresh,
                            double *scales);

/* Electron density from the LTE ionisation equilibrium of all elements: Solve_ne (rh/solvene.c:55-140), the
   computation behind pyrh.get_ne_from_nH (pyrh.pyx:396-425, rhf1d/pyrh_hse.c:555-677: Background(FALSE, TRUE) with
   SOLVE_NE = ONCE, hydrogen in LTE).  rhb200_set_elements hands over atmos.elements[] -- all elements of the
   periodic table, hydrogen first, rows RHB200_RE_* as in rhb200_set_lines, pf = ln U [npf_rows][npf] -- once;
   rhb200_solve_ne_batch solves n independent depth points: T [K], nHtot [m^-3] in, ne [m^-3] out (and starting
   guess when fromscratch == 0). */
int rhb200_set_elements(rhb200_ctx *ctx, int nelem, const double *elems, int npf_rows, int npf,
                        const double *pf, const double *Tpf);
int rhb200_solve_ne_batch(rhb200_ctx *ctx, size_t n, const double *T, const double *nHtot, double *ne,
                          int fromscratch);

/* pyrh.hse (pyrh.pyx:427-489; hse(), rhf1d/pyrh_hse.c:67-400) for a batch of columns: the gas pressure, electron
   density, total hydrogen density and mass density of an atmosphere in hydrostatic equilibrium, from its
   temperature run on a tau500 (atm_scale 0, scale = log10 tau500) or height (atm_scale 2, scale in km) grid and the
   gas pressure at the top.  Per layer, top down, the reference iterates {rho, get_ne() from scratch, LTE populations,
   ChemicalEquilibrium, pyrh_Background() = the 500 nm continuum without Metal_bf but with Thomson and Rayleigh
   scattering, pressure integration, new nHtot} until nHtot changes by <= 1 %; every step of that walk is a kernel
   over all columns here.  Needs rhb200_set_elements, the one-wavelength grid {500 nm} (rhb200_set_lines with
   nline = 0, rhb200_set_wavelengths), rhb200_set_continuum and rhb200_set_chemistry.
   scale, T [ncol][ndep]; pg_top [ncol] in Pa; outputs [ncol][ndep] in SI (m^-3, kg m^-3, Pa) like the reference's. */
int rhb200_hse_batch(rhb200_ctx *ctx, int ncol, int ndep, int atm_scale, const double *scale, const double *T,
                     const double *pg_top, double wght_per_H, double total_abund, double gravity,
                     double *ne, double *nHtot, double *rho, double *pg);

/* Finite-difference response functions of the LTE Stokes spectrum to the atmosphere rows (BASELINE config 3: T,
   v_LOS, B, inclination, azimuth per depth).  pyrh has no entry point for these: its callers perturb one row at
   one depth by +-delta and call pyrh.compute1d twice (2 x npar x ndep calls per column); this does the same for a
   batch, expanding the perturbed columns on the device and returning only
       rf [ncol][npar][ndep][4][nlambda] = (S(x_k + delta) - S(x_k - delta)) / (2 delta)
   with S exactly what rhb200_compute1d_batch returns for the perturbed column (so the result equals the one
   formed from the reference's own rhf1d() spectra bit for bit).  par_rows[p] = row of `atmosphere` (1 T, 3 v_z,
   4 v_mic, 5 B, 6 gamma, 7 chi, also 2 ne, 8 nH), par_delta[p] > 0 in the row's unit.  Other arguments as in
   rhb200_compute1d_batch; nlambda includes the reference wavelength.
   rhb200_rf_fd_depths_batch restricts the perturbations to the depths an inversion has its nodes at: work and the
   device-to-host traffic scale with nsel / ndep. */
int rhb200_rf_fd_depths_batch(rhb200_ctx *ctx, int ncol, int ndep, int nrow, double mu, int atm_scale,
                              const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                              int bc_top, int bc_bottom, int npar, const int *par_rows, const double *par_delta,
                              int nsel, const int *depths /* [nsel] depth indices, or NULL: all */,
                              double *rf /* [ncol][npar][nsel][4][nlambda] */);
int rhb200_rf_fd_batch(rhb200_ctx *ctx, int ncol, int ndep, int nrow, double mu, int atm_scale,
                       const double *atmosphere, int iref, double wght_per_H, double vmacro_tresh,
                       int bc_top, int bc_bottom, int npar, const int *par_rows, const double *par_delta,
                       double *rf);

/* Formal-solver selection = keyword.input S_INTERPOLATION / S_INTERPOLATION_STOKES (readvalue.c:366-404;
   enum values of inputs.h:26-27).  Applies to rhb200_lte_stokes_batch(_dev) (Stokes solver) and to
   rhb200_nlte_iterate / rhb200_nlte_formal (scalar solver).  Defaults: S_BEZIER3, DELO_BEZIER3. */
enum { RHB200_S_LINEAR = 0, RHB200_S_PARABOLIC = 1, RHB200_S_BEZIER3 = 2 };
enum { RHB200_DELO_PARABOLIC = 0, RHB200_DELO_BEZIER3 = 1 };
int rhb200_set_solvers(rhb200_ctx *ctx, int s_interpolation, int s_interpolation_stokes);

/* One scalar ray per entry with the chosen solver: Piecewise_Linear_1D (piecewise_1D.c:44-127),
   Piecewise_1D (piecewise_1D.c:134-253) or Piecewise_Bezier3_1D.  Layout as rhb200_bezier3_batch. */
int rhb200_scalar_ray_batch(rhb200_ctx *ctx, int solver, int nray, int ncol, int ndep, double muz, int to_obs,
                            int bc_top, int bc_bottom,
                            const int *ray_col, const double *ray_lambda,
                            const double *height, const double *T,
                            const double *chi, const double *S, double *I, double *Psi);

/* One polarised ray per entry: Piece_Stokes_1D (piecestokes_1D.c:49-174, DELO_PARABOLIC) or
   Piece_Stokes_Bezier3_1D.  Layout as rhb200_stokes_bezier3_batch. */
int rhb200_stokes_ray_batch(rhb200_ctx *ctx, int solver, int nray, int ncol, int ndep, double muz, int to_obs,
                            int bc_top, int bc_bottom,
                            const int *ray_col, const double *ray_lambda,
                            const double *height, const double *T,
                            const double *chi, const double *S, const double *chiQUV,
                            double *I, double *Psi);

/* Analytic log gf response function (get_atomic_rfs; bezier_1D.c:416-428, 477-490, 509-516): the two
   Piecewise_Bezier3_1D passes Formal() makes at one (wavelength, mu) in NO_STOKES mode (formal.c:167-283).
   chi_dn/S_dn: opacity and source function of the down-ray (to_obs = 0), chi_up/S_up of the up-ray
   [nray][ndep]; dchi, deta [nray][ndep][npar] = spectrum.dchi_c_lam / deta_c_lam (kurucz.c:696-699);
   out I [nray][ndep] (up-ray), dI [nray][ndep][npar]; atmos.atomic_rfs[nspect][mu][p] = dI[ray][0][p]
   (formal.c:278-282).  npar <= 16. */
int rhb200_bezier3_rf_batch(rhb200_ctx *ctx, int nray, int ncol, int ndep, double muz,
                            int bc_top, int bc_bottom,
                            const int *ray_col, const double *ray_lambda,
                            const double *height, const double *T,
                            const double *chi_dn, const double *S_dn, const double *chi_up, const double *S_up,
                            int npar, const double *dchi, const double *deta, double *I, double *dI);

/* Feautrier (feautrier.c:56-202, F_order = STANDARD): chi, S [nray][ndep]; out P [nray][ndep]
   (Feautrier mean intensity along the ray), Psi [nray][ndep] or NULL, Iem [nray] emergent intensity */
int rhb200_feautrier_batch(rhb200_ctx *ctx, int nray, int ncol, int ndep, double muz,
                           int bc_top, int bc_bottom, const int *ray_col, const double *ray_lambda,
                           const double *height, const double *T, const double *chi, const double *S,
                           double *P, double *Psi, double *Iem);

/* Voigt(a, v, &F, HUMLICEK) (voigt.c:381-419, humlicek.c): H, F for n (a, v) pairs;
   region[n] (1..4) may be NULL */
int rhb200_voigt_humlicek(rhb200_ctx *ctx, int n, const double *a, const double *v,
                          double *H, double *F, int *region);

/* Voigt(a, v, NULL, ARMSTRONG) (voigt.c:126-243): H for n (a, v) pairs; region[n] = 1,2,3 for
   VoigtK1/K2/K3 (may be NULL) */
int rhb200_voigt_armstrong(rhb200_ctx *ctx, int n, const double *a, const double *v, double *H, int *region);

/* exp / pow / sin / cos exactly as the device code evaluates them (test hook for
   the glibc-equivalence of the device math) */
int rhb200_math_probe(rhb200_ctx *ctx, int n, int func /*0 exp 1 sin 2 cos 3 pow 4 x/y via shared reciprocal 5 x/y 6 atan 7 log 8 log10*/,
                      const double *x, const double *y, double *out);

/* ---- NLTE: MALI iteration for active atoms (CRD, unpolarised radiation) ------------------
   Replaces Iterate() (iterate.c:48-143): per iteration initGammaAtom, solveSpectrum(eval_operator)
   = per wavelength Formal() with Opacity, Piecewise_Bezier3_1D / Feautrier, addtoCoupling,
   addtoGamma, addtoRates (formal.c:44-346, opacity.c:64-390, fillgamma.c), then updatePopulations =
   statEquil + Accelerate + MaxChange (statequil.c:177-216).  `ncol` independent columns share one
   atomic model / wavelength structure (the "plan"); each column stops at its own ITER_LIMIT.

   Transition rows (doubles), one per radiative transition, lines first then continua per atom: */
enum {
  RHB200_TR_ATOM = 0, RHB200_TR_TYPE /* 0 line, 1 continuum */, RHB200_TR_I, RHB200_TR_J,
  RHB200_TR_NBLUE, RHB200_TR_NLAMBDA,            /* AtomicLine/AtomicContinuum Nblue, Nlambda (atom.h:50-75) */
  RHB200_TR_AJI, RHB200_TR_BJI, RHB200_TR_BIJ, RHB200_TR_ISOFRAC,
  RHB200_TR_WOFF,                                /* offset of this transition in tr_lambda/tr_wlambda/tr_alpha */
  RHB200_TR_PHIROW,                              /* first row of line->phi[2*Nrays*Nlambda] in the phi table */
  RHB200_TR_KR, RHB200_TR_LINEIDX,               /* index in atom->line / row of wphi */
  RHB200_TR_LAMBDA0,                             /* line->lambda0 [nm] (used when profiles are computed on the device) */
  RHB200_TR_NFIELD = 16
};
typedef struct {
  int Nspect, Nrays, Ndep, Natom, Ntrans, moving;
  int Ngorder, Ngdelay, Ngperiod, isum;          /* keywords NG_ORDER, NG_DELAY, NG_PERIOD, I_SUM */
  int bc_top, bc_bottom;
  int ntrl, nphirow, nline;                      /* lengths of the concatenated tables below */
  const double *lambda;                          /* [Nspect] spectrum.lambda */
  const double *muz, *wmu;                       /* [Nrays] */
  const int    *atom_nlevel;                     /* [Natom] */
  const double *trans;                           /* [Ntrans][RHB200_TR_NFIELD] */
  const double *tr_lambda, *tr_wlambda, *tr_alpha;   /* [ntrl]: transition wavelengths, getwlambda_line/_cont weights, alpha */
  const int    *as_first;                        /* [Nspect+1] CSR of the active sets (spectrum.as[].art) */
  const int    *as_trans;                        /* [as_first[Nspect]] transition rows in the reference's order */
  const int    *bg_hasline;                      /* [Nspect] atmos.backgrflags[].hasline */
} rhb200_nlte_plan;
typedef struct {                                 /* per-column arrays, column-major blocks [ncol][...] */
  const double *T, *height;                      /* [ncol][Ndep] */
  const double *nstar;                           /* [ncol][sum Nlevel][Ndep] */
  const double *ntotal;                          /* [ncol][Natom][Ndep] */
  const double *C;                               /* [ncol][sum Nlevel^2][Ndep] collisional rates (atom->C) */
  const double *phi;                             /* [ncol][nphirow][Ndep]  line->phi (Profile(), profile.c:67) */
  const double *wphi;                            /* [ncol][nline][Ndep] */
  /* When phi == NULL the profiles are evaluated on the device (Profile(), field-free branch
     profile.c:311-323, single-component lines, VoigtArmstrong) from: */
  const double *adamp;                           /* [ncol][nline][Ndep]  Damping() (broad.c:273), host */
  const double *vbroad;                          /* [ncol][Natom][Ndep]  atom->vbroad */
  const double *vel;                             /* [ncol][Ndep]         geometry.vel [m/s] */
  const double *chi_c, *eta_c, *sca_c;           /* [ncol][Nspect][Ndep] background (spectrum.*_c_lam) */
  double *n;                                     /* [ncol][sum Nlevel][Ndep]  in: initial, out: converged */
  double *J;                                     /* [ncol][Nspect][Ndep]      in: after initScatter, out: final */
} rhb200_nlte_columns;
/* niter [ncol] iterations done; dpops_hist [ncol][NmaxIter] or NULL; when dump_iter >= 1 the Gamma
   matrices [ncol][sum Nl^2][Ndep] and rates {Rij [ncol][Ntrans][Ndep], Rji ...} of that iteration are
   copied out before statEquil (test hooks, may be NULL); phi_out/wphi_out receive the device-evaluated
   profiles [ncol][nphirow][Ndep] / [ncol][nline][Ndep] when cols->phi == NULL (may be NULL). */
int rhb200_nlte_iterate(rhb200_ctx *ctx, const rhb200_nlte_plan *plan, int ncol,
                        const rhb200_nlte_columns *cols, int NmaxScatter /* initScatter passes first, initscatter.c:62 */,
                        int NmaxIter, double iterLimit,
                        int *niter, double *dpops_hist, int dump_iter, double *gamma_dump,
                        double *rates_dump, double *phi_out, double *wphi_out);
/* solveSpectrum(FALSE, FALSE) (iterate.c:148-253) repeated up to npass times: the Lambda iteration of
   initScatter() (update_J = 1; a column stops when dJmax < dJlimit, initscatter.c:62-68), the extra
   scattering passes rhf1d() runs after Iterate() (update_J = 2: stops on dJmax <= dJlimit,
   pyrh_compute1dray.c:333-337) or the final
   formal solution of _solveray() (npass = 1, update_J = 0: J is used but not modified,
   pyrh_solveray.c:84-106).  Iem [ncol][Nspect][Nrays] = spectrum.I[nspect][mu] (may be NULL);
   cols->n is not modified, cols->J only when update_J. */
int rhb200_nlte_formal(rhb200_ctx *ctx, const rhb200_nlte_plan *plan, int ncol,
                       const rhb200_nlte_columns *cols, int npass, int update_J, double dJlimit,
                       double *Iem, int *npass_done);

/* ---- Wavelength sharding of ONE atmosphere across ranks ("config 4") ----
   Each rank formally solves a contiguous chunk of the sorted wavelengths (balanced by ray count,
   rhb200_nlte_shard_range) and the per-depth radiative rates are summed over ranks once per MALI
   iteration: Gamma [ncol][ngam][Ndep] (collisional part added on rank 0 only), Rij, Rji
   [ncol][Ntrans][Ndep]; dJmax [ncol] is max-reduced in the scattering passes; J and the emergent
   intensities are sum-reduced once at the end.  The library does not link a communication library:
   the host supplies the reduction, e.g. ncclAllReduce(buf, buf, count, ncclDouble, op, comm, stream)
   + cudaStreamSynchronize.  `fn` is called with a DEVICE pointer to `count` doubles after the
   library has drained its stream, must reduce in place over all ranks, and must have completed
   when it returns 0.  nrank = 1 switches sharding off. */
enum { RHB200_REDUCE_SUM = 0, RHB200_REDUCE_MAX = 1 };
typedef int (*rhb200_allreduce_fn)(void *user, double *device_buf, size_t count, int op);
int rhb200_nlte_set_shard(rhb200_ctx *ctx, int rank, int nrank, rhb200_allreduce_fn fn, void *user);
/* The same exchange with NCCL called by the library itself: libnccl is loaded at run time (dlopen; RHB200_NCCL_LIB names
   it, else libnccl.so.2), nothing is linked.  Per MALI iteration Gamma, Rij, Rji go out as ONE ncclGroup of three
   ncclAllReduce(.., ncclDouble, ncclSum, comm, stream) on the context's compute stream -- no host synchronisation, no
   callback; dJmax (ncclMax), J and the emergent intensities likewise.  `nccl_comm` is the host's ncclComm_t for this
   rank's GPU; or let the library create it: rank 0 calls rhb200_nccl_unique_id, ships the 128 bytes to the other
   ranks by any means (MPI_Bcast, a file, torch.distributed), every rank calls rhb200_nlte_set_shard_nccl_id. */
int rhb200_nlte_set_shard_nccl(rhb200_ctx *ctx, int rank, int nrank, void *nccl_comm);
int rhb200_nccl_unique_id(char id[128]);
int rhb200_nlte_set_shard_nccl_id(rhb200_ctx *ctx, int rank, int nrank, const char id[128]);
/* Order of the sums in addtoGamma / addtoRates (fillgamma.c:139-196, 375-461).  exact = 1: one thread per (column,
   transition, depth) adds the contributions of all wavelengths and rays in the reference's order -- Gamma, rates and
   populations then equal the reference's to the last bit, at the price of a long serial walk.  exact = 0 (default):
   every transition's wavelengths are cut into fixed segments of 16 (RHB200_NLTE_GAMMA_SEG) that are summed
   concurrently and then added in segment order: deterministic and independent of batch size, chunking and rank count,
   populations within ~1e-13 of the exact mode (north_star bar: 1e-6), several times faster.  The environment
   variable RHB200_NLTE_EXACT overrides. */
int rhb200_nlte_set_exact_rates(rhb200_ctx *ctx, int exact);
int rhb200_nlte_shard_range(const rhb200_nlte_plan *plan, int rank, int nrank, int *ns_lo, int *ns_hi);
/* ---- NLTE through the drop-in call: everything rhf1d() does per column for a working directory with ACTIVE atoms
   (pyrh_compute1dray.c:112-388 with input.solve_NLTE, STOKES_MODE = NO_STOKES, CRD), on the device, for a batch:
       unit conversion, Bproject                                   pyrh_compute1dray.c:248-299
       Background(): SetLTEQuantities = LTEpops + CollisionRate (ltepops.c:224-249, collision.c:450-946),
                     ChemicalEquilibrium, continuum, passive_bb, rlk_opacity, MolecularOpacity -- for the LAST ray
                     (mu = Nrays-1, up): pyrh keeps one background record per wavelength (readj.c:319-345)
       convertScales                                               multiatmos.c:100-177
       getProfiles (Damping + Profile), initSolution (LTE_POPULATIONS, J = 0), initScatter
       Iterate, the N_MAX_SCATTER passes after it                  pyrh_compute1dray.c:330-337
       _solveray(): Bproject, Background, getProfiles for the one ray mu, solveSpectrum(FALSE, FALSE), packing
   Needs on the context: rhb200_set_lines, rhb200_set_passive_lines (lines of the PASSIVE atoms only),
   rhb200_set_model_lines (all atoms), rhb200_set_wavelengths(plan->lambda), rhb200_set_continuum with solve_NLTE = 1
   (levels of ACTIVE atoms flagged in lev[][4], their continua in bf[][9], H_active), rhb200_set_chemistry.
   Collisional records, one per line of the atom files' collisional sections, in file order: */
enum {
  RHB200_CO_ATOM = 0,        /* ACTIVE-atom index */
  RHB200_CO_TYPE,            /* RHB200_CO_OMEGA ... */
  RHB200_CO_I, RHB200_CO_J,  /* i < j, levels of that atom */
  RHB200_CO_NT, RHB200_CO_TOFF,   /* slice of coll_T / coll_coef / coll_M */
  RHB200_CO_DE,              /* E[j] - E[i] [J] */
  RHB200_CO_NFIELD = 8
};
enum { RHB200_CO_OMEGA = 0, RHB200_CO_CE, RHB200_CO_CI, RHB200_CO_CP, RHB200_CO_CH, RHB200_CO_CH0, RHB200_CO_CHPLUS };
typedef struct {
  const int    *atom_model;       /* [plan->Natom] index of each ACTIVE atom among the model atoms of rhb200_set_continuum */
  int ncoll, ncolltab;
  const double *coll;             /* [ncoll][RHB200_CO_NFIELD] */
  const double *coll_T, *coll_coef, *coll_M;   /* [ncolltab]: temperature grids, coefficients, spline second derivatives
                                     (splineCoef, spline.c:31-66; unused for grids of <= 2 points: Linear) */
  const double *line_rows;        /* [plan->nline][RHB200_PL_NFIELD]: Damping() constants of the ACTIVE lines, line-index order;
                                     RHB200_PL_LEVEL_I/J are rows of the model-atom level table */
  int NmaxScatter, NmaxIter;      /* keywords N_MAX_SCATTER, N_MAX_ITER */
  double iterLimit;               /* ITER_LIMIT */
  const rhb200_nlte_plan *plan1;  /* the same plan with Nrays = 1 and the profile rows of one ray (muz/wmu are set by the call) */
  /* STOKES_MODE (zero-initialise for NO_STOKES).  stokes = 1 is FIELD_FREE: initScatter and Iterate() run with field-free
     profiles exactly as in NO_STOKES; adjustStokesMode() (zeeman.c:303-345) then recomputes the profiles of the polarizable
     lines with their Zeeman patterns (Profile(), profile.c:112-305) and the passes after Iterate() and _solveray()'s pass
     solve all four Stokes parameters where the active set or the background holds a polarised line (formal.c:86-217,
     opacity.c:262-296, stokesopac.c:28-82).  stokes = 2 is FULL_STOKES: polarised profiles and rays from initScatter on,
     Gamma from I_eff = I + Q + U + V - Psi (eta + eta_Q + eta_U + eta_V) where the active set holds a polarizable
     line (fillgamma.c:106-129).  stokes = 3 is POLARIZATION_FREE: Zeeman-broadened profiles from the start (profile.c:112),
     scalar transfer during the iterations, the four Stokes parameters afterwards with the same profiles. */
  int stokes;
  const int    *line_pol;         /* [plan->nline] line->polarizable (readatom.c:352-368) */
  const int    *line_zoff;        /* [plan->nline + 1] slice of each line in the pattern tables (0 components if not polarizable) */
  const int    *zq;               /* Zeeman(line), zeeman.c:186-281 (rhb200_zeeman): q, shift, strength */
  const double *zshift, *zstrength;
  /* angle-averaged partial redistribution (zero-initialise for CRD): line->PRD of every ACTIVE line (readatom.c:255-258:
     shape PRD and PRD_N_MAX_ITER > 0) and the keywords PRD_N_MAX_ITER, PRD_ITER_LIMIT.  After updatePopulations() of
     every MALI iteration Redistribute() (redistribute.c:38-106) runs PRDScatter() (scatter.c:51-290: Gouttebroze's GII,
     linear interpolation of J, no cross redistribution) for each PRD line and solveSpectrum(FALSE, TRUE) over their
     wavelengths, per column until the profile ratio rho changes by less than the limit.  PRD_NG_ORDER > 0,
     PRD_ANGLE_DEP and XRD are not implemented. */
  const int    *line_prd;         /* [plan->nline] or NULL */
  int PRD_NmaxIter;
  double PRDiterLimit;
} rhb200_nlte_front;
/* atmosphere [ncol][nrow][ndep] as in rhb200_compute1d_batch.  Out (any may be NULL): spectrum [ncol][Nspect] = spectrum.I[][0]
   of the final pass on plan->lambda (lambda_ref included; _solveray drops it), pops_n / pops_nstar [ncol][sum Nlevel][ndep]
   = AtomPops.n / .nstar, niter [ncol] iterations Iterate() took, passes [ncol][2] = solveSpectrum(FALSE, FALSE) passes of
   initScatter and of the loop after Iterate(), scales [ncol][3][ndep] = height, tau_ref, column mass. */
int rhb200_nlte_compute1d_batch(rhb200_ctx *ctx, const rhb200_nlte_plan *plan, const rhb200_nlte_front *front,
                                int ncol, int ndep, int nrow, double mu, int atm_scale, const double *atmosphere,
                                int iref, double wght_per_H, double vmacro_tresh,
                                double *spectrum, double *pops_n, double *pops_nstar, int *niter, int *passes,
                                double *scales);
/* The same with the emergent Stokes Q, U, V of the final pass: quv [ncol][3][Nspect] (zeros where Formal() solved for I
   alone; all zeros unless front->stokes). */
int rhb200_nlte_compute1d_stokes_batch(rhb200_ctx *ctx, const rhb200_nlte_plan *plan, const rhb200_nlte_front *front,
                                       int ncol, int ndep, int nrow, double mu, int atm_scale, const double *atmosphere,
                                       int iref, double wght_per_H, double vmacro_tresh,
                                       double *spectrum, double *quv, double *pops_n, double *pops_nstar, int *niter, int *passes,
                                       double *scales);
/* test hook: the per-column inputs the front end hands to Iterate() for the columns of the LAST
   rhb200_nlte_compute1d_batch call that fit in one chunk: which = 0 C, 1 nstar, 2 ntotal, 3 adamp, 4 vbroad,
   5 chi_c, 6 eta_c, 7 sca_c, 8 height, 9 J after the last pass, 10..12 chi_c/eta_c/sca_c of the final pass,
   13..15 phi / wphi / adamp of the final pass; out
   must hold the array ([ncol][...][ndep] in the layouts of rhb200_nlte_columns). */
int rhb200_nlte_front_debug(rhb200_ctx *ctx, int which, double *out, size_t count);

/* SolveLinearEq (ludcmp.c:36-86) for nsys systems: A [nsys][N][N] (untouched), b [nsys][N] in/out; N <= 32 */
int rhb200_solve_linear_eq_batch(rhb200_ctx *ctx, int nsys, int N, double *A, double *b, int improve);

/* ---- device memory helpers (for callers that keep inputs resident) -------- */
int rhb200_dev_alloc(rhb200_ctx *ctx, size_t bytes, void **dptr);
int rhb200_dev_free(rhb200_ctx *ctx, void *dptr);
int rhb200_host_alloc_pinned(size_t bytes, void **hptr);
int rhb200_host_free_pinned(void *hptr);
int rhb200_memcpy_h2d(rhb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int rhb200_memcpy_d2h(rhb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int rhb200_synchronize(rhb200_ctx *ctx);
/* write `bytes` of a scratch buffer (> L2) to evict L2 between timed iterations */
int rhb200_flush_l2(rhb200_ctx *ctx);

/* ---- instrumentation ------------------------------------------------------
   CUDA-event time (ms) and launch count per kernel family accumulated since the
   last rhb200_timing_reset(); which: 0 prep, 1 line opacity, 2 DELO-Bezier3,
   3 scalar Bezier3, 4 other, 5 NLTE Gamma/rates, 6 NLTE J, 7 statEquil, 8 Ng.  Mirrors the getCPU() labels the reference stubs
   out (rh/getcpu.c:60). */
enum { RHB200_K_PREP = 0, RHB200_K_OPACITY, RHB200_K_DELO, RHB200_K_BEZIER, RHB200_K_OTHER,
       RHB200_K_GAMMA /* addtoGamma/addtoRates/addtoCoupling */, RHB200_K_J, RHB200_K_STATEQ, RHB200_K_NG,
       RHB200_K_COUNT };
int rhb200_timing_enable(rhb200_ctx *ctx, int on);
int rhb200_timing_reset(rhb200_ctx *ctx);
int rhb200_timing_get(rhb200_ctx *ctx, int which, double *ms, long *launches);
/* CUDA-event stopwatch on the context's compute stream (bench.py times K steps with it) */
int rhb200_timer_begin(rhb200_ctx *ctx);
int rhb200_timer_end(rhb200_ctx *ctx, double *ms);
/* FP64 pipe micro-benchmark: runs a dependent-free DFMA kernel, returns TFLOP/s
   (2 flop per FMA).  Used by bench.py as the measured FP64 roofline denominator. */
int rhb200_fp64_peak(rhb200_ctx *ctx, double *tflops_fma, double *tflops_nofma);

#ifdef __cplusplus
}
#endif
#endif /* RHB200_H */
