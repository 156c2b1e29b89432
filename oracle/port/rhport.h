/* oracle/port/rhport.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's LTE/NLTE hot path (dvukadinovic/pyrh,
 * RH 1-D), written from the algorithm with the same arithmetic order so it is
 * bit-comparable with the compiled reference (oracle/_ref).  Every function
 * cites the reference file:line it follows.  It is pinned against the
 * reference itself by tests/test_oracle_port.py (golden dumps in tests/golden/
 * produced by oracle/gen_golden.py from oracle/_ref/liboracle_scalar.so).
 *
 * Nothing in pyrh_b200/ may include, link or call this.
 */
#ifndef RHPORT_H
#define RHPORT_H

#ifdef __cplusplus
extern "C" {
#endif

/* physical constants: rh/constant.h:28-73 */
#define RP_CLIGHT      2.99792458E+08
#define RP_HPLANCK     6.6260755E-34
#define RP_KBOLTZMANN  1.380658E-23
#define RP_AMU         1.6605402E-27
#define RP_M_ELECTRON  9.1093897E-31
#define RP_Q_ELECTRON  1.60217733E-19
#define RP_EPSILON_0   8.854187817E-12
#define RP_NM_TO_M     1.0E-09
#define RP_PI          3.14159265358979
#define RP_SQRTPI      1.77245385090551
#define RP_LARMOR      (RP_Q_ELECTRON / (4.0*RP_PI*RP_M_ELECTRON)) * RP_NM_TO_M
#define RP_Q_WING            20.0   /* kurucz.c:89 */
#define RP_MAX_GAUSS_DOPPLER 7.0    /* kurucz.c:92 */

/* line-table row layout (doubles) -- mirrors the fields of RLK_Line, atom.h:156-167 */
enum { RL_LAMBDA0 = 0, RL_GI, RL_GJ, RL_EI, RL_EJ, RL_BJI, RL_AJI, RL_BIJ,
       RL_GRAD, RL_GSTARK, RL_GVDW, RL_HFS_FRAC, RL_ISO_FRAC, RL_CROSS,
       RL_ALPHA, RL_POLARIZABLE, RL_VDWAALS, RL_ELEM, RL_STAGE, RL_ZOFF,
       RL_NCOMP, RL_NFIELD = 24 };
/* element-table row layout (doubles) -- Element, atom.h:139-145 */
enum { RE_WEIGHT = 0, RE_ABUND, RE_NSTAGE, RE_PFROW, RE_IONPOT0, RE_MAXSTAGE = 12,
       RE_NFIELD = RE_IONPOT0 + RE_MAXSTAGE };
enum { RP_UNSOLD = 0, RP_RIDDER = 1, RP_BARKLEM = 2, RP_KURUCZ = 3 };   /* atom.h:32 */
enum { RP_IRRADIATED = 0, RP_ZERO = 1, RP_THERMALIZED = 2 };            /* geometry.h:13 */

typedef struct {
  int nline, nelem, npf, matinv_simd;
  const double *lines;      /* [nline][RL_NFIELD], sorted ascending lambda0 */
  const int    *zq;         /* Zeeman components, concatenated */
  const double *zshift, *zstrength;
  const double *elems;      /* [nelem][RE_NFIELD] */
  const double *pf;         /* [rows][npf] ln U(T) table (abundance.c:161-203) */
  const double *Tpf;        /* [npf] */
  double vmicro_char;       /* keyword VMICRO_CHAR [m/s] */
  int magneto_optical;      /* must be 0 */
} rp_linetable;

typedef struct {
  int Ndep;
  const double *T, *ne, *vturb, *vel, *B, *cos_gamma, *cos_2chi, *sin_2chi,
               *nHtot, *np, *height;   /* SI units, as the reference holds them in Formal() */
  double muz;
  int moving;
} rp_column;

/* voigt.c:381-419 + humlicek.c + complex.c */
double rp_voigt_humlicek(double a, double v, double *F);
int    rp_humlicek_region(double a, double v);

/* ltepops.c:116-159 : n[stage][k] for one element */
void rp_ltepops_elem(const rp_linetable *lt, int ielem, const rp_column *col, double *n /*[nstage][Ndep]*/);
/* linear.c:22-58 */
void rp_linear(int Ntable, const double *xt, const double *yt, int N, const double *x, double *y);

/* kurucz.c:511-725 (+ RLKProfile :729-828) for one wavelength, one direction.
   elem_n: [nelem][RE_MAXSTAGE][Ndep].  chi, eta: [4][Ndep] zeroed+accumulated as the reference.
   returns flags: bit0 hasline, bit1 ispolarized */
int rp_rlk_opacity(const rp_linetable *lt, const rp_column *col, const double *elem_n,
                   double lambda, int to_obs, double *chi, double *eta);

/* bezier_1D.c:52-300 (+ stokesopac.c, bezier_aux.c, w3.c, planck.c) */
void rp_stokes_bezier3(int Ndep, const double *height, double muz, int to_obs,
                       const double *chi, const double *S /*[4][Ndep]*/,
                       const double *chiQUV /*[3][Ndep] summed numerators of K'*/,
                       const double *T, double lambda, int bc_top, int bc_bottom,
                       int matinv_simd, double *I /*[4][Ndep]*/, double *Psi /*[Ndep] or NULL*/);
void rp_matinv_scalar(float *m);
void rp_matinv_simd(float *m);
void rp_bezier3_coeffs(double dt, double *alpha, double *beta, double *gamma, double *theta, double *eps);
double rp_cent_deriv(double dsup, double dsdn, double chiup, double chic, double chidn);
void rp_w3(double dtau, double *w);
double rp_planck(double T, double lambda);

/* bezier_1D.c:306-541 scalar cubic Bezier (no RF) */
void rp_bezier3_scalar(int Ndep, const double *height, double muz, int to_obs,
                       const double *chi, const double *S, const double *T, double lambda,
                       int bc_top, int bc_bottom, double *I, double *Psi);

/* piecewise_1D.c:44-253, piecestokes_1D.c:49-174, w3.c:25 (rhport_piecewise.c) */
void rp_w2(double dtau, double *w);
void rp_piecewise_linear(int Ndep, const double *height, double muz, int to_obs,
                         const double *chi, const double *S, const double *T, double lambda,
                         int bc_top, int bc_bottom, double *I, double *Psi);
void rp_piecewise_parabolic(int Ndep, const double *height, double muz, int to_obs,
                            const double *chi, const double *S, const double *T, double lambda,
                            int bc_top, int bc_bottom, double *I, double *Psi);
void rp_stokes_parabolic(int Ndep, const double *height, double muz, int to_obs,
                         const double *chi_I, const double *S, const double *chiQUV,
                         const double *T, double lambda, int bc_top, int bc_bottom,
                         double *I, double *Psi);

/* the same with the log gf response function of the up-ray (bezier_1D.c:416-428, 477-490, 509-516) */
#define RP_MAXPAR 16
void rp_bezier3_scalar_rf(int Ndep, const double *height, double muz, int to_obs,
                          const double *chi, const double *S, const double *T, double lambda,
                          int bc_top, int bc_bottom, double *I, double *Psi,
                          int npar, const double *dchi, const double *deta, double *dI);

/* feautrier.c:56-202 (STANDARD order); returns emergent I */
double rp_feautrier(int Ndep, const double *height, double muz, const double *chi, const double *S,
                    const double *T, double lambda, int bc_top, int bc_bottom, double *P, double *Psi);

/* formal.c:157-275 restricted to LTE (solve_NLTE = FALSE, Nrays = 1, emergent up-ray only):
   cont: chi_ai, eta_ai [Nlambda][Ndep] (angle-independent background, background.c:343-465)
   out:  stokes [4][Nlambda] */
void rp_lte_stokes_column(const rp_linetable *lt, const rp_column *col,
                          int Nlambda, const double *lambda,
                          const double *chi_ai, const double *eta_ai,
                          int bc_top, int bc_bottom, double *stokes);

/* voigt.c:126-243 */
double rp_voigt_armstrong(double a, double v);
int    rp_armstrong_region(double a, double v);

/* opacity.c:711-916 (rhport_molecules.c) */
int rp_molecular_opacity(int N, int nmol, int nmline, const double *mlines,
                         const int *zq, const double *zshift, const double *zstrength,
                         double vmicro_char, double lambda, double muz, int moving, int to_obs,
                         const double *T, const double *vel, const double *B,
                         const double *cos_gamma, const double *cos_2chi, const double *sin_2chi,
                         const double *mol, double *chi, double *eta);

/* metal.c:174-344 (rhport_molecules.c) */
int rp_passive_bb(int N, int nline, const double *plines, const double *c_shift, const double *c_fraction,
                  double vmicro_char, double lambda, double muz, int moving, int to_obs,
                  const double *vel, const double *pcol, double *chi, double *eta);

/* ---- NLTE (rhport_nlte.c): see the struct there; driven from oracle/portdriver.py ---- */
void rp_solve_linear_eq(int N, double *A /*row-major, destroyed*/, double *b, int improve);
void rp_stat_equil(int Nl, int N, const double *Gamma, const double *ntotal, int isum, double *n);

#ifdef __cplusplus
}
#endif
#endif
