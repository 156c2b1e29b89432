/* oracle/port/rhport_nlte.c -- TEST INFRASTRUCTURE ONLY (see rhport.h).
 *
 * CPU restatement of the reference's NLTE (MALI) machinery for atoms, CRD, unpolarised:
 *   SolveLinearEq / LUdecomp / LUbacksubst   rh/ludcmp.c:36-177
 *   statEquil                                rh/statequil.c:40-103
 *   NgInit / Accelerate / MaxChange          rh/accelerate.c:34-147, rh/maxchange.c:32-50
 *   Opacity (active set)                     rh/opacity.c:64-390
 *   addtoCoupling / addtoGamma / addtoRates  rh/fillgamma.c:82-461
 *   Formal (scalar branches)                 rh/rhf1d/formal.c:44-346
 *   solveSpectrum / Iterate / updatePopulations   rh/iterate.c:48-253, rh/statequil.c:177-216
 * Operation order follows the reference expression by expression (build: -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "rhport.h"

/* ------------------------------------------------------------ ludcmp.c */
static void lu_decomp(int N, double *A, int *index)               /* ludcmp.c:92-150, A row-major */
{
  int i, j, k, imax = 0;
  double big, dum, sum, temp, *vv = malloc(N * sizeof(double));
  for (i = 0; i < N; i++) {
    big = 0.0;
    for (j = 0; j < N; j++) if ((temp = fabs(A[i*N+j])) > big) big = temp;
    vv[i] = 1.0 / big;
  }
  for (j = 0; j < N; j++) {
    for (i = 0; i < j; i++) {
      sum = A[i*N+j];
      for (k = 0; k < i; k++) sum -= A[i*N+k] * A[k*N+j];
      A[i*N+j] = sum;
    }
    big = 0.0;
    for (i = j; i < N; i++) {
      sum = A[i*N+j];
      for (k = 0; k < j; k++) sum -= A[i*N+k] * A[k*N+j];
      A[i*N+j] = sum;
      if ((dum = vv[i]*fabs(sum)) >= big) { big = dum; imax = i; }
    }
    if (j != imax) {
      for (k = 0; k < N; k++) { dum = A[imax*N+k]; A[imax*N+k] = A[j*N+k]; A[j*N+k] = dum; }
      vv[imax] = vv[j];
    }
    index[j] = imax;
    if (A[j*N+j] == 0.0) A[j*N+j] = 1.0e-20;
    dum = 1.0 / A[j*N+j];
    for (i = j+1; i < N; i++) A[i*N+j] *= dum;
  }
  free(vv);
}

static void lu_backsubst(int N, const double *A, const int *index, double *b)   /* ludcmp.c:156-177 */
{
  int i, j, ii = -1, ip;
  double sum;
  for (i = 0; i < N; i++) {
    ip = index[i];
    sum = b[ip];
    b[ip] = b[i];
    if (ii >= 0) { for (j = ii; j < i; j++) sum -= A[i*N+j] * b[j]; }
    else if (sum) ii = i;
    b[i] = sum;
  }
  for (i = N-1; i >= 0; i--) {
    sum = b[i];
    for (j = i+1; j < N; j++) sum -= A[i*N+j]*b[j];
    b[i] = sum / A[i*N+i];
  }
}

void rp_solve_linear_eq(int N, double *A, double *b, int improve)    /* ludcmp.c:36-86 */
{
  int i, j, *index = malloc(N * sizeof(int));
  double *A_copy = NULL, *b_copy = NULL, *residual = NULL;
  if (improve) {
    residual = malloc(N*sizeof(double)); b_copy = malloc(N*sizeof(double)); A_copy = malloc((size_t) N*N*sizeof(double));
    memcpy(b_copy, b, N*sizeof(double)); memcpy(A_copy, A, (size_t) N*N*sizeof(double));
  }
  lu_decomp(N, A, index);
  lu_backsubst(N, A, index, b);
  if (improve) {
    for (i = 0; i < N; i++) {
      residual[i] = b_copy[i];
      for (j = 0; j < N; j++) residual[i] -= A_copy[i*N+j] * b[j];
    }
    lu_backsubst(N, A, index, residual);
    for (i = 0; i < N; i++) b[i] += residual[i];
    free(residual); free(b_copy); free(A_copy);
  }
  free(index);
}

/* statEquil, statequil.c:40-103.  Gamma [Nl*Nl][N] (depth fastest), n [Nl][N] in/out */
void rp_stat_equil(int Nl, int N, const double *Gamma, const double *ntotal, int isum, double *n)
{
  int i, j, k, ie;
  double *Gk = malloc((size_t) Nl*Nl*sizeof(double)), *nk = malloc(Nl*sizeof(double)), nmax, GamDiag;
  for (k = 0; k < N; k++) {
    for (i = 0; i < Nl; i++) {
      nk[i] = n[i*N+k];
      for (j = 0; j < Nl; j++) Gk[i*Nl+j] = Gamma[(size_t)(i*Nl+j)*N + k];
    }
    if (isum == -1) {
      ie = 0; nmax = 0.0;
      for (i = 0; i < Nl; i++) if (nk[i] > nmax) { nmax = nk[i]; ie = i; }
    } else ie = isum;
    for (i = 0; i < Nl; i++) {
      GamDiag = 0.0; Gk[i*Nl+i] = 0.0; nk[i] = 0.0;
      for (j = 0; j < Nl; j++) GamDiag += Gk[j*Nl+i];
      Gk[i*Nl+i] = -GamDiag;
    }
    nk[ie] = ntotal[k];
    for (j = 0; j < Nl; j++) Gk[ie*Nl+j] = 1.0;
    rp_solve_linear_eq(Nl, Gk, nk, 1);
    for (i = 0; i < Nl; i++) n[i*N+k] = nk[i];
  }
  free(Gk); free(nk);
}

/* ------------------------------------------------------- accelerate.c */
typedef struct { int N, Ndelay, Norder, Nperiod, count; double *previous; } rp_ng;

static rp_ng *ng_init(int N, int Ndelay, int Norder, int Nperiod, const double *sol)   /* accelerate.c:34-62 */
{
  rp_ng *g = malloc(sizeof(rp_ng));
  g->N = N; g->Norder = Norder; g->Nperiod = Nperiod;
  g->Ndelay = (Ndelay > Norder + 2) ? Ndelay : Norder + 2;
  g->previous = calloc((size_t)(Norder+2)*N, sizeof(double));
  memcpy(g->previous, sol, N*sizeof(double));
  g->count = 1;
  return g;
}

static int ng_accelerate(rp_ng *g, double *sol)                      /* accelerate.c:68-147 */
{
  int i, j, k, Norder = g->Norder, N = g->N, ip, ipp, i0;
  i = g->count % (Norder + 2);
  memcpy(g->previous + (size_t) i*N, sol, N*sizeof(double));
  g->count++;
  if ((Norder > 0) && (g->count >= g->Ndelay) && !((g->count - g->Ndelay) % g->Nperiod)) {
    double *Delta = malloc((size_t)(Norder+1)*N*sizeof(double)), *weight = malloc(N*sizeof(double));
    double *A = calloc((size_t) Norder*Norder, sizeof(double)), *b = calloc(Norder, sizeof(double));
    for (i = 0; i <= Norder; i++) {
      ip  = (g->count - 1 - i) % (Norder + 2);
      ipp = (g->count - 2 - i) % (Norder + 2);
      for (k = 0; k < N; k++) Delta[(size_t) i*N+k] = g->previous[(size_t) ip*N+k] - g->previous[(size_t) ipp*N+k];
    }
    for (k = 0; k < N; k++) weight[k] = 1.0 / fabs(sol[k]);
    for (j = 0; j < Norder; j++) {
      for (k = 0; k < N; k++)
        b[j] += weight[k] * Delta[k]*(Delta[k] - Delta[(size_t)(j+1)*N+k]);
      for (i = 0; i < Norder; i++)
        for (k = 0; k < N; k++)
          A[i*Norder+j] += weight[k] * (Delta[(size_t)(j+1)*N+k] - Delta[k]) * (Delta[(size_t)(i+1)*N+k] - Delta[k]);
    }
    rp_solve_linear_eq(Norder, A, b, 1);
    i0 = (g->count - 1) % (Norder + 2);
    for (i = 0; i < Norder; i++) {
      ip = (g->count - 2 - i) % (Norder + 2);
      for (k = 0; k < N; k++)
        sol[k] += b[i] * (g->previous[(size_t) ip*N+k] - g->previous[(size_t) i0*N+k]);
    }
    memcpy(g->previous + (size_t) i0*N, sol, N*sizeof(double));
    free(Delta); free(weight); free(A); free(b);
    return 1;
  }
  return 0;
}

static double ng_maxchange(const rp_ng *g)                           /* maxchange.c:32-50 */
{
  double dmax = 0.0;
  if (g->count < 2) return dmax;
  const double *old = g->previous + (size_t)((g->count - 2) % (g->Norder + 2))*g->N;
  const double *new = g->previous + (size_t)((g->count - 1) % (g->Norder + 2))*g->N;
  for (int k = 0; k < g->N; k++)
    if (new[k]) { double d = fabs((new[k] - old[k]) / new[k]); if (d > dmax) dmax = d; }
  return dmax;
}

/* ------------------------------------------------------------- the MALI pass */
enum { TR_ATOM = 0, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA, TR_AJI, TR_BJI, TR_BIJ, TR_ISOFRAC,
       TR_WOFF, TR_PHIROW, TR_KR, TR_LINEIDX, TR_NFIELD = 16 };

typedef struct {
  int Nspect, Nrays, Ndep, Natom, Ntrans, moving, Ngorder, Ngdelay, Ngperiod, isum, bc_top, bc_bottom;
  const double *lambda, *muz, *wmu, *T, *height;
  const int *atom_nlevel;
  const double *trans, *tr_lambda, *tr_wlambda, *tr_alpha;
  const int *as_first, *as_trans, *bg_hasline;
  const double *nstar, *ntotal, *C, *phi, *wphi, *chi_c, *eta_c, *sca_c;
  double *n, *J;            /* state, in/out */
  double *Gamma, *Rij, *Rji;  /* [sum Nl^2][Ndep], [Ntrans][Ndep] work/output */
  int updateJ;                /* spectrum.updateJ */
  double *Iem;                /* [Nspect][Nrays] spectrum.I[nspect][mu] */
} rp_nlte;

#define MAXACT 64

/* one wavelength: Formal(nspect, eval_operator, FALSE), formal.c:44-346 (scalar, no PRD, no pol.) */
static double formal_lambda(const rp_nlte *P, int ns, int eval_operator,
                            const int *lev_off, const int *gam_off)
{
  const int N = P->Ndep, Nrays = P->Nrays;
  const double hc = RP_HPLANCK * RP_CLIGHT, fourPI = 4.0 * RP_PI, hc_4PI = hc / fourPI;
  const double nm3 = RP_NM_TO_M*RP_NM_TO_M*RP_NM_TO_M;
  const double twohc = 2.0*hc / nm3, hc_k = hc / (RP_KBOLTZMANN * RP_NM_TO_M);
  const int first = P->as_first[ns], nact = P->as_first[ns+1] - first;
  int n, k, mu, to_obs, m, boundbound = 0;
  double *Vij = malloc((size_t) MAXACT*N*sizeof(double)), *gij = malloc((size_t) MAXACT*N*sizeof(double)),
         *wla = malloc((size_t) MAXACT*N*sizeof(double));
  double *as_chi = malloc(N*sizeof(double)), *as_eta = malloc(N*sizeof(double));
  double *eta_atom = calloc((size_t) P->Natom*N, sizeof(double));
  int nlev_tot = lev_off[P->Natom];
  double *chi_up = calloc((size_t) nlev_tot*N, sizeof(double)), *chi_down = calloc((size_t) nlev_tot*N, sizeof(double)),
         *Uji_down = calloc((size_t) nlev_tot*N, sizeof(double));
  double *chi = malloc(N*sizeof(double)), *S = malloc(N*sizeof(double)), *I = malloc(N*sizeof(double)),
         *Psi = malloc(N*sizeof(double)), *Jdag = malloc(N*sizeof(double)), *Ieff = malloc(N*sizeof(double));
  double *J = P->J + (size_t) ns*N, twohnu3[MAXACT], dJmax = 0.0;
  const double *chi_c = P->chi_c + (size_t) ns*N, *eta_c = P->eta_c + (size_t) ns*N, *sca_c = P->sca_c + (size_t) ns*N;

  for (n = 0; n < nact; n++)
    if (P->trans[(size_t) P->as_trans[first+n]*TR_NFIELD + TR_TYPE] == 0) boundbound = 1;
  const int angle_dep = P->moving && (boundbound || P->bg_hasline[ns]);
  memcpy(Jdag, J, N*sizeof(double));
  if (P->updateJ) for (k = 0; k < N; k++) J[k] = 0.0;

  for (mu = 0; mu < Nrays; mu++) {
    for (to_obs = 0; to_obs <= (angle_dep ? 1 : 0); to_obs++) {
      const int initialize = (mu == 0 && to_obs == 0);
      const double wmu = angle_dep ? 0.5 * P->wmu[mu] : P->wmu[mu];
      if (initialize || (angle_dep && boundbound)) {                  /* Opacity(), opacity.c:64-390 */
        for (k = 0; k < N; k++) { as_chi[k] = 0.0; as_eta[k] = 0.0; }
        for (int a = 0; a < P->Natom; a++) {
          int any = 0;
          for (n = 0; n < nact; n++) if ((int) P->trans[(size_t) P->as_trans[first+n]*TR_NFIELD + TR_ATOM] == a) any = 1;
          if (any) for (k = 0; k < N; k++) eta_atom[(size_t) a*N+k] = 0.0;
        }
        for (n = 0; n < nact; n++) {
          const double *tr = P->trans + (size_t) P->as_trans[first+n]*TR_NFIELD;
          const int a = (int) tr[TR_ATOM], i = (int) tr[TR_I], j = (int) tr[TR_J], la = ns - (int) tr[TR_NBLUE];
          const double *n_i = P->n + (size_t)(lev_off[a] + i)*N, *n_j = P->n + (size_t)(lev_off[a] + j)*N;
          double *V = Vij + (size_t) n*N, *g = gij + (size_t) n*N, *w = wla + (size_t) n*N;
          if (tr[TR_TYPE] == 0) {
            const int lamu = 2*(Nrays*la + mu) + to_obs;
            const double *phi = P->phi + (size_t)((int) tr[TR_PHIROW] + lamu)*N;
            const double gijk = tr[TR_BJI] / tr[TR_BIJ], Bijxhc_4PI = hc_4PI * tr[TR_BIJ] * tr[TR_ISOFRAC];
            twohnu3[n] = tr[TR_AJI] / tr[TR_BJI];
            for (k = 0; k < N; k++) { g[k] = gijk; V[k] = Bijxhc_4PI * phi[k]; }
            if (initialize) {
              const double wlambda = P->tr_wlambda[(int) tr[TR_WOFF] + la];
              const double *wphi = P->wphi + (size_t)((int) tr[TR_LINEIDX])*N;
              for (k = 0; k < N; k++) w[k] = wlambda * wphi[k] / hc_4PI;
            }
          } else {
            const double lc = P->tr_lambda[(int) tr[TR_WOFF] + la];
            twohnu3[n] = twohc / (lc*lc*lc);
            if (initialize) {
              const double wlambda = P->tr_wlambda[(int) tr[TR_WOFF] + la], al = P->tr_alpha[(int) tr[TR_WOFF] + la];
              const double *ns_i = P->nstar + (size_t)(lev_off[a] + i)*N, *ns_j = P->nstar + (size_t)(lev_off[a] + j)*N;
              for (k = 0; k < N; k++) {
                V[k] = al;
                g[k] = ns_i[k] / ns_j[k] * exp(-hc_k / (lc * P->T[k]));
                w[k] = fourPI/RP_HPLANCK * (wlambda/lc);
              }
            }
          }
          if (twohnu3[n]) {
            double *ea = eta_atom + (size_t) a*N;
            for (k = 0; k < N; k++) {
              as_chi[k] += V[k] * (n_i[k] - g[k]*n_j[k]);
              ea[k] += twohnu3[n] * g[k] * V[k] * n_j[k];
            }
          }
        }
        for (int a = 0; a < P->Natom; a++) {
          int any = 0;
          for (n = 0; n < nact; n++) if ((int) P->trans[(size_t) P->as_trans[first+n]*TR_NFIELD + TR_ATOM] == a) any = 1;
          if (any) for (k = 0; k < N; k++) as_eta[k] += eta_atom[(size_t) a*N+k];
        }
      }
      if (eval_operator) {                                            /* addtoCoupling, fillgamma.c:251-332 */
        for (n = 0; n < nact; n++) {
          const double *tr = P->trans + (size_t) P->as_trans[first+n]*TR_NFIELD;
          const int a = (int) tr[TR_ATOM];
          memset(chi_up + (size_t)(lev_off[a] + (int) tr[TR_I])*N, 0, N*sizeof(double));
          memset(chi_down + (size_t)(lev_off[a] + (int) tr[TR_J])*N, 0, N*sizeof(double));
          memset(Uji_down + (size_t)(lev_off[a] + (int) tr[TR_J])*N, 0, N*sizeof(double));
        }
        for (n = 0; n < nact; n++) {
          const double *tr = P->trans + (size_t) P->as_trans[first+n]*TR_NFIELD;
          const int a = (int) tr[TR_ATOM], i = (int) tr[TR_I], j = (int) tr[TR_J];
          const double *n_i = P->n + (size_t)(lev_off[a] + i)*N, *n_j = P->n + (size_t)(lev_off[a] + j)*N;
          const double *V = Vij + (size_t) n*N, *g = gij + (size_t) n*N, *w = wla + (size_t) n*N;
          double *cu = chi_up + (size_t)(lev_off[a] + i)*N, *cd = chi_down + (size_t)(lev_off[a] + j)*N,
                 *ud = Uji_down + (size_t)(lev_off[a] + j)*N;
          if (twohnu3[n])
            for (k = 0; k < N; k++) {
              const double chicc = V[k] * w[k] * (n_i[k] - g[k]*n_j[k]);
              cu[k] += chicc; cd[k] += chicc;
              ud[k] += twohnu3[n] * g[k] * V[k];
            }
        }
      }
      for (k = 0; k < N; k++) {                                       /* formal.c:178-182 / :293-296 */
        chi[k] = as_chi[k] + chi_c[k];
        S[k] = (as_eta[k] + eta_c[k] + sca_c[k]*Jdag[k]) / chi[k];
      }
      if (angle_dep) {
        rp_bezier3_scalar(N, P->height, P->muz[mu], to_obs, chi, S, P->T, P->lambda[ns], P->bc_top, P->bc_bottom,
                          I, eval_operator ? Psi : NULL);
        if (P->Iem) P->Iem[(size_t) ns*Nrays + mu] = I[0];            /* formal.c:270 (last ray = up-ray) */
      } else {
        double Iplus = rp_feautrier(N, P->height, P->muz[mu], chi, S, P->T, P->lambda[ns], P->bc_top, P->bc_bottom,
                                    I, eval_operator ? Psi : NULL);
        if (P->Iem) P->Iem[(size_t) ns*Nrays + mu] = Iplus;           /* formal.c:299 */
      }
      if (eval_operator) {                                            /* addtoGamma, fillgamma.c:82-246 */
        for (k = 0; k < N; k++) Psi[k] /= chi[k];
        for (n = 0; n < nact; n++) {
          const double *tr = P->trans + (size_t) P->as_trans[first+n]*TR_NFIELD;
          const int a = (int) tr[TR_ATOM], i = (int) tr[TR_I], j = (int) tr[TR_J], Nl = P->atom_nlevel[a];
          const double *V = Vij + (size_t) n*N, *g = gij + (size_t) n*N, *w = wla + (size_t) n*N;
          const double *ea = eta_atom + (size_t) a*N;
          double *Gij = P->Gamma + (size_t)(gam_off[a] + i*Nl + j)*N, *Gji = P->Gamma + (size_t)(gam_off[a] + j*Nl + i)*N;
          for (k = 0; k < N; k++) Ieff[k] = I[k] - Psi[k] * ea[k];
          for (k = 0; k < N; k++) {
            const double wlamu = V[k] * w[k] * wmu;
            Gji[k] += Ieff[k] * wlamu;
            Gij[k] += (twohnu3[n] + Ieff[k]) * g[k] * wlamu;
          }
          { const double *cu = chi_up + (size_t)(lev_off[a] + i)*N, *ud = Uji_down + (size_t)(lev_off[a] + j)*N;
            for (k = 0; k < N; k++) Gij[k] -= cu[k] * Psi[k]*ud[k] * wmu; }
          for (m = 0; m < nact; m++) {
            const double *tm = P->trans + (size_t) P->as_trans[first+m]*TR_NFIELD;
            if ((int) tm[TR_ATOM] == a && (int) tm[TR_J] == i) {
              const double *cd = chi_down + (size_t)(lev_off[a] + j)*N, *ud = Uji_down + (size_t)(lev_off[a] + i)*N;
              for (k = 0; k < N; k++) Gji[k] += cd[k] * Psi[k]*ud[k] * wmu;
            }
          }
        }
      }
      if (!P->updateJ) continue;
      for (k = 0; k < N; k++) J[k] += wmu * I[k];                     /* formal.c:254-256 / :306 */
      for (n = 0; n < nact; n++) {                                    /* addtoRates, fillgamma.c:375-461 */
        const int t = P->as_trans[first+n];
        const double *V = Vij + (size_t) n*N, *g = gij + (size_t) n*N, *w = wla + (size_t) n*N;
        double *Rij = P->Rij + (size_t) t*N, *Rji = P->Rji + (size_t) t*N;
        for (k = 0; k < N; k++) {
          const double wlamu = V[k] * w[k] * wmu;
          Rij[k] += I[k] * wlamu;
          Rji[k] += g[k] * (twohnu3[n] + I[k]) * wlamu;
        }
      }
    }
  }
  if (P->updateJ)
    for (k = 0; k < N; k++) { double dJ = fabs(1.0 - Jdag[k]/J[k]); if (dJ > dJmax) dJmax = dJ; }
  free(Vij); free(gij); free(wla); free(as_chi); free(as_eta); free(eta_atom); free(chi_up); free(chi_down);
  free(Uji_down); free(chi); free(S); free(I); free(Psi); free(Jdag); free(Ieff);
  return dJmax;
}

static void offsets(const rp_nlte *P, int *lev_off, int *gam_off)
{
  lev_off[0] = gam_off[0] = 0;
  for (int a = 0; a < P->Natom; a++) {
    lev_off[a+1] = lev_off[a] + P->atom_nlevel[a];
    gam_off[a+1] = gam_off[a] + P->atom_nlevel[a]*P->atom_nlevel[a];
  }
}

/* solveSpectrum(eval_operator, FALSE), iterate.c:148-253 */
double rp_nlte_solve_spectrum(rp_nlte *P, int eval_operator)
{
  int lev_off[16], gam_off[16];
  double dJmax = 0.0;
  offsets(P, lev_off, gam_off);
  memset(P->Rij, 0, (size_t) P->Ntrans*P->Ndep*sizeof(double));     /* zeroRates */
  memset(P->Rji, 0, (size_t) P->Ntrans*P->Ndep*sizeof(double));
  for (int ns = 0; ns < P->Nspect; ns++) {
    double dJ = formal_lambda(P, ns, eval_operator, lev_off, gam_off);
    if (dJ > dJmax) dJmax = dJ;
  }
  return dJmax;
}

/* Iterate(NmaxIter, iterLimit), iterate.c:48-143.  n_hist [NmaxIter][nlev_tot*Ndep] and
   gamma_hist [NmaxIter][sum Nl^2 * Ndep] may be NULL; returns the number of iterations done */
int rp_nlte_iterate(rp_nlte *P, int NmaxIter, double iterLimit, double *n_hist, double *gamma_hist,
                    double *rates_hist, double *dpops_hist)
{
  int lev_off[16], gam_off[16], a, niter = 1, done = 0;
  offsets(P, lev_off, gam_off);
  const int N = P->Ndep;
  const size_t nlev = (size_t) lev_off[P->Natom]*N, ngam = (size_t) gam_off[P->Natom]*N;
  rp_ng **ng = malloc(P->Natom*sizeof(rp_ng *));
  for (a = 0; a < P->Natom; a++)
    ng[a] = ng_init(P->atom_nlevel[a]*N, P->Ngdelay, P->Ngorder, P->Ngperiod, P->n + (size_t) lev_off[a]*N);
  while (niter <= NmaxIter) {
    memcpy(P->Gamma, P->C, ngam*sizeof(double));                    /* initGammaAtom */
    rp_nlte_solve_spectrum(P, 1);
    if (gamma_hist) memcpy(gamma_hist + (size_t)(niter-1)*ngam, P->Gamma, ngam*sizeof(double));
    if (rates_hist) {
      memcpy(rates_hist + (size_t)(niter-1)*2*P->Ntrans*N, P->Rij, (size_t) P->Ntrans*N*sizeof(double));
      memcpy(rates_hist + (size_t)(niter-1)*2*P->Ntrans*N + (size_t) P->Ntrans*N, P->Rji, (size_t) P->Ntrans*N*sizeof(double));
    }
    double dpopsmax = 0.0;                                          /* updatePopulations */
    for (a = 0; a < P->Natom; a++) {
      double *na = P->n + (size_t) lev_off[a]*N;
      rp_stat_equil(P->atom_nlevel[a], N, P->Gamma + (size_t) gam_off[a]*N, P->ntotal + (size_t) a*N, P->isum, na);
      ng_accelerate(ng[a], na);
      double d = ng_maxchange(ng[a]);
      if (d > dpopsmax) dpopsmax = d;
    }
    if (n_hist) memcpy(n_hist + (size_t)(niter-1)*nlev, P->n, nlev*sizeof(double));
    if (dpops_hist) dpops_hist[niter-1] = dpopsmax;
    done = niter;
    if (dpopsmax < iterLimit) break;
    niter++;
  }
  for (a = 0; a < P->Natom; a++) { free(ng[a]->previous); free(ng[a]); }
  free(ng);
  return done;
}

/* Profile() for one un-polarised, single-component line in a moving atmosphere, profile.c:67-372
   (field-free branch :311-322, normalisation :323,358): phi [2*Nrays*Nlambda][N], wphi [N] */
double rp_voigt_armstrong(double a, double v);
void rp_profile_line(int N, int Nrays, int Nlambda, double lambda0, const double *lambda, const double *wlambda,
                     const double *adamp, const double *vbroad, const double *vel, const double *muz,
                     const double *wmu, double *phi, double *wphi)
{
  int la, mu, to_obs, k;
  for (k = 0; k < N; k++) wphi[k] = 0.0;
  for (la = 0; la < Nlambda; la++) {
    for (mu = 0; mu < Nrays; mu++) {
      const double wlamu = wlambda[la] * 0.5*wmu[mu];
      for (to_obs = 0; to_obs <= 1; to_obs++) {
        const double sign = to_obs ? 1.0 : -1.0;
        double *p = phi + (size_t)(2*(Nrays*la + mu) + to_obs)*N;
        for (k = 0; k < N; k++) {
          const double v = (lambda[la] - lambda0 - 0.0) * RP_CLIGHT / (vbroad[k] * lambda0);
          const double v_los = (muz[mu] * vel[k]) / vbroad[k];
          const double vk = v + sign * v_los;
          p[k] = 0.0 + rp_voigt_armstrong(adamp[k], vk) * 1.0 / (RP_SQRTPI * vbroad[k]);
          wphi[k] += p[k] * wlamu;
        }
      }
    }
  }
  for (k = 0; k < N; k++) wphi[k] = 1.0 / wphi[k];
}
