// rhb200_delo.cuh -- cubic DELO-Bezier polarised ray integrator, one thread per ray.
//
// Reference: Piece_Stokes_Bezier3_1D rh/rhf1d/bezier_1D.c:52-300, with
//   StokesK           rh/stokesopac.c:28-87   (MAGNETO_OPTICAL = FALSE)
//   cent_deriv(_mat/_vec), m4m, m4v, MatInv (scalar variant), Bezier3_coeffs
//                     rh/bezier_aux.c:34-116, 235-328, 333-359
//   w3                rh/w3.c:43-63
//   Planck            rh/planck.c:38-66
//
// B200 design notes:
//  * With magneto-optical terms off K' has only its first row/column non-zero,
//    K' = [[0,q,u,v],[q,0,0,0],[u,0,0,0],[v,0,0,0]].  The reference multiplies and
//    differentiates full 4x4 matrices; here only the 3 independent entries are kept
//    in registers (K, dK: 15 doubles instead of 80) and the matrix products are
//    written out.  Skipped terms are exact zeros (x*0, s+0), and the surviving terms
//    are evaluated in the reference's order, so every result rounds identically.
//  * cent_deriv(dsup,dsdn,a,b,c) needs f_{i-1} = (b-a)/dsup, which is bit-identical
//    to f_i of the previous depth step: it is carried instead of recomputed
//    (saves 8 of 23 FP64 divisions per ray-point).
//  * Md is formed in double, rounded to float and inverted with the reference's
//    scalar Cramer sequence in float (every op rounds to float; 1.0/det in double).
#pragma once
#include "rhb200_common.cuh"
#include "rhb200_math.cuh"
#include "rhb200_div.cuh"

namespace rhd {

#define RH_MAX0(x) (((x) > (0.0) ? (x) : (0.0)))   // MAX(x, 0.0), rh.h:36

__device__ __forceinline__ double planck(double T, double lambda) {   // planck.c:38-66
  const double hc_kla = (RH_HPLANCK * RH_CLIGHT) / (RH_KBOLTZMANN * RH_NM_TO_M * lambda);
  const double l = RH_NM_TO_M * lambda;
  const double twohnu3_c2 = (2.0*RH_HPLANCK*RH_CLIGHT) / (l*l*l);
  const double hc_Tkla = hc_kla / T;
  return (hc_Tkla <= 150.0) ? twohnu3_c2 / (rhm::rh_exp(hc_Tkla) - 1.0) : 0.0;
}

__device__ __forceinline__ void w3(double dtau, double &w0, double &w1) {   // w3.c:43-63 (w[2] unused here)
  if (dtau < 5.0E-4) {
    w0 = dtau*(1.0 - 0.5*dtau);
    w1 = (dtau*dtau)*(0.5 - dtau/3.0);
  } else if (dtau > 50.0) {
    w0 = w1 = 1.0;
  } else {
    const double e = rhm::rh_exp(-dtau);
    w0 = 1.0 - e;
    w1 = w0 - dtau*e;
  }
}

__device__ __forceinline__ void bezier3_coeffs(double dt, double &alpha, double &beta,
                                               double &gamma, double &theta, double &eps) {
  const double dt2 = dt*dt;                      // bezier_aux.c:333-359
  double dt3 = dt2*dt;
  if (dt >= 5.e-2) {
    eps = rhm::rh_exp(-dt);
    const rhdiv::Recip rdt3(dt3);                // x/dt3 and 1.0/dt3 share one reciprocal refinement
    alpha = rdt3.div(-6.0 + 6.0*dt - 3.0*dt2 + dt3 + 6.0*eps);
    dt3 = rdt3.div(1.0);
    beta  = (6.0 + (-6.0 - dt*(6.0 + dt*(3.0 + dt)))*eps) * dt3;
    gamma = 3.0 * (6.0 + (-4.0 + dt)*dt - 2.0*(3.0 + dt)*eps) * dt3;
    theta = 3.0 * (eps*(6.0 + dt2 + 4.0*dt) + 2.0*dt - 6.0) * dt3;
  } else {
    const double dt4 = dt2*dt2;
    eps   = 1.0 - dt + 0.5*dt2 - dt3/6.0 + dt4/24.0;
    alpha = 0.25*dt - 0.05*dt2 + dt3/120.0 - dt4/840.0;
    beta  = 0.25*dt - 0.20*dt2 + dt3/12.0  - dt4/42.0;
    gamma = 0.25*dt - 0.10*dt2 + dt3*0.025 - dt4/210.0;
    theta = 0.25*dt - 0.15*dt2 + dt3*0.05  - dt4/84.0;
  }
}

// Fritsch-Butland derivative given the two one-sided slopes (bezier_aux.c:34-51);
// ca = 0.333..*(1 + dsdn/(dsdn+dsup)) is shared by all quantities of a depth step
__device__ __forceinline__ double fb_deriv(double fim1, double fi, double ca) {
  return (fim1*fi > 0) ? (fim1*fi) / ((1.0 - ca)*fim1 + ca*fi) : 0.0;
}
__device__ __forceinline__ double fb_alpha(double dsup, double dsdn) {
  return 0.333333333333333333333333 * (1.0 + dsdn / (dsdn + dsup));
}

// 4x4 inverse in float, scalar variant of MatInv (bezier_aux.c:235-328): Cramer's rule on
// the transposed matrix, each operation rounded to float, reciprocal of det in double.
__device__ __forceinline__ void matinv_f32(const float m[16], float d[16]) {
  float s[16], t[12];
#pragma unroll
  for (int i = 0; i < 4; i++) { s[i] = m[i*4]; s[i+4] = m[i*4+1]; s[i+8] = m[i*4+2]; s[i+12] = m[i*4+3]; }
  t[0] = s[10]*s[15]; t[1] = s[11]*s[14]; t[2] = s[9]*s[15];  t[3] = s[11]*s[13];
  t[4] = s[9]*s[14];  t[5] = s[10]*s[13]; t[6] = s[8]*s[15];  t[7] = s[11]*s[12];
  t[8] = s[8]*s[14];  t[9] = s[10]*s[12]; t[10] = s[8]*s[13]; t[11] = s[9]*s[12];
  d[0]  = t[0]*s[5] + t[3]*s[6] + t[4]*s[7];   d[0] -= t[1]*s[5] + t[2]*s[6] + t[5]*s[7];
  d[1]  = t[1]*s[4] + t[6]*s[6] + t[9]*s[7];   d[1] -= t[0]*s[4] + t[7]*s[6] + t[8]*s[7];
  d[2]  = t[2]*s[4] + t[7]*s[5] + t[10]*s[7];  d[2] -= t[3]*s[4] + t[6]*s[5] + t[11]*s[7];
  d[3]  = t[5]*s[4] + t[8]*s[5] + t[11]*s[6];  d[3] -= t[4]*s[4] + t[9]*s[5] + t[10]*s[6];
  d[4]  = t[1]*s[1] + t[2]*s[2] + t[5]*s[3];   d[4] -= t[0]*s[1] + t[3]*s[2] + t[4]*s[3];
  d[5]  = t[0]*s[0] + t[7]*s[2] + t[8]*s[3];   d[5] -= t[1]*s[0] + t[6]*s[2] + t[9]*s[3];
  d[6]  = t[3]*s[0] + t[6]*s[1] + t[11]*s[3];  d[6] -= t[2]*s[0] + t[7]*s[1] + t[10]*s[3];
  d[7]  = t[4]*s[0] + t[9]*s[1] + t[10]*s[2];  d[7] -= t[5]*s[0] + t[8]*s[1] + t[11]*s[2];
  t[0] = s[2]*s[7];  t[1] = s[3]*s[6];  t[2] = s[1]*s[7];  t[3] = s[3]*s[5];
  t[4] = s[1]*s[6];  t[5] = s[2]*s[5];  t[6] = s[0]*s[7];  t[7] = s[3]*s[4];
  t[8] = s[0]*s[6];  t[9] = s[2]*s[4];  t[10] = s[0]*s[5]; t[11] = s[1]*s[4];
  d[8]  = t[0]*s[13] + t[3]*s[14] + t[4]*s[15];   d[8]  -= t[1]*s[13] + t[2]*s[14] + t[5]*s[15];
  d[9]  = t[1]*s[12] + t[6]*s[14] + t[9]*s[15];   d[9]  -= t[0]*s[12] + t[7]*s[14] + t[8]*s[15];
  d[10] = t[2]*s[12] + t[7]*s[13] + t[10]*s[15];  d[10] -= t[3]*s[12] + t[6]*s[13] + t[11]*s[15];
  d[11] = t[5]*s[12] + t[8]*s[13] + t[11]*s[14];  d[11] -= t[4]*s[12] + t[9]*s[13] + t[10]*s[14];
  d[12] = t[2]*s[10] + t[5]*s[11] + t[1]*s[9];    d[12] -= t[4]*s[11] + t[0]*s[9] + t[3]*s[10];
  d[13] = t[8]*s[11] + t[0]*s[8] + t[7]*s[10];    d[13] -= t[6]*s[10] + t[9]*s[11] + t[1]*s[8];
  d[14] = t[6]*s[9] + t[11]*s[11] + t[3]*s[8];    d[14] -= t[10]*s[11] + t[2]*s[8] + t[7]*s[9];
  d[15] = t[10]*s[10] + t[4]*s[8] + t[9]*s[9];    d[15] -= t[8]*s[9] + t[11]*s[10] + t[5]*s[8];
  float det = s[0]*d[0] + s[1]*d[1] + s[2]*d[2] + s[3]*d[3];
  det = (float) (1.0 / (double) det);
#pragma unroll
  for (int j = 0; j < 16; j++) d[j] *= det;
}

// c = Minv(float) * b, m4v of bezier_aux.c:109-116
__device__ __forceinline__ void m4v(const float a[16], const double b[4], double c[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) s += ((double) a[i*4+k]) * b[k];
    c[i] = s;
  }
}

// IO policy concept:
//   double chi(int k); void K(int k, double x[3]);   K'[0][1..3] at depth k (already / chi_I)
//   void S(int k, double s[4]);
//   void storeI(int k, const double I[4]); void storePsi(int k, double psi);
// MODE 1 additionally saves, MODE 2 starts from, the complete state of the sweep on entry to the step that arrives at
// depth k (DELO_NSTATE doubles at state[k*DELO_NSTATE]): a ray whose records differ from a saved ray's only at depth kp
// is bit-identical to it up to the step arriving at kp + 2 (to_obs; its stencil reaches chi at k - 2) and can resume there
// (finite-difference response functions, rhb200_rf_fd_batch).
#define DELO_NSTATE 44
template <class IO, int MODE = 0>
__device__ __forceinline__ void delo_bezier3_ray(IO &io, const int ndep, const double *__restrict__ z,
                                                 const double muz, const int to_obs,
                                                 const int bc_top, const int bc_bottom,
                                                 const double *__restrict__ T, const double lambda,
                                                 double *state = nullptr, const int kstart = -1)
{
  const double imu = 1.0 / muz;
  const int dk = to_obs ? -1 : 1;
  const int ks = to_obs ? ndep-1 : 0, ke = to_obs ? 0 : ndep-1;
  const rhdiv::Recip third(3.0);                 // x / 3.0 (IEEE quotient, not x * (1/3))

  double c_m = io.chi(ks), c_0 = io.chi(ks+dk);
  double z_m = z[ks], z_0 = z[ks+dk];
  double dtau_uw = 0.5 * imu * (c_m + c_0) * fabs(z_m - z_0);

  double I[4] = {0.0, 0.0, 0.0, 0.0};                 // bezier_1D.c:93-126
  if (to_obs) {
    if (bc_bottom == RHB200_BC_THERMALIZED) {
      const double B0 = planck(T[ndep-2], lambda), B1 = planck(T[ndep-1], lambda);
      I[0] = B1 - (B0 - B1) / dtau_uw;
    }
  }
  io.storeI(ks, I);
  io.storePsi(ks, 0.0);

  int k = ks + dk;
  double c_p = io.chi(k+dk), z_p = z[k+dk];
  double dsup = fabs(z_0 - z_m) * imu;
  double dsdn = fabs(z_p - z_0) * imu;
  double dchi_up = (c_0 - c_m) / dsup;
  double fchi = (c_p - c_0) / dsdn;                    // f_i of cent_deriv, carried forward
  double dchi_c = fb_deriv(dchi_up, fchi, fb_alpha(dsup, dsdn));

  {
    const double dsup3 = third.div(dsup);
    const double c2 = RH_MAX0(c_0 - dsup3 * dchi_c);
    const double c1 = RH_MAX0(c_m + dsup3 * dchi_up);
    dtau_uw = 0.25 * dsup * (c_0 + c_m + c1 + c2);
  }

  double Ku[3], K0[3], Kd[3], dKu[3], dK0[3], fK[3];
  double Su[4], S0[4], Sd[4], dSu[4], dS0[4], fS[4];
  io.K(ks, Ku);  io.K(k, K0);
  io.S(ks, Su);  io.S(k, S0);
  {
    const rhdiv::Recip ruw(dtau_uw);
#pragma unroll
    for (int n = 0; n < 4; n++) { dSu[n] = ruw.div(S0[n] - Su[n]); fS[n] = dSu[n]; }
#pragma unroll
    for (int n = 0; n < 3; n++) { dKu[n] = ruw.div(K0[n] - Ku[n]); fK[n] = dKu[n]; }
  }

#define DELO_STATE(OP) { double *q_ = state + (size_t) k * DELO_NSTATE; int n_ = 0;                                   \
    for (int m_ = 0; m_ < 4; m_++) { OP(I[m_]); OP(Su[m_]); OP(S0[m_]); OP(dSu[m_]); OP(fS[m_]); }                        \
    for (int m_ = 0; m_ < 3; m_++) { OP(Ku[m_]); OP(K0[m_]); OP(dKu[m_]); OP(fK[m_]); }                                    \
    OP(dtau_uw); OP(dsup); OP(dchi_up); OP(dchi_c); OP(fchi); OP(c_m); OP(c_0); OP(c_p); OP(z_m); OP(z_0); OP(z_p); }
#define DELO_PUT(x) q_[n_++] = (x)
#define DELO_GET(x) (x) = q_[n_++]
  if (MODE == 2 && kstart >= 0) {                  // resume: only for to_obs sweeps with k = kstart still inside the loop
    k = kstart;
    DELO_STATE(DELO_GET)
  }
  for (; k != ke; k += dk) {
    if (MODE == 1) DELO_STATE(DELO_PUT)
    io.prefetch(k + 4*dk, ndep);                   // pull the records of a later depth into L1
    dsdn = fabs(z_p - z_0) * imu;
    double dchi_dn, c_pp = 0.0, z_pp = 0.0, fnext = fchi;
    if (abs(k - ke) > 1) {
      c_pp = io.chi(k+2*dk); z_pp = z[k+2*dk];
      const double dsdn2 = fabs(z_pp - z_p) * imu;
      fnext = (c_pp - c_p) / dsdn2;
      dchi_dn = fb_deriv(fchi, fnext, fb_alpha(dsdn, dsdn2));
    } else
      dchi_dn = fchi;

    const double dsdn3 = third.div(dsdn);
    const double c2 = RH_MAX0(c_0 + dsdn3 * dchi_c);
    const double c1 = RH_MAX0(c_p - dsdn3 * dchi_dn);
    const double dtau_dw = 0.25 * dsdn * (c_0 + c_p + c1 + c2);
    const double dt = dtau_uw, dt03 = third.div(dt);

    double alpha, beta, gamma, theta, eps;
    bezier3_coeffs(dt, alpha, beta, gamma, theta, eps);
    io.storePsi(k, alpha + gamma);

    io.K(k+dk, Kd);
    io.S(k+dk, Sd);

    const double ca = fb_alpha(dtau_uw, dtau_dw);
    const rhdiv::Recip rdw(dtau_dw);             // seven slopes share the divisor dtau_dw
#pragma unroll
    for (int n = 0; n < 3; n++) {
      const double fi = rdw.div(Kd[n] - K0[n]);
      dK0[n] = fb_deriv(fK[n], fi, ca);
      fK[n] = fi;
    }
#pragma unroll
    for (int n = 0; n < 4; n++) {
      const double fi = rdw.div(Sd[n] - S0[n]);
      dS0[n] = fb_deriv(fS[n], fi, ca);
      fS[n] = fi;
    }

    // (Ku#Ku)[0][0], (K0#K0)[0][0]: 0 + q*q + u*u + v*v in k-order (m4m, bezier_aux.c:95-98)
    const double Mu00 = Ku[0]*Ku[0] + Ku[1]*Ku[1] + Ku[2]*Ku[2];
    const double A00  = K0[0]*K0[0] + K0[1]*K0[1] + K0[2]*K0[2];

    float Md[16];
    double Ma0[3], Mai[3][3];     // Ma[0][1..3] (= Ma[1..3][0]) and the lower-right block
    Md[0] = (float) (1.0 + gamma * (dt03 * A00));
    const double Ma00 = eps + theta * (dt03 * Mu00);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double md = alpha * K0[i] + gamma * (dt03 * (dK0[i] + K0[i]) + K0[i]);
      Md[1+i] = (float) md;  Md[4*(1+i)] = (float) md;
      Ma0[i] = theta * (dt03 * (dKu[i] + Ku[i]) - Ku[i]) - beta * Ku[i];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double dij = (i == j) ? 1.0 : 0.0;
        Md[4*(1+j) + 1+i] = (float) (dij + gamma * (dt03 * (K0[i]*K0[j])));
        Mai[j][i] = eps * dij + theta * (dt03 * (Ku[i]*Ku[j]));
      }
    }
    const double Mbd = beta + theta, Mcd = alpha + gamma;
    double V0[4], V1[4];
    {
      // row 0: sum over j in order, bezier_1D.c:232-240
      double v = (Ma00*I[0] + Mbd*Su[0]) + Mcd*S0[0];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double mb = theta * (0.0 - dt03 * Ku[j]);
        const double mc = gamma * (dt03 * K0[j]);
        v = v + ((Ma0[j]*I[1+j] + mb*Su[1+j]) + mc*S0[1+j]);
      }
      V0[0] = v + dt03 * (gamma * dS0[0] - theta * dSu[0]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double mb = theta * (0.0 - dt03 * Ku[i]);
      const double mc = gamma * (dt03 * K0[i]);
      double v = (Ma0[i]*I[0] + mb*Su[0]) + mc*S0[0];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        if (j == i) v = v + ((Mai[i][j]*I[1+j] + Mbd*Su[1+j]) + Mcd*S0[1+j]);
        else        v = v + Mai[i][j]*I[1+j];
      }
      V0[1+i] = v + dt03 * (gamma * dS0[1+i] - theta * dSu[1+i]);
    }

    float Mi[16];
    matinv_f32(Md, Mi);
    m4v(Mi, V0, V1);
#pragma unroll
    for (int n = 0; n < 4; n++) I[n] = V1[n];
    io.storeI(k, I);

#pragma unroll
    for (int n = 0; n < 4; n++) { Su[n] = S0[n]; S0[n] = Sd[n]; dSu[n] = dS0[n]; }
#pragma unroll
    for (int n = 0; n < 3; n++) { Ku[n] = K0[n]; K0[n] = Kd[n]; dKu[n] = dK0[n]; }
    dtau_uw = dtau_dw; dsup = dsdn; dchi_up = dchi_c; dchi_c = dchi_dn;
    fchi = fnext;
    c_m = c_0; c_0 = c_p; c_p = c_pp;
    z_m = z_0; z_0 = z_p; z_p = z_pp;
  }

  // linear DELO step in the last interval, bezier_1D.c:268-299
  dtau_uw = 0.5*imu * (c_0 + c_m) * fabs(z_0 - z_m);
  double w0, w1;
  w3(dtau_uw, w0, w1);
  io.storePsi(ke, w0 - w1 / dtau_uw);
  {
    const double a = -w1/dtau_uw, q = w0 - w1/dtau_uw, a_d = 1.0 - w0;
    float Md[16], Mi[16];
    double V0[4], V1[4];
#pragma unroll
    for (int n = 0; n < 4; n++) V0[n] = w0*S0[n] + w1 * -dSu[n];
    // A[n][m] = a*Ku[n][m], A[n][n] = 1-w0;  Md = q*K0, diag 1
    {
      double v = V0[0] + a_d * I[0];
#pragma unroll
      for (int m = 0; m < 3; m++) v = v + (a * Ku[m]) * I[1+m];
      V0[0] = v;
    }
#pragma unroll
    for (int n = 0; n < 3; n++) {
      double v = V0[1+n] + (a * Ku[n]) * I[0];
      V0[1+n] = v + a_d * I[1+n];
    }
#pragma unroll
    for (int j = 0; j < 16; j++) Md[j] = 0.0f;
    Md[0] = Md[5] = Md[10] = Md[15] = 1.0f;
#pragma unroll
    for (int n = 0; n < 3; n++) { const float f = (float) (q * K0[n]); Md[1+n] = f; Md[4*(1+n)] = f; }
    matinv_f32(Md, Mi);
    m4v(Mi, V0, V1);
    io.storeI(ke, V1);
  }
}

#undef DELO_STATE
#undef DELO_PUT
#undef DELO_GET

}  // namespace rhd
