// rhb200_lu.cuh -- SolveLinearEq / LUdecomp / LUbacksubst (rh/ludcmp.c:36-177) on thread-local
// storage: Crout LU with implicit row scaling, the reference's pivot tie rule (>=), the 1e-20
// singularity guard and one step of iterative improvement.  Row-major A[N*N], destroyed.
#pragma once

namespace rhlu {

template <int MAXN>
__device__ void solve_linear_eq(const int N, double *A, double *b, const bool improve)
{
  int index[MAXN];
  double vv[MAXN], A_copy[MAXN*MAXN], b_copy[MAXN], residual[MAXN];
  if (improve) {
    for (int i = 0; i < N; i++) { b_copy[i] = b[i]; for (int j = 0; j < N; j++) A_copy[i*N+j] = A[i*N+j]; }
  }
  // LUdecomp, ludcmp.c:92-150
  int imax = 0;
  for (int i = 0; i < N; i++) {
    double big = 0.0;
    for (int j = 0; j < N; j++) { const double temp = fabs(A[i*N+j]); if (temp > big) big = temp; }
    vv[i] = 1.0 / big;
  }
  for (int j = 0; j < N; j++) {
    for (int i = 0; i < j; i++) {
      double sum = A[i*N+j];
      for (int k = 0; k < i; k++) sum -= A[i*N+k] * A[k*N+j];
      A[i*N+j] = sum;
    }
    double big = 0.0;
    for (int i = j; i < N; i++) {
      double sum = A[i*N+j];
      for (int k = 0; k < j; k++) sum -= A[i*N+k] * A[k*N+j];
      A[i*N+j] = sum;
      const double dum = vv[i]*fabs(sum);
      if (dum >= big) { big = dum; imax = i; }
    }
    if (j != imax) {
      for (int k = 0; k < N; k++) { const double dum = A[imax*N+k]; A[imax*N+k] = A[j*N+k]; A[j*N+k] = dum; }
      vv[imax] = vv[j];
    }
    index[j] = imax;
    if (A[j*N+j] == 0.0) A[j*N+j] = 1.0e-20;
    const double dum = 1.0 / A[j*N+j];
    for (int i = j+1; i < N; i++) A[i*N+j] *= dum;
  }
  // LUbacksubst, ludcmp.c:156-177
  auto backsubst = [&](double *x) {
    int ii = -1;
    for (int i = 0; i < N; i++) {
      const int ip = index[i];
      double sum = x[ip];
      x[ip] = x[i];
      if (ii >= 0) { for (int j = ii; j < i; j++) sum -= A[i*N+j] * x[j]; }
      else if (sum != 0.0) ii = i;
      x[i] = sum;
    }
    for (int i = N-1; i >= 0; i--) {
      double sum = x[i];
      for (int j = i+1; j < N; j++) sum -= A[i*N+j]*x[j];
      x[i] = sum / A[i*N+i];
    }
  };
  backsubst(b);
  if (improve) {
    for (int i = 0; i < N; i++) {
      residual[i] = b_copy[i];
      for (int j = 0; j < N; j++) residual[i] -= A_copy[i*N+j] * b[j];
    }
    backsubst(residual);
    for (int i = 0; i < N; i++) b[i] += residual[i];
  }
}

}  // namespace rhlu
