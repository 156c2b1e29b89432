// rhb200_lu.cuh -- SolveLinearEq / LUdecomp / LUbacksubst (rh/ludcmp.c:36-177) on thread-local
// storage: Crout LU with implicit row scaling, the reference's pivot tie rule (>=), the 1e-20
// singularity guard and one step of iterative improvement.  Row-major A[N*N], destroyed.
#pragma once

namespace rhlu {

// matrix accessors: plain row-major thread-local storage, or one element per thread interleaved in shared
// memory (element e of thread t at p[e*stride], p already offset by t: conflict-free whatever e each thread asks for)
struct RowMajor { double *p; int N; __device__ __forceinline__ double &operator()(int i, int j) const { return p[i*N + j]; } };
struct Interleaved {
  double *p; int N, stride;
  __device__ __forceinline__ double &operator()(int i, int j) const { return p[(size_t) (i*N + j) * stride]; }
};

template <int MAXN, class Mat>
__device__ void solve_linear_eq_mat(const int N, const Mat M, double *b, const bool improve)
{
#define A_(i, j) M(i, j)
  int index[MAXN];
  double vv[MAXN], A_copy[MAXN*MAXN], b_copy[MAXN], residual[MAXN];
  if (improve) {
    for (int i = 0; i < N; i++) { b_copy[i] = b[i]; for (int j = 0; j < N; j++) A_copy[i*N+j] = A_(i, j); }
  }
  // LUdecomp, ludcmp.c:92-150
  int imax = 0;
  for (int i = 0; i < N; i++) {
    double big = 0.0;
    for (int j = 0; j < N; j++) { const double temp = fabs(A_(i, j)); if (temp > big) big = temp; }
    vv[i] = 1.0 / big;
  }
  for (int j = 0; j < N; j++) {
    for (int i = 0; i < j; i++) {
      double sum = A_(i, j);
      for (int k = 0; k < i; k++) sum -= A_(i, k) * A_(k, j);
      A_(i, j) = sum;
    }
    double big = 0.0;
    for (int i = j; i < N; i++) {
      double sum = A_(i, j);
      for (int k = 0; k < j; k++) sum -= A_(i, k) * A_(k, j);
      A_(i, j) = sum;
      const double dum = vv[i]*fabs(sum);
      if (dum >= big) { big = dum; imax = i; }
    }
    if (j != imax) {
      for (int k = 0; k < N; k++) { const double dum = A_(imax, k); A_(imax, k) = A_(j, k); A_(j, k) = dum; }
      vv[imax] = vv[j];
    }
    index[j] = imax;
    if (A_(j, j) == 0.0) A_(j, j) = 1.0e-20;
    const double dum = 1.0 / A_(j, j);
    for (int i = j+1; i < N; i++) A_(i, j) *= dum;
  }
  // LUbacksubst, ludcmp.c:156-177
  auto backsubst = [&](double *x) {
    int ii = -1;
    for (int i = 0; i < N; i++) {
      const int ip = index[i];
      double sum = x[ip];
      x[ip] = x[i];
      if (ii >= 0) { for (int j = ii; j < i; j++) sum -= A_(i, j) * x[j]; }
      else if (sum != 0.0) ii = i;
      x[i] = sum;
    }
    for (int i = N-1; i >= 0; i--) {
      double sum = x[i];
      for (int j = i+1; j < N; j++) sum -= A_(i, j)*x[j];
      x[i] = sum / A_(i, i);
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

#undef A_

template <int MAXN>
__device__ __forceinline__ void solve_linear_eq(const int N, double *A, double *b, const bool improve)
{
  solve_linear_eq_mat<MAXN>(N, RowMajor{A, N}, b, improve);
}

}  // namespace rhlu
