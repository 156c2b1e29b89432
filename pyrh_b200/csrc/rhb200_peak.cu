// rhb200_peak.cu -- FP64-pipe micro-benchmark used as the measured roofline denominator
// for the FP64-bound kernels ("measure both peaks on the box").
// Eight independent register chains per thread, no memory traffic in the timed loop.
#include "rhb200_common.cuh"

namespace {

template <bool FMA>
__global__ void __launch_bounds__(256) fp64_chain_kernel(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0,
         x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; i++) {
    if (FMA) {
      x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    } else {
      x0 = __dmul_rn(x0, a); x1 = __dadd_rn(x1, b); x2 = __dmul_rn(x2, a); x3 = __dadd_rn(x3, b);
      x4 = __dmul_rn(x4, a); x5 = __dadd_rn(x5, b); x6 = __dmul_rn(x6, a); x7 = __dadd_rn(x7, b);
    }
  }
  out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

template <bool FMA>
int time_chain(rhb200_ctx *ctx, double *d_out, int blocks, int iters, double *ops_per_s)
{
  cudaEvent_t e0, e1;
  RH_CUDA(cudaEventCreate(&e0)); RH_CUDA(cudaEventCreate(&e1));
  fp64_chain_kernel<FMA><<<blocks, 256, 0, ctx->stream>>>(d_out, 64, 0.999999, 1e-9);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    RH_CUDA(cudaEventRecord(e0, ctx->stream));
    fp64_chain_kernel<FMA><<<blocks, 256, 0, ctx->stream>>>(d_out, iters, 0.999999, 1e-9);
    RH_CUDA(cudaEventRecord(e1, ctx->stream));
    RH_CUDA(cudaEventSynchronize(e1));
    float ms; RH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  RH_CUDA(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ops_per_s = (double) blocks * 256.0 * iters * 8.0 / (best * 1e-3);   // FP64 instructions / s
  return RHB200_OK;
}

}  // namespace

int rh_fp64_peak(rhb200_ctx *ctx, double *tf_fma, double *tf_nofma)
{
  const int blocks = ctx->sm_count * 8, iters = 20000;
  double *d_out = nullptr;
  RH_CUDA(cudaMalloc(&d_out, (size_t) blocks * 256 * sizeof(double)));
  double ops = 0.0;
  int rc = time_chain<true>(ctx, d_out, blocks, iters, &ops);
  if (rc == RHB200_OK && tf_fma) *tf_fma = 2.0 * ops / 1e12;
  if (rc == RHB200_OK) rc = time_chain<false>(ctx, d_out, blocks, iters, &ops);
  if (rc == RHB200_OK && tf_nofma) *tf_nofma = ops / 1e12;
  cudaFree(d_out);
  return rc;
}
