// rhb200_math.cuh -- transcendental functions of the device path.
// The reference calls glibc libm (exp, pow, sin, cos); these wrappers are the
// single place where the device equivalents are chosen (DESIGN.md "libm").
#pragma once
#include <cuda_runtime.h>

namespace rhm {

__device__ __forceinline__ double rh_exp(double x) { return exp(x); }
__device__ __forceinline__ double rh_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double rh_sin(double x) { return sin(x); }
__device__ __forceinline__ double rh_cos(double x) { return cos(x); }

}  // namespace rhm
