// fp_strict.cuh -- explicit round-to-nearest fp64 operations: nvcc never contracts them into FMAs, so an expression
// written with them is evaluated in the reference's left-to-right order with the reference's roundings.
#pragma once
#include <cuda_runtime.h>

namespace pic1dp {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

}  // namespace pic1dp
