// Shared host/device helpers for the cream_b200 kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/cream_b200.h"

namespace cb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int check_last(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "cream_b200: %s: %s\n", what, cudaGetErrorString(e));
    return CREAM_ERR_CUDA;
  }
  return CREAM_OK;
}

#define CB_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      fprintf(stderr, "cream_b200: %s failed: %s\n", #expr, cudaGetErrorString(_e));  \
      return CREAM_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define CB_REQUIRE(cond, msg)                                             \
  do {                                                                    \
    if (!(cond)) {                                                        \
      fprintf(stderr, "cream_b200: invalid argument: %s (%s)\n", msg, #cond); \
      return CREAM_ERR_ARG;                                               \
    }                                                                     \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// exact (erf) GELU and its derivative, fp32 — the reference computes GELU in fp32
// (AutoFormer/model/supernet_transformer.py:14-18).
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Host-side TMA tensor-map factory (cached). dims/strides in elements; strides[0]
// is implicit (1). Returns nullptr on failure.
const CUtensorMap* get_tensor_map(const void* base, CUtensorMapDataType dtype, int rank,
                                  const uint64_t* dims, const uint64_t* strides_elems,
                                  const uint32_t* box, CUtensorMapSwizzle swizzle);

}  // namespace cb
