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

// erf-GELU and its derivative in fp32 — the reference computes GELU in fp32
// (AutoFormer/model/supernet_transformer.py:14-18).  erf is evaluated with Abramowitz-Stegun
// 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32 round-off level) on the MUFU rcp/ex2 units: about half
// the instructions of erff(), which matters because the fc1 epilogue is instruction-bound.
//   q(x) = P(t) * exp(-x^2/2), t = 1/(1 + p|x|/sqrt2)   =>   Phi(x) = x >= 0 ? 1 - q/2 : q/2
__device__ __forceinline__ void gelu_terms(float x, float& cdf, float& e) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170368f * x * x));  // exp(-x^2/2)
  const float hq = 0.5f * poly * e;
  cdf = x >= 0.f ? 1.0f - hq : hq;
}
__device__ __forceinline__ float gelu_f(float x) {
  float cdf, e;
  gelu_terms(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_f(float x) {
  float cdf, e;
  gelu_terms(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
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
