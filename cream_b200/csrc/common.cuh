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

// erf-GELU and its derivative in fp32 — the reference computes GELU in fp32 and casts the result
// back (AutoFormer/model/supernet_transformer.py:14-18); here the result is rounded to bf16 (2^-9
// relative), so the normal CDF is evaluated as an odd minimax polynomial on a clamped argument:
//   Phi(x) - 1/2      = z Q(z^2),  z = clamp(x, -4, 4),      |error| <= 7.5e-5   (deg 15)
//   dGELU(x) - 1/2    = z R(z^2),  z = clamp(x, -4.5, 4.5),  |error| <= 2.0e-4   (deg 19)
// (Chebyshev-basis fit, verified in fp32 Horner form over [-10, 10]; both reach exactly 1/2 at the
// clamp so the far tails are exact.)  No MUFU and no float->bf16->float round trips: the fc1 / fc2
// epilogues are instruction-issue bound, and erff() or the exp-based Abramowitz-Stegun form cost
// 1.5-2x the instructions plus two special-function ops per element.
__device__ __forceinline__ float gelu_f(float x) {
  const float z = fminf(fmaxf(x, -4.0f), 4.0f);
  const float u = z * z;
  float q = -1.580786280e-09f;
  q = fmaf(q, u, 1.217111105e-07f);
  q = fmaf(q, u, -4.100866136e-06f);
  q = fmaf(q, u, 8.066739247e-05f);
  q = fmaf(q, u, -1.048204373e-03f);
  q = fmaf(q, u, 9.664874524e-03f);
  q = fmaf(q, u, -6.617537886e-02f);
  q = fmaf(q, u, 3.988603354e-01f);
  return fmaf(x, z * q, 0.5f * x);   // x * (1/2 + (Phi - 1/2))
}
__device__ __forceinline__ float dgelu_f(float x) {
  const float z = fminf(fmaxf(x, -4.5f), 4.5f);
  const float u = z * z;
  float r = -2.210826834e-11f;
  r = fmaf(r, u, 2.521341358e-09f);
  r = fmaf(r, u, -1.268062277e-07f);
  r = fmaf(r, u, 3.721952680e-06f);
  r = fmaf(r, u, -7.122215902e-05f);
  r = fmaf(r, u, 9.405431920e-04f);
  r = fmaf(r, u, -8.815863170e-03f);
  r = fmaf(r, u, 5.860930681e-02f);
  r = fmaf(r, u, -2.649256885e-01f);
  r = fmaf(r, u, 7.976230979e-01f);
  return fmaf(z, r, 0.5f);           // Phi(x) + x phi(x)
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Adds n contiguous fp32 partial sums held in SHARED memory to global memory.  Called by every
// thread of the block after the sums are complete (it synchronises).  When the geometry allows
// it is ONE bulk reduce (the L2 adds whole 128-byte lines), not n same-address atomics: with a few
// hundred blocks folding into the same few hundred addresses, per-element atomics serialise in the
// L2 and cost more than the streaming pass that produced the sums.
__device__ __forceinline__ void block_add_to_global(float* dst, const float* s_src, int n) {
  __syncthreads();
  const bool bulk = (n & 3) == 0 && n > 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 &&
                    (static_cast<uint32_t>(__cvta_generic_to_shared(s_src)) & 15) == 0;
  if (bulk) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                   "r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_src))), "r"(n * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the read
    }
  } else {
    const int t = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int i = t; i < n; i += nt) atomicAdd(dst + i, s_src[i]);
  }
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The backward / forward of a block is a chain of ~25 dependent launches of 15-60 us each; with plain stream order
// every kernel pays its own ramp (CTA scheduling, barrier init, TMEM allocation, tensor-map fetch) after the previous
// kernel's LAST CTA has drained.  Kernels launched through launch_chain() carry
// cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs are placed on SMs as the previous kernel's CTAs exit
// and run their prologue there, then block in pdl_wait() (griddepcontrol.wait: returns when every prerequisite grid
// has COMPLETED and its writes are visible).  Contract for a kernel launched this way: pdl_trigger() first thing,
// and pdl_wait() before its first global-memory access of any kind (reads of the predecessor's output, and writes
// the predecessor may still be reading).  Completion is transitive - each grid waits for its predecessor before doing
// anything - so stream order of the data is unchanged.  CREAM_PDL=0 restores plain launches (the two device
// instructions are no-ops then).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // tensormap.cu: CREAM_PDL != "0"

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Host-side TMA tensor-map factory (cached). dims/strides in elements; strides[0]
// is implicit (1). Returns nullptr on failure.
const CUtensorMap* get_tensor_map(const void* base, CUtensorMapDataType dtype, int rank,
                                  const uint64_t* dims, const uint64_t* strides_elems,
                                  const uint32_t* box, CUtensorMapSwizzle swizzle);

}  // namespace cb
