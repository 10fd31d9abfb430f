// Shared helpers of the fused attention kernels: explicit shared-memory accessors (32-bit
// shared addresses, so the compiler can never fall back to generic loads), exp2, packed
// TMEM stores and index-byte extraction.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cb {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts_u32x4s(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f16(uint32_t a) {
  unsigned short h;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(a));
  return __half2float(__ushort_as_half(h));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}

// 32 fp32 accumulator values of one output row -> 32 bf16 (64 contiguous bytes): two 256-bit stores when the
// destination is 32-byte aligned (`wide`, checked on the host: base pointer and leading dimension), else four 128-bit ones.
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const uint32_t (&raw)[32], float mul, bool wide) {
  uint32_t u[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) u[k] = pack_bf16x2(mul * __uint_as_float(raw[2 * k]), mul * __uint_as_float(raw[2 * k + 1]));
  if (wide) {
    stg_256(dst, u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7]);
    stg_256(dst + 16, u[8], u[9], u[10], u[11], u[12], u[13], u[14], u[15]);
  } else {
    uint4* o4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; ++q) o4[q] = make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
  }
}
__host__ inline bool aligned_for_256bit(const void* base, int64_t ld_elems_bf16) {
  return (reinterpret_cast<uintptr_t>(base) & 31) == 0 && (ld_elems_bf16 % 16) == 0;
}

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[4], int k) {
  return (w[k >> 2] >> (8 * (k & 3))) & 0xFF;
}

// ---- AutoFormer relative position through the tensor cores (see attention_fwd.cu, softmax_plain) ----
constexpr int kFeat = 32;            // 14 rows + 14 columns + cls, padded to two K = 16 steps
constexpr int kIndChunk = kFeat * 128;   // bytes of one 64-key chunk of Ind (32 feature rows x 128 B)

// Ind[t][j] (bf16 0/1) for a 14 x 14 grid + cls; 128B-swizzled [feature rows][64-key chunks], written
// cooperatively by `nthreads` threads (tid in [0, nthreads)).
__device__ __forceinline__ void write_ind_matrix(uint32_t base, int tid, int nthreads, int G, int N) {
  for (int U = tid; U < 4 * kFeat * 8; U += nthreads) {
    const int c = U >> 8, t = (U >> 3) & 31, u = U & 7;
    uint32_t w[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      uint32_t word = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = c * 64 + u * 8 + e2 * 2 + e;
        bool one;
        if (j == 0) one = t == 2 * G;
        else if (j < N) one = (t == (j - 1) / G) || (t == G + (j - 1) % G);
        else one = false;
        if (one) word |= 0x3F80u << (16 * e);     // bf16 1.0
      }
      w[e2] = word;
    }
    const uint32_t a = base + c * kIndChunk + t * 128 + ((u ^ (t & 7)) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
}

// Dynamic shared memory base, required to be 1024-byte aligned (128B-swizzle atoms).
__device__ __forceinline__ void require_smem_alignment(const void* smem) {
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("cream_b200: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
}

}  // namespace cb
