// Sliced LayerNorm forward / backward (HBM-bound row kernels, one warp per token row).
//
// Replaces F.layer_norm on the sampled prefix of the supernet LayerNorm parameters
// (AutoFormer/model/module/layernorm_super.py:26-37; computed in fp32 under autocast).
// Input is the fp32 residual stream; the normalised output is written as bf16, i.e. directly
// as the A operand of the following sliced GEMM (the autocast cast of the reference is fused).
#include "common.cuh"

namespace cb {
namespace {

constexpr int kWarps = 4;
constexpr int kMaxPerLane = 24;  // supports E <= 768 (kernels are instantiated for 8 / 16 / 24)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool kOutF32, int kPer>
__global__ void __launch_bounds__(kWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, void* __restrict__ out, int64_t ldo,
              float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int E) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (E + 31) >> 5;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * kWarps) {
    const float* xr = x + r * ldx;
    float v[kPer];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = lane + i * 32;
      v[i] = (i < per && c < E) ? xr[c] : 0.f;
      s += v[i];
    }
    const float mu = warp_sum(s) / E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = lane + i * 32;
      const float d = (i < per && c < E) ? v[i] - mu : 0.f;
      q += d * d;
    }
    const float rs = rsqrtf(warp_sum(q) / E + eps);
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = lane + i * 32;
      if (i < per && c < E) {
        const float y = (v[i] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
        if (kOutF32) static_cast<float*>(out)[r * ldo + c] = y;
        else static_cast<__nv_bfloat16*>(out)[r * ldo + c] = __float2bfloat16_rn(y);
      }
    }
  }
}

// Vector twin of ln_fwd_kernel for E % 4 == 0 and 16-byte aligned rows: one 16-byte load and one
// 8-byte (bf16) / 16-byte (fp32) store per lane and 128 columns, all loads of a row in flight before
// the first reduction.  (The scalar kernel moves 4 / 2 bytes per lane per instruction.)
template <bool kOutF32, int kV>
__global__ void __launch_bounds__(kWarps * 32)
ln_fwd_vec_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, void* __restrict__ out, int64_t ldo,
                  float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int E) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invE = 1.0f / E;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * kWarps) {
    const float* xr = x + r * ldx;
    float4 v[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      v[i] = c < E ? __ldg(reinterpret_cast<const float4*>(xr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mu = warp_sum(s) * invE;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < E) {
        v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invE + eps);
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < E) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float y0 = fmaf(v[i].x * rs, g.x, b.x), y1 = fmaf(v[i].y * rs, g.y, b.y);
        const float y2 = fmaf(v[i].z * rs, g.z, b.z), y3 = fmaf(v[i].w * rs, g.w, b.w);
        if (kOutF32) {
          *reinterpret_cast<float4*>(static_cast<float*>(out) + r * ldo + c) = make_float4(y0, y1, y2, y3);
        } else {
          uint2 o;
          o.x = pack_bf16x2(y0, y1);
          o.y = pack_bf16x2(y2, y3);
          *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + r * ldo + c) = o;
        }
      }
    }
  }
}

// dX = resid_grad + rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)); dgamma/dbeta via
// per-lane register partials -> shared -> global atomics.
constexpr int kBwdWarps = 4;

template <bool kDyF32, int kPer>
__global__ void __launch_bounds__(kBwdWarps * 32)
ln_bwd_kernel(const void* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
              const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ resid_grad, int64_t ldrg,
              float* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
              float* __restrict__ dbeta, int64_t rows, int E) {
  extern __shared__ float sacc[];  // [2][E]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (E + 31) >> 5;
  for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float pg[kPer], pb[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kBwdWarps + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * kBwdWarps) {
    const float mu = mean[r], rs = rstd[r];
    const float* xr = x + r * ldx;
    float xh[kPer], dg[kPer];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = lane + i * 32;
      if (i < per && c < E) {
        const float d = kDyF32 ? static_cast<const float*>(dy)[r * lddy + c]
                               : __bfloat162float(static_cast<const __nv_bfloat16*>(dy)[r * lddy + c]);
        xh[i] = (xr[c] - mu) * rs;
        dg[i] = d * __ldg(gamma + c);
        pg[i] += d * xh[i];
        pb[i] += d;
        s1 += dg[i];
        s2 += dg[i] * xh[i];
      } else {
        xh[i] = 0.f;
        dg[i] = 0.f;
      }
    }
    s1 = warp_sum(s1) / E;
    s2 = warp_sum(s2) / E;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = lane + i * 32;
      if (i < per && c < E) {
        float g = rs * (dg[i] - s1 - xh[i] * s2);
        if (resid_grad) g += resid_grad[r * ldrg + c];
        dx[r * lddx + c] = g;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int c = lane + i * 32;
    if (i < per && c < E) {
      atomicAdd(&sacc[c], pg[i]);
      atomicAdd(&sacc[E + c], pb[i]);
    }
  }
  block_add_to_global(dgamma, sacc, E);
  block_add_to_global(dbeta, sacc + E, E);
}

// Vectorised backward (E % 4 == 0): each lane owns float4 groups c = 4*(lane + 32*i).  All loads
// of a row are issued back to back BEFORE any arithmetic (explicit load phase): with only 16
// resident warps per SM the kernel lives on memory-level parallelism inside a warp, and a
// load/accumulate interleaving serialises one DRAM round trip per element (measured: 12 % of HBM).
template <bool kDyF32, int kV>
__global__ void __launch_bounds__(kBwdWarps * 32, kV <= 3 ? 4 : 3)   // E <= 384: 16 warps / SM without spills
ln_bwd_vec_kernel(const void* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                  const float* __restrict__ gamma, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ resid_grad, int64_t ldrg,
                  float* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, int64_t rows, int E) {
  extern __shared__ float sacc[];  // [2][E] partial sums, then gamma [E] (kept out of the register file)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sgam = sacc + 2 * E;
  for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) sacc[i] = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) sgam[i] = __ldg(gamma + i);
  __syncthreads();
  float4 pg[kV], pb[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    pg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invE = 1.0f / E;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kBwdWarps + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * kBwdWarps) {
    float4 xv[kV], dv[kV], rg[kV];
    // ---- load phase ----
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < E) {
        xv[i] = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
        if (kDyF32) {
          dv[i] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(dy) + r * lddy + c));
        } else {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dy) + r * lddy + c));
          const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
          dv[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        rg[i] = resid_grad ? __ldg(reinterpret_cast<const float4*>(resid_grad + r * ldrg + c))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        xv[i] = dv[i] = rg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    // ---- arithmetic ----
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < E) {
        float4& xh = xv[i];
        xh.x = (xh.x - mu) * rs; xh.y = (xh.y - mu) * rs; xh.z = (xh.z - mu) * rs; xh.w = (xh.w - mu) * rs;
        pg[i].x += dv[i].x * xh.x; pg[i].y += dv[i].y * xh.y; pg[i].z += dv[i].z * xh.z; pg[i].w += dv[i].w * xh.w;
        pb[i].x += dv[i].x; pb[i].y += dv[i].y; pb[i].z += dv[i].z; pb[i].w += dv[i].w;
        const float4 gm = *reinterpret_cast<const float4*>(sgam + c);
        dv[i].x *= gm.x; dv[i].y *= gm.y; dv[i].z *= gm.z; dv[i].w *= gm.w;
        s1 += dv[i].x + dv[i].y + dv[i].z + dv[i].w;
        s2 += dv[i].x * xh.x + dv[i].y * xh.y + dv[i].z * xh.z + dv[i].w * xh.w;
      }
    }
    s1 = warp_sum(s1) * invE;
    s2 = warp_sum(s2) * invE;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < E) {
        float4 g;
        g.x = rs * (dv[i].x - s1 - xv[i].x * s2) + rg[i].x;
        g.y = rs * (dv[i].y - s1 - xv[i].y * s2) + rg[i].y;
        g.z = rs * (dv[i].z - s1 - xv[i].z * s2) + rg[i].z;
        g.w = rs * (dv[i].w - s1 - xv[i].w * s2) + rg[i].w;
        *reinterpret_cast<float4*>(dx + r * lddx + c) = g;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c < E) {
      atomicAdd(&sacc[c + 0], pg[i].x); atomicAdd(&sacc[c + 1], pg[i].y);
      atomicAdd(&sacc[c + 2], pg[i].z); atomicAdd(&sacc[c + 3], pg[i].w);
      atomicAdd(&sacc[E + c + 0], pb[i].x); atomicAdd(&sacc[E + c + 1], pb[i].y);
      atomicAdd(&sacc[E + c + 2], pb[i].z); atomicAdd(&sacc[E + c + 3], pb[i].w);
    }
  }
  block_add_to_global(dgamma, sacc, E);
  block_add_to_global(dbeta, sacc + E, E);
}

}  // namespace
}  // namespace cb

extern "C" int cream_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta,
                                   float eps, void* out, int64_t ldo, int out_f32, float* mean,
                                   float* rstd, int64_t rows, int E, void* stream_) {
  using namespace cb;
  if (rows == 0) return CREAM_OK;
  CB_REQUIRE(x && gamma && beta && out && rows > 0, "null pointer");
  CB_REQUIRE(E >= 1 && E <= 32 * kMaxPerLane, "embed dim must be <= 768");
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(rows, kWarps), kNumSMs * 16));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = (E & 3) == 0 && (ldx & 3) == 0 && (ldo & 3) == 0 && E <= 768 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                     reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (vec) {
#define CB_LN_FWD_VEC(F32, V) \
  ln_fwd_vec_kernel<F32, V><<<grid, kWarps * 32, 0, stream>>>(x, ldx, gamma, beta, eps, out, ldo, mean, rstd, rows, E)
    const int v = ceil_div(E, 128);
    if (out_f32) {
      if (v <= 2) CB_LN_FWD_VEC(true, 2); else if (v == 3) CB_LN_FWD_VEC(true, 3);
      else if (v == 4) CB_LN_FWD_VEC(true, 4); else CB_LN_FWD_VEC(true, 6);
    } else {
      if (v <= 2) CB_LN_FWD_VEC(false, 2); else if (v == 3) CB_LN_FWD_VEC(false, 3);
      else if (v == 4) CB_LN_FWD_VEC(false, 4); else CB_LN_FWD_VEC(false, 6);
    }
#undef CB_LN_FWD_VEC
    return check_last("ln_fwd_vec_kernel");
  }
#define CB_LN_FWD(F32, PER) \
  ln_fwd_kernel<F32, PER><<<grid, kWarps * 32, 0, stream>>>(x, ldx, gamma, beta, eps, out, ldo, mean, rstd, rows, E)
  if (out_f32) { if (E <= 256) CB_LN_FWD(true, 8); else if (E <= 512) CB_LN_FWD(true, 16); else CB_LN_FWD(true, 24); }
  else { if (E <= 256) CB_LN_FWD(false, 8); else if (E <= 512) CB_LN_FWD(false, 16); else CB_LN_FWD(false, 24); }
#undef CB_LN_FWD
  return check_last("ln_fwd_kernel");
}

extern "C" int cream_layernorm_bwd(const void* dy, int64_t lddy, int dy_f32, const float* x, int64_t ldx,
                                   const float* gamma, const float* mean, const float* rstd,
                                   const float* resid_grad, int64_t ldrg, float* dx, int64_t lddx,
                                   float* dgamma, float* dbeta, int64_t rows, int E, void* stream_) {
  using namespace cb;
  if (rows == 0) return CREAM_OK;
  CB_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "null pointer");
  CB_REQUIRE(E >= 1 && E <= 32 * kMaxPerLane, "embed dim must be <= 768");
  // few, fat blocks: every block ends with 2*E global atomics, so the block count bounds the
  // per-address contention on dgamma / dbeta (1184 blocks made this kernel 4x slower than HBM)
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(rows, kBwdWarps), kNumSMs * 4));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = 3 * E * sizeof(float);
  const bool vec_ok = (E % 4 == 0) && (ldx % 4 == 0) && (lddx % 4 == 0) && (lddy % 4 == 0) &&
                      (resid_grad == nullptr || ldrg % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dy) |
                        reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(resid_grad)) & 15) == 0;
  if (vec_ok) {
#define CB_LN_BWDV(F32, V)                                                                               \
  ln_bwd_vec_kernel<F32, V><<<grid, kBwdWarps * 32, smem, stream>>>(dy, lddy, x, ldx, gamma, mean, rstd,    \
                                                                resid_grad, ldrg, dx, lddx, dgamma, dbeta, rows, E)
    if (dy_f32) { if (E <= 256) CB_LN_BWDV(true, 2); else if (E <= 512) CB_LN_BWDV(true, 4); else CB_LN_BWDV(true, 6); }
    else { if (E <= 256) CB_LN_BWDV(false, 2); else if (E <= 512) CB_LN_BWDV(false, 4); else CB_LN_BWDV(false, 6); }
#undef CB_LN_BWDV
    return check_last("ln_bwd_vec_kernel");
  }
#define CB_LN_BWD(F32, PER)                                                                             \
  ln_bwd_kernel<F32, PER><<<grid, kBwdWarps * 32, smem, stream>>>(dy, lddy, x, ldx, gamma, mean, rstd, resid_grad, \
                                                            ldrg, dx, lddx, dgamma, dbeta, rows, E)
  if (dy_f32) { if (E <= 256) CB_LN_BWD(true, 8); else if (E <= 512) CB_LN_BWD(true, 16); else CB_LN_BWD(true, 24); }
  else { if (E <= 256) CB_LN_BWD(false, 8); else if (E <= 512) CB_LN_BWD(false, 16); else CB_LN_BWD(false, 24); }
#undef CB_LN_BWD
  return check_last("ln_bwd_kernel");
}
