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

// dX = resid_grad + rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)); dgamma/dbeta via
// per-lane register partials -> shared -> global atomics.
constexpr int kBwdWarps = 8;

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
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    atomicAdd(dgamma + c, sacc[c]);
    atomicAdd(dbeta + c, sacc[E + c]);
  }
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
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(rows, kBwdWarps), kNumSMs * 2));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = 2 * E * sizeof(float);
#define CB_LN_BWD(F32, PER)                                                                             \
  ln_bwd_kernel<F32, PER><<<grid, kBwdWarps * 32, smem, stream>>>(dy, lddy, x, ldx, gamma, mean, rstd, resid_grad, \
                                                            ldrg, dx, lddx, dgamma, dbeta, rows, E)
  if (dy_f32) { if (E <= 256) CB_LN_BWD(true, 8); else if (E <= 512) CB_LN_BWD(true, 16); else CB_LN_BWD(true, 24); }
  else { if (E <= 256) CB_LN_BWD(false, 8); else if (E <= 512) CB_LN_BWD(false, 16); else CB_LN_BWD(false, 24); }
#undef CB_LN_BWD
  return check_last("ln_bwd_kernel");
}
