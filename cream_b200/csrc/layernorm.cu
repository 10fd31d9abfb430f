// Sliced LayerNorm forward / backward (HBM-bound row kernels, one warp per token row).
//
// Replaces F.layer_norm on the sampled prefix of the supernet LayerNorm parameters
// (AutoFormer/model/module/layernorm_super.py:26-37; computed in fp32 under autocast).
// Input is the fp32 residual stream; the normalised output is written as bf16, i.e. directly
// as the A operand of the following sliced GEMM (the autocast cast of the reference is fused).
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

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
  pdl_trigger();
  pdl_wait();
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


// ---------------------------------------------------------------------------------------------------
// Pipelined backward: the rows of dy, x and the residual gradient travel global -> shared as bulk async
// copies (cp.async.bulk + mbarrier complete_tx) issued by a producer warp kStages tiles ahead of the eight
// consumer warps, so the bytes in flight per SM are set by the ring (>= 100 KB), not by how many warps the
// register file admits - the register-resident kernel above tops out at 16 warps / SM, each alternating
// between a load phase and an arithmetic phase, and reaches half of the HBM rate.  One row per consumer
// warp and tile; per-lane dgamma / dbeta partials as above.
// ---------------------------------------------------------------------------------------------------
constexpr int kPipeRows = 8;                 // rows per tile = consumer warps
constexpr int kPipeThreads = (kPipeRows + 1) * 32;

template <bool kDyF32, int kV, int kMinBlocks>
__global__ void __launch_bounds__(kPipeThreads, kMinBlocks)
ln_bwd_pipe_kernel(const void* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                   const float* __restrict__ gamma, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ resid_grad, int64_t ldrg,
                   float* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, int64_t rows, int E, int stages,
                   __nv_bfloat16* __restrict__ out_bf, int64_t ldob, const float* __restrict__ row_scale, int rows_per,
                   float* __restrict__ dbias) {
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  // layout: [stages][ x: R*E f32 | rg: R*E f32 | dy: R*E (f32 | bf16) ]  sacc[3E]  sgam[E]  full[stages] empty[stages]
  const int dy_bytes_row = E * (kDyF32 ? 4 : 2);
  const int stat_off = kPipeRows * (2 * E * 4 + dy_bytes_row);        // mean[8] | rstd[8] of the tile's rows
  const int stage_bytes = stat_off + 2 * kPipeRows * 4;
  float* sacc = reinterpret_cast<float*>(pipe_smem + static_cast<size_t>(stages) * stage_bytes);
  float* sgam = sacc + 3 * E;
  uint64_t* full = reinterpret_cast<uint64_t*>(sgam + E);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_rg = resid_grad != nullptr;
  pdl_trigger();
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) sacc[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kPipeRows); }
    fence_mbar_init();
  }
  pdl_wait();      // shared-memory setup above overlaps the previous kernel's tail; global memory from here on
  for (int i = threadIdx.x; i < E; i += blockDim.x) sgam[i] = __ldg(gamma + i);
  __syncthreads();
  const int64_t tiles = ceil_div64(rows, kPipeRows);

  if (warp == kPipeRows) {
    // ===================== producer warp =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      mbar_wait(&empty[stage], phase ^ 1);          // fresh barrier: parity 1 passes at once
      const int64_t r0 = t * kPipeRows;
      const int valid = static_cast<int>(rows - r0 < kPipeRows ? rows - r0 : kPipeRows);
      unsigned char* st = pipe_smem + static_cast<size_t>(stage) * stage_bytes;
      const bool stats = valid == kPipeRows;      // full tiles: the statistics ride with the data (16-byte aligned, 32 bytes each)
      if (lane == 0) mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(valid) * ((has_rg ? 2 : 1) * E * 4 + dy_bytes_row) +
                                                             (stats ? 2 * kPipeRows * 4 : 0));
      __syncwarp();
      // lane -> (array, row): 3 arrays x 8 rows.  An array whose rows are contiguous (pitch == E, the arena's case)
      // moves as ONE copy per tile - the copy engine's cost is per request, and 24 requests of 1-2 KB per tile
      // capped the kernel at 1.6 us per tile - otherwise one copy per row.
      const int arr = lane >> 3, rr = lane & 7;
      const unsigned char* src = nullptr;
      unsigned char* dst = nullptr;
      uint32_t row_bytes = 0;
      int64_t pitch_bytes = 0;
      if (arr == 0) { src = reinterpret_cast<const unsigned char*>(x + r0 * ldx); dst = st; row_bytes = E * 4; pitch_bytes = ldx * 4; }
      else if (arr == 1 && has_rg) { src = reinterpret_cast<const unsigned char*>(resid_grad + r0 * ldrg); dst = st + kPipeRows * E * 4; row_bytes = E * 4; pitch_bytes = ldrg * 4; }
      else if (arr == 2) { src = static_cast<const unsigned char*>(dy) + r0 * lddy * (kDyF32 ? 4 : 2); dst = st + 2 * kPipeRows * E * 4; row_bytes = dy_bytes_row; pitch_bytes = lddy * (kDyF32 ? 4 : 2); }
      if (arr == 3 && stats && rr < 2)
        bulk_load_1d(st + stat_off + rr * kPipeRows * 4, (rr == 0 ? mean : rstd) + r0, kPipeRows * 4, &full[stage]);
      if (src != nullptr) {
        if (pitch_bytes == row_bytes) {
          if (rr == 0) bulk_load_1d(dst, src, row_bytes * valid, &full[stage]);
        } else if (rr < valid) {
          bulk_load_1d(dst + rr * row_bytes, src + rr * pitch_bytes, row_bytes, &full[stage]);
        }
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== consumer warps: row `warp` of every tile =====================
    float4 pg[kV], pb[kV], pc[kV];
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      pg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float invE = 1.0f / E;
    int stage = 0;
    uint32_t phase = 0;
    int64_t t = blockIdx.x;
    float sc = 1.f;
    if (t < tiles && t * kPipeRows + warp < rows && row_scale) sc = __ldg(row_scale + (t * kPipeRows + warp) / rows_per);
    for (; t < tiles; t += gridDim.x) {
      const int64_t r = t * kPipeRows + warp;
      // statistics of the NEXT tile's row are requested before this tile's data is waited for
      const int64_t rn = (t + gridDim.x) * kPipeRows + warp;
      float sc_n = 1.f;
      if (t + gridDim.x < tiles && rn < rows && row_scale) sc_n = __ldg(row_scale + rn / rows_per);
      mbar_wait(&full[stage], phase);
      if (r < rows) {
        const unsigned char* st = pipe_smem + static_cast<size_t>(stage) * stage_bytes;
        const bool full_tile = (t + 1) * kPipeRows <= rows;
        const float mu = full_tile ? reinterpret_cast<const float*>(st + stat_off)[warp] : __ldg(mean + r);
        const float rs = full_tile ? reinterpret_cast<const float*>(st + stat_off)[kPipeRows + warp] : __ldg(rstd + r);
        const float* xs = reinterpret_cast<const float*>(st) + warp * E;
        const float* gs = reinterpret_cast<const float*>(st) + (kPipeRows + warp) * E;
        const unsigned char* ds = st + 2 * kPipeRows * E * 4 + warp * dy_bytes_row;
        float4 xh[kV], dv[kV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          const int c = 4 * (lane + 32 * i);
          if (c < E) {
            xh[i] = *reinterpret_cast<const float4*>(xs + c);
            if (kDyF32) {
              dv[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ds) + c);
            } else {
              const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(ds) + c);
              const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
              dv[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            xh[i].x = (xh[i].x - mu) * rs; xh[i].y = (xh[i].y - mu) * rs; xh[i].z = (xh[i].z - mu) * rs; xh[i].w = (xh[i].w - mu) * rs;
            pg[i].x += dv[i].x * xh[i].x; pg[i].y += dv[i].y * xh[i].y; pg[i].z += dv[i].z * xh[i].z; pg[i].w += dv[i].w * xh[i].w;
            pb[i].x += dv[i].x; pb[i].y += dv[i].y; pb[i].z += dv[i].z; pb[i].w += dv[i].w;
            const float4 gm = *reinterpret_cast<const float4*>(sgam + c);
            dv[i].x *= gm.x; dv[i].y *= gm.y; dv[i].z *= gm.z; dv[i].w *= gm.w;
            s1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
            s2 += (dv[i].x * xh[i].x + dv[i].y * xh[i].y) + (dv[i].z * xh[i].z + dv[i].w * xh[i].w);
          } else {
            xh[i] = dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        s1 = warp_sum(s1) * invE;
        s2 = warp_sum(s2) * invE;
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          const int c = 4 * (lane + 32 * i);
          if (c < E) {
            float4 g = has_rg ? *reinterpret_cast<const float4*>(gs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            g.x += rs * (dv[i].x - s1 - xh[i].x * s2);
            g.y += rs * (dv[i].y - s1 - xh[i].y * s2);
            g.z += rs * (dv[i].z - s1 - xh[i].z * s2);
            g.w += rs * (dv[i].w - s1 - xh[i].w * s2);
            *reinterpret_cast<float4*>(dx + r * lddx + c) = g;
            if (out_bf != nullptr) {
              // what cast_scale_kernel would produce from dx: bf16(DropPath scale * g) and its column sums
              const __nv_bfloat162 lo = __floats2bfloat162_rn(sc * g.x, sc * g.y);
              const __nv_bfloat162 hi = __floats2bfloat162_rn(sc * g.z, sc * g.w);
              uint2 o;
              o.x = *reinterpret_cast<const uint32_t*>(&lo);
              o.y = *reinterpret_cast<const uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(out_bf + r * ldob + c) = o;
              pc[i].x += __bfloat162float(lo.x); pc[i].y += __bfloat162float(lo.y);
              pc[i].z += __bfloat162float(hi.x); pc[i].w += __bfloat162float(hi.y);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);     // this warp's reads of the stage are done
      sc = sc_n;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < E) {
        atomicAdd(&sacc[c + 0], pg[i].x); atomicAdd(&sacc[c + 1], pg[i].y);
        atomicAdd(&sacc[c + 2], pg[i].z); atomicAdd(&sacc[c + 3], pg[i].w);
        atomicAdd(&sacc[E + c + 0], pb[i].x); atomicAdd(&sacc[E + c + 1], pb[i].y);
        atomicAdd(&sacc[E + c + 2], pb[i].z); atomicAdd(&sacc[E + c + 3], pb[i].w);
        if (dbias != nullptr) {
          atomicAdd(&sacc[2 * E + c + 0], pc[i].x); atomicAdd(&sacc[2 * E + c + 1], pc[i].y);
          atomicAdd(&sacc[2 * E + c + 2], pc[i].z); atomicAdd(&sacc[2 * E + c + 3], pc[i].w);
        }
      }
    }
  }
  block_add_to_global(dgamma, sacc, E);
  block_add_to_global(dbeta, sacc + E, E);
  if (dbias != nullptr) block_add_to_global(dbias, sacc + 2 * E, E);
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
  CB_CUDA_OK(launch_chain(ln_fwd_vec_kernel<F32, V>, dim3(grid), dim3(kWarps * 32), 0, stream, 1, x, ldx, gamma, beta, eps, out, ldo, mean, rstd, rows, E))
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

static int layernorm_bwd_impl(const void* dy, int64_t lddy, int dy_f32, const float* x, int64_t ldx,
                              const float* gamma, const float* mean, const float* rstd,
                              const float* resid_grad, int64_t ldrg, float* dx, int64_t lddx,
                              float* dgamma, float* dbeta, int64_t rows, int E, void* stream_,
                              void* out_bf16, int64_t ldob, const float* row_scale, int rows_per_scale, float* dbias,
                              bool* emitted) {
  using namespace cb;
  *emitted = false;
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
  // pipelined path: bulk async copies need 16-byte aligned rows of every array
  const int dyb = dy_f32 ? 4 : 2;
  static const auto tune = [](const char* n, int dflt) { const char* e = getenv(n); return e != nullptr && e[0] != 0 ? atoi(e) : dflt; };
  const bool pipe_ok = tune("CREAM_LN_PIPE", 1) != 0 && vec_ok && rows >= 4 * kPipeRows && (E * dyb) % 16 == 0 && (lddy * dyb) % 16 == 0 &&
                       ((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(rstd)) & 15) == 0;
  if (pipe_ok) {
    const size_t stage_bytes = static_cast<size_t>(kPipeRows) * (2 * E * 4 + E * dyb) + 2 * kPipeRows * 4;
    const size_t fixed = 4 * E * sizeof(float) + 2 * 8 * sizeof(uint64_t) + 128;
    int stages = static_cast<int>(std::min<size_t>(8, (200 * 1024 - fixed) / stage_bytes));
    if (stages >= 3) {
      const size_t smem_pipe = stages * stage_bytes + 4 * E * sizeof(float) + 2 * stages * sizeof(uint64_t);
      const bool emit = out_bf16 != nullptr && ldob % 4 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0;
      __nv_bfloat16* ob = emit ? static_cast<__nv_bfloat16*>(out_bf16) : nullptr;
      const int rp = rows_per_scale > 0 ? rows_per_scale : 1;
      *emitted = emit;
      const int gridp = static_cast<int>(std::min<int64_t>(ceil_div64(rows, kPipeRows), kNumSMs));
#define CB_LN_BWDP_(F32, V, MB)                                                                                         \
  do {                                                                                                                  \
    static bool attr_done = false;                                                                                      \
    if (!attr_done) {                                                                                                   \
      CB_CUDA_OK(cudaFuncSetAttribute(ln_bwd_pipe_kernel<F32, V, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024)); \
      attr_done = true;                                                                                                 \
    }                                                                                                                   \
    CB_CUDA_OK(launch_chain(ln_bwd_pipe_kernel<F32, V, MB>, dim3(gridp), dim3(kPipeThreads), smem_pipe, stream, 1,       \
        dy, lddy, x, ldx, gamma, mean, rstd, resid_grad, ldrg, dx, lddx, dgamma, dbeta, rows, E,                        \
        stages, ob, ldob,                                                                                               \
        emit ? row_scale : nullptr, rp, emit ? dbias : nullptr));                                                       \
  } while (0)
#define CB_LN_BWDP(F32, V) CB_LN_BWDP_(F32, V, 1)
      if (dy_f32) { if (E <= 256) CB_LN_BWDP(true, 2); else if (E <= 512) CB_LN_BWDP(true, 4); else CB_LN_BWDP(true, 6); }
      else { if (E <= 256) CB_LN_BWDP(false, 2); else if (E <= 512) CB_LN_BWDP(false, 4); else CB_LN_BWDP(false, 6); }
#undef CB_LN_BWDP_
#undef CB_LN_BWDP
      return check_last("ln_bwd_pipe_kernel");
    }
  }
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

extern "C" int cream_layernorm_bwd(const void* dy, int64_t lddy, int dy_f32, const float* x, int64_t ldx,
                                   const float* gamma, const float* mean, const float* rstd,
                                   const float* resid_grad, int64_t ldrg, float* dx, int64_t lddx,
                                   float* dgamma, float* dbeta, int64_t rows, int E, void* stream_) {
  bool emitted;
  return layernorm_bwd_impl(dy, lddy, dy_f32, x, ldx, gamma, mean, rstd, resid_grad, ldrg, dx, lddx, dgamma, dbeta, rows, E,
                            stream_, nullptr, 0, nullptr, 1, nullptr, &emitted);
}

extern "C" int cream_layernorm_bwd_cast(const void* dy, int64_t lddy, int dy_f32, const float* x, int64_t ldx,
                                        const float* gamma, const float* mean, const float* rstd,
                                        const float* resid_grad, int64_t ldrg, float* dx, int64_t lddx,
                                        float* dgamma, float* dbeta, int64_t rows, int E, void* out_bf16,
                                        int64_t ldo, const float* row_scale, int rows_per_scale, float* dbias,
                                        void* stream_) {
  CB_REQUIRE(out_bf16 != nullptr, "null pointer");
  bool emitted;
  const int rc = layernorm_bwd_impl(dy, lddy, dy_f32, x, ldx, gamma, mean, rstd, resid_grad, ldrg, dx, lddx, dgamma, dbeta,
                                    rows, E, stream_, out_bf16, ldo, row_scale, rows_per_scale, dbias, &emitted);
  if (rc != CREAM_OK || emitted || rows == 0) return rc;
  // shapes the pipelined kernel does not take: the two-pass form
  return cream_cast_scale(dx, lddx, out_bf16, ldo, row_scale, rows_per_scale, dbias, rows, E, stream_);
}
