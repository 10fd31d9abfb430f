// Memory-bound glue kernels of the sampled-subnet block engine (all HBM-bound; coalesced,
// 16-byte vectorised where alignment allows, grids sized in multiples of the SM count).
//
//   patch im2col        AutoFormer/model/module/embedding_super.py:33-40 (conv16x16/16 == GEMM)
//   token assembly      AutoFormer/model/supernet_transformer.py:150-155 (cls cat + pos add)
//   mean pooling        supernet_transformer.py:164-165 (gp)
//   grad cast/scale     DropPath backward + fp32->bf16 (model/utils.py:71-99)
//   bias gradient       column sums of dY
//   RPE table packing   multihead_super.py:32-35 / irpe.py:483-496 tables -> 64x64 bf16 packs
#include <cstdlib>
#include "common.cuh"

namespace cb {
namespace {

inline int grid_for(int64_t work, int threads, int waves = 8) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(work, threads), kNumSMs * waves)));
}

// images (B, C, H, W) fp32 -> patches (B*gh*gw, C*P*P) bf16, col = c*P*P + ky*P + kx
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int64_t ldo, int B, int C,
              int H, int W, int P) {
  const int gh = H / P, gw = W / P;
  const int K = C * P * P;
  const int K4 = K >> 2;
  const int64_t total = static_cast<int64_t>(B) * gh * gw * K4;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k4 = static_cast<int>(t % K4);
    const int64_t row = t / K4;
    const int k = k4 << 2;
    const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
    const int pw = static_cast<int>(row % gw), ph = static_cast<int>((row / gw) % gh);
    const int64_t b = row / (gw * gh);
    const float4 v = __ldg(reinterpret_cast<const float4*>(
        img + ((b * C + c) * H + (ph * P + ky)) * static_cast<int64_t>(W) + pw * P + kx));
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + row * ldo + k) = o;
  }
}

// x[b,0,:] = cls + pos[0] ; x[b,1+t,:] = patch[b*T+t,:] + pos[1+t]
__global__ void __launch_bounds__(256)
assemble_fwd_kernel(const __nv_bfloat16* __restrict__ patch, int64_t ldp, const float* __restrict__ cls,
                    const float* __restrict__ pos, int64_t ldpos, float* __restrict__ x, int64_t ldx,
                    int B, int N, int E) {
  const int64_t total = static_cast<int64_t>(B) * N * E;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(t % E);
    const int64_t r = t / E;
    const int n = static_cast<int>(r % N);
    const int64_t b = r / N;
    float v = pos ? pos[static_cast<int64_t>(n) * ldpos + e] : 0.f;
    v += (n == 0) ? cls[e] : __bfloat162float(patch[(b * (N - 1) + n - 1) * ldp + e]);
    x[r * ldx + e] = v;
  }
}

// dpatch (bf16) = g[b,1+t,:] ; dpos[n,:] += sum_b g[b,n,:] ; dcls += sum_b g[b,0,:]
__global__ void __launch_bounds__(256)
assemble_bwd_kernel(const float* __restrict__ g, int64_t ldg, __nv_bfloat16* __restrict__ dpatch,
                    int64_t ldp, float* __restrict__ dpos, int64_t ldpos, float* __restrict__ dcls,
                    int B, int N, int E) {
  const int64_t total = static_cast<int64_t>(N) * E;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(t % E);
    const int n = static_cast<int>(t / E);
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float v = g[(static_cast<int64_t>(b) * N + n) * ldg + e];
      acc += v;
      if (n > 0) dpatch[(static_cast<int64_t>(b) * (N - 1) + n - 1) * ldp + e] = __float2bfloat16_rn(v);
    }
    if (dpos) dpos[static_cast<int64_t>(n) * ldpos + e] += acc;
    if (n == 0 && dcls) dcls[e] += acc;
  }
}

// pooled[b,e] = mean_{n>=first} y[b,n,e]  (first = 1 for gp, pooling skips cls)
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ y, int64_t ldy, __nv_bfloat16* __restrict__ out, int64_t ldo,
                int B, int N, int E, int first, int count) {
  const int64_t total = static_cast<int64_t>(B) * E;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(t % E);
    const int64_t b = t / E;
    float acc = 0.f;
    for (int n = first; n < first + count; ++n) acc += y[(b * N + n) * ldy + e];
    out[b * ldo + e] = __float2bfloat16_rn(acc / count);
  }
}

// dy[b,n,e] = dpooled[b,e] / count for first <= n < first+count, else 0
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const __nv_bfloat16* __restrict__ dp, int64_t lddp, float* __restrict__ dy, int64_t lddy,
                int B, int N, int E, int first, int count) {
  const int64_t total = static_cast<int64_t>(B) * N * E;
  const float inv = 1.0f / count;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(t % E);
    const int64_t r = t / E;
    const int n = static_cast<int>(r % N);
    const int64_t b = r / N;
    dy[r * lddy + e] = (n >= first && n < first + count) ? inv * __bfloat162float(dp[b * lddp + e]) : 0.f;
  }
}

// out_bf16[r,c] = bf16(scale[r / rows_per] * in[r,c]); optional fused bias gradient
// dbias[c] += sum_r out[r,c].  Block = 32 column threads (x4 columns) x 8 row lanes; the row lanes
// are reduced through shared memory so that only gridDim.y atomics hit each dbias address.
template <int U>
__global__ void __launch_bounds__(256)
cast_scale_kernel(const float* __restrict__ in, int64_t ldi, __nv_bfloat16* __restrict__ out, int64_t ldo,
                  const float* __restrict__ scale, int rows_per, float* __restrict__ dbias, int64_t rows,
                  int cols) {
  __shared__ float red[8][32][4];
  const int c = (blockIdx.x * 32 + threadIdx.x) << 2;
  const bool live = c < cols;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.y) * 8;
  if (live) {
    for (int64_t r0 = static_cast<int64_t>(blockIdx.y) * 8 + threadIdx.y; r0 < rows; r0 += U * stride) {
      float4 v[U];
      float s[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * stride;
        if (r < rows) {
          v[u] = __ldg(reinterpret_cast<const float4*>(in + r * ldi + c));
          s[u] = scale ? __ldg(scale + r / rows_per) : 1.0f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * stride;
        if (r < rows) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(s[u] * v[u].x, s[u] * v[u].y);
          const __nv_bfloat162 hi = __floats2bfloat162_rn(s[u] * v[u].z, s[u] * v[u].w);
          uint2 o;
          o.x = *reinterpret_cast<const uint32_t*>(&lo);
          o.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(out + r * ldo + c) = o;
          a0 += __bfloat162float(lo.x); a1 += __bfloat162float(lo.y);
          a2 += __bfloat162float(hi.x); a3 += __bfloat162float(hi.y);
        }
      }
    }
  }
  if (dbias == nullptr) return;
  __shared__ __align__(16) float outv[128];
  red[threadIdx.y][threadIdx.x][0] = a0; red[threadIdx.y][threadIdx.x][1] = a1;
  red[threadIdx.y][threadIdx.x][2] = a2; red[threadIdx.y][threadIdx.x][3] = a3;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x][k];
      outv[threadIdx.x * 4 + k] = t;
    }
  }
  const int c0 = blockIdx.x * 128;
  block_add_to_global(dbias + c0, outv, min(128, cols - c0));
}

// dbias[c] += sum_r dy[r,c]  (bf16 input).  Block = 32 column threads (x8 columns) x 8 row lanes.
template <int kU>
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ dy, int64_t ld, float* __restrict__ dbias, int64_t rows,
              int cols) {
  constexpr int U = kU;
  __shared__ float red[8][32][8];
  const int c = (blockIdx.x * 32 + threadIdx.x) << 3;
  const bool live = c < cols;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t stride = static_cast<int64_t>(gridDim.y) * 8;
  const bool vec = (c + 8 <= cols) && ((ld & 7) == 0);
  if (live) {
    for (int64_t r0 = static_cast<int64_t>(blockIdx.y) * 8 + threadIdx.y; r0 < rows; r0 += U * stride) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * stride;
        v[u] = make_uint4(0, 0, 0, 0);
        if (r < rows) {
          if (vec) v[u] = __ldg(reinterpret_cast<const uint4*>(dy + r * ld + c));
          else {
            __nv_bfloat16 tmp[8];
            for (int k = 0; k < 8; ++k) tmp[k] = (c + k < cols) ? dy[r * ld + c + k] : __float2bfloat16_rn(0.f);
            v[u] = *reinterpret_cast<uint4*>(tmp);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float2 f0 = unpack_bf16x2(v[u].x), f1 = unpack_bf16x2(v[u].y), f2 = unpack_bf16x2(v[u].z), f3 = unpack_bf16x2(v[u].w);
        acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
        acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
      }
    }
  }
  __shared__ __align__(16) float outv[256];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.y][threadIdx.x][k] = acc[k];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x][k];
      outv[threadIdx.x * 8 + k] = t;
    }
  }
  const int c0 = blockIdx.x * 256;
  block_add_to_global(dbias + c0, outv, min(256, cols - c0));
}

struct PackSrc {
  const float* src;
  float* grad;
  int nb, row_off;
  int64_t stride_b, stride_d;
};

// dst[T][64][64] bf16: rows [row_off, row_off+nb) <- src[t][b][d] (any strides), zeros elsewhere
__global__ void __launch_bounds__(256)
pack_tables_kernel(__nv_bfloat16* __restrict__ dst, PackSrc s0, PackSrc s1, int64_t table_stride0,
                   int64_t table_stride1, int D) {
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    float v = 0.f;
    if (d < D) {
      if (s0.src && r >= s0.row_off && r < s0.row_off + s0.nb)
        v = s0.src[t * table_stride0 + (r - s0.row_off) * s0.stride_b + d * s0.stride_d];
      else if (s1.src && r >= s1.row_off && r < s1.row_off + s1.nb)
        v = s1.src[t * table_stride1 + (r - s1.row_off) * s1.stride_b + d * s1.stride_d];
    }
    dst[(static_cast<int64_t>(t) * 64 + r) * 64 + d] = __float2bfloat16_rn(v);
  }
}

// grad[t][b][d] += dpack[t][row_off+b][d]
__global__ void __launch_bounds__(256)
unpack_grads_kernel(const float* __restrict__ dpack, PackSrc s0, PackSrc s1, int64_t table_stride0,
                    int64_t table_stride1, int D) {
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    if (d >= D) continue;
    const float v = dpack[(static_cast<int64_t>(t) * 64 + r) * 64 + d];
    if (s0.grad && r >= s0.row_off && r < s0.row_off + s0.nb)
      s0.grad[t * table_stride0 + (r - s0.row_off) * s0.stride_b + d * s0.stride_d] += v;
    else if (s1.grad && r >= s1.row_off && r < s1.row_off + s1.nb)
      s1.grad[t * table_stride1 + (r - s1.row_off) * s1.stride_b + d * s1.stride_d] += v;
  }
}

constexpr int kMaxBatchPacks = 64;
struct PackBatch {
  const float* src0[kMaxBatchPacks];
  const float* src1[kMaxBatchPacks];
  float* grad0[kMaxBatchPacks];
  float* grad1[kMaxBatchPacks];
};

// One launch for every (layer, k|v) pack of a step: pack p <- rows [0,nb) from src0[p], rows
// [row_off1, row_off1+nb) from src1[p] (tables (nb, D) with element strides), zeros elsewhere.
__global__ void __launch_bounds__(256)
pack_tables_batch_kernel(__nv_bfloat16* __restrict__ dst, const __grid_constant__ PackBatch b, int nb,
                         int row_off1, int64_t stride_b, int64_t stride_d, int D) {
  const int p = blockIdx.x;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    float v = 0.f;
    if (d < D) {
      if (r < nb) v = b.src0[p][r * stride_b + d * stride_d];
      else if (b.src1[p] != nullptr && r >= row_off1 && r < row_off1 + nb) v = b.src1[p][(r - row_off1) * stride_b + d * stride_d];
    }
    dst[(static_cast<int64_t>(p) * 64 + r) * 64 + d] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(256)
unpack_grads_batch_kernel(const float* __restrict__ dpack, const __grid_constant__ PackBatch b, int nb,
                          int row_off1, int64_t stride_b, int64_t stride_d, int D) {
  const int p = blockIdx.x;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    if (d >= D) continue;
    const float v = dpack[(static_cast<int64_t>(p) * 64 + r) * 64 + d];
    if (r < nb) b.grad0[p][r * stride_b + d * stride_d] += v;
    else if (b.grad1[p] != nullptr && r >= row_off1 && r < row_off1 + nb) b.grad1[p][(r - row_off1) * stride_b + d * stride_d] += v;
  }
}

}  // namespace
}  // namespace cb

using namespace cb;

extern "C" int cream_patch_im2col(const float* images, void* out_bf16, int64_t ldo, int B, int C, int H,
                                  int W, int P, void* stream) {
  CB_REQUIRE(images && out_bf16 && B > 0, "null pointer");
  CB_REQUIRE(H % P == 0 && W % P == 0 && P % 4 == 0 && W % 4 == 0 && ldo % 4 == 0, "patch geometry");
  const int64_t work = static_cast<int64_t>(B) * (H / P) * (W / P) * (C * P * P / 4);
  im2col_kernel<<<grid_for(work, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, static_cast<__nv_bfloat16*>(out_bf16), ldo, B, C, H, W, P);
  return check_last("im2col_kernel");
}

extern "C" int cream_tokens_assemble_fwd(const void* patch_bf16, int64_t ldp, const float* cls,
                                         const float* pos, int64_t ldpos, float* x, int64_t ldx, int B,
                                         int N, int E, void* stream) {
  CB_REQUIRE(patch_bf16 && cls && x && B > 0 && N > 1 && E > 0, "bad args");
  assemble_fwd_kernel<<<grid_for(static_cast<int64_t>(B) * N * E, 256, 16), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch_bf16), ldp, cls, pos, ldpos, x, ldx, B, N, E);
  return check_last("assemble_fwd_kernel");
}

extern "C" int cream_tokens_assemble_bwd(const float* g, int64_t ldg, void* dpatch_bf16, int64_t ldp,
                                         float* dpos, int64_t ldpos, float* dcls, int B, int N, int E,
                                         void* stream) {
  CB_REQUIRE(g && dpatch_bf16 && B > 0 && N > 1 && E > 0, "bad args");
  assemble_bwd_kernel<<<grid_for(static_cast<int64_t>(N) * E, 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(
      g, ldg, static_cast<__nv_bfloat16*>(dpatch_bf16), ldp, dpos, ldpos, dcls, B, N, E);
  return check_last("assemble_bwd_kernel");
}

extern "C" int cream_pool_fwd(const float* y, int64_t ldy, void* out_bf16, int64_t ldo, int B, int N,
                              int E, int first, int count, void* stream) {
  CB_REQUIRE(y && out_bf16 && count > 0 && first >= 0 && first + count <= N, "bad args");
  pool_fwd_kernel<<<grid_for(static_cast<int64_t>(B) * E, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, ldy, static_cast<__nv_bfloat16*>(out_bf16), ldo, B, N, E, first, count);
  return check_last("pool_fwd_kernel");
}

extern "C" int cream_pool_bwd(const void* dpooled_bf16, int64_t lddp, float* dy, int64_t lddy, int B,
                              int N, int E, int first, int count, void* stream) {
  CB_REQUIRE(dpooled_bf16 && dy && count > 0, "bad args");
  pool_bwd_kernel<<<grid_for(static_cast<int64_t>(B) * N * E, 256, 16), 256, 0,
                    static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dpooled_bf16), lddp, dy, lddy, B, N, E, first, count);
  return check_last("pool_bwd_kernel");
}


extern "C" int cream_cast_scale(const float* in, int64_t ldi, void* out_bf16, int64_t ldo,
                                const float* row_scale, int rows_per_scale, float* dbias, int64_t rows,
                                int cols, void* stream) {
  if (rows == 0 || cols == 0) return CREAM_OK;
  CB_REQUIRE(in && out_bf16, "null pointer");
  CB_REQUIRE(ldi % 4 == 0 && ldo % 4 == 0 && cols % 4 == 0, "cols and pitches must be multiples of 4");
  // (unroll depth 2 / 4 / 8 and 2 - 8 blocks per SM were swept at the supernet shapes: within 5 % of each other)
  dim3 block(32, 8), grid(ceil_div(cols / 4, 32), static_cast<unsigned>(std::min<int64_t>(ceil_div64(rows, 16), kNumSMs * 2)));
  cast_scale_kernel<2><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      in, ldi, static_cast<__nv_bfloat16*>(out_bf16), ldo, row_scale, rows_per_scale > 0 ? rows_per_scale : 1, dbias, rows, cols);
  return check_last("cast_scale_kernel");
}

extern "C" int cream_bias_grad(const void* dy_bf16, int64_t ld, float* dbias, int64_t rows, int cols,
                               void* stream) {
  if (rows == 0 || cols == 0) return CREAM_OK;
  CB_REQUIRE(dy_bf16 && dbias && ld % 2 == 0, "bad args");
  dim3 block(32, 8), grid(ceil_div(ceil_div(cols, 8), 32), static_cast<unsigned>(std::min<int64_t>(ceil_div64(rows, 16), kNumSMs * 2)));
  colsum_kernel<2><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dy_bf16), ld, dbias, rows, cols);
  return check_last("colsum_kernel");
}

extern "C" int cream_pack_tables(void* dst_bf16, int num_tables, int head_dim, const float* src0, int nb0,
                                 int row_off0, int64_t stride_t0, int64_t stride_b0, int64_t stride_d0,
                                 const float* src1, int nb1, int row_off1, int64_t stride_t1,
                                 int64_t stride_b1, int64_t stride_d1, void* stream) {
  CB_REQUIRE(dst_bf16 && num_tables > 0 && head_dim > 0 && head_dim <= 64, "bad args");
  CB_REQUIRE(row_off0 + nb0 <= 64 && row_off1 + nb1 <= 64, "packed rows exceed 64");
  PackSrc a{src0, nullptr, nb0, row_off0, stride_b0, stride_d0};
  PackSrc b{src1, nullptr, nb1, row_off1, stride_b1, stride_d1};
  pack_tables_kernel<<<num_tables, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(dst_bf16), a, b, stride_t0, stride_t1, head_dim);
  return check_last("pack_tables_kernel");
}

extern "C" int cream_unpack_table_grads(const float* dpack, int num_tables, int head_dim, float* grad0,
                                        int nb0, int row_off0, int64_t stride_t0, int64_t stride_b0,
                                        int64_t stride_d0, float* grad1, int nb1, int row_off1,
                                        int64_t stride_t1, int64_t stride_b1, int64_t stride_d1,
                                        void* stream) {
  CB_REQUIRE(dpack && num_tables > 0 && head_dim > 0 && head_dim <= 64, "bad args");
  PackSrc a{nullptr, grad0, nb0, row_off0, stride_b0, stride_d0};
  PackSrc b{nullptr, grad1, nb1, row_off1, stride_b1, stride_d1};
  unpack_grads_kernel<<<num_tables, 256, 0, static_cast<cudaStream_t>(stream)>>>(dpack, a, b, stride_t0,
                                                                              stride_t1, head_dim);
  return check_last("unpack_grads_kernel");
}

extern "C" int cream_pack_tables_batch(void* dst_bf16, int n_packs, int head_dim, const float* const* src0_host,
                                       const float* const* src1_host, int nb, int row_off1, int64_t stride_b,
                                       int64_t stride_d, void* stream) {
  CB_REQUIRE(dst_bf16 && src0_host && n_packs > 0 && n_packs <= kMaxBatchPacks, "1..64 packs per call");
  CB_REQUIRE(head_dim > 0 && head_dim <= 64 && nb > 0 && nb <= 64, "pack geometry");
  if (src1_host != nullptr) CB_REQUIRE(row_off1 + nb <= 64 && row_off1 >= nb, "second table must fit rows [row_off1, 64)");
  PackBatch b{};
  for (int i = 0; i < n_packs; ++i) {
    b.src0[i] = src0_host[i];
    b.src1[i] = src1_host ? src1_host[i] : nullptr;
  }
  pack_tables_batch_kernel<<<n_packs, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(dst_bf16), b, nb, row_off1, stride_b, stride_d, head_dim);
  return check_last("pack_tables_batch_kernel");
}

extern "C" int cream_unpack_table_grads_batch(const float* dpack, int n_packs, int head_dim, float* const* grad0_host,
                                              float* const* grad1_host, int nb, int row_off1, int64_t stride_b,
                                              int64_t stride_d, void* stream) {
  CB_REQUIRE(dpack && grad0_host && n_packs > 0 && n_packs <= kMaxBatchPacks, "1..64 packs per call");
  CB_REQUIRE(head_dim > 0 && head_dim <= 64 && nb > 0 && nb <= 64, "pack geometry");
  if (grad1_host != nullptr) CB_REQUIRE(row_off1 + nb <= 64 && row_off1 >= nb, "second table must fit rows [row_off1, 64)");
  PackBatch b{};
  for (int i = 0; i < n_packs; ++i) {
    b.grad0[i] = grad0_host[i];
    b.grad1[i] = grad1_host ? grad1_host[i] : nullptr;
  }
  unpack_grads_batch_kernel<<<n_packs, 256, 0, static_cast<cudaStream_t>(stream)>>>(dpack, b, nb, row_off1, stride_b,
                                                                                  stride_d, head_dim);
  return check_last("unpack_grads_batch_kernel");
}
