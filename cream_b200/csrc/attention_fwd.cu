// Fused attention forward with relative-position terms, sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces, per layer, the five batched GEMMs + two index gathers + softmax of
//   AutoFormer/model/module/multihead_super.py:133-154  (AttentionSuper.forward)
//   iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:73-92 (RPEAttention.forward, rpe on k / v,
//   contextual or bias mode) + irpe.py:585-687 + rpe_ops/rpe_index_cuda.cu
// by ONE kernel that reads q,k,v once and writes o once.
//
// "Bucket formulation": a contextual RPE term  q_i . T[idx[i,j]]  is computed as
//   R = Q . Tpack^T  (one extra 128x64x64 MMA, Tpack = up to two tables packed in 64 rows)
// followed by a per-row gather  S[i,j] += R[i, idx[i,j]]  staged through shared memory and
// fused into the softmax pass; the value-side term  sum_j P[i,j] T'[idx[i,j]]  becomes
//   PB[i,b] = sum_{j: idx[i,j]=b} P[i,j]   (bucket sums, thread-local)   and
//   O = [P | PB] . [V ; T'pack]           (a single chain of tcgen05.mma, A from TMEM).
//
// CTA = (query tile of 128 rows, head, batch); 5 warps; 256 TMEM columns and ~101 KB of
// shared memory so that two CTAs are resident per SM and overlap each other's phases.
//   warp 0 (one lane): TMA loads, tcgen05.mma issue
//   warps 1..4       : one thread per query row: R copy-out, softmax, epilogue
// TMEM columns: S fp32 [0,Npad) -> P bf16x2 in place [0,Npad/2) | PB [Npad/2,Npad/2+32);
//               R fp32 [192,256) (dead before S cols >= 192 are produced); O fp32 [192,256).
#include <cstdlib>
#include <cstring>

#include "attention_common.cuh"

namespace cb {
namespace {

constexpr int kD = 64;              // head dim
constexpr int kNB = 64;             // packed bucket rows (two tables of <= 32, or one of <= 64)
constexpr int kThreads = 160;
constexpr int kRStride = 66;        // halfs per row of the staged R tile (bank spread)
constexpr int kPBStride = 65;       // floats per row of the bucket-sum tile (generic gather path)
constexpr int kPBStrideAF = 68;     // floats per row, 16-byte aligned rows (structured path)

struct AttnFwdParams {
  int B, H, N, Npad;
  float scale;
  int ctx_k, ctx_v;                    // contextual tables on K / V present
  int shared_tables;                   // 1: one table pack for all heads
  int af_grid, af_max_rel;             // AutoFormer structured mode (0 = generic index tables)
  int af_mma;                          // structured mode through the tensor cores (see softmax_plain)
  int gp_grid, gp_w, gp_skip;          // iRPE grid-product structured mode (0 = off)
  uint8_t lut_a[32], lut_b[32];        // bucket row / column components by (delta + grid - 1)
  const uint8_t* idx_a; const uint8_t* idx_b;   // K-side gather indices (N, ldi), values < 64
  const uint8_t* idx_va; const uint8_t* idx_vb; // V-side
  int ldi;
  const float* bias;                   // (H or 1, 64) pre-packed bias table, gathered by idx_a
  const float* dense; int64_t dense_sb, dense_sh, dense_si;   // dense additive logit term (generic path)
  int causal;                          // generic path: key j > query i masked
  int block_len;                       // generic path: block-diagonal attention over items of block_len tokens (0 = off)
  __nv_bfloat16* out; int64_t ldo;     // (B*N, H*64)
  int wide_out;                        // out rows are 32-byte aligned: 256-bit stores
  float* lse;                          // (B, H, N)
};

// ---------------------------------------------------------------------------------------
// Softmax over one query row held in TMEM (one thread per row), generic gather tables.
//   pass 1: t = scale*S + R[idx_a] + R[idx_b] (+ bias[idx_a]); running max; t written back
//   pass 2: p = exp(t - max); row sum; P -> bf16x2 in place; bucket sums PB[idx_v*] += p
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax_generic(const AttnFwdParams& p, uint32_t trow, uint32_t sr_row,
                                                uint32_t spb_row, uint32_t sbias, int row_c, const float* drow,
                                                float& mx_out, float& sum_out) {
  const int Npad = p.Npad;
  const int nchunks = Npad / 16;
  const uint8_t* ia = p.idx_a ? p.idx_a + static_cast<int64_t>(row_c) * p.ldi : nullptr;
  const uint8_t* ib = p.idx_b ? p.idx_b + static_cast<int64_t>(row_c) * p.ldi : nullptr;
  const bool use_bias = p.bias != nullptr;
  // visible keys of this row: [k_lo, k_hi) - its own block when items are packed, up to itself under the causal mask
  int k_lo = 0, k_hi = p.N;
  if (p.block_len > 0) { k_lo = (row_c / p.block_len) * p.block_len; k_hi = k_lo + p.block_len; }
  if (p.causal) k_hi = min(k_hi, row_c + 1);
  float mx = -INFINITY;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t raw[16];
    tmem_ld16(trow + c * 16, raw);
    uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
    if (ia) va = __ldg(reinterpret_cast<const uint4*>(ia + c * 16));
    if (ib) vb = __ldg(reinterpret_cast<const uint4*>(ib + c * 16));
    tmem_ld_wait();
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w};
    const uint32_t wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float t = p.scale * __uint_as_float(raw[k]);
      const uint32_t a_id = byte_of(wa, k);
      if (p.ctx_k) {
        if (ia) t += lds_f16(sr_row + 2 * a_id);
        if (ib) t += lds_f16(sr_row + 2 * byte_of(wb, k));
      }
      if (use_bias) t += lds_f32(sbias + 4 * a_id);
      if (drow != nullptr && c * 16 + k < p.N) t += __ldg(drow + c * 16 + k);
      if (c * 16 + k < k_lo || c * 16 + k >= k_hi) t = -INFINITY;
      mx = fmaxf(mx, t);
      raw[k] = __float_as_uint(t);
    }
    tmem_st16(trow + c * 16, raw);
  }
  tmem_st_wait();

  if (p.ctx_v) {
    for (int k = 0; k < kNB; ++k) sts_f32(spb_row + 4 * k, 0.f);
  }
  const uint8_t* iva = p.idx_va ? p.idx_va + static_cast<int64_t>(row_c) * p.ldi : nullptr;
  const uint8_t* ivb = p.idx_vb ? p.idx_vb + static_cast<int64_t>(row_c) * p.ldi : nullptr;
  const float mxl = mx * kLog2e;
  float sum = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t raw[16];
    tmem_ld16(trow + c * 16, raw);
    uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
    if (iva) va = __ldg(reinterpret_cast<const uint4*>(iva + c * 16));
    if (ivb) vb = __ldg(reinterpret_cast<const uint4*>(ivb + c * 16));
    tmem_ld_wait();
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w};
    const uint32_t wb[4] = {vb.x, vb.y, vb.z, vb.w};
    float pv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      pv[k] = fast_exp2(fmaf(__uint_as_float(raw[k]), kLog2e, -mxl));
      sum += pv[k];
    }
    if (p.ctx_v) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        // padded key columns have p == 0 exactly, so their (zero) ids are harmless
        if (iva) { const uint32_t a = spb_row + 4 * byte_of(wa, k); sts_f32(a, lds_f32(a) + pv[k]); }
        if (ivb) { const uint32_t a = spb_row + 4 * byte_of(wb, k); sts_f32(a, lds_f32(a) + pv[k]); }
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
    tmem_st8(trow + c * 8, pk);
  }
  if (p.ctx_v) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int k = 0; k < 16; ++k)
        pk[k] = pack_bf16x2(lds_f32(spb_row + 4 * (c * 32 + 2 * k)), lds_f32(spb_row + 4 * (c * 32 + 2 * k + 1)));
      tmem_st16(trow + Npad / 2 + c * 16, pk);
    }
  }
  mx_out = mx;
  sum_out = sum;
}

// ---------------------------------------------------------------------------------------
// AutoFormer-structured softmax (multihead_super.py:40-59): for a G x G patch grid + cls,
//   idx_v[i,j] = rj - ri + M1,  idx_h[i,j] = cj - ci + M1   (i,j >= 1; M1 = max_rel + 1;
//   the clamp never binds when G-1 <= max_rel), and bucket 0 for the cls row / column.
// The 2*G gather operands of a row live in registers; with the column loop fully unrolled
// (rj, cj) are compile-time, so the gather costs two FADDs per element and the value-side
// bucket sums are row / column sums kept in registers (no shared-memory read-modify-write).
// ---------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void softmax_af(const AttnFwdParams& p, uint32_t trow, uint32_t sr_row,
                                           uint32_t spb_row, int row, float& mx_out, float& sum_out) {
  constexpr int N = G * G + 1;
  constexpr int NPAD = (N + 15) / 16 * 16;
  constexpr int NCH = NPAD / 16;
  const int M1 = p.af_max_rel + 1;
  const bool patch = row >= 1 && row < N;
  const int qi = patch ? row - 1 : 0;
  const int ri = qi / G, ci = qi - ri * G;
  const float r0v = lds_f16(sr_row), r0h = lds_f16(sr_row + 2 * 32);
  const float r0 = r0v + r0h;
  float rv[G], rh[G];
#pragma unroll
  for (int t = 0; t < G; ++t) {
    rv[t] = patch ? lds_f16(sr_row + 2 * (M1 - ri + t)) : r0v;
    rh[t] = patch ? lds_f16(sr_row + 2 * (32 + M1 - ci + t)) : r0h;
  }
  float mx = -INFINITY;
  // accumulator reads run one chunk ahead of the arithmetic in both passes (tcgen05.wait::ld covers
  // every load issued before it, so the next chunk is requested right after the wait)
  uint32_t rb[2][16];
  tmem_ld16(trow, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t (&raw)[16] = rb[c & 1];
    tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld16(trow + (c + 1) * 16, rb[(c + 1) & 1]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = c * 16 + k;
      float t;
      if (j >= N) t = -INFINITY;
      else if (j == 0) t = fmaf(p.scale, __uint_as_float(raw[k]), r0);
      else t = fmaf(p.scale, __uint_as_float(raw[k]), rv[(j - 1) / G]) + rh[(j - 1) % G];
      mx = fmaxf(mx, t);
      raw[k] = __float_as_uint(t);
    }
    tmem_st16(trow + c * 16, raw);
  }
  tmem_st_wait();

  float prow[G], pcol[G], p0 = 0.f, sum = 0.f;
#pragma unroll
  for (int t = 0; t < G; ++t) { prow[t] = 0.f; pcol[t] = 0.f; }
  const float mxl = mx * kLog2e;
  tmem_ld16(trow, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t (&raw)[16] = rb[c & 1];
    tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld16(trow + (c + 1) * 16, rb[(c + 1) & 1]);
    float pv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = c * 16 + k;
      pv[k] = fast_exp2(fmaf(__uint_as_float(raw[k]), kLog2e, -mxl));
      sum += pv[k];
      if (j == 0) p0 = pv[k];
      else if (j < N) { prow[(j - 1) / G] += pv[k]; pcol[(j - 1) % G] += pv[k]; }
    }
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
    tmem_st8(trow + c * 8, pk);
  }
  // bucket sums -> (thread-private) shared row -> bf16x2 -> TMEM columns [NPAD/2, NPAD/2+32)
#pragma unroll
  for (int q = 0; q < 16; ++q) sts_f32x4(spb_row + 16 * q, make_float4(0.f, 0.f, 0.f, 0.f));
  if (patch) {
#pragma unroll
    for (int t = 0; t < G; ++t) {
      sts_f32(spb_row + 4 * (M1 - ri + t), prow[t]);
      sts_f32(spb_row + 4 * (32 + M1 - ci + t), pcol[t]);
    }
    sts_f32(spb_row, p0);
    sts_f32(spb_row + 4 * 32, p0);
  } else {
    sts_f32(spb_row, sum);
    sts_f32(spb_row + 4 * 32, sum);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t pk[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = lds_f32x4(spb_row + 4 * (c * 32 + 4 * q));
      pk[2 * q] = pack_bf16x2(v.x, v.y);
      pk[2 * q + 1] = pack_bf16x2(v.z, v.w);
    }
    tmem_st16(trow + NPAD / 2 + c * 16, pk);
  }
  mx_out = mx;
  sum_out = sum;
}

// ---------------------------------------------------------------------------------------
// iRPE product method on a G x G grid + cls (irpe.py:176-202), contextual table on keys only:
//   id(i, j) = A[rj - ri] * W + B[cj - ci]   for patch tokens,  the skip bucket when i or j is cls.
// The byte offsets of a row's 2 x G components live in registers; with the column loop fully
// unrolled one element costs an integer add, a 16-bit shared load and an FFMA - no index bytes
// from global memory, no byte extraction.  (V-side tables take the generic path.)
// ---------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void softmax_gridprod(const AttnFwdParams& p, uint32_t trow, uint32_t sr_row, uint32_t slut,
                                                 int row, float& mx_out, float& sum_out) {
  constexpr int N = G * G + 1;
  constexpr int NPAD = (N + 15) / 16 * 16;
  constexpr int NCH = NPAD / 16;
  const bool patch = row >= 1 && row < N;
  const int qi = patch ? row - 1 : 0;
  const int ri = qi / G, ci = qi - ri * G;
  const uint32_t skip_addr = sr_row + 2 * p.gp_skip;
  uint32_t base_a[G], off_b[G];
#pragma unroll
  for (int t = 0; t < G; ++t) {
    uint32_t a, b;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(a) : "r"(slut + (t - ri + G - 1)));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(slut + 32 + (t - ci + G - 1)));
    base_a[t] = patch ? sr_row + 2 * a * p.gp_w : skip_addr;
    off_b[t] = patch ? 2 * b : 0;
  }
  const float r0 = lds_f16(skip_addr);
  float mx = -INFINITY;
  uint32_t rb[2][16];
  tmem_ld16(trow, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t (&raw)[16] = rb[c & 1];
    tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld16(trow + (c + 1) * 16, rb[(c + 1) & 1]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = c * 16 + k;
      float t;
      if (j >= N) t = -INFINITY;
      else if (j == 0) t = fmaf(p.scale, __uint_as_float(raw[k]), r0);
      else t = fmaf(p.scale, __uint_as_float(raw[k]), lds_f16(base_a[(j - 1) / G] + off_b[(j - 1) % G]));
      mx = fmaxf(mx, t);
      raw[k] = __float_as_uint(t);
    }
    tmem_st16(trow + c * 16, raw);
  }
  tmem_st_wait();
  float sum = 0.f;
  const float mxl = mx * kLog2e;
  tmem_ld16(trow, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t (&raw)[16] = rb[c & 1];
    tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld16(trow + (c + 1) * 16, rb[(c + 1) & 1]);
    float pv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      pv[k] = fast_exp2(fmaf(__uint_as_float(raw[k]), kLog2e, -mxl));
      sum += pv[k];
    }
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
    tmem_st8(trow + c * 8, pk);
  }
  mx_out = mx;
  sum_out = sum;
}

// ---------------------------------------------------------------------------------------
// AutoFormer relative position THROUGH THE TENSOR CORES ("absolute-coordinate" formulation).
// For a query at grid position (ri, ci) the K-side term of key (rj, cj) is
//     Rv[rj - ri + M1] + Rh[cj - ci + M1]      (Rv | Rh = the two halves of R = Q . TKpack^T)
// i.e. a per-query vector a_i over 29 ABSOLUTE key features (14 grid rows, 14 grid columns, the cls
// key) times a 0/1 matrix Ind (features x keys) that does not depend on the query:
//     S += A . Ind,     A[i] = [ Rv[M1 - ri + t] (t < 14) | Rh[M1 - ci + t] (t < 14) | Rv[0] + Rh[0] | 0 0 0 ].
// A is built once per row (a shifted copy of 28 of its 64 R values), split into bf16 hi + lo parts
// (exact for the fp16-staged R) and added into the S accumulator by two K = 32 MMAs; the softmax is
// then a PLAIN softmax - no gather adds, no compile-time (rj, cj) - and the value-side bucket sums are
// the same matrix used the other way round:  PBabs = P . Ind^T  (one N = 32 MMA), un-shifted per row.
// ---------------------------------------------------------------------------------------
template <int NPAD>
__device__ __forceinline__ void softmax_plain(float scale, int N, uint32_t trow, float& mx_out, float& sum_out) {
  constexpr int NCH = NPAD / 16;
  float mx = -INFINITY;
  uint32_t rb[2][16];
  tmem_ld16(trow, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t (&raw)[16] = rb[c & 1];
    tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld16(trow + (c + 1) * 16, rb[(c + 1) & 1]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = c * 16 + k;
      if (j < N) mx = fmaxf(mx, __uint_as_float(raw[k]));
    }
  }
  mx *= scale;                                  // scale > 0
  const float sl = scale * kLog2e, mxl = mx * kLog2e;
  float sum = 0.f;
  tmem_ld16(trow, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t (&raw)[16] = rb[c & 1];
    tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld16(trow + (c + 1) * 16, rb[(c + 1) & 1]);
    float pv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = c * 16 + k;
      pv[k] = j < N ? fast_exp2(fmaf(__uint_as_float(raw[k]), sl, -mxl)) : 0.f;
      sum += pv[k];
    }
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
    tmem_st8(trow + c * 8, pk);
  }
  mx_out = mx;
  sum_out = sum;
}

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                const __grid_constant__ CUtensorMap map_tk, const __grid_constant__ CUtensorMap map_tv,
                const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  pdl_trigger();
  const int kv_bytes = p.Npad * 128;
  const int k_slot = max(kv_bytes, 11 * 1024);  // sQ + sTK + sK must hold the 128 x 68 fp32 PB overlay
  uint8_t* sQ = smem;
  uint8_t* sTK = sQ + 16384;
  uint8_t* sK = sTK + 8192;
  uint8_t* sV = sK + k_slot;         // [V ; TV] contiguous rows
  uint8_t* sTV = sV + kv_bytes;
  uint8_t* sR = sTV + 8192;          // fp16 [128][kRStride]
  uint8_t* sBias = sR + 128 * kRStride * 2;
  uint8_t* sLut = sBias + 64 * 4;    // 64 bytes: grid-product row / column components
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLut + 64);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_r = bars + 2;
  uint64_t* bar_rfree = bars + 3;
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_o = bars + 6;
  uint64_t* bar_pb = bars + 7;       // MMA -> rows: PBabs = P . Ind^T ready (tensor-core structured mode)
  uint64_t* bar_p2 = bars + 8;       // rows -> MMA: packed bucket sums written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int Npad = p.Npad;
  const int n_a = min(Npad, 192);     // S columns produced before R is released
  const int n_b = Npad - n_a;         // remaining S columns (overlap the R region)

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_r, 1);
    mbar_init(bar_rfree, 128);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_pb, 1);
    mbar_init(bar_p2, 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  pdl_wait();      // barrier init / TMEM allocation above overlap the previous kernel's tail; global memory from here on
  if (p.bias != nullptr && threadIdx.x >= 32 && threadIdx.x < 96)
    sts_f32(smem_u32(sBias) + 4 * (threadIdx.x - 32),
            p.bias[(p.shared_tables ? 0 : head) * 64 + threadIdx.x - 32]);
  if (p.gp_grid != 0 && threadIdx.x >= 96 && threadIdx.x < 160)
    sLut[threadIdx.x - 96] = threadIdx.x < 128 ? p.lut_a[threadIdx.x - 96] : p.lut_b[threadIdx.x - 128];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tab = p.shared_tables ? 0 : head;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- loads ----------------
      const int qcol = head * kD, kcol = (p.H + head) * kD, vcol = (2 * p.H + head) * kD;
      uint32_t qk_bytes = 16384 + kv_bytes + (p.ctx_k ? 8192 : 0);
      mbar_arrive_expect_tx(bar_qk, qk_bytes);
      tma_load_3d(sQ, &map_q, bar_qk, qcol, m0, b);
      tma_load_3d(sK, &map_kv, bar_qk, kcol, 0, b);
      if (p.ctx_k) tma_load_3d(sTK, &map_tk, bar_qk, 0, 0, tab);
      mbar_arrive_expect_tx(bar_v, kv_bytes + (p.ctx_v ? 8192 : 0));
      tma_load_3d(sV, &map_kv, bar_v, vcol, 0, b);
      if (p.ctx_v) tma_load_3d(sTV, &map_tv, bar_v, 0, 0, tab);

      // ---------------- S = Q K^T (+ R = Q TK^T) ----------------
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aTK = smem_u32(sTK);
      if (p.ctx_k) {
        const uint32_t id_r = umma_idesc_bf16(128, kNB, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + 192, umma_smem_desc_sw128(aQ + k * 32, 16, 1024),
                  umma_smem_desc_sw128(aTK + k * 32, 16, 1024), id_r, k > 0);
        umma_commit(bar_r);
      }
      {
        const uint32_t id_s = umma_idesc_bf16(128, n_a, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem, umma_smem_desc_sw128(aQ + k * 32, 16, 1024),
                  umma_smem_desc_sw128(aK + k * 32, 16, 1024), id_s, k > 0);
      }
      if (n_b > 0) {
        if (p.ctx_k) {
          mbar_wait(bar_rfree, 0);
          tc_fence_after();
        }
        const uint32_t id_s = umma_idesc_bf16(128, n_b, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + 192, umma_smem_desc_sw128(aQ + k * 32, 16, 1024),
                  umma_smem_desc_sw128(aK + 192 * 128 + k * 32, 16, 1024), id_s, k > 0);
      }
      const uint32_t aInd = smem_u32(sR);      // the 0/1 feature matrix replaces the staged R (tensor-core mode)
      if (p.af_mma) {
        // S += A_hi . Ind + A_lo . Ind   (A in TMEM columns [224, 240) / [240, 256), two K = 16 steps each)
        const uint32_t id_g = umma_idesc_bf16(128, Npad, 0, 1);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int k = 0; k < kFeat / 16; ++k)
            umma_ts(tmem, tmem + 224 + 16 * part + 8 * k, umma_smem_desc_sw128(aInd + k * 2048, kIndChunk, 1024), id_g, 1u);
      }
      umma_commit(bar_s);

      // ---------------- O = [P | PB] [V ; TV] ----------------
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t aV = smem_u32(sV);
      const uint32_t id_o = umma_idesc_bf16(128, kD, 0, 1);
      const int ksteps = (Npad + (p.ctx_v ? kNB : 0)) / 16;
      if (p.af_mma) {
        // bucket sums in absolute coordinates: PBabs[128 x 32] = P . Ind^T -> columns [112, 144)
        const uint32_t id_pb = umma_idesc_bf16(128, kFeat, 0, 0);
        for (int k = 0; k < Npad / 16; ++k)
          umma_ts(tmem + 112, tmem + 8 * k, umma_smem_desc_sw128(aInd + (k >> 2) * kIndChunk + (k & 3) * 32, 16, 1024), id_pb, k > 0);
        umma_commit(bar_pb);
        for (int k = 0; k < Npad / 16; ++k)      // O = P . V while the rows un-shift their bucket sums
          umma_ts(tmem + 192, tmem + 8 * k, umma_smem_desc_sw128(aV + k * 2048, 8192, 1024), id_o, k > 0);
        mbar_wait(bar_p2, 0);
        tc_fence_after();
        for (int k = Npad / 16; k < ksteps; ++k)  // O += PB . TV
          umma_ts(tmem + 192, tmem + 8 * k, umma_smem_desc_sw128(aV + k * 2048, 8192, 1024), id_o, 1u);
      } else {
        for (int k = 0; k < ksteps; ++k)
          umma_ts(tmem + 192, tmem + 8 * k, umma_smem_desc_sw128(aV + k * 2048, 8192, 1024), id_o, k > 0);
      }
      umma_commit(bar_o);
    }
  } else {
    // =================== softmax warps: one thread per query row ===================
    const int quarter = warp & 3;
    const int r_local = quarter * 32 + lane;
    const int row = m0 + r_local;                 // query index within the image
    const int row_c = min(row, p.N - 1);          // clamp for index-table reads of padding rows
    const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t sr_row = smem_u32(sR) + r_local * kRStride * 2;

    if (p.ctx_k) {
      mbar_wait(bar_r, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld32(trow + 192 + c * 32, raw);
        tmem_ld_wait();
        const float rs = p.af_mma ? 1.0f : p.scale;   // tensor-core mode adds the UNSCALED term into S
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __half2 h2 = __floats2half2_rn(rs * __uint_as_float(raw[2 * i]), rs * __uint_as_float(raw[2 * i + 1]));
          sts_u32(sr_row + 4 * (c * 16 + i), *reinterpret_cast<const uint32_t*>(&h2));
        }
      }
      if (p.af_mma) {
        // A[i]: this row's 29 absolute features, a shifted copy of its staged R values; bf16 hi + lo
        constexpr int G = 14;
        const int M1 = p.af_max_rel + 1;
        const bool patch = row >= 1 && row < p.N;
        const int qi = patch ? row - 1 : 0;
        const int ri = qi / G, ci = qi - ri * G;
        const float r0v = lds_f16(sr_row), r0h = lds_f16(sr_row + 2 * 32);
        float a[kFeat];
#pragma unroll
        for (int t = 0; t < G; ++t) {
          a[t] = patch ? lds_f16(sr_row + 2 * (M1 - ri + t)) : r0v;
          a[G + t] = patch ? lds_f16(sr_row + 2 * (32 + M1 - ci + t)) : r0h;
        }
        a[2 * G] = r0v + r0h;
#pragma unroll
        for (int t = 2 * G + 1; t < kFeat; ++t) a[t] = 0.f;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(a[2 * t]), h1 = __float2bfloat16_rn(a[2 * t + 1]);
          hi[t] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
          lo[t] = pack_bf16x2(a[2 * t] - __bfloat162float(h0), a[2 * t + 1] - __bfloat162float(h1));
        }
        tmem_st16(trow + 224, hi);
        tmem_st16(trow + 240, lo);
        tmem_st_wait();
        asm volatile("bar.sync 1, 128;" ::: "memory");          // every row has read its staged R
        write_ind_matrix(smem_u32(sR), threadIdx.x - 32, 128, G, p.N);
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(bar_rfree);
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();

    float mx, sum;
    if (p.af_mma) {
      constexpr int G = 14;
      softmax_plain<208>(p.scale, p.N, trow, mx, sum);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
      // un-shift the absolute bucket sums into this row's 64 packed buckets
      const uint32_t spb_row = smem_u32(smem) + r_local * kPBStrideAF * 4;
      const int M1 = p.af_max_rel + 1;
      const bool patch = row >= 1 && row < p.N;
      const int qi = patch ? row - 1 : 0;
      const int ri = qi / G, ci = qi - ri * G;
#pragma unroll
      for (int q = 0; q < 16; ++q) sts_f32x4(spb_row + 16 * q, make_float4(0.f, 0.f, 0.f, 0.f));
      mbar_wait(bar_pb, 0);
      tc_fence_after();
      uint32_t pb[32];
      tmem_ld32(trow + 112, pb);
      tmem_ld_wait();
      if (patch) {
#pragma unroll
        for (int t = 0; t < G; ++t) {
          sts_f32(spb_row + 4 * (M1 - ri + t), __uint_as_float(pb[t]));
          sts_f32(spb_row + 4 * (32 + M1 - ci + t), __uint_as_float(pb[G + t]));
        }
        sts_f32(spb_row, __uint_as_float(pb[2 * G]));
        sts_f32(spb_row + 4 * 32, __uint_as_float(pb[2 * G]));
      } else {
        sts_f32(spb_row, sum);
        sts_f32(spb_row + 4 * 32, sum);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = lds_f32x4(spb_row + 4 * (c * 32 + 4 * q));
          pk[2 * q] = pack_bf16x2(v.x, v.y);
          pk[2 * q + 1] = pack_bf16x2(v.z, v.w);
        }
        tmem_st16(trow + 208 / 2 + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p2);
    } else if (p.af_grid == 14) {
      softmax_af<14>(p, trow, sr_row, smem_u32(smem) + r_local * kPBStrideAF * 4, row, mx, sum);
    } else if (p.gp_grid == 14) {
      softmax_gridprod<14>(p, trow, sr_row, smem_u32(sLut), row, mx, sum);
    } else {
      const float* drow = p.dense ? p.dense + b * p.dense_sb + head * p.dense_sh + row_c * p.dense_si : nullptr;
      softmax_generic(p, trow, sr_row, smem_u32(smem) + r_local * kPBStride * 4, smem_u32(sBias), row_c, drow, mx, sum);
    }
    if (!p.af_mma) {
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
    }

    // ---- epilogue: O / sum -> bf16 -> global ; log-sum-exp ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    __nv_bfloat16* orow = p.out + (static_cast<int64_t>(b) * p.N + row) * p.ldo + head * kD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld32(trow + 192 + c * 32, raw);
      tmem_ld_wait();
      if (row < p.N) store_row32_bf16(orow + c * 32, raw, inv, p.wide_out != 0);
    }
    if (row < p.N && p.lse != nullptr)
      p.lse[(static_cast<int64_t>(b) * p.H + head) * p.N + row] = mx + __logf(sum);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

}  // namespace
}  // namespace cb

extern "C" int cream_attn_fwd(const cream_attn_desc* d, void* stream_) {
  using namespace cb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(d != nullptr && d->qkv != nullptr && d->out != nullptr, "null pointer");
  CB_REQUIRE(d->head_dim == kD, "head_dim must be 64");
  CB_REQUIRE(d->N >= 1 && d->N <= 208, "1 <= tokens <= 208");
  CB_REQUIRE(d->B >= 1 && d->H >= 1, "empty batch");
  CB_REQUIRE(d->ld_qkv % 8 == 0 && d->ld_out % 8 == 0, "leading dims % 8");
  const int Npad = round_up(d->N, 16);
  const int ldi = d->ld_idx;
  if (d->idx_a || d->idx_va) CB_REQUIRE(ldi >= Npad && ldi % 16 == 0, "index pitch must be >= Npad and % 16");
  const bool ctx_k = d->tk_pack != nullptr, ctx_v = d->tv_pack != nullptr;
  if (ctx_k) CB_REQUIRE(d->idx_a != nullptr, "K tables need idx_a");
  if (ctx_v) CB_REQUIRE(d->idx_va != nullptr, "V tables need idx_va");
  if (d->bias_pack) CB_REQUIRE(d->idx_a != nullptr, "bias needs idx_a");

  AttnFwdParams p{};
  p.B = d->B; p.H = d->H; p.N = d->N; p.Npad = Npad;
  p.scale = d->scale;
  p.ctx_k = ctx_k; p.ctx_v = ctx_v;
  p.shared_tables = d->tables_per_head ? 0 : 1;
  p.idx_a = d->idx_a; p.idx_b = d->idx_b; p.idx_va = d->idx_va; p.idx_vb = d->idx_vb;
  p.ldi = ldi;
  p.bias = d->bias_pack;
  p.out = static_cast<__nv_bfloat16*>(d->out); p.ldo = d->ld_out;
  p.wide_out = aligned_for_256bit(d->out, d->ld_out) ? 1 : 0;
  p.lse = d->lse;
  p.dense = d->dense_bias; p.dense_sb = d->dense_stride_b; p.dense_sh = d->dense_stride_h; p.dense_si = d->dense_stride_i;
  p.causal = d->causal;
  CB_REQUIRE(!d->causal || (d->af_grid == 0 && d->gp_grid == 0), "the causal mask runs on the generic gather path (no af / gp hint)");
  p.block_len = d->block_len;
  CB_REQUIRE(d->block_len >= 0 && (d->block_len == 0 || (d->N % d->block_len == 0 && d->af_grid == 0 && d->gp_grid == 0)),
             "block_len must divide N; packed items run on the generic gather path");
  // structured AutoFormer gather: square grid + cls, both tables, clamp never binding
  if (d->af_grid > 0 && d->dense_bias == nullptr) {
    CB_REQUIRE(d->af_grid * d->af_grid + 1 == d->N && ctx_k && ctx_v && !d->bias_pack, "af mode needs N = g*g+1 and both table packs");
    CB_REQUIRE(2 * d->af_max_rel + 2 <= 32, "af tables must fit 32 packed rows");
    if (d->af_grid == 14 && d->af_max_rel >= d->af_grid - 1) {
      p.af_grid = d->af_grid;
      p.af_max_rel = d->af_max_rel;
      static const bool use_mma = []() { const char* e = getenv("CREAM_AF_MMA"); return e != nullptr && e[0] == '1'; }();
      p.af_mma = use_mma ? 1 : 0;     // CREAM_AF_MMA=1 selects it; default: the register-arithmetic structured path
    }
  }

  // structured iRPE product gather: 14 x 14 grid + cls, one contextual table on keys, nothing else
  if (d->gp_grid == 14 && d->gp_grid * d->gp_grid + 1 == d->N && ctx_k && !ctx_v && d->idx_b == nullptr && !d->bias_pack &&
      d->dense_bias == nullptr && p.af_grid == 0 && d->gp_w >= 1 && d->gp_skip_id >= 0 && d->gp_skip_id < kNB) {
    bool ok = true;
    for (int t = 0; t < 2 * d->gp_grid - 1; ++t)
      ok = ok && d->gp_lut_a[t] * d->gp_w + d->gp_lut_b[t] < kNB && d->gp_lut_b[t] < d->gp_w;
    if (ok) {
      p.gp_grid = d->gp_grid; p.gp_w = d->gp_w; p.gp_skip = d->gp_skip_id;
      std::memcpy(p.lut_a, d->gp_lut_a, 32);
      std::memcpy(p.lut_b, d->gp_lut_b, 32);
    }
  }

  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d->H * kD), static_cast<uint64_t>(d->N),
                            static_cast<uint64_t>(d->B)};
  const uint64_t strides[3] = {1, static_cast<uint64_t>(d->ld_qkv),
                               static_cast<uint64_t>(d->N) * d->ld_qkv};
  const uint32_t box_q[3] = {64, 128, 1};
  const uint32_t box_kv[3] = {64, static_cast<uint32_t>(Npad), 1};
  const CUtensorMap* mq = get_tensor_map(d->qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides,
                                         box_q, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap* mkv = get_tensor_map(d->qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides,
                                          box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
  const int ntab = d->tables_per_head ? d->H : 1;
  const uint64_t tdims[3] = {64, 64, static_cast<uint64_t>(ntab)};
  const uint64_t tstrides[3] = {1, 64, 64 * 64};
  const uint32_t tbox[3] = {64, 64, 1};
  const CUtensorMap* mtk = ctx_k ? get_tensor_map(d->tk_pack, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, tdims,
                                                  tstrides, tbox, CU_TENSOR_MAP_SWIZZLE_128B)
                                 : mq;
  const CUtensorMap* mtv = ctx_v ? get_tensor_map(d->tv_pack, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, tdims,
                                                  tstrides, tbox, CU_TENSOR_MAP_SWIZZLE_128B)
                                 : mq;
  if (!mq || !mkv || !mtk || !mtv) return CREAM_ERR_CUDA;

  const size_t smem_bytes = 16384 + std::max<size_t>(Npad * 128, 11 * 1024) +
                            static_cast<size_t>(Npad) * 128 + 2 * 8192 + 128 * kRStride * 2 + 64 * 4 + 64 + 192;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    113 * 1024));
    attr_set = true;
  }
  CB_REQUIRE(smem_bytes <= 113 * 1024, "shared memory budget");
  dim3 grid(ceil_div(d->N, 128), d->H, d->B);
  CB_CUDA_OK(launch_chain(attn_fwd_kernel, grid, dim3(kThreads), smem_bytes, stream, 1, *mq, *mkv, *mtk, *mtv, p));
  return check_last("attn_fwd_kernel");
}
