// Fused attention backward with relative-position terms, sm_100a (tcgen05 + TMEM + TMA).
//
// Adjoint of attention_fwd.cu (autograd of AutoFormer/model/module/multihead_super.py:135-154
// and iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:73-92).  Two kernels:
//
//  bwd_rows (CTA = query tile x head x batch, one thread per query row i):
//      recompute T = scale*(Q K^T + R-gather) and P = exp(T - lse);
//      dP = dO V^T + gather(dO TV^T);  dT = P o (dP - delta_i),  delta_i = dO_i . O_i;
//      thread-local bucket sums  PB[i,b] = sum_j P[i,j][idx_v=b],  dR[i,b] = sum_j dT[i,j][idx_k=b];
//      dQ = scale * [dT | dR] . [K ; TK]   (tcgen05, A operand from TMEM);
//      [P | PB] and [dT | dR] (bf16) go to a workspace for the column-form products.
//  bwd_cols (persistent over head x batch):
//      [dV ; dTV] = [P | PB]^T dO        [dK ; dTK] = scale * [dT | dR]^T Q
//      (MN-major A operands streamed from the workspace by TMA; the two products run back to
//      back through one stage ring so each epilogue overlaps the other product's MMAs; table
//      gradients leave through 256-byte bulk reduce-adds, not per-element atomics).
//
// The workspace round trip (2 x B*H*N*(Npad+64) bf16) costs about 3x the algorithmic bytes of the
// fused ideal; a single fused kernel needs the dK/dV accumulators of all keys resident (CTA pairs).
#include <cstdlib>
#include <cstring>

#include "attention_common.cuh"

namespace cb {
namespace {

constexpr int kD = 64;
constexpr int kNB = 64;
constexpr int kRowsThreads = 288;  // warp 0: TMA + MMA issue; warps 1..8: two threads per query row
constexpr int kRowThreads = 256;
constexpr int kStride = 66;  // floats per row of the staged R / dPB / dR tiles (66: the structured
                             // gathers of 32 consecutive rows spread over the banks, <= 2-way)

struct BwdRowsParams {
  int B, H, N, Npad, ldw;
  int causal;                          // generic path: key j > query i masked
  int block_len;                       // generic path: block-diagonal attention over items of block_len tokens (0 = off)
  float scale;
  int ctx_k, ctx_v, shared_tables;
  int af_grid, af_max_rel;
  int af_mma;                          // AutoFormer structure through the tensor cores (see bwd_row_plain)
  int gp_grid, gp_w, gp_skip;          // iRPE grid-product structured mode (0 = off), see attention_fwd.cu
  uint8_t lut_a[32], lut_b[32];
  const uint8_t* idx_a; const uint8_t* idx_b; const uint8_t* idx_va; const uint8_t* idx_vb;
  int ldi;
  const float* bias;
  const __nv_bfloat16* out; int64_t ldo;     // forward output (for delta)
  const __nv_bfloat16* dout; int64_t lddo;
  const float* lse;
  __nv_bfloat16* dqkv; int64_t lddqkv;
  __nv_bfloat16* ws_p; __nv_bfloat16* ws_dt;  // (B*H*N, ldw)
  float* dbias;
  int wide_out;                        // dqkv rows are 32-byte aligned: 256-bit stores
  int stage;                           // structured paths: [P | PB] / [dT | dR] leave through per-warp staging tiles + TMA stores
  const float* dense; int64_t dense_sb, dense_sh, dense_si;   // dense additive logit term (generic path)
  float* ddense;                                              // (B,H,N,N) fp32: dS, or NULL
  long long* trace;   // CREAM_TRACE builds only
};

#ifdef CREAM_TRACE
// two traced CTAs: the first of the grid (first wave: every SM loads at once) and one from the middle of the grid
#define ROWS_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (blockIdx.z == 0 || blockIdx.z == gridDim.z / 2)) \
    p.trace[(blockIdx.z == 0 ? 0 : 64) + (slot)] = clock64(); } while (0)
#else
#define ROWS_TRACE(slot) do {} while (0)
#endif

__device__ __forceinline__ void rows_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kRowThreads) : "memory"); }

#ifndef CREAM_ABL
#define CREAM_ABL 0      // timing ablations (debug builds only; results are wrong when non-zero)
#endif
constexpr int kAfSplitCols = 112;                 // keys owned by the first thread of a row
constexpr int kAfSplitRows = 8;                   // = kAfSplitCols / 14 grid rows
constexpr int kDtHiCol = 464;                     // TMEM columns [464, 512): packed dT of the second half

struct RowCtx {
  const float* drow;    // this row of the dense additive logit term, or NULL
  float* ddrow;         // this row of its gradient, or NULL
  uint32_t trow;        // TMEM address of this thread's lane
  uint32_t s_r, s_dpb;  // shared addresses of this row's staged R (scaled) and dPB, fp32[kStride]
  uint32_t s_pb, s_dr;  // shared addresses of this row's PB (64 floats, rotated by sw) and dR (kStride)
  uint32_t s_bias, s_lut;
  int row, row_c, sw, half;
  float delta, lsel;
  int64_t wrow;
  uint64_t* bar_cols;   // rows -> MMA: this thread's packed dT is in TMEM (the dT . K part of dQ may start)
  // workspace staging (structured paths): this WARP's two tiles of 32 rows x 64 bytes (P at +0, dT at +2048),
  // SWIZZLE_64B layout, written row-per-lane and stored by TMA (rows >= N are clipped by the tensor map)
  uint32_t s_stage;
  int stage, lane, row0, bh;
  const CUtensorMap* map_sp; const CUtensorMap* map_sd;
};

// The workspace rows are 544 bytes apart, so a thread that owns a row can only ever write 32 contiguous bytes of it per
// store: 32 cache lines per warp instruction.  Measured (timing ablation, profiles/r02_README.md): those stores were
// 13 % of the backward.  Instead two consecutive 16-key chunks (64 bytes per row) of P and of dT are written to the
// warp's staging tiles - 16-byte chunk c of row r at chunk c ^ ((r >> 1) & 3), the layout CU_TENSOR_MAP_SWIZZLE_64B
// reads back, conflict-free for a row-per-lane writer - and ONE elected lane hands each tile to the TMA engine.
__device__ __forceinline__ void stage_wait_free(const RowCtx& x) {       // the previous TMA stores have read the tiles
  if (x.lane == 0) bulk_wait_read<0>();
  __syncwarp();
}
__device__ __forceinline__ void stage_put32(const RowCtx& x, int half_of_row, const uint32_t (&pk)[8], const uint32_t (&dk)[8]) {
  const uint32_t rowb = x.s_stage + x.lane * 64;
  const uint32_t sw = (static_cast<uint32_t>(x.lane) >> 1) & 3u;
  const uint32_t c0 = ((2u * half_of_row) ^ sw) << 4, c1 = ((2u * half_of_row + 1u) ^ sw) << 4;
  sts_u32x4s(rowb + c0, pk[0], pk[1], pk[2], pk[3]);
  sts_u32x4s(rowb + c1, pk[4], pk[5], pk[6], pk[7]);
  sts_u32x4s(rowb + 2048 + c0, dk[0], dk[1], dk[2], dk[3]);
  sts_u32x4s(rowb + 2048 + c1, dk[4], dk[5], dk[6], dk[7]);
}
__device__ __forceinline__ void stage_store(const RowCtx& x, int col) {   // both tiles -> workspace columns [col, col + 32)
  fence_proxy_async_smem();
  __syncwarp();
  if (x.lane == 0) {
    tma_store_3d_s(x.map_sp, x.s_stage, col, x.row0, x.bh);
    tma_store_3d_s(x.map_sd, x.s_stage + 2048, col, x.row0, x.bh);
    bulk_commit();
  }
}
// One 16-key chunk of a structured path: pairs of chunks (cc even, cc + 1) go through the staging tiles; the odd
// seventh chunk of the first half of a row is written directly.
// (cc is a compile-time constant at every call site: the column loops are fully unrolled.)
__device__ __forceinline__ void emit_chunk(const BwdRowsParams& p, const RowCtx& x, int cc, int hi, bool live, int64_t w0,
                                           const uint32_t (&pk)[8], const uint32_t (&dk)[8]) {
  if (CREAM_ABL & 1) return;
  if (x.stage && cc < 6) {
    if ((cc & 1) == 0 && cc >= 2) stage_wait_free(x);
    stage_put32(x, cc & 1, pk, dk);
    if (cc & 1) stage_store(x, kAfSplitCols * hi + (cc - 1) * 16);
  } else if (live) {
    stg_256(p.ws_p + w0 + cc * 16, pk);     // 16 keys = 32 bytes of this row, one 256-bit store each
    stg_256(p.ws_dt + w0 + cc * 16, dk);
  }
}

// Called by every row thread once its share of the packed dT has been written to TMEM: the tensor core starts
// dQ = dT . K while the bucket sums are still being folded and packed (dR . TK is added after that).
__device__ __forceinline__ void columns_done(const RowCtx& x) {
  tmem_st_wait();
  tc_fence_before();
  mbar_arrive(x.bar_cols);
}

// dT / P of one query row, generic gather tables (see attention_fwd.cu for the forward twin).
__device__ __forceinline__ void bwd_row_generic(const BwdRowsParams& p, const RowCtx& x) {
  const int Npad = p.Npad;
  if (x.half != 0) { columns_done(x); rows_barrier(); rows_barrier(); return; }   // keep the barrier schedule of the pair
  for (int k = 0; k < kNB; ++k) { sts_f32(x.s_pb + 4 * ((k + x.sw) & 63), 0.f); sts_f32(x.s_dr + 4 * k, 0.f); }
  const uint8_t* ia = p.idx_a ? p.idx_a + static_cast<int64_t>(x.row_c) * p.ldi : nullptr;
  const uint8_t* ib = p.idx_b ? p.idx_b + static_cast<int64_t>(x.row_c) * p.ldi : nullptr;
  const uint8_t* iva = p.idx_va ? p.idx_va + static_cast<int64_t>(x.row_c) * p.ldi : nullptr;
  const uint8_t* ivb = p.idx_vb ? p.idx_vb + static_cast<int64_t>(x.row_c) * p.ldi : nullptr;
  const bool use_bias = p.bias != nullptr;
  const bool want_dr = p.ctx_k || use_bias;
  int k_lo = 0, k_hi = x.row < p.N ? p.N : 0;      // visible keys of this row (see softmax_generic); none for padding rows
  if (p.block_len > 0 && x.row < p.N) { k_lo = (x.row / p.block_len) * p.block_len; k_hi = k_lo + p.block_len; }
  if (p.causal) k_hi = min(k_hi, x.row + 1);
  const int nchunks = Npad / 16;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t rt[16], rp[16];
    tmem_ld16(x.trow + c * 16, rt);
    tmem_ld16(x.trow + 256 + c * 16, rp);
    uint4 va = make_uint4(0, 0, 0, 0), vb = va, vva = va, vvb = va;
    if (ia) va = __ldg(reinterpret_cast<const uint4*>(ia + c * 16));
    if (ib) vb = __ldg(reinterpret_cast<const uint4*>(ib + c * 16));
    if (iva) vva = __ldg(reinterpret_cast<const uint4*>(iva + c * 16));
    if (ivb) vvb = __ldg(reinterpret_cast<const uint4*>(ivb + c * 16));
    tmem_ld_wait();
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    const uint32_t wva[4] = {vva.x, vva.y, vva.z, vva.w}, wvb[4] = {vvb.x, vvb.y, vvb.z, vvb.w};
    float pv[16], dt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t a_id = byte_of(wa, k), b_id = byte_of(wb, k);
      const uint32_t va_id = byte_of(wva, k), vb_id = byte_of(wvb, k);
      float t = p.scale * __uint_as_float(rt[k]);
      if (p.ctx_k) {
        if (ia) t += lds_f32(x.s_r + 4 * a_id);
        if (ib) t += lds_f32(x.s_r + 4 * b_id);
      }
      if (use_bias) t += lds_f32(x.s_bias + 4 * a_id);
      if (x.drow != nullptr && c * 16 + k < p.N) t += __ldg(x.drow + c * 16 + k);
      float pr = fast_exp2(fmaf(t, kLog2e, -x.lsel));
      if (c * 16 + k < k_lo || c * 16 + k >= k_hi) pr = 0.f;
      float dp = __uint_as_float(rp[k]);
      if (p.ctx_v) {
        if (iva) dp += lds_f32(x.s_dpb + 4 * va_id);
        if (ivb) dp += lds_f32(x.s_dpb + 4 * vb_id);
      }
      const float d = pr * (dp - x.delta);
      pv[k] = pr;
      dt[k] = d;
      if (x.ddrow != nullptr && c * 16 + k < p.N && x.row < p.N) x.ddrow[c * 16 + k] = d;
      if (p.ctx_v) {
        if (iva) { const uint32_t a = x.s_pb + 4 * ((va_id + x.sw) & 63); sts_f32(a, lds_f32(a) + pr); }
        if (ivb) { const uint32_t a = x.s_pb + 4 * ((vb_id + x.sw) & 63); sts_f32(a, lds_f32(a) + pr); }
      }
      if (want_dr) {
        if (ia) { const uint32_t a = x.s_dr + 4 * a_id; sts_f32(a, lds_f32(a) + d); }
        if (ib) { const uint32_t a = x.s_dr + 4 * b_id; sts_f32(a, lds_f32(a) + d); }
      }
    }
    uint32_t pk[8], dk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
      dk[k] = pack_bf16x2(dt[2 * k], dt[2 * k + 1]);
    }
    tmem_st8(x.trow + c * 8, dk);   // dT (bf16x2) in place over T, A operand of the dQ MMA
    if (x.row < p.N) {
      stg_256(p.ws_p + x.wrow + c * 16, pk);     // 16 keys = 32 bytes of this row, one 256-bit store each
      stg_256(p.ws_dt + x.wrow + c * 16, dk);
    }
  }
  columns_done(x);
  rows_barrier();
  rows_barrier();
}

// AutoFormer-structured twin (see softmax_af in attention_fwd.cu): the gather operand vectors and
// the bucket-sum vectors of a row are registers; afterwards the bucket sums are scattered ONCE
// into the shared rows the common tail reads.
//
// The two threads of a row split the keys 112 / 96: 112 = 8 grid rows, so key j0 + 112 has the
// same grid column as key j0 and a grid row 8 higher.  Both halves therefore run the SAME fully
// unrolled code on "local" grid rows 0..7 with their own operand registers (the kernel is
// instruction-fetch bound: one shared stream halves its footprint).  The only asymmetric element
// is local key 0: the cls key for half 0, grid position (7, 13) for half 1.

template <int G>
__device__ __forceinline__ void bwd_row_af(const BwdRowsParams& p, const RowCtx& x) {
  static_assert(G == 14 && kAfSplitCols == kAfSplitRows * G, "split is tied to the 14 x 14 grid");
  constexpr int N = G * G + 1;
  constexpr int LR = kAfSplitRows;                // local grid rows per half
  constexpr int NCC = kAfSplitCols / 16;          // 7 chunks for half 0, 6 for half 1
  const int M1 = p.af_max_rel + 1;
  const bool patch = x.row >= 1 && x.row < N;
  const bool live = x.row < N;
  const int hi = x.half;                           // 0 / 1
  const int qi = patch ? x.row - 1 : 0;
  const int ri = qi / G, ci = qi - ri * G;
  const int vb = M1 - ri + LR * hi;               // bucket of local grid row 0
  // Fold the per-row constants into the gather operands so one element costs
  //   e = T * (scale log2e) + rv[rj] + rh[cj] ; p = 2^e ; dT = p * (dP + gv[rj] + gh[cj]).
  // A padding row gets lse = +huge, i.e. p = 0 everywhere, without a per-element select; the local
  // grid rows of half 1 that do not exist (>= 6) get the same treatment.
  const float sl = p.scale * kLog2e;
  const float lsel = live ? x.lsel : 1e30f;
  const float r0v = lds_f32(x.s_r), r0h = lds_f32(x.s_r + 4 * 32);
  const float g0v = lds_f32(x.s_dpb), g0h = lds_f32(x.s_dpb + 4 * 32);
  float rv[LR], gv[LR], rh[G], gh[G];
#pragma unroll
  for (int t = 0; t < LR; ++t) {
    const bool exists = hi == 0 || t < G - LR;
    const float r = patch ? lds_f32(x.s_r + 4 * (exists ? vb + t : 0)) : r0v;
    const float g = patch ? lds_f32(x.s_dpb + 4 * (exists ? vb + t : 0)) : g0v;
    rv[t] = exists ? fmaf(r, kLog2e, -lsel) : -1e30f;
    gv[t] = exists ? g - x.delta : 0.f;
  }
#pragma unroll
  for (int t = 0; t < G; ++t) {
    rh[t] = (patch ? lds_f32(x.s_r + 4 * (32 + M1 - ci + t)) : r0h) * kLog2e;
    gh[t] = patch ? lds_f32(x.s_dpb + 4 * (32 + M1 - ci + t)) : g0h;
  }
  // local key 0: the cls key (bucket 0 of both tables) or grid position (7, 13)
  float e_first, g_first;
  if (hi == 0) {
    e_first = fmaf(r0v + r0h, kLog2e, -lsel);
    g_first = g0v + g0h - x.delta;
  } else {
    const float r7 = patch ? lds_f32(x.s_r + 4 * (M1 - ri + LR - 1)) : r0v;
    const float g7 = patch ? lds_f32(x.s_dpb + 4 * (M1 - ri + LR - 1)) : g0v;
    e_first = fmaf(r7, kLog2e, -lsel) + rh[G - 1];
    g_first = g7 - x.delta + gh[G - 1];
  }
  float prow[LR], drow[LR], pcol[G], dcol[G], pf = 0.f, df = 0.f;
#pragma unroll
  for (int t = 0; t < LR; ++t) { prow[t] = 0.f; drow[t] = 0.f; }
#pragma unroll
  for (int t = 0; t < G; ++t) { pcol[t] = 0.f; dcol[t] = 0.f; }

  const uint32_t t_in = x.trow + kAfSplitCols * hi;               // this half's T columns
  const uint32_t t_out = hi ? x.trow + kDtHiCol : x.trow;        // packed dT: in place, or the spare columns
  const int64_t w0 = x.wrow + kAfSplitCols * hi;
  // accumulator reads run one chunk ahead of the arithmetic (tcgen05.wait::ld covers every load
  // issued before it, so the next chunk is requested right after the wait)
  uint32_t rtb[2][16], rpb[2][16];
  tmem_ld16(t_in, rtb[0]);
  tmem_ld16(t_in + 256, rpb[0]);
#pragma unroll
  for (int cc = 0; cc < NCC; ++cc) {
    if (cc == NCC - 1 && hi) break;                                // half 1 owns 6 chunks (keys 112..207)
    uint32_t (&rt)[16] = rtb[cc & 1];
    uint32_t (&rp)[16] = rpb[cc & 1];
    tmem_ld_wait();
    if (cc + 1 < NCC && !(cc + 1 == NCC - 1 && hi)) {
      tmem_ld16(t_in + (cc + 1) * 16, rtb[(cc + 1) & 1]);
      tmem_ld16(t_in + 256 + (cc + 1) * 16, rpb[(cc + 1) & 1]);
    }
    float pv[16], dt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j0 = cc * 16 + k;                                  // local key
      float pr, d;
      if (j0 == 0) {
        pr = fast_exp2(fmaf(sl, __uint_as_float(rt[k]), e_first));
        d = pr * (__uint_as_float(rp[k]) + g_first);
        pf = pr; df = d;
      } else {
        const int rj = (j0 - 1) / G, cj = (j0 - 1) % G;
        pr = (CREAM_ABL & 8) ? fmaf(sl, __uint_as_float(rt[k]), rv[rj]) + rh[cj]
                             : fast_exp2(fmaf(sl, __uint_as_float(rt[k]), rv[rj]) + rh[cj]);
        d = pr * (__uint_as_float(rp[k]) + gv[rj] + gh[cj]);
        if (!(CREAM_ABL & 4)) {
        prow[rj] += pr; pcol[cj] += pr;
        drow[rj] += d;  dcol[cj] += d;
        }
      }
      pv[k] = pr;
      dt[k] = d;
    }
    uint32_t pk[8], dk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
      dk[k] = pack_bf16x2(dt[2 * k], dt[2 * k + 1]);
    }
    if (!(CREAM_ABL & 16)) tmem_st8(t_out + cc * 8, dk);     // half 0: over T columns it has already read; half 1: spare columns
    emit_chunk(p, x, cc, hi, live, w0, pk, dk);
  }
  columns_done(x);
  // this thread's share of the row totals (every key hits exactly one vertical bucket)
  float psum = pf, dsum = df;
#pragma unroll
  for (int t = 0; t < LR; ++t) { psum += prow[t]; dsum += drow[t]; }

  // scatter the register bucket sums into the shared rows read by the common tail: the thread
  // owning the first half initialises the row, its partner adds its partial sums after the barrier.
  auto put = [&](uint32_t a, float v) { if (hi) v += lds_f32(a); sts_f32(a, v); };
  if (hi == 0) {      // zero both bucket rows: PB (256 B, 16-byte aligned; the rotation only renames the slots) and dR (8-byte aligned)
#pragma unroll
    for (int k = 0; k < kNB / 4; ++k) sts_f32x4(x.s_pb + 16 * k, make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
    for (int k = 0; k < kNB / 2; ++k) sts_f32x2(x.s_dr + 8 * k, 0.f, 0.f);
  } else {
    rows_barrier();
  }
  if (patch) {
#pragma unroll
    for (int t = 0; t < LR; ++t) {
      if (hi && t >= G - LR) continue;
      put(x.s_pb + 4 * (((vb + t) + x.sw) & 63), prow[t]);
      put(x.s_dr + 4 * (vb + t), drow[t]);
    }
#pragma unroll
    for (int t = 0; t < G; ++t) {
      put(x.s_pb + 4 * (((32 + M1 - ci + t) + x.sw) & 63), pcol[t]);
      put(x.s_dr + 4 * (32 + M1 - ci + t), dcol[t]);
    }
    if (hi == 0) {           // cls key: bucket 0 of the vertical and of the horizontal table
      sts_f32(x.s_pb + 4 * ((0 + x.sw) & 63), pf);
      sts_f32(x.s_pb + 4 * ((32 + x.sw) & 63), pf);
      sts_f32(x.s_dr, df);
      sts_f32(x.s_dr + 4 * 32, df);
    } else {                 // grid position (7, 13)
      put(x.s_pb + 4 * (((M1 - ri + LR - 1) + x.sw) & 63), pf);
      put(x.s_pb + 4 * (((32 + M1 - ci + G - 1) + x.sw) & 63), pf);
      put(x.s_dr + 4 * (M1 - ri + LR - 1), df);
      put(x.s_dr + 4 * (32 + M1 - ci + G - 1), df);
    }
  } else {
    // cls query row: every key gathers bucket 0 of both tables.  The bucket sums of P are genuine (= 1 each);
    // those of dT are sum_j dT[0, j] = 0 exactly - only the rounding residual of delta would land there.
    put(x.s_pb + 4 * ((0 + x.sw) & 63), psum);
    put(x.s_pb + 4 * ((32 + x.sw) & 63), psum);
    (void)dsum;
  }
  if (hi == 0) rows_barrier();
  rows_barrier();
}

// iRPE product method on a 14 x 14 grid + cls, contextual table on keys only (see softmax_gridprod in
// attention_fwd.cu): id(i, j) = A[rj - ri] * W + B[cj - ci], the skip bucket when i or j is cls.
// Gather: the staged R row is read at  base[rj] + off[cj]  (one integer add per element, registers).
// Bucket sums dR[i, id] = sum_j dT[i, j] [id(i,j) = id]: A and B are monotone in the offsets, so the
// keys of one bucket form a RECTANGLE of the grid; column sums are kept in 14 registers over a run of
// grid rows with the same A and folded into the row's shared-memory buckets once per run - a plain
// store per rectangle, no per-element read-modify-write.  The two threads of a row (keys 0..111 /
// 112..207, see bwd_row_af) write to separate bucket rows (s_dr / the rotated s_pb), summed by the tail.
template <int G>
__device__ __forceinline__ void bwd_row_gridprod(const BwdRowsParams& p, const RowCtx& x) {
  static_assert(G == 14 && kAfSplitCols == kAfSplitRows * G, "split is tied to the 14 x 14 grid");
  constexpr int N = G * G + 1;
  constexpr int LR = kAfSplitRows;
  constexpr int NCC = kAfSplitCols / 16;
  const bool patch = x.row >= 1 && x.row < N;
  const bool live = x.row < N;
  const int hi = x.half;
  const int qi = patch ? x.row - 1 : 0;
  const int ri = qi / G, ci = qi - ri * G;
  const float sl = p.scale * kLog2e;
  const float lsel = live ? x.lsel : 1e30f;
  const uint32_t skip4 = 4u * p.gp_skip;
  auto lut = [&](int which, int delta) -> uint32_t {   // component of the bucket id for an offset
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(x.s_lut + 32 * which + (delta + G - 1)));
    return v;
  };
  // byte offsets into a 64-float bucket row: row part (A * W) per local grid row, column part per column
  uint32_t offa[LR], offb[G];
  const int rows_here = hi ? G - LR : LR;             // grid rows of this half: 8 / 6
#pragma unroll
  for (int t = 0; t < LR; ++t) {
    const int rj = min(t, rows_here - 1) + LR * hi;   // rows that do not exist repeat the last one (never flushed)
    offa[t] = patch ? 4u * lut(0, rj - ri) * p.gp_w : skip4;
  }
#pragma unroll
  for (int t = 0; t < G; ++t) offb[t] = patch ? 4u * lut(1, t - ci) : 0u;
  // local key 0: the cls key (skip bucket) for half 0, grid position (7, 13) for half 1
  const uint32_t first4 = (hi && patch) ? 4u * (lut(0, LR - 1 - ri) * p.gp_w + lut(1, G - 1 - ci)) : skip4;
  auto bucket_addr = [&](uint32_t off4) -> uint32_t {  // this half's private bucket row
    return hi ? x.s_pb + 4u * (((off4 >> 2) + x.sw) & 63u) : x.s_dr + off4;
  };
  if (hi) {
#pragma unroll
    for (int k = 0; k < kNB / 4; ++k) sts_f32x4(x.s_pb + 16 * k, make_float4(0.f, 0.f, 0.f, 0.f));
  } else {
#pragma unroll
    for (int k = 0; k < kNB / 2; ++k) sts_f32x2(x.s_dr + 8 * k, 0.f, 0.f);
  }

  float colacc[G], df = 0.f;
#pragma unroll
  for (int t = 0; t < G; ++t) colacc[t] = 0.f;
  auto flush = [&](uint32_t oa) {
    float acc = 0.f;
#pragma unroll
    for (int cj = 0; cj < G; ++cj) {
      acc += colacc[cj];
      colacc[cj] = 0.f;
      const bool end = (cj == G - 1) || (offb[(cj + 1) % G] != offb[cj]);
      if (end) { sts_f32(bucket_addr(oa + offb[cj]), acc); acc = 0.f; }
    }
  };

  const uint32_t t_in = x.trow + kAfSplitCols * hi;
  const uint32_t t_out = hi ? x.trow + kDtHiCol : x.trow;
  const int64_t w0 = x.wrow + kAfSplitCols * hi;
  uint32_t rtb[2][16], rpb[2][16];
  tmem_ld16(t_in, rtb[0]);
  tmem_ld16(t_in + 256, rpb[0]);
#pragma unroll
  for (int cc = 0; cc < NCC; ++cc) {
    if (cc == NCC - 1 && hi) break;
    uint32_t (&rt)[16] = rtb[cc & 1];
    uint32_t (&rp)[16] = rpb[cc & 1];
    tmem_ld_wait();
    if (cc + 1 < NCC && !(cc + 1 == NCC - 1 && hi)) {
      tmem_ld16(t_in + (cc + 1) * 16, rtb[(cc + 1) & 1]);
      tmem_ld16(t_in + 256 + (cc + 1) * 16, rpb[(cc + 1) & 1]);
    }
    float pv[16], dt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j0 = cc * 16 + k;
      float pr, d;
      if (j0 == 0) {
        const float e = fmaf(lds_f32(x.s_r + first4), kLog2e, -lsel);
        pr = fast_exp2(fmaf(sl, __uint_as_float(rt[k]), e));
        d = pr * (__uint_as_float(rp[k]) - x.delta);
        df = d;
      } else {
        const int rj = (j0 - 1) / G, cj = (j0 - 1) % G;
        const bool exists = hi == 0 || rj < G - LR;
        const float e = exists ? fmaf(lds_f32(x.s_r + offa[rj] + offb[cj]), kLog2e, -lsel) : -1e30f;
        pr = fast_exp2(fmaf(sl, __uint_as_float(rt[k]), e));
        d = pr * (__uint_as_float(rp[k]) - x.delta);
        colacc[cj] += d;
        if (cj == G - 1 && rj + 1 < LR) {               // end of a grid row: fold the run if A changes
          if (offa[rj + 1] != offa[rj]) flush(offa[rj]);
        }
      }
      pv[k] = pr;
      dt[k] = d;
    }
    uint32_t pk[8], dk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
      dk[k] = pack_bf16x2(dt[2 * k], dt[2 * k + 1]);
    }
    tmem_st8(t_out + cc * 8, dk);
    emit_chunk(p, x, cc, hi, live, w0, pk, dk);
  }
  columns_done(x);
  // The cls query row gathers ONE bucket for every key, i.e. adds a constant to its logits: its exact
  // contribution to dR is sum_j dT[0, j] = 0.  What a flash-style backward would add instead is the
  // rounding residual of delta = dO . O (bf16 O), scaled by the cls row's large gradient in DeiT - so
  // non-patch rows are left out (their bucket row stays zero).
  if (patch) {
    flush(offa[LR - 1]);                                // the last run (rows past the half's end repeat its offset)
    const uint32_t a = bucket_addr(first4);             // local key 0 may share a bucket with a rectangle
    sts_f32(a, lds_f32(a) + df);
  }
  rows_barrier();
}

// AutoFormer relative position through the tensor cores (the backward twin of softmax_plain in
// attention_fwd.cu).  The gather terms have already been ADDED into the accumulators by MMAs against the
// 0/1 feature matrix Ind (T = Q K^T + A_R . Ind,  dP = dO V^T + A_dPB . Ind), so a row is a plain
//     p = 2^(scale log2e T - lse log2e),   dT = p (dP - delta)
// over its keys: no gather adds, no compile-time (rj, cj), no bucket-sum accumulators - those come back
// from the tensor cores as PBabs = P . Ind^T and dRabs = dT . Ind^T.  Two threads per row (keys 0..111 /
// 112..207); each packs P and dT (bf16) in place over accumulator columns it has already read:
//     dT: half 0 -> [0, 56), half 1 -> [464, 512);    P: half 0 -> [256, 312), half 1 -> [112, 160).
constexpr int kPHiCol = 112;                      // TMEM columns of the second half's packed P
constexpr int kAbsPB = 160, kAbsDR = 56;          // TMEM columns of PBabs / dRabs (32 fp32 columns each)
constexpr int kDrPackCol = 320;                   // packed dR (A operand of the table part of dQ)
constexpr int kAvecR = 208, kAvecDPB = 240, kAvecDPBlo = 464;   // A vectors: R hi | lo at [208, 240), dPB hi [240, 256), lo [464, 480)

__device__ __forceinline__ void bwd_row_plain(const BwdRowsParams& p, const RowCtx& x) {
  constexpr int NCC = kAfSplitCols / 16;
  const bool live = x.row < p.N;
  const int hi = x.half;
  const float sl = p.scale * kLog2e;
  const float lsel = live ? x.lsel : 1e30f;
  const uint32_t t_in = x.trow + kAfSplitCols * hi;
  const uint32_t dt_out = hi ? x.trow + kDtHiCol : x.trow;
  const uint32_t p_out = hi ? x.trow + kPHiCol : x.trow + 256;
  const int64_t w0 = x.wrow + kAfSplitCols * hi;
  const int n_left = p.N - kAfSplitCols * hi;     // valid keys of this half: 112 / 85
  uint32_t rtb[2][16], rpb[2][16];
  tmem_ld16(t_in, rtb[0]);
  tmem_ld16(t_in + 256, rpb[0]);
#pragma unroll
  for (int cc = 0; cc < NCC; ++cc) {
    if (cc == NCC - 1 && hi) break;               // the second half owns 6 chunks (keys 112..207)
    uint32_t (&rt)[16] = rtb[cc & 1];
    uint32_t (&rp)[16] = rpb[cc & 1];
    tmem_ld_wait();
    if (cc + 1 < NCC && !(cc + 1 == NCC - 1 && hi)) {
      tmem_ld16(t_in + (cc + 1) * 16, rtb[(cc + 1) & 1]);
      tmem_ld16(t_in + 256 + (cc + 1) * 16, rpb[(cc + 1) & 1]);
    }
    float pv[16], dt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float pr = fast_exp2(fmaf(sl, __uint_as_float(rt[k]), -lsel));
      if (cc * 16 + k >= n_left) pr = 0.f;        // key padding (only the last chunk of the second half)
      pv[k] = pr;
      dt[k] = pr * (__uint_as_float(rp[k]) - x.delta);
    }
    uint32_t pk[8], dk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pk[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
      dk[k] = pack_bf16x2(dt[2 * k], dt[2 * k + 1]);
    }
    tmem_st8(dt_out + cc * 8, dk);
    tmem_st8(p_out + cc * 8, pk);
    if (live) {
      stg_256(p.ws_p + w0 + cc * 16, pk);     // 16 keys = 32 bytes of this row, one 256-bit store each
      stg_256(p.ws_dt + w0 + cc * 16, dk);
    }
  }
}

__global__ void __launch_bounds__(kRowsThreads, 1)
attn_bwd_rows_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_o,
                     const __grid_constant__ CUtensorMap map_tk, const __grid_constant__ CUtensorMap map_tv,
                     const __grid_constant__ CUtensorMap map_sp, const __grid_constant__ CUtensorMap map_sd,
                     const BwdRowsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  if (threadIdx.x == 0) ROWS_TRACE(40);
  pdl_trigger();
  const int kv_bytes = p.Npad * 128;
  const int v_slot = max(kv_bytes, 26 * 1024);
  uint8_t* sQ = smem;                       // 16 KB
  uint8_t* sdO = sQ + 16384;                // 16 KB
  uint8_t* sO = sdO + 16384;                // 16 KB: forward output rows, only for delta_i = dO_i . O_i
  uint8_t* sK = sO + 16384;                 // [K ; TK] contiguous rows
  uint8_t* sTK = sK + kv_bytes;
  uint8_t* sV = sTK + 8192;                 // [V ; TV]
  uint8_t* sTV = sV + v_slot;
  uint8_t* sR = sTV + 8192;                 // fp32 [128][kStride]
  uint8_t* sdPB = sR + 128 * kStride * 4;   // fp32 [128][kStride]
  uint8_t* sInd = sdPB + 128 * kStride * 4;  // 16 KB: 0/1 feature matrix (tensor-core structured mode; 1024-aligned at N = 197)
  uint8_t* sBias = sInd + 4 * kIndChunk;
  uint8_t* sDbias = sBias + 64 * 4;
  uint8_t* sLut = sDbias + 64 * 4;          // 64 bytes: grid-product row / column components
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLut + 64);
  uint64_t* bar_ld = bars + 0;
  uint64_t* bar_r = bars + 1;
  uint64_t* bar_rfree = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint64_t* bar_abs = bars + 6;              // MMA -> rows: PBabs / dRabs ready (tensor-core structured mode)
  uint64_t* bar_p2 = bars + 7;               // rows -> MMA: packed dR written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint64_t* bar_cols = bars + 9;             // rows -> MMA: packed dT written (register-arithmetic paths)
  // workspace staging tiles (p.stage): 4 KB per row warp; the first four warps reuse sO (dead once delta is known),
  // the other four a 16 KB region behind the barriers
  uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars + 16) + 1023) & ~static_cast<uintptr_t>(1023));
  // overlays: PB (128 x 64 fp32, each row rotated by 2*row) over sQ|sdO ; dR (128 x 65 fp32) over sV|sTV

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int Npad = p.Npad;
  const int tab = p.shared_tables ? 0 : head;
  const bool any_r = p.ctx_k || p.ctx_v;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    prefetch_tmap(&map_do);
    prefetch_tmap(&map_o);
    mbar_init(bar_ld, 1);
    mbar_init(bar_r, 1);
    mbar_init(bar_rfree, kRowThreads);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, kRowThreads);
    mbar_init(bar_o, 1);
    mbar_init(bar_abs, 1);
    mbar_init(bar_p2, kRowThreads);
    mbar_init(bar_cols, kRowThreads);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  pdl_wait();      // barrier init / TMEM allocation above overlap the previous kernel's tail; global memory from here on
  if (p.af_mma && threadIdx.x >= 32) {
    write_ind_matrix(smem_u32(sInd), threadIdx.x - 32, kRowThreads, 14, p.N);
    fence_proxy_async_smem();
  }
  if (threadIdx.x >= 32 && threadIdx.x < 96) {
    sts_f32(smem_u32(sBias) + 4 * (threadIdx.x - 32), p.bias ? p.bias[tab * 64 + threadIdx.x - 32] : 0.f);
    sts_f32(smem_u32(sDbias) + 4 * (threadIdx.x - 32), 0.f);
  }
  if (p.gp_grid != 0 && threadIdx.x >= 96 && threadIdx.x < 160)
    sLut[threadIdx.x - 96] = threadIdx.x < 128 ? p.lut_a[threadIdx.x - 96] : p.lut_b[threadIdx.x - 128];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int qcol = head * kD, kcol = (p.H + head) * kD, vcol = (2 * p.H + head) * kD;
      mbar_arrive_expect_tx(bar_ld, 3 * 16384 + 2 * kv_bytes + (p.ctx_k ? 8192 : 0) + (p.ctx_v ? 8192 : 0));
      tma_load_3d(sQ, &map_q, bar_ld, qcol, m0, b);
      tma_load_3d(sdO, &map_do, bar_ld, qcol, m0, b);
      tma_load_3d(sO, &map_o, bar_ld, qcol, m0, b);
      tma_load_3d(sK, &map_kv, bar_ld, kcol, 0, b);
      tma_load_3d(sV, &map_kv, bar_ld, vcol, 0, b);
      if (p.ctx_k) tma_load_3d(sTK, &map_tk, bar_ld, 0, 0, tab);
      if (p.ctx_v) tma_load_3d(sTV, &map_tv, bar_ld, 0, 0, tab);
      ROWS_TRACE(0);
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      ROWS_TRACE(1);
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t aTK = smem_u32(sTK), aTV = smem_u32(sTV);
      const uint32_t id64 = umma_idesc_bf16(128, kNB, 0, 0);
      const uint32_t idN = umma_idesc_bf16(128, Npad, 0, 0);
      if (p.ctx_k)
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + 0, umma_smem_desc_sw128(aQ + k * 32, 16, 1024),
                  umma_smem_desc_sw128(aTK + k * 32, 16, 1024), id64, k > 0);
      if (p.ctx_v)
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + 64, umma_smem_desc_sw128(adO + k * 32, 16, 1024),
                  umma_smem_desc_sw128(aTV + k * 32, 16, 1024), id64, k > 0);
      if (any_r) umma_commit(bar_r);
      for (int k = 0; k < 4; ++k)   // dP = dO V^T
        umma_ss(tmem + 256, umma_smem_desc_sw128(adO + k * 32, 16, 1024),
                umma_smem_desc_sw128(aV + k * 32, 16, 1024), idN, k > 0);
      if (any_r) {
        mbar_wait(bar_rfree, 0);
        tc_fence_after();
      }
      for (int k = 0; k < 4; ++k)   // T = Q K^T
        umma_ss(tmem + 0, umma_smem_desc_sw128(aQ + k * 32, 16, 1024),
                umma_smem_desc_sw128(aK + k * 32, 16, 1024), idN, k > 0);
      const uint32_t aInd = smem_u32(sInd);
      if (p.af_mma) {
        // T += A_R(hi, lo) . Ind ;  dP += A_dPB(hi, lo) . Ind     (A vectors in TMEM, two K = 16 steps each)
        const uint32_t id_g = umma_idesc_bf16(128, Npad, 0, 1);
        const uint32_t acol[4] = {kAvecR, kAvecR + 16, kAvecDPB, kAvecDPBlo};
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int k = 0; k < kFeat / 16; ++k)
            umma_ts(tmem + (v < 2 ? 0 : 256), tmem + acol[v] + 8 * k, umma_smem_desc_sw128(aInd + k * 2048, kIndChunk, 1024),
                    id_g, 1u);
      }
      umma_commit(bar_s);
      ROWS_TRACE(2);

      const uint32_t id_o = umma_idesc_bf16(128, kD, 0, 1);
      const int ksteps = (Npad + (p.ctx_k ? kNB : 0)) / 16;
      const int nk = Npad / 16;
      auto dq_step = [&](int k) {   // dQ (+)= [dT | dR] [K ; TK], one K = 16 step
        // structured paths: the packed dT of keys >= 112 sits in the spare columns (see bwd_row_af / bwd_row_gridprod)
        const bool hi_part = (p.af_grid != 0 || p.gp_grid != 0) && k >= kAfSplitCols / 16 && k < nk;
        const uint32_t a_col = hi_part ? kDtHiCol + 8 * (k - kAfSplitCols / 16) : 8 * k;
        umma_ts(tmem + 192, tmem + a_col, umma_smem_desc_sw128(aK + k * 2048, 8192, 1024), id_o, k > 0);
      };
      if (!p.af_mma) {
        mbar_wait(bar_cols, 0);     // every row thread has packed its dT: the key part runs under the bucket-sum tail
        tc_fence_after();
        for (int k = 0; k < nk; ++k) dq_step(k);
      }
      mbar_wait(bar_p, 0);
      tc_fence_after();
      ROWS_TRACE(3);
      if (p.af_mma) {
        const uint32_t id_abs = umma_idesc_bf16(128, kFeat, 0, 0);
        for (int k = 0; k < nk; ++k) {            // dQ = dT . K  (the table part follows the un-shifted dR)
          const uint32_t a_col = k >= kAfSplitCols / 16 ? kDtHiCol + 8 * (k - kAfSplitCols / 16) : 8 * k;
          umma_ts(tmem + 192, tmem + a_col, umma_smem_desc_sw128(aK + k * 2048, 8192, 1024), id_o, k > 0);
        }
        for (int k = 0; k < nk; ++k) {            // PBabs = P . Ind^T,  dRabs = dT . Ind^T
          const uint32_t d_col = k >= kAfSplitCols / 16 ? kDtHiCol + 8 * (k - kAfSplitCols / 16) : 8 * k;
          const uint32_t p_col = k >= kAfSplitCols / 16 ? kPHiCol + 8 * (k - kAfSplitCols / 16) : 256 + 8 * k;
          const uint64_t bd = umma_smem_desc_sw128(aInd + (k >> 2) * kIndChunk + (k & 3) * 32, 16, 1024);
          umma_ts(tmem + kAbsPB, tmem + p_col, bd, id_abs, k > 0);
          umma_ts(tmem + kAbsDR, tmem + d_col, bd, id_abs, k > 0);
        }
        umma_commit(bar_abs);
        mbar_wait(bar_p2, 0);
        tc_fence_after();
        for (int k = nk; k < ksteps; ++k)         // dQ += dR . TK
          umma_ts(tmem + 192, tmem + kDrPackCol + 8 * (k - nk), umma_smem_desc_sw128(aK + k * 2048, 8192, 1024), id_o, 1u);
      } else {
        for (int k = nk; k < ksteps; ++k) dq_step(k);   // the table part, behind the packed dR
      }
      umma_commit(bar_o);
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 1) >> 2;            // which column half of the row this thread owns
    const int r_local = quarter * 32 + lane;
    const int row = m0 + r_local;
    RowCtx x;
    x.half = half;
    x.trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    x.s_r = smem_u32(sR) + r_local * kStride * 4;
    x.s_dpb = smem_u32(sdPB) + r_local * kStride * 4;
    x.s_pb = smem_u32(sQ) + r_local * 64 * 4;
    x.s_dr = smem_u32(sV) + r_local * kStride * 4;
    x.s_bias = smem_u32(sBias);
    x.s_lut = smem_u32(sLut);
    x.row = row;
    x.row_c = min(row, p.N - 1);
    x.drow = p.dense ? p.dense + b * p.dense_sb + head * p.dense_sh + x.row_c * p.dense_si : nullptr;
    x.ddrow = p.ddense ? p.ddense + ((static_cast<int64_t>(b) * p.H + head) * p.N + x.row_c) * p.N : nullptr;
    x.sw = (2 * r_local) & 63;
    x.bar_cols = bar_cols;
    x.stage = p.stage;
    x.lane = lane;
    x.row0 = m0 + quarter * 32;
    x.bh = b * p.H + head;
    x.map_sp = &map_sp;
    x.map_sd = &map_sd;
    x.s_stage = half == 0 ? smem_u32(sO) + quarter * 4096 : smem_u32(sStage) + quarter * 4096;

    const int tslot = threadIdx.x == 32 ? 8 : (threadIdx.x == 160 ? 24 : -1);
#define RT(k) do { if (tslot >= 0) ROWS_TRACE(tslot + (k)); } while (0)
    // delta_i = dO_i . O_i from the TMA-staged tiles (same SWIZZLE_128B layout: 16-byte chunk c of row r sits at
    // chunk c ^ (r & 7)).  Reading the 2 x 128 bytes of a row straight from global - one row per lane - costs 32
    // L1 tag cycles per warp instruction and used to hold the row threads back for ~5 us per tile.
    float delta = 0.f, lse = 0.f;
    if (row < p.N) lse = p.lse[(static_cast<int64_t>(b) * p.H + head) * p.N + row];
    mbar_wait(bar_ld, 0);
    {
      const uint32_t o_row = smem_u32(sO) + r_local * 128, g_row = smem_u32(sdO) + r_local * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t off = static_cast<uint32_t>((q ^ (r_local & 7)) << 4);
        const uint4 a = lds_u32x4(o_row + off), g = lds_u32x4(g_row + off);
        const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
        const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
        delta += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y +
                 a3.x * g3.x + a3.y * g3.y;
      }
      if (row >= p.N) delta = 0.f;           // rows past N are zero-filled by TMA; keep them exactly inert
    }
    RT(0);
    if (any_r) {
      mbar_wait(bar_r, 0);
      tc_fence_after();
      RT(1);
      {
        const int c = half;                      // each thread of the pair stages 32 of the 64 buckets
        uint32_t raw_r[32], raw_g[32];
        if (p.ctx_k) tmem_ld32(x.trow + c * 32, raw_r);
        if (p.ctx_v) tmem_ld32(x.trow + 64 + c * 32, raw_g);
        tmem_ld_wait();
        if (!p.af_mma) {
          // R / dPB are in registers: the T = Q K^T MMAs may overwrite their columns while the values are staged
          tc_fence_before();
          mbar_arrive(bar_rfree);
        }
        if (p.ctx_k) {
          const float rs = p.af_mma ? 1.0f : p.scale;   // tensor-core mode adds the UNSCALED term into T
#pragma unroll
          for (int i = 0; i < 16; ++i)
            sts_f32x2(x.s_r + 4 * (c * 32 + 2 * i), rs * __uint_as_float(raw_r[2 * i]), rs * __uint_as_float(raw_r[2 * i + 1]));
        }
        if (p.ctx_v) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            sts_f32x2(x.s_dpb + 4 * (c * 32 + 2 * i), __uint_as_float(raw_g[2 * i]), __uint_as_float(raw_g[2 * i + 1]));
        }
      }
      if (p.af_mma) {
        // A vectors (absolute key features, see attention_fwd.cu): the first thread of a row builds the
        // logit-side vector from R, its partner the dP-side vector from dPB; bf16 hi + lo parts
        rows_barrier();                          // both halves of R / dPB staged
        constexpr int G = 14;
        const int M1 = p.af_max_rel + 1;
        const bool patch = row >= 1 && row < p.N;
        const int qi = patch ? row - 1 : 0;
        const int ri = qi / G, ci = qi - ri * G;
        const uint32_t src = half ? x.s_dpb : x.s_r;
        const float v0 = lds_f32(src), h0 = lds_f32(src + 4 * 32);
        float a[kFeat];
#pragma unroll
        for (int t = 0; t < G; ++t) {
          a[t] = patch ? lds_f32(src + 4 * (M1 - ri + t)) : v0;
          a[G + t] = patch ? lds_f32(src + 4 * (32 + M1 - ci + t)) : h0;
        }
        a[2 * G] = v0 + h0;
#pragma unroll
        for (int t = 2 * G + 1; t < kFeat; ++t) a[t] = 0.f;
        uint32_t hiw[16], low[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const __nv_bfloat16 b0 = __float2bfloat16_rn(a[2 * t]), b1 = __float2bfloat16_rn(a[2 * t + 1]);
          hiw[t] = pack_bf16x2(__bfloat162float(b0), __bfloat162float(b1));
          low[t] = pack_bf16x2(a[2 * t] - __bfloat162float(b0), a[2 * t + 1] - __bfloat162float(b1));
        }
        tmem_st16(x.trow + (half ? kAvecDPB : kAvecR), hiw);
        tmem_st16(x.trow + (half ? kAvecDPBlo : kAvecR + 16), low);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_rfree);
      } else {
        rows_barrier();                          // both halves of R / dPB staged before anyone gathers
      }
    }
    RT(2);

    x.delta = delta;
    x.lsel = lse * kLog2e;
    x.wrow = ((static_cast<int64_t>(b) * p.H + head) * p.N + row) * p.ldw;
    RT(3);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    RT(4);

    if (p.af_mma) {
      bwd_row_plain(p, x);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);                        // packed P / dT in place: dQ = dT K, PBabs, dRabs may start
      // un-shift the absolute bucket sums into this row's 64 packed buckets: the first thread of the row
      // takes PBabs -> PB (rotated shared row), its partner dRabs -> dR
      constexpr int G = 14;
      const int M1 = p.af_max_rel + 1;
      const bool patch = row >= 1 && row < p.N;
      const int qi = patch ? row - 1 : 0;
      const int ri = qi / G, ci = qi - ri * G;
      auto put = [&](int bucket, float v) {
        sts_f32(half ? x.s_dr + 4 * bucket : x.s_pb + 4 * ((bucket + x.sw) & 63), v);
      };
      for (int k = 0; k < kNB; ++k) put(k, 0.f);
      mbar_wait(bar_abs, 0);
      tc_fence_after();
      uint32_t ab[32];
      tmem_ld32(x.trow + (half ? kAbsDR : kAbsPB), ab);
      tmem_ld_wait();
      if (patch) {
#pragma unroll
        for (int t = 0; t < G; ++t) {
          put(M1 - ri + t, __uint_as_float(ab[t]));
          put(32 + M1 - ci + t, __uint_as_float(ab[G + t]));
        }
        put(0, __uint_as_float(ab[2 * G]));
        put(32, __uint_as_float(ab[2 * G]));
      } else if (half == 0) {                  // cls query row: bucket sums of P are genuine, those of dT exactly 0
        float tot = __uint_as_float(ab[2 * G]);
#pragma unroll
        for (int t = 0; t < G; ++t) tot += __uint_as_float(ab[t]);
        put(0, tot);
        put(32, tot);
      }
      rows_barrier();
    } else if (p.af_grid == 14) bwd_row_af<14>(p, x);
    else if (p.gp_grid == 14) bwd_row_gridprod<14>(p, x);
    else bwd_row_generic(p, x);
    RT(5);

    // bucket sums: PB -> workspace ; dR -> workspace + TMEM (A operand of the dQ MMA)
    {
      const int c = half;
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int b0 = c * 32 + 2 * k;
        // b0 and the rotation are even: the pair (b0, b0 + 1) is adjacent and 8-byte aligned in both rows
        const float2 qq = lds_f32x2(x.s_pb + 4 * ((b0 + x.sw) & 63)), rr = lds_f32x2(x.s_dr + 4 * b0);
        const float q0 = qq.x, q1 = qq.y, r0 = rr.x, r1 = rr.y;
        if (p.gp_grid != 0) {          // grid-product path: the two halves' dR rows; no value-side bucket sums
          pk[k] = 0u;
          dk[k] = pack_bf16x2(q0 + r0, q1 + r1);
        } else {
          pk[k] = pack_bf16x2(q0, q1);
          dk[k] = pack_bf16x2(r0, r1);
        }
      }
      if (p.ctx_k) tmem_st16(x.trow + (p.af_mma ? kDrPackCol : Npad / 2) + c * 16, dk);
      if (p.stage && !(CREAM_ABL & 2)) {
        // this thread's 32 buckets = 64 bytes = one full row of the warp's tiles: workspace columns [Npad + 32 c, + 32)
        stage_wait_free(x);
        const uint32_t rowb = x.s_stage + lane * 64;
        const uint32_t sw4 = (static_cast<uint32_t>(lane) >> 1) & 3u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t off = (static_cast<uint32_t>(q) ^ sw4) << 4;
          sts_u32x4s(rowb + off, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          sts_u32x4s(rowb + 2048 + off, dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
        }
        stage_store(x, Npad + c * 32);
      } else if (row < p.N && !(CREAM_ABL & 2)) {
        __nv_bfloat16* wp = p.ws_p + x.wrow + Npad + c * 32;
        __nv_bfloat16* wd = p.ws_dt + x.wrow + Npad + c * 32;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          stg_256(wp + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4], pk[8 * q + 5], pk[8 * q + 6], pk[8 * q + 7]);
          stg_256(wd + 16 * q, dk[8 * q], dk[8 * q + 1], dk[8 * q + 2], dk[8 * q + 3], dk[8 * q + 4], dk[8 * q + 5], dk[8 * q + 6], dk[8 * q + 7]);
        }
      }
    }
    if (p.dbias != nullptr && row < p.N && half == 0) {
      for (int k = 0; k < kNB; ++k) atomicAdd(reinterpret_cast<float*>(sDbias) + k, lds_f32(x.s_dr + 4 * k));
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(p.af_mma ? bar_p2 : bar_p);
    RT(6);

    const uint32_t trow = x.trow;
    mbar_wait(bar_o, 0);
    tc_fence_after();
    RT(7);
    __nv_bfloat16* qrow = p.dqkv + (static_cast<int64_t>(b) * p.N + row) * p.lddqkv + head * kD;
    {
      const int c = half;
      uint32_t raw[32];
      tmem_ld32(trow + 192 + c * 32, raw);
      tmem_ld_wait();
      if (row < p.N && !(CREAM_ABL & 2)) store_row32_bf16(qrow + c * 32, raw, p.scale, p.wide_out != 0);
    }
    if (p.stage && lane == 0) bulk_wait_all();   // this warp's TMA stores have completed (tiles read, writes performed)
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbias != nullptr && threadIdx.x < 64)
    atomicAdd(p.dbias + tab * 64 + threadIdx.x, lds_f32(smem_u32(sDbias) + 4 * threadIdx.x));
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
    if (lane == 0) ROWS_TRACE(41);
  }
}

// -------------------------------------------------------------------------------------------
// Column-form products.  One CTA per (head, batch):
//   acc_v[m, d] = sum_i Wp[i, m] dO[i, d]        m in [0, Npad+64)   ([P | PB]^T dO)
//   acc_k[m, d] = sum_i Wd[i, m] Q[i, d]                             ([dT | dR]^T Q)
// M is covered by 3 tiles of 128 (cols beyond Npad+64 are zero-filled by TMA).
// warp 0: TMA, warp 1: MMA, warps 2..5: epilogue.
// -------------------------------------------------------------------------------------------
constexpr int kColsThreads = 192;
constexpr int kColStages = 4;
constexpr int kColBoxes = 5;       // 64-column boxes staged per workspace: covers Npad + 64 <= 320 columns
constexpr int kColStageBytes = (kColBoxes + 1) * 8192;      // one product, 64 query rows: 5 A boxes + dO (or Q) = 48 KB
constexpr int kRedRowBytes = 272;  // one table-gradient row (64 fp32) + 16 B pad: conflict-free private rows
constexpr int kRedBytes = 2 * kNB * kRedRowBytes;            // [dTV | dTK] rows staged for the bulk reduce

struct BwdColsParams {
  int B, H, N, Npad, ldw;
  float scale;
  int shared_tables;
  __nv_bfloat16* dqkv; int64_t lddqkv;
  float* dtk; float* dtv;
  int wide_out;
};

__global__ void __launch_bounds__(kColsThreads, 1)
attn_bwd_cols_kernel(const __grid_constant__ CUtensorMap map_wp, const __grid_constant__ CUtensorMap map_wd,
                     const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_q,
                     const BwdColsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  pdl_trigger();
  uint8_t* red_rows = smem + kColStages * kColStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(red_rows + kRedBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kColStages;
  uint64_t* done = bars + 2 * kColStages;           // [2] MMA -> epilogue, one per product
  uint64_t* acc_free = done + 2;                    // [2] epilogue -> MMA: that product's accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = ceil_div(p.N, 64);
  const int mtiles = ceil_div(p.Npad + kNB, 128);  // <= 3
  const int nboxes = min(2 * mtiles, kColBoxes);    // 64-column boxes actually consumed per workspace
  const int items = p.B * p.H;                      // persistent: item = (batch, head)

  // The two products of an item run back to back through ONE stage ring (product 0: [P|PB]^T dO,
  // product 1: [dT|dR]^T Q), each into its own 192 accumulator columns, so the epilogue of one
  // product drains while the tensor core works on the other.
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_wp);
    prefetch_tmap(&map_wd);
    for (int s = 0; s < kColStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(&done[q], 1); mbar_init(&acc_free[q], 4); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();      // the rows kernel's workspace / dQ writes are complete and visible from here on

  if (warp == 0 && lane == 0) {
    // loads run ahead across products and items
    int it = 0;
    for (int w0 = blockIdx.x; w0 < items; w0 += gridDim.x) {
      const int w = items - 1 - w0;      // newest workspace rows first: they are still in L2 (see the epilogue loop)
      const int b = w / p.H, head = w - b * p.H;
      for (int prod = 0; prod < 2; ++prod) {
        const CUtensorMap* ma = prod ? &map_wd : &map_wp;
        const CUtensorMap* mb = prod ? &map_q : &map_do;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kColStages;
          mbar_wait(&empty[s], ((it / kColStages) & 1) ^ 1);
          uint8_t* st = smem + s * kColStageBytes;
          mbar_arrive_expect_tx(&full[s], (nboxes + 1) * 8192);
          for (int c = 0; c < nboxes; ++c) tma_load_3d(st + c * 8192, ma, &full[s], c * 64, kb * 64, w);
          tma_load_3d(st + kColBoxes * 8192, mb, &full[s], head * kD, kb * 64, b);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, kD, 1, 1);
    int it = 0, n = 0;
    for (int w0 = blockIdx.x; w0 < items; w0 += gridDim.x, ++n) {
      for (int prod = 0; prod < 2; ++prod) {
        mbar_wait(&acc_free[prod], (n & 1) ^ 1);     // the previous item's accumulators of this product were read
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kColStages;
          mbar_wait(&full[s], (it / kColStages) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * kColStageBytes);
          for (int mt = 0; mt < mtiles; ++mt) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // A: two 64-wide MN chunks (LBO = 8192) of 64 K-rows; B: one chunk.  The last M tile
              // covers only 16 valid columns; its second chunk is whatever follows in the stage
              // (finite bf16 data) and lands in accumulator lanes nobody reads.
              umma_ss(tmem + prod * 192 + mt * 64, umma_smem_desc_sw128(st + (2 * mt) * 8192 + k * 2048, 8192, 1024),
                      umma_smem_desc_sw128(st + kColBoxes * 8192 + k * 2048, 8192, 1024), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&done[prod]);
      }
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    int n = 0;
    // Items run from the LAST (batch, head) to the first: the rows kernel wrote the workspace in
    // ascending order, so its most recent ~100 MB are still resident in the 126 MB L2 when this kernel starts.
    for (int w0 = blockIdx.x; w0 < items; w0 += gridDim.x, ++n) {
      const int w = items - 1 - w0;
      const int b = w / p.H, head = w - b * p.H;
      const int tab = p.shared_tables ? 0 : head;
      for (int which = 0; which < 2; ++which) {          // 0: dV / dTV   1: dK / dTK
        mbar_wait(&done[which], n & 1);
        tc_fence_after();
        const float mul = which ? p.scale : 1.0f;
        float* dtab = which ? p.dtk : p.dtv;
        const int col0 = ((which ? 1 : 2) * p.H + head) * kD;
        for (int mt = 0; mt < mtiles; ++mt) {
          const int m = mt * 128 + quarter * 32 + lane;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t raw[32];
            tmem_ld32(trow + which * 192 + mt * 64 + c * 32, raw);
            tmem_ld_wait();
            if (m < p.N) {
              store_row32_bf16(p.dqkv + (static_cast<int64_t>(b) * p.N + m) * p.lddqkv + col0 + c * 32, raw, mul, p.wide_out != 0);
            } else if (m >= p.Npad && m < p.Npad + kNB && dtab != nullptr) {
              // table-gradient row (one bucket): staged in this thread's private shared row and
              // added to global memory by ONE 256-byte bulk reduce (the L2 does the fp32 adds on
              // whole lines) instead of 64 scattered per-element atomics
              const uint32_t srow = smem_u32(red_rows) + (which * kNB + (m - p.Npad)) * kRedRowBytes;
              if (c == 0) bulk_wait_read<0>();           // the previous item's reduce has left the row
#pragma unroll
              for (int q = 0; q < 8; ++q)
                sts_f32x4(srow + c * 128 + q * 16,
                          make_float4(mul * __uint_as_float(raw[4 * q + 0]), mul * __uint_as_float(raw[4 * q + 1]),
                                      mul * __uint_as_float(raw[4 * q + 2]), mul * __uint_as_float(raw[4 * q + 3])));
              if (c == 1) {
                fence_proxy_async_smem();
                bulk_reduce_add_f32(dtab + (static_cast<int64_t>(tab) * kNB + (m - p.Npad)) * kD, srow, kD * 4);
                bulk_commit();
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free[which]);
      }
    }
    bulk_wait_all();   // this thread's table-gradient reduces have completed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace
}  // namespace cb

extern "C" int64_t cream_attn_bwd_workspace_bytes(int B, int H, int N) {
  const int64_t ldw = cb::round_up(N, 16) + 64;
  return 2 * static_cast<int64_t>(B) * H * N * ldw * 2 + 256;
}

extern "C" int cream_attn_bwd(const cream_attn_desc* d, void* stream_) {
  using namespace cb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(d != nullptr && d->qkv && d->out && d->dout && d->dqkv && d->lse && d->workspace, "null pointer");
  CB_REQUIRE(d->head_dim == kD, "head_dim must be 64");
  CB_REQUIRE(d->N >= 1 && d->N <= 208, "1 <= tokens <= 208");
  CB_REQUIRE(d->ld_qkv % 8 == 0 && d->ld_out % 8 == 0 && d->ld_dout % 8 == 0 && d->ld_dqkv % 8 == 0, "leading dims % 8");
  const int Npad = round_up(d->N, 16);
  const int ldw = Npad + kNB;
  CB_REQUIRE(d->workspace_bytes >= cream_attn_bwd_workspace_bytes(d->B, d->H, d->N), "workspace too small");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(d->workspace) & 255) == 0, "workspace alignment");
  const bool ctx_k = d->tk_pack != nullptr, ctx_v = d->tv_pack != nullptr;
  if (ctx_k) CB_REQUIRE(d->idx_a != nullptr && d->dtk_pack != nullptr, "K tables need idx_a and dtk_pack");
  if (ctx_v) CB_REQUIRE(d->idx_va != nullptr && d->dtv_pack != nullptr, "V tables need idx_va and dtv_pack");
  if (d->idx_a || d->idx_va) CB_REQUIRE(d->ld_idx >= Npad && d->ld_idx % 16 == 0, "index pitch");

  __nv_bfloat16* ws_p = static_cast<__nv_bfloat16*>(d->workspace);
  const int64_t ws_elems = static_cast<int64_t>(d->B) * d->H * d->N * ldw;
  __nv_bfloat16* ws_dt = ws_p + ((ws_elems + 127) / 128) * 128;

  BwdRowsParams p{};
  p.B = d->B; p.H = d->H; p.N = d->N; p.Npad = Npad; p.ldw = ldw;
  p.scale = d->scale;
  p.ctx_k = ctx_k; p.ctx_v = ctx_v; p.shared_tables = d->tables_per_head ? 0 : 1;
  p.idx_a = d->idx_a; p.idx_b = d->idx_b; p.idx_va = d->idx_va; p.idx_vb = d->idx_vb; p.ldi = d->ld_idx;
  p.bias = d->bias_pack;
  if (d->af_grid == 14 && d->af_max_rel >= 13 && d->af_grid * d->af_grid + 1 == d->N && ctx_k && ctx_v &&
      !d->bias_pack && 2 * d->af_max_rel + 2 <= 32) {
    p.af_grid = d->af_grid;
    p.af_max_rel = d->af_max_rel;
  }
  p.out = static_cast<const __nv_bfloat16*>(d->out); p.ldo = d->ld_out;
  p.dout = static_cast<const __nv_bfloat16*>(d->dout); p.lddo = d->ld_dout;
  p.lse = d->lse;
  p.dqkv = static_cast<__nv_bfloat16*>(d->dqkv); p.lddqkv = d->ld_dqkv;
  p.wide_out = aligned_for_256bit(d->dqkv, d->ld_dqkv) ? 1 : 0;
  p.ws_p = ws_p; p.ws_dt = ws_dt;
  p.dbias = d->dbias_pack;
  p.dense = d->dense_bias; p.dense_sb = d->dense_stride_b; p.dense_sh = d->dense_stride_h; p.dense_si = d->dense_stride_i;
  p.causal = d->causal;
  CB_REQUIRE(!d->causal || (d->af_grid == 0 && d->gp_grid == 0), "the causal mask runs on the generic gather path (no af / gp hint)");
  p.block_len = d->block_len;
  CB_REQUIRE(d->block_len >= 0 && (d->block_len == 0 || (d->N % d->block_len == 0 && d->af_grid == 0 && d->gp_grid == 0)),
             "block_len must divide N; packed items run on the generic gather path");
  p.ddense = d->ddense;
  if (p.dense != nullptr || p.ddense != nullptr) p.af_grid = 0;   // generic gather path only
  if (p.af_grid != 0) {
    static const bool use_mma = []() { const char* e = getenv("CREAM_AF_MMA"); return e != nullptr && e[0] == '1'; }();
    p.af_mma = use_mma ? 1 : 0;     // CREAM_AF_MMA=1 selects it; default: the register-arithmetic structured path
  }
  if (d->gp_grid == 14 && d->gp_grid * d->gp_grid + 1 == d->N && ctx_k && !ctx_v && d->idx_b == nullptr && !d->bias_pack &&
      p.dense == nullptr && p.ddense == nullptr && p.af_grid == 0 && d->gp_w >= 1 && d->gp_skip_id >= 0 && d->gp_skip_id < kNB) {
    bool ok = true;
    for (int t = 0; t < 2 * d->gp_grid - 1; ++t)
      ok = ok && d->gp_lut_a[t] * d->gp_w + d->gp_lut_b[t] < kNB && d->gp_lut_b[t] < d->gp_w;
    if (ok) {
      p.gp_grid = d->gp_grid; p.gp_w = d->gp_w; p.gp_skip = d->gp_skip_id;
      std::memcpy(p.lut_a, d->gp_lut_a, 32);
      std::memcpy(p.lut_b, d->gp_lut_b, 32);
    }
  }

  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d->H * kD), static_cast<uint64_t>(d->N), static_cast<uint64_t>(d->B)};
  const uint64_t strides[3] = {1, static_cast<uint64_t>(d->ld_qkv), static_cast<uint64_t>(d->N) * d->ld_qkv};
  const uint32_t box_q[3] = {64, 128, 1};
  const uint32_t box_kv[3] = {64, static_cast<uint32_t>(Npad), 1};
  const uint32_t box_64[3] = {64, 64, 1};
  const CUtensorMap* mq = get_tensor_map(d->qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap* mkv = get_tensor_map(d->qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap* mq64 = get_tensor_map(d->qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box_64, CU_TENSOR_MAP_SWIZZLE_128B);
  const uint64_t ddims[3] = {static_cast<uint64_t>(d->H * kD), static_cast<uint64_t>(d->N), static_cast<uint64_t>(d->B)};
  const uint64_t dstrides[3] = {1, static_cast<uint64_t>(d->ld_dout), static_cast<uint64_t>(d->N) * d->ld_dout};
  const CUtensorMap* mdo = get_tensor_map(d->dout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ddims, dstrides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap* mdo64 = get_tensor_map(d->dout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ddims, dstrides, box_64, CU_TENSOR_MAP_SWIZZLE_128B);
  const uint64_t ostrides[3] = {1, static_cast<uint64_t>(d->ld_out), static_cast<uint64_t>(d->N) * d->ld_out};
  const CUtensorMap* mo = get_tensor_map(d->out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ddims, ostrides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
  const int ntab = d->tables_per_head ? d->H : 1;
  const uint64_t tdims[3] = {64, 64, static_cast<uint64_t>(ntab)};
  const uint64_t tstrides[3] = {1, 64, 64 * 64};
  const CUtensorMap* mtk = ctx_k ? get_tensor_map(d->tk_pack, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, tdims, tstrides, box_64, CU_TENSOR_MAP_SWIZZLE_128B) : mq;
  const CUtensorMap* mtv = ctx_v ? get_tensor_map(d->tv_pack, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, tdims, tstrides, box_64, CU_TENSOR_MAP_SWIZZLE_128B) : mq;
  const uint64_t wdims[3] = {static_cast<uint64_t>(ldw), static_cast<uint64_t>(d->N), static_cast<uint64_t>(d->B) * d->H};
  const uint64_t wstrides[3] = {1, static_cast<uint64_t>(ldw), static_cast<uint64_t>(d->N) * ldw};
  const CUtensorMap* mwp = get_tensor_map(ws_p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wdims, wstrides, box_64, CU_TENSOR_MAP_SWIZZLE_128B);
  const CUtensorMap* mwd = get_tensor_map(ws_dt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wdims, wstrides, box_64, CU_TENSOR_MAP_SWIZZLE_128B);
  // store-side maps of the workspace: 32 columns (64 bytes) x 32 rows per box, SWIZZLE_64B (the rows kernel's staging tiles)
  const uint32_t box_st[3] = {32, 32, 1};
  const CUtensorMap* msp = get_tensor_map(ws_p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wdims, wstrides, box_st, CU_TENSOR_MAP_SWIZZLE_64B);
  const CUtensorMap* msd = get_tensor_map(ws_dt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wdims, wstrides, box_st, CU_TENSOR_MAP_SWIZZLE_64B);
  if (!mq || !mkv || !mq64 || !mdo || !mdo64 || !mo || !mtk || !mtv || !mwp || !mwd || !msp || !msd) return CREAM_ERR_CUDA;
  {
    static const bool stage_on = []() { const char* e = getenv("CREAM_BWD_STAGE"); return e != nullptr && e[0] == '1'; }();   // measured slower: off unless asked for
    p.stage = (stage_on && (p.af_grid != 0 || p.gp_grid != 0) && !p.af_mma) ? 1 : 0;
  }

  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const size_t smem_rows = 3 * 16384 + static_cast<size_t>(Npad) * 128 + 8192 +
                           std::max<size_t>(Npad * 128, 26 * 1024) + 8192 + 2 * 128 * kStride * 4 + 4 * kIndChunk + 2 * 64 * 4 + 64 + 128 +
                           (p.stage ? 1024 + 4 * 4096 : 0);   // + the second half's staging tiles, 1024-aligned
  CB_REQUIRE(smem_rows <= 227 * 1024, "shared memory budget");
  dim3 grid(ceil_div(d->N, 128), d->H, d->B);
#ifdef CREAM_TRACE
  static long long* trace_dev = nullptr;
  if (getenv("CREAM_ATTN_TRACE") != nullptr) {
    if (trace_dev == nullptr) CB_CUDA_OK(cudaMalloc(&trace_dev, 128 * sizeof(long long)));
    CB_CUDA_OK(cudaMemsetAsync(trace_dev, 0, 128 * sizeof(long long), stream));
    p.trace = trace_dev;
  }
#endif
  CB_CUDA_OK(launch_chain(attn_bwd_rows_kernel, grid, dim3(kRowsThreads), smem_rows, stream, 1, *mq, *mkv, *mdo, *mo, *mtk, *mtv, *msp, *msd, p));
  int rc = check_last("attn_bwd_rows_kernel");
  if (rc) return rc;
#ifdef CREAM_TRACE
  if (p.trace != nullptr) {
    static int dumps = 0;
    if ((p.af_grid != 0 || p.gp_grid != 0) && dumps++ < 6) {
      long long hh[128];
      CB_CUDA_OK(cudaStreamSynchronize(stream));
      CB_CUDA_OK(cudaMemcpy(hh, trace_dev, sizeof(hh), cudaMemcpyDeviceToHost));
      for (int cta = 0; cta < 2; ++cta) {
        const long long* h = hh + 64 * cta;
        const long long t0 = h[0];
        fprintf(stderr, "ROWS TRACE cta %s (cycles from first TMA issue) B %d H %d N %d af %d gp %d\n", cta ? "mid-grid" : "first", p.B, p.H, p.N, p.af_grid, p.gp_grid);
        fprintf(stderr, "  kernel entry %lld | exit %lld\n", h[40] - t0, h[41] - t0);
        fprintf(stderr, "  mma thread: loads landed %lld | T,dP issued %lld | bar_p %lld\n", h[1] - t0, h[2] - t0, h[3] - t0);
        for (int q = 0; q < 2; ++q) {
          const long long* e = h + 8 + 16 * q;
          fprintf(stderr, "  row thread half %d: delta done %lld | R ready %lld | staged %lld | - %lld | T ready %lld | columns done %lld | tail done %lld | dQ ready %lld\n",
                  q, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[7] - t0);
        }
      }
    }
  }
#endif

  BwdColsParams c{};
  c.B = d->B; c.H = d->H; c.N = d->N; c.Npad = Npad; c.ldw = ldw;
  c.scale = d->scale;
  c.shared_tables = d->tables_per_head ? 0 : 1;
  c.dqkv = static_cast<__nv_bfloat16*>(d->dqkv); c.lddqkv = d->ld_dqkv;
  c.wide_out = p.wide_out;
  c.dtk = ctx_k ? d->dtk_pack : nullptr;
  c.dtv = ctx_v ? d->dtv_pack : nullptr;
  const size_t smem_cols = kColStages * kColStageBytes + kRedBytes + 256;
  const int grid2 = std::min(d->B * d->H, kNumSMs);
  CB_CUDA_OK(launch_chain(attn_bwd_cols_kernel, dim3(grid2), dim3(kColsThreads), smem_cols, stream, 1, *mwp, *mwd, *mdo64, *mq64, c));
  return check_last("attn_bwd_cols_kernel");
}
