// Sliced bf16 GEMM for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma
// (accumulators in TMEM, double buffered) -> epilogue warps (tcgen05.ld) -> TMA store.
//
// Replaces the cuBLAS calls behind F.linear on the sampled-subnet path
// (AutoFormer/model/module/Linear_super.py:52-54, qkv_super.py:53-55) and their
// autograd backward (dgrad / wgrad).  The sampled slice is carried by the TMA
// descriptor extents over the FULL supernet tensors, so no sliced weight copy exists
// and ragged K / N tails are zero-filled by the TMA unit.
//
// Persistent kernel: one CTA per SM, static round-robin tile schedule.
//   warp 0      : TMA producer (one lane)
//   warp 1      : tcgen05.mma issuer (one lane)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue.  The accumulator tile is drained in 128-byte-wide column blocks
//                 (64 bf16 / 32 fp32 columns): TMEM -> registers -> epilogue math -> 128B-swizzled
//                 shared staging tile -> ONE TMA store (or TMA reduce-add for the split-K weight
//                 gradient) per block, double buffered.  Residual / GELU pre-activation operands of
//                 the epilogue are TMA-prefetched into the same staging tile one block ahead, so the
//                 epilogue issues no strided global accesses and needs no bounds predicates (the
//                 tensor maps clip ragged M / N tails).
#include <cstdlib>
#include <vector>

#include "attention_common.cuh"  // explicit shared-memory accessors

namespace cb {

namespace {

constexpr int kBM = 128;          // UMMA M
constexpr int kBK = 64;           // K block = one 128-byte swizzle row of bf16
constexpr int kMaxStages = 8;
constexpr int kNumThreads = 384;  // 12 warps
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kEpiThreads = kNumEpiWarps * 32;
constexpr uint32_t kTmemCols = 512;  // 2 accumulator stages x 256 columns
constexpr int kEpiSlotBytes = kBM * 128;  // one 128 rows x 128 B staging tile

struct GemmKernelParams {
  int M, N, K, groups, BN;
  int num_mt, num_nt, split_k, kb_total, kb_per_split, total_work;
  int a_mn, b_mn, a_group_off, kpg;
  int b_group_rows;
  int num_stages;
  uint32_t stage_bytes, a_bytes, tx_bytes;
  int nb64;          // 64-column chunks of this CTA's B stage (MN-major B)
  int bn_cta;        // B rows (N) staged by one CTA: BN, or BN / 2 for a CTA pair
  int out_g_col;
  const float* bias;
  const float* row_scale;
  int rows_per_scale;
  float alpha;
  long long* trace;   // CREAM_TRACE builds only
};

#ifdef CREAM_TRACE
#define CB_TRACE(region, idx, k)                                                         \
  do {                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 96) p.trace[(region) * 1024 + (idx) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define CB_TRACE(region, idx, k) do {} while (0)
#endif

struct WorkItem {
  int mt, nt, g, kb0, kb1;
};

__device__ __forceinline__ WorkItem decode_work(const GemmKernelParams& p, int w) {
  // N-fastest: the (group, n-tile) siblings of one M tile run on neighbouring CTAs at the same
  // time, so the A tile (the activation stream) is fetched from HBM once and hit in L2 by the rest.
  WorkItem it;
  it.nt = w % p.num_nt;
  int rest = w / p.num_nt;
  it.g = rest % p.groups;
  rest /= p.groups;
  it.mt = rest % p.num_mt;
  const int ks = rest / p.num_mt;
  it.kb0 = ks * p.kb_per_split;
  it.kb1 = min(p.kb_total, it.kb0 + p.kb_per_split);
  return it;
}

// ---- TMA store / reduce (bulk async-group completion) ---------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void epi_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}

template <int EPI> struct EpiTraits {
  static constexpr bool kOutBf16 = EPI == CREAM_EPI_BF16 || EPI == CREAM_EPI_BF16_GELU || EPI == CREAM_EPI_BF16_DGELU;
  static constexpr int kCB = kOutBf16 ? 64 : 32;       // columns per 128-byte store block
  static constexpr int kPerThread = kCB / 2;           // columns per thread (two warps per lane quarter)
  static constexpr bool kLoads = EPI == CREAM_EPI_F32_RESID || EPI == CREAM_EPI_BF16_DGELU;
  static constexpr bool kBias = EPI != CREAM_EPI_F32_ATOMIC && EPI != CREAM_EPI_BF16_DGELU;
};

__device__ __forceinline__ uint4 lds_u32x4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32x4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int EPI, bool kPair>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                 const __grid_constant__ CUtensorMap tmap_out_h, const __grid_constant__ CUtensorMap tmap_aux_h,
                 const GemmKernelParams p) {
  using T = EpiTraits<EPI>;
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  pdl_trigger();   // the next kernel of the chain may be scheduled as this grid's CTAs retire (it blocks in pdl_wait)
  uint8_t* epi_slots = smem + p.num_stages * p.stage_bytes;             // 2 x 16 KB, 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_slots + 2 * kEpiSlotBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* ld_bar = tmem_empty + 2;                                    // [2] epilogue operand prefetch
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: CTAs (2c, 2c+1) form cluster c and share one 256-row tile; the even CTA leads
  // (issues the MMAs, owns the full / accumulator-empty barriers).
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int first_work = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int work_stride = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int kTileM = kPair ? 2 * kBM : kBM;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kNumEpiWarps * (kPair ? 2 : 1));
      mbar_init(&ld_bar[a], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kPair) tmem_alloc_pair<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above ran under the previous kernel's tail; no global access before this point

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int w = first_work; w < p.total_work; w += work_stride) {
      const WorkItem it = decode_work(p, w);
      const int m0 = it.mt * kTileM + static_cast<int>(rank) * kBM;      // this CTA's 128 rows of A
      const int n0 = it.nt * p.BN + static_cast<int>(rank) * p.bn_cta;    // this CTA's part of B
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * p.stage_bytes;
        uint8_t* sb = sa + p.a_bytes;
        if constexpr (!kPair) {
          mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes);
          if (!p.a_mn) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBK, m0);
          } else {
            const int c = it.g * p.a_group_off + m0;
            tma_load_2d(sa, &tmap_a, &full_bar[stage], c, kb * kBK);
            tma_load_2d(sa + 8192, &tmap_a, &full_bar[stage], c + 64, kb * kBK);
          }
          if (!p.b_mn) {
            tma_load_3d(sb, &tmap_b, &full_bar[stage], kb * kBK, n0, it.g);
          } else {
            const int krow = (kb / p.kpg) * p.b_group_rows + (kb % p.kpg) * kBK;
            for (int c = 0; c < p.nb64; ++c)
              tma_load_2d(sb + c * 8192, &tmap_b, &full_bar[stage], n0 + c * 64, krow);
          }
        } else {
          // all bytes of the pair are counted on the LEADER's barrier
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * p.tx_bytes);
          if (!p.a_mn) {
            tma_load_2d_pair(sa, &tmap_a, fb, kb * kBK, m0);
          } else {
            const int c = it.g * p.a_group_off + m0;
            tma_load_2d_pair(sa, &tmap_a, fb, c, kb * kBK);
            tma_load_2d_pair(sa + 8192, &tmap_a, fb, c + 64, kb * kBK);
          }
          if (!p.b_mn) {
            tma_load_3d_pair(sb, &tmap_b, fb, kb * kBK, n0, it.g);
          } else {
            const int krow = (kb / p.kpg) * p.b_group_rows + (kb % p.kpg) * kBK;
            for (int c = 0; c < p.nb64; ++c) tma_load_2d_pair(sb + c * 8192, &tmap_b, fb, n0 + c * 64, krow);
          }
        }
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    const uint32_t idesc = umma_idesc_bf16(kTileM, p.BN, p.a_mn, p.b_mn);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    [[maybe_unused]] int tile_no = 0;
    for (int w = first_work; w < p.total_work; w += work_stride) {
      const WorkItem it = decode_work(p, w);
      CB_TRACE(2, tile_no, 0);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      CB_TRACE(2, tile_no, 1);
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sb = sa + p.a_bytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // K-major: advance 32 bytes inside the 128B swizzle row; SBO = 8 rows * 128B.
          // MN-major: advance 16 K-rows (2048B); LBO = next 64-wide MN chunk (8192B).
          const uint64_t adesc = p.a_mn ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                        : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t bdesc = p.b_mn ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                        : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
          if constexpr (kPair) umma_ss_pair(d_tmem, adesc, bdesc, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
          else umma_ss(d_tmem, adesc, bdesc, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
        }
        // frees this stage in both CTAs of a pair
        if constexpr (kPair) umma_commit_pair(&empty_bar[stage]);
        else umma_commit(&empty_bar[stage]);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
      if constexpr (kPair) umma_commit_pair(&tmem_full[acc]);
      else umma_commit(&tmem_full[acc]);
      CB_TRACE(2, tile_no, 2);
      ++tile_no;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;        // which half of the block's columns
    const int r_local = quarter * 32 + lane;         // row within the tile
    const bool issuer = threadIdx.x == kEpiWarp0 * 32;
    const uint32_t slot0 = smem_u32(epi_slots);
    // this thread's four 16-byte chunks in a staging tile: chunk u of row r sits at (u ^ (r & 7))
    uint32_t chunk_full[4], chunk_half[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      chunk_full[q] = r_local * 128 + (((half * 4 + q) ^ (r_local & 7)) << 4);
      chunk_half[q] = r_local * 64 + (q << 4);   // dense 64-byte rows (no swizzle), half-0 threads only
    }
    // A tile whose width is an odd multiple of half a store block ends in a HALF block: it must not
    // spill into the neighbouring N tile, so it moves through the un-swizzled half-width maps.
    auto is_half_block = [&](const WorkItem& it, int cb) {
      const int n0 = it.nt * p.BN;
      const int tile_cols = min(p.BN, p.N - n0);
      return (tile_cols - cb * T::kCB) == T::kCB / 2 && (n0 + tile_cols) < p.N;
    };

    int acc = 0;
    uint32_t acc_phase = 0;
    int slot = 0;                       // staging tile used by the next store job
    uint32_t ld_phase[2] = {0, 0};
    [[maybe_unused]] int tile_no = 0, job_no = 0;

    // operand prefetch (RESID / DGELU): block stream = (work item, column block) pairs
    auto issue_load = [&](int w, int cb, int s) {
      const WorkItem it = decode_work(p, w);
      const bool hb = is_half_block(it, cb);
      mbar_arrive_expect_tx(&ld_bar[s], hb ? kEpiSlotBytes / 2 : kEpiSlotBytes);
      tma_load_3d(epi_slots + s * kEpiSlotBytes, hb ? &tmap_aux_h : &tmap_aux, &ld_bar[s],
                  it.nt * p.BN + cb * T::kCB, it.mt * kTileM + static_cast<int>(rank) * kBM, it.g);
    };
    if (T::kLoads && issuer && first_work < p.total_work) issue_load(first_work, 0, 0);

    // bias of the column block about to be drained (this thread's kPerThread columns), kept in
    // registers and requested one block ahead so its latency never sits on the critical path
    [[maybe_unused]] float bv[T::kPerThread];
    auto load_bias = [&](int col0, int g) {
#pragma unroll
      for (int i = 0; i < T::kPerThread; ++i) bv[i] = 0.f;
      if (p.bias == nullptr) return;
      const float* bp = p.bias + col0 + g * p.out_g_col;
      if (col0 + T::kPerThread <= p.N) {
#pragma unroll
        for (int q = 0; q < T::kPerThread / 4; ++q) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bp) + q);
          bv[4 * q + 0] = bb.x; bv[4 * q + 1] = bb.y; bv[4 * q + 2] = bb.z; bv[4 * q + 3] = bb.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < T::kPerThread; ++i)
          if (col0 + i < p.N) bv[i] = __ldg(bp + i);
      }
    };

    const uint32_t tmem_empty_leader[2] = {mapa_shared(smem_u32(&tmem_empty[0]), 0),
                                           mapa_shared(smem_u32(&tmem_empty[1]), 0)};
    for (int w = first_work; w < p.total_work; w += work_stride) {
      const WorkItem it = decode_work(p, w);
      const int m0 = it.mt * kTileM + static_cast<int>(rank) * kBM, n0 = it.nt * p.BN;
      if constexpr (T::kBias) load_bias(n0 + half * T::kPerThread, it.g);   // overlaps the wait below
      if (issuer) CB_TRACE(1, tile_no, 0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (issuer) CB_TRACE(1, tile_no, 1);
      ++tile_no;
      const int row = m0 + r_local;
      const int tile_cols = min(p.BN, p.N - n0);
      const int nblocks = (tile_cols + T::kCB - 1) / T::kCB;
      float rscale = 1.0f;
      if constexpr (EPI == CREAM_EPI_F32_RESID) {
        if (p.row_scale != nullptr && row < p.M) rscale = __ldg(p.row_scale + row / p.rows_per_scale);
      }
      for (int cb = 0; cb < nblocks; ++cb) {
        // ---- accumulator -> registers ----
        if (issuer) CB_TRACE(0, job_no, 0);
        float v[T::kPerThread];
        {
          const uint32_t taddr = tmem_base + acc * 256 + cb * T::kCB + half * T::kPerThread +
                                 (static_cast<uint32_t>(quarter * 32) << 16);
          if constexpr (T::kPerThread == 32) {
            uint32_t raw[32];
            tmem_ld32(taddr, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          } else {
            uint32_t raw[16];
            tmem_ld16(taddr, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
          }
        }
        if constexpr (T::kBias) {
#pragma unroll
          for (int i = 0; i < T::kPerThread; ++i) v[i] += bv[i];
          // the next block's bias is requested now and consumed after that block's accumulator read
          if (cb + 1 < nblocks) load_bias(n0 + (cb + 1) * T::kCB + half * T::kPerThread, it.g);
        }

        if (issuer) CB_TRACE(0, job_no, 1);
        // ---- acquire the staging tile ----
        if constexpr (T::kLoads) {
          if (issuer) {
            bulk_wait_read<0>();                         // the other tile's store has drained
            int nw = w, ncb = cb + 1;
            if (ncb >= nblocks) { nw = w + work_stride; ncb = 0; }
            if (nw < p.total_work) issue_load(nw, ncb, slot ^ 1);
          }
          mbar_wait(&ld_bar[slot], ld_phase[slot]);
          ld_phase[slot] ^= 1;
        }
        // (without loads: the barrier of the previous store job already proved this tile drained —
        //  the issuer waits for the older store before every barrier, see below)
        if (issuer) CB_TRACE(0, job_no, 2);
        const uint32_t sbase = slot0 + slot * kEpiSlotBytes;
        const bool hb = is_half_block(it, cb);
        const bool active = !hb || half == 0;             // half-1 threads hold columns of the next tile
        uint32_t chunk_off[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) chunk_off[q] = hb ? chunk_half[q] : chunk_full[q];

        // ---- epilogue math + write to the staging tile ----
        [[maybe_unused]] uint4 pre[4];
        if (active) {
        if constexpr (EPI == CREAM_EPI_F32_RESID) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 r = lds_f32x4(sbase + chunk_off[q]);
            sts_f32x4(sbase + chunk_off[q], make_float4(fmaf(rscale, v[4 * q + 0], r.x), fmaf(rscale, v[4 * q + 1], r.y),
                                                        fmaf(rscale, v[4 * q + 2], r.z), fmaf(rscale, v[4 * q + 3], r.w)));
          }
        } else if constexpr (EPI == CREAM_EPI_F32 || EPI == CREAM_EPI_F32_ATOMIC) {
          const float a = (EPI == CREAM_EPI_F32_ATOMIC) ? p.alpha : 1.0f;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts_f32x4(sbase + chunk_off[q], make_float4(a * v[4 * q + 0], a * v[4 * q + 1], a * v[4 * q + 2], a * v[4 * q + 3]));
        } else {
          if constexpr (EPI == CREAM_EPI_BF16_DGELU) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 u = lds_u32x4(sbase + chunk_off[q]);
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              v[8 * q + 0] *= dgelu_f(f0.x); v[8 * q + 1] *= dgelu_f(f0.y);
              v[8 * q + 2] *= dgelu_f(f1.x); v[8 * q + 3] *= dgelu_f(f1.y);
              v[8 * q + 4] *= dgelu_f(f2.x); v[8 * q + 5] *= dgelu_f(f2.y);
              v[8 * q + 6] *= dgelu_f(f3.x); v[8 * q + 7] *= dgelu_f(f3.y);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
            u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
            u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
            sts_u32x4(sbase + chunk_off[q], u);
            if constexpr (EPI == CREAM_EPI_BF16_GELU) pre[q] = u;
          }
        }
        }  // active
        fence_proxy_async_smem();
        if (issuer) CB_TRACE(0, job_no, 3);
        if constexpr (!T::kLoads) {
          if (issuer) bulk_wait_read<0>();               // previous store job has left its tile
        }
        if (issuer) CB_TRACE(0, job_no, 4);
        epi_barrier();
        if (issuer) CB_TRACE(0, job_no, 5);
        if (issuer) {
          const int c0 = n0 + cb * T::kCB;
          const CUtensorMap* mo = hb ? &tmap_out_h : &tmap_out;
          const CUtensorMap* mx = hb ? &tmap_aux_h : &tmap_aux;
          if constexpr (EPI == CREAM_EPI_F32_ATOMIC) tma_reduce_add_3d(mo, epi_slots + slot * kEpiSlotBytes, c0, m0, it.g);
          else if constexpr (EPI == CREAM_EPI_BF16_GELU) tma_store_3d(mx, epi_slots + slot * kEpiSlotBytes, c0, m0, it.g);
          else tma_store_3d(mo, epi_slots + slot * kEpiSlotBytes, c0, m0, it.g);
          bulk_commit();
        }
        if (issuer) CB_TRACE(0, job_no, 6);
        ++job_no;
        slot ^= 1;

        if constexpr (EPI == CREAM_EPI_BF16_GELU) {
          // second store job of the block: GELU of the bf16-rounded pre-activation (what backward sees)
          const uint32_t sb2 = slot0 + slot * kEpiSlotBytes;
          if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // pre[] holds the bf16 pairs just stored: widen them back instead of re-rounding v[]
            uint4 u;
            u.x = pack_bf16x2(gelu_f(__uint_as_float(pre[q].x << 16)), gelu_f(__uint_as_float(pre[q].x & 0xffff0000u)));
            u.y = pack_bf16x2(gelu_f(__uint_as_float(pre[q].y << 16)), gelu_f(__uint_as_float(pre[q].y & 0xffff0000u)));
            u.z = pack_bf16x2(gelu_f(__uint_as_float(pre[q].z << 16)), gelu_f(__uint_as_float(pre[q].z & 0xffff0000u)));
            u.w = pack_bf16x2(gelu_f(__uint_as_float(pre[q].w << 16)), gelu_f(__uint_as_float(pre[q].w & 0xffff0000u)));
            sts_u32x4(sb2 + chunk_off[q], u);
          }
          }
          fence_proxy_async_smem();
          if (issuer) bulk_wait_read<0>();
          epi_barrier();
          if (issuer) {
            tma_store_3d(hb ? &tmap_out_h : &tmap_out, epi_slots + slot * kEpiSlotBytes, n0 + cb * T::kCB, m0, it.g);
            bulk_commit();
          }
          slot ^= 1;
        }
      }
      // all TMEM reads of this accumulator are done
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_cluster(tmem_empty_leader[acc]);   // the leader's MMA warp waits on it
        else mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) bulk_wait_all();
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();   // no remote arrive / multicast commit may target an exited CTA
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int EPI, bool kPair>
int launch_gemm_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                     const CUtensorMap& toh, const CUtensorMap& txh, const GemmKernelParams& p, size_t smem_bytes,
                     int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, kPair>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  CB_CUDA_OK(launch_chain(gemm_bf16_kernel<EPI, kPair>, dim3(grid), dim3(kNumThreads), smem_bytes, stream, kPair ? 2 : 1,
                          ta, tb, to, tx, toh, txh, p));
  return check_last("gemm_bf16_kernel launch");
}

template <int EPI>
int launch_gemm(bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                const CUtensorMap& toh, const CUtensorMap& txh, const GemmKernelParams& p, size_t smem_bytes,
                int grid, cudaStream_t stream) {
  return pair ? launch_gemm_impl<EPI, true>(ta, tb, to, tx, toh, txh, p, smem_bytes, grid, stream)
              : launch_gemm_impl<EPI, false>(ta, tb, to, tx, toh, txh, p, smem_bytes, grid, stream);
}

}  // namespace

}  // namespace cb

extern "C" int cream_gemm_bf16(const cream_gemm_desc* d, void* stream_) {
  using namespace cb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(d != nullptr, "desc is null");
  CB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->groups >= 1, "empty GEMM");
  CB_REQUIRE(d->a != nullptr && d->b != nullptr && d->out != nullptr, "null operand");
  CB_REQUIRE(d->lda % 8 == 0 && d->ldb % 8 == 0, "bf16 leading dims must be multiples of 8");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->b) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
             "operands must be 16-byte aligned");
  CB_REQUIRE(d->epi >= 0 && d->epi <= CREAM_EPI_F32, "bad epilogue");
  const bool out_bf16 = d->epi == CREAM_EPI_BF16 || d->epi == CREAM_EPI_BF16_GELU ||
                        d->epi == CREAM_EPI_BF16_DGELU;
  if (out_bf16) {
    CB_REQUIRE(d->ldo % 8 == 0 && (d->out_g_col % 8) == 0, "bf16 out pitch/offset % 8");
  } else {
    CB_REQUIRE(d->ldo % 4 == 0 && (d->out_g_col % 4) == 0, "fp32 out pitch/offset % 4");
  }
  if (d->epi == CREAM_EPI_BF16_GELU || d->epi == CREAM_EPI_BF16_DGELU)
    CB_REQUIRE(d->aux != nullptr && d->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0, "aux required");
  if (d->epi == CREAM_EPI_F32_RESID)
    CB_REQUIRE(d->resid != nullptr && d->ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(d->resid) & 15) == 0, "resid");
  if (d->bias) CB_REQUIRE((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "bias must be 16-byte aligned");

  GemmKernelParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K; p.groups = d->groups;
  const bool out_is_bf16 = d->epi == CREAM_EPI_BF16 || d->epi == CREAM_EPI_BF16_GELU || d->epi == CREAM_EPI_BF16_DGELU;
  const int nt = ceil_div(d->N, 256);
  // tile width: a multiple of half a store block (32 bf16 / 16 fp32 columns)
  p.BN = std::min(256, round_up(ceil_div(d->N, nt), out_is_bf16 ? 32 : 16));
  p.num_nt = ceil_div(d->N, p.BN);
  // CTA pair (cta_group::2, 256-row tiles): each CTA stages half of B, which cuts the pair's L2 ->
  // shared-memory traffic by about a third; the mainloop of the path shapes is bound by exactly that.
  // Used when a pair tile is full enough and the extra wave quantisation does not eat the gain.
  // Measured on the supernet-S shapes (profiles/): +5..10 % where the mainloop dominates (K >= 768:
  // fc2 forward, fc1 / qkv dgrad, every wgrad; 8192^3 reaches 1397 TFLOP/s), a loss where the
  // epilogue dominates (GELU / dGELU, short K) or where 256-row tiles add a wave.
  bool pair = d->M >= 2 * kBM && p.BN >= 32 && (p.BN % 16) == 0 && d->K >= 768 &&
              d->epi != CREAM_EPI_BF16_GELU && d->epi != CREAM_EPI_BF16_DGELU;
  if (pair && d->epi != CREAM_EPI_F32_ATOMIC) {
    const int64_t t1 = static_cast<int64_t>(ceil_div(d->M, kBM)) * p.num_nt * d->groups;
    const int64_t t2 = static_cast<int64_t>(ceil_div(d->M, 2 * kBM)) * p.num_nt * d->groups;
    if (ceil_div64(t2, kNumSMs / 2) > ceil_div64(t1, kNumSMs)) pair = false;   // an extra wave
  }
  if (d->cta_pair == 1) pair = false;
  if (d->cta_pair == 2) {
    CB_REQUIRE(p.BN >= 32 && (p.BN % 16) == 0, "CTA pair needs a tile width that is a multiple of 16, >= 32");
    pair = true;
  }
  const int tile_m = pair ? 2 * kBM : kBM;
  const int sm_units = pair ? kNumSMs / 2 : kNumSMs;
  p.num_mt = ceil_div(d->M, tile_m);
  p.bn_cta = pair ? p.BN / 2 : p.BN;
  p.a_mn = d->a_mn ? 1 : 0;
  p.b_mn = d->b_mn ? 1 : 0;
  p.a_group_off = d->a_group_off;
  const int k_groups = std::max(1, d->k_groups);
  if (k_groups > 1) {
    CB_REQUIRE(p.b_mn == 1 && d->k_group_len % kBK == 0 && d->k_group_len * k_groups == d->K,
               "k-groups need MN-major B and 64-multiple group length");
    p.kpg = d->k_group_len / kBK;
  } else {
    p.kpg = 1 << 30;
  }
  p.b_group_rows = static_cast<int>(d->b_group_rows);
  p.kb_total = ceil_div(d->K, kBK);
  const int tiles = p.num_mt * p.num_nt * p.groups;
  int split = 1;
  if (d->epi == CREAM_EPI_F32_ATOMIC) {
    split = d->split_k > 0 ? d->split_k : std::max(1, sm_units / tiles);
    split = std::min(split, p.kb_total);
  }
  p.kb_per_split = ceil_div(p.kb_total, split);
  p.split_k = ceil_div(p.kb_total, p.kb_per_split);
  p.total_work = tiles * p.split_k;
  p.nb64 = ceil_div(p.bn_cta, 64);
  p.a_bytes = kBM * kBK * 2;
  p.stage_bytes = p.a_bytes + p.nb64 * 8192;
  p.tx_bytes = p.a_bytes + (p.b_mn ? p.nb64 * 8192 : p.bn_cta * kBK * 2);
  const size_t tail_bytes = 1024;  // barriers + tmem slot
  const size_t epi_bytes = 2 * kEpiSlotBytes;
  p.num_stages = std::min<int>(kMaxStages, (227 * 1024 - epi_bytes - tail_bytes) / p.stage_bytes);
  CB_REQUIRE(p.num_stages >= 2, "not enough shared memory for 2 stages");
  const size_t smem_bytes = static_cast<size_t>(p.num_stages) * p.stage_bytes + epi_bytes + tail_bytes;
  p.out_g_col = d->out_g_col;
  p.bias = d->bias;
  p.row_scale = d->row_scale; p.rows_per_scale = d->rows_per_scale > 0 ? d->rows_per_scale : 1;
  p.alpha = d->alpha == 0.0f ? 1.0f : d->alpha;

  // ---- tensor maps -----------------------------------------------------------
  const CUtensorMap *ta, *tb;
  if (!p.a_mn) {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->K), static_cast<uint64_t>(d->M)};
    const uint64_t strides[2] = {1, static_cast<uint64_t>(d->lda)};
    const uint32_t box[2] = {kBK, kBM};
    ta = get_tensor_map(d->a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    const uint64_t cols = static_cast<uint64_t>(p.groups - 1) * d->a_group_off + d->M;
    const uint64_t dims[2] = {cols, static_cast<uint64_t>(d->K)};
    const uint64_t strides[2] = {1, static_cast<uint64_t>(d->lda)};
    const uint32_t box[2] = {64, kBK};
    ta = get_tensor_map(d->a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (!p.b_mn) {
    const uint64_t dims[3] = {static_cast<uint64_t>(d->K), static_cast<uint64_t>(d->N),
                              static_cast<uint64_t>(p.groups)};
    const uint64_t gstride = p.groups > 1 ? static_cast<uint64_t>(d->b_group_rows) * d->ldb
                                          : static_cast<uint64_t>(d->N) * d->ldb;
    const uint64_t strides[3] = {1, static_cast<uint64_t>(d->ldb), gstride};
    const uint32_t box[3] = {kBK, static_cast<uint32_t>(p.bn_cta), 1};
    tb = get_tensor_map(d->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    const uint64_t krows = k_groups > 1
                               ? static_cast<uint64_t>(k_groups - 1) * d->b_group_rows + d->k_group_len
                               : static_cast<uint64_t>(d->K);
    const uint64_t dims[2] = {static_cast<uint64_t>(d->N), krows};
    const uint64_t strides[2] = {1, static_cast<uint64_t>(d->ldb)};
    const uint32_t box[2] = {64, kBK};
    tb = get_tensor_map(d->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  }
  // output (and epilogue operand) maps: element (n, m, g) at base + (m*row_mul + g*g_row)*ld + n + g*g_col
  const int row_mul = d->out_row_mul > 0 ? d->out_row_mul : 1;
  auto out_like_map = [&](const void* base, int64_t ld, bool bf16, bool half_width) -> const CUtensorMap* {
    const uint64_t row_stride = static_cast<uint64_t>(row_mul) * ld;
    uint64_t g_stride = static_cast<uint64_t>(d->out_g_row) * ld + d->out_g_col;
    if (p.groups == 1 || g_stride == 0) g_stride = row_stride;
    const uint64_t dims[3] = {static_cast<uint64_t>(d->N), static_cast<uint64_t>(d->M), static_cast<uint64_t>(p.groups)};
    const uint64_t strides[3] = {1, row_stride, g_stride};
    const uint32_t full = bf16 ? 64 : 32;
    const uint32_t box[3] = {half_width ? full / 2 : full, kBM, 1};
    return get_tensor_map(base, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims,
                          strides, box, half_width ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B);
  };
  const CUtensorMap* to = out_like_map(d->out, d->ldo, out_bf16, false);
  const CUtensorMap* toh = out_like_map(d->out, d->ldo, out_bf16, true);
  const CUtensorMap *tx = to, *txh = toh;
  if (d->epi == CREAM_EPI_BF16_GELU || d->epi == CREAM_EPI_BF16_DGELU) {
    tx = out_like_map(d->aux, d->ldaux, true, false);
    txh = out_like_map(d->aux, d->ldaux, true, true);
  }
  if (d->epi == CREAM_EPI_F32_RESID) {
    tx = out_like_map(d->resid, d->ldr, false, false);
    txh = out_like_map(d->resid, d->ldr, false, true);
  }
  if (!ta || !tb || !to || !tx || !toh || !txh) return CREAM_ERR_CUDA;

#ifdef CREAM_TRACE
  static long long* trace_dev = nullptr;
  const char* trace_env = getenv("CREAM_GEMM_TRACE");
  if (trace_env != nullptr) {
    if (trace_dev == nullptr) CB_CUDA_OK(cudaMalloc(&trace_dev, 3 * 1024 * sizeof(long long)));
    CB_CUDA_OK(cudaMemsetAsync(trace_dev, 0, 3 * 1024 * sizeof(long long), stream));
    p.trace = trace_dev;
  }
#endif
  const int grid = pair ? 2 * std::min(p.total_work, kNumSMs / 2) : std::min(p.total_work, kNumSMs);
#ifdef CREAM_TRACE
  if (trace_env != nullptr) {
    int rc;
    switch (d->epi) {
      case CREAM_EPI_BF16: rc = launch_gemm<CREAM_EPI_BF16>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream); break;
      case CREAM_EPI_BF16_GELU: rc = launch_gemm<CREAM_EPI_BF16_GELU>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream); break;
      case CREAM_EPI_F32_RESID: rc = launch_gemm<CREAM_EPI_F32_RESID>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream); break;
      case CREAM_EPI_BF16_DGELU: rc = launch_gemm<CREAM_EPI_BF16_DGELU>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream); break;
      default: rc = CREAM_ERR_ARG;
    }
    static int dumps = 0;
    if (rc == CREAM_OK && dumps < 2) {
      ++dumps;
      std::vector<long long> h(3 * 1024);
      CB_CUDA_OK(cudaStreamSynchronize(stream));
      CB_CUDA_OK(cudaMemcpy(h.data(), trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      const long long t0 = h[2 * 1024];
      fprintf(stderr, "TRACE epi %d M %d N %d K %d BN %d stages %d\n", d->epi, p.M, p.N, p.K, p.BN, p.num_stages);
      for (int t = 0; t < 12; ++t)
        fprintf(stderr, "  mma tile %2d: wait_empty %6lld..%6lld issued %6lld | epi tile: wait_full %6lld..%6lld\n", t,
                h[2048 + t * 8] - t0, h[2048 + t * 8 + 1] - t0, h[2048 + t * 8 + 2] - t0, h[1024 + t * 8] - t0,
                h[1024 + t * 8 + 1] - t0);
      for (int j = 0; j < 40; ++j) {
        const long long* e = &h[j * 8];
        fprintf(stderr, "  job %2d: start %6lld tmem+bias %5lld acquire %5lld math+sts %5lld waitrd %5lld barrier %5lld issue %5lld\n",
                j, e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4], e[6] - e[5]);
      }
    }
    return rc;
  }
#endif
  switch (d->epi) {
    case CREAM_EPI_BF16: return launch_gemm<CREAM_EPI_BF16>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream);
    case CREAM_EPI_BF16_GELU: return launch_gemm<CREAM_EPI_BF16_GELU>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream);
    case CREAM_EPI_F32_RESID: return launch_gemm<CREAM_EPI_F32_RESID>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream);
    case CREAM_EPI_BF16_DGELU: return launch_gemm<CREAM_EPI_BF16_DGELU>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream);
    case CREAM_EPI_F32_ATOMIC: return launch_gemm<CREAM_EPI_F32_ATOMIC>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream);
    default: return launch_gemm<CREAM_EPI_F32>(pair, *ta, *tb, *to, *tx, *toh, *txh, p, smem_bytes, grid, stream);
  }
}
