/*
 * cream_b200 — C ABI of the B200-native ViT block engine.
 *
 * Drop-in boundary for the attention-with-relative-position hot path of
 * microsoft/Cream (AutoFormer supernet, iRPE).  Plain pointers and sizes only; no
 * torch types.  All pointers are DEVICE pointers unless the name ends in `_host`.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * Every function returns CREAM_OK (0) or a CREAM_ERR_* code; nothing throws across
 * the boundary.  Citations `path:line` are relative to the reference checkout.
 *
 * Layout conventions
 *   tokens      : row-major (rows = B*N tokens, cols = features), bf16 activations
 *                 with a leading dimension that is a multiple of 8 elements
 *   residual    : fp32 row-major stream (the reference keeps it in fp32 under autocast)
 *   weights     : fp32 masters in the reference's full-supernet shapes; the engine
 *                 reads bf16 "shadows" of the same shape (QKV de-interleaved, see
 *                 cream_shadow_qkv) and addresses the sampled slice through the TMA
 *                 descriptor — no sliced copy is made.
 */
#ifndef CREAM_B200_H_
#define CREAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CREAM_OK 0
#define CREAM_ERR_ARG 1
#define CREAM_ERR_CUDA 2
#define CREAM_ERR_UNSUPPORTED 3

#define CREAM_DTYPE_F32 0
#define CREAM_DTYPE_F16 1
#define CREAM_DTYPE_BF16 2
#define CREAM_DTYPE_F64 3

/* Library version string, e.g. "cream_b200 0.1.0 (sm_100a)". */
const char* cream_version(void);

/* Version of the rpe_index operator contract; the reference asserts "1.2.0"
 * (iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.py:5-8, rpe_index.cpp:126-128). */
const char* cream_rpe_index_version(void);

/* ------------------------------------------------------------------------- *
 * rpe_index — replaces rpe_index_cpp.forward_gpu / backward_gpu
 * (iRPE/DeiT-with-iRPE/rpe_ops/rpe_index_cuda.cu:54-94 and :96-140).
 *   fwd: Y[b,h,i,j] = input[b,h,i,index[i,j]]       input strides s0..s3 (elements),
 *        index (Lq,Lk) int32 contiguous, Y (B,H,Lq,Lk) contiguous.
 *   bwd: grad_input[b,h,i,index[i,j]] += grad_output[b,h,i,j]; grad_input is
 *        (B,H,Lq,nb) contiguous and caller-zeroed (rpe_index.py:51).  Deterministic
 *        (segmented per-row reduction, no global atomics).
 * ------------------------------------------------------------------------- */
int cream_rpe_index_fwd(const void* input, const int32_t* index, void* out, int B, int H, int Lq,
                        int Lk, int num_buckets, int64_t s0, int64_t s1, int64_t s2, int64_t s3,
                        int dtype, void* stream);
int cream_rpe_index_bwd(void* grad_input, const void* grad_output, const int32_t* index, int B,
                        int H, int Lq, int Lk, int num_buckets, int dtype, void* stream);

/* ------------------------------------------------------------------------- *
 * Integer index tables (bit-exact contracts), written to HOST memory.
 * ------------------------------------------------------------------------- */
/* AutoFormer RelativePosition2D_super index pair for a g x g patch grid plus one
 * cls token (AutoFormer/model/module/multihead_super.py:40-59): n = g*g+1,
 * idx_v/idx_h are (n,n) int32 with values in {0} U [1, 2*max_rel+1]. */
int cream_autoformer_rel_index_host(int grid, int max_rel, int32_t* idx_v_host,
                                    int32_t* idx_h_host);

/* iRPE bucket ids (iRPE/DeiT-with-iRPE/irpe.py:19-52, 131-247, 291-415).
 * method: 0 euclidean, 1 quant, 3 product, 41 cross_rows, 42 cross_cols.
 * out is (skip+h*w, skip+h*w) int32; *num_buckets receives the bucket count
 * including the skip bucket. */
int cream_irpe_bucket_ids_host(int method, int height, int width, int skip, double alpha,
                               double beta, double gamma, int32_t* out_host, int* num_buckets);

/* ------------------------------------------------------------------------- *
 * Weight shadows (fp32 master -> bf16), refreshed once per optimizer step.
 * ------------------------------------------------------------------------- */
/* Plain cast, same shape: dst[r, c] = bf16(src[r, c]); ld in elements. */
int cream_shadow_cast(const float* src, void* dst_bf16, int64_t rows, int64_t cols, int64_t ld_src,
                      int64_t ld_dst, void* stream);
/* QKV de-interleave: reference row 3*j+i (i = 0:q,1:k,2:v; qkv_super.py:72-77) goes to
 * shadow row i*rows_per_group + j, so a sampled (E, heads) slice is three dense
 * row blocks.  src is (3*rows_per_group, cols). */
int cream_shadow_qkv(const float* src, void* dst_bf16, int64_t rows_per_group, int64_t cols,
                     int64_t ld_src, int64_t ld_dst, void* stream);

/* ------------------------------------------------------------------------- *
 * Sliced GEMM on tcgen05 tensor cores (TMA-fed, accumulators in TMEM).
 *   acc[g][m, n] = sum_k A[m, k] * B[g][n, k]       (bf16 x bf16 -> fp32)
 * The operand SLICE (sampled embed dim / heads / mlp ratio) is expressed by
 * M, N, K and the leading dimensions of the FULL supernet tensors.
 * ------------------------------------------------------------------------- */
#define CREAM_EPI_BF16 0        /* out_bf16 = acc + bias                              */
#define CREAM_EPI_BF16_GELU 1   /* aux_bf16 = acc + bias ; out_bf16 = gelu(aux)       */
#define CREAM_EPI_F32_RESID 2   /* out_f32 = resid + row_scale * (acc + bias)         */
#define CREAM_EPI_BF16_DGELU 3  /* out_bf16 = acc * gelu'(aux_bf16)                   */
#define CREAM_EPI_F32_ATOMIC 4  /* out_f32 += alpha * acc   (split-K weight gradient)  */
#define CREAM_EPI_F32 5         /* out_f32 = acc + bias                               */

typedef struct cream_gemm_desc {
  int M, N, K;        /* per-group extents; K = full contraction length            */
  int groups;         /* >= 1; QKV uses 3 (q,k,v blocks of the de-interleaved shadow) */
  /* A operand, bf16. a_mn = 0: A is (M, K) row-major, lda = row pitch.
   *                  a_mn = 1: A is stored (K, groups*a_group_off.. ) i.e. K rows of
   *                  M-contiguous data (used for dY^T in weight gradients);
   *                  a_group_off = column offset between groups. */
  const void* a; int64_t lda; int a_mn; int a_group_off;
  /* B operand, bf16. b_mn = 0: B[g] is (N, K) row-major at row g*b_group_rows.
   *                  b_mn = 1: B is stored (K, N) row-major (N contiguous); the
   *                  contraction may run over `k_groups` row blocks of length
   *                  k_group_len (multiple of 64) spaced b_group_rows apart. */
  const void* b; int64_t ldb; int b_mn; int64_t b_group_rows;
  int k_groups; int k_group_len;
  /* epilogue */
  int epi;
  void* out; int64_t ldo;      /* out[(m*out_row_mul + g*out_g_row) * ldo + n + g*out_g_col] */
  int out_row_mul, out_g_row, out_g_col;
  void* aux; int64_t ldaux;    /* GELU pre-activation, bf16, indexed like out          */
  const float* bias;           /* indexed by (n + g*out_g_col), may be NULL            */
  const float* resid; int64_t ldr;
  const float* row_scale; int rows_per_scale; /* DropPath per-sample scale, may be NULL */
  float alpha;
  int split_k;                 /* 0 = choose automatically (only for EPI_F32_ATOMIC)   */
} cream_gemm_desc;

int cream_gemm_bf16(const cream_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CREAM_B200_H_ */
