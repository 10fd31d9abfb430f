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

/* Bind the calling host thread to `device` inside the library's CUDA runtime (call once per
 * thread / device change before any other entry point; cheap and idempotent). */
int cream_bind_device(int device);

/* Version of the rpe_index operator contract; the reference asserts "1.2.0"
 * (iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.py:5-8, rpe_index.cpp:126-128). */
const char* cream_rpe_index_version(void);

/* ------------------------------------------------------------------------- *
 * rpe_index — replaces rpe_index_cpp.forward_gpu / backward_gpu
 * (iRPE/DeiT-with-iRPE/rpe_ops/rpe_index_cuda.cu:54-94 and :96-140).
 *   fwd: Y[b,h,i,j] = input[b,h,i,index[i,j]]       input strides s0..s3 (elements),
 *        index (Lq,Lk) int32 contiguous, Y (B,H,Lq,Lk) contiguous.
 *   bwd: grad_input[b,h,i,index[i,j]] += grad_output[b,h,i,j]; grad_input is
 *        (B,H,Lq,nb) contiguous and caller-zeroed (rpe_index.py:51).  Per-row
 *        shared-memory histogram, one coalesced read-modify-write per row (the
 *        reference issues one global atomic per element).
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
  int cta_pair;                /* 0 = choose automatically; 1 = one CTA per 128-row tile;
                                * 2 = CTA pairs (tcgen05 cta_group::2) on 256-row tiles    */
} cream_gemm_desc;

int cream_gemm_bf16(const cream_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused attention with relative-position terms (tcgen05 QK^T / PV, RPE bucket gather
 * fused into the softmax).  Replaces AttentionSuper.forward's attention core
 * (AutoFormer/model/module/multihead_super.py:135-154) and RPEAttention.forward's
 * (iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:73-92, irpe.py:585-687).
 *
 *   S[i,j] = scale * (q_i . k_j + sum_t q_i . TK_t[idx_t[i,j]]) + bias[idx_a[i,j]]
 *   P = softmax_j(S);  O[i] = sum_j P[i,j] * (v_j + sum_t TV_t[idx_vt[i,j]])
 *
 * qkv : bf16 (B*N, ld_qkv), columns [q: H*64 | k: H*64 | v: H*64] — the layout
 *       qkv(x).reshape(B, N, 3, H, 64) of the reference.
 * tk_pack / tv_pack : bf16 (T, 64, 64) "packed tables": row = bucket id (two tables of
 *       <= 32 buckets at rows [0,32) and [32,64), or one table of <= 64), col = head-dim
 *       channel; T = H if tables_per_head else 1.  NULL = no contextual term.
 * idx_* : uint8 (N, ld_idx) bucket ids ALREADY offset into the packed rows (< 64);
 *       ld_idx % 16 == 0, ld_idx >= roundup(N,16), padding entries 0.  idx_b / idx_vb
 *       are the second table's ids (AutoFormer horizontal table) or NULL.
 * bias_pack : fp32 (T, 64) bias-mode table gathered with idx_a, or NULL.
 * out : bf16 (B*N, ld_out) columns H*64;  lse: fp32 (B, H, N) log-sum-exp, or NULL.
 * ------------------------------------------------------------------------- */
typedef struct cream_attn_desc {
  int B, H, N, head_dim;
  float scale;
  const void* qkv; int64_t ld_qkv;
  void* out; int64_t ld_out;
  float* lse;
  const void* tk_pack; const void* tv_pack; int tables_per_head;
  const uint8_t* idx_a; const uint8_t* idx_b; const uint8_t* idx_va; const uint8_t* idx_vb;
  int ld_idx;
  const float* bias_pack;
  /* AutoFormer structure hint: the idx tables are those of a af_grid x af_grid patch grid + cls
   * with clamp af_max_rel (cream_autoformer_rel_index_host), v ids at packed rows [0,32) and
   * h ids at [32,64).  Lets the kernel replace the table gather by register arithmetic.
   * 0 = no hint (generic gather through idx_*). */
  int af_grid, af_max_rel;
  /* ---- backward only (cream_attn_bwd) ---- */
  const void* dout; int64_t ld_dout;   /* bf16 (B*N, ld_dout)                              */
  void* dqkv; int64_t ld_dqkv;         /* bf16 (B*N, ld_dqkv), same column layout as qkv    */
  float* dtk_pack; float* dtv_pack;    /* fp32 (T, 64, 64) accumulated (+=), caller-zeroed  */
  float* dbias_pack;                   /* fp32 (T, 64) accumulated, or NULL                 */
  void* workspace; int64_t workspace_bytes; /* see cream_attn_bwd_workspace_bytes          */
  /* ---- optional dense additive logit term (generic gather path only) ----
   * S[b,h,i,j] += dense_bias[b*stride_b + h*stride_h + i*stride_i + j]  (fp32, j contiguous; a
   * stride of 0 broadcasts).  Carries what the reference adds to the logits outside the bucket
   * tables: iRPE on queries, rpe_q(k*scale)^T (rpe_vision_transformer.py:82-83), TinyViT's per-head
   * attention_biases[:, idxs] (tiny_vit.py:281-283), an additive -inf mask (TinyCLIP text tower,
   * open_clip/model.py:756-762).  Backward: ddense (fp32, (B,H,N,N) contiguous, may be NULL)
   * receives dS = P * (dP - delta). */
  const float* dense_bias; int64_t dense_stride_b, dense_stride_h, dense_stride_i;
  float* ddense;
  /* ---- optional grid-product structure hint (iRPE product method on a g x g grid + cls, irpe.py:176-202):
   * for patch tokens i = (ri, ci), j = (rj, cj) the K-side bucket id is
   *     idx_a[i, j] = gp_lut_a[rj - ri + g - 1] * gp_w + gp_lut_b[cj - ci + g - 1]
   * and every pair that involves the cls token uses bucket gp_skip_id (the `skip` bucket, irpe.py:364-415).
   * Lets the kernel replace the index-table gather by per-thread registers, and the per-element
   * bucket-sum read-modify-writes of the backward by rectangle sums.  Used only when gp_grid > 0, the K
   * table is the single table of the pack (idx_b == NULL) and there is no V-side table, bias or dense term;
   * idx_a must still be the matching table (it is what the caller verified the structure against). */
  int gp_grid, gp_w, gp_skip_id;
  uint8_t gp_lut_a[32], gp_lut_b[32];
  /* 1: key j > query i is masked out (logit -inf) - the text tower's causal mask of
   * open_clip/model.py:756-762 computed from the coordinates instead of read from a dense (N, N) term.
   * Generic gather path only (no af / gp hint). */
  int causal;
  /* > 0: the N tokens are `N / block_len` independent items of block_len tokens laid end to end (block-diagonal
   * attention: key j is visible to query i only when i / block_len == j / block_len; `causal` then applies inside a
   * block).  Lets a caller whose sequences are short (TinyCLIP's 50-token image tower, open_clip/model.py:493-536)
   * hand TWO batch items to one 128-row tile: (B, N) tokens are passed as (B / 2, 2 N) with block_len = N, no copy.
   * N must be a multiple of block_len.  Generic gather path only. */
  int block_len;
} cream_attn_desc;

int cream_attn_fwd(const cream_attn_desc* desc, void* stream);
int cream_attn_bwd(const cream_attn_desc* desc, void* stream);
int64_t cream_attn_bwd_workspace_bytes(int B, int H, int N);

/* ------------------------------------------------------------------------- *
 * Sliced LayerNorm (AutoFormer/model/module/layernorm_super.py:26-37; fp32 statistics).
 *   fwd: out = LN(x[:, :E]) * gamma[:E] + beta[:E]; out bf16 (A operand of the next GEMM)
 *        or fp32 (out_f32 != 0); mean / rstd (rows) saved for backward (may be NULL).
 *   bwd: dx = resid_grad + dLN/dx (resid_grad may be NULL); dgamma/dbeta accumulated (+=).
 * ------------------------------------------------------------------------- */
int cream_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                        void* out, int64_t ldo, int out_f32, float* mean, float* rstd, int64_t rows,
                        int E, void* stream);
int cream_layernorm_bwd(const void* dy, int64_t lddy, int dy_f32, const float* x, int64_t ldx,
                        const float* gamma, const float* mean, const float* rstd,
                        const float* resid_grad, int64_t ldrg, float* dx, int64_t lddx, float* dgamma,
                        float* dbeta, int64_t rows, int E, void* stream);
/* cream_layernorm_bwd that ALSO emits what the next two GEMMs of the backward consume - out_bf16 =
 * bf16(row_scale[r / rows_per_scale] * dx) (the DropPath factor of the branch the gradient enters,
 * model/utils.py:71-99; NULL = 1) - and that branch's bias gradient dbias[c] += sum_r out_bf16[r, c]
 * (NULL = skip): the separate cream_cast_scale pass over dx disappears.  Same result as
 * cream_layernorm_bwd followed by cream_cast_scale (which is what runs for shapes the pipelined
 * kernel does not take). */
int cream_layernorm_bwd_cast(const void* dy, int64_t lddy, int dy_f32, const float* x, int64_t ldx,
                             const float* gamma, const float* mean, const float* rstd,
                             const float* resid_grad, int64_t ldrg, float* dx, int64_t lddx, float* dgamma,
                             float* dbeta, int64_t rows, int E, void* out_bf16, int64_t ldo,
                             const float* row_scale, int rows_per_scale, float* dbias, void* stream);

/* ------------------------------------------------------------------------- *
 * Patch embedding / token assembly / pooling glue.
 * ------------------------------------------------------------------------- */
/* images (B,C,H,W) fp32 -> (B*(H/P)*(W/P), C*P*P) bf16 patches; the 16x16/16 conv of
 * PatchembedSuper.forward (embedding_super.py:33-40) is then one sliced GEMM against
 * proj.weight.view(E*, C*P*P)[:E]. */
int cream_patch_im2col(const float* images, void* out_bf16, int64_t ldo, int B, int C, int H, int W,
                       int P, void* stream);
/* x[b,0] = cls[:E] + pos[0,:E]; x[b,1+t] = patch[b*T+t] + pos[1+t,:E]
 * (supernet_transformer.py:150-155); pos may be NULL (abs_pos = False). */
int cream_tokens_assemble_fwd(const void* patch_bf16, int64_t ldp, const float* cls, const float* pos,
                              int64_t ldpos, float* x, int64_t ldx, int B, int N, int E, void* stream);
int cream_tokens_assemble_bwd(const float* g, int64_t ldg, void* dpatch_bf16, int64_t ldp, float* dpos,
                              int64_t ldpos, float* dcls, int B, int N, int E, void* stream);
/* mean over tokens [first, first+count) (gp pooling, supernet_transformer.py:164-165). */
int cream_pool_fwd(const float* y, int64_t ldy, void* out_bf16, int64_t ldo, int B, int N, int E,
                   int first, int count, void* stream);
int cream_pool_bwd(const void* dpooled_bf16, int64_t lddp, float* dy, int64_t lddy, int B, int N, int E,
                   int first, int count, void* stream);
/* out_bf16 = bf16(row_scale[r / rows_per_scale] * in); dbias (+=) column sums of out, or NULL. */
int cream_cast_scale(const float* in, int64_t ldi, void* out_bf16, int64_t ldo, const float* row_scale,
                     int rows_per_scale, float* dbias, int64_t rows, int cols, void* stream);
/* dbias[c] += sum_r dy[r, c]. */
int cream_bias_grad(const void* dy_bf16, int64_t ld, float* dbias, int64_t rows, int cols, void* stream);

/* ------------------------------------------------------------------------- *
 * Relative-position table packs: (T, 64, 64) bf16, row = packed bucket id.
 * Up to two source tables src[t][b][d] with arbitrary element strides
 * (AutoFormer embeddings_table_v/h (30, 64); iRPE lookup_table_weight (H|1, 64, nb)
 * transposed or (H|1, nb, 64)).  unpack adds the packed gradient back (+=).
 * ------------------------------------------------------------------------- */
int cream_pack_tables(void* dst_bf16, int num_tables, int head_dim, const float* src0, int nb0,
                      int row_off0, int64_t stride_t0, int64_t stride_b0, int64_t stride_d0,
                      const float* src1, int nb1, int row_off1, int64_t stride_t1, int64_t stride_b1,
                      int64_t stride_d1, void* stream);
int cream_unpack_table_grads(const float* dpack, int num_tables, int head_dim, float* grad0, int nb0,
                             int row_off0, int64_t stride_t0, int64_t stride_b0, int64_t stride_d0,
                             float* grad1, int nb1, int row_off1, int64_t stride_t1, int64_t stride_b1,
                             int64_t stride_d1, void* stream);

/* Batched variants: one launch for all (layer, k|v) packs of a step (<= 64).  src*_host /
 * grad*_host are HOST arrays of n_packs DEVICE pointers to (nb, head_dim) tables with element
 * strides (stride_b, stride_d); table 0 goes to rows [0,nb), table 1 (may be NULL) to
 * rows [row_off1, row_off1+nb).  dst / dpack are (n_packs, 64, 64). */
int cream_pack_tables_batch(void* dst_bf16, int n_packs, int head_dim, const float* const* src0_host,
                            const float* const* src1_host, int nb, int row_off1, int64_t stride_b,
                            int64_t stride_d, void* stream);
int cream_unpack_table_grads_batch(const float* dpack, int n_packs, int head_dim, float* const* grad0_host,
                                   float* const* grad1_host, int nb, int row_off1, int64_t stride_b,
                                   int64_t stride_d, void* stream);

/* ------------------------------------------------------------------------- *
 * Whole sampled-subnet forward / backward in ONE call each (the native runtime under
 * Vision_TransformerSuper.forward, AutoFormer/model/supernet_transformer.py:147-172, 251-287, and
 * under the DeiT + iRPE VisionTransformer, iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:107-201):
 * patch-embed GEMM, token assembly, per block LN -> QKV GEMM -> fused attention + RPE -> proj GEMM
 * (+bias, DropPath, residual) -> LN -> fc1 GEMM (+bias, GELU) -> fc2 GEMM (+bias, DropPath,
 * residual), final LN, pooling, head GEMM - about 11 kernel launches per block enqueued from C++
 * with no allocation: every activation lives at a fixed offset of a caller-provided arena
 * (cream_vit_arena_bytes), so the TMA descriptors are cache hits from the second step on.
 *
 * All parameter pointers are the reference's FULL supernet tensors (fp32 masters, bf16 shadows);
 * the sampled slice is (E, heads, ffn) per layer.  Gradients are accumulated (+=) into full-size
 * fp32 tensors that the caller zeroed; layers >= depth are not touched (grad stays "None").
 * ------------------------------------------------------------------------- */
#define CREAM_VIT_MAX_DEPTH 32

typedef struct cream_vit_layer {
  int heads, ffn;                                     /* sampled head count / hidden width   */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;         /* attn_layer_norm / ffn_layer_norm    */
  const void* wqkv; const float* bqkv;                /* bf16 shadow (see qkv_interleaved)    */
  const void* wproj; const float* bproj;
  const void* wfc1; const float* bfc1;
  const void* wfc2; const float* bfc2;
  /* relative-position tables of this layer, fp32 masters: K side [0], [1]; V side [2], [3]
   * (AutoFormer: embeddings_table_v / _h of rel_pos_embed_k / _v; iRPE: lookup_table_weight in
   * [0] or [2], the second of a pair only for the cross method).  NULL = absent. */
  const float* tab[4];
  float* g_tab[4];
  const float* dp_scale;                              /* (2, B) DropPath factors or NULL     */
  float *g_ln1_g, *g_ln1_b, *g_ln2_g, *g_ln2_b;
  float *g_wqkv, *g_bqkv, *g_wproj, *g_bproj, *g_wfc1, *g_bfc1, *g_wfc2, *g_bfc2;
} cream_vit_layer;

typedef struct cream_vit_desc {
  int B, N, E, depth;                /* batch, tokens (1 + grid^2), sampled embed dim, layers */
  int num_classes, in_chans, img_size, patch_size;
  float eps;                         /* LayerNorm epsilon                                     */
  int pool_first, pool_count;        /* gp: (1, N-1); cls token: (0, 1)                       */
  int qkv_interleaved;               /* 1: AutoFormer qkv_super rows (de-interleaved shadow, rows_per_group =
                                      * qkv_group_rows); 0: plain [q; k; v] row blocks (DeiT)  */
  int qkv_group_rows;
  int64_t ld_wqkv, ld_wproj, ld_wfc1, ld_wfc2, ld_wpatch, ld_whead;   /* shadow row pitches    */
  int64_t ld_gqkv, ld_gproj, ld_gfc1, ld_gfc2, ld_gpatch, ld_ghead;   /* fp32 gradient row pitches
                                      * (= columns of the full master weights)                 */
  /* attention */
  float scale;
  int af_grid, af_max_rel;           /* AutoFormer structure hint (see cream_attn_desc)        */
  const uint8_t *idx_a, *idx_b, *idx_va, *idx_vb; int ld_idx;
  int gp_grid, gp_w, gp_skip_id;     /* grid-product structure hint (see cream_attn_desc)      */
  uint8_t gp_lut_a[32], gp_lut_b[32];
  int tab_nb, tab_row_off1;          /* buckets per table; packed row offset of the 2nd table  */
  int64_t tab_stride_b, tab_stride_d;/* element strides of a table: (bucket, channel)          */
  int64_t tabv_stride_b, tabv_stride_d; /* same for the V-side tables                          */
  /* embedding / head parameters */
  const float* images;               /* (B, C, img, img) fp32                                  */
  const void* wpatch; const float* bpatch;
  const float* cls; const float* pos; int64_t ld_pos;   /* pos may be NULL                     */
  const float *norm_g, *norm_b;
  const void* whead; const float* bhead;
  float* logits; int64_t ld_logits;  /* out: (B, num_classes) fp32                             */
  /* backward */
  const float* dlogits; int64_t ld_dlogits;   /* (B, num_classes) fp32                         */
  float *g_wpatch, *g_bpatch, *g_cls, *g_pos, *g_norm_g, *g_norm_b, *g_whead, *g_bhead;
  /* activations + scratch */
  void* arena; int64_t arena_bytes;
  cream_vit_layer layers[CREAM_VIT_MAX_DEPTH];
} cream_vit_desc;

/* Bytes of arena needed for `desc` (uses B, N, E, depth, num_classes, geometry and the per-layer
 * heads / ffn); size it once for the largest subnet of the search space. */
int64_t cream_vit_arena_bytes(const cream_vit_desc* desc);
int cream_vit_fwd(const cream_vit_desc* desc, void* stream);
/* Backward stages run in order head (0), layers depth-1 .. 0 (stages 1 .. depth), embedding
 * (stage depth+1).  One call may run any contiguous range [first_stage, last_stage], so a caller
 * can start the gradient all-reduce of a layer as soon as its stage has been enqueued. */
int cream_vit_bwd(const cream_vit_desc* desc, int first_stage, int last_stage, void* stream);
/* Kernels enqueued by the last cream_vit_fwd / cream_vit_bwd call of this thread. */
int cream_vit_last_launches(void);

/* Mean cross-entropy over the batch and its gradient in one kernel:
 * loss[0] = mean_b(logsumexp(logits[b]) - logits[b, target[b]]); dlogits = (softmax - onehot) / B. */
int cream_xent_fwd_bwd(const float* logits, int64_t ld, const int64_t* targets, float* loss,
                       float* dlogits, int64_t ldd, int B, int C, void* stream);

/* ------------------------------------------------------------------------- *
 * AdamW over a flat list of parameter segments, fused with the refresh of the bf16 weight shadows
 * (torch.optim.AdamW semantics: decoupled weight decay, bias correction from a PER-PARAMETER step
 * count, parameters without a gradient this step - identity layers - are skipped entirely).
 * `segs` is a DEVICE array of n_segs descriptors; `active` a device int array (1 = has a gradient
 * this step, the step counter is advanced by the kernel).
 * ------------------------------------------------------------------------- */
typedef struct cream_adamw_seg {
  float* p; const float* g; float* m; float* v;   /* fp32, `numel` (< 2^31) contiguous elements, 16-byte aligned */
  void* shadow;                                   /* bf16 shadow or NULL                        */
  int64_t numel;
  int32_t rows, cols;                             /* 2-D view (rows, cols) of p for the shadow  */
  int32_t shadow_ld;                              /* shadow row pitch (elements)                */
  int32_t qkv_group_rows;                         /* > 0: de-interleave rows (cream_shadow_qkv) */
  float weight_decay;
  int32_t step;                                   /* updated in place by the kernel             */
} cream_adamw_seg;
/* `blocks_dev`: n_blocks (segment, first element) int32 pairs, one per cream_adamw_chunk()-element chunk
 * of every segment, in any order (built once by the caller). */
int cream_adamw_step(cream_adamw_seg* segs_dev, const int32_t* active_dev, int n_segs, const int32_t* blocks_dev,
                     int n_blocks, float lr, float beta1, float beta2, float eps, void* stream);
int cream_adamw_chunk(void);

#ifdef __cplusplus
}
#endif
#endif /* CREAM_B200_H_ */
