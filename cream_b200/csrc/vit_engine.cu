// Native runtime of the sampled-subnet ViT: one C call per forward, one per (range of) backward
// stages.  Host code only - it sequences the library's own kernels (gemm_sm100.cu, attention_*.cu,
// layernorm.cu, elementwise.cu) over a caller-provided arena, so a training step costs a handful of
// C calls instead of ~320 Python -> ctypes launches with a torch allocation each.
//
// Mirrors Vision_TransformerSuper.forward after set_sample_config
// (AutoFormer/model/supernet_transformer.py:102-127, 147-172), TransformerEncoderLayer.forward
// (:251-287), AttentionSuper.forward (model/module/multihead_super.py:133-160) and their autograd
// backward; with qkv_interleaved = 0 and iRPE tables it is the DeiT + iRPE VisionTransformer
// (iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:107-201, RPEBlock :100-104).
#include <cstring>

#include <cstdlib>
#include "common.cuh"

namespace cb {
namespace {

thread_local int t_launches = 0;

#define VIT_TRY(expr)            \
  do {                           \
    const int _rc = (expr);      \
    if (_rc != CREAM_OK) return _rc; \
  } while (0)

inline int64_t up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------
// arena layout: every buffer at a fixed 256-byte aligned offset, a pure function of the descriptor
// ---------------------------------------------------------------------------------------------
struct Bf16Mat { __nv_bfloat16* p; int64_t ld; };
struct F32Mat { float* p; int64_t ld; };

struct LayerBufs {
  F32Mat x, x1;                    // residual stream entering the block / after the attention branch
  Bf16Mat ln1, ln2, qkv, att, hpre, act;
  float *mu1, *rs1, *mu2, *rs2, *lse;
};

struct Bufs {
  Bf16Mat cols, patch, pooled;
  F32Mat x_last, y;
  float *mu_f, *rs_f;
  __nv_bfloat16 *tk_packs, *tv_packs;   // (depth, 64, 64) bf16 each, or nullptr
  float *dtk_packs, *dtv_packs;         // (depth, 64, 64) fp32
  LayerBufs L[CREAM_VIT_MAX_DEPTH];
  // backward scratch (reused by every layer)
  Bf16Mat dl, dpooled, dy_bf, dh, dln, datt, dqkv, dpatch;
  F32Mat ga, gb, dy;
  void* attn_ws; int64_t attn_ws_bytes;
  int64_t total;
};

struct Bump {
  char* base; int64_t off;
  void* take(int64_t bytes) {
    void* p = base ? base + off : nullptr;
    off += up(bytes, 256);
    return p;
  }
  Bf16Mat bf16(int64_t rows, int64_t cols) {
    const int64_t ld = up(std::max<int64_t>(cols, 1), 8);
    return Bf16Mat{static_cast<__nv_bfloat16*>(take(rows * ld * 2)), ld};
  }
  F32Mat f32(int64_t rows, int64_t cols) {
    const int64_t ld = up(std::max<int64_t>(cols, 1), 4);
    return F32Mat{static_cast<float*>(take(rows * ld * 4)), ld};
  }
  float* vec(int64_t n) { return static_cast<float*>(take(n * 4)); }
};

bool has_k_tables(const cream_vit_desc& d) { return d.depth > 0 && d.layers[0].tab[0] != nullptr; }
bool has_v_tables(const cream_vit_desc& d) { return d.depth > 0 && d.layers[0].tab[2] != nullptr; }

void plan(const cream_vit_desc& d, Bufs& b) {
  Bump a{static_cast<char*>(d.arena), 0};
  const int64_t T = d.N - 1, M = static_cast<int64_t>(d.B) * d.N;
  const int64_t kdim = static_cast<int64_t>(d.in_chans) * d.patch_size * d.patch_size;
  b.cols = a.bf16(d.B * T, kdim);
  b.patch = a.bf16(d.B * T, d.E);
  b.tk_packs = has_k_tables(d) ? static_cast<__nv_bfloat16*>(a.take(static_cast<int64_t>(d.depth) * 64 * 64 * 2)) : nullptr;
  b.tv_packs = has_v_tables(d) ? static_cast<__nv_bfloat16*>(a.take(static_cast<int64_t>(d.depth) * 64 * 64 * 2)) : nullptr;
  b.dtk_packs = has_k_tables(d) ? a.vec(static_cast<int64_t>(d.depth) * 64 * 64) : nullptr;
  b.dtv_packs = has_v_tables(d) ? a.vec(static_cast<int64_t>(d.depth) * 64 * 64) : nullptr;
  F32Mat x = a.f32(M, d.E);
  int max_h = 1, max_ffn = 1;
  for (int i = 0; i < d.depth; ++i) {
    LayerBufs& l = b.L[i];
    const int qd = 64 * d.layers[i].heads, ffn = d.layers[i].ffn;
    max_h = std::max(max_h, d.layers[i].heads);
    max_ffn = std::max(max_ffn, ffn);
    l.x = x;
    l.ln1 = a.bf16(M, d.E);
    l.mu1 = a.vec(M); l.rs1 = a.vec(M);
    l.qkv = a.bf16(M, 3 * qd);
    l.att = a.bf16(M, qd);
    l.lse = a.vec(static_cast<int64_t>(d.B) * d.layers[i].heads * d.N);
    l.x1 = a.f32(M, d.E);
    l.ln2 = a.bf16(M, d.E);
    l.mu2 = a.vec(M); l.rs2 = a.vec(M);
    l.hpre = a.bf16(M, ffn);
    l.act = a.bf16(M, ffn);
    x = a.f32(M, d.E);
  }
  b.x_last = x;
  b.y = a.f32(M, d.E);
  b.mu_f = a.vec(M); b.rs_f = a.vec(M);
  b.pooled = a.bf16(d.B, d.E);
  // backward scratch
  b.dl = a.bf16(d.B, d.num_classes);
  b.dpooled = a.bf16(d.B, d.E);
  b.dy = a.f32(M, d.E);
  b.ga = a.f32(M, d.E);
  b.gb = a.f32(M, d.E);
  b.dy_bf = a.bf16(M, d.E);
  b.dh = a.bf16(M, max_ffn);
  b.dln = a.bf16(M, d.E);
  b.datt = a.bf16(M, 64 * max_h);
  b.dqkv = a.bf16(M, 3 * 64 * max_h);
  b.dpatch = a.bf16(d.B * T, d.E);
  b.attn_ws_bytes = cream_attn_bwd_workspace_bytes(d.B, max_h, d.N);
  b.attn_ws = a.take(b.attn_ws_bytes);
  b.total = a.off;
}

int validate(const cream_vit_desc* d) {
  CB_REQUIRE(d != nullptr, "desc is null");
  CB_REQUIRE(d->B >= 1 && d->N >= 2 && d->E >= 4 && d->E % 4 == 0, "bad geometry");
  CB_REQUIRE(d->depth >= 1 && d->depth <= CREAM_VIT_MAX_DEPTH, "1 <= depth <= CREAM_VIT_MAX_DEPTH");
  CB_REQUIRE(d->num_classes >= 4 && d->num_classes % 4 == 0, "num_classes must be a multiple of 4");
  CB_REQUIRE(d->patch_size >= 1 && d->img_size % d->patch_size == 0, "bad geometry");
  const int g = d->img_size / d->patch_size;
  CB_REQUIRE(g * g + 1 == d->N, "tokens must be 1 + (img_size / patch_size)^2");
  for (int i = 0; i < d->depth; ++i)
    CB_REQUIRE(d->layers[i].heads >= 1 && d->layers[i].ffn >= 8 && d->layers[i].ffn % 4 == 0, "bad layer slice");
  return CREAM_OK;
}

// ---------------------------------------------------------------------------------------------
// GEMM helpers (the slice is M, N, K plus the FULL tensors' pitches; see gemm_sm100.cu)
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  cream_gemm_desc g;
  GemmArgs() { std::memset(&g, 0, sizeof(g)); g.groups = 1; g.alpha = 1.0f; g.rows_per_scale = 1; }
};

int run_gemm(const cream_gemm_desc& g, cudaStream_t s) {
  ++t_launches;
  return cream_gemm_bf16(&g, s);
}

// y = x[:, :k] W[:n, :k]^T (+ bias) with epilogue `epi`
int linear_fwd(cudaStream_t s, int64_t M, int n, int k, Bf16Mat x, const void* w, int64_t ldw, const float* bias,
               int epi, void* out, int64_t ldo, void* aux = nullptr, int64_t ldaux = 0, const float* resid = nullptr,
               int64_t ldr = 0, const float* row_scale = nullptr, int rows_per_scale = 1) {
  GemmArgs a;
  a.g.M = static_cast<int>(M); a.g.N = n; a.g.K = k;
  a.g.a = x.p; a.g.lda = x.ld;
  a.g.b = w; a.g.ldb = ldw;
  a.g.epi = epi; a.g.out = out; a.g.ldo = ldo;
  a.g.aux = aux; a.g.ldaux = ldaux;
  a.g.bias = bias; a.g.resid = resid; a.g.ldr = ldr;
  a.g.row_scale = row_scale; a.g.rows_per_scale = rows_per_scale;
  return run_gemm(a.g, s);
}

// dx = dy[:, :n] W[:n, :k]   (B operand MN-major straight from the shadow)
int linear_dgrad(cudaStream_t s, int64_t M, int n, int k, Bf16Mat dy, const void* w, int64_t ldw, int epi,
                 Bf16Mat out, void* aux = nullptr, int64_t ldaux = 0) {
  GemmArgs a;
  a.g.M = static_cast<int>(M); a.g.N = k; a.g.K = n;
  a.g.a = dy.p; a.g.lda = dy.ld;
  a.g.b = w; a.g.ldb = ldw; a.g.b_mn = 1;
  a.g.epi = epi; a.g.out = out.p; a.g.ldo = out.ld;
  a.g.aux = aux; a.g.ldaux = ldaux;
  return run_gemm(a.g, s);
}

// dW[:n, :k] += dy^T x   (full-size fp32 gradient, row pitch ldg)
int linear_wgrad(cudaStream_t s, int64_t M, int n, int k, Bf16Mat dy, Bf16Mat x, float* dw, int64_t ldg) {
  GemmArgs a;
  a.g.M = n; a.g.N = k; a.g.K = static_cast<int>(M);
  a.g.a = dy.p; a.g.lda = dy.ld; a.g.a_mn = 1;
  a.g.b = x.p; a.g.ldb = x.ld; a.g.b_mn = 1;
  a.g.epi = CREAM_EPI_F32_ATOMIC; a.g.out = dw; a.g.ldo = ldg;
  return run_gemm(a.g, s);
}

int bias_grad(cudaStream_t s, Bf16Mat dy, int64_t rows, int cols, float* dbias) {
  ++t_launches;
  return cream_bias_grad(dy.p, dy.ld, dbias, rows, cols, s);
}

// The per-layer bias gradients of fc1 and qkv are leaves: nothing downstream in the backward reads them.  Their
// column-sum kernels (256 threads, 9 KB of shared memory) fit on an SM beside a resident GEMM CTA, and the GEMMs
// leave most of the HBM bandwidth unused, so they run on a side stream UNDER the weight / data gradient GEMMs that
// follow instead of in front of them.  fork: side waits for the producer on `s`; join (once per stage, before the
// buffers are rewritten): `s` waits for the side stream.  CREAM_SIDE_BIAS=0 keeps everything on `s`.
struct SideStream {
  cudaStream_t st = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool pending = false;
};
thread_local SideStream t_side[16];

bool side_enabled() {
  static const bool on = []() { const char* e = getenv("CREAM_SIDE_BIAS"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

int side_bias_grad(cudaStream_t s, Bf16Mat dy, int64_t rows, int cols, float* dbias) {
  int dev = 0;
  if (!side_enabled() || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return bias_grad(s, dy, rows, cols, dbias);
  SideStream& ss = t_side[dev];
  if (ss.st == nullptr) {
    CB_CUDA_OK(cudaStreamCreateWithFlags(&ss.st, cudaStreamNonBlocking));
    CB_CUDA_OK(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    CB_CUDA_OK(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
  }
  CB_CUDA_OK(cudaEventRecord(ss.fork, s));
  CB_CUDA_OK(cudaStreamWaitEvent(ss.st, ss.fork, 0));
  ss.pending = true;
  return bias_grad(ss.st, dy, rows, cols, dbias);
}

// The weight-gradient GEMMs are leaves too.  CREAM_SIDE_WGRAD=1 issues them on the side stream as well, so that
// their CTAs can take the SMs a main-chain kernel leaves idle in its last wave (both are one-CTA-per-SM kernels: the
// work is conserved, only the tails are filled).  Off by default; see DESIGN.md section 6 for the measurement.
bool side_wgrad_enabled() {
  static const bool on = []() { const char* e = getenv("CREAM_SIDE_WGRAD"); return e != nullptr && e[0] == '1'; }();
  return on;
}

cudaStream_t side_fork(cudaStream_t s) {     // the side stream, ordered after everything enqueued on `s` so far
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return s;
  SideStream& ss = t_side[dev];
  if (ss.st == nullptr) {
    if (cudaStreamCreateWithFlags(&ss.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess)
      return s;
  }
  if (cudaEventRecord(ss.fork, s) != cudaSuccess || cudaStreamWaitEvent(ss.st, ss.fork, 0) != cudaSuccess) return s;
  ss.pending = true;
  return ss.st;
}

int side_join(cudaStream_t s) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return CREAM_OK;
  SideStream& ss = t_side[dev];
  if (!ss.pending) return CREAM_OK;
  CB_CUDA_OK(cudaEventRecord(ss.join, ss.st));
  CB_CUDA_OK(cudaStreamWaitEvent(s, ss.join, 0));
  ss.pending = false;
  return CREAM_OK;
}

int pack_layer_tables(const cream_vit_desc& d, const Bufs& b, cudaStream_t s) {
  // K-side and V-side packs of every layer, one launch per side
  for (int side = 0; side < 2; ++side) {
    __nv_bfloat16* dst = side ? b.tv_packs : b.tk_packs;
    if (dst == nullptr) continue;
    const float* src0[CREAM_VIT_MAX_DEPTH];
    const float* src1[CREAM_VIT_MAX_DEPTH];
    bool second = false;
    for (int i = 0; i < d.depth; ++i) {
      src0[i] = d.layers[i].tab[2 * side];
      src1[i] = d.layers[i].tab[2 * side + 1];
      second = second || src1[i] != nullptr;
    }
    ++t_launches;
    VIT_TRY(cream_pack_tables_batch(dst, d.depth, 64, src0, second ? src1 : nullptr, d.tab_nb, d.tab_row_off1,
                                    side ? d.tabv_stride_b : d.tab_stride_b, side ? d.tabv_stride_d : d.tab_stride_d, s));
  }
  return CREAM_OK;
}

void fill_attn(const cream_vit_desc& d, const Bufs& b, int i, cream_attn_desc& a) {
  std::memset(&a, 0, sizeof(a));
  const LayerBufs& l = b.L[i];
  a.B = d.B; a.H = d.layers[i].heads; a.N = d.N; a.head_dim = 64;
  a.scale = d.scale;
  a.qkv = l.qkv.p; a.ld_qkv = l.qkv.ld;
  a.out = l.att.p; a.ld_out = l.att.ld;
  a.lse = l.lse;
  a.tk_pack = b.tk_packs ? b.tk_packs + static_cast<int64_t>(i) * 64 * 64 : nullptr;
  a.tv_pack = b.tv_packs ? b.tv_packs + static_cast<int64_t>(i) * 64 * 64 : nullptr;
  a.tables_per_head = 0;
  a.idx_a = a.tk_pack ? d.idx_a : nullptr; a.idx_b = a.tk_pack ? d.idx_b : nullptr;
  a.idx_va = a.tv_pack ? d.idx_va : nullptr; a.idx_vb = a.tv_pack ? d.idx_vb : nullptr;
  a.ld_idx = d.ld_idx;
  a.af_grid = d.af_grid; a.af_max_rel = d.af_max_rel;
  a.gp_grid = d.gp_grid; a.gp_w = d.gp_w; a.gp_skip_id = d.gp_skip_id;
  std::memcpy(a.gp_lut_a, d.gp_lut_a, sizeof(a.gp_lut_a));
  std::memcpy(a.gp_lut_b, d.gp_lut_b, sizeof(a.gp_lut_b));
}

}  // namespace
}  // namespace cb

extern "C" int cream_vit_last_launches(void) { return cb::t_launches; }

extern "C" int64_t cream_vit_arena_bytes(const cream_vit_desc* d) {
  using namespace cb;
  if (validate(d) != CREAM_OK) return -1;
  cream_vit_desc tmp = *d;
  tmp.arena = nullptr;
  Bufs b;
  plan(tmp, b);
  return b.total + 256;
}

extern "C" int cream_vit_fwd(const cream_vit_desc* d, void* stream_) {
  using namespace cb;
  t_launches = 0;
  VIT_TRY(validate(d));
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(d->arena != nullptr && (reinterpret_cast<uintptr_t>(d->arena) & 255) == 0, "arena must be 256-byte aligned");
  CB_REQUIRE(d->images && d->wpatch && d->cls && d->norm_g && d->norm_b && d->whead && d->logits, "null parameter");
  Bufs b;
  plan(*d, b);
  CB_REQUIRE(b.total <= d->arena_bytes, "arena too small (cream_vit_arena_bytes)");
  const int64_t T = d->N - 1, M = static_cast<int64_t>(d->B) * d->N;
  const int E = d->E;
  const int kdim = d->in_chans * d->patch_size * d->patch_size;

  // ---- patch embedding as a sliced GEMM over im2col patches (embedding_super.py:33-40) ----
  ++t_launches;
  VIT_TRY(cream_patch_im2col(d->images, b.cols.p, b.cols.ld, d->B, d->in_chans, d->img_size, d->img_size, d->patch_size, s));
  VIT_TRY(linear_fwd(s, d->B * T, E, kdim, b.cols, d->wpatch, d->ld_wpatch, d->bpatch, CREAM_EPI_BF16, b.patch.p, b.patch.ld));
  ++t_launches;
  VIT_TRY(cream_tokens_assemble_fwd(b.patch.p, b.patch.ld, d->cls, d->pos, d->ld_pos, b.L[0].x.p, b.L[0].x.ld, d->B, d->N, E, s));
  VIT_TRY(pack_layer_tables(*d, b, s));

  for (int i = 0; i < d->depth; ++i) {
    const cream_vit_layer& p = d->layers[i];
    const LayerBufs& l = b.L[i];
    const int h = p.heads, qd = 64 * h, ffn = p.ffn;
    const F32Mat x2 = (i + 1 < d->depth) ? b.L[i + 1].x : b.x_last;
    ++t_launches;
    VIT_TRY(cream_layernorm_fwd(l.x.p, l.x.ld, p.ln1_g, p.ln1_b, d->eps, l.ln1.p, l.ln1.ld, 0, l.mu1, l.rs1, M, E, s));
    {   // QKV: three row blocks of the shadow (qkv_super.py:45-55 de-interleaved, or plain [q;k;v])
      GemmArgs a;
      a.g.M = static_cast<int>(M); a.g.N = qd; a.g.K = E; a.g.groups = 3;
      a.g.a = l.ln1.p; a.g.lda = l.ln1.ld;
      a.g.b = p.wqkv; a.g.ldb = d->ld_wqkv; a.g.b_group_rows = d->qkv_group_rows;
      a.g.epi = CREAM_EPI_BF16; a.g.out = l.qkv.p; a.g.ldo = l.qkv.ld; a.g.out_g_col = qd;
      a.g.bias = p.bqkv;
      VIT_TRY(run_gemm(a.g, s));
    }
    cream_attn_desc at;
    fill_attn(*d, b, i, at);
    ++t_launches;
    VIT_TRY(cream_attn_fwd(&at, s));
    VIT_TRY(linear_fwd(s, M, E, qd, l.att, p.wproj, d->ld_wproj, p.bproj, CREAM_EPI_F32_RESID, l.x1.p, l.x1.ld, nullptr, 0,
                       l.x.p, l.x.ld, p.dp_scale, d->N));
    ++t_launches;
    VIT_TRY(cream_layernorm_fwd(l.x1.p, l.x1.ld, p.ln2_g, p.ln2_b, d->eps, l.ln2.p, l.ln2.ld, 0, l.mu2, l.rs2, M, E, s));
    VIT_TRY(linear_fwd(s, M, ffn, E, l.ln2, p.wfc1, d->ld_wfc1, p.bfc1, CREAM_EPI_BF16_GELU, l.act.p, l.act.ld, l.hpre.p, l.hpre.ld));
    VIT_TRY(linear_fwd(s, M, E, ffn, l.act, p.wfc2, d->ld_wfc2, p.bfc2, CREAM_EPI_F32_RESID, x2.p, x2.ld, nullptr, 0, l.x1.p,
                       l.x1.ld, p.dp_scale ? p.dp_scale + d->B : nullptr, d->N));
  }

  ++t_launches;
  VIT_TRY(cream_layernorm_fwd(b.x_last.p, b.x_last.ld, d->norm_g, d->norm_b, d->eps, b.y.p, b.y.ld, 1, b.mu_f, b.rs_f, M, E, s));
  ++t_launches;
  VIT_TRY(cream_pool_fwd(b.y.p, b.y.ld, b.pooled.p, b.pooled.ld, d->B, d->N, E, d->pool_first, d->pool_count, s));
  VIT_TRY(linear_fwd(s, d->B, d->num_classes, E, b.pooled, d->whead, d->ld_whead, d->bhead, CREAM_EPI_F32, d->logits, d->ld_logits));
  return CREAM_OK;
}

extern "C" int cream_vit_bwd(const cream_vit_desc* d, int first_stage, int last_stage, void* stream_) {
  using namespace cb;
  t_launches = 0;
  VIT_TRY(validate(d));
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(d->arena != nullptr && d->dlogits != nullptr, "null pointer");
  CB_REQUIRE(first_stage >= 0 && last_stage <= d->depth + 1 && first_stage <= last_stage, "bad stage range");
  Bufs b;
  plan(*d, b);
  CB_REQUIRE(b.total <= d->arena_bytes, "arena too small (cream_vit_arena_bytes)");
  const int64_t T = d->N - 1, M = static_cast<int64_t>(d->B) * d->N;
  const int E = d->E;
  // The residual-stream gradient enters every stage in `ga` and leaves it in `ga` (a layer goes
  // ga -> gb -> ga), so any contiguous stage range can be run by a separate call.
  const F32Mat g = b.ga, g1 = b.gb;

  for (int stage = first_stage; stage <= last_stage; ++stage) {
    if (stage == 0) {
      // ---- head: logits = pooled W^T + b; pooling; final LayerNorm ----
      ++t_launches;
      VIT_TRY(cream_cast_scale(d->dlogits, d->ld_dlogits, b.dl.p, b.dl.ld, nullptr, 1, nullptr, d->B, d->num_classes, s));
      if (d->g_bhead) VIT_TRY(bias_grad(s, b.dl, d->B, d->num_classes, d->g_bhead));
      VIT_TRY(linear_wgrad(s, d->B, d->num_classes, E, b.dl, b.pooled, d->g_whead, d->ld_ghead));
      VIT_TRY(linear_dgrad(s, d->B, d->num_classes, E, b.dl, d->whead, d->ld_whead, CREAM_EPI_BF16, b.dpooled));
      ++t_launches;
      VIT_TRY(cream_pool_bwd(b.dpooled.p, b.dpooled.ld, b.dy.p, b.dy.ld, d->B, d->N, E, d->pool_first, d->pool_count, s));
      {   // final LayerNorm backward; it also emits the bf16 DropPath-scaled gradient + fc2 bias gradient
          // that the last layer's FFN branch starts from
        const cream_vit_layer& nx = d->layers[d->depth - 1];
        ++t_launches;
        VIT_TRY(cream_layernorm_bwd_cast(b.dy.p, b.dy.ld, 1, b.x_last.p, b.x_last.ld, d->norm_g, b.mu_f, b.rs_f, nullptr, 0, g.p,
                                         g.ld, d->g_norm_g, d->g_norm_b, M, E, b.dy_bf.p, b.dy_bf.ld,
                                         nx.dp_scale ? nx.dp_scale + d->B : nullptr, d->N, nx.g_bfc2, s));
      }
      if (b.dtk_packs) { ++t_launches; CB_CUDA_OK(cudaMemsetAsync(b.dtk_packs, 0, static_cast<size_t>(d->depth) * 64 * 64 * 4, s)); }
      if (b.dtv_packs) { ++t_launches; CB_CUDA_OK(cudaMemsetAsync(b.dtv_packs, 0, static_cast<size_t>(d->depth) * 64 * 64 * 4, s)); }
      continue;
    }
    if (stage == d->depth + 1) {
      // ---- embedding: cls / pos gradients, patch projection ----
      ++t_launches;
      VIT_TRY(cream_tokens_assemble_bwd(g.p, g.ld, b.dpatch.p, b.dpatch.ld, d->g_pos, d->ld_pos, d->g_cls, d->B, d->N, E, s));
      if (d->g_bpatch) VIT_TRY(bias_grad(s, b.dpatch, d->B * T, E, d->g_bpatch));
      const int kdim = d->in_chans * d->patch_size * d->patch_size;
      VIT_TRY(linear_wgrad(s, d->B * T, E, kdim, b.dpatch, b.cols, d->g_wpatch, d->ld_gpatch));
      continue;
    }
    const int i = d->depth - stage;          // stages 1 .. depth run layers depth-1 .. 0
    const cream_vit_layer& p = d->layers[i];
    const LayerBufs& l = b.L[i];
    const int h = p.heads, qd = 64 * h, ffn = p.ffn;
    // ---- FFN branch: x2 = x1 + s * fc2(gelu(fc1(ln2))) ----
    // (b.dy_bf = bf16(DropPath scale * g) and the fc2 bias gradient were produced by the LayerNorm
    //  backward of the previous stage)
    const bool sw = side_wgrad_enabled();
    VIT_TRY(linear_wgrad(sw ? side_fork(s) : s, M, E, ffn, b.dy_bf, l.act, p.g_wfc2, d->ld_gfc2));
    VIT_TRY(linear_dgrad(s, M, E, ffn, b.dy_bf, p.wfc2, d->ld_wfc2, CREAM_EPI_BF16_DGELU, b.dh, l.hpre.p, l.hpre.ld));
    if (p.g_bfc1) VIT_TRY(side_bias_grad(s, b.dh, M, ffn, p.g_bfc1));
    VIT_TRY(linear_wgrad(sw ? side_fork(s) : s, M, ffn, E, b.dh, l.ln2, p.g_wfc1, d->ld_gfc1));
    VIT_TRY(linear_dgrad(s, M, ffn, E, b.dh, p.wfc1, d->ld_wfc1, CREAM_EPI_BF16, b.dln));
    if (sw) VIT_TRY(side_join(s));   // the side-stream weight gradients have read dy_bf (rewritten below)
    ++t_launches;   // ffn_layer_norm backward (+ residual gradient) -> g1, its bf16 DropPath-scaled copy, proj bias gradient
    VIT_TRY(cream_layernorm_bwd_cast(b.dln.p, b.dln.ld, 0, l.x1.p, l.x1.ld, p.ln2_g, l.mu2, l.rs2, g.p, g.ld, g1.p, g1.ld,
                                     p.g_ln2_g, p.g_ln2_b, M, E, b.dy_bf.p, b.dy_bf.ld, p.dp_scale, d->N, p.g_bproj, s));
    // ---- attention branch: x1 = x + s * proj(attn(qkv(ln1))) ----
    VIT_TRY(linear_wgrad(sw ? side_fork(s) : s, M, E, qd, b.dy_bf, l.att, p.g_wproj, d->ld_gproj));
    VIT_TRY(linear_dgrad(s, M, E, qd, b.dy_bf, p.wproj, d->ld_wproj, CREAM_EPI_BF16, b.datt));
    {
      cream_attn_desc at;
      fill_attn(*d, b, i, at);
      at.dout = b.datt.p; at.ld_dout = b.datt.ld;
      at.dqkv = b.dqkv.p; at.ld_dqkv = b.dqkv.ld;
      at.dtk_pack = b.dtk_packs ? b.dtk_packs + static_cast<int64_t>(i) * 64 * 64 : nullptr;
      at.dtv_pack = b.dtv_packs ? b.dtv_packs + static_cast<int64_t>(i) * 64 * 64 : nullptr;
      at.workspace = b.attn_ws; at.workspace_bytes = b.attn_ws_bytes;
      t_launches += 2;
      VIT_TRY(cream_attn_bwd(&at, s));
    }
    if (p.g_bqkv) VIT_TRY(side_bias_grad(s, b.dqkv, M, 3 * qd, p.g_bqkv));
    {   // dWqkv: the reference's interleaved rows 3j+i (qkv_super.py:72-77) or plain [q;k;v] row blocks
      GemmArgs a;
      a.g.M = qd; a.g.N = E; a.g.K = static_cast<int>(M); a.g.groups = 3;
      a.g.a = b.dqkv.p; a.g.lda = b.dqkv.ld; a.g.a_mn = 1; a.g.a_group_off = qd;
      a.g.b = l.ln1.p; a.g.ldb = l.ln1.ld; a.g.b_mn = 1;
      a.g.epi = CREAM_EPI_F32_ATOMIC; a.g.out = p.g_wqkv; a.g.ldo = d->ld_gqkv;
      if (d->qkv_interleaved) { a.g.out_row_mul = 3; a.g.out_g_row = 1; }
      else { a.g.out_row_mul = 1; a.g.out_g_row = d->qkv_group_rows; }
      VIT_TRY(run_gemm(a.g, sw ? side_fork(s) : s));
    }
    {   // dln1 = dqkv Wqkv: contraction over the three row blocks of the shadow
      GemmArgs a;
      a.g.M = static_cast<int>(M); a.g.N = E; a.g.K = 3 * qd;
      a.g.a = b.dqkv.p; a.g.lda = b.dqkv.ld;
      a.g.b = p.wqkv; a.g.ldb = d->ld_wqkv; a.g.b_mn = 1;
      a.g.k_groups = 3; a.g.k_group_len = qd; a.g.b_group_rows = d->qkv_group_rows;
      a.g.epi = CREAM_EPI_BF16; a.g.out = b.dln.p; a.g.ldo = b.dln.ld;
      VIT_TRY(run_gemm(a.g, s));
    }
    VIT_TRY(side_join(s));   // the side-stream column sums of this layer have read dh / dqkv
    ++t_launches;   // attn_layer_norm backward -> g; for i > 0 also what layer i-1's FFN branch starts from
    if (i > 0) {
      const cream_vit_layer& nx = d->layers[i - 1];
      VIT_TRY(cream_layernorm_bwd_cast(b.dln.p, b.dln.ld, 0, l.x.p, l.x.ld, p.ln1_g, l.mu1, l.rs1, g1.p, g1.ld, g.p, g.ld,
                                       p.g_ln1_g, p.g_ln1_b, M, E, b.dy_bf.p, b.dy_bf.ld,
                                       nx.dp_scale ? nx.dp_scale + d->B : nullptr, d->N, nx.g_bfc2, s));
    } else {
      VIT_TRY(cream_layernorm_bwd(b.dln.p, b.dln.ld, 0, l.x.p, l.x.ld, p.ln1_g, l.mu1, l.rs1, g1.p, g1.ld, g.p, g.ld, p.g_ln1_g,
                                  p.g_ln1_b, M, E, s));
    }
  }
  // table gradients of the layers this call covered: packed (64, 64) fp32 -> the reference's table tensors,
  // ONE launch per side for all of them (a single-block launch per layer costs ~12 us of latency each)
  {
    const int lo_stage = std::max(first_stage, 1), hi_stage = std::min(last_stage, d->depth);
    const int n_layers = hi_stage - lo_stage + 1;                 // layers depth-hi_stage .. depth-lo_stage
    if (n_layers > 0) {
      const int first_layer = d->depth - hi_stage;
      for (int side = 0; side < 2; ++side) {
        float* dp = side ? b.dtv_packs : b.dtk_packs;
        if (dp == nullptr || d->layers[first_layer].g_tab[2 * side] == nullptr) continue;
        float* g0[CREAM_VIT_MAX_DEPTH];
        float* g1[CREAM_VIT_MAX_DEPTH];
        bool second = false;
        for (int j = 0; j < n_layers; ++j) {
          g0[j] = d->layers[first_layer + j].g_tab[2 * side];
          g1[j] = d->layers[first_layer + j].g_tab[2 * side + 1];
          second = second || g1[j] != nullptr;
        }
        ++t_launches;
        VIT_TRY(cream_unpack_table_grads_batch(dp + static_cast<int64_t>(first_layer) * 64 * 64, n_layers, 64, g0,
                                               second ? g1 : nullptr, d->tab_nb, d->tab_row_off1,
                                               side ? d->tabv_stride_b : d->tab_stride_b,
                                               side ? d->tabv_stride_d : d->tab_stride_d, s));
      }
    }
  }
  return CREAM_OK;
}
