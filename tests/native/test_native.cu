// Stand-alone driver over the C ABI (no torch): correctness of the rpe_index and GEMM
// kernels against straightforward CPU loops, plus event timings.  Each case runs in its
// own process (a device trap poisons the context):  test_native <case> | list
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <random>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cream_b200.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static std::mt19937 rng(1234);
static float frand() { return std::uniform_real_distribution<float>(-1.f, 1.f)(rng); }
static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <typename T> static T* dmalloc(size_t n) {
  T* p;
  CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}
template <typename T> static void h2d(T* d, const std::vector<T>& h) {
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
}
template <typename T> static std::vector<T> d2h(const T* d, size_t n) {
  std::vector<T> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}
static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
  std::vector<__nv_bfloat16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
  return o;
}

static int report(const char* name, double err, double tol) {
  const bool ok = err <= tol && std::isfinite(err);
  printf("%-44s max_err %.3e (tol %.1e) %s\n", name, err, tol, ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}

// ----------------------------------------------------------------------------------
static int test_rpe_index(int B, int H, int L, int nb, bool timing) {
  std::vector<float> in((size_t)B * H * L * nb), gout((size_t)B * H * L * L);
  std::vector<int32_t> idx((size_t)L * L);
  for (auto& v : in) v = frand();
  for (auto& v : gout) v = frand();
  for (auto& v : idx) v = rng() % nb;
  float* d_in = dmalloc<float>(in.size());
  float* d_out = dmalloc<float>(gout.size());
  float* d_gout = dmalloc<float>(gout.size());
  float* d_gin = dmalloc<float>(in.size());
  int32_t* d_idx = dmalloc<int32_t>(idx.size());
  h2d(d_in, in); h2d(d_gout, gout); h2d(d_idx, idx);
  // input given in the transposed-view layout irpe.py:639-642 produces: strides (L*nb, B*L*nb, nb, 1)
  std::vector<float> in_t((size_t)B * H * L * nb);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int i = 0; i < L; ++i)
        for (int k = 0; k < nb; ++k)
          in_t[(((size_t)h * B + b) * L + i) * nb + k] = in[(((size_t)b * H + h) * L + i) * nb + k];
  float* d_in_t = dmalloc<float>(in_t.size());
  h2d(d_in_t, in_t);
  int rc = cream_rpe_index_fwd(d_in_t, d_idx, d_out, B, H, L, L, nb, (int64_t)L * nb, (int64_t)B * L * nb, nb,
                               1, CREAM_DTYPE_F32, nullptr);
  CK(cudaDeviceSynchronize());
  if (rc) { printf("rpe_index_fwd rc=%d\n", rc); return 1; }
  auto out = d2h(d_out, gout.size());
  double err = 0;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int i = 0; i < L; ++i)
        for (int j = 0; j < L; ++j) {
          const size_t row = ((size_t)b * H + h) * L + i;
          err = std::max(err, (double)std::fabs(out[row * L + j] - in[row * nb + idx[(size_t)i * L + j]]));
        }
  int fails = report("rpe_index fwd (strided input, bit-exact)", err, 0.0);
  CK(cudaMemset(d_gin, 0, in.size() * sizeof(float)));
  rc = cream_rpe_index_bwd(d_gin, d_gout, d_idx, B, H, L, L, nb, CREAM_DTYPE_F32, nullptr);
  CK(cudaDeviceSynchronize());
  if (rc) { printf("rpe_index_bwd rc=%d\n", rc); return 1; }
  auto gin = d2h(d_gin, in.size());
  std::vector<double> ref(in.size(), 0.0);
  for (size_t row = 0; row < (size_t)B * H * L; ++row) {
    const int i = row % L;
    for (int j = 0; j < L; ++j) ref[row * nb + idx[(size_t)i * L + j]] += gout[row * L + j];
  }
  err = 0;
  for (size_t k = 0; k < ref.size(); ++k) err = std::max(err, std::fabs(ref[k] - gin[k]));
  fails += report("rpe_index bwd (scatter-add)", err, 1e-4);
  if (timing) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 3; ++it)
      cream_rpe_index_fwd(d_in, d_idx, d_out, B, H, L, L, nb, (int64_t)H * L * nb, (int64_t)L * nb, nb, 1, CREAM_DTYPE_F32, nullptr);
    CK(cudaEventRecord(e0));
    const int iters = 20;
    for (int it = 0; it < iters; ++it)
      cream_rpe_index_fwd(d_in, d_idx, d_out, B, H, L, L, nb, (int64_t)H * L * nb, (int64_t)L * nb, nb, 1, CREAM_DTYPE_F32, nullptr);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)gout.size() * 4 + (double)in.size() * 4;
    printf("rpe_index fwd  B%d H%d L%d nb%d: %.1f us  %.0f GB/s (algorithmic)\n", B, H, L, nb, ms / iters * 1e3, bytes / (ms / iters * 1e-3) / 1e9);
    CK(cudaEventRecord(e0));
    for (int it = 0; it < iters; ++it)
      cream_rpe_index_bwd(d_gin, d_gout, d_idx, B, H, L, L, nb, CREAM_DTYPE_F32, nullptr);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes_b = (double)gout.size() * 4 + 2.0 * in.size() * 4;
    printf("rpe_index bwd  B%d H%d L%d nb%d: %.1f us  %.0f GB/s (algorithmic)\n", B, H, L, nb, ms / iters * 1e3, bytes_b / (ms / iters * 1e-3) / 1e9);
  }
  return fails;
}

// ----------------------------------------------------------------------------------
// Generic GEMM check.  Logical problem: for each group g: C_g[m,n] = sum_k A[m,k] * B_g[n,k].
struct GemmCase {
  const char* name;
  int M, N, K, groups;
  int a_mn, b_mn;
  int epi;
  int k_groups;       // for b_mn dgrad-QKV
  bool timing;
  int pad;            // extra leading-dimension padding (slice of a larger tensor)
};

static float gelu_h(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678f)); }
static float dgelu_h(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678f));
  return cdf + x * 0.39894228f * expf(-0.5f * x * x);
}

static int test_gemm(const GemmCase& c) {
  const int M = c.M, N = c.N, K = c.K, G = c.groups;
  // logical operands (bf16-rounded floats)
  // A: if groups apply to A (a_mn with groups>1) A_g differs per group, else shared.
  const bool a_grouped = c.a_mn && G > 1;
  const int AG = a_grouped ? G : 1;
  std::vector<float> A((size_t)AG * M * K), Bm((size_t)G * N * K);
  const bool b_grouped = !(c.a_mn && c.b_mn);  // wgrad: B (=X) shared across groups
  for (auto& v : A) v = bf16r(frand());
  for (auto& v : Bm) v = bf16r(frand());
  if (!b_grouped)
    for (int g = 1; g < G; ++g) std::copy(Bm.begin(), Bm.begin() + (size_t)N * K, Bm.begin() + (size_t)g * N * K);

  // ---- physical layouts ---------------------------------------------------------
  // A K-major : (M rows, lda) ; A MN-major: (K rows, lda) with column = g*a_group_off + m
  const int a_group_off = a_grouped ? ((M + 63) / 64) * 64 : 0;
  const int64_t lda = c.a_mn ? (int64_t)((AG - 1) * a_group_off + M + 7) / 8 * 8 + c.pad : (int64_t)(K + 7) / 8 * 8 + c.pad;
  std::vector<float> Ap((size_t)(c.a_mn ? K : M) * lda, 0.f);
  for (int g = 0; g < AG; ++g)
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) {
        const float v = A[((size_t)g * M + m) * K + k];
        if (c.a_mn) Ap[(size_t)k * lda + g * a_group_off + m] = v;
        else Ap[(size_t)m * lda + k] = v;
      }
  // B K-major: rows (g*b_group_rows + n), ldb ; B MN-major: (K rows [k-grouped], ldb) col = n
  const int kg = std::max(1, c.k_groups);
  const int kgl = K / kg;
  int64_t b_group_rows = 0, ldb;
  std::vector<float> Bp;
  if (!c.b_mn) {
    b_group_rows = N + 24;  // gap between groups, as in a sliced supernet tensor
    ldb = (int64_t)(K + 7) / 8 * 8 + c.pad;
    Bp.assign((size_t)(G * b_group_rows) * ldb, 7.f);  // 7 = poison outside the slice
    for (int g = 0; g < G; ++g)
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) Bp[((size_t)g * b_group_rows + n) * ldb + k] = Bm[((size_t)g * N + n) * K + k];
  } else {
    b_group_rows = kg > 1 ? kgl + 64 : 0;
    ldb = (int64_t)(N + 7) / 8 * 8 + c.pad;
    const size_t rows = kg > 1 ? (size_t)(kg - 1) * b_group_rows + kgl : (size_t)K;
    Bp.assign(rows * ldb, 7.f);
    for (int k = 0; k < K; ++k) {
      const size_t prow = kg > 1 ? (size_t)(k / kgl) * b_group_rows + (k % kgl) : (size_t)k;
      for (int n = 0; n < N; ++n) Bp[prow * ldb + n] = Bm[(size_t)n * K + k];
    }
  }
  auto Ab = to_bf16(Ap), Bb = to_bf16(Bp);
  __nv_bfloat16* dA = dmalloc<__nv_bfloat16>(Ab.size());
  __nv_bfloat16* dB = dmalloc<__nv_bfloat16>(Bb.size());
  h2d(dA, Ab); h2d(dB, Bb);

  // ---- output -------------------------------------------------------------------
  const bool out_bf16 = c.epi == CREAM_EPI_BF16 || c.epi == CREAM_EPI_BF16_GELU || c.epi == CREAM_EPI_BF16_DGELU;
  const bool wgrad_qkv = a_grouped;  // rows interleaved 3*m+g
  const int out_g_col = (!a_grouped && G > 1) ? ((N + 7) / 8 * 8) : 0;
  const int64_t out_cols = (!a_grouped && G > 1) ? (int64_t)G * out_g_col : N;
  const int64_t ldo = (out_cols + 7) / 8 * 8 + c.pad;
  const int64_t out_rows = wgrad_qkv ? (int64_t)G * M : M;
  const size_t out_elems = (size_t)out_rows * ldo;
  std::vector<float> bias(out_cols), resid, rscale;
  for (auto& v : bias) v = frand();
  float* d_bias = dmalloc<float>(bias.size()); h2d(d_bias, bias);
  void* d_out = nullptr; void* d_aux = nullptr; float* d_res = nullptr; float* d_rs = nullptr;
  std::vector<float> aux_h;
  if (out_bf16) { d_out = dmalloc<__nv_bfloat16>(out_elems); CK(cudaMemset(d_out, 0, out_elems * 2)); }
  else { d_out = dmalloc<float>(out_elems); CK(cudaMemset(d_out, 0, out_elems * 4)); }
  if (c.epi == CREAM_EPI_BF16_GELU || c.epi == CREAM_EPI_BF16_DGELU) {
    d_aux = dmalloc<__nv_bfloat16>(out_elems);
    aux_h.resize(out_elems);
    for (auto& v : aux_h) v = bf16r(2.f * frand());
    auto ab = to_bf16(aux_h);
    h2d((__nv_bfloat16*)d_aux, ab);
  }
  const int rows_per_scale = 7;
  if (c.epi == CREAM_EPI_F32_RESID) {
    resid.resize(out_elems); for (auto& v : resid) v = frand();
    d_res = dmalloc<float>(out_elems); h2d(d_res, resid);
    rscale.resize((M + rows_per_scale - 1) / rows_per_scale); for (auto& v : rscale) v = 0.5f + frand();
    d_rs = dmalloc<float>(rscale.size()); h2d(d_rs, rscale);
  }

  cream_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M; d.N = N; d.K = K; d.groups = G;
  d.a = dA; d.lda = lda; d.a_mn = c.a_mn; d.a_group_off = a_group_off;
  d.b = dB; d.ldb = ldb; d.b_mn = c.b_mn; d.b_group_rows = b_group_rows;
  d.k_groups = kg; d.k_group_len = kgl;
  d.epi = c.epi; d.out = d_out; d.ldo = ldo;
  d.out_row_mul = wgrad_qkv ? G : 1; d.out_g_row = wgrad_qkv ? 1 : 0; d.out_g_col = out_g_col;
  d.aux = d_aux; d.ldaux = ldo;
  d.bias = (c.epi == CREAM_EPI_F32_ATOMIC || c.epi == CREAM_EPI_BF16_DGELU) ? nullptr : d_bias;
  d.resid = d_res; d.ldr = ldo; d.row_scale = d_rs; d.rows_per_scale = rows_per_scale;
  d.alpha = 1.0f; d.split_k = 0;
  if (const char* e = getenv("CREAM_TEST_CTA_PAIR")) d.cta_pair = atoi(e);   // 1 = single CTA, 2 = CTA pairs

  int rc = cream_gemm_bf16(&d, nullptr);
  cudaError_t se = cudaDeviceSynchronize();
  if (rc || se != cudaSuccess) { printf("%-44s rc=%d cuda=%s FAIL\n", c.name, rc, cudaGetErrorString(se)); return 1; }

  std::vector<float> out(out_elems);
  if (out_bf16) {
    auto ob = d2h((__nv_bfloat16*)d_out, out_elems);
    for (size_t i = 0; i < out_elems; ++i) out[i] = __bfloat162float(ob[i]);
  } else out = d2h((float*)d_out, out_elems);
  std::vector<float> aux_out;
  if (c.epi == CREAM_EPI_BF16_GELU) {
    auto ab = d2h((__nv_bfloat16*)d_aux, out_elems);
    aux_out.resize(out_elems);
    for (size_t i = 0; i < out_elems; ++i) aux_out[i] = __bfloat162float(ab[i]);
  }

  // ---- reference (sampled rows for big problems) ------------------------------------
  double err = 0, err_aux = 0;
  const int row_step = M > 2048 ? M / 97 : 1;
  for (int g = 0; g < G; ++g)
    for (int m = 0; m < M; m += row_step)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        const float* a = &A[((size_t)(a_grouped ? g : 0) * M + m) * K];
        const float* b = &Bm[((size_t)g * N + n) * K];
        for (int k = 0; k < K; ++k) acc += (double)a[k] * b[k];
        const int64_t orow = wgrad_qkv ? (int64_t)m * G + g : m;
        const int64_t ocol = n + (int64_t)g * out_g_col;
        const size_t o = (size_t)orow * ldo + ocol;
        double ref;
        switch (c.epi) {
          case CREAM_EPI_BF16: ref = bf16r((float)(acc + bias[ocol])); break;
          case CREAM_EPI_F32: ref = acc + bias[ocol]; break;
          case CREAM_EPI_F32_ATOMIC: ref = acc; break;
          case CREAM_EPI_F32_RESID: ref = resid[o] + rscale[m / rows_per_scale] * (acc + bias[ocol]); break;
          case CREAM_EPI_BF16_GELU: {
            const float pre = bf16r((float)(acc + bias[ocol]));
            err_aux = std::max(err_aux, (double)std::fabs(pre - aux_out[o]));
            // compare GELU on the device's own rounded pre-activation to avoid 1-ulp bf16 flips
            ref = bf16r(gelu_h(aux_out[o]));
            break;
          }
          default: ref = bf16r((float)(acc * dgelu_h(aux_h[o]))); break;
        }
        err = std::max(err, std::fabs(ref - (double)out[o]));
      }
  const double scale = std::sqrt((double)K);
  const double tol = out_bf16 ? 0.02 * scale : 2e-4 * scale;
  int fails = report(c.name, err, tol);
  if (c.epi == CREAM_EPI_BF16_GELU) fails += report("  (gelu pre-activation aux)", err_aux, 0.02 * scale);
  // poison check: nothing outside the valid region may be written (sampled)
  if (!out_bf16 && c.epi != CREAM_EPI_F32_RESID) {
    double stray = 0;
    for (int64_t r = 0; r < out_rows; r += std::max<int64_t>(1, out_rows / 50))
      for (int64_t cc = out_cols; cc < ldo; ++cc) stray = std::max(stray, (double)std::fabs(out[(size_t)r * ldo + cc]));
    fails += report("  (no stray writes beyond N)", stray, 0.0);
  }

  if (c.timing) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 3; ++it) cream_gemm_bf16(&d, nullptr);
    const int iters = 20;
    CK(cudaEventRecord(e0));
    for (int it = 0; it < iters; ++it) cream_gemm_bf16(&d, nullptr);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * M * N * K * G;
    printf("  %-42s %.1f us  %.1f TFLOP/s\n", c.name, ms / iters * 1e3, fl / (ms / iters * 1e-3) / 1e12);
  }
  return fails;
}

// ----------------------------------------------------------------------------------
static int test_index_tables() {
  int fails = 0;
  const int g = 14, n = g * g + 1;
  std::vector<int32_t> v(n * n), h(n * n);
  if (cream_autoformer_rel_index_host(g, 14, v.data(), h.data())) return 1;
  int mn = 1 << 30, mx = -1;
  for (int i = 0; i < n * n; ++i) { mn = std::min({mn, v[i], h[i]}); mx = std::max({mx, v[i], h[i]}); }
  printf("autoformer rel index: min %d max %d  v[1,1]=%d v[1,15]=%d h[1,2]=%d\n", mn, mx, v[n + 1], v[n + 15], h[n + 2]);
  fails += !(mn == 0 && mx == 28 && v[n + 1] == 15 && v[n + 15] == 16 && h[n + 2] == 16);
  std::vector<int32_t> ids(n * n);
  int nb = 0;
  if (cream_irpe_bucket_ids_host(3, 14, 14, 1, 1.9, 3.8, 15.2, ids.data(), &nb)) return 1;
  const int golden[15] = {24, 23, 22, 22, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 17};
  int bad = nb != 50;
  for (int j = 0; j < 15; ++j) bad += ids[n + 1 + j] != golden[j];
  printf("irpe product bucket ids: nb=%d row1 %s\n", nb, bad ? "MISMATCH" : "matches SURVEY golden");
  return fails + bad;
}

int main(int argc, char** argv) {
  std::map<std::string, std::function<int()>> cases;
  cases["tables"] = [] { return test_index_tables(); };
  cases["rpe_small"] = [] { return test_rpe_index(3, 2, 50, 50, false); };
  cases["rpe_c2"] = [] { return test_rpe_index(32, 6, 197, 50, true); };
  auto G = [&](const char* key, GemmCase c) { cases[key] = [c] { return test_gemm(c); }; };
  //                      name                                  M     N     K   G amn bmn epi               kg timing pad
  G("gemm_kk_small",   {"gemm K/K  f32 128x64x64",            128,   64,   64, 1, 0, 0, CREAM_EPI_F32,        1, false, 0});
  G("gemm_kk_ragged",  {"gemm K/K  f32 300x200x136 (+pad)",   300,  200,  136, 1, 0, 0, CREAM_EPI_F32,        1, false, 40});
  G("gemm_kk_bf16",    {"gemm K/K  bf16 300x328x216",         300,  328,  216, 1, 0, 0, CREAM_EPI_BF16,       1, false, 8});
  G("gemm_qkv",        {"gemm K/K  bf16 qkv groups=3",        394,  320,  320, 3, 0, 0, CREAM_EPI_BF16,       1, false, 64});
  G("gemm_gelu",       {"gemm K/K  bf16+gelu 394x1120x320",   394, 1120,  320, 1, 0, 0, CREAM_EPI_BF16_GELU,  1, false, 0});
  G("gemm_resid",      {"gemm K/K  f32 resid 394x320x1120",   394,  320, 1120, 1, 0, 0, CREAM_EPI_F32_RESID,  1, false, 0});
  G("gemm_dgrad",      {"gemm K/MN bf16 dgrad 394x320x1120",  394,  320, 1120, 1, 0, 1, CREAM_EPI_BF16,       1, false, 16});
  G("gemm_dgrad_qkv",  {"gemm K/MN bf16 dgrad-qkv kg=3",      394,  320,  960, 1, 0, 1, CREAM_EPI_BF16,       3, false, 0});
  G("gemm_dgelu",      {"gemm K/MN bf16 dgelu 394x1120x320",  394, 1120,  320, 1, 0, 1, CREAM_EPI_BF16_DGELU, 1, false, 0});
  G("gemm_wgrad",      {"gemm MN/MN f32 atomic 1120x320x1576",1120, 320, 1576, 1, 1, 1, CREAM_EPI_F32_ATOMIC, 1, false, 0});
  G("gemm_wgrad_qkv",  {"gemm MN/MN f32 atomic qkv groups=3",  320,  320, 1576, 3, 1, 1, CREAM_EPI_F32_ATOMIC, 1, false, 0});
  G("perf_qkv",        {"perf fwd qkv  25216x(3x448)x448",   25216, 448,  448, 3, 0, 0, CREAM_EPI_BF16,       1, true, 0});
  G("perf_fc1",        {"perf fwd fc1  25216x1792x448 gelu", 25216, 1792, 448, 1, 0, 0, CREAM_EPI_BF16_GELU,  1, true, 0});
  G("perf_fc2",        {"perf fwd fc2  25216x448x1792 resid",25216, 448, 1792, 1, 0, 0, CREAM_EPI_F32_RESID,  1, true, 0});
  G("perf_dgrad",      {"perf dgrad    25216x448x1792",      25216, 448, 1792, 1, 0, 1, CREAM_EPI_BF16,       1, true, 0});
  G("perf_wgrad",      {"perf wgrad    1792x448x25216",       1792, 448, 25216, 1, 1, 1, CREAM_EPI_F32_ATOMIC, 1, true, 0});
  G("perf_big",        {"perf fwd      8192x8192x8192 (not a path shape)", 8192, 8192, 8192, 1, 0, 0, CREAM_EPI_BF16, 1, true, 0});
  if (argc < 2 || std::string(argv[1]) == "list") {
    for (auto& kv : cases) printf("%s\n", kv.first.c_str());
    return 0;
  }
  auto it = cases.find(argv[1]);
  if (it == cases.end()) { printf("unknown case %s\n", argv[1]); return 2; }
  printf("== %s (%s)\n", argv[1], cream_version());
  const int f = it->second();
  printf("== %s: %s\n", argv[1], f ? "FAILED" : "ok");
  return f ? 1 : 0;
}
