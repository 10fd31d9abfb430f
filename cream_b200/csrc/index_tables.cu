// Integer index tables of the relative-position path (host code; bit-exact contracts).
//
//   cream_autoformer_rel_index_host : AutoFormer/model/module/multihead_super.py:40-59
//   cream_irpe_bucket_ids_host      : iRPE/DeiT-with-iRPE/irpe.py:19-52 (piecewise_index),
//                                     :131-247 (methods), :291-415 (bucket ids + skip)
// The tables are built once per (grid, config) on the host and uploaded; the kernels
// only ever gather through them.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace {

// piecewise_index for one integer-valued offset, evaluated like the reference does on
// an int64 tensor: the log branch runs in fp32, round-half-even, clip(max=beta), then a
// truncating cast (irpe.py:34-51).
inline int piecewise_index_int(long x, float alpha, float beta, float gamma_over_alpha_log,
                               float beta_minus_alpha) {
  const long ax = x < 0 ? -x : x;
  if (static_cast<float>(ax) <= alpha) return static_cast<int>(x);
  const float sign = x > 0 ? 1.0f : -1.0f;
  float y = alpha + logf(static_cast<float>(ax) / alpha) / gamma_over_alpha_log * beta_minus_alpha;
  y = nearbyintf(y);  // default rounding mode = half to even, as torch.round
  if (y > beta) y = beta;
  return static_cast<int>(sign * y);  // truncation toward zero, as Tensor.to(long)
}

// float-valued input (euclidean method): round() inside the alpha band (irpe.py:45-47).
inline int piecewise_index_float(float x, float alpha, float beta, float gamma_over_alpha_log,
                                 float beta_minus_alpha) {
  const float ax = fabsf(x);
  if (ax <= alpha) return static_cast<int>(nearbyintf(x));
  const float sign = x > 0.f ? 1.0f : (x < 0.f ? -1.0f : 0.0f);
  float y = alpha + logf(ax / alpha) / gamma_over_alpha_log * beta_minus_alpha;
  y = nearbyintf(y);
  if (y > beta) y = beta;
  return static_cast<int>(sign * y);
}

}  // namespace

extern "C" int cream_autoformer_rel_index_host(int grid, int max_rel, int32_t* idx_v,
                                               int32_t* idx_h) {
  if (grid <= 0 || max_rel < 0 || idx_v == nullptr || idx_h == nullptr) return CREAM_ERR_ARG;
  const int n = grid * grid + 1;
  for (int q = 0; q < n; ++q) {
    for (int k = 0; k < n; ++k) {
      int v = 0, h = 0;  // row 0 / column 0 (cls token) are padded with 0
      if (q > 0 && k > 0) {
        const int qi = q - 1, ki = k - 1;
        int dv = ki / grid - qi / grid;
        int dh = ki % grid - qi % grid;
        dv = dv < -max_rel ? -max_rel : (dv > max_rel ? max_rel : dv);
        dh = dh < -max_rel ? -max_rel : (dh > max_rel ? max_rel : dh);
        v = dv + max_rel + 1;
        h = dh + max_rel + 1;
      }
      idx_v[q * n + k] = v;
      idx_h[q * n + k] = h;
    }
  }
  return CREAM_OK;
}

extern "C" int cream_irpe_bucket_ids_host(int method, int height, int width, int skip, double alpha_d,
                                          double beta_d, double gamma_d, int32_t* out,
                                          int* num_buckets) {
  if (height <= 0 || width <= 0 || skip < 0 || out == nullptr || num_buckets == nullptr)
    return CREAM_ERR_ARG;
  const float alpha = static_cast<float>(alpha_d), beta = static_cast<float>(beta_d);
  const float lg = static_cast<float>(std::log(gamma_d / alpha_d));
  const float bma = static_cast<float>(beta_d - alpha_d);
  const int beta_int = static_cast<int>(beta_d);
  const int S = 2 * beta_int + 1;
  int nb;
  switch (method) {
    case 3: nb = S * S; break;
    case 0: case 1: case 41: case 42: nb = S; break;
    default: return CREAM_ERR_UNSUPPORTED;
  }
  const int L = height * width, n = skip + L;
  for (int a = 0; a < L; ++a) {
    const int ra = a / width, ca = a % width;
    for (int b = 0; b < L; ++b) {
      const int rb = b / width, cb_ = b % width;
      const long dr = ra - rb, dc = ca - cb_;  // pos1 - pos2 (irpe.py:339-342)
      int id;
      switch (method) {
        case 3: {
          const int r = piecewise_index_int(dr, alpha, beta, lg, bma) + beta_int;
          const int c = piecewise_index_int(dc, alpha, beta, lg, bma) + beta_int;
          id = r * S + c;
          break;
        }
        case 0: {
          const float dis = nearbyintf(sqrtf(static_cast<float>(dr * dr + dc * dc)));
          id = piecewise_index_float(dis, alpha, beta, lg, bma) + beta_int;
          break;
        }
        case 1: id = piecewise_index_int(dr * dr + dc * dc, alpha, beta, lg, bma) + beta_int; break;
        case 41: id = piecewise_index_int(dr, alpha, beta, lg, bma) + beta_int; break;
        default: id = piecewise_index_int(dc, alpha, beta, lg, bma) + beta_int; break;
      }
      out[static_cast<size_t>(skip + a) * n + (skip + b)] = id;
    }
  }
  if (skip > 0) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        if (i < skip || j < skip) out[static_cast<size_t>(i) * n + j] = nb;
    nb += 1;
  }
  *num_buckets = nb;
  return CREAM_OK;
}
