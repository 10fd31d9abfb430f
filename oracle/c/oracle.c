/* ORACLE (test infrastructure only — never linked into the product library).
 *
 * Plain-C restatement of the two kernels at the native boundary of the reference:
 *   - rpe_index forward / backward   (iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.cpp:8-73, 82-124)
 *   - the attention core with relative-position terms, per (batch, head)
 *       AutoFormer  : AutoFormer/model/module/multihead_super.py:133-154
 *       iRPE (k/q/v): iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:73-92, irpe.py:585-687
 * written as direct loops with double accumulation.  Pinned through tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Y[b,h,i,j] = input[b,h,i,index[i,j]] ; all contiguous */
void oracle_rpe_index_fwd_f32(const float* in, const int32_t* index, float* out, int B, int H, int Lq,
                              int Lk, int nb) {
  for (long r = 0; r < (long)B * H * Lq; ++r) {
    const int i = (int)(r % Lq);
    for (int j = 0; j < Lk; ++j) out[r * Lk + j] = in[r * nb + index[(long)i * Lk + j]];
  }
}

/* grad_in[b,h,i,index[i,j]] += grad_out[b,h,i,j] ; grad_in caller-zeroed */
void oracle_rpe_index_bwd_f32(float* grad_in, const float* grad_out, const int32_t* index, int B, int H,
                              int Lq, int Lk, int nb) {
  for (long r = 0; r < (long)B * H * Lq; ++r) {
    const int i = (int)(r % Lq);
    for (int j = 0; j < Lk; ++j) grad_in[r * nb + index[(long)i * Lk + j]] += grad_out[r * Lk + j];
  }
}

/* Attention core for ONE (batch, head):  q,k,v (N, D) row-major with row stride ld.
 *   S[i,j] = scale * ( q_i . k_j  +  sum_t q_i . Tk[t][idx_k[t][i,j]] )  + bias[idx_b[i,j]]
 *          + scale_q * sum_t k_j . Tq[t][idx_q[t][j,i]]      (iRPE rpe_q, transposed)
 *   P = softmax_j(S) ;  O[i,:] = sum_j P[i,j] * ( v_j + sum_t Tv[t][idx_v[t][i,j]] )
 * ntk/ntq/ntv = number of (table, index) pairs on K / Q / V (AutoFormer: 2 on K and 2 on V
 * with the vertical and horizontal tables; iRPE: 0 or 1).  Tables are (nb, D) row-major.
 * For iRPE the caller passes q already scaled (scale = 1) exactly like the reference.  */
void oracle_attention_rpe_f64(const float* q, const float* k, const float* v, int ld, int N, int D,
                              double scale, int ntk, const float* const* Tk, const int32_t* const* idx_k,
                              int ntq, const float* const* Tq, const int32_t* const* idx_q, double scale_q,
                              int ntv, const float* const* Tv, const int32_t* const* idx_v,
                              const float* bias, const int32_t* idx_b, float* out, int ldo,
                              float* probs /* optional (N,N) */) {
  double* s = (double*)malloc(sizeof(double) * N);
  for (int i = 0; i < N; ++i) {
    double mx = -1e300;
    for (int j = 0; j < N; ++j) {
      double acc = 0;
      for (int d = 0; d < D; ++d) acc += (double)q[(long)i * ld + d] * k[(long)j * ld + d];
      for (int t = 0; t < ntk; ++t) {
        const float* row = Tk[t] + (long)idx_k[t][(long)i * N + j] * D;
        for (int d = 0; d < D; ++d) acc += (double)q[(long)i * ld + d] * row[d];
      }
      acc *= scale;
      for (int t = 0; t < ntq; ++t) {
        const float* row = Tq[t] + (long)idx_q[t][(long)j * N + i] * D;
        double a2 = 0;
        for (int d = 0; d < D; ++d) a2 += (double)k[(long)j * ld + d] * row[d];
        acc += scale_q * a2;
      }
      if (bias) acc += bias[idx_b[(long)i * N + j]];
      s[j] = acc;
      if (acc > mx) mx = acc;
    }
    double sum = 0;
    for (int j = 0; j < N; ++j) { s[j] = exp(s[j] - mx); sum += s[j]; }
    for (int j = 0; j < N; ++j) { s[j] /= sum; if (probs) probs[(long)i * N + j] = (float)s[j]; }
    for (int d = 0; d < D; ++d) {
      double acc = 0;
      for (int j = 0; j < N; ++j) {
        double val = v[(long)j * ld + d];
        for (int t = 0; t < ntv; ++t) val += Tv[t][(long)idx_v[t][(long)i * N + j] * D + d];
        acc += s[j] * val;
      }
      out[(long)i * ldo + d] = (float)acc;
    }
  }
  free(s);
}
