// Loss and optimizer kernels of the training step (HBM-bound byte movers).
//
//   cream_xent_fwd_bwd   mean cross-entropy + its gradient in one pass over the logits
//                        (criterion(outputs, targets) + the first backward node,
//                        AutoFormer/supernet_engine.py:74, 96-99)
//   cream_adamw_step     AdamW over EVERY parameter of the supernet in one launch, fused with the
//                        refresh of the bf16 weight shadows the GEMMs read (the reference re-casts each
//                        sampled slice on every forward under autocast, supernet_engine.py:65).
//                        torch.optim.AdamW semantics: decoupled weight decay, bias correction from a
//                        per-parameter step count, parameters without a gradient this step (identity
//                        layers: grad None under find_unused_parameters, supernet_train.py:288) skipped.
#include "common.cuh"

namespace cb {
namespace {

// one warp per sample; C classes streamed twice (max / sum-exp, then gradient)
__global__ void __launch_bounds__(128)
xent_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ targets, float* __restrict__ loss,
            float* __restrict__ dlogits, int64_t ldd, int B, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* row = logits + static_cast<int64_t>(warp) * ld;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += __expf(row[c] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const int64_t t = targets[warp];
  const float inv = 1.0f / sum, invB = 1.0f / B;
  if (dlogits != nullptr) {
    float* drow = dlogits + static_cast<int64_t>(warp) * ldd;
    for (int c = lane; c < C; c += 32) drow[c] = (__expf(row[c] - mx) * inv - (c == t ? 1.f : 0.f)) * invB;
  }
  if (lane == 0 && loss != nullptr && t >= 0 && t < C) atomicAdd(loss, (mx + __logf(sum) - row[t]) * invB);
}

// grid.y = segment; grid.x strides over the segment's elements (4 per thread where aligned)
__global__ void __launch_bounds__(256)
adamw_kernel(cream_adamw_seg* __restrict__ segs, const int32_t* __restrict__ active, float lr, float beta1, float beta2,
             float eps) {
  const cream_adamw_seg sg = segs[blockIdx.y];
  if (active != nullptr && active[blockIdx.y] == 0) return;
  const int step = sg.step + 1;
  // bias corrections in double, as torch.optim.AdamW computes them on the host (1 - 0.999^step needs it)
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  const float inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  const float decay = 1.0f - lr * sg.weight_decay;
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(sg.shadow);
  const int64_t n = sg.numel;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float g = sg.g[i];
    float p = sg.p[i] * decay;
    const float m = fmaf(beta1, sg.m[i], (1.0f - beta1) * g);
    const float v = fmaf(beta2, sg.v[i], (1.0f - beta2) * g * g);
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p -= step_size * (m / denom);
    sg.p[i] = p; sg.m[i] = m; sg.v[i] = v;
    if (sh != nullptr) {
      const int64_t r = i / sg.cols, c = i - r * sg.cols;
      int64_t dr = r;
      if (sg.qkv_group_rows > 0) dr = (r % 3) * sg.qkv_group_rows + r / 3;   // reference row 3j+i -> shadow row i*R+j
      sh[dr * sg.shadow_ld + c] = __float2bfloat16_rn(p);
    }
  }
  // sg.step is advanced by adamw_bump_kernel AFTER this grid (every block must read the same value)
}

__global__ void adamw_bump_kernel(cream_adamw_seg* __restrict__ segs, const int32_t* __restrict__ active, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (active == nullptr || active[i] != 0)) segs[i].step += 1;
}

}  // namespace
}  // namespace cb

extern "C" int cream_xent_fwd_bwd(const float* logits, int64_t ld, const int64_t* targets, float* loss, float* dlogits,
                                  int64_t ldd, int B, int C, void* stream_) {
  using namespace cb;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(logits && targets && B > 0 && C > 0, "bad args");
  if (loss) CB_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), s));
  xent_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(logits, ld, targets, loss, dlogits, ldd, B, C);
  return check_last("xent_kernel");
}

extern "C" int cream_adamw_step(cream_adamw_seg* segs_dev, const int32_t* active_dev, int n_segs, int64_t max_numel,
                                float lr, float beta1, float beta2, float eps, void* stream_) {
  using namespace cb;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(segs_dev && n_segs > 0 && n_segs <= 65535 && max_numel > 0, "bad args");
  // grid.x: enough blocks for the largest segment at ~8 elements per thread, capped at 2 waves
  const int gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(max_numel, 256 * 8), 2 * kNumSMs)));
  dim3 grid(gx, n_segs);
  adamw_kernel<<<grid, 256, 0, s>>>(segs_dev, active_dev, lr, beta1, beta2, eps);
  int rc = check_last("adamw_kernel");
  if (rc) return rc;
  adamw_bump_kernel<<<ceil_div(n_segs, 256), 256, 0, s>>>(segs_dev, active_dev, n_segs);
  return check_last("adamw_bump_kernel");
}
