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

// One block per 4096-element chunk of one segment (`blocks` maps block -> (segment, first element)),
// four consecutive elements per thread and iteration (16-byte accesses where the segment allows it),
// the bf16 shadow written in the same pass.
constexpr int kAdamChunk = 4096;
struct AdamBlock { int32_t seg; int32_t first; };

__global__ void __launch_bounds__(256)
adamw_kernel(cream_adamw_seg* __restrict__ segs, const AdamBlock* __restrict__ blocks, const int32_t* __restrict__ active,
             float lr, float beta1, float beta2, float eps) {
  const AdamBlock blk = blocks[blockIdx.x];
  if (active != nullptr && active[blk.seg] == 0) return;
  const cream_adamw_seg sg = segs[blk.seg];
  const int step = sg.step + 1;
  // bias corrections in double, as torch.optim.AdamW computes them on the host (1 - 0.999^step needs it)
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  const float inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  const float decay = 1.0f - lr * sg.weight_decay;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(sg.shadow);
  const int n = static_cast<int>(sg.numel);
  const int end = min(n, blk.first + kAdamChunk);
  const bool vec = (sg.cols & 3) == 0 && (blk.first & 3) == 0 && (sg.shadow_ld & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(sg.p) | reinterpret_cast<uintptr_t>(sg.g) | reinterpret_cast<uintptr_t>(sg.m) |
                     reinterpret_cast<uintptr_t>(sg.v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(sg.shadow) & 7) == 0;
  auto update = [&](float g, float& p, float& m, float& v) {
    p *= decay;
    m = fmaf(beta1, m, omb1 * g);
    v = fmaf(beta2, v, omb2 * g * g);
    p -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  };
  auto shadow_row = [&](int r) { return sg.qkv_group_rows > 0 ? (r % 3) * sg.qkv_group_rows + r / 3 : r; };   // 3j+i -> i*R+j
  if (vec) {
    for (int i = blk.first + 4 * threadIdx.x; i < end; i += 4 * blockDim.x) {
      const float4 g = *reinterpret_cast<const float4*>(sg.g + i);
      float4 p = *reinterpret_cast<const float4*>(sg.p + i);
      float4 m = *reinterpret_cast<const float4*>(sg.m + i);
      float4 v = *reinterpret_cast<const float4*>(sg.v + i);
      update(g.x, p.x, m.x, v.x); update(g.y, p.y, m.y, v.y); update(g.z, p.z, m.z, v.z); update(g.w, p.w, m.w, v.w);
      *reinterpret_cast<float4*>(sg.p + i) = p;
      *reinterpret_cast<float4*>(sg.m + i) = m;
      *reinterpret_cast<float4*>(sg.v + i) = v;
      if (sh != nullptr) {
        const int r = i / sg.cols, c = i - r * sg.cols;        // cols % 4 == 0: the four elements share a row
        uint2 o;
        o.x = pack_bf16x2(p.x, p.y);
        o.y = pack_bf16x2(p.z, p.w);
        *reinterpret_cast<uint2*>(sh + static_cast<int64_t>(shadow_row(r)) * sg.shadow_ld + c) = o;
      }
    }
  } else {
    for (int i = blk.first + threadIdx.x; i < end; i += blockDim.x) {
      float p = sg.p[i], m = sg.m[i], v = sg.v[i];
      update(sg.g[i], p, m, v);
      sg.p[i] = p; sg.m[i] = m; sg.v[i] = v;
      if (sh != nullptr) {
        const int r = i / sg.cols, c = i - r * sg.cols;
        sh[static_cast<int64_t>(shadow_row(r)) * sg.shadow_ld + c] = __float2bfloat16_rn(p);
      }
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

extern "C" int cream_adamw_step(cream_adamw_seg* segs_dev, const int32_t* active_dev, int n_segs,
                                const int32_t* blocks_dev, int n_blocks, float lr, float beta1, float beta2, float eps,
                                void* stream_) {
  using namespace cb;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(segs_dev && blocks_dev && n_segs > 0 && n_blocks > 0, "bad args");
  adamw_kernel<<<n_blocks, 256, 0, s>>>(segs_dev, reinterpret_cast<const AdamBlock*>(blocks_dev), active_dev, lr, beta1, beta2, eps);
  int rc = check_last("adamw_kernel");
  if (rc) return rc;
  adamw_bump_kernel<<<ceil_div(n_segs, 256), 256, 0, s>>>(segs_dev, active_dev, n_segs);
  return check_last("adamw_bump_kernel");
}

extern "C" int cream_adamw_chunk(void) { return cb::kAdamChunk; }
