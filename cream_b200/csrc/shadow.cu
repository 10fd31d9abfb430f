// fp32 master -> bf16 shadow casts (one pass per optimizer step over the supernet).
// The reference re-casts every sampled slice on every forward under autocast
// (AutoFormer/supernet_engine.py:65) and re-materialises the interleaved QKV slice with
// torch.cat on every set_sample_config (qkv_super.py:45-51,72-77); here the full tensor
// is cast once and the slice lives in the TMA descriptor.
#include "common.cuh"

namespace cb {
namespace {

// rows x cols, 4 elements per thread along cols; dst row = row_map(src row).
template <bool kQkv>
__global__ void __launch_bounds__(256)
shadow_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                   int64_t cols, int64_t ld_src, int64_t ld_dst, int64_t rows_per_group) {
  const int64_t cols4 = (cols + 3) >> 2;
  const int64_t total = rows * cols4;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = t / cols4;
    const int64_t c = (t - r * cols4) << 2;
    int64_t dr = r;
    if (kQkv) dr = (r % 3) * rows_per_group + r / 3;  // reference row 3*j+i -> shadow row i*R+j
    const float* s = src + r * ld_src + c;
    __nv_bfloat16* d = dst + dr * ld_dst + c;
    if (c + 4 <= cols && ((ld_src & 3) == 0) && ((ld_dst & 3) == 0)) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(s));
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(d) = o;
    } else {
      for (int k = 0; k < 4 && c + k < cols; ++k) d[k] = __float2bfloat16_rn(s[k]);
    }
  }
}

}  // namespace
}  // namespace cb

extern "C" const char* cream_version(void) { return "cream_b200 0.1.0 (sm_100a)"; }

// The library links its own CUDA runtime instance; a host thread that has not touched it yet
// (e.g. an autograd worker thread) must bind the caller's device before raw driver calls
// (cuTensorMapEncodeTiled) or launches are made from it.
extern "C" int cream_bind_device(int device) {
  using namespace cb;
  CB_CUDA_OK(cudaSetDevice(device));
  CB_CUDA_OK(cudaFree(nullptr));
  return CREAM_OK;
}

extern "C" int cream_shadow_cast(const float* src, void* dst, int64_t rows, int64_t cols,
                                 int64_t ld_src, int64_t ld_dst, void* stream_) {
  using namespace cb;
  if (rows * cols == 0) return CREAM_OK;
  CB_REQUIRE(src && dst && rows > 0 && cols > 0, "bad shadow_cast args");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
             "alignment");
  const int64_t work = rows * ((cols + 3) / 4);
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(work, 256), kNumSMs * 16));
  shadow_cast_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      src, static_cast<__nv_bfloat16*>(dst), rows, cols, ld_src, ld_dst, 0);
  return check_last("shadow_cast_kernel");
}

extern "C" int cream_shadow_qkv(const float* src, void* dst, int64_t rows_per_group, int64_t cols,
                                int64_t ld_src, int64_t ld_dst, void* stream_) {
  using namespace cb;
  if (rows_per_group * cols == 0) return CREAM_OK;
  CB_REQUIRE(src && dst && rows_per_group > 0 && cols > 0, "bad shadow_qkv args");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
             "alignment");
  const int64_t rows = 3 * rows_per_group;
  const int64_t work = rows * ((cols + 3) / 4);
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(work, 256), kNumSMs * 16));
  shadow_cast_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      src, static_cast<__nv_bfloat16*>(dst), rows, cols, ld_src, ld_dst, rows_per_group);
  return check_last("shadow_qkv_kernel");
}
