// rpe_index forward / backward for sm_100a.
//
// Replaces rpe_index_forward_gpu_kernel / rpe_index_backward_gpu_kernel
// (iRPE/DeiT-with-iRPE/rpe_ops/rpe_index_cuda.cu:24-52).  Both are pure HBM-bound
// byte movers: Y[b,h,i,j] = input[b,h,i,index[i,j]] and its scatter-add adjoint.
//
// One warp owns one (b,h,i) row at a time (persistent grid, 148 x k CTAs):
//   fwd: the row of `input` (<= num_buckets values, arbitrary strides) is staged in the
//        warp's slice of shared memory, then lanes stream index[i,:] (L2 resident,
//        shared by every (b,h)) and write Y[i,:] fully coalesced.  No integer divides
//        per element (the reference kernel does three per element).
//   bwd: lanes stream grad_output[i,:] coalesced and reduce into a per-warp shared-memory
//        bucket histogram (shared atomics, <= num_buckets addresses), which is then
//        added into grad_input[i,:] with one coalesced read-modify-write.  The reference
//        issues one GLOBAL atomic per (b,h,i,j) instead.
#include "common.cuh"

namespace cb {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxBucketsSmem = 1024;  // per-warp staging capacity (elements)

template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v);
template <> __device__ __forceinline__ float to_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double to_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_acc(typename Acc<T>::type v);
template <> __device__ __forceinline__ float from_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double from_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rpe_index_fwd_kernel(const T* __restrict__ in, const int32_t* __restrict__ index,
                     T* __restrict__ out, int64_t rows, int H, int Lq, int Lk, int nb, int64_t s0,
                     int64_t s1, int64_t s2, int64_t s3, int use_smem) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* srow = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * nb;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + warp;
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const int i = static_cast<int>(r % Lq);
    const int64_t bh = r / Lq;
    const int h = static_cast<int>(bh % H);
    const int64_t b = bh / H;
    const T* src = in + b * s0 + h * s1 + i * s2;
    const int32_t* irow = index + static_cast<int64_t>(i) * Lk;
    T* dst = out + r * Lk;
    if (use_smem) {
      for (int k = lane; k < nb; k += 32) srow[k] = src[k * s3];
      __syncwarp();
      for (int j = lane; j < Lk; j += 32) dst[j] = srow[__ldg(irow + j)];
      __syncwarp();
    } else {
      for (int j = lane; j < Lk; j += 32) dst[j] = __ldg(src + __ldg(irow + j) * s3);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rpe_index_bwd_kernel(T* __restrict__ grad_in, const T* __restrict__ grad_out,
                     const int32_t* __restrict__ index, int64_t rows, int Lq, int Lk, int nb) {
  using A = typename Acc<T>::type;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  A* hist = reinterpret_cast<A*>(smem_raw) + static_cast<size_t>(warp) * nb;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + warp;
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const int i = static_cast<int>(r % Lq);
    const int32_t* irow = index + static_cast<int64_t>(i) * Lk;
    const T* g = grad_out + r * Lk;
    for (int k = lane; k < nb; k += 32) hist[k] = A(0);
    __syncwarp();
    for (int j = lane; j < Lk; j += 32) atomicAdd(&hist[__ldg(irow + j)], to_acc<T>(g[j]));
    __syncwarp();
    T* gi = grad_in + r * nb;
    for (int k = lane; k < nb; k += 32) gi[k] = from_acc<T>(to_acc<T>(gi[k]) + hist[k]);
    __syncwarp();
  }
}

template <typename T>
int launch_fwd(const void* in, const int32_t* index, void* out, int B, int H, int Lq, int Lk, int nb,
               int64_t s0, int64_t s1, int64_t s2, int64_t s3, cudaStream_t stream) {
  const int64_t rows = static_cast<int64_t>(B) * H * Lq;
  const int use_smem = nb <= kMaxBucketsSmem ? 1 : 0;
  const size_t smem = use_smem ? sizeof(T) * nb * kWarpsPerBlock : 0;
  const int64_t want = ceil_div64(rows, kWarpsPerBlock);
  const int grid = static_cast<int>(std::min<int64_t>(want, kNumSMs * 8));
  rpe_index_fwd_kernel<T><<<grid, kWarpsPerBlock * 32, smem, stream>>>(
      static_cast<const T*>(in), index, static_cast<T*>(out), rows, H, Lq, Lk, nb, s0, s1, s2, s3,
      use_smem);
  return check_last("rpe_index_fwd_kernel");
}

template <typename T>
int launch_bwd(void* gin, const void* gout, const int32_t* index, int B, int H, int Lq, int Lk,
               int nb, cudaStream_t stream) {
  using A = typename Acc<T>::type;
  const int64_t rows = static_cast<int64_t>(B) * H * Lq;
  const size_t smem = sizeof(A) * nb * kWarpsPerBlock;
  if (smem > 48 * 1024) {
    fprintf(stderr, "cream_b200: rpe_index_bwd: num_buckets %d too large\n", nb);
    return CREAM_ERR_UNSUPPORTED;
  }
  const int64_t want = ceil_div64(rows, kWarpsPerBlock);
  const int grid = static_cast<int>(std::min<int64_t>(want, kNumSMs * 8));
  rpe_index_bwd_kernel<T><<<grid, kWarpsPerBlock * 32, smem, stream>>>(
      static_cast<T*>(gin), static_cast<const T*>(gout), index, rows, Lq, Lk, nb);
  return check_last("rpe_index_bwd_kernel");
}

}  // namespace
}  // namespace cb

extern "C" const char* cream_rpe_index_version(void) { return "1.2.0"; }

extern "C" int cream_rpe_index_fwd(const void* input, const int32_t* index, void* out, int B, int H,
                                   int Lq, int Lk, int nb, int64_t s0, int64_t s1, int64_t s2,
                                   int64_t s3, int dtype, void* stream_) {
  using namespace cb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(B >= 0 && H >= 0 && Lq >= 0 && Lk >= 0 && nb > 0, "negative size");
  if (static_cast<int64_t>(B) * H * Lq * Lk == 0) return CREAM_OK;
  CB_REQUIRE(input && index && out, "null pointer");
  switch (dtype) {
    case CREAM_DTYPE_F32: return launch_fwd<float>(input, index, out, B, H, Lq, Lk, nb, s0, s1, s2, s3, stream);
    case CREAM_DTYPE_F64: return launch_fwd<double>(input, index, out, B, H, Lq, Lk, nb, s0, s1, s2, s3, stream);
    case CREAM_DTYPE_F16: return launch_fwd<__half>(input, index, out, B, H, Lq, Lk, nb, s0, s1, s2, s3, stream);
    case CREAM_DTYPE_BF16: return launch_fwd<__nv_bfloat16>(input, index, out, B, H, Lq, Lk, nb, s0, s1, s2, s3, stream);
    default: return CREAM_ERR_UNSUPPORTED;
  }
}

extern "C" int cream_rpe_index_bwd(void* grad_input, const void* grad_output, const int32_t* index,
                                   int B, int H, int Lq, int Lk, int nb, int dtype, void* stream_) {
  using namespace cb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(B >= 0 && H >= 0 && Lq >= 0 && Lk >= 0 && nb > 0, "negative size");
  if (static_cast<int64_t>(B) * H * Lq * Lk == 0) return CREAM_OK;
  CB_REQUIRE(grad_input && grad_output && index, "null pointer");
  switch (dtype) {
    case CREAM_DTYPE_F32: return launch_bwd<float>(grad_input, grad_output, index, B, H, Lq, Lk, nb, stream);
    case CREAM_DTYPE_F64: return launch_bwd<double>(grad_input, grad_output, index, B, H, Lq, Lk, nb, stream);
    case CREAM_DTYPE_F16: return launch_bwd<__half>(grad_input, grad_output, index, B, H, Lq, Lk, nb, stream);
    case CREAM_DTYPE_BF16: return launch_bwd<__nv_bfloat16>(grad_input, grad_output, index, B, H, Lq, Lk, nb, stream);
    default: return CREAM_ERR_UNSUPPORTED;
  }
}
