// Host-side cache of TMA tensor maps.  The sampled-subnet slice (embed dim, heads,
// mlp ratio) only changes the extents of a map over the FULL supernet tensor, so a
// supernet needs a few dozen maps in total; they are encoded once and reused.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace cb {

bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("CREAM_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                              const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn resolve_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeFn>(p);
    } else {
      fprintf(stderr, "cream_b200: cuTensorMapEncodeTiled not available from the driver\n");
    }
  }
  return fn;
}

struct Key {
  uint64_t v[16];
  bool operator==(const Key& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) {
      h ^= x;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

std::mutex g_mu;
std::unordered_map<Key, CUtensorMap*, KeyHash> g_cache;

// Callers receive a pointer into a per-thread ring of COPIES, never into the cache itself, so
// evicting the cache cannot invalidate a map another entry point is about to launch with.
constexpr int kRing = 32;  // >= 2x the maps any single C-ABI call requests
struct alignas(64) MapSlot { CUtensorMap m; };
thread_local MapSlot t_ring[kRing];
thread_local unsigned t_ring_pos = 0;

const CUtensorMap* hand_out(const CUtensorMap* cached) {
  MapSlot& s = t_ring[t_ring_pos++ % kRing];
  std::memcpy(&s.m, cached, sizeof(CUtensorMap));
  return &s.m;
}

size_t elem_bytes(CUtensorMapDataType t) {
  switch (t) {
    case CU_TENSOR_MAP_DATA_TYPE_BFLOAT16:
    case CU_TENSOR_MAP_DATA_TYPE_FLOAT16: return 2;
    case CU_TENSOR_MAP_DATA_TYPE_FLOAT32: return 4;
    case CU_TENSOR_MAP_DATA_TYPE_UINT8: return 1;
    default: return 4;
  }
}

}  // namespace

const CUtensorMap* get_tensor_map(const void* base, CUtensorMapDataType dtype, int rank,
                                  const uint64_t* dims, const uint64_t* strides_elems,
                                  const uint32_t* box, CUtensorMapSwizzle swizzle) {
  Key key{};
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = (static_cast<uint64_t>(dtype) << 32) | (static_cast<uint64_t>(rank) << 8) |
             static_cast<uint64_t>(swizzle);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[7 + i] = strides_elems[i];
    key.v[12 + (i >> 1)] |= static_cast<uint64_t>(box[i]) << (32 * (i & 1));
  }
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) return hand_out(it->second);

  EncodeFn enc = resolve_encode();
  if (enc == nullptr) return nullptr;
  const size_t eb = elem_bytes(dtype);
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstrides[i - 1] = strides_elems[i] * eb;
  }
  CUtensorMap* m = nullptr;
  if (posix_memalign(reinterpret_cast<void**>(&m), 64, sizeof(CUtensorMap)) != 0) return nullptr;
  CUresult r = enc(m, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims,
                   gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "cream_b200: cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu] "
            "box [%u %u %u]\n",
            static_cast<int>(r), rank, (unsigned long long)dims[0],
            (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
            box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    free(m);
    return nullptr;
  }
  if (g_cache.size() > 16384) {  // bounded; weights need a few dozen maps, activation buffers recycle
    for (auto& kv : g_cache) free(kv.second);
    g_cache.clear();
  }
  g_cache.emplace(key, m);
  return hand_out(m);
}

}  // namespace cb
