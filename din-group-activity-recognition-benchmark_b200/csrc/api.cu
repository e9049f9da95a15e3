// api.cu — library-level entry points of libdin_sm100.so: versioning, thread-local error string,
// device queries and the tensor-map encoder shared by the TMA kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include <cudaTypedefs.h>

#include "din_common.cuh"

namespace {
thread_local char g_err[512] = {0};

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

// cuTensorMapEncodeTiled is a pure function of its arguments (the 128-byte descriptor holds the encoded address, extents,
// strides, box and swizzle; no driver state), so encoded maps are cached on exactly those arguments (SURVEY.md §8b:
// "cached CUtensorMaps keyed by pointer + shape").  A plan re-issues the same (pointer, shape) launches every step --
// workspaces are re-used and the caching allocator returns the same blocks -- so the two driver calls per
// convolution launch become two hash lookups.
struct TmapKey {
  uint64_t v[16];
  bool operator==(const TmapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; h ^= h >> 29; }
    return static_cast<size_t>(h);
  }
};
constexpr size_t kTmapCacheMax = 8192;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
std::mutex g_tmaps_mu;
uint64_t g_tmap_hits = 0, g_tmap_misses = 0;

void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) {
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
}
}  // namespace

int din_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int din_encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle) {
  std::call_once(g_encode_once, resolve_encode);
  if (!g_encode) return din_set_error(DIN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (rank < 1 || rank > 4) return din_set_error(DIN_ERR_INVALID_ARG, "din_encode_tmap: rank %d", rank);
  TmapKey key{};
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = (static_cast<uint64_t>(dtype) << 32) | (static_cast<uint64_t>(swizzle) << 8) | static_cast<uint64_t>(rank);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = i ? strides_bytes[i] : 0;
    key.v[10 + i] = (static_cast<uint64_t>(box[i]) << 32) | elem_strides[i];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      ++g_tmap_hits;
      return DIN_OK;
    }
  }
  // strides_bytes[0] is the (implied) element stride; the driver takes rank-1 outer strides.
  CUresult r = g_encode(out, dtype, static_cast<cuuint32_t>(rank), base,
                        reinterpret_cast<const cuuint64_t*>(dims),
                        reinterpret_cast<const cuuint64_t*>(strides_bytes + 1),
                        reinterpret_cast<const cuuint32_t*>(box),
                        reinterpret_cast<const cuuint32_t*>(elem_strides), CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return din_set_error(DIN_ERR_CUDA,
                         "cuTensorMapEncodeTiled failed (CUresult %d): rank %d base %p dims [%llu %llu %llu %llu] "
                         "box [%u %u %u %u]",
                         static_cast<int>(r), rank, base, (unsigned long long)dims[0],
                         (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                         (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
                         rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    if (g_tmaps.size() >= kTmapCacheMax) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
    ++g_tmap_misses;
  }
  return DIN_OK;
}

int din_num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

extern "C" {

// 2: + uint8 ingest, loss / metrics, backward entry points
// 3: + batched weight packing, dgrad with fused ReLU backward, ResNet-18 backward helpers, batch-statistics BatchNorm
// 4: + din_roi_align_nhwc_f16_f32out, din_tmap_cache_stats (tensor maps cached per pointer + shape)
int din_abi_version(void) { return 5; }

const char* din_last_error_string(void) { return g_err; }

int din_tmap_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  std::lock_guard<std::mutex> lk(g_tmaps_mu);
  if (hits) *hits = g_tmap_hits;
  if (misses) *misses = g_tmap_misses;
  return static_cast<int>(g_tmaps.size());
}

int din_device_sm_count(void) {
  int n = din_num_sms();
  if (n <= 0) return din_set_error(DIN_ERR_CUDA, "no CUDA device available");
  return n;
}

}  // extern "C"
