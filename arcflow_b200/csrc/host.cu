// arcflow_b200 — host-side utilities shared by all entry points: error string, TMA descriptor
// construction (driver entry point resolved at run time, so the library links without libcuda and
// loads on a box with no GPU), SM count.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace afb {

namespace {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

// Descriptor cache: the engine launches the same (pointer, shape, box) combinations every forward (1300+ launches x 4-6
// tensor maps per step), so an encoded CUtensorMap is kept per key instead of calling cuTensorMapEncodeTiled each time.
// A map depends on nothing but its key (the driver call is a pure function of these arguments).
struct TmapKey {
  uint64_t v[12];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : k.v) {
      h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
      h *= 0xD6E8FEB86659FD93ull;
    }
    return size_t(h ^ (h >> 32));
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
std::atomic<uint64_t> g_tmap_hits{0}, g_tmap_misses{0};
constexpr size_t TMAP_CACHE_MAX = 16384;
}  // namespace

void tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  *hits = g_tmap_hits.load(std::memory_order_relaxed);
  *misses = g_tmap_misses.load(std::memory_order_relaxed);
}

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

void count_launch(int n) { g_launches.fetch_add(uint64_t(n), std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_last_error("cuTensorMapEncodeTiled entry point unavailable (%s)",
                     e != cudaSuccess ? cudaGetErrorString(e) : "not found");
      return AFB_ERR_CUDA;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("TMA base pointer %p not 16-byte aligned", base);
    return AFB_ERR_INVALID;
  }
  if (rank < 1 || rank > 5) {
    set_last_error("TMA rank %d out of range", rank);
    return AFB_ERR_INVALID;
  }
  TmapKey key{};
  key.v[0] = reinterpret_cast<uintptr_t>(base);
  key.v[1] = uint64_t(rank);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i] | (uint64_t(box[i]) << 40);
    if (i > 0) key.v[7 + i - 1] = strides_bytes[i - 1];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return AFB_OK;
    }
  }
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_last_error("TMA stride %d (%llu bytes) not a multiple of 16", i,
                       (unsigned long long)gstr[i - 1]);
        return AFB_ERR_INVALID;
      }
    }
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank),
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu)",
                   int(r), rank, (unsigned long long)dims[0],
                   (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0));
    return AFB_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return AFB_OK;
}

}  // namespace afb
