// arcflow_b200 — host-side utilities shared by all entry points: error string, TMA descriptor
// construction (driver entry point resolved at run time, so the library links without libcuda and
// loads on a box with no GPU), SM count.
#include <stdarg.h>
#include <string.h>

#include <atomic>

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
}  // namespace

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
  return AFB_OK;
}

}  // namespace afb
