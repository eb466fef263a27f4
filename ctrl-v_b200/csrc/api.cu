// Library plumbing: thread-local error string, device check, TMA tensor-map encoding through
// the driver entry point (no link-time dependency on libcuda).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

static thread_local char g_err[512] = {0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTRLV_PDL");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;  // opt-in: measured neutral inside the captured step graph
  }
  return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return CTRLV_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return CTRLV_ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return CTRLV_OK;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(
        "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] "
        "strides [%llu %llu %llu] box [%u %u %u %u] base %p",
        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
        (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
        (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], rank > 1 ? box[1] : 0,
        rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return CTRLV_ERR_CUDA;
  }
  return CTRLV_OK;
}

}  // namespace ctrlv

extern "C" const char* ctrlv_last_error(void) { return ctrlv::g_err; }

extern "C" int64_t ctrlv_launch_count(void) { return (int64_t)ctrlv::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ctrlv_version(void) { return "ctrlv_b200 0.1 (sm_100a)"; }

extern "C" int ctrlv_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    ctrlv::set_last_error("no CUDA device visible (%s)", cudaGetErrorString(e));
    return CTRLV_ERR_CUDA;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    ctrlv::set_last_error("device compute capability %d.x is not sm_100 (B200)", major);
    return CTRLV_ERR_UNSUPPORTED;
  }
  return CTRLV_OK;
}
