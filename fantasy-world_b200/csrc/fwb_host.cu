// fwb_host.cu — error buffer, driver entry-point lookup for cuTensorMapEncodeTiled (no libcuda link
// dependency, so the library also loads on a GPU-less box), version/introspection entry points.
#include <stdarg.h>

#include "../../include/fwb200.h"
#include "fwb_host.h"

namespace fwb {

static thread_local char g_err[1024];
char* last_error_buf() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return FWB_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", gptr);
    return FWB_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) {
      gstr[i] = strides_bytes[i];
      if (gstr[i] % 16 != 0) {
        set_error("TMA stride %llu (dim %d) is not a multiple of 16 bytes", (unsigned long long)gstr[i], i + 1);
        return FWB_ERR_INVALID;
      }
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return FWB_ERR_CUDA;
  }
  return FWB_OK;
}

}  // namespace fwb

extern "C" {

const char* fwb_last_error(void) { return fwb::last_error_buf(); }

int fwb_abi_version(void) { return FWB_ABI_VERSION; }

int fwb_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    fwb::set_error("no CUDA device visible");
    return 0;
  }
  cudaDeviceProp p;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  if (p.major != 10) {
    fwb::set_error("device is sm_%d%d; this library is built for sm_100a only", p.major, p.minor);
    return 0;
  }
  return 1;
}

}  // extern "C"
