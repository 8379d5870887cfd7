// fwb_host.h — host-side helpers: error reporting for the C ABI, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace fwb {

// Thread-local last-error string, surfaced through fwb_last_error().
char* last_error_buf();
void set_error(const char* fmt, ...);

#define FWB_CHECK(cond, ...)      \
  do {                            \
    if (!(cond)) {                \
      fwb::set_error(__VA_ARGS__); \
      return FWB_ERR_INVALID;     \
    }                             \
  } while (0)

#define FWB_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      fwb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return FWB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

enum { FWB_OK = 0, FWB_ERR_INVALID = 1, FWB_ERR_CUDA = 2, FWB_ERR_UNSUPPORTED = 3 };

// bf16 tensor map with up to 4 dims.  dims[0] is the contiguous dimension.
// strides_bytes[i] is the byte stride of dims[i+1] (rank-1 entries).  Returns FWB_OK or an error code.
int make_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swz);

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// SM count of the CURRENT device (cached per device ordinal: a process may drive several GPUs).
inline int num_sms() {
  static int n[64] = {0};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a function: remember per device ordinal that it
// has been set (a process that drives several GPUs would otherwise fail its first > 48 KB launch on the second device).
struct AttrOnce {
  uint64_t done[4] = {0, 0, 0, 0};
  bool need(int dev) {
    if (dev < 0 || dev >= 256) return true;
    const bool n = !((done[dev >> 6] >> (dev & 63)) & 1);
    done[dev >> 6] |= 1ull << (dev & 63);
    return n;
  }
};

}  // namespace fwb
