// fwb_bringup.cu — single-CTA tcgen05 micro-test used to pin descriptor encodings on real hardware.
//
// D[128 x N] (fp32) = A[128 x K] * B, with
//   A: bf16 row-major [128][K]  — staged by TMA (SW128 boxes of 64 columns) or written to TMEM as packed bf16x2
//   B: bf16 [N][K] (K-major)    — TMA boxes {64 k, N rows}
//      or   [K][N] (MN-major)   — TMA boxes {64 n, K rows}  (the layout V has in attention)
// The descriptor fields that are easy to get wrong (LBO / SBO / per-K-step advance) are runtime parameters so
// one GPU session can sweep candidates.  Not on any product path.
#include "../../include/fwb200.h"
#include "fwb_common.cuh"
#include "fwb_host.h"

using namespace fwb;

struct BringupParams {
  int N, K;
  int a_in_tmem;
  int b_mn_major;
  uint32_t a_lbo, a_sbo, a_kadv;  // bytes
  uint32_t b_lbo, b_sbo, b_kadv;  // bytes
  uint32_t b_box_stride;          // bytes between consecutive TMA boxes of B in smem
  uint32_t a_tmem_kadv;           // TMEM columns per K=16 step when A is in TMEM
  int mma_n;                      // N of the MMA instruction (<= N; N = width of the loaded B / read-back D)
};

__global__ void __launch_bounds__(128, 1)
bringup_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __nv_bfloat16* __restrict__ Ag, float* __restrict__ D, BringupParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int K = p.K, N = p.N;
  uint8_t* sA = smem;                       // K/64 boxes of [128 x 128B]
  uint8_t* sB = smem + (K / 64) * 16384;    // boxes of B
  const uint32_t warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t d_tmem = tmem_base;          // columns [0, N)
  const uint32_t a_tmem = tmem_base + 256;    // columns [256, 256 + K/2)

  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (!p.a_in_tmem) bytes += 128 * K * 2;
    bytes += N * K * 2;
    mbar_arrive_expect_tx(&bar_load, bytes);
    if (!p.a_in_tmem)
      for (int kb = 0; kb < K / 64; ++kb) tma_load_2d(sA + kb * 16384, &tmA, &bar_load, kb * 64, 0);
    if (!p.b_mn_major) {
      for (int kb = 0; kb < K / 64; ++kb) tma_load_2d(sB + kb * p.b_box_stride, &tmB, &bar_load, kb * 64, 0);
    } else {
      for (int nb = 0; nb < N / 64; ++nb) tma_load_2d(sB + nb * p.b_box_stride, &tmB, &bar_load, nb * 64, 0);
    }
  }
  if (p.a_in_tmem) {
    // thread r owns TMEM lane r: write A[r][0..K) as packed bf16 pairs, 16 columns (32 elements) at a time
    const int r = threadIdx.x;
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const __nv_bfloat16* src = Ag + (size_t)r * K + 2 * (c0 + c);
        uint32_t lo = __bfloat16_as_ushort(src[0]);
        uint32_t hi = __bfloat16_as_ushort(src[1]);
        v[c] = lo | (hi << 16);
      }
      tmem_st16(a_tmem + ((warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, p.mma_n, 0, p.b_mn_major ? 1 : 0);
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t bdesc;
      if (!p.b_mn_major) {
        uint32_t addr = smem_u32(sB) + (kk / 4) * p.b_box_stride + (kk % 4) * p.b_kadv;
        bdesc = make_smem_desc(addr, p.b_lbo, p.b_sbo, SWZ_128B);
      } else {
        uint32_t addr = smem_u32(sB) + kk * p.b_kadv;
        bdesc = make_smem_desc(addr, p.b_lbo, p.b_sbo, SWZ_128B);
      }
      if (!p.a_in_tmem) {
        uint32_t addr = smem_u32(sA) + (kk / 4) * 16384 + (kk % 4) * p.a_kadv;
        uint64_t adesc = make_smem_desc(addr, p.a_lbo, p.a_sbo, SWZ_128B);
        umma_ss(d_tmem, adesc, bdesc, idesc, kk > 0);
      } else {
        umma_ts(d_tmem, a_tmem + kk * p.a_tmem_kadv, bdesc, idesc, kk > 0);
      }
    }
    tc_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  {
    const int r = threadIdx.x;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(d_tmem + ((warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) D[(size_t)r * N + c0 + c] = __uint_as_float(v[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static int bringup_impl(const void* A, const void* B, float* D, int N, int K, int a_in_tmem, int b_mn_major, const uint32_t* overrides,
                        int mma_n, cudaStream_t stream);

extern "C" int fwb_bringup_mma(const void* A, const void* B, float* D, int N, int K, int a_in_tmem, int b_mn_major,
                               const uint32_t* overrides /* 8 values or NULL */, cudaStream_t stream) {
  return bringup_impl(A, B, D, N, K, a_in_tmem, b_mn_major, overrides, N, stream);
}

// The PV configuration of the attention (A = P in TMEM, B = V MN-major in two 64-column SWIZZLE_128B boxes) with an MMA N that is
// narrower than the loaded tile: mma_n = 96 on a 128-wide V tile is the native head_dim-96 PV (columns [0, mma_n) of D are defined).
extern "C" int fwb_bringup_mma_pv_n(const void* A, const void* B, float* D, int N, int K, int mma_n, cudaStream_t stream) {
  FWB_CHECK(mma_n > 0 && mma_n <= N && mma_n % 16 == 0, "bringup: mma_n must be a multiple of 16 and <= N");
  return bringup_impl(A, B, D, N, K, 1, 1, nullptr, mma_n, stream);
}

static int bringup_impl(const void* A, const void* B, float* D, int N, int K, int a_in_tmem, int b_mn_major, const uint32_t* overrides,
                        int mma_n, cudaStream_t stream) {
  FWB_CHECK(N == 64 || N == 128 || N == 256, "bringup: N must be 64/128/256");
  FWB_CHECK(K % 64 == 0 && K >= 64 && K <= 256, "bringup: K must be a multiple of 64 in [64,256]");
  BringupParams p;
  p.N = N;
  p.K = K;
  p.a_in_tmem = a_in_tmem;
  p.b_mn_major = b_mn_major;
  p.a_lbo = 0;
  p.a_sbo = 1024;
  p.a_kadv = 32;
  if (!b_mn_major) {
    p.b_lbo = 0;
    p.b_sbo = 1024;
    p.b_kadv = 32;
    p.b_box_stride = N * 128;
  } else {
    p.b_box_stride = K * 128;   // one box = K rows x 128 B
    p.b_lbo = p.b_box_stride;   // stride between 64-wide column blocks
    p.b_sbo = 1024;             // stride between groups of 8 k-rows
    p.b_kadv = 16 * 128;        // 16 k-rows per MMA
  }
  p.a_tmem_kadv = 8;
  p.mma_n = mma_n;
  if (overrides) {
    if (overrides[0] != 0xFFFFFFFFu) p.a_lbo = overrides[0];
    if (overrides[1] != 0xFFFFFFFFu) p.a_sbo = overrides[1];
    if (overrides[2] != 0xFFFFFFFFu) p.a_kadv = overrides[2];
    if (overrides[3] != 0xFFFFFFFFu) p.b_lbo = overrides[3];
    if (overrides[4] != 0xFFFFFFFFu) p.b_sbo = overrides[4];
    if (overrides[5] != 0xFFFFFFFFu) p.b_kadv = overrides[5];
    if (overrides[6] != 0xFFFFFFFFu) p.b_box_stride = overrides[6];
    if (overrides[7] != 0xFFFFFFFFu) p.a_tmem_kadv = overrides[7];
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, 128};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (!b_mn_major) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)N};
    int rc = make_tmap_bf16(&tmB, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    uint64_t str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, (uint32_t)K};
    int rc = make_tmap_bf16(&tmB, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  size_t smem = 1024 + (size_t)(K / 64) * 16384 + (size_t)N * K * 2;
  FWB_CUDA(cudaFuncSetAttribute(bringup_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bringup_mma_kernel<<<1, 128, smem, stream>>>(tmA, tmB, reinterpret_cast<const __nv_bfloat16*>(A), D, p);
  FWB_CUDA(cudaGetLastError());
  return FWB_OK;
}
