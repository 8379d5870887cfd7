// fwb_gemm.cu — persistent tcgen05 GEMM with a fused column-affine epilogue.
//
//   out[M,N] = epilogue( A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major, nn.Linear layout) )
//
// Replaces every nn.Linear / 1x1x1 Conv3d call site on the hot path: DiT q/k/v/o
// (wan_video_dit.py:166-169), cross-attn projections (:216-225), FFN (:274-275), adapter projections
// (fusion/layer/block.py:340-346), VGGT qkv/proj (vggt/layers/attention.py:42,46), Mlp (vggt/layers/mlp.py:29-31),
// camera adapter MLPs (camera_control.py:27-51), projection_head (vggt/models/vggt.py:32).
//
// Design (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0      TMA producer: A box {64 k, 128 rows}, W box {64 k, BN rows}, SWIZZLE_128B, ring of ST stages
//   warp 1      MMA issuer:   tcgen05.mma cta_group::1 kind::f16, M=128 N=BN K=16, fp32 accumulators in TMEM,
//                             two accumulator buffers (2 x BN columns) so the epilogue of tile i overlaps tile i+1
//   warps 2..9  epilogue:     tcgen05.ld 32x32b.x32 -> bias/act/scale/shift/residual -> 16-byte global stores
// Roofline: tensor-pipe bound (2*M*N*K FLOP).
#include "../../include/fwb200.h"
#include "fwb_common.cuh"
#include "fwb_host.h"

using namespace fwb;

namespace {

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles;
  const float* bias;
  const float* scale1;
  const float* shift1;
  const float* scale2;
  const void* resid;
  long long resid_ld;
  int resid_f32;
  void* out;
  long long out_ld;
  int out_f32;
  int act;
  int round_flags;
};

constexpr int kGemmThreads = 320;  // 10 warps
constexpr int BM = 128;
constexpr int BK = 64;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kStageBytes = BM * BK * 2 + BN * BK * 2;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
};

__device__ __forceinline__ float act_apply(float y, int act) {
  switch (act) {
    case FWB_ACT_GELU_TANH: {
      // 0.5*y*(1+tanh(sqrt(2/pi)*(y+0.044715*y^3)))  (nn.GELU(approximate='tanh'), wan_video_dit.py:274)
      float u = 0.7978845608028654f * (y + 0.044715f * y * y * y);
      float t;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
      return 0.5f * y * (1.f + t);
    }
    case FWB_ACT_GELU_ERF: {
      // 0.5*y*(1+erf(y/sqrt(2)))  (nn.GELU(), vggt/layers/block.py:35); erf by Abramowitz-Stegun 7.1.26
      float x = fabsf(y) * 0.7071067811865476f;
      float t = __frcp_rn(1.f + 0.3275911f * x);
      float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
      float e = 1.f - poly * fast_exp2(-x * x * 1.4426950408889634f);
      e = copysignf(e, y);
      return 0.5f * y * (1.f + e);
    }
    case FWB_ACT_RELU:
      return fmaxf(y, 0.f);
    case FWB_ACT_SILU:
      return y / (1.f + __expf(-y));
    default:
      return y;
  }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int ST = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[ST], empty_bar[ST], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const uint32_t warp = warp_id_uniform();
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int nkb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);  // one arrive per epilogue warp
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / p.num_n_tiles, tn = tile % p.num_n_tiles;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % ST, ph = (it / ST) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, tm * BM);
          tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, tn * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % ST, ph = (it / ST) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * 32, 0, 1024, SWZ_128B);
            const uint64_t bdesc = make_smem_desc(sb + k * 32, 0, 1024, SWZ_128B);
            umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
        }
        tc_commit(&tfull_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const uint32_t e = warp - 2;          // 0..7
    const uint32_t q = warp & 3;          // TMEM lane quadrant this warp may access
    const uint32_t half = e >> 2;         // column half
    const uint32_t lane = lane_id();
    constexpr int COLS_PER_WARP = BN / 2;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int tm = tile / p.num_n_tiles, tn = tile % p.num_n_tiles;
      const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
        const int colbase = tn * BN + half * COLS_PER_WARP + c0;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((q * 32) << 16) + acc * BN + half * COLS_PER_WARP + c0, v);
        tmem_ld_wait();
        if (colbase >= p.N) continue;
        const bool full_chunk = (colbase + 32 <= p.N);
        float y[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) y[c] = __uint_as_float(v[c]);
        if (full_chunk) {
          if (p.bias) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + colbase + c));
              y[c] += b.x; y[c + 1] += b.y; y[c + 2] += b.z; y[c + 3] += b.w;
            }
          }
          if (p.round_flags & 1) {
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] = bf16_round(y[c]);
          }
          if (p.act != FWB_ACT_NONE) {
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] = act_apply(y[c], p.act);
            if (p.round_flags & 2) {
#pragma unroll
              for (int c = 0; c < 32; ++c) y[c] = bf16_round(y[c]);
            }
          }
          if (p.scale1) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 b = __ldg(reinterpret_cast<const float4*>(p.scale1 + colbase + c));
              y[c] *= b.x; y[c + 1] *= b.y; y[c + 2] *= b.z; y[c + 3] *= b.w;
            }
          }
          if (p.shift1) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 b = __ldg(reinterpret_cast<const float4*>(p.shift1 + colbase + c));
              y[c] += b.x; y[c + 1] += b.y; y[c + 2] += b.z; y[c + 3] += b.w;
            }
          }
          if (p.round_flags & 4) {
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] = bf16_round(y[c]);
          }
          if (p.scale2) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 b = __ldg(reinterpret_cast<const float4*>(p.scale2 + colbase + c));
              y[c] *= b.x; y[c + 1] *= b.y; y[c + 2] *= b.z; y[c + 3] *= b.w;
            }
            if (p.round_flags & 8) {
#pragma unroll
              for (int c = 0; c < 32; ++c) y[c] = bf16_round(y[c]);
            }
          }
          if (row_ok) {
            if (p.resid) {
              if (p.resid_f32) {
                const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.resid) +
                                                                 (size_t)row * p.resid_ld + colbase);
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                  float4 b = r[c / 4];
                  y[c] += b.x; y[c + 1] += b.y; y[c + 2] += b.z; y[c + 3] += b.w;
                }
              } else {
                const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) +
                                                               (size_t)row * p.resid_ld + colbase);
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                  uint4 b = r[c / 8];
                  const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    y[c + 2 * j] += __uint_as_float(w[j] << 16);
                    y[c + 2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                  }
                }
              }
            }
            if (p.out_f32) {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * p.out_ld + colbase);
#pragma unroll
              for (int c = 0; c < 32; c += 4) o[c / 4] = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
            } else {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.out_ld +
                                                  colbase);
#pragma unroll
              for (int c = 0; c < 32; c += 8) {
                uint4 w;
                w.x = pack_bf16x2(y[c], y[c + 1]);
                w.y = pack_bf16x2(y[c + 2], y[c + 3]);
                w.z = pack_bf16x2(y[c + 4], y[c + 5]);
                w.w = pack_bf16x2(y[c + 6], y[c + 7]);
                o[c / 8] = w;
              }
            }
          }
        } else if (row_ok) {
          // ragged N tail: scalar path
          for (int c = 0; c < 32; ++c) {
            const int col = colbase + c;
            if (col >= p.N) break;
            float t = y[c];
            if (p.bias) t += p.bias[col];
            if (p.round_flags & 1) t = bf16_round(t);
            if (p.act != FWB_ACT_NONE) {
              t = act_apply(t, p.act);
              if (p.round_flags & 2) t = bf16_round(t);
            }
            if (p.scale1) t *= p.scale1[col];
            if (p.shift1) t += p.shift1[col];
            if (p.round_flags & 4) t = bf16_round(t);
            if (p.scale2) {
              t *= p.scale2[col];
              if (p.round_flags & 8) t = bf16_round(t);
            }
            if (p.resid) {
              t += p.resid_f32 ? reinterpret_cast<const float*>(p.resid)[(size_t)row * p.resid_ld + col]
                               : __bfloat162float(
                                     reinterpret_cast<const __nv_bfloat16*>(p.resid)[(size_t)row * p.resid_ld + col]);
            }
            if (p.out_f32)
              reinterpret_cast<float*>(p.out)[(size_t)row * p.out_ld + col] = t;
            else
              reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)row * p.out_ld + col] = __float2bfloat16_rn(t);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    FWB_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = num_sms();
  if (grid > tiles) grid = tiles;
  gemm_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  FWB_CUDA(cudaGetLastError());
  return FWB_OK;
}

}  // namespace

extern "C" int fwb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                             const fwb_epilogue_t* ep, cudaStream_t stream) {
  FWB_CHECK(A && W && ep && ep->out, "gemm: null pointer");
  FWB_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  FWB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16-byte TMA rows): K=%d lda=%lld ldw=%lld", K, (long long)lda, (long long)ldw);
  FWB_CHECK(ep->out_ld >= N, "gemm: out_ld < N");
  const bool vec_ok = (N % 8 == 0) && (ep->out_ld % 8 == 0) && (!ep->resid || ep->resid_ld % 8 == 0) &&
                      ((reinterpret_cast<uintptr_t>(ep->out) & 15) == 0) &&
                      (!ep->resid || (reinterpret_cast<uintptr_t>(ep->resid) & 15) == 0);
  FWB_CHECK(vec_ok, "gemm: N, out_ld, resid_ld must be multiples of 8 and pointers 16-byte aligned (N=%d)", N);
  const float* vecs[4] = {ep->bias, ep->scale1, ep->shift1, ep->scale2};
  for (int i = 0; i < 4; ++i)
    FWB_CHECK(!vecs[i] || (reinterpret_cast<uintptr_t>(vecs[i]) & 15) == 0, "gemm: column vector %d not 16-byte aligned", i);

  const int BN = (N % 256 == 0 || N > 1024) ? 256 : 128;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_tiles = (M + BM - 1) / BM;
  p.num_n_tiles = (N + BN - 1) / BN;
  p.bias = ep->bias; p.scale1 = ep->scale1; p.shift1 = ep->shift1; p.scale2 = ep->scale2;
  p.resid = ep->resid; p.resid_ld = ep->resid_ld; p.resid_f32 = ep->resid_dtype == FWB_DT_F32;
  p.out = ep->out; p.out_ld = ep->out_ld; p.out_f32 = ep->out_dtype == FWB_DT_F32;
  p.act = ep->act; p.round_flags = ep->round_flags;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  return BN == 256 ? launch_gemm<256>(tmA, tmB, p, stream) : launch_gemm<128>(tmA, tmB, p, stream);
}
