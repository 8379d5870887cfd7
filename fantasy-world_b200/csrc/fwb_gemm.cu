// fwb_gemm.cu — persistent tcgen05 GEMMs with a fused column-affine epilogue.
//
//   out[M,N] = epilogue( A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major, nn.Linear layout) )
//
// Replaces every nn.Linear / 1x1x1 Conv3d call site on the hot path: DiT q/k/v/o
// (wan_video_dit.py:166-169), cross-attn projections (:216-225), FFN (:274-275), adapter projections
// (fusion/layer/block.py:340-346), VGGT qkv/proj (vggt/layers/attention.py:42,46), Mlp (vggt/layers/mlp.py:29-31),
// camera adapter MLPs (camera_control.py:27-51), projection_head (vggt/models/vggt.py:32).
//
// Two kernels, both persistent (static tile schedule), TMA (SWIZZLE_128B, 64-wide K blocks) -> smem ring ->
// tcgen05.mma with fp32 accumulators double-buffered in TMEM -> 8 epilogue warps (tcgen05.ld, bias / activation /
// column affine / residual, 16-byte global stores) overlapping the next tile's MMAs:
//   gemm2_kernel      CTA pair (cluster of 2, cta_group::2): 256 x 256 output tile per pair, each CTA stages its 128 rows
//                     of A and its 128 rows of W per K block (32 KB / stage instead of 48 KB -> 1.5x less L2->SM traffic,
//                     6 stages).  Used when N % 256 == 0 and M >= 256.
//   gemm_kernel<BN>   single CTA, 128 x BN tile (BN = 128 | 256): every other shape (ragged N, tiny M).
// Roofline: tensor-pipe bound, 2*M*N*K FLOP.
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

constexpr int kGemmThreads = 320;  // 10 warps: TMA, MMA, 8 epilogue
constexpr int BM = 128;
constexpr int BK = 64;

template <int ACT>
__device__ __forceinline__ float act_fn(float y) {
  if constexpr (ACT == FWB_ACT_GELU_TANH) {
    // 0.5*y*(1+tanh(sqrt(2/pi)*(y+0.044715*y^3)))  (nn.GELU(approximate='tanh'), wan_video_dit.py:274)
    float u = 0.7978845608028654f * (y + 0.044715f * y * y * y);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * y * (1.f + t);
  } else if constexpr (ACT == FWB_ACT_GELU_ERF) {
    // 0.5*y*(1+erf(y/sqrt(2)))  (nn.GELU(), vggt/layers/block.py:35); erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7)
    float x = fabsf(y) * 0.7071067811865476f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, x, 1.f)));
    float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    float e = fmaf(-poly, fast_exp2(-x * x * 1.4426950408889634f), 1.f);
    return 0.5f * y * (1.f + copysignf(e, y));
  } else if constexpr (ACT == FWB_ACT_RELU) {
    return fmaxf(y, 0.f);
  } else if constexpr (ACT == FWB_ACT_SILU) {
    return y / (1.f + __expf(-y));
  } else {
    return y;
  }
}

// Epilogue of one 32-row x 32-column chunk owned by one warp.  Thread `lane` arrives holding the fp32 accumulators of row
// row0+lane (TMEM 32x32b layout).  Column-wise math (bias, activation, affine) is done in that layout; the result is then
// transposed through a 4 KB per-warp shared-memory stage (16-byte chunks XOR-swizzled by row) so that the residual read
// and the store are coalesced: 4 (bf16) or 8 (fp32) lanes cover one row's contiguous 64 / 128 bytes instead of every lane
// touching a different row (the row-strided 16-byte stores were the bottleneck of the short-K shapes, profiles/r01_gemm2.md).
template <int ACT>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], int row0, int lane, int colbase,
                                               float* stage) {
  if (colbase >= p.N) return;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  float y[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) y[c] = __uint_as_float(v[c]);
  if (colbase + 32 <= p.N) {
    // Residual prefetch, in the coalesced (post-transpose) thread mapping used by the stores below: the loads are issued before
    // any math, so their HBM latency overlaps the column math and the transposition instead of sitting, once per row group,
    // between the transposition and every store (the short-K shapes — K <= 1152, fp32 geometry stream — were bound by exactly
    // that chain: ~8000 clk of epilogue per 256x256 tile against 8192 clk of MMA).
    float rres[32];
    if (p.resid) {
      if (p.out_f32) {
        const int piece = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row0 + it * 4 + (lane >> 3);
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < p.M) {
            const int col = colbase + piece * 4;
            if (p.resid_f32) {
              b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.resid) + (size_t)grow * p.resid_ld + col);
            } else {
              const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) + (size_t)grow * p.resid_ld + col);
              b = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                              __uint_as_float(u.y & 0xFFFF0000u));
            }
          }
          rres[4 * it] = b.x; rres[4 * it + 1] = b.y; rres[4 * it + 2] = b.z; rres[4 * it + 3] = b.w;
        }
      } else {
        const int piece = lane & 3;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int grow = row0 + it * 8 + (lane >> 2);
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (grow < p.M) {
            const int col = colbase + piece * 8;
            if (p.resid_f32) {
              const float* rp = reinterpret_cast<const float*>(p.resid) + (size_t)grow * p.resid_ld + col;
              b0 = *reinterpret_cast<const float4*>(rp);
              b1 = *reinterpret_cast<const float4*>(rp + 4);
            } else {
              const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) + (size_t)grow * p.resid_ld + col);
              b0 = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                               __uint_as_float(u.y & 0xFFFF0000u));
              b1 = make_float4(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xFFFF0000u), __uint_as_float(u.w << 16),
                               __uint_as_float(u.w & 0xFFFF0000u));
            }
          }
          rres[8 * it] = b0.x; rres[8 * it + 1] = b0.y; rres[8 * it + 2] = b0.z; rres[8 * it + 3] = b0.w;
          rres[8 * it + 4] = b1.x; rres[8 * it + 5] = b1.y; rres[8 * it + 6] = b1.z; rres[8 * it + 7] = b1.w;
        }
      }
    }
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
    if constexpr (ACT != FWB_ACT_NONE) {
#pragma unroll
      for (int c = 0; c < 32; ++c) y[c] = act_fn<ACT>(y[c]);
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
    // ---- transpose through the per-warp stage: row `lane`, 16-byte chunk c lives at chunk position c ^ (lane & 7)
    float4* srow = reinterpret_cast<float4*>(stage + lane * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) srow[c ^ (lane & 7)] = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
    __syncwarp();
    if (p.out_f32) {
      // 8 lanes per row (4 fp32 each), 4 rows per pass
      const int piece = lane & 7;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const int grow = row0 + r;
        float4 t = reinterpret_cast<const float4*>(stage + r * 32)[piece ^ (r & 7)];
        if (grow < p.M) {
          const int col = colbase + piece * 4;
          if (p.resid) {
            t.x += rres[4 * it]; t.y += rres[4 * it + 1]; t.z += rres[4 * it + 2]; t.w += rres[4 * it + 3];
          }
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)grow * p.out_ld + col) = t;
        }
      }
    } else {
      // 4 lanes per row (8 bf16 each), 8 rows per pass
      const int piece = lane & 3;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        const int grow = row0 + r;
        const float4* sr = reinterpret_cast<const float4*>(stage + r * 32);
        float4 a = sr[(2 * piece) ^ (r & 7)], b4 = sr[(2 * piece + 1) ^ (r & 7)];
        if (grow < p.M) {
          const int col = colbase + piece * 8;
          if (p.resid) {
            a.x += rres[8 * it]; a.y += rres[8 * it + 1]; a.z += rres[8 * it + 2]; a.w += rres[8 * it + 3];
            b4.x += rres[8 * it + 4]; b4.y += rres[8 * it + 5]; b4.z += rres[8 * it + 6]; b4.w += rres[8 * it + 7];
          }
          uint4 w;
          w.x = pack_bf16x2(a.x, a.y);
          w.y = pack_bf16x2(a.z, a.w);
          w.z = pack_bf16x2(b4.x, b4.y);
          w.w = pack_bf16x2(b4.z, b4.w);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)grow * p.out_ld + col) = w;
        }
      }
    }
    __syncwarp();   // the stage is reused by the next chunk
  } else if (row_ok) {
    // ragged N tail: scalar path
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
      const int col = colbase + c;
      if (col >= p.N) break;
      float t = __uint_as_float(v[c]);
      if (p.bias) t += p.bias[col];
      if (p.round_flags & 1) t = bf16_round(t);
      if (ACT != FWB_ACT_NONE) {
        t = act_fn<ACT>(t);
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
                         : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[(size_t)row * p.resid_ld + col]);
      }
      if (p.out_f32)
        reinterpret_cast<float*>(p.out)[(size_t)row * p.out_ld + col] = t;
      else
        reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)row * p.out_ld + col] = __float2bfloat16_rn(t);
    }
  }
}

constexpr int kStageFloats = 32 * 32;           // per-warp transpose stage (4 KB)
constexpr int kEpiStageBytes = 8 * kStageFloats * 4;

// =====================================================================================================================
// single-CTA kernel: 128 x BN tiles
// =====================================================================================================================
template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kStageBytes = BM * BK * 2 + BN * BK * 2;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + 1024;
};

template <int BN, int ACT>
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
    if (elect_one()) {  // ------------------------------ TMA producer ------------------------------
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
    if (elect_one()) {  // ------------------------------ MMA issuer --------------------------------
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
            umma_ss(d_tmem, make_smem_desc(sa + k * 32, 0, 1024, SWZ_128B), make_smem_desc(sb + k * 32, 0, 1024, SWZ_128B),
                    idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
        }
        tc_commit(&tfull_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const uint32_t e = warp - 2;   // 0..7
    const uint32_t q = warp & 3;   // TMEM lane quadrant this warp may access
    const uint32_t half = e >> 2;  // column half
    const uint32_t lane = lane_id();
    constexpr int COLS_PER_WARP = BN / 2;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int tm = tile / p.num_n_tiles, tn = tile % p.num_n_tiles;
      const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      float* stage = reinterpret_cast<float*>(smem + ST * Cfg::kStageBytes) + e * kStageFloats;
#pragma unroll 1
      for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((q * 32) << 16) + acc * BN + half * COLS_PER_WARP + c0, v);
        tmem_ld_wait();
        epilogue_chunk<ACT>(p, v, tm * BM + q * 32, lane, tn * BN + half * COLS_PER_WARP + c0, stage);
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

// =====================================================================================================================
// CTA-pair kernel: 256 x 256 tiles, cta_group::2
// =====================================================================================================================
constexpr int kStages2 = 6;
constexpr int kStageBytes2 = BM * BK * 2 + 128 * BK * 2;  // this CTA's 128 rows of A + its 128 rows of W
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + kEpiStageBytes + 1024;

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int ST = kStages2;
  constexpr int BN = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[ST], empty_bar[ST], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const uint32_t warp = warp_id_uniform();
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // 256 x 256 tiles
  const int nkb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_bar[s], 1);   // leader: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[s], 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 16);  // 8 epilogue warps in each CTA of the pair (leader's copy is the one waited on)
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc_2sm(&tmem_base_s, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocation of both CTAs visible before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {  // ------------------------------ TMA producer (both CTAs) ------------------------------
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tm = tile / p.num_n_tiles, tn = tile % p.num_n_tiles;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % ST, ph = (it / ST) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * kStageBytes2;
          uint8_t* sb = sa + BM * BK * 2;
          const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * kStageBytes2);
          tma_load_2d_2sm(sa, &tmA, lead_full, kb * BK, tm * 256 + rank * 128);
          tma_load_2d_2sm(sb, &tmB, lead_full, kb * BK, tn * 256 + rank * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {  // ------------------------------ MMA issuer (leader only) ----------------
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      uint32_t it = 0, lt = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
        const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % ST, ph = (it / ST) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes2);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_ss_2sm(d_tmem, make_smem_desc(sa + k * 32, 0, 1024, SWZ_128B),
                        make_smem_desc(sb + k * 32, 0, 1024, SWZ_128B), idesc, (kb | k) != 0);
          }
          tc_commit_2sm(&empty_bar[s], 3);  // frees the stage in BOTH CTAs
        }
        tc_commit_2sm(&tfull_bar[acc], 3);  // accumulators (each CTA's own 128 rows) complete
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ----------------------------------
    const uint32_t e = warp - 2;
    const uint32_t q = warp & 3;
    const uint32_t half = e >> 2;
    const uint32_t lane = lane_id();
    constexpr int COLS_PER_WARP = BN / 2;
    uint32_t lt = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
      const int tm = tile / p.num_n_tiles, tn = tile % p.num_n_tiles;
      const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      float* stage = reinterpret_cast<float*>(smem + ST * kStageBytes2) + e * kStageFloats;
#pragma unroll 1
      for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((q * 32) << 16) + acc * BN + half * COLS_PER_WARP + c0, v);
        tmem_ld_wait();
        epilogue_chunk<ACT>(p, v, tm * 256 + rank * 128 + q * 32, lane, tn * BN + half * COLS_PER_WARP + c0, stage);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's smem / barriers stay alive until every MMA and remote arrive has landed
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
template <int BN, int ACT>
int launch1(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static AttrOnce once;
  if (once.need(current_device()))
    FWB_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = num_sms();
  if (grid > tiles) grid = tiles;
  gemm_kernel<BN, ACT><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  FWB_CUDA(cudaGetLastError());
  return FWB_OK;
}

template <int ACT>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  static AttrOnce once;
  if (once.need(current_device()))
    FWB_CUDA(cudaFuncSetAttribute(gemm2_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int pairs = num_sms() / 2;
  if (pairs > tiles) pairs = tiles;
  gemm2_kernel<ACT><<<2 * pairs, kGemmThreads, kSmemBytes2, stream>>>(tmA, tmB, p);
  FWB_CUDA(cudaGetLastError());
  return FWB_OK;
}

template <int ACT>
int dispatch_shape(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  if (mode == 2) return launch2<ACT>(tmA, tmB, p, stream);
  if (mode == 1) return launch1<256, ACT>(tmA, tmB, p, stream);
  return launch1<128, ACT>(tmA, tmB, p, stream);
}

int g_force_mode = -1;  // tests: 0 = 1-CTA BN128, 1 = 1-CTA BN256, 2 = CTA pair, -1 = automatic

}  // namespace

extern "C" int fwb_gemm_set_mode(int mode) {
  g_force_mode = mode;
  return FWB_OK;
}

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
  FWB_CHECK(ep->act >= FWB_ACT_NONE && ep->act <= FWB_ACT_SILU, "gemm: unknown activation %d", ep->act);

  int mode;
  if (g_force_mode >= 0) mode = g_force_mode;
  else if (N % 256 == 0 && M >= 256) mode = 2;
  else mode = (N % 256 == 0 || N > 1024) ? 1 : 0;
  const int TM = mode == 2 ? 256 : BM;
  const int BN = mode == 0 ? 128 : 256;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_tiles = (M + TM - 1) / TM;
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
    uint32_t box[2] = {BK, (uint32_t)(mode == 2 ? 128 : BN)};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  switch (ep->act) {
    case FWB_ACT_GELU_TANH: return dispatch_shape<FWB_ACT_GELU_TANH>(mode, tmA, tmB, p, stream);
    case FWB_ACT_GELU_ERF: return dispatch_shape<FWB_ACT_GELU_ERF>(mode, tmA, tmB, p, stream);
    case FWB_ACT_RELU: return dispatch_shape<FWB_ACT_RELU>(mode, tmA, tmB, p, stream);
    case FWB_ACT_SILU: return dispatch_shape<FWB_ACT_SILU>(mode, tmA, tmB, p, stream);
    default: return dispatch_shape<FWB_ACT_NONE>(mode, tmA, tmB, p, stream);
  }
}
