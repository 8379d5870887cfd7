// fwb_attn.cu — non-causal flash attention forward on tcgen05 (sm_100a).
//
// Replaces:
//   flash_attention()                  FantasyWorld/diffsynth_wan21/models/wan_video_dit.py:28-66   (DiT self / cross)
//   F.scaled_dot_product_attention     FantasyWorld/fusion/layer/block.py:598-605                   (adapter, 2 directions)
//   F.scaled_dot_product_attention     FantasyWorld/vggt/layers/attention.py:61                     (VGGT frame / global)
//
// One CTA = 256 query rows (two 128-row Q tiles, ping-pong) x one (batch, head); it streams all KV tiles of 128 keys.
//   warps 0-3   softmax for Q tile 0 (thread r <-> query row r <-> TMEM lane r)
//   warps 4-7   softmax for Q tile 1
//   warp  8     MMA issuer (one elected thread): S_i = Q_i K_j^T (SS), O_i += P_i V_j (TS: P read from TMEM)
//   warp  9     TMA producer: Q once, K/V rings (SWIZZLE_128B boxes of 64 columns)
// TMEM (512 columns): S_0 | S_1 (fp32 128 cols each; P_i (bf16) overwrites the first 64 columns of S_i) | O_0 | O_1.
// Softmax: fp32, exp2 with the scale folded in, lazy rescaling of O (only when the row max grows by > 2^8),
// P rounded to bf16 before PV — the same rounding point as flash-attn / cuDNN SDPA.
// Variants measured on B200 and NOT kept (profiles/r01_attention_variants.md): two softmax threads per row (v2, same speed:
// the exponentials of a tile cost ~1024 clk of MUFU per SM sub-partition whatever the thread count), P handed over in
// four 32-key chunks (v3, slower: four mbarrier wake-ups per tile on the MMA thread cost more than the overlap buys),
// a SCALAR FMA-pipe exp2 polynomial for 25-75 % of the elements (slower: 8 issue slots per element), a speculative single-pass
// softmax and a two-half P hand-off (round 2, profiles/r02_attn_ab_emu{2,3}.log: no gain; removed).  What IS kept from that line
// of work is softmax_exp() below: packed f32x2 arithmetic (FFMA2 / FADD2) and a PACKED exp2 polynomial on a fraction of the pairs.
// A second kernel (attn2, further down) decouples S and P in TMEM and is the default at head_dim 64; both share the tile schedule
// with the key-split tail (attn_tail_merge_kernel) and the split-KV partial outputs.
// head_dim 96 (adapter) runs on the D=128 instance: TMA zero-fills columns 96..127 and QK^T skips the dead K-steps.
// Roofline: tensor-pipe bound, 4*B*H*Lq*Lk*D FLOP.
#include <math.h>

#include "../../include/fwb200.h"
#include "fwb_common.cuh"
#include "fwb_host.h"

using namespace fwb;

namespace {

constexpr int kAttnThreads = 320;
constexpr int BQ = 128;   // rows per Q tile
constexpr int BKV = 128;  // keys per KV tile

struct AttnParams {
  __nv_bfloat16* out;
  long long o_sb, o_sl, o_sh;
  int Lq, Lk, d_real;
  int short_kv;      // 1: one 128-row Q tile per CTA, two CTAs per SM (few KV tiles: prologue / epilogue dominate), nq counts 128-row blocks
  int pv_n;          // N of the PV MMAs: head_dim of the kernel instance, or 96 for head_dim 96 on the 128 instance (native width)
  float scale_log2;
  int accumulate;  // out = bf16(out + bf16(result))  (sum of two attentions sharing q: wan_video_dit.py:197-200)
  // split-KV mode (sequence parallel pipelining): write the subset-normalised result in fp32 and the row log-sum-exp
  // (base 2, scale folded in) so that fwb_attn_merge can combine the partial attentions over disjoint key subsets.
  float* part_out;   // [B, Lq, H, d_real] fp32 contiguous, or nullptr
  float* part_lse;   // [B, H, Lq] fp32, or nullptr
  int H;
  // tile schedule (1-D grid).  A tile = 256 query rows x one (batch, head); tiles [0, n_full) are processed by one CTA each
  // over all keys; each of the remaining `tail` tiles (the last, partly filled wave) is split over S CTAs along the keys
  // (stream-K for the tail) which write (normalised fp32 O, lse) to the workspace; attn_tail_merge_kernel combines them.
  int nq;            // query tiles per (batch, head)
  int n_full;        // tiles processed unsplit
  int S;             // key splits per tail tile (>= 2 when there is a split tail, else 1)
  float* ws_out;     // [tail, S, 256, d_real]
  float* ws_lse;     // [tail, S, 256]
};

// Optional timeline instrumentation (tools/attn_trace.py builds a separate library with -DFWB_ATTN_TRACE; never in libfwb200.so):
// SM clock stamps of one CTA's softmax warps and MMA thread for the first 64 KV tiles.
#ifdef FWB_ATTN_TRACE
__device__ long long g_trace[9][64][8];
__device__ int g_trace_cta = 300;
__device__ __forceinline__ long long trace_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define TRACE(slot, j, ev)                                                                             \
  do {                                                                                                 \
    if ((int)blockIdx.x == g_trace_cta && (j) < 64 && lane_id() == 0) g_trace[slot][j][ev] = trace_clock(); \
  } while (0)
#define TRACE1(slot, j, ev)                                                                  \
  do {                                                                                       \
    if ((int)blockIdx.x == g_trace_cta && (j) < 64) g_trace[slot][j][ev] = trace_clock();    \
  } while (0)
#else
#define TRACE(slot, j, ev) do { } while (0)
#define TRACE1(slot, j, ev) do { } while (0)
#endif

__device__ int g_attn1_pingpong = 0;   // v1 kernel: MUFU ping-pong of the two Q tiles' exp2 phases (fwb_attn_set_mufu_pingpong(1, .))

template <int D>
struct AttnCfg {
  static constexpr int kBoxes = D / 64;                   // 64-column TMA boxes per tile
  static constexpr int kTileBytes = kBoxes * BQ * 128;    // Q / K / V tile bytes (128 rows)
  static constexpr int kStages = (D == 128) ? 2 : 4;      // K ring depth == V ring depth
  static constexpr int kSmemBytes = 2 * kTileBytes + 2 * kStages * kTileBytes + 1024;
  static constexpr uint32_t kColS = 0;                    // S_i at columns i*128
  static constexpr uint32_t kColO = 256;                  // O_i at columns 256 + i*D
};

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2: one issue slot for two elements) ------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fadd2_rm(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Exponentials of one S tile row held in registers: p[c] = 2^(v[c] * sl2 - m), packed to bf16 pairs in pk[], returns sum_c p[c]
// (fp32, before the bf16 rounding, as flash-attn / cuDNN accumulate the row sum).
// Budget per SM sub-partition and KV tile (two softmax warps, 256 elements): MUFU.EX2 runs at 4 lanes/clk = 8 clk per warp
// instruction, i.e. 2048 clk for all exponentials — as much as the tile's MMA work at head_dim 128 and twice that at head_dim 64
// (tools/ubench/mufu.cu).  So (1) the scale/subtract and the row sum are done on PAIRS (FFMA2 / FADD2: half the issue slots), and
// (2) POLY of every 8 pairs take 2^x from the FMA / ALU pipes instead of the MUFU: Cody-Waite split with a round-toward-minus-
// infinity magic add (FADD2.RM: mantissa low bits = floor(x)), degree-3 minimax polynomial for 2^frac on [0,1) (max rel. error
// 8.6e-5, 45x below the bf16 rounding P receives next), exponent re-inserted with one LEA (shift-left-add) per element.
// Cost per pair: MUFU path 1 FFMA2 + 2 MUFU + 1 FADD2 + 1 F2FP; polynomial path 2 FMNMX + 4 FFMA2 + 3 FADD2 + 2 LEA + 1 FADD2 + 1 F2FP.
template <int N, int POLY>
__device__ __forceinline__ float softmax_exp(const uint32_t (&v)[N], float sl2, float m, uint32_t (&pk)[N / 2]) {
  const uint64_t S2 = pack2(sl2, sl2), M2 = pack2(-m, -m);
  const uint64_t MAGIC = pack2(12582912.f, 12582912.f);   // 1.5 * 2^23
  const uint64_t C3 = pack2(0.07706704f, 0.07706704f), C2 = pack2(0.22764499f, 0.22764499f), C1 = pack2(0.69511679f, 0.69511679f),
                 C0 = pack2(1.0f, 1.0f);
  uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < N; c += 2) {
    const int pair = c / 2;
    uint64_t x = ffma2(pack2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), S2, M2);
    float p0, p1;
    if (POLY > 0 && ((pair & 7) * POLY) % 8 < POLY) {
      float x0, x1;
      unpack2(x, x0, x1);
      x = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
      const uint64_t r = fadd2_rm(x, MAGIC);
      const uint64_t f = fsub2(x, fsub2(r, MAGIC));     // frac in [0, 1)
      uint64_t q = ffma2(f, C3, C2);
      q = ffma2(q, f, C1);
      q = ffma2(q, f, C0);
      float q0, q1, r0, r1;
      unpack2(q, q0, q1);
      unpack2(r, r0, r1);
      p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
      p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
    } else {
      float x0, x1;
      unpack2(x, x0, x1);
      p0 = fast_exp2(x0);
      p1 = fast_exp2(x1);
    }
    if (pair & 1) acc1 = fadd2(acc1, pack2(p0, p1));
    else acc0 = fadd2(acc0, pack2(p0, p1));
    pk[pair] = pack_bf16x2(p0, p1);
  }
  float s0, s1;
  unpack2(fadd2(acc0, acc1), s0, s1);
  return s0 + s1;
}

// MC (multicast pairs): the kernel runs as clusters of two CTAs that work on adjacent 256-row query blocks of the SAME (batch, head).
// They walk the K/V tiles in lock-step and share every tile: each CTA issues the TMA load of ONE of the tile's two 64-column boxes
// with `.multicast::cluster`, so a K/V tile crosses the L2 -> SM fabric once per pair instead of once per CTA (the kernel re-reads
// every K/V tile from L2 for each of the 128 query blocks of a head: 86 GB per DiT self-attention, profiles/r02_attn_d128.md), and a
// stage is refilled only when both CTAs' MMAs have released it (empty barriers count 2, commits multicast to both CTAs).
// NQ = Q tiles (128 rows each) per CTA.  2: the long-sequence configuration (two tiles ping-pong on one tensor pipe).  1: the
// short-key configuration — half the TMEM (S | O = 256 columns), one K and one V stage, 192 threads — so that TWO CTAs are resident
// per SM: with 2 - 13 KV tiles per CTA the prologue (barriers, TMEM, Q / first K load) and the epilogue are up to half of a CTA's
// life, and a second resident CTA runs its main loop under them (and plays the part of the second Q tile in between).
template <int D, int POLY, bool MC, int NQ>
__device__ __forceinline__ void attn_fwd_body(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p) {
  using Cfg = AttnCfg<D>;
  static_assert(NQ == 1 || NQ == 2, "one or two Q tiles per CTA");
  static_assert(!MC || NQ == 2, "multicast pairs use the two-tile configuration");
  static_assert(!MC || Cfg::kBoxes == 2, "multicast pairs: one 64-column box per CTA (head_dim 128 instance)");
  const uint32_t crank = MC ? cluster_ctarank() : 0;
  constexpr int ST = (NQ == 2) ? Cfg::kStages : Cfg::kStages / 2;   // NQ 1: 1 stage at head_dim 128, 2 at 64 (two CTAs share the SM's smem)
  constexpr uint32_t kMmaWarp = 4 * NQ, kTmaWarp = 4 * NQ + 1;
  constexpr uint32_t kColO = NQ * 128;                                // S_i at i*128, O_i at NQ*128 + i*D
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [NQ][kTileBytes]
  uint8_t* sK = smem + NQ * Cfg::kTileBytes;            // [ST][kTileBytes]
  uint8_t* sV = sK + ST * Cfg::kTileBytes;              // [ST][kTileBytes]
  __shared__ uint64_t q_full[2], k_full[ST], k_empty[ST], v_full[ST], v_empty[ST], s_full[2], p_full[2], o_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float pp_scratch[1 + 256];   // MUFU ping-pong pins (see attn2_kernel)

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const uint32_t pp_addr = smem_u32(pp_scratch);
  if (threadIdx.x == 0) pp_scratch[0] = 0.f;
  // tile schedule (see AttnParams): decoded from blockIdx wherever it is needed instead of being kept live across the KV loop
  auto decode = [&](int& tile, int& split, int& nsplit) {
    tile = blockIdx.x; split = 0; nsplit = 1;
    if (tile >= p.n_full) {
      const int r = tile - p.n_full;
      tile = p.n_full + r / p.S;
      split = r % p.S;
      nsplit = p.S;
    }
  };
  int qblock, head, batch, kv0, n_kv, j_ragged;
  {
    int tile, split, nsplit;
    decode(tile, split, nsplit);
    qblock = tile % p.nq; head = (tile / p.nq) % p.H; batch = tile / (p.nq * p.H);
    const int n_kv_all = (p.Lk + BKV - 1) / BKV;
    kv0 = (int)((long long)split * n_kv_all / nsplit);       // this CTA's KV tiles: [kv0, kv0 + n_kv)
    n_kv = (int)((long long)(split + 1) * n_kv_all / nsplit) - kv0;
    j_ragged = n_kv_all - 1 - kv0;                           // local index of the (possibly) partly filled last KV tile
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrive per softmax warp
      mbar_init(&o_full[i], 1);
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], MC ? 2 : 1);   // MC: released by the MMA threads of both CTAs of the pair
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], MC ? 2 : 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&tmem_base_s, 256 * NQ);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (MC) cluster_sync_all();   // the peer's barriers exist before any multicast load / commit targets them
  const uint32_t tmem_base = tmem_base_s;

  if (warp == kTmaWarp) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      for (int i = 0; i < NQ; ++i) {
        mbar_arrive_expect_tx(&q_full[i], Cfg::kTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_4d(sQ + i * Cfg::kTileBytes + b * 16384, &tmQ, &q_full[i], b * 64, head, (qblock * NQ + i) * BQ,
                      batch);
      }
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t s = j % ST, ph = (j / ST) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], Cfg::kTileBytes);
        if constexpr (MC) {   // this CTA's box for both CTAs; the other box arrives from the peer (it may land before the expect_tx above)
          tma_load_4d_mc(sK + s * Cfg::kTileBytes + crank * 16384, &tmK, &k_full[s], crank * 64, head, (kv0 + j) * BKV, batch, 3);
        } else {
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_4d(sK + s * Cfg::kTileBytes + b * 16384, &tmK, &k_full[s], b * 64, head, (kv0 + j) * BKV, batch);
        }
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], Cfg::kTileBytes);
        if constexpr (MC) {
          tma_load_4d_mc(sV + s * Cfg::kTileBytes + crank * 16384, &tmV, &v_full[s], crank * 64, head, (kv0 + j) * BKV, batch, 3);
        } else {
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_4d(sV + s * Cfg::kTileBytes + b * 16384, &tmV, &v_full[s], b * 64, head, (kv0 + j) * BKV, batch);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------ MMA issuer --------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(BQ, p.pv_n, 0, 1);  // B = V is MN-major (d contiguous); N = 96 skips the zero-padded columns
      const int ksteps_qk = (p.d_real + 15) / 16;
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);

      auto issue_qk = [&](int i, uint32_t ks) {
        const uint32_t qa = q_addr + i * Cfg::kTileBytes, ka = k_addr + ks * Cfg::kTileBytes;
        const uint32_t d_tmem = tmem_base + Cfg::kColS + i * 128;
        for (int kk = 0; kk < ksteps_qk; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(d_tmem, make_smem_desc(qa + off, 0, 1024, SWZ_128B), make_smem_desc(ka + off, 0, 1024, SWZ_128B),
                  idesc_qk, kk > 0);
        }
      };
      auto issue_pv = [&](int i, uint32_t vs, bool acc) {
        const uint32_t va = v_addr + vs * Cfg::kTileBytes;
        const uint32_t d_tmem = tmem_base + kColO + i * D;
        const uint32_t a_tmem = tmem_base + Cfg::kColS + i * 128;
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // 16 kv rows per step = 2048 B; LBO = stride between the 64-column d boxes, SBO = 8 kv rows
          umma_ts(d_tmem, a_tmem + kk * 8, make_smem_desc(va + kk * 2048, 16384, 1024, SWZ_128B), idesc_pv,
                  acc || kk > 0);
        }
      };

      auto release = [&](uint64_t* bar) {   // K / V stage consumed: tell the producer(s) that fill it
        if constexpr (MC) tc_commit_mc(bar, 3);
        else tc_commit(bar);
      };
      mbar_wait(&q_full[0], 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      tc_commit(&s_full[0]);
      if constexpr (NQ == 2) {
        mbar_wait(&q_full[1], 0);
        tc_fence_after();
        issue_qk(1, 0);
        tc_commit(&s_full[1]);
      }
      release(&k_empty[0]);

      for (int j = 0; j < n_kv; ++j) {
        const uint32_t vs = j % ST, vph = (j / ST) & 1;
        for (int i = 0; i < NQ; ++i) {
          TRACE1(8, j, i * 3 + 0);
          mbar_wait(&p_full[i], j & 1);
          TRACE1(8, j, i * 3 + 1);
          if (i == 0) mbar_wait(&v_full[vs], vph);
          tc_fence_after();
          issue_pv(i, vs, j > 0);
          if (i == NQ - 1) release(&v_empty[vs]);
          if (j + 1 < n_kv) {
            const uint32_t ks = (j + 1) % ST, kph = ((j + 1) / ST) & 1;
            if (i == 0) {
              mbar_wait(&k_full[ks], kph);
              tc_fence_after();
            }
            issue_qk(i, ks);  // overwrites S_i/P_i: ordered after PV_i(j) by in-order MMA execution
            tc_commit(&s_full[i]);
            if (i == NQ - 1) release(&k_empty[ks]);
          } else {
            tc_commit(&o_full[i]);
          }
          TRACE1(8, j, i * 3 + 2);
        }
      }
    }
  } else {
    // ------------------------------------ softmax + epilogue ------------------------------
    const int i = warp >> 2;                 // Q tile handled by this warpgroup
    const uint32_t quad = warp & 3;          // TMEM lane quadrant
    const uint32_t lane_off = (quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + Cfg::kColS + i * 128 + lane_off;
    const uint32_t o_tmem = tmem_base + kColO + i * D + lane_off;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY;  // running (stale-tolerant) row max in scaled log2 units
    float l_sum = 0.f;
    // optional MUFU ping-pong between warp w (Q tile 0) and warp w+4 (Q tile 1), as in attn2_kernel
    const bool pingpong = NQ == 2 && g_attn1_pingpong != 0;
    const uint32_t bar_mine = 1 + quad + 4 * i, bar_other = 1 + quad + 4 * (1 - i);
    if (pingpong && i == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");

    for (int j = 0; j < n_kv; ++j) {
      TRACE(warp, j, 0);
      mbar_wait(&s_full[i], j & 1);
      TRACE(warp, j, 1);
      tc_fence_after();
      uint32_t v[128];
      tmem_ld32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld32(s_tmem + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
      tmem_ld32(s_tmem + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
      tmem_ld_wait();
      TRACE(warp, j, 2);

      uint32_t pk[64];
      if (j == j_ragged) {
        const int valid = p.Lk - (kv0 + j) * BKV;
        if (valid < BKV) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= valid) v[c] = 0xFF800000u;  // -inf
        }
      }
      // (8 independent max chains instead of 4 measured no shorter: the phase is not bound by the FMNMX3 dependency chain)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(v[c]));
        mx1 = fmaxf(mx1, __uint_as_float(v[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[c + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(v[c + 3]));
      }
      const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
      const bool need = m_new > m_used + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float m_next = fmaxf(m_used, m_new);
        const float alpha = fast_exp2(m_used - m_next);  // m_used = -inf on the first tile -> 0
        l_sum *= alpha;
        m_used = m_next;
        if (j > 0) {
          // rescale this row of O.  PV_i(j-1) has completed: it was issued before S_i(j), whose commit we waited on.
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(o_tmem + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st16(o_tmem + c0, o);
          }
        }
      }
      TRACE(warp, j, 3);
      if (pingpong) {
        float z;
        asm volatile("bar.sync %1, 64;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(z) : "r"(bar_mine), "r"(pp_addr) : "memory");
        m_used += z;
      }
      const float blk_sum = softmax_exp<128, POLY>(v, sl2, m_used, pk);
      if (pingpong)
        asm volatile("st.volatile.shared.f32 [%0], %1;\n\tbar.arrive %2, 64;" ::"r"(pp_addr + 4 + 4 * threadIdx.x), "f"(blk_sum),
                     "r"(bar_other)
                     : "memory");
      l_sum += blk_sum;
      TRACE(warp, j, 4);
      tmem_st32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[i]);
      TRACE(warp, j, 5);
    }

    if (pingpong && i == 0) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");   // absorb tile 1's last arrive

    // epilogue: O / l  -> bf16 -> global
    mbar_wait(&o_full[i], 0);
    tc_fence_after();
    const int row = (qblock * NQ + i) * BQ + quad * 32 + lane;
    const float inv_l = 1.f / l_sum;
    __nv_bfloat16* orow = p.out + (long long)batch * p.o_sb + (long long)row * p.o_sl + (long long)head * p.o_sh;
    const int row_in_tile = i * BQ + quad * 32 + lane;
    float* ws_row = nullptr;
    int tile, split, nsplit;
    decode(tile, split, nsplit);
    if (nsplit > 1) {
      const long long slot = (long long)(tile - p.n_full) * p.S + split;
      ws_row = p.ws_out + (slot * (2 * BQ) + row_in_tile) * p.d_real;
      p.ws_lse[slot * (2 * BQ) + row_in_tile] = m_used + log2f(l_sum);
    } else if (p.part_lse && row < p.Lq) {
      p.part_lse[((long long)batch * p.H + head) * p.Lq + row] = m_used + log2f(l_sum);
    }
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c0, o);
      tmem_ld_wait();
      if (ws_row) {
        if (c0 < p.d_real) {
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(ws_row + c0 + c) =
                make_float4(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l, __uint_as_float(o[c + 2]) * inv_l,
                            __uint_as_float(o[c + 3]) * inv_l);
        }
      } else if (p.part_out) {
        if (row < p.Lq && c0 < p.d_real) {
          float* prow = p.part_out + (((long long)batch * p.Lq + row) * p.H + head) * p.d_real + c0;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(prow + c) = make_float4(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l,
                                                               __uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
        }
      } else if (row < p.Lq && c0 < p.d_real) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          float y[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) y[t] = __uint_as_float(o[c + t]) * inv_l;
          if (p.accumulate) {
            const uint4 prev = *reinterpret_cast<const uint4*>(orow + c0 + c);
            const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              y[2 * t] = bf16_round(y[2 * t]) + __uint_as_float(pw[t] << 16);
              y[2 * t + 1] = bf16_round(y[2 * t + 1]) + __uint_as_float(pw[t] & 0xFFFF0000u);
            }
          }
          uint4 w;
          w.x = pack_bf16x2(y[0], y[1]);
          w.y = pack_bf16x2(y[2], y[3]);
          w.z = pack_bf16x2(y[4], y[5]);
          w.w = pack_bf16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(orow + c0 + c) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();   // neither CTA may exit while the peer can still multicast into it or arrive on its barriers
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 256 * NQ);
}

template <int D, int POLY>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  attn_fwd_body<D, POLY, false, 2>(tmQ, tmK, tmV, p);
}

template <int D, int POLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAttnThreads, 1)
attn_fwd_mc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  attn_fwd_body<D, POLY, true, 2>(tmQ, tmK, tmV, p);
}

constexpr int kAttnShortThreads = 192;   // 4 softmax warps + MMA warp + TMA warp; two CTAs per SM

template <int D, int POLY>
__global__ void __launch_bounds__(kAttnShortThreads, 2)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  attn_fwd_body<D, POLY, false, 1>(tmQ, tmK, tmV, p);
}

int make_qkv_map(CUtensorMap* m, const fwb_tensor4_t* t, int B, int H, int L, int D, int box_rows = 128) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)t->sh * 2, (uint64_t)t->sl * 2, (uint64_t)t->sb * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_bf16(m, t->ptr, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Combine the S key-split partials of every tail tile and finish the row exactly as the unsplit epilogue would:
// bf16 out, or (split-KV mode) the fp32 subset-normalised result + lse.  One thread per 8 output elements.
__global__ void attn_tail_merge_kernel(const AttnParams p, int tail) {
  const int pieces = p.d_real / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)tail * (2 * BQ) * pieces) return;
  const int pc = (int)(gid % pieces);
  const int r = (int)((gid / pieces) % (2 * BQ));
  const int t = (int)(gid / ((long long)pieces * 2 * BQ));
  const int tile = p.n_full + t;
  const int qblock = tile % p.nq, head = (tile / p.nq) % p.H, batch = tile / (p.nq * p.H);
  const int row = qblock * 2 * BQ + r;
  if (row >= p.Lq) return;
  const float* lse = p.ws_lse + ((long long)t * p.S) * (2 * BQ) + r;
  const float* po = p.ws_out + (((long long)t * p.S) * (2 * BQ) + r) * p.d_real + pc * 8;
  float m = -INFINITY;
  for (int s = 0; s < p.S; ++s) m = fmaxf(m, lse[s * 2 * BQ]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wsum = 0.f;
  for (int s = 0; s < p.S; ++s) {
    const float w = fast_exp2(lse[s * 2 * BQ] - m);
    wsum += w;
    const float4 a = *reinterpret_cast<const float4*>(po + (long long)s * 2 * BQ * p.d_real);
    const float4 c = *reinterpret_cast<const float4*>(po + (long long)s * 2 * BQ * p.d_real + 4);
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * c.x; acc[5] += w * c.y; acc[6] += w * c.z; acc[7] += w * c.w;
  }
  const float inv = 1.f / wsum;
  if (p.part_out) {
    float* dst = p.part_out + (((long long)batch * p.Lq + row) * p.H + head) * p.d_real + pc * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
    if (pc == 0) p.part_lse[((long long)batch * p.H + head) * p.Lq + row] = m + log2f(wsum);
  } else {
    uint4 o;
    o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
    o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
    o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(p.out + (long long)batch * p.o_sb + (long long)row * p.o_sl + (long long)head * p.o_sh + pc * 8) = o;
  }
}

int g_attn_short_max_keys = 2048;   // default policy: <= this many keys -> the one-tile, two-CTAs-per-SM configuration of the aliased kernel
                                     // (fwb_attn_set_short_kv_max; 0 disables)
int g_attn_multicast = 1;   // aliased kernel, head_dim 128, >= 2048 keys: CTA pairs sharing K/V tiles by TMA multicast (fwb_attn_set_multicast);
                            // bit-identical, -1.0 % step time in the in-step A/B (profiles/r02_attention.md §7)

template <int D, int POLY>
int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B, int H,
                cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  if (p.short_kv) {
    // one Q tile per CTA, half the K/V stages: Q + (K + V) x stages  (97 KB at head_dim 128, 81 KB at 64 -> two CTAs per SM)
    constexpr int kShortSmem = Cfg::kTileBytes + 2 * (Cfg::kStages / 2) * Cfg::kTileBytes + 1024;
    static AttrOnce once_short;
    if (once_short.need(current_device()))
      FWB_CUDA(cudaFuncSetAttribute(attn_fwd_short_kernel<D, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kShortSmem));
    const long long tiles = (long long)p.nq * H * B;
    FWB_CHECK(tiles < (1ll << 31), "attn: grid too large");
    attn_fwd_short_kernel<D, POLY><<<(unsigned)tiles, kAttnShortThreads, kShortSmem, stream>>>(tq, tk, tv, p);
    FWB_CUDA(cudaGetLastError());
    return FWB_OK;
  }
  static AttrOnce once;
  if (once.need(current_device()))
    FWB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const long long n_tiles = (long long)p.nq * H * B;
  const long long tail = n_tiles - p.n_full;
  const long long grid = p.n_full + (p.S > 1 ? tail * p.S : tail);
  FWB_CHECK(grid < (1ll << 31), "attn: grid too large");
  bool mc = false;
  if constexpr (D == 128) {
    // multicast pairs need: both CTAs of a pair on the same (batch, head) -> an even number of query blocks; no key-split tail
    // ... and it only pays on long key sequences at the full head_dim (measured: -9 % at 512 keys — the pair runs in lock-step —, -1 % at
    // head_dim 96, +1 % isolated / see profiles/r02_attention.md for the in-step A/B at head_dim 128)
    mc = g_attn_multicast && p.S == 1 && (p.nq % 2) == 0 && (grid % 2) == 0 && p.d_real == 128 && p.Lk >= 2048;
    if (mc) {
      static AttrOnce once_mc;
      if (once_mc.need(current_device()))
        FWB_CUDA(cudaFuncSetAttribute(attn_fwd_mc_kernel<D, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      attn_fwd_mc_kernel<D, POLY><<<(unsigned)grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
    }
  }
  if (!mc) attn_fwd_kernel<D, POLY><<<(unsigned)grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  FWB_CUDA(cudaGetLastError());
  if (p.S > 1) {
    const long long total = tail * (2 * BQ) * (p.d_real / 8);
    attn_tail_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, (int)tail);
    FWB_CUDA(cudaGetLastError());
  }
  return FWB_OK;
}


// =====================================================================================================================
// attn2: decoupled pipeline.  The v1 kernel above keeps P aliased on S, so QK_i(j+1) can only be issued after PV_i(j): the
// chain softmax -> P hand-off -> PV + QK -> commit -> wake-up is serial per Q tile (measured with tools/attn_trace.py: the
// softmax warps wait ~40 % of the time and the tensor pipe idles ~45 %).  Here P_i has its own TMEM columns, the softmax
// releases S_i as soon as it sits in registers, and each Q tile has its own MMA-issuing warp:
//   warps 0-3 / 4-7   softmax of Q tile 0 / 1          warp 8 / 9   MMA issuer of Q tile 0 / 1          warp 10   TMA
//   MMA warp i:  for j: [S_i free] QK_i(j) -> s_full ;  [P_i(j-1) written] PV_i(j-1) -> p_free
//   softmax i :  for j: [s_full] S -> registers -> s_free ; max / exp2 ; [p_free(j-1)] P -> TMEM -> p_full
// so QK_i(j+1) runs on the tensor pipe while softmax_i(j) is still computing, and the softmax warps never wait in steady
// state.  TMEM (512 columns) has to hold O_0 O_1 S_0 P_0 S_1 P_1 = 2 * (D + 1.5 * BKV): head_dim 64 -> BKV 128 (512
// columns), head_dim 128 -> BKV 64 (448 columns).  K/V stages are released by both MMA warps (mbarrier count 2).
// =====================================================================================================================
constexpr int kAttn2Threads = 352;
__device__ int g_attn2_pingpong = 1;   // exp2 phases of the two Q tiles alternate through named barriers (fwb_attn_set_mufu_pingpong(2, .))

template <int D, int BK>
struct Attn2Cfg {
  static constexpr int kBoxes = D / 64;
  static constexpr int kQBoxBytes = BQ * 128;              // one 64-column box of a Q tile
  static constexpr int kQTileBytes = kBoxes * kQBoxBytes;
  static constexpr int kKVBoxBytes = BK * 128;             // one 64-column box of a K or V tile
  static constexpr int kKVTileBytes = kBoxes * kKVBoxBytes;
  static constexpr int kStages = 4;
  static constexpr int kSmemBytes = 2 * kQTileBytes + 2 * kStages * kKVTileBytes + 1024;
  static constexpr uint32_t kColO = 0;                     // O_i at i * D
  static constexpr uint32_t kColS = 2 * D;                 // S_i at 2D + i * 1.5 BK, P_i right behind S_i
  static constexpr uint32_t kTileCols = BK + BK / 2;
  static_assert(2 * D + 2 * kTileCols <= 512, "TMEM overflow");
};

template <int D, int BK, int POLY>
__global__ void __launch_bounds__(kAttn2Threads, 1)
attn2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = Attn2Cfg<D, BK>;
  constexpr int ST = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                     // [2][kQTileBytes]
  uint8_t* sK = smem + 2 * Cfg::kQTileBytes;              // [ST][kKVTileBytes]
  uint8_t* sV = sK + ST * Cfg::kKVTileBytes;              // [ST][kKVTileBytes]
  __shared__ uint64_t q_full[2], k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];
  __shared__ uint64_t s_full[2], s_free[2], p_full[2], p_free[2], o_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float pp_scratch[1 + 256];   // [0] = 0.0f (read behind the ping-pong barrier), [1 + tid] = sink for the block sums

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const uint32_t pp_addr = smem_u32(pp_scratch);
  if (threadIdx.x == 0) pp_scratch[0] = 0.f;
  auto decode = [&](int& tile, int& split, int& nsplit) {
    tile = blockIdx.x; split = 0; nsplit = 1;
    if (tile >= p.n_full) {
      const int r = tile - p.n_full;
      tile = p.n_full + r / p.S;
      split = r % p.S;
      nsplit = p.S;
    }
  };
  int qblock, head, batch, kv0, n_kv, j_ragged;
  {
    int tile, split, nsplit;
    decode(tile, split, nsplit);
    qblock = tile % p.nq; head = (tile / p.nq) % p.H; batch = tile / (p.nq * p.H);
    const int n_kv_all = (p.Lk + BK - 1) / BK;
    kv0 = (int)((long long)split * n_kv_all / nsplit);
    n_kv = (int)((long long)(split + 1) * n_kv_all / nsplit) - kv0;
    j_ragged = n_kv_all - 1 - kv0;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);   // one arrive per softmax warp
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_full[i], 1);
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);  // one commit per MMA warp
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 8) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 10) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      for (int i = 0; i < 2; ++i) {
        mbar_arrive_expect_tx(&q_full[i], Cfg::kQTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_4d(sQ + i * Cfg::kQTileBytes + b * Cfg::kQBoxBytes, &tmQ, &q_full[i], b * 64, head, (qblock * 2 + i) * BQ,
                      batch);
      }
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t s = j % ST, ph = (j / ST) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], Cfg::kKVTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_4d(sK + s * Cfg::kKVTileBytes + b * Cfg::kKVBoxBytes, &tmK, &k_full[s], b * 64, head, (kv0 + j) * BK, batch);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], Cfg::kKVTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_4d(sV + s * Cfg::kKVTileBytes + b * Cfg::kKVBoxBytes, &tmV, &v_full[s], b * 64, head, (kv0 + j) * BK, batch);
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------ MMA issuer of Q tile i ----------------------------
    if (elect_one()) {
      const int i = warp - 8;
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BK, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(BQ, p.pv_n, 0, 1);  // B = V is MN-major (d contiguous)
      const int ksteps_qk = (p.d_real + 15) / 16;
      const uint32_t qa = smem_u32(sQ) + i * Cfg::kQTileBytes, k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t s_tmem = tmem_base + Cfg::kColS + i * Cfg::kTileCols;
      const uint32_t p_tmem = s_tmem + BK;
      const uint32_t o_tmem = tmem_base + Cfg::kColO + i * D;

      mbar_wait(&q_full[i], 0);
      for (int j = 0; j <= n_kv; ++j) {
        if (j < n_kv) {
          // S_i(j) = Q_i K_j^T  — needs the softmax to have pulled S_i(j-1) into registers
          const uint32_t ks = j % ST;
          if (j > 0) mbar_wait(&s_free[i], (j - 1) & 1);
          mbar_wait(&k_full[ks], (j / ST) & 1);
          tc_fence_after();
          const uint32_t ka = k_addr + ks * Cfg::kKVTileBytes;
          for (int kk = 0; kk < ksteps_qk; ++kk) {
            umma_ss(s_tmem, make_smem_desc(qa + (kk >> 2) * Cfg::kQBoxBytes + (kk & 3) * 32, 0, 1024, SWZ_128B),
                    make_smem_desc(ka + (kk >> 2) * Cfg::kKVBoxBytes + (kk & 3) * 32, 0, 1024, SWZ_128B), idesc_qk, kk > 0);
          }
          tc_commit(&s_full[i]);
          tc_commit(&k_empty[ks]);
          TRACE1(8 + 0, j, i * 3 + 0);
        }
        if (j > 0) {
          // O_i += P_i(j-1) V_(j-1)
          const int jj = j - 1;
          const uint32_t vs = jj % ST;
          mbar_wait(&p_full[i], jj & 1);
          mbar_wait(&v_full[vs], (jj / ST) & 1);
          tc_fence_after();
          const uint32_t va = v_addr + vs * Cfg::kKVTileBytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // 16 kv rows per K-step = 2048 B; LBO = stride between the 64-column d boxes, SBO = 8 kv rows
            umma_ts(o_tmem, p_tmem + kk * 8, make_smem_desc(va + kk * 2048, Cfg::kKVBoxBytes, 1024, SWZ_128B), idesc_pv,
                    jj > 0 || kk > 0);
          }
          tc_commit(&p_free[i]);
          tc_commit(&v_empty[vs]);
          if (jj == n_kv - 1) tc_commit(&o_full[i]);
          TRACE1(8 + 0, jj, i * 3 + 1);
        }
      }
    }
  } else {
    // ------------------------------------ softmax + epilogue ------------------------------
    const int i = warp >> 2;
    const uint32_t quad = warp & 3;
    const uint32_t lane_off = (quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + Cfg::kColS + i * Cfg::kTileCols + lane_off;
    const uint32_t p_tmem = s_tmem + BK;
    const uint32_t o_tmem = tmem_base + Cfg::kColO + i * D + lane_off;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY;
    float l_sum = 0.f;
    // MUFU ping-pong.  Warp w (Q tile 0) and warp w+4 (Q tile 1) sit on the same SM sub-partition and share its MUFU
    // (4 lanes/clk: 8 clk per warp-wide ex2).  Left alone they fall into lock-step: the exp2 phases collide at half rate and
    // the MUFU idles while both do the row max and the TMEM traffic (tools/attn_trace.py).  Two named barriers per warp pair
    // make the exp2 phases alternate, so one tile's max / TMEM phase hides under the other's exponentials.
    const bool pingpong = g_attn2_pingpong != 0;
    const uint32_t bar_mine = 1 + quad + 4 * i, bar_other = 1 + quad + 4 * (1 - i);
    if (pingpong && i == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");   // tile 0 goes first

    for (int j = 0; j < n_kv; ++j) {
      TRACE(warp, j, 0);
      mbar_wait(&s_full[i], j & 1);
      TRACE(warp, j, 1);
      tc_fence_after();
      uint32_t v[BK];
#pragma unroll
      for (int c0 = 0; c0 < BK; c0 += 32) tmem_ld32(s_tmem + c0, *reinterpret_cast<uint32_t(*)[32]>(&v[c0]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[i]);   // S_i is in registers: the MMA warp may overwrite it with S_i(j+1)
      TRACE(warp, j, 2);

      if (j == j_ragged) {
        const int valid = p.Lk - (kv0 + j) * BK;
        if (valid < BK) {
#pragma unroll
          for (int c = 0; c < BK; ++c)
            if (c >= valid) v[c] = 0xFF800000u;  // -inf
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < BK; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(v[c]));
        mx1 = fmaxf(mx1, __uint_as_float(v[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[c + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(v[c + 3]));
      }
      const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
      const bool need = m_new > m_used + 8.0f;
      bool p_free_seen = false;
      if (__any_sync(0xffffffffu, need)) {
        const float m_next = fmaxf(m_used, m_new);
        const float alpha = fast_exp2(m_used - m_next);
        l_sum *= alpha;
        m_used = m_next;
        if (j > 0) {
          mbar_wait(&p_free[i], (j - 1) & 1);   // PV_i(j-1) has completed: O_i is quiescent
          p_free_seen = true;
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(o_tmem + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st16(o_tmem + c0, o);
          }
        }
      }
      // m_used is threaded through the barrier asm so that the exponentials (register-only math, which the compiler may
      // otherwise move across an asm volatile) stay behind it; same for the accumulators and the arrive below
      if (pingpong) {
        // ptxas moves register-only math (the exponentials) freely across a BAR; a volatile shared-memory load behind the
        // barrier that feeds m_used (+0) and a volatile store of the sums in front of the arrive pin the exp2 phase
        float z;
        asm volatile("bar.sync %1, 64;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(z) : "r"(bar_mine), "r"(pp_addr) : "memory");
        m_used += z;
      }
      TRACE(warp, j, 3);
      uint32_t pk[BK / 2];
      const float blk_sum = softmax_exp<BK, POLY>(v, sl2, m_used, pk);
      if (pingpong)
        asm volatile("st.volatile.shared.f32 [%0], %1;\n\tbar.arrive %2, 64;" ::"r"(pp_addr + 4 + 4 * threadIdx.x), "f"(blk_sum),
                     "r"(bar_other)
                     : "memory");
      l_sum += blk_sum;
      TRACE(warp, j, 4);
      if (j > 0 && !p_free_seen) {
        mbar_wait(&p_free[i], (j - 1) & 1);     // PV_i(j-1) has finished reading P_i
        tc_fence_after();
      }
#pragma unroll
      for (int c0 = 0; c0 < BK / 2; c0 += 32) tmem_st32(p_tmem + c0, *reinterpret_cast<uint32_t(*)[32]>(&pk[c0]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[i]);
      TRACE(warp, j, 5);
    }

    if (pingpong && i == 0) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");   // absorb tile 1's last arrive

    // epilogue: O / l  -> bf16 -> global (or fp32 partial / workspace, as in v1)
    mbar_wait(&o_full[i], 0);
    tc_fence_after();
    const int row = (qblock * 2 + i) * BQ + quad * 32 + lane;
    const float inv_l = 1.f / l_sum;
    __nv_bfloat16* orow = p.out + (long long)batch * p.o_sb + (long long)row * p.o_sl + (long long)head * p.o_sh;
    const int row_in_tile = i * BQ + quad * 32 + lane;
    float* ws_row = nullptr;
    int tile, split, nsplit;
    decode(tile, split, nsplit);
    if (nsplit > 1) {
      const long long slot = (long long)(tile - p.n_full) * p.S + split;
      ws_row = p.ws_out + (slot * (2 * BQ) + row_in_tile) * p.d_real;
      p.ws_lse[slot * (2 * BQ) + row_in_tile] = m_used + log2f(l_sum);
    } else if (p.part_lse && row < p.Lq) {
      p.part_lse[((long long)batch * p.H + head) * p.Lq + row] = m_used + log2f(l_sum);
    }
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c0, o);
      tmem_ld_wait();
      float* frow = ws_row ? ws_row : (p.part_out ? p.part_out + (((long long)batch * p.Lq + row) * p.H + head) * p.d_real : nullptr);
      if (frow) {
        if ((ws_row || row < p.Lq) && c0 < p.d_real) {
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(frow + c0 + c) =
                make_float4(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l, __uint_as_float(o[c + 2]) * inv_l,
                            __uint_as_float(o[c + 3]) * inv_l);
        }
      } else if (row < p.Lq && c0 < p.d_real) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          float y[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) y[t] = __uint_as_float(o[c + t]) * inv_l;
          if (p.accumulate) {
            const uint4 prev = *reinterpret_cast<const uint4*>(orow + c0 + c);
            const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              y[2 * t] = bf16_round(y[2 * t]) + __uint_as_float(pw[t] << 16);
              y[2 * t + 1] = bf16_round(y[2 * t + 1]) + __uint_as_float(pw[t] & 0xFFFF0000u);
            }
          }
          uint4 w;
          w.x = pack_bf16x2(y[0], y[1]);
          w.y = pack_bf16x2(y[2], y[3]);
          w.z = pack_bf16x2(y[4], y[5]);
          w.w = pack_bf16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(orow + c0 + c) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

template <int D, int BK, int POLY>
int launch_attn2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B, int H,
                 cudaStream_t stream) {
  using Cfg = Attn2Cfg<D, BK>;
  static AttrOnce once;
  if (once.need(current_device()))
    FWB_CUDA(cudaFuncSetAttribute(attn2_kernel<D, BK, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const long long n_tiles = (long long)p.nq * H * B;
  const long long tail = n_tiles - p.n_full;
  const long long grid = p.n_full + (p.S > 1 ? tail * p.S : tail);
  FWB_CHECK(grid < (1ll << 31), "attn: grid too large");
  attn2_kernel<D, BK, POLY><<<(unsigned)grid, kAttn2Threads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  FWB_CUDA(cudaGetLastError());
  if (p.S > 1) {
    const long long total = tail * (2 * BQ) * (p.d_real / 8);
    attn_tail_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, (int)tail);
    FWB_CUDA(cudaGetLastError());
  }
  return FWB_OK;
}

// POLY (pairs out of every 8 that take 2^x from the FMA-pipe polynomial) is a template parameter; built: 0 (MUFU only, the A/B
// reference) and 2 (the measured optimum, default).  3/8 and 4/8 were measured slower (profiles/r02_attention.md) and are not built.
template <int D>
int launch_attn_poly(int poly, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                     int H, cudaStream_t stream) {
  if (poly == 2) return launch_attn<D, 2>(tq, tk, tv, p, B, H, stream);
  return launch_attn<D, 0>(tq, tk, tv, p, B, H, stream);
}
template <int D, int BK>
int launch_attn2_poly(int poly, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                      int H, cudaStream_t stream) {
  if (poly == 2) return launch_attn2<D, BK, 2>(tq, tk, tv, p, B, H, stream);
  return launch_attn2<D, BK, 0>(tq, tk, tv, p, B, H, stream);
}

// out[b,l,h,:] = sum_s w_s part[s][b,l,h,:] / sum_s w_s,  w_s = 2^(lse[s][b,h,l] - max_s lse)   (fp32 in, bf16 out)
__global__ void attn_merge_kernel(const float* __restrict__ part, const float* __restrict__ lse, __nv_bfloat16* __restrict__ out,
                                  long long o_sb, long long o_sl, long long o_sh, int S, int B, int H, int L, int D) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 output elements
  const int pieces = D / 8;
  const long long total = (long long)B * L * H * pieces;
  if (gid >= total) return;
  const int pc = (int)(gid % pieces);
  const int h = (int)((gid / pieces) % H);
  const long long l = (gid / ((long long)pieces * H)) % L;
  const int b = (int)(gid / ((long long)pieces * H * L));
  const long long lse_idx = ((long long)b * H + h) * L + l, lse_stride = (long long)B * H * L;
  const long long part_idx = ((((long long)b * L + l) * H + h) * D) + pc * 8, part_stride = (long long)B * L * H * D;
  float m = -INFINITY;
  for (int s = 0; s < S; ++s) m = fmaxf(m, lse[lse_idx + s * lse_stride]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wsum = 0.f;
  for (int s = 0; s < S; ++s) {
    const float w = fast_exp2(lse[lse_idx + s * lse_stride] - m);
    wsum += w;
    const float4 a = *reinterpret_cast<const float4*>(part + part_idx + s * part_stride);
    const float4 c = *reinterpret_cast<const float4*>(part + part_idx + s * part_stride + 4);
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * c.x; acc[5] += w * c.y; acc[6] += w * c.z; acc[7] += w * c.w;
  }
  const float inv = 1.f / wsum;
  uint4 o;
  o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
  o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
  o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
  o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
  *reinterpret_cast<uint4*>(out + b * o_sb + l * o_sl + h * o_sh + pc * 8) = o;
}

// Process-wide tuning state behind the named setters of include/fwb200.h (A/B measurements, tests).  One process drives one GPU
// (DESIGN.md), so this is not per-stream state.
int g_attn_variant = 0;     // 0: per head_dim default (64 -> decoupled attn2, 96 / 128 -> v1: measured), 1: v1, 2: decoupled attn2
int g_attn_tail_split = 1;  // key-split tail of the tile schedule on / off
int g_attn_poly = -1;       // -1: per head_dim default, else 0 / 2 = pairs out of 8 on the exp2 polynomial
int g_attn_pv96 = 1;        // head_dim 96: PV MMAs with N = 96 instead of the zero-padded 128 (fwb_attn_set_pv_n96): bit-identical,
                            // +2.4 % on the adapter shapes (profiles/r02_attention.md)

// exp2 polynomial share per head_dim (pairs of 8); set from the A/B measurements in profiles/r02_attention.md
inline int default_poly(int D) { (void)D; return 2; }   // 2 of 8 pairs: +1 % (head_dim 128), +13 % (96), +10 % (64), r02_attn_sweep.log

// Tile schedule planner (pure host arithmetic).  n_tiles tiles on W SMs (one CTA per SM): the first floor(n/W)*W tiles run
// unsplit; each of the `tail` remaining tiles may be split S ways along the keys -> tail * S short CTAs running in
// ceil(tail * S / W) rounds of 1/S of a full tile each.  Chooses the S in 2..16 (at least 512 keys per split, slots must fit the
// workspace) with the smallest modelled time and uses it only for a predicted gain of at least 7 %.
void attn_plan(long long n_tiles, int Lk, int D, size_t ws_bytes, int W, int* n_full, int* S_out) {
  *n_full = (int)n_tiles;
  *S_out = 1;
  if (W <= 0 || n_tiles % W == 0) return;
  const long long full = (n_tiles / W) * W, tail = n_tiles - full;
  const int max_S = Lk / 512;
  const double plain = (double)(full / W + 1);
  const long long max_slots = (long long)(ws_bytes / ((size_t)2 * BQ * (D + 1) * sizeof(float)));
  double best = plain;
  int best_S = 1;
  for (int S = 2; S <= 16 && S <= max_S && tail * S <= max_slots; ++S) {
    const double cost = (double)(full / W) + (double)((tail * S + W - 1) / W) / S + 0.04;   // + short-CTA prologue / merge
    if (cost < best - 1e-9) {
      best = cost;
      best_S = S;
    }
  }
  if (best_S >= 2 && best <= 0.93 * plain) {
    *n_full = (int)full;
    *S_out = best_S;
  }
}

}  // namespace

extern "C" int fwb_attn_plan(int B, int H, int Lq, int Lk, int D, size_t workspace_bytes, int n_sms, int* n_unsplit_tiles,
                             int* key_splits) {
  FWB_CHECK(B > 0 && H > 0 && Lq > 0 && Lk > 0 && n_unsplit_tiles && key_splits, "attn_plan: bad arguments");
  attn_plan((long long)((Lq + 2 * BQ - 1) / (2 * BQ)) * H * B, Lk, D, workspace_bytes, n_sms, n_unsplit_tiles, key_splits);
  return FWB_OK;
}

#ifdef FWB_ATTN_TRACE
extern "C" int fwb_attn_trace_read(long long* host_out, int cta) {
  if (cta >= 0) {
    FWB_CUDA(cudaMemcpyToSymbol(g_trace_cta, &cta, sizeof(int)));
    return FWB_OK;
  }
  FWB_CUDA(cudaDeviceSynchronize());
  FWB_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 9 * 64 * 8));
  return FWB_OK;
}
#endif

extern "C" int fwb_attn_set_variant(int variant) {
  FWB_CHECK(variant >= 0 && variant <= 2, "attn_set_variant: 0 (default per head_dim), 1 (aliased S/P) or 2 (decoupled S/P)");
  g_attn_variant = variant;
  return FWB_OK;
}

extern "C" int fwb_attn_set_tail_split(int enabled) {
  g_attn_tail_split = enabled ? 1 : 0;
  return FWB_OK;
}

extern "C" int fwb_attn_set_exp2_poly(int pairs_of_8) {
  FWB_CHECK(pairs_of_8 == -1 || pairs_of_8 == 0 || pairs_of_8 == 2, "attn_set_exp2_poly: -1 (default), 0 or 2 pairs out of every 8");
  g_attn_poly = pairs_of_8;
  return FWB_OK;
}

extern "C" int fwb_attn_set_short_kv_max(int max_keys) {
  FWB_CHECK(max_keys >= 0, "attn_set_short_kv_max: >= 0");
  g_attn_short_max_keys = max_keys;
  return FWB_OK;
}

extern "C" int fwb_attn_set_multicast(int enabled) {
  g_attn_multicast = enabled ? 1 : 0;
  return FWB_OK;
}

extern "C" int fwb_attn_set_pv_n96(int enabled) {
  g_attn_pv96 = enabled ? 1 : 0;
  return FWB_OK;
}

extern "C" int fwb_attn_set_mufu_pingpong(int kernel, int enabled) {
  FWB_CHECK(kernel == 1 || kernel == 2, "attn_set_mufu_pingpong: kernel 1 (v1) or 2 (decoupled)");
  const int on = enabled ? 1 : 0;
  if (kernel == 1) {
    FWB_CUDA(cudaMemcpyToSymbol(g_attn1_pingpong, &on, sizeof(int)));
  } else {
    FWB_CUDA(cudaMemcpyToSymbol(g_attn2_pingpong, &on, sizeof(int)));
  }
  return FWB_OK;
}

static int attn_impl(const fwb_tensor4_t* q, const fwb_tensor4_t* k, const fwb_tensor4_t* v, const fwb_tensor4_t* out, int B, int H,
                     int Lq, int Lk, int D, float scale, int accumulate, float* part_out, float* part_lse, void* ws, size_t ws_bytes,
                     cudaStream_t stream);

extern "C" size_t fwb_attn_workspace_bytes(void) {
  // up to four rounds of split CTAs: 4 x (#SMs) slots of 256 rows x (128 + 1) floats  (78 MB on a 148-SM part)
  return (size_t)4 * (num_sms() > 0 ? num_sms() : 148) * 2 * BQ * (128 + 1) * sizeof(float);
}

extern "C" int fwb_attn_fwd(const fwb_tensor4_t* q, const fwb_tensor4_t* k, const fwb_tensor4_t* v,
                            const fwb_tensor4_t* out, int B, int H, int Lq, int Lk, int D, float scale,
                            int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  FWB_CHECK(out && out->ptr, "attn: null output");
  return attn_impl(q, k, v, out, B, H, Lq, Lk, D, scale, accumulate, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int fwb_attn_fwd_partial(const fwb_tensor4_t* q, const fwb_tensor4_t* k, const fwb_tensor4_t* v, float* part_out,
                                    float* part_lse, int B, int H, int Lq, int Lk, int D, float scale, void* workspace,
                                    size_t workspace_bytes, cudaStream_t stream) {
  FWB_CHECK(part_out && part_lse, "attn_partial: null output");
  FWB_CHECK((reinterpret_cast<uintptr_t>(part_out) & 15) == 0, "attn_partial: part_out must be 16-byte aligned");
  fwb_tensor4_t dummy = *q;   // only validated, never written in partial mode
  return attn_impl(q, k, v, &dummy, B, H, Lq, Lk, D, scale, 0, part_out, part_lse, workspace, workspace_bytes, stream);
}

extern "C" int fwb_attn_merge(const float* part, const float* lse, const fwb_tensor4_t* out, int S, int B, int H, int L, int D,
                              cudaStream_t stream) {
  FWB_CHECK(part && lse && out && out->ptr, "attn_merge: null pointer");
  FWB_CHECK(S >= 1 && B > 0 && H > 0 && L > 0 && D % 8 == 0, "attn_merge: bad shape");
  FWB_CHECK(out->sb % 8 == 0 && out->sl % 8 == 0 && out->sh % 8 == 0 && (reinterpret_cast<uintptr_t>(out->ptr) & 15) == 0,
            "attn_merge: output strides must be multiples of 8 elements and the pointer 16-byte aligned");
  const long long total = (long long)B * L * H * (D / 8);
  const long long blocks = (total + 255) / 256;
  attn_merge_kernel<<<(unsigned)blocks, 256, 0, stream>>>(part, lse, reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(out->ptr)),
                                                          out->sb, out->sl, out->sh, S, B, H, L, D);
  FWB_CUDA(cudaGetLastError());
  return FWB_OK;
}

static int attn_impl(const fwb_tensor4_t* q, const fwb_tensor4_t* k, const fwb_tensor4_t* v, const fwb_tensor4_t* out, int B, int H,
                     int Lq, int Lk, int D, float scale, int accumulate, float* part_out, float* part_lse, void* ws, size_t ws_bytes,
                     cudaStream_t stream) {
  FWB_CHECK(q && k && v && out && q->ptr && k->ptr && v->ptr && out->ptr, "attn: null pointer");
  FWB_CHECK(D == 64 || D == 96 || D == 128, "attn: head_dim %d unsupported (64, 96, 128)", D);
  FWB_CHECK(B > 0 && H > 0 && Lq > 0 && Lk > 0, "attn: empty problem B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  FWB_CHECK(H <= 65535 && B <= 65535, "attn: H and B must be <= 65535");
  const fwb_tensor4_t* ts[4] = {q, k, v, out};
  for (int i = 0; i < 4; ++i) {
    FWB_CHECK(ts[i]->sb % 8 == 0 && ts[i]->sl % 8 == 0 && ts[i]->sh % 8 == 0, "attn: strides must be multiples of 8 elements");
    FWB_CHECK((reinterpret_cast<uintptr_t>(ts[i]->ptr) & 15) == 0, "attn: pointers must be 16-byte aligned");
  }
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_map(&tq, q, B, H, Lq, D))) return rc;
  // default policy: few keys -> the short configuration of the aliased kernel (any head_dim); else decoupled at 64, aliased at 96 / 128
  const bool short_kv = g_attn_variant == 0 && g_attn_short_max_keys > 0 && Lk <= g_attn_short_max_keys;
  const int variant = short_kv ? 1 : (g_attn_variant ? g_attn_variant : (D == 64 ? 2 : 1));
  const int bk = (variant == 2) ? (D == 64 ? 128 : 64) : BKV;     // keys per KV tile of the kernel that will run
  if ((rc = make_qkv_map(&tk, k, B, H, Lk, D, bk))) return rc;
  if ((rc = make_qkv_map(&tv, v, B, H, Lk, D, bk))) return rc;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(out->ptr));
  p.o_sb = out->sb; p.o_sl = out->sl; p.o_sh = out->sh;
  p.Lq = Lq; p.Lk = Lk; p.d_real = D;
  p.pv_n = (D == 96) ? (g_attn_pv96 ? 96 : 128) : D;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.accumulate = accumulate;
  p.part_out = part_out;
  p.part_lse = part_lse;
  p.H = H;
  // ---- tile schedule: split the last, partly filled wave of tiles along the keys (needs a workspace; never in accumulate mode,
  // whose read-modify-write epilogue belongs to the unsplit CTA) ----
  p.short_kv = short_kv ? 1 : 0;
  p.nq = short_kv ? (Lq + BQ - 1) / BQ : (Lq + 2 * BQ - 1) / (2 * BQ);
  const long long n_tiles = (long long)p.nq * H * B;
  FWB_CHECK(n_tiles < (1ll << 30), "attn: too many tiles");
  p.n_full = (int)n_tiles;
  p.S = 1;
  p.ws_out = nullptr;
  p.ws_lse = nullptr;
  if (!short_kv && ws && !accumulate && g_attn_tail_split && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
    int n_full = 0, S = 1;
    attn_plan(n_tiles, Lk, D, ws_bytes, num_sms(), &n_full, &S);
    if (S >= 2) {
      p.n_full = n_full;
      p.S = S;
      p.ws_out = reinterpret_cast<float*>(ws);
      p.ws_lse = p.ws_out + (size_t)(n_tiles - n_full) * S * 2 * BQ * D;
    }
  }
  const int poly = g_attn_poly >= 0 ? g_attn_poly : default_poly(D);
  if (variant == 2) {
    if (D == 64) return launch_attn2_poly<64, 128>(poly, tq, tk, tv, p, B, H, stream);
    return launch_attn2_poly<128, 64>(poly, tq, tk, tv, p, B, H, stream);
  }
  if (D == 64) return launch_attn_poly<64>(poly, tq, tk, tv, p, B, H, stream);
  return launch_attn_poly<128>(poly, tq, tk, tv, p, B, H, stream);
}
