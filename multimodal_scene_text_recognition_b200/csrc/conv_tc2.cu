// bf16 implicit-GEMM convolution, CTA-pair variant: two SMs of one TPC cooperate on a 256 x BLOCK_N
// output tile with tcgen05.mma.cta_group::2 (UMMA M = 256).
//
// Why: the single-CTA kernel (conv_tc.cu) moves 48 KB of operands through shared memory per 128x256x64
// k-block twice (TMA write + tensor-core read = 192 B/clk/SM against ~128 B/clk of shared-memory
// bandwidth); ncu shows the tensor pipe active only ~80 % of the time on the 512->512 layers.  In the
// pair, each CTA stages its own 128 activation rows (A, 16 KB) and only HALF of the weight tile
// (B, N/2 rows), and the hardware feeds both tensor cores from the two halves: per-SM shared-memory
// traffic drops by a third, L2->SM traffic likewise, and the ring gets 6 stages instead of 4.
//
// Protocol (both CTAs run all three roles except MMA issue):
//   producer  waits its OWN empty barrier, issues its TMA loads with .cta_group::2 so that they
//             complete_tx on the LEADER's full barrier; the leader arms that barrier for both CTAs' bytes
//   MMA       leader only: waits full (own) and tempty (own, 8 arrivals = 4 epilogue warps x 2 CTAs),
//             issues tcgen05.mma.cta_group::2, commits with .multicast::cluster to BOTH CTAs'
//             empty / tfull barriers
//   epilogue  each CTA drains its own 128 TMEM lanes; the peer's warps arrive remotely on the leader's tempty
// Same arguments, epilogue and numerics as conv_gemm_bf16_kernel.
#include <stdlib.h>

#include "tc_common.cuh"

namespace mstr {

struct Tc2Params {
  int M, Cout;
  int num_m_pairs, num_n_tiles;
  int taps, S, cin_chunks;
  int cin2_chunks;          // extra k-blocks read from a second [M][C_in2] matrix (fused 1x1 branch), 0 = none
  int P, Q, sh, sw, ph, pw;
  int im2col;
  int relu;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out_bf16;
  float* out_f32;
};

constexpr int T2_THREADS = 320;      // producer, MMA issuer, 2 x 4 epilogue warps
constexpr int T2_BLOCK_K = 64;
constexpr int T2_A_BYTES = 128 * 128;   // 128 rows x 64 bf16

// Only BLOCK_N = 256, MT = 1 is instantiated.  The template also expresses 256 x 512 pair tiles (BLOCK_N = 512) and two
// 256-row accumulators per pair for C_out = 128 (BLOCK_N = 128, MT = 2); both were built, verified bit-identical and
// measured SLOWER on the bench step (conv stage 27.4 -> 30.3 ms; five C_out = 128 layers 4.47 -> 4.70 ms) and have been
// removed from the library - the measurements and what they say about the limiter are in DESIGN.md 4.
template <int BLOCK_N, int MT = 1, bool F32OUT = false>
struct Tc2Cfg {
  static constexpr int UMMA_N = BLOCK_N > 256 ? 256 : BLOCK_N;    // columns per tcgen05.mma
  static constexpr int NH = BLOCK_N / UMMA_N;                       // MMAs per k-step and accumulator
  static constexpr int NUM_ACC = (2 * MT * BLOCK_N <= 512) ? 2 : 1; // accumulator buffers in TMEM
  static constexpr int A_BYTES = MT * T2_A_BYTES;            // this CTA's 128 rows of each of the MT accumulators
  static constexpr int B_BYTES = (BLOCK_N / 2) * 128;       // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr bool BF16_ONLY = BLOCK_N > 256 || MT > 1;
  // output slab rings (tc_common.cuh), one per epilogue warp group: 24 KB each (3 bf16 slabs) for the common tile,
  // which has the room next to 5 operand stages; 32 KB each (2 f32 slabs) for the f32-output instantiation (the
  // BiLSTM input projections: K <= 512, 4 stages are plenty); 16 KB each (2 bf16 slabs) for the bf16-only variants
  static constexpr int EPI_GROUP_BYTES = BF16_ONLY ? EPI_SLAB_BYTES : (F32OUT ? 2 * EPI_SLAB_BYTES : 3 * EPI_SLAB_BYTES / 2);
  static constexpr int EPI_REGION = 2 * EPI_GROUP_BYTES;
  static constexpr int TAIL_BYTES = 4 * BLOCK_N * 4 + (2 * 8 + 4) * 8 + 16;
  static constexpr int RING_BUDGET = 227 * 1024 - 1024 - EPI_REGION - TAIL_BYTES;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = NUM_ACC * MT * BLOCK_N;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_REGION + TAIL_BYTES;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.cta_group::2"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c,
                                                    int w, int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.cta_group::2"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

template <int BLOCK_N, int MT = 1, bool F32OUT = false>
__global__ void __launch_bounds__(T2_THREADS, 1)
conv_gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmOb, const __grid_constant__ CUtensorMap tmOf,
                           const __grid_constant__ CUtensorMap tmA2, const Tc2Params p) {
  using Cfg = Tc2Cfg<BLOCK_N, MT, F32OUT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int PAIR_M = 256 * MT;         // output rows per pair and tile
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem base has the same offset in both CTAs, so this rounding is identical in both
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES;                            // output slab ring
  float* sScale = reinterpret_cast<float*>(sEpi + Cfg::EPI_REGION);            // [2][BLOCK_N]
  float* sShift = sScale + 2 * BLOCK_N;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sShift + 2 * BLOCK_N);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = p.num_m_pairs * p.num_n_tiles;
  const int num_k = p.taps * p.cin_chunks + p.cin2_chunks;
  const int first_tile = (int)cluster_id_x();
  const int tile_step = (int)num_clusters_x();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);    // leader's: one arrive.expect_tx (+ the bytes of both CTAs)
      mbar_init(&empty_bar[i], 1);   // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);   // one multicast commit
      mbar_init(&tempty_bar[i], 16); // leader's: 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                // barriers of BOTH CTAs initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        // this CTA stages rows [h * UMMA_N + rank * UMMA_N/2, + UMMA_N/2) of the weight tile for MMA h
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N + (int)rank * (Cfg::UMMA_N / 2);
        // accumulator a covers rows [tile_m0 + a * 256, + 256); this CTA stages its 128 rows of each
        const int m0 = (tile / p.num_n_tiles) * PAIR_M + (int)rank * 128;
        int cw[MT], chh[MT], cn[MT];
#pragma unroll
        for (int a = 0; a < MT; ++a) {
          cw[a] = chh[a] = cn[a] = 0;
          if (p.im2col) {
            const int ma = m0 + a * 256;
            const int q = ma % p.Q, pp = (ma / p.Q) % p.P;
            cn[a] = ma / (p.Q * p.P);
            cw[a] = q * p.sw - p.pw;
            chh[a] = pp * p.sh - p.ph;
          }
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
#pragma unroll
            for (int a = 0; a < MT; ++a) {
              uint8_t* dstA = sA + stage * Cfg::A_BYTES + a * T2_A_BYTES;
              if (p.im2col)
                tma2_load_im2col_4d(dstA, &tmA, bar, cc * T2_BLOCK_K, cw[a], chh[a], cn[a], (uint16_t)s, (uint16_t)r);
              else
                tma2_load_2d(dstA, &tmA, bar, cc * T2_BLOCK_K, m0 + a * 256);
            }
#pragma unroll
            for (int h = 0; h < Cfg::NH; ++h)
              tma2_load_2d(sB + stage * Cfg::B_BYTES + h * (Cfg::UMMA_N / 2) * 128, &tmB, bar,
                           (tap * p.cin_chunks + cc) * T2_BLOCK_K, n0 + h * Cfg::UMMA_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        // fused 1x1 branch: its input rows are the same output pixels, its weights follow in K
        for (int cc = 0; cc < p.cin2_chunks; ++cc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
#pragma unroll
          for (int a = 0; a < MT; ++a)
            tma2_load_2d(sA + stage * Cfg::A_BYTES + a * T2_A_BYTES, &tmA2, bar, cc * T2_BLOCK_K, m0 + a * 256);
#pragma unroll
          for (int h = 0; h < Cfg::NH; ++h)
            tma2_load_2d(sB + stage * Cfg::B_BYTES + h * (Cfg::UMMA_N / 2) * 128, &tmB, bar,
                         (p.taps * p.cin_chunks + cc) * T2_BLOCK_K, n0 + h * Cfg::UMMA_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, single thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, Cfg::UMMA_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * BLOCK_N);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < T2_BLOCK_K / 16; ++k)
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
              for (int h = 0; h < Cfg::NH; ++h)
                umma2_bf16(d_tmem + (uint32_t)(a * BLOCK_N + h * Cfg::UMMA_N),
                           make_smem_desc<128>(a_addr + a * T2_A_BYTES + k * 32),
                           make_smem_desc<128>(b_addr + h * (Cfg::UMMA_N / 2) * 128 + k * 32), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          umma2_commit_multicast(&empty_bar[stage]);
          if (kb == num_k - 1) umma2_commit_multicast(&tfull_bar[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == Cfg::NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs; warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    // two groups of four warps split the tile (MT = 2: accumulator g, otherwise column half g), each with its own
    // slab ring and named barriers - see conv_tc.cu
    const int eg = (warp - 2) >> 2;
    const int ep_tid = (threadIdx.x - 64) & 127;
    const int ep_all = threadIdx.x - 64;
    EpiStage est = make_epi_stage(sEpi + eg * Cfg::EPI_GROUP_BYTES, Cfg::EPI_GROUP_BYTES, F32OUT && p.out_f32 != nullptr);
    est.bar = 2 + 2 * eg;
    if (ep_tid == 0) {
      prefetch_tensormap(&tmOb);
      prefetch_tensormap(&tmOf);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
      const int m_tile = (tile / p.num_n_tiles) * PAIR_M + (int)rank * 128;
      float* sc = sScale + acc * BLOCK_N;
      float* sf = sShift + acc * BLOCK_N;
      for (int i = ep_all; i < BLOCK_N; i += 256) {
        sc[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.f;
        sf[i] = p.shift ? __ldg(p.shift + n0 + i) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int a_begin = MT == 2 ? eg : 0, a_end = MT == 2 ? eg + 1 : MT;
      const int c_begin = MT == 2 ? 0 : eg * (BLOCK_N / 2), c_end = MT == 2 ? BLOCK_N : (eg + 1) * (BLOCK_N / 2);
#pragma unroll 1
      for (int a = a_begin; a < a_end; ++a) {
        const int m0 = m_tile + a * 256;
        const int m = m0 + row;
        const bool valid = m < p.M;
        const size_t row_off = (size_t)m * p.Cout + n0;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * MT + a) * BLOCK_N + c0), v);
          tmem_ld_wait();
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaf(__uint_as_float(v[j]), sc[c0 + j], sf[c0 + j]);
          if (p.residual && valid) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 rv = __ldg(rp + u);
              const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(w4[e]);
                y[u * 8 + e * 2] += f.x;
                y[u * 8 + e * 2 + 1] += f.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
          }
          // registers -> swizzled smem slab -> TMA store (rows beyond M are clipped by the TMA unit; a slab that
          // starts beyond M stores nothing - every thread still takes part in the slab barriers)
          if (m0 < p.M) {
            if constexpr (F32OUT) {
              if (p.out_f32) epi_stage_store<true>(est, &tmOf, y, row, n0 + c0, m0, ep_tid);
            }
            if (p.out_bf16) epi_stage_store<false>(est, &tmOb, y, row, n0 + c0, m0, ep_tid);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[acc]), 0));
      if (++acc == Cfg::NUM_ACC) { acc = 0; acc_phase ^= 1; }
    }
    if (ep_tid == 0) bulk_wait_all();   // the last output slabs have left shared memory and are globally visible
  }

  // neither CTA may leave (or free TMEM) while its peer can still read its shared memory / signal its barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

static TmaEncoders g_enc2;

template <int BLOCK_N, int MT = 1, bool F32OUT = false>
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOb, const CUtensorMap& tmOf,
                      const CUtensorMap& tmA2, const Tc2Params& p, cudaStream_t st) {
  using Cfg = Tc2Cfg<BLOCK_N, MT, F32OUT>;
  static unsigned long long smem_set = 0;     // one per template instantiation
  {
    const int rc_attr = ensure_dynamic_smem((const void*)conv_gemm_bf16_2cta_kernel<BLOCK_N, MT, F32OUT>, Cfg::SMEM_BYTES, &smem_set);
    if (rc_attr != MSTR_OK) return rc_attr;
  }
  const int tiles = p.num_m_pairs * p.num_n_tiles;
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  cfg.blockDim = dim3(T2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MSTR_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_bf16_2cta_kernel<BLOCK_N, MT, F32OUT>, tmA, tmB, tmOb, tmOf, tmA2, p));
  MSTR_LAUNCH_CHECK("conv_gemm_bf16_2cta_kernel");
  return MSTR_OK;
}

// Called by mstr_conv_gemm_bf16 (conv_tc.cu) when the layer qualifies: C_in % 64 == 0, C_out % 256 == 0.
int conv_gemm_bf16_2cta(const void* x, const void* w, const float* scale, const float* shift, const void* residual,
                        void* out_bf16, float* out_f32, int B, int H, int W, int C_in, int C_out, int R, int S,
                        int stride_h, int stride_w, int pad_h, int pad_w, int relu, const void* x2, int C_in2,
                        cudaStream_t st) {
  int rc = load_tma_encoders(g_enc2);
  if (rc != MSTR_OK) return rc;
  Tc2Params p;
  p.P = (H + 2 * pad_h - R) / stride_h + 1;
  p.Q = (W + 2 * pad_w - S) / stride_w + 1;
  p.M = (int)((int64_t)B * p.P * p.Q);
  p.Cout = C_out;
  const int block_n = 256, mt = 1;
  p.num_m_pairs = ceil_div(p.M, 256 * mt);
  p.num_n_tiles = C_out / block_n;
  p.taps = R * S;
  p.S = S;
  p.cin_chunks = C_in / T2_BLOCK_K;
  p.cin2_chunks = x2 ? C_in2 / T2_BLOCK_K : 0;
  p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w;
  p.im2col = !(R == 1 && S == 1 && stride_h == 1 && stride_w == 1 && pad_h == 0 && pad_w == 0);
  p.relu = relu;
  p.scale = scale; p.shift = shift;
  p.residual = (const __nv_bfloat16*)residual;
  p.out_bf16 = (__nv_bfloat16*)out_bf16;
  p.out_f32 = out_f32;

  alignas(64) CUtensorMap tmA, tmB;
  const int64_t K = (int64_t)R * S * C_in + (x2 ? C_in2 : 0);
  const int umma_n = block_n > 256 ? 256 : block_n;
  rc = encode_tiled_2d_bf16(g_enc2, &tmB, w, (uint64_t)K, (uint64_t)C_out, T2_BLOCK_K, (uint32_t)(umma_n / 2),
                            CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  if (!p.im2col) {
    rc = encode_tiled_2d_bf16(g_enc2, &tmA, x, (uint64_t)C_in, (uint64_t)((int64_t)B * H * W), T2_BLOCK_K, 128,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != MSTR_OK) return rc;
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)C_in, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C_in * 2, (cuuint64_t)W * C_in * 2, (cuuint64_t)H * W * C_in * 2};
    int lower[2] = {-pad_w, -pad_h};
    int upper[2] = {pad_w - (S - 1), pad_h - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
    CUresult r = g_enc2.im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower,
                               upper, T2_BLOCK_K, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
    fix_small_im2col_map(&tmA, (int64_t)B * H * W * C_in * 2);
  }
  alignas(64) CUtensorMap tmA2 = tmA;
  if (x2) {   // [M][C_in2] matrix whose rows are the output pixels
    rc = encode_tiled_2d_bf16(g_enc2, &tmA2, x2, (uint64_t)C_in2, (uint64_t)p.M, T2_BLOCK_K, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != MSTR_OK) return rc;
  }
  alignas(64) CUtensorMap tmOb, tmOf;
  if (out_bf16) { rc = encode_output_2d(g_enc2, &tmOb, out_bf16, false, (uint64_t)C_out, (uint64_t)p.M); if (rc != MSTR_OK) return rc; }
  if (out_f32) { rc = encode_output_2d(g_enc2, &tmOf, out_f32, true, (uint64_t)C_out, (uint64_t)p.M); if (rc != MSTR_OK) return rc; }
  if (!out_bf16) tmOb = tmOf;
  if (!out_f32) tmOf = tmOb;
  if (out_f32) return launch_tc2<256, 1, true>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
  return launch_tc2<256>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
}

}  // namespace mstr
