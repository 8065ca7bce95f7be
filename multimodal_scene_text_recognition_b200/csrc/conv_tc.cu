// bf16 implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// out[m][n] = act( (sum_k A[m][k] W[n][k]) * scale[n] + shift[n] + residual[m][n] )
//   m = (b,p,q) output pixel, k = (r,s,c);  A is never materialised: the producer warp issues
//   TMA *im2col* loads (cp.async.bulk.tensor.4d...im2col) straight from the NHWC activation
//   tensor, one 128-pixel x BLOCK_K-channel box per filter tap, zero-filling the halo in hardware.
//   1x1 / linear layers use a plain 2-D tiled map instead.
// Reference call sites replaced: every nn.Conv2d(bias=False)+BatchNorm2d(eval)[+residual][+ReLU]
// with C_in >= 32 in modules/feature_extraction.py:32-47,109-150 and
// modules/transformation.py:52-57, and nn.Linear in modules/encoders.py:15,28 /
// modules/decoders.py:71,81.
//
// Kernel anatomy (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer  : STAGES-deep smem ring, mbarrier full/empty
//   warp 1      MMA issuer    : one thread issues tcgen05.mma.cta_group::1.kind::f16 (128 x BLOCK_N x 16),
//                               accumulators in TMEM (2 x BLOCK_N fp32 columns, double buffered)
//   warps 2..5  epilogue      : tcgen05.ld 32x32b.x32 -> folded BN, residual, ReLU -> swizzled smem slab -> TMA store
//                               (bf16 and/or f32 NHWC), overlapped with the next tile's MMAs through the second
//                               TMEM buffer
// This single-CTA kernel serves the layers the CTA-pair kernel (conv_tc2.cu) does not take: C_out = 128 layers
// (MT = 2: two 128-row accumulators per CTA share each weight tile) and small / odd shapes.
// Roofline: tensor pipe, fed from L2.  Algorithmic FLOPs = 2 * M * C_out * R*S*C_in per launch.
#include <stdlib.h>
#include <atomic>

#include "tc_common.cuh"

namespace mstr {

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct TcParams {
  int M, Cout;
  int num_m_tiles, num_n_tiles;
  int taps, S, cin_chunks;  // K loop = taps x cin_chunks blocks of BLOCK_K channels
  int cin2_chunks;          // + extra k-blocks from a second [M][C_in2] matrix (fused 1x1 branch), 0 = none
  int P, Q, sh, sw, ph, pw;
  int im2col;               // 1: A via im2col map; 0: A via 2-D tiled map [M][K]
  int relu;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out_bf16;
  float* out_f32;
};

constexpr int TC_BLOCK_M = 128;
constexpr int TC_THREADS = 320;     // producer, MMA issuer, 2 x 4 epilogue warps

constexpr int TC_MAX_STAGES = 8;   // barrier storage is sized for the maximum ring depth

// MT ("M tiles per CTA"): with MT = 2 a CTA owns 256 output rows as two 128-row accumulators that share every
// weight tile.  The C_out = 128 layers are bound by operand traffic from L2 (32 KB per 128x128x64 block);
// sharing the weight tile cuts that by a quarter and doubles the tensor-core work per barrier round trip.
template <int BLOCK_N, int BLOCK_K, int MT = 1>
struct TcCfg {
  static constexpr int ROW_BYTES = BLOCK_K * 2;
  static constexpr int A_HALF = TC_BLOCK_M * ROW_BYTES;
  static constexpr int A_BYTES = MT * A_HALF;
  static constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_SLAB = EPI_SLAB_BYTES;
  static constexpr int TAIL_BYTES = 4 * BLOCK_N * 4 /*scale/shift x2*/ + (2 * TC_MAX_STAGES + 5) * 8 + 16;
  static constexpr int RING_BUDGET = 227 * 1024 - 1024 /*align slack*/ - 2 * EPI_SLAB - TAIL_BYTES;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * MT * BLOCK_N < 32 ? 32 : 2 * MT * BLOCK_N;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 2 * EPI_SLAB + TAIL_BYTES;
};

template <int BLOCK_N, int BLOCK_K, int MT = 1>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOb, const __grid_constant__ CUtensorMap tmOf,
                      const __grid_constant__ CUtensorMap tmA2, const TcParams p) {
  using Cfg = TcCfg<BLOCK_N, BLOCK_K, MT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TILE_M = TC_BLOCK_M * MT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * Cfg::A_BYTES;
  uint8_t* sEpi = sA + STAGES * Cfg::STAGE_BYTES;       // 2 output slabs (1024-byte aligned: every part above is)
  float* sScale = reinterpret_cast<float*>(sEpi + 2 * Cfg::EPI_SLAB);          // [2][BLOCK_N]
  float* sShift = sScale + 2 * BLOCK_N;                                        // [2][BLOCK_N]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sShift + 2 * BLOCK_N);
  uint64_t* empty_bar = full_bar + TC_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + TC_MAX_STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_k = p.taps * p.cin_chunks + p.cin2_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
        const int m0 = (tile / p.num_n_tiles) * TILE_M;
        int cw[MT], chh[MT], cn[MT];
#pragma unroll
        for (int h = 0; h < MT; ++h) {
          cw[h] = chh[h] = cn[h] = 0;
          if (p.im2col) {
            const int mh = m0 + h * TC_BLOCK_M;
            const int q = mh % p.Q, pp = (mh / p.Q) % p.P;
            cn[h] = mh / (p.Q * p.P);
            cw[h] = q * p.sw - p.pw;
            chh[h] = pp * p.sh - p.ph;
          }
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
            for (int h = 0; h < MT; ++h) {
              uint8_t* dstA = sA + stage * Cfg::A_BYTES + h * Cfg::A_HALF;
              if (p.im2col)
                tma_load_im2col_4d(dstA, &tmA, &full_bar[stage], cc * BLOCK_K, cw[h], chh[h], cn[h], (uint16_t)s,
                                   (uint16_t)r);
              else
                tma_load_2d(dstA, &tmA, &full_bar[stage], cc * BLOCK_K, m0 + h * TC_BLOCK_M);
            }
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], (tap * p.cin_chunks + cc) * BLOCK_K, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        // fused 1x1 branch: its input rows are the same output pixels, its weights follow in K
        for (int cc = 0; cc < p.cin2_chunks; ++cc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
          for (int h = 0; h < MT; ++h)
            tma_load_2d(sA + stage * Cfg::A_BYTES + h * Cfg::A_HALF, &tmA2, &full_bar[stage], cc * BLOCK_K,
                        m0 + h * TC_BLOCK_M);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], (p.taps * p.cin_chunks + cc) * BLOCK_K, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * BLOCK_N);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t db = make_smem_desc<Cfg::ROW_BYTES>(b_addr + k * 32);
#pragma unroll
            for (int h = 0; h < MT; ++h) {
              const uint64_t da = make_smem_desc<Cfg::ROW_BYTES>(a_addr + h * Cfg::A_HALF + k * 32);
              umma_bf16(d_tmem + (uint32_t)(h * BLOCK_N), da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem slot once these MMAs retire
          if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    // Two groups of four warps (TMEM lane quarters 2,3,0,1 each) split the tile: with MT = 2 group g drains
    // accumulator g, otherwise column half g.  One group needed ~10k cycles per 256-row tile (ncu: TMEM load
    // latency, fence.proxy.async, the slab barriers and the per-tile fences add up to ~1.2k cycles per 32-column
    // step), more than the MMAs of every K <= 1152 layer; each group has its own slab ring and named barriers.
    const int eg = (warp - 2) >> 2;
    const int ep_tid = (threadIdx.x - 64) & 127;
    const int ep_all = threadIdx.x - 64;
    EpiStage est = make_epi_stage(sEpi + eg * Cfg::EPI_SLAB, Cfg::EPI_SLAB, p.out_f32 != nullptr);   // 1 f32 / 2 bf16 slabs
    est.bar = 2 + 2 * eg;
    if (ep_tid == 0) {
      prefetch_tensormap(&tmOb);
      prefetch_tensormap(&tmOf);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
      const int m0 = (tile / p.num_n_tiles) * TILE_M;
      float* sc = sScale + acc * BLOCK_N;
      float* sf = sShift + acc * BLOCK_N;
      for (int i = ep_all; i < BLOCK_N; i += 256) {
        sc[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.f;
        sf[i] = p.shift ? __ldg(p.shift + n0 + i) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int h_begin = MT == 2 ? eg : 0, h_end = MT == 2 ? eg + 1 : MT;
      const int c_begin = MT == 2 ? 0 : eg * (BLOCK_N / 2), c_end = MT == 2 ? BLOCK_N : (eg + 1) * (BLOCK_N / 2);
#pragma unroll 1
      for (int h = h_begin; h < h_end; ++h) {
        const int mh = m0 + h * TC_BLOCK_M;
        const int m = mh + row;
        const bool valid = m < p.M;
        const size_t row_off = (size_t)m * p.Cout + n0;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * MT + h) * BLOCK_N + c0), v);
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
          // rows beyond M are clipped by the TMA store (a tile that starts beyond M stores nothing)
          if (mh < p.M) {
            if (p.out_f32) epi_stage_store<true>(est, &tmOf, y, row, n0 + c0, mh, ep_tid);
            if (p.out_bf16) epi_stage_store<false>(est, &tmOb, y, row, n0 + c0, mh, ep_tid);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (ep_tid == 0) bulk_wait_all();   // the last output slabs have left shared memory and are globally visible
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
static TmaEncoders g_enc;
static std::atomic<int> g_gemm_route{0};   // test hook (mstr_debug_gemm_route): 0 = default dispatch, 1 / 2 force a kernel

template <int BLOCK_N, int BLOCK_K, int MT = 1>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOb, const CUtensorMap& tmOf,
                     const CUtensorMap& tmA2, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N, BLOCK_K, MT>;
  static unsigned long long smem_set = 0;     // one per template instantiation
  {
    const int rc_attr = ensure_dynamic_smem((const void*)conv_gemm_bf16_kernel<BLOCK_N, BLOCK_K, MT>, Cfg::SMEM_BYTES, &smem_set);
    if (rc_attr != MSTR_OK) return rc_attr;
  }
  TcParams q = p;
  q.num_m_tiles = ceil_div(p.M, TC_BLOCK_M * MT);
  const int tiles = q.num_m_tiles * q.num_n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  conv_gemm_bf16_kernel<BLOCK_N, BLOCK_K, MT><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, tmOb, tmOf, tmA2, q);
  MSTR_LAUNCH_CHECK("conv_gemm_bf16_kernel");
  return MSTR_OK;
}

// conv_tc2.cu: CTA-pair (cta_group::2) variant
int conv_gemm_bf16_2cta(const void* x, const void* w, const float* scale, const float* shift, const void* residual,
                        void* out_bf16, float* out_f32, int B, int H, int W, int C_in, int C_out, int R, int S,
                        int stride_h, int stride_w, int pad_h, int pad_w, int relu, const void* x2, int C_in2,
                        cudaStream_t st);

static int conv_gemm_bf16_impl(const void* x, const void* w, const float* scale, const float* shift,
                               const void* residual, void* out_bf16, float* out_f32, int B, int H, int W,
                               int C_in, int C_out, int R, int S, int stride_h, int stride_w, int pad_h,
                               int pad_w, int relu, const void* x2, int C_in2, mstr_stream_t stream) {
  MSTR_REQUIRE(x && w && (out_bf16 || out_f32), MSTR_ERR_BAD_ARG, "mstr_conv_gemm_bf16: null pointer");
  MSTR_REQUIRE(!x2 || (C_in2 > 0 && C_in2 % 64 == 0 && C_in % 64 == 0 && aligned16(x2)), MSTR_ERR_UNSUPPORTED,
               "mstr_conv_gemm_bf16_dual: the fused 1x1 branch needs C_in %% 64 == 0 and C_in2 %% 64 == 0");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0 && C_in > 0 && C_out > 0 && R > 0 && S > 0 && stride_h > 0 && stride_w > 0 &&
                   pad_h >= 0 && pad_w >= 0,
               MSTR_ERR_BAD_ARG, "mstr_conv_gemm_bf16: bad size");
  MSTR_REQUIRE(C_in % 32 == 0, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_bf16: C_in=%d must be a multiple of 32", C_in);
  MSTR_REQUIRE(C_out % 64 == 0, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_bf16: C_out=%d must be a multiple of 64", C_out);
  MSTR_REQUIRE(aligned16(x) && aligned16(w) && aligned16(residual) && aligned16(out_bf16) && aligned16(out_f32),
               MSTR_ERR_BAD_ARG, "mstr_conv_gemm_bf16: pointers must be 16-byte aligned");
  MSTR_REQUIRE(R <= 16 && S <= 16 && pad_h < 16 && pad_w < 16, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_bf16: filter too large");
  // CTA-pair kernel (conv_tc2.cu) for C_out % 256 == 0: +9 % on the 512->512 layers.  C_out = 128 layers stay on the
  // single-CTA kernel below (pairs measured 8 % slower there; DESIGN.md 4 lists the variants that were tried and removed)
  // ... except when the whole problem is at most ONE wave of 128 x 128 tiles (the linear layers of a decode step, small
  // shards): the single-CTA kernel's smaller tiles then finish sooner - 7.3 vs 10.9 us for [4096 x 256] x [256 x 256],
  // equal from two waves on - and the results are bit-identical (same K order per output element;
  // tests/test_gpu_kernels.py runs both on the same operands through mstr_debug_gemm_route).
  const int route = g_gemm_route.load(std::memory_order_relaxed);   // 0 = this rule, 1 = single CTA only, 2 = pairs wherever they apply
  const int64_t M_out = (int64_t)B * ((H + 2 * pad_h - R) / stride_h + 1) * ((W + 2 * pad_w - S) / stride_w + 1);
  const bool one_wave = C_out % 128 == 0 && ceil_div64(M_out, 128) * (C_out / 128) <= sm_count();
  const bool single = route == 1 || (route == 0 && one_wave);
  if (!single && C_in % 64 == 0 && C_out % 256 == 0 && (int64_t)B * H * W < (1ll << 31) && sm_count() >= 2) {
    const int64_t M2 = (int64_t)B * ((H + 2 * pad_h - R) / stride_h + 1) * ((W + 2 * pad_w - S) / stride_w + 1);
    MSTR_REQUIRE(M2 > 0 && M2 < (1ll << 31), MSTR_ERR_BAD_ARG, "mstr_conv_gemm_bf16: bad output size");
    return conv_gemm_bf16_2cta(x, w, scale, shift, residual, out_bf16, out_f32, B, H, W, C_in, C_out, R, S, stride_h,
                               stride_w, pad_h, pad_w, relu, x2, C_in2, (cudaStream_t)stream);
  }
  int rc = load_tma_encoders(g_enc);
  if (rc != MSTR_OK) return rc;

  TcParams p;
  p.P = (H + 2 * pad_h - R) / stride_h + 1;
  p.Q = (W + 2 * pad_w - S) / stride_w + 1;
  MSTR_REQUIRE(p.P > 0 && p.Q > 0, MSTR_ERR_BAD_ARG, "mstr_conv_gemm_bf16: empty output");
  const int64_t M64 = (int64_t)B * p.P * p.Q;
  MSTR_REQUIRE(M64 < (1ll << 31), MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_bf16: too many output pixels");
  p.M = (int)M64;
  p.Cout = C_out;
  const int block_k = (C_in % 64 == 0) ? 64 : 32;
  const int block_n = (C_out % 256 == 0 && !single) ? 256 : (C_out % 128 == 0 ? 128 : 64);
  p.num_m_tiles = ceil_div(p.M, TC_BLOCK_M);
  p.num_n_tiles = C_out / block_n;
  p.taps = R * S;
  p.S = S;
  p.cin_chunks = C_in / block_k;
  p.cin2_chunks = x2 ? C_in2 / block_k : 0;
  p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w;
  p.im2col = !(R == 1 && S == 1 && stride_h == 1 && stride_w == 1 && pad_h == 0 && pad_w == 0);
  p.relu = relu;
  p.scale = scale; p.shift = shift;
  p.residual = (const __nv_bfloat16*)residual;
  p.out_bf16 = (__nv_bfloat16*)out_bf16;
  p.out_f32 = out_f32;

  const CUtensorMapSwizzle swz = block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  alignas(64) CUtensorMap tmA, tmB;
  const int64_t K = (int64_t)R * S * C_in + (x2 ? C_in2 : 0);
  rc = encode_tiled_2d_bf16(g_enc, &tmB, w, (uint64_t)K, (uint64_t)C_out, (uint32_t)block_k, (uint32_t)block_n, swz);
  if (rc != MSTR_OK) return rc;
  if (!p.im2col) {
    // A is a plain [M][C_in] matrix (1x1 convolution / linear layer)
    rc = encode_tiled_2d_bf16(g_enc, &tmA, x, (uint64_t)C_in, (uint64_t)((int64_t)B * H * W), (uint32_t)block_k, TC_BLOCK_M, swz);
    if (rc != MSTR_OK) return rc;
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)C_in, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C_in * 2, (cuuint64_t)W * C_in * 2, (cuuint64_t)H * W * C_in * 2};
    // bounding box of filter-tap origins: [-pad, dim + pad - (filter-1) - 1]
    int lower[2] = {-pad_w, -pad_h};
    int upper[2] = {pad_w - (S - 1), pad_h - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
    CUresult r = g_enc.im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower,
                                 upper, (cuuint32_t)block_k, (cuuint32_t)TC_BLOCK_M, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
    fix_small_im2col_map(&tmA, (int64_t)B * H * W * C_in * 2);
  }

  alignas(64) CUtensorMap tmA2 = tmA;
  if (x2) {   // [M][C_in2] matrix whose rows are the output pixels
    rc = encode_tiled_2d_bf16(g_enc, &tmA2, x2, (uint64_t)C_in2, (uint64_t)p.M, (uint32_t)block_k, TC_BLOCK_M, swz);
    if (rc != MSTR_OK) return rc;
  }
  // output tensor maps for the epilogue's TMA stores (an absent output gets a copy of the other map; it is never used)
  alignas(64) CUtensorMap tmOb, tmOf;
  if (out_bf16) { rc = encode_output_2d(g_enc, &tmOb, out_bf16, false, (uint64_t)C_out, (uint64_t)p.M); if (rc != MSTR_OK) return rc; }
  if (out_f32) { rc = encode_output_2d(g_enc, &tmOf, out_f32, true, (uint64_t)C_out, (uint64_t)p.M); if (rc != MSTR_OK) return rc; }
  if (!out_bf16) tmOb = tmOf;
  if (!out_f32) tmOf = tmOb;
  cudaStream_t st = (cudaStream_t)stream;
  if (block_k == 64) {
    if (block_n == 256) return launch_tc<256, 64>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
    // two 128-row accumulators per CTA sharing each weight tile, once there are enough tiles to fill the device
    if (block_n == 128 && p.num_m_tiles >= 8 * sm_count())
      return launch_tc<128, 64, 2>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
    if (block_n == 128) return launch_tc<128, 64>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
    return launch_tc<64, 64>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
  } else {
    if (block_n == 256) return launch_tc<256, 32>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
    if (block_n == 128) return launch_tc<128, 32>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
    return launch_tc<64, 32>(tmA, tmB, tmOb, tmOf, tmA2, p, st);
  }
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_debug_gemm_route(int route) {
  mstr::g_gemm_route.store(route);
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_conv_gemm_bf16(
    const void* x, const void* w, const float* scale, const float* shift, const void* residual, void* out_bf16,
    float* out_f32, int B, int H, int W, int C_in, int C_out, int R, int S, int stride_h, int stride_w, int pad_h,
    int pad_w, int relu, mstr_stream_t stream) {
  return mstr::conv_gemm_bf16_impl(x, w, scale, shift, residual, out_bf16, out_f32, B, H, W, C_in, C_out, R, S, stride_h,
                                   stride_w, pad_h, pad_w, relu, nullptr, 0, stream);
}

extern "C" __attribute__((visibility("default"))) int mstr_conv_gemm_bf16_dual(
    const void* x, const void* x2, const void* w_cat, const float* scale, const float* shift, const void* residual,
    void* out_bf16, float* out_f32, int B, int H, int W, int C_in, int C_in2, int C_out, int R, int S, int stride_h,
    int stride_w, int pad_h, int pad_w, int relu, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x2, MSTR_ERR_BAD_ARG, "mstr_conv_gemm_bf16_dual: null x2");
  return conv_gemm_bf16_impl(x, w_cat, scale, shift, residual, out_bf16, out_f32, B, H, W, C_in, C_out, R, S, stride_h,
                             stride_w, pad_h, pad_w, relu, x2, C_in2, stream);
}
