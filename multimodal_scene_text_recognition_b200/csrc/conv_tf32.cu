// TF32 parity-mode implicit-GEMM convolution on the tensor cores: error-compensated 3xTF32.
//
// The north star asks for "a TF32 mode" whose results match the fp32 reference (ids identical,
// logits within 1e-4).  One TF32 product keeps 10 mantissa bits, which on this network flips ids on
// 28 % of the crops, so every operand is split into a TF32-exact head and a tail,
//     x = x_hi + x_lo,  x_hi = x with the low 13 mantissa bits cleared,  x_lo = x - x_hi  (exact),
// and each product is formed as  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi  with fp32 accumulation in TMEM;
// the dropped x_lo*w_lo term and the truncation of the tails are ~2^-21 relative.  Three
// tcgen05.mma.kind::tf32 (128x128x8) per K-step instead of one: 1/6 of the bf16 rate, ~10x the fp32
// SIMT kernel.
//
// Accumulation: the tensor core adds into its fp32 accumulator with truncation, so a single TMEM
// accumulation over K = 4608 drifts by ~2.5e-5 of the result (measured), an order of magnitude more than
// an fp32 FMA chain.  K is therefore cut into chunks of TF_CHUNK_KB k-blocks (256 values): each chunk
// accumulates in a fresh TMEM buffer and the epilogue threads add the chunk results into fp32 registers
// (round-to-nearest), the same two-level scheme fp8 GEMMs use.  The two TMEM buffers alternate per
// chunk, so the promotion of chunk i overlaps the MMAs of chunk i+1.
//
// Operands stay plain fp32 in global memory (NHWC activations, [C_out][R][S][C_in] weights; the
// weight split is done once on the host).  Per k-block of 32 channels the producer warp TMA-loads
// the raw fp32 activation tile (im2col map, 128 pixels x 128 bytes) plus the w_hi / w_lo tiles; four
// splitter warps turn the raw tile into x_hi / x_lo tiles IN PLACE of the swizzled layout (the split
// is element-wise, so the 128B-swizzle pattern written by TMA is preserved), then the MMA warp issues
// the twelve MMAs.  Epilogue as in conv_tc.cu (folded BN, residual, ReLU, smem slab, TMA store), fp32.
//
// Replaces the same reference call sites as conv_gemm_bf16_kernel; used when config.PRECISION ==
// 'tf32' for layers with C_in % 32 == 0 and C_out % 128 == 0 (the rest of that mode runs the fp32
// SIMT kernels).
#include "tc_common.cuh"

namespace mstr {

struct TfParams {
  int M, Cout;
  int num_m_tiles, num_n_tiles;
  int taps, S, cin_chunks;
  int P, Q, sh, sw, ph, pw;
  int im2col;
  int relu;
  const float* scale;
  const float* shift;
  const float* residual;
};

constexpr int TF_THREADS = 320;           // producer, MMA, 4 epilogue warps, 4 splitter warps
constexpr int TF_TILE = 128 * 128;        // 128 rows x 32 fp32 (one 128-byte swizzle span)
constexpr int TF_STAGE = 5 * TF_TILE;     // x_raw, x_hi, x_lo, w_hi, w_lo
constexpr int TF_STAGES = 2;
constexpr int TF_CHUNK_KB = 2;             // k-blocks (of 32 values) per TMEM accumulation chunk
constexpr int TF_SMEM_BYTES = 1024 + TF_STAGES * TF_STAGE + 2 * EPI_SLAB_BYTES + 4 * 128 * 4 + 16 * 8 + 16;

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: D = f32 (bits [4,6) = 1), A = B = TF32 (format 2 at [7,10) and [10,13)), K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(TF_THREADS, 1)
conv_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
                        const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CUtensorMap tmO,
                        const TfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sEpi = smem + TF_STAGES * TF_STAGE;
  float* sScale = reinterpret_cast<float*>(sEpi + 2 * EPI_SLAB_BYTES);   // [2][128]
  float* sShift = sScale + 2 * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sShift + 2 * 128);   // TMA landed (raw x, w_hi, w_lo)
  uint64_t* ready_bar = full_bar + TF_STAGES;                             // x split into hi / lo
  uint64_t* empty_bar = ready_bar + TF_STAGES;                            // MMAs done with the stage
  uint64_t* tfull_bar = empty_bar + TF_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_k = p.taps * p.cin_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmWhi);
    prefetch_tensormap(&tmWlo);
    prefetch_tensormap(&tmO);
    for (int i = 0; i < TF_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&ready_bar[i], 4);   // one arrive per splitter warp
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.num_n_tiles) * 128;
        const int m0 = (tile / p.num_n_tiles) * 128;
        int cw = 0, chh = 0, cn = 0;
        if (p.im2col) {
          const int q = m0 % p.Q, pp = (m0 / p.Q) % p.P;
          cn = m0 / (p.Q * p.P);
          cw = q * p.sw - p.pw;
          chh = pp * p.sh - p.ph;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            uint8_t* st = smem + stage * TF_STAGE;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], 3 * TF_TILE);
            if (p.im2col)
              tma_load_im2col_4d(st, &tmX, &full_bar[stage], cc * 32, cw, chh, cn, (uint16_t)s, (uint16_t)r);
            else
              tma_load_2d(st, &tmX, &full_bar[stage], cc * 32, m0);
            const int kcol = (tap * p.cin_chunks + cc) * 32;
            tma_load_2d(st + 3 * TF_TILE, &tmWhi, &full_bar[stage], kcol, n0);
            tma_load_2d(st + 4 * TF_TILE, &tmWlo, &full_bar[stage], kcol, n0);
            if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 128);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < num_k; kb0 += TF_CHUNK_KB) {
          const int kb1 = min(kb0 + TF_CHUNK_KB, num_k);
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&ready_bar[stage], phase);
            tc_fence_after();
            const uint32_t base = smem_u32(smem + stage * TF_STAGE);
            const uint32_t x_hi = base + 1 * TF_TILE, x_lo = base + 2 * TF_TILE;
            const uint32_t w_hi = base + 3 * TF_TILE, w_lo = base + 4 * TF_TILE;
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 8 fp32 = 32 bytes per K-step
              const uint64_t dxh = make_smem_desc<128>(x_hi + k * 32), dxl = make_smem_desc<128>(x_lo + k * 32);
              const uint64_t dwh = make_smem_desc<128>(w_hi + k * 32), dwl = make_smem_desc<128>(w_lo + k * 32);
              umma_tf32(d_tmem, dxl, dwh, idesc, (kb > kb0 || k > 0) ? 1u : 0u);   // small terms first
              umma_tf32(d_tmem, dxh, dwl, idesc, 1u);
              umma_tf32(d_tmem, dxh, dwh, idesc, 1u);
            }
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
            if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 6) {
    // ===================== splitter: raw fp32 tile -> TF32 head + tail (same swizzled positions) =====================
    const int sp_tid = threadIdx.x - 192;   // 0..127
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_k; ++kb) {
        uint8_t* st = smem + stage * TF_STAGE;
        mbar_wait(&full_bar[stage], phase);
        const uint4* raw = reinterpret_cast<const uint4*>(st);
        uint4* hi = reinterpret_cast<uint4*>(st + TF_TILE);
        float4* lo = reinterpret_cast<float4*>(st + 2 * TF_TILE);
#pragma unroll
        for (int i = 0; i < TF_TILE / 16 / 128; ++i) {
          const int idx = sp_tid + i * 128;
          const uint4 v = raw[idx];
          const uint4 h = make_uint4(v.x & 0xFFFFE000u, v.y & 0xFFFFE000u, v.z & 0xFFFFE000u, v.w & 0xFFFFE000u);
          hi[idx] = h;
          lo[idx] = make_float4(__uint_as_float(v.x) - __uint_as_float(h.x), __uint_as_float(v.y) - __uint_as_float(h.y),
                                __uint_as_float(v.z) - __uint_as_float(h.z), __uint_as_float(v.w) - __uint_as_float(h.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[stage]);
        if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int ep_tid = threadIdx.x - 64;
    EpiStage est = make_epi_stage(sEpi, 2 * EPI_SLAB_BYTES, true);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n0 = (tile % p.num_n_tiles) * 128;
      const int m0 = (tile / p.num_n_tiles) * 128;
      float* sc = sScale + acc * 128;
      float* sf = sShift + acc * 128;
      for (int i = ep_tid; i < 128; i += 128) {
        sc[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.f;
        sf[i] = p.shift ? __ldg(p.shift + n0 + i) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // two-level accumulation: add every K chunk's TMEM result into fp32 registers
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.f;
      for (int kb0 = 0; kb0 < num_k; kb0 += TF_CHUNK_KB) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128 + c0), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      const int m = m0 + row;
      const bool valid = m < p.M;
      const size_t row_off = (size_t)m * p.Cout + n0;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = fmaf(sum[c0 + j], sc[c0 + j], sf[c0 + j]);
        if (p.residual && valid) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual + row_off + c0);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 rv = __ldg(rp + u);
            y[4 * u] += rv.x; y[4 * u + 1] += rv.y; y[4 * u + 2] += rv.z; y[4 * u + 3] += rv.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        epi_stage_store<true>(est, &tmO, y, row, n0 + c0, m0, ep_tid);
      }
    }
    if (ep_tid == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static TmaEncoders g_enc_tf;

static int encode_f32_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 4};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_enc_tf.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed with CUresult %d", (int)r);
  return MSTR_OK;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_conv_gemm_tf32(
    const float* x, const float* w_hi, const float* w_lo, const float* scale, const float* shift,
    const float* residual, float* out, int B, int H, int W, int C_in, int C_out, int R, int S, int stride_h,
    int stride_w, int pad_h, int pad_w, int relu, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && w_hi && w_lo && out, MSTR_ERR_BAD_ARG, "mstr_conv_gemm_tf32: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0 && C_in > 0 && C_out > 0 && R > 0 && S > 0 && stride_h > 0 && stride_w > 0 &&
                   pad_h >= 0 && pad_w >= 0,
               MSTR_ERR_BAD_ARG, "mstr_conv_gemm_tf32: bad size");
  MSTR_REQUIRE(C_in % 32 == 0, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_tf32: C_in=%d must be a multiple of 32", C_in);
  MSTR_REQUIRE(C_out % 128 == 0, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_tf32: C_out=%d must be a multiple of 128", C_out);
  MSTR_REQUIRE(aligned16(x) && aligned16(w_hi) && aligned16(w_lo) && aligned16(residual) && aligned16(out),
               MSTR_ERR_BAD_ARG, "mstr_conv_gemm_tf32: pointers must be 16-byte aligned");
  MSTR_REQUIRE(R <= 16 && S <= 16 && pad_h < 16 && pad_w < 16, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_tf32: filter too large");
  int rc = load_tma_encoders(g_enc_tf);
  if (rc != MSTR_OK) return rc;

  TfParams p;
  p.P = (H + 2 * pad_h - R) / stride_h + 1;
  p.Q = (W + 2 * pad_w - S) / stride_w + 1;
  MSTR_REQUIRE(p.P > 0 && p.Q > 0, MSTR_ERR_BAD_ARG, "mstr_conv_gemm_tf32: empty output");
  const int64_t M64 = (int64_t)B * p.P * p.Q;
  MSTR_REQUIRE(M64 < (1ll << 31), MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_tf32: too many output pixels");
  p.M = (int)M64;
  p.Cout = C_out;
  p.num_m_tiles = ceil_div(p.M, 128);
  p.num_n_tiles = C_out / 128;
  p.taps = R * S;
  p.S = S;
  p.cin_chunks = C_in / 32;
  p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w;
  p.im2col = !(R == 1 && S == 1 && stride_h == 1 && stride_w == 1 && pad_h == 0 && pad_w == 0);
  p.relu = relu;
  p.scale = scale; p.shift = shift; p.residual = residual;

  alignas(64) CUtensorMap tmX, tmWhi, tmWlo, tmO;
  const int64_t K = (int64_t)R * S * C_in;
  rc = encode_f32_2d(&tmWhi, w_hi, (uint64_t)K, (uint64_t)C_out, 128);
  if (rc != MSTR_OK) return rc;
  rc = encode_f32_2d(&tmWlo, w_lo, (uint64_t)K, (uint64_t)C_out, 128);
  if (rc != MSTR_OK) return rc;
  rc = encode_output_2d(g_enc_tf, &tmO, out, true, (uint64_t)C_out, (uint64_t)p.M);
  if (rc != MSTR_OK) return rc;
  if (!p.im2col) {
    rc = encode_f32_2d(&tmX, x, (uint64_t)C_in, (uint64_t)((int64_t)B * H * W), 128);
    if (rc != MSTR_OK) return rc;
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)C_in, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C_in * 4, (cuuint64_t)W * C_in * 4, (cuuint64_t)H * W * C_in * 4};
    int lower[2] = {-pad_w, -pad_h};
    int upper[2] = {pad_w - (S - 1), pad_h - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
    CUresult r = g_enc_tf.im2col(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, lower,
                                 upper, 32, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeIm2col(f32) failed with CUresult %d", (int)r);
    fix_small_im2col_map(&tmX, (int64_t)B * H * W * C_in * 4);
  }
  static unsigned long long smem_set = 0;
  {
    const int rc_attr = ensure_dynamic_smem((const void*)conv_gemm_tf32x3_kernel, TF_SMEM_BYTES, &smem_set);
    if (rc_attr != MSTR_OK) return rc_attr;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  conv_gemm_tf32x3_kernel<<<grid, TF_THREADS, TF_SMEM_BYTES, (cudaStream_t)stream>>>(tmX, tmWhi, tmWlo, tmO, p);
  MSTR_LAUNCH_CHECK("conv_gemm_tf32x3_kernel");
  return MSTR_OK;
}
