// 3x3 / stride 1 / pad 1 convolution over a PADDED-FLAT activation on the 5th-generation tensor cores: every activation
// row is loaded ONCE per output tile and the nine filter taps are formed as shared-memory descriptor offsets.
//
// Why: the im2col kernels (conv_tc.cu / conv_tc2.cu) fetch one [128 pixels x C_in] box per filter tap, i.e. every
// activation row travels L2 -> shared memory nine times.  For the wide layers that is hidden behind the MMAs, for the
// narrow ones it is the limiter: conv0_2 (32 -> 64 channels at 32 x 100, modules/feature_extraction.py:61-63,113-115)
// ran as a 64 -> 128 convolution over horizontal pixel pairs (half of its MMA work structural zeros, 9 operand fetches
// per tile) at 1.15 ms per 4096-crop step for 0.48 TFLOP of useful work.
//
// Layout: the activation carries a one-pixel ZERO ring, [B][H+2][W+2][C] (what mstr_conv3x3_first writes with
// pad_out = 1), and is addressed as a flat matrix [NP = B (H+2) (W+2) positions][C].  With Wp = W + 2, the input of
// filter tap (r, s) for the output at flat position p is the row p + (r-1) Wp + (s-1) - the same offset for EVERY
// position, across image rows and across images (the ring supplies the zeros of the convolution's padding).  So:
//   * an output tile is 256 consecutive flat positions (two 128-row accumulators);
//   * the producer loads the rows [p0 - Wp - 1, p0 + 256 + Wp + 1) once (TMA tiled loads, rows outside the tensor are
//     zero-filled) into a slab of 64-byte rows (C_in = 32 bf16, 64B swizzle);
//   * for tap (r, s) the A operand of tcgen05.mma is the SAME slab at row offset r Wp + s: a shared-memory descriptor
//     whose start address is not aligned to the 8-row swizzle atom.  MEASURED on the B200 (tests/test_gpu_kernels.py::
//     test_padded_flat_conv_descriptor_base_offset): the tensor core applies the swizzle XOR to the bits of the final
//     shared-memory address, exactly as TMA did when it wrote the slab, so a row-shifted start address needs nothing
//     else - with the descriptor's "matrix base offset" field (bits 49-51) set to the phase of the start address the
//     result is WRONG (error 0.58 of the output range), with the field left at zero it is exact;
//   * the weights [9 taps][64 x 32] (36 KB) stay resident in shared memory for the life of the CTA;
//   * ring positions are computed like any other (10 % of the positions at 32 x 100) and written as ZEROS, so the
//     output is again a padded-flat activation; rows leave through swizzled slabs and TMA stores.
// Per 256-position tile: 30 KB of operand traffic (instead of 9 x 32 KB + 72 KB of weights) for 36 MMAs
// (128 x 64 x 16) - the kernel is bound by the tensor pipe, not by the L2 -> SM feed.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (four per accumulator); TMEM holds two
// tiles (2 x 2 x 64 columns) so the epilogue of tile t overlaps the MMAs of tile t + 1.
#include <atomic>

#include "tc_common.cuh"

namespace mstr {

constexpr int CF_CIN = 32;
constexpr int CF_COUT = 64;
constexpr int CF_TM = 256;                       // positions per tile (two accumulators of 128)
constexpr int CF_ROW_BYTES = CF_CIN * 2;         // 64
constexpr int CF_W_TAP_BYTES = CF_COUT * CF_ROW_BYTES;   // [64 output channels][32 input channels] bf16 = 4 KB
constexpr int CF_W_BYTES = 9 * CF_W_TAP_BYTES;   // 36 KB
constexpr int CF_EPI_WARPS = 8;
constexpr int CF_THREADS = 64 + 32 * CF_EPI_WARPS;
constexpr int CF_OUT_SLAB = 128 * CF_COUT * 2;   // [128 positions][64 channels] bf16: 128-byte rows, 128B swizzle = 16 KB
constexpr int CF_EPI_BYTES = 2 * 2 * CF_OUT_SLAB;             // two groups x two slabs
constexpr int CF_TAIL_BYTES = 16 * 8 + 16;
constexpr int CF_MAX_STAGES = 4;

struct CfParams {
  int NP, H, W, Hp, Wp;
  int num_tiles;
  int box_rows, nload, slab_bytes, stages;
  int relu;
  int desc_mode;                                 // 0: plain shifted start address (correct, default); 1: + base offset (experiment)
};
// Folded BatchNorm travels as a KERNEL PARAMETER: parameters live in the constant bank, and with the channel loop
// unrolled every scale / shift is a constant-bank operand of its FFMA - no load instruction.  (Staged in shared memory
// they cost 128 broadcast LDS per thread and tile: ncu showed the LSU shared-load pipe 24 % busy and the epilogue, at
// ~500 issued instructions per warp and tile, slower than the tile's MMAs.)
struct CfAffine {
  float scale[CF_COUT];
  float shift[CF_COUT];
};

__device__ __forceinline__ uint64_t cf_desc_a(uint32_t addr, int mode) {
  uint64_t d = make_smem_desc<CF_ROW_BYTES>(addr);
  if (mode) d |= (uint64_t)((addr >> 7) & 7u) << 49;     // experiment only: the "matrix base offset" field - wrong on B200
  return d;
}

__global__ void __launch_bounds__(CF_THREADS, 1)
conv3x3_flat_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CfAffine bn, const CfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                          // [9][64 rows][64 B], 64B swizzle
  uint8_t* sSlab = sW + CF_W_BYTES;                            // [stages][slab_bytes]
  uint8_t* sEpi = sSlab + p.stages * p.slab_bytes;             // 1024-aligned: slab_bytes is a multiple of 1024
  uint64_t* wfull_bar = reinterpret_cast<uint64_t*>(sEpi + CF_EPI_BYTES);
  uint64_t* sfull_bar = wfull_bar + 1;                         // [CF_MAX_STAGES]
  uint64_t* sempty_bar = sfull_bar + CF_MAX_STAGES;            // [CF_MAX_STAGES]
  uint64_t* tfull_bar = sempty_bar + CF_MAX_STAGES;            // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmO);
    mbar_init(wfull_bar, 1);
    for (int i = 0; i < CF_MAX_STAGES; ++i) { mbar_init(&sfull_bar[i], 1); mbar_init(&sempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], CF_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: the weights once, then one slab per tile =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull_bar, CF_W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + tap * CF_W_TAP_BYTES, &tmW, wfull_bar, tap * CF_CIN, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&sempty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&sfull_bar[stage], (uint32_t)(p.nload * p.box_rows * CF_ROW_BYTES));
        uint8_t* dst = sSlab + stage * p.slab_bytes;
        const int row0 = t * CF_TM - p.Wp - 1;                 // negative / beyond NP: zero-filled by TMA
        for (int i = 0; i < p.nload; ++i)
          tma_load_2d(dst + i * p.box_rows * CF_ROW_BYTES, &tmX, &sfull_bar[stage], 0, row0 + i * p.box_rows);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, CF_COUT);
      const uint32_t w_addr = smem_u32(sW);
      mbar_wait(wfull_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[buf], bphase ^ 1);               // the epilogue has drained this TMEM buffer
        mbar_wait(&sfull_bar[stage], phase);
        tc_fence_after();
        const uint32_t slab = smem_u32(sSlab + stage * p.slab_bytes);
#pragma unroll 1
        for (int a = 0; a < 2; ++a) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128 + a * CF_COUT);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int rowoff = a * 128 + (tap / 3) * p.Wp + (tap % 3);
            const uint32_t a_addr = slab + (uint32_t)rowoff * CF_ROW_BYTES;
            const uint32_t b_addr = w_addr + tap * CF_W_TAP_BYTES;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, cf_desc_a(a_addr + k * 32, p.desc_mode), make_smem_desc<CF_ROW_BYTES>(b_addr + k * 32), idesc,
                        (tap > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&sempty_bar[stage]);
        umma_commit(&tfull_bar[buf]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        buf ^= 1;
        if (buf == 0) bphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: warps 2..9, four per accumulator; TMEM lane = flat position =====================
    const int quarter = warp & 3;
    const int a = (warp - 2) >> 2;                             // accumulator = 128-position half of the tile
    const int row = quarter * 32 + lane;
    const int ep_tid = row;                                    // 0..127 within the group
    // Output path: one [128 x 64] bf16 slab (128-byte rows, 128B swizzle) per tile and group, handed to ONE TMA store;
    // two slabs per group alternate.  (The generic 32-column slab ring of tc_common.cuh costs two named barriers, a
    // proxy fence and a TMA store per 32 columns: ~1.2k cycles per step, 2.4k per tile - twice this tile's MMA time.)
    uint8_t* slabs = sEpi + a * (2 * CF_OUT_SLAB);
    const int bar_a = 2 + 2 * a, bar_b = 3 + 2 * a;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * CF_COUT);
    const int img = p.Hp * p.Wp;
    int buf = 0;
    uint32_t bphase = 0;
    int turn = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int m0 = t * CF_TM + a * 128;
      const int pos = m0 + row;
      const int pi = pos % img;
      const int yp = pi / p.Wp, xp = pi - yp * p.Wp;
      const bool interior = pos < p.NP && yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W;
      mbar_wait(&tfull_bar[buf], bphase);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_addr + (uint32_t)(buf * 128), v0);
      tmem_ld32(lane_addr + (uint32_t)(buf * 128 + 32), v1);
      tmem_ld_wait();
      tc_fence_before();                                       // this warp has read its part of the buffer
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      uint32_t packed[32];
      if (interior) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float r0 = fmaf(__uint_as_float(v0[2 * j]), bn.scale[2 * j], bn.shift[2 * j]);
          float r1 = fmaf(__uint_as_float(v0[2 * j + 1]), bn.scale[2 * j + 1], bn.shift[2 * j + 1]);
          float r2 = fmaf(__uint_as_float(v1[2 * j]), bn.scale[32 + 2 * j], bn.shift[32 + 2 * j]);
          float r3 = fmaf(__uint_as_float(v1[2 * j + 1]), bn.scale[32 + 2 * j + 1], bn.shift[32 + 2 * j + 1]);
          if (p.relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); r2 = fmaxf(r2, 0.f); r3 = fmaxf(r3, 0.f); }
          packed[j] = pack_bf16x2(r0, r1);
          packed[16 + j] = pack_bf16x2(r2, r3);
        }
      } else {
        // the ring is written as zeros: it is the next layer's padding
#pragma unroll
        for (int j = 0; j < 32; ++j) packed[j] = 0u;
      }
      uint8_t* slab = slabs + turn * CF_OUT_SLAB;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_a) : "memory");   // the store that last read this slab has drained it
      uint8_t* rp = slab + row * 128;                              // 128-byte rows, 128B swizzle: chunk ^= row & 7
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(rp + ((c ^ (row & 7)) << 4)) =
            make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_b) : "memory");   // slab complete
      if (ep_tid == 0) {
        tma_store_2d(&tmO, slab, 0, m0);                           // rows beyond NP are clipped by the tensor map
        bulk_commit();
        bulk_wait_read<1>();                                       // the slab written NEXT has been read
      }
      turn ^= 1;
      buf ^= 1;
      if (buf == 0) bphase ^= 1;
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

// ======================================================================================================
// The same convolution with the 2x2 / stride 2 max-pool that follows it (maxpool1, feature_extraction.py:66,116) in
// the epilogue: the pre-pool activation (1.7 GB at 4096 crops) never reaches HBM and the pool launch disappears.
//
// A tile is one PAIR of image rows (top row y = 2j+1 of the padded image, bottom row y+1) by 128 columns starting at
// padded column 1 + 128 c:  accumulator 0 holds the 128 positions of the top row, accumulator 1 the 128 positions one
// padded row (Wp positions) further - so TMEM lane i of both accumulators is the same image column and the vertical
// maximum is thread-local; lanes 2q and 2q+1 are the two columns of pooled pixel q: one shuffle.  The slab is the rows
// [p_top - Wp - 1, p_top + 128 + 2 Wp + 1) of the flat input; tap (r, s) of accumulator a reads it at row offset
// (a + r) Wp + s.  With W = 100 only 100 of the 128 lanes are image columns (the rest read the ring and the next rows:
// finite, unused), i.e. 28 % more MMA work than the un-pooled tiling - paid for by the stores it saves.
// Epilogue: SIXTEEN warps = TMEM lane quarter (warp % 4) x channel quarter; every warp reads ITS 16 channels of both
// accumulators (with eight warps x 32 channels the epilogue, ~2 us per tile, was twice the tile's MMA time).
// Arithmetic per output element is the un-pooled kernel's (same taps, same order, same BN / ReLU / rounding) followed
// by an exact maximum: results are bit-identical to conv3x3_flat_kernel + maxpool2x2_ring (tests/test_gpu_kernels.py).
// ======================================================================================================
constexpr int CFP_OUT_SLAB = 64 * CF_COUT * 2;   // [64 pooled pixels][64 channels] bf16 = 8 KB, 128-byte rows, 128B swizzle
constexpr int CFP_EPI_WARPS = 16;
constexpr int CFP_THREADS = 64 + 32 * CFP_EPI_WARPS;

struct CfPoolParams {
  int B, H, W, Hp, Wp;
  int chunks;                                    // 128-column chunks per row pair
  int num_tiles;                                 // B * H/2 * chunks
  int box_rows, nload, slab_bytes, stages;
  int relu;
};

// 16 channels (channel quarter CQ) of one column: BN + ReLU + bf16 rounding of the top and the bottom row's conv output,
// vertical maximum, horizontal maximum with the neighbouring lane.  CQ is a template parameter so that every scale /
// shift stays a constant-bank operand.
template <int CQ>
__device__ __forceinline__ void cfp_pool_quarter(const CfAffine& bn, const uint32_t (&top)[16], const uint32_t (&bot)[16],
                                                 const int relu, uint32_t (&packed)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t0 = fmaf(__uint_as_float(top[2 * j]), bn.scale[CQ * 16 + 2 * j], bn.shift[CQ * 16 + 2 * j]);
    float t1 = fmaf(__uint_as_float(top[2 * j + 1]), bn.scale[CQ * 16 + 2 * j + 1], bn.shift[CQ * 16 + 2 * j + 1]);
    float b0 = fmaf(__uint_as_float(bot[2 * j]), bn.scale[CQ * 16 + 2 * j], bn.shift[CQ * 16 + 2 * j]);
    float b1 = fmaf(__uint_as_float(bot[2 * j + 1]), bn.scale[CQ * 16 + 2 * j + 1], bn.shift[CQ * 16 + 2 * j + 1]);
    if (relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); b0 = fmaxf(b0, 0.f); b1 = fmaxf(b1, 0.f); }
    // round each conv output to bf16 first (what the un-pooled kernel stores), then the maximum: exact on bf16 values
    const uint32_t tp = pack_bf16x2(t0, t1), bp = pack_bf16x2(b0, b1);
    const float2 tf = unpack_bf16x2(tp), bf = unpack_bf16x2(bp);
    float m0 = fmaxf(tf.x, bf.x), m1 = fmaxf(tf.y, bf.y);
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    packed[j] = pack_bf16x2(m0, m1);
  }
}

__global__ void __launch_bounds__(CFP_THREADS, 1)
conv3x3_flat_pool_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CfAffine bn, const CfPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                          // [9][64 rows][64 B], 64B swizzle
  uint8_t* sSlab = sW + CF_W_BYTES;                            // [stages][slab_bytes]
  uint8_t* sEpi = sSlab + p.stages * p.slab_bytes;             // two output slabs
  uint64_t* wfull_bar = reinterpret_cast<uint64_t*>(sEpi + 2 * CFP_OUT_SLAB);
  uint64_t* sfull_bar = wfull_bar + 1;
  uint64_t* sempty_bar = sfull_bar + CF_MAX_STAGES;
  uint64_t* tfull_bar = sempty_bar + CF_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmO);
    mbar_init(wfull_bar, 1);
    for (int i = 0; i < CF_MAX_STAGES; ++i) { mbar_init(&sfull_bar[i], 1); mbar_init(&sempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], CFP_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int pairs = p.H / 2;
  // flat position of the first column (padded column 1 + 128 c) of the top row of tile t
  auto tile_top = [&](const int t, int& b, int& j, int& c) {
    c = t % p.chunks;
    const int bj = t / p.chunks;
    j = bj % pairs;
    b = bj / pairs;
    return (b * p.Hp + 2 * j + 1) * p.Wp + 1 + 128 * c;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull_bar, CF_W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + tap * CF_W_TAP_BYTES, &tmW, wfull_bar, tap * CF_CIN, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int b, j, c;
        const int row0 = tile_top(t, b, j, c) - p.Wp - 1;      // >= 0; rows beyond the tensor are zero-filled by TMA
        mbar_wait(&sempty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&sfull_bar[stage], (uint32_t)(p.nload * p.box_rows * CF_ROW_BYTES));
        uint8_t* dst = sSlab + stage * p.slab_bytes;
        for (int i = 0; i < p.nload; ++i)
          tma_load_2d(dst + i * p.box_rows * CF_ROW_BYTES, &tmX, &sfull_bar[stage], 0, row0 + i * p.box_rows);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, CF_COUT);
      const uint32_t w_addr = smem_u32(sW);
      mbar_wait(wfull_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[buf], bphase ^ 1);
        mbar_wait(&sfull_bar[stage], phase);
        tc_fence_after();
        const uint32_t slab = smem_u32(sSlab + stage * p.slab_bytes);
#pragma unroll 1
        for (int a = 0; a < 2; ++a) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128 + a * CF_COUT);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int rowoff = (a + tap / 3) * p.Wp + (tap % 3);
            const uint32_t a_addr = slab + (uint32_t)rowoff * CF_ROW_BYTES;
            const uint32_t b_addr = w_addr + tap * CF_W_TAP_BYTES;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, make_smem_desc<CF_ROW_BYTES>(a_addr + k * 32), make_smem_desc<CF_ROW_BYTES>(b_addr + k * 32), idesc,
                        (tap > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&sempty_bar[stage]);
        umma_commit(&tfull_bar[buf]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        buf ^= 1;
        if (buf == 0) bphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: warps 2..17 = TMEM lane quarter x channel quarter =====================
    const int quarter = warp & 3;
    const int cq = (warp - 2) >> 2;                            // 16 channels
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cq * 16);
    const int ep_tid = ((warp - 2) << 5) + lane;               // 0..511
    const int out_w = p.W / 2;
    const int prow = quarter * 16 + (lane >> 1);               // pooled pixel of the tile (even lanes store it)
    int buf = 0;
    uint32_t bphase = 0;
    int turn = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int b, j, c;
      tile_top(t, b, j, c);
      mbar_wait(&tfull_bar[buf], bphase);
      tc_fence_after();
      uint32_t top[16], bot[16];
      tmem_ld16(lane_addr + (uint32_t)(buf * 128), top);
      tmem_ld16(lane_addr + (uint32_t)(buf * 128 + CF_COUT), bot);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      uint32_t packed[8];
      switch (cq) {
        case 0: cfp_pool_quarter<0>(bn, top, bot, p.relu, packed); break;
        case 1: cfp_pool_quarter<1>(bn, top, bot, p.relu, packed); break;
        case 2: cfp_pool_quarter<2>(bn, top, bot, p.relu, packed); break;
        default: cfp_pool_quarter<3>(bn, top, bot, p.relu, packed); break;
      }
      uint8_t* slab = sEpi + turn * CFP_OUT_SLAB;
      asm volatile("bar.sync 2, 512;" ::: "memory");             // the store that last read this slab has drained it
      if ((lane & 1) == 0) {
        uint8_t* rp = slab + prow * 128;                         // 128-byte rows, 128B swizzle: chunk ^= row & 7
#pragma unroll
        for (int q = 0; q < 2; ++q)
          *reinterpret_cast<uint4*>(rp + (((cq * 2 + q) ^ (prow & 7)) << 4)) =
              make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 3, 512;" ::: "memory");             // slab complete
      if (ep_tid == 0) {
        // pooled row (b, j), pooled columns 64 c ..: the store box holds min(64, W/2) rows, so a chunk never runs
        // into the next pooled row
        tma_store_2d(&tmO, slab, 0, (b * pairs + j) * out_w + 64 * c);
        bulk_commit();
        bulk_wait_read<1>();
      }
      turn ^= 1;
      buf ^= 1;
      if (buf == 0) bphase ^= 1;
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

static TmaEncoders g_flat_enc;
static std::atomic<int> g_flat_desc_mode{0};

}  // namespace mstr

// Experiment switch for the A-descriptor base offset (tests only): 0 = leave zero (default, correct), 1 = set it.
extern "C" __attribute__((visibility("default"))) int mstr_debug_flat_desc_mode(int mode) {
  mstr::g_flat_desc_mode.store(mode ? 1 : 0);
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_conv3x3_flat_bf16(
    const void* x, const void* w, const float* scale_host, const float* shift_host, void* out, int B, int H, int W,
    int C_in, int C_out, int relu, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && w && out, MSTR_ERR_BAD_ARG, "mstr_conv3x3_flat_bf16: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0, MSTR_ERR_BAD_ARG, "mstr_conv3x3_flat_bf16: bad size");
  MSTR_REQUIRE(C_in == CF_CIN && C_out == CF_COUT, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_flat_bf16: only C_in = %d, C_out = %d (got %d, %d)", CF_CIN, CF_COUT, C_in, C_out);
  MSTR_REQUIRE(aligned16(x) && aligned16(w) && aligned16(out), MSTR_ERR_BAD_ARG,
               "mstr_conv3x3_flat_bf16: pointers must be 16-byte aligned");
  const int64_t NP64 = (int64_t)B * (H + 2) * (W + 2);
  MSTR_REQUIRE(NP64 + CF_TM < (1ll << 31), MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_flat_bf16: too many positions");
  int rc = load_tma_encoders(g_flat_enc);
  if (rc != MSTR_OK) return rc;
  CfParams p;
  p.NP = (int)NP64; p.H = H; p.W = W; p.Hp = H + 2; p.Wp = W + 2;
  p.num_tiles = ceil_div(p.NP, CF_TM);
  const int slab_rows = CF_TM + 2 * p.Wp + 2;
  p.nload = ceil_div(slab_rows, 256);
  p.box_rows = (ceil_div(slab_rows, p.nload) + 7) & ~7;
  MSTR_REQUIRE(p.box_rows <= 256, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_flat_bf16: W=%d too wide", W);
  p.slab_bytes = (p.nload * p.box_rows * CF_ROW_BYTES + 1023) & ~1023;
  const int budget = 227 * 1024 - 1024 - CF_W_BYTES - CF_EPI_BYTES - CF_TAIL_BYTES;
  p.stages = budget / p.slab_bytes;
  if (p.stages > CF_MAX_STAGES) p.stages = CF_MAX_STAGES;
  MSTR_REQUIRE(p.stages >= 2, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_flat_bf16: W=%d needs a slab of %d bytes", W, p.slab_bytes);
  p.relu = relu;
  p.desc_mode = g_flat_desc_mode.load(std::memory_order_relaxed);
  alignas(64) CUtensorMap tmX, tmW, tmO;
  rc = encode_tiled_2d_bf16(g_flat_enc, &tmX, x, CF_CIN, (uint64_t)p.NP, CF_CIN, (uint32_t)p.box_rows, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != MSTR_OK) return rc;
  rc = encode_tiled_2d_bf16(g_flat_enc, &tmW, w, 9 * CF_CIN, CF_COUT, CF_CIN, CF_COUT, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != MSTR_OK) return rc;
  rc = encode_tiled_2d_bf16(g_flat_enc, &tmO, out, CF_COUT, (uint64_t)p.NP, CF_COUT, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  CfAffine bn;
  for (int i = 0; i < CF_COUT; ++i) {
    bn.scale[i] = scale_host ? scale_host[i] : 1.f;
    bn.shift[i] = shift_host ? shift_host[i] : 0.f;
  }
  const int smem_bytes = 1024 + CF_W_BYTES + p.stages * p.slab_bytes + CF_EPI_BYTES + CF_TAIL_BYTES;
  static unsigned long long smem_set = 0;
  rc = ensure_dynamic_smem((const void*)conv3x3_flat_kernel, 227 * 1024, &smem_set);
  if (rc != MSTR_OK) return rc;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  conv3x3_flat_kernel<<<grid, CF_THREADS, smem_bytes, (cudaStream_t)stream>>>(tmX, tmW, tmO, bn, p);
  MSTR_LAUNCH_CHECK("conv3x3_flat_kernel");
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_conv3x3_flat_pool_bf16(
    const void* x, const void* w, const float* scale_host, const float* shift_host, void* out, int B, int H, int W,
    int C_in, int C_out, int relu, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && w && out, MSTR_ERR_BAD_ARG, "mstr_conv3x3_flat_pool_bf16: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0, MSTR_ERR_BAD_ARG, "mstr_conv3x3_flat_pool_bf16: bad size");
  MSTR_REQUIRE(C_in == CF_CIN && C_out == CF_COUT, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_flat_pool_bf16: only C_in = %d, C_out = %d (got %d, %d)", CF_CIN, CF_COUT, C_in, C_out);
  MSTR_REQUIRE(H % 2 == 0 && W % 2 == 0 && (W <= 128 || W % 128 == 0), MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_flat_pool_bf16: H=%d, W=%d (both even; W <= 128 or a multiple of 128)", H, W);
  MSTR_REQUIRE(aligned16(x) && aligned16(w) && aligned16(out), MSTR_ERR_BAD_ARG,
               "mstr_conv3x3_flat_pool_bf16: pointers must be 16-byte aligned");
  const int64_t NP64 = (int64_t)B * (H + 2) * (W + 2);
  MSTR_REQUIRE(NP64 + 1024 < (1ll << 31), MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_flat_pool_bf16: too many positions");
  int rc = load_tma_encoders(g_flat_enc);
  if (rc != MSTR_OK) return rc;
  CfPoolParams p;
  p.B = B; p.H = H; p.W = W; p.Hp = H + 2; p.Wp = W + 2;
  p.chunks = ceil_div(W, 128);
  p.num_tiles = B * (H / 2) * p.chunks;
  const int slab_rows = 128 + 3 * p.Wp + 2;
  p.nload = ceil_div(slab_rows, 256);
  p.box_rows = (ceil_div(slab_rows, p.nload) + 7) & ~7;
  MSTR_REQUIRE(p.box_rows <= 256, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_flat_pool_bf16: W=%d too wide", W);
  p.slab_bytes = (p.nload * p.box_rows * CF_ROW_BYTES + 1023) & ~1023;
  const int budget = 227 * 1024 - 1024 - CF_W_BYTES - 2 * CFP_OUT_SLAB - CF_TAIL_BYTES;
  p.stages = budget / p.slab_bytes;
  if (p.stages > CF_MAX_STAGES) p.stages = CF_MAX_STAGES;
  MSTR_REQUIRE(p.stages >= 2, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_flat_pool_bf16: W=%d needs a slab of %d bytes", W, p.slab_bytes);
  p.relu = relu;
  alignas(64) CUtensorMap tmX, tmW, tmO;
  rc = encode_tiled_2d_bf16(g_flat_enc, &tmX, x, CF_CIN, (uint64_t)NP64, CF_CIN, (uint32_t)p.box_rows, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != MSTR_OK) return rc;
  rc = encode_tiled_2d_bf16(g_flat_enc, &tmW, w, 9 * CF_CIN, CF_COUT, CF_CIN, CF_COUT, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != MSTR_OK) return rc;
  const int out_w = W / 2;
  rc = encode_tiled_2d_bf16(g_flat_enc, &tmO, out, CF_COUT, (uint64_t)B * (H / 2) * out_w, CF_COUT,
                            (uint32_t)(out_w < 64 ? out_w : 64), CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  CfAffine bn;
  for (int i = 0; i < CF_COUT; ++i) {
    bn.scale[i] = scale_host ? scale_host[i] : 1.f;
    bn.shift[i] = shift_host ? shift_host[i] : 0.f;
  }
  const int smem_bytes = 1024 + CF_W_BYTES + p.stages * p.slab_bytes + 2 * CFP_OUT_SLAB + CF_TAIL_BYTES;
  static unsigned long long smem_set = 0;
  rc = ensure_dynamic_smem((const void*)conv3x3_flat_pool_kernel, 227 * 1024, &smem_set);
  if (rc != MSTR_OK) return rc;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  conv3x3_flat_pool_kernel<<<grid, CFP_THREADS, smem_bytes, (cudaStream_t)stream>>>(tmX, tmW, tmO, bn, p);
  MSTR_LAUNCH_CHECK("conv3x3_flat_pool_kernel");
  return MSTR_OK;
}
