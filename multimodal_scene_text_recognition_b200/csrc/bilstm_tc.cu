// Persistent BiLSTM recurrence on the 5th-generation tensor cores (bf16 throughput mode).
//
// Replaces the recurrent half of nn.LSTM(bidirectional=True, batch_first=True)
// (modules/encoders.py:14,24; gate order i,f,g,o, zero initial state, the reverse direction runs
// t = T-1..0 with the *_reverse parameters).  The input projection x W_ih^T + b_ih + b_hh is a
// separate tcgen05 GEMM (mstr_conv_gemm_bf16); its fp32 result `gates_x` is added in the epilogue.
//
// One CTA owns a tile of 128 crops of one direction for ALL T steps (grid = ceil(B/128) x 2):
//   h_t   [128 crops x 256] bf16 lives in shared memory in the UMMA K-major 128B-swizzled layout
//         (double buffered), written by the epilogue threads and read by tcgen05.mma as operand A;
//   c_t   [128 x 256] fp32 lives in TMEM columns 0..255 for the whole sequence;
//   W_hh  (bf16, 512 KB per direction, L2 resident) is streamed by TMA through a 4-stage ring in
//         tiles of [128 gate rows x 64 k]; rows are pre-packed so that one 128-column accumulator
//         chunk holds the i,f,g,o gates of the same 32 hidden units;
//   gates [128 x 128] fp32 accumulators are double buffered in TMEM columns 256..511, so the MMAs
//         of chunk c+1 overlap the gate math of chunk c.
// Per step: 8 chunks x 4 k-blocks x 4 tcgen05.mma (128x128x16).  Warp roles: warp 0 = TMA producer,
// warp 1 = MMA issuer, warps 2..17 = epilogue (four warps per TMEM lane quarter, 8 units each: 16 warps
// hide the latency of the gates_x loads and TMEM round trips better than 8 did, 0.65 -> see profiles/).
// The gate non-linearities use tanh.approx (1 MUFU each; sigmoid(x) = 0.5 tanh(x/2) + 0.5), which is
// what bounds the kernel: 5 MUFU per hidden unit per crop per step at 16 MUFU/clk/SM.
#include "tc_common.cuh"

namespace mstr {

constexpr int LT_H = 256;          // hidden size (model.py:30)
constexpr int LT_M = 128;          // crops per CTA
constexpr int LT_CHUNK_UNITS = 32; // hidden units per accumulator chunk (x 4 gates = 128 columns)
constexpr int LT_CHUNKS = LT_H / LT_CHUNK_UNITS;   // 8
constexpr int LT_KB = LT_H / 64;   // 4 k-blocks of 64 bf16 (one 128-byte swizzle span)
constexpr int LT_STAGES = 4;
constexpr int LT_W_TILE_BYTES = 128 * 128;         // [128 gate rows][64 k] bf16
constexpr int LT_H_KB_BYTES = LT_M * 128;          // one k-block of h: 128 rows x 128 bytes
constexpr int LT_H_BUF_BYTES = LT_KB * LT_H_KB_BYTES;  // 64 KB
constexpr int LT_THREADS = 576;
constexpr int LT_EPI_WARPS = 16;
constexpr int LT_SMEM_BYTES = 1024 + 2 * LT_H_BUF_BYTES + LT_STAGES * LT_W_TILE_BYTES + 256;
constexpr uint32_t LT_ACC_COL0 = 256;  // TMEM: [0,256) cell state, [256,384) / [384,512) accumulators

struct LstmTcParams {
  const float* gates_x;         // [B][T][2][4H]
  __nv_bfloat16* out;           // [B][T][2H]
  int B, T;
};

__global__ void __launch_bounds__(LT_THREADS, 1)
bilstm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sH = smem;                                   // [2][LT_KB][128 rows][128 B]
  uint8_t* sW = smem + 2 * LT_H_BUF_BYTES;              // [STAGES][128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sW + LT_STAGES * LT_W_TILE_BYTES);
  uint64_t* empty_bar = full_bar + LT_STAGES;
  uint64_t* tfull_bar = empty_bar + LT_STAGES;          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                 // [2]
  uint64_t* hready_bar = tempty_bar + 2;                // [LT_KB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hready_bar + LT_KB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.y;
  const int crop0 = blockIdx.x * LT_M;
  const int T = p.T;

  // h_0 = 0 in buffer 0
  for (int i = threadIdx.x; i < LT_H_BUF_BYTES / 16; i += LT_THREADS)
    reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmW);
    for (int i = 0; i < LT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], LT_EPI_WARPS); }
    for (int i = 0; i < LT_KB; ++i) mbar_init(&hready_bar[i], 2 * LT_EPI_WARPS);  // 2 chunks x 16 warps per k-block
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: W_hh tiles, (step, chunk, kb) order =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < T; ++step)
        for (int chunk = 0; chunk < LT_CHUNKS; ++chunk)
          for (int kb = 0; kb < LT_KB; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], LT_W_TILE_BYTES);
            tma_load_2d(sW + stage * LT_W_TILE_BYTES, &tmW, &full_bar[stage], kb * 64,
                        dir * 4 * LT_H + chunk * 128);
            if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(LT_M, 128);
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;  // running chunk counter
      for (int step = 0; step < T; ++step) {
        const uint32_t h_addr = smem_u32(sH + (step & 1) * LT_H_BUF_BYTES);
        for (int chunk = 0; chunk < LT_CHUNKS; ++chunk, ++n) {
          const int buf = n & 1;
          mbar_wait(&tempty_bar[buf], ((n >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + LT_ACC_COL0 + (uint32_t)(buf * 128);
          for (int kb = 0; kb < LT_KB; ++kb) {
            // h_t k-block kb is complete once both chunks of step t-1 that produce it have been written
            if (step > 0 && chunk == 0) mbar_wait(&hready_bar[kb], (step - 1) & 1);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = h_addr + kb * LT_H_KB_BYTES;
            const uint32_t b_addr = smem_u32(sW + stage * LT_W_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, make_smem_desc<128>(a_addr + k * 32), make_smem_desc<128>(b_addr + k * 32), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == LT_KB - 1) umma_commit(&tfull_bar[buf]);
            if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue: gate math, c in TMEM, h -> smem (next A operand) + global ===========
    const int ew = warp - 2;                 // 0..15
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int uq = ew >> 2;                  // which 8 of the chunk's 32 units
    const int row = quarter * 32 + lane;     // crop row within the tile
    const int crop = crop0 + row;
    const bool valid = crop < p.B;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int n = 0;
    for (int step = 0; step < T; ++step) {
      const int t = dir ? (T - 1 - step) : step;
      uint8_t* h_next = sH + ((step + 1) & 1) * LT_H_BUF_BYTES;
      const float* gx_row = p.gates_x + (((size_t)(valid ? crop : 0) * T + t) * 2 + dir) * (4 * LT_H);
      __nv_bfloat16* out_row = p.out + ((size_t)(valid ? crop : 0) * T + t) * (2 * LT_H) + dir * LT_H;
      for (int chunk = 0; chunk < LT_CHUNKS; ++chunk, ++n) {
        const int buf = n & 1;
        const int unit0 = chunk * LT_CHUNK_UNITS + uq * 8;       // first of this thread's 8 hidden units
        // input-projection gates for these units: issued before waiting for the MMA
        float4 gx[4][2];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int v = 0; v < 2; ++v)
            gx[g][v] = valid ? __ldg(reinterpret_cast<const float4*>(gx_row + g * LT_H + unit0) + v)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(&tfull_bar[buf], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t acc_addr = lane_addr + LT_ACC_COL0 + (uint32_t)(buf * 128 + uq * 8);
        uint32_t a[4][8], cst[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld8(acc_addr + (uint32_t)(g * 32), a[g]);
        if (step > 0) tmem_ld8(lane_addr + (uint32_t)unit0, cst);
        tmem_ld_wait();
        // accumulator values are in registers: hand the buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        float hv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int v = j >> 2;
          const float gi = reinterpret_cast<const float*>(&gx[0][v])[j & 3] + __uint_as_float(a[0][j]);
          const float gf = reinterpret_cast<const float*>(&gx[1][v])[j & 3] + __uint_as_float(a[1][j]);
          const float gg = reinterpret_cast<const float*>(&gx[2][v])[j & 3] + __uint_as_float(a[2][j]);
          const float go = reinterpret_cast<const float*>(&gx[3][v])[j & 3] + __uint_as_float(a[3][j]);
          const float c_prev = step > 0 ? __uint_as_float(cst[j]) : 0.f;
          const float c_new = fmaf(sigmoid_approx(gf), c_prev, sigmoid_approx(gi) * tanh_approx(gg));
          cst[j] = __float_as_uint(c_new);
          hv[j] = sigmoid_approx(go) * tanh_approx(c_new);
        }
        tmem_st8(lane_addr + (uint32_t)unit0, cst);
        const uint4 hpk = make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]),
                                     pack_bf16x2(hv[6], hv[7]));
        // h_{t+1}: one 16-byte chunk of row `row` in k-block unit0/64 (128B swizzle: chunk ^= row & 7)
        const int kb = unit0 >> 6;
        const int c0 = (unit0 & 63) >> 3;
        *reinterpret_cast<uint4*>(h_next + kb * LT_H_KB_BYTES + row * 128 + ((c0 ^ (row & 7)) << 4)) = hpk;
        if (valid) *reinterpret_cast<uint4*>(out_row + unit0) = hpk;
        tmem_st_wait();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hready_bar[kb]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static TmaEncoders g_lstm_enc;

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_bilstm_recurrent_bf16(
    const float* gates_x, const void* w_hh_packed, void* out_bf16, int B, int T, int Hd, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(gates_x && w_hh_packed && out_bf16, MSTR_ERR_BAD_ARG, "mstr_bilstm_recurrent_bf16: null pointer");
  MSTR_REQUIRE(B > 0 && T > 0, MSTR_ERR_BAD_ARG, "mstr_bilstm_recurrent_bf16: bad size");
  MSTR_REQUIRE(Hd == LT_H, MSTR_ERR_UNSUPPORTED, "mstr_bilstm_recurrent_bf16: hidden size %d != %d", Hd, LT_H);
  MSTR_REQUIRE(aligned16(gates_x) && aligned16(w_hh_packed) && aligned16(out_bf16), MSTR_ERR_BAD_ARG,
               "mstr_bilstm_recurrent_bf16: pointers must be 16-byte aligned");
  int rc = load_tma_encoders(g_lstm_enc);
  if (rc != MSTR_OK) return rc;
  alignas(64) CUtensorMap tmW;
  rc = encode_tiled_2d_bf16(g_lstm_enc, &tmW, w_hh_packed, LT_H, 2 * 4 * LT_H, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MSTR_CUDA(cudaFuncSetAttribute(bilstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM_BYTES));
    attr_set = true;
  }
  LstmTcParams p{gates_x, (__nv_bfloat16*)out_bf16, B, T};
  dim3 grid(ceil_div(B, LT_M), 2);
  bilstm_tc_kernel<<<grid, LT_THREADS, LT_SMEM_BYTES, (cudaStream_t)stream>>>(tmW, p);
  MSTR_LAUNCH_CHECK("bilstm_tc_kernel");
  return MSTR_OK;
}
