// Cluster-cooperative BiLSTM recurrence on the 5th-generation tensor cores (bf16 throughput mode): W_hh STATIONARY in
// shared memory, the hidden state exchanged between the CTAs of a cluster through distributed shared memory.
//
// Same contract as bilstm_tc.cu (mstr_bilstm_recurrent_bf16; nn.LSTM(bidirectional=True, batch_first=True),
// modules/encoders.py:14,24): gates_x [B][T][2][4H] f32 (input projection + both biases), w_hh_packed
// [2][4H][H] bf16 (rows ordered chunk*128 + gate*32 + unit), out [B][T][2H] bf16.
//
// A cluster of 8 CTAs owns one tile of 128 crops of one direction for all T steps; CTA `rank` owns the gate
// chunk `rank` = hidden units [rank*32, rank*32+32) x gates i,f,g,o (128 gate rows):
//   W slice  [128 gate rows x 256 k] bf16 = 64 KB, loaded ONCE by TMA (bilstm_tc.cu streams all 512 KB of W_hh of
//            a direction from L2 every step), as 8 k-blocks of 32 (64-byte rows, 64B swizzle);
//   h_t      [128 crops x 256] bf16, the A operand, double buffered in EVERY CTA of the cluster, in the same k-blocks
//            of 32 units: k-block j IS the 8 KB slice CTA j produces.  The epilogue threads write their 16 bytes
//            into the local copy of the slice, fence (fence.proxy.async), meet at a CTA-local barrier, and seven
//            threads then push the slice to the seven peers with cp.async.bulk.shared::cluster, each completing
//            8 KB of transaction count on the PEER's `hfull` mbarrier; the CTA's own arrive.expect_tx(7 x 8 KB)
//            covers the local slice.  The MMA thread waits for `hfull` (acquire.cluster) and issues the step's 16
//            tcgen05.mma.  (A first version wrote 16 bytes per thread into all 8 CTAs with st.shared::cluster -
//            4096 remote stores per CTA and step plus a cluster-scope fence: 6 us per step.)
//   gates    three [128 x 128] f32 accumulator buffers in TMEM (step % 3).  gates_x does not depend on h, so the
//            loader warps pre-fill the accumulators up to two steps AHEAD (TMA tile -> shared memory -> registers ->
//            tcgen05.st); the MMAs accumulate on top;
//   c_t      lives in the registers of the epilogue threads (a thread always owns the same 8 units of one crop).
// Per step and CTA: 16 MMAs (128x128x16) + one chunk of gate math, instead of 8 chunks walked serially by one CTA:
// the serial path per step drops from ~17.7 us to a few us, which is what the latency-bound recurrence needs -
// above all at small batches, where bilstm_tc.cu uses 4 CTAs for a batch of 256.
//
// Warp roles: warp 0 = TMA producer (W slice once, then the gates_x tiles), warp 1 = MMA issuer, warps 2..17 =
// epilogue (four per TMEM lane quarter, 8 units each), warps 18..21 = gates_x loaders (one per lane quarter).
#include <atomic>

#include "tc_common.cuh"

namespace mstr {

constexpr int LC_H = 256;
constexpr int LC_M = 128;                      // crops per cluster
constexpr int LC_CLUSTER = 8;                  // CTAs per cluster = gate chunks of 32 hidden units
constexpr int LC_KB = LC_H / 32;               // 8 k-blocks of 32 bf16: k-block j = the hidden units CTA j produces
constexpr int LC_KTILE_BYTES = 128 * 64;       // [128 rows][64 B], 64B swizzle: one k-block of W or of h (8 KB)
constexpr int LC_TILE_BYTES = 128 * 128;       // [128 rows][128 B] f32, 128B swizzle: one gates_x pass (16 KB)
constexpr int LC_H_BUF_BYTES = LC_KB * LC_KTILE_BYTES;   // 64 KB
constexpr int LC_GX_STAGES = 2;
constexpr int LC_ACC = 3;                      // accumulator buffers in TMEM
constexpr int LC_EPI_WARPS = 16;
constexpr int LC_LOAD_WARPS = 4;
constexpr int LC_THREADS = 32 * (2 + LC_EPI_WARPS + LC_LOAD_WARPS);   // 704
constexpr int LC_SMEM_BYTES = 1024 + LC_KB * LC_KTILE_BYTES + 2 * LC_H_BUF_BYTES + LC_GX_STAGES * LC_TILE_BYTES + 256;

struct LstmClusterParams {
  const void* gates_x;          // [B][T][2][4H] f32, or bf16 when gx_bf16
  __nv_bfloat16* out;           // [B][T][2H]
  int B, T;
  int gx_bf16;
};

__device__ __forceinline__ uint32_t lc_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t lc_cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void lc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t lc_map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// wait with cluster-scope acquire: the arrivals come from the other CTAs of the cluster
__device__ __forceinline__ void lc_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();     // a protocol bug must surface as a trapped launch, not a hang
  }
}
// bulk copy of `bytes` from this CTA's shared memory into a peer CTA's shared memory (async proxy), completing
// `bytes` of transaction count on the PEER's mbarrier
__device__ __forceinline__ void lc_bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                     uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}

__global__ void __launch_bounds__(LC_THREADS, 1)
bilstm_cluster_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmGx,
                      const LstmClusterParams p) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem base has the same offset in every CTA of the cluster, so this rounding is identical in all of them
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                    // [LC_KB][128 gate rows][64 B]
  uint8_t* sH = sW + LC_KB * LC_KTILE_BYTES;             // [2][LC_KB][128 crops][64 B]
  uint8_t* sG = sH + 2 * LC_H_BUF_BYTES;                 // [LC_GX_STAGES][128 crops][128 B] f32
  uint64_t* wfull_bar = reinterpret_cast<uint64_t*>(sG + LC_GX_STAGES * LC_TILE_BYTES);
  uint64_t* gfull_bar = wfull_bar + 1;                   // [LC_GX_STAGES]
  uint64_t* gempty_bar = gfull_bar + LC_GX_STAGES;       // [LC_GX_STAGES]
  uint64_t* tinit_bar = gempty_bar + LC_GX_STAGES;       // [LC_ACC] accumulator holds its gates_x tile
  uint64_t* tfull_bar = tinit_bar + LC_ACC;              // [LC_ACC] MMAs of the step have completed
  uint64_t* tempty_bar = tfull_bar + LC_ACC;             // [LC_ACC] epilogue has drained the accumulator
  uint64_t* hfull_bar = tempty_bar + LC_ACC;             // [2] all 8 slices of h in buffer b have arrived
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hfull_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = lc_cluster_ctarank();            // gate chunk
  const int tile = (int)lc_cluster_id_x();               // (crop tile, direction)
  const int dir = tile & 1;
  const int crop0 = (tile >> 1) * LC_M;
  const int T = p.T;

  // h_0 = 0 in buffer 0 (local copy of every CTA)
  for (int i = threadIdx.x; i < LC_H_BUF_BYTES / 16; i += LC_THREADS)
    reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmGx);
    mbar_init(wfull_bar, 1);
    for (int i = 0; i < LC_GX_STAGES; ++i) { mbar_init(&gfull_bar[i], 1); mbar_init(&gempty_bar[i], LC_LOAD_WARPS); }
    for (int i = 0; i < LC_ACC; ++i) {
      mbar_init(&tinit_bar[i], LC_LOAD_WARPS);
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], LC_EPI_WARPS);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&hfull_bar[i], 1);   // own arrive.expect_tx; the 7 peer slices complete_tx
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  lc_cluster_sync();               // barriers of all 8 CTAs are initialised before any remote store / arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: W slice once, then gates_x tiles (step, gate) =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull_bar, LC_KB * LC_KTILE_BYTES);
      for (int kb = 0; kb < LC_KB; ++kb)
        tma_load_2d(sW + kb * LC_KTILE_BYTES, &tmW, wfull_bar, kb * 32, dir * 4 * LC_H + (int)rank * 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        for (int g = 0; g < 4; ++g) {
          mbar_wait_quiet(&gempty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&gfull_bar[stage], p.gx_bf16 ? LC_TILE_BYTES / 2 : LC_TILE_BYTES);
          tma_load_2d(sG + stage * LC_TILE_BYTES, &tmGx, &gfull_bar[stage],
                      (t * 2 + dir) * 4 * LC_H + g * LC_H + (int)rank * 32, crop0);
          if (++stage == LC_GX_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(LC_M, 128);
      mbar_wait_quiet(wfull_bar, 0);
      tc_fence_after();
      for (int step = 0; step < T; ++step) {
        const int acc = step % LC_ACC;
        const uint32_t acc_par = (uint32_t)((step / LC_ACC) & 1);
        if (step > 0) lc_mbar_wait_cluster(&hfull_bar[step & 1], (uint32_t)(((step - 1) >> 1) & 1));   // h_step complete
        mbar_wait_quiet(&tinit_bar[acc], acc_par);                                                     // gates_x in place
        tc_fence_after();
        const uint32_t h_addr = smem_u32(sH + (step & 1) * LC_H_BUF_BYTES);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
#pragma unroll
        for (int kb = 0; kb < LC_KB; ++kb) {
          const uint32_t a_addr = h_addr + kb * LC_KTILE_BYTES;
          const uint32_t b_addr = smem_u32(sW + kb * LC_KTILE_BYTES);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(d_tmem, make_smem_desc<64>(a_addr + k * 32), make_smem_desc<64>(b_addr + k * 32), idesc, 1u);
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp < 2 + LC_EPI_WARPS) {
    // ===================== epilogue: gate math, c in registers, h -> all CTAs of the cluster + global ============
    const int ew = warp - 2;                 // 0..15
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int uq = ew >> 2;                  // which 8 of this CTA's 32 units
    const int row = quarter * 32 + lane;     // crop row within the tile
    const int crop = crop0 + row;
    const bool valid = crop < p.B;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int unit0 = (int)rank * 32 + uq * 8;            // first of this thread's 8 hidden units
    // this CTA's slice of h is k-block `rank` of the buffer: [128 rows][64 B], 64B swizzle (chunk ^= (row >> 1) & 3);
    // this thread owns 16-byte chunk uq of its row
    const uint32_t slice_off = (uint32_t)((int)rank * LC_KTILE_BYTES);
    const uint32_t h_off = slice_off + (uint32_t)(row * 64 + ((uq ^ ((row >> 1) & 3)) << 4));
    const uint32_t ep_tid = (uint32_t)(threadIdx.x - 64);   // 0..511
    float cst[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cst[j] = 0.f;
    for (int step = 0; step < T; ++step) {
      const int t = dir ? (T - 1 - step) : step;
      const int acc = step % LC_ACC;
      mbar_wait_quiet(&tfull_bar[acc], (uint32_t)((step / LC_ACC) & 1));
      tc_fence_after();
      uint32_t a[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld8(lane_addr + (uint32_t)(acc * 128 + g * 32 + uq * 8), a[g]);
      tmem_ld_wait();
      // gate pre-activations are in registers: hand the accumulator back to the loaders
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      float hv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gi = __uint_as_float(a[0][j]), gf = __uint_as_float(a[1][j]);
        const float gg = __uint_as_float(a[2][j]), go = __uint_as_float(a[3][j]);
        const float c_new = fmaf(sigmoid_approx(gf), cst[j], sigmoid_approx(gi) * tanh_approx(gg));
        cst[j] = c_new;
        hv[j] = sigmoid_approx(go) * tanh_approx(c_new);
      }
      const uint4 hpk = make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]),
                                   pack_bf16x2(hv[6], hv[7]));
      if (step + 1 < T) {
        const int nb = (step + 1) & 1;
        uint8_t* hb = sH + nb * LC_H_BUF_BYTES;
        *reinterpret_cast<uint4*>(hb + h_off) = hpk;           // own copy of the slice (also the source of the bulk copies)
        fence_proxy_async_smem();
        // the bulk copies this thread issued one step ago (from the OTHER buffer, which is rewritten next step) have
        // finished reading shared memory: every thread passes the barrier below only after this
        if (ep_tid < LC_CLUSTER) bulk_wait_read<0>();
        asm volatile("bar.sync 1, 512;" ::: "memory");         // the whole 8 KB slice is written and fenced
        if (ep_tid < LC_CLUSTER) {
          const uint32_t src = smem_u32(hb) + slice_off;
          if (ep_tid == rank) {
            // own barrier: the local slice is in place, and 7 peer slices of 8 KB are expected
            mbar_arrive_expect_tx(&hfull_bar[nb], (LC_CLUSTER - 1) * LC_KTILE_BYTES);
          } else {
            lc_bulk_copy_to_peer(lc_map_to_cta(src, ep_tid), src, LC_KTILE_BYTES,
                                 lc_map_to_cta(smem_u32(&hfull_bar[nb]), ep_tid));
            bulk_commit();
          }
        }
      }
      // the output copy goes out after the slice has been published (a store in flight would stall the fence)
      if (valid) *reinterpret_cast<uint4*>(p.out + ((size_t)crop * T + t) * (2 * LC_H) + dir * LC_H + unit0) = hpk;
    }
    if (ep_tid < LC_CLUSTER) bulk_wait_read<0>();   // this CTA's shared memory is no longer being read by a bulk copy
  } else {
    // ===================== gates_x loaders: shared-memory tile -> registers -> accumulator buffer ==========
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int crop = crop0 + row;
    // L2 prefetch of this crop's 128-byte segments is left to the TMA unit's own L2 promotion: the loaders run up to
    // two steps ahead of the MMAs, which hides the DRAM latency of the gates_x tiles
    (void)crop;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t srow = smem_u32(sG) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    int stage = 0;
    uint32_t phase = 0;
    for (int step = 0; step < T; ++step) {
      const int acc = step % LC_ACC;
      mbar_wait_quiet(&tempty_bar[acc], (uint32_t)(((step / LC_ACC) & 1) ^ 1));   // epilogue of step-3 has drained it
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t r[32];
        mbar_wait_quiet(&gfull_bar[stage], phase);
        if (p.gx_bf16) {
          lds_row32_bf16_as_f32(smem_u32(sG) + (uint32_t)stage * LC_TILE_BYTES, row, r);
        } else {
          const uint32_t base = srow + (uint32_t)stage * LC_TILE_BYTES;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            lds_v4(base + (((uint32_t)c ^ sw) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        }
        tmem_st32(lane_addr + (uint32_t)(acc * 128 + g * 32), r);
        // the store has issued, so the shared-memory reads that feed it have completed: release the stage
        __syncwarp();
        if (lane == 0) mbar_arrive(&gempty_bar[stage]);
        if (++stage == LC_GX_STAGES) { stage = 0; phase ^= 1; }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tinit_bar[acc]);
    }
  }

  // no CTA may leave (or free TMEM) while a peer can still write its shared memory or signal its barriers
  tc_fence_before();
  lc_cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static TmaEncoders g_lstm_cluster_enc;

static cudaLaunchConfig_t lc_launch_config(int tiles, cudaStream_t stream, cudaLaunchAttribute* attr) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LC_CLUSTER * tiles, 1, 1);
  cfg.blockDim = dim3(LC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = LC_SMEM_BYTES;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LC_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cfg;
}

static int lc_prepare() {
  static unsigned long long smem_set = 0;
  return ensure_dynamic_smem((const void*)bilstm_cluster_kernel, LC_SMEM_BYTES, &smem_set);
}

// How many 8-CTA clusters of this kernel the device can hold at once (each CTA takes a whole SM and the 8 SMs of a
// cluster must sit in one GPC): 13 on the B200s of this pool.  0 if the query fails.
int bilstm_cluster_capacity() {
  static std::atomic<int> cached[64];             // per device ordinal; value + 1 (0 = not queried yet)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  const int have = cached[dev].load(std::memory_order_relaxed);
  if (have > 0) return have - 1;
  int n = 0;
  if (lc_prepare() == MSTR_OK) {
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = lc_launch_config(64, nullptr, attr);
    if (cudaOccupancyMaxActiveClusters(&n, bilstm_cluster_kernel, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
  }
  cached[dev].store(n + 1, std::memory_order_relaxed);
  return n;
}

// Called by mstr_bilstm_recurrent_bf16 (bilstm_tc.cu).
int bilstm_recurrent_bf16_cluster(const void* gates_x, int gx_bf16, const void* w_hh_packed, void* out_bf16, int B, int T,
                                  cudaStream_t stream) {
  int rc = load_tma_encoders(g_lstm_cluster_enc);
  if (rc != MSTR_OK) return rc;
  alignas(64) CUtensorMap tmW, tmGx;
  rc = encode_tiled_2d_bf16(g_lstm_cluster_enc, &tmW, w_hh_packed, LC_H, 2 * 4 * LC_H, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != MSTR_OK) return rc;
  {
    // gates_x viewed as [B rows][T*2*4H f32 columns]; box = 32 columns x 128 crops, rows beyond B are zero filled
    cuuint64_t dims[2] = {(cuuint64_t)T * 2 * 4 * LC_H, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)T * 2 * 4 * LC_H * (gx_bf16 ? 2 : 4)};
    cuuint32_t box[2] = {32, LC_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_lstm_cluster_enc.tiled(&tmGx, gx_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                          const_cast<void*>(gates_x), dims,
                                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          gx_bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeTiled(gates_x) failed with CUresult %d", (int)r);
  }
  rc = lc_prepare();
  if (rc != MSTR_OK) return rc;
  LstmClusterParams p{gates_x, (__nv_bfloat16*)out_bf16, B, T, gx_bf16};
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = lc_launch_config(ceil_div(B, LC_M) * 2, stream, attr);
  MSTR_CUDA(cudaLaunchKernelEx(&cfg, bilstm_cluster_kernel, tmW, tmGx, p));
  MSTR_LAUNCH_CHECK("bilstm_cluster_kernel");
  return MSTR_OK;
}

}  // namespace mstr
