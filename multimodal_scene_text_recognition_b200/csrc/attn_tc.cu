// Attention decoder for the bf16 throughput mode: the three GEMMs of every decode step run on the tensor cores, the
// Bahdanau attention on all SMs, and the whole greedy (or teacher-forced) loop is ONE C-ABI call
// (mstr_attn_decode_bf16).  A step has three phases:
//
//   head_i   [hp_i | logits_{i-1}] = h_i [W_h2h ; W_gen]^T      tcgen05 GEMM, N = 384 (3 tiles of 128)
//            hp tiles: += b_h2h, fp32, out through swizzled shared-memory slabs and TMA stores;
//            logits tile: transposed through shared memory, then one warp per crop row adds b_gen, stores the row with
//            coalesced stores and finds the argmax with shuffles -> id, next token, [s] bookkeeping
//   ctx_i    e_t = w_score . tanh(proj[b,t,:] + hp_i[b,:]), alpha = softmax_T(e), ctx = sum_t alpha_t enc[b,t,:]
//            the time steps of a crop are split into FOUR interleaved subsets, each reduced with its own online
//            softmax (16-byte bf16 loads one iteration ahead, lane = 8 channels, shuffle reductions, tanh.approx) and
//            combined in a fixed order; 1, 2 or 4 warps share the four subsets of a crop, so the arithmetic - and the
//            result, bit for bit - does not depend on how many warps a batch size gets; writes ctx (bf16) into the left
//            half of the [ctx | h] operand buffer
//   cell_i   gates = [ctx_i | h_i] [W_ctx | W_hh]^T + W_tok[token_i] + bias -> (h_{i+1}, c_{i+1})
//            tcgen05 GEMM, N = 1024 with rows packed (chunk of 16 units, gate, unit) so that a 64- or 128-column tile
//            carries i,f,g,o of the same units and the LSTM cell is thread-local in the epilogue; c stays f32 in global
//            memory (L2 resident, transposed [H][B]), h_{i+1} is written as bf16 into the right half of the other
//            [ctx | h] buffer
// Every GEMM epilogue runs on EIGHT warps (two per TMEM lane quarter, each half of the tile's columns).
//
// Two drivers run the same phase bodies (same tiles, MMA order and epilogue arithmetic: bit-identical results):
//   * attn_decode_persistent_kernel (default): ONE cooperative launch, one CTA per SM, resident for all steps; the
//     phases are distributed over the grid and separated by grid-wide barriers; TMEM, mbarriers and tensor maps are
//     set up once;
//   * the launch chain (mstr_set_decode_persistent(0)): three launches per step with programmatic dependent launch -
//     the cross-check of the persistent kernel's barriers and proxy fences (tests/test_gpu_kernels.py).
//
// Replaces modules/decoders.py:26-64 (Attention.forward greedy :52-62 and teacher-forced :41-50 branches), :79-101
// (AttentionCell.forward), :17/:59 (generator), :19-24 (_char_to_onehot: a one-hot times W_ih is a row gather of
// W_tok).  proj = i2h(enc) is loop invariant (:81) and computed once by the caller.  Differences to the fp32 SIMT
// kernel (attn_decode.cu): proj / enc / h / ctx are bf16 operands (fp32 accumulation everywhere), tanh.approx in the
// attention scores.
#include <atomic>

#include "tc_common.cuh"

namespace mstr {

constexpr int AT_H = 256;
constexpr int AT_BLOCK = 128;      // BLOCK_M = BLOCK_N
constexpr int AT_STAGES = 3;
constexpr int AT_TILE_BYTES = 128 * 128;   // [128 rows][64 k] bf16, 128B swizzle
constexpr int AT_EPI_WARPS = 8;            // two per TMEM lane quarter
constexpr int AT_EPI_THREADS = 32 * AT_EPI_WARPS;
constexpr int AT_THREADS = 64 + AT_EPI_THREADS;   // launch chain: producer, MMA issuer, 8 epilogue warps
constexpr int AT_STG_LD = AT_BLOCK + 1;    // odd row stride of the transposing stage (conflict-free row-wise writes)
constexpr int AT_STG_BYTES = AT_BLOCK * AT_STG_LD * 4 + 512;   // 66560: stage, or 4 f32 output slabs + 128 bias values
constexpr int AT_SMEM_BYTES = 1024 + AT_STAGES * 2 * AT_TILE_BYTES + 256 + 512;
static_assert(4 * EPI_SLAB_BYTES + 512 <= AT_STG_BYTES, "hp slabs + bias fit in the stage");
static_assert(AT_STG_BYTES <= AT_STAGES * 2 * AT_TILE_BYTES, "the launch chain reuses its operand stages as the stage");

struct DecodeState {
  // operand / state buffers (device)
  __nv_bfloat16* xh[2];   // [B][512]: cols 0..255 ctx_i, cols 256..511 h_i   (double buffered by step parity)
  float* hp;              // [B][256]
  float* cst;             // [256][B] (transposed: the cell epilogue's lane = crop accesses are coalesced)
  int32_t* tok;           // [B] token fed to the current step
  int32_t* done;          // [B] crop has emitted [s]
  int32_t* live;          // [1] number of crops still decoding (early exit); live[1] = grid-barrier arrival counter
  // parameters
  const float* b_h2h; const float* b_gen; const float* w_score; const float* w_tok; const float* bias;
  const int64_t* text;    // optional teacher forcing [B][steps]
  // outputs
  float* logits; int32_t* ids; int32_t* lengths; float* alpha_out;
  int B, T, C, steps, early_exit;
};

// Programmatic dependent launch (launch chain): every kernel of the chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be scheduled (and run their prologue:
// barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel drains; griddep_wait() blocks
// until the previous grid has completed and its writes are visible, and must precede every dependent access.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// A teacher-forcing token outside [0, C) would index past W_tok (the reference's scatter_ raises, decoders.py:23): it is
// mapped to 0 = [GO]; the host side validates the text under config.CHECK_INPUTS.
__device__ __forceinline__ int at_token(const int64_t t, const int C) { return (t < 0 || t >= C) ? 0 : (int)t; }

__device__ __forceinline__ float at_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float at_tanh(float x) { return fmaf(2.f, at_sigmoid(2.f * x), -1.f); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

// One of the eight epilogue warps: TMEM lane quarter (= warp index mod 4, a hardware rule), column half, lane.
struct EpiWho {
  int quarter, half, lane;
  __device__ __forceinline__ int row() const { return quarter * 32 + lane; }       // crop row of the tile
  __device__ __forceinline__ int tid() const { return half * 128 + quarter * 32 + lane; }   // 0..255
};

// ======================================================================================================
// phase bodies shared by the persistent kernel and the launch chain
// ======================================================================================================

// ---- head epilogue, column tiles 0 and 1: hp_i = h_i W_h2h^T + b_h2h -----------------------------------
// The accumulator layout gives every thread one crop ROW, and rows are 1 KB apart in global memory: direct stores cost
// 32 L1 wavefronts per instruction, and a transposing stage with per-row copy loops was measured at 3.3 us per tile.
// Each column half goes through its own pair of swizzled [128 x 32] f32 slabs and leaves by TMA store
// (tc_common.cuh: epi_stage_store); the issuing threads wait for their stores before the phase ends, because the
// attention phase of OTHER CTAs reads hp.  `stage_prologue` (bias -> shared memory) runs before the accumulator is
// awaited.  Called by all 256 epilogue threads.
template <typename WaitAcc, typename Release>
__device__ __forceinline__ void at_head_hp_epilogue(const DecodeState& s, const CUtensorMap* tmHp, uint8_t* stg,
                                                    const uint32_t taddr, const EpiWho& w, const int m0, const int n_tile,
                                                    WaitAcc wait_acc, Release release) {
  float* sBias = reinterpret_cast<float*>(stg + 4 * EPI_SLAB_BYTES);
  if (w.tid() < AT_BLOCK) sBias[w.tid()] = __ldg(s.b_h2h + n_tile * AT_BLOCK + w.tid());
  wait_acc();                                            // also orders the bias staging before its use (bar.sync inside)
  EpiStage est{stg + w.half * 2 * EPI_SLAB_BYTES, EPI_SLAB_BYTES, 2, 0, 2 + 2 * w.half};
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int c0 = w.half * 64 + c * 32;
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
    if (c == 1) {
      tc_fence_before();
      release();
    }
    float y[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(v[j]) + sBias[c0 + j];
    epi_stage_store<true>(est, tmHp, y, w.row(), n_tile * AT_BLOCK + c0, m0, w.quarter * 32 + w.lane);
  }
  if (w.quarter == 0 && w.lane == 0) {                   // the thread that issued this half's stores
    bulk_wait_all();
    fence_proxy_async_global();
  }
}

// ---- head epilogue, column tile 2: logits of decode step (step-1) = h_step W_gen^T + b_gen ---------------
// Phase 1: the raw accumulator is transposed into the stage (lane = row writes with an odd stride).  Phase 2: one warp
// per crop row, lanes = classes: + b_gen, coalesced stores of the row, argmax by shuffles (the lowest index wins ties,
// like torch.max), then the lanes do the per-crop bookkeeping of their rows.  (The first version kept the argmax
// thread-local in phase 1: 97 dependent load-add-compare-store rounds per thread, 6 us per tile.)
template <typename WaitAcc, typename Release>
__device__ __forceinline__ void at_head_logits_epilogue(const DecodeState& s, float* stg, const uint32_t taddr, const EpiWho& w,
                                                        const int m0, const int step, WaitAcc wait_acc, Release release) {
  constexpr int LD = AT_STG_LD;
  constexpr int ROWS = AT_BLOCK / AT_EPI_WARPS;          // 16 rows per warp in phase 2
  const int ls = step - 1;
  const int w8 = w.half * 4 + w.quarter;                 // 0..7
  const int r0 = w8 * ROWS;
  // per-row state of phase 2, lane L < ROWS <-> row r0 + L: loaded before the accumulator is awaited
  const int my_crop = m0 + r0 + w.lane;
  const bool my_valid = w.lane < ROWS && my_crop < s.B;
  const int my_done = my_valid ? __ldcg(s.done + my_crop) : 0;
  float bg[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bg[q] = (q * 32 + w.lane < s.C) ? __ldg(s.b_gen + q * 32 + w.lane) : 0.f;
  wait_acc();
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int c0 = w.half * 64 + c * 32;
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) stg[w.row() * LD + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  release();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  int my_best = 0;
#pragma unroll 4
  for (int rr = 0; rr < ROWS; ++rr) {
    const int crop = m0 + r0 + rr;
    if (crop >= s.B) break;                              // uniform over the warp
    const int row_done = __shfl_sync(0xffffffffu, my_done, rr);
    float y[4];
    float best = -INFINITY;
    int besti = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int cls = q * 32 + w.lane;
      y[q] = stg[(r0 + rr) * LD + cls] + bg[q];
      if (cls < s.C && y[q] > best) { best = y[q]; besti = cls; }     // ascending classes: the lowest index wins ties
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (!(s.early_exit && row_done)) {
      float* lrow = s.logits + ((size_t)crop * s.steps + ls) * s.C;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q * 32 + w.lane < s.C) lrow[q * 32 + w.lane] = y[q];
    }
    if (w.lane == rr) my_best = besti;
  }
  if (my_valid) {
    const bool was_done = s.early_exit && my_done != 0;
    if (was_done) {
      s.ids[(size_t)my_crop * s.steps + ls] = 0;         // argmax of the zero-initialised logits row
    } else {
      s.ids[(size_t)my_crop * s.steps + ls] = my_best;
      if (my_best == 1 && my_done == 0) {
        if (s.lengths && __ldcg(s.lengths + my_crop) == s.steps) s.lengths[my_crop] = ls;
        if (s.early_exit) {
          s.done[my_crop] = 1;
          atomicSub(s.live, 1);
        }
      }
    }
    if (step < s.steps) s.tok[my_crop] = s.text ? at_token(s.text[(size_t)my_crop * s.steps + step], s.C) : my_best;
  }
}

// ---- cell epilogue: this tile = 32 (N64: 16) hidden units x 4 gates of 128 crops; this warp: 16 (8) of the units ----
// Weight rows (= accumulator columns) are ordered [chunk of 16 units][gate][unit] (ops.pack_gate_rows_bf16): a 128-column
// tile is two such chunks, one per warp half; a 64-column tile (N64, small batches: half the MMA and half the epilogue
// per CTA, twice the CTAs) is one chunk whose 16 units the two warp halves share.
// Everything the LSTM cell needs besides the accumulator - the token's row of W_tok (+ bias) and the previous cell
// state - does not depend on this phase's GEMM, and the per-lane (row = crop, rows 4 KB apart) gathers cost 32 L1
// wavefronts per instruction: ALL of them are issued before the accumulator is awaited (`wait_acc` blocks on the TMA
// loads and the MMAs, which do not use the load/store path), so that what follows the wait is only TMEM loads, the
// gate math and the stores.  (Loading one group of 4 units ahead left ~3 of the 5 us of this epilogue exposed.)
template <bool N64, typename WaitAcc, typename Release>
__device__ __forceinline__ void at_cell_epilogue(const DecodeState& s, const uint32_t taddr, const EpiWho& w, const int m0,
                                                 const int n_tile, const int step, WaitAcc wait_acc, Release release) {
  constexpr int UNITS = N64 ? 8 : 16;                    // units of this warp
  constexpr int NSUB = UNITS / 4;
  const int crop = m0 + w.row();
  const bool valid = crop < s.B;
  const int unit0 = N64 ? n_tile * 16 + w.half * 8 : n_tile * 32 + w.half * 16;
  const uint32_t col0 = N64 ? (uint32_t)(w.half * 8) : (uint32_t)(w.half * 64);   // + gate * 16 + unit within the chunk
  const int tok = valid ? __ldcg(s.tok + crop) : 0;
  const float* wt = s.w_tok + (size_t)tok * (4 * AT_H);
  float* ccol = s.cst + (size_t)unit0 * s.B + (valid ? crop : 0);   // cell state is kept [H][B]: lane = crop is coalesced
  __nv_bfloat16* hrow = s.xh[(step + 1) & 1] + (size_t)(valid ? crop : 0) * (2 * AT_H) + AT_H + unit0;
  float4 add4[NSUB][4];                                  // [group of 4 units][gate]
  float cprev[UNITS];
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = g * AT_H + unit0 + sub * 4;
      float4 t = valid ? __ldg(reinterpret_cast<const float4*>(wt + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (s.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(s.bias + col));
        t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
      }
      add4[sub][g] = t;
    }
  }
#pragma unroll
  for (int j = 0; j < UNITS; ++j) cprev[j] = valid ? __ldcg(ccol + (size_t)j * s.B) : 0.f;
  wait_acc();
  uint32_t hpk[4];
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
    uint32_t a[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) tmem_ld4(taddr + col0 + (uint32_t)(g * 16 + sub * 4), a[g]);
    tmem_ld_wait();
    if (sub == NSUB - 1) {                                    // the accumulator is in registers: hand it back
      tc_fence_before();
      release();
    }
    const float add[4][4] = {{add4[sub][0].x, add4[sub][0].y, add4[sub][0].z, add4[sub][0].w},
                             {add4[sub][1].x, add4[sub][1].y, add4[sub][1].z, add4[sub][1].w},
                             {add4[sub][2].x, add4[sub][2].y, add4[sub][2].z, add4[sub][2].w},
                             {add4[sub][3].x, add4[sub][3].y, add4[sub][3].z, add4[sub][3].w}};
    float hv[4], cn4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = __uint_as_float(a[0][j]) + add[0][j];
      const float gf = __uint_as_float(a[1][j]) + add[1][j];
      const float gg = __uint_as_float(a[2][j]) + add[2][j];
      const float go = __uint_as_float(a[3][j]) + add[3][j];
      const float cn = fmaf(at_sigmoid(gf), cprev[sub * 4 + j], at_sigmoid(gi) * at_tanh(gg));
      cn4[j] = cn;
      hv[j] = at_sigmoid(go) * at_tanh(cn);
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ccol[(size_t)(sub * 4 + j) * s.B] = cn4[j];
    }
    hpk[(sub & 1) * 2] = pack_bf16x2(hv[0], hv[1]);
    hpk[(sub & 1) * 2 + 1] = pack_bf16x2(hv[2], hv[3]);
    if ((sub & 1) && valid) *reinterpret_cast<uint4*>(hrow + (sub - 1) * 4) = make_uint4(hpk[0], hpk[1], hpk[2], hpk[3]);
  }
}

// ---- Bahdanau attention of one crop ----------------------------------------------------------------
// The time steps are split into AT_VSUB = 4 interleaved subsets (t = v, v + 4, ...).  Each subset is reduced in a single
// pass with an online softmax: every iteration scores four time steps of proj and folds them into a running (max,
// denominator, context) while the eight 16-byte loads of the NEXT iteration are already in flight.  The partial
// results go to shared memory and are combined in the fixed order v = 0..3.  W = 1, 2 or 4 warps share the four subsets
// of a crop (warp `sub` takes v = sub, sub + W, ...): the arithmetic is the same for every W, so a crop's result does not
// depend on the batch it arrived in.  The dependent chain per crop is 2 * 4 / W memory round trips - with few crops
// per SM that chain IS the phase time (one warp walking all 26 steps: 7 round trips, 9 us).
// `part`: [AT_VSUB][AT_PART] floats of this crop group (and round parity); `bar_id`: named barrier of the group (W > 1).
constexpr int AT_VSUB = 4;
constexpr int AT_PART = 4 + AT_H;
template <int W>
__device__ __forceinline__ void at_attention_crop(const __nv_bfloat16* __restrict__ enc, const __nv_bfloat16* __restrict__ proj,
                                                  const DecodeState& s, const int step, const int crop, const int lane,
                                                  const int sub, float* part, const int bar_id) {
  const int T = s.T;
  float hp[8], ws[8];
  {
    const float4* hp4 = reinterpret_cast<const float4*>(s.hp + (size_t)crop * AT_H + lane * 8);
    const float4* ws4 = reinterpret_cast<const float4*>(s.w_score + lane * 8);
    const float4 a = __ldcg(hp4), b = __ldcg(hp4 + 1), c = __ldg(ws4), d = __ldg(ws4 + 1);
    hp[0] = a.x; hp[1] = a.y; hp[2] = a.z; hp[3] = a.w; hp[4] = b.x; hp[5] = b.y; hp[6] = b.z; hp[7] = b.w;
    ws[0] = c.x; ws[1] = c.y; ws[2] = c.z; ws[3] = c.w; ws[4] = d.x; ws[5] = d.y; ws[6] = d.z; ws[7] = d.w;
  }
  const uint4* pj = reinterpret_cast<const uint4*>(proj + (size_t)crop * T * AT_H) + lane;   // 32 uint4 per time step
  const uint4* en = reinterpret_cast<const uint4*>(enc + (size_t)crop * T * AT_H) + lane;
  constexpr int MAX_SLOTS = 3;   // T <= 96
  float e_slot[MAX_SLOTS];       // score of time step t lives in lane t % 32, slot t / 32 (for alpha_out)
#pragma unroll
  for (int k = 0; k < MAX_SLOTS; ++k) e_slot[k] = -INFINITY;
#pragma unroll 1
  for (int v = sub; v < AT_VSUB; v += W) {
    const int nj = v < T ? (T - v + AT_VSUB - 1) / AT_VSUB : 0;      // time steps of this subset: t = v + 4 j
    float mx = -INFINITY, den = 0.f;
    float cx[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cx[q] = 0.f;
    uint4 np[4], ne[4];
    auto load_iter = [&](const int j0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = j0 + u < nj;
        const size_t t = (size_t)(v + AT_VSUB * (j0 + u));
        np[u] = in ? __ldg(pj + t * 32) : make_uint4(0u, 0u, 0u, 0u);
        ne[u] = in ? __ldg(en + t * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    if (nj > 0) load_iter(0);
    for (int j0 = 0; j0 < nj; j0 += 4) {
      uint4 vp[4], ve[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { vp[u] = np[u]; ve[u] = ne[u]; }
      if (j0 + 4 < nj) load_iter(j0 + 4);
      float e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w4[4] = {vp[u].x, vp[u].y, vp[u].z, vp[u].w};
        float p = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w4[q]);
          p = fmaf(ws[2 * q], tanh_approx(f.x + hp[2 * q]), p);
          p = fmaf(ws[2 * q + 1], tanh_approx(f.y + hp[2 * q + 1]), p);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        const bool in = j0 + u < nj;
        const int t = v + AT_VSUB * (j0 + u);
        e[u] = in ? p : -INFINITY;
        if (in && (t & 31) == lane) {
#pragma unroll
          for (int k = 0; k < MAX_SLOTS; ++k)
            if ((t >> 5) == k) e_slot[k] = p;
        }
      }
      // online softmax update with the four new scores (all lanes hold the same e[], mx, den)
      const float m_new = fmaxf(fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])), mx);
      const float rescale = __expf(mx - m_new);          // 0 on the first iteration (mx = -inf)
      den *= rescale;
#pragma unroll
      for (int q = 0; q < 8; ++q) cx[q] *= rescale;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float a = __expf(e[u] - m_new);            // 0 for padded time steps (e = -inf)
        den += a;
        const uint32_t w4[4] = {ve[u].x, ve[u].y, ve[u].z, ve[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w4[q]);
          cx[2 * q] = fmaf(a, f.x, cx[2 * q]);
          cx[2 * q + 1] = fmaf(a, f.y, cx[2 * q + 1]);
        }
      }
      mx = m_new;
    }
    float* mine = part + v * AT_PART;
    if (lane == 0) { mine[0] = mx; mine[1] = den; }
    *reinterpret_cast<float4*>(mine + 4 + lane * 8) = make_float4(cx[0], cx[1], cx[2], cx[3]);
    *reinterpret_cast<float4*>(mine + 8 + lane * 8) = make_float4(cx[4], cx[5], cx[6], cx[7]);
  }
  // combine the four partial results (every warp of the group reads all of them)
  if constexpr (W > 1) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * W) : "memory");
  else __syncwarp();
  float m = -INFINITY;
#pragma unroll
  for (int v = 0; v < AT_VSUB; ++v) m = fmaxf(m, part[v * AT_PART]);
  float den = 0.f;
  float cx[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) cx[q] = 0.f;
#pragma unroll
  for (int v = 0; v < AT_VSUB; ++v) {
    const float sc = __expf(part[v * AT_PART] - m);      // 0 for a subset without time steps (max = -inf)
    den = fmaf(part[v * AT_PART + 1], sc, den);
    const float4 c0 = *reinterpret_cast<const float4*>(part + v * AT_PART + 4 + lane * 8);
    const float4 c1 = *reinterpret_cast<const float4*>(part + v * AT_PART + 8 + lane * 8);
    cx[0] = fmaf(c0.x, sc, cx[0]); cx[1] = fmaf(c0.y, sc, cx[1]); cx[2] = fmaf(c0.z, sc, cx[2]); cx[3] = fmaf(c0.w, sc, cx[3]);
    cx[4] = fmaf(c1.x, sc, cx[4]); cx[5] = fmaf(c1.y, sc, cx[5]); cx[6] = fmaf(c1.z, sc, cx[6]); cx[7] = fmaf(c1.w, sc, cx[7]);
  }
  const float inv = 1.f / den;
  if (s.alpha_out) {
#pragma unroll
    for (int k = 0; k < MAX_SLOTS; ++k) {
      const int t = k * 32 + lane;
      if (t < T && ((t % AT_VSUB) % W) == sub) s.alpha_out[((size_t)crop * s.steps + step) * T + t] = __expf(e_slot[k] - m) * inv;
    }
  }
  if (sub == 0) {
    __nv_bfloat16* dst = s.xh[step & 1] + (size_t)crop * (2 * AT_H) + lane * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(cx[0] * inv, cx[1] * inv), pack_bf16x2(cx[2] * inv, cx[3] * inv),
                                                pack_bf16x2(cx[4] * inv, cx[5] * inv), pack_bf16x2(cx[6] * inv, cx[7] * inv));
  }
  if constexpr (W == 1) __syncwarp();                    // the partials are rewritten by this warp's next crop
}

// Warps per crop in the attention phase: as many as keep every SM's warps busy in one round (the result does not
// depend on it).
static inline int attention_warps_per_crop(int B, int sms) {
  const int warps = sms * 16;
  return 4 * B <= warps ? 4 : (2 * B <= warps ? 2 : 1);
}

// ======================================================================================================
// launch chain
// ======================================================================================================
// EPI 0: head (hp + logits/argmax), EPI 1: LSTM cell.  One 128x128 output tile per CTA; K = num_k * 64.
template <int EPI>
__global__ void __launch_bounds__(AT_THREADS, 1)
rnn_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmHp, const DecodeState s, const int step, const int num_k) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + AT_STAGES * AT_TILE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * AT_STAGES * AT_TILE_BYTES);
  uint64_t* empty_bar = full_bar + AT_STAGES;
  uint64_t* tfull_bar = empty_bar + AT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m0 = blockIdx.y * AT_BLOCK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (EPI == 0) prefetch_tensormap(&tmHp);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                                        // the previous kernel of the chain has completed
  griddep_launch();
  const bool idle = s.early_exit && *s.live == 0;        // every crop has emitted [s]

  if (idle) {
    // nothing to do for this step
  } else if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], 2 * AT_TILE_BYTES);
        tma_load_2d(sA + stage * AT_TILE_BYTES, &tmA, &full_bar[stage], kb * 64, m0);
        tma_load_2d(sB + stage * AT_TILE_BYTES, &tmB, &full_bar[stage], kb * 64, n_tile * AT_BLOCK);
        if (++stage == AT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(AT_BLOCK, AT_BLOCK);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + stage * AT_TILE_BYTES);
        const uint32_t b_addr = smem_u32(sB + stage * AT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, make_smem_desc<128>(a_addr + k * 32), make_smem_desc<128>(b_addr + k * 32), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == num_k - 1) umma_commit(tfull_bar);
        if (++stage == AT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM lane = crop row; warps 2..9 =====================
    const EpiWho w{warp & 3, (warp - 2) >> 2, lane};
    const uint32_t taddr = tmem_base + ((uint32_t)(w.quarter * 32) << 16);
    auto nothing = [] {};
    if constexpr (EPI == 0) {
      // the operand stages are idle once the accumulator is complete: they double as the stage - so nothing may be
      // written there before the MMAs have finished (the bias staging of the hp tiles comes after the wait)
      auto wait_first = [&] { mbar_wait(tfull_bar, 0); tc_fence_after(); };
      if (n_tile < 2) {
        wait_first();
        at_head_hp_epilogue(s, &tmHp, smem, taddr, w, m0, n_tile, [&] { asm volatile("bar.sync 1, 256;" ::: "memory"); }, nothing);
      } else if (step > 0) {
        at_head_logits_epilogue(s, reinterpret_cast<float*>(smem), taddr, w, m0, step, wait_first, nothing);
      }
    } else {
      at_cell_epilogue<false>(s, taddr, w, m0, n_tile, step, [&] { mbar_wait(tfull_bar, 0); tc_fence_after(); }, nothing);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// Attention phase of the launch chain: 4 warps per block = 4 / W crops.
constexpr int AT_CTX_WARPS = 4;
template <int W>
__global__ void __launch_bounds__(32 * AT_CTX_WARPS, 4)
attn_context_kernel(const __nv_bfloat16* __restrict__ enc, const __nv_bfloat16* __restrict__ proj,
                    const DecodeState s, const int step) {
  __shared__ __align__(16) float part[(AT_CTX_WARPS / W) * AT_VSUB * AT_PART];
  griddep_wait();
  griddep_launch();
  if (s.early_exit && *s.live == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crop = blockIdx.x * (AT_CTX_WARPS / W) + warp / W;
  if (crop >= s.B) return;                               // uniform over the W warps of a group
  at_attention_crop<W>(enc, proj, s, step, crop, lane, warp % W, part + (warp / W) * AT_VSUB * AT_PART, 2 + warp / W);
}

// token_0 = [GO] (0) or text[:, 0]; lengths = steps; done = 0; live = B; ids zero-filled.
__global__ void decode_init_kernel(const DecodeState s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    *s.live = s.B;
    s.live[1] = 0;          // arrival counter of the persistent kernel's grid barrier
  }
  if (i < 256) s.live[2 + i] = 0;   // PD_DYN_STEPS crop counters of the persistent kernel's attention phase
  if (i < s.B) {
    s.tok[i] = s.text ? at_token(s.text[(size_t)i * s.steps], s.C) : 0;
    s.done[i] = 0;
    if (s.lengths) s.lengths[i] = s.steps;
  }
  for (int j = i; j < s.B * s.steps; j += gridDim.x * blockDim.x) s.ids[j] = 0;
}

// ======================================================================================================
// PERSISTENT decoder: ONE cooperative launch runs all steps.
//
// The launch chain pays, per step, three launch latencies, three prologues (barrier init, TMEM allocation,
// tensor-map prefetch) and three pipeline drains.  Here one CTA per SM stays resident for the whole decode; the
// three phases of a step are distributed over the grid and separated by a grid-wide barrier (one atomic per CTA on an
// L2 word).  What changes with the persistence:
//  * buffers that one CTA writes and another reads in a later phase (hp, [ctx|h], c, tok, done, live, lengths) are
//    read with ld.global.cg (L2; the L1 of the reading SM may still hold the line of the previous step);
//  * the [ctx|h] operand is written with ordinary stores and read by TMA (async proxy) in the next phase: writers
//    execute fence.proxy.async.global before the barrier, the TMA-issuing thread after it; hp goes the other way
//    (TMA stores, read with ordinary loads): the storing thread waits for its bulk group and fences;
//  * the attention phase walks its crops in alternating direction (even steps ascending, odd steps descending): every
//    step touches all of proj+enc (109 MB at 4096 crops, about the L2 capacity), and with a fixed direction an LRU
//    cache holds exactly the lines that are needed LAST; zig-zag reuse starts with the most recently used lines.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 8..15 = GEMM epilogues; all 16 warps run the attention phase.
// ======================================================================================================
constexpr int PD_WARPS = 16;
constexpr int PD_DYN_STEPS = 256;   // per-step crop counters of the attention phase in the workspace (live[2 .. 2+256))
constexpr int PD_THREADS = 32 * PD_WARPS;
constexpr int PD_SMEM_BYTES = 1024 + AT_STAGES * 2 * AT_TILE_BYTES + AT_STG_BYTES + 256;
static_assert(PD_WARPS * AT_VSUB * AT_PART * 4 <= AT_STG_BYTES, "attention partials of all warps (W = 1) live in the stage");

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Phase trace (diagnostic, mstr_debug_decode_trace): %globaltimer of CTA 0 at every barrier entry (this CTA's phase
// work is done) and exit (every CTA's is), the entry time of EVERY CTA at the first PD_TRACE_BARRIERS barriers, and a
// fine trace of the GEMM tiles of CTAs 0..3 in step PD_FINE_STEP - phase work, barrier wait, barrier latency,
// stragglers and the stages of a tile can be told apart (tools/tune_seq.py prints them).
constexpr int PD_TRACE_MAX = 512;
constexpr int PD_TRACE_BARRIERS = 48;
constexpr int PD_TRACE_CTAS = 160;
constexpr int PD_FINE_STEP = 6;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ unsigned long long g_pd_trace[PD_TRACE_MAX];
__device__ unsigned long long g_pd_trace_cta[PD_TRACE_BARRIERS * PD_TRACE_CTAS];
// [phase (0 hp tiles, 1 cell, 2 logits tiles)][owner of tile 0..3][16 slots]: 0 phase entry, 1 first TMA issued,
// 2 first operand stage arrived, 3 last MMA committed, 4 accumulator complete (epilogue woke), 6 accumulator released,
// 7 epilogue done
__device__ unsigned long long g_pd_fine[3 * 4 * 16];
#define PD_FINE(phase_id, slot)                                                                         \
  do {                                                                                                  \
    if (step == PD_FINE_STEP && first < 4) g_pd_fine[((phase_id) * 4 + first) * 16 + (slot)] = globaltimer_ns(); \
  } while (0)

// Grid-wide barrier (all CTAs are co-resident: cooperative launch), split into ARRIVE (this CTA's work of the phase is
// done and visible) and WAIT (every CTA has arrived).  `target` is this CTA's running arrival target.  Work a CTA does
// between its arrive and its wait is work nobody waits for at THIS barrier (the logits tiles, see the kernel).
__device__ __forceinline__ void pd_barrier_arrive(unsigned int* counter, unsigned int& target, int& trace_n) {
  fence_proxy_async_global();              // this thread's global stores -> later async-proxy (TMA) reads of any CTA
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t_in = globaltimer_ns();
    const int nb = (int)(target / gridDim.x);
    if (nb < PD_TRACE_BARRIERS && blockIdx.x < PD_TRACE_CTAS) g_pd_trace_cta[nb * PD_TRACE_CTAS + blockIdx.x] = t_in;
    if (blockIdx.x == 0 && trace_n < PD_TRACE_MAX) g_pd_trace[trace_n++] = t_in;
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void pd_barrier_wait(unsigned int* counter, const unsigned int target, int& trace_n) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(counter) < target) {
      if (clock64() - t0 > 4000000000ll) {
        printf("mstr: grid barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    fence_proxy_async_global();
    if (blockIdx.x == 0 && trace_n < PD_TRACE_MAX) g_pd_trace[trace_n++] = globaltimer_ns();
  }
  __syncthreads();
}
__device__ __forceinline__ void pd_grid_barrier(unsigned int* counter, unsigned int& target, int& trace_n) {
  pd_barrier_arrive(counter, target, trace_n);
  pd_barrier_wait(counter, target, trace_n);
}

struct PdPipe {            // running ring / accumulator state of one role, carried across tiles, phases and steps
  int stage = 0;
  uint32_t phase = 0;
  uint32_t acc_phase = 0;
};

struct PdMaps {
  CUtensorMap h[2], xh[2], head, cat, cat64, hp;   // cat64: the cell weights with a 64-row box (128 x 64 tiles)
};

// One GEMM phase over this CTA's tiles: first, first + gridDim.x, ... of the n_count x m_tiles tiles (column tile
// fastest).  EPI 0: head, n_count column tiles starting at n_first; EPI 1: cell.
template <int EPI, int BN = AT_BLOCK>
__device__ __forceinline__ void pd_gemm_phase(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmHp,
                                              const DecodeState& s, const int step, const int num_k, const int first,
                                              const int n_first,
                                              const int n_count, const int m_tiles, uint8_t* sA, uint8_t* sB, uint8_t* stg,
                                              uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, const uint32_t tmem_base, PdPipe& pipe) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = n_count * m_tiles;
  const int FINE_ID = EPI == 1 ? 1 : (n_first == 2 ? 2 : 0);
  if (warp == 0) {
    if (lane == 0) {
      PD_FINE(FINE_ID, 0);       // (thread 0 executed fence.proxy.async.global when it left the last grid barrier)
      for (int tile = first; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = n_first + tile % n_count, m0 = (tile / n_count) * AT_BLOCK;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[pipe.stage], pipe.phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[pipe.stage], AT_TILE_BYTES + BN * 128);
          tma_load_2d(sA + pipe.stage * AT_TILE_BYTES, tmA, &full_bar[pipe.stage], kb * 64, m0);
          tma_load_2d(sB + pipe.stage * AT_TILE_BYTES, tmB, &full_bar[pipe.stage], kb * 64, n_tile * BN);
          if (kb == 0) PD_FINE(FINE_ID, 1);
          if (++pipe.stage == AT_STAGES) { pipe.stage = 0; pipe.phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(AT_BLOCK, BN);
      for (int tile = first; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar, pipe.acc_phase ^ 1);       // the epilogue of the previous tile has drained the accumulator
        tc_fence_after();
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[pipe.stage], pipe.phase);
          if (kb == 0) PD_FINE(FINE_ID, 2);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + pipe.stage * AT_TILE_BYTES);
          const uint32_t b_addr = smem_u32(sB + pipe.stage * AT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, make_smem_desc<128>(a_addr + k * 32), make_smem_desc<128>(b_addr + k * 32), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[pipe.stage]);
          if (kb == num_k - 1) { umma_commit(tfull_bar); PD_FINE(FINE_ID, 3); }
          if (++pipe.stage == AT_STAGES) { pipe.stage = 0; pipe.phase ^= 1; }
        }
        pipe.acc_phase ^= 1;
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: TMEM lane = crop row; warps 8..15 =====================
    const EpiWho w{warp & 3, (warp - 8) >> 2, lane};
    const uint32_t taddr = tmem_base + ((uint32_t)(w.quarter * 32) << 16);
    const bool tr = warp == 8 && lane == 0;
    auto release = [&] {                                 // one arrival per epilogue warp hands the accumulator back
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      if (tr) PD_FINE(FINE_ID, 6);
    };
    for (int tile = first; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = n_first + tile % n_count, m0 = (tile / n_count) * AT_BLOCK;
      auto wait_acc = [&] {
        mbar_wait(tfull_bar, pipe.acc_phase);
        if (tr) PD_FINE(FINE_ID, 4);
        tc_fence_after();
      };
      if constexpr (EPI == 0) {
        asm volatile("bar.sync 1, 256;" ::: "memory");     // the previous tile's rows / bias have left the stage
        if (n_tile < 2) {
          at_head_hp_epilogue(s, tmHp, stg, taddr, w, m0, n_tile,
                              [&] { wait_acc(); asm volatile("bar.sync 1, 256;" ::: "memory"); }, release);
        } else {
          at_head_logits_epilogue(s, reinterpret_cast<float*>(stg), taddr, w, m0, step, wait_acc, release);
        }
      } else {
        at_cell_epilogue<BN == 64>(s, taddr, w, m0, n_tile, step, wait_acc, release);
      }
      if (tr) PD_FINE(FINE_ID, 7);
      pipe.acc_phase ^= 1;
    }
  }
  __syncwarp();
}

template <int W>
__global__ void __launch_bounds__(PD_THREADS, 1)
attn_decode_persistent_kernel(const __grid_constant__ PdMaps maps, const __nv_bfloat16* __restrict__ enc,
                              const __nv_bfloat16* __restrict__ proj, const DecodeState s, unsigned int* gbar) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + AT_STAGES * AT_TILE_BYTES;
  uint8_t* stg = smem + 2 * AT_STAGES * AT_TILE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + AT_STG_BYTES);
  uint64_t* empty_bar = full_bar + AT_STAGES;
  uint64_t* tfull_bar = empty_bar + AT_STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.h[0]); prefetch_tensormap(&maps.h[1]);
    prefetch_tensormap(&maps.xh[0]); prefetch_tensormap(&maps.xh[1]);
    prefetch_tensormap(&maps.head); prefetch_tensormap(&maps.cat); prefetch_tensormap(&maps.cat64); prefetch_tensormap(&maps.hp);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, AT_EPI_WARPS);                   // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (s.B + AT_BLOCK - 1) / AT_BLOCK;
  // attention: groups of W adjacent warps, one crop per group and round
  constexpr int GROUPS = PD_WARPS / W;
  const int group = warp / W, sub = warp % W;
  const int gg = blockIdx.x * GROUPS + group, GG = gridDim.x * GROUPS;
  const int my_crops = gg < s.B ? (s.B - gg + GG - 1) / GG : 0;
  float* part_base = reinterpret_cast<float*>(stg);
  unsigned int target = 0;
  int trace_n = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) g_pd_trace[trace_n++] = globaltimer_ns();
  PdPipe pipe;
  // Tile ownership.  Cell tiles go to CTAs 0, 1, ...; head tiles to CTAs G-1, G-2, ... (the logits tiles first, then
  // the hp tiles): the attention phase hands out crops from CTA 0 upwards, so the top CTAs have the fewest (at batches
  // up to 2048 none), and a CTA never holds a head and a cell tile while 11 m_tiles <= G.
  const int G = (int)gridDim.x;
  const int from_top = G - 1 - (int)blockIdx.x;
  const int first_logits = from_top;
  const int first_hp = (from_top - m_tiles % G + G) % G;
  // The attention phase needs hp_i (head column tiles 0, 1) but not the logits / argmax of step i-1 (tile 2: the
  // slowest tile of the phase, 6.9 us against 3.3 us) - only the cell phase needs the token.  So a CTA ARRIVES at the
  // barrier behind the head phase once its hp tiles are done, works on its logits tiles, and only then waits: nobody
  // waits for a logits tile before the barrier that follows the attention phase.  With early exit the all-crops-done
  // test therefore moves behind that second barrier (the attention phase of the step has then run once more, which
  // only shows in alpha_out: with early exit AND alpha_out the barrier is not split).
  const bool split = !(s.early_exit && s.alpha_out != nullptr);
  for (int i = 0; i <= s.steps; ++i) {
    const int step = i;
    // ---- head_i: hp_i (column tiles 0,1; not at i = steps) and logits_{i-1} (tile 2; nothing to store at i = 0)
    if (i < s.steps)
      pd_gemm_phase<0>(&maps.h[i & 1], &maps.head, &maps.hp, s, step, AT_H / 64, first_hp, 0, 2, m_tiles, sA, sB, stg, full_bar,
                       empty_bar, tfull_bar, tempty_bar, tmem_base, pipe);
    if (split && i < s.steps) pd_barrier_arrive(gbar, target, trace_n);
    if (i > 0)
      pd_gemm_phase<0>(&maps.h[i & 1], &maps.head, &maps.hp, s, step, AT_H / 64, first_logits, 2, 1, m_tiles, sA, sB, stg, full_bar,
                       empty_bar, tfull_bar, tempty_bar, tmem_base, pipe);
    if (i == s.steps) break;
    if (split) pd_barrier_wait(gbar, target, trace_n);
    else pd_grid_barrier(gbar, target, trace_n);
    if (!split && s.early_exit && __ldcg(s.live) == 0) break;   // every crop has emitted [s]: the remaining logits stay zero
    // ---- attention_i over this group's crops, zig-zag over the steps.  Partials: W = 1: a private block per warp;
    // W > 1: a block per group and round parity (the next round may start before a slow warp has read this one)
    if (W == 1 && s.B > GG && i < PD_DYN_STEPS) {
      // more crops than warps: the first round is static (logical index gg), the rest is handed out through a per-step
      // counter, fetched while the current crop is being worked on, so the phase ends within one crop time (~12 us at
      // 4096 crops) of the first idle warp whatever B is.  Logical index l <-> crop l (even steps) / B-1-l (odd steps):
      // the zig-zag described above.  Tried on top of this and measured slower at 4096 crops (25.3 us per step): L2
      // evict-first / no-L1-allocate loads (25.9), one bulk L2 prefetch of the crop's proj + enc block per crop (29.2),
      // an evict_last / evict_first split of the crops so that 65 % of proj + enc stays in L2 (25.9), and the crop's two
      // blocks bulk-copied into shared memory (five 26 KB buffers per SM, one warp each) and scored from there (45.3; 5.4
      // instead of 4.4 us at 512 crops).  None of the memory-side changes helps because the phase is not waiting for memory:
      // a crop is ~1500 dependent-ish warp instructions (208 tanh on the MUFU pipe, 130 shuffle steps, the online
      // softmax), 28 crops per SM per step, and 16 warps per SM is all the latency hiding there is at 128 registers.
      int* counter = s.live + 2 + i;
      float* part = part_base + (size_t)warp * AT_VSUB * AT_PART;
      int l = gg;
      while (l < s.B) {
        const int fetched = lane == 0 ? atomicAdd(counter, 1) : 0;
        at_attention_crop<W>(enc, proj, s, step, (i & 1) ? s.B - 1 - l : l, lane, sub, part, 2 + group);
        l = GG + __shfl_sync(0xffffffffu, fetched, 0);
      }
    } else {
      for (int kk = 0; kk < my_crops; ++kk) {
        const int k = (i & 1) ? (my_crops - 1 - kk) : kk;
        float* part = part_base + (size_t)((W == 1 ? warp : (kk & 1) * GROUPS + group)) * AT_VSUB * AT_PART;
        at_attention_crop<W>(enc, proj, s, step, gg + k * GG, lane, sub, part, 2 + group);
      }
    }
    if (threadIdx.x == 0) { prefetch_tensormap(&maps.xh[i & 1]); prefetch_tensormap(&maps.cat); }
    pd_grid_barrier(gbar, target, trace_n);
    if (split && s.early_exit && __ldcg(s.live) == 0) break;
    // ---- cell_i
    // 128 x 64 tiles while they all fit the grid at once (<= 1152 crops on 148 SMs): the 128 x 128 x 16 MMA of one CTA
    // runs at about half rate (8 KB of shared-memory operands per 64 cycles = the whole port), so a tile's 32 MMAs and
    // its epilogue are the phase; halving the tile halves both and uses twice the SMs.  Same K order: same bits.
    if (16 * m_tiles <= G)
      pd_gemm_phase<1, 64>(&maps.xh[i & 1], &maps.cat64, &maps.hp, s, step, 2 * AT_H / 64, (int)blockIdx.x, 0, 4 * AT_H / 64, m_tiles,
                           sA, sB, stg, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base, pipe);
    else
      pd_gemm_phase<1>(&maps.xh[i & 1], &maps.cat, &maps.hp, s, step, 2 * AT_H / 64, (int)blockIdx.x, 0, 4 * AT_H / AT_BLOCK, m_tiles,
                       sA, sB, stg, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base, pipe);
    pd_grid_barrier(gbar, target, trace_n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static TmaEncoders g_attn_enc;

static inline size_t decode_ws_bytes(int Bg) {
  // xh[2][Bg][2H] bf16 | hp [Bg][H] f32 | c [Bg][H] f32 | tok [Bg] | done [Bg] | live, barrier counter, 256 per-step crop
  // counters; rounded to 256 bytes
  size_t b = (size_t)Bg * AT_H * (2 * 2 * 2 + 4 + 4) + (size_t)Bg * 8 + 16 + 4 * 256;
  return (b + 255) & ~(size_t)255;
}

struct DecodeArgs {
  const __nv_bfloat16* enc; const __nv_bfloat16* proj; const void* w_head; const void* w_cat;
  const float* b_h2h; const float* b_gen; const float* w_score; const float* w_tok; const float* bias;
  const int64_t* text; float* logits; int32_t* ids; int32_t* lengths; float* alpha_out;
  int T, C, steps, early_exit;
};

// Driver selection: -1 = default (persistent), 1 = persistent, 0 = launch chain.  Measured (B200, 26 steps, persistent /
// chain): 256 crops 0.49 / 0.51 ms, 512 crops 0.50 / 0.55, 2048 crops 0.78 / 0.83, 4096 crops 1.28 / 1.40 - the
// persistent kernel wins at every batch size since the logits tiles left the critical path (split barrier) and the
// attention phase hands out crops dynamically; the chain stays as the cross-check.  Results are bit-identical.
static std::atomic<int> g_persistent{-1};
constexpr int PD_MAX_BATCH = 1 << 20;

template <int W>
static int launch_persistent(const PdMaps& maps, const __nv_bfloat16* enc, const __nv_bfloat16* proj, const DecodeState& s,
                             cudaStream_t st) {
  static unsigned long long smem_set_pd = 0;
  int rc = ensure_dynamic_smem((const void*)attn_decode_persistent_kernel<W>, PD_SMEM_BYTES, &smem_set_pd);
  if (rc != MSTR_OK) return rc;
  int per_sm = 0;
  MSTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_decode_persistent_kernel<W>, PD_THREADS, PD_SMEM_BYTES));
  MSTR_REQUIRE(per_sm >= 1, MSTR_ERR_UNSUPPORTED, "attn_decode_persistent_kernel does not fit on an SM");
  cudaLaunchAttribute coop[1];
  coop[0].id = cudaLaunchAttributeCooperative;
  coop[0].val.cooperative = 1;
  cudaLaunchConfig_t c = {};
  c.gridDim = dim3(sm_count()); c.blockDim = dim3(PD_THREADS); c.dynamicSmemBytes = PD_SMEM_BYTES; c.stream = st;
  c.attrs = coop; c.numAttrs = 1;
  MSTR_CUDA(cudaLaunchKernelEx(&c, attn_decode_persistent_kernel<W>, maps, enc, proj, s,
                               reinterpret_cast<unsigned int*>(s.live + 1)));
  MSTR_LAUNCH_CHECK("attn_decode_persistent_kernel");
  return MSTR_OK;
}

// State block in the workspace, then either the persistent launch or the chain of launches on `st`.
static int decode_run(const DecodeArgs& a, uint8_t* ws, int B, cudaStream_t st, bool persistent) {
  DecodeState s;
  const size_t xh_bytes = (size_t)B * 2 * AT_H * 2;
  s.xh[0] = reinterpret_cast<__nv_bfloat16*>(ws);
  s.xh[1] = reinterpret_cast<__nv_bfloat16*>(ws + xh_bytes);
  s.hp = reinterpret_cast<float*>(ws + 2 * xh_bytes);
  s.cst = s.hp + (size_t)B * AT_H;
  s.tok = reinterpret_cast<int32_t*>(s.cst + (size_t)B * AT_H);
  s.done = s.tok + B;
  s.live = s.done + B;
  s.b_h2h = a.b_h2h; s.b_gen = a.b_gen; s.w_score = a.w_score; s.w_tok = a.w_tok; s.bias = a.bias;
  s.text = a.text;
  s.logits = a.logits;
  s.ids = a.ids;
  s.lengths = a.lengths;
  s.alpha_out = a.alpha_out;
  s.B = B; s.T = a.T; s.C = a.C; s.steps = a.steps; s.early_exit = a.early_exit;
  // h_0 = c_0 = 0 (decoders.py:36-38); ctx halves are overwritten before use
  MSTR_CUDA(cudaMemsetAsync(ws, 0, 2 * xh_bytes + (size_t)B * AT_H * 8, st));
  decode_init_kernel<<<ceil_div(B, 256), 256, 0, st>>>(s);
  MSTR_LAUNCH_CHECK("decode_init_kernel");

  // tensor maps: h view [B][256] with row stride 512 elements; [ctx|h] view [B][512]; weights; hp as a TMA-store target
  alignas(64) CUtensorMap tmH[2], tmXH[2], tmHead, tmCat, tmHp;
  int rc;
  for (int p = 0; p < 2; ++p) {
    cuuint64_t dims[2] = {(cuuint64_t)AT_H, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * AT_H * 2)};
    cuuint32_t box[2] = {64, AT_BLOCK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_attn_enc.tiled(&tmH[p], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, s.xh[p] + AT_H, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeTiled(h) failed with CUresult %d", (int)r);
    rc = encode_tiled_2d_bf16(g_attn_enc, &tmXH[p], s.xh[p], 2 * AT_H, B, 64, AT_BLOCK, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != MSTR_OK) return rc;
  }
  rc = encode_tiled_2d_bf16(g_attn_enc, &tmHead, a.w_head, AT_H, 3 * AT_BLOCK, 64, AT_BLOCK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  rc = encode_tiled_2d_bf16(g_attn_enc, &tmCat, a.w_cat, 2 * AT_H, 4 * AT_H, 64, AT_BLOCK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  alignas(64) CUtensorMap tmCat64;
  rc = encode_tiled_2d_bf16(g_attn_enc, &tmCat64, a.w_cat, 2 * AT_H, 4 * AT_H, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  rc = encode_output_2d(g_attn_enc, &tmHp, s.hp, true, (uint64_t)AT_H, (uint64_t)B);
  if (rc != MSTR_OK) return rc;

  const int m_tiles = ceil_div(B, AT_BLOCK);
  const int W = attention_warps_per_crop(B, sm_count());
  if (persistent) {
    PdMaps maps;
    maps.h[0] = tmH[0]; maps.h[1] = tmH[1]; maps.xh[0] = tmXH[0]; maps.xh[1] = tmXH[1]; maps.head = tmHead; maps.cat = tmCat; maps.cat64 = tmCat64;
    maps.hp = tmHp;
    if (W == 4) return launch_persistent<4>(maps, a.enc, a.proj, s, st);
    if (W == 2) return launch_persistent<2>(maps, a.enc, a.proj, s, st);
    return launch_persistent<1>(maps, a.enc, a.proj, s, st);
  }
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl[0].val.programmaticStreamSerializationAllowed = 1;
  auto config = [&](dim3 grid, int threads, size_t smem) {
    cudaLaunchConfig_t c = {};
    c.gridDim = grid; c.blockDim = dim3(threads); c.dynamicSmemBytes = smem; c.stream = st;
    c.attrs = pdl; c.numAttrs = 1;
    return c;
  };
  for (int i = 0; i <= a.steps; ++i) {
    // head_i: hp_i (tiles 0,1) and logits_{i-1} (tile 2); the last iteration only needs the logits tile
    {
      cudaLaunchConfig_t c = config(dim3(3, m_tiles), AT_THREADS, AT_SMEM_BYTES);
      MSTR_CUDA(cudaLaunchKernelEx(&c, rnn_gemm_kernel<0>, tmH[i & 1], tmHead, tmHp, s, i, AT_H / 64));
      MSTR_LAUNCH_CHECK("rnn_gemm_kernel<head>");
    }
    if (i == a.steps) break;
    {
      cudaLaunchConfig_t c = config(dim3(ceil_div(B, AT_CTX_WARPS / W)), 32 * AT_CTX_WARPS, 0);
      if (W == 4) MSTR_CUDA(cudaLaunchKernelEx(&c, attn_context_kernel<4>, a.enc, a.proj, s, i));
      else if (W == 2) MSTR_CUDA(cudaLaunchKernelEx(&c, attn_context_kernel<2>, a.enc, a.proj, s, i));
      else MSTR_CUDA(cudaLaunchKernelEx(&c, attn_context_kernel<1>, a.enc, a.proj, s, i));
      MSTR_LAUNCH_CHECK("attn_context_kernel");
    }
    {
      cudaLaunchConfig_t c = config(dim3(4 * AT_H / AT_BLOCK, m_tiles), AT_THREADS, AT_SMEM_BYTES);
      MSTR_CUDA(cudaLaunchKernelEx(&c, rnn_gemm_kernel<1>, tmXH[i & 1], tmCat, tmHp, s, i, 2 * AT_H / 64));
      MSTR_LAUNCH_CHECK("rnn_gemm_kernel<cell>");
    }
  }
  return MSTR_OK;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_debug_decode_trace(unsigned long long* out_host, int n,
                                                                              unsigned long long* per_cta_host, int n_cta) {
  using namespace mstr;
  MSTR_REQUIRE(out_host && n > 0 && n <= PD_TRACE_MAX, MSTR_ERR_BAD_ARG, "mstr_debug_decode_trace: bad argument");
  MSTR_CUDA(cudaMemcpyFromSymbol(out_host, g_pd_trace, sizeof(unsigned long long) * (size_t)n));
  if (per_cta_host && n_cta == 3 * 4 * 16) {
    MSTR_CUDA(cudaMemcpyFromSymbol(per_cta_host, g_pd_fine, sizeof(unsigned long long) * (size_t)n_cta));
  } else if (per_cta_host) {
    MSTR_REQUIRE(n_cta == PD_TRACE_BARRIERS * PD_TRACE_CTAS, MSTR_ERR_BAD_ARG, "mstr_debug_decode_trace: per-CTA buffer must hold %d values",
                 PD_TRACE_BARRIERS * PD_TRACE_CTAS);
    MSTR_CUDA(cudaMemcpyFromSymbol(per_cta_host, g_pd_trace_cta, sizeof(unsigned long long) * (size_t)n_cta));
  }
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_set_decode_persistent(int on) {
  mstr::g_persistent.store(on < 0 ? -1 : (on ? 1 : 0));
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int64_t mstr_attn_decode_bf16_workspace_bytes(int B, int Hd) {
  (void)Hd;
  using namespace mstr;
  return (int64_t)decode_ws_bytes((B + AT_BLOCK - 1) / AT_BLOCK * AT_BLOCK);
}

extern "C" __attribute__((visibility("default"))) int mstr_attn_decode_bf16(
    const void* enc_bf16, const void* proj_bf16, const void* w_head, const float* b_h2h, const float* b_gen,
    const float* w_score, const void* w_cat, const float* w_tok, const float* bias, const int64_t* text,
    void* workspace, float* logits, int32_t* ids, int32_t* lengths, float* alpha_out, int B, int T, int Hd, int C,
    int steps, int early_exit, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(enc_bf16 && proj_bf16 && w_head && b_h2h && b_gen && w_score && w_cat && w_tok && workspace && logits && ids,
               MSTR_ERR_BAD_ARG, "mstr_attn_decode_bf16: null pointer");
  MSTR_REQUIRE(B > 0 && T > 0 && steps > 0 && C > 0, MSTR_ERR_BAD_ARG, "mstr_attn_decode_bf16: bad size");
  MSTR_REQUIRE(Hd == AT_H, MSTR_ERR_UNSUPPORTED, "mstr_attn_decode_bf16: hidden size %d != %d", Hd, AT_H);
  MSTR_REQUIRE(C <= AT_BLOCK, MSTR_ERR_UNSUPPORTED, "mstr_attn_decode_bf16: num_classes %d > %d", C, AT_BLOCK);
  MSTR_REQUIRE(T <= 96, MSTR_ERR_UNSUPPORTED, "mstr_attn_decode_bf16: T=%d > 96", T);
  MSTR_REQUIRE(aligned16(enc_bf16) && aligned16(proj_bf16) && aligned16(w_head) && aligned16(w_cat) &&
                   aligned16(w_tok) && aligned16(bias) && aligned16(workspace) && aligned16(w_score) && aligned16(b_h2h),
               MSTR_ERR_BAD_ARG, "mstr_attn_decode_bf16: pointers must be 16-byte aligned");
  if (text) early_exit = 0;   // the teacher-forced branch always runs every step (decoders.py:41-50)
  int rc = load_tma_encoders(g_attn_enc);
  if (rc != MSTR_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  static unsigned long long smem_set_head = 0, smem_set_cell = 0;
  rc = ensure_dynamic_smem((const void*)rnn_gemm_kernel<0>, AT_SMEM_BYTES, &smem_set_head);
  if (rc != MSTR_OK) return rc;
  rc = ensure_dynamic_smem((const void*)rnn_gemm_kernel<1>, AT_SMEM_BYTES, &smem_set_cell);
  if (rc != MSTR_OK) return rc;

  DecodeArgs a{static_cast<const __nv_bfloat16*>(enc_bf16), static_cast<const __nv_bfloat16*>(proj_bf16), w_head, w_cat,
               b_h2h, b_gen, w_score, w_tok, bias, text, logits, ids, lengths, alpha_out, T, C, steps, early_exit};
  const int mode = g_persistent.load(std::memory_order_relaxed);
  return decode_run(a, static_cast<uint8_t*>(workspace), B, st, mode < 0 ? B <= PD_MAX_BATCH : mode != 0);
}
