// Attention decoder for the bf16 throughput mode: the three GEMMs of every decode step run on the
// tensor cores, the Bahdanau attention on all SMs, and the whole greedy loop is enqueued by ONE
// C-ABI call (mstr_attn_decode_bf16) as three launches per step:
//
//   head_i   [hp_i | logits_{i-1}] = h_i [W_h2h ; W_gen]^T      tcgen05 GEMM, N = 384 (3 tiles of 128)
//            epilogue: hp_i += b_h2h (f32, for the attention kernel); logits_{i-1} += b_gen -> stored,
//            thread-local argmax over the 97 classes (TMEM lane = crop) -> id, next token, [s] bookkeeping
//   ctx_i    e_t = w_score . tanh(proj[b,t,:] + hp_i[b,:]), alpha = softmax_T(e), ctx = sum_t alpha_t enc[b,t,:]
//            one warp per crop, 16-byte bf16 loads, shuffle reductions, tanh.approx; writes ctx (bf16) into
//            the left half of the [ctx | h] operand buffer
//   cell_i   gates = [ctx_i | h_i] [W_ctx | W_hh]^T + W_tok[token_i] + bias -> (h_{i+1}, c_{i+1})
//            tcgen05 GEMM, N = 1024 with rows packed (chunk of 32 units, gate, unit) so that one
//            128-column tile carries i,f,g,o of the same units and the LSTM cell is thread-local in the
//            epilogue; c stays f32 in global memory (L2 resident), h_{i+1} is written as bf16 into the
//            right half of the other [ctx | h] buffer
//
// Replaces modules/decoders.py:26-64 (Attention.forward greedy :52-62 and teacher-forced :41-50
// branches), :79-101 (AttentionCell.forward), :17/:59 (generator), :19-24 (_char_to_onehot: a one-hot
// times W_ih is a row gather of W_tok).  proj = i2h(enc) is loop invariant (:81) and computed once by
// the caller.  Differences to the fp32 SIMT kernel (attn_decode.cu): proj / enc / h / ctx are bf16
// operands (fp32 accumulation everywhere), tanh.approx in the attention scores.
#include <atomic>

#include "tc_common.cuh"

namespace mstr {

constexpr int AT_H = 256;
constexpr int AT_BLOCK = 128;      // BLOCK_M = BLOCK_N
constexpr int AT_STAGES = 3;
constexpr int AT_TILE_BYTES = 128 * 128;   // [128 rows][64 k] bf16, 128B swizzle
constexpr int AT_THREADS = 192;
constexpr int AT_SMEM_BYTES = 1024 + AT_STAGES * 2 * AT_TILE_BYTES + 256;

struct DecodeState {
  // operand / state buffers (device)
  __nv_bfloat16* xh[2];   // [B][512]: cols 0..255 ctx_i, cols 256..511 h_i   (double buffered by step parity)
  float* hp;              // [B][256]
  float* cst;             // [256][B] (transposed: the cell epilogue's lane = crop accesses are coalesced)
  int32_t* tok;           // [B] token fed to the current step
  int32_t* done;          // [B] crop has emitted [s]
  int32_t* live;          // [1] number of crops still decoding (early exit)
  // parameters
  const float* b_h2h; const float* b_gen; const float* w_score; const float* w_tok; const float* bias;
  const int64_t* text;    // optional teacher forcing [B][steps]
  // outputs
  float* logits; int32_t* ids; int32_t* lengths; float* alpha_out;
  int B, T, C, steps, early_exit;
};

// Programmatic dependent launch: every kernel of the decode chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be scheduled (and run their prologue:
// barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel drains; griddep_wait() blocks
// until the previous grid has completed and its writes are visible, and must precede every dependent access.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float at_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float at_tanh(float x) { return fmaf(2.f, at_sigmoid(2.f * x), -1.f); }

// EPI 0: head (hp + logits/argmax), EPI 1: LSTM cell.  One 128x128 output tile per CTA; K = num_k * 64.
template <int EPI>
__global__ void __launch_bounds__(AT_THREADS, 2)
rnn_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const DecodeState s, const int step, const int num_k) {
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
    // ===================== epilogue: TMEM lane = crop row =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int crop = m0 + row;
    const bool valid = crop < s.B;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if constexpr (EPI == 0) {
      // The accumulator layout gives every thread one crop ROW, and rows are 1 KB (hp) / 10 KB (logits) apart in
      // global memory: a direct store costs 32 L1 wavefronts per instruction (an elimination run priced these
      // stores at 0.35 of the decoder's 1.9 ms).  The tile is therefore transposed through the (now idle)
      // operand stages: row-wise writes with an odd stride (conflict free), then each warp copies whole rows
      // with 128-byte coalesced stores.
      float* stg = reinterpret_cast<float*>(smem);
      const int ew = warp - 2;               // 0..3
      if (n_tile < 2) {
        // hp_i = h_i W_h2h^T + b_h2h  (consumed by attn_context_kernel of this step)
        constexpr int LD = AT_BLOCK + 1;
#pragma unroll 1
        for (int c0 = 0; c0 < AT_BLOCK; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[row * LD + c0 + j] = __uint_as_float(v[j]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float bias4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bias4[q] = __ldg(s.b_h2h + n_tile * AT_BLOCK + q * 32 + lane);
#pragma unroll 4
        for (int r = ew * 32; r < ew * 32 + 32; ++r) {
          if (m0 + r >= s.B) break;
          float* dst = s.hp + (size_t)(m0 + r) * AT_H + n_tile * AT_BLOCK;
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q * 32 + lane] = stg[r * LD + q * 32 + lane] + bias4[q];
        }
      } else if (step > 0) {
        // logits of decode step (step-1) = h_step W_gen^T + b_gen; argmax feeds the token of `step`
        constexpr int LD = 101;              // >= num_classes rounded up to odd; C <= 128 is checked by the host
        const int ld = (s.C | 1) > LD ? (s.C | 1) : LD;
        int* skip = reinterpret_cast<int*>(stg + AT_BLOCK * 129);   // per row: do not store (crop finished / out of range)
        const int ls = step - 1;
        const bool was_done = valid && s.early_exit && s.done[crop];
        float best = -INFINITY;
        int besti = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < AT_BLOCK; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int cls = c0 + j;
            if (cls < s.C) {
              const float y = __uint_as_float(v[j]) + __ldg(s.b_gen + cls);
              stg[row * ld + cls] = y;
              if (y > best) { best = y; besti = cls; }      // ascending: the lowest index wins ties
            }
          }
        }
        skip[row] = (!valid || was_done) ? 1 : 0;
        if (valid) {
          if (was_done) {
            s.ids[(size_t)crop * s.steps + ls] = 0;          // argmax of the zero-initialised logits row
          } else {
            s.ids[(size_t)crop * s.steps + ls] = besti;
            if (besti == 1 && !s.done[crop]) {
              if (s.lengths && s.lengths[crop] == s.steps) s.lengths[crop] = ls;
              if (s.early_exit) {
                s.done[crop] = 1;
                atomicSub(s.live, 1);
              }
            }
          }
          if (step < s.steps) s.tok[crop] = s.text ? (int)s.text[(size_t)crop * s.steps + step] : besti;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 2
        for (int r = ew * 32; r < ew * 32 + 32; ++r) {
          if (skip[r]) continue;
          float* lrow = s.logits + ((size_t)(m0 + r) * s.steps + ls) * s.C;
          for (int cls = lane; cls < s.C; cls += 32) lrow[cls] = stg[r * ld + cls];
        }
      }
    } else {
      // LSTM cell: this tile = 32 hidden units x 4 gates of 128 crops
      const int unit0 = n_tile * 32;
      const int tok = valid ? s.tok[crop] : 0;
      const float* wt = s.w_tok + (size_t)tok * (4 * AT_H);
      float* ccol = s.cst + (size_t)unit0 * s.B + (valid ? crop : 0);   // cell state is kept [H][B]: lane = crop is coalesced
      __nv_bfloat16* hrow = s.xh[(step + 1) & 1] + (size_t)(valid ? crop : 0) * (2 * AT_H) + AT_H + unit0;
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub) {            // 8 units at a time
        uint32_t a[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld8(taddr + (uint32_t)(g * 32 + sub * 8), a[g]);
        tmem_ld_wait();
        if (valid) {
          float add[4][8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = g * AT_H + unit0 + sub * 8;
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(wt + col));
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(wt + col) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(s.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(s.bias + col) + 1);
            add[g][0] = t0.x + b0.x; add[g][1] = t0.y + b0.y; add[g][2] = t0.z + b0.z; add[g][3] = t0.w + b0.w;
            add[g][4] = t1.x + b1.x; add[g][5] = t1.y + b1.y; add[g][6] = t1.z + b1.z; add[g][7] = t1.w + b1.w;
          }
          float cp[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) cp[j] = ccol[(size_t)(sub * 8 + j) * s.B];
          float hv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gi = __uint_as_float(a[0][j]) + add[0][j];
            const float gf = __uint_as_float(a[1][j]) + add[1][j];
            const float gg = __uint_as_float(a[2][j]) + add[2][j];
            const float go = __uint_as_float(a[3][j]) + add[3][j];
            const float cn = fmaf(at_sigmoid(gf), cp[j], at_sigmoid(gi) * at_tanh(gg));
            cp[j] = cn;
            hv[j] = at_sigmoid(go) * at_tanh(cn);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) ccol[(size_t)(sub * 8 + j) * s.B] = cp[j];
          *reinterpret_cast<uint4*>(hrow + sub * 8) =
              make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]),
                         pack_bf16x2(hv[6], hv[7]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// Bahdanau attention of one decode step: one warp per crop, lane = 8 channels.
// Single pass with an online softmax: every iteration loads four time steps of BOTH proj and enc (eight 16-byte
// loads in flight per lane), scores them and folds them into a running (max, denominator, context).  The earlier
// two-pass version (scores over all of proj, softmax, then enc) was a chain of 14 dependent memory round trips per
// warp; with all 4096 warps resident at once that chain IS the kernel time (26 us at 52 % of the HBM peak).
// 4 warps per block, 7 blocks per SM (72 registers): 28 warps x 148 SMs hold all 4096 crops of the bench batch at once.
constexpr int AT_CTX_WARPS = 4;
__global__ void __launch_bounds__(32 * AT_CTX_WARPS, 7)
attn_context_kernel(const __nv_bfloat16* __restrict__ enc, const __nv_bfloat16* __restrict__ proj,
                    const DecodeState s, const int step) {
  griddep_wait();
  griddep_launch();
  if (s.early_exit && *s.live == 0) return;
  const int lane = threadIdx.x & 31;
  const int crop = blockIdx.x * AT_CTX_WARPS + (threadIdx.x >> 5);
  if (crop >= s.B) return;
  const int T = s.T;
  float hp[8], ws[8];
  {
    const float4* hp4 = reinterpret_cast<const float4*>(s.hp + (size_t)crop * AT_H + lane * 8);
    const float4* ws4 = reinterpret_cast<const float4*>(s.w_score + lane * 8);
    const float4 a = hp4[0], b = hp4[1], c = __ldg(ws4), d = __ldg(ws4 + 1);
    hp[0] = a.x; hp[1] = a.y; hp[2] = a.z; hp[3] = a.w; hp[4] = b.x; hp[5] = b.y; hp[6] = b.z; hp[7] = b.w;
    ws[0] = c.x; ws[1] = c.y; ws[2] = c.z; ws[3] = c.w; ws[4] = d.x; ws[5] = d.y; ws[6] = d.z; ws[7] = d.w;
  }
  const uint4* pj = reinterpret_cast<const uint4*>(proj + (size_t)crop * T * AT_H) + lane;   // 32 uint4 per time step
  const uint4* en = reinterpret_cast<const uint4*>(enc + (size_t)crop * T * AT_H) + lane;
  constexpr int MAX_SLOTS = 3;   // T <= 96
  float e_slot[MAX_SLOTS];       // score of time step t lives in lane t % 32, slot t / 32 (for alpha_out)
#pragma unroll
  for (int k = 0; k < MAX_SLOTS; ++k) e_slot[k] = -INFINITY;
  float mx = -INFINITY, den = 0.f;
  float cx[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) cx[q] = 0.f;
  for (int t0 = 0; t0 < T; t0 += 4) {
    uint4 vp[4], ve[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in = t0 + u < T;
      vp[u] = in ? __ldg(pj + (size_t)(t0 + u) * 32) : make_uint4(0u, 0u, 0u, 0u);
      ve[u] = in ? __ldg(en + (size_t)(t0 + u) * 32) : make_uint4(0u, 0u, 0u, 0u);
    }
    float e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w4[4] = {vp[u].x, vp[u].y, vp[u].z, vp[u].w};
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16x2(w4[q]);
        part = fmaf(ws[2 * q], tanh_approx(f.x + hp[2 * q]), part);
        part = fmaf(ws[2 * q + 1], tanh_approx(f.y + hp[2 * q + 1]), part);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const int t = t0 + u;
      e[u] = t < T ? part : -INFINITY;
      if (t < T && (t & 31) == lane) {
#pragma unroll
        for (int k = 0; k < MAX_SLOTS; ++k)
          if ((t >> 5) == k) e_slot[k] = part;
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
  const float inv = 1.f / den;
  if (s.alpha_out) {
#pragma unroll
    for (int k = 0; k < MAX_SLOTS; ++k) {
      const int t = k * 32 + lane;
      if (t < T) s.alpha_out[((size_t)crop * s.steps + step) * T + t] = __expf(e_slot[k] - mx) * inv;
    }
  }
  __nv_bfloat16* dst = s.xh[step & 1] + (size_t)crop * (2 * AT_H) + lane * 8;
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(cx[0] * inv, cx[1] * inv), pack_bf16x2(cx[2] * inv, cx[3] * inv),
                                              pack_bf16x2(cx[4] * inv, cx[5] * inv), pack_bf16x2(cx[6] * inv, cx[7] * inv));
}

// =====================================================================================================
// PERSISTENT decoder: ONE cooperative launch runs all steps.
//
// The launch chain above pays, per step, three launch latencies, three prologues (barrier init, TMEM allocation,
// tensor-map prefetch) and three pipeline drains: at 256 crops the decoder took 0.83 ms for ~0.1 ms of work.  Here one
// CTA per SM stays resident for the whole decode; the three phases of a step - head GEMM tiles, attention over crops,
// cell GEMM tiles - are distributed over the grid and separated by a grid-wide barrier (one atomic per CTA on an L2
// word, ~1 us).  TMEM, the mbarriers and the tensor maps are set up once.  The phase bodies are the ones of the chain
// kernels (same tiles, same MMA order, same epilogue arithmetic), so the results are bit-identical to the chain.
//
// What changes with the persistence:
//  * buffers that one CTA writes and another reads in a later phase (hp, [ctx|h], c, tok, done, live, lengths) are
//    read with ld.global.cg (L2; the L1 of the reading SM may still hold the line of the previous step);
//  * the [ctx|h] operand is written with ordinary stores and read by TMA (async proxy) in the next phase: writers
//    execute fence.proxy.async.global before the barrier, the TMA-issuing thread after it;
//  * the attention phase walks its crops in alternating direction (even steps ascending, odd steps descending): every
//    step touches all of proj+enc (109 MB at 4096 crops, about the L2 capacity), and with a fixed direction an LRU
//    cache holds exactly the lines that are needed LAST; zig-zag reuse starts with the most recently used lines.
// =====================================================================================================
constexpr int PD_WARPS = 16;
constexpr int PD_THREADS = 32 * PD_WARPS;
constexpr int PD_STG_BYTES = AT_BLOCK * (AT_BLOCK + 1) * 4 + 512;      // transposing stage of the head epilogue + skip flags
constexpr int PD_SMEM_BYTES = 1024 + AT_STAGES * 2 * AT_TILE_BYTES + PD_STG_BYTES + 256;

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Phase trace of CTA 0 (diagnostic, mstr_debug_decode_trace): %globaltimer at every barrier entry (this CTA's phase
// work is done) and exit (every CTA's is), so phase work, barrier wait and barrier latency can be told apart.
constexpr int PD_TRACE_MAX = 512;
__device__ unsigned long long g_pd_trace[PD_TRACE_MAX];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Grid-wide barrier (all CTAs are co-resident: cooperative launch).  `target` is this CTA's running arrival target.
__device__ __forceinline__ void pd_grid_barrier(unsigned int* counter, unsigned int& target, int& trace_n) {
  fence_proxy_async_global();              // this thread's global stores -> later async-proxy (TMA) reads of any CTA
  __syncthreads();
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0 && trace_n < PD_TRACE_MAX) g_pd_trace[trace_n++] = globaltimer_ns();
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (ld_acquire_gpu(counter) < target) {
      if (clock64() - t0 > 4000000000ll) {
        printf("mstr: grid barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
    fence_proxy_async_global();
    if (blockIdx.x == 0 && trace_n < PD_TRACE_MAX) g_pd_trace[trace_n++] = globaltimer_ns();
  }
  __syncthreads();
}

struct PdPipe {            // running ring / accumulator state of one role, carried across tiles, phases and steps
  int stage = 0;
  uint32_t phase = 0;
  uint32_t acc_phase = 0;
};

// One GEMM phase over this CTA's tiles.  EPI 0: head, n_tiles_n column tiles starting at n_first; EPI 1: cell.
template <int EPI>
__device__ __forceinline__ void pd_gemm_phase(const CUtensorMap* tmA, const CUtensorMap* tmB, const DecodeState& s,
                                              const int step, const int num_k, const int n_first, const int n_count,
                                              const int m_tiles, uint8_t* sA, uint8_t* sB, float* stg, uint64_t* full_bar,
                                              uint64_t* empty_bar, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                              const uint32_t tmem_base, PdPipe& pipe) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = n_count * m_tiles;
  if (warp == 0) {
    if (lane == 0) {
      fence_proxy_async_global();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = n_first + tile % n_count, m0 = (tile / n_count) * AT_BLOCK;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[pipe.stage], pipe.phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[pipe.stage], 2 * AT_TILE_BYTES);
          tma_load_2d(sA + pipe.stage * AT_TILE_BYTES, tmA, &full_bar[pipe.stage], kb * 64, m0);
          tma_load_2d(sB + pipe.stage * AT_TILE_BYTES, tmB, &full_bar[pipe.stage], kb * 64, n_tile * AT_BLOCK);
          if (++pipe.stage == AT_STAGES) { pipe.stage = 0; pipe.phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(AT_BLOCK, AT_BLOCK);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar, pipe.acc_phase ^ 1);       // the epilogue of the previous tile has drained the accumulator
        tc_fence_after();
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[pipe.stage], pipe.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + pipe.stage * AT_TILE_BYTES);
          const uint32_t b_addr = smem_u32(sB + pipe.stage * AT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, make_smem_desc<128>(a_addr + k * 32), make_smem_desc<128>(b_addr + k * 32), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[pipe.stage]);
          if (kb == num_k - 1) umma_commit(tfull_bar);
          if (++pipe.stage == AT_STAGES) { pipe.stage = 0; pipe.phase ^= 1; }
        }
        pipe.acc_phase ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: TMEM lane = crop row (warps 4..7 = lane quarters 0..3) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = n_first + tile % n_count, m0 = (tile / n_count) * AT_BLOCK;
      const int crop = m0 + row;
      const bool valid = crop < s.B;
      mbar_wait(tfull_bar, pipe.acc_phase);
      tc_fence_after();
      if constexpr (EPI == 0) {
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the previous tile's rows have left the stage
        if (n_tile < 2) {
          constexpr int LD = AT_BLOCK + 1;
#pragma unroll 1
          for (int c0 = 0; c0 < AT_BLOCK; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[row * LD + c0 + j] = __uint_as_float(v[j]);
          }
          tc_fence_before();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (lane == 0) mbar_arrive(tempty_bar);
          float bias4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) bias4[q] = __ldg(s.b_h2h + n_tile * AT_BLOCK + q * 32 + lane);
#pragma unroll 4
          for (int r = quarter * 32; r < quarter * 32 + 32; ++r) {
            if (m0 + r >= s.B) break;
            float* dst = s.hp + (size_t)(m0 + r) * AT_H + n_tile * AT_BLOCK;
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q * 32 + lane] = stg[r * LD + q * 32 + lane] + bias4[q];
          }
        } else {
          constexpr int LD = 101;
          const int ld = (s.C | 1) > LD ? (s.C | 1) : LD;
          int* skip = reinterpret_cast<int*>(stg + AT_BLOCK * 129);
          const int ls = step - 1;
          const bool is_done = valid && __ldcg(s.done + crop) != 0;
          const bool was_done = s.early_exit && is_done;
          float best = -INFINITY;
          int besti = 0;
#pragma unroll 1
          for (int c0 = 0; c0 < AT_BLOCK; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cls = c0 + j;
              if (cls < s.C) {
                const float y = __uint_as_float(v[j]) + __ldg(s.b_gen + cls);
                stg[row * ld + cls] = y;
                if (y > best) { best = y; besti = cls; }      // ascending: the lowest index wins ties
              }
            }
          }
          tc_fence_before();
          skip[row] = (!valid || was_done) ? 1 : 0;
          if (valid) {
            if (was_done) {
              s.ids[(size_t)crop * s.steps + ls] = 0;
            } else {
              s.ids[(size_t)crop * s.steps + ls] = besti;
              if (besti == 1 && !is_done) {
                if (s.lengths && __ldcg(s.lengths + crop) == s.steps) s.lengths[crop] = ls;
                if (s.early_exit) {
                  s.done[crop] = 1;
                  atomicSub(s.live, 1);
                }
              }
            }
            if (step < s.steps) s.tok[crop] = s.text ? (int)s.text[(size_t)crop * s.steps + step] : besti;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (lane == 0) mbar_arrive(tempty_bar);
#pragma unroll 2
          for (int r = quarter * 32; r < quarter * 32 + 32; ++r) {
            if (skip[r]) continue;
            float* lrow = s.logits + ((size_t)(m0 + r) * s.steps + ls) * s.C;
            for (int cls = lane; cls < s.C; cls += 32) lrow[cls] = stg[r * ld + cls];
          }
        }
      } else {
        const int unit0 = n_tile * 32;
        const int tok = valid ? __ldcg(s.tok + crop) : 0;
        const float* wt = s.w_tok + (size_t)tok * (4 * AT_H);
        float* ccol = s.cst + (size_t)unit0 * s.B + (valid ? crop : 0);
        __nv_bfloat16* hrow = s.xh[(step + 1) & 1] + (size_t)(valid ? crop : 0) * (2 * AT_H) + AT_H + unit0;
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {
          uint32_t a[4][8];
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld8(taddr + (uint32_t)(g * 32 + sub * 8), a[g]);
          tmem_ld_wait();
          if (sub == 3) {                                   // the accumulator is in registers: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
          }
          if (valid) {
            float add[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = g * AT_H + unit0 + sub * 8;
              const float4 t0 = __ldg(reinterpret_cast<const float4*>(wt + col));
              const float4 t1 = __ldg(reinterpret_cast<const float4*>(wt + col) + 1);
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(s.bias + col));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(s.bias + col) + 1);
              add[g][0] = t0.x + b0.x; add[g][1] = t0.y + b0.y; add[g][2] = t0.z + b0.z; add[g][3] = t0.w + b0.w;
              add[g][4] = t1.x + b1.x; add[g][5] = t1.y + b1.y; add[g][6] = t1.z + b1.z; add[g][7] = t1.w + b1.w;
            }
            float cp[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cp[j] = __ldcg(ccol + (size_t)(sub * 8 + j) * s.B);
            float hv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float gi = __uint_as_float(a[0][j]) + add[0][j];
              const float gf = __uint_as_float(a[1][j]) + add[1][j];
              const float gg = __uint_as_float(a[2][j]) + add[2][j];
              const float go = __uint_as_float(a[3][j]) + add[3][j];
              const float cn = fmaf(at_sigmoid(gf), cp[j], at_sigmoid(gi) * at_tanh(gg));
              cp[j] = cn;
              hv[j] = at_sigmoid(go) * at_tanh(cn);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) ccol[(size_t)(sub * 8 + j) * s.B] = cp[j];
            *reinterpret_cast<uint4*>(hrow + sub * 8) =
                make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]),
                           pack_bf16x2(hv[6], hv[7]));
          }
        }
      }
      pipe.acc_phase ^= 1;
    }
  }
  __syncwarp();
}

// Attention of one crop by one warp (lane = 8 channels): the body of attn_context_kernel.
__device__ __forceinline__ void pd_attention_crop(const __nv_bfloat16* __restrict__ enc, const __nv_bfloat16* __restrict__ proj,
                                                  const DecodeState& s, const int step, const int crop, const int lane) {
  const int T = s.T;
  float hp[8], ws[8];
  {
    const float4* hp4 = reinterpret_cast<const float4*>(s.hp + (size_t)crop * AT_H + lane * 8);
    const float4* ws4 = reinterpret_cast<const float4*>(s.w_score + lane * 8);
    const float4 a = __ldcg(hp4), b = __ldcg(hp4 + 1), c = __ldg(ws4), d = __ldg(ws4 + 1);
    hp[0] = a.x; hp[1] = a.y; hp[2] = a.z; hp[3] = a.w; hp[4] = b.x; hp[5] = b.y; hp[6] = b.z; hp[7] = b.w;
    ws[0] = c.x; ws[1] = c.y; ws[2] = c.z; ws[3] = c.w; ws[4] = d.x; ws[5] = d.y; ws[6] = d.z; ws[7] = d.w;
  }
  const uint4* pj = reinterpret_cast<const uint4*>(proj + (size_t)crop * T * AT_H) + lane;
  const uint4* en = reinterpret_cast<const uint4*>(enc + (size_t)crop * T * AT_H) + lane;
  constexpr int MAX_SLOTS = 3;   // T <= 96
  float e_slot[MAX_SLOTS];
#pragma unroll
  for (int k = 0; k < MAX_SLOTS; ++k) e_slot[k] = -INFINITY;
  float mx = -INFINITY, den = 0.f;
  float cx[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) cx[q] = 0.f;
  for (int t0 = 0; t0 < T; t0 += 4) {
    uint4 vp[4], ve[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in = t0 + u < T;
      vp[u] = in ? __ldg(pj + (size_t)(t0 + u) * 32) : make_uint4(0u, 0u, 0u, 0u);
      ve[u] = in ? __ldg(en + (size_t)(t0 + u) * 32) : make_uint4(0u, 0u, 0u, 0u);
    }
    float e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w4[4] = {vp[u].x, vp[u].y, vp[u].z, vp[u].w};
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16x2(w4[q]);
        part = fmaf(ws[2 * q], tanh_approx(f.x + hp[2 * q]), part);
        part = fmaf(ws[2 * q + 1], tanh_approx(f.y + hp[2 * q + 1]), part);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const int t = t0 + u;
      e[u] = t < T ? part : -INFINITY;
      if (t < T && (t & 31) == lane) {
#pragma unroll
        for (int k = 0; k < MAX_SLOTS; ++k)
          if ((t >> 5) == k) e_slot[k] = part;
      }
    }
    const float m_new = fmaxf(fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])), mx);
    const float rescale = __expf(mx - m_new);
    den *= rescale;
#pragma unroll
    for (int q = 0; q < 8; ++q) cx[q] *= rescale;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float a = __expf(e[u] - m_new);
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
  const float inv = 1.f / den;
  if (s.alpha_out) {
#pragma unroll
    for (int k = 0; k < MAX_SLOTS; ++k) {
      const int t = k * 32 + lane;
      if (t < T) s.alpha_out[((size_t)crop * s.steps + step) * T + t] = __expf(e_slot[k] - mx) * inv;
    }
  }
  __nv_bfloat16* dst = s.xh[step & 1] + (size_t)crop * (2 * AT_H) + lane * 8;
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(cx[0] * inv, cx[1] * inv), pack_bf16x2(cx[2] * inv, cx[3] * inv),
                                              pack_bf16x2(cx[4] * inv, cx[5] * inv), pack_bf16x2(cx[6] * inv, cx[7] * inv));
}

struct PdMaps {
  CUtensorMap h[2], xh[2], head, cat;
};

__global__ void __launch_bounds__(PD_THREADS, 1)
attn_decode_persistent_kernel(const __grid_constant__ PdMaps maps, const __nv_bfloat16* __restrict__ enc,
                              const __nv_bfloat16* __restrict__ proj, const DecodeState s, unsigned int* gbar) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + AT_STAGES * AT_TILE_BYTES;
  float* stg = reinterpret_cast<float*>(smem + 2 * AT_STAGES * AT_TILE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * AT_STAGES * AT_TILE_BYTES + PD_STG_BYTES);
  uint64_t* empty_bar = full_bar + AT_STAGES;
  uint64_t* tfull_bar = empty_bar + AT_STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.h[0]); prefetch_tensormap(&maps.h[1]);
    prefetch_tensormap(&maps.xh[0]); prefetch_tensormap(&maps.xh[1]);
    prefetch_tensormap(&maps.head); prefetch_tensormap(&maps.cat);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);                              // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (s.B + AT_BLOCK - 1) / AT_BLOCK;
  const int gw = blockIdx.x * PD_WARPS + warp, GW = gridDim.x * PD_WARPS;
  const int my_crops = gw < s.B ? (s.B - gw + GW - 1) / GW : 0;
  unsigned int target = 0;
  int trace_n = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) g_pd_trace[trace_n++] = globaltimer_ns();
  PdPipe pipe;
  for (int i = 0; i <= s.steps; ++i) {
    // ---- head_i: hp_i (column tiles 0,1) and logits_{i-1} (tile 2; nothing to store at i = 0, only tile 2 at i = steps)
    const int n_first = (i == s.steps) ? 2 : 0;
    const int n_count = (i == s.steps) ? 1 : (i == 0 ? 2 : 3);
    pd_gemm_phase<0>(&maps.h[i & 1], &maps.head, s, i, AT_H / 64, n_first, n_count, m_tiles, sA, sB, stg, full_bar,
                     empty_bar, tfull_bar, tempty_bar, tmem_base, pipe);
    if (i == s.steps) break;
    pd_grid_barrier(gbar, target, trace_n);
    if (s.early_exit && __ldcg(s.live) == 0) break;        // every crop has emitted [s]: the remaining logits stay zero
    // ---- attention_i over this warp's crops, zig-zag over the steps
    for (int kk = 0; kk < my_crops; ++kk) {
      const int k = (i & 1) ? (my_crops - 1 - kk) : kk;
      pd_attention_crop(enc, proj, s, i, gw + k * GW, lane);
    }
    pd_grid_barrier(gbar, target, trace_n);
    // ---- cell_i
    pd_gemm_phase<1>(&maps.xh[i & 1], &maps.cat, s, i, 2 * AT_H / 64, 0, 4 * AT_H / AT_BLOCK, m_tiles, sA, sB, stg, full_bar,
                     empty_bar, tfull_bar, tempty_bar, tmem_base, pipe);
    pd_grid_barrier(gbar, target, trace_n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// token_0 = [GO] (0) or text[:, 0]; lengths = steps; done = 0; live = B; ids zero-filled.
__global__ void decode_init_kernel(const DecodeState s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    *s.live = s.B;
    s.live[1] = 0;          // arrival counter of the persistent kernel's grid barrier
  }
  if (i < s.B) {
    s.tok[i] = s.text ? (int)s.text[(size_t)i * s.steps] : 0;
    s.done[i] = 0;
    if (s.lengths) s.lengths[i] = s.steps;
  }
  for (int j = i; j < s.B * s.steps; j += gridDim.x * blockDim.x) s.ids[j] = 0;
}

static TmaEncoders g_attn_enc;

static inline size_t group_ws_bytes(int Bg) {
  // xh[2][Bg][2H] bf16 | hp [Bg][H] f32 | c [Bg][H] f32 | tok [Bg] | done [Bg] | live, rounded to 256 bytes
  size_t b = (size_t)Bg * AT_H * (2 * 2 * 2 + 4 + 4) + (size_t)Bg * 8 + 16;
  return (b + 255) & ~(size_t)255;
}

struct DecodeArgs {
  const __nv_bfloat16* enc; const __nv_bfloat16* proj; const void* w_head; const void* w_cat;
  const float* b_h2h; const float* b_gen; const float* w_score; const float* w_tok; const float* bias;
  const int64_t* text; float* logits; int32_t* ids; int32_t* lengths; float* alpha_out;
  int T, C, steps, early_exit;
};

static std::atomic<int> g_persistent{1};    // 1 = one cooperative launch for all steps (default); 0 = launch chain

// One crop group [g0, g0 + Bg): its own state block and either the persistent launch or the chain of launches on `st`.
static int decode_group(const DecodeArgs& a, uint8_t* ws, int g0, int Bg, cudaStream_t st, bool persistent) {
  DecodeState s;
  const size_t xh_bytes = (size_t)Bg * 2 * AT_H * 2;
  s.xh[0] = reinterpret_cast<__nv_bfloat16*>(ws);
  s.xh[1] = reinterpret_cast<__nv_bfloat16*>(ws + xh_bytes);
  s.hp = reinterpret_cast<float*>(ws + 2 * xh_bytes);
  s.cst = s.hp + (size_t)Bg * AT_H;
  s.tok = reinterpret_cast<int32_t*>(s.cst + (size_t)Bg * AT_H);
  s.done = s.tok + Bg;
  s.live = s.done + Bg;
  s.b_h2h = a.b_h2h; s.b_gen = a.b_gen; s.w_score = a.w_score; s.w_tok = a.w_tok; s.bias = a.bias;
  s.text = a.text ? a.text + (size_t)g0 * a.steps : nullptr;
  s.logits = a.logits + (size_t)g0 * a.steps * a.C;
  s.ids = a.ids + (size_t)g0 * a.steps;
  s.lengths = a.lengths ? a.lengths + g0 : nullptr;
  s.alpha_out = a.alpha_out ? a.alpha_out + (size_t)g0 * a.steps * a.T : nullptr;
  s.B = Bg; s.T = a.T; s.C = a.C; s.steps = a.steps; s.early_exit = a.early_exit;
  // h_0 = c_0 = 0 (decoders.py:36-38); ctx halves are overwritten before use
  MSTR_CUDA(cudaMemsetAsync(ws, 0, 2 * xh_bytes + (size_t)Bg * AT_H * 8, st));
  decode_init_kernel<<<ceil_div(Bg, 256), 256, 0, st>>>(s);
  MSTR_LAUNCH_CHECK("decode_init_kernel");

  // tensor maps: h view [Bg][256] with row stride 512 elements; [ctx|h] view [Bg][512]; weights
  alignas(64) CUtensorMap tmH[2], tmXH[2], tmHead, tmCat;
  int rc;
  for (int p = 0; p < 2; ++p) {
    cuuint64_t dims[2] = {(cuuint64_t)AT_H, (cuuint64_t)Bg};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * AT_H * 2)};
    cuuint32_t box[2] = {64, AT_BLOCK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_attn_enc.tiled(&tmH[p], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, s.xh[p] + AT_H, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSTR_REQUIRE(r == CUDA_SUCCESS, MSTR_ERR_CUDA, "cuTensorMapEncodeTiled(h) failed with CUresult %d", (int)r);
    rc = encode_tiled_2d_bf16(g_attn_enc, &tmXH[p], s.xh[p], 2 * AT_H, Bg, 64, AT_BLOCK, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != MSTR_OK) return rc;
  }
  rc = encode_tiled_2d_bf16(g_attn_enc, &tmHead, a.w_head, AT_H, 3 * AT_BLOCK, 64, AT_BLOCK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;
  rc = encode_tiled_2d_bf16(g_attn_enc, &tmCat, a.w_cat, 2 * AT_H, 4 * AT_H, 64, AT_BLOCK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != MSTR_OK) return rc;

  const int m_tiles = ceil_div(Bg, AT_BLOCK);
  const __nv_bfloat16* enc = a.enc + (size_t)g0 * a.T * AT_H;
  const __nv_bfloat16* proj = a.proj + (size_t)g0 * a.T * AT_H;
  if (persistent) {
    static unsigned long long smem_set_pd = 0;
    rc = ensure_dynamic_smem((const void*)attn_decode_persistent_kernel, PD_SMEM_BYTES, &smem_set_pd);
    if (rc != MSTR_OK) return rc;
    int per_sm = 0;
    MSTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_decode_persistent_kernel, PD_THREADS, PD_SMEM_BYTES));
    MSTR_REQUIRE(per_sm >= 1, MSTR_ERR_UNSUPPORTED, "attn_decode_persistent_kernel does not fit on an SM");
    PdMaps maps;
    maps.h[0] = tmH[0]; maps.h[1] = tmH[1]; maps.xh[0] = tmXH[0]; maps.xh[1] = tmXH[1]; maps.head = tmHead; maps.cat = tmCat;
    cudaLaunchAttribute coop[1];
    coop[0].id = cudaLaunchAttributeCooperative;
    coop[0].val.cooperative = 1;
    cudaLaunchConfig_t c = {};
    c.gridDim = dim3(sm_count()); c.blockDim = dim3(PD_THREADS); c.dynamicSmemBytes = PD_SMEM_BYTES; c.stream = st;
    c.attrs = coop; c.numAttrs = 1;
    MSTR_CUDA(cudaLaunchKernelEx(&c, attn_decode_persistent_kernel, maps, enc, proj, s,
                                 reinterpret_cast<unsigned int*>(s.live + 1)));
    MSTR_LAUNCH_CHECK("attn_decode_persistent_kernel");
    return MSTR_OK;
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
      MSTR_CUDA(cudaLaunchKernelEx(&c, rnn_gemm_kernel<0>, tmH[i & 1], tmHead, s, i, AT_H / 64));
      MSTR_LAUNCH_CHECK("rnn_gemm_kernel<head>");
    }
    if (i == a.steps) break;
    {
      cudaLaunchConfig_t c = config(dim3(ceil_div(Bg, AT_CTX_WARPS)), 32 * AT_CTX_WARPS, 0);
      MSTR_CUDA(cudaLaunchKernelEx(&c, attn_context_kernel, enc, proj, s, i));
      MSTR_LAUNCH_CHECK("attn_context_kernel");
    }
    {
      cudaLaunchConfig_t c = config(dim3(4 * AT_H / AT_BLOCK, m_tiles), AT_THREADS, AT_SMEM_BYTES);
      MSTR_CUDA(cudaLaunchKernelEx(&c, rnn_gemm_kernel<1>, tmXH[i & 1], tmCat, s, i, 2 * AT_H / 64));
      MSTR_LAUNCH_CHECK("rnn_gemm_kernel<cell>");
    }
  }
  return MSTR_OK;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_debug_decode_trace(unsigned long long* out_host, int n) {
  using namespace mstr;
  MSTR_REQUIRE(out_host && n > 0 && n <= PD_TRACE_MAX, MSTR_ERR_BAD_ARG, "mstr_debug_decode_trace: bad argument");
  MSTR_CUDA(cudaMemcpyFromSymbol(out_host, g_pd_trace, sizeof(unsigned long long) * (size_t)n));
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_set_decode_persistent(int on) {
  mstr::g_persistent.store(on ? 1 : 0);
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int64_t mstr_attn_decode_bf16_workspace_bytes(int B, int Hd) {
  (void)Hd;
  using namespace mstr;
  return (int64_t)group_ws_bytes((B + AT_BLOCK - 1) / AT_BLOCK * AT_BLOCK);
}

extern "C" __attribute__((visibility("default"))) int mstr_attn_decode_bf16(
    const void* enc_bf16, const void* proj_bf16, const void* w_head, const float* b_h2h, const float* b_gen,
    const float* w_score, const void* w_cat, const float* w_tok, const float* bias, const int64_t* text,
    void* workspace, float* logits, int32_t* ids, int32_t* lengths, float* alpha_out, int B, int T, int Hd, int C,
    int steps, int early_exit, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(enc_bf16 && proj_bf16 && w_head && b_h2h && b_gen && w_score && w_cat && w_tok && bias && workspace &&
                   logits && ids,
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
  return decode_group(a, static_cast<uint8_t*>(workspace), 0, B, st, g_persistent.load(std::memory_order_relaxed) != 0);
}
