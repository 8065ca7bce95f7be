// Persistent attention decoder: every greedy step of Attention.forward for a tile of crops in ONE
// launch - Bahdanau attention, LSTMCell, generator, argmax feedback, early exit on [s].
//
// Replaces modules/decoders.py:26-64 (Attention.forward: greedy branch :52-62, teacher-forced
// branch :41-50), :79-101 (AttentionCell.forward), :17/:59 (generator) and the per-step
// _char_to_onehot CPU->GPU round trip (:19-24; a one-hot times W_ih is a row gather of W_tok).
// The `semantics` argument of the reference is not an input here because the reference never
// reads it (decoders.py:79-101).
//
// Per step, for the 32 crops of a CTA (warp w owns crops 4w..4w+3 end to end, so all per-crop
// dependencies are warp-local and only the weight ring needs block barriers):
//   A  hp     = h W_h2h^T + b_h2h                                   streamed GEMM, 256 rows
//   B  e_t    = w_score . tanh(proj[b,t,:] + hp), alpha = softmax_T(e), ctx = sum_t alpha_t enc[b,t,:]
//             (lane = 8 channels, 4 time steps of 16-byte loads in flight, shuffle reductions)
//   C  gates  = ctx W_ctx^T + h W_hh^T + W_tok[token] + bias  ->  (h', c')   streamed GEMM, 1024 rows, K = 512
//   D  logits = h' W_gen^T + b_gen  -> argmax (lowest index on ties, like torch.max on CPU)
// Weights (2.4 MB fp32) stream from L2 through the 3-deep cp.async ring of stream_gemm.cuh; proj
// and enc (2 x T x 1 KB per crop) are re-read from L2/HBM every step.
// Roofline: fp32 FMA pipe for A/C/D (628 kMAC per crop-step), HBM/L2 for B.
#include "stream_gemm.cuh"

namespace mstr {

// A teacher-forcing token outside [0, C) would index past W_tok (the reference's scatter_ raises, decoders.py:23): it is
// mapped to 0 = [GO]; the host side validates the text under config.CHECK_INPUTS.
__device__ __forceinline__ int ad_token(const int64_t t, const int C) { return (t < 0 || t >= C) ? 0 : (int)t; }


constexpr int AD_H = 256;
constexpr int AD_LD = AD_H + 4;
constexpr int AD_MAXC = 128;        // num_classes <= 128 (one generator pass)
constexpr int AD_MAX_TSLOTS = 3;    // T <= 96
// crops per thread AD_CPT = 4, 2 or 1 (8 AD_CPT crops per CTA): the largest that still gives every SM a CTA.  Every CTA
// streams all the weights per step whatever its tile, so at small batches small tiles cost L2 traffic, not time (256
// crops: 32 CTAs of 8 crops instead of 8 of 32); a crop's arithmetic does not depend on the choice.
template <int CPT> constexpr int ad_smem_floats() { return 3 * 8 * CPT * AD_LD + 8 * CPT * AD_H + SG_NBUF * SG_WBUF_FLOATS; }

struct AttnParams {
  const float* enc; const float* proj;
  const float* w_h2h; const float* b_h2h; const float* w_score;
  const float* w_ctx; const float* w_tok; const float* w_hh; const float* bias;
  const float* w_gen; const float* b_gen;
  const int64_t* text;
  const float* h0; const float* c0;            // optional initial state [B][H] (null = zeros)
  float* logits; int32_t* ids; int32_t* lengths;
  float* alpha_out;                            // optional [B][steps][T] attention weights
  float* h_out; float* c_out;                  // optional final state [B][H]
  int B, T, C, steps, early_exit;
};

__device__ __forceinline__ float ad_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ad_tanh(float x) { return fmaf(2.f, ad_sigmoid(2.f * x), -1.f); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int AD_CPT>
__global__ void __launch_bounds__(256, 1) attn_decode_kernel(const AttnParams p) {
  constexpr int AD_BT = 8 * AD_CPT;
  extern __shared__ __align__(16) float smem_f[];
  float* h_cur = smem_f;                       // [BT][LD]
  float* h_nxt = h_cur + AD_BT * AD_LD;        // [BT][LD]
  float* xbuf = h_nxt + AD_BT * AD_LD;         // [BT][LD]  hp -> ctx -> logits staging (per-crop, warp-local)
  float* cst = xbuf + AD_BT * AD_LD;           // [BT][H]
  float* wring = cst + AD_BT * AD_H;
  __shared__ int s_tok[AD_BT];
  __shared__ int s_done[AD_BT];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lrow0 = warp * AD_CPT;
  const int crop0 = blockIdx.x * AD_BT + lrow0;
  const int T = p.T;

  for (int i = tid; i < 3 * AD_BT * AD_LD + AD_BT * AD_H; i += 256) smem_f[i] = 0.f;
  __syncthreads();
  if (p.h0 || p.c0) {
    for (int i = tid; i < AD_BT * AD_H; i += 256) {
      const int r = i / AD_H, u = i % AD_H;
      const int crop = blockIdx.x * AD_BT + r;
      if (crop < p.B) {
        if (p.h0) h_cur[r * AD_LD + u] = p.h0[(size_t)crop * AD_H + u];
        if (p.c0) cst[r * AD_H + u] = p.c0[(size_t)crop * AD_H + u];
      }
    }
  }
  if (tid < AD_BT) {
    const int crop = blockIdx.x * AD_BT + tid;
    s_tok[tid] = (p.text && crop < p.B) ? ad_token(p.text[(size_t)crop * p.steps], p.C) : 0;  // [GO] = 0
    s_done[tid] = (crop < p.B) ? 0 : 1;
  }

  SgMatrix m_h2h{p.w_h2h, AD_H, AD_H, 32, 128, AD_H};
  SgMatrix m_ctx{p.w_ctx, AD_H, AD_H, AD_H, 32, 4 * AD_H};
  SgMatrix m_hh{p.w_hh, AD_H, AD_H, AD_H, 32, 4 * AD_H};
  SgMatrix m_gen{p.w_gen, AD_H, AD_H, 32, 128, p.C};
  constexpr int KCH = AD_H / SG_KC;                       // 8 chunks per 256-wide K
  constexpr int CH_A = 2 * KCH, CH_C = 8 * 2 * KCH, CH_D = KCH;
  constexpr int CH_STEP = CH_A + CH_C + CH_D;             // 152
  const int total_chunks = p.steps * CH_STEP;
  auto issue = [&](int idx) {
    if (idx < total_chunks) {
      const int i = idx % CH_STEP;
      float* buf = wring + (idx % SG_NBUF) * SG_WBUF_FLOATS;
      if (i < CH_A) {
        sg_issue_chunk(m_h2h, i / KCH, i % KCH, buf);
      } else if (i < CH_A + CH_C) {
        const int j = i - CH_A, pass = j / (2 * KCH), kk = j % (2 * KCH);
        if (kk < KCH) sg_issue_chunk(m_ctx, pass, kk, buf);
        else sg_issue_chunk(m_hh, pass, kk - KCH, buf);
      } else {
        sg_issue_chunk(m_gen, 0, i - CH_A - CH_C, buf);
      }
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);
  int chunk = 0;
  auto run_chunk = [&](const float* xs, float (&acc)[4][AD_CPT]) {
    cp_async_wait<1>();
    __syncthreads();
    issue(chunk + 2);
    sg_compute_chunk<AD_CPT>(wring + (chunk % SG_NBUF) * SG_WBUF_FLOATS, xs, AD_LD, acc);
    ++chunk;
  };

  // per-lane constants of the attention phase: channels lane*4 + 128*i, i = 0,1
  float4 sc_r[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) sc_r[i] = __ldg(reinterpret_cast<const float4*>(p.w_score + lane * 4 + 128 * i));

  for (int step = 0; step < p.steps; ++step) {
    // ---------------- A: hp = h W_h2h^T + b ----------------
    for (int pass = 0; pass < 2; ++pass) {
      float acc[4][AD_CPT];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float b = __ldg(p.b_h2h + pass * 128 + g * 32 + lane);
#pragma unroll
        for (int c = 0; c < AD_CPT; ++c) acc[g][c] = b;
      }
      for (int kc = 0; kc < KCH; ++kc) run_chunk(h_cur + lrow0 * AD_LD + kc * SG_KC, acc);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int c = 0; c < AD_CPT; ++c) xbuf[(lrow0 + c) * AD_LD + pass * 128 + g * 32 + lane] = acc[g][c];
    }
    __syncwarp();

    // ---------------- B: attention over T for each of this warp's crops ----------------
    for (int c = 0; c < AD_CPT; ++c) {
      const int crop = crop0 + c;
      float* xrow = xbuf + (lrow0 + c) * AD_LD;
      if (crop >= p.B) {   // warp-uniform
        __syncwarp();
        continue;
      }
      float4 hp_r[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) hp_r[i] = *reinterpret_cast<const float4*>(xrow + lane * 4 + 128 * i);
      const float4* pj = reinterpret_cast<const float4*>(p.proj + (size_t)crop * T * AD_H) + lane;
      const float4* en = reinterpret_cast<const float4*>(p.enc + (size_t)crop * T * AD_H) + lane;
      float e_slot[AD_MAX_TSLOTS];
#pragma unroll
      for (int s = 0; s < AD_MAX_TSLOTS; ++s) e_slot[s] = -INFINITY;
      for (int t0 = 0; t0 < T; t0 += 4) {
        float4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < 2; ++i)
            v[u][i] = (t0 + u < T) ? __ldg(pj + (size_t)(t0 + u) * (AD_H / 4) + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            part = fmaf(sc_r[i].x, ad_tanh(v[u][i].x + hp_r[i].x), part);
            part = fmaf(sc_r[i].y, ad_tanh(v[u][i].y + hp_r[i].y), part);
            part = fmaf(sc_r[i].z, ad_tanh(v[u][i].z + hp_r[i].z), part);
            part = fmaf(sc_r[i].w, ad_tanh(v[u][i].w + hp_r[i].w), part);
          }
          const float e = warp_sum(part);
          const int t = t0 + u;
          if (t < T && (t & 31) == lane) {
#pragma unroll
            for (int s = 0; s < AD_MAX_TSLOTS; ++s)
              if ((t >> 5) == s) e_slot[s] = e;
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int s = 0; s < AD_MAX_TSLOTS; ++s) mx = fmaxf(mx, e_slot[s]);
      mx = warp_max(mx);
      float den = 0.f;
#pragma unroll
      for (int s = 0; s < AD_MAX_TSLOTS; ++s) {
        e_slot[s] = (e_slot[s] == -INFINITY) ? 0.f : expf(e_slot[s] - mx);
        den += e_slot[s];
      }
      den = warp_sum(den);
      const float inv = 1.f / den;
      if (p.alpha_out) {
#pragma unroll
        for (int s = 0; s < AD_MAX_TSLOTS; ++s) {
          const int t = s * 32 + lane;
          if (t < T) p.alpha_out[((size_t)crop * p.steps + step) * T + t] = e_slot[s] * inv;
        }
      }
      float4 cx[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
      for (int t0 = 0; t0 < T; t0 += 4) {
        float4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < 2; ++i)
            v[u][i] = (t0 + u < T) ? __ldg(en + (size_t)(t0 + u) * (AD_H / 4) + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = t0 + u;
          float a = 0.f;
#pragma unroll
          for (int s = 0; s < AD_MAX_TSLOTS; ++s) {
            const float cand = __shfl_sync(0xffffffffu, e_slot[s], t & 31);
            if ((t >> 5) == s) a = cand;
          }
          a = (t < T) ? a * inv : 0.f;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            cx[i].x = fmaf(a, v[u][i].x, cx[i].x); cx[i].y = fmaf(a, v[u][i].y, cx[i].y);
            cx[i].z = fmaf(a, v[u][i].z, cx[i].z); cx[i].w = fmaf(a, v[u][i].w, cx[i].w);
          }
        }
      }
      __syncwarp();   // every lane has read hp before it is overwritten with ctx
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(xrow + lane * 4 + 128 * i) = cx[i];
      __syncwarp();
    }

    // ---------------- C: LSTMCell ----------------
    int tok[AD_CPT];
#pragma unroll
    for (int c = 0; c < AD_CPT; ++c) tok[c] = s_tok[lrow0 + c];
    for (int pass = 0; pass < 8; ++pass) {
      const int unit = pass * 32 + lane;
      float acc[4][AD_CPT];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float b = __ldg(p.bias + g * AD_H + unit);
#pragma unroll
        for (int c = 0; c < AD_CPT; ++c)
          acc[g][c] = b + __ldg(p.w_tok + (size_t)tok[c] * (4 * AD_H) + g * AD_H + unit);
      }
      for (int kc = 0; kc < KCH; ++kc) run_chunk(xbuf + lrow0 * AD_LD + kc * SG_KC, acc);
      for (int kc = 0; kc < KCH; ++kc) run_chunk(h_cur + lrow0 * AD_LD + kc * SG_KC, acc);
#pragma unroll
      for (int c = 0; c < AD_CPT; ++c) {
        const float ig = ad_sigmoid(acc[0][c]), fg = ad_sigmoid(acc[1][c]);
        const float gg = ad_tanh(acc[2][c]), og = ad_sigmoid(acc[3][c]);
        float* cp = cst + (lrow0 + c) * AD_H + unit;
        const float cn = fmaf(fg, *cp, ig * gg);
        *cp = cn;
        h_nxt[(lrow0 + c) * AD_LD + unit] = og * ad_tanh(cn);
      }
    }
    __syncwarp();

    // ---------------- D: generator + argmax ----------------
    {
      float acc[4][AD_CPT];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int cls = g * 32 + lane;
        const float b = cls < p.C ? __ldg(p.b_gen + cls) : 0.f;
#pragma unroll
        for (int c = 0; c < AD_CPT; ++c) acc[g][c] = b;
      }
      for (int kc = 0; kc < KCH; ++kc) run_chunk(h_nxt + lrow0 * AD_LD + kc * SG_KC, acc);
#pragma unroll
      for (int c = 0; c < AD_CPT; ++c) {
        const int crop = crop0 + c;
        const bool live = crop < p.B && !(p.early_exit && s_done[lrow0 + c]);
        float best = -INFINITY;
        int besti = 0x7fffffff;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cls = g * 32 + lane;
          if (cls < p.C) {
            if (live) p.logits[((size_t)crop * p.steps + step) * p.C + cls] = acc[g][c];
            if (acc[g][c] > best) { best = acc[g][c]; besti = cls; }   // ascending cls: lowest index wins ties
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0 && crop < p.B) {
          if (live) {
            p.ids[(size_t)crop * p.steps + step] = besti;
            if (besti == 1 && !s_done[lrow0 + c]) {
              if (p.lengths && p.lengths[crop] == p.steps) p.lengths[crop] = step;
              if (p.early_exit) s_done[lrow0 + c] = 1;
            }
          } else {
            p.ids[(size_t)crop * p.steps + step] = 0;   // argmax of the zero-initialised logits row
          }
          s_tok[lrow0 + c] = (p.text && step + 1 < p.steps) ? ad_token(p.text[(size_t)crop * p.steps + step + 1], p.C) : besti;
        }
      }
    }
    { float* tmp = h_cur; h_cur = h_nxt; h_nxt = tmp; }
    if (p.early_exit) {
      // block-uniform: leave when every crop of the tile has emitted [s]
      const int all_done = __syncthreads_and(tid < AD_BT ? s_done[tid] : 1);
      if (all_done) break;
    } else {
      __syncwarp();
    }
  }
  cp_async_wait<0>();
  if (p.h_out || p.c_out) {
    __syncthreads();
    for (int i = tid; i < AD_BT * AD_H; i += 256) {
      const int r = i / AD_H, u = i % AD_H;
      const int crop = blockIdx.x * AD_BT + r;
      if (crop < p.B) {
        if (p.h_out) p.h_out[(size_t)crop * AD_H + u] = h_cur[r * AD_LD + u];
        if (p.c_out) p.c_out[(size_t)crop * AD_H + u] = cst[r * AD_H + u];
      }
    }
  }
}

// lengths[b] = steps (no [s] seen yet); also zero-fills ids when the decoder may exit early.
__global__ void attn_decode_init_kernel(int32_t* lengths, int32_t* ids, int B, int steps, int zero_ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (lengths && i < B) lengths[i] = steps;
  if (zero_ids)
    for (int j = i; j < B * steps; j += gridDim.x * blockDim.x) ids[j] = 0;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_attn_decode(const float* enc, const float* proj, const float* w_h2h, const float* b_h2h,
                                const float* w_score, const float* w_ctx, const float* w_tok, const float* w_hh,
                                const float* bias, const float* w_gen, const float* b_gen, const int64_t* text,
                                const float* h0, const float* c0, float* logits, int32_t* ids, int32_t* lengths,
                                float* alpha_out, float* h_out, float* c_out, int B, int T, int Hd, int C,
                                int steps, int early_exit, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(enc && proj && w_h2h && b_h2h && w_score && w_ctx && w_tok && w_hh && bias && w_gen && b_gen &&
                   logits && ids,
               MSTR_ERR_BAD_ARG, "mstr_attn_decode: null pointer");
  MSTR_REQUIRE(B > 0 && T > 0 && steps > 0 && C > 0, MSTR_ERR_BAD_ARG, "mstr_attn_decode: bad size");
  MSTR_REQUIRE(Hd == AD_H, MSTR_ERR_UNSUPPORTED, "mstr_attn_decode: hidden size %d != %d", Hd, AD_H);
  MSTR_REQUIRE(C <= AD_MAXC, MSTR_ERR_UNSUPPORTED, "mstr_attn_decode: num_classes %d > %d", C, AD_MAXC);
  MSTR_REQUIRE(T <= 32 * AD_MAX_TSLOTS, MSTR_ERR_UNSUPPORTED, "mstr_attn_decode: T=%d > %d", T, 32 * AD_MAX_TSLOTS);
  MSTR_REQUIRE(aligned16(enc) && aligned16(proj) && aligned16(w_h2h) && aligned16(w_ctx) && aligned16(w_hh) &&
                   aligned16(w_gen) && aligned16(w_score),
               MSTR_ERR_BAD_ARG, "mstr_attn_decode: pointers must be 16-byte aligned");
  if (text) early_exit = 0;  // the teacher-forced branch always runs every step (decoders.py:41-50)
  cudaStream_t st = (cudaStream_t)stream;
  attn_decode_init_kernel<<<ceil_div(B, 256), 256, 0, st>>>(lengths, ids, B, steps, early_exit);
  MSTR_LAUNCH_CHECK("attn_decode_init_kernel");
  AttnParams p{enc, proj, w_h2h, b_h2h, w_score, w_ctx, w_tok, w_hh, bias, w_gen, b_gen, text, h0, c0,
               logits, ids, lengths, alpha_out, h_out, c_out, B, T, C, steps, early_exit};
#define MSTR_AD_LAUNCH(CPT)                                                                            \
  do {                                                                                                 \
    static unsigned long long smem_set = 0;                                                            \
    const int smem = ad_smem_floats<CPT>() * (int)sizeof(float);                                       \
    const int rc_attr = ensure_dynamic_smem((const void*)attn_decode_kernel<CPT>, smem, &smem_set);    \
    if (rc_attr != MSTR_OK) return rc_attr;                                                            \
    attn_decode_kernel<CPT><<<ceil_div(B, 8 * CPT), 256, smem, st>>>(p);                               \
  } while (0)
  const int sms = sm_count();
  if (ceil_div(B, 32) >= sms) MSTR_AD_LAUNCH(4);
  else if (ceil_div(B, 16) >= sms) MSTR_AD_LAUNCH(2);
  else MSTR_AD_LAUNCH(1);
#undef MSTR_AD_LAUNCH
  MSTR_LAUNCH_CHECK("attn_decode_kernel");
  return MSTR_OK;
}
