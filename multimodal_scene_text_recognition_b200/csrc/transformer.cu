// Small kernels of the transformer encoder / decoder family: everything between the GEMMs.
// The GEMMs themselves (in_proj, out_proj, linear1/2, hid_to_emb, emb_to_classes) run on the conv-GEMM kernels
// (1x1 "convolutions": conv_tc2.cu / conv_tc.cu on tcgen05 in bf16 mode, conv_tf32.cu / conv_simt.cu in the parity
// modes) with the bias, the ReLU and the residual in their epilogues.
//
//   layer_norm_kernel      nn.LayerNorm over the last dimension (encoders.py:78,166-167; decoders.py:112,364-366)
//   add_pos_kernel         PositionalEncoding.forward: x + pe[:T] (encoders.py:162-165), eval mode (dropout = identity)
//   embed_tokens_kernel    nn.Embedding + PositionalEncoding of the decoder's targets (decoders.py:117,256-262)
//   mha_full_kernel        softmax(q k^T / sqrt(dh) [+ causal mask]) v for whole sequences, nn.MultiheadAttention's core
//                          (encoders.py:186-187; decoders.py:441-442 with tgt_mask, :450-451 over the memory), one warp
//                          per (crop, head), lane = query position
//   mha_step_kernel        the same for ONE query per crop against Lk cached keys: the greedy loop with a K/V cache
//                          (decoders.py:250-279 recomputes every earlier position at every step; with the causal
//                          mask those rows do not change, so the cache gives the same numbers) and the
//                          memory attention (decoders.py:450-451); one warp per crop, every row one coalesced request
//   pair_add_relu_kernel,  get_relevant_semantic of the PRE_ENCODER_MLP / PRE_DECODER_MLP fusion (encoders.py:89-111,
//   relevance_reduce_kernel  decoders.py:147-170): first MLP layer of every (column, object) pair from the two halves of its
//                          GEMM, and last MLP layer + softmax over the objects + weighted sum of the object vectors
//   gelu_kernel            exact (erf) GELU of BERT's intermediate layer (encoders.py:34-66 via transformers' BertModel)
//   argmax_tokens_kernel   t_output[-1].max(1), targets[t+1] = char_index, output[t] = t_output[t] (decoders.py:276-279)
//
// All arithmetic is fp32; T = float (parity modes) or __nv_bfloat16 (throughput mode) is the storage type.
#include "common.cuh"

namespace mstr {

// 16-byte chunk helpers: N consecutive elements of T as f32
template <typename T> struct Chunk;
template <> struct Chunk<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf16x2(w[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                              pack_bf16x2(v[6], v[7]));
  }
};

// One warp per row.  VEC (D a multiple of 32 16-byte chunks): chunk c of lane l is elements (c * 32 + l) * N ..., so every
// load / store instruction of the warp covers 512 consecutive bytes - of x, of the output AND of gamma / beta (float4).
// (A first version gave every lane D / 32 consecutive channels and read gamma / beta with scalar loads at a 64-byte
// lane stride: 32 L1 wavefronts per instruction, 229 us for [106496 x 512] where the traffic needs 30.)  Otherwise
// lane-strided scalar accesses.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  T* __restrict__ out, int64_t rows, int D, float eps) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int N = Chunk<T>::N;
  const T* xr = x + row * D;
  float v[32];                                            // D <= 1024
  const int n = D >> 5;                                   // channels per lane
  float s = 0.f;
  if (VEC) {
#pragma unroll
    for (int c = 0; c < 32 / N; ++c)
      if (c * N < n) {
        float t[N];
        Chunk<T>::load(xr + (c * 32 + lane) * N, t);
#pragma unroll
        for (int u = 0; u < N; ++u) { v[c * N + u] = t[u]; s += t[u]; }
      }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n) { v[i] = Elem<T>::load(xr + i * 32 + lane); s += v[i]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < n) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)D + eps);         // biased variance, like torch
  T* orow = out + row * D;
  if (VEC) {
#pragma unroll
    for (int c = 0; c < 32 / N; ++c)
      if (c * N < n) {
        const int e0 = (c * 32 + lane) * N;
        float t[N];
#pragma unroll
        for (int u = 0; u < N; u += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + e0 + u));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + e0 + u));
          t[u] = fmaf((v[c * N + u] - mean) * rstd, g.x, bb.x);
          t[u + 1] = fmaf((v[c * N + u + 1] - mean) * rstd, g.y, bb.y);
          t[u + 2] = fmaf((v[c * N + u + 2] - mean) * rstd, g.z, bb.z);
          t[u + 3] = fmaf((v[c * N + u + 3] - mean) * rstd, g.w, bb.w);
        }
        Chunk<T>::store(orow + e0, t);
      }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n) {
        const int c = i * 32 + lane;
        Elem<T>::store(orow + c, fmaf((v[i] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c)));
      }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
add_pos_kernel(const TI* __restrict__ x, const float* __restrict__ pe, TO* __restrict__ out, int64_t total, int T, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % D);
  const int t = (int)((i / D) % T);
  Elem<TO>::store(out + i, Elem<TI>::load(x + i) + __ldg(pe + (int64_t)t * D + d));
}

// tokens [B*T] (row-major [B][T]); out[b][t][:] = emb[token] + pe[t0 + t]
template <typename TO>
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ emb, const float* __restrict__ pe,
                    TO* __restrict__ out, int64_t total, int T, int D, int C, int t0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % D);
  const int64_t bt = i / D;
  const int t = (int)(bt % T);
  int tok = tokens[bt];
  if (tok < 0 || tok >= C) tok = 0;                       // the reference's nn.Embedding would raise
  Elem<TO>::store(out + i, __ldg(emb + (int64_t)tok * D + d) + __ldg(pe + (int64_t)(t0 + t) * D + d));
}

// One warp per (crop, head); lane = query position (Lq, Lk <= 32).  K and V of the head are staged in shared memory as
// f32.  q row i of crop b at q + (b*Lq + i)*ldq, key j at k + (b*Lk + j)*ldkv (likewise v); head h at column h*DH.
template <typename T, int DH>
__global__ void __launch_bounds__(128)
mha_full_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, const T* __restrict__ v, int ldkv,
                T* __restrict__ out, int B, int Lq, int Lk, int heads, int causal, float scale) {
  extern __shared__ float smem[];
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  if (item >= (int64_t)B * heads) return;                 // whole warps leave together
  const int b = (int)(item / heads), h = (int)(item % heads);
  float* sk = smem + (size_t)warp_in_block * 2 * 32 * DH;
  float* sv = sk + 32 * DH;
  const T* kb = k + (int64_t)b * Lk * ldkv + h * DH;
  const T* vb = v + (int64_t)b * Lk * ldkv + h * DH;
  {
    constexpr int N = Chunk<T>::N, PER_ROW = DH / N;    // 16-byte chunks
    for (int i = lane; i < Lk * PER_ROW; i += 32) {
      const int j = i / PER_ROW, d = (i % PER_ROW) * N;
      float t[N];
      Chunk<T>::load(kb + (int64_t)j * ldkv + d, t);
#pragma unroll
      for (int u = 0; u < N; ++u) sk[j * DH + d + u] = t[u];
      Chunk<T>::load(vb + (int64_t)j * ldkv + d, t);
#pragma unroll
      for (int u = 0; u < N; ++u) sv[j * DH + d + u] = t[u];
    }
  }
  __syncwarp();
  if (lane >= Lq) return;
  const T* qr = q + ((int64_t)b * Lq + lane) * ldq + h * DH;
  float acc[DH];
  {
    constexpr int N = Chunk<T>::N;
#pragma unroll
    for (int c = 0; c < DH / N; ++c) {
      float t[N];
      Chunk<T>::load(qr + c * N, t);
#pragma unroll
      for (int u = 0; u < N; ++u) acc[c * N + u] = t[u] * scale;         // q is scaled first, like torch
    }
  }
  // K / V rows are read from shared memory as float4 (all lanes the same address: one broadcast wavefront per 16 bytes)
  const float4* sk4 = reinterpret_cast<const float4*>(sk);
  const float4* sv4 = reinterpret_cast<const float4*>(sv);
  float s[32];
  float mx = -INFINITY;
  const int last = causal ? lane : Lk - 1;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s[j] = -INFINITY;
    if (j < Lk && j <= last) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four partial sums: shorter dependent chains
#pragma unroll
      for (int d = 0; d < DH / 4; ++d) {
        const float4 kv = sk4[j * (DH / 4) + d];
        a0 = fmaf(acc[4 * d], kv.x, a0);
        a1 = fmaf(acc[4 * d + 1], kv.y, a1);
        a2 = fmaf(acc[4 * d + 2], kv.z, a2);
        a3 = fmaf(acc[4 * d + 3], kv.w, a3);
      }
      const float a = (a0 + a1) + (a2 + a3);
      s[j] = a;
      mx = fmaxf(mx, a);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s[j] = (j < Lk && j <= last) ? expf(s[j] - mx) : 0.f;
    den += s[j];
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;             // the query is no longer needed: reuse as the output accumulator
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < Lk && j <= last) {
      const float p = s[j] * inv;
#pragma unroll
      for (int d = 0; d < DH / 4; ++d) {
        const float4 vv = sv4[j * (DH / 4) + d];
        acc[4 * d] = fmaf(p, vv.x, acc[4 * d]);
        acc[4 * d + 1] = fmaf(p, vv.y, acc[4 * d + 1]);
        acc[4 * d + 2] = fmaf(p, vv.z, acc[4 * d + 2]);
        acc[4 * d + 3] = fmaf(p, vv.w, acc[4 * d + 3]);
      }
    }
  }
  T* orow = out + ((int64_t)b * Lq + lane) * heads * DH + h * DH;
  constexpr int N = Chunk<T>::N;
#pragma unroll
  for (int c = 0; c < DH / N; ++c) {
    float t[N];
#pragma unroll
    for (int u = 0; u < N; ++u) t[u] = acc[c * N + u];
    Chunk<T>::store(orow + c * N, t);
  }
}

// ONE query per crop.  One warp per crop covers all heads: the E = heads * DH channels of a row are split over the 32
// lanes in contiguous pieces of E / 32 channels (32 / heads lanes per head), so every key / value row is ONE coalesced
// request of the warp (16-byte loads).  Per key: a partial dot product per lane, reduced over the lanes of the head
// with log2(32 / heads) shuffles; the scores of all keys stay in registers (softmax without further shuffles); then the
// weighted sum of the value rows, each lane for its own channels.  (A first version - one warp per (crop, head), lane =
// key, 2-byte loads with a row stride between lanes - took 90 us per launch at 4096 crops; this one a few.)
// Optionally appends this step's k / v rows (k_new, v_new: [B][ld_new]) to the caches as row Lk - 1 first.
template <typename T, int DH, int HEADS>
__global__ void __launch_bounds__(128)
mha_step_kernel(const T* __restrict__ q, int ldq, T* __restrict__ kc, T* __restrict__ vc, int ldkv, int64_t kv_batch_stride,
                const T* __restrict__ k_new, const T* __restrict__ v_new, int ld_new, T* __restrict__ out, int B, int Lk,
                float scale) {
  constexpr int E = DH * HEADS, PER_LANE = E / 32, N = Chunk<T>::N, CH = PER_LANE / N, LPH = 32 / HEADS;
  static_assert(PER_LANE % N == 0 && CH >= 1 && (32 % HEADS) == 0, "E / 32 channels per lane in 16-byte chunks");
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int c0 = lane * PER_LANE;                       // this lane's channels: c0 .. c0 + PER_LANE - 1 (one head)
  T* kb = kc + b * kv_batch_stride + c0;
  T* vb = vc + b * kv_batch_stride + c0;
  float qv[PER_LANE];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float t[N];
    Chunk<T>::load(q + b * ldq + c0 + c * N, t);
#pragma unroll
    for (int i = 0; i < N; ++i) qv[c * N + i] = t[i] * scale;     // q is scaled first, like torch
  }
  if (k_new) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float t[N];
      Chunk<T>::load(k_new + b * ld_new + c0 + c * N, t);
      Chunk<T>::store(kb + (int64_t)(Lk - 1) * ldkv + c * N, t);
      Chunk<T>::load(v_new + b * ld_new + c0 + c * N, t);
      Chunk<T>::store(vb + (int64_t)(Lk - 1) * ldkv + c * N, t);
    }
  }
  float s[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s[j] = -INFINITY;
    if (j < Lk) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float t[N];
        Chunk<T>::load(kb + (int64_t)j * ldkv + c * N, t);
#pragma unroll
        for (int i = 0; i < N; ++i) a = fmaf(qv[c * N + i], t[i], a);
      }
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      s[j] = a;
      mx = fmaxf(mx, a);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s[j] = j < Lk ? expf(s[j] - mx) : 0.f;
    den += s[j];
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) qv[i] = 0.f;       // the query is no longer needed: reuse as the output accumulator
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < Lk) {
      const float p = s[j] * inv;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float t[N];
        Chunk<T>::load(vb + (int64_t)j * ldkv + c * N, t);
#pragma unroll
        for (int i = 0; i < N; ++i) qv[c * N + i] = fmaf(p, t[i], qv[c * N + i]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = qv[c * N + i];
    Chunk<T>::store(out + b * E + c0 + c * N, t);
  }
}

// H[b][t][o][:] = relu(A[b][t][:] + S[b][o][:]): the first layer of the relevance MLP on cat(column, object) is linear,
// so its two halves are computed once per column / per object (two small GEMMs, bias in A) instead of once per pair.
template <typename T>
__global__ void __launch_bounds__(256)
pair_add_relu_kernel(const T* __restrict__ A, const T* __restrict__ S, T* __restrict__ H, int64_t chunks, int Tn, int On,
                     int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= chunks) return;
  constexpr int N = Chunk<T>::N;
  const int cpr = C / N;                                 // chunks per row
  const int c = (int)(i % cpr);
  const int64_t row = i / cpr;                           // (b, t, o)
  const int o = (int)(row % On);
  const int64_t bt = row / On;
  const int64_t b = bt / Tn;
  float a[N], sv[N];
  Chunk<T>::load(A + bt * C + c * N, a);
  Chunk<T>::load(S + (b * On + o) * C + c * N, sv);
#pragma unroll
  for (int u = 0; u < N; ++u) a[u] = fmaxf(a[u] + sv[u], 0.f);
  Chunk<T>::store(H + row * C + c * N, a);
}

// One warp per (crop, column): score_o = H2[b][t][o][:] . w3 + b3, p = softmax over the objects,
// out[b][t][:] = sum_o p_o sem[b][o][:]  (On <= 64, E % 32 == 0, E <= 512).
template <typename T>
__global__ void __launch_bounds__(256)
relevance_reduce_kernel(const T* __restrict__ H2, const float* __restrict__ w3, float b3, const float* __restrict__ sem,
                        T* __restrict__ out, int64_t rows, int Tn, int On, int C, int E) {
  const int64_t bt = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bt >= rows) return;
  constexpr int N = Chunk<T>::N;
  const int64_t b = bt / Tn;
  float sc[2] = {-INFINITY, -INFINITY};                  // score of object o lives in lane o % 32, slot o / 32
  for (int o = 0; o < On; ++o) {
    const T* h = H2 + (bt * On + o) * C;
    float acc = 0.f;
    for (int c = lane; c < C / N; c += 32) {
      float v[N];
      Chunk<T>::load(h + c * N, v);
#pragma unroll
      for (int u = 0; u < N; u += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w3 + c * N + u));
        acc = fmaf(v[u], w.x, acc); acc = fmaf(v[u + 1], w.y, acc); acc = fmaf(v[u + 2], w.z, acc); acc = fmaf(v[u + 3], w.w, acc);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((o & 31) == lane) sc[o >> 5] = acc + b3;
  }
  float mx = fmaxf(sc[0], sc[1]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  const float e0 = lane < On ? expf(sc[0] - mx) : 0.f;
  const float e1 = lane + 32 < On ? expf(sc[1] - mx) : 0.f;
  float den = e0 + e1;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) den += __shfl_xor_sync(0xffffffffu, den, s);
  const float inv = 1.f / den;
  float acc[16];                                         // E <= 512: E / 32 channels per lane, lane-strided
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (int o = 0; o < On; ++o) {
    const float p = __shfl_sync(0xffffffffu, (o < 32 ? e0 : e1), o & 31) * inv;
    const float* sv = sem + (b * On + o) * E;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k * 32 < E) acc[k] = fmaf(p, __ldg(sv + k * 32 + lane), acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k * 32 < E) Elem<T>::store(out + bt * E + k * 32 + lane, acc[k]);
}

// y = 0.5 x (1 + erf(x / sqrt(2))), 16 bytes per thread
template <typename T>
__global__ void __launch_bounds__(256)
gelu_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t chunks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= chunks) return;
  constexpr int N = Chunk<T>::N;
  float v[N];
  Chunk<T>::load(x + i * N, v);
#pragma unroll
  for (int u = 0; u < N; ++u) v[u] = 0.5f * v[u] * (1.f + erff(v[u] * 0.70710678118654752f));
  Chunk<T>::store(out + i * N, v);
}

// One warp per crop: logits row (f32, ld >= C) -> output row, argmax (the lowest index wins ties, like torch.max).
__global__ void __launch_bounds__(256)
argmax_tokens_kernel(const float* __restrict__ logits, int ld, float* __restrict__ logits_out, int64_t out_row_stride,
                     int32_t* __restrict__ ids, int ids_row_stride, int32_t* __restrict__ next_tok, int next_stride,
                     int B, int C) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float best = -INFINITY;
  int besti = 0;
  for (int c = lane; c < C; c += 32) {
    const float y = logits[b * ld + c];
    if (logits_out) logits_out[b * out_row_stride + c] = y;
    if (y > best) { best = y; besti = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if (lane == 0) {
    if (ids) ids[b * ids_row_stride] = besti;
    if (next_tok) next_tok[b * next_stride] = besti;
  }
}

}  // namespace mstr

#define MSTR_EXPORT extern "C" __attribute__((visibility("default")))

MSTR_EXPORT int mstr_layer_norm(const void* x, const float* gamma, const float* beta, void* out, int64_t rows, int D,
                                float eps, int dtype, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && gamma && beta && out, MSTR_ERR_BAD_ARG, "mstr_layer_norm: null pointer");
  MSTR_REQUIRE(rows > 0 && D > 0, MSTR_ERR_BAD_ARG, "mstr_layer_norm: bad size");
  MSTR_REQUIRE(D % 32 == 0 && D <= 1024, MSTR_ERR_UNSUPPORTED, "mstr_layer_norm: D=%d must be a multiple of 32, <= 1024", D);
  const int blocks = (int)ceil_div64(rows * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  const bool al = aligned16(x) && aligned16(out) && aligned16(gamma) && aligned16(beta);
  if (dtype == MSTR_BF16) {
    if (al && D % 256 == 0)
      layer_norm_kernel<bf, true><<<blocks, 256, 0, st>>>(static_cast<const bf*>(x), gamma, beta, static_cast<bf*>(out), rows, D, eps);
    else
      layer_norm_kernel<bf, false><<<blocks, 256, 0, st>>>(static_cast<const bf*>(x), gamma, beta, static_cast<bf*>(out), rows, D, eps);
  } else {
    if (al && D % 128 == 0)
      layer_norm_kernel<float, true><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), gamma, beta, static_cast<float*>(out), rows, D, eps);
    else
      layer_norm_kernel<float, false><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), gamma, beta, static_cast<float*>(out), rows, D, eps);
  }
  MSTR_LAUNCH_CHECK("layer_norm_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_add_pos_encoding(const void* x, int x_dtype, const float* pe, void* out, int out_dtype, int B, int T, int D,
                                      mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && pe && out, MSTR_ERR_BAD_ARG, "mstr_add_pos_encoding: null pointer");
  MSTR_REQUIRE(B > 0 && T > 0 && D > 0, MSTR_ERR_BAD_ARG, "mstr_add_pos_encoding: bad size");
  const int64_t total = (int64_t)B * T * D;
  const int blocks = (int)ceil_div64(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  if (x_dtype == MSTR_BF16 && out_dtype == MSTR_BF16)
    add_pos_kernel<bf, bf><<<blocks, 256, 0, st>>>(static_cast<const bf*>(x), pe, static_cast<bf*>(out), total, T, D);
  else if (x_dtype == MSTR_F32 && out_dtype == MSTR_BF16)
    add_pos_kernel<float, bf><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), pe, static_cast<bf*>(out), total, T, D);
  else if (x_dtype == MSTR_F32 && out_dtype == MSTR_F32)
    add_pos_kernel<float, float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), pe, static_cast<float*>(out), total, T, D);
  else
    MSTR_REQUIRE(false, MSTR_ERR_UNSUPPORTED, "mstr_add_pos_encoding: unsupported dtype pair %d -> %d", x_dtype, out_dtype);
  MSTR_LAUNCH_CHECK("add_pos_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_embed_tokens(const int32_t* tokens, const float* emb, const float* pe, void* out, int out_dtype, int B, int T,
                                  int D, int C, int t0, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(tokens && emb && pe && out, MSTR_ERR_BAD_ARG, "mstr_embed_tokens: null pointer");
  MSTR_REQUIRE(B > 0 && T > 0 && D > 0 && C > 0 && t0 >= 0, MSTR_ERR_BAD_ARG, "mstr_embed_tokens: bad size");
  const int64_t total = (int64_t)B * T * D;
  const int blocks = (int)ceil_div64(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MSTR_BF16)
    embed_tokens_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(tokens, emb, pe, static_cast<__nv_bfloat16*>(out), total, T, D, C, t0);
  else
    embed_tokens_kernel<float><<<blocks, 256, 0, st>>>(tokens, emb, pe, static_cast<float*>(out), total, T, D, C, t0);
  MSTR_LAUNCH_CHECK("embed_tokens_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_mha(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int B, int Lq, int Lk,
                         int heads, int dh, int causal, int dtype, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(q && k && v && out, MSTR_ERR_BAD_ARG, "mstr_mha: null pointer");
  MSTR_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && heads > 0, MSTR_ERR_BAD_ARG, "mstr_mha: bad size");
  MSTR_REQUIRE(Lq <= 32 && Lk <= 32 && (dh == 32 || dh == 64), MSTR_ERR_UNSUPPORTED,
               "mstr_mha: Lq=%d, Lk=%d (<= 32), head dim %d (32 or 64)", Lq, Lk, dh);
  MSTR_REQUIRE(!causal || Lq == Lk, MSTR_ERR_BAD_ARG, "mstr_mha: the causal mask needs Lq == Lk");
  MSTR_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && (ldkv * (dtype == MSTR_BF16 ? 2 : 4)) % 16 == 0 &&
                   (ldq * (dtype == MSTR_BF16 ? 2 : 4)) % 16 == 0, MSTR_ERR_BAD_ARG,
               "mstr_mha: pointers and row strides must be multiples of 16 bytes");
  const float scale = 1.0f / sqrtf((float)dh);
  const int warps = 4;
  const int blocks = (int)ceil_div64((int64_t)B * heads, warps);
  const size_t smem = (size_t)warps * 2 * 32 * dh * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
#define MSTR_MHA_FULL(T, DH)                                                                                      \
  do {                                                                                                            \
    static unsigned long long done = 0;                                                                           \
    int rc = ensure_dynamic_smem((const void*)mha_full_kernel<T, DH>, (int)smem, &done);                          \
    if (rc != MSTR_OK) return rc;                                                                                 \
    mha_full_kernel<T, DH><<<blocks, 32 * warps, smem, st>>>(static_cast<const T*>(q), ldq, static_cast<const T*>(k), \
                                                            static_cast<const T*>(v), ldkv, static_cast<T*>(out), B, Lq, \
                                                            Lk, heads, causal, scale);                            \
  } while (0)
  if (dtype == MSTR_BF16) { if (dh == 64) MSTR_MHA_FULL(bf, 64); else MSTR_MHA_FULL(bf, 32); }
  else { if (dh == 64) MSTR_MHA_FULL(float, 64); else MSTR_MHA_FULL(float, 32); }
#undef MSTR_MHA_FULL
  MSTR_LAUNCH_CHECK("mha_full_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_mha_step(const void* q, int ldq, void* k_cache, void* v_cache, int ldkv, int64_t kv_batch_stride,
                              const void* k_new, const void* v_new, int ld_new, void* out, int B, int Lk, int heads, int dh,
                              int dtype, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(q && k_cache && v_cache && out, MSTR_ERR_BAD_ARG, "mstr_mha_step: null pointer");
  MSTR_REQUIRE((k_new == nullptr) == (v_new == nullptr), MSTR_ERR_BAD_ARG, "mstr_mha_step: k_new and v_new go together");
  MSTR_REQUIRE(B > 0 && Lk > 0 && heads > 0, MSTR_ERR_BAD_ARG, "mstr_mha_step: bad size");
  MSTR_REQUIRE(Lk <= 32 && heads == 8 && (dh == 32 || dh == 64), MSTR_ERR_UNSUPPORTED,
               "mstr_mha_step: Lk=%d (<= 32), %d heads (8), head dim %d (32 or 64)", Lk, heads, dh);
  const int esz = dtype == MSTR_BF16 ? 2 : 4;
  MSTR_REQUIRE(aligned16(q) && aligned16(k_cache) && aligned16(v_cache) && aligned16(out) && aligned16(k_new) && aligned16(v_new) &&
                   (ldq * esz) % 16 == 0 && (ldkv * esz) % 16 == 0 && (kv_batch_stride * esz) % 16 == 0 && (ld_new * esz) % 16 == 0,
               MSTR_ERR_BAD_ARG, "mstr_mha_step: pointers and strides must be multiples of 16 bytes");
  const float scale = 1.0f / sqrtf((float)dh);
  const int blocks = ceil_div(B, 4);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
#define MSTR_MHA_STEP(T, DH)                                                                                          \
  mha_step_kernel<T, DH, 8><<<blocks, 128, 0, st>>>(static_cast<const T*>(q), ldq, static_cast<T*>(k_cache),          \
                                                    static_cast<T*>(v_cache), ldkv, kv_batch_stride,                  \
                                                    static_cast<const T*>(k_new), static_cast<const T*>(v_new), ld_new, \
                                                    static_cast<T*>(out), B, Lk, scale)
  if (dtype == MSTR_BF16) { if (dh == 64) MSTR_MHA_STEP(bf, 64); else MSTR_MHA_STEP(bf, 32); }
  else { if (dh == 64) MSTR_MHA_STEP(float, 64); else MSTR_MHA_STEP(float, 32); }
#undef MSTR_MHA_STEP
  MSTR_LAUNCH_CHECK("mha_step_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_pair_add_relu(const void* a, const void* s, void* h, int B, int T, int O, int C, int dtype,
                                   mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(a && s && h, MSTR_ERR_BAD_ARG, "mstr_pair_add_relu: null pointer");
  const int per = dtype == MSTR_BF16 ? 8 : 4;
  MSTR_REQUIRE(B > 0 && T > 0 && O > 0 && C > 0 && C % per == 0 && aligned16(a) && aligned16(s) && aligned16(h), MSTR_ERR_BAD_ARG,
               "mstr_pair_add_relu: bad size / alignment (C must be a multiple of %d)", per);
  const int64_t chunks = (int64_t)B * T * O * (C / per);
  const int blocks = (int)ceil_div64(chunks, 256);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  if (dtype == MSTR_BF16)
    pair_add_relu_kernel<bf><<<blocks, 256, 0, st>>>(static_cast<const bf*>(a), static_cast<const bf*>(s), static_cast<bf*>(h), chunks, T, O, C);
  else
    pair_add_relu_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(s),
                                                        static_cast<float*>(h), chunks, T, O, C);
  MSTR_LAUNCH_CHECK("pair_add_relu_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_relevance_reduce(const void* h2, const float* w3, float b3, const float* sem, void* out, int B, int T, int O,
                                      int C, int E, int dtype, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(h2 && w3 && sem && out, MSTR_ERR_BAD_ARG, "mstr_relevance_reduce: null pointer");
  const int per = dtype == MSTR_BF16 ? 8 : 4;
  MSTR_REQUIRE(B > 0 && T > 0 && O > 0 && C > 0 && C % per == 0 && aligned16(h2) && aligned16(w3), MSTR_ERR_BAD_ARG,
               "mstr_relevance_reduce: bad size / alignment");
  MSTR_REQUIRE(O <= 64 && E % 32 == 0 && E <= 512, MSTR_ERR_UNSUPPORTED, "mstr_relevance_reduce: O=%d (<= 64), E=%d (multiple of 32, <= 512)", O, E);
  const int64_t rows = (int64_t)B * T;
  const int blocks = (int)ceil_div64(rows * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  typedef __nv_bfloat16 bf;
  if (dtype == MSTR_BF16)
    relevance_reduce_kernel<bf><<<blocks, 256, 0, st>>>(static_cast<const bf*>(h2), w3, b3, sem, static_cast<bf*>(out), rows, T, O, C, E);
  else
    relevance_reduce_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(h2), w3, b3, sem, static_cast<float*>(out), rows, T, O, C, E);
  MSTR_LAUNCH_CHECK("relevance_reduce_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_gelu(const void* x, void* out, int64_t n, int dtype, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && out, MSTR_ERR_BAD_ARG, "mstr_gelu: null pointer");
  const int per = dtype == MSTR_BF16 ? 8 : 4;
  MSTR_REQUIRE(n > 0 && n % per == 0 && aligned16(x) && aligned16(out), MSTR_ERR_BAD_ARG,
               "mstr_gelu: n must be a positive multiple of %d, pointers 16-byte aligned", per);
  const int64_t chunks = n / per;
  const int blocks = (int)ceil_div64(chunks, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSTR_BF16)
    gelu_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), chunks);
  else
    gelu_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(out), chunks);
  MSTR_LAUNCH_CHECK("gelu_kernel");
  return MSTR_OK;
}

MSTR_EXPORT int mstr_argmax_tokens(const float* logits, int ld, float* logits_out, int64_t out_row_stride, int32_t* ids,
                                   int ids_row_stride, int32_t* next_tok, int next_stride, int B, int C, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(logits, MSTR_ERR_BAD_ARG, "mstr_argmax_tokens: null pointer");
  MSTR_REQUIRE(B > 0 && C > 0 && ld >= C, MSTR_ERR_BAD_ARG, "mstr_argmax_tokens: bad size");
  const int blocks = (int)ceil_div64((int64_t)B * 32, 256);
  argmax_tokens_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, ld, logits_out, out_row_stride, ids, ids_row_stride,
                                                                 next_tok, next_stride, B, C);
  MSTR_LAUNCH_CHECK("argmax_tokens_kernel");
  return MSTR_OK;
}
