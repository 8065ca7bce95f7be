// Persistent BiLSTM recurrence: all T steps of one direction for a tile of crops in ONE launch.
//
// Replaces the recurrent half of nn.LSTM(bidirectional=True, batch_first=True)
// (modules/encoders.py:14,24; gate order i,f,g,o, zero initial state, the reverse direction runs
// t = T-1..0 with the *_reverse parameters).  The input projection is a separate GEMM
// (mstr_conv_gemm_*), its result gates_x already contains b_ih + b_hh.
//
// grid = (ceil(B / (8 CPT)), 2 directions) with CPT = 7, 4, 2 or 1 crops per thread, the largest that still gives every
// SM a CTA: 74 x 2 = 148 CTAs of 56 crops at B = 4096; at B = 256 CPT = 1 gives 64 CTAs of 8 crops instead of 10 of 56
// (every CTA streams all of W_hh per step whatever its tile, so small tiles cost L2 traffic, not time, while the
// recurrence is latency bound).  A crop's arithmetic does not depend on CPT.
// Per CTA: h_t and c_t of its 8 CPT crops stay in shared memory for the whole sequence; W_hh
// (1024 x 256 fp32 = 1 MB per direction, L2 resident) is streamed through a 3-deep cp.async ring
// (stream_gemm.cuh).  A thread owns the four gate rows of one hidden unit for 7 crops, so the
// sigmoid/tanh cell update is thread-local.  fp32 FMA throughout (parity mode and, in round 1,
// also the bf16 mode's recurrence).  Roofline: fp32 FMA pipe, 2*4H*H MAC per crop-step.
#include "stream_gemm.cuh"

namespace mstr {

constexpr int LS_H = 256;             // hidden size (model.py:30)
constexpr int LS_HLD = LS_H + 4;      // padded h row
template <int CPT> constexpr int ls_smem_floats() { return 2 * 8 * CPT * LS_HLD + 8 * CPT * LS_H + SG_NBUF * SG_WBUF_FLOATS; }

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, fast_sigmoid(2.f * x), -1.f); }

template <int LS_CPT>                 // crops per thread; 8 LS_CPT crops per CTA
__global__ void __launch_bounds__(256, 1)
bilstm_recurrent_kernel(const float* __restrict__ gates_x, const float* __restrict__ w_hh,
                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int B, int T) {
  constexpr int LS_BT = 8 * LS_CPT;
  extern __shared__ __align__(16) float smem_f[];
  float* h_cur = smem_f;
  float* h_nxt = h_cur + LS_BT * LS_HLD;
  float* cst = h_nxt + LS_BT * LS_HLD;
  float* wring = cst + LS_BT * LS_H;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dir = blockIdx.y;
  const int crop0 = blockIdx.x * LS_BT + warp * LS_CPT;  // first crop of this warp
  const int lrow0 = warp * LS_CPT;                       // its row in the smem tiles

  for (int i = tid; i < 2 * LS_BT * LS_HLD + LS_BT * LS_H; i += 256) smem_f[i] = 0.f;

  SgMatrix m;
  m.w = w_hh + (size_t)dir * 4 * LS_H * LS_H;
  m.ld = LS_H; m.k_len = LS_H; m.gate_stride = LS_H; m.pass_stride = 32; m.rows_valid = 4 * LS_H;

  constexpr int CHUNKS_PER_PASS = LS_H / SG_KC;      // 8
  constexpr int PASSES = LS_H / 32;                  // 8
  constexpr int CHUNKS_PER_STEP = PASSES * CHUNKS_PER_PASS;
  const int total_chunks = T * CHUNKS_PER_STEP;
  auto issue = [&](int idx) {
    if (idx < total_chunks) {
      const int in_step = idx % CHUNKS_PER_STEP;
      sg_issue_chunk(m, in_step / CHUNKS_PER_PASS, in_step % CHUNKS_PER_PASS, wring + (idx % SG_NBUF) * SG_WBUF_FLOATS);
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);

  // gates_x prefetch registers for the upcoming pass
  float gxn[4][LS_CPT];
  auto prefetch_gx = [&](int step, int pass) {
    const int t = dir ? (T - 1 - step) : step;
    const int unit = pass * 32 + lane;
#pragma unroll
    for (int c = 0; c < LS_CPT; ++c) {
      const int crop = crop0 + c;
      const bool ok = (crop < B) && (step < T);
      const float* src = gates_x + (((size_t)(ok ? crop : 0) * T + (ok ? t : 0)) * 2 + dir) * (4 * LS_H) + unit;
#pragma unroll
      for (int g = 0; g < 4; ++g) gxn[g][c] = ok ? __ldg(src + g * LS_H) : 0.f;
    }
  };
  prefetch_gx(0, 0);

  int chunk = 0;
  for (int step = 0; step < T; ++step) {
    const int t = dir ? (T - 1 - step) : step;
    for (int pass = 0; pass < PASSES; ++pass) {
      float acc[4][LS_CPT];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int c = 0; c < LS_CPT; ++c) acc[g][c] = gxn[g][c];
      if (pass + 1 < PASSES) prefetch_gx(step, pass + 1); else prefetch_gx(step + 1, 0);

      for (int kc = 0; kc < CHUNKS_PER_PASS; ++kc, ++chunk) {
        cp_async_wait<1>();
        __syncthreads();
        issue(chunk + 2);
        sg_compute_chunk<LS_CPT>(wring + (chunk % SG_NBUF) * SG_WBUF_FLOATS,
                                 h_cur + lrow0 * LS_HLD + kc * SG_KC, LS_HLD, acc);
      }

      const int unit = pass * 32 + lane;
#pragma unroll
      for (int c = 0; c < LS_CPT; ++c) {
        const float ig = fast_sigmoid(acc[0][c]);
        const float fg = fast_sigmoid(acc[1][c]);
        const float gg = fast_tanh(acc[2][c]);
        const float og = fast_sigmoid(acc[3][c]);
        float* cp = cst + (lrow0 + c) * LS_H + unit;
        const float cn = fmaf(fg, *cp, ig * gg);
        *cp = cn;
        const float hn = og * fast_tanh(cn);
        h_nxt[(lrow0 + c) * LS_HLD + unit] = hn;
        const int crop = crop0 + c;
        if (crop < B) {
          const size_t o = ((size_t)crop * T + t) * (2 * LS_H) + dir * LS_H + unit;
          if (out_f32) out_f32[o] = hn;
          if (out_bf16) out_bf16[o] = __float2bfloat16_rn(hn);
        }
      }
    }
    // h_nxt becomes visible to all warps at the first pipeline barrier of the next step
    float* tmp = h_cur; h_cur = h_nxt; h_nxt = tmp;
  }
  cp_async_wait<0>();
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_bilstm_recurrent(const float* gates_x, const float* w_hh, float* out_f32, void* out_bf16,
                                     int B, int T, int Hd, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(gates_x && w_hh && (out_f32 || out_bf16), MSTR_ERR_BAD_ARG, "mstr_bilstm_recurrent: null pointer");
  MSTR_REQUIRE(B > 0 && T > 0, MSTR_ERR_BAD_ARG, "mstr_bilstm_recurrent: bad size");
  MSTR_REQUIRE(Hd == LS_H, MSTR_ERR_UNSUPPORTED, "mstr_bilstm_recurrent: hidden size %d != %d", Hd, LS_H);
  MSTR_REQUIRE(aligned16(w_hh), MSTR_ERR_BAD_ARG, "mstr_bilstm_recurrent: w_hh must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* ob = (__nv_bfloat16*)out_bf16;
#define MSTR_LS_LAUNCH(CPT)                                                                                         \
  do {                                                                                                              \
    static unsigned long long smem_set = 0;                                                                         \
    const int smem = ls_smem_floats<CPT>() * (int)sizeof(float);                                                    \
    const int rc_attr = ensure_dynamic_smem((const void*)bilstm_recurrent_kernel<CPT>, smem, &smem_set);            \
    if (rc_attr != MSTR_OK) return rc_attr;                                                                         \
    bilstm_recurrent_kernel<CPT><<<dim3(ceil_div(B, 8 * CPT), 2), 256, smem, st>>>(gates_x, w_hh, out_f32, ob, B, T); \
  } while (0)
  // the largest tile that still gives every SM a CTA (2 directions)
  const int sms = sm_count();
  if (2 * ceil_div(B, 56) >= sms) MSTR_LS_LAUNCH(7);
  else if (2 * ceil_div(B, 32) >= sms) MSTR_LS_LAUNCH(4);
  else if (2 * ceil_div(B, 16) >= sms) MSTR_LS_LAUNCH(2);
  else MSTR_LS_LAUNCH(1);
#undef MSTR_LS_LAUNCH
  MSTR_LAUNCH_CHECK("bilstm_recurrent_kernel");
  return MSTR_OK;
}
