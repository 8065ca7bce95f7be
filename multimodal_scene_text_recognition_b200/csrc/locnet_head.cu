// Head of the localisation network in ONE launch: AdaptiveAvgPool2d(1) -> Linear(512, 256) + ReLU -> Linear(256, 2F)
// (modules/transformation.py:57,60-61,76-81).  It produces the 2F fiducial coordinates the TPS grid is built from, so
// every product is an fp32 FMA in all precision modes (the three launches it replaces - mean_hw, a 3xTF32 GEMM with one
// under-filled wave, an fp32 GEMM with N = 40 - took ~130 us at 4096 crops and ~90 us at 512, almost all of it latency;
// this kernel: 125 us and 28 us).
//
// One CTA of 1024 threads per tile of CT = 8 crops (16 crops per CTA halve the L2 reads of the fc1 weight and measured the
// same 125 us at 4096 crops, 39 instead of 28 us at 512):
//   1. pooling: 128 threads per crop, four consecutive channels each: a position row of a crop is one
//      coalesced 1 KB (bf16) / 2 KB (f32) request; the pooled vectors [CT][512] go to shared memory;
//   2. fc1: thread = (K quarter q, output j): partial dot products over k in [128 q, 128 q + 128) for all crops of the
//      tile, weights read as w1t[k][j] (K-major copy: coalesced over j, L2 resident), pooled values broadcast from shared
//      memory; the four partials are summed in shared memory in the fixed order q = 0..3, + bias, ReLU;
//   3. fc2: thread = (crop, output o): dot over 256 with w2t[:][o] (K-major copy, 40 KB, brought into shared memory
//      by cp.async while the tile is pooled: read from L2 by a cold CTA it was 32 dependent round trips - 37 -> 28 us
//      at 512 crops).
// Every crop's arithmetic has a fixed order that does not depend on the tile or the batch it is in.
#include "common.cuh"

namespace mstr {

constexpr int LH_C = 512;      // channels of the last localisation convolution
constexpr int LH_H1 = 256;     // hidden width of localization_fc1
constexpr int LH_THREADS = 1024;

__device__ __forceinline__ void lh_add4(const float* p, float& a0, float& a1, float& a2, float& a3) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
}
__device__ __forceinline__ void lh_add4(const __nv_bfloat16* p, float& a0, float& a1, float& a2, float& a3) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  a0 += __uint_as_float(v.x << 16); a1 += __uint_as_float(v.x & 0xffff0000u);
  a2 += __uint_as_float(v.y << 16); a3 += __uint_as_float(v.y & 0xffff0000u);
}

template <typename T, int CT>
__global__ void __launch_bounds__(LH_THREADS, 1)
locnet_head_kernel(const T* __restrict__ feat, const float* __restrict__ w1t, const float* __restrict__ b1,
                   const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ out, int B, int HW,
                   int O) {
  extern __shared__ __align__(16) float lh_smem[];
  float (*s_pool)[LH_C] = reinterpret_cast<float (*)[LH_C]>(lh_smem);                            // [CT][512]
  float (*s_part)[CT][LH_H1] = reinterpret_cast<float (*)[CT][LH_H1]>(lh_smem + CT * LH_C);      // [4][CT][256]
  float (*s_hid)[LH_H1] = reinterpret_cast<float (*)[LH_H1]>(lh_smem);   // reuses the pooled tile after fc1
  float* s_w2 = lh_smem + CT * LH_C + 4 * CT * LH_H1;                   // [256][O], lands while the tile is pooled
  const int tid = threadIdx.x;
  const int crop0 = blockIdx.x * CT;
  for (int i = tid; i < LH_H1 * O / 4; i += LH_THREADS)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_w2 + 4 * i)),
                 "l"(w2t + 4 * i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---- 1. pooling: 128 threads per crop, 4 channels each, 8 crops per pass
  for (int c = tid >> 7; c < CT; c += LH_THREADS / 128) {
    const int ch = (tid & 127) * 4;
    const int crop = crop0 + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (crop < B) {
      const T* p = feat + ((size_t)crop * HW) * LH_C + ch;
#pragma unroll 16
      for (int i = 0; i < HW; ++i, p += LH_C) lh_add4(p, a0, a1, a2, a3);
    }
    const float inv = 1.f / (float)HW;
    *reinterpret_cast<float4*>(&s_pool[c][ch]) = make_float4(a0 * inv, a1 * inv, a2 * inv, a3 * inv);
  }
  __syncthreads();
  // ---- 2. fc1: K quarter q, output j
  {
    const int q = tid >> 8, j = tid & 255;
    float acc[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = 0.f;
    const float* w = w1t + (size_t)(q * 128) * LH_H1 + j;
#pragma unroll 4   // 16 weight loads in flight per thread
    for (int k = 0; k < 128; k += 4) {
      const float w0 = __ldg(w + (size_t)k * LH_H1), w1 = __ldg(w + (size_t)(k + 1) * LH_H1);
      const float w2 = __ldg(w + (size_t)(k + 2) * LH_H1), w3 = __ldg(w + (size_t)(k + 3) * LH_H1);
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const float4 p = *reinterpret_cast<const float4*>(&s_pool[c][q * 128 + k]);   // warp-wide broadcast
        acc[c] = fmaf(p.x, w0, acc[c]); acc[c] = fmaf(p.y, w1, acc[c]);
        acc[c] = fmaf(p.z, w2, acc[c]); acc[c] = fmaf(p.w, w3, acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) s_part[q][c][j] = acc[c];
  }
  __syncthreads();
  for (int i = tid; i < CT * LH_H1; i += LH_THREADS) {
    const int c = i >> 8, j = i & 255;
    const float v = ((s_part[0][c][j] + s_part[1][c][j]) + s_part[2][c][j]) + s_part[3][c][j] + __ldg(b1 + j);
    s_hid[c][j] = fmaxf(v, 0.f);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- 3. fc2: thread = (crop, output o), weights K-major in shared memory so a warp reads consecutive o
  for (int i = tid; i < CT * O; i += LH_THREADS) {
    const int c = i / O, o = i % O;
    const int crop = crop0 + c;
    if (crop >= B) continue;
    float a = 0.f;
#pragma unroll 8
    for (int k = 0; k < LH_H1; ++k) a = fmaf(s_hid[c][k], s_w2[k * O + o], a);
    out[(size_t)crop * O + o] = a + __ldg(b2 + o);
  }
}

template <typename T, int CT>
static int lh_launch(const T* feat, const float* w1t, const float* b1, const float* w2t, const float* b2, float* out,
                     int B, int HW, int O, cudaStream_t st) {
  const int smem = (CT * LH_C + 4 * CT * LH_H1 + LH_H1 * O) * (int)sizeof(float);
  constexpr int smem_max = (CT * LH_C + 4 * CT * LH_H1 + LH_H1 * 128) * (int)sizeof(float);   // O <= 128
  static unsigned long long smem_set = 0;
  const int rc = ensure_dynamic_smem((const void*)locnet_head_kernel<T, CT>, smem_max, &smem_set);
  if (rc != MSTR_OK) return rc;
  locnet_head_kernel<T, CT><<<ceil_div(B, CT), LH_THREADS, smem, st>>>(feat, w1t, b1, w2t, b2, out, B, HW, O);
  MSTR_LAUNCH_CHECK("locnet_head_kernel");
  return MSTR_OK;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_locnet_head(const void* feat, int dtype, const float* w1t,
                                                                       const float* b1, const float* w2t, const float* b2,
                                                                       float* out, int B, int HW, int C, int H1, int O,
                                                                       mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(feat && w1t && b1 && w2t && b2 && out, MSTR_ERR_BAD_ARG, "mstr_locnet_head: null pointer");
  MSTR_REQUIRE(B > 0 && HW > 0 && O > 0, MSTR_ERR_BAD_ARG, "mstr_locnet_head: bad size");
  MSTR_REQUIRE(C == LH_C && H1 == LH_H1 && O <= 128 && O % 4 == 0, MSTR_ERR_UNSUPPORTED,
               "mstr_locnet_head: written for %d channels, %d hidden units, <= 128 outputs in multiples of 4 (got %d, %d, %d)", LH_C, LH_H1, C, H1, O);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSTR_BF16)
    return lh_launch<__nv_bfloat16, 8>(static_cast<const __nv_bfloat16*>(feat), w1t, b1, w2t, b2, out, B, HW, O, st);
  MSTR_REQUIRE(dtype == MSTR_F32, MSTR_ERR_BAD_ARG, "mstr_locnet_head: unknown dtype %d", dtype);
  return lh_launch<float, 8>(static_cast<const float*>(feat), w1t, b1, w2t, b2, out, B, HW, O, st);
}
