// Direct 3x3 convolution for the two layers that read the image itself (C_in = 1..4):
// conv0_1/bn0_1/ReLU of the ResNet (modules/feature_extraction.py:58-60,111-113) and
// conv.0/BN/ReLU/MaxPool2d(2,2) of the localisation network (modules/transformation.py:49-52).
//
// K = 9*C_in (9 or 27) is far below a tensor-core tile, and the layer is HBM-bound on its
// NHWC output (~9 FLOP/B), so it is SIMT (see the kernel comment for the thread mapping).
#include "common.cuh"

namespace mstr {

constexpr int kFirstMaxCin = 4;
constexpr int kFirstMaxCout = 64;
template <bool POOL> struct FirstCg { static constexpr int value = 32; };   // output channels per thread

// One thread = one (pooled) output pixel x 32 output channels.  grid = (pixel tiles of 256, crop,
// channel group).  The 3x3 input patch is loaded into registers with coalesced loads along W; the folded weights sit in shared memory as [tap][32] and are read with broadcast
// 16-byte loads (all lanes of a warp use the same tap at the same time), so the inner loop is
// 32 FMAs per 8 LDS.128.  The 32 results go out as one contiguous 64-byte (bf16) / 128-byte (f32) run.
template <typename T, bool POOL, int CIN>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     T* __restrict__ out, int H, int W, int Cout) {
  constexpr int kFirstCg = FirstCg<POOL>::value;
  __shared__ __align__(16) float sW[CIN * 9][kFirstCg];
  __shared__ float sScale[kFirstCg], sShift[kFirstCg];
  constexpr int K = CIN * 9;
  const int cg0 = blockIdx.z * kFirstCg;
  for (int i = threadIdx.x; i < K * kFirstCg; i += blockDim.x) {
    const int o = i % kFirstCg, k = i / kFirstCg;   // w is [Cout][Cin][3][3] => [o][k]
    sW[k][o] = w[(cg0 + o) * K + k];
  }
  if (threadIdx.x < kFirstCg) {
    sScale[threadIdx.x] = scale[cg0 + threadIdx.x];
    sShift[threadIdx.x] = shift[cg0 + threadIdx.x];
  }
  __syncthreads();

  const int P = POOL ? H / 2 : H, Q = POOL ? W / 2 : W;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= P * Q) return;
  const int p = pix / Q, q = pix - p * Q;
  const int b = blockIdx.y;
  float best[kFirstCg];
#pragma unroll
  for (int j = 0; j < kFirstCg; ++j) best[j] = 0.f;   // ReLU output is >= 0, so 0 is the identity of the max
  // pooled variant: the four conv pixels of the 2x2 window one after the other (not unrolled: keeping
  // 4 x 32 accumulators live would spill); their patches overlap and hit L1
#pragma unroll 1
  for (int sub = 0; sub < (POOL ? 4 : 1); ++sub) {
    const int y0 = (POOL ? 2 * p + (sub >> 1) : p) - 1, x0 = (POOL ? 2 * q + (sub & 1) : q) - 1;
    float in[CIN][3][3];
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* plane = x + ((size_t)b * CIN + c) * H * W;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y0 + dy;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xs = x0 + dx;
          in[c][dy][dx] = (yy >= 0 && yy < H && xs >= 0 && xs < W) ? __ldg(plane + yy * W + xs) : 0.f;
        }
      }
    }
    float acc[kFirstCg];
#pragma unroll
    for (int j = 0; j < kFirstCg; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float v = in[c][dy][dx];
          const float4* wr = reinterpret_cast<const float4*>(&sW[c * 9 + dy * 3 + dx][0]);
#pragma unroll
          for (int u = 0; u < kFirstCg / 4; ++u) {
            const float4 wv = wr[u];
            acc[4 * u + 0] = fmaf(v, wv.x, acc[4 * u + 0]);
            acc[4 * u + 1] = fmaf(v, wv.y, acc[4 * u + 1]);
            acc[4 * u + 2] = fmaf(v, wv.z, acc[4 * u + 2]);
            acc[4 * u + 3] = fmaf(v, wv.w, acc[4 * u + 3]);
          }
        }
#pragma unroll
    for (int j = 0; j < kFirstCg; ++j) best[j] = fmaxf(best[j], fmaf(acc[j], sScale[j], sShift[j]));
  }

  T* dst = out + (((size_t)b * P + p) * Q + q) * Cout + cg0;
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int u = 0; u < kFirstCg / 8; ++u) {
      uint4 o;
      o.x = pack_bf16x2(best[8 * u + 0], best[8 * u + 1]); o.y = pack_bf16x2(best[8 * u + 2], best[8 * u + 3]);
      o.z = pack_bf16x2(best[8 * u + 4], best[8 * u + 5]); o.w = pack_bf16x2(best[8 * u + 6], best[8 * u + 7]);
      reinterpret_cast<uint4*>(dst)[u] = o;
    }
  } else {
#pragma unroll
    for (int u = 0; u < kFirstCg / 4; ++u)
      reinterpret_cast<float4*>(dst)[u] = make_float4(best[4 * u], best[4 * u + 1], best[4 * u + 2], best[4 * u + 3]);
  }
}

template <typename T, bool POOL>
static void launch_first(int cin, dim3 grid, cudaStream_t st, const float* x, const float* w, const float* scale,
                         const float* shift, T* out, int H, int W, int Cout) {
  switch (cin) {
    case 1: conv3x3_first_kernel<T, POOL, 1><<<grid, 256, 0, st>>>(x, w, scale, shift, out, H, W, Cout); break;
    case 2: conv3x3_first_kernel<T, POOL, 2><<<grid, 256, 0, st>>>(x, w, scale, shift, out, H, W, Cout); break;
    case 3: conv3x3_first_kernel<T, POOL, 3><<<grid, 256, 0, st>>>(x, w, scale, shift, out, H, W, Cout); break;
    default: conv3x3_first_kernel<T, POOL, 4><<<grid, 256, 0, st>>>(x, w, scale, shift, out, H, W, Cout); break;
  }
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_conv3x3_first(const float* x, const float* w, const float* scale,
                                  const float* shift, void* out, int out_dtype, int B, int C_in,
                                  int H, int W, int C_out, int pool, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && w && scale && shift && out, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: bad size");
  MSTR_REQUIRE(C_in >= 1 && C_in <= kFirstMaxCin, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: C_in=%d not in [1,%d] (use mstr_conv_gemm_*)", C_in, kFirstMaxCin);
  const int kFirstCg = pool ? FirstCg<true>::value : FirstCg<false>::value;
  MSTR_REQUIRE(C_out % kFirstCg == 0, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: C_out=%d must be a multiple of %d", C_out, kFirstCg);
  MSTR_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: pooled variant needs even H, W");
  MSTR_REQUIRE(aligned16(out), MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: out must be 16-byte aligned");
  const int P = pool ? H / 2 : H, Q = pool ? W / 2 : W;
  MSTR_REQUIRE(B <= 65535, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_first: batch %d > 65535 (split the batch)", B);
  dim3 grid(ceil_div(P * Q, 256), B, C_out / kFirstCg);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MSTR_BF16) {
    if (pool) launch_first<__nv_bfloat16, true>(C_in, grid, st, x, w, scale, shift, (__nv_bfloat16*)out, H, W, C_out);
    else      launch_first<__nv_bfloat16, false>(C_in, grid, st, x, w, scale, shift, (__nv_bfloat16*)out, H, W, C_out);
  } else if (out_dtype == MSTR_F32) {
    if (pool) launch_first<float, true>(C_in, grid, st, x, w, scale, shift, (float*)out, H, W, C_out);
    else      launch_first<float, false>(C_in, grid, st, x, w, scale, shift, (float*)out, H, W, C_out);
  } else {
    MSTR_REQUIRE(false, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: unknown dtype %d", out_dtype);
  }
  MSTR_LAUNCH_CHECK("conv3x3_first_kernel");
  return MSTR_OK;
}
