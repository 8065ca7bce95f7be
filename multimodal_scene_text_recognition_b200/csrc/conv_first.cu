// Direct 3x3 convolution for the two layers that read the image itself (C_in = 1..4):
// conv0_1/bn0_1/ReLU of the ResNet (modules/feature_extraction.py:58-60,111-113) and
// conv.0/BN/ReLU/MaxPool2d(2,2) of the localisation network (modules/transformation.py:49-52).
//
// K = 9*C_in (9 or 27) is far below a tensor-core tile, and the layer is HBM-bound on its
// NHWC output (~9 FLOP/B), so it is SIMT: one thread produces 8 consecutive output channels of
// one (optionally 2x2-pooled) output pixel and stores them with a single 16-byte (bf16) or two
// 16-byte (f32) coalesced stores; the folded weights live in shared memory as [k][C_out].
#include "common.cuh"

namespace mstr {

constexpr int kFirstMaxCin = 4;
constexpr int kFirstMaxCout = 64;

template <typename T, bool POOL>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     T* __restrict__ out, int B, int Cin, int H, int W, int Cout) {
  __shared__ __align__(16) float sW[kFirstMaxCin * 9][kFirstMaxCout];
  __shared__ float sScale[kFirstMaxCout], sShift[kFirstMaxCout];
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int o = i % Cout, k = i / Cout;  // w is [Cout][Cin][3][3] => [o][k]
    sW[k][o] = w[o * K + k];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
    sScale[i] = scale[i];
    sShift[i] = shift[i];
  }
  __syncthreads();

  const int P = POOL ? H / 2 : H, Q = POOL ? W / 2 : W;
  const int groups = Cout / 8;
  const int64_t total = (int64_t)B * P * Q * groups;
  constexpr int NPIX = POOL ? 4 : 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    int64_t r = i / groups;
    const int q = (int)(r % Q); r /= Q;
    const int p = (int)(r % P);
    const int b = (int)(r / P);
    float best[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) best[v] = -INFINITY;
#pragma unroll
    for (int sub = 0; sub < NPIX; ++sub) {
      const int y = POOL ? 2 * p + (sub >> 1) : p;
      const int xx = POOL ? 2 * q + (sub & 1) : q;
      float acc[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) acc[v] = 0.f;
      for (int c = 0; c < Cin; ++c) {
        const float* plane = x + ((int64_t)b * Cin + c) * H * W;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int yy = y + dy - 1;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int xs = xx + dx - 1;
            const float v = (yy >= 0 && yy < H && xs >= 0 && xs < W) ? __ldg(plane + yy * W + xs) : 0.f;
            const float4* wr = reinterpret_cast<const float4*>(&sW[c * 9 + dy * 3 + dx][g * 8]);
            const float4 w0 = wr[0], w1 = wr[1];
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
            acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
            acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const float yv = fmaxf(fmaf(acc[v], sScale[g * 8 + v], sShift[g * 8 + v]), 0.f);
        best[v] = fmaxf(best[v], yv);
      }
    }
    T* dst = out + (((int64_t)b * P + p) * Q + q) * Cout + g * 8;
    if constexpr (sizeof(T) == 2) {
      uint4 o;
      o.x = pack_bf16x2(best[0], best[1]); o.y = pack_bf16x2(best[2], best[3]);
      o.z = pack_bf16x2(best[4], best[5]); o.w = pack_bf16x2(best[6], best[7]);
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      reinterpret_cast<float4*>(dst)[0] = make_float4(best[0], best[1], best[2], best[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(best[4], best[5], best[6], best[7]);
    }
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
  MSTR_REQUIRE(C_out % 8 == 0 && C_out <= kFirstMaxCout, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: C_out=%d must be a multiple of 8 and <= %d", C_out, kFirstMaxCout);
  MSTR_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: pooled variant needs even H, W");
  MSTR_REQUIRE(aligned16(out), MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: out must be 16-byte aligned");
  const int P = pool ? H / 2 : H, Q = pool ? W / 2 : W;
  const int64_t items = (int64_t)B * P * Q * (C_out / 8);
  int64_t blocks = ceil_div64(items, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MSTR_BF16) {
    if (pool) conv3x3_first_kernel<__nv_bfloat16, true><<<(int)blocks, 256, 0, st>>>(x, w, scale, shift, (__nv_bfloat16*)out, B, C_in, H, W, C_out);
    else      conv3x3_first_kernel<__nv_bfloat16, false><<<(int)blocks, 256, 0, st>>>(x, w, scale, shift, (__nv_bfloat16*)out, B, C_in, H, W, C_out);
  } else if (out_dtype == MSTR_F32) {
    if (pool) conv3x3_first_kernel<float, true><<<(int)blocks, 256, 0, st>>>(x, w, scale, shift, (float*)out, B, C_in, H, W, C_out);
    else      conv3x3_first_kernel<float, false><<<(int)blocks, 256, 0, st>>>(x, w, scale, shift, (float*)out, B, C_in, H, W, C_out);
  } else {
    MSTR_REQUIRE(false, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: unknown dtype %d", out_dtype);
  }
  MSTR_LAUNCH_CHECK("conv3x3_first_kernel");
  return MSTR_OK;
}
