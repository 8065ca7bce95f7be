// fp32 implicit-GEMM convolution with folded-BN / residual / ReLU epilogue (exact-parity mode).
//
// out[m][n] = act( (sum_k A[m][k] * Wt[n][k]) * scale[n] + shift[n] + residual[m][n] )
//   m = (b, p, q) output pixel, k = (r, s, c), A gathered on the fly from the NHWC input.
// Reference call sites: nn.Conv2d(bias=False)+BatchNorm2d(eval)+ReLU of
// modules/feature_extraction.py:32-47,109-150 and modules/transformation.py:52-57; with
// R=S=H=W=1 it is nn.Linear (modules/encoders.py:15,28, modules/decoders.py:71,
// modules/transformation.py:60-61).
//
// This is the fp32-FMA path used for the parity mode ("TF32 mode" of the north star is specified
// as bit-identical ids / 1e-4 logits, which single-pass TF32 cannot reach on this network -
// SURVEY.md 7.3-1).  128x128x16 tiles, 256 threads, 8x8 outputs per thread, register-prefetch
// double buffering.  Roofline: fp32 FMA pipe (not tensor cores); throughput mode is conv_tc.cu.
#include "common.cuh"

namespace mstr {

constexpr int SBM = 128, SBN = 128, SBK = 16, SLD = SBK + 4;  // row stride 20 floats: conflict-free LDS.128

struct ConvGeom {
  int B, H, W, Cin, Cout, R, S, sh, sw, ph, pw, P, Q;
  int M, K;
};

__global__ void __launch_bounds__(256)
conv_gemm_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ residual, float* __restrict__ out, ConvGeom g, int relu) {
  __shared__ __align__(16) float As[2][SBM][SLD];
  __shared__ __align__(16) float Bs[2][SBN][SLD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;

  // loader role: rows lr and lr+64, float4 column lc
  const int lr = tid >> 2, lc = (tid & 3) * 4;
  int pix_b[2], pix_y[2], pix_x[2];
  bool pix_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = m0 + lr + 64 * h;
    pix_ok[h] = m < g.M;
    const int mm = pix_ok[h] ? m : 0;
    const int q = mm % g.Q, p = (mm / g.Q) % g.P;
    pix_b[h] = mm / (g.Q * g.P);
    pix_y[h] = p * g.sh - g.ph;
    pix_x[h] = q * g.sw - g.pw;
  }
  const float* wrow[2];
  bool w_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = n0 + lr + 64 * h;
    w_ok[h] = n < g.Cout;
    wrow[h] = w + (size_t)(w_ok[h] ? n : 0) * g.K + lc;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = g.K / SBK;
  float4 ra[2], rb[2];
  auto load_tile = [&](int kt) {
    const int k0 = kt * SBK;
    const int tap = k0 / g.Cin, c0 = k0 - tap * g.Cin;
    const int r = tap / g.S, s = tap - r * g.S;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int y = pix_y[h] + r, xx = pix_x[h] + s;
      if (pix_ok[h] && y >= 0 && y < g.H && xx >= 0 && xx < g.W)
        ra[h] = __ldg(reinterpret_cast<const float4*>(
            x + (((size_t)pix_b[h] * g.H + y) * g.W + xx) * g.Cin + c0 + lc));
      else
        ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = w_ok[h] ? __ldg(reinterpret_cast<const float4*>(wrow[h] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&As[buf][lr + 64 * h][lc]) = ra[h];
      *reinterpret_cast<float4*>(&Bs[buf][lr + 64 * h][lc]) = rb[h];
    }
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
    for (int kk = 0; kk < SBK; kk += 4) {
      float4 a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(&As[buf][ty + 16 * i][kk]);
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(&Bs[buf][tx + 16 * j][kk]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // epilogue: folded BN, residual, ReLU
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + tx + 16 * j;
    if (n >= g.Cout) continue;
    const float sc = scale ? __ldg(scale + n) : 1.f;
    const float sf = shift ? __ldg(shift + n) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty + 16 * i;
      if (m >= g.M) continue;
      float v = fmaf(acc[i][j], sc, sf);
      if (residual) v += __ldg(residual + (size_t)m * g.Cout + n);
      if (relu) v = fmaxf(v, 0.f);
      out[(size_t)m * g.Cout + n] = v;
    }
  }
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_conv_gemm_f32(const float* x, const float* w, const float* scale,
                                  const float* shift, const float* residual, float* out, int B, int H,
                                  int W, int C_in, int C_out, int R, int S, int stride_h, int stride_w,
                                  int pad_h, int pad_w, int relu, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && w && out, MSTR_ERR_BAD_ARG, "mstr_conv_gemm_f32: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0 && C_in > 0 && C_out > 0 && R > 0 && S > 0 && stride_h > 0 &&
                   stride_w > 0 && pad_h >= 0 && pad_w >= 0,
               MSTR_ERR_BAD_ARG, "mstr_conv_gemm_f32: bad size");
  MSTR_REQUIRE(C_in % SBK == 0, MSTR_ERR_UNSUPPORTED,
               "mstr_conv_gemm_f32: C_in=%d must be a multiple of %d (use mstr_conv3x3_first for image layers)",
               C_in, SBK);
  MSTR_REQUIRE(aligned16(x) && aligned16(w), MSTR_ERR_BAD_ARG, "mstr_conv_gemm_f32: x and w must be 16-byte aligned");
  ConvGeom g;
  g.B = B; g.H = H; g.W = W; g.Cin = C_in; g.Cout = C_out; g.R = R; g.S = S;
  g.sh = stride_h; g.sw = stride_w; g.ph = pad_h; g.pw = pad_w;
  g.P = (H + 2 * pad_h - R) / stride_h + 1;
  g.Q = (W + 2 * pad_w - S) / stride_w + 1;
  MSTR_REQUIRE(g.P > 0 && g.Q > 0, MSTR_ERR_BAD_ARG, "mstr_conv_gemm_f32: empty output");
  const int64_t M64 = (int64_t)B * g.P * g.Q;
  MSTR_REQUIRE(M64 < (1ll << 31), MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_f32: too many output pixels");
  g.M = (int)M64;
  g.K = R * S * C_in;
  dim3 grid(ceil_div(g.M, SBM), ceil_div(C_out, SBN));
  MSTR_REQUIRE(grid.y <= 65535u, MSTR_ERR_UNSUPPORTED, "mstr_conv_gemm_f32: C_out too large");
  conv_gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, scale, shift, residual, out, g, relu);
  MSTR_LAUNCH_CHECK("conv_gemm_f32_kernel");
  return MSTR_OK;
}
