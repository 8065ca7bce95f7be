// NHWC pooling / reduction / cast helpers (all HBM-bound, 16-byte vectorised along C).
//
//   maxpool_nhwc : nn.MaxPool2d incl. the asymmetric k2 s(2,1) p(0,1) of maxpool3
//                  (modules/feature_extraction.py:66,72,78); padding behaves as -inf.
//   mean_hw      : AdaptiveAvgPool2d(1) of the localisation network (modules/transformation.py:57).
//   mean_h       : permute + AdaptiveAvgPool2d((None,1)) + squeeze of model.py:92-93.
#include <float.h>

#include "common.cuh"

namespace mstr {

// one thread = one output pixel x 8 (bf16) or 4 (f32) channels
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
maxpool_nhwc_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int P,
                    int Q, int kh, int kw, int sh, int sw, int ph, int pw) {
  const int cv = C / VEC;
  const int64_t total = (int64_t)B * P * Q * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * VEC;
    int64_t r = i / cv;
    const int q = (int)(r % Q); r /= Q;
    const int p = (int)(r % P);
    const int b = (int)(r / P);
    float m[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) m[v] = -FLT_MAX;
    bool any = false;
    for (int dy = 0; dy < kh; ++dy) {
      const int y = p * sh - ph + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = 0; dx < kw; ++dx) {
        const int xx = q * sw - pw + dx;
        if (xx < 0 || xx >= W) continue;
        any = true;
        const T* src = x + (((int64_t)b * H + y) * W + xx) * C + c;
        uint4 raw = *reinterpret_cast<const uint4*>(src);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int v = 0; v < VEC; ++v) m[v] = fmaxf(m[v], Elem<T>::load(e + v));
      }
    }
    uint4 o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int v = 0; v < VEC; ++v) Elem<T>::store(oe + v, any ? m[v] : -INFINITY);
    *reinterpret_cast<uint4*>(out + (((int64_t)b * P + p) * Q + q) * C + c) = o;
  }
}

// Fast path for the network's pools (bf16, 2x2 window, any stride / padding): one thread = one output pixel x
// 8 channels; the four 16-byte taps are ALL issued before the first max (the generic kernel's runtime loops
// serialise load -> max -> load, four DRAM latencies per item: 4.2 TB/s), out-of-range taps read as -inf.
// bf16 max is exact, so the packed HMNMX2 path gives the same bits as the fp32 path.
// `ring`: the input is a padded-flat activation [B][H + 2 ring][W + 2 ring][C] (conv_flat.cu); its ring is skipped.
__global__ void __launch_bounds__(256)
maxpool2x2_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H, int W, int cv, int P, int Q,
                       int sh, int sw, int ph, int pw, int total, int ring) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c = i % cv;
  int r = i / cv;
  const int q = r % Q; r /= Q;
  const int p = r % P;
  const int b = r / P;
  const int y0 = p * sh - ph, x0 = q * sw - pw;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  uint4 v[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int y = y0 + (t >> 1), xx = x0 + (t & 1);
    const bool ok = (y >= 0) && (y < H) && (xx >= 0) && (xx < W);
    v[t] = ok ? __ldg(x + ((size_t)(b * (H + 2 * ring) + y + ring) * (W + 2 * ring) + xx + ring) * cv + c) : ninf;
  }
  auto mx = [](uint32_t a, uint32_t bb) {
    __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&bb));
    return *reinterpret_cast<uint32_t*>(&r2);
  };
  uint4 o;
  o.x = mx(mx(v[0].x, v[1].x), mx(v[2].x, v[3].x));
  o.y = mx(mx(v[0].y, v[1].y), mx(v[2].y, v[3].y));
  o.z = mx(mx(v[0].z, v[1].z), mx(v[2].z, v[3].z));
  o.w = mx(mx(v[0].w, v[1].w), mx(v[2].w, v[3].w));
  out[i] = o;
}

// [B][HW][C] -> [B][C] f32 mean; one thread per (b, c), coalesced across c.
template <typename T>
__global__ void __launch_bounds__(256)
mean_hw_kernel(const T* __restrict__ x, float* __restrict__ out, int B, int HW, int C) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * C) return;
  const int c = (int)(i % C);
  const int b = (int)(i / C);
  const T* src = x + (int64_t)b * HW * C + c;
  float acc = 0.f;
  for (int j = 0; j < HW; ++j) acc += Elem<T>::load(src + (int64_t)j * C);
  out[i] = acc / (float)HW;
}

// [B][H][W][C] -> [B][W][C] mean over H.
template <typename T>
__global__ void __launch_bounds__(256)
mean_h_kernel(const T* __restrict__ x, T* __restrict__ out, float* __restrict__ out_f32, int B, int H,
              int W, int C) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * W * C) return;
  const int c = (int)(i % C);
  const int w = (int)((i / C) % W);
  const int b = (int)(i / ((int64_t)C * W));
  float acc = 0.f;
  for (int h = 0; h < H; ++h) acc += Elem<T>::load(x + (((int64_t)b * H + h) * W + w) * C + c);
  acc /= (float)H;
  if (out) Elem<T>::store(out + i, acc);
  if (out_f32) out_f32[i] = acc;
}

template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_kernel(const S* __restrict__ x, D* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    Elem<D>::store(out + i, Elem<S>::load(x + i));
}

static int grid_for(int64_t work_items, int threads) {
  int64_t blocks = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_maxpool_nhwc(const void* x, void* out, int dtype, int B, int H, int W, int C,
                                 int kh, int kw, int sh, int sw, int ph, int pw, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && out, MSTR_ERR_BAD_ARG, "mstr_maxpool_nhwc: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0,
               MSTR_ERR_BAD_ARG, "mstr_maxpool_nhwc: bad size");
  MSTR_REQUIRE(aligned16(x) && aligned16(out), MSTR_ERR_BAD_ARG, "mstr_maxpool_nhwc: pointers must be 16-byte aligned");
  const int P = (H + 2 * ph - kh) / sh + 1, Q = (W + 2 * pw - kw) / sw + 1;
  MSTR_REQUIRE(P > 0 && Q > 0, MSTR_ERR_BAD_ARG, "mstr_maxpool_nhwc: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSTR_BF16) {
    MSTR_REQUIRE(C % 8 == 0, MSTR_ERR_UNSUPPORTED, "mstr_maxpool_nhwc: C %% 8 != 0 for bf16");
    const int64_t items = (int64_t)B * P * Q * (C / 8);
    if (kh == 2 && kw == 2 && ph <= 1 && pw <= 1 && items < (int64_t)1 << 31 && (int64_t)B * H * W * (C / 8) < (int64_t)1 << 31) {
      // every window has at least one in-range tap when the padding is smaller than the window
      maxpool2x2_bf16_kernel<<<(int)ceil_div64(items, 256), 256, 0, st>>>(
          (const uint4*)x, (uint4*)out, H, W, C / 8, P, Q, sh, sw, ph, pw, (int)items, 0);
    } else {
      maxpool_nhwc_kernel<__nv_bfloat16, 8><<<grid_for(items, 256), 256, 0, st>>>(
          (const __nv_bfloat16*)x, (__nv_bfloat16*)out, B, H, W, C, P, Q, kh, kw, sh, sw, ph, pw);
    }
  } else if (dtype == MSTR_F32) {
    MSTR_REQUIRE(C % 4 == 0, MSTR_ERR_UNSUPPORTED, "mstr_maxpool_nhwc: C %% 4 != 0 for f32");
    const int64_t items = (int64_t)B * P * Q * (C / 4);
    maxpool_nhwc_kernel<float, 4><<<grid_for(items, 256), 256, 0, st>>>(
        (const float*)x, (float*)out, B, H, W, C, P, Q, kh, kw, sh, sw, ph, pw);
  } else {
    MSTR_REQUIRE(false, MSTR_ERR_BAD_ARG, "mstr_maxpool_nhwc: unknown dtype %d", dtype);
  }
  MSTR_LAUNCH_CHECK("maxpool_nhwc_kernel");
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_maxpool2x2_ring_bf16(const void* x_ring, void* out, int B, int H, int W,
                                                                                 int C, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x_ring && out, MSTR_ERR_BAD_ARG, "mstr_maxpool2x2_ring_bf16: null pointer");
  MSTR_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && C % 8 == 0, MSTR_ERR_BAD_ARG, "mstr_maxpool2x2_ring_bf16: bad size");
  MSTR_REQUIRE(aligned16(x_ring) && aligned16(out), MSTR_ERR_BAD_ARG, "mstr_maxpool2x2_ring_bf16: pointers must be 16-byte aligned");
  const int P = H / 2, Q = W / 2;
  const int64_t items = (int64_t)B * P * Q * (C / 8);
  MSTR_REQUIRE(items < (int64_t)1 << 31 && (int64_t)B * (H + 2) * (W + 2) * (C / 8) < (int64_t)1 << 31, MSTR_ERR_UNSUPPORTED,
               "mstr_maxpool2x2_ring_bf16: tensor too large");
  maxpool2x2_bf16_kernel<<<(int)ceil_div64(items, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x_ring, (uint4*)out, H, W, C / 8, P, Q, 2, 2, 0, 0, (int)items, 1);
  MSTR_LAUNCH_CHECK("maxpool2x2_bf16_kernel");
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_mean_hw(const void* x, float* out, int dtype, int B, int H, int W, int C,
                            mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && out, MSTR_ERR_BAD_ARG, "mstr_mean_hw: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, MSTR_ERR_BAD_ARG, "mstr_mean_hw: bad size");
  const int64_t items = (int64_t)B * C;
  const int blocks = (int)ceil_div64(items, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSTR_BF16)
    mean_hw_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, out, B, H * W, C);
  else if (dtype == MSTR_F32)
    mean_hw_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, out, B, H * W, C);
  else
    MSTR_REQUIRE(false, MSTR_ERR_BAD_ARG, "mstr_mean_hw: unknown dtype %d", dtype);
  MSTR_LAUNCH_CHECK("mean_hw_kernel");
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_mean_h(const void* x, void* out, float* out_f32, int dtype, int B, int H, int W,
                           int C, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && (out || out_f32), MSTR_ERR_BAD_ARG, "mstr_mean_h: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, MSTR_ERR_BAD_ARG, "mstr_mean_h: bad size");
  const int64_t items = (int64_t)B * W * C;
  const int blocks = (int)ceil_div64(items, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSTR_BF16)
    mean_h_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, out_f32, B, H, W, C);
  else if (dtype == MSTR_F32)
    mean_h_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)out, out_f32, B, H, W, C);
  else
    MSTR_REQUIRE(false, MSTR_ERR_BAD_ARG, "mstr_mean_h: unknown dtype %d", dtype);
  MSTR_LAUNCH_CHECK("mean_h_kernel");
  return MSTR_OK;
}

extern "C" __attribute__((visibility("default"))) int mstr_cast(const void* x, void* out, int src_dtype, int dst_dtype, int64_t n,
                         mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(x && out && n > 0, MSTR_ERR_BAD_ARG, "mstr_cast: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = grid_for(n, 256);
  if (src_dtype == MSTR_F32 && dst_dtype == MSTR_BF16)
    cast_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)out, n);
  else if (src_dtype == MSTR_BF16 && dst_dtype == MSTR_F32)
    cast_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)out, n);
  else
    MSTR_REQUIRE(false, MSTR_ERR_UNSUPPORTED, "mstr_cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  MSTR_LAUNCH_CHECK("cast_kernel");
  return MSTR_OK;
}
