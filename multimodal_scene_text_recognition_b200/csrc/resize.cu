// Batched crop -> model-input preprocessing on the GPU: the loader's `transforms.Resize((32, 100))` +
// `transforms.ToTensor()` on grayscale word crops (reference: coco_dataset.py:27,40-41 applied to the crop
// of get_sample, :239-243; SURVEY.md section 8 row f2 - at 100k crops/s/GPU the reference's PIL loader with
// num_workers=0 is 3-4 orders of magnitude too slow).
//
// torchvision's Resize on a PIL image is Pillow's ImagingResample (src/libImaging/Resample.c): an antialiased
// separable triangle filter in FIXED POINT for 8-bit images - double-precision weights normalised per output
// index and rounded to 22 fractional bits, a horizontal pass into an 8-bit intermediate, then a vertical pass,
// each with round-half-up and clipping; a pass is skipped when that dimension already matches.  This kernel
// reproduces that arithmetic bit for bit (checked against Pillow through oracle/resize_oracle.py):
// the coefficient tables are built per crop in shared memory with IEEE double operations in Pillow's order
// (no FMA contraction: explicit __dmul_rn / __dadd_rn / __ddiv_rn), the two passes are integer, and ToTensor
// is a true fp32 division by 255.
//
// One CTA per crop (ragged: every crop has its own height / width / offset into one packed uint8 buffer).
// Thread t < out_w builds the horizontal table row t, threads out_w.. build the vertical rows; then every
// thread produces output pixels (yy, xx): for each vertical tap it resamples that input row at xx
// horizontally (recomputed instead of staged: a row is shared by ~2 output rows, and the whole job is a few
// hundred integer MACs per output pixel) and accumulates vertically.  The stage is bound by its HBM reads of
// the packed crops (each byte is read once from DRAM, ~4x from L1/L2).
#include "common.cuh"

namespace mstr {

constexpr int kResizePrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS
constexpr int kResizeThreads = 256;

// precompute_coeffs + normalize_coeffs_8bpc for output index xx of an axis resampled from in_size to out_size
__device__ __forceinline__ void resize_coeffs(int in_size, int out_size, int xx, int ksize, int* bounds, int* kk) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                       // bilinear: support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > ksize) xmax = ksize;                           // cannot happen for sizes within the host's bounds
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (t < 0.0) t = -t;
    const double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < xmax; ++x) {
    double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (t < 0.0) t = -t;
    double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    kk[x] = (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << kResizePrecisionBits)));   // weights are >= 0
  }
  bounds[0] = xmin;
  bounds[1] = xmax;
}

__device__ __forceinline__ int resize_clip8(int v) {
  v >>= kResizePrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(kResizeThreads)
resize_to_tensor_u8_kernel(const uint8_t* __restrict__ pixels, const int64_t* __restrict__ offsets,
                           const int32_t* __restrict__ sizes_hw, float* __restrict__ out, int out_h, int out_w,
                           int ksize_h, int ksize_v, int max_in_h, int max_in_w, int* __restrict__ status) {
  extern __shared__ int smem_i[];
  int* bnd_h = smem_i;                       // [out_w][2]
  int* bnd_v = bnd_h + 2 * out_w;            // [out_h][2]
  int* kk_h = bnd_v + 2 * out_h;             // [out_w][ksize_h]
  int* kk_v = kk_h + out_w * ksize_h;        // [out_h][ksize_v]
  const int b = blockIdx.x;
  const int h = sizes_hw[2 * b], w = sizes_hw[2 * b + 1];
  float* dst = out + (size_t)b * out_h * out_w;
  if (h < 1 || w < 1 || h > max_in_h || w > max_in_w) {     // the tables are sized for max_in_h x max_in_w
    if (threadIdx.x == 0 && status) atomicExch(status, b + 1);
    for (int i = threadIdx.x; i < out_h * out_w; i += kResizeThreads) dst[i] = __int_as_float(0x7fc00000);
    return;
  }
  const uint8_t* src = pixels + offsets[b];
  const bool do_h = w != out_w, do_v = h != out_h;          // Pillow skips a pass whose size already matches
  for (int t = threadIdx.x; t < out_w + out_h; t += kResizeThreads) {
    if (t < out_w) {
      if (do_h) resize_coeffs(w, out_w, t, ksize_h, bnd_h + 2 * t, kk_h + t * ksize_h);
    } else if (do_v) {
      const int yy = t - out_w;
      resize_coeffs(h, out_h, yy, ksize_v, bnd_v + 2 * yy, kk_v + yy * ksize_v);
    }
  }
  __syncthreads();
  const int half = 1 << (kResizePrecisionBits - 1);
  for (int i = threadIdx.x; i < out_h * out_w; i += kResizeThreads) {
    const int yy = i / out_w, xx = i - yy * out_w;
    const int y0 = do_v ? bnd_v[2 * yy] : yy, ny = do_v ? bnd_v[2 * yy + 1] : 1;
    const int x0 = do_h ? bnd_h[2 * xx] : xx, nx = do_h ? bnd_h[2 * xx + 1] : 1;
    const int* kh = kk_h + xx * ksize_h;
    const int* kv = kk_v + yy * ksize_v;
    int acc = half;
    int v = 0;
    for (int j = 0; j < ny; ++j) {
      const uint8_t* rowp = src + (size_t)(y0 + j) * w + x0;
      int hv;
      if (do_h) {
        int s = half;
        for (int k = 0; k < nx; ++k) s += (int)rowp[k] * kh[k];
        hv = resize_clip8(s);                                // 8-bit intermediate image of the horizontal pass
      } else {
        hv = rowp[0];
      }
      if (do_v) acc += hv * kv[j];
      else v = hv;
    }
    if (do_v) v = resize_clip8(acc);
    dst[i] = __fdiv_rn((float)v, 255.0f);                    // ToTensor: uint8 -> float32 / 255
  }
}

static inline int resize_ksize(int in_size, int out_size) {
  double scale = (double)in_size / out_size;
  if (scale < 1.0) scale = 1.0;
  return (int)ceil(scale) * 2 + 1;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_resize_to_tensor_u8(
    const void* pixels, const int64_t* offsets, const int32_t* sizes_hw, float* out, int* status, int B, int out_h,
    int out_w, int max_in_h, int max_in_w, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(pixels && offsets && sizes_hw && out, MSTR_ERR_BAD_ARG, "mstr_resize_to_tensor_u8: null pointer");
  MSTR_REQUIRE(B > 0 && out_h > 0 && out_w > 0 && max_in_h > 0 && max_in_w > 0, MSTR_ERR_BAD_ARG,
               "mstr_resize_to_tensor_u8: non-positive size");
  MSTR_REQUIRE(max_in_h <= 16384 && max_in_w <= 16384, MSTR_ERR_UNSUPPORTED,
               "mstr_resize_to_tensor_u8: crops larger than 16384 pixels per side are not supported");
  const int ksize_h = resize_ksize(max_in_w, out_w), ksize_v = resize_ksize(max_in_h, out_h);
  const size_t smem = sizeof(int) * ((size_t)2 * (out_w + out_h) + (size_t)out_w * ksize_h + (size_t)out_h * ksize_v);
  MSTR_REQUIRE(smem <= 200 * 1024, MSTR_ERR_UNSUPPORTED,
               "mstr_resize_to_tensor_u8: coefficient tables need %zu bytes of shared memory (crop %dx%d -> %dx%d)", smem,
               max_in_h, max_in_w, out_h, out_w);
  if (smem > 48 * 1024)
    MSTR_CUDA(cudaFuncSetAttribute(resize_to_tensor_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resize_to_tensor_u8_kernel<<<B, kResizeThreads, smem, (cudaStream_t)stream>>>(
      static_cast<const uint8_t*>(pixels), offsets, sizes_hw, out, out_h, out_w, ksize_h, ksize_v, max_in_h, max_in_w,
      status);
  MSTR_LAUNCH_CHECK("resize_to_tensor_u8_kernel");
  return MSTR_OK;
}
