// Direct 3x3 convolution for the two layers that read the image itself (C_in = 1..4):
// conv0_1/bn0_1/ReLU of the ResNet (modules/feature_extraction.py:58-60,111-113) and
// conv.0/BN/ReLU/MaxPool2d(2,2) of the localisation network (modules/transformation.py:49-52).
//
// K = 9*C_in (9 or 27) is far below a tensor-core tile and the layer is bound by its NHWC output
// write, so it is SIMT.  Register blocking: one thread computes a strip of 4 horizontally adjacent
// conv pixels (x 2 rows in the pooled variant) for 8 output channels.  Per filter tap the 8 weights
// come from shared memory with two broadcast 16-byte loads and feed 32 (64) FMAs - the first version
// (1 pixel x 32 channels per thread) issued one shared load per 4 FMAs and ran at 25 % of the FMA
// pipe.  The input strip (3 or 4 rows x 6 columns per channel) is read once into registers with
// coalesced loads.  Lanes of a warp cover consecutive channel groups first, so the 16-byte bf16
// stores of 4 (8) neighbouring lanes form one contiguous 64-byte (128-byte) run of the NHWC output.
// Round 2, measured and not kept (0.85 ms for the two launches at 4096 crops either way): requesting image b + 1's
// window before image b is computed (load latency is not what the warps wait for), and a [half][group][4] weight
// order in shared memory that removes the 2-way bank conflicts ncu counts on the weight loads at Cout = 64 (38 M of
// 74 M wavefronts; the LSU pipe is busy, but not the limiter).  The 1 -> 64 launch runs at 47 % of the fp32 FMA rate
// (7.55 GFMA in 0.45 ms; 288 of a thread's ~630 instructions are packed FMAs), the 1 -> 32 launch at 2.5 TB/s of
// output writes.
#include "common.cuh"

namespace mstr {

constexpr int kFirstMaxCin = 4;
constexpr int kFirstMaxCout = 64;
constexpr int kImagesPerBlock = 8;

// PADOUT (not pooled): the output is written as a padded-flat activation [B][H+2][W+2][Cout] whose one-pixel ring is
// ZERO (the layout conv_flat.cu consumes): every thread stores its four pixels one row down / one column right, and the
// threads on the image border also store the ring pixels next to their strip.
template <typename T, bool POOL, int CIN, bool PADOUT = false>
__global__ void __launch_bounds__(256, (CIN == 1 && !POOL) ? 3 : 2)
conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     T* __restrict__ out, int B, int H, int W, int Cout) {
  __shared__ __align__(16) float sW[CIN * 9][kFirstMaxCout];
  __shared__ __align__(16) float sScale[kFirstMaxCout], sShift[kFirstMaxCout];
  constexpr int K = CIN * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int o = i % Cout, k = i / Cout;   // w is [Cout][Cin][3][3] => [o][k]
    sW[k][o] = w[o * K + k];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
    sScale[i] = scale[i];
    sShift[i] = shift[i];
  }
  __syncthreads();

  constexpr int PH = POOL ? 2 : 1;             // conv rows per thread
  const int groups = Cout >> 3;                // channel groups of 8
  const int strips = W >> 2;                   // 4-pixel strips per row
  const int rows = H / PH;
  const int item = blockIdx.x * 256 + threadIdx.x;
  if (item >= rows * strips * groups) return;
  const int g = item % groups;
  const int rest = item / groups;
  const int sx = rest % strips, ry = rest / strips;
  const int y0 = ry * PH - 1, x0 = sx * 4 - 1;
  // Row / edge validity does not depend on the image: computed once, then kImagesPerBlock images are
  // processed by the same thread (the weight staging above and the index arithmetic were ~40 % of the
  // instructions when every block handled a single image).  The 4 interior columns of a strip are one
  // aligned 16-byte load (W % 4 == 0), the two edge columns are predicated scalars.
  bool rowok[PH + 2];
#pragma unroll
  for (int dy = 0; dy < PH + 2; ++dy) rowok[dy] = (y0 + dy >= 0) && (y0 + dy < H);
  const bool left_ok = x0 >= 0, right_ok = x0 + 5 < W;
  const int base = y0 * W + x0 + 1;            // first interior column of row y0
  const int b_begin = blockIdx.y * kImagesPerBlock;
  const int b_end = min(b_begin + kImagesPerBlock, B);
  for (int b = b_begin; b < b_end; ++b) {
  float in[CIN][PH + 2][6];
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const float* plane = x + ((size_t)b * CIN + c) * H * W + base;
#pragma unroll
    for (int dy = 0; dy < PH + 2; ++dy) {
      const float* rowp = plane + dy * W;
      float4 mid = make_float4(0.f, 0.f, 0.f, 0.f);
      float lft = 0.f, rgt = 0.f;
      if (rowok[dy]) {
        mid = __ldg(reinterpret_cast<const float4*>(rowp));
        if (left_ok) lft = __ldg(rowp - 1);
        if (right_ok) rgt = __ldg(rowp + 4);
      }
      in[c][dy][0] = lft; in[c][dy][1] = mid.x; in[c][dy][2] = mid.y; in[c][dy][3] = mid.z; in[c][dy][4] = mid.w;
      in[c][dy][5] = rgt;
    }
  }

  // Packed fp32 FMAs (fma.rn.f32x2, two IEEE FMAs per instruction: same bits as scalar FMAs): ncu showed
  // the scalar version issue bound (63 % issue-active with 2.5 eligible-but-not-selected warps per issue,
  // FMA pipe 35 %), 690 instructions per thread of which 288 were FFMA.  Accumulators are channel pairs.
  unsigned long long acc2[PH][4][4];
#pragma unroll
  for (int r = 0; r < PH; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[r][p][j] = 0ull;
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(&sW[c * 9 + dy * 3 + dx][g * 8]);
        const ulonglong2 w01 = wr[0], w23 = wr[1];
        const unsigned long long wv[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
        for (int r = 0; r < PH; ++r)
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = in[c][r + dy][p + dx];
            unsigned long long v2;
            asm("mov.b64 %0, {%1, %1};" : "=l"(v2) : "f"(v));
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][p][j]) : "l"(v2), "l"(wv[j]));
          }
      }
  float acc[PH][4][8];
#pragma unroll
  for (int r = 0; r < PH; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[r][p][2 * j]), "=f"(acc[r][p][2 * j + 1]) : "l"(acc2[r][p][j]));

  float sc[8], sf[8];
  {
    const float4 a0 = *reinterpret_cast<const float4*>(&sScale[g * 8]), a1 = *reinterpret_cast<const float4*>(&sScale[g * 8 + 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&sShift[g * 8]), b1 = *reinterpret_cast<const float4*>(&sShift[g * 8 + 4]);
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sf[0] = b0.x; sf[1] = b0.y; sf[2] = b0.z; sf[3] = b0.w; sf[4] = b1.x; sf[5] = b1.y; sf[6] = b1.z; sf[7] = b1.w;
  }
  auto store8 = [&](T* dst, const float (&v)[8]) {
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    } else {
      reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  };
  if constexpr (POOL) {
    // BN + ReLU, then the 2x2 max: two pooled pixels per thread
    const int Q = W >> 1;
    T* dst = out + (((size_t)b * rows + ry) * Q + sx * 2) * Cout + g * 8;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float m = 0.f;   // ReLU output is >= 0: 0 is the identity of the max
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int p = 0; p < 2; ++p) m = fmaxf(m, fmaf(acc[r][2 * q + p][j], sc[j], sf[j]));
        v[j] = m;
      }
      store8(dst + (size_t)q * Cout, v);
    }
  } else if constexpr (PADOUT) {
    const int Wp = W + 2;
    T* dst = out + (((size_t)b * (H + 2) + ry + 1) * Wp + sx * 4 + 1) * Cout + g * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(acc[0][p][j], sc[j], sf[j]), 0.f);
      store8(dst + (size_t)p * Cout, v);
    }
    // the zero ring: left / right column of this row, and the row above / below (with its corners)
    const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool first_col = sx == 0, last_col = sx == strips - 1;
    if (first_col) store8(dst - Cout, zero);
    if (last_col) store8(dst + (size_t)4 * Cout, zero);
    if (ry == 0 || ry == H - 1) {
      T* edge = (ry == 0) ? dst - (size_t)Wp * Cout : dst + (size_t)Wp * Cout;
      if (ry == 0 && ry == H - 1) {   // H == 1: both rows
        T* below = dst + (size_t)Wp * Cout;
#pragma unroll
        for (int p = 0; p < 4; ++p) store8(below + (size_t)p * Cout, zero);
        if (first_col) store8(below - Cout, zero);
        if (last_col) store8(below + (size_t)4 * Cout, zero);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) store8(edge + (size_t)p * Cout, zero);
      if (first_col) store8(edge - Cout, zero);
      if (last_col) store8(edge + (size_t)4 * Cout, zero);
    }
  } else {
    T* dst = out + (((size_t)b * H + ry) * W + sx * 4) * Cout + g * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(acc[0][p][j], sc[j], sf[j]), 0.f);
      store8(dst + (size_t)p * Cout, v);
    }
  }
  }   // images
}

template <typename T, bool POOL, bool PADOUT = false>
static void launch_first(int cin, dim3 grid, cudaStream_t st, const float* x, const float* w, const float* scale,
                         const float* shift, T* out, int B, int H, int W, int Cout) {
  switch (cin) {
    case 1: conv3x3_first_kernel<T, POOL, 1, PADOUT><<<grid, 256, 0, st>>>(x, w, scale, shift, out, B, H, W, Cout); break;
    case 2: conv3x3_first_kernel<T, POOL, 2, PADOUT><<<grid, 256, 0, st>>>(x, w, scale, shift, out, B, H, W, Cout); break;
    case 3: conv3x3_first_kernel<T, POOL, 3, PADOUT><<<grid, 256, 0, st>>>(x, w, scale, shift, out, B, H, W, Cout); break;
    default: conv3x3_first_kernel<T, POOL, 4, PADOUT><<<grid, 256, 0, st>>>(x, w, scale, shift, out, B, H, W, Cout); break;
  }
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_conv3x3_first(const float* x, const float* w, const float* scale,
                                  const float* shift, void* out, int out_dtype, int B, int C_in,
                                  int H, int W, int C_out, int pool, mstr_stream_t stream) {
  using namespace mstr;
  const bool pad_out = pool == 2;          // pool: 0 = plain, 1 = fused 2x2 max-pool, 2 = padded-flat bf16 output with a zero ring
  if (pad_out) pool = 0;
  MSTR_REQUIRE(!pad_out || out_dtype == MSTR_BF16, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_first: the padded-flat output is bf16 only");
  MSTR_REQUIRE(x && w && scale && shift && out, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: null pointer");
  MSTR_REQUIRE(B > 0 && H > 0 && W > 0, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: bad size");
  MSTR_REQUIRE(C_in >= 1 && C_in <= kFirstMaxCin, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: C_in=%d not in [1,%d] (use mstr_conv_gemm_*)", C_in, kFirstMaxCin);
  MSTR_REQUIRE(C_out % 8 == 0 && C_out <= kFirstMaxCout, MSTR_ERR_UNSUPPORTED,
               "mstr_conv3x3_first: C_out=%d must be a multiple of 8 and <= %d", C_out, kFirstMaxCout);
  MSTR_REQUIRE(W % 4 == 0, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_first: W=%d must be a multiple of 4", W);
  MSTR_REQUIRE(!pool || H % 2 == 0, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_first: pooled variant needs even H");
  MSTR_REQUIRE(aligned16(out) && aligned16(x), MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: x and out must be 16-byte aligned");
  MSTR_REQUIRE(B <= 65535 * kImagesPerBlock, MSTR_ERR_UNSUPPORTED, "mstr_conv3x3_first: batch %d too large (split the batch)", B);
  const int rows = pool ? H / 2 : H;
  dim3 grid(ceil_div(rows * (W / 4) * (C_out / 8), 256), ceil_div(B, kImagesPerBlock), 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MSTR_BF16) {
    if (pool) launch_first<__nv_bfloat16, true>(C_in, grid, st, x, w, scale, shift, (__nv_bfloat16*)out, B, H, W, C_out);
    else if (pad_out) launch_first<__nv_bfloat16, false, true>(C_in, grid, st, x, w, scale, shift, (__nv_bfloat16*)out, B, H, W, C_out);
    else      launch_first<__nv_bfloat16, false>(C_in, grid, st, x, w, scale, shift, (__nv_bfloat16*)out, B, H, W, C_out);
  } else if (out_dtype == MSTR_F32) {
    if (pool) launch_first<float, true>(C_in, grid, st, x, w, scale, shift, (float*)out, B, H, W, C_out);
    else      launch_first<float, false>(C_in, grid, st, x, w, scale, shift, (float*)out, B, H, W, C_out);
  } else {
    MSTR_REQUIRE(false, MSTR_ERR_BAD_ARG, "mstr_conv3x3_first: unknown dtype %d", out_dtype);
  }
  MSTR_LAUNCH_CHECK("conv3x3_first_kernel");
  return MSTR_OK;
}
