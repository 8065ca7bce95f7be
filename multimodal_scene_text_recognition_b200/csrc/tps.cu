// Fused TPS grid generator + bilinear border sampler.
//
// Reference path: GridGenerator.build_P_prime (modules/transformation.py:154-163) materialises
// P_hat and inv_delta_C once per crop (repeat), runs two torch.bmm and then F.grid_sample
// (modules/transformation.py:33-34).  Here one kernel does all three: a CTA owns a tile of 256
// target pixels and a group of 16 crops; every thread keeps its pixel's basis row in registers
// (read once, coalesced, from the transposed basis [K][H*W]), the per-crop 2xK coefficient matrix
// T is staged in shared memory (mode 0 also stages inv_delta_C there and forms T = inv_delta_C
// [C';0] in the reference's order), and each crop then costs 2K FMAs + 4 gathers per channel.
//
// Roofline: HBM.  Algorithmic bytes per crop = C*H*W*4 read + C*H*W*4 written + F*2*4 (C');
// the basis (K*H*W*4 bytes, shared by all crops) stays in L2 and is re-read once per 16 crops.
#include "common.cuh"

namespace mstr {

constexpr int kTpsThreads = 256;
constexpr int kTpsGroup = 16;  // crops per CTA
constexpr int kTpsMaxK = 24;   // F + 3 <= 24 (the reference uses F = 20, model.py:23)

// KFIX > 0: the basis size K is the compile-time constant KFIX (23 = F+3 and 20 = F of the reference geometry):
// the per-pixel dot products are then straight-line FMAs on vectorised shared-memory loads; with a run-time K
// every FMA was guarded (ncu: 250 instructions per pixel per crop, 128 registers, 23 % warps active).
template <int KMAX, int KFIX>
__global__ void __launch_bounds__(kTpsThreads, KFIX > 0 ? 3 : 2)
tps_rectify_kernel(const float* __restrict__ img, const float* __restrict__ c_prime,
                   const float* __restrict__ basis, const float* __restrict__ inv_delta_C,
                   float* __restrict__ out, float* __restrict__ grid_out, int B, int C, int H, int W,
                   int F, int mode) {
  __shared__ __align__(16) float sT[kTpsGroup][KMAX][2];
  __shared__ float sInv[KMAX * KMAX];
  const int n = H * W;
  const int K = KFIX > 0 ? KFIX : ((mode == 0) ? F + 3 : F);
  const int b0 = blockIdx.y * kTpsGroup;
  const int nb = min(kTpsGroup, B - b0);
  const int tid = threadIdx.x;

  if (mode == 0) {
    for (int i = tid; i < K * K; i += kTpsThreads) sInv[i] = inv_delta_C[i];
    __syncthreads();
    // T[b][i][d] = sum_{j<F} inv_delta_C[i][j] * C'[b][j][d]   (rows F..F+2 of [C';0] are zero)
    for (int o = tid; o < nb * K * 2; o += kTpsThreads) {
      const int d = o & 1, i = (o >> 1) % K, b = (o >> 1) / K;
      const float* cp = c_prime + (size_t)(b0 + b) * F * 2 + d;
      float acc = 0.f;
      for (int j = 0; j < F; ++j) acc = fmaf(sInv[i * K + j], cp[j * 2], acc);
      sT[b][i][d] = acc;
    }
  } else {
    for (int o = tid; o < nb * K * 2; o += kTpsThreads) {
      const int d = o & 1, i = (o >> 1) % K, b = (o >> 1) / K;
      sT[b][i][d] = c_prime[(size_t)(b0 + b) * F * 2 + i * 2 + d];
    }
  }
  __syncthreads();

  const int p = blockIdx.x * kTpsThreads + tid;
  if (p >= n) return;
  constexpr int KR = KFIX > 0 ? KFIX : KMAX;
  float row[KR];
#pragma unroll
  for (int k = 0; k < KR; ++k) row[k] = (k < K) ? __ldg(basis + (size_t)k * n + p) : 0.f;

  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  for (int b = 0; b < nb; ++b) {
    float gx = 0.f, gy = 0.f;
    if constexpr (KFIX > 0) {
      // same summation order as the guarded loop (k ascending); two basis functions per 16-byte broadcast load
#pragma unroll
      for (int k = 0; k + 1 < KFIX; k += 2) {
        const float4 t = *reinterpret_cast<const float4*>(&sT[b][k][0]);
        gx = fmaf(row[k], t.x, gx);
        gy = fmaf(row[k], t.y, gy);
        gx = fmaf(row[k + 1], t.z, gx);
        gy = fmaf(row[k + 1], t.w, gy);
      }
      if constexpr (KFIX & 1) {
        const float2 t = *reinterpret_cast<const float2*>(&sT[b][KFIX - 1][0]);
        gx = fmaf(row[KFIX - 1], t.x, gx);
        gy = fmaf(row[KFIX - 1], t.y, gy);
      }
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          gx = fmaf(row[k], sT[b][k][0], gx);
          gy = fmaf(row[k], sT[b][k][1], gy);
        }
      }
    }
    const size_t crop = (size_t)(b0 + b);
    if (grid_out) reinterpret_cast<float2*>(grid_out)[crop * n + p] = make_float2(gx, gy);
    if (!img) continue;  // grid-only call (GridGenerator.build_P_prime on its own)
    // grid_sample(bilinear, border, align_corners=True)
    float ix = (gx + 1.f) * 0.5f * wm1;
    float iy = (gy + 1.f) * 0.5f * hm1;
    ix = fminf(fmaxf(ix, 0.f), wm1);
    iy = fminf(fmaxf(iy, 0.f), hm1);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx1 = ix - x0f, wy1 = iy - y0f;
    const float wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const bool vx1 = (x0 + 1 <= W - 1), vy1 = (y0 + 1 <= H - 1);
    const int x1 = vx1 ? x0 + 1 : x0, y1 = vy1 ? y0 + 1 : y0;
    const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
    for (int c = 0; c < C; ++c) {
      const float* src = img + (crop * C + c) * (size_t)n;
      const float v00 = __ldg(src + y0 * W + x0);
      const float v01 = vx1 ? __ldg(src + y0 * W + x1) : 0.f;
      const float v10 = vy1 ? __ldg(src + y1 * W + x0) : 0.f;
      const float v11 = (vx1 && vy1) ? __ldg(src + y1 * W + x1) : 0.f;
      out[(crop * C + c) * (size_t)n + p] = v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11;
    }
  }
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_tps_rectify(const float* img, const float* c_prime, const float* basis,
                                const float* inv_delta_C, float* out, float* grid_out, int B, int C,
                                int H, int W, int F, int mode, mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(c_prime && basis, MSTR_ERR_BAD_ARG, "mstr_tps_rectify: null pointer");
  MSTR_REQUIRE((img && out) || (!img && !out && grid_out), MSTR_ERR_BAD_ARG,
               "mstr_tps_rectify: give img and out, or neither together with grid_out");
  MSTR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && F > 0, MSTR_ERR_BAD_ARG,
               "mstr_tps_rectify: non-positive size");
  MSTR_REQUIRE(mode == 0 || mode == 1, MSTR_ERR_BAD_ARG, "mstr_tps_rectify: mode must be 0 or 1");
  MSTR_REQUIRE(mode == 1 || inv_delta_C, MSTR_ERR_BAD_ARG, "mstr_tps_rectify: mode 0 needs inv_delta_C");
  MSTR_REQUIRE(F + 3 <= kTpsMaxK, MSTR_ERR_UNSUPPORTED, "mstr_tps_rectify: F + 3 > %d", kTpsMaxK);
  MSTR_REQUIRE(!grid_out || (reinterpret_cast<uintptr_t>(grid_out) & 7u) == 0, MSTR_ERR_BAD_ARG,
               "mstr_tps_rectify: grid_out must be 8-byte aligned");
  dim3 grid(ceil_div(H * W, kTpsThreads), ceil_div(B, kTpsGroup));
  MSTR_REQUIRE(grid.y <= 65535u, MSTR_ERR_UNSUPPORTED, "mstr_tps_rectify: batch too large");
  const int K = (mode == 0) ? F + 3 : F;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 23)
    tps_rectify_kernel<kTpsMaxK, 23><<<grid, kTpsThreads, 0, st>>>(img, c_prime, basis, inv_delta_C, out, grid_out, B, C,
                                                                  H, W, F, mode);
  else if (K == 20)
    tps_rectify_kernel<kTpsMaxK, 20><<<grid, kTpsThreads, 0, st>>>(img, c_prime, basis, inv_delta_C, out, grid_out, B, C,
                                                                  H, W, F, mode);
  else
    tps_rectify_kernel<kTpsMaxK, 0><<<grid, kTpsThreads, 0, st>>>(img, c_prime, basis, inv_delta_C, out, grid_out, B, C,
                                                                 H, W, F, mode);
  MSTR_LAUNCH_CHECK("tps_rectify_kernel");
  return MSTR_OK;
}
