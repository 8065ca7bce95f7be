// Shared building block of the two recurrent kernels (bilstm.cu, attn_decode.cu):
//
//   acc[g][c] += sum_k Wt[row(g)][k] * X[c][k]        g = 0..3 (four weight rows per thread),
//                                                     c = 0..CPT-1 (crops of this warp)
//
// A CTA of 256 threads works on a tile of 8*CPT crops whose activations X live in shared memory
// ([crop][K] rows, padded), while the weight matrix is *streamed* from L2 through a 3-deep
// cp.async ring of [128 rows][32 k] chunks.  lane -> row-within-pass, warp -> crop group, so every
// weight value fetched from shared memory is reused for CPT crops and every activation value for
// 4 rows (32*CPT FMAs per 4+CPT 16-byte shared loads).  The four rows of a thread are the i,f,g,o
// gate rows of one hidden unit when the matrix is an LSTM weight, which makes the whole cell
// update thread-local.
#pragma once
#include "common.cuh"

namespace mstr {

constexpr int SG_ROWS = 128;           // weight rows per pass (4 per thread x 32 lanes)
constexpr int SG_KC = 32;              // k values per streamed chunk
constexpr int SG_WLD = SG_KC + 4;      // padded chunk row stride (floats): conflict-free LDS.128
constexpr int SG_NBUF = 3;
constexpr int SG_WBUF_FLOATS = SG_ROWS * SG_WLD;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Describes one streamed matrix: row-major [rows][ld] fp32 in global memory; logical row of
// (pass, g, lane) is g*gate_stride + pass*pass_stride + lane; rows >= rows_valid read as zero.
struct SgMatrix {
  const float* w;
  int ld;           // leading dimension (floats)
  int k_len;        // K extent streamed (multiple of SG_KC)
  int gate_stride;  // distance between the 4 rows of a thread
  int pass_stride;  // distance between passes
  int rows_valid;
};

// Issue the cp.async loads of chunk (pass, kc) of matrix `m` into ring buffer `buf`.
__device__ __forceinline__ void sg_issue_chunk(const SgMatrix& m, int pass, int kc, float* buf) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + 256 * i;
    const int row = idx >> 3, seg = idx & 7;
    const int g = row >> 5, j = row & 31;
    const int grow = g * m.gate_stride + pass * m.pass_stride + j;
    float* dst = buf + row * SG_WLD + seg * 4;
    if (grow < m.rows_valid)
      cp_async16(dst, m.w + (size_t)grow * m.ld + kc * SG_KC + seg * 4);
    else
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Multiply one resident chunk: acc[g][c] += W[g*32+lane][kk] * X[c][kk], kk in [0, SG_KC).
// xs points at X[first crop of this warp][k offset of the chunk]; x_ld is the crop row stride.
template <int CPT>
__device__ __forceinline__ void sg_compute_chunk(const float* __restrict__ wbuf, const float* __restrict__ xs,
                                                 int x_ld, float (&acc)[4][CPT]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < SG_KC; kk += 4) {
    float4 wv[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) wv[g] = *reinterpret_cast<const float4*>(wbuf + (g * 32 + lane) * SG_WLD + kk);
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + c * x_ld + kk);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        acc[g][c] = fmaf(wv[g].x, xv.x, acc[g][c]);
        acc[g][c] = fmaf(wv[g].y, xv.y, acc[g][c]);
        acc[g][c] = fmaf(wv[g].z, xv.z, acc[g][c]);
        acc[g][c] = fmaf(wv[g].w, xv.w, acc[g][c]);
      }
    }
  }
}

}  // namespace mstr
