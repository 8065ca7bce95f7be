// Semantic-vector embedding: Linear_Embedding.forward (modules/semantic_vectors.py:33-49).
//   mode 0 'overlap': out[b][j][:] = table[(long)ids[b][j]]
//   mode 1 'scene'  : out[b][j][:] = softmax(ious[b], dim=1)[j] * table[(long)ids[b][j]]
// One warp per (b, j) row; HBM/L2 gather, 16-byte vectorised along D.
// The decoder never reads the result (modules/decoders.py:79-101 ignores `semantics`); it is
// computed because Model.forward computes it (model.py:96).
#include "common.cuh"

namespace mstr {

__global__ void __launch_bounds__(256)
embed_semantics_kernel(const float* __restrict__ ids, const float* __restrict__ table,
                       const float* __restrict__ ious, float* __restrict__ out, int B, int n, int D,
                       int rows, int mode) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * n) return;
  const int b = warp / n, j = warp % n;
  float weight = 1.f;
  if (mode == 1) {
    // softmax over the n ious of crop b (n <= 64 in practice; loop is warp-strided for any n)
    const float* io = ious + (int64_t)b * n;
    float m = -INFINITY;
    for (int t = lane; t < n; t += 32) m = fmaxf(m, io[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int t = lane; t < n; t += 32) s += expf(io[t] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    weight = expf(io[j] - m) / s;
  }
  long long id = (long long)ids[(int64_t)b * n + j];  // .long(): truncation toward zero
  if (id < 0 || id >= rows) id = 0;                   // the reference would raise; ids are validated on the host
  const float4* src = reinterpret_cast<const float4*>(table + id * D);
  float4* dst = reinterpret_cast<float4*>(out + ((int64_t)b * n + j) * D);
  for (int v = lane; v < D / 4; v += 32) {
    float4 t = __ldg(src + v);
    t.x *= weight; t.y *= weight; t.z *= weight; t.w *= weight;
    dst[v] = t;
  }
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_embed_semantics(const float* ids, const float* table, const float* ious,
                                    float* out, int B, int n, int D, int rows, int mode,
                                    mstr_stream_t stream) {
  using namespace mstr;
  MSTR_REQUIRE(ids && table && out, MSTR_ERR_BAD_ARG, "mstr_embed_semantics: null pointer");
  MSTR_REQUIRE(mode == 0 || (mode == 1 && ious), MSTR_ERR_BAD_ARG, "mstr_embed_semantics: bad mode / missing ious");
  MSTR_REQUIRE(B > 0 && n > 0 && D > 0 && rows > 0 && D % 4 == 0, MSTR_ERR_BAD_ARG, "mstr_embed_semantics: bad size");
  MSTR_REQUIRE(aligned16(table) && aligned16(out), MSTR_ERR_BAD_ARG, "mstr_embed_semantics: 16-byte alignment");
  const int64_t warps = (int64_t)B * n;
  const int blocks = (int)ceil_div64(warps * 32, 256);
  embed_semantics_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ids, table, ious, out, B, n, D, rows, mode);
  MSTR_LAUNCH_CHECK("embed_semantics_kernel");
  return MSTR_OK;
}
