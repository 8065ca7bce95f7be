// Shared host/device helpers for the mstr_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mstr_b200.h"

namespace mstr {

// ---- host-side error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MSTR_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) {                                    \
      ::mstr::set_error(__VA_ARGS__);                 \
      return (code);                                  \
    }                                                 \
  } while (0)

#define MSTR_CUDA(expr)                                                             \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::mstr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                        __FILE__, __LINE__);                                        \
      return MSTR_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

// Check the launch that was just enqueued (no sync: only configuration errors show up here).
#define MSTR_LAUNCH_CHECK(name)                                                      \
  do {                                                                               \
    ::mstr::count_launch();                                                          \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::mstr::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
      return MSTR_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device), thread safe.
// `done_mask` is a per-kernel bit mask of the device ordinals already configured.
int ensure_dynamic_smem(const void* kernel, int bytes, unsigned long long* done_mask);

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

__device__ __forceinline__ float sigmoid_accurate(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float load(const float* p) { return *p; }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

}  // namespace mstr
