// Shared device/host helpers for libsnb (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/snb.h"

namespace snb {

// ---- host-side error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SNB_CHECK_ARG(cond, code, ...)      \
  do {                                      \
    if (!(cond)) {                          \
      snb::set_error(__VA_ARGS__);          \
      return (code);                        \
    }                                       \
  } while (0)

#define SNB_CUDA(expr)                                    \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return snb::cuda_fail(_e, #expr); \
  } while (0)

void count_launch();

inline int launch_status(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return 0;
}

int num_sms();

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(__expf(x)); }

}  // namespace snb
