// Shared helpers for the sos-b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SOS_OK 0
#define SOS_ERR_ARG -1
#define SOS_ERR_CUDA -2
#define SOS_ERR_UNSUPPORTED -3

void sos_set_error(const char* fmt, ...);

#define SOS_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      sos_set_error(__VA_ARGS__);                 \
      return SOS_ERR_ARG;                         \
    }                                             \
  } while (0)

#define SOS_CHECK_LAUNCH(name)                                                  \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      sos_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return SOS_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

static inline int sos_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container.  tcgen05 kind::tf32 TRUNCATES the low 13
// mantissa bits of its operands, a systematic -2^-12 relative bias per operand; producers of tensor-core operands
// round here instead so that the MMA sees exactly representable values.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// Half-precision operand scaling: the power of two s that brings a tensor's RMS to ~1 (fp16 keeps its 11 significant bits
// between 6.1e-5 and 65504; max <= sqrt(n) * RMS, so n < 2^32 elements cannot overflow).  Returns the exponent e, s = 2^e.
__device__ __forceinline__ int half_scale_exp(float sumsq, float count) {
  const float ms = sumsq / count;
  if (!(ms > 0.f) || !(ms < 3e38f)) return 0;
  int e = (int)rintf(-0.5f * log2f(ms));
  return max(-100, min(100, e));
}
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((e + 127) << 23); }
// Two floats -> packed half2, round to nearest even, SATURATING at +-65504: an out-of-range activation must not turn into an
// infinity (and then NaNs) inside a tensor-core operand.
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first source operand -> upper half
  return r;
}
#define SOS_ACT_MASK 15
#define SOS_ACT_ROUND_TF32 16

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of `v`; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < (int)((blockDim.x + 31) >> 5)) ? smem32[lane] : 0.f;
    v = warp_sum(v);
  }
  __syncthreads();
  return v;
}
