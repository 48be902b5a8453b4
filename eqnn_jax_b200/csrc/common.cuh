// Shared helpers for the eqnn_b200 C-ABI library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/eqnn_b200.h"

int eqnn_fail(int code, const char* msg);  // sets the thread-local error string, returns code
void eqnn_count_launches(int n);           // process-wide count of kernels enqueued by this library

#define EQNN_CHECK_ARG(cond, msg)                        \
  do {                                                   \
    if (!(cond)) return eqnn_fail(EQNN_ERR_ARG, msg);    \
  } while (0)

#define EQNN_CHECK_LAUNCH_N(n)                                           \
  do {                                                                   \
    eqnn_count_launches(n);                                              \
    cudaError_t _e = cudaPeekAtLastError();                              \
    if (_e != cudaSuccess) return eqnn_fail(EQNN_ERR_CUDA, cudaGetErrorString(_e)); \
  } while (0)

#define EQNN_CUDA(call)                                                  \
  do {                                                                   \
    cudaError_t _e = (call);                                             \
    if (_e != cudaSuccess) return eqnn_fail(EQNN_ERR_CUDA, cudaGetErrorString(_e)); \
  } while (0)

#define EQNN_CHECK_LAUNCH() EQNN_CHECK_LAUNCH_N(1)

static inline cudaStream_t as_stream(eqnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// gelu (tanh approximation, flax.linen.gelu default) written as x * sigmoid(2u),
// u = sqrt(2/pi) (x + 0.044715 x^3); identical to 0.5 x (1 + tanh u).
// ex2.approx + rcp.approx: relative error ~1e-7, well inside the 1e-5 parity budget
__device__ __forceinline__ float sigmoid_acc(float t) { return __fdividef(1.0f, 1.0f + __expf(-t)); }
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u2 = 2.0f * k0 * (x + k1 * x * x * x);
  return x * sigmoid_acc(u2);
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float u2 = 2.0f * k0 * (x + k1 * x * x2);
  float s = sigmoid_acc(u2);
  float du2 = 2.0f * k0 * (1.0f + 3.0f * k1 * x2);
  return s + x * s * (1.0f - s) * du2;
}
// gelu and its derivative from ONE sigmoid (one exponential instead of two)
__device__ __forceinline__ void gelu_tanh_both(float x, float& g, float& gp) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float u2 = 2.0f * k0 * (x + k1 * x * x2);
  const float s = sigmoid_acc(u2);
  const float du2 = 2.0f * k0 * (1.0f + 3.0f * k1 * x2);
  g = x * s;
  gp = s + x * s * (1.0f - s) * du2;
}
