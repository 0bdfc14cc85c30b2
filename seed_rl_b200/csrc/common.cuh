// Shared helpers for libseedrl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/seedrl_b200.h"

namespace seedrl {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;

inline int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}



#define SEEDRL_CHECK_ARG(cond, msg)                                         \
  do {                                                                      \
    if (!(cond))                                                            \
      return ::seedrl::set_error(SEEDRL_ERR_INVALID_ARGUMENT,               \
                                 std::string(__func__) + ": " + (msg));     \
  } while (0)

#define SEEDRL_CHECK_LAUNCH()                                               \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess)                                                 \
      return ::seedrl::set_error(                                           \
          SEEDRL_ERR_INTERNAL, std::string(__func__) + ": CUDA launch: " +  \
                                   cudaGetErrorString(e__));                \
  } while (0)

#define SEEDRL_CUDA(call)                                                   \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess)                                                 \
      return ::seedrl::set_error(                                           \
          SEEDRL_ERR_INTERNAL,                                              \
          std::string(__func__) + ": " #call ": " + cudaGetErrorString(e__)); \
  } while (0)

// Optional per-category kernel timing (CUDA events on the launching stream), used by
// bench.py's separate profiling pass -- never inside a timed region.
enum ProfCat { PC_CONV_FWD = 0, PC_CONV_DGRAD, PC_CONV_WGRAD, PC_POOL, PC_GEMM, PC_LSTM_PW,
               PC_LOSS, PC_ADAM, PC_VTRACE, PC_MISC, PC_COUNT };
extern bool g_prof_on;
extern int g_conv_cat;
// One event AFTER each launch; a kernel's time = gap to the previous event on the stream.
void prof_mark_(int cat, cudaStream_t st);

inline void count_launch(int cat = PC_MISC, cudaStream_t st = 0) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on) prof_mark_(cat, st);
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t ceil_div_sz(size_t a, size_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace seedrl
