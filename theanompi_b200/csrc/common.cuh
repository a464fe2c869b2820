// Common helpers for the sm_100a extension (torch-free: raw pointers + cudaStream_t).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <atomic>
#include <stdexcept>
#include <string>

namespace tmpi {

// every launcher bumps this; Python reads it for bench.py's "gpu_launches"
extern std::atomic<unsigned long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

inline void check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    throw std::runtime_error(std::string("tmpi_native: ") + what + ": " + cudaGetErrorString(e));
  }
}
#define TMPI_CHECK_LAUNCH(name) ::tmpi::check_cuda(cudaGetLastError(), name)

// TMPI_DEBUG_CAPTURE=1: after every launch verify that an ongoing stream capture is still valid and name the
// first op that invalidated it (CUDA only reports "a previous error" at capture end).
inline void check_capture(cudaStream_t st, const char* name) {
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("TMPI_DEBUG_CAPTURE"); dbg = (e && e[0] == '1') ? 1 : 0; }
  if (!dbg) return;
  cudaStreamCaptureStatus s = cudaStreamCaptureStatusNone;
  cudaError_t e = cudaStreamIsCapturing(st, &s);
  if (e != cudaSuccess || s == cudaStreamCaptureStatusInvalidated)
    throw std::runtime_error(std::string("tmpi_native: stream capture invalidated at/before ") + name + ": " + cudaGetErrorString(e));
}

constexpr int kArenaBlock = 1024;   // must match parallel/arena.py BLOCK
constexpr int kMaxGroups = 8;

struct GroupTable {
  float lr_mult[kMaxGroups];
  float wd[kMaxGroups];
  int exch[kMaxGroups];
};

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f_to_bf16(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of 8 bf16
struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

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
#endif

// TMPI_DETERMINISTIC=1: bit-reproducible training steps — no split-K (gradient tiles are otherwise combined with fp32 atomics in
// arrival order) and every cross-CTA atomic reduction (bias gradients, batch-norm statistics) collapses to one CTA per channel
// group with a fixed summation order.  Slower; meant for debugging / regression runs.
inline bool deterministic_mode() {
  static const bool v = [] { const char* e = getenv("TMPI_DETERMINISTIC"); return e && e[0] == '1'; }();
  return v;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace tmpi
