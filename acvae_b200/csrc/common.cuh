// Shared helpers for the acvae_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

namespace acvae {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

// Kernel probe (acvae_set_kernel_probe): CUDA events recorded on the launching stream right before / after every
// launch whose kernel name contains `name`.  bench.py uses it to time ONE kernel live inside a step.
struct KernelProbe {
  char name[96];
  cudaEvent_t e0, e1;
  int active;
  int hits;
};
extern KernelProbe g_probe;
inline bool probe_match(const char* kernel) { return g_probe.active && strstr(kernel, g_probe.name) != nullptr; }

inline int set_error(const char* what, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), "%s%s%s", what, detail[0] ? ": " : "", detail);
  return -1;
}

// Every kernel launch in the library goes through this macro: it counts the
// launch (bench.py's `gpu_launches`) and turns a launch error into a C-ABI
// error code instead of an exception.
#define ACVAE_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
  do {                                                                              \
    const bool probe__ = acvae::probe_match(#kernel);                               \
    if (probe__) cudaEventRecord(acvae::g_probe.e0, (stream));                      \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                     \
    if (probe__) { cudaEventRecord(acvae::g_probe.e1, (stream)); ++acvae::g_probe.hits; } \
    acvae::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    cudaError_t e__ = cudaPeekAtLastError();                                        \
    if (e__ != cudaSuccess) return acvae::set_error(#kernel, cudaGetErrorString(e__)); \
  } while (0)

#define ACVAE_CHECK(expr)                                                           \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) return acvae::set_error(#expr, cudaGetErrorString(e__)); \
  } while (0)

#define ACVAE_TRY(expr)            \
  do {                             \
    int r__ = (expr);              \
    if (r__ != 0) return r__;      \
  } while (0)

#define ACVAE_REQUIRE(cond, msg)                        \
  do {                                                  \
    if (!(cond)) return acvae::set_error("invalid argument", msg); \
  } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// tanh for the additive-attention scores (attn_model.py:32): T = Te*A evaluations per query row make the
// attention kernels MUFU / issue bound, and libdevice tanhf costs ~100 issue slots per warp there (branchy, two
// paths).  tanh(x) = 1 - 2 / (e^{2x} + 1) with one ex2.approx and one rcp.approx: five instructions, no clamp needed
// (e^{2x} = inf gives 1, e^{2x} = 0 gives -1), absolute error < 2e-7 over the whole range (the final subtraction only
// loses RELATIVE accuracy near 0, where tanh itself is ~x).  Used consistently by the forward, its backward and the
// sampling kernels.
constexpr float kTwoLog2e = 2.885390081777927f;      // 2 * log2(e)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float attn_tanh(float x) {
  return fmaf(-2.0f, rcp_approx(ex2_approx(x * kTwoLog2e) + 1.0f), 1.0f);
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

// Block-wide reductions for blockDim.x <= 1024 (multiple of 32); `red` is a
// 33-float shared scratch.  All threads receive the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float x = lane < nw ? red[lane] : 0.0f;
    x = warp_sum(x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float x = lane < nw ? red[lane] : -INFINITY;
    x = warp_max(x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}

// Per-device caches: function attributes (cudaFuncSetAttribute), side streams and capability probes belong to the device
// that was current when they were set; every such cache in the library is an array indexed by the current device.
constexpr int kMaxDevices = 16;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-supplied workspace.
struct Arena {
  char* base;
  size_t off = 0;
  explicit Arena(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(n * sizeof(T));
    return r;
  }
};

}  // namespace acvae
