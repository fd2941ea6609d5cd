// Grid-barrier micro-benchmark behind recurrent.cuh: cost per barrier of 128 co-resident CTAs x 256 threads.
//   mode 0: atomic counter (fence, atomicAdd, acquire-load spin)        -- cooperative-groups style
//   mode 1: per-CTA flags, volatile polling by warp 0, fence before store and after poll
//   mode 2: mode 1 without any fence (NOT correct, timing only: isolates the fence cost)
//   mode 3: mode 1 + payload: every thread stores one float before and ld.cg-loads 8 floats after the barrier
//   mode 4: red.release.gpu + volatile poll (the form recurrent.cuh uses)
//   mode 5: red.release.gpu + ld.relaxed.gpu poll
//   mode 6: four arrival counters on separate 128-byte lines (CTA b -> counter b & 3), lanes 0..3 of warp 0 poll one each
//   mode 7: mode 4 with __nanosleep(32) between polls
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_gridbar profiles/ubench_gridbar.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bar_counter(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
  }
  __syncthreads();
}
template <bool FENCE>
__device__ __forceinline__ void bar_flags(unsigned* flags, unsigned& epoch, unsigned n) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      if (FENCE) __threadfence();
      *reinterpret_cast<volatile unsigned*>(flags + blockIdx.x) = epoch;
    }
    bool done;
    do {
      done = true;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned idx = threadIdx.x + 32u * i;
        if (idx < n) done = done && (*reinterpret_cast<volatile unsigned*>(flags + idx) >= epoch);
      }
      done = __all_sync(0xffffffffu, done);
    } while (!done);
    if (FENCE) __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void bar_red(unsigned* counter, unsigned& target, unsigned n, int variant) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    if (variant == 0) {
      while (*reinterpret_cast<volatile unsigned*>(counter) < target) {}
    } else if (variant == 1) {
      unsigned v;
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    } else {
      while (*reinterpret_cast<volatile unsigned*>(counter) < target) __nanosleep(32);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void bar_red4(unsigned* counters, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x < 32) {
    target += n / 4;
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counters + (blockIdx.x & 3) * 32) : "memory");
    bool done;
    do {
      done = threadIdx.x >= 4 || *reinterpret_cast<volatile unsigned*>(counters + threadIdx.x * 32) >= target;
      done = __all_sync(0xffffffffu, done);
    } while (!done);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k(int mode, int iters, unsigned* sync, float* buf, long long* out) {
  unsigned st = 0;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 3) buf[(size_t)(i & 1) * gridDim.x * 256 + blockIdx.x * 256 + threadIdx.x] = acc + i;
    if (mode == 4) bar_red(sync, st, gridDim.x, 0);
    else if (mode == 5) bar_red(sync, st, gridDim.x, 1);
    else if (mode == 6) bar_red4(sync, st, gridDim.x);
    else if (mode == 7) bar_red(sync, st, gridDim.x, 2);
    else if (mode == 0) bar_counter(sync, st, gridDim.x);
    else if (mode == 2) bar_flags<false>(sync, st, gridDim.x);
    else bar_flags<true>(sync, st, gridDim.x);
    if (mode == 3) {
      const float* src = buf + (size_t)(i & 1) * gridDim.x * 256 + ((blockIdx.x + 1) % gridDim.x) * 256;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += __ldcg(src + ((threadIdx.x + 32 * q) & 255));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.f) out[0] = 0;
}

int main() {
  unsigned* sync; float* buf; long long* out;
  cudaMalloc(&sync, 4096); cudaMalloc(&buf, 2 * 148 * 256 * 4); cudaMallocManaged(&out, 148 * 8);
  for (int grid : {32, 128}) {
    for (int mode = 0; mode < 8; ++mode) {
      int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(sync, 0, 4096);
        void* args[] = {&mode, &iters, &sync, &buf, &out};
        cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(256), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (rep == 1) printf("grid=%3d mode=%d: %7.1f clk/barrier  [%s]\n", grid, mode, (double)out[0] / iters, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
