// Grid-barrier micro-benchmark behind recurrent.cuh: cost per barrier of 128 co-resident CTAs x 256 threads.
//   mode 0: atomic counter (fence, atomicAdd, acquire-load spin)        -- cooperative-groups style
//   mode 1: per-CTA flags, volatile polling by warp 0, fence before store and after poll
//   mode 2: mode 1 without any fence (NOT correct, timing only: isolates the fence cost)
//   mode 3: mode 1 + payload: every thread stores one float before and ld.cg-loads 8 floats after the barrier
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

__global__ void __launch_bounds__(256) k(int mode, int iters, unsigned* sync, float* buf, long long* out) {
  unsigned st = 0;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 3) buf[(size_t)(i & 1) * gridDim.x * 256 + blockIdx.x * 256 + threadIdx.x] = acc + i;
    if (mode == 0) bar_counter(sync, st, gridDim.x);
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
    for (int mode = 0; mode < 4; ++mode) {
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
