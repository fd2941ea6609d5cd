// State-broadcast micro-benchmark behind recurrent.cuh: after a grid barrier every one of the 128 CTAs reads the
// SAME [rows x K] fp32 state block from L2 with ld.global.cg (the phase exchange of the persistent chain kernels).
//   mode 0: barrier only
//   mode 1: every CTA reads the same 64 KB (32 rows x 2 x 256 floats), thread = (row tid/8, 8-way K slice)
//   mode 2: every CTA reads its OWN 64 KB (same traffic, no shared lines)
//   mode 3: every CTA reads the same 32 KB (half the rows twice as many CTAs would share)
//   mode 4: mode 1 with the row order rotated by the CTA index (de-synchronises the hot lines)
//   mode 5: every CTA reads the same 16 KB
//   mode 7: packed writes (each CTA stores 4 full 128-byte lines), reads as mode 1;  mode 8: scattered writes,
//           linear reads;  mode 9: packed writes, linear reads;  mode 10: mode 6 with ~1600 cycles of ALU work
//           between the stores and the barrier (is the cost the write acknowledgement?)
//   mode 6: mode 1, preceded by the scattered 4-byte stores of the real kernel (each CTA writes its 4 columns)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_bcast profiles/ubench_bcast.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ldcg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {}
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k(int mode, int iters, unsigned* sync, float* buf, long long* out) {
  unsigned target = 0;
  float acc = 0.f;
  const int tid = threadIdx.x, n = tid >> 3, kp = tid & 7;
  long long tl = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float* base = buf + (size_t)(i & 1) * (1 << 22);     // two alternating 16 MB regions
    if ((mode == 6 || mode == 8 || mode == 10) && kp < 4) base[(size_t)n * 512 + (kp >> 1) * 256 + blockIdx.x * 2 + (kp & 1)] = acc + i;
    if ((mode == 7 || mode == 9) && tid < 128) base[(size_t)blockIdx.x * 128 + tid] = acc + i;      // packed: 4 full lines per CTA
    if (mode == 10) { float d = acc; for (int z = 0; z < 400; ++z) d = fmaf(d, 1.0001f, 0.5f); acc += d * 1e-30f; }
    grid_sync(sync, target, gridDim.x);
    const long long a = clock64();
    if (mode >= 1) {
      const float* src = base;
      int row = n;
      if (mode == 2) src += (size_t)blockIdx.x * 16384;
      if (mode == 4) row = (n + blockIdx.x) & 31;
      float4 v[16];
      const int nld = mode == 3 ? 8 : (mode == 5 ? 4 : 16);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q < nld) v[q] = (mode == 8 || mode == 9) ? ldcg4(src + (size_t)(q * 256 + tid) * 4)     // linear
                                                     : ldcg4(src + (size_t)row * 512 + (q * 8 + kp) * 4);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q < nld) acc += v[q].x + v[q].y + v[q].z + v[q].w;
    }
    tl += clock64() - a;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = tl; }
  if (acc == 123.f) out[0] = 0;
}

int main() {
  unsigned* sync; float* buf; long long* out;
  cudaMalloc(&sync, 4096); cudaMalloc(&buf, 2 * (size_t)(1 << 22) * 4); cudaMallocManaged(&out, 148 * 16);
  cudaMemset(buf, 0, 2 * (size_t)(1 << 22) * 4);
  for (int grid : {64, 128}) {
    for (int mode = 0; mode < 11; ++mode) {
      int iters = 1000;
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(sync, 0, 4096);
        void* args[] = {&mode, &iters, &sync, &buf, &out};
        cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(256), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (rep == 1)
          printf("grid=%3d mode=%d: %7.1f clk/iter, loads %7.1f clk (thread 0 of CTA 0)  [%s]\n", grid, mode,
                 (double)out[0] / iters, (double)out[1] / iters, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
