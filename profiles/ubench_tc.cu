// Micro-benchmarks behind the tcgen05 GEMM design (tc_gemm.cuh): what one SM can do, measured with clock64.
//   tma   : 2 x {32 fp32, 128 rows} 128B-swizzled boxes per iteration, S iterations in flight
//   mma   : back-to-back tcgen05.mma kind::tf32 128x128x8 from shared memory, one commit at the end
//   commit: 4 MMAs + tcgen05.commit + mbarrier wait per iteration (round trip of the stage-free handshake)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tc profiles/ubench_tc.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../acvae_b200/csrc/tc_gemm.cuh"

namespace acvae {
thread_local char g_err[512] = {0};
std::atomic<unsigned long long> g_launches{0};
}  // namespace acvae
using namespace acvae;

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                     int stages, int iters, int rows_a, int kblocks, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 6 * 32768);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m0 = (blockIdx.x * 128) % rows_a;
    long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      const int st = i % stages, it = i / stages;
      if (it > 0) mbar_wait(&full[st], (it - 1) & 1);
      if (i < iters) {
        const int kb = (i % kblocks) * 32;
        mbar_expect_tx(&full[st], 32768);
        tma_load_2d(smem + st * 32768, &mapA, &full[st], kb, m0);
        tma_load_2d(smem + st * 32768 + 16384, &mapB, &full[st], kb, 0);
      }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, int per_commit, int nacc, int ncols, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a = smem_u32(smem), b = a + 16384;
    long long t0 = clock64();
    int phase = 0;
    for (int i = 0; i < iters; ++i) {
      const int j = i & 3;
      tc_mma_tf32(tmem + (uint32_t)((i % nacc) * ncols), tc_smem_desc(a + j * 32, 16, 1024, 2), tc_smem_desc(b + j * 32, 16, 1024, 2), idesc, i >= nacc);
      if (per_commit > 0 && (i % per_commit) == per_commit - 1) {
        tc_commit(bar);
        mbar_wait(bar, phase);
        phase ^= 1;
      }
    }
    if (per_commit <= 0) { tc_commit(bar); mbar_wait(bar, 0); }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
}

int main() {
  const int rows = 16384, K = 1024;
  float *A, *B;
  cudaMalloc(&A, (size_t)rows * K * 4);
  cudaMalloc(&B, (size_t)128 * K * 4);
  cudaMemset(A, 0, (size_t)rows * K * 4);
  cudaMemset(B, 0, (size_t)128 * K * 4);
  long long* out;
  cudaMallocManaged(&out, 148 * 8);
  CUtensorMap ma, mb;
  if (!tc_make_map(&ma, A, rows, K, K, 128) || !tc_make_map(&mb, B, 128, K, K, 128)) { printf("map failed\n"); return 1; }
  const int smem = 6 * 32768 + 2048;
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int grid : {1, 16, 148}) {
    for (int stages : {1, 2, 3, 4, 6}) {
      const int iters = 512;
      tma_kernel<<<grid, 128, smem>>>(ma, mb, stages, iters, rows, K / 32, out);
      cudaDeviceSynchronize();
      tma_kernel<<<grid, 128, smem>>>(ma, mb, stages, iters, rows, K / 32, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
      printf("tma grid=%3d stages=%d: %7.1f clk/iter (32 KB)  -> %6.1f B/clk/SM  [%s]\n", grid, stages, (double)mx / iters,
             32768.0 * iters / mx, cudaGetErrorString(e));
    }
  }
  for (int ncols : {128, 256}) {
    for (int nacc : {1, 2, 4}) {
      if (nacc * ncols > 512) continue;
      for (int per : {0, 12}) {
        const int iters = 1200;
        mma_kernel<<<1, 128, 70000>>>(iters, per, nacc, ncols, out);
        cudaDeviceSynchronize();
        mma_kernel<<<1, 128, 70000>>>(iters, per, nacc, ncols, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mma tf32 128x%dx8, %d accumulators round-robin, commit+wait every %2d: %7.1f clk/MMA  [%s]\n", ncols, nacc, per,
               (double)out[0] / iters, cudaGetErrorString(e));
      }
    }
  }
  printf("sm clock attr: %d kHz\n", clk_khz);
  return 0;
}
