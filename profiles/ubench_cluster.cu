// Cluster / DSMEM exchange micro-benchmark behind recurrent_cluster.cuh: what does one phase exchange of a recurrent
// chain cost when the CTAs that share the state are one thread-block cluster (distributed shared memory) instead of a
// cooperative grid exchanging through L2 (profiles/ubench_gridbar.cu: 2400 cycles per barrier + 1200-2400 per re-read)?
//   mode 0: barrier.cluster.arrive.release + barrier.cluster.wait.acquire, nothing else
//   mode 1: all-gather: every CTA stores its `tile` bytes into ALL CTAs' buffers with st.shared::cluster.v4.f32,
//           barrier.cluster, then every thread reads 16 bytes of the gathered buffer (double-buffered by parity)
//   mode 2: the same all-gather with st.async ... mbarrier::complete_tx::bytes and a local mbarrier wait (no cluster barrier)
//   mode 3: reduce-scatter shape: every CTA stores a distinct 128-byte slice into each peer (4-byte stores), barrier.cluster,
//           every CTA sums the C slices it received
//   mode 4: mode 1 with the pull model: store locally, barrier.cluster, ld.shared::cluster from every peer
//   mode 5: all-gather with ONE cp.async.bulk.shared::cluster.shared::cta per destination (thread d copies this CTA's tile
//           into CTA d; completion on the destination's mbarrier), local mbarrier wait -- no per-word stores at all
//   mode 6: mode 5 with a 256-byte tile, mode 7: mode 5 with a 1024-byte tile, mode 8: 128-byte tile (reduce-scatter slices)
// Reported: SM cycles per exchange (clock64 of thread 0 of CTA 0 over `iters` exchanges), for cluster sizes 8 and 16,
// with 1 and with `maxc` clusters resident; and cudaOccupancyMaxActiveClusters for the launch configuration.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_cluster profiles/ubench_cluster.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_v4(unsigned addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(unsigned addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cluster_v4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_async_v4(unsigned addr, float4 v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_to_cluster(unsigned dst_cluster_addr, unsigned src_cta_addr, unsigned bytes, unsigned mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (long long spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1ll << 24)) __trap();
  }
}

constexpr int kThreads = 256;
constexpr int kMaxC = 16;
constexpr int kTileFloats = 128;          // 512 bytes per CTA per exchange (R = 8 rows x 16 units)
constexpr int kMaxTileFloats = 256;

template <int C>
__global__ void __launch_bounds__(kThreads) bench_kernel(int mode, int iters, long long* out, float* sink) {
  __shared__ __align__(16) float buf[2][kMaxC * kMaxTileFloats];  // gathered state, double-buffered
  __shared__ __align__(16) float mine[kMaxTileFloats];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x;
  const unsigned rank = cluster_rank();
  if (tid == 0) { mbar_init(smem_u32(&mbar[0]), 1); mbar_init(smem_u32(&mbar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = tid; i < 2 * kMaxC * kMaxTileFloats; i += kThreads) (&buf[0][0])[i] = 0.0f;
  if (tid < kMaxTileFloats) mine[tid] = (float)(rank * 1000 + tid);
  const int bulk_floats = mode == 6 ? 64 : (mode == 7 ? 256 : (mode == 8 ? 32 : 128));
  __syncthreads();
  cl_arrive(); cl_wait();
  if (mode == 2 && tid == 0) { mbar_expect(smem_u32(&mbar[0]), C * kTileFloats * 4); }
  if (mode >= 5 && tid == 0) { mbar_expect(smem_u32(&mbar[0]), C * bulk_floats * 4); }
  float acc = 0.0f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    if (mode == 0) {
      cl_arrive(); cl_wait();
    } else if (mode == 1) {
      // 32 float4 per destination x C destinations, 2 (C = 16) or 1 (C = 8) per thread
      for (int e = tid; e < C * (kTileFloats / 4); e += kThreads) {
        const int dst = e / (kTileFloats / 4), q = e % (kTileFloats / 4);
        float4 v = reinterpret_cast<const float4*>(mine)[q];
        v.x += acc;
        st_cluster_v4(mapa(smem_u32(&buf[par][rank * kTileFloats + q * 4]), dst), v);
      }
      cl_arrive(); cl_wait();
      const float4 r = reinterpret_cast<const float4*>(&buf[par][0])[tid % (C * kTileFloats / 4)];
      acc = r.x * 1e-9f;
    } else if (mode == 2) {
      if (tid == 0 && it + 1 < iters) mbar_expect(smem_u32(&mbar[par ^ 1]), C * kTileFloats * 4);   // arm the next phase's barrier
      for (int e = tid; e < C * (kTileFloats / 4); e += kThreads) {
        const int dst = e / (kTileFloats / 4), q = e % (kTileFloats / 4);
        float4 v = reinterpret_cast<const float4*>(mine)[q];
        v.x += acc;
        st_async_v4(mapa(smem_u32(&buf[par][rank * kTileFloats + q * 4]), dst), v, mapa(smem_u32(&mbar[par]), dst));
      }
      mbar_wait(smem_u32(&mbar[par]), (it >> 1) & 1);
      const float4 r = reinterpret_cast<const float4*>(&buf[par][0])[tid % (C * kTileFloats / 4)];
      acc = r.x * 1e-9f;
      __syncthreads();   // every thread has read the buffer before this CTA sends again (WAR two phases later is covered by the protocol)
    } else if (mode == 3) {
      // 32 floats to each peer with 4-byte stores
      for (int e = tid; e < C * 32; e += kThreads) {
        const int dst = e / 32, q = e % 32;
        st_cluster_f32(mapa(smem_u32(&buf[par][rank * 32 + q]), dst), mine[q] + acc);
      }
      cl_arrive(); cl_wait();
      float s = 0.0f;
      if (tid < 32)
        for (int c = 0; c < C; ++c) s += buf[par][c * 32 + tid];
      acc = s * 1e-9f;
    } else if (mode == 4) {
      if (tid < kTileFloats / 4) reinterpret_cast<float4*>(&buf[par][rank * kTileFloats])[tid] = reinterpret_cast<const float4*>(mine)[tid];
      cl_arrive(); cl_wait();
      float s = 0.0f;
      for (int e = tid; e < C * (kTileFloats / 4); e += kThreads) {
        const int src = e / (kTileFloats / 4), q = e % (kTileFloats / 4);
        const float4 r = ld_cluster_v4(mapa(smem_u32(&buf[par][src * kTileFloats + q * 4]), src));
        s += r.x;
      }
      acc = s * 1e-9f;
    } else {
      // stage (generic proxy) -> fence -> one bulk copy per destination
      if (tid == 0 && it + 1 < iters) mbar_expect(smem_u32(&mbar[par ^ 1]), C * bulk_floats * 4);
      if (tid < bulk_floats) mine[tid] = (float)tid + acc;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid < C)
        bulk_copy_to_cluster(mapa(smem_u32(&buf[par][rank * bulk_floats]), tid), smem_u32(mine), bulk_floats * 4, mapa(smem_u32(&mbar[par]), tid));
      mbar_wait(smem_u32(&mbar[par]), (it >> 1) & 1);
      const float4 r = reinterpret_cast<const float4*>(&buf[par][0])[tid % (C * bulk_floats / 4)];
      acc = r.x * 1e-9f;
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  cl_arrive(); cl_wait();
  if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

template <int C>
static void run(int nclusters, int iters, long long* d_out, float* d_sink) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * C);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (C > 8) cudaFuncSetAttribute(bench_kernel<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int maxc = -1;
  cudaError_t eo = cudaOccupancyMaxActiveClusters(&maxc, bench_kernel<C>, &cfg);
  printf("cluster size %2d, %d clusters: cudaOccupancyMaxActiveClusters = %d (%s)\n", C, nclusters, maxc, cudaGetErrorString(eo));
  const char* names[9] = {"barrier.cluster only", "all-gather st.shared::cluster.v4 + barrier", "all-gather st.async + mbarrier",
                          "reduce-scatter 4-byte stores + barrier", "all-gather pull (ld.shared::cluster) + barrier",
                          "all-gather bulk copy 512 B + mbarrier", "all-gather bulk copy 256 B + mbarrier",
                          "all-gather bulk copy 1024 B + mbarrier", "all-gather bulk copy 128 B + mbarrier"};
  // how many clusters of this size are co-resident when a CTA takes a whole SM (180 KB of dynamic shared memory)?
  {
    cudaLaunchConfig_t big = cfg;
    big.dynamicSmemBytes = 180 * 1024;
    cudaFuncSetAttribute(bench_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
    int mc = -1;
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&mc, bench_kernel<C>, &big);
    printf("  with 180 KB dynamic shared memory per CTA (one CTA per SM): max active clusters = %d (%s)\n", mc, cudaGetErrorString(e2));
  }
  for (int mode = 0; mode < 9; ++mode) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<C>, mode, iters, d_out, d_sink);
      if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  run failed: %s\n", cudaGetErrorString(e)); exit(1); }
      cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    }
    printf("  mode %d  %-48s %8.1f cycles per exchange\n", mode, names[mode], (double)h / iters);
  }
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 64);
  const int iters = 2000;
  run<8>(1, iters, d_out, d_sink);
  run<8>(16, iters, d_out, d_sink);
  run<16>(1, iters, d_out, d_sink);
  run<16>(8, iters, d_out, d_sink);
  run<16>(16, iters, d_out, d_sink);     // more clusters than GPCs: shows whether clusters queue (they must not deadlock)
  return 0;
}
