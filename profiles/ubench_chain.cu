// In-kernel phase timing of the persistent posterior chain (recurrent.cuh): clock64 stamps of CTA 0 / thread 0.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_chain profiles/ubench_chain.cu
#include <cstdio>
#include <vector>
#include "../acvae_b200/csrc/recurrent.cuh"
namespace acvae { thread_local char g_err[512] = {0}; std::atomic<unsigned long long> g_launches{0}; KernelProbe g_probe{}; }
using namespace acvae;
int main() {
  const int N = 32, T = 19, E = kChainE;
  PostChainFwd p{};
  p.N = N; p.T = T;
  auto alloc = [](size_t n) { float* q; cudaMalloc(&q, n * 4); cudaMemset(q, 0, n * 4); return q; };
  for (int d = 0; d < 2; ++d) { p.gx[d] = alloc((size_t)N * T * 3 * E); p.whh[d] = alloc(3 * E * E); p.bhh[d] = alloc(3 * E); p.gq[d] = alloc((size_t)N * T * 4 * E); }
  int* lens; cudaMalloc(&lens, N * 4); std::vector<int> hl(N, T); cudaMemcpy(lens, hl.data(), N * 4, cudaMemcpyHostToDevice);
  p.lens = lens; p.ho = alloc((size_t)N * T * 2 * E);
  unsigned* bar; cudaMalloc(&bar, 1024); p.bar = bar;
  long long* tr; cudaMallocManaged(&tr, T * 8 * 8); p.trace = tr;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(bar, 0, 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    int rc = launch_chain(post_chain_fwd_kernel, 0, 0, "post", p);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("rep %d rc=%d %s: %.1f us total, %.2f us/step\n", rep, rc, cudaGetErrorString(e), ms * 1e3, ms * 1e3 / T);
  }
  printf("step: loads  fma  reduce  pointwise+store  grid_sync  (cycles, thread 0 of CTA 0)\n");
  for (int s = 1; s < T - 1; ++s)
    printf("%2d: %6lld %6lld %6lld %6lld %6lld   step total %6lld\n", s, tr[s*8+5]-tr[s*8+0], tr[s*8+1]-tr[s*8+5], tr[s*8+2]-tr[s*8+1], tr[s*8+3]-tr[s*8+2], tr[s*8+4]-tr[s*8+3], tr[s*8+4]-tr[s*8+0]);
  return 0;
}
