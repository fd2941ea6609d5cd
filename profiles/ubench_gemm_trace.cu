// Where a small tcgen05 GEMM launch spends its time: clock64 stamps of CTA (0,0,0) of tc_gemm_kernel<EPI_PLAIN> on the
// weight-gradient shape of the train step (dW[768,256] = dG^T[608,768] . X[608,256], both operands MN-major).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_gemm_trace profiles/ubench_gemm_trace.cu
#include <cstdio>
#include <vector>
#include "../acvae_b200/csrc/train.cuh"
namespace acvae { thread_local char g_err[512] = {0}; std::atomic<unsigned long long> g_launches{0}; KernelProbe g_probe{}; }
using namespace acvae;
int main() {
  const int M = 768, U = 256, R = 608;
  float *dy, *x, *dw; long long* trd; long long tr[16];
  cudaMalloc(&dy, (size_t)R * M * 4); cudaMalloc(&x, (size_t)R * U * 4); cudaMalloc(&dw, (size_t)M * U * 4);
  cudaMemset(dy, 0, (size_t)R * M * 4); cudaMemset(x, 0, (size_t)R * U * 4);
  cudaMalloc(&trd, 16 * 8);
  const char* names[] = {"entry", "setup done (barriers, TMEM alloc, sync)", "first TMA stage landed", "first stage split",
                         "last chunk committed", "accumulators drained", "staged to smem", "epilogue done", "TMEM freed"};
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemset(trd, 0, 16 * 8);
    tc_trace_ptr() = rep >= 2 ? trd : nullptr;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    int rc = linear_bwd_weight(M, U, R, dy, M, x, U, dw, U, 0);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(tr, trd, 16 * 8, cudaMemcpyDeviceToHost);
    printf("rep %d rc=%d (%s) %s: %.1f us (events around memset + launch)\n", rep, rc, g_err, cudaGetErrorString(e), ms * 1e3);
    if (rep >= 2)
      for (int i = 1; i < 9; ++i) printf("  %-44s +%6lld cycles (at %6lld)\n", names[i], tr[i] - tr[i - 1], tr[i] - tr[0]);
  }
  return 0;
}
