// Data-parallel optimizer tail as ONE fused compute + collective over NVLink peer memory (no NCCL on this path).
//
// Reference semantics (runners/pytorch_runner_vae.py:204-207, 321-324): DistributedDataParallel averages the gradients
// over ranks inside loss.backward(); then clip_grad_norm_ on the averaged gradients and Adam.step() on every rank.
// With NCCL that is all-reduce(32 MB) [~155 us at 8 GPUs, serial] + clip/Adam over all 8 M parameters [40 us] per rank.
// Here every rank owns 1/W of the flat buffers (ZeRO-1 style) and the exchange is fused with the arithmetic:
//   dp_reduce_kernel  waits until every peer's gradients are final (flags in peer memory), then reduces ITS shard:
//                     g[i] = (1/W) sum_q G_q[i], the peers' gradients read straight over NVLink (ranks summed in a
//                     fixed order), accumulates the shard's sum of squares and publishes it to every peer;
//   dp_adam_kernel    waits for the W partial norms -> global norm -> clip coefficient, runs Adam on its shard and
//                     writes the updated parameters into EVERY rank's parameter buffer (the all-gather), then signals
//                     "done" and waits for the peers' "done" (nobody starts the next forward on a half-written buffer,
//                     nobody overwrites gradients a peer is still reading).
// Per rank and step: (W-1)/W of 32 MB read and written over NVLink (28 + 28 MB at W = 8), Adam on 1/W of the parameters;
// every rank ends with bit-identical parameters (each shard is computed once and broadcast).
// Buffers are mapped with CUDA IPC (acvae_ipc_export / acvae_ipc_open); one process per GPU, kernels of different
// ranks run on DIFFERENT GPUs and only wait on flags the peers write (what NCCL's kernels do); a protocol bug traps
// after 60 s of wall clock instead of hanging.
#pragma once
#include "optim.cuh"
#include "recurrent.cuh"     // globaltimer_ns

namespace acvae {

constexpr int kDpMaxWorld = 16;
constexpr int kDpBlocks = 592;       // 4 CTAs per SM x 148 SMs
constexpr int kDpThreads = 256;

// Lives in IPC-shared memory of every rank; slot q is written by rank q.
struct DpComm {
  unsigned ready[kDpMaxWorld];       // rank q's gradients of epoch e are final
  unsigned normed[kDpMaxWorld];      // rank q's shard sum of squares of epoch e is in normsq[q]
  unsigned done[kDpMaxWorld];        // rank q has written its parameter shard of epoch e into this rank's buffer
  float normsq[kDpMaxWorld];
};

struct DpParams {
  int world, rank;
  long long n4;                       // float4 words per SHARD
  const float4* grads[kDpMaxWorld];   // every rank's flat gradient buffer (peer-mapped; [rank] is local)
  float4* params[kDpMaxWorld];        // every rank's flat parameter buffer
  DpComm* comm[kDpMaxWorld];          // every rank's communication block
  float4* gshard;                     // [n4] local: the reduced (averaged) gradient shard
  float4 *m, *v;                      // [n4] local: Adam moments of the shard
  float* partial;                     // [kDpBlocks] local scratch
  unsigned* ticket;                   // [2] local, zero between calls
  const float* hyper;                 // device {max_norm, lr, beta1, beta2, eps, weight_decay}
  const int* step;                    // device counter of completed steps (epoch = step + 1)
  float* total_norm;                  // out (may be NULL)
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ void dp_wait_flag(const unsigned* flag, unsigned epoch) {
  const unsigned long long t0 = globaltimer_ns();
  while ((int)(ld_volatile_u32(flag) - epoch) < 0) {
    __nanosleep(100);
    if (globaltimer_ns() - t0 > 60000000000ull) __trap();
  }
}
__device__ __forceinline__ float4 ld_peer4(const float4* p) {
  float4 v;                            // .cv: never served from a stale L1 line (peer memory is not cached in the local L2)
  asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kDpThreads) dp_reduce_kernel(const __grid_constant__ DpParams a) {
  __shared__ float red[33];
  __shared__ bool last;
  const int W = a.world, tid = threadIdx.x;
  const unsigned epoch = (unsigned)a.step[0] + 1u;
  // my gradients are final (they were written by earlier kernels of this stream): tell every peer once
  if (blockIdx.x == 0 && tid < W) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(&a.comm[tid]->ready[a.rank]) = epoch;
  }
  if (tid < W) dp_wait_flag(&a.comm[a.rank]->ready[tid], epoch);
  __syncthreads();
  const float invW = 1.0f / (float)W;
  const long long base = (long long)a.rank * a.n4;
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int q = 0; q < W; ++q) {                     // fixed rank order: the same sum on whichever rank owns the shard
      const float4 g = ld_peer4(a.grads[q] + base + i);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    acc.x *= invW; acc.y *= invW; acc.z *= invW; acc.w *= invW;
    a.gshard[i] = acc;
    s = fmaf(acc.x, acc.x, s); s = fmaf(acc.y, acc.y, s); s = fmaf(acc.z, acc.z, s); s = fmaf(acc.w, acc.w, s);
  }
  s = block_sum_opt(s, red);
  if (tid == 0) {
    a.partial[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(&a.ticket[0], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {                                          // the last CTA adds the partials in a fixed order and publishes the shard norm
    __threadfence();
    float t = 0.0f;
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) t += a.partial[i];
    t = block_sum_opt(t, red);
    if (tid < W) {
      *reinterpret_cast<volatile float*>(&a.comm[tid]->normsq[a.rank]) = t;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(&a.comm[tid]->normed[a.rank]) = epoch;
    }
    if (tid == 0) a.ticket[0] = 0;
  }
}

__global__ void __launch_bounds__(kDpThreads) dp_adam_kernel(const __grid_constant__ DpParams a) {
  __shared__ float red[33];
  __shared__ bool last;
  const int W = a.world, tid = threadIdx.x;
  const int t = a.step[0] + 1;
  const unsigned epoch = (unsigned)t;
  if (tid < W) dp_wait_flag(&a.comm[a.rank]->normed[tid], epoch);
  __syncthreads();
  float tot = 0.0f;
  for (int q = 0; q < W; ++q) tot += *reinterpret_cast<const volatile float*>(&a.comm[a.rank]->normsq[q]);
  const float norm = sqrtf(tot);
  const float max_norm = a.hyper[0], lr = a.hyper[1], b1 = a.hyper[2], b2 = a.hyper[3], eps = a.hyper[4], wd = a.hyper[5];
  float coef = 1.0f;
  if (max_norm > 0.0f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  if (blockIdx.x == 0 && tid == 0 && a.total_norm) a.total_norm[0] = norm;
  const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  auto upd = [&](float& p, float g, float& m, float& v) {
    g *= coef;
    if (wd != 0.0f) g = fmaf(wd, p, g);
    m = fmaf(1.0f - b1, g - m, m);
    v = fmaf(1.0f - b2, g * g, b2 * v);
    p -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  };
  const long long base = (long long)a.rank * a.n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = a.params[a.rank][base + i], g = a.gshard[i], m = a.m[i], v = a.v[i];
    upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
    a.m[i] = m; a.v[i] = v;
#pragma unroll 4
    for (int q = 0; q < W; ++q) a.params[q][base + i] = p;            // the all-gather: my shard into every rank's buffer
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    last = atomicAdd(&a.ticket[1], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    // every CTA of this rank has pushed its stores: tell the peers, then hold the stream until every peer has done the same
    if (tid < W) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(&a.comm[tid]->done[a.rank]) = epoch;
      dp_wait_flag(&a.comm[a.rank]->done[tid], epoch);
    }
    if (tid == 0) a.ticket[1] = 0;
  }
}

}  // namespace acvae
