// Factorised additive (Bahdanau) attention over the frame memory.
//
// Reference: models/attn_model.py:20-46 (`Seq2SeqAttention.forward`) computes
//   tanh(Linear([q ‖ mem_j]))·v  for every frame j at every decode step, i.e. it
// re-projects the whole memory through h2attn each call.  Here the memory half
//   P[c,j,:] = mem[c,j,:]·W[:,Dq:]^T + b          (once per batch, GEMM)
// is precomputed and a call only needs  s_j = v · tanh(P[c,j,:] + q·W[:,:Dq]^T),
// the masked softmax over j (mask value -1e10 => exactly zero weight in fp32)
// and ctx = sum_j w_j mem[c,j,:]   (SURVEY.md Appendix A.3).
//
// Layout: P [clips,Te,A], mem [clips,Te,E] row-major; a CTA owns one query row
// and streams the clip's P and mem rows with fully coalesced loads (a warp reads
// one frame's A (or E) contiguous floats).  HBM/L2-bandwidth bound.
#pragma once
#include "common.cuh"

namespace acvae {

struct AttnFwdParams {
  int rows, Te, A, E, Dq;
  int rows_per_clip;                 // clip(r) = r / rows_per_clip
  const float* qp_in; long long ld_qp_in;      // precomputed q·Wq^T [rows,A] or NULL
  const float* q; long long ld_q;               // query rows (NULL => q = 0)
  const int* q_gather; long long q_gather_stride;  // optional: q row index = q_gather[r*stride] (embedding lookup)
  const float* wq; long long ldwq;              // Wq[a*ldwq + k]
  float* qp_out; long long ld_qp_out;           // saved projection (backward) or NULL
  const float* P; const float* mem; const float* v; const int* mem_lens;
  float* ctx; long long ld_ctx;                 // [rows,E]
  float* w_out; long long ld_w;                 // saved weights [rows,Te] or NULL
  float* aw_out; long long aw_ld_r, aw_ld_j;    // user-visible weights (vae_model.py:868 layout) or NULL
  const int* live;                              // optional device flag: return at once when *live == 0
};

// dynamic smem: A (qp) + Te (scores) + Dq (q) + 33 floats
__global__ void __launch_bounds__(256) attn_fwd_kernel(const __grid_constant__ AttnFwdParams p) {
  extern __shared__ float sm[];
  float* qp = sm;
  float* sc = qp + p.A;
  float* qs = sc + p.Te;
  float* red = qs + p.Dq;
  if (p.live && *p.live == 0) return;
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int clip = r / p.rows_per_clip;
  const int len = min(p.mem_lens[clip], p.Te);

  // 1. query projection
  if (p.qp_in) {
    for (int a = tid; a < p.A; a += blockDim.x) qp[a] = p.qp_in[(long long)r * p.ld_qp_in + a];
  } else if (p.q) {
    const long long qrow = p.q_gather ? (long long)p.q_gather[r * p.q_gather_stride] : (long long)r;
    for (int k = tid; k < p.Dq; k += blockDim.x) qs[k] = p.q[qrow * p.ld_q + k];
    __syncthreads();
    for (int a = wid; a < p.A; a += nw) {
      const float* wr = p.wq + (long long)a * p.ldwq;
      float s = 0.0f;
      for (int k = lane; k < p.Dq; k += 32) s = fmaf(qs[k], __ldg(wr + k), s);
      s = warp_sum(s);
      if (lane == 0) qp[a] = s;
    }
  } else {
    for (int a = tid; a < p.A; a += blockDim.x) qp[a] = 0.0f;
  }
  __syncthreads();
  if (p.qp_out)
    for (int a = tid; a < p.A; a += blockDim.x) p.qp_out[(long long)r * p.ld_qp_out + a] = qp[a];

  // 2. scores: one warp per frame
  const float* Pc = p.P + (long long)clip * p.Te * p.A;
  for (int j = wid; j < len; j += nw) {
    const float* pr = Pc + (long long)j * p.A;
    float s = 0.0f;
    for (int a = lane; a < p.A; a += 32) s = fmaf(__ldg(p.v + a), tanhf(pr[a] + qp[a]), s);
    s = warp_sum(s);
    if (lane == 0) sc[j] = s;
  }
  __syncthreads();

  // 3. masked softmax over valid frames (masked frames get exactly 0)
  float mx = -INFINITY;
  for (int j = tid; j < len; j += blockDim.x) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, red);
  float sum = 0.0f;
  for (int j = tid; j < len; j += blockDim.x) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = 1.0f / sum;
  for (int j = tid; j < p.Te; j += blockDim.x) {
    const float w = j < len ? sc[j] * inv : 0.0f;
    if (j < len) sc[j] = w;
    if (p.w_out) p.w_out[(long long)r * p.ld_w + j] = w;
    if (p.aw_out) p.aw_out[(long long)r * p.aw_ld_r + (long long)j * p.aw_ld_j] = w;
  }
  __syncthreads();

  // 4. context
  const float* mc = p.mem + (long long)clip * p.Te * p.E;
  for (int e = tid; e < p.E; e += blockDim.x) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int j = 0;
    for (; j + 4 <= len; j += 4) {
      a0 = fmaf(sc[j], mc[(long long)j * p.E + e], a0);
      a1 = fmaf(sc[j + 1], mc[(long long)(j + 1) * p.E + e], a1);
      a2 = fmaf(sc[j + 2], mc[(long long)(j + 2) * p.E + e], a2);
      a3 = fmaf(sc[j + 3], mc[(long long)(j + 3) * p.E + e], a3);
    }
    for (; j < len; ++j) a0 = fmaf(sc[j], mc[(long long)j * p.E + e], a0);
    p.ctx[(long long)r * p.ld_ctx + e] = (a0 + a1) + (a2 + a3);
  }
}

inline int launch_attn_fwd(const AttnFwdParams& p, cudaStream_t st) {
  if (p.rows <= 0) return 0;
  const size_t smem = (size_t)(p.A + p.Te + p.Dq + 33) * sizeof(float);
  ACVAE_LAUNCH(attn_fwd_kernel, p.rows, 256, smem, st, p);
  return 0;
}

// ---- backward, in-chain part: d(ctx) -> d(scores) (saved) and d(q·Wq^T) -------
struct AttnBwdQParams {
  int rows, Te, A, E, rows_per_clip;
  const float* dctx; long long ld_dctx;     // [rows,E]
  const float* w; long long ld_w;           // saved softmax weights [rows,Te]
  const float* qp; long long ld_qp;         // saved projections [rows,A]
  const float* P; const float* mem; const float* v; const int* mem_lens;
  float* ds; long long ld_ds;               // out: d(score) [rows,Te]
  float* dqp; long long ld_dqp;             // out: d(q·Wq^T) [rows,A]
};

__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const __grid_constant__ AttnBwdQParams p) {
  extern __shared__ float sm[];
  float* dw = sm;            // Te
  float* dc = dw + p.Te;     // E
  float* red = dc + p.E;     // 33
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int clip = r / p.rows_per_clip;
  const int len = min(p.mem_lens[clip], p.Te);
  for (int e = tid; e < p.E; e += blockDim.x) dc[e] = p.dctx[(long long)r * p.ld_dctx + e];
  __syncthreads();
  const float* mc = p.mem + (long long)clip * p.Te * p.E;
  for (int j = wid; j < len; j += nw) {
    float s = 0.0f;
    for (int e = lane; e < p.E; e += 32) s = fmaf(dc[e], mc[(long long)j * p.E + e], s);
    s = warp_sum(s);
    if (lane == 0) dw[j] = s;
  }
  __syncthreads();
  const float* wr = p.w + (long long)r * p.ld_w;
  float dot = 0.0f;
  for (int j = tid; j < len; j += blockDim.x) dot = fmaf(wr[j], dw[j], dot);
  dot = block_sum(dot, red);
  for (int j = tid; j < p.Te; j += blockDim.x) {
    const float d = j < len ? wr[j] * (dw[j] - dot) : 0.0f;
    if (j < len) dw[j] = d;
    p.ds[(long long)r * p.ld_ds + j] = d;
  }
  __syncthreads();
  const float* Pc = p.P + (long long)clip * p.Te * p.A;
  for (int a = tid; a < p.A; a += blockDim.x) {
    const float q = p.qp[(long long)r * p.ld_qp + a];
    const float va = __ldg(p.v + a);
    float acc = 0.0f;
    for (int j = 0; j < len; ++j) {
      const float th = tanhf(Pc[(long long)j * p.A + a] + q);
      acc = fmaf(dw[j] * va, 1.0f - th * th, acc);
    }
    p.dqp[(long long)r * p.ld_dqp + a] = acc;
  }
}

inline int launch_attn_bwd_q(const AttnBwdQParams& p, cudaStream_t st) {
  if (p.rows <= 0) return 0;
  const size_t smem = (size_t)(p.Te + p.E + 33) * sizeof(float);
  ACVAE_LAUNCH(attn_bwd_q_kernel, p.rows, 256, smem, st, p);
  return 0;
}

// ---- backward, deferred part: accumulate dP, dmem, dv over all rows of a clip --
// grid (ceil(Te/JT), clips); no atomics on dP/dmem: one CTA owns (clip, frame chunk).
struct AttnBwdAccParams {
  int clips, Te, A, E, rows_per_clip;
  const float* ds; long long ld_ds;        // [rows,Te]
  const float* w; long long ld_w;          // [rows,Te]
  const float* qp; long long ld_qp;        // [rows,A]
  const float* dctx; long long ld_dctx;    // [rows,E]
  const float* P; const float* v; const int* mem_lens;
  float* dP;                               // [clips,Te,A] (overwritten)
  float* dmem; int dmem_accumulate;        // [clips,Te,E]
  float* dv;                               // [A], atomically accumulated (zeroed by the caller)
};

constexpr int kAttnJT = 4;

__global__ void __launch_bounds__(256) attn_bwd_acc_kernel(const __grid_constant__ AttnBwdAccParams p) {
  const int clip = blockIdx.y;
  const int j0 = blockIdx.x * kAttnJT;
  const int tid = threadIdx.x;
  const int len = min(p.mem_lens[clip], p.Te);
  const int r0 = clip * p.rows_per_clip;
  // dP and dv: threads over a
  for (int a = tid; a < p.A; a += blockDim.x) {
    const float va = __ldg(p.v + a);
    float dvacc = 0.0f;
    for (int jj = 0; jj < kAttnJT; ++jj) {
      const int j = j0 + jj;
      if (j >= p.Te) break;
      float acc = 0.0f;
      if (j < len) {
        const float pj = p.P[((long long)clip * p.Te + j) * p.A + a];
        for (int i = 0; i < p.rows_per_clip; ++i) {
          const long long r = r0 + i;
          const float d = p.ds[r * p.ld_ds + j];
          const float th = tanhf(pj + p.qp[r * p.ld_qp + a]);
          acc = fmaf(d * va, 1.0f - th * th, acc);
          dvacc = fmaf(d, th, dvacc);
        }
      }
      p.dP[((long long)clip * p.Te + j) * p.A + a] = acc;
    }
    atomicAdd(p.dv + a, dvacc);
  }
  // dmem: threads over e
  for (int e = tid; e < p.E; e += blockDim.x) {
    for (int jj = 0; jj < kAttnJT; ++jj) {
      const int j = j0 + jj;
      if (j >= p.Te) break;
      float acc = 0.0f;
      if (j < len) {
        for (int i = 0; i < p.rows_per_clip; ++i) {
          const long long r = r0 + i;
          acc = fmaf(p.w[r * p.ld_w + j], p.dctx[r * p.ld_dctx + e], acc);
        }
      }
      float* dst = p.dmem + ((long long)clip * p.Te + j) * p.E + e;
      *dst = p.dmem_accumulate ? *dst + acc : acc;
    }
  }
}

inline int launch_attn_bwd_acc(const AttnBwdAccParams& p, cudaStream_t st) {
  if (p.clips <= 0) return 0;
  dim3 grid((p.Te + kAttnJT - 1) / kAttnJT, p.clips);
  ACVAE_LAUNCH(attn_bwd_acc_kernel, grid, 256, 0, st, p);
  return 0;
}

}  // namespace acvae
