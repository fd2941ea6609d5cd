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
  const float* qp_in; long long ld_qp_in;      // q·Wq^T [rows,A] (skinny GEMM) or NULL => zero query
  const float* P; const float* mem; const float* v; const int* mem_lens;
  float* ctx; long long ld_ctx;                 // [rows,E]
  float* w_out; long long ld_w;                 // saved weights [rows,Te] or NULL
  float* aw_out; long long aw_ld_r, aw_ld_j;    // user-visible weights (vae_model.py:868 layout) or NULL
  const int* live;                              // optional device flag: return at once when *live == 0
};

// The clip's projected memory P[clip] and memory mem[clip] are streamed through a 4-deep ring of
// 32-frame shared-memory tiles with cp.async: all tiles of a 62-frame clip are in flight at once, so
// a CTA pays one L2/HBM round trip, not one per frame.
constexpr int kAttnJT = 32;      // frames per tile
constexpr int kAttnNB = 4;       // ring depth

__device__ __forceinline__ void attn_cp16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void attn_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void attn_wait_dyn(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::); break;
    default: asm volatile("cp.async.wait_group 3;\n" ::); break;
  }
}

// Ring schedule shared by forward and backward: `ntile` tiles per pass, two passes over the clip
// (src0 then src1), widths W0 / W1 floats per frame (multiples of 4, 16-byte aligned rows).
struct AttnRing {
  const float *src0, *src1;
  int W0, W1, len, ntile, WMAX;
  float* bufs;
  int nb = kAttnNB;   // ring depth actually used (<= kAttnNB)
  __device__ __forceinline__ void issue(int c) const {
    const bool second = c >= ntile;
    const int t = second ? c - ntile : c;
    const float* src = (second ? src1 : src0) + (long long)t * kAttnJT * (second ? W1 : W0);
    const int W = second ? W1 : W0;
    const int nf = min(kAttnJT, len - t * kAttnJT);
    float* dst = bufs + (size_t)(c % nb) * kAttnJT * WMAX;
    const int n4 = nf * W / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) attn_cp16(dst + i * 4, src + i * 4);   // tile stored densely [nf][W]
    attn_commit();
  }
  __device__ __forceinline__ const float* tile(int c) const { return bufs + (size_t)(c % nb) * kAttnJT * WMAX; }
};

// dynamic smem: A (qp) + Te (scores) + 64 (reduction scratch) + ring
__global__ void __launch_bounds__(256) attn_fwd_kernel(const __grid_constant__ AttnFwdParams p) {
  extern __shared__ __align__(16) float sm[];
  if (p.live && *p.live == 0) return;
  const int WMAX = max(p.A, p.E);
  float* bufs = sm;
  float* qp = bufs + (size_t)kAttnNB * kAttnJT * WMAX;
  float* sc = qp + p.A;
  float* red = sc + p.Te;
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int clip = r / p.rows_per_clip;
  const int len = max(1, min(p.mem_lens[clip], p.Te));
  AttnRing ring{p.P + (long long)clip * p.Te * p.A, p.mem + (long long)clip * p.Te * p.E, p.A, p.E, len,
                (len + kAttnJT - 1) / kAttnJT, WMAX, bufs};
  const int nchunk = 2 * ring.ntile;
  int issued = 0;
  for (; issued < min(nchunk, kAttnNB); ++issued) ring.issue(issued);

  for (int a = tid; a < p.A; a += blockDim.x) qp[a] = p.qp_in ? p.qp_in[(long long)r * p.ld_qp_in + a] : 0.0f;
  __syncthreads();

  // pass 1: scores, one warp per frame
  for (int c = 0; c < ring.ntile; ++c) {
    attn_wait_dyn(issued - 1 - c);
    __syncthreads();
    const float* tile = ring.tile(c);
    const int nf = min(kAttnJT, len - c * kAttnJT);
    for (int jj = wid; jj < nf; jj += nw) {
      const float* pr = tile + jj * p.A;
      float s = 0.0f;
      for (int a = lane; a < p.A; a += 32) s = fmaf(__ldg(p.v + a), attn_tanh(pr[a] + qp[a]), s);
      s = warp_sum(s);
      if (lane == 0) sc[c * kAttnJT + jj] = s;
    }
    __syncthreads();
    if (issued < nchunk) { ring.issue(issued); ++issued; }
  }
  // masked softmax over valid frames (masked frames: exp(-1e10 - max) == 0 exactly in fp32)
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
  // pass 2: context, one thread per feature
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // supports E <= 4 * blockDim.x
  for (int c = ring.ntile; c < nchunk; ++c) {
    attn_wait_dyn(issued - 1 - c);
    __syncthreads();
    const float* tile = ring.tile(c);
    const int j0 = (c - ring.ntile) * kAttnJT;
    const int nf = min(kAttnJT, len - j0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * blockDim.x;
      if (e < p.E) {
        float a0 = 0.f, a1 = 0.f;
        int jj = 0;
        for (; jj + 2 <= nf; jj += 2) {
          a0 = fmaf(sc[j0 + jj], tile[jj * p.E + e], a0);
          a1 = fmaf(sc[j0 + jj + 1], tile[(jj + 1) * p.E + e], a1);
        }
        if (jj < nf) a0 = fmaf(sc[j0 + jj], tile[jj * p.E + e], a0);
        acc[q] += a0 + a1;
      }
    }
    __syncthreads();
    if (issued < nchunk) { ring.issue(issued); ++issued; }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = tid + q * blockDim.x;
    if (e < p.E) p.ctx[(long long)r * p.ld_ctx + e] = acc[q];
  }
}

// ---- forward for K query rows that share one clip (diverse sampling: K captions per clip) ------------------------
// One CTA per clip: the clip's P and mem tiles are streamed ONCE through the cp.async ring and reused by all
// R = rows_per_clip queries (the one-CTA-per-row kernel above re-reads them R times from L2: 1.3 GB per call at
// 10 450 rows).  tanh-bound: R * Te * A evaluations per clip.
constexpr int kAttnMaxR = 16;
// dynamic smem: ring + R*A (qp) + R*Te (scores) + 64
// RT > 0: rows_per_clip known at compile time (no predicated-off iterations: at R = 10 the 16-way unrolled runtime
// form issued 25 instructions per tanh, 130 M warp instructions per call); RT == 0: any R <= kAttnMaxR.
template <int RT>
__global__ void __launch_bounds__(256) attn_fwd_multi_kernel(const __grid_constant__ AttnFwdParams p) {
  extern __shared__ __align__(16) float sm[];
  if (p.live && *p.live == 0) return;
  constexpr int RU = RT > 0 ? RT : kAttnMaxR;            // unroll bound
  const int WMAX = max(p.A, p.E), R = RT > 0 ? RT : p.rows_per_clip, A = p.A, E = p.E, Te = p.Te;
  constexpr int NB = 2;                                  // shallow ring: 3 CTAs per SM hide the tanh latency
  float* bufs = sm;
  float* qp = bufs + (size_t)NB * kAttnJT * WMAX;        // [R][A]
  float* sc = qp + R * A;                                // [R][Te]
  const int clip = blockIdx.x;
  const int r0 = clip * R;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int len = max(1, min(p.mem_lens[clip], Te));
  AttnRing ring{p.P + (long long)clip * Te * A, p.mem + (long long)clip * Te * E, A, E, len, (len + kAttnJT - 1) / kAttnJT, WMAX, bufs, NB};
  const int nchunk = 2 * ring.ntile;
  int issued = 0;
  for (; issued < min(nchunk, NB); ++issued) ring.issue(issued);
  // queries pre-scaled by 2 log2(e): the score loop is then  x = P' + q' ; t = ex2(x) ; r = rcp(t + 1) ; s += (-2 v) r
  // (tanh = 1 - 2 r, so sum_a v_a tanh = sum_a v_a + sum_a (-2 v_a) r_a): 4 instructions + 1 LDS per evaluation
  for (int i = tid; i < R * A; i += blockDim.x) {
    const int r = i / A, a = i % A;
    qp[i] = p.qp_in ? kTwoLog2e * p.qp_in[(long long)(r0 + r) * p.ld_qp_in + a] : 0.0f;
  }
  float vsum = 0.0f;
  for (int a = lane; a < A; a += 32) vsum += __ldg(p.v + a);
  vsum = warp_sum(vsum);
  __syncthreads();
  // pass 1: scores; a warp owns a frame, lanes stride A, all R queries per loaded P value
  for (int c = 0; c < ring.ntile; ++c) {
    attn_wait_dyn(issued - 1 - c);
    __syncthreads();
    const float* tile = ring.tile(c);
    const int nf = min(kAttnJT, len - c * kAttnJT);
    for (int jj = wid; jj < nf; jj += nw) {
      const float* pr = tile + jj * A;
      float s[RU];
#pragma unroll
      for (int r = 0; r < RU; ++r) s[r] = 0.0f;
#pragma unroll 2
      for (int a = lane; a < A; a += 32) {
        const float pv = kTwoLog2e * pr[a], vv = -2.0f * __ldg(p.v + a);
#pragma unroll
        for (int r = 0; r < RU; ++r)
          if (RT > 0 || r < R) s[r] = fmaf(vv, rcp_approx(ex2_approx(pv + qp[r * A + a]) + 1.0f), s[r]);
      }
#pragma unroll
      for (int r = 0; r < RU; ++r)
        if (RT > 0 || r < R) {
          const float t = warp_sum(s[r]) + vsum;
          if (lane == 0) sc[r * Te + c * kAttnJT + jj] = t;
        }
    }
    __syncthreads();
    if (issued < nchunk) { ring.issue(issued); ++issued; }
  }
  // masked softmax, one warp per query row
  for (int r = wid; r < R; r += nw) {
    float* sr = sc + r * Te;
    float mx = -INFINITY;
    for (int j = lane; j < len; j += 32) mx = fmaxf(mx, sr[j]);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < len; j += 32) { const float e = expf(sr[j] - mx); sr[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Te; j += 32) {
      const float w = j < len ? sr[j] * inv : 0.0f;
      if (j < len) sr[j] = w;
      if (p.w_out) p.w_out[(long long)(r0 + r) * p.ld_w + j] = w;
    }
  }
  // pass 2: context; a thread owns features e, e + 256, ... for all R queries
  for (int e0 = 0; e0 < E; e0 += blockDim.x) {   // E <= 256 in practice: one sweep of the ring
    float acc[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) acc[r] = 0.0f;
    const int e = e0 + tid;
    for (int c = ring.ntile; c < nchunk; ++c) {
      attn_wait_dyn(issued - 1 - c);
      __syncthreads();
      const float* tile = ring.tile(c);
      const int j0 = (c - ring.ntile) * kAttnJT;
      const int nf = min(kAttnJT, len - j0);
      if (e < E) {
        for (int jj = 0; jj < nf; ++jj) {
          const float mv = tile[jj * E + e];
#pragma unroll
          for (int r = 0; r < RU; ++r)
            if (RT > 0 || r < R) acc[r] = fmaf(sc[r * Te + j0 + jj], mv, acc[r]);
        }
      }
      __syncthreads();
      if (issued < nchunk) { ring.issue(issued); ++issued; }
    }
    if (e < E) {
#pragma unroll
      for (int r = 0; r < RU; ++r)
        if (RT > 0 || r < R) p.ctx[(long long)(r0 + r) * p.ld_ctx + e] = acc[r];
    }
  }
}

inline size_t attn_smem_bytes(int A, int E, int Te, int extra) {
  const int WMAX = A > E ? A : E;
  return ((size_t)kAttnNB * kAttnJT * WMAX + A + Te + 64 + extra) * sizeof(float);
}

inline int launch_attn_fwd(const AttnFwdParams& p, cudaStream_t st) {
  if (p.rows <= 0) return 0;
  ACVAE_REQUIRE(p.E <= 1024 && p.A % 4 == 0 && p.E % 4 == 0, "attention: E <= 1024 and A, E multiples of 4");
  if (p.rows_per_clip > 1 && p.rows_per_clip <= kAttnMaxR && p.rows % p.rows_per_clip == 0 && p.E <= 256 && !p.aw_out) {
    const int WMAX = p.A > p.E ? p.A : p.E;
    const size_t smem_m = ((size_t)2 * kAttnJT * WMAX + (size_t)p.rows_per_clip * (p.A + p.Te) + 64) * sizeof(float);
    if (smem_m <= 227 * 1024) {
      // compile-time row counts for the common captions-per-clip / beam sizes, the runtime form for the rest
#define ACVAE_ATTN_MULTI(RT)                                                                                             \
      {                                                                                                                  \
        static size_t configured_dev[kMaxDevices] = {0};                                                                 \
        size_t& configured_m = configured_dev[current_device()];                                                         \
        if (smem_m > configured_m) {                                                                                     \
          ACVAE_CHECK(cudaFuncSetAttribute(attn_fwd_multi_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m)); \
          configured_m = smem_m;                                                                                         \
        }                                                                                                                \
        ACVAE_LAUNCH(attn_fwd_multi_kernel<RT>, p.rows / p.rows_per_clip, 256, smem_m, st, p);                           \
      }
      switch (p.rows_per_clip) {
        case 2: ACVAE_ATTN_MULTI(2) break;
        case 3: ACVAE_ATTN_MULTI(3) break;
        case 4: ACVAE_ATTN_MULTI(4) break;
        case 5: ACVAE_ATTN_MULTI(5) break;
        case 6: ACVAE_ATTN_MULTI(6) break;
        case 8: ACVAE_ATTN_MULTI(8) break;
        case 10: ACVAE_ATTN_MULTI(10) break;
        default: ACVAE_ATTN_MULTI(0) break;
      }
#undef ACVAE_ATTN_MULTI
      return 0;
    }
  }
  const size_t smem = attn_smem_bytes(p.A, p.E, p.Te, 0);
  ACVAE_REQUIRE(smem <= 227 * 1024, "attention tile ring exceeds shared memory");
  static size_t configured_dev[kMaxDevices] = {0};
  size_t& configured = configured_dev[current_device()];
  if (smem > configured) {
    ACVAE_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
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

// dynamic smem: ring + A (qp) + Te (dw/ds) + 64 + E (dctx)
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const __grid_constant__ AttnBwdQParams p) {
  extern __shared__ __align__(16) float sm[];
  const int WMAX = max(p.A, p.E);
  float* bufs = sm;
  float* qp = bufs + (size_t)kAttnNB * kAttnJT * WMAX;
  float* dw = qp + p.A;
  float* red = dw + p.Te;
  float* dc = red + 64;
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int clip = r / p.rows_per_clip;
  const int len = max(1, min(p.mem_lens[clip], p.Te));
  // pass 1 streams mem (dw_j = dctx . mem_j), pass 2 streams P (tanh derivative)
  AttnRing ring{p.mem + (long long)clip * p.Te * p.E, p.P + (long long)clip * p.Te * p.A, p.E, p.A, len,
                (len + kAttnJT - 1) / kAttnJT, WMAX, bufs};
  const int nchunk = 2 * ring.ntile;
  int issued = 0;
  for (; issued < min(nchunk, kAttnNB); ++issued) ring.issue(issued);
  for (int e = tid; e < p.E; e += blockDim.x) dc[e] = p.dctx[(long long)r * p.ld_dctx + e];
  for (int a = tid; a < p.A; a += blockDim.x) qp[a] = p.qp[(long long)r * p.ld_qp + a];
  __syncthreads();
  for (int c = 0; c < ring.ntile; ++c) {
    attn_wait_dyn(issued - 1 - c);
    __syncthreads();
    const float* tile = ring.tile(c);
    const int nf = min(kAttnJT, len - c * kAttnJT);
    for (int jj = wid; jj < nf; jj += nw) {
      float s = 0.0f;
      for (int e = lane; e < p.E; e += 32) s = fmaf(dc[e], tile[jj * p.E + e], s);
      s = warp_sum(s);
      if (lane == 0) dw[c * kAttnJT + jj] = s;
    }
    __syncthreads();
    if (issued < nchunk) { ring.issue(issued); ++issued; }
  }
  const float* wr = p.w + (long long)r * p.ld_w;
  float dot = 0.0f;
  for (int j = tid; j < len; j += blockDim.x) dot = fmaf(wr[j], dw[j], dot);
  dot = block_sum(dot, red);
  for (int j = tid; j < p.Te; j += blockDim.x) {
    const float d = j < len ? wr[j] * (dw[j] - dot) : 0.0f;
    if (j < len) dw[j] = d;
    p.ds[(long long)r * p.ld_ds + j] = d;
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = ring.ntile; c < nchunk; ++c) {
    attn_wait_dyn(issued - 1 - c);
    __syncthreads();
    const float* tile = ring.tile(c);
    const int j0 = (c - ring.ntile) * kAttnJT;
    const int nf = min(kAttnJT, len - j0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = tid + q * blockDim.x;
      if (a < p.A) {
        const float qa = qp[a], va = __ldg(p.v + a);
        float s = 0.0f;
        for (int jj = 0; jj < nf; ++jj) {
          const float th = attn_tanh(tile[jj * p.A + a] + qa);
          s = fmaf(dw[j0 + jj] * va, 1.0f - th * th, s);
        }
        acc[q] += s;
      }
    }
    __syncthreads();
    if (issued < nchunk) { ring.issue(issued); ++issued; }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int a = tid + q * blockDim.x;
    if (a < p.A) p.dqp[(long long)r * p.ld_dqp + a] = acc[q];
  }
}

inline int launch_attn_bwd_q(const AttnBwdQParams& p, cudaStream_t st) {
  if (p.rows <= 0) return 0;
  ACVAE_REQUIRE(p.E <= 1024 && p.A <= 1024 && p.A % 4 == 0 && p.E % 4 == 0, "attention: A, E <= 1024, multiples of 4");
  const size_t smem = attn_smem_bytes(p.A, p.E, p.Te, p.E);
  ACVAE_REQUIRE(smem <= 227 * 1024, "attention tile ring exceeds shared memory");
  static size_t configured_dev[kMaxDevices] = {0};
  size_t& configured = configured_dev[current_device()];
  if (smem > configured) {
    ACVAE_CHECK(cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
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

constexpr int kAttnAccJT = 4;

__global__ void __launch_bounds__(256) attn_bwd_acc_kernel(const __grid_constant__ AttnBwdAccParams p) {
  const int clip = blockIdx.y;
  const int j0 = blockIdx.x * kAttnAccJT;
  const int tid = threadIdx.x;
  const int len = min(p.mem_lens[clip], p.Te);
  const int r0 = clip * p.rows_per_clip;
  // dP and dv: threads over a
  for (int a = tid; a < p.A; a += blockDim.x) {
    const float va = __ldg(p.v + a);
    float dvacc = 0.0f;
    for (int jj = 0; jj < kAttnAccJT; ++jj) {
      const int j = j0 + jj;
      if (j >= p.Te) break;
      float acc = 0.0f;
      if (j < len) {
        const float pj = p.P[((long long)clip * p.Te + j) * p.A + a];
        for (int i = 0; i < p.rows_per_clip; ++i) {
          const long long r = r0 + i;
          const float d = p.ds[r * p.ld_ds + j];
          const float th = attn_tanh(pj + p.qp[r * p.ld_qp + a]);
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
    for (int jj = 0; jj < kAttnAccJT; ++jj) {
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
  dim3 grid((p.Te + kAttnAccJT - 1) / kAttnAccJT, p.clips);
  ACVAE_LAUNCH(attn_bwd_acc_kernel, grid, 256, 0, st, p);
  return 0;
}

// ---- backward of ALL rows of a clip together (hoisted schedule) -------------------------------------------------------
// attn_bwd_q_kernel (one CTA per query row) + attn_bwd_acc_kernel (one CTA per clip and frame chunk) evaluate the same
// 1 - tanh^2(P[j,a] + q[t,a]) terms twice and re-stream the clip's P / mem once per row: 46 + 42 us for the prior's
// 608 rows at CFG1, on the critical tail of the backward.  Here a clip's rows are handled together, in two launches:
//   attn_dalpha_kernel    d alpha[t,j] = d ctx[t,:] . mem[n,j,:]        grid (Te/8, clips), a warp per frame
//                         (skipped when `ds_in` is given: the decoder's chain kernels have produced d s and d qp already)
//   attn_bwd_clip_kernel  grid (8, clips); CTA (s, n) owns attention columns / memory columns [s*A/8, (s+1)*A/8):
//     softmax backward d s[t,j] = w (d alpha - sum w d alpha) (every CTA of the clip, redundantly: T x Te values);
//     ONE evaluation of g = 1 - tanh^2 per (t, j, a) feeds d qp[t,a] = v_a sum_j d s g, d P[j,a] = v_a sum_t d s g and
//     d v[a] = sum d s tanh;   d mem[n,j,e] (+)= sum_t w[t,j] d ctx[t,e]
struct AttnBwdClipParams {
  int clips, T, Te, A, E;
  const float* dctx;        // [clips*T, E]
  const float* w;           // [clips*T, Te] saved softmax weights
  const float* qp;          // [clips*T, A]  saved query projections
  const float *P, *mem, *v; // [clips,Te,A], [clips,Te,E], [A]
  const int* mem_lens;
  const float* ds_in;       // optional [clips*T, Te]: d score already known
  float* ds_out;            // [clips*T, Te]: scratch that carries d alpha between the two launches (used when ds_in == NULL)
  float* dqp;               // [clips*T, A]  (written when ds_in == NULL)
  float* dP;                // [clips,Te,A] (overwritten)
  float* dmem; int dmem_accumulate;   // [clips,Te,E]
  float* dv;                // [A], atomically accumulated (zeroed by the caller)
};
constexpr int kAbcSplit = 8;              // CTAs per clip: 256 CTAs of 30 K registers / 60 KB, two per SM (4 -> 8: 41 -> ~22 us, profiles/r2)
constexpr int kAbcMaxFr = 12;             // frames per thread in the tanh pass: Te <= (256 / AS) * 12
inline size_t attn_bwd_clip_smem(int T, int Te, int A, int E) {
  const int ld = Te + 1, AS = A / kAbcSplit;
  return ((size_t)2 * T * ld + (size_t)T * E + (size_t)T * AS + (size_t)256 * T + (size_t)Te * AS + 64) * sizeof(float);
}

// d alpha[t, j] for one clip and 8 frames: warp = frame, lanes over E (E <= 256), four rows' shuffle trees in flight at once
__global__ void __launch_bounds__(256) attn_dalpha_kernel(int T, int Te, int E, const float* __restrict__ dctx, const float* __restrict__ mem,
                                                          const int* __restrict__ mem_lens, float* __restrict__ dalpha) {
  extern __shared__ __align__(16) float sm[];              // [T][E] d ctx rows of the clip
  const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, j = blockIdx.x * 8 + (tid >> 5);
  const int len = max(1, min(mem_lens[n], Te));
  const long long r0 = (long long)n * T;
  for (int i = tid; i < T * E; i += 256) sm[i] = dctx[r0 * E + i];
  __syncthreads();
  if (j >= Te) return;
  if (j >= len) {
    for (int t = lane; t < T; t += 32) dalpha[(r0 + t) * Te + j] = 0.0f;
    return;
  }
  float m[8];
  const float* mr = mem + ((long long)n * Te + j) * E;
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = lane + 32 * k < E ? __ldg(mr + lane + 32 * k) : 0.0f;
  for (int t0 = 0; t0 < T; t0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = t0 + q < T ? t0 + q : T - 1;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (lane + 32 * k < E) acc[q] = fmaf(sm[t * E + lane + 32 * k], m[k], acc[q]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    if (lane < 4 && t0 + lane < T) dalpha[(r0 + t0 + lane) * Te + j] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
  }
}

__global__ void __launch_bounds__(256) attn_bwd_clip_kernel(const __grid_constant__ AttnBwdClipParams p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T, Te = p.Te, A = p.A, E = p.E, AS = A / kAbcSplit, ES = E / kAbcSplit, ld = Te + 1;
  float* dss = sm;                        // [T][ld]  d alpha, then d score
  float* ws = dss + T * ld;               // [T][ld]  softmax weights
  float* dcs = ws + T * ld;               // [T][E]   d ctx rows of the clip
  float* qs = dcs + T * E;                // [T][AS]  2 log2 e * q[t, my columns]
  float* red = qs + T * AS;               // [G][T][AS] partial d qp per frame group (G = 256 / AS)
  float* Pss = red + 256 * T;             // [Te][AS] 2 log2 e * P[n, :, my columns]
  const int s = blockIdx.x, n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int len = max(1, min(p.mem_lens[n], Te));
  const long long r0 = (long long)n * T;
  const float* dsrc = p.ds_in ? p.ds_in : p.ds_out;           // d score, or d alpha left by attn_dalpha_kernel
  for (int i = tid; i < T * E; i += 256) dcs[i] = p.dctx[r0 * E + i];
  for (int i = tid; i < T * Te; i += 256) {
    const int t = i / Te, j = i % Te;
    ws[t * ld + j] = j < len ? p.w[(r0 + t) * Te + j] : 0.0f;
    dss[t * ld + j] = j < len ? dsrc[(r0 + t) * Te + j] : 0.0f;
  }
  for (int i = tid; i < T * AS; i += 256) { const int t = i / AS, a = i % AS; qs[i] = kTwoLog2e * p.qp[(r0 + t) * A + s * AS + a]; }
  for (int i = tid; i < len * AS; i += 256) { const int j = i / AS, a = i % AS; Pss[i] = kTwoLog2e * __ldg(p.P + ((long long)n * Te + j) * A + s * AS + a); }
  __syncthreads();
  if (!p.ds_in) {
    // softmax backward per row (a warp per row)
    for (int t = wid; t < T; t += 8) {
      float dot = 0.0f;
      for (int j = lane; j < len; j += 32) dot = fmaf(ws[t * ld + j], dss[t * ld + j], dot);
      dot = warp_sum(dot);
      for (int j = lane; j < Te; j += 32) {
        const float d = j < len ? ws[t * ld + j] * (dss[t * ld + j] - dot) : 0.0f;
        dss[t * ld + j] = d;
      }
    }
    __syncthreads();     // (d score stays on chip: the four CTAs of a clip all read d alpha from the scratch, nobody may overwrite it)
  }
  // ---- tanh pass: thread = (column a of my slice, frame group jq): frames j = jq + G*f, f < kAbcMaxFr; rows t in the outer loop ----
  {
    const int G = 256 / AS;                                    // frame groups (4 at A = 256)
    const int a = tid % AS, jq = tid / AS, ag = s * AS + a;
    const float va = __ldg(p.v + ag);
    float dvacc = 0.0f;
    float accP[kAbcMaxFr], pj[kAbcMaxFr];
#pragma unroll
    for (int f = 0; f < kAbcMaxFr; ++f) { accP[f] = 0.0f; const int j = jq + G * f; pj[f] = j < len ? Pss[j * AS + a] : 0.0f; }
    for (int t = 0; t < T; ++t) {
      const float qv = qs[t * AS + a];
      const float* dr = dss + t * ld + jq;
      float dq = 0.0f;
#pragma unroll
      for (int f = 0; f < kAbcMaxFr; ++f) {
        if (jq + G * f < len) {
          const float r = rcp_approx(ex2_approx(pj[f] + qv) + 1.0f);     // tanh = 1 - 2 r, 1 - tanh^2 = 4 r (1 - r)
          const float d = dr[G * f];
          const float g = 4.0f * d * fmaf(-r, r, r);
          accP[f] += g; dq += g;
          dvacc = fmaf(d, fmaf(-2.0f, r, 1.0f), dvacc);
        }
      }
      if (!p.ds_in) red[(jq * T + t) * AS + a] = dq;
    }
#pragma unroll
    for (int f = 0; f < kAbcMaxFr; ++f) {
      const int j = jq + G * f;
      if (j < Te) p.dP[((long long)n * Te + j) * A + ag] = j < len ? va * accP[f] : 0.0f;
    }
    atomicAdd(p.dv + ag, dvacc);
    if (!p.ds_in) {
      __syncthreads();
      for (int i = tid; i < T * AS; i += 256) {
        const int t = i / AS, aa = i % AS;
        float sum = 0.0f;
        for (int g2 = 0; g2 < G; ++g2) sum += red[(g2 * T + t) * AS + aa];
        p.dqp[(r0 + t) * A + s * AS + aa] = __ldg(p.v + s * AS + aa) * sum;
      }
    }
  }
  // ---- d mem[n, j, my e slice] (+)= sum_t w[t,j] d ctx[t,e] ----
  {
    const int G = 256 / ES;
    const int e = tid % ES, jq = tid / ES, eg = s * ES + e;
    for (int j = jq; j < Te; j += G) {
      float acc = 0.0f;
      if (j < len)
        for (int t = 0; t < T; ++t) acc = fmaf(ws[t * ld + j], dcs[t * E + eg], acc);
      float* dst = p.dmem + ((long long)n * Te + j) * E + eg;
      *dst = p.dmem_accumulate ? *dst + acc : acc;
    }
  }
}
// returns 1 if launched, 0 if the shape is not covered (caller falls back to attn_bwd_q + attn_bwd_acc)
inline int launch_attn_bwd_clip(const AttnBwdClipParams& p, cudaStream_t st) {
  if (p.clips <= 0) return 1;
  const int AS = p.A / kAbcSplit, ES = p.E / kAbcSplit;
  if (p.A % (4 * kAbcSplit) || p.E % (4 * kAbcSplit) || AS > 256 || ES > 256 || 256 % AS || 256 % ES || 256 / AS > 8 || p.E > 256 ||
      p.Te > (256 / AS) * kAbcMaxFr)
    return 0;
  const size_t smem = attn_bwd_clip_smem(p.T, p.Te, p.A, p.E);
  const size_t smem_a = (size_t)p.T * p.E * sizeof(float);
  if (smem > 200 * 1024 || smem_a > 200 * 1024) return 0;
  static size_t configured_dev[kMaxDevices] = {0}, configured_a[kMaxDevices] = {0};
  size_t& configured = configured_dev[current_device()];
  if (smem > configured) {
    ACVAE_CHECK(cudaFuncSetAttribute(attn_bwd_clip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  if (!p.ds_in) {
    size_t& conf_a = configured_a[current_device()];
    if (smem_a > conf_a) {
      ACVAE_CHECK(cudaFuncSetAttribute(attn_dalpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
      conf_a = smem_a;
    }
    ACVAE_LAUNCH(attn_dalpha_kernel, dim3((p.Te + 7) / 8, p.clips), 256, smem_a, st, p.T, p.Te, p.E, p.dctx, p.mem, p.mem_lens, p.ds_out);
  }
  ACVAE_LAUNCH(attn_bwd_clip_kernel, dim3(kAbcSplit, p.clips), 256, smem, st, p);
  return 1;
}

}  // namespace acvae
