// Element-wise, gather/scatter and small reduction kernels of the hot path:
// word selection, vocabulary-statistics reduce, masked pooling (global
// constraint), Gaussian KL, label-smoothed CE rows, and the pointwise halves of
// the GRU / LSTM / Gaussian-head backward.  All HBM/L2-bound, coalesced along
// the feature dimension.
#pragma once
#include "common.cuh"

namespace acvae {

// ---- word selection (reference vae_model.py:826-832) ---------------------------
// words[n,t] = caps[n,t] for teacher-forced steps; <start> for a free step 0;
// other free steps are filled by the previous step's vocab reduce.
__global__ void words_init_kernel(int N, int T, int L, const int* __restrict__ caps, unsigned long long tf_mask,
                                  int start_idx, int* __restrict__ words) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * T) return;
  const int n = i / T, t = i % T;
  if ((tf_mask >> t) & 1ull) words[i] = caps[(long long)n * L + t];
  else words[i] = start_idx;      // t == 0: the start token; free steps: a VALID placeholder (the batched gathers of the hoisted
                                  // schedule read every entry) until the previous step's arg-max overwrites it
}

// ---- vocabulary statistics: combine per-tile partials -------------------------
struct VocabReduceParams {
  int M, ntiles;
  const float* pmax; const float* pexp; const float* psum; const float* pbest; const int* parg;
  float* lse; float* lsum; float* logprob; long long ld_row;   // row m -> [m*ld_row]
  long long* seqs; long long ld_seqs;                           // int64 out or NULL
  int* next_word; long long ld_next;                            // int32 out or NULL (word fed at the next step)
  int* unfinished; int end_idx;                                 // sampling: stop bookkeeping or NULL
  int* active_count;                                            // sampling: device counter of unfinished rows
  const int* live;                                              // optional: skip when *live == 0
};

__global__ void __launch_bounds__(128) vocab_reduce_kernel(const __grid_constant__ VocabReduceParams p) {
  // one warp per row: lanes stride the per-tile partials
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= p.M) return;
  if (p.live && *p.live == 0) return;
  const long long o = (long long)m * p.ntiles;
  float gmax = -INFINITY;
  for (int i = lane; i < p.ntiles; i += 32) gmax = fmaxf(gmax, p.pmax[o + i]);
  gmax = warp_max(gmax);
  float se = 0.0f, ss = 0.0f, best = -INFINITY, bl = 0.0f;
  int arg = 0x7fffffff;
  for (int i = lane; i < p.ntiles; i += 32) {
    se += p.pexp[o + i] * expf(p.pmax[o + i] - gmax);
    ss += p.psum[o + i];
    const float b = p.pbest[(o + i) * 2];
    const int a = p.parg[o + i];
    if (b > best || (b == best && a < arg)) { best = b; bl = p.pbest[(o + i) * 2 + 1]; arg = a; }
  }
  se = warp_sum(se);
  ss = warp_sum(ss);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
    const float ol = __shfl_xor_sync(0xffffffffu, bl, off);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; bl = ol; }
  }
  if (lane != 0) return;
  const float lse = gmax + logf(se);
  if (p.lse) p.lse[m * p.ld_row] = lse;
  if (p.lsum) p.lsum[m * p.ld_row] = ss;
  if (p.logprob) p.logprob[m * p.ld_row] = bl - lse;
  int w = arg;
  if (p.unfinished) {
    // reference vae_model.py:712-718: a row stays finished once it emitted <end>
    const int u = p.unfinished[m] && (arg != p.end_idx);
    p.unfinished[m] = u;
    if (!u) w = p.end_idx;
    if (u && p.active_count) atomicAdd(p.active_count, 1);
  }
  if (p.seqs) p.seqs[m * p.ld_seqs] = (long long)w;
  if (p.next_word) p.next_word[m * p.ld_next] = w;
}

// ---- gather embedding rows ------------------------------------------------------
__global__ void embed_gather_kernel(int rows, int E, const float* __restrict__ table, const int* __restrict__ idx,
                                    float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * E) return;
  const int r = (int)(i / E), e = (int)(i % E);
  out[i] = table[(long long)idx[r] * E + e];
}

// grad_table[idx[r], :] += d[r, :]   (dense embedding gradient, as nn.Embedding)
__global__ void embed_scatter_add_kernel(int rows, int E, const float* __restrict__ d, long long ldd,
                                         const int* __restrict__ idx, float* __restrict__ grad_table) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * E) return;
  const int r = (int)(i / E), e = (int)(i % E);
  atomicAdd(grad_table + (long long)idx[r] * E + e, d[(long long)r * ldd + e]);
}

// ---- masked mean + max pooling over time (utils/train_util.py:207-231) ----------
// x [N,T,D], lens [N] (valid steps) -> pool [N,D] = mean_{t<len} + max_{t<len}; argmax saved.
__global__ void pool_fwd_kernel(int N, int T, int D, const float* __restrict__ x, const int* __restrict__ lens,
                                int len_off, float* __restrict__ pool, int* __restrict__ amax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  const int n = i / D, d = i % D;
  const int len = min(lens[n] + len_off, T);
  float s = 0.0f, mx = -INFINITY;
  int am = 0;
  for (int t = 0; t < len; ++t) {
    const float v = x[((long long)n * T + t) * D + d];
    s += v;
    if (v > mx) { mx = v; am = t; }
  }
  pool[i] = s / (float)len + mx;
  amax[i] = am;
}

// dx[n,t,d] = base[n,t,d] (or 0) + dpool[n,d]/len (t<len) + dpool[n,d]*[t==amax]
__global__ void pool_bwd_kernel(int N, int T, int D, const float* __restrict__ dpool, const int* __restrict__ lens,
                                int len_off, const int* __restrict__ amax, const float* __restrict__ base,
                                float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * T * D) return;
  const int d = (int)(i % D);
  const int t = (int)((i / D) % T);
  const int n = (int)(i / ((long long)D * T));
  const int len = min(lens[n] + len_off, T);
  float v = base ? base[i] : 0.0f;
  if (dpool && t < len) {
    const float g = dpool[(long long)n * D + d];
    v += g / (float)len;
    if (amax[(long long)n * D + d] == t) v += g;
  }
  dx[i] = v;
}

// ---- Gaussian KL (utils/train_util.py:259-266) -----------------------------------
// One block per group of rows, deterministic two-stage sum.
__global__ void __launch_bounds__(256) kl_partial_kernel(long long n_elem, const float* __restrict__ mq,
                                                         const float* __restrict__ lq, const float* __restrict__ mp,
                                                         const float* __restrict__ lp, float* __restrict__ partial) {
  __shared__ float red[33];
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += (long long)gridDim.x * blockDim.x) {
    const float d = mq[i] - mp[i];
    s += 0.5f * lp[i] - 0.5f * lq[i] + (expf(lq[i]) + d * d) / (2.0f * expf(lp[i])) - 0.5f;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = scale * sum(partial[0..n))   (single block, deterministic)
__global__ void __launch_bounds__(256) final_sum_kernel(int n, const float* __restrict__ partial, float scale,
                                                        const float* __restrict__ denom, float* __restrict__ out) {
  __shared__ float red[33];
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = denom ? scale * s / denom[0] : scale * s;
}

__global__ void kl_bwd_kernel(long long n_elem, float inv_rows, const float* __restrict__ mq, const float* __restrict__ lq,
                              const float* __restrict__ mp, const float* __restrict__ lp, const float* __restrict__ dkl,
                              float* __restrict__ dmq, float* __restrict__ dlq, float* __restrict__ dmp,
                              float* __restrict__ dlp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elem) return;
  const float g = dkl[0] * inv_rows;
  const float d = mq[i] - mp[i];
  const float vq = expf(lq[i]), ivp = expf(-lp[i]);
  const float gm = g * d * ivp;
  dmq[i] = gm;
  dmp[i] = -gm;
  dlq[i] = g * (-0.5f + 0.5f * vq * ivp);
  dlp[i] = g * (0.5f - 0.5f * (vq + d * d) * ivp);
}

// ---- label-smoothed CE rows from vocabulary statistics ---------------------------
// row loss = -[(on-off)*(logit_y - lse) + off*(sum_j logit_j - V*lse)]   (train_util.py:244-251)
struct CeRowsParams {
  int M, V, E;
  const float* hidden; long long ld_h;
  const float* cls_w; const float* cls_b;
  const int* targets; const float* row_w;
  const float* lse; const float* lsum;
  float on, off;
  float* row_loss;   // [M] weighted
  float* row_cnt;    // [M] weights (for the mean denominator)
};
__global__ void __launch_bounds__(256) ce_rows_kernel(const __grid_constant__ CeRowsParams p) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= p.M) return;
  const int y = p.targets[m];
  const float* h = p.hidden + (long long)m * p.ld_h;
  const float* w = p.cls_w + (long long)y * p.E;
  float s = 0.0f;
  for (int k = lane; k < p.E; k += 32) s = fmaf(h[k], __ldg(w + k), s);
  s = warp_sum(s);
  if (lane == 0) {
    const float ly = s + p.cls_b[y];
    const float lse = p.lse[m];
    const float rw = p.row_w ? p.row_w[m] : 1.0f;
    const float loss = -((p.on - p.off) * (ly - lse) + p.off * (p.lsum[m] - (float)p.V * lse));
    p.row_loss[m] = rw * loss;
    p.row_cnt[m] = rw;
  }
}

// gscale[0] = d_loss[0] / sum(row_cnt)
__global__ void __launch_bounds__(256) ce_gscale_kernel(int M, const float* __restrict__ row_w,
                                                        const float* __restrict__ dloss, float* __restrict__ gscale) {
  __shared__ float red[33];
  float s = 0.0f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) s += row_w ? row_w[i] : 1.0f;
  s = block_sum(s, red);
  if (threadIdx.x == 0) gscale[0] = dloss[0] / s;
}

// out[c] = sum_r x[r*ld + c]   (bias gradients); block = 32 columns x 32 row-lanes
__global__ void __launch_bounds__(1024) colsum_kernel(long long rows, int cols, const float* __restrict__ x, long long ld,
                                                      float* __restrict__ out, int accumulate) {
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.0f;
  if (c < cols) {
#pragma unroll 8
    for (long long r = ry; r < rows; r += 32) s += x[r * ld + c];
  }
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][cx];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ---- GRU pointwise backward --------------------------------------------------------
// gates saved gate-major [rows, 4U] = (r, z, n, gh_n).  Produces dGi (r,z,n) and dGh (r,z,n*r)
// and the direct carry dh*z.  Rows with t >= lens[n] (packed posterior) produce zeros.
struct GruBwdParams {
  int N, U;
  const float* dh_ext; long long ld_dh_ext;    // upstream grad of h_t (or NULL)
  const float* dh_carry;                       // [N,U] (or NULL)
  const float* gates; long long ld_gates;
  const float* hprev; long long ld_hprev;      // or NULL (= 0)
  const int* lens; int len_off; int t;         // optional mask
  float* dgi; long long ld_dgi;                // [.,3U]
  float* dgh; long long ld_dgh;                // [.,3U]
  float* dh_out;                               // [N,U]: dh * z
};
__global__ void gru_bwd_kernel(const __grid_constant__ GruBwdParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N * p.U) return;
  const int n = i / p.U, u = i % p.U;
  const int U = p.U;
  float* gi = p.dgi + (long long)n * p.ld_dgi + u;
  float* gh = p.dgh + (long long)n * p.ld_dgh + u;
  if (p.lens && p.t >= p.lens[n] + p.len_off) {
    gi[0] = gi[U] = gi[2 * U] = 0.0f;
    gh[0] = gh[U] = gh[2 * U] = 0.0f;
    p.dh_out[i] = 0.0f;
    return;
  }
  float dh = p.dh_carry ? p.dh_carry[i] : 0.0f;
  if (p.dh_ext) dh += p.dh_ext[(long long)n * p.ld_dh_ext + u];
  const float* g = p.gates + (long long)n * p.ld_gates + u;
  const float r = g[0], z = g[U], nn = g[2 * U], ghn = g[3 * U];
  const float hp = p.hprev ? p.hprev[(long long)n * p.ld_hprev + u] : 0.0f;
  const float dn = dh * (1.0f - z);
  const float dz = dh * (hp - nn);
  const float dan = dn * (1.0f - nn * nn);
  const float dar = dan * ghn * r * (1.0f - r);
  const float daz = dz * z * (1.0f - z);
  gi[0] = dar; gi[U] = daz; gi[2 * U] = dan;
  gh[0] = dar; gh[U] = daz; gh[2 * U] = dan * r;
  p.dh_out[i] = dh * z;
}

// ---- LSTM pointwise backward ----------------------------------------------------------
struct LstmBwdParams {
  int N, U;
  const float* dh;                              // [N,U] total grad of h_t
  const float* dc_carry;                        // [N,U] or NULL
  const float* gates; long long ld_gates;       // (i,f,g,o) gate-major
  const float* c; long long ld_c;               // c_t
  const float* cprev; long long ld_cprev;       // c_{t-1} or NULL
  float* dg; long long ld_dg;                   // [.,4U]
  float* dc_out;                                // [N,U]
};
__global__ void lstm_bwd_kernel(const __grid_constant__ LstmBwdParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N * p.U) return;
  const int n = i / p.U, u = i % p.U;
  const int U = p.U;
  const float* g = p.gates + (long long)n * p.ld_gates + u;
  const float ig = g[0], fg = g[U], gg = g[2 * U], og = g[3 * U];
  const float tc = tanhf(p.c[(long long)n * p.ld_c + u]);
  const float cp = p.cprev ? p.cprev[(long long)n * p.ld_cprev + u] : 0.0f;
  const float dh = p.dh[i];
  float dc = dh * og * (1.0f - tc * tc);
  if (p.dc_carry) dc += p.dc_carry[i];
  float* dg = p.dg + (long long)n * p.ld_dg + u;
  dg[0] = dc * gg * ig * (1.0f - ig);
  dg[U] = dc * cp * fg * (1.0f - fg);
  dg[2 * U] = dc * ig * (1.0f - gg * gg);
  dg[3 * U] = dh * tc * og * (1.0f - og);
  p.dc_out[i] = dc * fg;
}

// ---- Gaussian head + reparameterisation backward ------------------------------------------
// dz = sum of up to three sources; dML = [dmean_ext + dz | dlog_ext + dz*eps*0.5*exp(0.5*log)]
struct HeadBwdParams {
  long long rows; int U;
  const float* dz0; long long ld_dz0;
  const float* dz1; long long ld_dz1;
  const float* dz2; long long ld_dz2;
  const float* dmean; long long ld_dmean;
  const float* dlog; long long ld_dlog;
  const float* eps; long long ld_eps;
  const float* logv; long long ld_logv;
  float* dml; long long ld_dml;                 // [rows, 2U]
  // per-row source switch for batched use: dz1 applies only where bit (row % period) of flag_mask == want
  unsigned long long flag_mask; int period; int want; int use_flags;
};
__global__ void head_bwd_kernel(const __grid_constant__ HeadBwdParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.rows * p.U) return;
  const long long r = i / p.U;
  const int u = (int)(i % p.U);
  float dz = 0.0f;
  if (p.dz0) dz += p.dz0[r * p.ld_dz0 + u];
  if (p.dz1 && (!p.use_flags || (int)((p.flag_mask >> (r % p.period)) & 1ull) == p.want)) dz += p.dz1[r * p.ld_dz1 + u];
  if (p.dz2) dz += p.dz2[r * p.ld_dz2 + u];
  float dm = dz, dl = dz * p.eps[r * p.ld_eps + u] * 0.5f * expf(0.5f * p.logv[r * p.ld_logv + u]);
  if (p.dmean) dm += p.dmean[r * p.ld_dmean + u];
  if (p.dlog) dl += p.dlog[r * p.ld_dlog + u];
  p.dml[r * p.ld_dml + u] = dm;
  p.dml[r * p.ld_dml + p.U + u] = dl;
}

// y[i] += x[i]
__global__ void add_inplace_kernel(long long n, float* __restrict__ y, const float* __restrict__ x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// copy with row strides: dst[r*ldd + c] = src[r*lds + c]
__global__ void copy2d_kernel(long long rows, int cols, const float* __restrict__ src, long long lds,
                              float* __restrict__ dst, long long ldd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i % cols);
  dst[r * ldd + c] = src[r * lds + c];
}

// ---- recurrent cells on precomputed pre-activations (large-batch sampling: the GEMMs run on the tensor cores) ----
// LSTM (text_encoder.py:253): pre [N,4E] gate-major (i,f,g,o) incl. b_ih; rows strided by ld_*.
__global__ void lstm_cell_kernel(int N, int E, const float* __restrict__ pre, const float* __restrict__ b_hh,
                                 const float* __restrict__ c_prev, long long ld_cprev, float* __restrict__ c_out,
                                 float* __restrict__ h_out, long long ld_out, const int* __restrict__ live) {
  if (live && *live == 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * E) return;
  const long long n = i / E;
  const int u = (int)(i % E);
  const float* pr = pre + n * 4 * E + u;
  const float ig = sigmoidf_(pr[0] + b_hh[u]), fg = sigmoidf_(pr[E] + b_hh[E + u]);
  const float gg = tanhf(pr[2 * E] + b_hh[2 * E + u]), og = sigmoidf_(pr[3 * E] + b_hh[3 * E + u]);
  const float cp = c_prev ? c_prev[n * ld_cprev + u] : 0.0f;
  const float cn = fg * cp + ig * gg;
  c_out[n * ld_out + u] = cn;
  h_out[n * ld_out + u] = og * tanhf(cn);
}
// GRU (decoder.py:194): pre_x [N,3E] (r,z,n) incl. b_ih; pre_h [N,3E] without bias (NULL = zero state).
__global__ void gru_cell_kernel(int N, int E, const float* __restrict__ pre_x, const float* __restrict__ pre_h,
                                const float* __restrict__ b_hh, const float* __restrict__ h_prev, long long ld_hprev,
                                float* __restrict__ h_out, long long ld_out, const int* __restrict__ live) {
  if (live && *live == 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * E) return;
  const long long n = i / E;
  const int u = (int)(i % E);
  const float* px = pre_x + n * 3 * E + u;
  float hr = b_hh[u], hz = b_hh[E + u], hn = b_hh[2 * E + u];
  if (pre_h) { const float* ph = pre_h + n * 3 * E + u; hr += ph[0]; hz += ph[E]; hn += ph[2 * E]; }
  const float rg = sigmoidf_(px[0] + hr), zg = sigmoidf_(px[E] + hz);
  const float ng = tanhf(px[2 * E] + rg * hn);
  const float hp = h_prev ? h_prev[n * ld_hprev + u] : 0.0f;
  h_out[n * ld_out + u] = (1.0f - zg) * ng + zg * hp;
}
// Gaussian head + reparameterisation (text_encoder.py:255-262): ml [N,2E] = (mean | log) incl. bias.
__global__ void head_cell_kernel(int N, int E, const float* __restrict__ ml, const float* __restrict__ eps,
                                 float* __restrict__ pm, float* __restrict__ pl, float* __restrict__ pz, long long ld_out,
                                 const int* __restrict__ live) {
  if (live && *live == 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * E) return;
  const long long n = i / E;
  const int u = (int)(i % E);
  const float mean = ml[n * 2 * E + u], lg = ml[n * 2 * E + E + u];
  pm[n * ld_out + u] = mean;
  pl[n * ld_out + u] = lg;
  pz[n * ld_out + u] = eps[n * E + u] * expf(0.5f * lg) + mean;
}

inline int grid1d(long long n, int block = 256) { return (int)((n + block - 1) / block); }

}  // namespace acvae
