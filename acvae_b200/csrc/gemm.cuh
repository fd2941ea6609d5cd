// Segmented fp32 SIMT GEMM with fused recurrent-cell / head / vocabulary epilogues.
//
//   acc[m, (g,u)] = sum over segments s, k < K_s of  A_s(m,k) * B_s(g,u,k)
//
// A "segment" is one block of the K dimension with its own operand pointers,
// so a concatenated input such as the decoder's [emb | ctx | z] (reference
// models/decoder.py:188) or [x | h] of a recurrent cell is never materialised.
// Output columns are "virtual": column c = u*G + g interleaves the G gates of
// unit u so that the epilogue sees all gates of a unit together and can apply
// the LSTM / GRU / Gaussian-head arithmetic before anything is written to HBM.
//
// This is the exact-fp32 path (CUDA-core FFMA): it carries the 1e-4 parity bar
// of BASELINE.json.  Tensor-core (tcgen05) tiles for the batched contractions
// live in tc_gemm.cuh.
#pragma once
#include "common.cuh"

namespace acvae {

constexpr int kMaxSeg = 4;
constexpr int kMaxGate = 4;

struct GemmSeg {
  const float* a;            // A(m,k) = a[row(m)*lda + k]   (a_trans: a[k*lda + m])
  long long lda;
  const int* gather;         // optional row gather: row(m) = gather[m*gather_stride]
  long long gather_stride;
  int a_trans;
  const float* w[kMaxGate];  // B(g,u,k) = w[g][u*ldw + k]   (w_trans: w[g][k*ldw + u]); NULL: no contribution
  long long ldw;
  int w_trans;
  int K;
  int k_zero_period, k_zero_rem;  // if period > 0: k with k % period == rem contribute nothing (shifted operands)
};

enum EpiKind {
  EPI_PLAIN = 0,   // C_g[m*ldc + u] = scale*acc + bias_g[u] (+ C_g if accumulate)
  EPI_LSTM = 1,    // G=4 (i,f,g,o): nn.LSTM cell
  EPI_GRU = 2,     // G=4 (r,z,n_x,n_h): nn.GRU cell, optional packed-sequence length mask
  EPI_HEAD = 3,    // G=2 (mean,log): Gaussian head + reparameterisation
  EPI_STATS = 4,   // G=1: per-row partial max/argmax/sum-exp/sum over this CTA's columns
  EPI_DLOGITS = 5  // G=1: label-smoothed softmax-CE gradient wrt logits
};

struct EpiParams {
  // EPI_PLAIN
  float* c[kMaxGate];
  long long ldc;
  const float* bias[kMaxGate];
  int accumulate;
  float scale;
  // cells: precomputed input-side pre-activations (x-part incl. b_ih) or NULL
  const float* gx; long long ld_gx;
  const float* b_ih; const float* b_hh;          // [G*U] gate-major
  const float* prev; long long ld_prev;           // c_{t-1} (LSTM) / h_{t-1} (GRU) or NULL (= 0)
  float* gates; long long ld_gates;               // saved activations for the backward, gate-major [.., G*U]
  float* out0; long long ld_out0;                 // LSTM: c_t      GRU: h_t (0 at padded rows)   HEAD: mean
  float* out1; long long ld_out1;                 // LSTM: h_t      GRU: optional copy            HEAD: log
  float* out2; long long ld_out2;                 //                                              HEAD: z
  const float* eps; long long ld_eps;             // HEAD: noise
  const int* lens; int t;                         // GRU: row active iff t < lens[m]
  // EPI_STATS
  float* pmax; float* pexp; float* psum; float* pbest; int* parg;   // [M, ntiles]
  const float* noise; long long ld_noise; float inv_temp;           // optional Gumbel-max: uniform u
  // EPI_DLOGITS
  const float* lse; const int* targets; const float* row_w; const float* gscale; // gscale: device [1] = dLoss / count
  float smooth_off, smooth_on;
};

struct GemmParams {
  int M, U, G, nseg;
  const int* live;   // optional device flag: the kernel returns at once when *live == 0
  GemmSeg seg[kMaxSeg];
  EpiParams epi;
};

__device__ __forceinline__ float gumbel_from_u(float u) {
  // reference models/word_model.py:188-190 (eps = 1e-20)
  return -logf(-logf(u + 1e-20f) + 1e-20f);
}

template <int BM, int BN, int EPI>
__global__ void __launch_bounds__(256) gemm_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ float Cs[BM][BN + 1];

  if (p.live && *p.live == 0) return;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM;
  const int c0 = blockIdx.x * BN;
  const int G = p.G;
  const int NC = p.U * G;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& sg = p.seg[s];
    for (int k0 = 0; k0 < sg.K; k0 += BK) {
      // ---- A tile -> As[k][m]
      if (!sg.a_trans) {
        for (int e = tid; e < BM * BK; e += 256) {
          const int m = e / BK, k = e % BK;
          const int gm = m0 + m, gk = k0 + k;
          float v = 0.0f;
          if (gm < p.M && gk < sg.K) {
            const long long row = sg.gather ? (long long)sg.gather[gm * sg.gather_stride] : (long long)gm;
            v = __ldg(sg.a + row * sg.lda + gk);
          }
          As[k][m] = v;
        }
      } else {
        for (int e = tid; e < BM * BK; e += 256) {
          const int k = e / BM, m = e % BM;
          const int gm = m0 + m, gk = k0 + k;
          float v = 0.0f;
          if (gm < p.M && gk < sg.K && !(sg.k_zero_period && gk % sg.k_zero_period == sg.k_zero_rem))
            v = __ldg(sg.a + (long long)gk * sg.lda + gm);
          As[k][m] = v;
        }
      }
      // ---- B tile -> Bs[k][c]
      if (!sg.w_trans) {
        for (int e = tid; e < BN * BK; e += 256) {
          const int c = e / BK, k = e % BK;
          const int gc = c0 + c, gk = k0 + k;
          float v = 0.0f;
          if (gc < NC && gk < sg.K) {
            const int g = gc % G, u = gc / G;
            const float* wp = sg.w[g];
            if (wp) v = __ldg(wp + (long long)u * sg.ldw + gk);
          }
          Bs[k][c] = v;
        }
      } else {
        for (int e = tid; e < BN * BK; e += 256) {
          const int k = e / BN, c = e % BN;
          const int gc = c0 + c, gk = k0 + k;
          float v = 0.0f;
          if (gc < NC && gk < sg.K && !(sg.k_zero_period && gk % sg.k_zero_period == sg.k_zero_rem)) {
            const int g = gc % G, u = gc / G;
            const float* wp = sg.w[g];
            if (wp) v = __ldg(wp + (long long)gk * sg.ldw + u);
          }
          Bs[k][c] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) Cs[ty * TM + i][tx * TN + j] = acc[i][j];
  __syncthreads();

  const EpiParams& ep = p.epi;
  const int U = p.U;

  if constexpr (EPI == EPI_PLAIN) {
    // iterate gate-major so that stores of one gate are coalesced along u
    const int UT = BN / G;  // units per tile (BN is a multiple of G for all supported G)
    for (int e = tid; e < G * BM * UT; e += 256) {
      const int gg = e / (BM * UT);
      const int r = (e / UT) % BM;
      const int ul = e % UT;
      const int gm = m0 + r;
      const int u = c0 / G + ul;
      if (gm < p.M && u < U && ep.c[gg]) {
        float v = ep.scale * Cs[r][ul * G + gg];
        if (ep.bias[gg]) v += __ldg(ep.bias[gg] + u);
        float* dst = ep.c[gg] + (long long)gm * ep.ldc + u;
        if (ep.accumulate) v += *dst;
        *dst = v;
      }
    }
  } else if constexpr (EPI == EPI_LSTM) {
    const int UT = BN / 4;
    for (int e = tid; e < BM * UT; e += 256) {
      const int r = e / UT, ul = e % UT;
      const int gm = m0 + r, u = c0 / 4 + ul;
      if (gm >= p.M || u >= U) continue;
      float pre[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v = Cs[r][ul * 4 + g];
        if (ep.gx) v += ep.gx[(long long)gm * ep.ld_gx + g * U + u];
        if (ep.b_ih) v += __ldg(ep.b_ih + g * U + u);
        if (ep.b_hh) v += __ldg(ep.b_hh + g * U + u);
        pre[g] = v;
      }
      const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
      const float cp = ep.prev ? ep.prev[(long long)gm * ep.ld_prev + u] : 0.0f;
      const float cn = fg * cp + ig * gg;
      const float hn = og * tanhf(cn);
      float* gs = ep.gates + (long long)gm * ep.ld_gates + u;
      gs[0] = ig; gs[U] = fg; gs[2 * U] = gg; gs[3 * U] = og;
      ep.out0[(long long)gm * ep.ld_out0 + u] = cn;
      ep.out1[(long long)gm * ep.ld_out1 + u] = hn;
    }
  } else if constexpr (EPI == EPI_GRU) {
    const int UT = BN / 4;
    for (int e = tid; e < BM * UT; e += 256) {
      const int r = e / UT, ul = e % UT;
      const int gm = m0 + r, u = c0 / 4 + ul;
      if (gm >= p.M || u >= U) continue;
      float xr = Cs[r][ul * 4 + 0], xz = Cs[r][ul * 4 + 1], xn = Cs[r][ul * 4 + 2], hn_ = Cs[r][ul * 4 + 3];
      if (ep.gx) {
        const float* gx = ep.gx + (long long)gm * ep.ld_gx + u;
        xr += gx[0]; xz += gx[U]; xn += gx[2 * U];
      } else if (ep.b_ih) {
        xr += __ldg(ep.b_ih + u); xz += __ldg(ep.b_ih + U + u); xn += __ldg(ep.b_ih + 2 * U + u);
      }
      xr += __ldg(ep.b_hh + u); xz += __ldg(ep.b_hh + U + u); hn_ += __ldg(ep.b_hh + 2 * U + u);
      const float rg = sigmoidf_(xr), zg = sigmoidf_(xz);
      const float ng = tanhf(xn + rg * hn_);
      const float hp = ep.prev ? ep.prev[(long long)gm * ep.ld_prev + u] : 0.0f;
      float hnew = (1.0f - zg) * ng + zg * hp;
      if (ep.lens && ep.t >= ep.lens[gm]) hnew = 0.0f;  // packed sequence: padded outputs are zero
      float* gs = ep.gates + (long long)gm * ep.ld_gates + u;
      gs[0] = rg; gs[U] = zg; gs[2 * U] = ng; gs[3 * U] = hn_;
      ep.out0[(long long)gm * ep.ld_out0 + u] = hnew;
      if (ep.out1) ep.out1[(long long)gm * ep.ld_out1 + u] = hnew;
    }
  } else if constexpr (EPI == EPI_HEAD) {
    const int UT = BN / 2;
    for (int e = tid; e < BM * UT; e += 256) {
      const int r = e / UT, ul = e % UT;
      const int gm = m0 + r, u = c0 / 2 + ul;
      if (gm >= p.M || u >= U) continue;
      float mean = Cs[r][ul * 2 + 0], lg = Cs[r][ul * 2 + 1];
      if (ep.bias[0]) mean += __ldg(ep.bias[0] + u);
      if (ep.bias[1]) lg += __ldg(ep.bias[1] + u);
      const float e_ = ep.eps[(long long)gm * ep.ld_eps + u];
      const float z = e_ * expf(0.5f * lg) + mean;
      ep.out0[(long long)gm * ep.ld_out0 + u] = mean;
      ep.out1[(long long)gm * ep.ld_out1 + u] = lg;
      ep.out2[(long long)gm * ep.ld_out2 + u] = z;
    }
  } else if constexpr (EPI == EPI_STATS) {
    // one warp handles rows r = warp, warp+8, ...; lanes stride the BN columns
    const int lane = tid & 31, wid = tid >> 5;
    const int ntiles = gridDim.x;
    for (int r = wid; r < BM; r += 8) {
      const int gm = m0 + r;
      if (gm >= p.M) continue;  // warp-uniform
      float vmax = -INFINITY, vsum = 0.0f, best = -INFINITY, bestlogit = 0.0f;
      int barg = 0x7fffffff;
      for (int c = lane; c < BN; c += 32) {
        const int u = c0 + c;
        if (u < U) {
          float v = Cs[r][c] + (ep.bias[0] ? __ldg(ep.bias[0] + u) : 0.0f);
          Cs[r][c] = v;
          vmax = fmaxf(vmax, v);
          vsum += v;
          float key = v;
          if (ep.noise) key = v * ep.inv_temp + gumbel_from_u(ep.noise[(long long)gm * ep.ld_noise + u]);
          if (key > best) { best = key; barg = u; bestlogit = v; }
        }
      }
      const float wmax = warp_max(vmax);
      float vexp = 0.0f;
      for (int c = lane; c < BN; c += 32)
        if (c0 + c < U) vexp += expf(Cs[r][c] - wmax);
      vexp = warp_sum(vexp);
      vsum = warp_sum(vsum);
      // arg-best with lowest-index tie break (torch.max returns the first maximum)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, barg, o);
        const float ol = __shfl_xor_sync(0xffffffffu, bestlogit, o);
        if (ob > best || (ob == best && oa < barg)) { best = ob; barg = oa; bestlogit = ol; }
      }
      if (lane == 0) {
        const long long o = (long long)gm * ntiles + blockIdx.x;
        ep.pmax[o] = wmax; ep.pexp[o] = vexp; ep.psum[o] = vsum; ep.pbest[o * 2] = best; ep.pbest[o * 2 + 1] = bestlogit;
        ep.parg[o] = barg;
      }
    }
  } else if constexpr (EPI == EPI_DLOGITS) {
    const float gs = *ep.gscale;
    for (int e = tid; e < BM * BN; e += 256) {
      const int r = e / BN, c = e % BN;
      const int gm = m0 + r, u = c0 + c;
      if (gm >= p.M || u >= U) continue;
      const float v = Cs[r][c] + (ep.bias[0] ? __ldg(ep.bias[0] + u) : 0.0f);
      const float pr = expf(v - ep.lse[gm]);
      const float td = (u == ep.targets[gm]) ? ep.smooth_on : ep.smooth_off;
      const float rw = ep.row_w ? ep.row_w[gm] : 1.0f;
      ep.c[0][(long long)gm * ep.ldc + u] = gs * rw * (pr - td);
    }
  }
}

// Host-side launch: picks a tile by M and issues the kernel.
template <int EPI>
inline int launch_gemm(const GemmParams& p, cudaStream_t st) {
  if (p.M <= 0 || p.U <= 0) return 0;
  const int NC = p.U * p.G;
  if (EPI == EPI_STATS || p.M > 48) {
    if (EPI != EPI_STATS && (long long)((p.M + 63) / 64) * ((NC + 63) / 64) < 96) {
      dim3 grid((NC + 31) / 32, (p.M + 31) / 32);
      ACVAE_LAUNCH((gemm_kernel<32, 32, EPI>), grid, 256, 0, st, p);
    } else {
      dim3 grid((NC + 63) / 64, (p.M + 63) / 64);
      ACVAE_LAUNCH((gemm_kernel<64, 64, EPI>), grid, 256, 0, st, p);
    }
  } else if (p.M > 16) {
    dim3 grid((NC + 31) / 32, (p.M + 31) / 32);
    ACVAE_LAUNCH((gemm_kernel<32, 32, EPI>), grid, 256, 0, st, p);
  } else {
    dim3 grid((NC + 31) / 32, (p.M + 15) / 16);
    ACVAE_LAUNCH((gemm_kernel<16, 32, EPI>), grid, 256, 0, st, p);
  }
  return 0;
}

// ---- small builders ---------------------------------------------------------
inline GemmSeg seg_plain(const float* a, long long lda, const float* w, long long ldw, int K) {
  GemmSeg s{};
  s.a = a; s.lda = lda; s.w[0] = w; s.ldw = ldw; s.K = K;
  return s;
}
inline GemmSeg seg_gather(const float* table, long long ld, const int* idx, long long idx_stride,
                          const float* w, long long ldw, int K) {
  GemmSeg s = seg_plain(table, ld, w, ldw, K);
  s.gather = idx; s.gather_stride = idx_stride;
  return s;
}
// gates g = 0..G-1 read weight rows [g*U, (g+1)*U) of a [G*U, ldw] matrix, columns col0..col0+K
inline GemmSeg seg_gates(const float* a, long long lda, const float* w, long long ldw, int col0, int K, int U, int G) {
  GemmSeg s{};
  s.a = a; s.lda = lda; s.ldw = ldw; s.K = K;
  for (int g = 0; g < G; ++g) s.w[g] = w + (long long)g * U * ldw + col0;
  return s;
}

}  // namespace acvae
