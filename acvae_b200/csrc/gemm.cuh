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
  long long ldwg[kMaxGate];  // optional per-gate leading dimension (0 => ldw)
  int w_trans;
  int K;
  int a_vec, w_vec;               // set by the launcher: 16-byte vector loads are legal for this operand
  int k_zero_period, k_zero_rem;  // if period > 0: k with k % period == rem contribute nothing (shifted operands)
  int w_row_shift;                // w[0] already points `w_row_shift` K-rows away from the real tensor (TMA path un-shifts)
};

enum EpiKind {
  EPI_PLAIN = 0,   // C_g[m*ldc + u] = scale*acc + bias_g[u] (+ C_g if accumulate)
  EPI_LSTM = 1,    // G=4 (i,f,g,o): nn.LSTM cell
  EPI_GRU = 2,     // G=4 (r,z,n_x,n_h): nn.GRU cell, optional packed-sequence length mask
  EPI_HEAD = 3,    // G=2 (mean,log): Gaussian head + reparameterisation
  EPI_STATS = 4,   // G=1: per-row partial max/argmax/sum-exp/sum over this CTA's columns
  EPI_DLOGITS = 5, // G=1: label-smoothed softmax-CE gradient wrt logits
  EPI_GRU_BWD = 6, // G=1: acc = dh_{t-1} partial; fused GRU pointwise backward of step t-1
  EPI_LSTM_BWD = 7,// G=1: acc = dML_t.W_head; fused LSTM pointwise backward of step t
  EPI_HEAD_BWD = 8 // G=2: (d last_z, d h_{t-1}); fused Gaussian-head/reparam backward of step t-1
};

struct EpiParams {
  // EPI_PLAIN
  float* c[kMaxGate];
  long long ldc;
  const float* bias[kMaxGate];
  int accumulate;
  int atomic;       // split-K (tc_gemm.cuh): several CTAs add partial tiles into a pre-zeroed / accumulated C
  int free_order;   // the caller accepts an order-dependent last bit (gradient GEMMs): any number of K splits.
                    // Otherwise at most two partials are added onto zeros (commutative, hence bit-reproducible)
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
  const float* noise; long long ld_noise; float inv_temp;           // optional Gumbel-max: uniform u ...
  int noise_is_gumbel;                                              // ... or the Gumbel variate itself (precomputed)
  const unsigned long long* rng; int rng_step;                      // ... or drawn here: device {seed, call}, see philox_uniform4
  // EPI_DLOGITS
  const float* lse; const int* targets; const float* row_w; const float* gscale; // gscale: device [1] = dLoss / count
  float smooth_off, smooth_on;
  // backward-chain epilogues: generic inputs x0..x5 / outputs y0..y2 (meaning documented per epilogue)
  const float* x0; long long ld_x0; const float* x1; long long ld_x1; const float* x2; long long ld_x2;
  const float* x3; long long ld_x3; const float* x4; long long ld_x4; const float* x5; long long ld_x5;
  float* y0; long long ld_y0; float* y1; long long ld_y1; float* y2; long long ld_y2;
};

struct GemmParams {
  int M, U, G, nseg;
  const int* live;   // optional device flag: the kernel returns at once when *live == 0
  GemmSeg seg[kMaxSeg];
  EpiParams epi;
};

// ---- counter-based sampling noise (no [T, N, V] tensor) ------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), key = the 64-bit seed,
// counter = (sequence row, word group, decode step, call number).  Word w of the vocabulary belongs to group
// (w / 128) * 32 + ((w % 128) / 64) * 16 + w % 16 and takes output word (w % 64) / 16 of it: the four outputs of one call
// are four columns of the SAME 64-column half tile, 16 apart, so the thread-per-row epilogue of the persistent tcgen05
// kernel (one thread = one row x 64 columns) uses every output it computes, and the lane-per-column epilogues get theirs
// with one call per lane and a shuffle.  u = (x >> 8) * 2^-24 in [0, 1).  The same function in numpy
// (oracle/philox_ref.py) lets the tests reproduce the device's draws exactly.
__device__ __forceinline__ void philox4x32_10(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                              unsigned (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    c1 = (unsigned)p1; c3 = (unsigned)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox_uniform4(const unsigned long long* rng, int step, int row, int group, float (&u)[4]) {
  const unsigned long long seed = rng[0];
  unsigned x[4];
  philox4x32_10((unsigned)seed, (unsigned)(seed >> 32), (unsigned)row, (unsigned)group, (unsigned)step, (unsigned)rng[1], x);
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = (float)(x[i] >> 8) * 5.9604644775390625e-8f;
}

__device__ __forceinline__ float gumbel_from_u(float u) {
  // reference models/word_model.py:188-190 (eps = 1e-20)
  return -logf(-logf(u + 1e-20f) + 1e-20f);
}
// the same variate for device-drawn noise, with the hardware logarithm (lg2.approx, 2^-21 absolute error on the inner log):
// the two exact logf cost ~50 instructions per logit, more than the vocabulary GEMM itself; the injected-noise path, which is
// what is pinned against the reference token by token, keeps the exact form
__device__ __forceinline__ float gumbel_from_u_fast(float u) {
  return -__logf(-__logf(u + 1e-20f) + 1e-20f);
}


// ---- epilogue shared by both kernels: Cs is the staged accumulator tile [BM][ldcs] -------------
template <int EPI, int BM, int BN, int NT = 256>
__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, const float* __restrict__ Cs_, int ldcs, int m0, int c0,
                                              int tid, int tile_x, int ntiles_x) {
  // 2-D view helper
  struct View { const float* b; int ld; __device__ const float* operator[](int r) const { return b + r * ld; } };
  // EPI_STATS rewrites the tile in place (adds the bias), hence the const_cast below
  struct MView { float* b; int ld; __device__ float* operator[](int r) const { return b + r * ld; } };
  const MView Cs{const_cast<float*>(Cs_), ldcs};
  const int G = p.G;
  const EpiParams& ep = p.epi;
  const int U = p.U;

  if constexpr (EPI == EPI_PLAIN) {
    if (G == 1 && (ldcs & 3) == 0 && (ep.ldc & 3) == 0 && (c0 & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.c[0]) & 15) == 0 &&
        c0 + BN <= U) {
      // vector path (16-byte aligned staging rows and output rows): one 16-byte store / read-modify-write /
      // red.global.add.v4.f32 per 4 columns.  The scalar split-K epilogue was RED-issue bound: 64 RED.32 per
      // thread took 8300 cycles of a 13 000-cycle launch (profiles/r1/gemm_trace.log).
      float* __restrict__ cbase = ep.c[0];
      const float* __restrict__ bias = ep.bias[0];
      const bool add_bias = bias && (!ep.atomic || blockIdx.z == 0);
      for (int e = tid; e < BM * (BN / 4); e += NT) {
        const int r = e / (BN / 4), c = (e % (BN / 4)) * 4;
        const int gm = m0 + r, u = c0 + c;
        if (gm >= p.M) continue;
        float4 v = *reinterpret_cast<const float4*>(Cs[r] + c);
        v.x *= ep.scale; v.y *= ep.scale; v.z *= ep.scale; v.w *= ep.scale;
        if (add_bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + u));
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        float* dst = cbase + (long long)gm * ep.ldc + u;
        if (ep.atomic) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
          if (ep.accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(dst);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *reinterpret_cast<float4*>(dst) = v;
        }
      }
    } else if (G == 1) {
      // common case: compile-time index arithmetic, coalesced 128-byte row segments
      float* __restrict__ cbase = ep.c[0];
      const float* __restrict__ bias = ep.bias[0];
      for (int e = tid; e < BM * BN; e += NT) {
        const int r = e / BN, c = e % BN;
        const int gm = m0 + r, u = c0 + c;
        if (gm < p.M && u < U) {
          float v = ep.scale * Cs[r][c];
          float* dst = cbase + (long long)gm * ep.ldc + u;
          if (ep.atomic) {
            if (bias && blockIdx.z == 0) v += __ldg(bias + u);
            atomicAdd(dst, v);
          } else {
            if (bias) v += __ldg(bias + u);
            if (ep.accumulate) v += *dst;
            *dst = v;
          }
        }
      }
    } else {
      // iterate gate-major so that stores of one gate are coalesced along u
      const int UT = BN / G;  // units per tile (BN is a multiple of G for all supported G)
      for (int e = tid; e < G * BM * UT; e += NT) {
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
    // one warp handles rows r = warp, warp + NT/32, ...; lanes stride the BN columns.  The sampling noise comes
    // straight from HBM (it is used once): the loads of kRows rows are issued together, otherwise every row costs
    // a full memory round trip and the epilogue dwarfs the GEMM (measured 1.06 ms vs 0.26 ms at M = 10450).
    constexpr int kRows = 4, kCols = (BN + 31) / 32, NW = NT / 32;
    const int lane = tid & 31, wid = tid >> 5;
    const int ntiles = ntiles_x;
    for (int rb = wid * kRows; rb < BM; rb += NW * kRows) {
      float nz[kRows][kCols];
      const bool noisy = ep.noise || ep.rng;
      if (ep.noise) {
#pragma unroll
        for (int i = 0; i < kRows; ++i)
#pragma unroll
          for (int q = 0; q < kCols; ++q) {
            const int gm = m0 + rb + i, u = c0 + lane + 32 * q;
            nz[i][q] = (rb + i < BM && gm < p.M && u < U && lane + 32 * q < BN) ? ep.noise[(long long)gm * ep.ld_noise + u] : 0.5f;
          }
      } else if (ep.rng) {
        // the Philox word layout is defined on 64-column half tiles: EPI_STATS only ever runs with 128- or 64-wide tiles
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
          float uu[4];
          if constexpr (BN == 128) {
            // lane L draws group 32 * tile + L (half L / 16, columns L % 16 + {0, 16, 32, 48} of it); column lane + 32 q lives
            // in half q / 2 at offset lane + 32 (q % 2): group of lane (q / 2) * 16 + lane % 16, output word lane / 16 + 2 (q % 2)
            philox_uniform4(ep.rng, ep.rng_step, m0 + rb + i, ((c0 >> 7) << 5) + lane, uu);
#pragma unroll
            for (int q = 0; q < kCols; ++q) {
              const int src = (q >> 1) * 16 + (lane & 15);
              const float u0 = __shfl_sync(0xffffffffu, uu[2 * (q & 1)], src), u1 = __shfl_sync(0xffffffffu, uu[2 * (q & 1) + 1], src);
              nz[i][q] = lane < 16 ? u0 : u1;
            }
          } else if constexpr (BN == 64) {
            philox_uniform4(ep.rng, ep.rng_step, m0 + rb + i, ((c0 >> 7) << 5) + ((c0 & 64) >> 2) + (lane & 15), uu);
#pragma unroll
            for (int q = 0; q < kCols; ++q) nz[i][q] = lane < 16 ? uu[2 * q] : uu[2 * q + 1];
          } else {
#pragma unroll
            for (int q = 0; q < kCols; ++q) nz[i][q] = 0.5f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int r = rb + i;
        const int gm = m0 + r;
        if (r >= BM || gm >= p.M) continue;  // warp-uniform
        float vmax = -INFINITY, vsum = 0.0f, best = -INFINITY, bestlogit = 0.0f;
        int barg = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < kCols; ++q) {
          const int c = lane + 32 * q;
          const int u = c0 + c;
          if (c < BN && u < U) {
            float v = Cs[r][c] + (ep.bias[0] ? __ldg(ep.bias[0] + u) : 0.0f);
            Cs[r][c] = v;
            vmax = fmaxf(vmax, v);
            vsum += v;
            float key = v;
            if (noisy) key = v * ep.inv_temp + (ep.noise_is_gumbel ? nz[i][q] : (ep.rng ? gumbel_from_u_fast(nz[i][q]) : gumbel_from_u(nz[i][q])));
            if (key > best) { best = key; barg = u; bestlogit = v; }
          }
        }
        const float wmax = warp_max(vmax);
        float vexp = 0.0f;
#pragma unroll
        for (int q = 0; q < kCols; ++q) {
          const int c = lane + 32 * q;
          if (c < BN && c0 + c < U) vexp += expf(Cs[r][c] - wmax);
        }
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
          const long long o = (long long)gm * ntiles + tile_x;
          ep.pmax[o] = wmax; ep.pexp[o] = vexp; ep.psum[o] = vsum; ep.pbest[o * 2] = best; ep.pbest[o * 2 + 1] = bestlogit;
          ep.parg[o] = barg;
        }
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
  } else if constexpr (EPI == EPI_GRU_BWD) {
    // x0: carry_in = dh_t * z_t [M,U] (or NULL); x1: upstream d h_{t-1} (ld_x1) (or NULL);
    // gates/ld_gates: saved (r,z,n,gh_n) of step t-1; prev/ld_prev: h_{t-2} (or NULL);
    // lens/t: packed-sequence mask for step t-1.  y0: dGi [.,3U]; y1: dGh [.,3U]; y2: carry_out [M,U].
    for (int e = tid; e < BM * BN; e += 256) {
      const int r = e / BN, c = e % BN;
      const int gm = m0 + r, u = c0 + c;
      if (gm >= p.M || u >= U) continue;
      float* gi = ep.y0 + (long long)gm * ep.ld_y0 + u;
      float* gh = ep.y1 + (long long)gm * ep.ld_y1 + u;
      if (ep.lens && ep.t >= ep.lens[gm]) {
        gi[0] = gi[U] = gi[2 * U] = 0.0f; gh[0] = gh[U] = gh[2 * U] = 0.0f;
        ep.y2[(long long)gm * U + u] = 0.0f;
        continue;
      }
      float dh = Cs[r][c];
      if (ep.x0) dh += ep.x0[(long long)gm * U + u];
      if (ep.x1) dh += ep.x1[(long long)gm * ep.ld_x1 + u];
      const float* g = ep.gates + (long long)gm * ep.ld_gates + u;
      const float rr = g[0], z = g[U], nn = g[2 * U], ghn = g[3 * U];
      const float hp = ep.prev ? ep.prev[(long long)gm * ep.ld_prev + u] : 0.0f;
      const float dn = dh * (1.0f - z), dz = dh * (hp - nn);
      const float dan = dn * (1.0f - nn * nn);
      const float dar = dan * ghn * rr * (1.0f - rr), daz = dz * z * (1.0f - z);
      gi[0] = dar; gi[U] = daz; gi[2 * U] = dan;
      gh[0] = dar; gh[U] = daz; gh[2 * U] = dan * rr;
      ep.y2[(long long)gm * U + u] = dh * z;
    }
  } else if constexpr (EPI == EPI_LSTM_BWD) {
    // x0: dh carry [M,U] (or NULL); x1: dc carry in [M,U] (or NULL); gates: saved (i,f,g,o) of step t;
    // x2/ld_x2: c_t; x3/ld_x3: c_{t-1} (or NULL).  y0: dG [.,4U]; y1: dc carry out [M,U].
    for (int e = tid; e < BM * BN; e += 256) {
      const int r = e / BN, c = e % BN;
      const int gm = m0 + r, u = c0 + c;
      if (gm >= p.M || u >= U) continue;
      float dh = Cs[r][c];
      if (ep.x0) dh += ep.x0[(long long)gm * U + u];
      const float* g = ep.gates + (long long)gm * ep.ld_gates + u;
      const float ig = g[0], fg = g[U], gg = g[2 * U], og = g[3 * U];
      const float tc = tanhf(ep.x2[(long long)gm * ep.ld_x2 + u]);
      const float cp = ep.x3 ? ep.x3[(long long)gm * ep.ld_x3 + u] : 0.0f;
      float dc = dh * og * (1.0f - tc * tc);
      if (ep.x1) dc += ep.x1[(long long)gm * U + u];
      float* dg = ep.y0 + (long long)gm * ep.ld_y0 + u;
      dg[0] = dc * gg * ig * (1.0f - ig);
      dg[U] = dc * cp * fg * (1.0f - fg);
      dg[2 * U] = dc * ig * (1.0f - gg * gg);
      dg[3 * U] = dh * tc * og * (1.0f - og);
      ep.y1[(long long)gm * U + u] = dc * fg;
    }
  } else if constexpr (EPI == EPI_HEAD_BWD) {
    // columns interleave (dz, dh) per unit.  x0/ld_x0: upstream d p_z[t-1] (or NULL); x1: d p_means[t-1] (or NULL);
    // x2: d p_logs[t-1] (or NULL); x3/ld_x3: eps[t-1]; x4/ld_x4: p_logs[t-1].
    // y0/ld_y0: dML[t-1] [.,2U]; y1: dh carry out [M,U].
    const int UT = BN / 2;
    for (int e = tid; e < BM * UT; e += 256) {
      const int r = e / UT, ul = e % UT;
      const int gm = m0 + r, u = c0 / 2 + ul;
      if (gm >= p.M || u >= U) continue;
      float dz = Cs[r][ul * 2 + 0];
      const float dh = Cs[r][ul * 2 + 1];
      if (ep.x0) dz += ep.x0[(long long)gm * ep.ld_x0 + u];
      float dm = dz;
      float dl = dz * ep.x3[(long long)gm * ep.ld_x3 + u] * 0.5f * expf(0.5f * ep.x4[(long long)gm * ep.ld_x4 + u]);
      if (ep.x1) dm += ep.x1[(long long)gm * ep.ld_x1 + u];
      if (ep.x2) dl += ep.x2[(long long)gm * ep.ld_x2 + u];
      float* dml = ep.y0 + (long long)gm * ep.ld_y0 + u;
      dml[0] = dm; dml[U] = dl;
      ep.y1[(long long)gm * U + u] = dh;
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int n = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int n = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// =================================================================================================
// Skinny kernel: the recurrent chain's GEMMs have M = batch <= 32..64 rows.  One CTA owns all 32 rows
// of a BN-column slab and the FULL K; K is split over the 8 warps (split-K inside the CTA, reduced
// through shared memory), operands are staged with double-buffered cp.async in KC-wide chunks.  The
// point is latency: a handful of L2 round trips per launch instead of K/16 synchronous tile loads.
// =================================================================================================
constexpr int kSkinnyKC = 128;
constexpr int kSkinnyLD = kSkinnyKC + 4;
constexpr int kSkinnyMaxStages = 8;

__device__ __forceinline__ void cp_async_wait_dyn(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::); break;
    case 3: asm volatile("cp.async.wait_group 3;\n" ::); break;
    case 4: asm volatile("cp.async.wait_group 4;\n" ::); break;
    case 5: asm volatile("cp.async.wait_group 5;\n" ::); break;
    case 6: asm volatile("cp.async.wait_group 6;\n" ::); break;
    default: asm volatile("cp.async.wait_group 7;\n" ::); break;
  }
}

// Stage s of the ring: A tile [32][LD] followed by W tile [BN][LD].  Up to 8 stages (K <= 1024) are all
// issued before the first wait, so the whole operand set of a recurrent-step GEMM is one burst of
// cp.async traffic and the CTA pays a single L2 round trip.
template <int BN, int EPI>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(const __grid_constant__ GemmParams p, int nstages) {
  constexpr int BM = 32, KC = kSkinnyKC, LD = kSkinnyLD, NJ = BN / 4;
  constexpr int STAGE = (BM + BN) * LD;
  extern __shared__ __align__(16) float sk_smem[];
  __shared__ long long rowoff[kMaxSeg][BM];
  if (p.live && *p.live == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
  const int G = p.G, NC = p.U * G;

  for (int i = tid; i < p.nseg * BM; i += 256) {
    const int s = i / BM, r = i % BM;
    const GemmSeg& sg = p.seg[s];
    const int gm = m0 + r;
    long long off = -1;
    if (gm < p.M) off = (sg.gather ? (long long)sg.gather[gm * sg.gather_stride] : (long long)gm) * sg.lda;
    rowoff[s][r] = off;
  }
  __syncthreads();

  int nchunk = 0;
  for (int s = 0; s < p.nseg; ++s) nchunk += (p.seg[s].K + KC - 1) / KC;

  auto issue = [&](int chunk) {
    float* As = sk_smem + (size_t)(chunk % nstages) * STAGE;
    float* Ws = As + BM * LD;
    int s = 0, k0 = 0, c = chunk;
    for (; s < p.nseg; ++s) {
      const int n = (p.seg[s].K + KC - 1) / KC;
      if (c < n) { k0 = c * KC; break; }
      c -= n;
    }
    const GemmSeg& sg = p.seg[s];
#pragma unroll
    for (int i = 0; i < (BM * KC / 4) / 256; ++i) {
      const int idx = tid + i * 256;
      const int r = idx / (KC / 4), k4 = (idx % (KC / 4)) * 4;
      const int gk = k0 + k4;
      const long long off = rowoff[s][r];
      const bool ok = off >= 0 && gk < sg.K;
      cp_async16(As + r * LD + k4, ok ? (const void*)(sg.a + off + gk) : (const void*)sg.a, ok);
    }
    if (!sg.w_trans) {
#pragma unroll
      for (int i = 0; i < (BN * KC / 4) / 256; ++i) {
        const int idx = tid + i * 256;
        const int c_ = idx / (KC / 4), k4 = (idx % (KC / 4)) * 4;
        const int gc = c0 + c_, gk = k0 + k4;
        const int g = gc % G, u = gc / G;
        const float* wp = gc < NC ? sg.w[g] : nullptr;
        const bool ok = wp != nullptr && gk < sg.K;
        cp_async16(Ws + c_ * LD + k4, ok ? (const void*)(wp + (long long)u * (sg.ldwg[g] ? sg.ldwg[g] : sg.ldw) + gk) : (const void*)sg.a, ok);
      }
    } else {
#pragma unroll
      for (int i = 0; i < (BN * KC) / 256; ++i) {
        const int idx = tid + i * 256;
        const int k = idx / BN, c_ = idx % BN;
        const int gc = c0 + c_, gk = k0 + k;
        const int g = gc % G, u = gc / G;
        const float* wp = gc < NC ? sg.w[g] : nullptr;
        const bool ok = wp != nullptr && gk < sg.K;
        cp_async4(Ws + c_ * LD + k, ok ? (const void*)(wp + (long long)gk * (sg.ldwg[g] ? sg.ldwg[g] : sg.ldw) + u) : (const void*)sg.a, ok);
      }
    }
    cp_async_commit();
  };

  float acc[4][NJ];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.0f;
  const int rg = lane >> 2, cg = lane & 3;

  int issued = 0;
  for (; issued < min(nchunk, nstages); ++issued) issue(issued);
  for (int ch = 0; ch < nchunk; ++ch) {
    cp_async_wait_dyn(issued - 1 - ch);
    __syncthreads();
    const float* As = sk_smem + (size_t)(ch % nstages) * STAGE;
    const float* Ws = As + BM * LD;
#pragma unroll
    for (int kk = 0; kk < KC / 32; ++kk) {
      const int k = wid * (KC / 8) + kk * 4;
      float4 a[4], w[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (rg + 8 * i) * LD + k);
#pragma unroll
      for (int j = 0; j < NJ; ++j) w[j] = *reinterpret_cast<const float4*>(Ws + (cg + 4 * j) * LD + k);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
        }
    }
    if (issued < nchunk) {        // ring reuse (K > nstages * KC): refill the slot just consumed
      __syncthreads();
      issue(issued);
      ++issued;
    }
  }
  __syncthreads();
  // cross-warp split-K reduction through shared memory (reuses stage 0 / stage 1)
  float* part = sk_smem;                       // [8][BM][BN]
  static_assert(8 * BM * BN <= STAGE, "partial buffer does not fit in one stage");
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) part[(wid * BM + rg + 8 * i) * BN + cg + 4 * j] = acc[i][j];
  __syncthreads();
  float* Cs = sk_smem + STAGE;                 // [BM][BN+1]  (the launcher always provides >= 2 stages)
  for (int e = tid; e < BM * BN; e += 256) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w * BM * BN + e];
    Cs[(e / BN) * (BN + 1) + (e % BN)] = s;
  }
  __syncthreads();
  gemm_epilogue<EPI, BM, BN>(p, Cs, BN + 1, m0, c0, tid, blockIdx.x, gridDim.x);
}

template <int BN, int EPI>
inline int launch_skinny(const GemmParams& p, int NC, cudaStream_t st) {
  int nchunk = 0;
  for (int s = 0; s < p.nseg; ++s) nchunk += (p.seg[s].K + kSkinnyKC - 1) / kSkinnyKC;
  int nstages = nchunk < 2 ? 2 : (nchunk > kSkinnyMaxStages ? kSkinnyMaxStages : nchunk);
  const size_t smem = (size_t)nstages * (32 + BN) * kSkinnyLD * sizeof(float);
  static size_t configured_dev[kMaxDevices] = {0};
  size_t& configured = configured_dev[current_device()];
  if (smem > configured) {
    ACVAE_CHECK(cudaFuncSetAttribute(skinny_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((NC + BN - 1) / BN, (p.M + 31) / 32);
  ACVAE_LAUNCH((skinny_gemm_kernel<BN, EPI>), grid, 256, smem, st, p, nstages);
  return 0;
}

// =================================================================================================
// Tiled kernel for the batched contractions (M = N*T rows or weight-gradient shapes): 64x64x16 tiles,
// 4x4 register blocking, next tile prefetched into registers while the current one is multiplied.
// =================================================================================================
template <int EPI>
__global__ void __launch_bounds__(256) tiled_gemm_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  __shared__ float Cs[BM][BN + 1];
  if (p.live && *p.live == 0) return;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
  const int G = p.G, NC = p.U * G;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // per-thread load coordinates: non-transposed operand -> (row = tid/4, k4 = (tid%4)*4);
  //                              transposed operand     -> (k = tid/16, x4 = (tid%16)*4)
  const int nt_r = tid >> 2, nt_k = (tid & 3) * 4;
  const int tr_k = tid >> 4, tr_x = (tid & 15) * 4;

  int buf = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& sg = p.seg[s];
    // hoisted row pointers for this segment
    const float* a_row = nullptr;
    if (!sg.a_trans) {
      const int gm = m0 + nt_r;
      if (gm < p.M) a_row = sg.a + (sg.gather ? (long long)sg.gather[gm * sg.gather_stride] : (long long)gm) * sg.lda;
    }
    const float* w_row = nullptr;
    if (!sg.w_trans) {
      const int gc = c0 + nt_r;
      if (gc < NC) { const int g_ = gc % G; const float* wp = sg.w[g_]; if (wp) w_row = wp + (long long)(gc / G) * (sg.ldwg[g_] ? sg.ldwg[g_] : sg.ldw); }
    }
    const int ntile = (sg.K + BK - 1) / BK;
    float4 ra, rb;
    auto kz = [&](int gk) { return sg.k_zero_period && gk % sg.k_zero_period == sg.k_zero_rem; };
    auto fetch = [&](int kt) {
      const int k0 = kt * BK;
      ra = make_float4(0.f, 0.f, 0.f, 0.f); rb = ra;
      if (!sg.a_trans) {
        const int gk = k0 + nt_k;
        if (a_row && gk < sg.K) {
          if (sg.a_vec) ra = __ldg(reinterpret_cast<const float4*>(a_row + gk));
          else { ra.x = __ldg(a_row + gk); if (gk + 1 < sg.K) ra.y = __ldg(a_row + gk + 1);
                 if (gk + 2 < sg.K) ra.z = __ldg(a_row + gk + 2); if (gk + 3 < sg.K) ra.w = __ldg(a_row + gk + 3); }
        }
      } else {
        const int gk = k0 + tr_k, gm = m0 + tr_x;
        if (gk < sg.K && gm < p.M && !kz(gk)) {
          const float* src = sg.a + (long long)gk * sg.lda + gm;
          if (sg.a_vec && gm + 3 < p.M) ra = __ldg(reinterpret_cast<const float4*>(src));
          else { ra.x = __ldg(src); if (gm + 1 < p.M) ra.y = __ldg(src + 1); if (gm + 2 < p.M) ra.z = __ldg(src + 2);
                 if (gm + 3 < p.M) ra.w = __ldg(src + 3); }
        }
      }
      if (!sg.w_trans) {
        const int gk = k0 + nt_k;
        if (w_row && gk < sg.K) {
          if (sg.w_vec) rb = __ldg(reinterpret_cast<const float4*>(w_row + gk));
          else { rb.x = __ldg(w_row + gk); if (gk + 1 < sg.K) rb.y = __ldg(w_row + gk + 1);
                 if (gk + 2 < sg.K) rb.z = __ldg(w_row + gk + 2); if (gk + 3 < sg.K) rb.w = __ldg(w_row + gk + 3); }
        }
      } else {
        const int gk = k0 + tr_k, gc = c0 + tr_x;
        if (gk < sg.K && gc < NC && !kz(gk)) {
          if (G == 1) {
            const float* src = sg.w[0] + (long long)gk * sg.ldw + gc;
            if (sg.w_vec && gc + 3 < NC) rb = __ldg(reinterpret_cast<const float4*>(src));
            else { rb.x = __ldg(src); if (gc + 1 < NC) rb.y = __ldg(src + 1); if (gc + 2 < NC) rb.z = __ldg(src + 2);
                   if (gc + 3 < NC) rb.w = __ldg(src + 3); }
          } else {
            float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c_ = gc + q;
              if (c_ < NC) { const int g_ = c_ % G; const float* wp = sg.w[g_]; if (wp) t4[q] = __ldg(wp + (long long)gk * (sg.ldwg[g_] ? sg.ldwg[g_] : sg.ldw) + c_ / G); }
            }
            rb = make_float4(t4[0], t4[1], t4[2], t4[3]);
          }
        }
      }
    };
    auto stash = [&](int b) {
      if (!sg.a_trans) { As[b][nt_k][nt_r] = ra.x; As[b][nt_k + 1][nt_r] = ra.y; As[b][nt_k + 2][nt_r] = ra.z; As[b][nt_k + 3][nt_r] = ra.w; }
      else *reinterpret_cast<float4*>(&As[b][tr_k][tr_x]) = ra;
      if (!sg.w_trans) { Bs[b][nt_k][nt_r] = rb.x; Bs[b][nt_k + 1][nt_r] = rb.y; Bs[b][nt_k + 2][nt_r] = rb.z; Bs[b][nt_k + 3][nt_r] = rb.w; }
      else *reinterpret_cast<float4*>(&Bs[b][tr_k][tr_x]) = rb;
    };
    if (ntile > 0) { fetch(0); stash(buf); }
    __syncthreads();
    for (int kt = 0; kt < ntile; ++kt) {
      if (kt + 1 < ntile) fetch(kt + 1);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (kt + 1 < ntile) stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  gemm_epilogue<EPI, BM, BN>(p, &Cs[0][0], BN + 1, m0, c0, tid, blockIdx.x, gridDim.x);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Host-side launch: annotates vector-load legality, then picks the kernel by shape.
template <int EPI>
inline int launch_gemm_simt(GemmParams p, cudaStream_t st) {
  if (p.M <= 0 || p.U <= 0) return 0;
  const int NC = p.U * p.G;
  bool skinny_ok = EPI != EPI_STATS && EPI != EPI_DLOGITS && p.M <= 64;
  for (int s = 0; s < p.nseg; ++s)
    for (int g = 0; g < p.G; ++g)
      if (p.seg[s].ldwg[g] % 4 != 0) skinny_ok = false;
  for (int s = 0; s < p.nseg; ++s) {
    GemmSeg& sg = p.seg[s];
    bool wal = true;
    for (int g = 0; g < p.G; ++g) wal = wal && (sg.w[g] == nullptr || aligned16(sg.w[g]));
    if (!sg.a_trans) sg.a_vec = aligned16(sg.a) && sg.lda % 4 == 0 && sg.K % 4 == 0;
    else sg.a_vec = aligned16(sg.a) && sg.lda % 4 == 0;
    if (!sg.w_trans) sg.w_vec = wal && sg.ldw % 4 == 0 && sg.K % 4 == 0;
    else sg.w_vec = wal && sg.ldw % 4 == 0 && p.G == 1;
    skinny_ok = skinny_ok && !sg.a_trans && sg.a_vec && (sg.w_trans || sg.w_vec) && sg.k_zero_period == 0;
  }
  if (skinny_ok) {
    if (NC >= 512) return launch_skinny<16, EPI>(p, NC, st);
    return launch_skinny<8, EPI>(p, NC, st);
  } else {
    dim3 grid((NC + 63) / 64, (p.M + 63) / 64);
    ACVAE_LAUNCH((tiled_gemm_kernel<EPI>), grid, 256, 0, st, p);
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
