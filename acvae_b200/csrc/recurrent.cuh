// Persistent recurrent-chain kernels: a whole T-step chain of the training step in ONE cooperative launch.
//
// At batch 32 every recurrent step is a [32 x K] . [K x cols] product followed by a pointwise cell -- a few
// MFLOP that a launch-per-step schedule spends 9-22 us on (launch, weight fetch from L2, drain).  Here the
// chain runs inside one kernel of kChainCtas CTAs:
//   * the weight matrix is COLUMN-partitioned: CTA b owns hidden units {2b, 2b+1} (all their gates) and keeps
//     that weight slice resident in shared memory for all T steps -- weights are read from L2 exactly once;
//   * per step a CTA reads the [N, K] state rows written by all CTAs in the previous phase (L2, ld.cg),
//     computes its columns for all N rows (a warp owns four rows, a lane a 1/32 slice of K; reduce-scatter over the
//     warp), applies the cell arithmetic in registers and writes its units;
//   * phases are separated by a grid barrier (one global atomic counter);
//   * per-unit carries (dh*z, dc*f, ...) never leave registers: the unit partition is the same in every phase.
// The decoder chains add a per-clip phase: CTA c keeps clip c's projected memory P_d[c] and memory mem[c]
// resident in shared memory (Te*(A+E)*4 bytes) and runs the additive attention / its backward for that clip.
//
// Reference arithmetic: models/text_encoder.py:182-216 (posterior biGRU), :247-268 (prior LSTM + head),
// models/decoder.py:175-203 + models/attn_model.py:20-46 (decoder step); same saved activations and the same
// formulas as the launch-per-step schedule in train.cuh / train_fast.cuh (which remains the general path).
#pragma once
#include "common.cuh"

namespace acvae {

constexpr int kChainE = 256;                 // E == H == Hq == A handled by the persistent kernels
constexpr int kChainCtas = kChainE / 2;      // 2 hidden units per CTA
constexpr int kChainThreads = 256;           // 8 warps: warp = rows 4w .. 4w+3, lane = K slice (quad mapping below)
constexpr int kChainMaxN = kChainThreads / 8;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct GridBar {
  unsigned* counter;  // one arrival counter (zeroed before the launch); the target grows by nctas per barrier
  unsigned target;
  unsigned nctas;
};
// Barrier over all CTAs of the (cooperatively launched, hence co-resident) grid: bar.sync; thread 0: release
// increment of one counter, acquire polling of the counter; bar.sync.  Measured on B200 with 128 CTAs
// (profiles/ubench_gridbar.cu): 2400-2500 cycles for this form, 5200 for per-CTA flags polled by every CTA (hot L2
// lines), 8800 with a store before / loads after it.  Writes made by any thread of a CTA before the barrier
// happen-before every thread of every CTA after it (release / acquire on the counter, bar.sync on both sides).
// A protocol bug traps (after 10 s of wall clock) instead of hanging the GPU.
__device__ __forceinline__ void grid_sync(GridBar& gb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gb.target += gb.nctas;
    // release-increment without a return value: orders the CTA's earlier writes (bar.sync makes them
    // happen-before this thread) and does not wait for the atomic's round trip
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(gb.counter) : "memory");
    // acquire poll: pairs with the producers' release-increments, so the CTA's later reads of their data are ordered by
    // the memory model (the ld.global.cg data loads below stay as an optimisation -- L2 hits, no L1 pollution -- and are
    // no longer what correctness rests on).  Same cost as the volatile poll (profiles/r1/ubench_gridbar.log: 2430 vs 2400).
    unsigned seen;
    const unsigned long long t0 = globaltimer_ns();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(gb.counter) : "memory");
      // a protocol bug must abort instead of hanging the GPU; the limit is wall-clock (10 s), so preemption, MPS
      // time-slicing or a debugger do not trip it
      if (seen < gb.target && globaltimer_ns() - t0 > 10000000000ull) __trap();
    } while (seen < gb.target);
  }
  __syncthreads();
}

// A step's phase is latency-bound: ONE L2 round trip for all of its operands, not one per load.  The row loads
// are therefore issued as a block of `asm volatile` ld.global.cg (program order is kept, so nothing is sunk
// next to its use) into registers, and consumed afterwards.
__device__ __forceinline__ float4 ldcg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldcg1(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
  return v;
}
// ---- "quad" mapping (4 rows per warp, 32-way K split) ---------------------------------------------------------
// A warp = one group of four batch rows; lane = K slice (float4s {lane, lane+32, ...} of a row).  Every weight word
// a lane reads from shared memory is used for four rows from registers: the row-per-8-lanes mapping above re-reads
// each weight word in all four row groups of a warp (LDS.128 is served per quarter warp), which made the phases
// shared-memory bound: 2800 cycles of LDS for 770 cycles of FMA per posterior step (profiles/ubench_chain.cu).
template <int K>
__device__ __forceinline__ void quadload(const float* __restrict__ src, int lane, float4 (&a)[K / 128]) {
#pragma unroll
  for (int i = 0; i < K / 128; ++i) a[i] = ldcg4(src + (i * 32 + lane) * 4);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& w, float acc) {
  acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc); acc = fmaf(a.z, w.z, acc); return fmaf(a.w, w.w, acc);
}
// Reduce-scatter over the 32 lanes of a warp: every lane holds partial sums v[q][g] of 16 "combos" q with G_ values
// each; afterwards lanes 2q and 2q+1 hold the full sums of combo q in v[0][*].  48 shuffles for 16 x 3 values
// instead of 5 x 48 for a butterfly all-reduce; only static register indices.
template <int G_>
__device__ __forceinline__ void reduce_scatter16(float (&v)[16][G_], int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool hi = (lane & (half * 2)) != 0;
#pragma unroll
    for (int q = 0; q < half; ++q)
#pragma unroll
      for (int g = 0; g < G_; ++g) {
        const float keep = hi ? v[q + half][g] : v[q][g];
        const float send = hi ? v[q][g] : v[q + half][g];
        v[q][g] = keep + __shfl_xor_sync(0xffffffffu, send, half * 2);
      }
  }
#pragma unroll
  for (int g = 0; g < G_; ++g) v[0][g] += __shfl_xor_sync(0xffffffffu, v[0][g], 1);
}

// 8 combos q = (row r, unit j) x G_ values: afterwards lanes 4q .. 4q+3 hold combo q in v[0][*]
template <int G_>
__device__ __forceinline__ void reduce_scatter8(float (&v)[8][G_], int lane) {
#pragma unroll
  for (int half = 4; half >= 1; half >>= 1) {
    const bool hi = (lane & (half * 4)) != 0;
#pragma unroll
    for (int q = 0; q < half; ++q)
#pragma unroll
      for (int g = 0; g < G_; ++g) {
        const float keep = hi ? v[q + half][g] : v[q][g];
        const float send = hi ? v[q][g] : v[q + half][g];
        v[q][g] = keep + __shfl_xor_sync(0xffffffffu, send, half * 4);
      }
  }
#pragma unroll
  for (int g = 0; g < G_; ++g) {
    v[0][g] += __shfl_xor_sync(0xffffffffu, v[0][g], 2);
    v[0][g] += __shfl_xor_sync(0xffffffffu, v[0][g], 1);
  }
}
// v[r*2 + j][g] += sum_k a_r[k] * W[g*2 + j][k] over this lane's K slice; W: shared [2*G_][ldw]
template <int G_, int K>
__device__ __forceinline__ void quadfma(const float4 (&a)[4][K / 128], const float* W, int ldw, int lane, float (&v)[8][G_]) {
#pragma unroll
  for (int c = 0; c < 2 * G_; ++c)
#pragma unroll
    for (int i = 0; i < K / 128; ++i) {
      const float4 w = *(reinterpret_cast<const float4*>(W + c * ldw) + i * 32 + lane);
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r * 2 + (c & 1)][c >> 1] = dot4(a[r][i], w, v[r * 2 + (c & 1)][c >> 1]);
    }
}
// rows n0 .. n0+3 (row stride `stride` floats, K contiguous floats each) against W; loads first, then the FMAs
template <int K>
__device__ __forceinline__ void quadrows(int n0, int N, const float* __restrict__ src0, long long stride, int lane, float4 (&a)[4][K / 128]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (n0 + r < N) quadload<K>(src0 + r * stride, lane, a[r]);
    else {
#pragma unroll
      for (int i = 0; i < K / 128; ++i) a[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
template <int G_, int K>
__device__ __forceinline__ void quaddot(int n0, int N, const float* __restrict__ src0, long long stride, const float* W, int ldw,
                                        int lane, float (&v)[8][G_]) {
  float4 a[4][K / 128];
  quadrows<K>(n0, N, src0, stride, lane, a);
  __syncwarp();
  quadfma<G_, K>(a, W, ldw, lane, v);
}
template <int G_>
__device__ __forceinline__ void zero8(float (&v)[8][G_]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int g = 0; g < G_; ++g) v[q][g] = 0.0f;
}
// Thread roles of the quad mapping inside a chain kernel: the warp's rows, and -- for the lanes that run the
// pointwise epilogue after reduce_scatter8 -- the (row, unit) they own for the whole kernel (so per-unit carries
// stay in that lane's registers).
struct QuadRole {
  int lane, n0;     // K slice, first row of the warp
  int n, j, u;      // epilogue row, unit index within the CTA (0/1), global unit
  bool epi;         // this lane runs the epilogue
};
__device__ __forceinline__ QuadRole quad_role(int N, int u0) {
  QuadRole q;
  q.lane = threadIdx.x & 31; q.n0 = (threadIdx.x >> 5) * 4;
  const int c = q.lane >> 2;
  q.n = q.n0 + (c >> 1); q.j = c & 1; q.u = u0 + q.j;
  q.epi = (q.lane & 3) == 0 && q.n < N;
  return q;
}

// =====================================================================================================
// posterior biGRU forward, both directions in one kernel (text_encoder.py:188-191)
// =====================================================================================================
struct PostChainFwd {
  int N, T;
  const float* gx[2];    // [N,T,3E] input-side pre-activations incl. b_ih
  const float* whh[2];   // [3E,E]
  const float* bhh[2];   // [3E]
  const int* lens;       // [N] valid steps
  float* ho;             // [N,T,2E]
  float* gq[2];          // [N,T,4E] saved (r,z,n,gh_n)
  unsigned* bar;
  long long* trace;      // optional [T][8] clock64 stamps of thread 0 of CTA 0 (profiles/ubench_chain.cu) or NULL
};
__global__ void __launch_bounds__(kChainThreads) post_chain_fwd_kernel(const __grid_constant__ PostChainFwd p) {
  constexpr int E = kChainE;
  __shared__ __align__(16) float W[2][6][E];            // W[dir][gate*2 + j][k] = W_hh[dir][gate*E + u0 + j][k]
  const int tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
  const int u0 = blockIdx.x * 2;
  for (int i = tid; i < 2 * 6 * E; i += kChainThreads) {
    const int dir = i / (6 * E), r = (i / E) % 6, k = i % E;
    W[dir][r][k] = p.whh[dir][(long long)((r >> 1) * E + u0 + (r & 1)) * E + k];
  }
  __syncthreads();
  GridBar gb{p.bar, 0u, gridDim.x};
  const int T = p.T;
  const int n0 = wq * 4;                                 // this warp's rows n0 .. n0+3
  // epilogue role of the even lanes after the reduce-scatter: combo q = lane >> 1 = (row r, direction, unit j)
  const int q = lane >> 1, er = q >> 2, dir = (q >> 1) & 1, j = q & 1, u = u0 + j;
  const int n = n0 + er;
  const bool epi = (lane & 1) == 0 && n < p.N;
  const int len = epi ? p.lens[n] : 0;
  const float bh_r = p.bhh[dir][u], bh_z = p.bhh[dir][E + u], bh_n = p.bhh[dir][2 * E + u];
  for (int s = 0; s < T; ++s) {
    float v[16][3];                                      // v[r*4 + dir*2 + j][gate]
#pragma unroll
    for (int c = 0; c < 16; ++c) { v[c][0] = 0.0f; v[c][1] = 0.0f; v[c][2] = 0.0f; }
    const int t = dir ? T - 1 - s : s;
    const int tp = dir ? t + 1 : t - 1;
    const bool tr = p.trace && blockIdx.x == 0 && tid == 0;
    if (tr) p.trace[s * 8 + 0] = clock64();
    float gxr = 0.f, gxz = 0.f, gxn = 0.f, hp = 0.f;
    if (epi) {                                           // pointwise operands: issued with the row loads
      const float* gx = p.gx[dir] + ((long long)n * T + t) * 3 * E + u;
      gxr = ldcg1(gx); gxz = ldcg1(gx + E); gxn = ldcg1(gx + 2 * E);
      if (s > 0) hp = ldcg1(p.ho + ((long long)n * T + tp) * 2 * E + dir * E + u);
    }
    if (s > 0) {
      float4 a[4][2][E / 128];                           // [row][dir][float4]
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (n0 + r < p.N) {
          quadload<E>(p.ho + ((long long)(n0 + r) * T + (s - 1)) * 2 * E, lane, a[r][0]);
          quadload<E>(p.ho + ((long long)(n0 + r) * T + (T - s)) * 2 * E + E, lane, a[r][1]);
        } else {
#pragma unroll
          for (int i = 0; i < E / 128; ++i) { a[r][0][i] = make_float4(0.f, 0.f, 0.f, 0.f); a[r][1][i] = a[r][0][i]; }
        }
      }
      __syncwarp();
      if (tr) {                                          // stamp 5: this thread's row loads have landed
        float chk = 0.0f;
#pragma unroll
        for (int r = 0; r < 4; ++r) chk += a[r][0][0].x + a[r][1][0].x + a[r][0][1].x + a[r][1][1].x;
        if (chk == 1234.5f) p.trace[s * 8 + 6] = 0;
        p.trace[s * 8 + 5] = clock64();
      }
#pragma unroll
      for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
          for (int i = 0; i < E / 128; ++i) {
            const float4 w = *(reinterpret_cast<const float4*>(&W[d][c][0]) + i * 32 + lane);
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r * 4 + d * 2 + (c & 1)][c >> 1] = dot4(a[r][d][i], w, v[r * 4 + d * 2 + (c & 1)][c >> 1]);
          }
      if (tr) p.trace[s * 8 + 1] = clock64();
      reduce_scatter16<3>(v, lane);
    }
    if (tr) p.trace[s * 8 + 2] = clock64();
    if (epi) {
      const float hr = v[0][0] + bh_r, hz = v[0][1] + bh_z, hn = v[0][2] + bh_n;
      const float rg = sigmoidf_(gxr + hr), zg = sigmoidf_(gxz + hz);
      const float ng = tanhf(gxn + rg * hn);
      float hnew = (1.0f - zg) * ng + zg * hp;
      if (t >= len) hnew = 0.0f;                         // packed sequence: padded outputs are zero
      float* gs = p.gq[dir] + ((long long)n * T + t) * 4 * E + u;
      gs[0] = rg; gs[E] = zg; gs[2 * E] = ng; gs[3 * E] = hn;
      p.ho[((long long)n * T + t) * 2 * E + dir * E + u] = hnew;
    }
    if (tr) p.trace[s * 8 + 3] = clock64();
    if (s + 1 < T) grid_sync(gb);
    if (tr) p.trace[s * 8 + 4] = clock64();
  }
}

// =====================================================================================================
// prior LSTM + Gaussian head forward (text_encoder.py:253-262); word attention and the input-side gate
// pre-activations are hoisted (train_fast.cuh)
// =====================================================================================================
struct PriorChainFwd {
  int N, T;
  const float* gx;       // [N,T,4E] = [xe | ctx] . W_ih[:, :2E]^T + b_ih
  const float* wih;      // [4E,3E] (columns 2E..3E act on last_z)
  const float* whh;      // [4E,E]
  const float* bhh;      // [4E]
  const float* head_w;   // [2E,E]
  const float* head_b;   // [2E]
  const float* eps;      // [T,N,E]
  float *gates, *c, *h;  // [N,T,4E], [N,T,E], [N,T,E]
  float *pm, *pl, *pz;   // [N,T,E]
  unsigned* bar;
  int t0, t1;            // steps [t0, t1) (t1 == 0: T); t0 > 0 resumes from the saved state of step t0 - 1 (stand-alone kernel only)
};
// The prior chain as phase functions: used by the stand-alone kernel below and, merged phase by phase, by the
// decoder kernel (the two chains are independent and have the same barrier structure, and two cooperative
// kernels never overlap on the device, so sharing the barriers hides the prior chain completely).
struct PriorFwdRegs { float bh[4]; float hb_m, hb_l, c_prev, eps_t; };
constexpr int kPriorFwdSmemFloats = 8 * 2 * kChainE + 4 * kChainE;   // Wl[8][2E] | Wh[4][E]

__device__ __forceinline__ void prior_fwd_setup(const PriorChainFwd& p, float* Wl, float* Wh, PriorFwdRegs& r, int u0, int u) {
  constexpr int E = kChainE;
  for (int i = threadIdx.x; i < 8 * 2 * E; i += kChainThreads) {        // rows g*2+j over K = [last_z | h]
    const int rr = i / (2 * E), k = i % (2 * E);
    const long long wr = (long long)((rr >> 1) * E + u0 + (rr & 1));
    Wl[i] = k < E ? p.wih[wr * 3 * E + 2 * E + k] : p.whh[wr * E + (k - E)];
  }
  for (int i = threadIdx.x; i < 4 * E; i += kChainThreads) {            // rows (mean j0, mean j1, log j0, log j1)
    const int rr = i / E, k = i % E;
    Wh[i] = p.head_w[(long long)((rr >> 1) * E + u0 + (rr & 1)) * E + k];
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) r.bh[g] = p.bhh[g * E + u];
  r.hb_m = p.head_b[u]; r.hb_l = p.head_b[E + u];
  r.c_prev = 0.0f; r.eps_t = 0.0f;
}
// LSTM cell of step t (needs z_{t-1}, h_{t-1} of all units: one barrier after the previous head phase).
// Split into the operand loads and the arithmetic so that a caller can put independent work (the decoder's attention)
// between the two: the L2 round trip of the rows then overlaps that work instead of following it.
struct PriorLstmOps { float4 a0[4][kChainE / 128], a1[4][kChainE / 128]; float gxv[4]; };
__device__ __forceinline__ void prior_fwd_lstm_load(const PriorChainFwd& p, PriorFwdRegs& r, int t, const QuadRole& q, PriorLstmOps& o) {
  constexpr int E = kChainE;
  const int T = p.T, N = p.N;
#pragma unroll
  for (int g = 0; g < 4; ++g) o.gxv[g] = 0.0f;
  if (q.epi) {
    const float* gx = p.gx + ((long long)q.n * T + t) * 4 * E + q.u;
#pragma unroll
    for (int g = 0; g < 4; ++g) o.gxv[g] = ldcg1(gx + g * E);
    r.eps_t = ldcg1(p.eps + ((long long)t * N + q.n) * E + q.u);
  }
  if (t > 0) {
    quadrows<E>(q.n0, N, p.pz + ((long long)q.n0 * T + t - 1) * E, (long long)T * E, q.lane, o.a0);
    quadrows<E>(q.n0, N, p.h + ((long long)q.n0 * T + t - 1) * E, (long long)T * E, q.lane, o.a1);
  }
}
__device__ __forceinline__ void prior_fwd_lstm_compute(const PriorChainFwd& p, const float* Wl, PriorFwdRegs& r, int t, const QuadRole& q,
                                                       const PriorLstmOps& o) {
  constexpr int E = kChainE;
  const int T = p.T;
  float v[8][4];
  zero8(v);
  if (t > 0) {
    __syncwarp();
    quadfma<4, E>(o.a0, Wl, 2 * E, q.lane, v);
    quadfma<4, E>(o.a1, Wl + E, 2 * E, q.lane, v);
    reduce_scatter8<4>(v, q.lane);
  }
  if (q.epi) {
    const float ig = sigmoidf_(v[0][0] + o.gxv[0] + r.bh[0]);
    const float fg = sigmoidf_(v[0][1] + o.gxv[1] + r.bh[1]);
    const float gg = tanhf(v[0][2] + o.gxv[2] + r.bh[2]);
    const float og = sigmoidf_(v[0][3] + o.gxv[3] + r.bh[3]);
    const float cn = fg * r.c_prev + ig * gg;
    const float hn = og * tanhf(cn);
    r.c_prev = cn;
    float* gs = p.gates + ((long long)q.n * T + t) * 4 * E + q.u;
    gs[0] = ig; gs[E] = fg; gs[2 * E] = gg; gs[3 * E] = og;
    p.c[((long long)q.n * T + t) * E + q.u] = cn;
    p.h[((long long)q.n * T + t) * E + q.u] = hn;
  }
}
__device__ __forceinline__ void prior_fwd_lstm(const PriorChainFwd& p, const float* Wl, PriorFwdRegs& r, int t, const QuadRole& q) {
  PriorLstmOps o;
  prior_fwd_lstm_load(p, r, t, q, o);
  prior_fwd_lstm_compute(p, Wl, r, t, q, o);
}
// Gaussian head + reparameterisation of step t (needs h_t of all units: one barrier after the LSTM phase)
__device__ __forceinline__ void prior_fwd_head_load(const PriorChainFwd& p, int t, const QuadRole& q, float4 (&a)[4][kChainE / 128]) {
  constexpr int E = kChainE;
  quadrows<E>(q.n0, p.N, p.h + ((long long)q.n0 * p.T + t) * E, (long long)p.T * E, q.lane, a);
}
__device__ __forceinline__ void prior_fwd_head_compute(const PriorChainFwd& p, const float* Wh, PriorFwdRegs& r, int t, const QuadRole& q,
                                                       const float4 (&a)[4][kChainE / 128]) {
  constexpr int E = kChainE;
  const int T = p.T;
  float v[8][2];
  zero8(v);
  __syncwarp();
  quadfma<2, E>(a, Wh, E, q.lane, v);
  reduce_scatter8<2>(v, q.lane);
  if (q.epi) {
    const float mean = v[0][0] + r.hb_m;
    const float lg = v[0][1] + r.hb_l;
    const long long o = ((long long)q.n * T + t) * E + q.u;
    p.pm[o] = mean; p.pl[o] = lg; p.pz[o] = r.eps_t * expf(0.5f * lg) + mean;
  }
}
__device__ __forceinline__ void prior_fwd_head(const PriorChainFwd& p, const float* Wh, PriorFwdRegs& r, int t, const QuadRole& q) {
  float4 a[4][kChainE / 128];
  prior_fwd_head_load(p, t, q, a);
  prior_fwd_head_compute(p, Wh, r, t, q, a);
}

// two CTAs per SM (<= 128 registers): next to the cluster decoder chain (64 whole SMs) the 128 CTAs fit on the other 84 SMs
__global__ void __launch_bounds__(kChainThreads, 2) prior_chain_fwd_kernel(const __grid_constant__ PriorChainFwd p) {
  __shared__ __align__(16) float Wsm[kPriorFwdSmemFloats];
  float* Wl = Wsm; float* Wh = Wsm + 8 * 2 * kChainE;
  const int u0 = blockIdx.x * 2;
  const QuadRole q = quad_role(p.N, u0);
  PriorFwdRegs r;
  prior_fwd_setup(p, Wl, Wh, r, u0, q.u);
  __syncthreads();
  GridBar gb{p.bar, 0u, gridDim.x};
  const int t1 = p.t1 > 0 ? p.t1 : p.T;
  if (p.t0 > 0 && q.epi) r.c_prev = p.c[((long long)q.n * p.T + p.t0 - 1) * kChainE + q.u];    // resume (h, z are read from global)
  for (int t = p.t0; t < t1; ++t) {
    prior_fwd_lstm(p, Wl, r, t, q);
    grid_sync(gb);
    prior_fwd_head(p, Wh, r, t, q);
    if (t + 1 < t1) grid_sync(gb);
  }
}

// =====================================================================================================
// posterior biGRU backward (BPTT with the packed-sequence mask), both directions in one kernel
// =====================================================================================================
struct PostChainBwd {
  int N, T;
  const float* dho;      // [N,T,2E] upstream gradient of the GRU outputs
  const float* whh[2];   // [3E,E]
  const float* gq[2];    // [N,T,4E]
  const float* ho;       // [N,T,2E]
  const int* lens;
  float* dgi[2];         // [N,T,3E]
  float* dgh[2];         // [N,T,3E]
  unsigned* bar;
};
__global__ void __launch_bounds__(kChainThreads) post_chain_bwd_kernel(const __grid_constant__ PostChainBwd p) {
  constexpr int E = kChainE;
  __shared__ __align__(16) float Wt[2][2][3 * E];   // Wt[dir][j][c] = W_hh[c][u0+j]
  const int tid = threadIdx.x, lane = tid & 31, n0 = (tid >> 5) * 4;
  const int u0 = blockIdx.x * 2;
  for (int i = tid; i < 2 * 2 * 3 * E; i += kChainThreads) {
    const int dir = i / (2 * 3 * E), jj = (i / (3 * E)) & 1, c = i % (3 * E);
    Wt[dir][jj][c] = p.whh[dir][(long long)c * E + u0 + jj];
  }
  __syncthreads();
  GridBar gb{p.bar, 0u, gridDim.x};
  const int T = p.T, N = p.N;
  // epilogue role after reduce_scatter16: combo = lane >> 1 = (row r, direction, unit j)
  const int cq = lane >> 1, dir = (cq >> 1) & 1, j = cq & 1, u = u0 + j, n = n0 + (cq >> 2);
  const bool epi = (lane & 1) == 0 && n < N;
  const int len = epi ? p.lens[n] : 0;
  float carry = 0.0f;
  for (int s = T - 1; s >= 0; --s) {
    float v[16][1];                                    // v[r*4 + dir*2 + j]
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c][0] = 0.0f;
    const int t = dir ? T - 1 - s : s;
    const int tp = dir ? t + 1 : t - 1;
    float dho = 0.f, rr = 0.f, z = 0.f, nn = 0.f, ghn = 0.f, hp = 0.f;
    if (epi) {                                         // pointwise operands: issued with the row loads
      dho = ldcg1(p.dho + ((long long)n * T + t) * 2 * E + dir * E + u);
      const float* g = p.gq[dir] + ((long long)n * T + t) * 4 * E + u;
      rr = ldcg1(g); z = ldcg1(g + E); nn = ldcg1(g + 2 * E); ghn = ldcg1(g + 3 * E);
      if (s > 0) hp = ldcg1(p.ho + ((long long)n * T + tp) * 2 * E + dir * E + u);
    }
    if (s < T - 1) {
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int ts = d ? T - 2 - s : s + 1;          // the step this direction processed in the previous phase
        float4 a[4][3 * E / 128];
        quadrows<3 * E>(n0, N, p.dgh[d] + ((long long)n0 * T + ts) * 3 * E, (long long)T * 3 * E, lane, a);
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int i = 0; i < 3 * E / 128; ++i) {
            const float4 w = *(reinterpret_cast<const float4*>(&Wt[d][jj][0]) + i * 32 + lane);
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r * 4 + d * 2 + jj][0] = dot4(a[r][i], w, v[r * 4 + d * 2 + jj][0]);
          }
      }
      reduce_scatter16<1>(v, lane);
    }
    if (epi) {
      float* gi = p.dgi[dir] + ((long long)n * T + t) * 3 * E + u;
      float* gh = p.dgh[dir] + ((long long)n * T + t) * 3 * E + u;
      if (t >= len) {
        gi[0] = gi[E] = gi[2 * E] = 0.0f;
        gh[0] = gh[E] = gh[2 * E] = 0.0f;
        carry = 0.0f;
      } else {
        float dh = dho;
        if (s < T - 1) dh += carry + v[0][0];
        const float dn = dh * (1.0f - z), dz = dh * (hp - nn);
        const float dan = dn * (1.0f - nn * nn);
        const float dar = dan * ghn * rr * (1.0f - rr), daz = dz * z * (1.0f - z);
        gi[0] = dar; gi[E] = daz; gi[2 * E] = dan;
        gh[0] = dar; gh[E] = daz; gh[2 * E] = dan * rr;
        carry = dh * z;
      }
    }
    if (s > 0) grid_sync(gb);
  }
}

// =====================================================================================================
// prior backward: BPTT through the LSTM, the Gaussian head and the last_z chain (vae_model.py:869)
// =====================================================================================================
struct PriorChainBwd {
  int N, T;
  const float *d_pz, *d_pm, *d_pl;   // [N,T,E] upstream gradients or NULL
  const float* eps;                  // [T,N,E]
  const float* p_logs;               // [N,T,E]
  const float* head_w;               // [2E,E]
  const float* wih;                  // [4E,3E]
  const float* whh;                  // [4E,E]
  const float* gates;                // [N,T,4E]
  const float* c;                    // [N,T,E]
  float* dml;                        // [N,T,2E]
  float* dg;                         // [N,T,4E]
  unsigned* bar;
};
struct PriorBwdRegs { float dh_carry, dc_carry; };
constexpr int kPriorBwdSmemFloats = 2 * 2 * kChainE + 4 * 4 * kChainE;   // WA[2][2E] | WB[4][4E]

__device__ __forceinline__ void prior_bwd_setup(const PriorChainBwd& p, float* WA, float* WB, PriorBwdRegs& r, int u0) {
  constexpr int E = kChainE;
  for (int i = threadIdx.x; i < 2 * 2 * E; i += kChainThreads) {        // WA[j][c] = W_head[c][u0+j]
    const int jj = i / (2 * E), c = i % (2 * E);
    WA[i] = p.head_w[(long long)c * E + u0 + jj];
  }
  for (int i = threadIdx.x; i < 4 * 4 * E; i += kChainThreads) {        // (d last_z j0, j1, d h j0, j1)
    const int rr = i / (4 * E), c = i % (4 * E);
    WB[i] = rr < 2 ? p.wih[(long long)c * 3 * E + 2 * E + u0 + rr] : p.whh[(long long)c * E + u0 + (rr - 2)];
  }
  r.dh_carry = 0.0f; r.dc_carry = 0.0f;
}
// head backward of the LAST step (no chain input): must be followed by a barrier before prior_bwd_lstm(T-1)
__device__ __forceinline__ void prior_bwd_first(const PriorChainBwd& p, const QuadRole& q) {
  constexpr int E = kChainE;
  const int T = p.T, N = p.N, t = T - 1;
  if (q.epi) {
    const long long o = ((long long)q.n * T + t) * E + q.u;
    float dz = p.d_pz ? p.d_pz[o] : 0.0f;
    float dm = dz;
    float dl = dz * p.eps[((long long)t * N + q.n) * E + q.u] * 0.5f * expf(0.5f * p.p_logs[o]);
    if (p.d_pm) dm += p.d_pm[o];
    if (p.d_pl) dl += p.d_pl[o];
    float* d = p.dml + ((long long)q.n * T + t) * 2 * E + q.u;
    d[0] = dm; d[E] = dl;
  }
}
// dh_t = dML_t . W_head (+ carry); LSTM pointwise backward of step t
__device__ __forceinline__ void prior_bwd_lstm(const PriorChainBwd& p, const float* WA, PriorBwdRegs& r, int t, const QuadRole& q) {
  constexpr int E = kChainE;
  const int T = p.T;
  float v[8][1];
  zero8(v);
  float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cc = 0.f, cp = 0.f;
  if (q.epi) {
    const float* g = p.gates + ((long long)q.n * T + t) * 4 * E + q.u;
    ig = ldcg1(g); fg = ldcg1(g + E); gg = ldcg1(g + 2 * E); og = ldcg1(g + 3 * E);
    cc = ldcg1(p.c + ((long long)q.n * T + t) * E + q.u);
    if (t > 0) cp = ldcg1(p.c + ((long long)q.n * T + t - 1) * E + q.u);
  }
  quaddot<1, 2 * E>(q.n0, p.N, p.dml + ((long long)q.n0 * T + t) * 2 * E, (long long)T * 2 * E, WA, 2 * E, q.lane, v);
  reduce_scatter8<1>(v, q.lane);
  if (q.epi) {
    const float dh = v[0][0] + r.dh_carry;
    const float tc = tanhf(cc);
    const float dc = dh * og * (1.0f - tc * tc) + r.dc_carry;
    float* dg = p.dg + ((long long)q.n * T + t) * 4 * E + q.u;
    dg[0] = dc * gg * ig * (1.0f - ig);
    dg[E] = dc * cp * fg * (1.0f - fg);
    dg[2 * E] = dc * ig * (1.0f - gg * gg);
    dg[3 * E] = dh * tc * og * (1.0f - og);
    r.dc_carry = dc * fg;
  }
}
// [d last_z | d h_{t-1}] = dG_t . [W_ih[:, 2E:3E] | W_hh]; head backward of step t-1   (t > 0)
__device__ __forceinline__ void prior_bwd_head(const PriorChainBwd& p, const float* WB, PriorBwdRegs& r, int t, const QuadRole& q) {
  constexpr int E = kChainE;
  const int T = p.T, N = p.N;
  float v[8][2];                                       // (d last_z, d h) per (row, unit)
  zero8(v);
  float u_pz = 0.f, u_pm = 0.f, u_pl = 0.f, e_ = 0.f, lv = 0.f;
  if (q.epi) {
    const long long o = ((long long)q.n * T + t - 1) * E + q.u;
    if (p.d_pz) u_pz = ldcg1(p.d_pz + o);
    if (p.d_pm) u_pm = ldcg1(p.d_pm + o);
    if (p.d_pl) u_pl = ldcg1(p.d_pl + o);
    e_ = ldcg1(p.eps + ((long long)(t - 1) * N + q.n) * E + q.u);
    lv = ldcg1(p.p_logs + o);
  }
  // two halves of K = 4E: 16 instead of 32 float4 of operands in flight
  quaddot<2, 2 * E>(q.n0, N, p.dg + ((long long)q.n0 * T + t) * 4 * E, (long long)T * 4 * E, WB, 4 * E, q.lane, v);
  quaddot<2, 2 * E>(q.n0, N, p.dg + ((long long)q.n0 * T + t) * 4 * E + 2 * E, (long long)T * 4 * E, WB + 2 * E, 4 * E, q.lane, v);
  reduce_scatter8<2>(v, q.lane);
  if (q.epi) {
    r.dh_carry = v[0][1];
    const float dz = v[0][0] + u_pz;
    float* d = p.dml + ((long long)q.n * T + t - 1) * 2 * E + q.u;
    d[0] = dz + u_pm;
    d[E] = dz * e_ * 0.5f * expf(0.5f * lv) + u_pl;
  }
}

__global__ void __launch_bounds__(kChainThreads, 2) prior_chain_bwd_kernel(const __grid_constant__ PriorChainBwd p) {
  __shared__ __align__(16) float Wsm[kPriorBwdSmemFloats];
  float* WA = Wsm; float* WB = Wsm + 2 * 2 * kChainE;
  const int u0 = blockIdx.x * 2;
  const QuadRole q = quad_role(p.N, u0);
  PriorBwdRegs r;
  prior_bwd_setup(p, WA, WB, r, u0);
  __syncthreads();
  GridBar gb{p.bar, 0u, gridDim.x};
  prior_bwd_first(p, q);
  grid_sync(gb);
  for (int t = p.T - 1; t >= 0; --t) {
    prior_bwd_lstm(p, WA, r, t, q);
    if (t == 0) break;
    grid_sync(gb);
    prior_bwd_head(p, WB, r, t, q);
    grid_sync(gb);
  }
}

// =====================================================================================================
// decoder forward chain: query projection -> additive attention over the clip's frames -> GRU
// (decoder.py:183-199, attn_model.py:20-46); the [emb | z] half of the gate pre-activations is hoisted
// =====================================================================================================
struct DecChainFwd {
  int N, T, Te;
  const float* gx;        // [N,T,3E] = [emb | z] . W_ih[:, {0:E, 2E:3E}]^T + b_ih
  const float* attn_w;    // [A,2E] (query columns first)
  const float* attn_v;    // [A]
  const float* wih;       // [3E,3E]
  const float* whh;       // [3E,E]
  const float* bhh;       // [3E]
  const float *Pd, *mem;  // [N,Te,A], [N,Te,E]
  const int* mem_lens;
  float *qp, *w, *ctx, *gates, *out;   // [N,T,A], [N,T,Te], [N,T,E], [N,T,4E], [N,T,E]
  float* aw;              // [N,Te,T] user-visible attention weights or NULL
  float* part;            // [N, kChainCtas, A] per-CTA partial query projections of the step in flight
  unsigned* bar;
  long long* trace;       // optional [T][16] clock64 stamps of thread 0 of CTA 0 (profiles/chain_trace.py) or NULL
};
// process-wide trace destination picked up by train_fast.cuh (profiling only)
inline long long*& chain_trace_ptr() { static long long* p = nullptr; return p; }
inline size_t dec_chain_fwd_smem(int Te) { return ((size_t)Te * 2 * kChainE + 2 * kChainE + 12 * kChainE + 2 * kChainE + Te + 64 + 4 * kChainE) * sizeof(float); }

// Two phases (grid barriers) per step:
//   A (clip CTAs): query projection of the clip = sum of the 128 per-CTA partials written by phase G of the previous
//     step, additive attention, context;                                  [+ prior LSTM cell of the step, all CTAs]
//   G (all CTAs): GRU cell of units {u0, u0+1} for all rows, then THIS CTA's partial of the next step's query
//     projection, q.Wq^T restricted to k in {u0, u0+1} for all A outputs  [+ prior Gaussian head of the step].
// The partial-sum form removes the separate query-projection phase (one barrier + one state broadcast per step):
// a unit-partitioned product needs every unit of h_t, the partials need only the two this CTA has just computed,
// and the clip's CTA adds them in a fixed order (deterministic).
// `pp.N > 0`: the (independent) prior chain of the same step count runs inside the same phases.
__global__ void __launch_bounds__(kChainThreads) dec_chain_fwd_kernel(const __grid_constant__ DecChainFwd p,
                                                                      const __grid_constant__ PriorChainFwd pp) {
  constexpr int E = kChainE, A = kChainE;
  extern __shared__ __align__(16) float dsm[];
  __shared__ __align__(16) float Wprior[kPriorFwdSmemFloats];
  const bool prior = pp.N > 0;
  const int Te = p.Te, T = p.T, N = p.N;
  float* Ps = dsm;                        // [Te][A]   clip blockIdx.x
  float* Ms = Ps + (size_t)Te * A;        // [Te][E]
  float* hs = Ms + (size_t)Te * E;        // [N][2] h_t of this CTA's two units (+ padding up to 2E floats)
  float* Wg = hs + 2 * E;                 // [12][E]: rows 0..5 (r,z,n)x2 on ctx, rows 6..11 on h
  float* qps = Wg + 12 * E;               // [A]
  float* vs = qps + A;                    // [A]
  float* psum = vs + A;                   // [4][A] partial query projections, one row per quarter of the CTAs (16-byte aligned)
  float* sc = psum + 4 * A;               // [Te]
  float* red = sc + Te;                   // [64]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int u0 = blockIdx.x * 2;
  const QuadRole q = quad_role(N, u0);
  const int clip = blockIdx.x;
  const bool own_clip = clip < N;
  const int len = own_clip ? max(1, min(p.mem_lens[clip], Te)) : 0;
  if (own_clip) {
    // P_d is kept pre-scaled by 2 log2(e): the score loop is  t = ex2(P' + q') ; r = rcp(t + 1) ; s += (-2 v) r
    // (tanh(x) = 1 - 2 / (e^{2x} + 1), so  sum_a v_a tanh(x_a) = sum_a v_a + sum_a (-2 v_a) r_a)
    const float4* sp = reinterpret_cast<const float4*>(p.Pd + (long long)clip * Te * A);
    const float4* sm_ = reinterpret_cast<const float4*>(p.mem + (long long)clip * Te * E);
    for (int i = tid; i < len * A / 4; i += kChainThreads) {
      float4 v = sp[i];
      v.x *= kTwoLog2e; v.y *= kTwoLog2e; v.z *= kTwoLog2e; v.w *= kTwoLog2e;
      reinterpret_cast<float4*>(Ps)[i] = v;
    }
    for (int i = tid; i < len * E / 4; i += kChainThreads) reinterpret_cast<float4*>(Ms)[i] = sm_[i];
  }
  vs[tid] = p.attn_v[tid];                 // blockDim == A
  float v2[A / 32], vsum = 0.0f;           // this lane's slice of -2 v, and sum_a v_a
#pragma unroll
  for (int i = 0; i < A / 32; ++i) { const float x = p.attn_v[lane + 32 * i]; v2[i] = -2.0f * x; vsum += x; }
  vsum = warp_sum(vsum);
  // query-projection weights of output a = tid for this CTA's two input units (attn_model.py:31: query columns first)
  const float wq0 = p.attn_w[(long long)tid * 2 * E + u0], wq1 = p.attn_w[(long long)tid * 2 * E + u0 + 1];
  for (int i = tid; i < 12 * E; i += kChainThreads) {
    const int r = i / E, k = i % E, rr = r % 6;
    const long long wr = (long long)((rr >> 1) * E + u0 + (rr & 1));
    Wg[i] = r < 6 ? p.wih[wr * 3 * E + E + k] : p.whh[wr * E + k];
  }
  const int n = q.n, u = q.u;
  PriorFwdRegs pr;
  if (prior) prior_fwd_setup(pp, Wprior, Wprior + 8 * 2 * E, pr, u0, u);
  __syncthreads();
  GridBar gb{p.bar, 0u, gridDim.x};
  const float bh_r = p.bhh[u], bh_z = p.bhh[E + u], bh_n = p.bhh[2 * E + u];
  for (int t = 0; t < T; ++t) {
    // the prior LSTM's operand rows are requested first: their L2 round trip overlaps the attention below
    const bool tr = p.trace && blockIdx.x == 0 && tid == 0;
    if (tr) p.trace[t * 16 + 0] = clock64();
    // ---- A: attention of clip `clip` (query projection = sum of the partials; zero query at t = 0, decoder.py:94-98) ----
    // query projection of the clip: 128 partial rows of A floats, summed in a fixed order (deterministic): thread
    // (quarter bq, columns 4*aq..4*aq+3) adds 32 rows with 16-byte loads -- ALL 32 requested before the first add, one
    // L2 round trip (8 in flight cost four: 4500 cycles, profiles/chain_trace.py) -- the four quarters meet in shared memory
    if (own_clip && t > 0) {
      const int aq = tid & 63, bq = tid >> 6;
      const float4* pq = reinterpret_cast<const float4*>(p.part + ((long long)clip * kChainCtas + bq * 32) * A) + aq;
      float4 x[32];
#pragma unroll
      for (int b = 0; b < 32; ++b) x[b] = ldcg4(reinterpret_cast<const float*>(pq + (long long)b * (A / 4)));
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
      for (int b = 0; b < 32; b += 2) {
        s0.x += x[b].x; s0.y += x[b].y; s0.z += x[b].z; s0.w += x[b].w;
        s1.x += x[b + 1].x; s1.y += x[b + 1].y; s1.z += x[b + 1].z; s1.w += x[b + 1].w;
      }
      reinterpret_cast<float4*>(psum + bq * A)[aq] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
    }
    // the prior LSTM's operand rows are requested next: their L2 round trip overlaps the attention below
    PriorLstmOps lops;
    if (prior) prior_fwd_lstm_load(pp, pr, t, q, lops);
    if (own_clip) {
      if (t > 0) __syncthreads();
      const float qv = t > 0 ? (psum[tid] + psum[A + tid]) + (psum[2 * A + tid] + psum[3 * A + tid]) : 0.0f;
      qps[tid] = qv;
      p.qp[((long long)clip * T + t) * A + tid] = qv;          // saved for the backward
      __syncthreads();
      if (tr) p.trace[t * 16 + 1] = clock64();
      {
        float qreg[A / 32];
#pragma unroll
        for (int i = 0; i < A / 32; ++i) qreg[i] = kTwoLog2e * qps[lane + 32 * i];
        for (int jj = wid; jj < len; jj += 2 * (kChainThreads / 32)) {     // two frames per pass: 16 independent chains
          const int j2 = jj + kChainThreads / 32;
          const float* pr0 = Ps + (size_t)jj * A + lane;
          const float* pr1 = Ps + (size_t)(j2 < len ? j2 : jj) * A + lane;
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int i = 0; i < A / 32; ++i) {
            s0 = fmaf(v2[i], rcp_approx(ex2_approx(pr0[32 * i] + qreg[i]) + 1.0f), s0);
            s1 = fmaf(v2[i], rcp_approx(ex2_approx(pr1[32 * i] + qreg[i]) + 1.0f), s1);
          }
          s0 = warp_sum(s0); s1 = warp_sum(s1);
          if (lane == 0) { sc[jj] = s0 + vsum; if (j2 < len) sc[j2] = s1 + vsum; }
        }
      }
      __syncthreads();
      if (tr) p.trace[t * 16 + 9] = clock64();
      if (wid == 0) {                                        // masked softmax over <= Te frames by one warp (no block reductions)
        float mx = -INFINITY;
        for (int jj = lane; jj < len; jj += 32) mx = fmaxf(mx, sc[jj]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int jj = lane; jj < len; jj += 32) { const float e = expf(sc[jj] - mx); sc[jj] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int jj = lane; jj < len; jj += 32) sc[jj] *= inv;
      }
      __syncthreads();
      if (tr) p.trace[t * 16 + 10] = clock64();
      float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
      int jj = 0;
      for (; jj + 4 <= len; jj += 4) {
        c0 = fmaf(sc[jj], Ms[(size_t)jj * E + tid], c0);
        c1 = fmaf(sc[jj + 1], Ms[(size_t)(jj + 1) * E + tid], c1);
        c2 = fmaf(sc[jj + 2], Ms[(size_t)(jj + 2) * E + tid], c2);
        c3 = fmaf(sc[jj + 3], Ms[(size_t)(jj + 3) * E + tid], c3);
      }
      for (; jj < len; ++jj) c0 = fmaf(sc[jj], Ms[(size_t)jj * E + tid], c0);
      p.ctx[((long long)clip * T + t) * E + tid] = (c0 + c1) + (c2 + c3);
      // the saved / user-visible attention weights leave the chip off the critical path (sc is read-only until the next step)
      for (int j2 = tid; j2 < Te; j2 += kChainThreads) {
        const float wv = j2 < len ? sc[j2] : 0.0f;
        p.w[((long long)clip * T + t) * Te + j2] = wv;
        if (p.aw) p.aw[((long long)clip * Te + j2) * T + t] = wv;
      }
    }
    if (tr) p.trace[t * 16 + 2] = clock64();
    if (prior) prior_fwd_lstm_compute(pp, Wprior, pr, t, q, lops);
    if (tr) p.trace[t * 16 + 3] = clock64();
    grid_sync(gb);
    if (tr) p.trace[t * 16 + 4] = clock64();
    // ---- G: GRU cell, units {u0, u0+1} ----
    float vx[8][3], vh[8][3];                        // (r, z, n) pre-activations from ctx_t and from h_{t-1}
    zero8(vx); zero8(vh);
    float gxr = 0.f, gxz = 0.f, gxn = 0.f, hp = 0.f;
    if (q.epi) {
      const float* gx = p.gx + ((long long)n * T + t) * 3 * E + u;
      gxr = ldcg1(gx); gxz = ldcg1(gx + E); gxn = ldcg1(gx + 2 * E);
      if (t > 0) hp = ldcg1(p.out + ((long long)n * T + t - 1) * E + u);
    }
    float4 hrow[4][E / 128];                          // the prior head's operand rows, requested with the GRU's
    {
      float4 a0[4][E / 128], a1[4][E / 128];
      quadrows<E>(q.n0, N, p.ctx + ((long long)q.n0 * T + t) * E, (long long)T * E, q.lane, a0);
      if (t > 0) quadrows<E>(q.n0, N, p.out + ((long long)q.n0 * T + t - 1) * E, (long long)T * E, q.lane, a1);
      if (prior) prior_fwd_head_load(pp, t, q, hrow);
      __syncwarp();
      quadfma<3, E>(a0, Wg, E, q.lane, vx);
      if (t > 0) quadfma<3, E>(a1, Wg + 6 * E, E, q.lane, vh);
    }
    reduce_scatter8<3>(vx, q.lane);
    if (t > 0) reduce_scatter8<3>(vh, q.lane);
    if (tr) p.trace[t * 16 + 5] = clock64();
    if (q.epi) {
      const float hn = vh[0][2] + bh_n;
      const float rg = sigmoidf_(vx[0][0] + gxr + vh[0][0] + bh_r);
      const float zg = sigmoidf_(vx[0][1] + gxz + vh[0][1] + bh_z);
      const float ng = tanhf(vx[0][2] + gxn + rg * hn);
      float* gs = p.gates + ((long long)n * T + t) * 4 * E + u;
      gs[0] = rg; gs[E] = zg; gs[2 * E] = ng; gs[3 * E] = hn;
      const float hnew = (1.0f - zg) * ng + zg * hp;
      p.out[((long long)n * T + t) * E + u] = hnew;
      hs[n * 2 + q.j] = hnew;
    }
    if (prior) prior_fwd_head_compute(pp, Wprior + 8 * 2 * E, pr, t, q, hrow);
    if (tr) p.trace[t * 16 + 6] = clock64();
    if (t + 1 < T) {
      // this CTA's partial of the next step's query projection: part[n][b][a] = Wq[a][u0] h[n][u0] + Wq[a][u0+1] h[n][u0+1]
      __syncthreads();
      float* dst = p.part + (long long)blockIdx.x * A + tid;
      for (int r = 0; r < N; ++r)
        dst[(long long)r * kChainCtas * A] = fmaf(wq1, hs[r * 2 + 1], wq0 * hs[r * 2]);
      if (tr) p.trace[t * 16 + 7] = clock64();
      grid_sync(gb);
      if (tr) p.trace[t * 16 + 8] = clock64();
    }
  }
}

// =====================================================================================================
// decoder backward chain: GRU pointwise backward -> d ctx -> attention backward (d score, d query proj)
// =====================================================================================================
struct DecChainBwd {
  int N, T, Te;
  const float* dout;      // [N,T,E] upstream gradient of the GRU outputs (incl. the pooled global head)
  const float* attn_w;    // [A,2E]
  const float* attn_v;    // [A]
  const float* wih;       // [3E,3E]
  const float* whh;       // [3E,E]
  const float *Pd, *mem;
  const int* mem_lens;
  const float *qp, *w, *gates, *out;   // saved by the forward
  float *dgi, *dgh;       // [N,T,3E]
  float *dctx, *ds, *dqp; // [N,T,E], [N,T,Te], [N,T,A]
  unsigned* bar;
  long long* trace;       // optional [T][16] clock64 stamps of thread 0 of CTA 0 or NULL (profiles/chain_trace.py)
};
inline size_t dec_chain_bwd_smem(int Te) { return ((size_t)Te * 2 * kChainE + 2 * 4 * kChainE + 2 * 3 * kChainE + 2 * kChainE + Te + 64) * sizeof(float); }

// `pp.N > 0`: the prior backward chain runs inside the same phases (LSTM backward next to the GRU backward, head
// backward next to the d ctx projection).
// (With the row-per-8-lanes mapping MERGE = true spilled ~800 bytes per thread; with the quad mapping it does not.
// ACVAE_MERGE_BWD=0 selects MERGE = false plus prior_chain_bwd_kernel.)
template <bool MERGE>
__global__ void __launch_bounds__(kChainThreads) dec_chain_bwd_kernel(const __grid_constant__ DecChainBwd p,
                                                                      const __grid_constant__ PriorChainBwd pp) {
  constexpr int E = kChainE, A = kChainE;
  extern __shared__ __align__(16) float dsm[];
  __shared__ __align__(16) float Wprior[MERGE ? kPriorBwdSmemFloats : 4];
  const bool prior = MERGE && pp.N > 0;
  const int Te = p.Te, T = p.T, N = p.N;
  float* Ps = dsm;                        // [Te][A]
  float* Ms = Ps + (size_t)Te * A;        // [Te][E]
  float* W1 = Ms + (size_t)Te * E;        // [2][4E]: W1[j][c] = c < 3E ? W_hh[c][u0+j] : Wq[c-3E][u0+j]
  float* W2 = W1 + 2 * 4 * E;             // [2][3E]: W2[j][c] = W_ih[c][E + u0+j]
  float* qps = W2 + 2 * 3 * E;            // [A]
  float* dcs = qps + A;                   // [E]
  float* dw = dcs + E;                    // [Te]
  float* red = dw + Te;                   // [64]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int u0 = blockIdx.x * 2;
  const QuadRole q = quad_role(N, u0);
  const int clip = blockIdx.x;
  const bool own_clip = clip < N;
  const int len = own_clip ? max(1, min(p.mem_lens[clip], Te)) : 0;
  if (own_clip) {
    const float4* sp = reinterpret_cast<const float4*>(p.Pd + (long long)clip * Te * A);
    const float4* sm_ = reinterpret_cast<const float4*>(p.mem + (long long)clip * Te * E);
    // P_d pre-scaled by 2 log2(e) (see the forward kernel): r = 1 / (e^{2x} + 1) = rcp(ex2(P' + q') + 1), 1 - tanh^2 = 4 r (1 - r)
    for (int i = tid; i < len * A / 4; i += kChainThreads) {
      float4 v = sp[i];
      v.x *= kTwoLog2e; v.y *= kTwoLog2e; v.z *= kTwoLog2e; v.w *= kTwoLog2e;
      reinterpret_cast<float4*>(Ps)[i] = v;
    }
    for (int i = tid; i < len * E / 4; i += kChainThreads) reinterpret_cast<float4*>(Ms)[i] = sm_[i];
  }
  for (int i = tid; i < 2 * 4 * E; i += kChainThreads) {
    const int jj = i / (4 * E), c = i % (4 * E);
    W1[i] = c < 3 * E ? p.whh[(long long)c * E + u0 + jj] : p.attn_w[(long long)(c - 3 * E) * 2 * E + u0 + jj];
  }
  for (int i = tid; i < 2 * 3 * E; i += kChainThreads) {
    const int jj = i / (3 * E), c = i % (3 * E);
    W2[i] = p.wih[(long long)c * 3 * E + E + u0 + jj];
  }
  const int n = q.n, u = q.u;
  PriorBwdRegs pr;
  if (prior) prior_bwd_setup(pp, Wprior, Wprior + 2 * 2 * E, pr, u0);
  __syncthreads();
  GridBar gb{p.bar, 0u, gridDim.x};
  if (prior) {
    prior_bwd_first(pp, q);
    grid_sync(gb);
  }
  const float va = p.attn_v[tid];
  float carry = 0.0f;
  for (int t = T - 1; t >= 0; --t) {
    const bool tr = p.trace && blockIdx.x == 0 && tid == 0;
    if (tr) p.trace[t * 16 + 0] = clock64();
    // ---- B1: dh_t (units u0, u0+1) and the GRU pointwise backward of step t ----
    // (requesting all 6E floats of row operands of this phase at once was tried: the backward phases are L2-bandwidth
    // bound -- every CTA reads every row, 130-230 KB per CTA and phase, 25 MB per phase over all CTAs -- and the chunked
    // form below overlaps one chunk's FMAs with the next chunk's loads: 5500 vs 7000 cycles, profiles/chain_trace.py)
    float va2[8][1];
    zero8(va2);
    float dh = 0.f, rr = 0.f, z = 0.f, nn = 0.f, ghn = 0.f, hp = 0.f;
    if (q.epi) {
      dh = ldcg1(p.dout + ((long long)n * T + t) * E + u);
      const float* g = p.gates + ((long long)n * T + t) * 4 * E + u;
      rr = ldcg1(g); z = ldcg1(g + E); nn = ldcg1(g + 2 * E); ghn = ldcg1(g + 3 * E);
      if (t > 0) hp = ldcg1(p.out + ((long long)n * T + t - 1) * E + u);
    }
    if (t < T - 1) {
      quaddot<1, 2 * E>(q.n0, N, p.dgh + ((long long)q.n0 * T + t + 1) * 3 * E, (long long)T * 3 * E, W1, 4 * E, q.lane, va2);
      {
        float4 a0[4][E / 128], a1[4][A / 128];
        quadrows<E>(q.n0, N, p.dgh + ((long long)q.n0 * T + t + 1) * 3 * E + 2 * E, (long long)T * 3 * E, q.lane, a0);
        quadrows<A>(q.n0, N, p.dqp + ((long long)q.n0 * T + t + 1) * A, (long long)T * A, q.lane, a1);
        __syncwarp();
        quadfma<1, E>(a0, W1 + 2 * E, 4 * E, q.lane, va2);
        quadfma<1, A>(a1, W1 + 3 * E, 4 * E, q.lane, va2);
      }
      reduce_scatter8<1>(va2, q.lane);
    }
    if (q.epi) {
      if (t < T - 1) dh += carry + va2[0][0];
      const float dn = dh * (1.0f - z), dz = dh * (hp - nn);
      const float dan = dn * (1.0f - nn * nn);
      const float dar = dan * ghn * rr * (1.0f - rr), daz = dz * z * (1.0f - z);
      float* gi = p.dgi + ((long long)n * T + t) * 3 * E + u;
      float* gh = p.dgh + ((long long)n * T + t) * 3 * E + u;
      gi[0] = dar; gi[E] = daz; gi[2 * E] = dan;
      gh[0] = dar; gh[E] = daz; gh[2 * E] = dan * rr;
      carry = dh * z;
    }
    if (tr) p.trace[t * 16 + 1] = clock64();
    if (prior) prior_bwd_lstm(pp, Wprior, pr, t, q);
    if (tr) p.trace[t * 16 + 2] = clock64();
    grid_sync(gb);
    if (tr) p.trace[t * 16 + 3] = clock64();
    // ---- B1.5: d ctx_t = dGi_t . W_ih[:, E:2E], columns {u0, u0+1} ----
    float vb2[8][1];
    zero8(vb2);
    quaddot<1, 2 * E>(q.n0, N, p.dgi + ((long long)q.n0 * T + t) * 3 * E, (long long)T * 3 * E, W2, 3 * E, q.lane, vb2);
    quaddot<1, E>(q.n0, N, p.dgi + ((long long)q.n0 * T + t) * 3 * E + 2 * E, (long long)T * 3 * E, W2 + 2 * E, 3 * E, q.lane, vb2);
    reduce_scatter8<1>(vb2, q.lane);
    if (q.epi) p.dctx[((long long)n * T + t) * E + u] = vb2[0][0];
    if (tr) p.trace[t * 16 + 4] = clock64();
    if (prior && t > 0) prior_bwd_head(pp, Wprior + 2 * 2 * E, pr, t, q);
    if (tr) p.trace[t * 16 + 5] = clock64();
    grid_sync(gb);
    if (tr) p.trace[t * 16 + 6] = clock64();
    // ---- B2: attention backward of clip `clip` ----
    if (own_clip) {
      // all global operands of the phase in one round trip: d ctx, the saved query projection, the saved weights
      const float dcv = ldcg1(p.dctx + ((long long)clip * T + t) * E + tid);
      const float qa = kTwoLog2e * p.qp[((long long)clip * T + t) * A + tid];
      const float wreg = tid < len ? p.w[((long long)clip * T + t) * Te + tid] : 0.0f;      // Te <= 83 < blockDim
      dcs[tid] = dcv;
      if (tid < len) qps[tid] = wreg;          // the weights, staged for the dot product below (qps is free: q lives in qa)
      __syncthreads();
      {
        float dreg[E / 32];
#pragma unroll
        for (int i = 0; i < E / 32; ++i) dreg[i] = dcs[lane + 32 * i];
        for (int jj = wid; jj < len; jj += 2 * (kChainThreads / 32)) {       // d w_j = d ctx . mem_j, two frames per pass
          const int j2 = jj + kChainThreads / 32;
          const float* m0 = Ms + (size_t)jj * E + lane;
          const float* m1 = Ms + (size_t)(j2 < len ? j2 : jj) * E + lane;
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int i = 0; i < E / 32; ++i) { s0 = fmaf(dreg[i], m0[32 * i], s0); s1 = fmaf(dreg[i], m1[32 * i], s1); }
          s0 = warp_sum(s0); s1 = warp_sum(s1);
          if (lane == 0) { dw[jj] = s0; if (j2 < len) dw[j2] = s1; }
        }
      }
      __syncthreads();
      // softmax backward d s_j = w_j (d w_j - sum_k w_k d w_k): the dot product by every warp for itself (<= 3 terms a lane)
      float dot = 0.0f;
      for (int jj = lane; jj < len; jj += 32) dot = fmaf(qps[jj], dw[jj], dot);
      dot = warp_sum(dot);
      __syncthreads();
      float dsv = 0.0f;
      if (tid < len) { dsv = wreg * (dw[tid] - dot); dw[tid] = 4.0f * dsv; }          // dw := 4 d s (the factor of 4 r (1 - r))
      if (tid < Te) p.ds[((long long)clip * T + t) * Te + tid] = dsv;
      __syncthreads();
      // d qp_a = v_a sum_j d s_j (1 - tanh^2(P_ja + q_a)), four independent chains
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* pc = Ps + tid;
      int jj = 0;
      for (; jj + 4 <= len; jj += 4) {
        const float r0 = rcp_approx(ex2_approx(pc[(size_t)jj * A] + qa) + 1.0f);
        const float r1 = rcp_approx(ex2_approx(pc[(size_t)(jj + 1) * A] + qa) + 1.0f);
        const float r2 = rcp_approx(ex2_approx(pc[(size_t)(jj + 2) * A] + qa) + 1.0f);
        const float r3 = rcp_approx(ex2_approx(pc[(size_t)(jj + 3) * A] + qa) + 1.0f);
        s0 = fmaf(dw[jj], fmaf(-r0, r0, r0), s0);
        s1 = fmaf(dw[jj + 1], fmaf(-r1, r1, r1), s1);
        s2 = fmaf(dw[jj + 2], fmaf(-r2, r2, r2), s2);
        s3 = fmaf(dw[jj + 3], fmaf(-r3, r3, r3), s3);
      }
      for (; jj < len; ++jj) {
        const float r0 = rcp_approx(ex2_approx(pc[(size_t)jj * A] + qa) + 1.0f);
        s0 = fmaf(dw[jj], fmaf(-r0, r0, r0), s0);
      }
      p.dqp[((long long)clip * T + t) * A + tid] = va * ((s0 + s1) + (s2 + s3));
    }
    if (tr) p.trace[t * 16 + 7] = clock64();
    if (t > 0) grid_sync(gb);
    if (tr) p.trace[t * 16 + 8] = clock64();
  }
}

// Holds a stream for `ns` nanoseconds (one thread): used to give the decoder's cluster chain a head start over the
// prior's cooperative chain, see train_fast.cuh.
__global__ void stream_delay_kernel(unsigned ns) {
  const unsigned long long t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < ns) __nanosleep(200);
}

// ---- host side -------------------------------------------------------------------------------------
template <typename Kern, typename... P>
inline int launch_chain(Kern kern, size_t smem, cudaStream_t st, const char* name, const P&... p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kChainCtas);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  // the chains are the critical path of the step: their CTAs are placed before those of the batched GEMMs that
  // are queued on the side streams at the same time (a cooperative grid needs all its SMs at once)
  static int prio = 1;
  if (prio > 0) { int lo = 0, hi = 0; prio = (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess) ? hi : 0; }
  at[1].id = cudaLaunchAttributePriority;
  at[1].val.priority = prio;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  const bool probe = probe_match(name);
  if (probe) cudaEventRecord(g_probe.e0, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p...);
  if (probe) { cudaEventRecord(g_probe.e1, st); ++g_probe.hits; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return set_error(name, cudaGetErrorString(e));
  return 0;
}

// Can the persistent chains run this problem?  (E == A == 256, batch <= 32 rows, the clip's projected memory
// and memory fit in shared memory, and all kChainCtas CTAs are co-resident.)
inline bool chain_supported(int N, int T, int Te, int E, int A) {
  if (E != kChainE || A != kChainE || N > kChainMaxN || N > kChainCtas || T < 1) return false;
  static int ok_dev[kMaxDevices], max_te_dev[kMaxDevices];
  static bool probed[kMaxDevices] = {false};
  const int cur = current_device();
  int& ok = ok_dev[cur];
  int& max_te = max_te_dev[cur];
  if (!probed[cur]) { probed[cur] = true; ok = -1; max_te = 0; }
  if (ok < 0) {
    ok = 0;
    const char* env = getenv("ACVAE_DISABLE_CHAIN");
    int dev = 0, sms = 0, coop = 0, optin = 0;
    if (!(env && env[0] == '1') && cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && coop && sms >= kChainCtas) {
      // largest Te whose resident clip fits next to the weight slices (the merged prior phases keep 20 KB of
      // weights in static shared memory)
      const int dyn_max = optin - 24 * 1024;
      max_te = (int)(((size_t)dyn_max / sizeof(float) - (2 * 4 + 2 * 3 + 2 + 12 + 2 + 4) * kChainE - 256) / (2 * kChainE + 1));   // ~83 on B200
      if (cudaFuncSetAttribute(dec_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) == cudaSuccess &&
          cudaFuncSetAttribute(dec_chain_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) == cudaSuccess &&
          cudaFuncSetAttribute(dec_chain_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) == cudaSuccess)
        ok = 1;
    }
  }
  return ok == 1 && Te <= max_te;
}

}  // namespace acvae
