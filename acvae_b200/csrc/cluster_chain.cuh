// Recurrent chains of the training step on thread-block CLUSTERS: state exchange through distributed shared memory.
//
// recurrent.cuh runs a chain as one cooperative grid of 128 CTAs that exchange the recurrent state through L2: every
// phase pays a grid barrier (2400 cycles = three dependent L2 trips) plus a re-read of state lines other SMs have just
// written (1200-2400 cycles), and 0.69 ms of the 1.2 ms step was that exchange (profiles/r1/chain_trace.log).
// Batch rows are independent, so here the ROWS are partitioned over clusters and only the hidden units over the CTAs of
// a cluster:
//   * a cluster of 8 CTAs owns R batch rows (R = 8 posterior, 4 decoder); CTA `rank` owns 32 hidden units (all gates);
//   * its weight slice lives in REGISTERS for the whole chain (96-128 floats per thread) -- shared memory is left for the
//     clip-resident attention operands;
//   * the state travels CTA to CTA with st.async (remote shared-memory store that completes on the DESTINATION's
//     mbarrier): the consumer waits on a local mbarrier, no cluster-wide barrier, no L2 round trip.  Measured
//     (profiles/ubench_cluster.cu, profiles/r2/ubench_cluster.log): all-gather of 512 B per CTA over 8 CTAs = 750-800
//     cycles per exchange (barrier.cluster alone: 630; the same all-gather with st.shared::cluster + barrier.cluster:
//     1100-1350; cp.async.bulk per destination: 1050; a grid barrier + re-read through L2: 3600-4800);
//   * clusters never wait for each other, so the launch is NOT cooperative: chains of different networks overlap
//     on disjoint SMs and any number of row groups is legal (they queue).  Co-residency on B200 with one CTA per SM:
//     15 clusters of 8, 7 clusters of 16 (one GPC has fewer than 16 SMs) -- hence clusters of 8.
// Exchange protocol (every hop is all-to-all inside the cluster): receive buffers and mbarriers are double-buffered by
// step parity.  A CTA can run at most one hop ahead of its slowest peer (it needs that peer's data of hop i before it can
// send hop i+1), so data of use k+1 of a (buffer, barrier) pair -- two steps later -- can never arrive before use k has
// been consumed.  The consumer arms its barrier with mbarrier.arrive.expect_tx; remote complete_tx that arrive before
// the arming only drive the transaction count negative (the phase cannot complete before the local arrive).
// A final barrier.cluster keeps every CTA's shared memory alive until all peers have stopped sending.
//
// Reference arithmetic: models/text_encoder.py:182-216 (posterior biGRU), models/decoder.py:175-203 +
// models/attn_model.py:20-46 (decoder step); same saved activations as recurrent.cuh / train.cuh.
#pragma once
#include "recurrent.cuh"

namespace acvae {

#define CL_STAMP(slot) do { if (tr) tr[(slot)] = clock64(); } while (0)
constexpr int kClE = 256;            // E == H == Hq == A handled here
constexpr int kClC = 8;              // CTAs per cluster
constexpr int kClU = kClE / kClC;    // hidden units per CTA
constexpr int kClThreads = 256;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cl_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_st_async_v4(uint32_t addr, float4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void cl_st_async_f32(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(addr), "f"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Every thread that reads the received data waits itself (the completed phase makes the st.async payload visible to
// the waiting thread).  A protocol bug traps after 10 s of wall clock instead of hanging the GPU.
__device__ __forceinline__ void cl_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spin & 1023) == 1023) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000ull) __trap();
    }
  }
}
// One exchange channel: receive buffer + mbarrier per step parity, all in THIS CTA's shared memory.
struct ClHop {
  uint32_t buf[2];   // shared::cta addresses of the two receive buffers
  uint32_t bar[2];   // shared::cta addresses of the two mbarriers
  uint32_t bytes;    // transaction bytes this CTA receives per use
  __device__ __forceinline__ void arm(int par) const { cl_mbar_expect(bar[par], bytes); }
  __device__ __forceinline__ void wait(int par, int use) const { cl_mbar_wait(bar[par], (uint32_t)(use & 1)); }
};

// Cell epilogue roles after reduce_scatter16<G>(v): lanes 2q and 2q+1 both hold combo q = (row r = q >> 2, unit j = q & 3)
// of the warp's four units.  The four units of a row sit in lanes 8r + 2j (+1): gather4 hands every lane of the group of
// eight the float4 {unit 0..3} of the even lanes (pass 0) or of the odd lanes (pass 1).
__device__ __forceinline__ float4 gather4(float mine, int lane) {
  // lanes of one row group: l8 = lane & 7 = 2j + p.  Pull from the lanes 8r + 2k + p, k = 0..3.
  const int base = (lane & ~7) | (lane & 1);
  float4 o;
  o.x = __shfl_sync(0xffffffffu, mine, base);
  o.y = __shfl_sync(0xffffffffu, mine, base + 2);
  o.z = __shfl_sync(0xffffffffu, mine, base + 4);
  o.w = __shfl_sync(0xffffffffu, mine, base + 6);
  return o;
}

// v[r*4 + j][g] += sum_k a[r][i].k * w[g][j][i].k : four rows x (G gates x four units), weights in registers
template <int G_>
__device__ __forceinline__ void regfma4(const float4 (&a)[4][2], const float4 (&w)[G_][4][2], float (&v)[16][G_]) {
#pragma unroll
  for (int g = 0; g < G_; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r * 4 + j][g] = dot4(a[r][i], w[g][j][i], v[r * 4 + j][g]);
}
template <int G_>
__device__ __forceinline__ void zero16(float (&v)[16][G_]) {
#pragma unroll
  for (int q = 0; q < 16; ++q)
#pragma unroll
    for (int g = 0; g < G_; ++g) v[q][g] = 0.0f;
}
// rows `r0 .. r0+3` of a gathered state buffer [rows][E] in shared memory: this lane's K slice (float4 lane, lane + 32)
__device__ __forceinline__ void ldrows4(const float* H, int r0, int lane, float4 (&a)[4][2]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float4* row = reinterpret_cast<const float4*>(H + (r0 + r) * kClE);
    a[r][0] = row[lane]; a[r][1] = row[lane + 32];
  }
}

// =====================================================================================================================
// posterior biGRU forward (text_encoder.py:188-191): cluster = (direction, group of 8 rows)
// =====================================================================================================================
constexpr int kPostR = 8;
__global__ void __cluster_dims__(kClC, 1, 1) __launch_bounds__(kClThreads, 1) post_cl_fwd_kernel(const __grid_constant__ PostChainFwd p) {
  constexpr int E = kClE, R = kPostR;
  __shared__ __align__(16) float Hb[2][R * E];               // h of the previous step, all units (gathered)
  __shared__ __align__(8) unsigned long long bars[2];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rank = (int)cl_rank(), cid = (int)cl_id();
  const int dir = cid & 1, n0 = (cid >> 1) * R;
  const int T = p.T, N = p.N;
  const int uw = rank * kClU + 4 * w;                          // first of this warp's four units
  if (tid == 0) {
    cl_mbar_init(cl_smem(&bars[0]), 1); cl_mbar_init(cl_smem(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W_hh rows of this warp's units, this lane's K slice: 3 gates x 4 units x 2 float4 = 96 registers
  float4 wr[3][4][2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4* row = reinterpret_cast<const float4*>(p.whh[dir] + (long long)(g * E + uw + j) * E);
      wr[g][j][0] = __ldg(row + lane); wr[g][j][1] = __ldg(row + lane + 32);
    }
  // epilogue role: lane = 2q + pass; combo q = (row-in-pass, unit); pass 0 = rows 0..3, pass 1 = rows 4..7
  const int q = lane >> 1, pass = lane & 1, j = q & 3;
  const int rloc = pass * 4 + (q >> 2), n = n0 + rloc, u = uw + j;
  const bool live = n < N;
  const int len = live ? p.lens[n] : 0;
  const float bh_r = p.bhh[dir][u], bh_z = p.bhh[dir][E + u], bh_n = p.bhh[dir][2 * E + u];
  ClHop hop{{cl_smem(&Hb[0][0]), cl_smem(&Hb[1][0])}, {cl_smem(&bars[0]), cl_smem(&bars[1])}, (uint32_t)(kClC * R * kClU * 4)};
  __syncthreads();
  cl_sync_all();                                               // every peer's barriers are initialised before anyone sends
  float hprev = 0.0f;                                          // h_{s-1} of this lane's (row, unit): stays in a register
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const int par = s & 1;
    long long* tr = (p.trace && blockIdx.x == 0 && tid == 0) ? p.trace + s * 16 : nullptr;   // profiling only
    CL_STAMP(0);
    if (tid == 0 && s + 1 < T) hop.arm(par);
    float gxr = 0.f, gxz = 0.f, gxn = 0.f;
    if (live) {
      const float* gx = p.gx[dir] + ((long long)n * T + t) * 3 * E + u;
      gxr = __ldg(gx); gxz = __ldg(gx + E); gxn = __ldg(gx + 2 * E);
    }
    float res[3] = {0.f, 0.f, 0.f};
    if (s > 0) {
      hop.wait(par ^ 1, (s - 1) >> 1);
      CL_STAMP(1);
      const float* H = &Hb[par ^ 1][0];
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        float4 a[4][2];
        ldrows4(H, ps * 4, lane, a);
        float v[16][3];
        zero16(v);
        regfma4<3>(a, wr, v);
        reduce_scatter16<3>(v, lane);
        if (ps == pass) { res[0] = v[0][0]; res[1] = v[0][1]; res[2] = v[0][2]; }
      }
    }
    CL_STAMP(2);
    const float hn = res[2] + bh_n;
    const float rg = sigmoidf_(gxr + res[0] + bh_r), zg = sigmoidf_(gxz + res[1] + bh_z);
    const float ng = tanhf(gxn + rg * hn);
    float hnew = (1.0f - zg) * ng + zg * hprev;
    if (t >= len) hnew = 0.0f;                                 // packed sequence: padded outputs are zero
    hprev = hnew;
    CL_STAMP(3);
    if (s + 1 < T) {
      // all-gather: the eight lanes of a row group all hold the row's float4 of this warp's units; lane l8 sends it to CTA l8
      const float4 hv = gather4(hnew, lane);
      // lanes 2j+p of a group carry pass p: rows differ between even and odd lanes, so BOTH send -- 16 float4 per warp and
      // destination pair; lane (l8) -> destinations l8 >> 1 and (l8 >> 1) + 4
      const uint32_t off = (uint32_t)((rloc * E + uw) * 4);
      const int d0 = (lane & 7) >> 1;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint32_t dst = (uint32_t)(d0 + 4 * k);
        cl_st_async_v4(cl_mapa(hop.buf[par] + off, dst), hv, cl_mapa(hop.bar[par], dst));
      }
    }
    // saved activations leave after the exchange is on its way (the sends are the critical path of the next step)
    if (live) {
      float* gs = p.gq[dir] + ((long long)n * T + t) * 4 * E + u;
      gs[0] = rg; gs[E] = zg; gs[2 * E] = ng; gs[3 * E] = hn;
      p.ho[((long long)n * T + t) * 2 * E + dir * E + u] = hnew;
    }
    CL_STAMP(4);
  }
  cl_sync_all();                                               // nobody exits while a peer may still write into its shared memory
}

// =====================================================================================================================
// posterior biGRU backward (BPTT with the packed-sequence mask): cluster = (direction, group of 8 rows).
// Split-K form: CTA `rank` holds dGh of ITS 96 gate columns (just produced), multiplies them with its 96 rows of W_hh for
// ALL 256 output units (thread = output unit, weights in registers) and the partial sums are reduce-scattered: the
// 32 units x 8 rows of CTA d travel as two float4 per thread to CTA d.  No all-gather of the 768-wide dGh is needed.
// =====================================================================================================================
__global__ void __cluster_dims__(kClC, 1, 1) __launch_bounds__(kClThreads, 1) post_cl_bwd_kernel(const __grid_constant__ PostChainBwd p) {
  constexpr int E = kClE, R = kPostR, U = kClU;
  __shared__ __align__(16) float G[2][R][3 * U];             // dGh of this CTA's columns (gate-major), by step parity
  __shared__ __align__(16) float RS[2][kClC][U][R];          // partial dh of this CTA's units from every peer
  __shared__ __align__(8) unsigned long long bars[2];
  const int tid = threadIdx.x;
  const int rank = (int)cl_rank(), cid = (int)cl_id();
  const int dir = cid & 1, n0 = (cid >> 1) * R;
  const int T = p.T, N = p.N;
  const int u0 = rank * U;
  if (tid == 0) {
    cl_mbar_init(cl_smem(&bars[0]), 1); cl_mbar_init(cl_smem(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // product role: thread = output unit `tid`; W_hh[c][tid] for this CTA's 96 rows c (gate-major)
  float wr[3 * U];
#pragma unroll
  for (int m = 0; m < 3 * U; ++m) wr[m] = __ldg(p.whh[dir] + (long long)((m / U) * E + u0 + (m % U)) * E + tid);
  // pointwise role: thread = (row rloc = warp, unit k = lane) of this CTA's units
  const int rloc = tid >> 5, k = tid & 31, n = n0 + rloc, u = u0 + k;
  const bool live = n < N;
  const int len = live ? p.lens[n] : 0;
  ClHop hop{{cl_smem(&RS[0][0][0][0]), cl_smem(&RS[1][0][0][0])}, {cl_smem(&bars[0]), cl_smem(&bars[1])}, (uint32_t)(kClC * U * R * 4)};
  __syncthreads();
  cl_sync_all();
  float carry = 0.0f;
  for (int b = 0; b < T; ++b) {
    const int s = T - 1 - b;                                   // forward position being differentiated
    const int t = dir ? T - 1 - s : s;
    const int tp = dir ? t + 1 : t - 1;
    const int par = b & 1;
    if (tid == 0 && b + 1 < T) hop.arm(par);
    float dho = 0.f, rr = 0.f, z = 0.f, nn = 0.f, ghn = 0.f, hp = 0.f;
    if (live) {
      dho = __ldg(p.dho + ((long long)n * T + t) * 2 * E + dir * E + u);
      const float* g = p.gq[dir] + ((long long)n * T + t) * 4 * E + u;
      rr = __ldg(g); z = __ldg(g + E); nn = __ldg(g + 2 * E); ghn = __ldg(g + 3 * E);
      if (s > 0) hp = __ldg(p.ho + ((long long)n * T + tp) * 2 * E + dir * E + u);
    }
    float dh = dho;
    if (b > 0) {
      hop.wait(par ^ 1, (b - 1) >> 1);
      float sum = 0.0f;
#pragma unroll
      for (int src = 0; src < kClC; ++src) sum += RS[par ^ 1][src][k][rloc];
      dh += carry + sum;
    }
    float dar = 0.f, daz = 0.f, dan = 0.f;
    if (t >= len) {
      carry = 0.0f;
    } else {
      const float dn = dh * (1.0f - z), dz = dh * (hp - nn);
      dan = dn * (1.0f - nn * nn);
      dar = dan * ghn * rr * (1.0f - rr); daz = dz * z * (1.0f - z);
      carry = dh * z;
    }
    if (b + 1 < T) {
      G[par][rloc][k] = dar; G[par][rloc][U + k] = daz; G[par][rloc][2 * U + k] = dan * rr;
      __syncthreads();
      // partial dh_{s-1}[row][unit tid] over this CTA's 96 columns
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.0f;
#pragma unroll
      for (int m = 0; m < 3 * U; m += 4)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 g4 = *reinterpret_cast<const float4*>(&G[par][r][m]);
          acc[r] = fmaf(g4.x, wr[m], acc[r]); acc[r] = fmaf(g4.y, wr[m + 1], acc[r]);
          acc[r] = fmaf(g4.z, wr[m + 2], acc[r]); acc[r] = fmaf(g4.w, wr[m + 3], acc[r]);
        }
      const uint32_t dst = (uint32_t)(tid >> 5);               // unit tid belongs to CTA tid / 32
      const uint32_t off = (uint32_t)(((rank * U + (tid & 31)) * R) * 4);
      const uint32_t rb = cl_mapa(hop.buf[par] + off, dst), mb = cl_mapa(hop.bar[par], dst);
      cl_st_async_v4(rb, make_float4(acc[0], acc[1], acc[2], acc[3]), mb);
      cl_st_async_v4(rb + 16, make_float4(acc[4], acc[5], acc[6], acc[7]), mb);
    }
    if (live) {                                                // saved gradients: after the exchange is on its way
      float* gi = p.dgi[dir] + ((long long)n * T + t) * 3 * E + u;
      float* gh = p.dgh[dir] + ((long long)n * T + t) * 3 * E + u;
      gi[0] = dar; gi[E] = daz; gi[2 * E] = dan;
      gh[0] = dar; gh[E] = daz; gh[2 * E] = dan * rr;
    }
  }
  cl_sync_all();
}

// =====================================================================================================================
// decoder forward chain (decoder.py:183-199, attn_model.py:20-46): cluster = group of kDecR rows, CTA = 32 hidden units
// AND 32 attention columns.  Per step three all-to-all hops:
//   A  all-gather h_{t-1} [R,256] (float4 per lane)                      -> query projection q[r, my 32 a] and the
//                                                                           h-part of the GRU gates (weights in registers)
//   RS reduce-scatter of the partial scores sum_{a in my 32} v_a tanh(P[r,j,a] + q[r,a]) over the R*Te (row, frame) pairs
//   AG all-gather of the summed scores                                   -> every CTA runs the masked softmax of all R rows
// The context never crosses CTAs: with Mg[n,j,:] = mem[n,j,:] . W_ih[:, E:2E]^T hoisted into one batched GEMM, the ctx
// part of the gate pre-activations of MY units is sum_j alpha[r,j] Mg[r,j, my columns] (K = Te instead of K = E, and no
// all-gather of ctx).  ctx itself (needed by the weight gradients) is recomputed after the chain from the saved weights
// (attn_ctx_kernel).  Shared memory per CTA: Mg tile R*Te*100 + P tile R*Te*36 + Wq slice 32*256 floats (~180 KB, Te 62).
// =====================================================================================================================
constexpr int kDecR = 4;
constexpr int kMgLd = 3 * kClU + 4;    // 100: row stride of the Mg tile (conflict-free for lanes = (row, unit, frame parity))
constexpr int kPsLd = kClU + 4;        // 36: row stride of the P tile (conflict-free 16-byte loads by lanes = frames)
constexpr int kAlLd = 104;             // row stride of the per-row softmax weights (Te <= 96; rows 8 banks apart: lanes = (row, frame parity))

struct DecClFwd {
  int N, T, Te;
  const float* gx;        // [N,T,3E] = [emb | z] . W_ih[:, {0:E, 2E:3E}]^T + b_ih
  const float* attn_w;    // [A,2E] (query columns first)
  const float* attn_v;    // [A]
  const float* whh;       // [3E,E]
  const float* bhh;       // [3E]
  const float *Pd, *Mg;   // [N,Te,A], [N,Te,3E]
  const int* mem_lens;
  float *qp, *w, *gates, *out;   // saved: [N,T,A], [N,T,Te], [N,T,4E], [N,T,E]
  float* aw;              // [N,Te,T] user-visible attention weights or NULL
  long long* trace;       // optional [T][16] clock64 stamps of thread 0 of CTA 0 (profiles/cluster_trace.py) or NULL
  int t0, t1;             // steps [t0, t1) of the chain (t1 == 0: T); t0 > 0 resumes from the saved state of step t0 - 1 --
                          // scheduled sampling cuts the chain at every free step, whose input word is the previous step's arg-max
};
inline int dec_cl_slice(int Te) { return (kDecR * Te + kClC - 1) / kClC; }          // (row, frame) pairs owned per CTA in the score reduction
inline size_t dec_cl_fwd_smem(int Te) {
  const size_t sl = dec_cl_slice(Te);
  return ((size_t)kDecR * Te * (kMgLd + kPsLd) + (size_t)kClU * kClE + 2 * kDecR * kClE /*Hb*/ + 2 * kClC * sl /*RSb*/ +
          2 * kClC * sl /*SC*/ + kDecR * kAlLd /*alpha rows*/ + kDecR * kClU /*qs*/ + 2 * kClU /*v2s, lens*/ + 16 /*barriers*/) * sizeof(float);
}

__global__ void __cluster_dims__(kClC, 1, 1) __launch_bounds__(kClThreads, 1) dec_cl_fwd_kernel(const __grid_constant__ DecClFwd p) {
  constexpr int E = kClE, R = kDecR, U = kClU;
  extern __shared__ __align__(16) float dsm[];
  const int Te = p.Te, T = p.T, N = p.N;
  const int SL = (R * Te + kClC - 1) / kClC;                   // score-slice length
  float* Mgs = dsm;                                            // [R][Te][kMgLd]   gate-major columns g*32 + k of my units
  float* Ps = Mgs + (size_t)R * Te * kMgLd;                    // [R][Te][kPsLd]   my 32 attention columns, x 2 log2 e
  float* Wq = Ps + (size_t)R * Te * kPsLd;                     // [32][E]          query-projection rows of my columns
  float* Hb = Wq + U * E;                                      // [2][R][E]
  float* RSb = Hb + 2 * R * E;                                 // [2][C][SL]
  float* SC = RSb + 2 * kClC * SL;                             // [2][C*SL]
  float* al = SC + 2 * kClC * SL;                              // [R][kAlLd]: softmax weights of every row (Te <= 96)
  float* qs = al + R * kAlLd;                                  // [R][32]  2 log2 e x query projection, my columns
  float* v2s = qs + R * U;                                     // [32] -2 v_a
  int* lens_s = reinterpret_cast<int*>(v2s + U);               // [R] valid frames per row (0 for rows past N)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(v2s + 2 * U);   // [3][2]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rank = (int)cl_rank(), cid = (int)cl_id();
  const int n0 = cid * R, u0 = rank * U, uw = u0 + 4 * w;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) cl_mbar_init(cl_smem(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- one-time loads ----
  for (int i = tid; i < R * Te * 3 * (U / 4); i += kClThreads) {       // Mg tile: 3 gates x 8 float4 per (row, frame)
    const int c4 = i % (U / 4), g = (i / (U / 4)) % 3, rj = i / (3 * (U / 4));
    const int r = rj / Te, jj = rj % Te, n = n0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = __ldg(reinterpret_cast<const float4*>(p.Mg + ((long long)n * Te + jj) * 3 * E + g * E + u0) + c4);
    *reinterpret_cast<float4*>(Mgs + (size_t)rj * kMgLd + g * U + c4 * 4) = v;
  }
  for (int i = tid; i < R * Te * (U / 4); i += kClThreads) {           // P tile, pre-scaled by 2 log2 e (see attn_tanh)
    const int c4 = i % (U / 4), rj = i / (U / 4);
    const int r = rj / Te, jj = rj % Te, n = n0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = __ldg(reinterpret_cast<const float4*>(p.Pd + ((long long)n * Te + jj) * E + u0) + c4);
    v.x *= kTwoLog2e; v.y *= kTwoLog2e; v.z *= kTwoLog2e; v.w *= kTwoLog2e;
    *reinterpret_cast<float4*>(Ps + (size_t)rj * kPsLd + c4 * 4) = v;
  }
  for (int i = tid; i < U * E / 4; i += kClThreads) {                  // Wq[a][k] = attn_w[(u0 + a) * 2E + k], k < E
    const int a = i / (E / 4), k4 = i % (E / 4);
    reinterpret_cast<float4*>(Wq)[i] = __ldg(reinterpret_cast<const float4*>(p.attn_w + (long long)(u0 + a) * 2 * E) + k4);
  }
  float vpart = 0.0f;                                                  // sum of v_a over my columns (added once per score)
  if (tid < U) { const float x = p.attn_v[u0 + tid]; v2s[tid] = -2.0f * x; }
  if (tid < R) lens_s[tid] = n0 + tid < N ? max(1, min(p.mem_lens[n0 + tid], Te)) : 0;
  for (int a = 0; a < U; ++a) vpart += p.attn_v[u0 + a];
  // W_hh rows of this warp's four units, this lane's K slice (96 registers)
  float4 wr[3][4][2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4* row = reinterpret_cast<const float4*>(p.whh + (long long)(g * E + uw + j) * E);
      wr[g][j][0] = __ldg(row + lane); wr[g][j][1] = __ldg(row + lane + 32);
    }
  // ---- roles ----
  // cell: lane = 2q + half, combo q = (row r4 = q >> 2, unit j = q & 3); both halves hold the sums, each accumulates the
  // frames of its parity in the ctx part
  const int q = lane >> 1, half = lane & 1, j = q & 3, r4 = q >> 2;
  const int n = n0 + r4, u = uw + j, kcol = 4 * w + j;                // kcol: unit index inside the CTA
  const bool live = n < N;
  const float bh_r = p.bhh[u], bh_z = p.bhh[E + u], bh_n = p.bhh[2 * E + u];
  ClHop hopA{{cl_smem(Hb), cl_smem(Hb + R * E)}, {cl_smem(&bars[0]), cl_smem(&bars[1])}, (uint32_t)(kClC * R * U * 4)};
  const int mycount = max(0, min(SL, R * Te - rank * SL));           // pairs of my score slice
  ClHop hopR{{cl_smem(RSb), cl_smem(RSb + kClC * SL)}, {cl_smem(&bars[2]), cl_smem(&bars[3])}, (uint32_t)(kClC * mycount * 4)};
  ClHop hopG{{cl_smem(SC), cl_smem(SC + kClC * SL)}, {cl_smem(&bars[4]), cl_smem(&bars[5])}, (uint32_t)(R * Te * 4)};
  const int t0 = p.t0, nsteps = (p.t1 > 0 ? p.t1 : T) - p.t0;
  float hprev = 0.0f;
  if (t0 > 0) {
    // resume: h_{t0-1} of my rows from the saved outputs, into the buffer the first step's query projection reads
    hprev = live ? p.out[((long long)n * T + t0 - 1) * E + u] : 0.0f;
    for (int i = tid; i < R * E; i += kClThreads) {
      const int r = i / E, k = i - r * E;
      Hb[R * E + i] = n0 + r < N ? p.out[((long long)(n0 + r) * T + t0 - 1) * E + k] : 0.0f;
    }
  }
  __syncthreads();
  cl_sync_all();
  for (int s = 0; s < nsteps; ++s) {
    const int t = t0 + s, par = s & 1;       // the hop protocol counts the steps of THIS launch
    long long* tr = (p.trace && blockIdx.x == 0 && tid == 0) ? p.trace + t * 16 : nullptr;   // profiling only
    CL_STAMP(0);
    if (tid == 0) {
      if (s + 1 < nsteps) hopA.arm(par);
      hopR.arm(par); hopG.arm(par);
    }
    float gxr = 0.f, gxz = 0.f, gxn = 0.f;
    if (live) {
      const float* gx = p.gx + ((long long)n * T + t) * 3 * E + u;
      gxr = __ldg(gx); gxz = __ldg(gx + E); gxn = __ldg(gx + 2 * E);
    }
    // ---- A: h_{t-1} gathered -> query projection of my 32 columns (zero query at t = 0, decoder.py:94-98) ----
    float4 a[4][2];
    if (t > 0) {
      if (s > 0) hopA.wait(par ^ 1, (s - 1) >> 1);
      CL_STAMP(1);
      ldrows4(Hb + (par ^ 1) * R * E, 0, lane, a);
      float v[16][1];
      zero16(v);
#pragma unroll
      for (int jq = 0; jq < 4; ++jq) {
        const float4* wrow = reinterpret_cast<const float4*>(Wq + (4 * w + jq) * E);
        const float4 w0 = wrow[lane], w1 = wrow[lane + 32];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r * 4 + jq][0] = dot4(a[r][1], w1, dot4(a[r][0], w0, v[r * 4 + jq][0]));
      }
      reduce_scatter16<1>(v, lane);
      if (half == 0) {
        qs[r4 * U + kcol] = v[0][0] * kTwoLog2e;
        if (live) p.qp[((long long)n * T + t) * E + u] = v[0][0];
      }
    } else if (half == 0) {
      qs[r4 * U + kcol] = 0.0f;
      if (live) p.qp[((long long)n * T + t) * E + u] = 0.0f;
    }
    __syncthreads();
    CL_STAMP(2);
    // ---- partial scores of my columns, one (row, frame) pair per thread and pass; reduce-scatter ----
    for (int i = tid; i < R * Te; i += kClThreads) {
      const int sr = i / Te, sj = i - sr * Te;
      float sc = 0.0f;
      if (sj < lens_s[sr]) {
        const float4* pr = reinterpret_cast<const float4*>(Ps + (size_t)i * kPsLd);
        const float4* qr = reinterpret_cast<const float4*>(qs + sr * U);
        const float4* vr = reinterpret_cast<const float4*>(v2s);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < U / 4; ++c4) {
          const float4 x = pr[c4], qq = qr[c4], vv = vr[c4];
          s0 = fmaf(vv.x, rcp_approx(ex2_approx(x.x + qq.x) + 1.0f), s0);
          s1 = fmaf(vv.y, rcp_approx(ex2_approx(x.y + qq.y) + 1.0f), s1);
          s2 = fmaf(vv.z, rcp_approx(ex2_approx(x.z + qq.z) + 1.0f), s2);
          s3 = fmaf(vv.w, rcp_approx(ex2_approx(x.w + qq.w) + 1.0f), s3);
        }
        sc = ((s0 + s1) + (s2 + s3)) + vpart;
      }
      const uint32_t dst = (uint32_t)(i / SL);
      const uint32_t off = (uint32_t)((rank * SL + i % SL) * 4);
      cl_st_async_f32(cl_mapa(hopR.buf[par] + off, dst), sc, cl_mapa(hopR.bar[par], dst));
    }
    CL_STAMP(3);
    // ---- h part of the GRU gates while the scores travel ----
    float hres[3] = {0.f, 0.f, 0.f};
    if (t > 0) {
      float v[16][3];
      zero16(v);
      regfma4<3>(a, wr, v);
      reduce_scatter16<3>(v, lane);
      hres[0] = v[0][0]; hres[1] = v[0][1]; hres[2] = v[0][2];
    }
    // ---- RS landed: sum the eight partials of my slice and send the sums to every CTA ----
    CL_STAMP(4);
    hopR.wait(par, s >> 1);
    CL_STAMP(5);
    for (int i = tid; i < kClC * mycount; i += kClThreads) {
      const int dst = i / mycount, k = i - dst * mycount;
      const float* rs = RSb + par * kClC * SL + k;
      float sum = 0.0f;
#pragma unroll
      for (int src = 0; src < kClC; ++src) sum += rs[src * SL];
      cl_st_async_f32(cl_mapa(hopG.buf[par] + (uint32_t)((rank * SL + k) * 4), (uint32_t)dst), sum, cl_mapa(hopG.bar[par], (uint32_t)dst));
    }
    // ---- AG landed: masked softmax of every row (warp r = row r) ----
    hopG.wait(par, s >> 1);
    CL_STAMP(6);
    if (w < R) {
      const int nr = n0 + w;
      const int len = lens_s[w];
      const float* sc = SC + par * kClC * SL + w * Te;
      float x[3];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 3; ++i) { const int jj = lane + 32 * i; x[i] = jj < len ? sc[jj] : -INFINITY; mx = fmaxf(mx, x[i]); }
      mx = warp_max(mx);
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i) { const int jj = lane + 32 * i; x[i] = jj < len ? expf(x[i] - mx) : 0.0f; sum += x[i]; }
      sum = warp_sum(sum);
      const float inv = len > 0 ? 1.0f / sum : 0.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int jj = lane + 32 * i;
        if (jj < Te) {
          const float wv = x[i] * inv;
          al[w * kAlLd + jj] = wv;
          if (rank == w && nr < N) {                                 // CTA r saves row r (every CTA holds all rows)
            p.w[((long long)nr * T + t) * Te + jj] = wv;
            if (p.aw) p.aw[((long long)nr * Te + jj) * T + t] = wv;
          }
        }
      }
    }
    __syncthreads();
    CL_STAMP(7);
    // ---- ctx part of the gates: sum_j alpha[r, j] Mg[r, j, (gate, my unit)], frames of my parity; GRU cell ----
    {
      const float* mrow = Mgs + (size_t)(r4 * Te) * kMgLd + kcol;
      const float* arow = al + r4 * kAlLd;
      // four frames per pass with all twelve operands requested before the first FMA (two accumulator sets): the plain loop
      // paid one shared-memory round trip per frame (2000 cycles for 31 frames, profiles/cluster_trace.py)
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, y0 = 0.f, y1 = 0.f, y2 = 0.f;
      int jj = half;
      for (; jj + 6 < Te; jj += 8) {
        const float w0 = arow[jj], w1 = arow[jj + 2], w2 = arow[jj + 4], w3 = arow[jj + 6];
        const float* m0 = mrow + (size_t)jj * kMgLd;
        const float* m1 = m0 + 2 * kMgLd; const float* m2 = m0 + 4 * kMgLd; const float* m3 = m0 + 6 * kMgLd;
        const float a0 = m0[0], a1 = m0[U], a2 = m0[2 * U], b0 = m1[0], b1 = m1[U], b2 = m1[2 * U];
        const float c0 = m2[0], c1 = m2[U], c2 = m2[2 * U], d0 = m3[0], d1 = m3[U], d2 = m3[2 * U];
        x0 = fmaf(w0, a0, x0); x1 = fmaf(w0, a1, x1); x2 = fmaf(w0, a2, x2);
        y0 = fmaf(w1, b0, y0); y1 = fmaf(w1, b1, y1); y2 = fmaf(w1, b2, y2);
        x0 = fmaf(w2, c0, x0); x1 = fmaf(w2, c1, x1); x2 = fmaf(w2, c2, x2);
        y0 = fmaf(w3, d0, y0); y1 = fmaf(w3, d1, y1); y2 = fmaf(w3, d2, y2);
      }
      for (; jj < Te; jj += 2) {
        const float wv = arow[jj];
        const float* m = mrow + (size_t)jj * kMgLd;
        x0 = fmaf(wv, m[0], x0); x1 = fmaf(wv, m[U], x1); x2 = fmaf(wv, m[2 * U], x2);
      }
      x0 += y0; x1 += y1; x2 += y2;
      x0 += __shfl_xor_sync(0xffffffffu, x0, 1); x1 += __shfl_xor_sync(0xffffffffu, x1, 1); x2 += __shfl_xor_sync(0xffffffffu, x2, 1);
      CL_STAMP(8);
      const float hn = hres[2] + bh_n;
      const float rg = sigmoidf_(x0 + gxr + hres[0] + bh_r);
      const float zg = sigmoidf_(x1 + gxz + hres[1] + bh_z);
      const float ng = tanhf(x2 + gxn + rg * hn);
      const float hnew = (1.0f - zg) * ng + zg * hprev;
      hprev = hnew;
      if (s + 1 < nsteps) {
        // all-gather h_t: the eight lanes of a row group hold the row's four units; lane l8 sends the float4 to CTA l8
        const float4 hv = gather4(hnew, lane);
        const uint32_t dst = (uint32_t)(lane & 7);
        cl_st_async_v4(cl_mapa(hopA.buf[par] + (uint32_t)((r4 * E + uw) * 4), dst), hv, cl_mapa(hopA.bar[par], dst));
      }
      if (live && half == 0) {                                   // saved activations: after the exchange is on its way
        float* gs = p.gates + ((long long)n * T + t) * 4 * E + u;
        gs[0] = rg; gs[E] = zg; gs[2 * E] = ng; gs[3 * E] = hn;
        p.out[((long long)n * T + t) * E + u] = hnew;
      }
      CL_STAMP(9);
    }
  }
  cl_sync_all();
}

// =====================================================================================================================
// decoder backward chain (BPTT through the GRU, the attention weights and the query projection), same partition.
// Per step three all-to-all hops:
//   RS/AG  d alpha[r,j] = sum_c dGi[r,c] Mg[r,j,c] (= d ctx . mem_j): partial over MY 96 gate columns, reduce-scattered
//          and all-gathered like the forward scores; every CTA then runs the softmax backward of all R rows
//   D      d h_{t-1}[r,:] = dGh_t . W_hh + d qp_t . Wq in split-K form: MY 96 rows of W_hh and 32 rows of Wq (thread = output
//          unit, 128 weights in registers) against the dGh / d qp columns this CTA has just produced; the partial sums
//          are reduce-scattered (one float4 = four rows per thread)
// d ctx (needed by the memory gradients) is formed after the chain by one batched GEMM dGi . W_ih[:, E:2E].
// =====================================================================================================================
struct DecClBwd {
  int N, T, Te;
  const float* dout;      // [N,T,E] upstream gradient of the GRU outputs (incl. the pooled global head)
  const float* attn_w;    // [A,2E]
  const float* attn_v;    // [A]
  const float* whh;       // [3E,E]
  const float *Pd, *Mg;   // [N,Te,A], [N,Te,3E]
  const int* mem_lens;
  const float *qp, *w, *gates, *out;   // saved by the forward
  float *dgi, *dgh;       // [N,T,3E]
  float *ds, *dqp;        // [N,T,Te], [N,T,A]
};
inline size_t dec_cl_bwd_smem(int Te) {
  const size_t sl = dec_cl_slice(Te);
  return ((size_t)kDecR * Te * (kMgLd + kPsLd) + 2 * kDecR * 4 * kClU /*G*/ + 2 * kClC * kClU * kDecR /*RSD*/ + 2 * kClC * sl /*RSb*/ +
          2 * kClC * sl /*DA*/ + 2 * kDecR * 96 /*al, dsS*/ + 2 * kDecR * kClU /*qs, Q*/ + 2 * kClU /*lens*/ + 16 /*barriers*/) * sizeof(float);
}

__global__ void __cluster_dims__(kClC, 1, 1) __launch_bounds__(kClThreads, 1) dec_cl_bwd_kernel(const __grid_constant__ DecClBwd p) {
  constexpr int E = kClE, R = kDecR, U = kClU;
  extern __shared__ __align__(16) float dsm[];
  const int Te = p.Te, T = p.T, N = p.N;
  const int SL = (R * Te + kClC - 1) / kClC;
  float* Mgs = dsm;                                            // [R][Te][kMgLd]
  float* Ps = Mgs + (size_t)R * Te * kMgLd;                    // [R][Te][kPsLd]  x 2 log2 e
  float* G = Ps + (size_t)R * Te * kPsLd;                      // [2][R][4U]: dar | daz | dan | dan*r of my units
  float* RSD = G + 2 * R * 4 * U;                              // [2][C][U][R] partial dh of my units from every peer
  float* RSb = RSD + 2 * kClC * U * R;                         // [2][C][SL]   partial d alpha of my slice
  float* DA = RSb + 2 * kClC * SL;                             // [2][C*SL]    d alpha, all rows
  float* al = DA + 2 * kClC * SL;                              // [R][96] saved attention weights of the step
  float* dsS = al + R * 96;                                    // [R][96] 4 * d score
  float* qs = dsS + R * 96;                                    // [R][32] 2 log2 e * saved query projection, my columns
  float* Q = qs + R * U;                                       // [R][32] d qp of my columns
  int* lens_s = reinterpret_cast<int*>(Q + R * U);             // [R]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(Q + R * U + U);   // [3][2]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rank = (int)cl_rank(), cid = (int)cl_id();
  const int n0 = cid * R, u0 = rank * U;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) cl_mbar_init(cl_smem(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < R * Te * 3 * (U / 4); i += kClThreads) {
    const int c4 = i % (U / 4), g = (i / (U / 4)) % 3, rj = i / (3 * (U / 4));
    const int r = rj / Te, jj = rj % Te, n = n0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = __ldg(reinterpret_cast<const float4*>(p.Mg + ((long long)n * Te + jj) * 3 * E + g * E + u0) + c4);
    *reinterpret_cast<float4*>(Mgs + (size_t)rj * kMgLd + g * U + c4 * 4) = v;
  }
  for (int i = tid; i < R * Te * (U / 4); i += kClThreads) {
    const int c4 = i % (U / 4), rj = i / (U / 4);
    const int r = rj / Te, jj = rj % Te, n = n0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = __ldg(reinterpret_cast<const float4*>(p.Pd + ((long long)n * Te + jj) * E + u0) + c4);
    v.x *= kTwoLog2e; v.y *= kTwoLog2e; v.z *= kTwoLog2e; v.w *= kTwoLog2e;
    *reinterpret_cast<float4*>(Ps + (size_t)rj * kPsLd + c4 * 4) = v;
  }
  if (tid < R) lens_s[tid] = n0 + tid < N ? max(1, min(p.mem_lens[n0 + tid], Te)) : 0;
  // product role: thread = output unit tid; rows of W_hh (my 96 gate columns) and of Wq (my 32 attention columns)
  float wr[4 * U];
#pragma unroll
  for (int m = 0; m < 3 * U; ++m) wr[m] = __ldg(p.whh + (long long)((m / U) * E + u0 + (m % U)) * E + tid);
#pragma unroll
  for (int a = 0; a < U; ++a) wr[3 * U + a] = __ldg(p.attn_w + (long long)(u0 + a) * 2 * E + tid);
  // pointwise role (threads < 128): (row = warp, unit = lane)
  const bool pw = tid < R * U;
  const int pn = n0 + w, pu = u0 + lane;
  const bool plive = pw && pn < N;
  // d qp role: tid = r*64 + a*2 + half
  const int qr_ = tid >> 6, qa = (tid & 63) >> 1, qh = tid & 1;
  const float va = p.attn_v[u0 + qa];
  const bool qlive = n0 + qr_ < N;
  const int mycount = max(0, min(SL, R * Te - rank * SL));
  ClHop hopD{{cl_smem(RSD), cl_smem(RSD + kClC * U * R)}, {cl_smem(&bars[0]), cl_smem(&bars[1])}, (uint32_t)(kClC * U * R * 4)};
  ClHop hopR{{cl_smem(RSb), cl_smem(RSb + kClC * SL)}, {cl_smem(&bars[2]), cl_smem(&bars[3])}, (uint32_t)(kClC * mycount * 4)};
  ClHop hopG{{cl_smem(DA), cl_smem(DA + kClC * SL)}, {cl_smem(&bars[4]), cl_smem(&bars[5])}, (uint32_t)(R * Te * 4)};
  __syncthreads();
  cl_sync_all();
  float carry = 0.0f;
  for (int b = 0; b < T; ++b) {
    const int t = T - 1 - b;
    const int par = b & 1;
    if (tid == 0) {
      if (t > 0) hopD.arm(par);
      hopR.arm(par); hopG.arm(par);
    }
    // ---- operands of the step from global memory (saved by the forward) ----
    float dh = 0.f, rr = 0.f, z = 0.f, nn = 0.f, ghn = 0.f, hp = 0.f;
    if (plive) {
      dh = __ldg(p.dout + ((long long)pn * T + t) * E + pu);
      const float* g = p.gates + ((long long)pn * T + t) * 4 * E + pu;
      rr = __ldg(g); z = __ldg(g + E); nn = __ldg(g + 2 * E); ghn = __ldg(g + 3 * E);
      if (t > 0) hp = __ldg(p.out + ((long long)pn * T + t - 1) * E + pu);
    }
    if (w < R) {                                               // warp r: attention weights of row r
      const int nr = n0 + w;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int jj = lane + 32 * i;
        if (jj < Te) al[w * 96 + jj] = (nr < N && jj < lens_s[w]) ? __ldg(p.w + ((long long)nr * T + t) * Te + jj) : 0.0f;
      }
    } else {                                                   // warps 4..7: saved query projections of my columns
      const int r = w - R, nr = n0 + r;
      qs[r * U + lane] = nr < N ? kTwoLog2e * __ldg(p.qp + ((long long)nr * T + t) * E + u0 + lane) : 0.0f;
    }
    // ---- D: dh_t of my units = upstream + carry + reduce-scattered partial sums; GRU pointwise backward ----
    if (pw) {
      if (b > 0) {
        hopD.wait(par ^ 1, (b - 1) >> 1);
        const float* rs = RSD + (par ^ 1) * kClC * U * R + lane * R + w;
        float sum = 0.0f;
#pragma unroll
        for (int src = 0; src < kClC; ++src) sum += rs[src * U * R];
        dh += carry + sum;
      }
      const float dn = dh * (1.0f - z), dz = dh * (hp - nn);
      const float dan = dn * (1.0f - nn * nn);
      const float dar = dan * ghn * rr * (1.0f - rr), daz = dz * z * (1.0f - z);
      carry = dh * z;
      float* g = G + (par * R + w) * 4 * U + lane;
      g[0] = dar; g[U] = daz; g[2 * U] = dan; g[3 * U] = dan * rr;
    }
    __syncthreads();
    // ---- partial d alpha over my 96 gate columns, one (row, frame) pair per thread and pass; reduce-scatter ----
    for (int i = tid; i < R * Te; i += kClThreads) {
      const int sr = i / Te, sj = i - sr * Te;
      float da = 0.0f;
      if (sj < lens_s[sr]) {
        const float4* m = reinterpret_cast<const float4*>(Mgs + (size_t)i * kMgLd);
        const float4* g = reinterpret_cast<const float4*>(G + (par * R + sr) * 4 * U);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 3 * U / 4; ++c4) {
          const float4 x = m[c4], y = g[c4];
          s0 = fmaf(x.x, y.x, s0); s1 = fmaf(x.y, y.y, s1); s2 = fmaf(x.z, y.z, s2); s3 = fmaf(x.w, y.w, s3);
        }
        da = (s0 + s1) + (s2 + s3);
      }
      const uint32_t dst = (uint32_t)(i / SL);
      cl_st_async_f32(cl_mapa(hopR.buf[par] + (uint32_t)((rank * SL + i % SL) * 4), dst), da, cl_mapa(hopR.bar[par], dst));
    }
    if (plive) {                                               // saved gradients (from the shared-memory copy): off the critical path
      const float* g = G + (par * R + w) * 4 * U + lane;
      float* gi = p.dgi + ((long long)pn * T + t) * 3 * E + pu;
      float* gh = p.dgh + ((long long)pn * T + t) * 3 * E + pu;
      const float dar = g[0], daz = g[U], dan = g[2 * U], danr = g[3 * U];
      gi[0] = dar; gi[E] = daz; gi[2 * E] = dan;
      gh[0] = dar; gh[E] = daz; gh[2 * E] = danr;
    }
    // ---- dGh part of the partial dh_{t-1} while d alpha travels: (dar, daz, dan*r) . my 96 rows of W_hh ----
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
    if (t > 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* g = G + (par * R + r) * 4 * U;
#pragma unroll
        for (int m = 0; m < 2 * U; m += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(g + m);
          acc[r] = fmaf(g4.x, wr[m], acc[r]); acc[r] = fmaf(g4.y, wr[m + 1], acc[r]);
          acc[r] = fmaf(g4.z, wr[m + 2], acc[r]); acc[r] = fmaf(g4.w, wr[m + 3], acc[r]);
        }
#pragma unroll
        for (int m = 0; m < U; m += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(g + 3 * U + m);
          acc[r] = fmaf(g4.x, wr[2 * U + m], acc[r]); acc[r] = fmaf(g4.y, wr[2 * U + m + 1], acc[r]);
          acc[r] = fmaf(g4.z, wr[2 * U + m + 2], acc[r]); acc[r] = fmaf(g4.w, wr[2 * U + m + 3], acc[r]);
        }
      }
    }
    // ---- RS landed: sum the eight partials of my slice, all-gather the sums ----
    hopR.wait(par, b >> 1);
    for (int i = tid; i < kClC * mycount; i += kClThreads) {
      const int dst = i / mycount, k = i - dst * mycount;
      const float* rs = RSb + par * kClC * SL + k;
      float sum = 0.0f;
#pragma unroll
      for (int src = 0; src < kClC; ++src) sum += rs[src * SL];
      cl_st_async_f32(cl_mapa(hopG.buf[par] + (uint32_t)((rank * SL + k) * 4), (uint32_t)dst), sum, cl_mapa(hopG.bar[par], (uint32_t)dst));
    }
    // ---- AG landed: softmax backward d s_j = w_j (d alpha_j - sum_k w_k d alpha_k), warp r = row r ----
    hopG.wait(par, b >> 1);
    if (w < R) {
      const int nr = n0 + w, len = lens_s[w];
      const float* da = DA + par * kClC * SL + w * Te;
      float x[3], wv[3], dot = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int jj = lane + 32 * i;
        wv[i] = jj < len ? al[w * 96 + jj] : 0.0f;
        x[i] = jj < len ? da[jj] : 0.0f;
        dot = fmaf(wv[i], x[i], dot);
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int jj = lane + 32 * i;
        if (jj < Te) {
          const float dsv = wv[i] * (x[i] - dot);
          dsS[w * 96 + jj] = 4.0f * dsv;                       // the factor of 1 - tanh^2 = 4 r (1 - r)
          if (rank == w && nr < N) p.ds[((long long)nr * T + t) * Te + jj] = dsv;
        }
      }
    }
    __syncthreads();
    // ---- d qp[r, a] = v_a sum_j d s_j (1 - tanh^2(P[r,j,a] + q[r,a])) for my 32 columns, frames of my parity ----
    {
      const int len = lens_s[qr_];
      const float qv = qs[qr_ * U + qa];
      const float* pc = Ps + (size_t)(qr_ * Te) * kPsLd + qa;
      const float* dsr = dsS + qr_ * 96;
      float s0 = 0.f, s1 = 0.f;
      int jj = qh;
      for (; jj + 2 < len; jj += 4) {
        const float r0 = rcp_approx(ex2_approx(pc[(size_t)jj * kPsLd] + qv) + 1.0f);
        const float r1 = rcp_approx(ex2_approx(pc[(size_t)(jj + 2) * kPsLd] + qv) + 1.0f);
        s0 = fmaf(dsr[jj], fmaf(-r0, r0, r0), s0);
        s1 = fmaf(dsr[jj + 2], fmaf(-r1, r1, r1), s1);
      }
      for (; jj < len; jj += 2) {
        const float r0 = rcp_approx(ex2_approx(pc[(size_t)jj * kPsLd] + qv) + 1.0f);
        s0 = fmaf(dsr[jj], fmaf(-r0, r0, r0), s0);
      }
      float sdq = s0 + s1;
      sdq += __shfl_xor_sync(0xffffffffu, sdq, 1);
      const float dq = va * sdq;
      if (qh == 0) {
        Q[qr_ * U + qa] = dq;
        if (qlive) p.dqp[((long long)(n0 + qr_) * T + t) * E + u0 + qa] = dq;
      }
    }
    if (t > 0) {
      __syncthreads();
      // ---- d qp part of the partial dh_{t-1}; reduce-scatter: unit tid belongs to CTA tid / 32 ----
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* qrow = Q + r * U;
#pragma unroll
        for (int a = 0; a < U; a += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(qrow + a);
          acc[r] = fmaf(q4.x, wr[3 * U + a], acc[r]); acc[r] = fmaf(q4.y, wr[3 * U + a + 1], acc[r]);
          acc[r] = fmaf(q4.z, wr[3 * U + a + 2], acc[r]); acc[r] = fmaf(q4.w, wr[3 * U + a + 3], acc[r]);
        }
      }
      const uint32_t dst = (uint32_t)(tid >> 5);
      const uint32_t off = (uint32_t)(((rank * U + (tid & 31)) * R) * 4);
      cl_st_async_v4(cl_mapa(hopD.buf[par] + off, dst), make_float4(acc[0], acc[1], acc[2], acc[3]), cl_mapa(hopD.bar[par], dst));
    }
  }
  cl_sync_all();
}

// ctx[n,t,:] = sum_j w[n,t,j] mem[n,j,:] from the saved attention weights (the cluster chain never forms the context:
// see above); needed by the weight gradient dW_ih[:, E:2E] = dGi^T . ctx and by VAEModel's rnn_input.
__global__ void __launch_bounds__(256) attn_ctx_kernel(int T, int Te, int E, const float* __restrict__ w, const float* __restrict__ mem,
                                                       const int* __restrict__ mem_lens, float* __restrict__ ctx) {
  const int n = blockIdx.x, t = blockIdx.y;
  const int len = max(1, min(mem_lens[n], Te));
  const float* wr = w + ((long long)n * T + t) * Te;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float* m = mem + (long long)n * Te * E + e;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    int jj = 0;
    for (; jj + 4 <= len; jj += 4) {
      c0 = fmaf(wr[jj], m[(long long)jj * E], c0); c1 = fmaf(wr[jj + 1], m[(long long)(jj + 1) * E], c1);
      c2 = fmaf(wr[jj + 2], m[(long long)(jj + 2) * E], c2); c3 = fmaf(wr[jj + 3], m[(long long)(jj + 3) * E], c3);
    }
    for (; jj < len; ++jj) c0 = fmaf(wr[jj], m[(long long)jj * E], c0);
    ctx[((long long)n * T + t) * E + e] = (c0 + c1) + (c2 + c3);
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
template <typename Kern, typename... P>
inline int launch_cluster_chain(Kern kern, int nclusters, size_t smem, cudaStream_t st, const char* name, const P&... p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * kClC);
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  // the chains are the critical path of the step: their CTAs are placed before those of the batched GEMMs queued on the
  // side streams (the cluster dimension itself is a compile-time attribute of the kernels)
  static int prio_dev[kMaxDevices];
  static bool have[kMaxDevices] = {false};
  const int dev = current_device();
  if (!have[dev]) { int lo = 0, hi = 0; prio_dev[dev] = (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess) ? hi : 0; have[dev] = true; }
  at[0].id = cudaLaunchAttributePriority;
  at[0].val.priority = prio_dev[dev];
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const bool probe = probe_match(name);
  if (probe) cudaEventRecord(g_probe.e0, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p...);
  if (probe) { cudaEventRecord(g_probe.e1, st); ++g_probe.hits; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return set_error(name, cudaGetErrorString(e));
  return 0;
}

inline int post_cl_clusters(int N) { return 2 * ((N + kPostR - 1) / kPostR); }
inline int dec_cl_clusters(int N) { return (N + kDecR - 1) / kDecR; }

// 0 = cooperative grid chains of recurrent.cuh, 1 = cluster chains (default).  ACVAE_CHAIN_IMPL=coop is a profiling switch
// (A/B timing of the two exchange mechanisms); the product default never reads anything else.
inline bool cluster_chain_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ACVAE_CHAIN_IMPL"); v = (e && e[0] == 'c' && e[1] == 'o') ? 0 : 1; }
  return v == 1;
}
// Can the cluster chains run this problem?  (E == A == 256; the decoder's clip tiles fit in shared memory; the device
// schedules clusters of 8 with that much shared memory.)
inline bool cluster_chain_supported(int N, int T, int Te, int E, int A) {
  if (!cluster_chain_enabled() || E != kClE || A != kClE || T < 1 || N < 1 || Te > 96) return false;
  static int ok_dev[kMaxDevices], optin_dev[kMaxDevices];
  static bool probed[kMaxDevices] = {false};
  const int cur = current_device();
  if (!probed[cur]) {
    probed[cur] = true; ok_dev[cur] = 0; optin_dev[cur] = 0;
    int cc = 0, optin = 0;
    if (cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, cur) == cudaSuccess && cc >= 9 &&
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cur) == cudaSuccess &&
        cudaFuncSetAttribute(dec_cl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess &&
        cudaFuncSetAttribute(dec_cl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess) {
      ok_dev[cur] = 1; optin_dev[cur] = optin;
    }
  }
  const size_t need = dec_cl_fwd_smem(Te) > dec_cl_bwd_smem(Te) ? dec_cl_fwd_smem(Te) : dec_cl_bwd_smem(Te);
  return ok_dev[cur] == 1 && need <= (size_t)optin_dev[cur];
}

}  // namespace acvae
