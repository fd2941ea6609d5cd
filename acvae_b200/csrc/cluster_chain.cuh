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
    if (tid == 0 && s + 1 < T) hop.arm(par);
    float gxr = 0.f, gxz = 0.f, gxn = 0.f;
    if (live) {
      const float* gx = p.gx[dir] + ((long long)n * T + t) * 3 * E + u;
      gxr = __ldg(gx); gxz = __ldg(gx + E); gxn = __ldg(gx + 2 * E);
    }
    float res[3] = {0.f, 0.f, 0.f};
    if (s > 0) {
      hop.wait(par ^ 1, (s - 1) >> 1);
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
    const float hn = res[2] + bh_n;
    const float rg = sigmoidf_(gxr + res[0] + bh_r), zg = sigmoidf_(gxz + res[1] + bh_z);
    const float ng = tanhf(gxn + rg * hn);
    float hnew = (1.0f - zg) * ng + zg * hprev;
    if (t >= len) hnew = 0.0f;                                 // packed sequence: padded outputs are zero
    hprev = hnew;
    if (live) {
      float* gs = p.gq[dir] + ((long long)n * T + t) * 4 * E + u;
      gs[0] = rg; gs[E] = zg; gs[2 * E] = ng; gs[3 * E] = hn;
      p.ho[((long long)n * T + t) * 2 * E + dir * E + u] = hnew;
    }
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
    if (live) {
      float* gi = p.dgi[dir] + ((long long)n * T + t) * 3 * E + u;
      float* gh = p.dgh[dir] + ((long long)n * T + t) * 3 * E + u;
      gi[0] = dar; gi[E] = daz; gi[2 * E] = dan;
      gh[0] = dar; gh[E] = daz; gh[2 * E] = dan * rr;
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
  }
  cl_sync_all();
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

}  // namespace acvae
