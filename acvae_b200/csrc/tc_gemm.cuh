// tcgen05 / TMEM / TMA GEMM for the batched (non-recurrent) contractions of the step, fp32-accurate.
//
//   C[M, NC] = sum_seg  A_s[M, K_s] . B_s[NC, K_s]^T        (same GemmParams / epilogues as gemm.cuh)
//
// Tensor cores have no fp32 input format; kind::tf32 keeps 10 mantissa bits, which would break the
// 1e-4 parity bar.  Every operand word x is therefore split in shared memory into
//     hi = x truncated to tf32 (10 explicit mantissa bits) = what the tensor core reads from the raw fp32 word
//     lo = (x - hi) rounded to tf32                   (x - hi is exact in fp32, |lo| < 2^-10 |x|)
// and each 128x128x8 step issues three MMAs  hi*hi + lo*hi + hi*lo  (error ~2^-21 relative, i.e. fp32
// grade; the dropped lo*lo term is ~2^-20 smaller than hi*hi).  The split is position-wise, so it is
// oblivious to the 128-byte swizzle TMA wrote the tile with.
//
// The tensor core's fp32 accumulation truncates, and the bias grows linearly with the number of chained
// MMAs (measured 7e-9 * K relative on B200, profiles/prec_probe.py).  The accumulator therefore lives in
// TMEM only for kTcChunk k-blocks (64 values of k); two TMEM buffers alternate, and four accumulator
// warps drain each finished chunk with tcgen05.ld and add it to fp32 registers with round-to-nearest
// while the next chunk is being multiplied.
//
// Pipeline (one 128x128 output tile per CTA, 512 threads):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor.2d into a 3-stage ring, mbarrier complete_tx
//   warps 4..7  : split workers -- hi/lo split of the landed stage
//   warps 8..15 : accumulators  -- drain TMEM chunk by chunk into registers (warp w: lane quadrant w%4,
//                                  column half (w-8)/4), stage the tile for the epilogue
//   warp 1      : MMA issuer    -- one elected lane, tcgen05.mma.cta_group::1.kind::tf32, accumulator
//                                  in TMEM (128 lanes x 128 fp32 columns), tcgen05.commit frees the stage
//   all warps   : fused epilogue of gemm.cuh on the staged tile (bias / accumulate / vocab statistics /
//                 label-smoothed CE gradient), so logits are still never written to HBM.
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); both are loaded with
// 128B-swizzled TMA boxes and described to the MMA with the matching canonical layouts, so forward
// (x.W^T), backward-data (dY.W) and backward-weight (dY^T.X) GEMMs need no transposed copies.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "gemm.cuh"

namespace acvae {

constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32, kTcStages = 3;
constexpr int kTcChunk = 2;      // k-blocks accumulated inside TMEM before the partial sum is drained to fp32 registers
constexpr int kTcThreads = 512;  // warps 0-3: TMA / MMA / idle, 4-7: hi-lo split workers, 8-15: accumulator warps
constexpr int kTcTileBytes = kTcBM * kTcBK * 4;                 // 16 KB (A or B, hi or lo)
constexpr int kTcStageBytes = 4 * kTcTileBytes;                 // A_hi, A_lo, B_hi, B_lo
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kTcLdCs = kTcBN + 4;   // staging row stride: 16-byte aligned rows, conflict-free 128-bit stores by row-per-lane warps
static_assert(kTcBM * kTcLdCs * 4 <= kTcStages * kTcStageBytes, "staging tile must fit in the operand ring");

struct TcSeg {
  int a_mn_major, b_mn_major;      // 0: K-major (row-major [rows,K]); 1: MN-major (row-major [K,rows])
  int K;
  int a_row_shift, b_row_shift;    // added to the K-row coordinate of an MN-major operand (shifted recurrent inputs)
  int k_zero_period, k_zero_rem;   // K rows with (k % period == rem) contribute nothing
};
struct TcParams {
  int nseg;
  TcSeg seg[2];
  long long* trace;    // optional clock64 stamps of CTA (0,0,0) (profiles/ubench_gemm_trace.cu) or NULL
  int single_pass;     // reduced precision (acvae_set_precision(1)): ONE kind::tf32 MMA per k-step on the raw fp32 words,
                       // no hi/lo split pass (10-bit mantissa products, fp32 accumulation; within the 2e-2 bf16 tolerance)
};
// process-wide arithmetic mode of the batched contractions: 0 = fp32-grade 3xTF32 (default), 1 = single-pass TF32
inline int& tc_precision_mode() { static int m = 0; return m; }
// Split-K policy of the calling thread: 0 = default (latency: as many K splits as fill the machine, >= 2 k-blocks
// each); n > 0 = at least n k-blocks per CTA (throughput: the fixed prologue / epilogue of a CTA, ~3 us, is amortised
// over more k-blocks -- used for the weight-gradient GEMMs that only fill idle SMs next to the persistent chains).
inline int& tc_min_kblk_override() { thread_local int v = 0; return v; }
struct TcThroughputScope {
  int saved;
  explicit TcThroughputScope(int min_kblk) : saved(tc_min_kblk_override()) { tc_min_kblk_override() = min_kblk; }
  ~TcThroughputScope() { tc_min_kblk_override() = saved; }
};
// process-wide trace destination picked up by try_launch_tc (micro-benchmark only)
inline long long*& tc_trace_ptr() { static long long* p = nullptr; return p; }
#define TC_STAMP(slot) do { if (tr) tp.trace[slot] = clock64(); } while (0)

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (spin > (1ll << 24)) __trap();   // a protocol bug must abort, never hang the GPU
  }
}
// Same wait for the roles that are off the critical path (accumulator warps): sleep between polls so the spin
// does not compete for issue slots with the MMA / TMA issuing warps on the same scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done) __nanosleep(64);
    if (spin > (1ll << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// One lane of a converged warp (elect.sync).  The issuing loops below are executed by ALL lanes of their warp and
// only the tcgen05 / TMA instruction itself is predicated on the elected lane: operands are then provably
// warp-uniform and live in uniform registers.  Issuing from inside `if (lane == 0)` makes ptxas wrap every
// UTCHMMA / UTMALDG in an R2UR + BRA.U.ANY waterfall loop, and the single issuing thread becomes the bottleneck of
// the whole pipeline (measured: 1.3 us per k-block, independent of the number of MMAs; profiles/ubench_tc.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
// accumulate = compile-time constant (no predicate register traffic in the issue loop)
template <int ACC>
__device__ __forceinline__ void tc_mma_tf32_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tc_ld4_nowait(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// x = hi + lo.  kind::tf32 reads 32-bit words and ignores the 13 low mantissa bits, so the RAW fp32 word already
// acts as hi = trunc_tf32(x) and needs no rewrite; only lo = x - trunc_tf32(x) (exact in fp32, same sign,
// |lo| < 2^-10 |x|) is written, rounded to nearest tf32 so that its own representation error (2^-21 |x|) is
// unbiased.  Halves the shared-memory store traffic of the split pass (the pipeline is shared-memory bound).
__device__ __forceinline__ uint32_t tc_lo(uint32_t x) {
  const float r = __uint_as_float(x) - __uint_as_float(x & 0xffffe000u);
  return (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}

// UMMA shared-memory descriptor, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;  // 2 = SWIZZLE_128B (16-byte atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms)
  return d;
}

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ TcParams tp,
               const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1) {
  if (p.live && *p.live == 0) return;   // sampling early stop: uniform over the grid, before any barrier / TMEM allocation
  extern __shared__ uint8_t tc_smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (keeps the shared address space visible to the
  // compiler: the split workers' accesses become LDS/STS instead of generic LD/ST)
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* full = bars;                       // TMA bytes landed           (count 1 + tx)
  uint64_t* ready = bars + kTcStages;          // hi/lo split done           (count 4: one per worker warp)
  uint64_t* empty = bars + 2 * kTcStages;      // MMAs reading the stage retired (tcgen05.commit)
  uint64_t* cfull = bars + 3 * kTcStages;      // [2] TMEM chunk buffer complete (tcgen05.commit)
  uint64_t* cempty = cfull + 2;                // [2] TMEM chunk buffer drained  (count 8: one per accumulator warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + 2);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kTcBM, c0 = blockIdx.x * kTcBN;
  const bool tr = tp.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
  if (tid == 0) TC_STAMP(0);

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&cfull[b], 1); mbar_init(&cempty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (wid == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  if (tid == 0) TC_STAMP(1);

  const int nblk0 = (tp.seg[0].K + kTcBK - 1) / kTcBK;
  const int nblk1 = tp.nseg > 1 ? (tp.seg[1].K + kTcBK - 1) / kTcBK : 0;
  // split-K: grid.z CTAs share one output tile, each owns a contiguous range of k-blocks and adds its partial
  // tile with atomics (few-tile, long-K shapes such as dH = dlogits.W would otherwise use 10 of 148 SMs)
  const int per_split = (nblk0 + nblk1 + (int)gridDim.z - 1) / (int)gridDim.z;
  const int i0 = (int)blockIdx.z * per_split;
  const int total = min(nblk0 + nblk1, i0 + per_split) - i0;      // k-blocks of this CTA (>= 1 by construction)

  float acc_reg[kTcBN / 2];   // only the accumulator warps (8..15) touch it
  if (wid == 0) {
    // ===== TMA producer: the whole warp runs the loop, one elected lane issues =====
    int st = 0, par = 0;                        // parity of the `empty` phase to wait for (first pass: no wait)
    for (int i = 0; i < total; ++i) {
      if (i >= kTcStages) mbar_wait(&empty[st], par);
      const bool s1 = i0 + i >= nblk0;
      const TcSeg& sg = tp.seg[s1 ? 1 : 0];
      const int kb = (s1 ? i0 + i - nblk0 : i0 + i) * kTcBK;
      const CUtensorMap* ma = s1 ? &mapA1 : &mapA0;
      const CUtensorMap* mb = s1 ? &mapB1 : &mapB0;
      uint8_t* sa = smem + st * kTcStageBytes;
      uint8_t* sb = sa + 2 * kTcTileBytes;
      if (elect_one()) {
        mbar_expect_tx(&full[st], 2 * kTcTileBytes);
        if (!sg.a_mn_major) tma_load_2d(sa, ma, &full[st], kb, m0);                      // box {32 k, 128 rows}
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_load_2d(sa + q * 4096, ma, &full[st], m0 + q * 32, kb + sg.a_row_shift);  // box {32 m, 32 k}
        }
        if (!sg.b_mn_major) tma_load_2d(sb, mb, &full[st], kb, c0);
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_load_2d(sb + q * 4096, mb, &full[st], c0 + q * 32, kb + sg.b_row_shift);
        }
      }
      __syncwarp();
      if (++st == kTcStages) { st = 0; if (i >= kTcStages) par ^= 1; }
    }
  } else if (wid == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
    // Per-segment descriptor templates (everything but the 14-bit start-address field) and k-step strides.
    // K-major: 8-row groups 1024 B apart (SBO), LBO unused.  MN-major tf32 must use the 32-byte-atom swizzle
    // (cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem layout"):
    // 32-element MN chunks 4096 B apart (LBO), 4-row K groups 512 B apart (SBO).
    uint64_t a_tmpl[2], b_tmpl[2];
    uint32_t a_step[2], b_step[2], idesc[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const TcSeg& sg = tp.seg[s];
      a_tmpl[s] = tc_smem_desc(0, sg.a_mn_major ? 4096u : 16u, sg.a_mn_major ? 512u : 1024u, sg.a_mn_major ? 1u : 2u);
      b_tmpl[s] = tc_smem_desc(0, sg.b_mn_major ? 4096u : 16u, sg.b_mn_major ? 512u : 1024u, sg.b_mn_major ? 1u : 2u);
      a_step[s] = (sg.a_mn_major ? 1024u : 32u) >> 4;
      b_step[s] = (sg.b_mn_major ? 1024u : 32u) >> 4;
      idesc[s] = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)sg.a_mn_major << 15) | ((uint32_t)sg.b_mn_major << 16) |
                 ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
    }
    const uint32_t ring = smem_u32(smem) >> 4;
    int st = 0, par = 0, in_chunk = 0, chunk = 0;
    for (int i = 0; i < total; ++i) {
      const int cb = chunk & 1;
      if (in_chunk == 0 && chunk >= 2) mbar_wait(&cempty[cb], ((chunk >> 1) - 1) & 1);
      mbar_wait(&ready[st], par);
      tc_fence_after();
      if (elect_one()) {
        const int s = i0 + i >= nblk0 ? 1 : 0;
        const uint64_t dah = a_tmpl[s] + (uint64_t)(ring + (uint32_t)(st * (kTcStageBytes >> 4)));
        const uint64_t dal = dah + (kTcTileBytes >> 4);
        const uint64_t dbh = b_tmpl[s] + (uint64_t)(ring + (uint32_t)(st * (kTcStageBytes >> 4)) + 2 * (kTcTileBytes >> 4));
        const uint64_t dbl = dbh + (kTcTileBytes >> 4);
        const uint32_t tmem_c = tmem_d + (uint32_t)cb * kTcBN;
        const uint32_t as = a_step[s], bs = b_step[s], id = idesc[s];
        if (tp.single_pass) {
          if (in_chunk == 0) tc_mma_tf32_c<0>(tmem_c, dah, dbh, id);
          else tc_mma_tf32_c<1>(tmem_c, dah, dbh, id);
#pragma unroll
          for (int j = 1; j < kTcBK / 8; ++j) tc_mma_tf32_c<1>(tmem_c, dah + j * as, dbh + j * bs, id);
        } else {
          // small terms first, the dominant hi*hi product last
          if (in_chunk == 0) tc_mma_tf32_c<0>(tmem_c, dal, dbh, id);
          else tc_mma_tf32_c<1>(tmem_c, dal, dbh, id);
          tc_mma_tf32_c<1>(tmem_c, dah, dbl, id);
          tc_mma_tf32_c<1>(tmem_c, dah, dbh, id);
#pragma unroll
          for (int j = 1; j < kTcBK / 8; ++j) {
            tc_mma_tf32_c<1>(tmem_c, dal + j * as, dbh + j * bs, id);
            tc_mma_tf32_c<1>(tmem_c, dah + j * as, dbl + j * bs, id);
            tc_mma_tf32_c<1>(tmem_c, dah + j * as, dbh + j * bs, id);
          }
        }
        tc_commit(&empty[st]);
        if (in_chunk == kTcChunk - 1 || i == total - 1) tc_commit(&cfull[cb]);
      }
      __syncwarp();
      if (++st == kTcStages) { st = 0; par ^= 1; }
      if (++in_chunk == kTcChunk) { in_chunk = 0; ++chunk; }
    }
  } else if (wid >= 8) {
// ===== accumulator warps: drain finished TMEM chunks into fp32 registers (round-to-nearest adds) =====
    const int q = wid & 3;                                     // TMEM lane quadrant this warp may access
    const int ch = (wid - 8) >> 2;                             // which 64-column half it owns
    const int nchunks = (total + kTcChunk - 1) / kTcChunk;
#pragma unroll
    for (int j = 0; j < kTcBN / 2; ++j) acc_reg[j] = 0.0f;
    for (int c = 0; c < nchunks; ++c) {
      const int cb = c & 1;
      mbar_wait_backoff(&cfull[cb], (c >> 1) & 1);
      if (wid == 8 && c == nchunks - 1) TC_STAMP(4);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < kTcBN / 2; cc += 32) {
        uint32_t v[32];
        tc_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * kTcBN + ch * (kTcBN / 2) + cc), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc_reg[cc + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&cempty[cb]);
      if (wid == 8 && c == nchunks - 1) TC_STAMP(5);
    }
  } else if (wid >= 4) {
    // ===== split workers: x -> (hi, lo), 128 threads over 2 x 4096 words =====
    const int wt = tid - 128;
    int st = 0, par = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[st], par);
      if (wt == 0 && i == 0) TC_STAMP(2);
      const bool s1 = i0 + i >= nblk0;
      const TcSeg& sg = tp.seg[s1 ? 1 : 0];
      const int kb = (s1 ? i0 + i - nblk0 : i0 + i) * kTcBK;
      uint4* hiA = reinterpret_cast<uint4*>(smem + st * kTcStageBytes);
      uint4* loA = hiA + kTcTileBytes / 16;
      uint4* hiB = loA + kTcTileBytes / 16;
      uint4* loB = hiB + kTcTileBytes / 16;
      const bool kmask = sg.k_zero_period > 0 && sg.b_mn_major;
      if (!tp.single_pass) {
#pragma unroll
        for (int q0 = 0; q0 < kTcTileBytes / 16; q0 += 128 * 4) {
          uint4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = hiA[q0 + u * 128 + wt];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            loA[q0 + u * 128 + wt] = make_uint4(tc_lo(x[u].x), tc_lo(x[u].y), tc_lo(x[u].z), tc_lo(x[u].w));
        }
      }
      if (tp.single_pass && kmask) {                   // only the masked K rows of B have to be cleared
        for (int q = wt; q < kTcTileBytes / 16; q += 128) {
          const int krow = kb + ((q >> 3) & 31);
          if (krow % sg.k_zero_period == sg.k_zero_rem) hiB[q] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int q0 = 0; !tp.single_pass && q0 < kTcTileBytes / 16; q0 += 128 * 4) {
        uint4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = hiB[q0 + u * 128 + wt];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u * 128 + wt;
          if (kmask) {
            const int krow = kb + ((q >> 3) & 31);            // MN-major tile: 128-byte row r of each 4 KB slab is K row r
            if (krow % sg.k_zero_period == sg.k_zero_rem) { x[u] = make_uint4(0u, 0u, 0u, 0u); hiB[q] = x[u]; }
          }
          loB[q] = make_uint4(tc_lo(x[u].x), tc_lo(x[u].y), tc_lo(x[u].z), tc_lo(x[u].w));
        }
      }
      fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (wt == 0 && i == 0) TC_STAMP(3);
      if (lane == 0) mbar_arrive(&ready[st]);
      if (++st == kTcStages) { st = 0; par ^= 1; }
    }
  }
  // every role has passed its last use of the operand ring before the staging tile overwrites it
  __syncthreads();
  if (wid >= 8) {
    float* Cs = reinterpret_cast<float*>(smem);               // [128][132] staging tile over the (now idle) operand ring
    const int row = (wid & 3) * 32 + lane, col0 = ((wid - 8) >> 2) * (kTcBN / 2);
#pragma unroll
    for (int j = 0; j < kTcBN / 2; j += 4)
      *reinterpret_cast<float4*>(Cs + row * kTcLdCs + col0 + j) = make_float4(acc_reg[j], acc_reg[j + 1], acc_reg[j + 2], acc_reg[j + 3]);
  }
  __syncthreads();
  if (tid == 0) TC_STAMP(6);
  if constexpr (EPI == EPI_STATS) {   // row-per-warp epilogue: all 16 warps
    gemm_epilogue<EPI, kTcBM, kTcBN, kTcThreads>(p, reinterpret_cast<const float*>(smem), kTcLdCs, m0, c0, tid, blockIdx.x, gridDim.x);
  } else if constexpr (EPI == EPI_PLAIN) {   // all 16 warps
    gemm_epilogue<EPI, kTcBM, kTcBN, kTcThreads>(p, reinterpret_cast<const float*>(smem), kTcLdCs, m0, c0, tid, blockIdx.x, gridDim.x);
  } else if (tid < 256) {             // the other element-wise epilogues stride by 256 threads
    gemm_epilogue<EPI, kTcBM, kTcBN>(p, reinterpret_cast<const float*>(smem), kTcLdCs, m0, c0, tid, blockIdx.x, gridDim.x);
  }
  __syncthreads();
  if (tid == 0) TC_STAMP(7);
  if (wid == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(256));
    TC_STAMP(8);
  }
}

// =====================================================================================================
// Persistent variant for the forward GEMMs with >= 75 output tiles (sampling: M = 10 450 sequences, 656 .. 2870 tiles).
// One CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the operand ring, the split workers and the MMA
// issuer run straight through the tile boundaries, and the accumulator lives in TMEM for the WHOLE K range of a tile
// (K <= 512 on this path: the tensor core's truncating accumulation costs 7e-9 * K relative, profiles/prec_probe.py),
// in one of FOUR 128-column buffers (all 512 TMEM columns).  When a tile's last MMA retires, the eight epilogue warps
// work on it (thread = one output row x 64 columns) while the next tiles are being multiplied -- the one-tile kernel above
// pays prologue + main loop + epilogue serially per tile (23 us per vocabulary tile of which 6 are MMAs).
//   EPI_PLAIN: the 64 values go to registers, the buffer is handed back at once, the stores follow.
//   EPI_STATS: row statistics are thread-local in this layout (no staging tile, no shuffles): a SMALL loop walks the row
//   eight columns at a time straight from TMEM (an unrolled 64-column body with its 32 inlined Philox calls does not fit
//   the instruction cache: ncu stall_no_instruction 6.1 per issue, profiles/r2), keeps an online max / sum-exp, and each
//   thread emits the partial of a 64-column half tile (same partial layout as the 64-wide SIMT tiles).
// =====================================================================================================
// EPI_STATS runs TWO epilogue groups of eight warps (24 warps, 80 registers): with the sampling noise drawn in place the
// epilogue of a tile (Philox + two logarithms + exp per logit: ~18 us) is twice its main loop (8.7 us); the groups take
// alternate tiles (buffers 0, 2 / 1, 3), so a tile leaves the SM every ~9 us.  EPI_PLAIN keeps one group (128 registers).
template <int EPI> __host__ __device__ constexpr int tc_persist_threads() { return EPI == EPI_STATS ? 768 : kTcThreads; }
template <int EPI>
__global__ void __launch_bounds__(tc_persist_threads<EPI>(), 1)
tc_gemm_persist_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ TcParams tp,
                       const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
                       const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
                       int tiles_n, int total_tiles) {
  if (p.live && *p.live == 0) return;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* full = bars;
  uint64_t* ready = bars + kTcStages;
  uint64_t* empty = bars + 2 * kTcStages;
  uint64_t* afull = bars + 3 * kTcStages;      // [4] accumulator buffer complete (tcgen05.commit)
  uint64_t* aempty = afull + 4;                // [4] accumulator buffer consumed (count 8: one per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 4);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 4; ++b) { mbar_init(&afull[b], 1); mbar_init(&aempty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (wid == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int nblk0 = (tp.seg[0].K + kTcBK - 1) / kTcBK;
  const int nblk1 = tp.nseg > 1 ? (tp.seg[1].K + kTcBK - 1) / kTcBK : 0;
  const int total = nblk0 + nblk1;             // k-blocks per tile

  if (wid == 0) {
    // ===== TMA producer =====
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kTcBM, c0 = (tile % tiles_n) * kTcBN;
      for (int i = 0; i < total; ++i, ++it) {
        const int st = it % kTcStages;
        if (it >= kTcStages) mbar_wait(&empty[st], (it / kTcStages - 1) & 1);
        const bool s1 = i >= nblk0;
        const int kb = (s1 ? i - nblk0 : i) * kTcBK;
        const CUtensorMap* ma = s1 ? &mapA1 : &mapA0;
        const CUtensorMap* mb = s1 ? &mapB1 : &mapB0;
        uint8_t* sa = smem + st * kTcStageBytes;
        uint8_t* sb = sa + 2 * kTcTileBytes;
        if (elect_one()) {
          mbar_expect_tx(&full[st], 2 * kTcTileBytes);
          tma_load_2d(sa, ma, &full[st], kb, m0);
          tma_load_2d(sb, mb, &full[st], kb, c0);
        }
        __syncwarp();
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer =====
    const uint64_t tmpl = tc_smem_desc(0, 16u, 1024u, 2u);          // K-major operands only on this path
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
    const uint32_t ring = smem_u32(smem) >> 4;
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 3;
      if (lt >= 4) mbar_wait(&aempty[buf], ((lt >> 2) - 1) & 1);
      tc_fence_after();
      const uint32_t tmem_c = tmem_d + (uint32_t)buf * kTcBN;
      for (int i = 0; i < total; ++i, ++it) {
        const int st = it % kTcStages;
        mbar_wait(&ready[st], (it / kTcStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dah = tmpl + (uint64_t)(ring + (uint32_t)(st * (kTcStageBytes >> 4)));
          const uint64_t dal = dah + (kTcTileBytes >> 4);
          const uint64_t dbh = dah + 2 * (kTcTileBytes >> 4);
          const uint64_t dbl = dbh + (kTcTileBytes >> 4);
          if (tp.single_pass) {
            tc_mma_tf32(tmem_c, dah, dbh, idesc, i > 0 ? 1u : 0u);
#pragma unroll
            for (int j = 1; j < kTcBK / 8; ++j) tc_mma_tf32_c<1>(tmem_c, dah + j * 2, dbh + j * 2, idesc);
          } else {
            tc_mma_tf32(tmem_c, dal, dbh, idesc, i > 0 ? 1u : 0u);
            tc_mma_tf32_c<1>(tmem_c, dah, dbl, idesc);
            tc_mma_tf32_c<1>(tmem_c, dah, dbh, idesc);
#pragma unroll
            for (int j = 1; j < kTcBK / 8; ++j) {
              tc_mma_tf32_c<1>(tmem_c, dal + j * 2, dbh + j * 2, idesc);
              tc_mma_tf32_c<1>(tmem_c, dah + j * 2, dbl + j * 2, idesc);
              tc_mma_tf32_c<1>(tmem_c, dah + j * 2, dbh + j * 2, idesc);
            }
          }
          tc_commit(&empty[st]);
          if (i == total - 1) tc_commit(&afull[buf]);
        }
        __syncwarp();
      }
    }
  } else if (wid >= 8) {
    // ===== epilogue warps =====
    const int q = wid & 3, ch = ((wid - 8) >> 2) & 1, grp = (wid - 8) >> 3;
    constexpr int kGroups = tc_persist_threads<EPI>() / 256 - 1;      // 1 or 2 groups of eight warps
    const EpiParams& ep = p.epi;
    const int U = p.U;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      if (kGroups == 2 && (lt & 1) != grp) continue;                  // the other group's tile
      const int buf = lt & 3;
      const int tile_x = tile % tiles_n;
      const int m0 = (tile / tiles_n) * kTcBM, c0 = tile_x * kTcBN + ch * (kTcBN / 2);
      const int gm = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kTcBN + ch * (kTcBN / 2));
      mbar_wait_backoff(&afull[buf], (lt >> 2) & 1);
      tc_fence_after();
      if constexpr (EPI == EPI_PLAIN) {
        float acc[kTcBN / 2];
#pragma unroll
        for (int cc = 0; cc < kTcBN / 2; cc += 32) {
          uint32_t v[32];
          tc_ld32(taddr + cc, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cc + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aempty[buf]);          // the values are in registers: the buffer is free again
        if (gm >= p.M || c0 >= U) continue;
        float* __restrict__ dst = ep.c[0] + (long long)gm * ep.ldc + c0;
        const float* __restrict__ bias = ep.bias[0];
        const bool vec = (ep.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.c[0]) & 15) == 0 && c0 + kTcBN / 2 <= U;
        if (vec) {
#pragma unroll
          for (int j = 0; j < kTcBN / 2; j += 4) {
            float4 v = make_float4(acc[j] * ep.scale, acc[j + 1] * ep.scale, acc[j + 2] * ep.scale, acc[j + 3] * ep.scale);
            if (bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (ep.accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(dst + j);
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *reinterpret_cast<float4*>(dst + j) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < kTcBN / 2; ++j) {
            if (c0 + j < U) {
              float v = acc[j] * ep.scale;
              if (bias) v += __ldg(bias + c0 + j);
              if (ep.accumulate) v += dst[j];
              dst[j] = v;
            }
          }
        }
      } else if constexpr (EPI == EPI_STATS) {
        float m = -INFINITY, sexp = 0.0f, vsum = 0.0f, best = -INFINITY, bestlogit = 0.0f;
        int barg = 0x7fffffff;
        const bool row_ok = gm < p.M && c0 < U;
        const bool noisy = ep.noise || ep.rng;
        const float* __restrict__ nrow = ep.noise ? ep.noise + (long long)gm * ep.ld_noise + c0 : nullptr;
        const float* __restrict__ bias = ep.bias[0];
#pragma unroll 1
        for (int g0 = 0; g0 < 16; g0 += 4) {
          // sixteen columns per pass: g0 + k + 16 h (k, h = 0..3).  Philox group (tile, half, g0 + k) yields the four
          // variates of columns g0 + k + {0, 16, 32, 48} (gemm.cuh philox_uniform4): every output word is used.
          uint32_t a[4][4];
          __syncwarp();                                // the TMEM loads are warp-collective (.sync.aligned)
#pragma unroll
          for (int h = 0; h < 4; ++h) tc_ld4_nowait(taddr + 16 * h + g0, a[h]);
          tc_ld_wait();
          if (row_ok) {
          float v[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float uu[4];
            if (ep.rng) philox_uniform4(ep.rng, ep.rng_step, gm, ((c0 >> 7) << 5) + ch * 16 + g0 + k, uu);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int jj = g0 + k + 16 * h, u = c0 + jj, e = k * 4 + h;
              if (u < U) {
                v[e] = __uint_as_float(a[h][k]) + (bias ? __ldg(bias + u) : 0.0f);
                float key = v[e];
                if (ep.rng) key = v[e] * ep.inv_temp + gumbel_from_u_fast(uu[h]);
                else if (noisy) key = v[e] * ep.inv_temp + (ep.noise_is_gumbel ? nrow[jj] : gumbel_from_u(nrow[jj]));
                vsum += v[e];
                if (key > best || (key == best && u < barg)) { best = key; barg = u; bestlogit = v[e]; }
              } else {
                v[e] = -INFINITY;
              }
            }
          }
          // online log-sum-exp: one rescale per sixteen columns (column c0 is always valid, so the maximum is finite)
          float gmax = v[0];
#pragma unroll
          for (int e = 1; e < 16; ++e) gmax = fmaxf(gmax, v[e]);
          const float nm = fmaxf(m, gmax);
          float add = 0.0f;
#pragma unroll
          for (int e = 0; e < 16; ++e) add += expf(v[e] - nm);
          sexp = sexp * expf(m - nm) + add;
          m = nm;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aempty[buf]);
        if (row_ok) {
          const int ntiles = (U + kTcBN / 2 - 1) / (kTcBN / 2);
          const long long o = (long long)gm * ntiles + tile_x * 2 + ch;
          ep.pmax[o] = m; ep.pexp[o] = sexp; ep.psum[o] = vsum; ep.pbest[o * 2] = best; ep.pbest[o * 2 + 1] = bestlogit;
          ep.parg[o] = barg;
        }
      }
    }
  } else if (wid >= 4) {
    // ===== split workers =====
    const int wt = tid - 128;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int i = 0; i < total; ++i, ++it) {
        const int st = it % kTcStages;
        mbar_wait(&full[st], (it / kTcStages) & 1);
        if (!tp.single_pass) {
          uint4* hiA = reinterpret_cast<uint4*>(smem + st * kTcStageBytes);
          uint4* loA = hiA + kTcTileBytes / 16;
          uint4* hiB = loA + kTcTileBytes / 16;
          uint4* loB = hiB + kTcTileBytes / 16;
#pragma unroll
          for (int q0 = 0; q0 < kTcTileBytes / 16; q0 += 128 * 4) {
            uint4 x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { x[u] = hiA[q0 + u * 128 + wt]; y[u] = hiB[q0 + u * 128 + wt]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              loA[q0 + u * 128 + wt] = make_uint4(tc_lo(x[u].x), tc_lo(x[u].y), tc_lo(x[u].z), tc_lo(x[u].w));
              loB[q0 + u * 128 + wt] = make_uint4(tc_lo(y[u].x), tc_lo(y[u].y), tc_lo(y[u].z), tc_lo(y[u].w));
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[st]);
      }
    }
  }
  __syncthreads();
  if (wid == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(512));
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] (cols contiguous, leading dimension ld elements), box {32 cols, box_rows}, 128B swizzle
inline bool tc_make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows,
                        bool atom32 = false) {
  PFN_encodeTiled fn = tc_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ACVAE_DISABLE_TC");
    v = (e && e[0] == '1') ? 0 : (tc_encode_fn() ? 1 : 0);
  }
  return v == 1;
}

// Extra description of shifted operands: the tensor map is built on `base` (the real allocation) and the
// K-row coordinate is shifted, so rows outside [0, R) are zero-filled by TMA instead of being read.
struct TcShift { int a_shift = 0, b_shift = 0; };

// Returns 1 if the GEMM was issued on the tensor cores, 0 if the caller must use the SIMT kernels, <0 on error.
template <int EPI>
inline int try_launch_tc(const GemmParams& p, cudaStream_t st) {
  if (!tc_enabled()) return 0;
  if (p.G != 1 || p.nseg < 1 || p.nseg > 2 || p.M < 96) return 0;
  const int NC = p.U;
  TcParams tp{};
  tp.nseg = p.nseg;
  tp.trace = tc_trace_ptr();
  tp.single_pass = tc_precision_mode() == 1;
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& sg = p.seg[s];
    if (sg.gather || sg.K < 8) return 0;
    if (sg.lda % 4 || sg.ldw % 4 || !aligned16(sg.a) || !aligned16(sg.w[0])) return 0;
    if (sg.k_zero_period && !(sg.a_trans && sg.w_trans)) return 0;
    TcSeg& t = tp.seg[s];
    t.K = sg.K; t.a_mn_major = sg.a_trans; t.b_mn_major = sg.w_trans;
    t.k_zero_period = sg.k_zero_period; t.k_zero_rem = sg.k_zero_rem;
    t.a_row_shift = 0; t.b_row_shift = sg.w_row_shift;
    const float* wbase = sg.w[0] - (long long)sg.w_row_shift * sg.ldw;   // undo the pointer shift: map the real tensor
    bool ok = true;
    if (!sg.a_trans) ok = ok && tc_make_map(&maps[2 * s], sg.a, p.M, sg.K, sg.lda, kTcBM);
    else ok = ok && tc_make_map(&maps[2 * s], sg.a, sg.K, p.M, sg.lda, kTcBK, true);
    if (!sg.w_trans) ok = ok && tc_make_map(&maps[2 * s + 1], sg.w[0], NC, sg.K, sg.ldw, kTcBN);
    else ok = ok && tc_make_map(&maps[2 * s + 1], wbase, sg.K, NC, sg.ldw, kTcBK, true);
    if (!ok) return 0;
  }
  if (p.nseg == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }
  static bool configured_dev[kMaxDevices] = {false};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    ACVAE_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    configured = true;
  }
  dim3 grid((NC + kTcBN - 1) / kTcBN, (p.M + kTcBM - 1) / kTcBM);
  GemmParams pl = p;
  if constexpr (EPI == EPI_PLAIN || EPI == EPI_STATS) {
    // many-tile forward GEMMs: the persistent kernel (accumulator in TMEM for the whole K range, epilogue overlapped)
    static int persist = -1, min_tiles = 75;   // measured: sampling at 1310 sequences 3.74 -> 3.63 ms from 149 -> 75, nothing below; the train step is indifferent
    if (persist < 0) {
      const char* e = getenv("ACVAE_TC_PERSIST"); persist = (e && e[0] == '0') ? 0 : 1;
      const char* m = getenv("ACVAE_TC_PERSIST_MIN_TILES"); if (m) min_tiles = atoi(m);
    }
    int nblk = 0;
    bool kmajor = true;
    for (int s = 0; s < p.nseg; ++s) {
      nblk += (p.seg[s].K + kTcBK - 1) / kTcBK;
      kmajor = kmajor && !p.seg[s].a_trans && !p.seg[s].w_trans && !p.seg[s].k_zero_period;
    }
    const int tiles = (int)(grid.x * grid.y);
    if (persist && kmajor && !tp.trace && !p.epi.atomic && nblk <= 16 && tiles >= min_tiles) {
      static bool configured_p[kMaxDevices] = {false};
      bool& cfg = configured_p[current_device()];
      if (!cfg) {
        ACVAE_CHECK(cudaFuncSetAttribute(tc_gemm_persist_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        cfg = true;
      }
      ACVAE_LAUNCH((tc_gemm_persist_kernel<EPI>), dim3(tiles < 148 ? tiles : 148), tc_persist_threads<EPI>(), kTcSmemBytes, st, pl, tp,
                   maps[0], maps[1], maps[2], maps[3], (int)grid.x, tiles);
      return EPI == EPI_STATS ? 2 : 1;
    }
  }
  if constexpr (EPI == EPI_PLAIN) {
    int nblk = 0;
    for (int s = 0; s < p.nseg; ++s) nblk += (p.seg[s].K + kTcBK - 1) / kTcBK;
    const int tiles = (int)(grid.x * grid.y);
    int splits = 148 / tiles;                       // fill the machine once
    // k-blocks per CTA below which splitting stops.  Alone, a launch is fastest split down to 2 k-blocks per CTA
    // (profiles/r1/gemm_split.log: 16 -> 11 us); inside the step every split CTA owns an SM the chains and the other GEMMs are
    // short of, and the optimum moves to ~6 (profiles/r2/tc_min_kblk.log: 0.871 / 0.848 / 0.844 ms per step at 2 / 4 / 6).
    static int min_blk = 0, min_blk_bwd = 0;
    if (!min_blk) {
      const char* e = getenv("ACVAE_TC_MIN_KBLK"); min_blk = e ? atoi(e) : 6; if (min_blk < 1) min_blk = 1;
      const char* b = getenv("ACVAE_TC_MIN_KBLK_BWD"); min_blk_bwd = b ? atoi(b) : min_blk; if (min_blk_bwd < 1) min_blk_bwd = 1;
    }
    const int mb = tc_min_kblk_override() > 0 ? tc_min_kblk_override() : (p.epi.free_order ? min_blk_bwd : min_blk);
    if (splits > nblk / mb) splits = nblk / mb;
    if (!p.epi.free_order) splits = p.epi.accumulate ? 1 : (splits > 2 ? 2 : splits);   // forward / sampling: reproducible
    if (splits >= 2) {
      const int per = (nblk + splits - 1) / splits;
      splits = (nblk + per - 1) / per;              // no empty CTA
      grid.z = splits;
      pl.epi.atomic = 1;
      if (!p.epi.accumulate)
        ACVAE_CHECK(cudaMemset2DAsync(p.epi.c[0], (size_t)p.epi.ldc * sizeof(float), 0, (size_t)NC * sizeof(float), (size_t)p.M, st));
    }
  }
  ACVAE_LAUNCH((tc_gemm_kernel<EPI>), grid, kTcThreads, kTcSmemBytes, st, pl, tp, maps[0], maps[1], maps[2], maps[3]);
  return 1;
}

// Dispatcher used by the whole library: tensor cores for the batched G == 1 contractions, SIMT kernels for
// the skinny recurrent-step GEMMs and everything the TMA path cannot describe (gathers, gate interleave).
template <int EPI>
inline int launch_gemm(const GemmParams& p, cudaStream_t st, int* used_tc = nullptr) {
  if (used_tc) *used_tc = 0;
  if (p.M <= 0 || p.U <= 0) return 0;
  if constexpr (EPI == EPI_PLAIN || EPI == EPI_STATS || EPI == EPI_DLOGITS) {
    const int r = try_launch_tc<EPI>(p, st);
    if (r < 0) return r;
    if (r >= 1) { if (used_tc) *used_tc = r; return 0; }   // 2: EPI_STATS partials are per 64-column half tile
  }
  return launch_gemm_simt<EPI>(p, st);
}

}  // namespace acvae
