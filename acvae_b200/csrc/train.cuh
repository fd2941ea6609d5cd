// Host-side orchestration of the training forward / backward of the latent
// word-decoding step (reference models/vae_model.py:700-869 and its autograd).
// Everything here enqueues kernels on the caller's stream; no allocation, no
// host synchronisation.
#pragma once
#include "../../include/acvae_b200.h"
#include "attention.cuh"
#include "tc_gemm.cuh"
#include "pointwise.cuh"
#include "streams.cuh"

namespace acvae {

constexpr int kStartIdx = 1;  // models/word_model.py:20
constexpr int kVocabTile = 64;

// Workspace carve-up shared by acvae_train_fwd and acvae_train_bwd: the forward
// leaves its saved activations here, the backward consumes them.
struct TrainWs {
  // memory
  float *mem, *Pp, *Pd;
  int *words, *qids, *steplens;
  // embeddings gathered per (n,t)
  float *xq, *xp, *xd;
  // posterior
  float *gxq[2], *gq[2], *ho; int* amax_q;
  // prior
  float *qp_p, *w_p, *ctx_p, *gates_p, *c_p, *h_p;
  // decoder
  float *qp_d, *w_d, *ctx_d, *gates_d, *pool_d; int* amax_d;
  float* part_d;    // [min(N,32), 128, max(A,E)] per-CTA partial projections of the persistent decoder chains (recurrent.cuh)
  float* Mg;        // [N,Te,3E] mem . W_ih[:, E:2E]^T: per-frame context share of the decoder's gate pre-activations (cluster_chain.cuh)
  // vocab partials
  float *pmax, *pexp, *psum, *pbest; int* parg;
  // backward scratch
  float *dout, *dgi_d, *dgh_d, *dh_carry, *dctx_d, *ds_d, *dqp_d, *dxz_d, *dxe_d;
  float *dml_p, *dg_p, *dhp_carry, *dcp_carry, *dzp_carry, *dxe_p, *dctx_p, *ds_p, *dqp_p;
  float *dml_q, *dho, *dgi_q[2], *dgh_q[2], *dhq_carry, *dxq, *dzq_carry;
  float *dPp, *dPd, *dmem, *dmem2, *dpool;
  float *zsel, *dpz;  // [N,T,E] hoisted schedule with dis flags: the z each decoder step consumed; d p_z incl. the decoder's share
  unsigned* bars;   // grid-barrier counters of the persistent chain kernels (recurrent.cuh)
  size_t bytes;
};

inline TrainWs carve_train_ws(const acvae_dims& d, void* base) {
  Arena ar(base);
  TrainWs w{};
  const size_t N = d.N, T = d.T, Te = d.Te, E = d.E, A = d.A;
  const size_t NT = N * T;
  const int ntiles = (d.V + kVocabTile - 1) / kVocabTile;
  w.mem = ar.take<float>(N * Te * E); w.Pp = ar.take<float>(N * Te * E); w.Pd = ar.take<float>(N * Te * A);
  w.words = ar.take<int>(NT); w.qids = ar.take<int>(NT); w.steplens = ar.take<int>(N);
  w.xq = ar.take<float>(NT * E); w.xp = ar.take<float>(NT * E); w.xd = ar.take<float>(NT * E);
  for (int k = 0; k < 2; ++k) { w.gxq[k] = ar.take<float>(NT * 3 * E); w.gq[k] = ar.take<float>(NT * 4 * E); }
  w.ho = ar.take<float>(NT * 2 * E); w.amax_q = ar.take<int>(N * 2 * E);
  w.qp_p = ar.take<float>(NT * E); w.w_p = ar.take<float>(NT * Te); w.ctx_p = ar.take<float>(NT * E);
  w.gates_p = ar.take<float>(NT * 4 * E); w.c_p = ar.take<float>(NT * E); w.h_p = ar.take<float>(NT * E);
  w.qp_d = ar.take<float>(NT * A); w.w_d = ar.take<float>(NT * Te); w.ctx_d = ar.take<float>(NT * E);
  w.part_d = ar.take<float>((N < 32 ? N : 32) * 128 * (A > E ? A : E));
  w.Mg = ar.take<float>(Te <= 96 ? N * Te * 3 * E : 0);
  w.gates_d = ar.take<float>(NT * 4 * E); w.pool_d = ar.take<float>(N * E); w.amax_d = ar.take<int>(N * E);
  w.pmax = ar.take<float>(NT * ntiles); w.pexp = ar.take<float>(NT * ntiles); w.psum = ar.take<float>(NT * ntiles);
  w.pbest = ar.take<float>(NT * ntiles * 2); w.parg = ar.take<int>(NT * ntiles);
  // backward
  w.dout = ar.take<float>(NT * E); w.dgi_d = ar.take<float>(NT * 3 * E); w.dgh_d = ar.take<float>(NT * 3 * E);
  w.dh_carry = ar.take<float>(N * E); w.dctx_d = ar.take<float>(NT * E); w.ds_d = ar.take<float>(NT * Te);
  w.dqp_d = ar.take<float>(NT * A); w.dxz_d = ar.take<float>(NT * E); w.dxe_d = ar.take<float>(NT * E);
  w.dml_p = ar.take<float>(NT * 2 * E); w.dg_p = ar.take<float>(NT * 4 * E); w.dhp_carry = ar.take<float>(N * E);
  w.dcp_carry = ar.take<float>(N * E); w.dzp_carry = ar.take<float>(N * E); w.dxe_p = ar.take<float>(NT * E);
  w.dctx_p = ar.take<float>(NT * E); w.ds_p = ar.take<float>(NT * Te); w.dqp_p = ar.take<float>(NT * E);
  w.dml_q = ar.take<float>(NT * 2 * E); w.dho = ar.take<float>(NT * 2 * E);
  for (int k = 0; k < 2; ++k) { w.dgi_q[k] = ar.take<float>(NT * 3 * E); w.dgh_q[k] = ar.take<float>(NT * 3 * E); }
  w.dhq_carry = ar.take<float>(N * E); w.dxq = ar.take<float>(NT * E); w.dzq_carry = ar.take<float>(N * E);
  w.zsel = ar.take<float>(NT * E); w.dpz = ar.take<float>(NT * E);
  w.dPp = ar.take<float>(N * Te * E); w.dPd = ar.take<float>(N * Te * A); w.dmem = ar.take<float>(N * Te * E);
  w.dmem2 = ar.take<float>(N * Te * E);
  w.dpool = ar.take<float>(N * 2 * E);
  w.bars = ar.take<unsigned>(8 * 128);
  w.bytes = ar.off;
  return w;
}

inline unsigned long long flag_mask(const uint8_t* flags, int T) {
  unsigned long long m = 0;
  for (int t = 0; t < T && t < 64; ++t)
    if (flags[t]) m |= 1ull << t;
  return m;
}

inline int check_dims(const acvae_dims* d) {
  ACVAE_REQUIRE(d != nullptr, "dims is NULL");
  ACVAE_REQUIRE(d->N > 0 && d->Te > 0 && d->T > 0 && d->E > 0 && d->A > 0 && d->V > 0 && d->Eenc > 0, "non-positive dimension");
  ACVAE_REQUIRE(d->E % 4 == 0 && d->A % 4 == 0 && d->Eenc % 4 == 0, "E, A, Eenc must be multiples of 4");
  ACVAE_REQUIRE(d->mem_rep >= 1 && d->N % d->mem_rep == 0, "N must be a multiple of mem_rep");
  ACVAE_REQUIRE(d->variant == 0 || d->variant == 1, "variant must be 0 (hybrid) or 1 (vae)");
  return 0;
}

// ---- plain GEMM helpers -----------------------------------------------------------
// C[M,Nn] (ldc) = A[M,K](lda) . W[Nn,K](ldw)^T + bias   (nn.Linear forward)
inline int linear_fwd(int M, int Nn, int K, const float* a, long long lda, const float* w, long long ldw,
                      const float* bias, float* c, long long ldc, cudaStream_t st, int accumulate = 0) {
  GemmParams p{};
  p.M = M; p.U = Nn; p.G = 1; p.nseg = 1;
  p.seg[0] = seg_plain(a, lda, w, ldw, K);
  p.epi.c[0] = c; p.epi.ldc = ldc; p.epi.bias[0] = bias; p.epi.scale = 1.0f; p.epi.accumulate = accumulate;
  return launch_gemm<EPI_PLAIN>(p, st);
}
// dX[M,K] (lddx) (+)= dY[M,Nn](lddy) . W[Nn,.](ldw)            (nn.Linear backward-data; pass w + col0)
inline int linear_bwd_data(int M, int K, int Nn, const float* dy, long long lddy, const float* w, long long ldw,
                           float* dx, long long lddx, cudaStream_t st, int accumulate = 0) {
  GemmParams p{};
  p.M = M; p.U = K; p.G = 1; p.nseg = 1;
  GemmSeg s{};
  s.a = dy; s.lda = lddy; s.w[0] = w; s.ldw = ldw; s.w_trans = 1; s.K = Nn;
  p.seg[0] = s;
  p.epi.c[0] = dx; p.epi.ldc = lddx; p.epi.scale = 1.0f; p.epi.accumulate = accumulate; p.epi.free_order = 1;
  return launch_gemm<EPI_PLAIN>(p, st);
}
// dW[Nn,K] (lddw) = dY[R,Nn](lddy)^T . X[R,K](ldx)             (nn.Linear backward-weight)
// k_zero_period/rem: rows r with r % period == rem are skipped (shifted recurrent operands).
inline int linear_bwd_weight(int Nn, int K, int R, const float* dy, long long lddy, const float* x, long long ldx,
                             float* dw, long long lddw, cudaStream_t st, int k_zero_period = 0, int k_zero_rem = 0,
                             int x_row_shift = 0, int accumulate = 0) {
  GemmParams p{};
  p.M = Nn; p.U = K; p.G = 1; p.nseg = 1;
  GemmSeg s{};
  s.a = dy; s.lda = lddy; s.a_trans = 1; s.w[0] = x; s.ldw = ldx; s.w_trans = 1; s.K = R;
  s.k_zero_period = k_zero_period; s.k_zero_rem = k_zero_rem; s.w_row_shift = x_row_shift;
  p.seg[0] = s;
  p.epi.c[0] = dw; p.epi.ldc = lddw; p.epi.scale = 1.0f; p.epi.free_order = 1; p.epi.accumulate = accumulate;
  return launch_gemm<EPI_PLAIN>(p, st);
}
inline int colsum(long long rows, int cols, const float* x, long long ld, float* out, cudaStream_t st, int accumulate = 0) {
  ACVAE_LAUNCH(colsum_kernel, (cols + 31) / 32, 1024, 0, st, rows, cols, x, ld, out, accumulate);
  return 0;
}
inline int gather_rows(int rows, int E, const float* table, const int* idx, float* out, cudaStream_t st) {
  ACVAE_LAUNCH(embed_gather_kernel, grid1d((long long)rows * E), 256, 0, st, rows, E, table, idx, out);
  return 0;
}
inline int scatter_rows(int rows, int E, const float* d, long long ldd, const int* idx, float* grad, cudaStream_t st) {
  ACVAE_LAUNCH(embed_scatter_add_kernel, grid1d((long long)rows * E), 256, 0, st, rows, E, d, ldd, idx, grad);
  return 0;
}

// ---- H1: memory projection (vae_model.py:743-744) + memory halves of both attentions --
inline int memory_prepare(const acvae_dims& d, const acvae_weights& w, const float* audio, float* mem, float* Pp,
                          float* Pd, cudaStream_t st) {
  const int clips = d.N / d.mem_rep;
  const int R = clips * d.Te;
  if (w.ln_w) {
    ACVAE_TRY(linear_fwd(R, d.E, d.Eenc, audio, d.Eenc, w.ln_w, d.Eenc, w.ln_b, mem, d.E, st));
  } else {
    ACVAE_REQUIRE(d.Eenc == d.E, "ln weights missing but Eenc != E");
    ACVAE_CHECK(cudaMemcpyAsync(mem, audio, sizeof(float) * (size_t)R * d.E, cudaMemcpyDeviceToDevice, st));
  }
  // h2attn.weight is [A, Dq+E]; columns [Dq:] act on the memory (attn_model.py:31 puts the query first)
  // the prior's attention width is the memory width E (text_encoder.py:225: Seq2SeqAttention(E, E, E));
  // only the decoder's is configurable (decoder.py:172 attn_size)
  ACVAE_TRY(linear_fwd(R, d.E, d.E, mem, d.E, w.p_attn_w + d.E, 2 * d.E, w.p_attn_b, Pp, d.E, st));
  ACVAE_TRY(linear_fwd(R, d.A, d.E, mem, d.E, w.d_attn_w + d.E, 2 * d.E, w.d_attn_b, Pd, d.A, st));
  return 0;
}

// ---- vocabulary statistics over `M` hidden rows ---------------------------------------
struct VocabStatsArgs {
  int M, V, E;
  const float* hidden; long long ld_h;
  const float* cls_w; const float* cls_b;
  float *pmax, *pexp, *psum, *pbest; int* parg;       // partial buffers [M, ntiles]
  const float* noise; long long ld_noise; float inv_temp; int noise_is_gumbel;
  const unsigned long long* rng; int rng_step;          // counter-based noise drawn in the epilogue (gemm.cuh philox_uniform4)
  const int* live;
  VocabReduceParams red;                                // outputs (M/ntiles/partials filled in here)
};
inline int vocab_stats(VocabStatsArgs a, cudaStream_t st) {
  GemmParams p{};
  p.M = a.M; p.U = a.V; p.G = 1; p.nseg = 1;
  p.seg[0] = seg_plain(a.hidden, a.ld_h, a.cls_w, a.E, a.E);
  p.epi.bias[0] = a.cls_b;
  p.epi.pmax = a.pmax; p.epi.pexp = a.pexp; p.epi.psum = a.psum; p.epi.pbest = a.pbest; p.epi.parg = a.parg;
  p.epi.noise = a.noise; p.epi.ld_noise = a.ld_noise; p.epi.inv_temp = a.inv_temp; p.epi.noise_is_gumbel = a.noise_is_gumbel;
  p.epi.rng = a.rng; p.epi.rng_step = a.rng_step;
  p.live = a.live; a.red.live = a.live;
  int used_tc = 0;
  ACVAE_TRY(launch_gemm<EPI_STATS>(p, st, &used_tc));
  const int tile = used_tc == 1 ? kTcBN : kVocabTile;  // column width of one partial of the kernel that ran (persistent tcgen05: half tiles)
  a.red.M = a.M; a.red.ntiles = (a.V + tile - 1) / tile;
  a.red.pmax = a.pmax; a.red.pexp = a.pexp; a.red.psum = a.psum; a.red.pbest = a.pbest; a.red.parg = a.parg;
  ACVAE_LAUNCH(vocab_reduce_kernel, (a.M + 3) / 4, 128, 0, st, a.red);
  return 0;
}

// ---- one prior step (text_encoder.py:247-268) ---------------------------------------------
// State buffers are [N,S,*]: S = T history slots in training (slot = t), S = 2 ring slots in
// sampling.  Row n of slot s lives at base + (n*S + s)*width.
struct StepCtx {
  const acvae_dims& d; const acvae_weights& w; cudaStream_t st;
  const int* mem_lens; const float *mem, *Pp, *Pd;
  const int* live;   // optional device flag: kernels return immediately when *live == 0 (sampling early stop)
};
struct StepBufs {
  int S;
  float *qp_p, *w_p, *ctx_p, *gates_p, *c_p, *h_p, *pm, *pl, *pz;   // prior
  float *qp_d, *w_d, *ctx_d, *gates_d, *hd;                          // decoder (hd = GRU hidden = `outputs`)
};

inline int prior_step(const StepCtx& c, const StepBufs& b, int slot, int slot_prev /* <0: zero state */,
                      const int* words, long long words_stride, const float* eps_t /*[N,E]*/) {
  const int N = c.d.N, E = c.d.E, A = c.d.A, Te = c.d.Te;
  const long long S = b.S;
  // query projection q.Wq^T with the word embedding as the query (text_encoder.py:249-251)
  GemmParams qg{};
  qg.M = N; qg.U = E; qg.G = 1; qg.nseg = 1; qg.live = c.live;
  qg.seg[0] = seg_gather(c.w.p_emb, E, words, words_stride, c.w.p_attn_w, 2 * E, E);
  qg.epi.c[0] = b.qp_p + (long long)slot * E; qg.epi.ldc = S * E; qg.epi.scale = 1.0f;
  ACVAE_TRY(launch_gemm<EPI_PLAIN>(qg, c.st));
  AttnFwdParams a{};
  a.rows = N; a.Te = Te; a.A = E; a.E = E; a.Dq = E; a.rows_per_clip = c.d.mem_rep; a.live = c.live;
  a.qp_in = b.qp_p + (long long)slot * E; a.ld_qp_in = S * E;
  a.P = c.Pp; a.mem = c.mem; a.v = c.w.p_attn_v; a.mem_lens = c.mem_lens;
  a.ctx = b.ctx_p + (long long)slot * E; a.ld_ctx = S * E;
  a.w_out = b.w_p + (long long)slot * Te; a.ld_w = S * Te;
  ACVAE_TRY(launch_attn_fwd(a, c.st));

  GemmParams g{};
  g.M = N; g.U = E; g.G = 4; g.live = c.live;
  int ns = 0;
  g.seg[ns] = seg_gates(c.w.p_emb, E, c.w.p_wih, 3 * E, 0, E, E, 4);
  g.seg[ns].gather = words; g.seg[ns].gather_stride = words_stride; ++ns;
  g.seg[ns++] = seg_gates(b.ctx_p + (long long)slot * E, S * E, c.w.p_wih, 3 * E, E, E, E, 4);
  if (slot_prev >= 0) {
    g.seg[ns++] = seg_gates(b.pz + (long long)slot_prev * E, S * E, c.w.p_wih, 3 * E, 2 * E, E, E, 4);  // last_z (vae_model.py:869)
    g.seg[ns++] = seg_gates(b.h_p + (long long)slot_prev * E, S * E, c.w.p_whh, E, 0, E, E, 4);
  }
  g.nseg = ns;
  g.epi.b_ih = c.w.p_bih; g.epi.b_hh = c.w.p_bhh;
  g.epi.prev = slot_prev >= 0 ? b.c_p + (long long)slot_prev * E : nullptr; g.epi.ld_prev = S * E;
  g.epi.gates = b.gates_p + (long long)slot * 4 * E; g.epi.ld_gates = S * 4 * E;
  g.epi.out0 = b.c_p + (long long)slot * E; g.epi.ld_out0 = S * E;
  g.epi.out1 = b.h_p + (long long)slot * E; g.epi.ld_out1 = S * E;
  ACVAE_TRY(launch_gemm<EPI_LSTM>(g, c.st));

  GemmParams h{};
  h.M = N; h.U = E; h.G = 2; h.nseg = 1; h.live = c.live;
  h.seg[0] = seg_gates(b.h_p + (long long)slot * E, S * E, c.w.p_head_w, E, 0, E, E, 2);
  h.epi.bias[0] = c.w.p_head_b; h.epi.bias[1] = c.w.p_head_b + E;
  h.epi.eps = eps_t; h.epi.ld_eps = E;
  h.epi.out0 = b.pm + (long long)slot * E; h.epi.ld_out0 = S * E;
  h.epi.out1 = b.pl + (long long)slot * E; h.epi.ld_out1 = S * E;
  h.epi.out2 = b.pz + (long long)slot * E; h.epi.ld_out2 = S * E;
  ACVAE_TRY(launch_gemm<EPI_HEAD>(h, c.st));
  return 0;
}

// ---- one decoder step (decoder.py:175-203) ---------------------------------------------------
inline int decoder_step(const StepCtx& c, const StepBufs& b, int slot, int slot_prev, const int* words,
                        long long words_stride, const float* z, long long ld_z,
                        float* aw_out, long long aw_ld_r, long long aw_ld_j) {
  const int N = c.d.N, E = c.d.E, A = c.d.A, Te = c.d.Te;
  const long long S = b.S;
  const float* hprev = slot_prev >= 0 ? b.hd + (long long)slot_prev * E : nullptr;  // zero at t=0 (decoder.py:94-98)
  // query projection h_{t-1}.Wq^T (decoder.py:186); zero query at t = 0
  GemmParams qg{};
  qg.M = N; qg.U = A; qg.G = 1; qg.nseg = hprev ? 1 : 0; qg.live = c.live;
  if (hprev) qg.seg[0] = seg_plain(hprev, S * E, c.w.d_attn_w, 2 * E, E);
  qg.epi.c[0] = b.qp_d + (long long)slot * A; qg.epi.ldc = S * A; qg.epi.scale = 1.0f;
  ACVAE_TRY(launch_gemm<EPI_PLAIN>(qg, c.st));
  AttnFwdParams a{};
  a.rows = N; a.Te = Te; a.A = A; a.E = E; a.Dq = E; a.rows_per_clip = c.d.mem_rep; a.live = c.live;
  a.qp_in = b.qp_d + (long long)slot * A; a.ld_qp_in = S * A;
  a.P = c.Pd; a.mem = c.mem; a.v = c.w.d_attn_v; a.mem_lens = c.mem_lens;
  a.ctx = b.ctx_d + (long long)slot * E; a.ld_ctx = S * E;
  a.w_out = b.w_d + (long long)slot * Te; a.ld_w = S * Te;
  a.aw_out = aw_out; a.aw_ld_r = aw_ld_r; a.aw_ld_j = aw_ld_j;
  ACVAE_TRY(launch_attn_fwd(a, c.st));

  GemmParams g{};
  g.M = N; g.U = E; g.G = 4; g.live = c.live;
  int ns = 0;
  auto xseg = [&](const float* ap, long long lda, int col0) {
    GemmSeg s = seg_gates(ap, lda, c.w.d_wih, 3 * E, col0, E, E, 3);
    s.w[3] = nullptr;  // n_h gets nothing from x
    return s;
  };
  g.seg[ns] = xseg(c.w.d_emb, E, 0);
  g.seg[ns].gather = words; g.seg[ns].gather_stride = words_stride; ++ns;
  g.seg[ns++] = xseg(b.ctx_d + (long long)slot * E, S * E, E);
  g.seg[ns++] = xseg(z, ld_z, 2 * E);
  if (hprev) {
    GemmSeg s = seg_gates(hprev, S * E, c.w.d_whh, E, 0, E, E, 3);
    s.w[3] = s.w[2]; s.w[2] = nullptr;  // W_hn feeds the separate n_h accumulator
    g.seg[ns++] = s;
  }
  g.nseg = ns;
  g.epi.b_ih = c.w.d_bih; g.epi.b_hh = c.w.d_bhh;
  g.epi.prev = hprev; g.epi.ld_prev = S * E;
  g.epi.gates = b.gates_d + (long long)slot * 4 * E; g.epi.ld_gates = S * 4 * E;
  g.epi.out0 = b.hd + (long long)slot * E; g.epi.ld_out0 = S * E;
  ACVAE_TRY(launch_gemm<EPI_GRU>(g, c.st));
  return 0;
}

// ---- H2: posterior (text_encoder.py:182-216 hybrid / :121-154 AR) ----------------------
inline int posterior_fwd(const acvae_dims& d, const acvae_weights& w, const acvae_train_io& io, TrainWs& ws,
                         cudaStream_t st) {
  const int N = d.N, T = d.T, E = d.E, NT = N * T;
  ACVAE_TRY(gather_rows(NT, E, w.q_emb, ws.qids, ws.xq, st));                      // :184
  for (int dir = 0; dir < 2; ++dir)
    ACVAE_TRY(linear_fwd(NT, 3 * E, E, ws.xq, E, w.q_wih[dir], E, w.q_bih[dir], ws.gxq[dir], 3 * E, st));
  // packed bidirectional GRU, zero initial state, zero outputs at padded steps (:188-191)
  for (int s = 0; s < T; ++s) {
    for (int dir = 0; dir < 2; ++dir) {
      const int t = dir == 0 ? s : T - 1 - s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      const bool has_prev = s > 0;
      GemmParams g{};
      g.M = N; g.U = E; g.G = 4; g.nseg = 0;
      const float* hp = has_prev ? ws.ho + (long long)tp * 2 * E + dir * E : nullptr;
      if (has_prev) {
        GemmSeg sg = seg_gates(hp, (long long)T * 2 * E, w.q_whh[dir], E, 0, E, E, 3);
        sg.w[3] = sg.w[2]; sg.w[2] = nullptr;
        g.seg[0] = sg; g.nseg = 1;
      }
      g.epi.gx = ws.gxq[dir] + (long long)t * 3 * E; g.epi.ld_gx = (long long)T * 3 * E;
      g.epi.b_hh = w.q_bhh[dir];
      g.epi.prev = hp; g.epi.ld_prev = (long long)T * 2 * E;
      g.epi.lens = ws.steplens; g.epi.t = t;
      g.epi.gates = ws.gq[dir] + (long long)t * 4 * E; g.epi.ld_gates = (long long)T * 4 * E;
      g.epi.out0 = ws.ho + (long long)t * 2 * E + dir * E; g.epi.ld_out0 = (long long)T * 2 * E;
      ACVAE_TRY(launch_gemm<EPI_GRU>(g, st));
    }
  }
  if (d.variant == 0) {
    GemmParams h{};
    h.M = NT; h.U = E; h.G = 2; h.nseg = 1;
    h.seg[0] = seg_gates(ws.ho, 2 * E, w.q_head_w, 2 * E, 0, 2 * E, E, 2);          // :193-195
    h.epi.bias[0] = w.q_head_b; h.epi.bias[1] = w.q_head_b + E;
    h.epi.eps = io.eps_q; h.epi.ld_eps = E;                                          // :196
    h.epi.out0 = io.q_means; h.epi.out1 = io.q_logs; h.epi.out2 = io.q_z;
    h.epi.ld_out0 = h.epi.ld_out1 = h.epi.ld_out2 = E;
    ACVAE_TRY(launch_gemm<EPI_HEAD>(h, st));
    ACVAE_LAUNCH(pool_fwd_kernel, grid1d((long long)N * 2 * E), 256, 0, st, N, T, 2 * E, ws.ho, ws.steplens, 0,
                 io.q_means_utt, ws.amax_q);                                         // :199-201
  } else {
    // autoregressive posterior: [ho_t ‖ z_{t-1}] -> mean/log, one noise draw per step (:137-150)
    for (int t = 0; t < T; ++t) {
      GemmParams h{};
      h.M = N; h.U = E; h.G = 2;
      int ns = 0;
      h.seg[ns++] = seg_gates(ws.ho + (long long)t * 2 * E, (long long)T * 2 * E, w.q_head_w, 3 * E, 0, 2 * E, E, 2);
      if (t > 0) h.seg[ns++] = seg_gates(io.q_z + (long long)(t - 1) * E, (long long)T * E, w.q_head_w, 3 * E, 2 * E, E, E, 2);
      h.nseg = ns;
      h.epi.bias[0] = w.q_head_b; h.epi.bias[1] = w.q_head_b + E;
      h.epi.eps = io.eps_q + (long long)t * N * E; h.epi.ld_eps = E;
      h.epi.out0 = io.q_means + (long long)t * E; h.epi.out1 = io.q_logs + (long long)t * E;
      h.epi.out2 = io.q_z + (long long)t * E;
      h.epi.ld_out0 = h.epi.ld_out1 = h.epi.ld_out2 = (long long)T * E;
      ACVAE_TRY(launch_gemm<EPI_HEAD>(h, st));
    }
  }
  return 0;
}

__global__ void steplens_kernel(int N, const int* __restrict__ cap_lens, int* __restrict__ steplens) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) steplens[i] = cap_lens[i] - 1;   // text_encoder.py:186, vae_model.py:703
}
__global__ void qids_kernel(int N, int T, int L, const int* __restrict__ caps, int* __restrict__ qids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * T) qids[i] = caps[(long long)(i / T) * L + (i % T)];
}

inline int train_fwd(const acvae_dims& d, const acvae_weights& w, const acvae_train_io& io, void* workspace,
                     cudaStream_t st) {
  TrainWs ws = carve_train_ws(d, workspace);
  const int N = d.N, T = d.T, E = d.E, NT = N * T;
  bool all_tf = true;
  for (int t = 0; t < T; ++t) all_tf = all_tf && io.tf_flags[t];

  ACVAE_LAUNCH(steplens_kernel, grid1d(N), 256, 0, st, N, io.cap_lens, ws.steplens);
  ACVAE_LAUNCH(qids_kernel, grid1d(NT), 256, 0, st, N, T, d.L, io.caps_ids, ws.qids);
  // words for teacher-forced steps and a free step 0 (vae_model.py:826-832)
  ACVAE_LAUNCH(words_init_kernel, grid1d(NT), 256, 0, st, N, T, d.L, io.caps_ids, flag_mask(io.tf_flags, T), kStartIdx,
               ws.words);
  ACVAE_TRY(wait_input_event(st));
  ACVAE_TRY(memory_prepare(d, w, io.audio_embeds, ws.mem, ws.Pp, ws.Pd, st));
  ACVAE_TRY(posterior_fwd(d, w, io, ws, st));

  StepCtx c{d, w, st, io.mem_lens, ws.mem, ws.Pp, ws.Pd, nullptr};
  StepBufs b{T, ws.qp_p, ws.w_p, ws.ctx_p, ws.gates_p, ws.c_p, ws.h_p, io.p_means, io.p_logs, io.p_z,
             ws.qp_d, ws.w_d, ws.ctx_d, ws.gates_d, io.outputs};
  for (int t = 0; t < T; ++t) {
    ACVAE_TRY(prior_step(c, b, t, t - 1, ws.words + t, T, io.eps_p + (long long)t * N * E));
    const float* z = (io.dis_flags[t] ? io.p_z : io.q_z) + (long long)t * E;        // vae_model.py:800-806
    ACVAE_TRY(decoder_step(c, b, t, t - 1, ws.words + t, T, z, (long long)T * E,
                           io.attn_weights ? io.attn_weights + t : nullptr, (long long)d.Te * T, T));
    if (!all_tf) {
      // greedy word of this step (word_model.py:177-179), fed to step t+1 when that step is free
      VocabStatsArgs v{};
      v.M = N; v.V = d.V; v.E = E; v.hidden = io.outputs + (long long)t * E; v.ld_h = (long long)T * E;
      v.cls_w = w.cls_w; v.cls_b = w.cls_b;
      v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
      v.red.lse = io.logit_lse + t; v.red.lsum = io.logit_sum + t; v.red.logprob = io.sampled_logprobs + t;
      v.red.ld_row = T; v.red.seqs = (long long*)io.seqs + t; v.red.ld_seqs = T;
      if (t + 1 < T && !io.tf_flags[t + 1]) { v.red.next_word = ws.words + t + 1; v.red.ld_next = T; }
      ACVAE_TRY(vocab_stats(v, st));
    }
  }
  if (all_tf) {
    VocabStatsArgs v{};
    v.M = NT; v.V = d.V; v.E = E; v.hidden = io.outputs; v.ld_h = E;
    v.cls_w = w.cls_w; v.cls_b = w.cls_b;
    v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
    v.red.lse = io.logit_lse; v.red.lsum = io.logit_sum; v.red.logprob = io.sampled_logprobs; v.red.ld_row = 1;
    v.red.seqs = (long long*)io.seqs; v.red.ld_seqs = 1;
    ACVAE_TRY(vocab_stats(v, st));
  }
  if (d.variant == 0) {
    // H8 global-constraint head (vae_model.py:722-729)
    ACVAE_LAUNCH(pool_fwd_kernel, grid1d((long long)N * E), 256, 0, st, N, T, E, io.outputs, ws.steplens, 0, ws.pool_d,
                 ws.amax_d);
    ACVAE_TRY(linear_fwd(N, 2 * E, E, ws.pool_d, E, w.g_w, E, w.g_b, io.p_means_utt, 2 * E, st));
  }
  // embeddings actually consumed (needed by the backward and by rnn_input)
  ACVAE_TRY(gather_rows(NT, E, w.p_emb, ws.words, ws.xp, st));
  ACVAE_TRY(gather_rows(NT, E, w.d_emb, ws.words, ws.xd, st));
  if (io.rnn_input) {
    // vae_model.py:187: [embed ‖ ctx ‖ z]
    ACVAE_LAUNCH(copy2d_kernel, grid1d((long long)NT * E), 256, 0, st, (long long)NT, E, ws.xd, (long long)E, io.rnn_input, (long long)3 * E);
    ACVAE_LAUNCH(copy2d_kernel, grid1d((long long)NT * E), 256, 0, st, (long long)NT, E, ws.ctx_d, (long long)E, io.rnn_input + E, (long long)3 * E);
    for (int t = 0; t < T; ++t) {
      const float* z = (io.dis_flags[t] ? io.p_z : io.q_z) + (long long)t * E;
      ACVAE_LAUNCH(copy2d_kernel, grid1d((long long)N * E), 256, 0, st, (long long)N, E, z, (long long)T * E,
                   io.rnn_input + (long long)t * 3 * E + 2 * E, (long long)T * 3 * E);
    }
  }
  if (io.logits) ACVAE_TRY(linear_fwd(NT, d.V, E, io.outputs, E, w.cls_w, E, w.cls_b, io.logits, d.V, st));
  return 0;
}

// ===================================== backward ================================================
inline int train_bwd(const acvae_dims& d, const acvae_weights& w, const acvae_train_io& io,
                     const acvae_train_grads_in& gi, acvae_weight_grads& gw, float* d_audio, void* workspace,
                     cudaStream_t st) {
  TrainWs ws = carve_train_ws(d, workspace);
  const int N = d.N, T = d.T, E = d.E, A = d.A, Te = d.Te, NT = N * T, V = d.V;
  const long long s1 = T;
  auto zero = [&](float* p, size_t n) { return cudaMemsetAsync(p, 0, n * sizeof(float), st); };

  // ---- global head (vae_model.py:722-729) ------------------------------------------------
  const float* dpool = nullptr;
  if (d.variant == 0 && gi.d_p_means_utt) {
    ACVAE_TRY(linear_bwd_data(N, E, 2 * E, gi.d_p_means_utt, 2 * E, w.g_w, E, ws.dpool, E, st));
    ACVAE_TRY(linear_bwd_weight(2 * E, E, N, gi.d_p_means_utt, 2 * E, ws.pool_d, E, gw.g_w, E, st));
    ACVAE_TRY(colsum(N, 2 * E, gi.d_p_means_utt, 2 * E, gw.g_b, st));
    dpool = ws.dpool;
  } else if (d.variant == 0) {
    ACVAE_CHECK(zero(gw.g_w, (size_t)2 * E * E)); ACVAE_CHECK(zero(gw.g_b, (size_t)2 * E));
  }
  ACVAE_LAUNCH(pool_bwd_kernel, grid1d((long long)NT * E), 256, 0, st, N, T, E, dpool, ws.steplens, 0, ws.amax_d,
               gi.d_outputs, ws.dout);

  // ---- decoder BPTT (decoder.py:175-203 reversed) ----------------------------------------
  for (int t = T - 1; t >= 0; --t) {
    GruBwdParams g{};
    g.N = N; g.U = E;
    g.dh_ext = ws.dout + (long long)t * E; g.ld_dh_ext = s1 * E;
    g.dh_carry = t < T - 1 ? ws.dh_carry : nullptr;
    g.gates = ws.gates_d + (long long)t * 4 * E; g.ld_gates = s1 * 4 * E;
    g.hprev = t > 0 ? io.outputs + (long long)(t - 1) * E : nullptr; g.ld_hprev = s1 * E;
    g.dgi = ws.dgi_d + (long long)t * 3 * E; g.ld_dgi = s1 * 3 * E;
    g.dgh = ws.dgh_d + (long long)t * 3 * E; g.ld_dgh = s1 * 3 * E;
    g.dh_out = ws.dh_carry;
    ACVAE_LAUNCH(gru_bwd_kernel, grid1d((long long)N * E), 256, 0, st, g);
    // d ctx_t = dGi . W_ih[:, E:2E]
    ACVAE_TRY(linear_bwd_data(N, E, 3 * E, ws.dgi_d + (long long)t * 3 * E, s1 * 3 * E, w.d_wih + E, 3 * E,
                              ws.dctx_d + (long long)t * E, s1 * E, st));
    AttnBwdQParams a{};
    a.rows = N; a.Te = Te; a.A = A; a.E = E; a.rows_per_clip = 1;
    a.dctx = ws.dctx_d + (long long)t * E; a.ld_dctx = s1 * E;
    a.w = ws.w_d + (long long)t * Te; a.ld_w = s1 * Te;
    a.qp = ws.qp_d + (long long)t * A; a.ld_qp = s1 * A;
    a.P = ws.Pd; a.mem = ws.mem; a.v = w.d_attn_v; a.mem_lens = io.mem_lens;
    a.ds = ws.ds_d + (long long)t * Te; a.ld_ds = s1 * Te;
    a.dqp = ws.dqp_d + (long long)t * A; a.ld_dqp = s1 * A;
    ACVAE_TRY(launch_attn_bwd_q(a, st));
    if (t > 0) {
      // dh_{t-1} = dh*z (already in dh_carry) + dGh . W_hh + dqp . Wq
      GemmParams p{};
      p.M = N; p.U = E; p.G = 1; p.nseg = 2;
      GemmSeg s0{}; s0.a = ws.dgh_d + (long long)t * 3 * E; s0.lda = s1 * 3 * E; s0.w[0] = w.d_whh; s0.ldw = E; s0.w_trans = 1; s0.K = 3 * E;
      GemmSeg s1_{}; s1_.a = ws.dqp_d + (long long)t * A; s1_.lda = s1 * A; s1_.w[0] = w.d_attn_w; s1_.ldw = 2 * E; s1_.w_trans = 1; s1_.K = A;
      p.seg[0] = s0; p.seg[1] = s1_;
      p.epi.c[0] = ws.dh_carry; p.epi.ldc = E; p.epi.scale = 1.0f; p.epi.accumulate = 1;
      ACVAE_TRY(launch_gemm<EPI_PLAIN>(p, st));
    }
  }
  // batched remainders of the decoder
  ACVAE_TRY(linear_bwd_data(NT, E, 3 * E, ws.dgi_d, 3 * E, w.d_wih, 3 * E, ws.dxe_d, E, st));          // d embed
  ACVAE_TRY(linear_bwd_data(NT, E, 3 * E, ws.dgi_d, 3 * E, w.d_wih + 2 * E, 3 * E, ws.dxz_d, E, st));  // d z fed to the decoder
  ACVAE_CHECK(zero(gw.d_emb, (size_t)V * E));
  ACVAE_TRY(scatter_rows(NT, E, ws.dxe_d, E, ws.words, gw.d_emb, st));
  ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, ws.xd, E, gw.d_wih, 3 * E, st));
  ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, ws.ctx_d, E, gw.d_wih + E, 3 * E, st));
  // z column block of W_ih: the z that was fed at step t is q_z or the prior's z (vae_model.py:800-806)
  {
    bool any_dis = false, any_q = false;
    for (int t = 0; t < T; ++t) { any_dis = any_dis || io.dis_flags[t]; any_q = any_q || !io.dis_flags[t]; }
    if (!any_dis) {
      ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, io.q_z, E, gw.d_wih + 2 * E, 3 * E, st));
    } else if (!any_q) {
      ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, io.p_z, E, gw.d_wih + 2 * E, 3 * E, st));
    } else {
      // mixed: one [3E,E] += dGi_t^T . z_t GEMM per step
      for (int t = 0; t < T; ++t) {
        GemmParams p{};
        p.M = 3 * E; p.U = E; p.G = 1; p.nseg = 1;
        GemmSeg s{};
        s.a = ws.dgi_d + (long long)t * 3 * E; s.lda = s1 * 3 * E; s.a_trans = 1;
        s.w[0] = (io.dis_flags[t] ? io.p_z : io.q_z) + (long long)t * E; s.ldw = s1 * E; s.w_trans = 1; s.K = N;
        p.seg[0] = s;
        p.epi.c[0] = gw.d_wih + 2 * E; p.epi.ldc = 3 * E; p.epi.scale = 1.0f; p.epi.accumulate = t > 0;
        ACVAE_TRY(launch_gemm<EPI_PLAIN>(p, st));
      }
    }
  }
  ACVAE_TRY(colsum(NT, 3 * E, ws.dgi_d, 3 * E, gw.d_bih, st));
  ACVAE_TRY(colsum(NT, 3 * E, ws.dgh_d, 3 * E, gw.d_bhh, st));
  // W_hh and the attention query half see h_{t-1}: shifted rows, t == 0 rows skipped
  ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgh_d, 3 * E, io.outputs - E, E, gw.d_whh, E, st, T, 0, -1));
  ACVAE_TRY(linear_bwd_weight(A, E, NT, ws.dqp_d, A, io.outputs - E, E, gw.d_attn_w, 2 * E, st, T, 0, -1));

  // ---- prior BPTT (text_encoder.py:247-268 reversed; chain through last_z, vae_model.py:869) ----
  for (int t = T - 1; t >= 0; --t) {
    HeadBwdParams h{};
    h.rows = N; h.U = E;
    if (gi.d_p_z) { h.dz0 = gi.d_p_z + (long long)t * E; h.ld_dz0 = s1 * E; }
    if (io.dis_flags[t]) { h.dz1 = ws.dxz_d + (long long)t * E; h.ld_dz1 = s1 * E; }
    if (t < T - 1) { h.dz2 = ws.dzp_carry; h.ld_dz2 = E; }
    if (gi.d_p_means) { h.dmean = gi.d_p_means + (long long)t * E; h.ld_dmean = s1 * E; }
    if (gi.d_p_logs) { h.dlog = gi.d_p_logs + (long long)t * E; h.ld_dlog = s1 * E; }
    h.eps = io.eps_p + (long long)t * N * E; h.ld_eps = E;
    h.logv = io.p_logs + (long long)t * E; h.ld_logv = s1 * E;
    h.dml = ws.dml_p + (long long)t * 2 * E; h.ld_dml = s1 * 2 * E;
    ACVAE_LAUNCH(head_bwd_kernel, grid1d((long long)N * E), 256, 0, st, h);
    // dh_t = dML . W_head (+ carry from step t+1)
    ACVAE_TRY(linear_bwd_data(N, E, 2 * E, ws.dml_p + (long long)t * 2 * E, s1 * 2 * E, w.p_head_w, E, ws.dhp_carry, E,
                              st, t < T - 1));
    LstmBwdParams l{};
    l.N = N; l.U = E; l.dh = ws.dhp_carry; l.dc_carry = t < T - 1 ? ws.dcp_carry : nullptr;
    l.gates = ws.gates_p + (long long)t * 4 * E; l.ld_gates = s1 * 4 * E;
    l.c = ws.c_p + (long long)t * E; l.ld_c = s1 * E;
    l.cprev = t > 0 ? ws.c_p + (long long)(t - 1) * E : nullptr; l.ld_cprev = s1 * E;
    l.dg = ws.dg_p + (long long)t * 4 * E; l.ld_dg = s1 * 4 * E;
    l.dc_out = ws.dcp_carry;
    ACVAE_LAUNCH(lstm_bwd_kernel, grid1d((long long)N * E), 256, 0, st, l);
    if (t > 0) {
      // [d last_z | d h_{t-1}] = dG . [W_ih[:, 2E:3E] | W_hh]
      ACVAE_TRY(linear_bwd_data(N, E, 4 * E, ws.dg_p + (long long)t * 4 * E, s1 * 4 * E, w.p_wih + 2 * E, 3 * E,
                                ws.dzp_carry, E, st));
      ACVAE_TRY(linear_bwd_data(N, E, 4 * E, ws.dg_p + (long long)t * 4 * E, s1 * 4 * E, w.p_whh, E, ws.dhp_carry, E, st));
    }
  }
  ACVAE_TRY(linear_bwd_data(NT, E, 4 * E, ws.dg_p, 4 * E, w.p_wih, 3 * E, ws.dxe_p, E, st));            // d word embedding
  ACVAE_TRY(linear_bwd_data(NT, E, 4 * E, ws.dg_p, 4 * E, w.p_wih + E, 3 * E, ws.dctx_p, E, st));      // d ctx
  {
    AttnBwdQParams a{};
    a.rows = NT; a.Te = Te; a.A = E; a.E = E; a.rows_per_clip = T;
    a.dctx = ws.dctx_p; a.ld_dctx = E; a.w = ws.w_p; a.ld_w = Te; a.qp = ws.qp_p; a.ld_qp = E;
    a.P = ws.Pp; a.mem = ws.mem; a.v = w.p_attn_v; a.mem_lens = io.mem_lens;
    a.ds = ws.ds_p; a.ld_ds = Te; a.dqp = ws.dqp_p; a.ld_dqp = E;
    ACVAE_TRY(launch_attn_bwd_q(a, st));
  }
  ACVAE_TRY(linear_bwd_data(NT, E, E, ws.dqp_p, E, w.p_attn_w, 2 * E, ws.dxe_p, E, st, 1));            // query = word embedding
  ACVAE_CHECK(zero(gw.p_emb, (size_t)V * E));
  ACVAE_TRY(scatter_rows(NT, E, ws.dxe_p, E, ws.words, gw.p_emb, st));
  ACVAE_TRY(linear_bwd_weight(E, E, NT, ws.dqp_p, E, ws.xp, E, gw.p_attn_w, 2 * E, st));
  ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, ws.xp, E, gw.p_wih, 3 * E, st));
  ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, ws.ctx_p, E, gw.p_wih + E, 3 * E, st));
  ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, io.p_z - E, E, gw.p_wih + 2 * E, 3 * E, st, T, 0, -1));
  ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, ws.h_p - E, E, gw.p_whh, E, st, T, 0, -1));
  ACVAE_TRY(colsum(NT, 4 * E, ws.dg_p, 4 * E, gw.p_bih, st));
  ACVAE_TRY(colsum(NT, 4 * E, ws.dg_p, 4 * E, gw.p_bhh, st));
  ACVAE_TRY(linear_bwd_weight(2 * E, E, NT, ws.dml_p, 2 * E, ws.h_p, E, gw.p_head_w, E, st));
  ACVAE_TRY(colsum(NT, 2 * E, ws.dml_p, 2 * E, gw.p_head_b, st));

  // ---- deferred attention accumulation: dP, dmem, dv ----------------------------------------
  ACVAE_CHECK(zero(gw.p_attn_v, E)); ACVAE_CHECK(zero(gw.d_attn_v, A));
  {
    AttnBwdAccParams a{};
    a.clips = N; a.Te = Te; a.A = E; a.E = E; a.rows_per_clip = T;
    a.ds = ws.ds_p; a.ld_ds = Te; a.w = ws.w_p; a.ld_w = Te; a.qp = ws.qp_p; a.ld_qp = E;
    a.dctx = ws.dctx_p; a.ld_dctx = E; a.P = ws.Pp; a.v = w.p_attn_v; a.mem_lens = io.mem_lens;
    a.dP = ws.dPp; a.dmem = ws.dmem; a.dmem_accumulate = 0; a.dv = gw.p_attn_v;
    ACVAE_TRY(launch_attn_bwd_acc(a, st));
    a.A = A; a.ld_qp = A;
    a.ds = ws.ds_d; a.w = ws.w_d; a.qp = ws.qp_d; a.dctx = ws.dctx_d; a.P = ws.Pd; a.v = w.d_attn_v;
    a.dP = ws.dPd; a.dmem_accumulate = 1; a.dv = gw.d_attn_v;
    ACVAE_TRY(launch_attn_bwd_acc(a, st));
  }

  // ---- posterior backward (text_encoder.py:182-216 / :121-154) ---------------------------------
  if (d.variant == 0) {
    HeadBwdParams h{};
    h.rows = NT; h.U = E;
    if (gi.d_q_z) { h.dz0 = gi.d_q_z; h.ld_dz0 = E; }
    h.dz1 = ws.dxz_d; h.ld_dz1 = E;
    // rows are (n,t): the decoder's dz reaches q_z only at steps that were not fed the prior's z
    h.flag_mask = flag_mask(io.dis_flags, T); h.period = T; h.want = 0; h.use_flags = 1;
    if (gi.d_q_means) { h.dmean = gi.d_q_means; h.ld_dmean = E; }
    if (gi.d_q_logs) { h.dlog = gi.d_q_logs; h.ld_dlog = E; }
    h.eps = io.eps_q; h.ld_eps = E; h.logv = io.q_logs; h.ld_logv = E;
    h.dml = ws.dml_q; h.ld_dml = 2 * E;
    ACVAE_LAUNCH(head_bwd_kernel, grid1d((long long)NT * E), 256, 0, st, h);
    // d ho = pooled-utterance path + dML . W_tml
    const float* dq_utt = gi.d_q_means_utt;
    ACVAE_LAUNCH(pool_bwd_kernel, grid1d((long long)NT * 2 * E), 256, 0, st, N, T, 2 * E, dq_utt, ws.steplens, 0,
                 ws.amax_q, (const float*)nullptr, ws.dho);
    ACVAE_TRY(linear_bwd_data(NT, 2 * E, 2 * E, ws.dml_q, 2 * E, w.q_head_w, 2 * E, ws.dho, 2 * E, st, 1));
    ACVAE_TRY(linear_bwd_weight(2 * E, 2 * E, NT, ws.dml_q, 2 * E, ws.ho, 2 * E, gw.q_head_w, 2 * E, st));
    ACVAE_TRY(colsum(NT, 2 * E, ws.dml_q, 2 * E, gw.q_head_b, st));
  } else {
    // AR posterior: reverse chain through z_{t-1}
    for (int t = T - 1; t >= 0; --t) {
      HeadBwdParams h{};
      h.rows = N; h.U = E;
      if (gi.d_q_z) { h.dz0 = gi.d_q_z + (long long)t * E; h.ld_dz0 = s1 * E; }
      if (!io.dis_flags[t]) { h.dz1 = ws.dxz_d + (long long)t * E; h.ld_dz1 = s1 * E; }
      if (t < T - 1) { h.dz2 = ws.dzq_carry; h.ld_dz2 = E; }
      if (gi.d_q_means) { h.dmean = gi.d_q_means + (long long)t * E; h.ld_dmean = s1 * E; }
      if (gi.d_q_logs) { h.dlog = gi.d_q_logs + (long long)t * E; h.ld_dlog = s1 * E; }
      h.eps = io.eps_q + (long long)t * N * E; h.ld_eps = E;
      h.logv = io.q_logs + (long long)t * E; h.ld_logv = s1 * E;
      h.dml = ws.dml_q + (long long)t * 2 * E; h.ld_dml = s1 * 2 * E;
      ACVAE_LAUNCH(head_bwd_kernel, grid1d((long long)N * E), 256, 0, st, h);
      if (t > 0)
        ACVAE_TRY(linear_bwd_data(N, E, 2 * E, ws.dml_q + (long long)t * 2 * E, s1 * 2 * E, w.q_head_w + 2 * E, 3 * E,
                                  ws.dzq_carry, E, st));
    }
    ACVAE_TRY(linear_bwd_data(NT, 2 * E, 2 * E, ws.dml_q, 2 * E, w.q_head_w, 3 * E, ws.dho, 2 * E, st));
    ACVAE_TRY(linear_bwd_weight(2 * E, 2 * E, NT, ws.dml_q, 2 * E, ws.ho, 2 * E, gw.q_head_w, 3 * E, st));
    ACVAE_TRY(linear_bwd_weight(2 * E, E, NT, ws.dml_q, 2 * E, io.q_z - E, E, gw.q_head_w + 2 * E, 3 * E, st, T, 0, -1));
    ACVAE_TRY(colsum(NT, 2 * E, ws.dml_q, 2 * E, gw.q_head_b, st));
  }
  // biGRU BPTT with the packed-sequence mask
  for (int dir = 0; dir < 2; ++dir) {
    for (int s = T - 1; s >= 0; --s) {
      const int t = dir == 0 ? s : T - 1 - s;        // forward visited t at position s
      const int tp = dir == 0 ? t - 1 : t + 1;       // where h_prev of this step lives
      GruBwdParams g{};
      g.N = N; g.U = E;
      g.dh_ext = ws.dho + (long long)t * 2 * E + dir * E; g.ld_dh_ext = s1 * 2 * E;
      g.dh_carry = s < T - 1 ? ws.dhq_carry : nullptr;
      g.gates = ws.gq[dir] + (long long)t * 4 * E; g.ld_gates = s1 * 4 * E;
      g.hprev = s > 0 ? ws.ho + (long long)tp * 2 * E + dir * E : nullptr; g.ld_hprev = s1 * 2 * E;
      g.lens = ws.steplens; g.len_off = 0; g.t = t;
      g.dgi = ws.dgi_q[dir] + (long long)t * 3 * E; g.ld_dgi = s1 * 3 * E;
      g.dgh = ws.dgh_q[dir] + (long long)t * 3 * E; g.ld_dgh = s1 * 3 * E;
      g.dh_out = ws.dhq_carry;
      ACVAE_LAUNCH(gru_bwd_kernel, grid1d((long long)N * E), 256, 0, st, g);
      if (s > 0)
        ACVAE_TRY(linear_bwd_data(N, E, 3 * E, ws.dgh_q[dir] + (long long)t * 3 * E, s1 * 3 * E, w.q_whh[dir], E,
                                  ws.dhq_carry, E, st, 1));
    }
    ACVAE_TRY(linear_bwd_data(NT, E, 3 * E, ws.dgi_q[dir], 3 * E, w.q_wih[dir], E, ws.dxq, E, st, dir));
    ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_q[dir], 3 * E, ws.xq, E, gw.q_wih[dir], E, st));
    ACVAE_TRY(colsum(NT, 3 * E, ws.dgi_q[dir], 3 * E, gw.q_bih[dir], st));
    ACVAE_TRY(colsum(NT, 3 * E, ws.dgh_q[dir], 3 * E, gw.q_bhh[dir], st));
    if (dir == 0)
      ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgh_q[0], 3 * E, ws.ho - 2 * E, 2 * E, gw.q_whh[0], E, st, T, 0, -1));
    else
      ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgh_q[1], 3 * E, ws.ho + 2 * E + E, 2 * E, gw.q_whh[1], E, st, T, T - 1, 1));
  }
  ACVAE_CHECK(zero(gw.q_emb, (size_t)V * E));
  ACVAE_TRY(scatter_rows(NT, E, ws.dxq, E, ws.qids, gw.q_emb, st));

  // ---- memory backward: attention memory halves, ln (vae_model.py:743-744) -----------------------
  const int R = N * Te;
  ACVAE_TRY(linear_bwd_data(R, E, E, ws.dPp, E, w.p_attn_w + E, 2 * E, ws.dmem, E, st, 1));
  ACVAE_TRY(linear_bwd_data(R, E, A, ws.dPd, A, w.d_attn_w + E, 2 * E, ws.dmem, E, st, 1));
  ACVAE_TRY(linear_bwd_weight(E, E, R, ws.dPp, E, ws.mem, E, gw.p_attn_w + E, 2 * E, st));
  ACVAE_TRY(linear_bwd_weight(A, E, R, ws.dPd, A, ws.mem, E, gw.d_attn_w + E, 2 * E, st));
  ACVAE_TRY(colsum(R, E, ws.dPp, E, gw.p_attn_b, st));
  ACVAE_TRY(colsum(R, A, ws.dPd, A, gw.d_attn_b, st));
  if (w.ln_w) {
    if (d_audio) ACVAE_TRY(linear_bwd_data(R, d.Eenc, E, ws.dmem, E, w.ln_w, d.Eenc, d_audio, d.Eenc, st));
    ACVAE_TRY(linear_bwd_weight(E, d.Eenc, R, ws.dmem, E, io.audio_embeds, d.Eenc, gw.ln_w, d.Eenc, st));
    ACVAE_TRY(colsum(R, E, ws.dmem, E, gw.ln_b, st));
  } else if (d_audio) {
    ACVAE_CHECK(cudaMemcpyAsync(d_audio, ws.dmem, sizeof(float) * (size_t)R * E, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

}  // namespace acvae
