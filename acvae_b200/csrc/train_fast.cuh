// Fast path of the training step for the default configuration (every step teacher-forced,
// dis_ratio == 0, Hybrid_VAEModel): same arithmetic and the same saved activations as the general path
// in train.cuh, re-scheduled for B200:
//   * everything that does not depend on a recurrent state is hoisted out of the T-step loops and
//     batched over all N*T rows (prior word attention, all input-side gate pre-activations);
//   * the four recurrent chains (posterior forward / reverse, prior, decoder) run on forked streams;
//   * pointwise backward steps are fused into the epilogue of the GEMM that produces their input, so
//     a reverse step is 3 launches (decoder), 2 (prior) or 1 per direction (posterior).
#pragma once
#include "cluster_chain.cuh"
#include "streams.cuh"
#include "train.cuh"

namespace acvae {

// The hoisted multi-stream schedule covers the hybrid variant.  Scheduled sampling (some tf_flags false: vae_model.py:826-832 feeds
// the previous step's arg-max word instead of the caption's) and prior-z replacement (dis_flags: vae_model.py:800-806, the decoder
// consumes the prior's sample at that step) are covered where the decoder chain runs on clusters next to the stand-alone prior chain:
// both chains are cut at every free step, and a segment with dis steps runs the prior first (see train_fwd_fast / train_bwd_fast).
inline bool fast_path_ok(const acvae_dims& d, const acvae_train_io& io) {
  if (d.variant != 0 || d.mem_rep != 1) return false;
  bool all_tf = true;
  for (int t = 0; t < d.T; ++t) {
    if (io.dis_flags[t]) all_tf = false;                // (either kind of flag needs the segmented chains)
    if (t > 0 && !io.tf_flags[t]) all_tf = false;       // step 0 always starts from <start>
  }
  if (!all_tf && !(cluster_chain_supported(d.N, d.T, d.Te, d.E, d.A) && chain_supported(d.N, d.T, d.Te, d.E, d.A))) return false;
  return aux() != nullptr;
}

// z each decoder step consumed: the posterior's sample, or the prior's where dis_flags[t] (vae_model.py:800-806)
__global__ void zsel_kernel(long long n, int T, int E, unsigned long long dis_mask, const float* __restrict__ q_z,
                            const float* __restrict__ p_z, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = (int)((i / E) % T);
  out[i] = ((dis_mask >> t) & 1ull) ? p_z[i] : q_z[i];
}
// d p_z = upstream gradient (or 0) + the decoder's d z at the steps where it consumed the prior's sample
__global__ void dpz_kernel(long long n, int T, int E, unsigned long long dis_mask, const float* __restrict__ up,
                           const float* __restrict__ dxz, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = (int)((i / E) % T);
  out[i] = (up ? up[i] : 0.0f) + (((dis_mask >> t) & 1ull) ? dxz[i] : 0.0f);
}

// rows of two embedding tables for ONE step's words: out[n, t, :] = table[words[n, t], :]  (buffers are [N, T, E])
__global__ void gather_step2_kernel(int N, int T, int t, int E, const float* __restrict__ t0, const float* __restrict__ t1,
                                    const int* __restrict__ words, float* __restrict__ o0, float* __restrict__ o1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * E) return;
  const int n = i / E, e = i - n * E;
  const long long src = (long long)words[n * T + t] * E + e, dst = ((long long)n * T + t) * E + e;
  o0[dst] = t0[src]; o1[dst] = t1[src];
}

inline int train_fwd_fast(const acvae_dims& d, const acvae_weights& w, const acvae_train_io& io, void* workspace,
                          cudaStream_t st_user) {
  TrainWs ws = carve_train_ws(d, workspace);
  Aux* ax = aux();
  // the critical path runs on a high-priority stream forked from (and, at the end, joined to) the caller's
  cudaStream_t st = ax->s[kAuxMain];
  ACVAE_TRY(stream_dep(st_user, st, ax));
  const int N = d.N, T = d.T, E = d.E, A = d.A, Te = d.Te, NT = N * T;
  const long long s1 = T;
  cudaStream_t sq0 = ax->s[0], sq1 = ax->s[1], sp = ax->s[2];

  // Persistent recurrent-chain kernels.  `cl`: posterior and decoder chains on thread-block clusters (cluster_chain.cuh,
  // state exchange through distributed shared memory); `coop`: the cooperative-grid chains of recurrent.cuh (exchange
  // through L2) -- used for the prior next to the cluster chains, and for every chain where the cluster form does not apply.
  const bool cl = cluster_chain_supported(N, T, Te, E, A);
  const bool coop = chain_supported(N, T, Te, E, A);
  const bool post_chain = cl || coop, prior_chain = coop, dec_chain = cl || coop;
  if (coop) ACVAE_CHECK(cudaMemsetAsync(ws.bars, 0, 8 * 128 * sizeof(unsigned), st));
  ACVAE_LAUNCH(steplens_kernel, grid1d(N), 256, 0, st, N, io.cap_lens, ws.steplens);
  ACVAE_LAUNCH(qids_kernel, grid1d(NT), 256, 0, st, N, T, d.L, io.caps_ids, ws.qids);
  ACVAE_LAUNCH(words_init_kernel, grid1d(NT), 256, 0, st, N, T, d.L, io.caps_ids, flag_mask(io.tf_flags, T), kStartIdx,
               ws.words);
  ACVAE_TRY(stream_dep(st, sq0, ax));
  ACVAE_TRY(stream_dep(st, sq1, ax));
  // the prior's word-only inputs (embedding rows, query projection) need no memory: ahead of everything on the prior's stream,
  // so that its batched word attention runs UNDER the posterior chain instead of across its end (attention CTAs fill every
  // SM for 40 us; the posterior head GEMM behind the chain then waits ~20 us for a free SM)
  ACVAE_TRY(stream_dep(st, sp, ax));
  ACVAE_TRY(gather_rows(NT, E, w.p_emb, ws.words, ws.xp, sp));
  ACVAE_TRY(linear_fwd(NT, E, E, ws.xp, E, w.p_attn_w, 2 * E, nullptr, ws.qp_p, E, sp));

  // ---- posterior (text_encoder.py:182-216): the two directions are independent chains --------------
  ACVAE_TRY(gather_rows(NT, E, w.q_emb, ws.qids, ws.xq, sq0));
  ACVAE_TRY(stream_dep(sq0, sq1, ax));
  cudaStream_t sq[2] = {sq0, sq1};
  for (int dir = 0; dir < 2; ++dir) {
    ACVAE_TRY(linear_fwd(NT, 3 * E, E, ws.xq, E, w.q_wih[dir], E, w.q_bih[dir], ws.gxq[dir], 3 * E, sq[dir]));
    if (post_chain) continue;
    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? s : T - 1 - s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      GemmParams g{};
      g.M = N; g.U = E; g.G = 4; g.nseg = 0;
      const float* hp = s > 0 ? ws.ho + (long long)tp * 2 * E + dir * E : nullptr;
      if (hp) {
        GemmSeg sg = seg_gates(hp, s1 * 2 * E, w.q_whh[dir], E, 0, E, E, 3);
        sg.w[3] = sg.w[2]; sg.w[2] = nullptr;
        g.seg[0] = sg; g.nseg = 1;
      }
      g.epi.gx = ws.gxq[dir] + (long long)t * 3 * E; g.epi.ld_gx = s1 * 3 * E;
      g.epi.b_hh = w.q_bhh[dir];
      g.epi.prev = hp; g.epi.ld_prev = s1 * 2 * E;
      g.epi.lens = ws.steplens; g.epi.t = t;
      g.epi.gates = ws.gq[dir] + (long long)t * 4 * E; g.epi.ld_gates = s1 * 4 * E;
      g.epi.out0 = ws.ho + (long long)t * 2 * E + dir * E; g.epi.ld_out0 = s1 * 2 * E;
      ACVAE_TRY(launch_gemm<EPI_GRU>(g, sq[dir]));
    }
  }
  ACVAE_TRY(stream_dep(sq1, sq0, ax));
  cudaEvent_t ev_qz = ax->ev();
  if (post_chain) {
    PostChainFwd pc{};
    pc.N = N; pc.T = T; pc.lens = ws.steplens; pc.ho = ws.ho; pc.bar = ws.bars + 0 * 128;
    for (int dir = 0; dir < 2; ++dir) { pc.gx[dir] = ws.gxq[dir]; pc.whh[dir] = w.q_whh[dir]; pc.bhh[dir] = w.q_bhh[dir]; pc.gq[dir] = ws.gq[dir]; }
    if (cl) pc.trace = chain_trace_ptr() ? chain_trace_ptr() + 2LL * T * 16 : nullptr;
    if (cl) ACVAE_TRY(launch_cluster_chain(post_cl_fwd_kernel, post_cl_clusters(N), 0, sq0, "post_cl_fwd_kernel", pc));
    else ACVAE_TRY(launch_chain(post_chain_fwd_kernel, 0, sq0, "post_chain_fwd_kernel", pc));
  }
  {
    GemmParams h{};
    h.M = NT; h.U = E; h.G = 2; h.nseg = 1;
    h.seg[0] = seg_gates(ws.ho, 2 * E, w.q_head_w, 2 * E, 0, 2 * E, E, 2);
    h.epi.bias[0] = w.q_head_b; h.epi.bias[1] = w.q_head_b + E;
    h.epi.eps = io.eps_q; h.epi.ld_eps = E;
    h.epi.out0 = io.q_means; h.epi.out1 = io.q_logs; h.epi.out2 = io.q_z;
    h.epi.ld_out0 = h.epi.ld_out1 = h.epi.ld_out2 = E;
    if (NT >= 96 && tc_enabled()) {
      // tensor cores: [mean | log] = ho . W^T + b into a scratch (dml_q is free during the forward), then the
      // reparameterisation as a small pointwise kernel
      ACVAE_TRY(linear_fwd(NT, 2 * E, 2 * E, ws.ho, 2 * E, w.q_head_w, 2 * E, w.q_head_b, ws.dml_q, 2 * E, sq0));
      ACVAE_LAUNCH(head_cell_kernel, grid1d((long long)NT * E), 256, 0, sq0, NT, E, (const float*)ws.dml_q, io.eps_q,
                   io.q_means, io.q_logs, io.q_z, (long long)E, (const int*)nullptr);
    } else {
      ACVAE_TRY(launch_gemm<EPI_HEAD>(h, sq0));
    }
    ACVAE_CHECK(cudaEventRecord(ev_qz, sq0));      // q_z is final: the decoder does not wait for the pooling below
    ACVAE_LAUNCH(pool_fwd_kernel, grid1d((long long)N * 2 * E), 256, 0, sq0, N, T, 2 * E, ws.ho, ws.steplens, 0,
                 io.q_means_utt, ws.amax_q);
  }

  // ---- memory (vae_model.py:743-744 + factorised attention halves) -------------------------------------
  ACVAE_TRY(wait_input_event(st));       // the audio copy overlaps the posterior chain, which does not read it
  ACVAE_TRY(memory_prepare(d, w, io.audio_embeds, ws.mem, ws.Pp, ws.Pd, st));
  ACVAE_TRY(stream_dep(st, sp, ax));
  // cluster decoder chain: the context's share of the gate pre-activations per FRAME, Mg = mem . W_ih[:, E:2E]^T, so that
  // the chain needs neither the K = E context product nor an exchange of the context (cluster_chain.cuh).  96 exclusive CTAs
  // that nothing needs before the decoder chain starts: on a low-priority stream, behind the prior's attention
  cudaStream_t s_mg = ax->s[kAuxFan0];

  // ---- prior (text_encoder.py:247-268): word attention and input-side gates batched over (n,t) ---------
  {
    AttnFwdParams a{};
    a.rows = NT; a.Te = Te; a.A = E; a.E = E; a.Dq = E; a.rows_per_clip = T;
    a.qp_in = ws.qp_p; a.ld_qp_in = E;
    a.P = ws.Pp; a.mem = ws.mem; a.v = w.p_attn_v; a.mem_lens = io.mem_lens;
    a.ctx = ws.ctx_p; a.ld_ctx = E; a.w_out = ws.w_p; a.ld_w = Te;
    ACVAE_TRY(launch_attn_fwd(a, sp));
    if (cl) {     // behind the attention (which must not be slowed down: it has to be gone when the posterior chain ends)
      ACVAE_TRY(stream_dep(sp, s_mg, ax));
      ACVAE_TRY(linear_fwd(N * Te, 3 * E, E, ws.mem, E, w.d_wih + E, 3 * E, nullptr, ws.Mg, 3 * E, s_mg));
    }
    // gx_p = [xe | ctx] . W_ih[:, :2E]^T + b_ih   (4E columns, gate-major), written into the gate buffer
    GemmParams g{};
    g.M = NT; g.U = 4 * E; g.G = 1; g.nseg = 2;
    g.seg[0] = seg_plain(ws.xp, E, w.p_wih, 3 * E, E);
    g.seg[1] = seg_plain(ws.ctx_p, E, w.p_wih + E, 3 * E, E);
    g.epi.c[0] = ws.dg_p; g.epi.ldc = 4 * E; g.epi.bias[0] = w.p_bih; g.epi.scale = 1.0f;   // dg_p doubles as gx_p in the forward
    ACVAE_TRY(launch_gemm<EPI_PLAIN>(g, sp));
  }
  PriorChainFwd ppc{};      // cooperative chains: the prior chain runs inside the decoder's persistent kernel (same barriers)
  if (prior_chain) {
    ppc.N = N; ppc.T = T; ppc.gx = ws.dg_p; ppc.wih = w.p_wih; ppc.whh = w.p_whh; ppc.bhh = w.p_bhh;
    ppc.head_w = w.p_head_w; ppc.head_b = w.p_head_b; ppc.eps = io.eps_p;
    ppc.gates = ws.gates_p; ppc.c = ws.c_p; ppc.h = ws.h_p; ppc.pm = io.p_means; ppc.pl = io.p_logs; ppc.pz = io.p_z;
    // next to the cluster decoder chain the prior is its own (cooperative, 128-CTA) kernel, launched below right AFTER
    // the decoder's chain
    if (cl) ppc.bar = ws.bars + 1 * 128;
  }
  for (int t = 0; t < T && !prior_chain; ++t) {
    GemmParams g{};
    g.M = N; g.U = E; g.G = 4; g.nseg = 0;
    if (t > 0) {
      g.seg[0] = seg_gates(io.p_z + (long long)(t - 1) * E, s1 * E, w.p_wih, 3 * E, 2 * E, E, E, 4);   // last_z (vae_model.py:869)
      g.seg[1] = seg_gates(ws.h_p + (long long)(t - 1) * E, s1 * E, w.p_whh, E, 0, E, E, 4);
      g.nseg = 2;
    }
    g.epi.gx = ws.dg_p + (long long)t * 4 * E; g.epi.ld_gx = s1 * 4 * E;
    g.epi.b_hh = w.p_bhh;
    g.epi.prev = t > 0 ? ws.c_p + (long long)(t - 1) * E : nullptr; g.epi.ld_prev = s1 * E;
    g.epi.gates = ws.gates_p + (long long)t * 4 * E; g.epi.ld_gates = s1 * 4 * E;
    g.epi.out0 = ws.c_p + (long long)t * E; g.epi.ld_out0 = s1 * E;
    g.epi.out1 = ws.h_p + (long long)t * E; g.epi.ld_out1 = s1 * E;
    ACVAE_TRY(launch_gemm<EPI_LSTM>(g, sp));
    GemmParams h{};
    h.M = N; h.U = E; h.G = 2; h.nseg = 1;
    h.seg[0] = seg_gates(ws.h_p + (long long)t * E, s1 * E, w.p_head_w, E, 0, E, E, 2);
    h.epi.bias[0] = w.p_head_b; h.epi.bias[1] = w.p_head_b + E;
    h.epi.eps = io.eps_p + (long long)t * N * E; h.epi.ld_eps = E;
    h.epi.out0 = io.p_means + (long long)t * E; h.epi.ld_out0 = s1 * E;
    h.epi.out1 = io.p_logs + (long long)t * E; h.epi.ld_out1 = s1 * E;
    h.epi.out2 = io.p_z + (long long)t * E; h.epi.ld_out2 = s1 * E;
    ACVAE_TRY(launch_gemm<EPI_HEAD>(h, sp));
  }

  // ---- decoder (decoder.py:175-203): needs q_z from the posterior -------------------------------------------
  ACVAE_TRY(gather_rows(NT, E, w.d_emb, ws.words, ws.xd, st));
  ACVAE_CHECK(cudaStreamWaitEvent(st, ev_qz, 0));
  {
    // gx_d = [emb | q_z] . W_ih[:, {0:E, 2E:3E}]^T + b_ih  (3E columns), kept in dgi_d until the backward overwrites it
    GemmParams g{};
    g.M = NT; g.U = 3 * E; g.G = 1; g.nseg = 2;
    g.seg[0] = seg_plain(ws.xd, E, w.d_wih, 3 * E, E);
    g.seg[1] = seg_plain(io.q_z, E, w.d_wih + 2 * E, 3 * E, E);
    g.epi.c[0] = ws.dgi_d; g.epi.ldc = 3 * E; g.epi.bias[0] = w.d_bih; g.epi.scale = 1.0f;
    ACVAE_TRY(launch_gemm<EPI_PLAIN>(g, st));
  }
  if (cl) {
    DecClFwd dc{};
    dc.N = N; dc.T = T; dc.Te = Te; dc.gx = ws.dgi_d; dc.attn_w = w.d_attn_w; dc.attn_v = w.d_attn_v; dc.whh = w.d_whh; dc.bhh = w.d_bhh;
    dc.Pd = ws.Pd; dc.Mg = ws.Mg; dc.mem_lens = io.mem_lens; dc.qp = ws.qp_d; dc.w = ws.w_d; dc.gates = ws.gates_d; dc.out = io.outputs;
    dc.aw = io.attn_weights; dc.trace = chain_trace_ptr();
    // The two chains are independent (dis_ratio == 0) and share the machine: a cluster-chain CTA needs a WHOLE SM (255
    // registers, 180 KB), a prior CTA half of one.  Order matters: when the prior's 128 CTAs come first they take one SM
    // each and the decoder's clusters trickle in two at a time on the 20 SMs left (214 instead of 108 us); when the
    // decoder's 8 clusters come first they take 64 SMs and the prior packs two CTAs per SM onto the other 84.  So the
    // prior waits for the decoder's inputs and is held back a few microseconds behind the decoder's launch.
    ACVAE_TRY(stream_dep(s_mg, st, ax));
    // Scheduled sampling: a free step t (tf_flags[t] == 0) is fed the arg-max word of step t-1 (vae_model.py:826-832), which
    // exists only once the decoder has produced step t-1.  Both chains are cut there: vocabulary arg-max of the N rows of step
    // t-1, then the word-dependent hoisted inputs of step t alone (embedding rows, the prior's query projection / attention /
    // input gates, the decoder's input gates), then the chains resume from their saved state.  Teacher-forced steps keep the
    // batched inputs computed above; with every step teacher-forced this is one segment = the schedule of the benchmark.
    int seg_begin = 0;
    for (int seg_end = 1; seg_end <= T; ++seg_end) {
      if (seg_end < T && io.tf_flags[seg_end]) continue;
      const int a = seg_begin;
      if (a > 0) {
        VocabStatsArgs v{};
        v.M = N; v.V = d.V; v.E = E; v.hidden = io.outputs + (long long)(a - 1) * E; v.ld_h = (long long)T * E;
        v.cls_w = w.cls_w; v.cls_b = w.cls_b;
        v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
        v.red.lse = io.logit_lse + (a - 1); v.red.lsum = io.logit_sum + (a - 1); v.red.logprob = io.sampled_logprobs + (a - 1);
        v.red.ld_row = T; v.red.seqs = (long long*)io.seqs + (a - 1); v.red.ld_seqs = T;
        v.red.next_word = ws.words + a; v.red.ld_next = T;
        ACVAE_TRY(vocab_stats(v, st));
        ACVAE_LAUNCH(gather_step2_kernel, grid1d((long long)N * E), 256, 0, st, N, T, a, E, w.d_emb, w.p_emb, (const int*)ws.words,
                     ws.xd, ws.xp);
        ACVAE_TRY(stream_dep(st, sp, ax));
        {   // decoder input gates of step a
          GemmParams g{};
          g.M = N; g.U = 3 * E; g.G = 1; g.nseg = 2;
          g.seg[0] = seg_plain(ws.xd + (long long)a * E, s1 * E, w.d_wih, 3 * E, E);
          g.seg[1] = seg_plain(io.q_z + (long long)a * E, s1 * E, w.d_wih + 2 * E, 3 * E, E);
          g.epi.c[0] = ws.dgi_d + (long long)a * 3 * E; g.epi.ldc = s1 * 3 * E; g.epi.bias[0] = w.d_bih; g.epi.scale = 1.0f;
          ACVAE_TRY(launch_gemm<EPI_PLAIN>(g, st));
        }
        {   // prior: query projection, word attention, input gates of step a
          ACVAE_TRY(linear_fwd(N, E, E, ws.xp + (long long)a * E, s1 * E, w.p_attn_w, 2 * E, nullptr, ws.qp_p + (long long)a * E, s1 * E, sp));
          AttnFwdParams at{};
          at.rows = N; at.Te = Te; at.A = E; at.E = E; at.Dq = E; at.rows_per_clip = 1;
          at.qp_in = ws.qp_p + (long long)a * E; at.ld_qp_in = s1 * E;
          at.P = ws.Pp; at.mem = ws.mem; at.v = w.p_attn_v; at.mem_lens = io.mem_lens;
          at.ctx = ws.ctx_p + (long long)a * E; at.ld_ctx = s1 * E; at.w_out = ws.w_p + (long long)a * Te; at.ld_w = s1 * Te;
          ACVAE_TRY(launch_attn_fwd(at, sp));
          GemmParams g{};
          g.M = N; g.U = 4 * E; g.G = 1; g.nseg = 2;
          g.seg[0] = seg_plain(ws.xp + (long long)a * E, s1 * E, w.p_wih, 3 * E, E);
          g.seg[1] = seg_plain(ws.ctx_p + (long long)a * E, s1 * E, w.p_wih + E, 3 * E, E);
          g.epi.c[0] = ws.dg_p + (long long)a * 4 * E; g.epi.ldc = s1 * 4 * E; g.epi.bias[0] = w.p_bih; g.epi.scale = 1.0f;
          ACVAE_TRY(launch_gemm<EPI_PLAIN>(g, sp));
        }
      }
      dc.t0 = a; dc.t1 = seg_end;
      bool seg_dis = false;
      for (int t = a; t < seg_end; ++t) seg_dis = seg_dis || io.dis_flags[t];
      if (seg_dis) {
        // the decoder consumes the prior's sample at some step of this segment (vae_model.py:800-806): the prior's segment
        // first, then the decoder's input gates of those steps from p_z, then the decoder's segment
        ppc.t0 = a; ppc.t1 = seg_end;
        if (a > 0) ACVAE_CHECK(cudaMemsetAsync(ppc.bar, 0, 128 * sizeof(unsigned), sp));
        ACVAE_TRY(launch_chain(prior_chain_fwd_kernel, 0, sp, "prior_chain_fwd_kernel", ppc));
        ACVAE_TRY(stream_dep(sp, st, ax));
        for (int t = a; t < seg_end; ++t) {
          if (!io.dis_flags[t]) continue;
          GemmParams g{};
          g.M = N; g.U = 3 * E; g.G = 1; g.nseg = 2;
          g.seg[0] = seg_plain(ws.xd + (long long)t * E, s1 * E, w.d_wih, 3 * E, E);
          g.seg[1] = seg_plain(io.p_z + (long long)t * E, s1 * E, w.d_wih + 2 * E, 3 * E, E);
          g.epi.c[0] = ws.dgi_d + (long long)t * 3 * E; g.epi.ldc = s1 * 3 * E; g.epi.bias[0] = w.d_bih; g.epi.scale = 1.0f;
          ACVAE_TRY(launch_gemm<EPI_PLAIN>(g, st));
        }
        ACVAE_TRY(launch_cluster_chain(dec_cl_fwd_kernel, dec_cl_clusters(N), dec_cl_fwd_smem(Te), st, "dec_cl_fwd_kernel", dc));
        seg_begin = seg_end;
        continue;
      }
      if (prior_chain) ACVAE_TRY(stream_dep(st, sp, ax));
      ACVAE_TRY(launch_cluster_chain(dec_cl_fwd_kernel, dec_cl_clusters(N), dec_cl_fwd_smem(Te), st, "dec_cl_fwd_kernel", dc));
      if (prior_chain) {
        ppc.t0 = a; ppc.t1 = seg_end;
        if (a > 0) ACVAE_CHECK(cudaMemsetAsync(ppc.bar, 0, 128 * sizeof(unsigned), sp));     // the grid barrier counts from zero
        ACVAE_LAUNCH(stream_delay_kernel, 1, 1, 0, sp, 8000u);
        ACVAE_TRY(launch_chain(prior_chain_fwd_kernel, 0, sp, "prior_chain_fwd_kernel", ppc));
      }
      seg_begin = seg_end;
    }
    // the context itself (weight gradients, rnn_input) from the saved weights, off the critical stream
    ACVAE_TRY(stream_dep(st, sq1, ax));
    ACVAE_LAUNCH(attn_ctx_kernel, dim3(N, T), 256, 0, sq1, T, Te, E, (const float*)ws.w_d, (const float*)ws.mem, io.mem_lens, ws.ctx_d);
  } else if (coop) {
    DecChainFwd dc{};
    dc.N = N; dc.T = T; dc.Te = Te; dc.gx = ws.dgi_d; dc.attn_w = w.d_attn_w; dc.attn_v = w.d_attn_v;
    dc.wih = w.d_wih; dc.whh = w.d_whh; dc.bhh = w.d_bhh; dc.Pd = ws.Pd; dc.mem = ws.mem; dc.mem_lens = io.mem_lens;
    dc.qp = ws.qp_d; dc.w = ws.w_d; dc.ctx = ws.ctx_d; dc.gates = ws.gates_d; dc.out = io.outputs; dc.aw = io.attn_weights;
    dc.bar = ws.bars + 2 * 128; dc.part = ws.part_d; dc.trace = chain_trace_ptr();
    ACVAE_TRY(stream_dep(sp, st, ax));      // the prior's hoisted inputs (gx_p) are produced on sp
    ACVAE_TRY(launch_chain(dec_chain_fwd_kernel, dec_chain_fwd_smem(Te), st, "dec_chain_fwd_kernel", dc, ppc));
  }
  for (int t = 0; t < T && !dec_chain; ++t) {
    const float* hprev = t > 0 ? io.outputs + (long long)(t - 1) * E : nullptr;
    GemmParams qg{};
    qg.M = N; qg.U = A; qg.G = 1; qg.nseg = hprev ? 1 : 0;
    if (hprev) qg.seg[0] = seg_plain(hprev, s1 * E, w.d_attn_w, 2 * E, E);
    qg.epi.c[0] = ws.qp_d + (long long)t * A; qg.epi.ldc = s1 * A; qg.epi.scale = 1.0f;
    ACVAE_TRY(launch_gemm<EPI_PLAIN>(qg, st));
    AttnFwdParams a{};
    a.rows = N; a.Te = Te; a.A = A; a.E = E; a.Dq = E; a.rows_per_clip = 1;
    a.qp_in = ws.qp_d + (long long)t * A; a.ld_qp_in = s1 * A;
    a.P = ws.Pd; a.mem = ws.mem; a.v = w.d_attn_v; a.mem_lens = io.mem_lens;
    a.ctx = ws.ctx_d + (long long)t * E; a.ld_ctx = s1 * E;
    a.w_out = ws.w_d + (long long)t * Te; a.ld_w = s1 * Te;
    if (io.attn_weights) { a.aw_out = io.attn_weights + t; a.aw_ld_r = (long long)Te * T; a.aw_ld_j = T; }
    ACVAE_TRY(launch_attn_fwd(a, st));
    GemmParams g{};
    g.M = N; g.U = E; g.G = 4;
    int ns = 0;
    {
      GemmSeg s = seg_gates(ws.ctx_d + (long long)t * E, s1 * E, w.d_wih, 3 * E, E, E, E, 3);
      s.w[3] = nullptr;
      g.seg[ns++] = s;
    }
    if (hprev) {
      GemmSeg s = seg_gates(hprev, s1 * E, w.d_whh, E, 0, E, E, 3);
      s.w[3] = s.w[2]; s.w[2] = nullptr;
      g.seg[ns++] = s;
    }
    g.nseg = ns;
    g.epi.gx = ws.dgi_d + (long long)t * 3 * E; g.epi.ld_gx = s1 * 3 * E;
    g.epi.b_hh = w.d_bhh;
    g.epi.prev = hprev; g.epi.ld_prev = s1 * E;
    g.epi.gates = ws.gates_d + (long long)t * 4 * E; g.epi.ld_gates = s1 * 4 * E;
    g.epi.out0 = io.outputs + (long long)t * E; g.epi.ld_out0 = s1 * E;
    ACVAE_TRY(launch_gemm<EPI_GRU>(g, st));
  }
  // vocabulary statistics over all rows (greedy word, lse, sum; logits never stored)
  {
    VocabStatsArgs v{};
    v.M = NT; v.V = d.V; v.E = E; v.hidden = io.outputs; v.ld_h = E;
    v.cls_w = w.cls_w; v.cls_b = w.cls_b;
    v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
    v.red.lse = io.logit_lse; v.red.lsum = io.logit_sum; v.red.logprob = io.sampled_logprobs; v.red.ld_row = 1;
    v.red.seqs = (long long*)io.seqs; v.red.ld_seqs = 1;
    ACVAE_TRY(vocab_stats(v, st));
  }
  // global-constraint head (vae_model.py:722-729): needs only the decoder's outputs -- next to the vocabulary GEMM, not behind it
  cudaStream_t s_g = cl ? sq1 : st;
  ACVAE_LAUNCH(pool_fwd_kernel, grid1d((long long)N * E), 256, 0, s_g, N, T, E, io.outputs, ws.steplens, 0, ws.pool_d,
               ws.amax_d);
  ACVAE_TRY(linear_fwd(N, 2 * E, E, ws.pool_d, E, w.g_w, E, w.g_b, io.p_means_utt, 2 * E, s_g));
  if (io.logits) ACVAE_TRY(linear_fwd(NT, d.V, E, io.outputs, E, w.cls_w, E, w.cls_b, io.logits, d.V, st));
  ACVAE_TRY(stream_dep(sp, st, ax));
  ACVAE_TRY(stream_dep(sq0, st, ax));            // the posterior's pooling
  if (cl) ACVAE_TRY(stream_dep(sq1, st, ax));
  ACVAE_TRY(stream_dep(st, st_user, ax));
  return 0;
}

// k-blocks per CTA of the weight-gradient GEMMs on the fan streams (see TcThroughputScope).  A tc_gemm CTA owns its SM
// (197 KB of shared memory, 62 K registers) and spends ~3 us in prologue + epilogue whatever its K range, so deep split-K
// wastes SM time the backward is short of next to the chains, while no split leaves the K = N*Te contractions as 60 us
// stragglers: measured step 0.912 / 0.902 / 0.938 / 0.975 ms at 6 / 12 / 24 / 64 (profiles/r2/fan_min_kblk.log).
inline int fan_min_kblk() {
  static int v = 0;
  if (!v) { const char* e = getenv("ACVAE_FAN_MIN_KBLK"); v = e ? atoi(e) : 12; if (v < 1) v = 1; }
  return v;
}

// ===================================== backward ===================================================
inline int train_bwd_fast(const acvae_dims& d, const acvae_weights& w, const acvae_train_io& io,
                          const acvae_train_grads_in& gi, acvae_weight_grads& gw, float* d_audio, void* workspace,
                          cudaStream_t st_user) {
  TrainWs ws = carve_train_ws(d, workspace);
  Aux* ax = aux();
  cudaStream_t st = ax->s[kAuxMain];          // high-priority critical-path stream (see train_fwd_fast)
  ACVAE_TRY(stream_dep(st_user, st, ax));
  const int N = d.N, T = d.T, E = d.E, A = d.A, Te = d.Te, NT = N * T, V = d.V;
  const long long s1 = T;
  cudaStream_t sp = ax->s[2], sx = ax->s[3], sq0 = ax->s[0], sq1 = ax->s[1];
  auto zero = [&](float* p, size_t n, cudaStream_t s) { return cudaMemsetAsync(p, 0, n * sizeof(float), s); };
  const bool cl = cluster_chain_supported(N, T, Te, E, A);   // posterior / decoder chains on clusters (cluster_chain.cuh)
  const bool coop = chain_supported(N, T, Te, E, A);         // cooperative-grid chains (recurrent.cuh)
  const bool post_chain = cl || coop, dec_chain = cl || coop;
  // cooperative chains only: the prior's backward chain runs inside the decoder's persistent kernel (same barriers), so
  // the two chains cost max(...) instead of their sum (two cooperative kernels never overlap on the device).  Next to the
  // cluster decoder chain the prior is its own kernel on its own stream.
  static int merge_env = -1;
  if (merge_env < 0) { const char* e = getenv("ACVAE_MERGE_BWD"); merge_env = (e && e[0] == '0') ? 0 : 1; }
  const bool merge_bwd = !cl && coop && merge_env == 1;
  if (coop) ACVAE_CHECK(cudaMemsetAsync(ws.bars + 4 * 128, 0, 4 * 128 * sizeof(unsigned), st));
  // prior-z replacement (dis_flags, vae_model.py:800-806): the decoder's d z goes to the prior at those steps, so the prior's
  // backward chain waits for the decoder's; zin = the z each decoder step consumed (operand of the z block of d W_ih)
  bool any_dis = false;
  for (int t = 0; t < T; ++t) any_dis = any_dis || io.dis_flags[t];
  const unsigned long long dis_mask = flag_mask(io.dis_flags, T);
  const float* zin = io.q_z;
  if (any_dis) {
    ACVAE_LAUNCH(zsel_kernel, grid1d((long long)NT * E), 256, 0, st, (long long)NT * E, T, E, dis_mask, io.q_z, io.p_z, ws.zsel);
    zin = ws.zsel;
  }
  // (Starting the prior's backward chain earlier -- right behind the KL gradients, under the cross-entropy gradient GEMMs --
  // was tried: the cooperative kernel holds all 148 SMs, the GEMMs it overlaps take 130 instead of 27 us and the step gets
  // 30 us longer, gpurun_out r2f timeline.)
  ACVAE_TRY(stream_dep(st, sp, ax));

  // ================= prior BPTT on its own stream (KL gradients only: dis_ratio == 0) ====================
  PriorChainBwd ppc{};
  if (coop) {
    ppc.N = N; ppc.T = T; ppc.d_pz = gi.d_p_z; ppc.d_pm = gi.d_p_means; ppc.d_pl = gi.d_p_logs; ppc.eps = io.eps_p;
    ppc.p_logs = io.p_logs; ppc.head_w = w.p_head_w; ppc.wih = w.p_wih; ppc.whh = w.p_whh; ppc.gates = ws.gates_p; ppc.c = ws.c_p;
    ppc.dml = ws.dml_p; ppc.dg = ws.dg_p; ppc.bar = ws.bars + 4 * 128;
    if (!merge_bwd && !cl) ACVAE_TRY(launch_chain(prior_chain_bwd_kernel, 0, sp, "prior_chain_bwd_kernel", ppc));
  } else
  {
    // step T-1 head backward (standalone), then per step: [dh GEMM + LSTM pointwise] -> [dz|dh GEMM + head pointwise]
    HeadBwdParams h{};
    h.rows = N; h.U = E;
    const int t = T - 1;
    if (gi.d_p_z) { h.dz0 = gi.d_p_z + (long long)t * E; h.ld_dz0 = s1 * E; }
    if (gi.d_p_means) { h.dmean = gi.d_p_means + (long long)t * E; h.ld_dmean = s1 * E; }
    if (gi.d_p_logs) { h.dlog = gi.d_p_logs + (long long)t * E; h.ld_dlog = s1 * E; }
    h.eps = io.eps_p + (long long)t * N * E; h.ld_eps = E;
    h.logv = io.p_logs + (long long)t * E; h.ld_logv = s1 * E;
    h.dml = ws.dml_p + (long long)t * 2 * E; h.ld_dml = s1 * 2 * E;
    ACVAE_LAUNCH(head_bwd_kernel, grid1d((long long)N * E), 256, 0, sp, h);
  }
  for (int t = T - 1; t >= 0 && !coop; --t) {
    GemmParams p{};
    p.M = N; p.U = E; p.G = 1; p.nseg = 1;
    GemmSeg s{};
    s.a = ws.dml_p + (long long)t * 2 * E; s.lda = s1 * 2 * E; s.w[0] = w.p_head_w; s.ldw = E; s.w_trans = 1; s.K = 2 * E;
    p.seg[0] = s;
    p.epi.x0 = t < T - 1 ? ws.dhp_carry : nullptr;
    p.epi.x1 = t < T - 1 ? ws.dcp_carry : nullptr;
    p.epi.gates = ws.gates_p + (long long)t * 4 * E; p.epi.ld_gates = s1 * 4 * E;
    p.epi.x2 = ws.c_p + (long long)t * E; p.epi.ld_x2 = s1 * E;
    p.epi.x3 = t > 0 ? ws.c_p + (long long)(t - 1) * E : nullptr; p.epi.ld_x3 = s1 * E;
    p.epi.y0 = ws.dg_p + (long long)t * 4 * E; p.epi.ld_y0 = s1 * 4 * E;
    p.epi.y1 = ws.dcp_carry;
    ACVAE_TRY(launch_gemm<EPI_LSTM_BWD>(p, sp));
    if (t > 0) {
      GemmParams q{};
      q.M = N; q.U = E; q.G = 2; q.nseg = 1;
      GemmSeg s2{};
      s2.a = ws.dg_p + (long long)t * 4 * E; s2.lda = s1 * 4 * E; s2.K = 4 * E; s2.w_trans = 1;
      s2.w[0] = w.p_wih + 2 * E; s2.ldwg[0] = 3 * E;      // d last_z
      s2.w[1] = w.p_whh; s2.ldwg[1] = E;                  // d h_{t-1}
      s2.ldw = E;
      q.seg[0] = s2;
      const int tm = t - 1;
      if (gi.d_p_z) { q.epi.x0 = gi.d_p_z + (long long)tm * E; q.epi.ld_x0 = s1 * E; }
      if (gi.d_p_means) { q.epi.x1 = gi.d_p_means + (long long)tm * E; q.epi.ld_x1 = s1 * E; }
      if (gi.d_p_logs) { q.epi.x2 = gi.d_p_logs + (long long)tm * E; q.epi.ld_x2 = s1 * E; }
      q.epi.x3 = io.eps_p + (long long)tm * N * E; q.epi.ld_x3 = E;
      q.epi.x4 = io.p_logs + (long long)tm * E; q.epi.ld_x4 = s1 * E;
      q.epi.y0 = ws.dml_p + (long long)tm * 2 * E; q.epi.ld_y0 = s1 * 2 * E;
      q.epi.y1 = ws.dhp_carry;
      ACVAE_TRY(launch_gemm<EPI_HEAD_BWD>(q, sp));
    }
  }
  // prior batched remainders (after the prior chain: at once in the launch-per-step schedule, after the merged
  // decoder+prior kernel in chain mode)
  cudaEvent_t ev_dec_mem = ax->ev();        // recorded on sx once the decoder's half of the memory backward is complete
  auto prior_remainders = [&]() -> int {
    // weight / bias gradients that need only the chain's dg_p / dml_p: fanned over four streams so that they fill
    // the SMs the decoder's persistent kernel leaves free instead of queueing behind the attention backward
    {
      cudaStream_t* f = &ax->s[kAuxPriorFan0];
      TcThroughputScope throughput(fan_min_kblk());
      for (int i = 0; i < 6; ++i) ACVAE_TRY(stream_dep(sp, f[i], ax));
      ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, ws.xp, E, gw.p_wih, 3 * E, f[0]));
      ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, ws.ctx_p, E, gw.p_wih + E, 3 * E, f[1]));
      ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, io.p_z - E, E, gw.p_wih + 2 * E, 3 * E, f[2], T, 0, -1));
      ACVAE_TRY(linear_bwd_weight(4 * E, E, NT, ws.dg_p, 4 * E, ws.h_p - E, E, gw.p_whh, E, f[3], T, 0, -1));
      ACVAE_TRY(linear_bwd_weight(2 * E, E, NT, ws.dml_p, 2 * E, ws.h_p, E, gw.p_head_w, E, f[4]));
      ACVAE_TRY(colsum(NT, 4 * E, ws.dg_p, 4 * E, gw.p_bih, f[5]));
      ACVAE_CHECK(cudaMemcpyAsync(gw.p_bhh, gw.p_bih, sizeof(float) * 4 * E, cudaMemcpyDeviceToDevice, f[5]));
      ACVAE_TRY(colsum(NT, 2 * E, ws.dml_p, 2 * E, gw.p_head_b, f[5]));
      // d xe_p, the half that needs only the chain (the attention-query half is added behind the attention backward)
      ACVAE_TRY(stream_dep(sp, f[6], ax));
      ACVAE_TRY(linear_bwd_data(NT, E, 4 * E, ws.dg_p, 4 * E, w.p_wih, 3 * E, ws.dxe_p, E, f[6]));
      ACVAE_CHECK(zero(gw.p_emb, (size_t)V * E, f[6]));
    }
    // critical first: d ctx -> attention backward -> per-clip accumulation (the memory backward waits for it)
    ACVAE_TRY(linear_bwd_data(NT, E, 4 * E, ws.dg_p, 4 * E, w.p_wih + E, 3 * E, ws.dctx_p, E, sp));
    ACVAE_CHECK(zero(gw.p_attn_v, E, sp));
    {
      // all T rows of a clip in one pass (d score, d qp, d P, d mem, d v); the row-per-CTA + per-clip-accumulation pair
      // for the shapes the fused kernel does not cover
      AttnBwdClipParams c{};
      c.clips = N; c.T = T; c.Te = Te; c.A = E; c.E = E; c.dctx = ws.dctx_p; c.w = ws.w_p; c.qp = ws.qp_p; c.P = ws.Pp; c.mem = ws.mem;
      c.v = w.p_attn_v; c.mem_lens = io.mem_lens; c.ds_out = ws.ds_p; c.dqp = ws.dqp_p; c.dP = ws.dPp; c.dmem = ws.dmem;
      c.dmem_accumulate = 0; c.dv = gw.p_attn_v;
      const int fused = launch_attn_bwd_clip(c, sp);
      if (fused < 0) return fused;
      if (!fused) {
        AttnBwdQParams a{};
        a.rows = NT; a.Te = Te; a.A = E; a.E = E; a.rows_per_clip = T;
        a.dctx = ws.dctx_p; a.ld_dctx = E; a.w = ws.w_p; a.ld_w = Te; a.qp = ws.qp_p; a.ld_qp = E;
        a.P = ws.Pp; a.mem = ws.mem; a.v = w.p_attn_v; a.mem_lens = io.mem_lens;
        a.ds = ws.ds_p; a.ld_ds = Te; a.dqp = ws.dqp_p; a.ld_dqp = E;
        ACVAE_TRY(launch_attn_bwd_q(a, sp));
        AttnBwdAccParams b{};
        b.clips = N; b.Te = Te; b.A = E; b.E = E; b.rows_per_clip = T;
        b.ds = ws.ds_p; b.ld_ds = Te; b.w = ws.w_p; b.ld_w = Te; b.qp = ws.qp_p; b.ld_qp = E;
        b.dctx = ws.dctx_p; b.ld_dctx = E; b.P = ws.Pp; b.v = w.p_attn_v; b.mem_lens = io.mem_lens;
        b.dP = ws.dPp; b.dmem = ws.dmem; b.dmem_accumulate = 0; b.dv = gw.p_attn_v;
        ACVAE_TRY(launch_attn_bwd_acc(b, sp));
      }
    }
    // the rest: embedding / attention-query weight gradients (need dqp_p from the attention backward); off sp, which
    // carries the prior's half of the memory backward next (the step's last dependency chain)
    {
      cudaStream_t fe = ax->s[kAuxPriorFan0 + 6], fw = ax->s[kAuxPriorFan0 + 7];
      TcThroughputScope throughput(fan_min_kblk());
      ACVAE_TRY(stream_dep(sp, fe, ax));
      ACVAE_TRY(stream_dep(sp, fw, ax));
      ACVAE_TRY(linear_bwd_data(NT, E, E, ws.dqp_p, E, w.p_attn_w, 2 * E, ws.dxe_p, E, fe, 1));
      ACVAE_TRY(scatter_rows(NT, E, ws.dxe_p, E, ws.words, gw.p_emb, fe));
      ACVAE_TRY(linear_bwd_weight(E, E, NT, ws.dqp_p, E, ws.xp, E, gw.p_attn_w, 2 * E, fw));
    }
    return 0;
  };
  if (!merge_bwd && !cl) ACVAE_TRY(prior_remainders());

  // ================= decoder BPTT on the main stream ========================================================
  const float* dpool = nullptr;
  if (gi.d_p_means_utt) {
    ACVAE_TRY(linear_bwd_data(N, E, 2 * E, gi.d_p_means_utt, 2 * E, w.g_w, E, ws.dpool, E, st));
    {   // the head's own weight / bias gradients feed nothing in the step: off the critical stream (joined with the fan at the end)
      cudaStream_t fg = ax->s[kAuxFan0 + 7];
      ACVAE_TRY(stream_dep(st_user, fg, ax));
      ACVAE_TRY(linear_bwd_weight(2 * E, E, N, gi.d_p_means_utt, 2 * E, ws.pool_d, E, gw.g_w, E, fg));
      ACVAE_TRY(colsum(N, 2 * E, gi.d_p_means_utt, 2 * E, gw.g_b, fg));
    }
    dpool = ws.dpool;
  } else {
    ACVAE_CHECK(zero(gw.g_w, (size_t)2 * E * E, st)); ACVAE_CHECK(zero(gw.g_b, (size_t)2 * E, st));
    ACVAE_TRY(stream_dep(st_user, ax->s[kAuxFan0 + 7], ax));     // every fan stream is forked before the join below
  }
  ACVAE_LAUNCH(pool_bwd_kernel, grid1d((long long)NT * E), 256, 0, st, N, T, E, dpool, ws.steplens, 0, ws.amax_d,
               gi.d_outputs, ws.dout);
  if (cl) {
    DecClBwd dc{};
    dc.N = N; dc.T = T; dc.Te = Te; dc.dout = ws.dout; dc.attn_w = w.d_attn_w; dc.attn_v = w.d_attn_v; dc.whh = w.d_whh;
    dc.Pd = ws.Pd; dc.Mg = ws.Mg; dc.mem_lens = io.mem_lens; dc.qp = ws.qp_d; dc.w = ws.w_d; dc.gates = ws.gates_d; dc.out = io.outputs;
    dc.dgi = ws.dgi_d; dc.dgh = ws.dgh_d; dc.ds = ws.ds_d; dc.dqp = ws.dqp_d;
    // decoder first, the prior's cooperative chain a few microseconds behind it on its own stream (see train_fwd_fast)
    if (coop) ACVAE_TRY(stream_dep(st, sp, ax));
    ACVAE_TRY(launch_cluster_chain(dec_cl_bwd_kernel, dec_cl_clusters(N), dec_cl_bwd_smem(Te), st, "dec_cl_bwd_kernel", dc));
    if (coop && !any_dis) {
      ACVAE_LAUNCH(stream_delay_kernel, 1, 1, 0, sp, 8000u);
      ACVAE_TRY(launch_chain(prior_chain_bwd_kernel, 0, sp, "prior_chain_bwd_kernel", ppc));
    }
    if (!any_dis) ACVAE_TRY(prior_remainders());
  } else if (coop) {
    DecChainBwd dc{};
    dc.N = N; dc.T = T; dc.Te = Te; dc.dout = ws.dout; dc.attn_w = w.d_attn_w; dc.attn_v = w.d_attn_v; dc.wih = w.d_wih;
    dc.whh = w.d_whh; dc.Pd = ws.Pd; dc.mem = ws.mem; dc.mem_lens = io.mem_lens; dc.qp = ws.qp_d; dc.w = ws.w_d;
    dc.gates = ws.gates_d; dc.out = io.outputs; dc.dgi = ws.dgi_d; dc.dgh = ws.dgh_d; dc.dctx = ws.dctx_d; dc.ds = ws.ds_d;
    dc.dqp = ws.dqp_d; dc.bar = ws.bars + 5 * 128;
    dc.trace = chain_trace_ptr() ? chain_trace_ptr() + (long long)T * 16 : nullptr;
    if (merge_bwd) {
      ACVAE_TRY(launch_chain(dec_chain_bwd_kernel<true>, dec_chain_bwd_smem(Te), st, "dec_chain_bwd_kernel", dc, ppc));
      ACVAE_TRY(stream_dep(st, sp, ax));
      ACVAE_TRY(prior_remainders());
    } else {
      PriorChainBwd none{};
      ACVAE_TRY(launch_chain(dec_chain_bwd_kernel<false>, dec_chain_bwd_smem(Te), st, "dec_chain_bwd_kernel", dc, none));
    }
  } else
  {
    GruBwdParams g{};
    const int t = T - 1;
    g.N = N; g.U = E;
    g.dh_ext = ws.dout + (long long)t * E; g.ld_dh_ext = s1 * E;
    g.gates = ws.gates_d + (long long)t * 4 * E; g.ld_gates = s1 * 4 * E;
    g.hprev = t > 0 ? io.outputs + (long long)(t - 1) * E : nullptr; g.ld_hprev = s1 * E;
    g.dgi = ws.dgi_d + (long long)t * 3 * E; g.ld_dgi = s1 * 3 * E;
    g.dgh = ws.dgh_d + (long long)t * 3 * E; g.ld_dgh = s1 * 3 * E;
    g.dh_out = ws.dh_carry;
    ACVAE_LAUNCH(gru_bwd_kernel, grid1d((long long)N * E), 256, 0, st, g);
  }
  for (int t = T - 1; t >= 0 && !dec_chain; --t) {
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
      // dh_{t-1} = dh_t*z_t (carry) + dGh_t.W_hh + dqp_t.Wq + upstream; fused GRU pointwise backward of step t-1
      GemmParams p{};
      p.M = N; p.U = E; p.G = 1; p.nseg = 2;
      GemmSeg s0{}; s0.a = ws.dgh_d + (long long)t * 3 * E; s0.lda = s1 * 3 * E; s0.w[0] = w.d_whh; s0.ldw = E; s0.w_trans = 1; s0.K = 3 * E;
      GemmSeg s1_{}; s1_.a = ws.dqp_d + (long long)t * A; s1_.lda = s1 * A; s1_.w[0] = w.d_attn_w; s1_.ldw = 2 * E; s1_.w_trans = 1; s1_.K = A;
      p.seg[0] = s0; p.seg[1] = s1_;
      const int tm = t - 1;
      p.epi.x0 = ws.dh_carry;
      p.epi.x1 = ws.dout + (long long)tm * E; p.epi.ld_x1 = s1 * E;
      p.epi.gates = ws.gates_d + (long long)tm * 4 * E; p.epi.ld_gates = s1 * 4 * E;
      p.epi.prev = tm > 0 ? io.outputs + (long long)(tm - 1) * E : nullptr; p.epi.ld_prev = s1 * E;
      p.epi.y0 = ws.dgi_d + (long long)tm * 3 * E; p.epi.ld_y0 = s1 * 3 * E;
      p.epi.y1 = ws.dgh_d + (long long)tm * 3 * E; p.epi.ld_y1 = s1 * 3 * E;
      p.epi.y2 = ws.dh_carry;
      ACVAE_TRY(launch_gemm<EPI_GRU_BWD>(p, st));
    }
  }
  // d z fed to the decoder -> d q_z (needed by the posterior backward): first thing after the chain
  ACVAE_TRY(linear_bwd_data(NT, E, 3 * E, ws.dgi_d, 3 * E, w.d_wih + 2 * E, 3 * E, ws.dxz_d, E, st));
  if (cl && any_dis) {
    // the prior's backward chain behind the decoder's: d p_z = upstream + the decoder's d z at the dis steps
    ACVAE_TRY(stream_dep(st, sp, ax));
    ACVAE_LAUNCH(dpz_kernel, grid1d((long long)NT * E), 256, 0, sp, (long long)NT * E, T, E, dis_mask, gi.d_p_z, (const float*)ws.dxz_d,
                 ws.dpz);
    ppc.d_pz = ws.dpz;
    ACVAE_TRY(launch_chain(prior_chain_bwd_kernel, 0, sp, "prior_chain_bwd_kernel", ppc));
    ACVAE_TRY(prior_remainders());
  }
  ACVAE_TRY(stream_dep(st, sx, ax));
  // ---- side stream sx (overlaps the posterior chains) ----
  // critical first: the decoder's per-clip attention accumulation (into its own buffer: no ordering against the
  // prior's), then the memory backward as soon as the prior's accumulation is done too
  // cluster chain: d ctx of all steps in one batched contraction (the chain itself works on d alpha directly)
  if (cl) ACVAE_TRY(linear_bwd_data(NT, E, 3 * E, ws.dgi_d, 3 * E, w.d_wih + E, 3 * E, ws.dctx_d, E, sx));
  ACVAE_CHECK(zero(gw.d_attn_v, A, sx));
  {
    AttnBwdClipParams c{};
    c.clips = N; c.T = T; c.Te = Te; c.A = A; c.E = E; c.dctx = ws.dctx_d; c.w = ws.w_d; c.qp = ws.qp_d; c.P = ws.Pd; c.mem = ws.mem;
    c.v = w.d_attn_v; c.mem_lens = io.mem_lens; c.ds_in = ws.ds_d; c.dP = ws.dPd; c.dmem = ws.dmem2; c.dmem_accumulate = 0; c.dv = gw.d_attn_v;
    const int fused = launch_attn_bwd_clip(c, sx);
    if (fused < 0) return fused;
    if (!fused) {
      AttnBwdAccParams a{};
      a.clips = N; a.Te = Te; a.A = A; a.E = E; a.rows_per_clip = T;
      a.ds = ws.ds_d; a.ld_ds = Te; a.w = ws.w_d; a.ld_w = Te; a.qp = ws.qp_d; a.ld_qp = A;
      a.dctx = ws.dctx_d; a.ld_dctx = E; a.P = ws.Pd; a.v = w.d_attn_v; a.mem_lens = io.mem_lens;
      a.dP = ws.dPd; a.dmem = ws.dmem2; a.dmem_accumulate = 0; a.dv = gw.d_attn_v;
      ACVAE_TRY(launch_attn_bwd_acc(a, sx));
    }
  }
  // Memory backward (attention memory halves, ln: vae_model.py:743-744), split by linearity into the decoder's and the
  // prior's contribution: d mem = (d mem_dec + dP_d.W_d) + (d mem_prior + dP_p.W_p), and d ln.weight, d ln.bias, d audio are
  // linear in d mem.  The decoder's half is complete ~150 us before the prior's (whose attention backward is a batched
  // kernel AFTER the chain), so the step's last dependency chain is one accumulating pass over the prior's half instead
  // of the whole memory backward.
  {
    const int R = N * Te;
    {
      cudaStream_t f1 = ax->s[kAuxFan0 + 6];
      TcThroughputScope throughput(fan_min_kblk());
      ACVAE_TRY(stream_dep(sx, f1, ax));
      ACVAE_TRY(linear_bwd_weight(A, E, R, ws.dPd, A, ws.mem, E, gw.d_attn_w + E, 2 * E, f1));
      ACVAE_TRY(colsum(R, A, ws.dPd, A, gw.d_attn_b, f1));
    }
    ACVAE_TRY(linear_bwd_data(R, E, A, ws.dPd, A, w.d_attn_w + E, 2 * E, ws.dmem2, E, sx, 1));
    if (w.ln_w) {
      if (d_audio) ACVAE_TRY(linear_bwd_data(R, d.Eenc, E, ws.dmem2, E, w.ln_w, d.Eenc, d_audio, d.Eenc, sx));
      ACVAE_TRY(linear_bwd_weight(E, d.Eenc, R, ws.dmem2, E, io.audio_embeds, d.Eenc, gw.ln_w, d.Eenc, sx));
      ACVAE_TRY(colsum(R, E, ws.dmem2, E, gw.ln_b, sx));
    } else if (d_audio) {
      ACVAE_CHECK(cudaMemcpyAsync(d_audio, ws.dmem2, sizeof(float) * (size_t)R * E, cudaMemcpyDeviceToDevice, sx));
    }
    ACVAE_CHECK(cudaEventRecord(ev_dec_mem, sx));
    // the prior's half, on its own stream behind its attention accumulation
    {
      cudaStream_t f0 = ax->s[kAuxFan0 + 6];
      TcThroughputScope throughput(fan_min_kblk());
      ACVAE_TRY(stream_dep(sp, f0, ax));
      ACVAE_TRY(linear_bwd_weight(E, E, R, ws.dPp, E, ws.mem, E, gw.p_attn_w + E, 2 * E, f0));
      ACVAE_TRY(colsum(R, E, ws.dPp, E, gw.p_attn_b, f0));
    }
    ACVAE_TRY(linear_bwd_data(R, E, E, ws.dPp, E, w.p_attn_w + E, 2 * E, ws.dmem, E, sp, 1));
    ACVAE_CHECK(cudaStreamWaitEvent(sp, ev_dec_mem, 0));
    if (w.ln_w) {
      if (d_audio) ACVAE_TRY(linear_bwd_data(R, d.Eenc, E, ws.dmem, E, w.ln_w, d.Eenc, d_audio, d.Eenc, sp, 1));
      ACVAE_TRY(linear_bwd_weight(E, d.Eenc, R, ws.dmem, E, io.audio_embeds, d.Eenc, gw.ln_w, d.Eenc, sp, 0, 0, 0, 1));
      ACVAE_TRY(colsum(R, E, ws.dmem, E, gw.ln_b, sp, 1));
    } else if (d_audio) {
      ACVAE_LAUNCH(add_inplace_kernel, grid1d((long long)R * E), 256, 0, sp, (long long)R * E, d_audio, (const float*)ws.dmem);
    }
  }
  // decoder embedding / weight / bias gradients: independent of each other and of the memory backward, so they fan
  // out over four more streams (each GEMM is a ~10 us launch of a few dozen CTAs; in one stream they serialise)
  {
    cudaStream_t* f = &ax->s[kAuxFan0];
    TcThroughputScope throughput(fan_min_kblk());
    for (int i = 0; i < 6; ++i) ACVAE_TRY(stream_dep(st, f[i], ax));
    ACVAE_TRY(linear_bwd_data(NT, E, 3 * E, ws.dgi_d, 3 * E, w.d_wih, 3 * E, ws.dxe_d, E, f[0]));
    ACVAE_CHECK(zero(gw.d_emb, (size_t)V * E, f[0]));
    ACVAE_TRY(scatter_rows(NT, E, ws.dxe_d, E, ws.words, gw.d_emb, f[0]));
    ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, ws.xd, E, gw.d_wih, 3 * E, f[1]));
    ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, ws.ctx_d, E, gw.d_wih + E, 3 * E, f[2]));
    ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_d, 3 * E, zin, E, gw.d_wih + 2 * E, 3 * E, f[3]));
    ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgh_d, 3 * E, io.outputs - E, E, gw.d_whh, E, f[4], T, 0, -1));
    ACVAE_TRY(linear_bwd_weight(A, E, NT, ws.dqp_d, A, io.outputs - E, E, gw.d_attn_w, 2 * E, f[5], T, 0, -1));
    ACVAE_TRY(colsum(NT, 3 * E, ws.dgi_d, 3 * E, gw.d_bih, f[5]));
    ACVAE_TRY(colsum(NT, 3 * E, ws.dgh_d, 3 * E, gw.d_bhh, f[4]));
    if (bucket_event()) {
      // every decoder.* gradient is final once fan streams 0..7 (7: deferred classifier gradients, acvae_defer_classifier_grads)
      // and sx (attention v / bias / memory-side weights) get here
      for (int i = 0; i < 6; ++i) ACVAE_TRY(stream_dep(f[i], f[6], ax));
      ACVAE_TRY(stream_dep(f[7], f[6], ax));
      ACVAE_TRY(stream_dep(sx, f[6], ax));
      ACVAE_TRY(stream_dep(st_user, f[6], ax));       // the classifier's gradients (loss backward, caller's stream)
      ACVAE_CHECK(cudaEventRecord(bucket_event(), f[6]));
    }
    for (int i = 0; i < kAuxFanN; ++i) ACVAE_TRY(stream_dep(f[i], sx, ax));
  }

  // ================= posterior backward (main + one side stream per direction) ==============================
  {
    HeadBwdParams h{};
    h.rows = NT; h.U = E;
    if (gi.d_q_z) { h.dz0 = gi.d_q_z; h.ld_dz0 = E; }
    h.dz1 = ws.dxz_d; h.ld_dz1 = E;
    if (any_dis) { h.flag_mask = dis_mask; h.period = T; h.want = 0; h.use_flags = 1; }   // d z of the decoder: the non-dis steps only
    if (gi.d_q_means) { h.dmean = gi.d_q_means; h.ld_dmean = E; }
    if (gi.d_q_logs) { h.dlog = gi.d_q_logs; h.ld_dlog = E; }
    h.eps = io.eps_q; h.ld_eps = E; h.logv = io.q_logs; h.ld_logv = E;
    h.dml = ws.dml_q; h.ld_dml = 2 * E;
    ACVAE_LAUNCH(head_bwd_kernel, grid1d((long long)NT * E), 256, 0, st, h);
    ACVAE_LAUNCH(pool_bwd_kernel, grid1d((long long)NT * 2 * E), 256, 0, st, N, T, 2 * E, gi.d_q_means_utt, ws.steplens, 0,
                 ws.amax_q, (const float*)nullptr, ws.dho);
    ACVAE_TRY(linear_bwd_data(NT, 2 * E, 2 * E, ws.dml_q, 2 * E, w.q_head_w, 2 * E, ws.dho, 2 * E, st, 1));
  }
  ACVAE_TRY(stream_dep(st, sq0, ax));
  ACVAE_TRY(stream_dep(st, sq1, ax));
  ACVAE_TRY(linear_bwd_weight(2 * E, 2 * E, NT, ws.dml_q, 2 * E, ws.ho, 2 * E, gw.q_head_w, 2 * E, st));
  ACVAE_TRY(colsum(NT, 2 * E, ws.dml_q, 2 * E, gw.q_head_b, st));
  cudaStream_t sq[2] = {sq0, sq1};
  float* carry[2] = {ws.dhq_carry, ws.dzq_carry};   // one carry buffer per direction
  if (post_chain) {
    PostChainBwd pc{};
    pc.N = N; pc.T = T; pc.dho = ws.dho; pc.ho = ws.ho; pc.lens = ws.steplens; pc.bar = ws.bars + 6 * 128;
    for (int dir = 0; dir < 2; ++dir) { pc.whh[dir] = w.q_whh[dir]; pc.gq[dir] = ws.gq[dir]; pc.dgi[dir] = ws.dgi_q[dir]; pc.dgh[dir] = ws.dgh_q[dir]; }
    if (cl) ACVAE_TRY(launch_cluster_chain(post_cl_bwd_kernel, post_cl_clusters(N), 0, sq0, "post_cl_bwd_kernel", pc));
    else ACVAE_TRY(launch_chain(post_chain_bwd_kernel, 0, sq0, "post_chain_bwd_kernel", pc));
    ACVAE_TRY(stream_dep(sq0, sq1, ax));
  }
  for (int dir = 0; dir < 2; ++dir) {
    cudaStream_t s_ = sq[dir];
    if (!post_chain) {
      const int s = T - 1;
      const int t = dir == 0 ? s : T - 1 - s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      GruBwdParams g{};
      g.N = N; g.U = E;
      g.dh_ext = ws.dho + (long long)t * 2 * E + dir * E; g.ld_dh_ext = s1 * 2 * E;
      g.gates = ws.gq[dir] + (long long)t * 4 * E; g.ld_gates = s1 * 4 * E;
      g.hprev = s > 0 ? ws.ho + (long long)tp * 2 * E + dir * E : nullptr; g.ld_hprev = s1 * 2 * E;
      g.lens = ws.steplens; g.len_off = 0; g.t = t;
      g.dgi = ws.dgi_q[dir] + (long long)t * 3 * E; g.ld_dgi = s1 * 3 * E;
      g.dgh = ws.dgh_q[dir] + (long long)t * 3 * E; g.ld_dgh = s1 * 3 * E;
      g.dh_out = carry[dir];
      ACVAE_LAUNCH(gru_bwd_kernel, grid1d((long long)N * E), 256, 0, s_, g);
    }
    for (int s = T - 1; s > 0 && !post_chain; --s) {
      const int t = dir == 0 ? s : T - 1 - s;            // step whose dGh is propagated
      const int tm = dir == 0 ? t - 1 : t + 1;            // the step before it in this direction's forward order
      const int sm = s - 1;                               // its position in forward order
      const int tmp = dir == 0 ? tm - 1 : tm + 1;         // where ITS h_prev lives
      GemmParams p{};
      p.M = N; p.U = E; p.G = 1; p.nseg = 1;
      GemmSeg sg{};
      sg.a = ws.dgh_q[dir] + (long long)t * 3 * E; sg.lda = s1 * 3 * E; sg.w[0] = w.q_whh[dir]; sg.ldw = E; sg.w_trans = 1; sg.K = 3 * E;
      p.seg[0] = sg;
      p.epi.x0 = carry[dir];
      p.epi.x1 = ws.dho + (long long)tm * 2 * E + dir * E; p.epi.ld_x1 = s1 * 2 * E;
      p.epi.gates = ws.gq[dir] + (long long)tm * 4 * E; p.epi.ld_gates = s1 * 4 * E;
      p.epi.prev = sm > 0 ? ws.ho + (long long)tmp * 2 * E + dir * E : nullptr; p.epi.ld_prev = s1 * 2 * E;
      p.epi.lens = ws.steplens; p.epi.t = tm;
      p.epi.y0 = ws.dgi_q[dir] + (long long)tm * 3 * E; p.epi.ld_y0 = s1 * 3 * E;
      p.epi.y1 = ws.dgh_q[dir] + (long long)tm * 3 * E; p.epi.ld_y1 = s1 * 3 * E;
      p.epi.y2 = carry[dir];
      ACVAE_TRY(launch_gemm<EPI_GRU_BWD>(p, s_));
    }
    // weight / bias gradients of this direction: in chain mode each on its own fan stream (nothing but the embedding
    // gradient below is left on the critical stream after the last chain), otherwise behind the direction's BPTT
    cudaStream_t* f = &ax->s[kAuxFan0];
    cudaStream_t s_wih = post_chain ? f[dir * 3] : s_, s_whh = post_chain ? f[dir * 3 + 1] : s_, s_b = post_chain ? f[dir * 3 + 2] : s_;
    if (post_chain)
      for (int i = 0; i < 3; ++i) ACVAE_TRY(stream_dep(sq0, f[dir * 3 + i], ax));
    ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgi_q[dir], 3 * E, ws.xq, E, gw.q_wih[dir], E, s_wih));
    ACVAE_TRY(colsum(NT, 3 * E, ws.dgi_q[dir], 3 * E, gw.q_bih[dir], s_b));
    ACVAE_TRY(colsum(NT, 3 * E, ws.dgh_q[dir], 3 * E, gw.q_bhh[dir], s_b));
    if (dir == 0)
      ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgh_q[0], 3 * E, ws.ho - 2 * E, 2 * E, gw.q_whh[0], E, s_whh, T, 0, -1));
    else
      ACVAE_TRY(linear_bwd_weight(3 * E, E, NT, ws.dgh_q[1], 3 * E, ws.ho + 2 * E + E, 2 * E, gw.q_whh[1], E, s_whh, T, T - 1, 1));
  }
  ACVAE_TRY(stream_dep(sq0, st, ax));
  ACVAE_TRY(stream_dep(sq1, st, ax));
  {
    // d x_q = d gi_fwd . W_ih_fwd + d gi_bwd . W_ih_bwd as ONE two-segment contraction (K = 3E + 3E)
    GemmParams g{};
    g.M = NT; g.U = E; g.G = 1; g.nseg = 2;
    for (int dir = 0; dir < 2; ++dir) {
      GemmSeg sg{};
      sg.a = ws.dgi_q[dir]; sg.lda = 3 * E; sg.w[0] = w.q_wih[dir]; sg.ldw = E; sg.w_trans = 1; sg.K = 3 * E;
      g.seg[dir] = sg;
    }
    g.epi.c[0] = ws.dxq; g.epi.ldc = E; g.epi.scale = 1.0f; g.epi.free_order = 1;
    ACVAE_TRY(launch_gemm<EPI_PLAIN>(g, st));
  }
  ACVAE_CHECK(zero(gw.q_emb, (size_t)V * E, st));
  ACVAE_TRY(scatter_rows(NT, E, ws.dxq, E, ws.qids, gw.q_emb, st));
  if (post_chain)
    for (int i = 0; i < 6; ++i) ACVAE_TRY(stream_dep(ax->s[kAuxFan0 + i], st, ax));
  ACVAE_TRY(stream_dep(sp, st, ax));
  ACVAE_TRY(stream_dep(sx, st, ax));
  ACVAE_TRY(stream_dep(st, st_user, ax));
  return 0;
}

}  // namespace acvae
