// Inference loops with latents drawn from the autoregressive prior:
//   - decode_sample: greedy / multinomial / gumbel stepwise decoding
//     (reference models/vae_model.py:700-720, 880-894; models/word_model.py:173-207)
//   - beam_search:   per-clip beam search with per-beam prior noise
//     (reference models/vae_model.py:896-995)
// State lives in a 2-slot ring [N,2,*]; nothing is kept per step.  `mem_rep`
// consecutive sequences share one clip's memory and projected memory, so K
// captions per clip read the clip's frames from L2 once per step instead of
// K tiled copies (the reference tiles the clip K times through the encoder,
// runners/pytorch_runner_vae.py:101-104).
#pragma once
#include "streams.cuh"
#include "train.cuh"

namespace acvae {

struct SampleWs {
  float *mem, *Pp, *Pd;
  int *words, *unfinished, *active;
  float *qp_p, *w_p, *ctx_p, *gates_p, *c_p, *h_p, *pm, *pl, *pz, *qp_d, *w_d, *ctx_d, *gates_d, *hd;
  float *pmax, *pexp, *psum, *pbest; int* parg;
  float *xe_p, *xe_d, *pre, *pre_h;   // tensor-core step: gathered embeddings, gate pre-activations
  float *pre_d, *q_tab;               // tensor-core step: the decoder's own pre-activations; prior query projection of every word [V,E]
  float* gum[2];                      // tensor-core step: Gumbel variates of the current / next step [N,V]
  size_t bytes;
};

inline SampleWs carve_sample_ws(const acvae_dims& d, void* base) {
  Arena ar(base);
  SampleWs w{};
  const size_t N = d.N, Te = d.Te, E = d.E, A = d.A, clips = d.N / d.mem_rep;
  const size_t nt = (d.V + kVocabTile - 1) / kVocabTile;
  w.mem = ar.take<float>(clips * Te * E); w.Pp = ar.take<float>(clips * Te * E); w.Pd = ar.take<float>(clips * Te * A);
  w.words = ar.take<int>(N * 2); w.unfinished = ar.take<int>(N); w.active = ar.take<int>((size_t)d.T + 1);
  w.qp_p = ar.take<float>(N * 2 * E); w.w_p = ar.take<float>(N * 2 * Te); w.ctx_p = ar.take<float>(N * 2 * E);
  w.gates_p = ar.take<float>(N * 2 * 4 * E); w.c_p = ar.take<float>(N * 2 * E); w.h_p = ar.take<float>(N * 2 * E);
  w.pm = ar.take<float>(N * 2 * E); w.pl = ar.take<float>(N * 2 * E); w.pz = ar.take<float>(N * 2 * E);
  w.qp_d = ar.take<float>(N * 2 * A); w.w_d = ar.take<float>(N * 2 * Te); w.ctx_d = ar.take<float>(N * 2 * E);
  w.gates_d = ar.take<float>(N * 2 * 4 * E); w.hd = ar.take<float>(N * 2 * E);
  w.pmax = ar.take<float>(N * nt); w.pexp = ar.take<float>(N * nt); w.psum = ar.take<float>(N * nt);
  w.pbest = ar.take<float>(N * nt * 2); w.parg = ar.take<int>(N * nt);
  w.xe_p = ar.take<float>(N * E); w.xe_d = ar.take<float>(N * E); w.pre = ar.take<float>(N * 4 * E); w.pre_h = ar.take<float>(N * 3 * E);
  w.pre_d = ar.take<float>(N >= 256 ? N * 3 * E : 0); w.q_tab = ar.take<float>(N >= 256 ? (size_t)d.V * E : 0);
  const size_t gum = N >= 256 ? N * (size_t)d.V : 0;   // only the large-batch tensor-core step uses it
  w.gum[0] = ar.take<float>(gum); w.gum[1] = ar.take<float>(gum);
  w.bytes = ar.off;
  return w;
}

__global__ void sample_init_kernel(int N, int T, int start_idx, int end_idx, int* __restrict__ words,
                                   int* __restrict__ unfinished, int* __restrict__ active, long long* __restrict__ seqs,
                                   float* __restrict__ logprobs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { words[i * 2] = start_idx; words[i * 2 + 1] = start_idx; unfinished[i] = 1; }
  if (i <= T) active[i] = 0;
  if (i < N * T) { seqs[i] = end_idx; logprobs[i] = 0.0f; }   // vae_model.py:764,767
}

__global__ void sample_nsteps_kernel(int T, const int* __restrict__ active, int* __restrict__ n_steps) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int n = T;
    for (int t = 0; t < T; ++t)
      if (active[t] == 0) { n = t + 1; break; }   // vae_model.py:719-720
    *n_steps = n;
  }
}

// ---- one decode step with every contraction on the tensor cores (large sequence counts) ----------------------
// Same arithmetic as prior_step / decoder_step; the gate GEMMs write pre-activations ([N,4E] / [N,3E], L2-resident)
// and small pointwise kernels apply the cells.  C = sum of up to two (A_s, W_s) segments, optionally accumulated.
inline int tc_linear2(int M, int Nn, const float* a0, long long lda0, const float* w0, long long ldw0, int K0,
                      const float* a1, long long lda1, const float* w1, long long ldw1, int K1, const float* bias,
                      float* c, long long ldc, int accumulate, const int* live, cudaStream_t st) {
  GemmParams p{};
  p.M = M; p.U = Nn; p.G = 1; p.nseg = a1 ? 2 : 1; p.live = live;
  p.seg[0] = seg_plain(a0, lda0, w0, ldw0, K0);
  if (a1) p.seg[1] = seg_plain(a1, lda1, w1, ldw1, K1);
  p.epi.c[0] = c; p.epi.ldc = ldc; p.epi.bias[0] = bias; p.epi.scale = 1.0f; p.epi.accumulate = accumulate;
  return launch_gemm<EPI_PLAIN>(p, st);
}
// rows of three tables for the step's input words: prior / decoder word embeddings and the prior's query projection
__global__ void gather3_kernel(int rows, int E, const float* __restrict__ t0, const float* __restrict__ t1,
                               const float* __restrict__ t2, const int* __restrict__ idx, long long idx_stride,
                               float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2, long long ld_o2,
                               const int* __restrict__ live) {
  if (live && *live == 0) return;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= (long long)rows * E) return;
  const int r = (int)(i / E), e = (int)(i % E);
  const long long src = (long long)idx[r * idx_stride] * E + e;
  *reinterpret_cast<float4*>(o0 + i) = *reinterpret_cast<const float4*>(t0 + src);
  *reinterpret_cast<float4*>(o1 + i) = *reinterpret_cast<const float4*>(t1 + src);
  *reinterpret_cast<float4*>(o2 + (long long)r * ld_o2 + e) = *reinterpret_cast<const float4*>(t2 + src);
}

// One decode step (text_encoder.py:247-268, decoder.py:175-203, vae_model.py:808) as three concurrent branches:
//   s2 : the recurrent half of the prior's gates, [z_{t-1} | h_{t-1}] . W^T + b_ih      (ready when the step starts)
//   s1 : everything of the decoder that needs only h^dec_{t-1}: query projection, attention, [xe | ctx] . W_ih^T, h . W_hh^T
//   st : word gather (the prior's query projection is a row of a per-call table: it depends on the word id only) ->
//        prior attention -> gates += [xe | ctx] . W^T -> LSTM cell -> head -> z_t -> decoder gates += z_t . W^T -> GRU cell
// 10 dependent kernels per step instead of 16.
inline int sample_step_tc(const StepCtx& c, const StepBufs& b, SampleWs& ws, Aux* ax, int slot, int prev, const int* words,
                          long long words_stride, const float* eps_t) {
  const int N = c.d.N, E = c.d.E, A = c.d.A, Te = c.d.Te;
  const long long S = b.S;
  const acvae_weights& w = c.w;
  cudaStream_t st = c.st, s1 = ax->s[1], s2 = ax->s[2];
  const int* live = c.live;
  const float* hprev = prev >= 0 ? b.hd + (long long)prev * E : nullptr;
  if (prev >= 0) {
    ACVAE_TRY(stream_dep(st, s2, ax));
    ACVAE_TRY(tc_linear2(N, 4 * E, b.pz + (long long)prev * E, S * E, w.p_wih + 2 * E, 3 * E, E, b.h_p + (long long)prev * E,
                         S * E, w.p_whh, E, E, w.p_bih, ws.pre, 4 * E, 0, live, s2));
  }
  ACVAE_LAUNCH(gather3_kernel, grid1d((long long)N * E / 4), 256, 0, st, N, E, w.p_emb, w.d_emb, (const float*)ws.q_tab, words,
               words_stride, ws.xe_p, ws.xe_d, b.qp_p + (long long)slot * E, S * E, live);
  // ---- decoder branch ----
  ACVAE_TRY(stream_dep(st, s1, ax));
  if (hprev)
    ACVAE_TRY(tc_linear2(N, A, hprev, S * E, w.d_attn_w, 2 * E, E, nullptr, 0, nullptr, 0, 0, nullptr,
                         b.qp_d + (long long)slot * A, S * A, 0, live, s1));
  {
    AttnFwdParams a{};
    a.rows = N; a.Te = Te; a.A = A; a.E = E; a.Dq = E; a.rows_per_clip = c.d.mem_rep; a.live = live;
    a.qp_in = hprev ? b.qp_d + (long long)slot * A : nullptr; a.ld_qp_in = S * A;
    a.P = c.Pd; a.mem = c.mem; a.v = w.d_attn_v; a.mem_lens = c.mem_lens;
    a.ctx = b.ctx_d + (long long)slot * E; a.ld_ctx = S * E;
    a.w_out = nullptr;
    ACVAE_TRY(launch_attn_fwd(a, s1));
  }
  ACVAE_TRY(tc_linear2(N, 3 * E, ws.xe_d, E, w.d_wih, 3 * E, E, b.ctx_d + (long long)slot * E, S * E, w.d_wih + E, 3 * E, E,
                       w.d_bih, ws.pre_d, 3 * E, 0, live, s1));
  if (hprev)
    ACVAE_TRY(tc_linear2(N, 3 * E, hprev, S * E, w.d_whh, E, E, nullptr, 0, nullptr, 0, 0, nullptr, ws.pre_h, 3 * E, 0, live, s1));
  // ---- prior (critical path) ----
  {
    AttnFwdParams a{};
    a.rows = N; a.Te = Te; a.A = E; a.E = E; a.Dq = E; a.rows_per_clip = c.d.mem_rep; a.live = live;
    a.qp_in = b.qp_p + (long long)slot * E; a.ld_qp_in = S * E;
    a.P = c.Pp; a.mem = c.mem; a.v = w.p_attn_v; a.mem_lens = c.mem_lens;
    a.ctx = b.ctx_p + (long long)slot * E; a.ld_ctx = S * E;
    a.w_out = nullptr;
    ACVAE_TRY(launch_attn_fwd(a, st));
  }
  if (prev >= 0) ACVAE_TRY(stream_dep(s2, st, ax));
  ACVAE_TRY(tc_linear2(N, 4 * E, ws.xe_p, E, w.p_wih, 3 * E, E, b.ctx_p + (long long)slot * E, S * E, w.p_wih + E, 3 * E, E,
                       prev >= 0 ? nullptr : w.p_bih, ws.pre, 4 * E, prev >= 0 ? 1 : 0, live, st));
  ACVAE_LAUNCH(lstm_cell_kernel, grid1d((long long)N * E), 256, 0, st, N, E, (const float*)ws.pre, w.p_bhh,
               prev >= 0 ? (const float*)(b.c_p + (long long)prev * E) : (const float*)nullptr, S * E,
               b.c_p + (long long)slot * E, b.h_p + (long long)slot * E, S * E, live);
  ACVAE_TRY(tc_linear2(N, 2 * E, b.h_p + (long long)slot * E, S * E, w.p_head_w, E, E, nullptr, 0, nullptr, 0, 0, w.p_head_b,
                       ws.pre, 2 * E, 0, live, st));
  ACVAE_LAUNCH(head_cell_kernel, grid1d((long long)N * E), 256, 0, st, N, E, (const float*)ws.pre, eps_t,
               b.pm + (long long)slot * E, b.pl + (long long)slot * E, b.pz + (long long)slot * E, S * E, live);
  // ---- decoder, fed the prior's sample (vae_model.py:808) ----
  ACVAE_TRY(stream_dep(s1, st, ax));
  ACVAE_TRY(tc_linear2(N, 3 * E, b.pz + (long long)slot * E, S * E, w.d_wih + 2 * E, 3 * E, E, nullptr, 0, nullptr, 0, 0,
                       nullptr, ws.pre_d, 3 * E, 1, live, st));
  ACVAE_LAUNCH(gru_cell_kernel, grid1d((long long)N * E), 256, 0, st, N, E, (const float*)ws.pre_d,
               hprev ? (const float*)ws.pre_h : (const float*)nullptr, w.d_bhh, hprev, S * E, b.hd + (long long)slot * E, S * E,
               live);
  return 0;
}

// the next call draws from a fresh counter range (also when the call is a replayed CUDA graph)
__global__ void rng_advance_kernel(unsigned long long* rng) { rng[1] += 1ull; }

// g = -log(-log(u + 1e-20) + 1e-20) (word_model.py:188-190), elementwise at full occupancy on a side stream:
// inside the vocabulary GEMM's epilogue the two logs per logit cost more than the GEMM itself.
__global__ void gumbel_kernel(long long n, const float* __restrict__ u, float* __restrict__ g) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const bool vec = ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  if (vec && i + 3 < n) {
    const float4 x = *reinterpret_cast<const float4*>(u + i);
    *reinterpret_cast<float4*>(g + i) = make_float4(gumbel_from_u(x.x), gumbel_from_u(x.y), gumbel_from_u(x.z), gumbel_from_u(x.w));
  } else {
    for (long long k = i; k < n && k < i + 4; ++k) g[k] = gumbel_from_u(u[k]);
  }
}

inline bool sample_tc_ok(const acvae_dims& d) {
  static int dis = -1;
  if (dis < 0) { const char* e = getenv("ACVAE_DISABLE_SAMPLE_TC"); dis = (e && e[0] == '1') ? 1 : 0; }
  return !dis && tc_enabled() && d.N >= 256 && d.E % 32 == 0 && d.A % 32 == 0;
}

inline int decode_sample(const acvae_dims& d, const acvae_weights& w, const acvae_sample_io& io, void* workspace,
                         cudaStream_t st) {
  SampleWs ws = carve_sample_ws(d, workspace);
  const int N = d.N, T = d.T, E = d.E;
  ACVAE_TRY(wait_input_event(st));
  ACVAE_TRY(memory_prepare(d, w, io.audio_embeds, ws.mem, ws.Pp, ws.Pd, st));
  ACVAE_LAUNCH(sample_init_kernel, grid1d((long long)N * T + T + 1), 256, 0, st, N, T, io.start_idx, io.end_idx, ws.words,
               ws.unfinished, ws.active, (long long*)io.seqs, io.sampled_logprobs);
  StepBufs b{2, ws.qp_p, ws.w_p, ws.ctx_p, ws.gates_p, ws.c_p, ws.h_p, ws.pm, ws.pl, ws.pz,
             ws.qp_d, ws.w_d, ws.ctx_d, ws.gates_d, ws.hd};
  const bool use_tc = sample_tc_ok(d);
  const bool draw = io.method != 0 && !io.u;                 // noise drawn in the vocabulary epilogue (Philox), no u tensor
  Aux* axs = use_tc ? aux() : nullptr;                        // side streams of the tensor-core step's branches
  if (use_tc && !axs) return set_error("decode_sample", "could not create the side streams");
  Aux* ax = (use_tc && io.method != 0 && io.u) ? axs : nullptr;      // injected u: side stream, Gumbel variates one step ahead
  // the prior's query projection depends on the input word only: one row per vocabulary entry, once per call
  if (use_tc)
    ACVAE_TRY(tc_linear2(d.V, E, w.p_emb, E, w.p_attn_w, 2 * E, E, nullptr, 0, nullptr, 0, 0, nullptr, ws.q_tab, E, 0, nullptr, st));
  const long long nv = (long long)N * d.V;
  auto gumbel_ahead = [&](int t) -> int {
    cudaStream_t sg = ax->s[0];
    ACVAE_TRY(stream_dep(st, sg, ax));     // gum[t & 1] was last read by step t-2 (already enqueued on st)
    ACVAE_LAUNCH(gumbel_kernel, grid1d((nv + 3) / 4), 256, 0, sg, nv, io.u + (long long)t * nv, ws.gum[t & 1]);
    return 0;
  };
  if (ax) ACVAE_TRY(gumbel_ahead(0));
  for (int t = 0; t < T; ++t) {
    const int slot = t & 1, prev = t > 0 ? (t - 1) & 1 : -1;
    const int* live = t > 0 ? ws.active + (t - 1) : nullptr;
    StepCtx c{d, w, st, io.mem_lens, ws.mem, ws.Pp, ws.Pd, live};
    if (use_tc) {
      ACVAE_TRY(sample_step_tc(c, b, ws, axs, slot, prev, ws.words + slot, 2, io.eps_p + (long long)t * N * E));
    } else {
      ACVAE_TRY(prior_step(c, b, slot, prev, ws.words + slot, 2, io.eps_p + (long long)t * N * E));
      // inference: the decoder consumes the prior's sample (vae_model.py:808)
      ACVAE_TRY(decoder_step(c, b, slot, prev, ws.words + slot, 2, ws.pz + (long long)slot * E, 2LL * E, nullptr, 0, 0));
    }
    VocabStatsArgs v{};
    v.M = N; v.V = d.V; v.E = E; v.hidden = ws.hd + (long long)slot * E; v.ld_h = 2LL * E;
    v.cls_w = w.cls_w; v.cls_b = w.cls_b; v.live = live;
    v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
    if (io.method != 0 && ax) {
      ACVAE_TRY(stream_dep(ax->s[0], st, ax));             // this step's variates are ready
      if (t + 1 < T) ACVAE_TRY(gumbel_ahead(t + 1));
      v.noise = ws.gum[t & 1]; v.ld_noise = d.V; v.noise_is_gumbel = 1;
      v.inv_temp = io.method == 1 ? 1.0f / io.temp : 1.0f;
    } else if (io.method != 0 && io.u) {
      v.noise = io.u + (long long)t * N * d.V; v.ld_noise = d.V;
      v.inv_temp = io.method == 1 ? 1.0f / io.temp : 1.0f;   // word_model.py:187-198
    } else if (draw) {
      v.rng = reinterpret_cast<const unsigned long long*>(io.rng_state); v.rng_step = t;
      v.inv_temp = io.method == 1 ? 1.0f / io.temp : 1.0f;
    }
    v.red.logprob = io.sampled_logprobs + t; v.red.ld_row = T;
    v.red.seqs = (long long*)io.seqs + t; v.red.ld_seqs = T;
    v.red.next_word = ws.words + (slot ^ 1); v.red.ld_next = 2;
    v.red.unfinished = ws.unfinished; v.red.end_idx = io.end_idx; v.red.active_count = ws.active + t;
    ACVAE_TRY(vocab_stats(v, st));
    auto keep = [&](float* dst, const float* ring) -> int {
      if (!dst) return 0;
      ACVAE_LAUNCH(copy2d_kernel, grid1d((long long)N * E), 256, 0, st, (long long)N, E, ring + (long long)slot * E,
                   2LL * E, dst + (long long)t * E, (long long)T * E);
      return 0;
    };
    ACVAE_TRY(keep(io.p_means, ws.pm)); ACVAE_TRY(keep(io.p_logs, ws.pl));
    ACVAE_TRY(keep(io.p_z, ws.pz)); ACVAE_TRY(keep(io.outputs, ws.hd));
  }
  if (io.n_steps) ACVAE_LAUNCH(sample_nsteps_kernel, 1, 32, 0, st, T, (const int*)ws.active, io.n_steps);
  if (draw) ACVAE_LAUNCH(rng_advance_kernel, 1, 1, 0, st, reinterpret_cast<unsigned long long*>(io.rng_state));
  return 0;
}

// ================================ beam search ====================================================
struct BeamWs {
  float *mem, *Pp, *Pd;
  int *words, *prev, *hist[2];
  float *top_lp, *logits;
  float *qp_p, *w_p, *ctx_p, *gates_p, *c_p, *h_p, *pm, *pl, *pz, *qp_d, *w_d, *ctx_d, *gates_d, *hd;
  size_t bytes;
};

inline BeamWs carve_beam_ws(const acvae_dims& d, int beam, void* base) {
  Arena ar(base);
  BeamWs w{};
  const size_t clips = d.N, R = (size_t)d.N * beam, Te = d.Te, E = d.E, A = d.A, T = d.T;
  w.mem = ar.take<float>(clips * Te * E); w.Pp = ar.take<float>(clips * Te * E); w.Pd = ar.take<float>(clips * Te * A);
  w.words = ar.take<int>(R * 2); w.prev = ar.take<int>(R); w.hist[0] = ar.take<int>(R * T); w.hist[1] = ar.take<int>(R * T);
  w.top_lp = ar.take<float>(R); w.logits = ar.take<float>(R * d.V);
  w.qp_p = ar.take<float>(R * 2 * E); w.w_p = ar.take<float>(R * 2 * Te); w.ctx_p = ar.take<float>(R * 2 * E);
  w.gates_p = ar.take<float>(R * 2 * 4 * E); w.c_p = ar.take<float>(R * 2 * E); w.h_p = ar.take<float>(R * 2 * E);
  w.pm = ar.take<float>(R * 2 * E); w.pl = ar.take<float>(R * 2 * E); w.pz = ar.take<float>(R * 2 * E);
  w.qp_d = ar.take<float>(R * 2 * A); w.w_d = ar.take<float>(R * 2 * Te); w.ctx_d = ar.take<float>(R * 2 * E);
  w.gates_d = ar.take<float>(R * 2 * 4 * E); w.hd = ar.take<float>(R * 2 * E);
  w.bytes = ar.off;
  return w;
}

__global__ void beam_init_kernel(int R, int start_idx, int* __restrict__ words, float* __restrict__ top_lp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) { words[i * 2] = start_idx; words[i * 2 + 1] = start_idx; top_lp[i] = 0.0f; }   // vae_model.py:929,962
}

// One CTA per clip: log-softmax each beam row, add the running beam score, take the top `beam`
// of the beam*V candidates (vae_model.py:909-916), extend the hypotheses (:917-921).
__global__ void __launch_bounds__(256) beam_topk_kernel(int beam, int V, int T, int t, float* __restrict__ logits,
                                                        float* __restrict__ top_lp, int* __restrict__ prev,
                                                        int* __restrict__ words /*[R,2] slot 0*/,
                                                        const int* __restrict__ hist_in, int* __restrict__ hist_out) {
  __shared__ float red[33];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ float s_lse[32];
  __shared__ float s_newlp[32];
  __shared__ int s_newidx[32];
  const int clip = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* lg = logits + (long long)clip * beam * V;
  for (int b = 0; b < beam; ++b) {
    float mx = -INFINITY;
    for (int v = tid; v < V; v += blockDim.x) mx = fmaxf(mx, lg[(long long)b * V + v]);
    mx = block_max(mx, red);
    float se = 0.0f;
    for (int v = tid; v < V; v += blockDim.x) se += expf(lg[(long long)b * V + v] - mx);
    se = block_sum(se, red);
    if (tid == 0) s_lse[b] = mx + logf(se);
  }
  __syncthreads();
  // candidates in place: score = top_lp[b] + logit - lse[b]
  for (int b = 0; b < beam; ++b) {
    const float off = top_lp[clip * beam + b] - s_lse[b];
    for (int v = tid; v < V; v += blockDim.x) lg[(long long)b * V + v] += off;
  }
  __syncthreads();
  const int total = beam * V;
  for (int k = 0; k < beam; ++k) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < total; i += blockDim.x) {
      const float x = lg[i];
      if (x > best) { best = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      float bb = s_val[0]; int ii = s_idx[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (s_val[w] > bb || (s_val[w] == bb && s_idx[w] < ii)) { bb = s_val[w]; ii = s_idx[w]; }
      s_newlp[k] = bb; s_newidx[k] = ii;
      lg[ii] = -INFINITY;   // exclude from the next pass
    }
    __syncthreads();
  }
  for (int k = tid; k < beam; k += blockDim.x) {
    const int idx = s_newidx[k];
    const int r = clip * beam + k;
    top_lp[r] = s_newlp[k];
    prev[r] = idx / V;
    words[r * 2] = idx % V;
  }
  for (int i = tid; i < beam * (t + 1); i += blockDim.x) {
    const int k = i / (t + 1), tt = i % (t + 1);
    const int idx = s_newidx[k];
    const int src = clip * beam + idx / V;
    hist_out[(long long)(clip * beam + k) * T + tt] = tt < t ? hist_in[(long long)src * T + tt] : idx % V;
  }
}

// state[(clip,k), slot 0] = state[(clip, prev[k]), slot 1]   (vae_model.py:963-969)
__global__ void beam_reindex_kernel(int R, int beam, int E, const int* __restrict__ prev, float* __restrict__ hd,
                                    float* __restrict__ h_p, float* __restrict__ c_p, float* __restrict__ pz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * E) return;
  const int r = (int)(i / E), e = (int)(i % E);
  const int src = (r / beam) * beam + prev[r];
  const long long so = ((long long)src * 2 + 1) * E + e, dst_o = ((long long)r * 2) * E + e;
  hd[dst_o] = hd[so]; h_p[dst_o] = h_p[so]; c_p[dst_o] = c_p[so]; pz[dst_o] = pz[so];
}

__global__ void beam_final_kernel(int N, int beam, int T, const int* __restrict__ hist, long long* __restrict__ seqs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * T) return;
  const int n = i / T, t = i % T;
  seqs[i] = hist[(long long)(n * beam) * T + t];   // top beam (vae_model.py:986)
}

inline int beam_search(const acvae_dims& d0, const acvae_weights& w, const float* audio, const int* mem_lens,
                       const float* eps_b, int beam, int start_idx, int64_t* seqs, void* workspace, cudaStream_t st) {
  BeamWs ws = carve_beam_ws(d0, beam, workspace);
  const int clips = d0.N, T = d0.T, E = d0.E, V = d0.V, R = clips * beam;
  acvae_dims dm = d0;                       // memory: one row per clip
  ACVAE_TRY(wait_input_event(st));
  ACVAE_TRY(memory_prepare(dm, w, audio, ws.mem, ws.Pp, ws.Pd, st));
  acvae_dims d = d0;
  d.N = R; d.mem_rep = beam;                // decode rows: beam hypotheses per clip share its memory
  ACVAE_LAUNCH(beam_init_kernel, grid1d(R), 256, 0, st, R, start_idx, ws.words, ws.top_lp);
  StepBufs b{2, ws.qp_p, ws.w_p, ws.ctx_p, ws.gates_p, ws.c_p, ws.h_p, ws.pm, ws.pl, ws.pz,
             ws.qp_d, ws.w_d, ws.ctx_d, ws.gates_d, ws.hd};
  StepCtx c{d, w, st, mem_lens, ws.mem, ws.Pp, ws.Pd, nullptr};
  for (int t = 0; t < T; ++t) {
    const int prev = t > 0 ? 0 : -1;        // slot 0 holds the re-indexed state, slot 1 the fresh step
    ACVAE_TRY(prior_step(c, b, 1, prev, ws.words, 2, eps_b + (long long)t * R * E));
    ACVAE_TRY(decoder_step(c, b, 1, prev, ws.words, 2, ws.pz + E, 2LL * E, nullptr, 0, 0));
    ACVAE_TRY(linear_fwd(R, V, E, ws.hd + E, 2LL * E, w.cls_w, E, w.cls_b, ws.logits, V, st));
    ACVAE_LAUNCH(beam_topk_kernel, clips, 256, 0, st, beam, V, T, t, ws.logits, ws.top_lp, ws.prev, ws.words,
                 (const int*)ws.hist[t & 1], ws.hist[(t + 1) & 1]);
    ACVAE_LAUNCH(beam_reindex_kernel, grid1d((long long)R * E), 256, 0, st, R, beam, E, (const int*)ws.prev, ws.hd,
                 ws.h_p, ws.c_p, ws.pz);
  }
  ACVAE_LAUNCH(beam_final_kernel, grid1d((long long)clips * T), 256, 0, st, clips, beam, T, (const int*)ws.hist[T & 1],
               (long long*)seqs);
  return 0;
}

// ================================ diverse beam search ============================================
// CaptionModel.diverse_beam_search (word_model.py:297-394) with Hybrid_VAEModel.dbs_step (vae_model.py:997-1048).
// Rows r = (clip*G + g)*bdash + k.  Group g runs g steps behind group 0; at global step t every active group
// advances one local step.  The model step (prior, decoder, vocabulary projection) of ALL clips and groups is one
// batched launch sequence -- it depends only on the selections of global step t-1 -- while the selection itself is
// sequential over the groups of a clip (group g is penalised with the words groups < g hold at the same position
// AFTER their own selection at this global step), so one CTA per clip walks its groups in order.
struct DbsWs {
  float *mem, *Pp, *Pd;
  int *words, *prev, *hist;
  float *top_lp, *logits;
  float *qp_p, *w_p, *ctx_p, *gates_p, *c_p, *h_p, *pm, *pl, *pz, *qp_d, *w_d, *ctx_d, *gates_d, *hd;
  double* done_score;   // [clips*G, bdash] best finished hypotheses per group, sorted by score (stable)
  int *done_len, *done_seq, *done_cnt;
  size_t state_off, state_bytes;   // the recurrent-state block zeroed before the first step
  size_t bytes;
};

inline DbsWs carve_dbs_ws(const acvae_dims& d, int G, int bdash, void* base) {
  Arena ar(base);
  DbsWs w{};
  const size_t clips = d.N, R = (size_t)d.N * G * bdash, Te = d.Te, E = d.E, A = d.A, T = d.T;
  w.mem = ar.take<float>(clips * Te * E); w.Pp = ar.take<float>(clips * Te * E); w.Pd = ar.take<float>(clips * Te * A);
  w.words = ar.take<int>(R * 2); w.prev = ar.take<int>(R); w.hist = ar.take<int>(R * T);
  w.top_lp = ar.take<float>(R); w.logits = ar.take<float>(R * d.V);
  w.qp_p = ar.take<float>(R * 2 * E); w.w_p = ar.take<float>(R * 2 * Te); w.ctx_p = ar.take<float>(R * 2 * E);
  w.gates_p = ar.take<float>(R * 2 * 4 * E);
  w.state_off = ar.off;
  w.c_p = ar.take<float>(R * 2 * E); w.h_p = ar.take<float>(R * 2 * E); w.pz = ar.take<float>(R * 2 * E); w.hd = ar.take<float>(R * 2 * E);
  w.state_bytes = ar.off - w.state_off;
  w.pm = ar.take<float>(R * 2 * E); w.pl = ar.take<float>(R * 2 * E);
  w.qp_d = ar.take<float>(R * 2 * A); w.w_d = ar.take<float>(R * 2 * Te); w.ctx_d = ar.take<float>(R * 2 * E);
  w.gates_d = ar.take<float>(R * 2 * 4 * E);
  w.done_score = ar.take<double>(R); w.done_len = ar.take<int>(R); w.done_seq = ar.take<int>(R * T); w.done_cnt = ar.take<int>(clips * G);
  w.bytes = ar.off;
  return w;
}

__global__ void dbs_init_kernel(int R, int groups, int start_idx, int* __restrict__ words, float* __restrict__ top_lp,
                                int* __restrict__ done_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) { words[i * 2] = start_idx; words[i * 2 + 1] = start_idx; top_lp[i] = 0.0f; }   // vae_model.py:1008, word_model.py:331
  if (i < groups) done_cnt[i] = 0;
}

struct DbsParams {
  int G, bdash, V, T, t, end_idx;
  float temperature, lambda;
  float* logits; float* top_lp; int* prev; int* words; int* hist;
  double* done_score; int* done_len; int* done_seq; int* done_cnt;
};

constexpr int kDbsMaxBeam = 32;

// One CTA per clip; groups in order (word_model.py:337-386).
__global__ void __launch_bounds__(256) dbs_select_kernel(const __grid_constant__ DbsParams p) {
  extern __shared__ int s_hist[];            // [bdash][T] staging of the group's hypotheses while they are re-ordered
  __shared__ float red[33];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ float s_newlp[kDbsMaxBeam];
  __shared__ int s_newidx[kDbsMaxBeam];
  __shared__ int s_tok[kDbsMaxBeam];
  __shared__ float s_cnt[kDbsMaxBeam];
  __shared__ int s_ntok;
  const int clip = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = p.G, bdash = p.bdash, V = p.V, T = p.T, t = p.t;
  for (int g = 0; g < G; ++g) {
    if (!(g <= t && t <= T + g - 1)) continue;                 // word_model.py:338 (uniform over the CTA)
    const int lt = t - g;
    const int r0 = (clip * G + g) * bdash;
    float* lg = p.logits + (long long)r0 * V;
    const int nb = lt == 0 ? 1 : bdash;                        // first local step: all rows are equal, row 0 is used (:357-359)
    // ---- words the earlier groups hold at this position, with multiplicities (add_diversity, :298-313) ----
    if (tid == 0) {
      int n = 0;
      for (int pc = 0; pc < g; ++pc)
        for (int k = 0; k < bdash; ++k) {
          const int tok = p.hist[(long long)((clip * G + pc) * bdash + k) * T + lt];
          int j = 0;
          while (j < n && s_tok[j] != tok) ++j;
          if (j == n) { s_tok[n] = tok; s_cnt[n] = 0.0f; ++n; }
          s_cnt[j] += 1.0f;
        }
      s_ntok = n;
    }
    // ---- log_softmax(log_softmax(logits) / temperature) per row (:353-354), in place ----
    for (int b = 0; b < nb; ++b) {
      float* row = lg + (long long)b * V;
      float mx = -INFINITY;
      for (int v = tid; v < V; v += blockDim.x) mx = fmaxf(mx, row[v]);
      mx = block_max(mx, red);
      float se = 0.0f;
      for (int v = tid; v < V; v += blockDim.x) se += expf(row[v] - mx);
      se = block_sum(se, red);
      const float lse = logf(se);
      float mx2 = -INFINITY;
      for (int v = tid; v < V; v += blockDim.x) {
        const float x = __fdiv_rn(__fsub_rn(__fsub_rn(row[v], mx), lse), p.temperature);
        row[v] = x;
        mx2 = fmaxf(mx2, x);
      }
      mx2 = block_max(mx2, red);
      float se2 = 0.0f;
      for (int v = tid; v < V; v += blockDim.x) se2 += expf(row[v] - mx2);
      se2 = block_sum(se2, red);
      const float lse2 = logf(se2);
      for (int v = tid; v < V; v += blockDim.x) row[v] = __fsub_rn(__fsub_rn(row[v], mx2), lse2);
    }
    __syncthreads();
    // ---- diversity penalty, then the running score of the hypothesis (:355-356) ----
    for (int i = tid; i < s_ntok * nb; i += blockDim.x) {
      const int b = i / s_ntok, j = i % s_ntok;
      float* x = lg + (long long)b * V + s_tok[j];
      *x = __fsub_rn(*x, __fmul_rn(s_cnt[j], p.lambda));
    }
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
      const float base = p.top_lp[r0 + b];
      float* row = lg + (long long)b * V;
      for (int v = tid; v < V; v += blockDim.x) row[v] = __fadd_rn(base, row[v]);
    }
    __syncthreads();
    // ---- top bdash of the nb*V candidates, sorted (:357-362) ----
    const int total = nb * V;
    for (int k = 0; k < bdash; ++k) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = tid; i < total; i += blockDim.x) {
        const float x = lg[i];
        if (x > best) { best = x; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
      __syncthreads();
      if (tid == 0) {
        float bb = s_val[0]; int ii = s_idx[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
          if (s_val[w] > bb || (s_val[w] == bb && s_idx[w] < ii)) { bb = s_val[w]; ii = s_idx[w]; }
        s_newlp[k] = bb; s_newidx[k] = ii;
        lg[ii] = -INFINITY;   // exclude from the next pass
      }
      __syncthreads();
    }
    // ---- re-order and extend the group's hypotheses (:363-371) ----
    for (int i = tid; i < bdash * lt; i += blockDim.x) s_hist[i] = p.hist[(long long)(r0 + i / lt) * T + i % lt];
    __syncthreads();
    for (int i = tid; i < bdash * (lt + 1); i += blockDim.x) {
      const int k = i / (lt + 1), tt = i % (lt + 1);
      const int idx = s_newidx[k];
      p.hist[(long long)(r0 + k) * T + tt] = tt < lt ? s_hist[(idx / V) * lt + tt] : idx % V;
    }
    __syncthreads();
    // ---- finished hypotheses (:373-385): kept sorted by score, ties in insertion order (== sorted(...)[:bdash], :387) ----
    if (tid == 0) {
      const int grp = clip * G + g;
      int cnt = p.done_cnt[grp];
      for (int k = 0; k < bdash; ++k) {
        const int idx = s_newidx[k];
        const int word = idx % V;
        float lp = s_newlp[k];
        const bool is_end = word == p.end_idx || t == T + g - 1;
        if (is_end) {
          const double score = (double)lp / (double)(lt + 1);
          int pos = cnt < bdash ? cnt : bdash;
          while (pos > 0 && score > p.done_score[(long long)grp * bdash + pos - 1]) --pos;
          if (pos < bdash) {
            const int last = cnt < bdash ? cnt : bdash - 1;
            for (int j = last; j > pos; --j) {
              p.done_score[(long long)grp * bdash + j] = p.done_score[(long long)grp * bdash + j - 1];
              p.done_len[grp * bdash + j] = p.done_len[grp * bdash + j - 1];
              for (int tt = 0; tt < T; ++tt)
                p.done_seq[(long long)(grp * bdash + j) * T + tt] = p.done_seq[(long long)(grp * bdash + j - 1) * T + tt];
            }
            p.done_score[(long long)grp * bdash + pos] = score;
            p.done_len[grp * bdash + pos] = lt + 1;
            for (int tt = 0; tt <= lt; ++tt) p.done_seq[(long long)(grp * bdash + pos) * T + tt] = p.hist[(long long)(r0 + k) * T + tt];
            if (cnt < bdash) ++cnt;
          }
          lp -= 1000.0f;                                        // :385
        }
        p.top_lp[r0 + k] = lp;
        p.prev[r0 + k] = idx / V;
        p.words[(r0 + k) * 2] = word;
      }
      p.done_cnt[grp] = cnt;
    }
    __syncthreads();
  }
}

// state[(clip,g,k), slot 0] = state[(clip,g,prev[k]), slot 1] for the groups that stepped at global step t
// (vae_model.py:1015-1024); groups that have not started keep their zero state and <start> word
__global__ void dbs_reindex_kernel(int R, int G, int bdash, int T, int t, int E, const int* __restrict__ prev,
                                   float* __restrict__ hd, float* __restrict__ h_p, float* __restrict__ c_p,
                                   float* __restrict__ pz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * E) return;
  const int r = (int)(i / E), e = (int)(i % E);
  const int g = (r / bdash) % G;
  if (!(g <= t && t <= T + g - 1)) return;
  const int src = (r / bdash) * bdash + prev[r];
  const long long so = ((long long)src * 2 + 1) * E + e, dst_o = ((long long)r * 2) * E + e;
  hd[dst_o] = hd[so]; h_p[dst_o] = h_p[so]; c_p[dst_o] = c_p[so]; pz[dst_o] = pz[so];
}

// seqs [clips, n_out, T] (END-filled): group_nbest -> every group's bdash best in group order (:388-389),
// else the best of each group (:390-391)
__global__ void dbs_final_kernel(int clips, int G, int bdash, int T, int n_out, int nbest, int end_idx,
                                 const int* __restrict__ done_len, const int* __restrict__ done_seq,
                                 const int* __restrict__ done_cnt, long long* __restrict__ seqs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= clips * n_out * T) return;
  const int tt = i % T, o = (i / T) % n_out, clip = i / (T * n_out);
  int g, j;
  if (nbest) { g = o / bdash; j = o % bdash; } else { g = o; j = 0; }
  long long v = end_idx;
  if (g < G) {
    const int grp = clip * G + g;
    if (j < done_cnt[grp] && tt < done_len[grp * bdash + j]) v = done_seq[(long long)(grp * bdash + j) * T + tt];
  }
  seqs[i] = v;
}

inline int diverse_beam_search(const acvae_dims& d0, const acvae_weights& w, const float* audio, const int* mem_lens,
                               const float* eps_g, int beam, int G, float lambda, float temperature, int nbest, int start_idx,
                               int end_idx, int64_t* seqs, void* workspace, cudaStream_t st) {
  const int bdash = beam / G;
  DbsWs ws = carve_dbs_ws(d0, G, bdash, workspace);
  const int clips = d0.N, T = d0.T, E = d0.E, V = d0.V, R = clips * G * bdash;
  acvae_dims dm = d0;                       // memory: one row per clip
  ACVAE_TRY(wait_input_event(st));
  ACVAE_TRY(memory_prepare(dm, w, audio, ws.mem, ws.Pp, ws.Pd, st));
  acvae_dims d = d0;
  d.N = R; d.mem_rep = G * bdash;           // decode rows: all hypotheses of a clip share its memory
  ACVAE_CHECK(cudaMemsetAsync(static_cast<char*>(workspace) + ws.state_off, 0, ws.state_bytes, st));
  const int ninit = R > clips * G ? R : clips * G;
  ACVAE_LAUNCH(dbs_init_kernel, grid1d(ninit), 256, 0, st, R, clips * G, start_idx, ws.words, ws.top_lp, ws.done_cnt);
  StepBufs b{2, ws.qp_p, ws.w_p, ws.ctx_p, ws.gates_p, ws.c_p, ws.h_p, ws.pm, ws.pl, ws.pz,
             ws.qp_d, ws.w_d, ws.ctx_d, ws.gates_d, ws.hd};
  StepCtx c{d, w, st, mem_lens, ws.mem, ws.Pp, ws.Pd, nullptr};
  DbsParams p{};
  p.G = G; p.bdash = bdash; p.V = V; p.T = T; p.end_idx = end_idx; p.temperature = temperature; p.lambda = lambda;
  p.logits = ws.logits; p.top_lp = ws.top_lp; p.prev = ws.prev; p.words = ws.words; p.hist = ws.hist;
  p.done_score = ws.done_score; p.done_len = ws.done_len; p.done_seq = ws.done_seq; p.done_cnt = ws.done_cnt;
  const size_t smem = sizeof(int) * (size_t)bdash * T;
  for (int t = 0; t < T + G - 1; ++t) {                                   // word_model.py:336
    // slot 0 holds the (re-indexed or initial zero) state, slot 1 receives the fresh step
    ACVAE_TRY(prior_step(c, b, 1, 0, ws.words, 2, eps_g + (long long)t * R * E));
    ACVAE_TRY(decoder_step(c, b, 1, 0, ws.words, 2, ws.pz + E, 2LL * E, nullptr, 0, 0));
    ACVAE_TRY(linear_fwd(R, V, E, ws.hd + E, 2LL * E, w.cls_w, E, w.cls_b, ws.logits, V, st));
    p.t = t;
    ACVAE_LAUNCH(dbs_select_kernel, clips, 256, smem, st, p);
    ACVAE_LAUNCH(dbs_reindex_kernel, grid1d((long long)R * E), 256, 0, st, R, G, bdash, T, t, E, (const int*)ws.prev, ws.hd,
                 ws.h_p, ws.c_p, ws.pz);
  }
  const int n_out = nbest ? beam : G;
  ACVAE_LAUNCH(dbs_final_kernel, grid1d((long long)clips * n_out * T), 256, 0, st, clips, G, bdash, T, n_out, nbest, end_idx,
               (const int*)ws.done_len, (const int*)ws.done_seq, (const int*)ws.done_cnt, (long long*)seqs);
  return 0;
}

// ================================ diversity statistics ===========================================
// Div-1 / Div-2 of the K captions decoded per clip (utils/div_utils.py:11-29: distinct n-grams over the clip's captions
// divided by its token count) and the presence flags behind the global distinct-word count gDiv-1 (:31-44, n = 1),
// straight from the id tensor the sampling loop leaves on the device.  A caption is the ids up to the first <end>,
// <start> skipped (runners/base_runner.py:146-157).  One CTA per clip; the K*L tokens live in shared memory and an
// n-gram counts when no earlier position of the clip holds the same one (quadratic in K*L <= a few hundred).
__global__ void __launch_bounds__(256) diversity_stats_kernel(int K, int L, int V, int start_idx, int end_idx,
                                                             const long long* __restrict__ seqs, double* __restrict__ div1,
                                                             double* __restrict__ div2, int* __restrict__ vocab_flags) {
  extern __shared__ int dsm_i[];
  int* tok = dsm_i;                 // [K*L] compacted tokens of the clip, caption after caption
  int* cap = tok + K * L;           // [K*L] caption index of each compacted token
  __shared__ int s_n, s_uni, s_bi;
  const int clip = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int n = 0;
    for (int k = 0; k < K; ++k)
      for (int l = 0; l < L; ++l) {
        const int w = (int)seqs[((long long)clip * K + k) * L + l];
        if (w == end_idx) break;
        if (w == start_idx) continue;
        tok[n] = w; cap[n] = k; ++n;
      }
    s_n = n; s_uni = 0; s_bi = 0;
  }
  __syncthreads();
  const int n = s_n;
  int uni = 0, bi = 0;
  for (int p = tid; p < n; p += blockDim.x) {
    const int w = tok[p];
    if (vocab_flags && w >= 0 && w < V) vocab_flags[w] = 1;
    bool first = true;
    for (int q = 0; q < p && first; ++q) first = tok[q] != w;
    uni += first;
    if (p + 1 < n && cap[p + 1] == cap[p]) {          // a bigram never crosses captions
      const int w2 = tok[p + 1];
      bool f2 = true;
      for (int q = 0; q < p && f2; ++q) f2 = !(tok[q] == w && q + 1 < n && cap[q + 1] == cap[q] && tok[q + 1] == w2);
      bi += f2;
    }
  }
  atomicAdd(&s_uni, uni);
  atomicAdd(&s_bi, bi);
  __syncthreads();
  if (tid == 0) {
    div1[clip] = (double)s_uni / (1e-6 + (double)n);
    div2[clip] = (double)s_bi / (1e-6 + (double)n);
  }
}

// mBLEU statistics of the K captions decoded per clip (utils/diverse_mutil.py:35-51, eval_div_stats): for every caption i
// of a clip the sufficient statistics of BLEU-1..4 with caption i as the candidate and the clip's other K-1 captions as
// the references -- out[clip][i] = {testlen, reflen (closest reference length), guess[4], correct[4]} (pycocoevalcap
// BleuScorer: cook_test / cook_refs with clipped counts).  The corpus sums over clips and the BLEU formula (40 numbers)
// are left to the caller.  One CTA per clip; an n-gram of the candidate is "correct" when it is at most the m-th
// occurrence of that n-gram in the candidate, m = the largest count of the n-gram in any single reference.
__global__ void __launch_bounds__(256) mbleu_stats_kernel(int K, int L, int start_idx, int end_idx, const long long* __restrict__ seqs,
                                                          int* __restrict__ out) {
  extern __shared__ int msm[];
  int* tok = msm;                    // [K*L] compacted tokens, caption after caption
  int* off = tok + K * L;            // [K+1] first token of caption k
  int* corr = off + K + 1;           // [K][4]
  const int clip = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int n = 0;
    for (int k = 0; k < K; ++k) {
      off[k] = n;
      for (int l = 0; l < L; ++l) {
        const int w = (int)seqs[((long long)clip * K + k) * L + l];
        if (w == end_idx) break;
        if (w == start_idx) continue;
        tok[n++] = w;
      }
    }
    off[K] = n;
  }
  for (int i = tid; i < K * 4; i += blockDim.x) corr[i] = 0;
  __syncthreads();
  const int items = K * 4 * L;       // (candidate i, order n - 1, position p)
  for (int it = tid; it < items; it += blockDim.x) {
    const int p = it % L, n = (it / L) % 4 + 1, i = it / (4 * L);
    const int len = off[i + 1] - off[i];
    if (p + n > len) continue;
    const int* g = tok + off[i] + p;
    auto same = [&](const int* a) { bool eq = true; for (int x = 0; x < n; ++x) eq = eq && a[x] == g[x]; return eq; };
    int occ = 1;
    for (int q = 0; q < p; ++q) occ += same(tok + off[i] + q) ? 1 : 0;
    int maxref = 0;
    for (int r = 0; r < K; ++r) {
      if (r == i) continue;
      const int lr = off[r + 1] - off[r];
      int c = 0;
      for (int q = 0; q + n <= lr; ++q) c += same(tok + off[r] + q) ? 1 : 0;
      maxref = c > maxref ? c : maxref;
    }
    if (occ <= maxref) atomicAdd(&corr[i * 4 + n - 1], 1);
  }
  __syncthreads();
  for (int i = tid; i < K; i += blockDim.x) {
    const int len = off[i + 1] - off[i];
    int best_d = 0x7fffffff, best_l = 0x7fffffff;       // closest reference length, ties to the shorter (min of (|l - len|, l))
    for (int r = 0; r < K; ++r) {
      if (r == i) continue;
      const int lr = off[r + 1] - off[r];
      const int dd = lr > len ? lr - len : len - lr;
      if (dd < best_d || (dd == best_d && lr < best_l)) { best_d = dd; best_l = lr; }
    }
    int* o = out + ((long long)clip * K + i) * 10;
    o[0] = len; o[1] = best_l;
    for (int n = 1; n <= 4; ++n) { o[1 + n] = len - n + 1 > 0 ? len - n + 1 : 0; o[5 + n] = corr[i * 4 + n - 1]; }
  }
}

}  // namespace acvae
