/*
 * acvae_b200.h -- C ABI of the B200-native AC-VAE latent word-decoding step.
 *
 * The reference (XinMing0411/AC-VAE) is pure Python/PyTorch and has no FFI or
 * operator registry; its boundary for this path is the Python class contract
 * of `models/vae_model.py`.  This header is the C-ABI
 * that our Python mirror of that contract (`acvae_b200/`) binds with ctypes.
 * Every entry point below names the reference code it replaces (paths are
 * relative to the upstream repository root).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (torch tensors),
 *     borrowed for the duration of the call.  fp32 unless the name says ids
 *     (int32) or seqs (int64).  Row-major; the last index is contiguous.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     performs no allocation and no host synchronisation, and is re-entrant
 *     per stream.  Scratch comes from the caller-supplied workspace whose
 *     size `acvae_*_workspace_bytes` reports.
 *   - Return value 0 = ok, <0 = error; `acvae_last_error()` returns a
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - Shape constraint inherited from the reference: E == H == Hq == prior
 *     hidden size (models/decoder.py:171, models/text_encoder.py:240-245,
 *     models/vae_model.py:693,726); E, A, Eenc multiples of 4.
 */
#ifndef ACVAE_B200_H_
#define ACVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACVAE_ABI_VERSION 2

typedef struct acvae_dims {
  int32_t N;     /* sequences in the batch (clips x captions-per-clip)        */
  int32_t Te;    /* padded encoder frames of the memory                        */
  int32_t T;     /* decode steps = max(cap_lens) - 1 (train) or max_length     */
  int32_t E;     /* embedding = latent = hidden width (E == H == Hq)           */
  int32_t A;     /* decoder attention width (decoder.py:172); the prior's is E   */
  int32_t V;     /* vocabulary                                                 */
  int32_t Eenc;  /* audio-encoder width; ln is skipped when Eenc == E and ln_w == NULL */
  int32_t L;     /* row stride (columns) of caps_ids; >= T + 1 for training    */
  int32_t mem_rep; /* sequences per clip sharing one memory row (sampling K); 1 in training */
  int32_t variant; /* 0 = Hybrid_VAEModel, 1 = VAEModel (AR posterior, no global head) */
} acvae_dims;

/* Weights, named after the reference state_dict. */
typedef struct acvae_weights {
  const float *ln_w, *ln_b;                       /* ln.weight [E,Eenc], ln.bias [E]                 vae_model.py:695-697 */
  const float *q_emb;                             /* qnet.word_embedding.weight [V,E]                text_encoder.py:24   */
  const float *q_wih[2], *q_whh[2];               /* qnet.network.weight_{ih,hh}_l0[_reverse] [3E,E] text_encoder.py:166  */
  const float *q_bih[2], *q_bhh[2];               /* qnet.network.bias_{ih,hh}_l0[_reverse] [3E]                          */
  const float *q_head_w, *q_head_b;               /* hybrid: qnet.token_mean_log [2E,2E]; vae: qnet.mean_log_out [2E,3E]  */
  const float *p_emb;                             /* pnet.word_embedding.weight [V,E]                                     */
  const float *p_attn_w, *p_attn_b, *p_attn_v;    /* pnet.word_attn.h2attn [E,2E],[E]; .v [E]        text_encoder.py:225  */
  const float *p_wih, *p_whh, *p_bih, *p_bhh;     /* pnet.network LSTM [4E,3E],[4E,E],[4E],[4E]      text_encoder.py:229  */
  const float *p_head_w, *p_head_b;               /* pnet.mean_log_out [2E,E],[2E]                   text_encoder.py:236  */
  const float *d_emb;                             /* decoder.word_embeddings.weight [V,E]            decoder.py:22        */
  const float *d_attn_w, *d_attn_b, *d_attn_v;    /* decoder.attn.h2attn [A,2E],[A]; .v [A]          decoder.py:173       */
  const float *d_wih, *d_whh, *d_bih, *d_bhh;     /* decoder.model GRU [3E,3E],[3E,E],[3E],[3E]      decoder.py:39-44     */
  const float *cls_w, *cls_b;                     /* decoder.classifier [V,E],[V]                    decoder.py:45-46     */
  const float *g_w, *g_b;                         /* mean_log_out [2E,E],[2E] (hybrid only)          vae_model.py:693     */
} acvae_weights;

/* Gradients: same fields, writable.  Backward entry points OVERWRITE them. */
typedef struct acvae_weight_grads {
  float *ln_w, *ln_b, *q_emb, *q_wih[2], *q_whh[2], *q_bih[2], *q_bhh[2], *q_head_w, *q_head_b;
  float *p_emb, *p_attn_w, *p_attn_b, *p_attn_v, *p_wih, *p_whh, *p_bih, *p_bhh, *p_head_w, *p_head_b;
  float *d_emb, *d_attn_w, *d_attn_b, *d_attn_v, *d_wih, *d_whh, *d_bih, *d_bhh, *cls_w, *cls_b, *g_w, *g_b;
} acvae_weight_grads;

/* Inputs and user-visible outputs of one training forward
 * (Hybrid_VAEModel.forward 4-input branch, vae_model.py:732-750, with the
 * encoder output given; output dict of vae_model.py:762-790, 850-869). */
typedef struct acvae_train_io {
  /* inputs */
  const float   *audio_embeds;   /* [N,Te,Eenc] encoder output (models/encoder.py:672-707 contract) */
  const int32_t *mem_lens;       /* [N] valid frames per clip (audio_embeds_lens)                   */
  const int32_t *caps_ids;       /* [N,L] caption token ids (caps.long(), <start>=1 ... <end>=2, pad 0) */
  const int32_t *cap_lens;       /* [N] caption lengths incl. <start>/<end>, sorted descending      */
  const float   *eps_q;          /* hybrid: [N,T,E] posterior noise (text_encoder.py:196); vae: [T,N,E] (:143) */
  const float   *eps_p;          /* [T,N,E] prior noise, one draw per step (text_encoder.py:259)    */
  const uint8_t *tf_flags;       /* HOST [T]: 1 = feed caps[:,t] (random.random() < ss_ratio, vae_model.py:826) */
  const uint8_t *dis_flags;      /* HOST [T]: 1 = decoder consumes the prior's z (vae_model.py:802-806) */
  /* outputs (all [N,T,*] unless noted) */
  float   *q_means, *q_logs, *q_z;       /* [N,T,E]                                   */
  float   *q_means_utt;                  /* [N,2E]  (hybrid)                          */
  float   *p_means, *p_logs, *p_z;       /* [N,T,E]                                   */
  float   *outputs;                      /* [N,T,E] decoder GRU hidden states         */
  float   *p_means_utt;                  /* [N,2E]  global-constraint head (hybrid)   */
  float   *attn_weights;                 /* [N,Te,T] decoder attention (vae_model.py:868) */
  float   *rnn_input;                    /* [N,T,3E] (vae variant, vae_model.py:187) or NULL */
  int64_t *seqs;                         /* [N,T] greedy argmax per step              */
  float   *sampled_logprobs;             /* [N,T] log-prob of the argmax              */
  float   *logit_lse;                    /* [N,T] log-sum-exp of the step's logits    */
  float   *logit_sum;                    /* [N,T] sum_j logits (label smoothing term) */
  float   *logits;                       /* [N,T,V] or NULL: materialise only on request */
} acvae_train_io;

/* Upstream gradients for the training backward. NULL = zero. */
typedef struct acvae_train_grads_in {
  const float *d_q_means, *d_q_logs, *d_q_z;     /* [N,T,E] */
  const float *d_p_means, *d_p_logs, *d_p_z;     /* [N,T,E] */
  const float *d_outputs;                        /* [N,T,E] */
  const float *d_q_means_utt, *d_p_means_utt;    /* [N,2E]  */
} acvae_train_grads_in;

const char *acvae_last_error(void);
int acvae_abi_version(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t acvae_launch_count(void);
/* Measurement hook (bench.py): record `ev_start` / `ev_stop` (cudaEvent_t, created WITH timing) on the launching
 * stream right before / after every launch of a kernel whose name contains `kernel_name` (e.g.
 * "dec_chain_fwd_kernel"); the last matching launch wins.  NULL or "" disables the probe.  Not for use under
 * CUDA-graph capture.  acvae_kernel_probe_hits() = matching launches since the probe was set. */
int acvae_set_kernel_probe(const char *kernel_name, void *ev_start, void *ev_stop);
int acvae_kernel_probe_hits(void);

/* ---- dense contraction (the building block of every batched GEMM of the step) ----------------
 * C[M,N] (ldc) = op(A) . op(B)^T [+ bias[N]] [+ C]   in fp32-grade accuracy.
 *   a_trans = 0: A is row-major [M,K] (lda);   1: A is row-major [K,M] (lda)
 *   b_trans = 0: B is row-major [N,K] (ldb) -- nn.Linear weight layout;  1: B is row-major [K,N] (ldb)
 * Large shapes run on the tcgen05 tensor cores with a 3xTF32 hi/lo split (tc_gemm.cuh); small or
 * unaligned shapes on the CUDA-core kernels.  *used_tc (host int, may be NULL) reports which. */
int acvae_gemm(int32_t M, int32_t N, int32_t K, const float *A, int64_t lda, int32_t a_trans, const float *B,
               int64_t ldb, int32_t b_trans, const float *bias, float *C, int64_t ldc, int32_t accumulate,
               int32_t *used_tc, void *stream);

/* ---- training ---------------------------------------------------------- */
size_t acvae_train_workspace_bytes(const acvae_dims *d);

/* H1: memory projection + both attentions' memory halves, once per batch.
 * Replaces vae_model.py:743-744 and the per-step re-projection of the whole
 * memory in attn_model.py:29-32 (factorised).
 * mem [Nc,Te,E], Pp [Nc,Te,E], Pd [Nc,Te,A] with Nc = N / mem_rep clips. */
int acvae_memory_prepare(const acvae_dims *d, const acvae_weights *w, const float *audio_embeds,
                         float *mem, float *Pp, float *Pd, void *stream);

/* Full training forward: H1 + H2 (posterior) + T x {H3 prior, H4 z choice,
 * H5 decoder, H6 greedy word, H7 bookkeeping} + H8 global head + vocab
 * statistics (argmax / lse / sum) without storing logits.
 * Replaces Hybrid_VAEModel.forward (vae_model.py:732-750) / VAEModel.forward.
 * Two schedules behind this one entry (same results, chosen per call from dims and flags):
 *   - hoisted (csrc/train_fast.cuh): everything that does not depend on a recurrent state batched over all N*T rows, each
 *     recurrent chain one persistent kernel (clusters / cooperative grids); hybrid variant, E == A == 256, Te <= 96, N <= 32.
 *     Free steps (tf_flags[t] == 0) and dis steps cut the chains and resume them from saved state.
 *   - general (csrc/train.cuh): one launch sequence per step; every variant and shape. */
int acvae_train_fwd(const acvae_dims *d, const acvae_weights *w, const acvae_train_io *io,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Reverse-time BPTT of acvae_train_fwd given upstream gradients; writes all
 * weight gradients except cls_w/cls_b (see acvae_vocab_ce_bwd) and
 * d_audio_embeds [N,Te,Eenc].  Replaces autograd over vae_model.py:700-869. */
int acvae_train_bwd(const acvae_dims *d, const acvae_weights *w, const acvae_train_io *io,
                    const acvae_train_grads_in *gin, acvae_weight_grads *gw, float *d_audio_embeds,
                    void *workspace, size_t workspace_bytes, void *stream);

/* ---- vocabulary projection + (label-smoothed) cross-entropy -------------
 * Replaces decoder.py:199 + pack_padded_sequence (pytorch_runner_vae.py:94-95)
 * + LabelSmoothingLoss.forward (utils/train_util.py:244-251).
 * hidden [M,E] rows (already packed or not), targets [M] int32.
 * loss_out[0] = mean over rows with weight row_w[m] (NULL = all rows)       */
size_t acvae_vocab_workspace_bytes(int32_t M, int32_t V, int32_t E);
/* Materialise logits [M,V] (compat path for callers that really want them). */
int acvae_vocab_logits(int32_t M, int32_t V, int32_t E, const float *hidden, const float *cls_w,
                       const float *cls_b, float *logits, void *stream);
/* Backward of acvae_vocab_logits given d_logits [M,V]: d_hidden [M,E], d_cls_w [V,E], d_cls_b [V]
 * (any output may be NULL). */
int acvae_vocab_logits_bwd(int32_t M, int32_t V, int32_t E, const float *hidden, const float *cls_w,
                           const float *d_logits, float *d_hidden, float *d_cls_w, float *d_cls_b,
                           void *stream);
/* Per-row log-sum-exp, sum of logits, greedy argmax (word_model.py:177-179) and its
 * log-probability, computed tile by tile without storing logits.  Outputs may be NULL. */
int acvae_vocab_stats(int32_t M, int32_t V, int32_t E, const float *hidden, const float *cls_w,
                      const float *cls_b, float *row_lse, float *row_sum, int64_t *row_argmax,
                      float *row_logprob, void *workspace, size_t workspace_bytes, void *stream);
/* Label-smoothed CE.  have_stats != 0: row_lse/row_sum are INPUTS (e.g. from acvae_train_fwd's
 * logit_lse/logit_sum gathered to the packed rows); otherwise they are computed here and returned. */
int acvae_vocab_ce_fwd(int32_t M, int32_t V, int32_t E, const float *hidden, const float *cls_w,
                       const float *cls_b, const int32_t *targets, const float *row_w, float smoothing,
                       int32_t have_stats, float *row_lse, float *row_sum, float *loss_out,
                       void *workspace, size_t workspace_bytes, void *stream);
int acvae_vocab_ce_bwd(int32_t M, int32_t V, int32_t E, const float *hidden, const float *cls_w,
                       const float *cls_b, const int32_t *targets, const float *row_w, float smoothing,
                       const float *row_lse, const float *d_loss /* device scalar */,
                       float *d_hidden, float *d_cls_w, float *d_cls_b,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- Gaussian KL (utils/train_util.py:259-266) ---------------------------
 * kl_out[0] = mean over `rows` positions of sum_d KL(q || p).              */
int acvae_kl_fwd(int64_t rows, int32_t E, const float *q_mean, const float *q_log, const float *p_mean,
                 const float *p_log, float *kl_out, void *workspace /* >= 148 floats */, size_t workspace_bytes,
                 void *stream);
int acvae_kl_bwd(int64_t rows, int32_t E, const float *q_mean, const float *q_log, const float *p_mean,
                 const float *p_log, const float *d_kl /* device scalar */, float *d_q_mean, float *d_q_log,
                 float *d_p_mean, float *d_p_log, void *stream);

/* ---- diverse sampling loop ------------------------------------------------
 * Replaces Hybrid_VAEModel.inference_forward -> stepwise_forward with
 * caps=None (vae_model.py:880-894, 700-720) and sample_next_word
 * (word_model.py:173-207).  N sequences, `mem_rep` consecutive sequences
 * share one clip's memory (audio_embeds is [N/mem_rep,Te,Eenc]).
 * method: 0 greedy, 1 multinomial-as-Gumbel-max, 2 gumbel (word_model.py:173-207).  The uniforms behind the Gumbel
 * variates are either injected (`u`, what the parity tests do: the reference draws torch.rand_like(logits) per step) or
 * drawn inside the vocabulary GEMM's epilogue from a counter-based generator (u == NULL): Philox4x32-10 keyed by
 * rng_state[0] with counter (sequence, word group, step, rng_state[1]) -- no [T,N,V] tensor exists and every call
 * increments rng_state[1] on the device, so a replayed CUDA graph draws fresh noise (csrc/gemm.cuh philox_uniform4). */
typedef struct acvae_sample_io {
  const float   *audio_embeds;   /* [N/mem_rep,Te,Eenc] */
  const int32_t *mem_lens;       /* [N/mem_rep]         */
  const float   *eps_p;          /* [T,N,E] prior noise */
  const float   *u;              /* [T,N,V] uniforms in [0,1), or NULL: greedy / drawn from rng_state */
  int32_t method;
  float   temp;
  int32_t start_idx, end_idx;
  int64_t *seqs;                 /* [N,T] END-filled after a row finishes */
  float   *sampled_logprobs;     /* [N,T] */
  float   *p_means, *p_logs, *p_z, *outputs;  /* [N,T,E] or NULL */
  int32_t *n_steps;              /* device scalar: steps executed before every row had finished */
  uint64_t *rng_state;           /* device [2] = {seed, calls}; required when method != 0 and u == NULL */
} acvae_sample_io;
size_t acvae_sample_workspace_bytes(const acvae_dims *d);
int acvae_decode_sample(const acvae_dims *d, const acvae_weights *w, const acvae_sample_io *io,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- beam search with prior latents (vae_model.py:896-995) ---------------
 * N clips, `beam` hypotheses each; eps_b [T, N*beam, E] (step-major: the caller permutes the
 * reference's per-clip draw order); seqs [N,T] = top beam. */
size_t acvae_beam_workspace_bytes(const acvae_dims *d, int32_t beam);
int acvae_beam_search(const acvae_dims *d, const acvae_weights *w, const float *audio_embeds,
                      const int32_t *mem_lens, const float *eps_b, int32_t beam, int32_t start_idx,
                      int64_t *seqs, void *workspace, size_t workspace_bytes, void *stream);

/* Deferred classifier gradients.  With the switch on, acvae_vocab_ce_bwd produces d_hidden on the caller's stream and
 * d_cls_w / d_cls_b -- which nothing in the step reads before the optimizer -- on a side stream, so the ~20 us they take no
 * longer sit between the loss and acvae_train_bwd on the step's critical path.  acvae_train_bwd, acvae_clip_adam*,
 * acvae_dp_clip_adam and acvae_join_deferred join that stream back into the stream they are given; a caller that reads the
 * two gradients any other way calls acvae_join_deferred(stream) first.  (The reference has no counterpart: autograd
 * computes the three gradients of the classifier back to back, decoder.py:199.)                                           */
int acvae_defer_classifier_grads(int32_t on);
int acvae_join_deferred(void *stream);

/* ---- fused optimizer tail --------------------------------
 * Global-norm gradient clipping (runners/pytorch_runner_vae.py:322, clip_grad_norm_ semantics:
 * coef = min(1, max_norm / (||g||_2 + 1e-6)); max_norm <= 0 disables it) chained with the Adam
 * update (:324; torch.optim.Adam, no amsgrad) over FLAT fp32 buffers of n elements (n % 4 == 0,
 * 16-byte aligned; padding elements must be zero in `grads`).  `step` is a device counter of
 * completed steps (read, then incremented), so the call is CUDA-graph capturable.
 * total_norm (device scalar, may be NULL) receives the pre-clip norm.                          */
size_t acvae_clip_adam_workspace_bytes(void);
int acvae_clip_adam(int64_t n, float *params, float *grads, float *exp_avg, float *exp_avg_sq,
                    float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int32_t *step, float *total_norm, int32_t write_clipped_grads,
                    void *workspace, size_t workspace_bytes, void *stream);
/* The same update with every hyper-parameter read from DEVICE memory when the kernel runs:
 * hyper[6] = {max_norm, lr, beta1, beta2, eps, weight_decay}.  Replaces the reference's per-iteration
 * `scheduler.step()` -> `optimizer.param_groups[i]["lr"]` path (utils/lr_scheduler.py:5-86,
 * runners/pytorch_runner_vae.py:239-257, 305) for a step that is replayed as a CUDA graph: the schedule only
 * rewrites the 24 bytes behind `hyper`, nothing is re-captured.                                            */
int acvae_clip_adam_dev(int64_t n, float *params, float *grads, float *exp_avg, float *exp_avg_sq,
                        const float *hyper, int32_t *step, float *total_norm, int32_t write_clipped_grads,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- mBLEU sufficient statistics (utils/diverse_mutil.py:35-51, eval_div_stats) --------------------
 * For every caption i of a clip, BLEU-1..4 statistics with caption i as the candidate and the clip's other
 * K-1 captions as the references (pycocoevalcap BleuScorer with option "closest"):
 * stats[clip][i][10] = {testlen, reflen, guess[4], correct[4]}.  Captions are the ids up to the first
 * <end>, <start> skipped (runners/base_runner.py:146-157).  seqs [clips, K, L] int64, K >= 2.           */
int acvae_mbleu_stats(int32_t clips, int32_t K, int32_t L, const int64_t *seqs, int32_t start_idx,
                      int32_t end_idx, int32_t *stats, void *stream);

/* ---- encoder hand-off -------------------------------------------------------
 * Replaces the tail of Cnn10.forward, models/encoder.py:691 `x = torch.mean(x, dim=3)` and :700
 * `x = x.transpose(1, 2).contiguous()`, by ONE pass over the last convolution block's output
 * fmap [N, C, Te, F] (contiguous) that writes audio_embeds [N, Te, C] -- the row-major frame memory
 * acvae_memory_prepare / acvae_train_fwd / acvae_decode_sample read through TMA.  pooled (optional,
 * [N, C]) = max over frames + mean over frames of the same means (:693-695, the operand of
 * `embed_pooled`; the VAE decoder ignores it).  bwd: d_fmap[n,c,j,f] = d_audio_embeds[n,j,c] / F.     */
int acvae_encoder_handoff_fwd(int32_t N, int32_t C, int32_t Te, int32_t F, const float *fmap,
                              float *audio_embeds, float *pooled, void *stream);
int acvae_encoder_handoff_bwd(int32_t N, int32_t C, int32_t Te, int32_t F, const float *d_audio_embeds,
                              float *d_fmap, void *stream);

/* ---- data-parallel optimizer tail as one fused compute + collective over NVLink peer memory --------
 * Replaces, for one process per GPU on one node: DistributedDataParallel's gradient all-reduce
 * (runners/pytorch_runner_vae.py:204-207, inside loss.backward() :321) + clip_grad_norm_ (:322) +
 * optimizer.step() (:324).  Every rank owns 1/world of the flat buffers: it averages ITS shard of the
 * gradients reading the peers' buffers over NVLink (fixed rank order), publishes the shard's sum of
 * squares, clips with the global norm, runs Adam on the shard and writes the new parameters into
 * every rank's parameter buffer.  grads / params / comm: HOST arrays of `world` device pointers, entry
 * [rank] local, the others mapped with acvae_ipc_open from the peers' acvae_ipc_export handles; comm
 * blocks (acvae_dp_comm_bytes each) and the workspace must be zero before the first call.
 * n = floats per flat buffer (multiple of 4*world); grad_shard / exp_avg / exp_avg_sq: local [n/world].
 * hyper: device {max_norm, lr, beta1, beta2, eps, weight_decay}; step as in acvae_clip_adam.          */
int acvae_ipc_export(const void *ptr, void *handle64, int64_t *offset);
int acvae_ipc_open(const void *handle64, int64_t offset, void **peer_ptr);
/* Unmaps every peer allocation acvae_ipc_open has mapped in this process (mappings are cached per handle: a handle names a
 * whole allocation and two tensors of a peer may share one).  Call after the last step, before the peers free their buffers. */
int acvae_ipc_close_all(void);
size_t acvae_dp_comm_bytes(void);
size_t acvae_dp_workspace_bytes(void);
int acvae_dp_clip_adam(int32_t world, int32_t rank, int64_t n, const void *const *grads,
                       void *const *params, void *const *comm, float *grad_shard, float *exp_avg,
                       float *exp_avg_sq, const float *hyper, int32_t *step, float *total_norm,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- loss composition (runners/pytorch_runner_vae.py:315-320) as one node ------------------
 * terms[4] = {loss, ce, kl, mse}; loss = ce + kl_weight*kl + alpha*mean((q_utt - p_utt)^2) (nn.MSELoss, :318).
 * ce / kl are the device scalars of acvae_vocab_ce_fwd / acvae_kl_fwd.  q_utt == p_utt == NULL: no global term.
 * bwd: d_q_utt / d_p_utt [n]; scal[2] = {d_loss, d_loss*kl_weight} = the device scalars acvae_vocab_ce_bwd /
 * acvae_kl_bwd take as d_loss / d_kl.                                                                       */
int acvae_loss_combine_fwd(int64_t n, const float *q_utt, const float *p_utt, const float *ce, const float *kl,
                           float kl_weight, float alpha, float *terms, void *stream);
int acvae_loss_combine_bwd(int64_t n, const float *q_utt, const float *p_utt, const float *d_loss, float kl_weight,
                           float alpha, float *d_q_utt, float *d_p_utt, float *scal, void *stream);

/* ---- diverse beam search with prior latents --------------------------------------------------
 * Replaces CaptionModel.diverse_beam_search (models/word_model.py:297-394) driven by
 * Hybrid_VAEModel.dbs_step / prepare_dbs_decoder_input / dbs_process_step (models/vae_model.py:997-1048).
 * N clips; group_size groups of bdash = beam_size / group_size hypotheses, group g running g steps behind
 * group 0; d->T = max_length.  eps_g [T + group_size - 1, N*group_size*bdash, E]: prior noise of global step t
 * for row (clip*group_size + g)*bdash + k (the caller permutes the reference's draw order; rows of groups
 * inactive at t are ignored).  seqs [N, group_nbest ? beam_size : group_size, T], END-filled.           */
size_t acvae_dbs_workspace_bytes(const acvae_dims *d, int32_t beam_size, int32_t group_size);
int acvae_diverse_beam_search(const acvae_dims *d, const acvae_weights *w, const float *audio_embeds,
                              const int32_t *mem_lens, const float *eps_g, int32_t beam_size, int32_t group_size,
                              float diversity_lambda, float temperature, int32_t group_nbest, int32_t start_idx,
                              int32_t end_idx, int64_t *seqs, void *workspace, size_t workspace_bytes, void *stream);

/* ---- diversity statistics of the decoded captions -------------------------
 * Replaces utils/div_utils.py:11-44 (compute_div_n for n = 1, 2; compute_global_div_n for n = 1) as used by
 * utils/diverse_mutil.py:25-29, on the id tensor seqs [clips, K, L] of the sampling loop (a caption = ids up to the
 * first <end>, <start> skipped: runners/base_runner.py:146-157).  div1 / div2 [clips] (fp64, as numpy computes them):
 * distinct uni- / bigrams of the clip's K captions over its token count; vocab_flags [V] (may be NULL; zeroed by the
 * caller): set to 1 for every word that occurs -- their sum is gDiv-1. */
int acvae_diversity_stats(int32_t clips, int32_t K, int32_t L, int32_t V, const int64_t *seqs, int32_t start_idx,
                          int32_t end_idx, double *div1, double *div2, int32_t *vocab_flags, void *stream);

/* ---- overlapping the input copy --------------------------------------------------------------
 * `cuda_event` (a cudaEvent_t, or NULL to switch off): recorded by the caller after the host-to-device copy of a step's
 * audio embeddings, on whatever stream carries it.  acvae_train_fwd waits for it (cudaEventWaitExternal: an external
 * event-wait node under stream capture) right before the first kernel that reads the audio, so the copy runs under the
 * posterior chain, which does not.  The caller keeps the usual WAR discipline on the audio buffer. */
int acvae_set_input_event(void *cuda_event);
/* Data-parallel overlap (reference: DistributedDataParallel's bucketed all-reduce inside loss.backward(),
 * runners/pytorch_runner_vae.py:204-207, 321): an event every acvae_train_bwd records at the point where all
 * decoder.* weight gradients (word embeddings, GRU, attention; the classifier's are final before the call)
 * are final, so their all-reduce can run under the rest of the backward.  NULL switches it off.          */
int acvae_set_bucket_event(void *cuda_event);

/* profiling only: [T][16] int64 device buffer for clock64 stamps of the decoder forward chain (CTA 0), or NULL */
int acvae_debug_set_chain_trace(void *device_buffer);

/* ---- arithmetic mode of the batched contractions (process-wide) ------------------------------
 * 0 (default): fp32-grade products on the tensor cores (three kind::tf32 MMAs per k-step, chunked fp32
 * accumulation) -- the mode of the 1e-4 parity tests.  1: single-pass TF32 (one MMA per k-step, no operand split):
 * the reduced-precision class BASELINE.json quotes as "bf16" (2e-2).  Recurrent chains, attention and pointwise
 * arithmetic are fp32 in both modes. */
int acvae_set_precision(int32_t mode);
int acvae_get_precision(void);

#ifdef __cplusplus
}
#endif
#endif /* ACVAE_B200_H_ */
