// C ABI of libacvae_b200.so (see include/acvae_b200.h).  Thin: argument checks,
// workspace carve-up and kernel enqueueing; no torch types, no exceptions.
#include "../../include/acvae_b200.h"
#include <stdlib.h>

#include <map>
#include <mutex>
#include <string>

#include "dp_optim.cuh"
#include "handoff.cuh"
#include "optim.cuh"
#include "sample.cuh"
#include "train_fast.cuh"

namespace acvae {
thread_local char g_err[512] = {0};
std::atomic<unsigned long long> g_launches{0};
KernelProbe g_probe{};

struct VocabWs {
  float *pmax, *pexp, *psum, *pbest; int* parg;
  float *row_loss, *row_cnt, *scal, *dlogits;
  size_t bytes;
};
static VocabWs carve_vocab_ws(int M, int V, void* base, bool with_dlogits) {
  Arena ar(base);
  VocabWs w{};
  const size_t nt = (V + kVocabTile - 1) / kVocabTile;
  w.pmax = ar.take<float>((size_t)M * nt); w.pexp = ar.take<float>((size_t)M * nt); w.psum = ar.take<float>((size_t)M * nt);
  w.pbest = ar.take<float>((size_t)M * nt * 2); w.parg = ar.take<int>((size_t)M * nt);
  w.row_loss = ar.take<float>(M); w.row_cnt = ar.take<float>(M); w.scal = ar.take<float>(8);
  if (with_dlogits) w.dlogits = ar.take<float>((size_t)M * V);
  w.bytes = ar.off;
  return w;
}

}  // namespace acvae

using namespace acvae;

extern "C" {

const char* acvae_last_error(void) { return g_err; }
int acvae_abi_version(void) { return ACVAE_ABI_VERSION; }
uint64_t acvae_launch_count(void) { return g_launches.load(); }

int acvae_set_kernel_probe(const char* kernel_name, void* ev_start, void* ev_stop) {
  if (!kernel_name || !kernel_name[0]) { g_probe.active = 0; return 0; }
  ACVAE_REQUIRE(ev_start && ev_stop, "kernel probe needs two CUDA events");
  ACVAE_REQUIRE(strlen(kernel_name) < sizeof(g_probe.name), "kernel name too long");
  strcpy(g_probe.name, kernel_name);
  g_probe.e0 = (cudaEvent_t)ev_start; g_probe.e1 = (cudaEvent_t)ev_stop;
  g_probe.hits = 0;
  g_probe.active = 1;
  return 0;
}
int acvae_kernel_probe_hits(void) { return g_probe.hits; }

int acvae_gemm(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, int32_t a_trans, const float* B, int64_t ldb,
               int32_t b_trans, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t* used_tc, void* stream) {
  ACVAE_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, "bad argument");
  GemmParams p{};
  p.M = M; p.U = N; p.G = 1; p.nseg = 1;
  GemmSeg s{};
  s.a = A; s.lda = lda; s.a_trans = a_trans; s.w[0] = B; s.ldw = ldb; s.w_trans = b_trans; s.K = K;
  p.seg[0] = s;
  p.epi.c[0] = C; p.epi.ldc = ldc; p.epi.bias[0] = bias; p.epi.scale = 1.0f; p.epi.accumulate = accumulate;
  int tc = 0;
  ACVAE_TRY(launch_gemm<EPI_PLAIN>(p, (cudaStream_t)stream, &tc));
  if (used_tc) *used_tc = tc ? 1 : 0;
  return 0;
}

size_t acvae_train_workspace_bytes(const acvae_dims* d) {
  if (check_dims(d) != 0) return 0;
  return carve_train_ws(*d, nullptr).bytes;
}

int acvae_memory_prepare(const acvae_dims* d, const acvae_weights* w, const float* audio_embeds, float* mem, float* Pp,
                         float* Pd, void* stream) {
  ACVAE_TRY(check_dims(d));
  ACVAE_REQUIRE(w && audio_embeds && mem && Pp && Pd, "NULL pointer");
  ACVAE_TRY(wait_input_event((cudaStream_t)stream));
  return memory_prepare(*d, *w, audio_embeds, mem, Pp, Pd, (cudaStream_t)stream);
}

int acvae_train_fwd(const acvae_dims* d, const acvae_weights* w, const acvae_train_io* io, void* workspace,
                    size_t workspace_bytes, void* stream) {
  ACVAE_TRY(check_dims(d));
  ACVAE_REQUIRE(w && io && workspace, "NULL pointer");
  ACVAE_REQUIRE(d->mem_rep == 1, "training requires mem_rep == 1");
  ACVAE_REQUIRE(d->T <= 64, "T > 64 decode steps not supported");
  ACVAE_REQUIRE(d->L >= d->T + 1, "caps row stride L must be >= T + 1");
  ACVAE_REQUIRE(workspace_bytes >= carve_train_ws(*d, nullptr).bytes, "workspace too small");
  ACVAE_REQUIRE(io->tf_flags && io->dis_flags, "tf_flags / dis_flags are required (host arrays of T bytes)");
  ACVAE_REQUIRE(io->audio_embeds && io->mem_lens && io->caps_ids && io->cap_lens && io->eps_q && io->eps_p, "NULL input");
  ACVAE_REQUIRE(io->q_means && io->q_logs && io->q_z && io->p_means && io->p_logs && io->p_z && io->outputs &&
                    io->seqs && io->sampled_logprobs && io->logit_lse && io->logit_sum,
                "NULL output");
  ACVAE_REQUIRE(d->variant == 1 || (io->q_means_utt && io->p_means_utt), "hybrid variant needs *_means_utt outputs");
  if (!getenv("ACVAE_DISABLE_FAST") && fast_path_ok(*d, *io)) return train_fwd_fast(*d, *w, *io, workspace, (cudaStream_t)stream);
  return train_fwd(*d, *w, *io, workspace, (cudaStream_t)stream);
}

int acvae_train_bwd(const acvae_dims* d, const acvae_weights* w, const acvae_train_io* io,
                    const acvae_train_grads_in* gin, acvae_weight_grads* gw, float* d_audio_embeds, void* workspace,
                    size_t workspace_bytes, void* stream) {
  ACVAE_TRY(check_dims(d));
  ACVAE_REQUIRE(w && io && gin && gw && workspace, "NULL pointer");
  ACVAE_REQUIRE(workspace_bytes >= carve_train_ws(*d, nullptr).bytes, "workspace too small");
  ACVAE_REQUIRE(io->tf_flags && io->dis_flags, "tf_flags / dis_flags are required");
  if (!getenv("ACVAE_DISABLE_FAST") && fast_path_ok(*d, *io)) {
    ACVAE_TRY(train_bwd_fast(*d, *w, *io, *gin, *gw, d_audio_embeds, workspace, (cudaStream_t)stream));
    return join_deferred_cls_grads((cudaStream_t)stream);     // (already joined through the fan: clears the pending mark)
  }
  ACVAE_TRY(join_deferred_cls_grads((cudaStream_t)stream));
  ACVAE_TRY(train_bwd(*d, *w, *io, *gin, *gw, d_audio_embeds, workspace, (cudaStream_t)stream));
  if (bucket_event()) ACVAE_CHECK(cudaEventRecord(bucket_event(), (cudaStream_t)stream));
  return 0;
}

size_t acvae_vocab_workspace_bytes(int32_t M, int32_t V, int32_t E) {
  (void)E;
  if (M <= 0 || V <= 0) return 0;
  return carve_vocab_ws(M, V, nullptr, true).bytes;
}

int acvae_vocab_logits(int32_t M, int32_t V, int32_t E, const float* hidden, const float* cls_w, const float* cls_b,
                       float* logits, void* stream) {
  ACVAE_REQUIRE(M > 0 && V > 0 && E > 0 && hidden && cls_w && logits, "bad argument");
  return linear_fwd(M, V, E, hidden, E, cls_w, E, cls_b, logits, V, (cudaStream_t)stream);
}

int acvae_vocab_logits_bwd(int32_t M, int32_t V, int32_t E, const float* hidden, const float* cls_w,
                           const float* d_logits, float* d_hidden, float* d_cls_w, float* d_cls_b, void* stream) {
  ACVAE_REQUIRE(M > 0 && V > 0 && E > 0 && hidden && cls_w && d_logits, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_hidden) ACVAE_TRY(linear_bwd_data(M, E, V, d_logits, V, cls_w, E, d_hidden, E, st));
  if (d_cls_w) ACVAE_TRY(linear_bwd_weight(V, E, M, d_logits, V, hidden, E, d_cls_w, E, st));
  if (d_cls_b) ACVAE_TRY(colsum(M, V, d_logits, V, d_cls_b, st));
  return 0;
}

int acvae_vocab_stats(int32_t M, int32_t V, int32_t E, const float* hidden, const float* cls_w, const float* cls_b,
                      float* row_lse, float* row_sum, int64_t* row_argmax, float* row_logprob, void* workspace,
                      size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(M > 0 && V > 0 && E > 0 && hidden && cls_w && workspace, "bad argument");
  ACVAE_REQUIRE(workspace_bytes >= carve_vocab_ws(M, V, nullptr, false).bytes, "workspace too small");
  VocabWs ws = carve_vocab_ws(M, V, workspace, false);
  VocabStatsArgs v{};
  v.M = M; v.V = V; v.E = E; v.hidden = hidden; v.ld_h = E; v.cls_w = cls_w; v.cls_b = cls_b;
  v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
  v.red.lse = row_lse; v.red.lsum = row_sum; v.red.logprob = row_logprob; v.red.ld_row = 1;
  v.red.seqs = (long long*)row_argmax; v.red.ld_seqs = 1;
  return vocab_stats(v, (cudaStream_t)stream);
}

int acvae_vocab_ce_fwd(int32_t M, int32_t V, int32_t E, const float* hidden, const float* cls_w, const float* cls_b,
                       const int32_t* targets, const float* row_w, float smoothing, int32_t have_stats, float* row_lse,
                       float* row_sum, float* loss_out, void* workspace, size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(M > 0 && V > 1 && E > 0 && hidden && cls_w && cls_b && targets && row_lse && row_sum && loss_out && workspace,
                "bad argument");
  ACVAE_REQUIRE(workspace_bytes >= carve_vocab_ws(M, V, nullptr, false).bytes, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  VocabWs ws = carve_vocab_ws(M, V, workspace, false);
  if (!have_stats) {
    VocabStatsArgs v{};
    v.M = M; v.V = V; v.E = E; v.hidden = hidden; v.ld_h = E; v.cls_w = cls_w; v.cls_b = cls_b;
    v.pmax = ws.pmax; v.pexp = ws.pexp; v.psum = ws.psum; v.pbest = ws.pbest; v.parg = ws.parg;
    v.red.lse = row_lse; v.red.lsum = row_sum; v.red.ld_row = 1;
    ACVAE_TRY(vocab_stats(v, st));
  }
  CeRowsParams c{};
  c.M = M; c.V = V; c.E = E; c.hidden = hidden; c.ld_h = E; c.cls_w = cls_w; c.cls_b = cls_b; c.targets = targets;
  c.row_w = row_w; c.lse = row_lse; c.lsum = row_sum;
  c.on = 1.0f - smoothing; c.off = smoothing / (float)(V - 1);   // utils/train_util.py:237,249-250
  c.row_loss = ws.row_loss; c.row_cnt = ws.row_cnt;
  ACVAE_LAUNCH(ce_rows_kernel, (M + 7) / 8, 256, 0, st, c);
  ACVAE_LAUNCH(final_sum_kernel, 1, 256, 0, st, M, (const float*)ws.row_cnt, 1.0f, (const float*)nullptr, ws.scal);
  ACVAE_LAUNCH(final_sum_kernel, 1, 256, 0, st, M, (const float*)ws.row_loss, 1.0f, (const float*)ws.scal, loss_out);
  return 0;
}

int acvae_vocab_ce_bwd(int32_t M, int32_t V, int32_t E, const float* hidden, const float* cls_w, const float* cls_b,
                       const int32_t* targets, const float* row_w, float smoothing, const float* row_lse,
                       const float* d_loss, float* d_hidden, float* d_cls_w, float* d_cls_b, void* workspace,
                       size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(M > 0 && V > 1 && E > 0 && hidden && cls_w && cls_b && targets && row_lse && d_loss && workspace, "bad argument");
  ACVAE_REQUIRE(workspace_bytes >= carve_vocab_ws(M, V, nullptr, true).bytes, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  VocabWs ws = carve_vocab_ws(M, V, workspace, true);
  ACVAE_LAUNCH(ce_gscale_kernel, 1, 256, 0, st, M, row_w, d_loss, ws.scal);
  GemmParams p{};
  p.M = M; p.U = V; p.G = 1; p.nseg = 1;
  p.seg[0] = seg_plain(hidden, E, cls_w, E, E);
  p.epi.bias[0] = cls_b; p.epi.c[0] = ws.dlogits; p.epi.ldc = V;
  p.epi.lse = row_lse; p.epi.targets = targets; p.epi.row_w = row_w; p.epi.gscale = ws.scal;
  p.epi.smooth_on = 1.0f - smoothing; p.epi.smooth_off = smoothing / (float)(V - 1);
  ACVAE_TRY(launch_gemm<EPI_DLOGITS>(p, st));
  cudaStream_t sw = st;
  if (cls_defer_flag() && (d_cls_w || d_cls_b)) {
    // the weight / bias gradients feed nothing before the optimizer: off the caller's stream (which carries the step's critical
    // path into acvae_train_bwd), joined back by acvae_train_bwd / the optimizer entries
    Aux* ax = aux();
    ACVAE_REQUIRE(ax, "no side streams");
    sw = ax->s[kAuxFan0 + 7];
    ACVAE_TRY(stream_dep(st, sw, ax));
    *cls_defer_pending() = true;
  }
  if (d_hidden) ACVAE_TRY(linear_bwd_data(M, E, V, ws.dlogits, V, cls_w, E, d_hidden, E, st));
  if (d_cls_w) ACVAE_TRY(linear_bwd_weight(V, E, M, ws.dlogits, V, hidden, E, d_cls_w, E, sw));
  if (d_cls_b) ACVAE_TRY(colsum(M, V, ws.dlogits, V, d_cls_b, sw));
  return 0;
}

int acvae_defer_classifier_grads(int32_t on) { cls_defer_flag() = on != 0; return 0; }
int acvae_join_deferred(void* stream) { return join_deferred_cls_grads((cudaStream_t)stream); }

int acvae_kl_fwd(int64_t rows, int32_t E, const float* q_mean, const float* q_log, const float* p_mean,
                 const float* p_log, float* kl_out, void* workspace, size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(rows > 0 && E > 0 && q_mean && q_log && p_mean && p_log && kl_out && workspace, "bad argument");
  ACVAE_REQUIRE(workspace_bytes >= 148 * sizeof(float), "workspace too small (need 148 floats)");
  const long long n = (long long)rows * E;
  int blocks = (int)((n + 2047) / 2048);
  blocks = blocks < 1 ? 1 : (blocks > 148 ? 148 : blocks);
  // deterministic two-stage sum: per-block partials, then one block adds them in a fixed order
  ACVAE_LAUNCH(kl_partial_kernel, blocks, 256, 0, (cudaStream_t)stream, n, q_mean, q_log, p_mean, p_log, (float*)workspace);
  ACVAE_LAUNCH(final_sum_kernel, 1, 256, 0, (cudaStream_t)stream, blocks, (const float*)workspace, 1.0f / (float)rows,
               (const float*)nullptr, kl_out);
  return 0;
}

int acvae_kl_bwd(int64_t rows, int32_t E, const float* q_mean, const float* q_log, const float* p_mean,
                 const float* p_log, const float* d_kl, float* d_q_mean, float* d_q_log, float* d_p_mean,
                 float* d_p_log, void* stream) {
  ACVAE_REQUIRE(rows > 0 && E > 0 && q_mean && q_log && p_mean && p_log && d_kl && d_q_mean && d_q_log && d_p_mean && d_p_log,
                "bad argument");
  const long long n = (long long)rows * E;
  ACVAE_LAUNCH(kl_bwd_kernel, grid1d(n), 256, 0, (cudaStream_t)stream, n, 1.0f / (float)rows, q_mean, q_log, p_mean,
               p_log, d_kl, d_q_mean, d_q_log, d_p_mean, d_p_log);
  return 0;
}

size_t acvae_sample_workspace_bytes(const acvae_dims* d) {
  if (check_dims(d) != 0) return 0;
  return carve_sample_ws(*d, nullptr).bytes;
}

int acvae_decode_sample(const acvae_dims* d, const acvae_weights* w, const acvae_sample_io* io, void* workspace,
                        size_t workspace_bytes, void* stream) {
  ACVAE_TRY(check_dims(d));
  ACVAE_REQUIRE(w && io && workspace, "NULL pointer");
  ACVAE_REQUIRE(io->audio_embeds && io->mem_lens && io->eps_p && io->seqs && io->sampled_logprobs, "NULL pointer in io");
  ACVAE_REQUIRE(io->method == 0 || io->u || io->rng_state, "method sample/gumbel needs uniform noise u or an rng_state to draw it from");
  ACVAE_REQUIRE(io->method >= 0 && io->method <= 2, "unknown sampling method");
  ACVAE_REQUIRE(io->temp > 0.0f, "temp must be positive");
  ACVAE_REQUIRE(workspace_bytes >= carve_sample_ws(*d, nullptr).bytes, "workspace too small");
  return decode_sample(*d, *w, *io, workspace, (cudaStream_t)stream);
}

size_t acvae_beam_workspace_bytes(const acvae_dims* d, int32_t beam) {
  if (check_dims(d) != 0 || beam <= 0) return 0;
  return carve_beam_ws(*d, beam, nullptr).bytes;
}

int acvae_beam_search(const acvae_dims* d, const acvae_weights* w, const float* audio_embeds, const int32_t* mem_lens,
                      const float* eps_b, int32_t beam, int32_t start_idx, int64_t* seqs, void* workspace,
                      size_t workspace_bytes, void* stream) {
  ACVAE_TRY(check_dims(d));
  ACVAE_REQUIRE(w && audio_embeds && mem_lens && eps_b && seqs && workspace, "NULL pointer");
  ACVAE_REQUIRE(beam >= 1 && beam <= 32, "beam must be in [1, 32]");
  ACVAE_REQUIRE(d->mem_rep == 1, "beam search takes one row per clip (mem_rep == 1)");
  ACVAE_REQUIRE(workspace_bytes >= carve_beam_ws(*d, beam, nullptr).bytes, "workspace too small");
  return beam_search(*d, *w, audio_embeds, mem_lens, eps_b, beam, start_idx, seqs, workspace, (cudaStream_t)stream);
}

size_t acvae_clip_adam_workspace_bytes(void) { return sizeof(float) * kOptBlocks; }

static int clip_adam_impl(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, float max_norm, float lr,
                          float beta1, float beta2, float eps, float weight_decay, const float* hyper, int32_t* step,
                          float* total_norm, int32_t write_clipped_grads, void* workspace, size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(n > 0 && n % 4 == 0, "n must be a positive multiple of 4 (pad the flat buffers)");
  ACVAE_REQUIRE(params && grads && exp_avg && exp_avg_sq && step && workspace, "NULL pointer");
  ACVAE_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq), "flat buffers must be 16-byte aligned");
  ACVAE_REQUIRE(workspace_bytes >= sizeof(float) * kOptBlocks, "workspace too small");
  ACVAE_TRY(join_deferred_cls_grads((cudaStream_t)stream));
  const long long n4 = n / 4;
  int blocks = (int)((n4 + kOptThreads - 1) / kOptThreads);
  blocks = blocks < 1 ? 1 : (blocks > kOptBlocks ? kOptBlocks : blocks);
  cudaStream_t st = (cudaStream_t)stream;
  ACVAE_LAUNCH(sumsq_partial_kernel, blocks, kOptThreads, 0, st, n4, (const float4*)grads, (float*)workspace);
  ClipAdamParams a{};
  a.n4 = n4; a.p = (float4*)params; a.g = (float4*)grads; a.m = (float4*)exp_avg; a.v = (float4*)exp_avg_sq;
  a.partial = (const float*)workspace; a.npartial = blocks; a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2;
  a.eps = eps; a.weight_decay = weight_decay; a.hyper = hyper; a.step = step; a.total_norm = total_norm;
  a.write_grad = write_clipped_grads;
  ACVAE_LAUNCH(clip_adam_kernel, blocks, kOptThreads, 0, st, a);
  ACVAE_LAUNCH(step_advance_kernel, 1, 1, 0, st, step);
  return 0;
}

int acvae_clip_adam(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, float max_norm, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int32_t* step, float* total_norm,
                    int32_t write_clipped_grads, void* workspace, size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(lr >= 0.0f && beta1 >= 0.0f && beta1 < 1.0f && beta2 >= 0.0f && beta2 < 1.0f && eps >= 0.0f, "bad hyper-parameter");
  return clip_adam_impl(n, params, grads, exp_avg, exp_avg_sq, max_norm, lr, beta1, beta2, eps, weight_decay, nullptr, step,
                        total_norm, write_clipped_grads, workspace, workspace_bytes, stream);
}

// The same update with every hyper-parameter read from DEVICE memory at run time: hyper[6] = {max_norm, lr, beta1, beta2,
// eps, weight_decay}.  A captured CUDA graph of the step then follows an LR schedule (utils/lr_scheduler.py:5-86, stepped
// every iteration by the runner, pytorch_runner_vae.py:241-257, 305) without re-capture: the host only rewrites the
// (pinned) source of the 24-byte copy that feeds `hyper`.
int acvae_clip_adam_dev(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, const float* hyper,
                        int32_t* step, float* total_norm, int32_t write_clipped_grads, void* workspace,
                        size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(hyper, "hyper (device vector of 6 floats) is NULL");
  return clip_adam_impl(n, params, grads, exp_avg, exp_avg_sq, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, hyper, step, total_norm,
                        write_clipped_grads, workspace, workspace_bytes, stream);
}

// ---- CUDA IPC plumbing of the fused data-parallel optimizer (one process per GPU on one node) ----
typedef CUresult (*PFN_getRange)(CUdeviceptr*, size_t*, CUdeviceptr);
static PFN_getRange addr_range_fn() {
  static PFN_getRange fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_getRange>(f);
  }
  return fn;
}

int acvae_ipc_export(const void* ptr, void* handle64, int64_t* offset) {
  ACVAE_REQUIRE(ptr && handle64 && offset, "NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PFN_getRange fn = addr_range_fn();
  ACVAE_REQUIRE(fn, "cuMemGetAddressRange is not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) return set_error("cuMemGetAddressRange", "pointer is not a device allocation");
  cudaIpcMemHandle_t h;
  ACVAE_CHECK(cudaIpcGetMemHandle(&h, (void*)base));          // fails for stream-ordered / VMM allocations (expandable segments)
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((CUdeviceptr)ptr - base);
  return 0;
}

// A handle names a whole allocation and may be opened once per process: two tensors of a peer that live in the same allocation
// (torch's caching allocator carves many tensors out of one cudaMalloc block) share the mapping.
namespace {
struct IpcCache {
  std::mutex mu;
  std::map<std::string, void*> open;      // 64-byte handle -> mapped base
};
IpcCache& ipc_cache() { static IpcCache c; return c; }
}  // namespace

int acvae_ipc_open(const void* handle64, int64_t offset, void** peer_ptr) {
  ACVAE_REQUIRE(handle64 && peer_ptr && offset >= 0, "bad argument");
  IpcCache& c = ipc_cache();
  std::lock_guard<std::mutex> g(c.mu);
  const std::string key(static_cast<const char*>(handle64), 64);
  auto it = c.open.find(key);
  void* base = nullptr;
  if (it != c.open.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    ACVAE_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c.open.emplace(key, base);
  }
  *peer_ptr = static_cast<char*>(base) + offset;
  return 0;
}

int acvae_ipc_close_all(void) {
  IpcCache& c = ipc_cache();
  std::lock_guard<std::mutex> g(c.mu);
  for (auto& kv : c.open) cudaIpcCloseMemHandle(kv.second);
  c.open.clear();
  return 0;
}

size_t acvae_dp_comm_bytes(void) { return align_up(sizeof(DpComm)); }
size_t acvae_dp_workspace_bytes(void) { return sizeof(float) * kDpBlocks + 64; }

int acvae_dp_clip_adam(int32_t world, int32_t rank, int64_t n, const void* const* grads, void* const* params, void* const* comm,
                       float* grad_shard, float* exp_avg, float* exp_avg_sq, const float* hyper, int32_t* step, float* total_norm,
                       void* workspace, size_t workspace_bytes, void* stream) {
  ACVAE_REQUIRE(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, "bad world / rank");
  ACVAE_REQUIRE(n > 0 && n % (4LL * world) == 0, "n must be a positive multiple of 4 * world (pad the flat buffers)");
  ACVAE_REQUIRE(grads && params && comm && grad_shard && exp_avg && exp_avg_sq && hyper && step && workspace, "NULL pointer");
  ACVAE_REQUIRE(workspace_bytes >= acvae_dp_workspace_bytes(), "workspace too small");
  ACVAE_TRY(join_deferred_cls_grads((cudaStream_t)stream));
  DpParams a{};
  a.world = world; a.rank = rank; a.n4 = n / 4 / world;
  for (int q = 0; q < world; ++q) {
    ACVAE_REQUIRE(grads[q] && params[q] && comm[q] && aligned16(grads[q]) && aligned16(params[q]), "peer buffers must be non-NULL and 16-byte aligned");
    a.grads[q] = (const float4*)grads[q]; a.params[q] = (float4*)params[q]; a.comm[q] = (DpComm*)comm[q];
  }
  ACVAE_REQUIRE((a.n4 * 16) % 16 == 0 && aligned16(grad_shard) && aligned16(exp_avg) && aligned16(exp_avg_sq), "shard buffers must be 16-byte aligned");
  a.gshard = (float4*)grad_shard; a.m = (float4*)exp_avg; a.v = (float4*)exp_avg_sq;
  a.partial = (float*)workspace; a.ticket = (unsigned*)((char*)workspace + sizeof(float) * kDpBlocks);
  a.hyper = hyper; a.step = step; a.total_norm = total_norm;
  int blocks = (int)((a.n4 + kDpThreads - 1) / kDpThreads);
  blocks = blocks < 1 ? 1 : (blocks > kDpBlocks ? kDpBlocks : blocks);
  cudaStream_t st = (cudaStream_t)stream;
  ACVAE_LAUNCH(dp_reduce_kernel, blocks, kDpThreads, 0, st, a);
  ACVAE_LAUNCH(dp_adam_kernel, blocks, kDpThreads, 0, st, a);
  ACVAE_LAUNCH(step_advance_kernel, 1, 1, 0, st, step);
  return 0;
}

int acvae_loss_combine_fwd(int64_t n, const float* q_utt, const float* p_utt, const float* ce, const float* kl, float kl_weight,
                           float alpha, float* terms, void* stream) {
  ACVAE_REQUIRE(ce && kl && terms, "NULL pointer");
  ACVAE_REQUIRE((q_utt == nullptr) == (p_utt == nullptr) && (q_utt == nullptr || n > 0), "bad global-term arguments");
  ACVAE_LAUNCH(loss_combine_kernel, 1, 1024, 0, (cudaStream_t)stream, (long long)n, q_utt, p_utt, ce, kl, kl_weight, alpha, terms);
  return 0;
}

int acvae_loss_combine_bwd(int64_t n, const float* q_utt, const float* p_utt, const float* d_loss, float kl_weight, float alpha,
                           float* d_q_utt, float* d_p_utt, float* scal, void* stream) {
  ACVAE_REQUIRE(d_loss && scal, "NULL pointer");
  ACVAE_REQUIRE(q_utt == nullptr || (p_utt && d_q_utt && d_p_utt && n > 0), "bad global-term arguments");
  const long long cnt = q_utt ? n : 1;
  ACVAE_LAUNCH(loss_combine_bwd_kernel, grid1d(cnt), 256, 0, (cudaStream_t)stream, (long long)n, q_utt, p_utt, d_loss, kl_weight,
               alpha, d_q_utt, d_p_utt, scal);
  return 0;
}

size_t acvae_dbs_workspace_bytes(const acvae_dims* d, int32_t beam_size, int32_t group_size) {
  if (check_dims(d) != 0 || group_size <= 0 || beam_size < group_size) return 0;
  return carve_dbs_ws(*d, group_size, beam_size / group_size, nullptr).bytes;
}

int acvae_diverse_beam_search(const acvae_dims* d, const acvae_weights* w, const float* audio_embeds, const int32_t* mem_lens,
                              const float* eps_g, int32_t beam_size, int32_t group_size, float diversity_lambda,
                              float temperature, int32_t group_nbest, int32_t start_idx, int32_t end_idx, int64_t* seqs,
                              void* workspace, size_t workspace_bytes, void* stream) {
  ACVAE_TRY(check_dims(d));
  ACVAE_REQUIRE(w && audio_embeds && mem_lens && eps_g && seqs && workspace, "NULL pointer");
  ACVAE_REQUIRE(group_size >= 1 && beam_size >= group_size && beam_size <= kDbsMaxBeam, "need 1 <= group_size <= beam_size <= 32");
  ACVAE_REQUIRE(temperature > 0.0f, "temperature must be positive");
  ACVAE_REQUIRE(d->mem_rep == 1, "diverse beam search takes one row per clip (mem_rep == 1)");
  ACVAE_REQUIRE(workspace_bytes >= carve_dbs_ws(*d, group_size, beam_size / group_size, nullptr).bytes, "workspace too small");
  return diverse_beam_search(*d, *w, audio_embeds, mem_lens, eps_g, beam_size, group_size, diversity_lambda, temperature,
                             group_nbest != 0, start_idx, end_idx, seqs, workspace, (cudaStream_t)stream);
}

int acvae_encoder_handoff_fwd(int32_t N, int32_t C, int32_t Te, int32_t F, const float* fmap, float* audio_embeds, float* pooled,
                              void* stream) {
  ACVAE_REQUIRE(N > 0 && C > 0 && Te > 0 && F > 0 && fmap && audio_embeds, "bad argument");
  return encoder_handoff_fwd(N, C, Te, F, fmap, audio_embeds, pooled, (cudaStream_t)stream);
}

int acvae_encoder_handoff_bwd(int32_t N, int32_t C, int32_t Te, int32_t F, const float* d_audio_embeds, float* d_fmap, void* stream) {
  ACVAE_REQUIRE(N > 0 && C > 0 && Te > 0 && F > 0 && d_audio_embeds && d_fmap, "bad argument");
  return encoder_handoff_bwd(N, C, Te, F, d_audio_embeds, d_fmap, (cudaStream_t)stream);
}

// profiling only (profiles/chain_trace.py): device buffer of [T][16] int64 that thread 0 of CTA 0 of the decoder
// forward chain fills with clock64 stamps; NULL switches it off
int acvae_debug_set_chain_trace(void* device_buffer) {
  chain_trace_ptr() = static_cast<long long*>(device_buffer);
  return 0;
}

// Arithmetic of the batched contractions, process-wide: 0 = fp32-grade (3xTF32, default), 1 = single-pass TF32
// (reduced precision: products with 10-bit mantissas, fp32 accumulation -- the "bf16" tolerance class of BASELINE.json).
// The recurrent chains, attention and all pointwise arithmetic stay fp32 in both modes.
int acvae_set_precision(int32_t mode) {
  ACVAE_REQUIRE(mode == 0 || mode == 1, "precision mode must be 0 (fp32-grade) or 1 (single-pass tf32)");
  tc_precision_mode() = mode;
  return 0;
}
int acvae_get_precision(void) { return tc_precision_mode(); }

int acvae_diversity_stats(int32_t clips, int32_t K, int32_t L, int32_t V, const int64_t* seqs, int32_t start_idx, int32_t end_idx,
                          double* div1, double* div2, int32_t* vocab_flags, void* stream) {
  ACVAE_REQUIRE(clips > 0 && K > 0 && L > 0 && V > 0 && seqs && div1 && div2, "bad argument");
  const size_t smem = sizeof(int) * 2 * (size_t)K * L;
  ACVAE_REQUIRE(smem <= 48 * 1024, "K * L too large for the per-clip shared-memory table (<= 6144 tokens)");
  ACVAE_LAUNCH(diversity_stats_kernel, clips, 256, smem, (cudaStream_t)stream, K, L, V, start_idx, end_idx, (const long long*)seqs,
               div1, div2, vocab_flags);
  return 0;
}

int acvae_mbleu_stats(int32_t clips, int32_t K, int32_t L, const int64_t* seqs, int32_t start_idx, int32_t end_idx, int32_t* stats,
                      void* stream) {
  ACVAE_REQUIRE(clips > 0 && K >= 2 && L > 0 && seqs && stats, "bad argument (mBLEU needs K >= 2 captions per clip)");
  const size_t smem = sizeof(int) * ((size_t)K * L + K + 1 + (size_t)K * 4);
  ACVAE_REQUIRE(smem <= 48 * 1024, "K * L too large for the per-clip shared-memory table");
  ACVAE_LAUNCH(mbleu_stats_kernel, clips, 256, smem, (cudaStream_t)stream, K, L, start_idx, end_idx, (const long long*)seqs, stats);
  return 0;
}

// Optional: a CUDA event the caller records after the host-to-device copy of the next call's audio embeddings (on any
// stream).  Every entry point that reads audio_embeds (acvae_train_fwd in both schedules, acvae_memory_prepare,
// acvae_decode_sample, acvae_beam_search, acvae_diverse_beam_search) waits for it right before its first read -- the
// hoisted training schedule only after the posterior chain, so the copy overlaps it; NULL (default) restores plain
// stream order.  The event must outlive every captured graph.
// Optional: an event every acvae_train_bwd records once all decoder.* weight gradients are final (see streams.cuh);
// NULL (default) switches it off.  The step-by-step schedule records it at its end.
int acvae_set_bucket_event(void* cuda_event) {
  bucket_event() = static_cast<cudaEvent_t>(cuda_event);
  return 0;
}

int acvae_set_input_event(void* cuda_event) {
  input_ready_event() = static_cast<cudaEvent_t>(cuda_event);
  return 0;
}

}  // extern "C"
