// Auxiliary streams/events for intra-call concurrency.  The recurrent chains (posterior directions,
// prior, decoder) are latency-bound and occupy a few dozen SMs each; running them -- and the batched
// GEMMs that do not depend on them -- on forked streams lets them overlap.  Fork/join is expressed with
// events, so it is legal under CUDA-graph stream capture (the side streams join the capture).
// Created once per process on first use (the device current at that time).
#pragma once
#include <mutex>

#include "common.cuh"

namespace acvae {

// 0,1: posterior directions; 2: prior; 3: memory backward; 4..11: weight-gradient fan of the decoder / posterior;
// 12..19: weight-gradient fan of the prior (its own streams: stream order is FIFO, so work that becomes ready at different
// times must not share a stream); 20: main chain
constexpr int kAuxStreams = 21;
constexpr int kAuxMain = 20;
constexpr int kAuxFan0 = 4, kAuxFanN = 16, kAuxPriorFan0 = 12;
constexpr int kAuxEvents = 128;

struct Aux {
  cudaStream_t s[kAuxStreams];
  cudaEvent_t e[kAuxEvents];
  int next_event = 0;
  bool ok = false;
  std::mutex mu;      // the event ring is shared by every host thread that drives this device
  cudaEvent_t ev() { std::lock_guard<std::mutex> g(mu); cudaEvent_t r = e[next_event]; next_event = (next_event + 1) % kAuxEvents; return r; }
};

// One set of side streams / events per device, created on first use with that device current (streams and events belong
// to the context they were created in).
inline Aux* aux() {
  static Aux all[kMaxDevices];
  static std::mutex init_mu;
  Aux& a = all[current_device()];
  std::lock_guard<std::mutex> g(init_mu);
  if (!a.ok) {
    // Priorities: the recurrent chains and the few batched kernels between them are the critical path (highest);
    // the memory backward is next; the weight-gradient fan only fills idle SMs (lowest).  The caller's stream has
    // the default (lowest) priority, so the critical work runs on s[kAuxMain], forked from / joined to it.
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { lo = 0; hi = 0; }
    for (int i = 0; i < kAuxStreams; ++i) {
      const int prio = (i >= kAuxFan0 && i < kAuxFan0 + kAuxFanN) ? lo : (i == 3 ? (hi + lo) / 2 : hi);
      if (cudaStreamCreateWithPriority(&a.s[i], cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
    }
    for (int i = 0; i < kAuxEvents; ++i)
      if (cudaEventCreateWithFlags(&a.e[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    a.ok = true;
  }
  return &a;
}

// Optional "inputs ready" event of the caller (acvae_set_input_event): recorded by the caller on whatever stream carries
// the host-to-device copy of the step's audio embeddings.  The forward waits for it right before the first kernel that
// reads them, with cudaEventWaitExternal -- legal inside stream capture, where it becomes an external event-wait node --
// so the copy runs under the posterior chain (which does not read the audio) instead of in front of the step.
inline cudaEvent_t& input_ready_event() { static cudaEvent_t e = nullptr; return e; }

// Optional "decoder gradients final" event (acvae_set_bucket_event): recorded by every acvae_train_bwd at the point where all
// decoder.* weight gradients (and, before it in the caller's stream, the classifier's) are final -- ~0.2 ms before the end of
// the backward.  A data-parallel caller starts the all-reduce of that bucket behind it, under the rest of the backward.
inline cudaEvent_t& bucket_event() { static cudaEvent_t e = nullptr; return e; }

// Deferred classifier gradients (acvae_defer_classifier_grads): the loss backward produces d hidden on the caller's stream and
// d W_cls / d b_cls -- which nothing in the step reads before the optimizer -- on fan stream 7; every later entry point that
// could consume them (acvae_train_bwd, the optimizer entries) joins that stream back into the caller's.
inline bool& cls_defer_flag() { static bool f = false; return f; }
inline bool* cls_defer_pending() { static bool p[kMaxDevices] = {false}; return &p[current_device()]; }
inline int join_deferred_cls_grads(cudaStream_t user);

// Make `st` wait for the caller's "inputs ready" event, if one is set: called by EVERY entry point right before its first
// kernel that reads audio_embeds (both training schedules, the sampling / beam / diverse-beam loops, acvae_memory_prepare).
inline int wait_input_event(cudaStream_t st) {
  if (!input_ready_event()) return 0;
  // under stream capture the wait must be an EXTERNAL event-wait node (the event is recorded outside the graph, before
  // each replay); outside capture that flag is invalid and a plain wait does the same
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  ACVAE_CHECK(cudaStreamIsCapturing(st, &cs));
  ACVAE_CHECK(cudaStreamWaitEvent(st, input_ready_event(), cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : 0));
  return 0;
}

// `to` waits for everything enqueued so far on `from`
inline int stream_dep(cudaStream_t from, cudaStream_t to, Aux* a) {
  cudaEvent_t e = a->ev();
  ACVAE_CHECK(cudaEventRecord(e, from));
  ACVAE_CHECK(cudaStreamWaitEvent(to, e, 0));
  return 0;
}

inline int join_deferred_cls_grads(cudaStream_t user) {
  bool* pending = cls_defer_pending();
  if (!*pending) return 0;
  Aux* ax = aux();
  if (!ax) return set_error("join_deferred_cls_grads", "no side streams");
  ACVAE_TRY(stream_dep(ax->s[kAuxFan0 + 7], user, ax));
  *pending = false;
  return 0;
}

}  // namespace acvae
