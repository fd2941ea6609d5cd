// Auxiliary streams/events for intra-call concurrency.  The recurrent chains (posterior directions,
// prior, decoder) are latency-bound and occupy a few dozen SMs each; running them -- and the batched
// GEMMs that do not depend on them -- on forked streams lets them overlap.  Fork/join is expressed with
// events, so it is legal under CUDA-graph stream capture (the side streams join the capture).
// Created once per process on first use (the device current at that time).
#pragma once
#include "common.cuh"

namespace acvae {

constexpr int kAuxStreams = 13;     // 0,1: posterior directions; 2: prior; 3: memory backward; 4..11: weight-gradient fan; 12: main chain
constexpr int kAuxMain = 12;
constexpr int kAuxFan0 = 4, kAuxFanN = 8;
constexpr int kAuxEvents = 128;

struct Aux {
  cudaStream_t s[kAuxStreams];
  cudaEvent_t e[kAuxEvents];
  int next_event = 0;
  bool ok = false;
  cudaEvent_t ev() { cudaEvent_t r = e[next_event]; next_event = (next_event + 1) % kAuxEvents; return r; }
};

inline Aux* aux() {
  static Aux a;
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

// `to` waits for everything enqueued so far on `from`
inline int stream_dep(cudaStream_t from, cudaStream_t to, Aux* a) {
  cudaEvent_t e = a->ev();
  ACVAE_CHECK(cudaEventRecord(e, from));
  ACVAE_CHECK(cudaStreamWaitEvent(to, e, 0));
  return 0;
}

}  // namespace acvae
