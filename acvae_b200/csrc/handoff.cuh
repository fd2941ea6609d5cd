// Encoder hand-off: the tail of the audio encoder fused with the entry of the latent decoding step.
//
// Cnn10.forward (models/encoder.py:691-700) ends with
//     x = torch.mean(x, dim=3)                 # [N, C, Te, F] -> [N, C, Te]      (read 4x, write 1x)
//     ...
//     x = x.transpose(1, 2).contiguous()       # -> audio_embeds [N, Te, C]       (read 1x, write 1x)
// and Hybrid_VAEModel.forward then reads audio_embeds again for `ln` (vae_model.py:743-744).  Here the frequency mean and
// the transpose are ONE pass over the convolution's feature map that writes the frame memory directly in the row-major
// [N*Te, C] layout the step's tensor-core GEMMs (ln, the two attention memory projections, the decoder's Mg) take as
// their K-major A operand through TMA -- no [N, C, Te] intermediate, no transposed copy.  The backward is the matching
// single pass d fmap[n,c,j,f] = d audio_embeds[n,j,c] / F.
// HBM-bound: N*C*Te*(F + 1) floats per call (16 MB + 4 MB at N = 32, C = 512, Te = 62, F = 4).
#pragma once
#include "common.cuh"

namespace acvae {

constexpr int kHoC = 32;     // channels per CTA (one 128-byte output segment per frame)

// grid (C / 32, N), 256 threads: warp w handles channels c0 + w, c0 + w + 8, ...; a lane handles frames lane, lane + 32, ...
__global__ void __launch_bounds__(256) encoder_handoff_fwd_kernel(int C, int Te, int F, const float* __restrict__ fmap,
                                                                  float* __restrict__ out, float* __restrict__ pooled) {
  extern __shared__ float tile[];                              // [Te][kHoC + 1]
  const int n = blockIdx.y, c0 = blockIdx.x * kHoC;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float invF = 1.0f / (float)F;
  for (int cc = w; cc < kHoC; cc += 8) {
    const int c = c0 + cc;
    float mx = -INFINITY, sm = 0.0f;
    if (c < C) {
      const float* src = fmap + ((long long)n * C + c) * Te * F;
      for (int j = lane; j < Te; j += 32) {
        float s = 0.0f;
        if (F == 4 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
          s = (v.x + v.y) + (v.z + v.w);
        } else {
          for (int f = 0; f < F; ++f) s += __ldg(src + (long long)j * F + f);
        }
        s *= invF;
        tile[j * (kHoC + 1) + cc] = s;
        mx = fmaxf(mx, s); sm += s;
      }
    }
    if (pooled) {                                              // max over frames + mean over frames (encoder.py:693-695)
      mx = warp_max(mx); sm = warp_sum(sm);
      if (lane == 0 && c < C) pooled[(long long)n * C + c] = mx + sm / (float)Te;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Te * kHoC; i += blockDim.x) {
    const int j = i / kHoC, cc = i % kHoC;
    if (c0 + cc < C) out[((long long)n * Te + j) * C + c0 + cc] = tile[j * (kHoC + 1) + cc];
  }
}

// d fmap[n,c,j,f] = d out[n,j,c] / F
__global__ void __launch_bounds__(256) encoder_handoff_bwd_kernel(int C, int Te, int F, const float* __restrict__ dout,
                                                                  float* __restrict__ dfmap) {
  extern __shared__ float tile[];                              // [Te][kHoC + 1]
  const int n = blockIdx.y, c0 = blockIdx.x * kHoC;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float invF = 1.0f / (float)F;
  for (int i = threadIdx.x; i < Te * kHoC; i += blockDim.x) {
    const int j = i / kHoC, cc = i % kHoC;
    tile[j * (kHoC + 1) + cc] = c0 + cc < C ? __ldg(dout + ((long long)n * Te + j) * C + c0 + cc) * invF : 0.0f;
  }
  __syncthreads();
  for (int cc = w; cc < kHoC; cc += 8) {
    const int c = c0 + cc;
    if (c >= C) continue;
    float* dst = dfmap + ((long long)n * C + c) * Te * F;
    for (int i = lane; i < Te * F; i += 32) dst[i] = tile[(i / F) * (kHoC + 1) + cc];
  }
}

inline int encoder_handoff_fwd(int N, int C, int Te, int F, const float* fmap, float* out, float* pooled, cudaStream_t st) {
  const size_t smem = (size_t)Te * (kHoC + 1) * sizeof(float);
  ACVAE_REQUIRE(smem <= 48 * 1024, "encoder hand-off: too many frames for one shared-memory tile (Te <= 372)");
  ACVAE_LAUNCH(encoder_handoff_fwd_kernel, dim3((C + kHoC - 1) / kHoC, N), 256, smem, st, C, Te, F, fmap, out, pooled);
  return 0;
}
inline int encoder_handoff_bwd(int N, int C, int Te, int F, const float* dout, float* dfmap, cudaStream_t st) {
  const size_t smem = (size_t)Te * (kHoC + 1) * sizeof(float);
  ACVAE_REQUIRE(smem <= 48 * 1024, "encoder hand-off: too many frames for one shared-memory tile (Te <= 372)");
  ACVAE_LAUNCH(encoder_handoff_bwd_kernel, dim3((C + kHoC - 1) / kHoC, N), 256, smem, st, C, Te, F, dout, dfmap);
  return 0;
}

}  // namespace acvae
