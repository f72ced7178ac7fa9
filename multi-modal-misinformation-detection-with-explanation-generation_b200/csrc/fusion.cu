// K5: fusion judge -- Linear(5,64)+ReLU (+Dropout=identity in eval) + Linear(64,32)+ReLU +
// Linear(32,2) + softmax + verdict, one warp per sample, a single launch.
// Replaces misinfo_forensics.py:83-90,106-108 (fusion_layer / forward_fusion), :587-608
// (fusion_verdict) and, with a modality mask, the fallback rule of :884-899.
//
// 28 B in / 12 B out and 4 864 flop per sample: neither roof is reachable; the point is
// one launch instead of ~8 and no .item() syncs.  The weights (10 120 B) live in device memory
// TRANSPOSED (mmf_fusion_load does it once on the host), so that lane j reads column j of every
// layer with coalesced 128-byte loads straight into REGISTERS (81 per lane) -- no shared-memory
// staging, no block barrier, and no shared-memory read per multiply in the sample loop; layer-2
// inputs travel by warp shuffle.
#include "common.cuh"

namespace mmf {

// device layout of the parameters (floats): lane j's column of every layer is contiguous across lanes
constexpr int FP_W0T = 0;        // [5][64]   w0t[i][j] = W0[j][i]
constexpr int FP_B0 = 320;       // [64]
constexpr int FP_W3T = 384;      // [64][32]  w3t[i][j] = W3[j][i]
constexpr int FP_B3 = 2432;      // [32]
constexpr int FP_W5 = 2464;      // [2][32]
constexpr int FP_B5 = 2528;      // [2]

// ASSEMBLE: the fusion inputs are built here as well (analyze()'s score assembly, misinfo_forensics.py:794-809):
// x[s] = [ai, misinfo, deepfake, clip_sim, vault_disc] from head (n,3), sim (n), disc (n) with the scores of a skipped
// modality zeroed (sim / disc masked in place, x written out) -- one launch instead of two for the batched path.
template <bool ASSEMBLE>
__global__ void __launch_bounds__(256) fusion_judge_kernel(const float* __restrict__ params,
                                                           const float* __restrict__ x,
                                                           const unsigned char* __restrict__ modality,
                                                           long long n, float* __restrict__ out_probs,
                                                           int* __restrict__ out_verdict,
                                                           float* __restrict__ out_conf,
                                                           const float* __restrict__ head, float* __restrict__ sim,
                                                           float* __restrict__ disc, float* __restrict__ x_out) {
  const int lane = threadIdx.x & 31;
  // this lane's columns of the three layers (independent loads, all in flight at once)
  float w0a[5], w0b[5], w3[64];
#pragma unroll
  for (int i = 0; i < 5; ++i) { w0a[i] = __ldg(params + FP_W0T + i * 64 + lane); w0b[i] = __ldg(params + FP_W0T + i * 64 + 32 + lane); }
#pragma unroll
  for (int i = 0; i < 64; ++i) w3[i] = __ldg(params + FP_W3T + i * 32 + lane);
  const float b0a = __ldg(params + FP_B0 + lane), b0b = __ldg(params + FP_B0 + 32 + lane), b3 = __ldg(params + FP_B3 + lane);
  const float w5a = __ldg(params + FP_W5 + lane), w5b = __ldg(params + FP_W5 + 32 + lane);
  const float b5a = __ldg(params + FP_B5), b5b = __ldg(params + FP_B5 + 1);
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; s < n; s += warps) {
    const int mod = modality ? modality[s] : 3;
    float xl;
    if (ASSEMBLE) {
      const bool has_text = mod & 1, has_vis = mod & 2;
      xl = 0.f;
      if (lane < 2) xl = has_text ? head[s * 3 + lane] : 0.f;
      else if (lane == 2) xl = has_vis ? head[s * 3 + 2] : 0.f;
      else if (lane == 3) { xl = (has_text && has_vis) ? sim[s] : 0.f; sim[s] = xl; }   // analyze() skips the steps whose
      else if (lane == 4) { xl = has_vis ? disc[s] : 0.f; disc[s] = xl; }               // modality is missing -> 0.0
      if (lane < 5) x_out[s * 5 + lane] = xl;
    } else {
      xl = (lane < 5) ? x[s * 5 + lane] : 0.f;
    }
    float xi[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) xi[i] = __shfl_sync(FULL, xl, i);
    float real, fake;
    if (mod == 3) {                                   // warp-uniform: one sample per warp
      float h1a = b0a, h1b = b0b;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        h1a = fmaf(w0a[i], xi[i], h1a);
        h1b = fmaf(w0b[i], xi[i], h1b);
      }
      h1a = fmaxf(h1a, 0.f);
      h1b = fmaxf(h1b, 0.f);
      float h2 = b3;
#pragma unroll
      for (int i = 0; i < 32; ++i) h2 = fmaf(w3[i], __shfl_sync(FULL, h1a, i), h2);
#pragma unroll
      for (int i = 0; i < 32; ++i) h2 = fmaf(w3[32 + i], __shfl_sync(FULL, h1b, i), h2);
      h2 = fmaxf(h2, 0.f);
      const float l0 = warp_sum(w5a * h2) + b5a;
      const float l1 = warp_sum(w5b * h2) + b5b;
      const float m = fmaxf(l0, l1);                  // torch.softmax(dim=1), :598
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      const float inv = 1.0f / (e0 + e1);
      real = e0 * inv;
      fake = e1 * inv;
      if (l0 != l0 || l1 != l1) real = fake = l0 + l1;   // NaN in -> NaN out, like torch
    } else {                                          // misinfo_forensics.py:884-899
      fake = (mod == 1) ? xi[1] : (mod == 2) ? fmaxf(xi[2], xi[4]) : 0.5f;
      fake = fmaxf(0.0f, fminf(1.0f, fake));
      real = 1.0f - fake;
    }
    if (lane == 0) {
      const int verdict = fake > 0.5f ? 1 : 0;        // :605
      out_probs[s * 2 + 0] = real;
      out_probs[s * 2 + 1] = fake;
      if (out_verdict) out_verdict[s] = verdict;
      if (out_conf) out_conf[s] = verdict ? fake : real;
    }
  }
}

}  // namespace mmf

static int fusion_launch(mmf_handle* h, const float* x, const uint8_t* modality, int64_t n, float* out_probs,
                         int32_t* out_verdict, float* out_conf, cudaStream_t st, const char* who) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n < 0 || (n > 0 && (!x || !out_probs))) return mmf_set_error(h, MMF_ERR_BAD_ARG, "%s: bad argument", who);
  if (!h->fusion_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "%s: fusion weights not loaded", who);
  if (n == 0) return MMF_OK;
  const long long want = (n + 7) / 8;
  const int grid = (int)(want < (long long)h->sm_count * 4 ? want : (long long)h->sm_count * 4);
  mmf::fusion_judge_kernel<false><<<grid, 256, 0, st>>>(h->fusion_params, x, modality, n, out_probs, out_verdict, out_conf,
                                                        nullptr, nullptr, nullptr, nullptr);
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

// score assembly + verdict in one launch (the batched host entry, api.cu); modality_in may be null (= both present)
int mmf_assemble_verdict(mmf_handle* h, const float* head, const uint8_t* modality_in, int64_t n, float* sim, float* disc,
                         float* x, float* out_probs, int32_t* out_verdict, float* out_conf, cudaStream_t st) {
  if (!h->fusion_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "assemble_verdict: fusion weights not loaded");
  if (n <= 0) return MMF_OK;
  const long long want = (n + 7) / 8;
  const int grid = (int)(want < (long long)h->sm_count * 4 ? want : (long long)h->sm_count * 4);
  mmf::fusion_judge_kernel<true><<<grid, 256, 0, st>>>(h->fusion_params, nullptr, modality_in, n, out_probs, out_verdict,
                                                       out_conf, head, sim, disc, x);
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

extern "C" int mmf_fusion_load(mmf_handle* h, const float* params_host) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (!params_host) return mmf_set_error(h, MMF_ERR_BAD_ARG, "fusion_load: null params");
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  if (!h->fusion_params) MMF_CUDA_OK(h, cudaMalloc(&h->fusion_params, MMF_FUSION_PARAMS * sizeof(float)));
  // params_host: [W0 (64,5) | b0 64 | W3 (32,64) | b3 32 | W5 (2,32) | b5 2]  (nn.Linear: weight (out,in)); the device
  // copy holds W0 and W3 transposed (see the FP_* offsets).  Synchronous on purpose: the trainer mutates the weights
  // in place and calls this again.
  float t[MMF_FUSION_PARAMS];
  memcpy(t, params_host, sizeof t);
  for (int j = 0; j < 64; ++j)
    for (int i = 0; i < 5; ++i) t[mmf::FP_W0T + i * 64 + j] = params_host[j * 5 + i];
  for (int j = 0; j < 32; ++j)
    for (int i = 0; i < 64; ++i) t[mmf::FP_W3T + i * 32 + j] = params_host[384 + j * 64 + i];
  MMF_CUDA_OK(h, cudaMemcpy(h->fusion_params, t, sizeof t, cudaMemcpyHostToDevice));
  h->fusion_loaded = true;
  return MMF_OK;
}

extern "C" int mmf_fusion_forward(mmf_handle* h, const float* x, int64_t n, float* out_probs, int32_t* out_verdict,
                                  float* out_confidence, mmf_stream_t stream) {
  return fusion_launch(h, x, nullptr, n, out_probs, out_verdict, out_confidence, (cudaStream_t)stream, "fusion_forward");
}

extern "C" int mmf_verdict_batch(mmf_handle* h, const float* scores, const uint8_t* modality, int64_t n,
                                 float* out_probs, int32_t* out_verdict, float* out_confidence,
                                 mmf_stream_t stream) {
  if (h && n > 0 && !modality) return mmf_set_error(h, MMF_ERR_BAD_ARG, "verdict_batch: null modality");
  return fusion_launch(h, scores, modality, n, out_probs, out_verdict, out_confidence, (cudaStream_t)stream, "verdict_batch");
}
