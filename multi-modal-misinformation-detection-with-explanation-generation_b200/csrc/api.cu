// C-ABI glue: handle lifetime, error strings, scratch, search dispatch, host-buffer entry.
#include "common.cuh"

#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>

int mmf_stream_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                      float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc, cudaStream_t st);
int mmf_mma_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                   float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc, cudaStream_t st);
int mmf_mma_supported(const mmf_handle* h, int64_t n_queries, int top_k);
void mmf_mma_destroy(mmf_handle* h);
extern "C" int mmf_exchange_detach(mmf_handle* h);
extern "C" int mmf_shard_finalize(mmf_handle* h);
int mmf_fill_empty(mmf_handle* h, int64_t n_queries, int top_k, float* out_scores, int64_t* out_rows,
                   uint64_t* out_packed, float* out_disc, cudaStream_t st);

int mmf_set_error(mmf_handle* h, int status, const char* fmt, ...) {
  if (h) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    h->last_error = buf;
  }
  return status;
}

int mmf_ensure_scratch(mmf_handle* h, size_t bytes, cudaStream_t stream) {
  const int a = h->scratch_sel;
  if (bytes <= h->scratch_arena_bytes[a]) return MMF_OK;
  // growth is rare (first call / larger batch).  It frees memory that earlier calls may still be using, so it
  // waits for the device (cudaFree does anyway) -- which a stream capture does not allow: size the scratch by
  // running the step once before capturing it.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
    return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "scratch must grow to %zu bytes during a stream capture: run the step once before capturing", bytes);
  cudaGetLastError();
  MMF_CUDA_OK(h, cudaDeviceSynchronize());
  if (h->scratch_arena[a]) MMF_CUDA_OK(h, cudaFree(h->scratch_arena[a]));
  h->scratch_arena[a] = nullptr;
  h->scratch_arena_bytes[a] = 0;
  const size_t want = std::max(bytes + bytes / 4, (size_t)1 << 22);
  if (cudaMalloc(&h->scratch_arena[a], want) != cudaSuccess) {
    cudaGetLastError();
    return mmf_set_error(h, MMF_ERR_NOMEM, "cannot allocate %zu bytes of scratch", want);
  }
  h->scratch_arena_bytes[a] = want;
  MMF_CUDA_OK(h, cudaMemsetAsync(h->scratch_arena[a], 0, 65536, stream));   // arrival counters start at 0
  return MMF_OK;
}

// device I/O + pinned staging of one host-entry slot; growth waits for the slot's previous user (it is not busy:
// collected) and for nothing else
static int ensure_slot(mmf_handle* h, mmf_host_slot& sl, size_t io_bytes, size_t pinned_bytes) {
  if (!sl.ev_in) {
    MMF_CUDA_OK(h, cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    MMF_CUDA_OK(h, cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    MMF_CUDA_OK(h, cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
  }
  if (pinned_bytes > sl.pinned_bytes) {
    if (sl.pinned) MMF_CUDA_OK(h, cudaFreeHost(sl.pinned));
    sl.pinned = nullptr;
    sl.pinned_bytes = 0;
    const size_t want = std::max(pinned_bytes + pinned_bytes / 4, (size_t)1 << 20);
    MMF_CUDA_OK(h, cudaMallocHost(&sl.pinned, want));
    sl.pinned_bytes = want;
  }
  if (io_bytes > sl.io_bytes) {
    if (sl.io) MMF_CUDA_OK(h, cudaFree(sl.io));
    sl.io = nullptr;
    sl.io_bytes = 0;
    const size_t want = io_bytes + io_bytes / 4;
    if (cudaMalloc(&sl.io, want) != cudaSuccess) {
      cudaGetLastError();
      return mmf_set_error(h, MMF_ERR_NOMEM, "cannot allocate %zu bytes of I/O buffer", want);
    }
    sl.io_bytes = want;
  }
  return MMF_OK;
}

extern "C" const char* mmf_version(void) { return "mmf_b200 0.1.0 (sm_100a)"; }
extern "C" int mmf_arch(void) { return 100; }

extern "C" const char* mmf_status_string(int status) {
  switch (status) {
    case MMF_OK: return "ok";
    case MMF_ERR_BAD_ARG: return "bad argument";
    case MMF_ERR_CUDA: return "CUDA error";
    case MMF_ERR_NOT_LOADED: return "not loaded";
    case MMF_ERR_NO_DEVICE: return "no CUDA device";
    case MMF_ERR_UNSUPPORTED: return "unsupported";
    case MMF_ERR_NOMEM: return "out of device memory";
    case MMF_ERR_NCCL: return "NCCL error";
  }
  return "unknown status";
}

extern "C" int mmf_create(int device_ordinal, mmf_handle** out) {
  if (!out) return MMF_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return MMF_ERR_NO_DEVICE; }
  if (device_ordinal < 0 || device_ordinal >= n) return MMF_ERR_BAD_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return MMF_ERR_CUDA;
  if (prop.major != 10) return MMF_ERR_UNSUPPORTED;      // sm_100a code only; no other path exists
  mmf_handle* h = new (std::nothrow) mmf_handle();
  if (!h) return MMF_ERR_NOMEM;
  h->device = device_ordinal;
  h->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device_ordinal) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return MMF_ERR_CUDA;
  }
  // the only place the environment is read: MMF_OPT_SCREEN, MMF_OPT_FUSED_PUSH, MMF_OPT_DEBUG, MMF_OPT_FORCE_CG,
  // MMF_OPT_FLAT_SCHEDULE (same names as mmf_set_option, upper case)
  for (const char* name : {"screen", "fused_push", "debug", "force_cg", "flat_schedule"}) {
    char env[64] = "MMF_OPT_";
    size_t n = strlen(env);
    for (const char* c = name; *c && n + 1 < sizeof env; ++c) env[n++] = (char)toupper((unsigned char)*c);
    env[n] = 0;
    const char* v = getenv(env);
    if (v && *v) mmf_set_option(h, name, atoi(v));
  }
  *out = h;
  return MMF_OK;
}

extern "C" int mmf_set_option(mmf_handle* h, const char* name, int value) {
  if (!h || !name) return MMF_ERR_BAD_ARG;
  mmf_options& o = h->opt;
  if (!strcmp(name, "screen")) o.screen = value != 0;
  else if (!strcmp(name, "fused_push")) o.fused_push = value != 0;
  else if (!strcmp(name, "debug")) o.debug = value;
  else if (!strcmp(name, "force_cg")) o.force_cg = (value == 1 || value == 2) ? value : 0;
  else if (!strcmp(name, "flat_schedule")) o.flat_schedule = value != 0;
  else return mmf_set_error(h, MMF_ERR_BAD_ARG, "set_option: unknown option '%s'", name);
  return MMF_OK;
}

extern "C" int mmf_get_option(const mmf_handle* h, const char* name, int* value) {
  if (!h || !name || !value) return MMF_ERR_BAD_ARG;
  const mmf_options& o = h->opt;
  if (!strcmp(name, "screen")) *value = o.screen;
  else if (!strcmp(name, "fused_push")) *value = o.fused_push;
  else if (!strcmp(name, "debug")) *value = o.debug;
  else if (!strcmp(name, "force_cg")) *value = o.force_cg;
  else if (!strcmp(name, "flat_schedule")) *value = o.flat_schedule;
  else return MMF_ERR_BAD_ARG;
  return MMF_OK;
}

extern "C" int mmf_destroy(mmf_handle* h) {
  if (!h) return MMF_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  mmf_shard_finalize(h);
  mmf_exchange_detach(h);
  mmf_mma_destroy(h);
  if (h->vault) cudaFree(h->vault);
  if (h->fusion_params) cudaFree(h->fusion_params);
  if (h->vault_nan_rows_dev) cudaFree(h->vault_nan_rows_dev);
  for (int a = 0; a < 2; ++a)
    if (h->scratch_arena[a]) cudaFree(h->scratch_arena[a]);
  for (mmf_host_slot& sl : h->slot) {
    if (sl.pinned) cudaFreeHost(sl.pinned);
    if (sl.io) cudaFree(sl.io);
    for (cudaEvent_t e : {sl.ev_in, sl.ev_done, sl.ev_out})
      if (e) cudaEventDestroy(e);
  }
  for (cudaStream_t st : {h->own_stream, h->h2d_stream, h->d2h_stream})
    if (st) cudaStreamDestroy(st);
  delete h;
  return MMF_OK;
}

extern "C" const char* mmf_last_error(const mmf_handle* h) { return h ? h->last_error.c_str() : "null handle"; }
extern "C" int64_t mmf_launch_count(const mmf_handle* h) { return h ? h->launches : 0; }
extern "C" int64_t mmf_collective_count(const mmf_handle* h) { return h ? h->collectives : 0; }

static int search_dispatch(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                           int algo, float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc,
                           cudaStream_t st, const char* who) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_queries < 0 || top_k <= 0 || (n_queries > 0 && !queries))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "%s: bad argument (n_queries=%lld top_k=%d)", who, (long long)n_queries, top_k);
  if (top_k > MMF_MAX_TOP_K)
    return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "%s: top_k %d > %d", who, top_k, MMF_MAX_TOP_K);
  if (!h->vault_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "%s: no vault loaded", who);
  if (algo != MMF_ALGO_AUTO && algo != MMF_ALGO_STREAM && algo != MMF_ALGO_MMA)
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "%s: unknown algo %d", who, algo);
  if (n_queries == 0) return MMF_OK;
  if (h->vault_rows == 0) return mmf_fill_empty(h, n_queries, top_k, out_scores, out_rows, out_packed, out_disc, st);
  const int64_t chunk = 65536;
  for (int64_t q0 = 0; q0 < n_queries; q0 += chunk) {
    const int64_t nq = std::min(chunk, n_queries - q0);
    bool use_mma;
    if (algo == MMF_ALGO_MMA) {
      if (!mmf_mma_supported(h, nq, top_k))
        return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "%s: tcgen05 path unavailable for this shape", who);
      use_mma = true;
    } else if (algo == MMF_ALGO_STREAM) {
      use_mma = false;
    } else {
      use_mma = nq >= 16 && mmf_mma_supported(h, nq, top_k);   // below that the search is HBM-bound on CUDA cores
    }
    float* os = out_scores ? out_scores + q0 * top_k : nullptr;
    int64_t* orow = out_rows ? out_rows + q0 * top_k : nullptr;
    uint64_t* op = out_packed ? out_packed + q0 * top_k : nullptr;
    float* od = out_disc ? out_disc + q0 : nullptr;
    const int rc = use_mma ? mmf_mma_search(h, queries + q0 * MMF_DIM, nq, top_k, threshold, os, orow, op, od, st)
                           : mmf_stream_search(h, queries + q0 * MMF_DIM, nq, top_k, threshold, os, orow, op, od, st);
    if (rc != MMF_OK) return rc;
  }
  return MMF_OK;
}

// local search with packed output, for the peer-memory exchange (exchange.cu)
int mmf_search_dispatch_packed(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int algo,
                               uint64_t* out_packed, cudaStream_t st, const char* who) {
  return search_dispatch(h, queries, n_queries, top_k, 0.0, algo, nullptr, nullptr, out_packed, nullptr, st, who);
}

extern "C" int mmf_vault_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                                int algo, float* out_scores, int64_t* out_rows, float* out_discrepancy,
                                mmf_stream_t stream) {
  if (h && n_queries > 0 && (!out_scores || !out_rows))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search: null output");
  return search_dispatch(h, queries, n_queries, top_k, threshold, algo, out_scores, out_rows, nullptr,
                         out_discrepancy, (cudaStream_t)stream, "vault_search");
}

extern "C" int mmf_vault_search_candidates(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int algo,
                                           uint64_t* out_packed, mmf_stream_t stream) {
  if (h && n_queries > 0 && !out_packed)
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_candidates: null output");
  return search_dispatch(h, queries, n_queries, top_k, 0.0, algo, nullptr, nullptr, out_packed, nullptr,
                         (cudaStream_t)stream, "vault_search_candidates");
}

extern "C" int mmf_vault_search_host(mmf_handle* h, const float* queries_host, int64_t n_queries, int top_k,
                                     double threshold, int algo, float* out_scores_host, int64_t* out_rows_host,
                                     float* out_discrepancy_host) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_queries < 0 || top_k <= 0 || top_k > MMF_MAX_TOP_K ||
      (n_queries > 0 && (!queries_host || !out_scores_host || !out_rows_host)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_host: bad argument");
  if (!h->vault_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "vault_search_host: no vault loaded");
  if (n_queries == 0) return MMF_OK;
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  // pinned staging: [queries | scores | rows | disc]; device I/O buffers mirror it
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t bq = al((size_t)n_queries * MMF_DIM * 4), bs = al((size_t)n_queries * top_k * 4);
  const size_t br = al((size_t)n_queries * top_k * 8), bd = al((size_t)n_queries * 4);
  const size_t total = bq + bs + br + bd;
  int rc = ensure_pinned(h, total);
  if (rc != MMF_OK) return rc;
  if (total > h->io_bytes) {
    if (h->io) MMF_CUDA_OK(h, cudaFree(h->io));
    h->io = nullptr;
    h->io_bytes = 0;
    MMF_CUDA_OK(h, cudaMalloc(&h->io, total));
    h->io_bytes = total;
  }
  char* hp = (char*)h->pinned;
  char* dp = (char*)h->io;
  memcpy(hp, queries_host, (size_t)n_queries * MMF_DIM * 4);
  MMF_CUDA_OK(h, cudaMemcpyAsync(dp, hp, bq, cudaMemcpyHostToDevice, st));
  float* d_scores = (float*)(dp + bq);
  int64_t* d_rows = (int64_t*)(dp + bq + bs);
  float* d_disc = (float*)(dp + bq + bs + br);
  rc = search_dispatch(h, (const float*)dp, n_queries, top_k, threshold, algo, d_scores, d_rows, nullptr, d_disc, st,
                       "vault_search_host");
  if (rc != MMF_OK) return rc;
  MMF_CUDA_OK(h, cudaMemcpyAsync(hp + bq, dp + bq, bs + br + bd, cudaMemcpyDeviceToHost, st));
  MMF_CUDA_OK(h, cudaStreamSynchronize(st));
  memcpy(out_scores_host, hp + bq, (size_t)n_queries * top_k * 4);
  memcpy(out_rows_host, hp + bq + bs, (size_t)n_queries * top_k * 8);
  if (out_discrepancy_host) memcpy(out_discrepancy_host, hp + bq + bs + br, (size_t)n_queries * 4);
  return MMF_OK;
}

// ---- batched analyze downstream of the encoders, HOST buffers in and out ---------------------------------
// The whole hot path for a batch in one call: H2D of the embeddings, caption/image cosine (K1), vault search
// (K2 / K3), score assembly with the modality rules of misinfo_forensics.py:794-809, fusion judge / fallback
// verdict (K5), ONE D2H of all results, ONE stream synchronisation.  Same results as mmf_b200.score_batch
// (pipeline.py), which issues the same kernels from Python with one torch op and one .cpu() sync per tensor.
namespace mmf {
// x[i] = [ai, misinfo, deepfake, clip_sim, vault_disc] with the skipped modalities zeroed; sim / disc masked in place
__global__ void __launch_bounds__(256) assemble_scores_kernel(const float* __restrict__ head, const unsigned char* __restrict__ mod_in,
                                                              long long n, float* __restrict__ sim, float* __restrict__ disc,
                                                              float* __restrict__ x, unsigned char* __restrict__ mod_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int mod = mod_in ? mod_in[i] : 3;
  const bool has_text = mod & 1, has_vis = mod & 2;
  const float s = (has_text && has_vis) ? sim[i] : 0.f;      // analyze() skips the steps whose modality is missing
  const float d = has_vis ? disc[i] : 0.f;
  sim[i] = s;
  disc[i] = d;
  x[i * 5 + 0] = has_text ? head[i * 3 + 0] : 0.f;
  x[i * 5 + 1] = has_text ? head[i * 3 + 1] : 0.f;
  x[i * 5 + 2] = has_vis ? head[i * 3 + 2] : 0.f;
  x[i * 5 + 3] = s;
  x[i * 5 + 4] = d;
  mod_out[i] = (unsigned char)mod;
}
}  // namespace mmf

extern "C" int mmf_score_batch_host(mmf_handle* h, const float* text_host, const float* image_host, const float* head_host,
                                    const uint8_t* modality_host, int64_t n, int top_k, double threshold, int algo,
                                    float* out_clip_similarity, float* out_vault_discrepancy, float* out_vault_scores,
                                    int64_t* out_vault_rows, float* out_scores5, float* out_probs, int32_t* out_verdict,
                                    float* out_confidence) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n < 0 || top_k <= 0 || top_k > MMF_MAX_TOP_K ||
      (n > 0 && (!text_host || !image_host || !head_host || !out_probs)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_host: bad argument (n=%lld top_k=%d)", (long long)n, top_k);
  if (!h->fusion_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "score_batch_host: fusion weights not loaded");
  if (n == 0) return MMF_OK;
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  // device buffer: inputs, then ONE contiguous block of outputs (mirrored by the pinned staging buffer)
  const size_t b_emb = al((size_t)n * MMF_DIM * 4), b_head = al((size_t)n * 3 * 4), b_mod = al((size_t)n);
  const size_t o_sim = 0, o_disc = o_sim + al((size_t)n * 4), o_vs = o_disc + al((size_t)n * 4);
  const size_t o_vr = o_vs + al((size_t)n * top_k * 4), o_x = o_vr + al((size_t)n * top_k * 8);
  const size_t o_probs = o_x + al((size_t)n * 5 * 4), o_verdict = o_probs + al((size_t)n * 2 * 4);
  const size_t o_conf = o_verdict + al((size_t)n * 4), out_bytes = o_conf + al((size_t)n * 4);
  const size_t in_bytes = 2 * b_emb + b_head + 2 * b_mod;
  const size_t total = in_bytes + out_bytes;
  int rc = ensure_pinned(h, out_bytes);
  if (rc != MMF_OK) return rc;
  if (total > h->io_bytes) {
    MMF_CUDA_OK(h, cudaStreamSynchronize(st));
    if (h->io) MMF_CUDA_OK(h, cudaFree(h->io));
    h->io = nullptr;
    h->io_bytes = 0;
    MMF_CUDA_OK(h, cudaMalloc(&h->io, total));
    h->io_bytes = total;
  }
  char* dp = (char*)h->io;
  float* d_text = (float*)dp;
  float* d_img = (float*)(dp + b_emb);
  float* d_head = (float*)(dp + 2 * b_emb);
  unsigned char* d_mod_in = (unsigned char*)(dp + 2 * b_emb + b_head);
  unsigned char* d_mod = d_mod_in + b_mod;
  char* dout = dp + in_bytes;
  float* d_sim = (float*)(dout + o_sim);
  float* d_disc = (float*)(dout + o_disc);
  float* d_vs = (float*)(dout + o_vs);
  int64_t* d_vr = (int64_t*)(dout + o_vr);
  float* d_x = (float*)(dout + o_x);
  float* d_probs = (float*)(dout + o_probs);
  int32_t* d_verdict = (int32_t*)(dout + o_verdict);
  float* d_conf = (float*)(dout + o_conf);

  // straight from the caller's buffers: asynchronous DMA when they are pinned, staged by the driver otherwise
  MMF_CUDA_OK(h, cudaMemcpyAsync(d_text, text_host, (size_t)n * MMF_DIM * 4, cudaMemcpyHostToDevice, st));
  MMF_CUDA_OK(h, cudaMemcpyAsync(d_img, image_host, (size_t)n * MMF_DIM * 4, cudaMemcpyHostToDevice, st));
  MMF_CUDA_OK(h, cudaMemcpyAsync(d_head, head_host, (size_t)n * 3 * 4, cudaMemcpyHostToDevice, st));
  if (modality_host) MMF_CUDA_OK(h, cudaMemcpyAsync(d_mod_in, modality_host, (size_t)n, cudaMemcpyHostToDevice, st));

  rc = mmf_cosine_pairs(h, d_text, d_img, n, MMF_DIM, 0.0, d_sim, nullptr, st);
  if (rc != MMF_OK) return rc;
  if (h->vault_loaded) {
    rc = search_dispatch(h, d_img, n, top_k, threshold, algo, d_vs, d_vr, nullptr, d_disc, st, "score_batch_host");
  } else {      // reference: vault_loaded == False -> zero discrepancy, no matches (misinfo_forensics.py:422-428)
    rc = mmf_fill_empty(h, n, top_k, d_vs, d_vr, nullptr, d_disc, st);
  }
  if (rc != MMF_OK) return rc;
  mmf::assemble_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_head, modality_host ? d_mod_in : nullptr, n, d_sim,
                                                                          d_disc, d_x, d_mod);
  MMF_LAUNCH_OK(h);
  rc = mmf_verdict_batch(h, d_x, d_mod, n, d_probs, d_verdict, d_conf, st);
  if (rc != MMF_OK) return rc;
  char* hp = (char*)h->pinned;
  MMF_CUDA_OK(h, cudaMemcpyAsync(hp, dout, out_bytes, cudaMemcpyDeviceToHost, st));
  MMF_CUDA_OK(h, cudaStreamSynchronize(st));
  if (out_clip_similarity) memcpy(out_clip_similarity, hp + o_sim, (size_t)n * 4);
  if (out_vault_discrepancy) memcpy(out_vault_discrepancy, hp + o_disc, (size_t)n * 4);
  if (out_vault_scores) memcpy(out_vault_scores, hp + o_vs, (size_t)n * top_k * 4);
  if (out_vault_rows) memcpy(out_vault_rows, hp + o_vr, (size_t)n * top_k * 8);
  if (out_scores5) memcpy(out_scores5, hp + o_x, (size_t)n * 5 * 4);
  memcpy(out_probs, hp + o_probs, (size_t)n * 2 * 4);
  if (out_verdict) memcpy(out_verdict, hp + o_verdict, (size_t)n * 4);
  if (out_confidence) memcpy(out_confidence, hp + o_conf, (size_t)n * 4);
  return MMF_OK;
}
