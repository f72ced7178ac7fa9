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

extern "C" int mmf_set_option(mmf_handle* h, const char* name, int value);

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
  for (const char* name : {"screen", "fused_push", "debug", "force_cg", "flat_schedule", "epi_parity", "stream_tma", "lockstep"}) {
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
  else if (!strcmp(name, "epi_parity")) o.epi_parity = value < 0 ? -1 : (value != 0);
  else if (!strcmp(name, "lockstep")) o.lockstep = value != 0;
  else if (!strcmp(name, "stream_tma")) o.stream_tma = value != 0;
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
  else if (!strcmp(name, "epi_parity")) *value = o.epi_parity;
  else if (!strcmp(name, "lockstep")) *value = o.lockstep;
  else if (!strcmp(name, "stream_tma")) *value = o.stream_tma;
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

// RAII: the host-buffer entry points run on the handle's own streams with their own scratch arena (arena 1), so
// that they never share counters or candidate lists with an asynchronous search in flight on a caller's stream
struct HostArena {
  mmf_handle* h;
  explicit HostArena(mmf_handle* hh) : h(hh) { h->scratch_sel = 1; }
  ~HostArena() { h->scratch_sel = 0; }
};

extern "C" int mmf_vault_search_host(mmf_handle* h, const float* queries_host, int64_t n_queries, int top_k,
                                     double threshold, int algo, float* out_scores_host, int64_t* out_rows_host,
                                     float* out_discrepancy_host) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_queries < 0 || top_k <= 0 || top_k > MMF_MAX_TOP_K ||
      (n_queries > 0 && (!queries_host || !out_scores_host || !out_rows_host)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_host: bad argument");
  if (!h->vault_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "vault_search_host: no vault loaded");
  if (n_queries == 0) return MMF_OK;
  mmf_host_slot& sl = h->slot[0];
  if (sl.busy) return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_host: a submitted batch is pending in slot 0 (collect it first)");
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  // device: [queries | scores | rows | disc]; the pinned staging mirrors the outputs
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t bq = al((size_t)n_queries * MMF_DIM * 4), bs = al((size_t)n_queries * top_k * 4);
  const size_t br = al((size_t)n_queries * top_k * 8), bd = al((size_t)n_queries * 4);
  int rc = ensure_slot(h, sl, bq + bs + br + bd, bs + br + bd);
  if (rc != MMF_OK) return rc;
  HostArena arena(h);
  char* hp = (char*)sl.pinned;
  char* dp = (char*)sl.io;
  MMF_CUDA_OK(h, cudaMemcpyAsync(dp, queries_host, (size_t)n_queries * MMF_DIM * 4, cudaMemcpyHostToDevice, st));
  float* d_scores = (float*)(dp + bq);
  int64_t* d_rows = (int64_t*)(dp + bq + bs);
  float* d_disc = (float*)(dp + bq + bs + br);
  rc = search_dispatch(h, (const float*)dp, n_queries, top_k, threshold, algo, d_scores, d_rows, nullptr, d_disc, st,
                       "vault_search_host");
  if (rc != MMF_OK) return rc;
  MMF_CUDA_OK(h, cudaMemcpyAsync(hp, dp + bq, bs + br + bd, cudaMemcpyDeviceToHost, st));
  MMF_CUDA_OK(h, cudaStreamSynchronize(st));
  memcpy(out_scores_host, hp, (size_t)n_queries * top_k * 4);
  memcpy(out_rows_host, hp + bs, (size_t)n_queries * top_k * 8);
  if (out_discrepancy_host) memcpy(out_discrepancy_host, hp + bs + br, (size_t)n_queries * 4);
  return MMF_OK;
}

// ---- batched analyze downstream of the encoders, HOST buffers in and out ---------------------------------
// The whole hot path for a batch: H2D of the embeddings, caption/image cosine (K1), vault search (K2 / K3), score
// assembly with the modality rules of misinfo_forensics.py:794-809 + fusion judge / fallback verdict (K5, one
// launch), ONE D2H of all results.  Two slots and three streams (H2D, compute, D2H) let a caller keep two batches in
// flight: the copies of one overlap the kernels of the other (mmf_score_batch_submit / _collect); the synchronous
// mmf_score_batch_host is submit + collect on slot 0.
int mmf_assemble_verdict(mmf_handle* h, const float* head, const uint8_t* modality_in, int64_t n, float* sim, float* disc,
                         float* x, float* out_probs, int32_t* out_verdict, float* out_conf, cudaStream_t st);

namespace {
struct BatchLayout {       // byte offsets inside a slot's device buffer: inputs, then the output block
  size_t text, img, head, mod, in_bytes;
  size_t o_sim, o_disc, o_vs, o_vr, o_x, o_probs, o_verdict, o_conf, out_bytes;
  BatchLayout(int64_t n, int top_k) {
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b_emb = al((size_t)n * MMF_DIM * 4);
    text = 0; img = b_emb; head = 2 * b_emb; mod = head + al((size_t)n * 3 * 4); in_bytes = mod + al((size_t)n);
    o_sim = 0; o_disc = o_sim + al((size_t)n * 4); o_vs = o_disc + al((size_t)n * 4);
    o_vr = o_vs + al((size_t)n * top_k * 4); o_x = o_vr + al((size_t)n * top_k * 8);
    o_probs = o_x + al((size_t)n * 5 * 4); o_verdict = o_probs + al((size_t)n * 2 * 4);
    o_conf = o_verdict + al((size_t)n * 4); out_bytes = o_conf + al((size_t)n * 4);
  }
};
}  // namespace

// Score assembly + verdict for device-resident scores (the tail of the batched path as ONE launch): see the header.
extern "C" int mmf_verdict_assemble(mmf_handle* h, const float* head, const uint8_t* modality, int64_t n, float* clip_similarity,
                                    float* vault_discrepancy, float* out_scores5, float* out_probs, int32_t* out_verdict,
                                    float* out_confidence, mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n < 0 || (n > 0 && (!head || !clip_similarity || !vault_discrepancy || !out_scores5 || !out_probs)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "verdict_assemble: bad argument");
  return mmf_assemble_verdict(h, head, modality, n, clip_similarity, vault_discrepancy, out_scores5, out_probs, out_verdict,
                              out_confidence, (cudaStream_t)stream);
}

// The whole path for a batch, DEVICE buffers in and out, asynchronous on `stream`: see the header.
extern "C" int mmf_score_batch(mmf_handle* h, const float* text, const float* image, const float* head, const uint8_t* modality,
                               int64_t n, int top_k, double threshold, int algo, float* out_clip_similarity,
                               float* out_vault_discrepancy, float* out_vault_scores, int64_t* out_vault_rows, float* out_scores5,
                               float* out_probs, int32_t* out_verdict, float* out_confidence, mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n < 0 || top_k <= 0 || top_k > MMF_MAX_TOP_K ||
      (n > 0 && (!text || !image || !head || !out_clip_similarity || !out_vault_discrepancy || !out_vault_scores || !out_vault_rows ||
                 !out_scores5 || !out_probs)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch: bad argument (n=%lld top_k=%d)", (long long)n, top_k);
  if (!h->fusion_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "score_batch: fusion weights not loaded");
  if (n == 0) return MMF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = mmf_cosine_pairs(h, text, image, n, MMF_DIM, 0.0, out_clip_similarity, nullptr, stream);
  if (rc != MMF_OK) return rc;
  if (h->vault_loaded)
    rc = search_dispatch(h, image, n, top_k, threshold, algo, out_vault_scores, out_vault_rows, nullptr, out_vault_discrepancy, st,
                         "score_batch");
  else
    rc = mmf_fill_empty(h, n, top_k, out_vault_scores, out_vault_rows, nullptr, out_vault_discrepancy, st);
  if (rc != MMF_OK) return rc;
  return mmf_assemble_verdict(h, head, modality, n, out_clip_similarity, out_vault_discrepancy, out_scores5, out_probs, out_verdict,
                              out_confidence, st);
}

extern "C" int mmf_score_batch_submit(mmf_handle* h, int slot, const float* text_host, const float* image_host,
                                      const float* head_host, const uint8_t* modality_host, int64_t n, int top_k,
                                      double threshold, int algo) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (slot < 0 || slot > 1 || n < 0 || top_k <= 0 || top_k > MMF_MAX_TOP_K || (n > 0 && (!text_host || !image_host || !head_host)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_submit: bad argument (slot=%d n=%lld top_k=%d)", slot, (long long)n, top_k);
  if (!h->fusion_loaded) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "score_batch_submit: fusion weights not loaded");
  mmf_host_slot& sl = h->slot[slot];
  if (sl.busy) return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_submit: slot %d has not been collected", slot);
  sl.n = n;
  sl.top_k = top_k;
  if (n == 0) { sl.busy = true; return MMF_OK; }
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  const BatchLayout L(n, top_k);
  int rc = ensure_slot(h, sl, L.in_bytes + L.out_bytes, L.out_bytes);
  if (rc != MMF_OK) return rc;                       // (every error path leaves the slot free)
  HostArena arena(h);
  char* dp = (char*)sl.io;
  char* dout = dp + L.in_bytes;
  float* d_text = (float*)(dp + L.text);
  float* d_img = (float*)(dp + L.img);
  float* d_head = (float*)(dp + L.head);
  uint8_t* d_mod = (uint8_t*)(dp + L.mod);
  // H2D straight from the caller's buffers (asynchronous DMA when they are pinned; they must stay valid until the
  // batch is collected).  The image embeddings go first: the search needs nothing else.
  cudaStream_t up = h->h2d_stream, st = h->own_stream, down = h->d2h_stream;
  MMF_CUDA_OK(h, cudaMemcpyAsync(d_img, image_host, (size_t)n * MMF_DIM * 4, cudaMemcpyHostToDevice, up));
  MMF_CUDA_OK(h, cudaMemcpyAsync(d_text, text_host, (size_t)n * MMF_DIM * 4, cudaMemcpyHostToDevice, up));
  MMF_CUDA_OK(h, cudaMemcpyAsync(d_head, head_host, (size_t)n * 3 * 4, cudaMemcpyHostToDevice, up));
  if (modality_host) MMF_CUDA_OK(h, cudaMemcpyAsync(d_mod, modality_host, (size_t)n, cudaMemcpyHostToDevice, up));
  MMF_CUDA_OK(h, cudaEventRecord(sl.ev_in, up));
  MMF_CUDA_OK(h, cudaStreamWaitEvent(st, sl.ev_in, 0));
  float* d_sim = (float*)(dout + L.o_sim);
  float* d_disc = (float*)(dout + L.o_disc);
  rc = mmf_cosine_pairs(h, d_text, d_img, n, MMF_DIM, 0.0, d_sim, nullptr, st);
  if (rc != MMF_OK) return rc;
  if (h->vault_loaded) {
    rc = search_dispatch(h, d_img, n, top_k, threshold, algo, (float*)(dout + L.o_vs), (int64_t*)(dout + L.o_vr), nullptr,
                         d_disc, st, "score_batch_submit");
  } else {      // reference: vault_loaded == False -> zero discrepancy, no matches (misinfo_forensics.py:422-428)
    rc = mmf_fill_empty(h, n, top_k, (float*)(dout + L.o_vs), (int64_t*)(dout + L.o_vr), nullptr, d_disc, st);
  }
  if (rc != MMF_OK) return rc;
  rc = mmf_assemble_verdict(h, d_head, modality_host ? d_mod : nullptr, n, d_sim, d_disc, (float*)(dout + L.o_x),
                            (float*)(dout + L.o_probs), (int32_t*)(dout + L.o_verdict), (float*)(dout + L.o_conf), st);
  if (rc != MMF_OK) return rc;
  MMF_CUDA_OK(h, cudaEventRecord(sl.ev_done, st));
  MMF_CUDA_OK(h, cudaStreamWaitEvent(down, sl.ev_done, 0));
  MMF_CUDA_OK(h, cudaMemcpyAsync(sl.pinned, dout, L.out_bytes, cudaMemcpyDeviceToHost, down));
  MMF_CUDA_OK(h, cudaEventRecord(sl.ev_out, down));
  sl.busy = true;
  return MMF_OK;
}

extern "C" int mmf_score_batch_collect(mmf_handle* h, int slot, float* out_clip_similarity, float* out_vault_discrepancy,
                                       float* out_vault_scores, int64_t* out_vault_rows, float* out_scores5,
                                       float* out_probs, int32_t* out_verdict, float* out_confidence) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (slot < 0 || slot > 1) return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_collect: bad slot %d", slot);
  mmf_host_slot& sl = h->slot[slot];
  if (!sl.busy) return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_collect: nothing submitted in slot %d", slot);
  sl.busy = false;
  const int64_t n = sl.n;
  if (n == 0) return MMF_OK;
  if (!out_probs) return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_collect: null out_probs");
  MMF_CUDA_OK(h, cudaEventSynchronize(sl.ev_out));
  const BatchLayout L(n, sl.top_k);
  const char* hp = (const char*)sl.pinned;
  if (out_clip_similarity) memcpy(out_clip_similarity, hp + L.o_sim, (size_t)n * 4);
  if (out_vault_discrepancy) memcpy(out_vault_discrepancy, hp + L.o_disc, (size_t)n * 4);
  if (out_vault_scores) memcpy(out_vault_scores, hp + L.o_vs, (size_t)n * sl.top_k * 4);
  if (out_vault_rows) memcpy(out_vault_rows, hp + L.o_vr, (size_t)n * sl.top_k * 8);
  if (out_scores5) memcpy(out_scores5, hp + L.o_x, (size_t)n * 5 * 4);
  memcpy(out_probs, hp + L.o_probs, (size_t)n * 2 * 4);
  if (out_verdict) memcpy(out_verdict, hp + L.o_verdict, (size_t)n * 4);
  if (out_confidence) memcpy(out_confidence, hp + L.o_conf, (size_t)n * 4);
  return MMF_OK;
}

extern "C" int mmf_score_batch_host(mmf_handle* h, const float* text_host, const float* image_host, const float* head_host,
                                    const uint8_t* modality_host, int64_t n, int top_k, double threshold, int algo,
                                    float* out_clip_similarity, float* out_vault_discrepancy, float* out_vault_scores,
                                    int64_t* out_vault_rows, float* out_scores5, float* out_probs, int32_t* out_verdict,
                                    float* out_confidence) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n > 0 && !out_probs) return mmf_set_error(h, MMF_ERR_BAD_ARG, "score_batch_host: bad argument (null out_probs)");
  const int rc = mmf_score_batch_submit(h, 0, text_host, image_host, head_host, modality_host, n, top_k, threshold, algo);
  if (rc != MMF_OK) return rc;
  return mmf_score_batch_collect(h, 0, out_clip_similarity, out_vault_discrepancy, out_vault_scores, out_vault_rows,
                                 out_scores5, out_probs, out_verdict, out_confidence);
}
