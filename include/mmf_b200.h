/* libmmf_b200.so -- C ABI of the B200-native scoring hot path.
 *
 * The reference (yashingle-ai/Multi-Modal-Misinformation-Detection-with-Explanation-
 * Generation) has NO plugin / FFI boundary: the hot path is inline torch / NumPy
 * arithmetic inside Python methods.  Each entry point below therefore names the inline
 * reference lines it replaces (paths relative to the reference checkout); the binding a
 * maintainer adds on the reference side is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no C++ types or exceptions cross the boundary; every call returns an int
 *    status (MMF_OK == 0, negatives are errors) and mmf_last_error() has the detail;
 *  - unless a name ends in _host, data pointers are DEVICE pointers on the handle's
 *    device (torch.Tensor.data_ptr() passes zero-copy) and the call is ASYNCHRONOUS on
 *    the cudaStream_t given as `stream` (torch.cuda.current_stream().cuda_stream); no
 *    hidden synchronisation, with one exception: when a call needs more scratch than any
 *    earlier one (first call, larger batch) the scratch is re-allocated, which waits for
 *    the device -- and is refused during a stream capture: run a step once before
 *    capturing it.  *_host calls take host pointers, include the H2D / D2H
 *    copies and return after the result is in host memory;
 *  - a handle is not thread-safe: one handle per (process, device), calls serialised by
 *    the caller (the reference is single-threaded and synchronous);
 *  - there is no CPU fallback: without a CUDA device mmf_create fails with
 *    MMF_ERR_NO_DEVICE.
 */
#ifndef MMF_B200_H
#define MMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmf_handle mmf_handle;
typedef void* mmf_stream_t; /* cudaStream_t */

enum mmf_status {
  MMF_OK = 0,
  MMF_ERR_BAD_ARG = -1,
  MMF_ERR_CUDA = -2,
  MMF_ERR_NOT_LOADED = -3, /* vault / fusion weights not loaded (reference: vault_loaded == False) */
  MMF_ERR_NO_DEVICE = -4,
  MMF_ERR_UNSUPPORTED = -5,
  MMF_ERR_NOMEM = -6,
  MMF_ERR_NCCL = -7 /* the NCCL library is missing or a collective failed (row-sharded search only) */
};

enum mmf_dtype { MMF_F32 = 0, MMF_F16 = 1, MMF_BF16 = 2, MMF_F64 = 3 };

/* How the vault shard is kept in HBM.
 *  MMF_VAULT_FP32: fp32-exact.  Rows are L2-normalised in fp32 and stored as two fp16
 *                  planes hi/lo (v*2^8 = hi + lo, 22+ significant bits, 4 B/element --
 *                  the same bytes as fp32) so that the SAME resident copy feeds both the
 *                  HBM-streaming kernel (reconstructs hi+lo in fp32) and the tcgen05
 *                  kernel (3 f16 MMA passes, fp32 accumulate).  Tolerance 1e-5.
 *  MMF_VAULT_BF16: rows normalised in fp32, rounded to bf16 (2 B/element). Tolerance 1e-2. */
enum mmf_vault_mode { MMF_VAULT_FP32 = 0, MMF_VAULT_BF16 = 1 };

/* Which search kernel mmf_vault_search uses. AUTO: streaming kernel for small query
 * batches (HBM-bound), tcgen05 kernel for large ones (tensor-bound). */
enum mmf_search_algo { MMF_ALGO_AUTO = 0, MMF_ALGO_STREAM = 1, MMF_ALGO_MMA = 2 };

#define MMF_MAX_TOP_K 256
#define MMF_FUSION_PARAMS 2530 /* 64*5+64 + 32*64+32 + 2*32+2, misinfo_forensics.py:83-90 */

const char* mmf_version(void);
int mmf_arch(void); /* 100: built for sm_100a only */
const char* mmf_status_string(int status);

/* One handle per (process, device).  Owns the uploaded vault shard, the packed fusion
 * weights and scratch; nothing else. */
int mmf_create(int device_ordinal, mmf_handle** out);
int mmf_destroy(mmf_handle* h);
const char* mmf_last_error(const mmf_handle* h);

/* ---- CLIP caption<->image consistency ------------------------------------------------
 * Replaces misinfo_forensics.py:399-404 (analyze_consistency: normalise both embeddings,
 * dot) and clip_similarity_engine.py:103-111 (same cosine + `sim >= threshold` label);
 * also misinfo_forensics.py:481-484 (caption<->headline text similarity).
 * a, b: (n_pairs, dim) fp32 row-major; out_sim: (n_pairs) fp32;
 * out_match: (n_pairs) uint8, 1 where (double)sim >= match_threshold; may be NULL. */
int mmf_cosine_pairs(mmf_handle* h, const float* a, const float* b, int64_t n_pairs, int dim,
                     double match_threshold, float* out_sim, uint8_t* out_match, mmf_stream_t stream);

/* ---- Truth Vault ---------------------------------------------------------------------
 * mmf_vault_load replaces the per-query renormalisation of misinfo_forensics.py:443-445
 * (Vn = V / ||V||, hoisted: done once at load) and uploads this rank's row shard.
 * rows: (n_rows, dim) row-major of src_dtype, host pointer (rows_on_device == 0) or device
 * pointer.  row_offset: global id of the shard's first row (row-sharding).
 * dim must be 512 (CLIP ViT-B/32 projection, misinfo_forensics.py:78-79). Synchronous. */
int mmf_vault_load(mmf_handle* h, const void* rows, int rows_on_device, int64_t n_rows, int dim,
                   int src_dtype, int vault_mode, int64_t row_offset);
int mmf_vault_unload(mmf_handle* h);
int mmf_vault_info(const mmf_handle* h, int64_t* n_rows, int* dim, int* vault_mode, int64_t* row_offset);

/* Replaces misinfo_forensics.py:438-440 (query normalise), :446 (similarities), :449-450
 * (argsort top-k, descending; ties: higher row id first; NaN ranks first like np.argsort)
 * and :463-464 (discrepancy = top if (double)top > threshold else 0).
 * queries: (n_queries, dim) fp32, un-normalised embeddings.
 * out_scores (n_queries, top_k) fp32, out_rows (n_queries, top_k) int64 GLOBAL row ids;
 * if top_k > n_rows the tail is filled with NaN / -1 (the reference returns n_rows items).
 * out_discrepancy (n_queries) fp32, may be NULL. */
int mmf_vault_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                     int algo, float* out_scores, int64_t* out_rows, float* out_discrepancy,
                     mmf_stream_t stream);

/* Same, host buffers: H2D of the queries, search, D2H of the results, stream sync. */
int mmf_vault_search_host(mmf_handle* h, const float* queries_host, int64_t n_queries, int top_k,
                          double threshold, int algo, float* out_scores_host, int64_t* out_rows_host,
                          float* out_discrepancy_host);

/* Row-sharded search: this rank's local top-k as packed candidates, one
 * uint64 per candidate = (order-preserving score key << 32) | global row id; 0 = empty.
 * out_packed: (n_queries, top_k) uint64.  The caller all-gathers the buffers of all
 * ranks (NCCL) and calls mmf_topk_merge. */
int mmf_vault_search_candidates(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int algo,
                                uint64_t* out_packed, mmf_stream_t stream);

/* packed: (n_lists, n_queries, k_in) uint64 candidates (e.g. the all-gathered shards).
 * Selects the global top_k per query under the same total order, so the result is
 * independent of the sharding.  Outputs as mmf_vault_search.  Any lists are accepted; lists
 * sorted descending with their empty slots last (what every search entry of this library
 * writes) take a faster path (binary-search ranks instead of a radix select). */
int mmf_topk_merge(mmf_handle* h, const uint64_t* packed, int n_lists, int64_t n_queries, int k_in, int top_k,
                   double threshold, float* out_scores, int64_t* out_rows, float* out_discrepancy,
                   mmf_stream_t stream);

/* ---- Row-sharded search with the collective owned by the library -------------------
 * Sharded form of misinfo_forensics.py:443-450: rank r holds rows [row_offset, row_offset + n_rows) of the vault
 * (mmf_vault_load), every rank searches its shard, ONE ncclAllGather moves the packed per-shard top-k candidates
 * (8 B each) over NVLink and every rank merges them -- identical results on every rank, bit-identical to the
 * unsharded search.  NCCL is bound at run time (dlopen of libnccl.so.2, or $MMF_NCCL_LIB); a process that
 * already carries one (PyTorch) shares it.
 * Bootstrap: rank 0 calls mmf_shard_unique_id and hands the MMF_SHARD_ID_BYTES bytes to its peers out of band
 * (torch.distributed store, MPI, a file ...); then EVERY rank calls mmf_shard_init (collective, blocking).
 * world == 1 needs no id and no NCCL.  One shard group per handle; mmf_destroy finalises it. */
#define MMF_SHARD_ID_BYTES 128 /* sizeof(ncclUniqueId) */
int mmf_shard_unique_id(void* id_out);
int mmf_shard_init(mmf_handle* h, int rank, int world, const void* unique_id);
int mmf_shard_finalize(mmf_handle* h);
int mmf_shard_info(const mmf_handle* h, int* rank, int* world, int* nccl_version);
/* mmf_vault_search over ALL shards: local search (top_k candidates per query, global row ids) + all-gather +
 * merge, asynchronous on `stream` (the collective runs on that stream too); every rank must call it with the
 * same queries, n_queries and top_k.  Outputs as mmf_vault_search, the same on every rank. */
int mmf_vault_search_sharded(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                             int algo, float* out_scores, int64_t* out_rows, float* out_discrepancy,
                             mmf_stream_t stream);
/* The collective alone, for callers that time or overlap the phases themselves (bench.py does): n_keys packed
 * candidates of this rank (mmf_vault_search_candidates) -> out_gathered (world, n_keys) in rank order.
 * Then mmf_topk_merge(h, out_gathered, world, n_queries, k_in, ...). */
int mmf_shard_all_gather(mmf_handle* h, const uint64_t* local_packed, int64_t n_keys, uint64_t* out_gathered,
                         mmf_stream_t stream);

/* ---- Candidate exchange over NVLink peer memory (row-sharded search) ------------------
 * Alternative to "mmf_vault_search_candidates + NCCL all-gather + mmf_topk_merge": every rank stores its
 * candidates straight into the peers' gather buffers and the merge kernel waits on per-rank flags, so the
 * sharded search is three launches and no library collective (csrc/exchange.cu).
 * The buffers are SYMMETRIC MEMORY owned by the caller (e.g. torch.distributed._symmetric_memory): one
 * allocation of bytes_per_rank on every rank, mapped into every process; peer_ptrs is a HOST array of `world`
 * device addresses (peer_ptrs[r] = rank r's buffer as seen from this process; peer_ptrs[rank] = the local one).
 * mmf_exchange_layout says how many bytes an exchange of (n_queries, k_in) candidates per rank needs.
 * Protocol: all ranks attach, the caller barriers, then all ranks call mmf_vault_search_exchange the same
 * number of times with the same (n_queries, top_k, k_local).  world <= 16.  Synchronous: attach / detach. */
int mmf_exchange_layout(int world, int64_t n_queries, int k_in, int64_t* gather_bytes_per_parity, int64_t* bytes_needed);
int mmf_exchange_attach(mmf_handle* h, int rank, int world, const uint64_t* peer_ptrs, int64_t bytes_per_rank);
int mmf_exchange_detach(mmf_handle* h);
/* The two phases of an exchange, for callers that want to put other work between them (and for single-device
 * checks, which must enqueue every rank's push before any rank's merge -- see csrc/exchange.cu):
 * mmf_vault_search_push: local search (k_local candidates per query) + push into every rank's buffer + flag;
 * never waits.  mmf_vault_exchange_merge: device-side wait for all ranks' candidates of the pending exchange +
 * merge into the global top_k (top_k >= k_local).  One exchange may be pending per handle. */
int mmf_vault_search_push(mmf_handle* h, const float* queries, int64_t n_queries, int k_local, int algo,
                          mmf_stream_t stream);
int mmf_vault_exchange_merge(mmf_handle* h, int top_k, double threshold, float* out_scores, int64_t* out_rows,
                             float* out_discrepancy, mmf_stream_t stream);
/* Both phases in one call: local search (k_local = min(top_k, rows per rank) candidates per query) + push +
 * wait + merge.  Outputs as mmf_vault_search, identical on every rank; asynchronous on `stream`. */
int mmf_vault_search_exchange(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int k_local,
                              double threshold, int algo, float* out_scores, int64_t* out_rows,
                              float* out_discrepancy, mmf_stream_t stream);

/* ---- Fusion judge --------------------------------------------------------------------
 * Weights of MultiModalMisinfoDetector.fusion_layer (misinfo_forensics.py:83-90), in the
 * order and nn.Linear layout of the .pth (train_fusion_judge.py:259-267):
 * [0.weight (64,5) | 0.bias (64) | 3.weight (32,64) | 3.bias (32) | 5.weight (2,32) | 5.bias (2)]
 * = MMF_FUSION_PARAMS fp32, HOST pointer.  Call again after the trainer mutates them. */
int mmf_fusion_load(mmf_handle* h, const float* params_host);

/* Replaces misinfo_forensics.py:587-608: x (n,5) fp32 = [ai, misinfo, deepfake,
 * clip_similarity, vault_discrepancy] -> probs (n,2) fp32 [real, fake] (softmax),
 * verdict (n) int32 = fake > 0.5, confidence (n) fp32.  out_verdict/out_confidence may be NULL. */
int mmf_fusion_forward(mmf_handle* h, const float* x, int64_t n, float* out_probs, int32_t* out_verdict,
                       float* out_confidence, mmf_stream_t stream);

/* Batched MisinfoForensics.analyze downstream of the encoders (misinfo_forensics.py:
 * 866-900): per row, modality bit0 = has text, bit1 = has image/video.
 *   both  -> fusion judge on [ai, misinfo, deepfake, clip_sim, vault_disc]
 *   text  -> fake = misinfo;  visual -> fake = max(deepfake, vault_disc);  none -> 0.5
 * (fallbacks clamped to [0,1], real = 1 - fake).  scores: (n,5) fp32 as above.
 * Outputs as mmf_fusion_forward. */
int mmf_verdict_batch(mmf_handle* h, const float* scores, const uint8_t* modality, int64_t n, float* out_probs,
                      int32_t* out_verdict, float* out_confidence, mmf_stream_t stream);

/* Score assembly + verdict in ONE launch, device buffers (batched misinfo_forensics.py:794-809 + :866-900):
 * head (n,3) = [ai, misinfo, deepfake]; modality (n) uint8 (bit0 text, bit1 visual) or NULL = both;
 * clip_similarity / vault_discrepancy (n) are read AND masked in place (a skipped modality's score is 0);
 * out_scores5 (n,5) = the fusion inputs; outputs as mmf_fusion_forward. */
int mmf_verdict_assemble(mmf_handle* h, const float* head, const uint8_t* modality, int64_t n, float* clip_similarity,
                         float* vault_discrepancy, float* out_scores5, float* out_probs, int32_t* out_verdict,
                         float* out_confidence, mmf_stream_t stream);

/* The whole path for a batch, DEVICE buffers in and out, asynchronous on `stream`: caption/image cosine, vault
 * search of the image embeddings (zero discrepancy / NaN, -1 matches when no vault is loaded), score assembly +
 * fusion judge / fallback verdict.  Shapes as mmf_score_batch_host; every output but verdict / confidence is
 * required.  Single-GPU / replica vaults (row-sharded: mmf_cosine_pairs + mmf_vault_search_sharded +
 * mmf_verdict_assemble). */
int mmf_score_batch(mmf_handle* h, const float* text, const float* image, const float* head, const uint8_t* modality,
                    int64_t n, int top_k, double threshold, int algo, float* out_clip_similarity,
                    float* out_vault_discrepancy, float* out_vault_scores, int64_t* out_vault_rows, float* out_scores5,
                    float* out_probs, int32_t* out_verdict, float* out_confidence, mmf_stream_t stream);

/* The whole path for a batch with HOST buffers in and out (the end-to-end call): H2D of the embeddings,
 * caption/image cosine, vault search (zero discrepancy / no matches when no vault is loaded, misinfo_forensics.py:
 * 422-428), score assembly with the skipped-modality zeros of :794-809 + fusion judge / fallback verdict, one D2H,
 * one synchronisation; returns when the results are in host memory.  Pinned host buffers make the copies
 * asynchronous DMA.  text/image: (n,512) fp32; head: (n,3) fp32 = [ai, misinfo, deepfake]; modality: (n) uint8
 * (bit0 text, bit1 visual) or NULL = both.  Outputs (any but out_probs may be NULL): clip_similarity (n),
 * vault_discrepancy (n), vault_scores (n,top_k), vault_rows (n,top_k) int64, scores5 (n,5) = the fusion inputs,
 * probs (n,2) [real,fake], verdict (n) int32, confidence (n).  Single-GPU / replica vaults only. */
int mmf_score_batch_host(mmf_handle* h, const float* text_host, const float* image_host, const float* head_host,
                         const uint8_t* modality_host, int64_t n, int top_k, double threshold, int algo,
                         float* out_clip_similarity, float* out_vault_discrepancy, float* out_vault_scores,
                         int64_t* out_vault_rows, float* out_scores5, float* out_probs, int32_t* out_verdict,
                         float* out_confidence);
/* The same call split in two, with two slots (0, 1): submit enqueues H2D + kernels + D2H on the handle's own
 * streams and returns at once; collect waits for that slot and copies the results out.  A caller that submits batch
 * i+1 before collecting batch i overlaps the copies of one with the kernels of the other (throughput serving).
 * The input buffers of a slot must stay valid until it is collected; mmf_score_batch_host == submit + collect on
 * slot 0.  The host entry points use their own streams and their own scratch, so they may be mixed with the
 * asynchronous entry points; do not call mmf_fusion_load / mmf_vault_load while a slot is pending. */
int mmf_score_batch_submit(mmf_handle* h, int slot, const float* text_host, const float* image_host,
                           const float* head_host, const uint8_t* modality_host, int64_t n, int top_k,
                           double threshold, int algo);
int mmf_score_batch_collect(mmf_handle* h, int slot, float* out_clip_similarity, float* out_vault_discrepancy,
                            float* out_vault_scores, int64_t* out_vault_rows, float* out_scores5, float* out_probs,
                            int32_t* out_verdict, float* out_confidence);

/* Switches of a handle (A/B and triage; production needs none).  Read once from the environment at mmf_create
 * (MMF_OPT_<NAME>), never on the search path.  Names: "screen" (fp32-exact vaults, top_k <= 16: 1 = screened
 * search, default; 0 = 3-pass kernel), "fused_push" (peer-memory exchange: the search pushes its winners itself,
 * default 1), "debug" (tcgen05 search triage bits), "force_cg" (1 / 2 CTAs per MMA, 0 = automatic),
 * "flat_schedule", "epi_parity" (-1 = by strip length), "stream_tma" (1 = TMA-staged streaming kernel, default),
 * "lockstep" (1 = producers of the tcgen05 search that sweep the same vault tiles pace each other, default). */
int mmf_set_option(mmf_handle* h, const char* name, int value);
int mmf_get_option(const mmf_handle* h, const char* name, int* value);

/* Host-only self check of the tcgen05 search's work decomposition for a (n_queries, n_rows) problem on a
 * device with sm_count SMs: MMF_OK iff every (query-tile group, vault tile) unit is scheduled exactly once,
 * strip ids are unique and the load is balanced.  Needs no GPU (used by the CPU test-suite). */
int mmf_mma_plan_check(int64_t n_queries, int64_t n_rows, int sm_count, int64_t* out_units, int* out_pairs,
                       int* out_cg);

/* Host-only model of the experimental histogram bound of the tcgen05 search (env MMF_MMA_BOUND=hist, large
 * top_k): the lower bound of the top_k-th best of `scores` (n fp32 HOST values) that the per-query score
 * histogram yields, -inf when it yields none.  Same binning code as the kernel; needs no GPU. */
int mmf_mma_hist_bound(const float* scores, int64_t n, int top_k, float* out_bound);

/* The error bound eps (score units) of the screened fp32-exact search: |one-pass fp16 hi-plane score - exact
 * score| <= eps for unit-norm rows and queries; rows within 2*eps of the k-th best approximate score are
 * re-scored exactly.  Host-only. */
double mmf_mma_screen_eps(void);

/* Number of kernel launches this handle has issued (for bench.py's gpu_launches), and of library collectives. */
int64_t mmf_launch_count(const mmf_handle* h);
int64_t mmf_collective_count(const mmf_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MMF_B200_H */
