// Row-sharded Truth-Vault search with the collective OWNED BY THE LIBRARY:
// every rank searches its own row shard (mmf_vault_load with row_offset), ONE ncclAllGather moves the per-shard
// top-k candidates (8 B each: order-preserving score key << 32 | global row id) over NVLink, and every rank merges
// the `world` lists under the same total order -- so the result is bit-identical to the unsharded search, on every
// rank, and a plain C caller needs nothing but this library and an out-of-band way to hand 128 bytes to its peers.
// Sharded form of misinfo_forensics.py:443-450 (similarities over all rows + argsort top-k).
//
// NCCL is bound at run time (dlopen): a process that already carries an NCCL (PyTorch does) shares that copy, a
// plain C program gets the system libnccl.so.2; programs that never shard never load it.  Only six entry points
// are used and their ABI has been stable since NCCL 2.0, so no NCCL header is needed at build time.
#include "common.cuh"

#include <dlfcn.h>
#include <cstring>
#include <new>

int mmf_search_dispatch_packed(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int algo,
                               uint64_t* out_packed, cudaStream_t st, const char* who);

namespace {

struct NcclId { char internal[MMF_SHARD_ID_BYTES]; };          // ncclUniqueId
typedef void* NcclComm;                                        // ncclComm_t
constexpr int kNcclUint64 = 5;                                 // ncclDataType_t::ncclUint64

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  const char* why = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* env = getenv("MMF_NCCL_LIB");
  if (env && *env) api.lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses
  if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) { api.why = "libnccl.so.2 not found (set MMF_NCCL_LIB)"; return &api; }
  *(void**)&api.GetUniqueId = dlsym(api.lib, "ncclGetUniqueId");
  *(void**)&api.CommInitRank = dlsym(api.lib, "ncclCommInitRank");
  *(void**)&api.CommDestroy = dlsym(api.lib, "ncclCommDestroy");
  *(void**)&api.AllGather = dlsym(api.lib, "ncclAllGather");
  *(void**)&api.GetErrorString = dlsym(api.lib, "ncclGetErrorString");
  *(void**)&api.GetVersion = dlsym(api.lib, "ncclGetVersion");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString)
    api.why = "libnccl.so.2 lacks an expected entry point";
  return &api;
}

struct ShardState {
  int rank = 0, world = 1;
  NcclComm comm = nullptr;
  uint64_t* local = nullptr;       // this rank's candidates of the current search (n_queries, top_k)
  uint64_t* gather = nullptr;      // (world, n_queries, top_k)
  size_t cap_keys = 0;             // keys `local` can hold; `gather` holds world times that
};

int nccl_fail(mmf_handle* h, const NcclApi* a, int rc, const char* what) {
  return mmf_set_error(h, MMF_ERR_NCCL, "%s failed: %s", what, a->GetErrorString ? a->GetErrorString(rc) : "?");
}

}  // namespace

extern "C" int mmf_shard_unique_id(void* id_out) {
  if (!id_out) return MMF_ERR_BAD_ARG;
  NcclApi* a = nccl_api();
  if (a->why) return MMF_ERR_NCCL;
  NcclId id;
  if (a->GetUniqueId(&id) != 0) return MMF_ERR_NCCL;
  memcpy(id_out, &id, sizeof id);
  return MMF_OK;
}

extern "C" int mmf_shard_init(mmf_handle* h, int rank, int world, const void* unique_id) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !unique_id))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "shard_init: bad argument (rank=%d world=%d)", rank, world);
  if (h->shard_state) return mmf_set_error(h, MMF_ERR_BAD_ARG, "shard_init: already initialised (mmf_shard_finalize first)");
  ShardState* s = new (std::nothrow) ShardState();
  if (!s) return mmf_set_error(h, MMF_ERR_NOMEM, "shard_init: out of host memory");
  s->rank = rank;
  s->world = world;
  if (world > 1) {
    NcclApi* a = nccl_api();
    if (a->why) { delete s; return mmf_set_error(h, MMF_ERR_NCCL, "shard_init: %s", a->why); }
    if (cudaSetDevice(h->device) != cudaSuccess) { delete s; return mmf_set_error(h, MMF_ERR_CUDA, "shard_init: cudaSetDevice failed"); }
    NcclId id;
    memcpy(&id, unique_id, sizeof id);
    const int rc = a->CommInitRank(&s->comm, world, id, rank);      // collective: returns when all ranks have joined
    if (rc != 0) { delete s; return nccl_fail(h, a, rc, "ncclCommInitRank"); }
  }
  h->shard_state = s;
  return MMF_OK;
}

extern "C" int mmf_shard_finalize(mmf_handle* h) {
  if (!h || !h->shard_state) return MMF_OK;
  ShardState* s = (ShardState*)h->shard_state;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (s->comm) nccl_api()->CommDestroy(s->comm);
  if (s->local) cudaFree(s->local);
  if (s->gather) cudaFree(s->gather);
  delete s;
  h->shard_state = nullptr;
  return MMF_OK;
}

extern "C" int mmf_shard_info(const mmf_handle* h, int* rank, int* world, int* nccl_version) {
  if (!h) return MMF_ERR_BAD_ARG;
  const ShardState* s = (const ShardState*)h->shard_state;
  if (rank) *rank = s ? s->rank : 0;
  if (world) *world = s ? s->world : 1;
  if (nccl_version) {
    *nccl_version = 0;
    NcclApi* a = (s && s->comm) ? nccl_api() : nullptr;
    if (a && a->GetVersion) a->GetVersion(nccl_version);
  }
  return MMF_OK;
}

// buffers for (n_queries, top_k) candidates per rank; growth is rare (first call / larger batch) and synchronises
static int shard_reserve(mmf_handle* h, ShardState* s, size_t keys, cudaStream_t st) {
  if (keys <= s->cap_keys) return MMF_OK;
  MMF_CUDA_OK(h, cudaStreamSynchronize(st));
  if (s->local) MMF_CUDA_OK(h, cudaFree(s->local));
  if (s->gather) MMF_CUDA_OK(h, cudaFree(s->gather));
  s->local = s->gather = nullptr;
  s->cap_keys = 0;
  const size_t want = keys + keys / 4;
  if (cudaMalloc(&s->local, want * 8) != cudaSuccess || cudaMalloc(&s->gather, want * 8 * (size_t)s->world) != cudaSuccess) {
    cudaGetLastError();
    return mmf_set_error(h, MMF_ERR_NOMEM, "sharded search: cannot allocate the candidate buffers (%zu keys x %d ranks)", want, s->world);
  }
  s->cap_keys = want;
  return MMF_OK;
}

extern "C" int mmf_shard_all_gather(mmf_handle* h, const uint64_t* local_packed, int64_t n_keys, uint64_t* out_gathered,
                                    mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  ShardState* s = (ShardState*)h->shard_state;
  if (!s) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "shard_all_gather: mmf_shard_init has not been called");
  if (n_keys < 0 || (n_keys > 0 && (!local_packed || !out_gathered)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "shard_all_gather: bad argument");
  if (n_keys == 0) return MMF_OK;
  if (s->world == 1) {
    if (out_gathered != local_packed)
      MMF_CUDA_OK(h, cudaMemcpyAsync(out_gathered, local_packed, (size_t)n_keys * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MMF_OK;
  }
  NcclApi* a = nccl_api();
  const int rc = a->AllGather(local_packed, out_gathered, (size_t)n_keys, kNcclUint64, s->comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(h, a, rc, "ncclAllGather");
  h->collectives++;
  return MMF_OK;
}

extern "C" int mmf_vault_search_sharded(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                                        int algo, float* out_scores, int64_t* out_rows, float* out_discrepancy,
                                        mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  ShardState* s = (ShardState*)h->shard_state;
  if (!s) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "vault_search_sharded: mmf_shard_init has not been called");
  if (n_queries < 0 || top_k < 1 || top_k > MMF_MAX_TOP_K || (n_queries > 0 && (!queries || !out_scores || !out_rows)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_sharded: bad argument (n_queries=%lld top_k=%d)", (long long)n_queries, top_k);
  if (n_queries == 0) return MMF_OK;
  if (s->world == 1) return mmf_vault_search(h, queries, n_queries, top_k, threshold, algo, out_scores, out_rows, out_discrepancy, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t keys = (size_t)n_queries * top_k;
  int rc = shard_reserve(h, s, keys, st);
  if (rc != MMF_OK) return rc;
  // every rank sends exactly top_k slots per query (a shard with fewer rows pads with empty keys), so the
  // gathered layout is (world, n_queries, top_k) whatever the shard sizes are
  rc = mmf_search_dispatch_packed(h, queries, n_queries, top_k, algo, s->local, st, "vault_search_sharded");
  if (rc != MMF_OK) return rc;
  rc = mmf_shard_all_gather(h, s->local, (int64_t)keys, s->gather, stream);
  if (rc != MMF_OK) return rc;
  return mmf_topk_merge(h, s->gather, s->world, n_queries, top_k, top_k, threshold, out_scores, out_rows, out_discrepancy, stream);
}
