// Row-sharded Truth-Vault search: the candidate exchange over NVLink peer memory.
//
// Every rank searches its own row shard and ends up with its local top-k per query as packed candidates
// ((order-preserving score key << 32) | global row id, topk.cuh).  The global result is the top-k of the
// union, so the ranks must see each other's candidates: 8 B x top_k x queries per rank (C4: 3.3 MB).
// The first implementation all-gathers them with NCCL (vault.py: exchange_candidates) and merges in a
// second kernel.  This file does the same exchange with plain stores into the peers' memory:
//
//   exchange_push_kernel     every rank WRITES its candidates straight into slot [rank] of every peer's
//                            gather buffer (symmetric memory mapped into this process: ordinary global
//                            stores that travel over NVLink / NVSwitch), then the last block to finish
//                            publishes the epoch in flag [rank] of every peer (st.release.sys);
//   exchange_wait_merge_kernel
//                            one block per query: waits (ld.acquire.sys) until all `world` flags of THIS
//                            rank carry the epoch, then selects + sorts the global top-k from the `world`
//                            lists (block_select_topk, the discrepancy rule fused) -- the merge starts the
//                            moment the last peer's candidates land, with no host round trip, no NCCL
//                            kernel and no intermediate copy.
//
// Buffer of one rank (identical layout on all ranks; `bytes_per_rank` as attached):
//   [0, 1024)            flags: u32 flag[2][MMF_XCHG_MAX_WORLD], flag[parity][src] = epoch of the last exchange
//                        whose candidates from rank `src` are complete in gather[parity]
//   [1024, ...)          gather[2][world][n_queries][top_k] u64: parity p starts at 1024 + p * half, half = the
//                        1 KB-aligned half of the rest of the buffer (fixed at attach; a call must fit one half)
// Two parities (epoch & 1): a peer may already push exchange e+1 while this rank still merges exchange e.
// It cannot get to e+2 before this rank has pushed e+1 -- which this rank does after its merge of e, in stream
// order -- so two buffers are enough and no back-signal is needed.
//
// Progress: a push never waits, and every rank enqueues its push before its wait-merge, so the spinning
// blocks of the wait-merge kernel cannot keep a needed kernel off the device -- PROVIDED every rank has its own
// GPU.  Never run two ranks of this protocol on one device at the same time (nothing guarantees that the kernel
// that is waited for gets to run; B200_PROFILING.md reports Xid 109 for that pattern): the single-device checks
// (tools/cabi_selftest, tests) enqueue ALL pushes before ANY merge on one stream, so no kernel ever waits.
//
// Selected with TruthVault(exchange="p2p"); the library-owned NCCL all-gather (shard.cu) is the default exchange.
#include "common.cuh"
#include "topk.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#define MMF_XCHG_HEADER 1024      // MMF_XCHG_MAX_WORLD: common.cuh

int mmf_search_dispatch_packed(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int algo,
                               uint64_t* out_packed, cudaStream_t st, const char* who);

namespace mmf {

struct ExchangePeers {
  unsigned char* base[MMF_XCHG_MAX_WORLD];     // peer r's buffer as mapped in this process
};

struct ExchangeState {
  int rank = 0, world = 0;
  size_t bytes_per_rank = 0;
  ExchangePeers peers;
  u32 epoch = 0;                                // exchanges done so far (all ranks call in the same order)
  u64* local = nullptr;                         // this rank's packed candidates of the current call
  size_t local_bytes = 0;
  u32* done = nullptr;                          // ticket counter of the push kernel (device, self-resetting)
  int64_t pending_queries = 0;                  // > 0 between mmf_vault_search_push and mmf_vault_exchange_merge
  int pending_k = 0;
  size_t pending_gather_off = 0;
};

__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
  u32 v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// n = n_queries * top_k keys of this rank -> slot [rank] of gather[parity] on every rank (own copy included,
// so that the merge reads one layout).  Stores to one peer are contiguous: 8 B per thread, coalesced.
__global__ void __launch_bounds__(256) exchange_push_kernel(const u64* local, long long n, ExchangePeers peers,
                                                            int rank, int world, size_t slot_off, int parity, u32 epoch,
                                                            u32* done) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 key = local[i];
#pragma unroll 1
    for (int r = 0; r < world; ++r) {
      const int peer = (rank + r) % world;                     // start with the own copy, spread the links
      u64* dst = reinterpret_cast<u64*>(peers.base[peer] + slot_off) + i;
      if (dst != local + i) *dst = key;                        // (the candidates may already sit in the own slot)
    }
  }
  __threadfence_system();                                      // this thread's peer stores, system scope
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const u32 ticket = atomicAdd(done, 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // every block's stores are ordered before its ticket (fence + barrier + atomic); the last block publishes
  if (threadIdx.x == 0) *done = 0;                             // self-reset for the next exchange
  __threadfence_system();
  if ((int)threadIdx.x < world) {
    u32* flag = reinterpret_cast<u32*>(peers.base[threadIdx.x]) + parity * MMF_XCHG_MAX_WORLD + rank;
    st_release_sys(flag, epoch);
  }
}

// One block per query: wait for all ranks' candidates of this epoch, then merge the `world` lists.
__global__ void __launch_bounds__(256) exchange_wait_merge_kernel(const unsigned char* own, int world,
                                                                  size_t gather_off, int parity, u32 epoch,
                                                                  long long n_queries, int k_in, int top_k,
                                                                  double threshold, float* out_scores,
                                                                  long long* out_rows, float* out_disc) {
  __shared__ SelectSmem sel;
  __shared__ u64 staging[4096];
  if ((int)threadIdx.x < world) {
    const u32* flag = reinterpret_cast<const u32*>(own) + parity * MMF_XCHG_MAX_WORLD + threadIdx.x;
    // epochs only grow; (int) difference tolerates the 32-bit wrap
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
  const long long qg = blockIdx.x;
  CandidateLists src;
  src.lists = reinterpret_cast<const u64*>(own + gather_off) + qg * k_in;
  src.counts = nullptr;
  src.n_lists = world;
  src.k_in = k_in;
  src.list_stride = n_queries * k_in;
  src.count_stride = 0;
  block_select_topk(src, top_k, sel, staging, 4096, 0ull, out_scores ? out_scores + qg * top_k : nullptr,
                    out_rows ? out_rows + qg * top_k : nullptr, nullptr, out_disc ? out_disc + qg : nullptr, threshold);
}

}  // namespace mmf

using namespace mmf;

// Byte offsets inside a rank's buffer for an exchange of (n_queries, k_in) candidates per rank.
// Host-only, also exported for the CPU test-suite.
extern "C" int mmf_exchange_layout(int world, int64_t n_queries, int k_in, int64_t* gather_bytes_per_parity,
                                   int64_t* bytes_needed) {
  if (world < 1 || world > MMF_XCHG_MAX_WORLD || n_queries < 0 || k_in < 1) return MMF_ERR_BAD_ARG;
  const int64_t per_parity = (((int64_t)world * n_queries * k_in * 8) + 1023) / 1024 * 1024;
  if (gather_bytes_per_parity) *gather_bytes_per_parity = per_parity;
  if (bytes_needed) *bytes_needed = MMF_XCHG_HEADER + 2 * per_parity;
  return MMF_OK;
}

extern "C" int mmf_exchange_attach(mmf_handle* h, int rank, int world, const uint64_t* peer_ptrs, int64_t bytes_per_rank) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (world < 1 || world > MMF_XCHG_MAX_WORLD || rank < 0 || rank >= world || !peer_ptrs || bytes_per_rank < MMF_XCHG_HEADER)
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "exchange_attach: bad argument (rank=%d world=%d bytes=%lld)", rank, world,
                         (long long)bytes_per_rank);
  for (int r = 0; r < world; ++r)
    if (!peer_ptrs[r]) return mmf_set_error(h, MMF_ERR_BAD_ARG, "exchange_attach: null pointer for rank %d", r);
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  ExchangeState* x = (ExchangeState*)h->xchg_state;
  if (!x) {
    x = new (std::nothrow) ExchangeState();
    if (!x) return mmf_set_error(h, MMF_ERR_NOMEM, "exchange_attach: out of host memory");
    h->xchg_state = x;
    MMF_CUDA_OK(h, cudaMalloc(&x->done, 256));
    MMF_CUDA_OK(h, cudaMemset(x->done, 0, 256));
  }
  x->rank = rank;
  x->world = world;
  x->bytes_per_rank = (size_t)bytes_per_rank;
  x->epoch = 0;
  x->pending_queries = 0;
  // this rank's candidates of one exchange can never exceed one slot of the attached buffer: size the local
  // buffer for that now, so that the (asynchronous) search path never allocates or synchronises
  const size_t local_cap = ((size_t)bytes_per_rank - MMF_XCHG_HEADER) / 2 / (size_t)world + 1024;
  if (local_cap > x->local_bytes) {
    MMF_CUDA_OK(h, cudaDeviceSynchronize());
    if (x->local) MMF_CUDA_OK(h, cudaFree(x->local));
    x->local = nullptr;
    x->local_bytes = 0;
    MMF_CUDA_OK(h, cudaMalloc(&x->local, local_cap));
    x->local_bytes = local_cap;
  }
  memset(&x->peers, 0, sizeof x->peers);
  for (int r = 0; r < world; ++r) x->peers.base[r] = reinterpret_cast<unsigned char*>((uintptr_t)peer_ptrs[r]);
  // the flags of THIS rank start at 0 (= no exchange seen); the caller barriers after attaching
  MMF_CUDA_OK(h, cudaMemset(x->peers.base[rank], 0, MMF_XCHG_HEADER));
  MMF_CUDA_OK(h, cudaDeviceSynchronize());
  return MMF_OK;
}

extern "C" int mmf_exchange_detach(mmf_handle* h) {
  if (!h || !h->xchg_state) return MMF_OK;
  ExchangeState* x = (ExchangeState*)h->xchg_state;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (x->local) cudaFree(x->local);
  if (x->done) cudaFree(x->done);
  delete x;
  h->xchg_state = nullptr;
  return MMF_OK;
}

// Phase 1 of an exchange: local search of this rank's shard + push of its candidates into every rank's gather
// buffer + publication of this rank's epoch flag.  Never waits for anybody.  Asynchronous on `stream`.
extern "C" int mmf_vault_search_push(mmf_handle* h, const float* queries, int64_t n_queries, int k_local, int algo,
                                     mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  ExchangeState* x = (ExchangeState*)h->xchg_state;
  if (!x) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "vault_search_push: no peer buffers attached");
  if (n_queries <= 0 || k_local < 1 || k_local > MMF_MAX_TOP_K || !queries)
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_push: bad argument (n_queries=%lld k_local=%d)",
                         (long long)n_queries, k_local);
  if (x->pending_queries != 0)
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_push: the previous exchange has not been merged yet");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t per_parity = 0, need = 0;
  mmf_exchange_layout(x->world, n_queries, k_local, &per_parity, &need);
  // the two parities live at FIXED offsets (the halves of the attached buffer), whatever the batch size of a call:
  // a peer may already push exchange e+1 while this rank still merges e, so e+1 must never land on bytes that the
  // layout of e put into gather[parity of e] -- which a per-call offset would allow when the batch size changes
  const size_t half = ((x->bytes_per_rank - MMF_XCHG_HEADER) / 2) & ~(size_t)1023;
  if ((size_t)need > x->bytes_per_rank || (size_t)per_parity > half)
    return mmf_set_error(h, MMF_ERR_NOMEM, "vault_search_push: %lld bytes of peer buffer needed, %zu attached",
                         (long long)need, x->bytes_per_rank);
  const size_t local_bytes = (size_t)n_queries * k_local * 8;
  if (local_bytes > x->local_bytes)      // cannot happen after the size check above (attach sized it for one slot)
    return mmf_set_error(h, MMF_ERR_NOMEM, "vault_search_push: local candidate buffer too small");
  const u32 epoch = x->epoch + 1;
  const int parity = (int)(epoch & 1u);
  const size_t gather_off = MMF_XCHG_HEADER + (size_t)parity * half;
  const size_t slot_off = gather_off + (size_t)x->rank * local_bytes;
  // local search of this rank's shard -> packed candidates with GLOBAL row ids.
  // Option "fused_push" (default on): the search writes them into this rank's own slot and, where its merge
  // tail supports it (tcgen05 search, top_k > 16), pushes them to the peers and publishes the flags itself.
  const bool want_fused = h->opt.fused_push != 0 && n_queries <= 65536;
  mmf_push_ctx ctx;
  memset(&ctx, 0, sizeof ctx);
  for (int r = 0; r < x->world; ++r) ctx.base[r] = x->peers.base[r];
  ctx.rank = x->rank; ctx.world = x->world; ctx.parity = parity; ctx.epoch = epoch; ctx.slot_off = slot_off; ctx.done = x->done;
  u64* packed = want_fused ? reinterpret_cast<u64*>(x->peers.base[x->rank] + slot_off) : x->local;
  h->push_fused = false;
  h->push_ctx = want_fused ? &ctx : nullptr;
  int rc = mmf_search_dispatch_packed(h, queries, n_queries, k_local, algo, (uint64_t*)packed, st, "vault_search_push");
  h->push_ctx = nullptr;
  if (rc != MMF_OK) return rc;
  if (!h->push_fused) {
    const long long n = (long long)n_queries * k_local;
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
    exchange_push_kernel<<<blocks, 256, 0, st>>>(packed, n, x->peers, x->rank, x->world, slot_off, parity, epoch, x->done);
    MMF_LAUNCH_OK(h);
  }
  x->epoch = epoch;
  x->pending_queries = n_queries;
  x->pending_k = k_local;
  x->pending_gather_off = gather_off;
  return MMF_OK;
}

// Phase 2: wait (on the device) until every rank's candidates of the pending exchange have landed, merge them.
extern "C" int mmf_vault_exchange_merge(mmf_handle* h, int top_k, double threshold, float* out_scores, int64_t* out_rows,
                                        float* out_discrepancy, mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  ExchangeState* x = (ExchangeState*)h->xchg_state;
  if (!x) return mmf_set_error(h, MMF_ERR_NOT_LOADED, "vault_exchange_merge: no peer buffers attached");
  if (x->pending_queries == 0) return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_exchange_merge: no exchange pending");
  if (top_k < x->pending_k || top_k > MMF_MAX_TOP_K || !out_scores || !out_rows)
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_exchange_merge: bad argument (top_k=%d, %d candidates per rank)", top_k,
                         x->pending_k);
  const int64_t n_queries = x->pending_queries;
  exchange_wait_merge_kernel<<<(unsigned)n_queries, 256, 0, (cudaStream_t)stream>>>(
      x->peers.base[x->rank], x->world, x->pending_gather_off, (int)(x->epoch & 1u), x->epoch, (long long)n_queries, x->pending_k,
      top_k, threshold, out_scores, (long long*)out_rows, out_discrepancy);
  x->pending_queries = 0;
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

extern "C" int mmf_vault_search_exchange(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, int k_local,
                                         double threshold, int algo, float* out_scores, int64_t* out_rows,
                                         float* out_discrepancy, mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_queries < 0 || top_k < 1 || top_k > MMF_MAX_TOP_K || k_local < 1 || k_local > top_k ||
      (n_queries > 0 && (!queries || !out_scores || !out_rows)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_search_exchange: bad argument (n_queries=%lld top_k=%d k_local=%d)",
                         (long long)n_queries, top_k, k_local);
  if (n_queries == 0) return MMF_OK;
  const int rc = mmf_vault_search_push(h, queries, n_queries, k_local, algo, stream);
  if (rc != MMF_OK) return rc;
  return mmf_vault_exchange_merge(h, top_k, threshold, out_scores, out_rows, out_discrepancy, stream);
}
