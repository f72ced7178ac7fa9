// Streaming top-k building blocks shared by the vault search kernels.
//
// Producers never sort.  They keep, per query, an unsorted candidate buffer of packed
// keys (common.cuh: pack_key) and a threshold tau that is a LOWER BOUND on the current
// k-th best score; an element is appended only when !(s < tau) (NaN and ties pass).  When
// a buffer nears capacity one warp compacts it to the exact top-k (warp_compact) and
// raises tau.  A final block-wide radix select + bitonic sort (block_select_topk) turns
// any number of candidate lists into the sorted top-k.  Keys are unique (row id in the
// low word), so the selected set and its order are fully deterministic.
#pragma once
#include "common.cuh"

namespace mmf {

// k-th largest of the warp-distributed keys (KPL per lane, 0 = empty): radix-4 descent on the score
// word (16 rounds, the 3 candidate thresholds of a round are counted independently with ballots, so a
// round costs one ballot latency, not three dependent reductions), then -- only if several candidates
// tie on the score -- a bitwise descent on the row word.
template <int KPL>
__device__ __forceinline__ int warp_count_ge(const u64 (&key)[KPL], u32 cand) {
  int c = 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) c += __popc(__ballot_sync(FULL, (u32)(key[r] >> 32) >= cand));
  return c;
}

template <int KPL>
__device__ __forceinline__ u64 warp_kth_largest(const u64 (&key)[KPL], int k) {
  u32 t_hi = 0;
#pragma unroll 1
  for (int shift = 30; shift >= 0; shift -= 2) {
    const int c1 = warp_count_ge<KPL>(key, t_hi | (1u << shift));
    const int c2 = warp_count_ge<KPL>(key, t_hi | (2u << shift));
    const int c3 = warp_count_ge<KPL>(key, t_hi | (3u << shift));
    t_hi |= (c3 >= k ? 3u : c2 >= k ? 2u : c1 >= k ? 1u : 0u) << shift;
  }
  int gt = 0, eq = 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const u32 hi = (u32)(key[r] >> 32);
    gt += __popc(__ballot_sync(FULL, hi > t_hi));
    eq += __popc(__ballot_sync(FULL, hi == t_hi && key[r] != 0));
  }
  const int need = k - gt;                 // how many of the tied candidates survive
  u32 t_lo = 0;
  if (eq > need && t_hi != 0) {            // warp-uniform
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const u32 cand = t_lo | (1u << bit);
      int c = 0;
#pragma unroll
      for (int r = 0; r < KPL; ++r) c += __popc(__ballot_sync(FULL, (u32)(key[r] >> 32) == t_hi && (u32)key[r] >= cand));
      if (c >= need) t_lo = cand;
    }
  }
  return ((u64)t_hi << 32) | t_lo;
}

// Compacts buf[0..cnt) (shared or global memory, cnt <= 32*KPL) to its top-k in place.
// Whole warp participates.  Returns the new count; *tau_out = score of the k-th best when
// the buffer held >= k candidates (else unchanged).
template <int KPL>
__device__ __forceinline__ int warp_compact(u64* buf, int cnt, int k, float* tau_out) {
  const int lane = threadIdx.x & 31;
  if (cnt <= k) return cnt;
  u64 key[KPL];
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const int i = r * 32 + lane;
    key[r] = (i < cnt) ? buf[i] : 0ull;
  }
  __syncwarp();
  const u64 t = warp_kth_largest<KPL>(key, k);
  int base = 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const bool keep = key[r] >= t && key[r] != 0;
    const u32 m = __ballot_sync(FULL, keep);
    if (keep) buf[base + __popc(m & ((1u << lane) - 1))] = key[r];
    base += __popc(m);
  }
  __syncwarp();
  *tau_out = okey_inv((u32)(t >> 32));
  return base;
}

// Screened search (approximate scores, see vault_mma.cu VAR_SCREEN): like warp_compact, but every candidate
// whose score lies within `margin` of the k-th best survives, because the exact re-scoring may still lift it
// into the top-k.  If more than `limit` candidates would survive, the band does not fit the list: the plain
// top-k is kept and *overflow is set (the caller flags the search for the exact redo).
template <int KPL>
__device__ __forceinline__ int warp_compact_band(u64* buf, int cnt, int k, float margin, int limit, float* tau_out,
                                                 bool* overflow) {
  const int lane = threadIdx.x & 31;
  if (cnt <= k) return cnt;
  u64 key[KPL];
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const int i = r * 32 + lane;
    key[r] = (i < cnt) ? buf[i] : 0ull;
  }
  __syncwarp();
  const u64 t = warp_kth_largest<KPL>(key, k);
  const float ts = okey_inv((u32)(t >> 32));
  u64 cut = (u64)okey(ts - margin) << 32;
  int n = 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) n += __popc(__ballot_sync(FULL, key[r] >= cut && key[r] != 0));
  if (n > limit) { cut = t; *overflow = true; }      // warp-uniform
  int base = 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const bool keep = key[r] >= cut && key[r] != 0;
    const u32 m = __ballot_sync(FULL, keep);
    if (keep) buf[base + __popc(m & ((1u << lane) - 1))] = key[r];
    base += __popc(m);
  }
  __syncwarp();
  *tau_out = ts;
  return base;
}

// ---- block-wide exact selection ---------------------------------------------------------
#define MMF_SELECT_MAX_LISTS 1024

struct SelectSmem {
  u32 hist[256];
  u32 sel_digit, sel_above, n_win, n_staged;
  u64 win[MMF_MAX_TOP_K];
  int counts[MMF_SELECT_MAX_LISTS];
};

// Candidate source: n_lists lists of up to k_in packed keys; list l starts at
// lists + l*list_stride; if counts != nullptr only the first counts[l*count_stride]
// entries of list l are valid; key 0 is always skipped.
struct CandidateLists {
  const u64* lists;
  const int* counts;
  int n_lists;
  int k_in;
  long long list_stride;
  long long count_stride;
  const int* slots = nullptr;   // optional (shared memory, n_lists <= MMF_SELECT_MAX_LISTS): list l lives at
                                // lists + slots[l]*list_stride (counts + slots[l]*count_stride); null = identity
};

// One 8-bit radix-select pass over `n` keys produced by `key_at(i)`: histogram of the digit at
// `pass` among keys matching (mask, prefix), then the digit where the suffix count reaches `need`.
// Returns false when there are fewer than `need` keys in total (only possible on the first pass).
template <typename KeyAt>
__device__ __forceinline__ bool radix_pass(KeyAt key_at, long long n, int pass, u64& prefix, u64& mask, u32& need,
                                           SelectSmem& sm) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < 256; i += nthr) sm.hist[i] = 0;
  __syncthreads();
  for (long long i = tid; i < n; i += nthr) {
    const u64 key = key_at(i);
    if (key != 0 && (key & mask) == prefix) atomicAdd(&sm.hist[(key >> (8 * pass)) & 255], 1u);
  }
  __syncthreads();
  if (tid < 32) {                             // warp 0: find the digit where the suffix count reaches `need`
    u32 loc[8], s = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { loc[b] = sm.hist[tid * 8 + b]; s += loc[b]; }
    u32 incl = s;                             // suffix sum over lanes >= tid
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_down_sync(FULL, incl, o);
      if (tid + o < 32) incl += v;
    }
    const u32 above_lane = incl - s;          // candidates in higher lanes' bins
    const u32 tot = __shfl_sync(FULL, incl, 0);
    if (tot < need) {
      if (tid == 0) sm.sel_digit = 0xFFFFFFFFu;
    } else if (above_lane < need && incl >= need) {     // exactly one lane
      u32 above = above_lane;
      int d = 7;
      for (; d >= 0; --d) {
        if (above + loc[d] >= need) break;
        above += loc[d];
      }
      sm.sel_digit = tid * 8 + d;
      sm.sel_above = above;
    }
  }
  __syncthreads();
  const u32 digit = sm.sel_digit, above = sm.sel_above;
  __syncthreads();
  if (digit == 0xFFFFFFFFu) return false;
  need -= above;
  prefix |= (u64)digit << (8 * pass);
  mask |= 0xFFull << (8 * pass);
  return true;
}

// Selects the top_k largest keys and writes them sorted descending.  All threads of the
// block must call it.  top_k <= MMF_MAX_TOP_K.  Slots beyond the number of valid
// candidates get score NaN / row -1 / key 0.
// `staging` (shared memory, `staging_cap` keys) receives the valid candidates in ONE parallel
// sweep over global memory (counts first, then 32-entry chunks, all loads independent), so
// the 8 radix passes and the winner gather run out of shared memory; if the candidates do
// not fit, the passes fall back to re-reading global memory (correct, slower).
// `min_key`: keys below it are known not to be in the top-k (a published lower bound of the k-th
// best, e.g. g_tau << 32) and are dropped while staging.
__device__ __forceinline__ void block_select_topk(const CandidateLists& src, int top_k, SelectSmem& sm, u64* staging,
                                                  int staging_cap, u64 min_key, float* out_scores, long long* out_rows,
                                                  u64* out_packed, float* out_disc, double threshold) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  const long long total = (long long)src.n_lists * src.k_in;
  const bool small_lists = src.n_lists <= MMF_SELECT_MAX_LISTS;
  if (tid == 0) sm.n_staged = 0;
  if (small_lists)
    for (int l = tid; l < src.n_lists; l += nthr)
      sm.counts[l] = src.counts ? min(src.counts[(src.slots ? src.slots[l] : l) * src.count_stride], src.k_in) : src.k_in;
  __syncthreads();
  bool staged = small_lists;
  if (small_lists) {
    auto stage = [&](u64 key) {
      if (key != 0 && key >= min_key) {
        const u32 pos = atomicAdd(&sm.n_staged, 1u);
        if (pos < (u32)staging_cap) staging[pos] = key;
      }
    };
    if (src.n_lists * 2 >= nthr) {
      // many short lists: one thread per list, 8 independent loads in flight per thread
      for (int l = tid; l < src.n_lists; l += nthr) {
        const u64* lp = src.lists + (src.slots ? src.slots[l] : l) * src.list_stride;
        const int n = sm.counts[l];
        for (int j0 = 0; j0 < n; j0 += 8) {
          u64 key[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) key[e] = (j0 + e < n) ? lp[j0 + e] : 0ull;
#pragma unroll
          for (int e = 0; e < 8; ++e) stage(key[e]);
        }
      }
    } else {
      // few long lists: a warp per 32-entry chunk, coalesced
      const int chunks_per_list = (src.k_in + 31) >> 5;
      const int n_chunks = src.n_lists * chunks_per_list;
      for (int c = tid >> 5; c < n_chunks; c += nthr >> 5) {
        const int l = c / chunks_per_list, j = (c - l * chunks_per_list) * 32 + lane;
        stage(j < sm.counts[l] ? src.lists[(src.slots ? src.slots[l] : l) * src.list_stride + j] : 0ull);
      }
    }
    __syncthreads();
    staged = sm.n_staged <= (u32)staging_cap;
  }
  const long long n_keys = staged ? (long long)sm.n_staged : total;
  auto key_at = [&](long long i) -> u64 {
    if (staged) return staging[i];
    const int l = (int)(i / src.k_in), j = (int)(i % src.k_in);
    const long long sl = src.slots ? src.slots[l] : l;
    if (src.counts && j >= src.counts[sl * src.count_stride]) return 0ull;
    return src.lists[sl * src.list_stride + j];
  };

  u64 prefix = 0, mask = 0;
  u32 need = top_k;
  bool all = false;                           // fewer valid candidates than top_k: take all
  for (int pass = 7; pass >= 0 && !all; --pass) all = !radix_pass(key_at, n_keys, pass, prefix, mask, need, sm);
  const u64 t = all ? 1ull : prefix;          // winners: key >= t (keys unique -> exactly top_k of them)
  // gather winners, pad, bitonic sort descending
  int kp2 = 1;
  while (kp2 < top_k) kp2 <<= 1;
  for (int i = tid; i < kp2; i += nthr) sm.win[i] = 0;
  if (tid == 0) sm.n_win = 0;
  __syncthreads();
  for (long long i = tid; i < n_keys; i += nthr) {
    const u64 key = key_at(i);
    if (key != 0 && key >= t) {
      const u32 pos = atomicAdd(&sm.n_win, 1u);
      if (pos < (u32)top_k) sm.win[pos] = key;
    }
  }
  __syncthreads();
  for (int size = 2; size <= kp2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kp2; i += nthr) {
        const int j = i ^ stride;
        if (j > i) {
          const u64 a = sm.win[i], b = sm.win[j];
          const bool desc = ((i & size) == 0);
          if (desc ? (a < b) : (a > b)) { sm.win[i] = b; sm.win[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < top_k; i += nthr) {
    const u64 key = sm.win[i];
    if (out_packed) out_packed[i] = key;
    if (out_scores) out_scores[i] = key ? okey_inv((u32)(key >> 32)) : __int_as_float(0x7FC00000);
    if (out_rows) out_rows[i] = key ? (long long)(u32)key : -1ll;
  }
  if (tid == 0 && out_disc) {
    const u64 key = sm.win[0];
    *out_disc = key ? discrepancy_rule(okey_inv((u32)(key >> 32)), threshold) : 0.0f;
  }
  __syncthreads();
}


// ---- fast tail (mma_merge_kernel / mma_rerank_kernel <.., FAST>; MMF_MERGE_FAST=0 selects the first form) ----
// The three steps of block_select_topk as separate pieces, so that a kernel can stage once and pick the
// cheapest selection for what was staged.

// Step 1: the staging sweep of block_select_topk.  Returns the number of candidates that passed `min_key`
// (they are in staging[0..n) only if n <= staging_cap).  All threads of the block must call it;
// src.n_lists <= MMF_SELECT_MAX_LISTS.
__device__ __forceinline__ u32 stage_candidates(const CandidateLists& src, SelectSmem& sm, u64* staging, int staging_cap,
                                                u64 min_key) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  if (tid == 0) sm.n_staged = 0;
  for (int l = tid; l < src.n_lists; l += nthr)
    sm.counts[l] = src.counts ? min(src.counts[(src.slots ? src.slots[l] : l) * src.count_stride], src.k_in) : src.k_in;
  __syncthreads();
  auto stage = [&](u64 key) {
    if (key != 0 && key >= min_key) {
      const u32 pos = atomicAdd(&sm.n_staged, 1u);
      if (pos < (u32)staging_cap) staging[pos] = key;
    }
  };
  if (src.n_lists * 2 >= nthr) {
    for (int l = tid; l < src.n_lists; l += nthr) {
      const u64* lp = src.lists + (src.slots ? src.slots[l] : l) * src.list_stride;
      const int n = sm.counts[l];
      for (int j0 = 0; j0 < n; j0 += 8) {
        u64 key[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) key[e] = (j0 + e < n) ? lp[j0 + e] : 0ull;
#pragma unroll
        for (int e = 0; e < 8; ++e) stage(key[e]);
      }
    }
  } else {
    const int chunks_per_list = (src.k_in + 31) >> 5;
    const int n_chunks = src.n_lists * chunks_per_list;
    for (int c = tid >> 5; c < n_chunks; c += nthr >> 5) {
      const int l = c / chunks_per_list, j = (c - l * chunks_per_list) * 32 + lane;
      stage(j < sm.counts[l] ? src.lists[(src.slots ? src.slots[l] : l) * src.list_stride + j] : 0ull);
    }
  }
  __syncthreads();
  const u32 n = sm.n_staged;
  // every thread must have READ the count before anyone goes on: the caller may enter block_select_topk next, whose
  // first statement lets thread 0 reset sm.n_staged (a warp that read 0 here took a different branch than the rest:
  // the lost-candidates bug of the first fast tails, seen whenever more than RANK_SELECT_MAX keys were staged)
  __syncthreads();
  return n;
}

// Step 3: sm.win[0..top_k) (sorted descending, 0 = empty) -> the output arrays, as block_select_topk writes them.
__device__ __forceinline__ void write_topk_outputs(SelectSmem& sm, int top_k, float* out_scores, long long* out_rows,
                                                   u64* out_packed, float* out_disc, double threshold) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < top_k; i += nthr) {
    const u64 key = sm.win[i];
    if (out_packed) out_packed[i] = key;
    if (out_scores) out_scores[i] = key ? okey_inv((u32)(key >> 32)) : __int_as_float(0x7FC00000);
    if (out_rows) out_rows[i] = key ? (long long)(u32)key : -1ll;
  }
  if (tid == 0 && out_disc) {
    const u64 key = sm.win[0];
    *out_disc = key ? discrepancy_rule(okey_inv((u32)(key >> 32)), threshold) : 0.0f;
  }
  __syncthreads();
}

// Step 2 for SHORT inputs: rank by counting.  keys[0..n) in shared memory, unique, 0 = empty.  Every thread
// ranks its keys against all others (n broadcast reads each) and drops the winners straight into their
// sorted slot: two barriers instead of the 8 radix passes + bitonic sort.  sm.win[0..top_k) on return.
__device__ __forceinline__ void block_rank_select(const u64* keys, int n, int top_k, SelectSmem& sm) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < top_k; i += nthr) sm.win[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += nthr) {
    const u64 key = keys[i];
    if (key == 0) continue;
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += keys[j] > key;
    if (rank < top_k) sm.win[rank] = key;
  }
  __syncthreads();
}
constexpr int RANK_SELECT_MAX = 512;      // above this the radix select wins

}  // namespace mmf
