// K2: HBM-streaming Truth-Vault search for small query batches (batch-1 latency mode),
// fused with the top-k so the similarity vector never reaches HBM.
// Replaces misinfo_forensics.py:446 (similarities = Vn @ q) and :449-450 (argsort top-k).
//
// Roofline: HBM.  Algorithmic bytes per query = n_rows * 512 * elem (2.048 GB at 1M rows
// fp32-exact, 1.024 GB bf16); arithmetic is ~1 FMA per 4 bytes, far below the ridge.
//
// Shape: grid = (row ranges, query chunks of QC).  A warp owns whole rows (2 KB, four
// coalesced 512 B requests per lane-quad), keeps the QC normalised queries in registers
// and reduces with xor shuffles; per-block candidate buffers + thresholds live in shared
// memory (topk.cuh), thresholds are shared grid-wide through one atomicMax word per
// query, and the LAST block to finish a query chunk merges the per-block top-k lists, so
// one launch produces the final sorted result.
#include "common.cuh"
#include "topk.cuh"

#include <algorithm>
#include <cstdlib>

namespace mmf {

struct StreamParams {
  const void* vault;
  long long n_rows;
  u32 row_base;
  const float* qn;        // (n_queries, 512) normalised queries
  int n_queries;
  int top_k;
  long long rows_per_cta; // multiple of 32
  u32* g_tau;             // (n_queries) shared threshold keys, zeroed by the prep kernel
  u64* part_keys;         // (gridDim.x, n_queries, top_k)
  int* part_cnt;          // (gridDim.x, n_queries)
  u32* done;              // (gridDim.y) arrival counters, self-resetting
  float* out_scores;
  long long* out_rows;
  u64* out_packed;
  float* out_disc;
  double threshold;
};
// One warp per query: q / ||q|| (misinfo_forensics.py:439) + reset of the shared threshold.
__global__ void __launch_bounds__(256) query_prep_kernel(const float* __restrict__ q, int n_queries,
                                                         float* __restrict__ qn, u32* __restrict__ g_tau) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_queries) return;
  float v[MMF_DIM / 32], ss = 0.f;
#pragma unroll
  for (int j = 0; j < MMF_DIM / 32; ++j) { v[j] = q[(long long)w * MMF_DIM + j * 32 + lane]; ss = fmaf(v[j], v[j], ss); }
  const float norm = sqrtf(warp_sum(ss));
#pragma unroll
  for (int j = 0; j < MMF_DIM / 32; ++j) qn[(long long)w * MMF_DIM + j * 32 + lane] = v[j] / norm;
  if (lane == 0) g_tau[w] = 0;
}

__device__ __forceinline__ void unpack8(const uint4& w, float (&f)[8], bool bf16) {
  const u32 x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bf16) {
      f[2 * i] = __uint_as_float(x[i] << 16);
      f[2 * i + 1] = __uint_as_float(x[i] & 0xFFFF0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&x[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

// (A screened variant of this kernel -- hi planes only + exact re-scoring of a band, like the tcgen05 search --
// was measured on a B200 in round 2 and deleted: 0.64 ms against 0.37 ms per query.  With one query per block the
// kernel is bound by bytes in flight per SM, not by bytes, and the band bookkeeping cost more than the lo planes.)
template <int QC, bool BF16, int KPL>
__global__ void __launch_bounds__(256) vault_stream_topk_kernel(const StreamParams p) {
  constexpr int U = (QC <= 2) ? 4 : 2; // rows per warp in flight at once: 16 (8) independent 128-bit loads per lane
  constexpr int SUB = 4 / U;           // sub-rounds per 32-row interval (8 warps x U rows each)
  constexpr int C = 32 * KPL;          // candidate capacity per query
  constexpr int LIMIT = C - 64;        // compaction trigger (two 32-row intervals of slack)
  constexpr int NLD = BF16 ? 2 : 4;    // 128-bit loads per lane per row
  constexpr int POOL = (QC * C > 2048) ? QC * C : 2048;   // candidate buffers, later the merge's staging area
  __shared__ u64 pool[POOL];
  u64 (*buf)[C] = reinterpret_cast<u64 (*)[C]>(pool);
  __shared__ int cnt[QC];
  __shared__ volatile float tau[QC];
  __shared__ u32 tau_key[QC];
  __shared__ int s_last;
  __shared__ SelectSmem sel;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.y * QC;
  const int nq = min(QC, p.n_queries - q0);
  const int k = p.top_k;

  if (tid < QC) { cnt[tid] = 0; tau[tid] = -INFINITY; tau_key[tid] = 0; }

  // lane owns elements 8*lane..+7 and 256+8*lane..+7 of every row
  float q[QC][16];
#pragma unroll
  for (int qi = 0; qi < QC; ++qi) {
#pragma unroll
    for (int e = 0; e < 16; ++e)
      q[qi][e] = (qi < nq) ? p.qn[(long long)(q0 + qi) * MMF_DIM + (e >> 3) * 256 + lane * 8 + (e & 7)] : 0.f;
  }
  __syncthreads();

  const long long row_begin = (long long)blockIdx.x * p.rows_per_cta;
  const long long row_end = min(p.n_rows, row_begin + p.rows_per_cta);
  const uint4* vault = reinterpret_cast<const uint4*>(p.vault);
  constexpr int ROW_U4 = BF16 ? 64 : 128;   // uint4 per stored row

  for (long long base = row_begin; base < row_end; base += 32) {
    u32 g_seen = 0;
    if (warp < nq && lane == 0) g_seen = *reinterpret_cast<volatile u32*>(p.g_tau + q0 + warp);
#pragma unroll
    for (int sub = 0; sub < SUB; ++sub) {
      const long long r0 = base + sub * (8 * U) + warp * U;
      uint4 ld[U][NLD];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r0 + u < row_end) {
          const uint4* rp = vault + (r0 + u) * ROW_U4;
#pragma unroll
          for (int c = 0; c < NLD; ++c) ld[u][c] = ldg_stream(rp + c * 32 + lane);
        } else {
#pragma unroll
          for (int c = 0; c < NLD; ++c) ld[u][c] = make_uint4(0, 0, 0, 0);
        }
      }
      float acc[U][QC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float f[8];
          unpack8(ld[u][c], f, BF16);
          if (!BF16) {
            float g[8];
            unpack8(ld[u][c + 2], g, false);   // lo plane
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += g[e];   // exact: hi + lo fits 24 bits
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) v[c * 8 + e] = f[e];
        }
#pragma unroll
        for (int qi = 0; qi < QC; ++qi) {
          float a = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) a = fmaf(v[e], q[qi][e], a);
          acc[u][qi] = a;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int qi = 0; qi < QC; ++qi) acc[u][qi] = warp_sum(acc[u][qi]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (r0 + u < row_end) {
#pragma unroll
            for (int qi = 0; qi < QC; ++qi) {
              if (qi < nq) {
                const float s = BF16 ? acc[u][qi] : acc[u][qi] * MMF_SPLIT_INV_SCALE;
                if (!(s < tau[qi])) {
                  const int pos = atomicAdd(&cnt[qi], 1);
                  buf[qi][pos] = pack_key(s, p.row_base + (u32)(r0 + u));
                }
              }
            }
          }
        }
      }
    }
    // adopt a better threshold found by another block (any block's k-th best is a valid lower bound)
    if (warp < nq && lane == 0 && g_seen > tau_key[warp]) { tau_key[warp] = g_seen; tau[warp] = okey_inv(g_seen); }
    // the predicate may miss appends still in flight; LIMIT leaves a second interval of slack
    const int over = __syncthreads_or(tid < nq && cnt[tid] > LIMIT);
    if (over) {
      if (warp < nq && cnt[warp] > k) {
        float t = 0.f;
        const int c = warp_compact<KPL>(buf[warp], cnt[warp], k, &t);
        if (lane == 0) {
          cnt[warp] = c;
          const u32 tk = okey(t);
          if (tk > tau_key[warp]) { tau_key[warp] = tk; tau[warp] = t; atomicMax(p.g_tau + q0 + warp, tk); }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (warp < nq) {
    float t;
    const int c = warp_compact<KPL>(buf[warp], cnt[warp], k, &t);
    u64* dst = p.part_keys + ((long long)blockIdx.x * p.n_queries + q0 + warp) * k;
    for (int i = lane; i < c; i += 32) dst[i] = buf[warp][i];
    if (lane == 0) p.part_cnt[(long long)blockIdx.x * p.n_queries + q0 + warp] = c;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const u32 ticket = atomicAdd(p.done + blockIdx.y, 1u);
    s_last = (ticket == gridDim.x - 1);
    if (s_last) p.done[blockIdx.y] = 0;     // self-reset for the next search
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int qi = 0; qi < nq; ++qi) {
    const long long qg = q0 + qi;
    CandidateLists src;
    src.lists = p.part_keys + qg * k;
    src.counts = p.part_cnt + qg;
    src.n_lists = gridDim.x;
    src.k_in = k;
    src.list_stride = (long long)p.n_queries * k;
    src.count_stride = p.n_queries;
    block_select_topk(src, k, sel, pool, POOL, (u64)(*reinterpret_cast<volatile u32*>(p.g_tau + qg)) << 32, p.out_scores ? p.out_scores + qg * k : nullptr,
                      p.out_rows ? p.out_rows + qg * k : nullptr, p.out_packed ? p.out_packed + qg * k : nullptr,
                      p.out_disc ? p.out_disc + qg : nullptr, p.threshold);
  }
}

// ---- TMA-staged form of the same kernel ------------------------------------------------------------------
// The vault rows of a block are one contiguous byte range, so a producer thread streams them through a shared-memory
// ring with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx: UBLKCP in SASS), 64 KB = 32 fp32-exact rows per
// stage, 3 stages: 128 KB in flight per SM while one stage is read -- against the 4 rows x 2 KB per warp that
// registers allow above.  (First cut: 10 stages of 16 KB, one row per warp and stage -- 0.525 ms per query against
// 0.367: every warp paid a barrier wait, a shuffle reduction and a hand-off per ROW, in lock-step with the others.)
// tools/tma_stream_micro.cu: this loader alone streams whole rows at 7.3 TB/s (the register kernel: 5.5 TB/s).
// 8 consumer warps take the rows of a stage (warp w: rows w*RPW .., 4 at a time), read their 16 elements per lane from shared
// memory (conflict-free 128-bit reads) and do EXACTLY the arithmetic of the kernel above, in the same order, so the
// two kernels return bit-identical scores; candidate buffers, thresholds, compaction and the last-block merge are
// the same code.  Consumers never meet at a block barrier in the steady state: a named barrier every CHECK stages
// lets them agree on a compaction (rare).
constexpr int TMA_STAGE_BYTES = 65536, TMA_STAGES = 3, TMA_THREADS = 288;

__device__ __forceinline__ u32 st_smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(u64* bar, u32 parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "SWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SDONE_%=;\n\t"
      "bra SWAIT_%=;\n\t"
      "SDONE_%=:\n\t}" ::"r"(st_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_bulk_load(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(st_smem_u32(dst)), "l"(src), "r"(bytes), "r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int QC, bool BF16, int KPL>
__global__ void __launch_bounds__(TMA_THREADS, 1) vault_stream_tma_kernel(const StreamParams p) {
  constexpr int ROW_BYTES = BF16 ? 1024 : 2048;
  constexpr int SROWS = TMA_STAGE_BYTES / ROW_BYTES;     // rows per stage: 32 (fp32-exact) / 64 (bf16)
  constexpr int RPW = SROWS / 8;                         // rows per consumer warp per stage: 4 / 8
  constexpr int RCH = QC >= 8 ? 2 : 4;                   // ... processed 4 (2) at a time: 16 (8) independent 128-bit reads per lane
  constexpr int CHECK = 1;                               // stages between compaction checks (<= 64 appends per query in between)
  constexpr int C = 32 * KPL;
  constexpr int LIMIT = C - 64;
  constexpr int POOL = (QC * C > 2048) ? QC * C : 2048;
  extern __shared__ __align__(128) unsigned char ring_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ring_raw) + 127) & ~(uintptr_t)127);
  __shared__ u64 pool[POOL];
  u64 (*buf)[C] = reinterpret_cast<u64 (*)[C]>(pool);
  __shared__ int cnt[QC];
  __shared__ volatile float tau[QC];
  __shared__ u32 tau_key[QC];
  __shared__ int s_last;
  __shared__ volatile int s_over[3];
  __shared__ SelectSmem sel;
  __shared__ __align__(8) u64 full_bar[TMA_STAGES], empty_bar[TMA_STAGES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.y * QC;
  const int nq = min(QC, p.n_queries - q0);
  const int k = p.top_k;
  if (tid < QC) { cnt[tid] = 0; tau[tid] = -INFINITY; tau_key[tid] = 0; }
  if (tid < 3) s_over[tid] = 0;
  if (tid == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) { st_mbar_init(full_bar + s, 1); st_mbar_init(empty_bar + s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long row_begin = (long long)blockIdx.x * p.rows_per_cta;
  const long long row_end = min(p.n_rows, row_begin + p.rows_per_cta);
  const long long n_stages = row_end > row_begin ? (row_end - row_begin + SROWS - 1) / SROWS : 0;
  const unsigned char* vault = reinterpret_cast<const unsigned char*>(p.vault);

  if (warp == 8) {
    // ===== producer =====
    if (lane == 0) {
      for (long long it = 0; it < n_stages; ++it) {
        const int s = (int)(it % TMA_STAGES);
        st_mbar_wait(empty_bar + s, (u32)(((it / TMA_STAGES) & 1) ^ 1));
        const long long row = row_begin + it * SROWS;
        const u32 bytes = (u32)(min((long long)SROWS, row_end - row) * ROW_BYTES);
        st_mbar_expect_tx(full_bar + s, bytes);
        st_bulk_load(ring + (size_t)s * TMA_STAGE_BYTES, vault + row * ROW_BYTES, bytes, full_bar + s);
      }
    }
  } else {
    // ===== consumers: lane owns elements 8*lane..+7 and 256+8*lane..+7 of every row (as the register kernel) =====
    float q[QC][16];
#pragma unroll
    for (int qi = 0; qi < QC; ++qi) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        q[qi][e] = (qi < nq) ? p.qn[(long long)(q0 + qi) * MMF_DIM + (e >> 3) * 256 + lane * 8 + (e & 7)] : 0.f;
    }
    for (long long it = 0; it < n_stages; ++it) {
      const int s = (int)(it % TMA_STAGES);
      u32 g_seen = 0;
      if (it % CHECK == 0 && warp < nq && lane == 0) g_seen = *reinterpret_cast<volatile u32*>(p.g_tau + q0 + warp);
      st_mbar_wait(full_bar + s, (u32)((it / TMA_STAGES) & 1));
      const uint4* stage = reinterpret_cast<const uint4*>(ring + (size_t)s * TMA_STAGE_BYTES);
#pragma unroll 1
      for (int r0 = 0; r0 < RPW; r0 += RCH) {
        float acc[RCH][QC];
#pragma unroll
        for (int r = 0; r < RCH; ++r) {
          const uint4* rp = stage + (warp * RPW + r0 + r) * (ROW_BYTES / 16);
          float v[16];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float f[8];
            unpack8(rp[c * 32 + lane], f, BF16);
            if (!BF16) {
              float g[8];
              unpack8(rp[(c + 2) * 32 + lane], g, false);   // lo plane
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += g[e];   // exact: hi + lo fits 24 bits
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[c * 8 + e] = f[e];
          }
#pragma unroll
          for (int qi = 0; qi < QC; ++qi) {
            float a = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) a = fmaf(v[e], q[qi][e], a);
            acc[r][qi] = a;
          }
        }
        if (r0 + RCH == RPW) {                              // last rows of this warp read: the slot may be refilled
          __syncwarp();
          if (lane == 0) st_mbar_arrive(empty_bar + s);
        }
#pragma unroll
        for (int r = 0; r < RCH; ++r)
#pragma unroll
          for (int qi = 0; qi < QC; ++qi) acc[r][qi] = warp_sum(acc[r][qi]);
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < RCH; ++r) {
            const long long row = row_begin + it * SROWS + warp * RPW + r0 + r;
            if (row < row_end) {
#pragma unroll
              for (int qi = 0; qi < QC; ++qi) {
                if (qi < nq) {
                  const float sc = BF16 ? acc[r][qi] : acc[r][qi] * MMF_SPLIT_INV_SCALE;
                  if (!(sc < tau[qi])) {
                    const int pos = atomicAdd(&cnt[qi], 1);
                    buf[qi][pos] = pack_key(sc, p.row_base + (u32)row);
                  }
                }
              }
            }
          }
        }
      }
      if (it % CHECK == CHECK - 1 || it + 1 == n_stages) {
        // adopt a better threshold found by another block; agree on a compaction (three rotating flags: the one
        // written at check c is cleared after the barrier of check c+1, when nobody can still be reading it)
        const int chk = (int)((it / CHECK) % 3);
        if (warp < nq && lane == 0 && g_seen > tau_key[warp]) { tau_key[warp] = g_seen; tau[warp] = okey_inv(g_seen); }
        if (tid < nq && cnt[tid] > LIMIT) s_over[chk] = 1;
        consumer_barrier();
        const int over = s_over[chk];
        if (tid == 0) s_over[(chk + 2) % 3] = 0;
        if (over) {
          if (warp < nq && cnt[warp] > k) {
            float t = 0.f;
            const int c = warp_compact<KPL>(buf[warp], cnt[warp], k, &t);
            if (lane == 0) {
              cnt[warp] = c;
              const u32 tk = okey(t);
              if (tk > tau_key[warp]) { tau_key[warp] = tk; tau[warp] = t; atomicMax(p.g_tau + q0 + warp, tk); }
            }
          }
          consumer_barrier();
        }
      }
    }
  }
  __syncthreads();
  if (warp < nq) {
    float t;
    const int c = warp_compact<KPL>(buf[warp], cnt[warp], k, &t);
    u64* dst = p.part_keys + ((long long)blockIdx.x * p.n_queries + q0 + warp) * k;
    for (int i = lane; i < c; i += 32) dst[i] = buf[warp][i];
    if (lane == 0) p.part_cnt[(long long)blockIdx.x * p.n_queries + q0 + warp] = c;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const u32 ticket = atomicAdd(p.done + blockIdx.y, 1u);
    s_last = (ticket == gridDim.x - 1);
    if (s_last) p.done[blockIdx.y] = 0;     // self-reset for the next search
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int qi = 0; qi < nq; ++qi) {
    const long long qg = q0 + qi;
    CandidateLists src;
    src.lists = p.part_keys + qg * k;
    src.counts = p.part_cnt + qg;
    src.n_lists = gridDim.x;
    src.k_in = k;
    src.list_stride = (long long)p.n_queries * k;
    src.count_stride = p.n_queries;
    block_select_topk(src, k, sel, pool, POOL, (u64)(*reinterpret_cast<volatile u32*>(p.g_tau + qg)) << 32, p.out_scores ? p.out_scores + qg * k : nullptr,
                      p.out_rows ? p.out_rows + qg * k : nullptr, p.out_packed ? p.out_packed + qg * k : nullptr,
                      p.out_disc ? p.out_disc + qg : nullptr, p.threshold);
  }
}

// (n_lists, n_queries, k_in) packed candidates -> per-query sorted top-k.  One block per query.
__global__ void __launch_bounds__(256) topk_merge_kernel(const u64* __restrict__ packed, int n_lists,
                                                         long long n_queries, int k_in, int top_k, double threshold,
                                                         float* out_scores, long long* out_rows, u64* out_packed,
                                                         float* out_disc) {
  __shared__ SelectSmem sel;
  __shared__ u64 staging[4096];
  const long long qg = blockIdx.x;
  // Fast path, the sharded search's case: every list is sorted descending with its empty slots (key 0) at the end --
  // what every search entry of this library writes.  Then the global rank of element j of list l is j plus, for every
  // other list, the number of its keys that beat this one: a binary search each (ties go to the lower list index, so
  // ranks are unique whatever the caller passes), and a winner drops straight into its sorted slot.  8 lists x 100:
  // ~50 shared-memory reads per key that can still win instead of 8 radix passes + a bitonic sort.
  // The lists are checked while they are staged; anything else takes the general path below.
  const int n_total = n_lists * k_in;
  if (n_lists > 1 && n_total <= 4096) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < n_total; i += nthr) {
      const int l = i / k_in, j = i - l * k_in;
      staging[i] = packed[(long long)l * n_queries * k_in + qg * k_in + j];
    }
    for (int i = tid; i < top_k; i += nthr) sel.win[i] = 0;
    __syncthreads();
    int bad = 0;
    for (int i = tid; i < n_total; i += nthr) {
      const int j = i % k_in;
      if (j > 0 && staging[i - 1] < staging[i]) bad = 1;
    }
    if (!__syncthreads_or(bad)) {
      // a cheap lower bound of the top_k-th best first: every list holds at least `per` keys >= the smallest of the
      // lists' per-th entries, per * n_lists >= top_k of them in all -- nothing below it needs a rank (8 x 100 -> ~150)
      const int per = (top_k + n_lists - 1) / n_lists;
      u64 floor_key = ~0ull;
      if (per <= k_in) { for (int o = 0; o < n_lists; ++o) floor_key = min(floor_key, staging[o * k_in + per - 1]); }
      else floor_key = 0ull;
      for (int i = tid; i < n_total; i += nthr) {
        const u64 key = staging[i];
        if (key == 0 || key < floor_key) continue;
        const int l = i / k_in;
        int rank = i - l * k_in;
        for (int o = 0; o < n_lists; ++o) {
          if (o == l) continue;
          const u64* lst = staging + o * k_in;
          int lo = 0, hi = k_in;                       // first position whose key does not beat `key`
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const u64 x = lst[mid];
            if (x > key || (x == key && o < l)) lo = mid + 1; else hi = mid;
          }
          rank += lo;
        }
        if (rank < top_k) sel.win[rank] = key;
      }
      __syncthreads();
      write_topk_outputs(sel, top_k, out_scores ? out_scores + qg * top_k : nullptr, out_rows ? out_rows + qg * top_k : nullptr,
                         out_packed ? out_packed + qg * top_k : nullptr, out_disc ? out_disc + qg : nullptr, threshold);
      return;
    }
  }
  CandidateLists src;
  src.lists = packed + qg * k_in;
  src.counts = nullptr;
  src.n_lists = n_lists;
  src.k_in = k_in;
  src.list_stride = n_queries * k_in;
  src.count_stride = 0;
  block_select_topk(src, top_k, sel, staging, 4096, 0ull, out_scores ? out_scores + qg * top_k : nullptr,
                    out_rows ? out_rows + qg * top_k : nullptr, out_packed ? out_packed + qg * top_k : nullptr,
                    out_disc ? out_disc + qg : nullptr, threshold);
}

__global__ void fill_empty_kernel(long long n_queries, int top_k, float* out_scores, long long* out_rows,
                                  u64* out_packed, float* out_disc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_queries * top_k) {
    if (out_scores) out_scores[i] = __int_as_float(0x7FC00000);
    if (out_rows) out_rows[i] = -1;
    if (out_packed) out_packed[i] = 0;
  }
  if (i < n_queries && out_disc) out_disc[i] = 0.f;
}

}  // namespace mmf

using namespace mmf;

template <int QC, bool BF16>
static void launch_stream_k(int kpl, dim3 grid, cudaStream_t st, const StreamParams& p) {
  if (kpl == 4) vault_stream_topk_kernel<QC, BF16, 4><<<grid, 256, 0, st>>>(p);
  else if (kpl == 8) vault_stream_topk_kernel<QC, BF16, 8><<<grid, 256, 0, st>>>(p);
  else vault_stream_topk_kernel<QC, BF16, 16><<<grid, 256, 0, st>>>(p);
}

template <int QC, bool BF16>
static int launch_stream_tma_k(mmf_handle* h, int kpl, dim3 grid, cudaStream_t st, const StreamParams& p) {
  const int smem = TMA_STAGES * TMA_STAGE_BYTES + 256;
#define MMF_TMA_CASE(KPL_)                                                                                              \
  do {                                                                                                                  \
    auto kern = vault_stream_tma_kernel<QC, BF16, KPL_>;                                                                \
    static unsigned long long attr_set = 0;                                                                             \
    if (h->device >= 64 || !((attr_set >> h->device) & 1ull)) {                                                         \
      MMF_CUDA_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                    \
      if (h->device < 64) attr_set |= 1ull << h->device;                                                                \
    }                                                                                                                   \
    kern<<<grid, TMA_THREADS, smem, st>>>(p);                                                                           \
  } while (0)
  if (kpl == 4) MMF_TMA_CASE(4);
  else if (kpl == 8) MMF_TMA_CASE(8);
  else MMF_TMA_CASE(16);
#undef MMF_TMA_CASE
  return MMF_OK;
}

template <bool BF16>
static int launch_stream_tma_q(mmf_handle* h, int qc, int kpl, dim3 grid, cudaStream_t st, const StreamParams& p) {
  if (qc == 1) return launch_stream_tma_k<1, BF16>(h, kpl, grid, st, p);
  if (qc == 2) return launch_stream_tma_k<2, BF16>(h, kpl, grid, st, p);
  if (qc == 4) return launch_stream_tma_k<4, BF16>(h, kpl, grid, st, p);
  return launch_stream_tma_k<8, BF16>(h, kpl, grid, st, p);
}

template <bool BF16>
static void launch_stream_q(int qc, int kpl, dim3 grid, cudaStream_t st, const StreamParams& p) {
  if (qc == 1) launch_stream_k<1, BF16>(kpl, grid, st, p);
  else if (qc == 2) launch_stream_k<2, BF16>(kpl, grid, st, p);
  else if (qc == 4) launch_stream_k<4, BF16>(kpl, grid, st, p);
  else launch_stream_k<8, BF16>(kpl, grid, st, p);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Streaming search over the resident shard.  Exactly one of (out_scores/out_rows) or out_packed
// is normally given; any may be null.
int mmf_stream_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                      float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc, cudaStream_t st) {
  const int Q = (int)n_queries;
  const int qc = Q <= 1 ? 1 : Q <= 2 ? 2 : Q <= 4 ? 4 : 8;
  const int gy = (Q + qc - 1) / qc;
  const int kpl = top_k <= 64 ? 4 : top_k <= 192 ? 8 : 16;
  // TMA-staged kernel (option "stream_tma", default): one persistent block per SM and query chunk sweeps a contiguous
  // slab of rows through a 160 KB shared-memory ring; the register kernel: 4 blocks per SM
  // (8 queries per block with top_k > 192 would need 32 KB of candidate buffers next to the 192 KB ring: register kernel)
  const bool tma = h->opt.stream_tma != 0 && !(qc == 8 && kpl == 16);
  const long long target = (long long)h->sm_count * (tma ? 1 : 4);
  long long gx = std::max<long long>(1, (target + gy - 1) / gy);
  gx = std::min<long long>(gx, std::max<long long>(1, (h->vault_rows + 511) / 512));
  long long rows_per_cta = align_up((size_t)((h->vault_rows + gx - 1) / gx), 32);
  gx = (h->vault_rows + rows_per_cta - 1) / rows_per_cta;

  // scratch: [done 64 KB | qn | g_tau | part_cnt | part_keys]
  const size_t off_qn = 65536;
  const size_t off_tau = off_qn + align_up((size_t)Q * MMF_DIM * 4, 256);
  const size_t off_cnt = off_tau + align_up((size_t)Q * 4, 256);
  const size_t off_keys = off_cnt + align_up((size_t)gx * Q * 4, 256);
  const size_t total = off_keys + (size_t)gx * Q * top_k * 8;
  int rc = mmf_ensure_scratch(h, total, st);
  if (rc != MMF_OK) return rc;
  char* s = (char*)h->scratch();
  float* qn = (float*)(s + off_qn);
  u32* g_tau = (u32*)(s + off_tau);

  query_prep_kernel<<<(Q + 7) / 8, 256, 0, st>>>(queries, Q, qn, g_tau);
  MMF_LAUNCH_OK(h);

  StreamParams p;
  p.vault = h->vault;
  p.n_rows = h->vault_rows;
  p.row_base = (u32)h->vault_row_offset;
  p.qn = qn;
  p.n_queries = Q;
  p.top_k = top_k;
  p.rows_per_cta = rows_per_cta;
  p.g_tau = g_tau;
  p.part_keys = (u64*)(s + off_keys);
  p.part_cnt = (int*)(s + off_cnt);
  p.done = (u32*)s;
  p.out_scores = out_scores;
  p.out_rows = (long long*)out_rows;
  p.out_packed = (u64*)out_packed;
  p.out_disc = out_disc;
  p.threshold = threshold;
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (tma) {
    rc = h->vault_mode == MMF_VAULT_BF16 ? launch_stream_tma_q<true>(h, qc, kpl, grid, st, p)
                                         : launch_stream_tma_q<false>(h, qc, kpl, grid, st, p);
    if (rc != MMF_OK) return rc;
  } else if (h->vault_mode == MMF_VAULT_BF16) {
    launch_stream_q<true>(qc, kpl, grid, st, p);
  } else {
    launch_stream_q<false>(qc, kpl, grid, st, p);
  }
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

int mmf_fill_empty(mmf_handle* h, int64_t n_queries, int top_k, float* out_scores, int64_t* out_rows,
                   uint64_t* out_packed, float* out_disc, cudaStream_t st) {
  const long long n = std::max<long long>(n_queries * top_k, n_queries);
  fill_empty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n_queries, top_k, out_scores, (long long*)out_rows,
                                                                 (u64*)out_packed, out_disc);
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

extern "C" int mmf_topk_merge(mmf_handle* h, const uint64_t* packed, int n_lists, int64_t n_queries, int k_in,
                              int top_k, double threshold, float* out_scores, int64_t* out_rows, float* out_discrepancy,
                              mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_lists <= 0 || n_queries < 0 || k_in <= 0 || top_k <= 0 || top_k > MMF_MAX_TOP_K || (n_queries > 0 && !packed))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "topk_merge: bad argument (n_lists=%d n_queries=%lld k_in=%d top_k=%d)",
                         n_lists, (long long)n_queries, k_in, top_k);
  if (n_queries == 0) return MMF_OK;
  topk_merge_kernel<<<(unsigned)n_queries, 256, 0, (cudaStream_t)stream>>>(
      (const u64*)packed, n_lists, n_queries, k_in, top_k, threshold, out_scores, (long long*)out_rows, nullptr,
      out_discrepancy);
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}
