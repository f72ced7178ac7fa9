// K3: batched Truth-Vault search on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), with the
// top-k fused into the epilogue so the (queries x rows) score matrix never leaves the SM.
// Replaces the batched form of misinfo_forensics.py:446 (Vn @ q) and :449-450 (argsort top-k).
//
// Operands (K-major, 128B-swizzled tiles of [rows][64 elements], 128 B per row, fed by TMA):
//   A = queries (M = 128 per tile), normalised by the prep kernel.  The query tile of a strip
//       is RESIDENT on the SM (the v1 kernel streamed it with every vault tile and was bound
//       by L2->SM operand traffic): 8 k-block tiles = 128 KB of shared memory.
//   B = vault rows, streamed from HBM through a 3-stage mbarrier ring.
//   MMF_VAULT_BF16: one bf16 plane each, N = 256 rows per tile, 4 UMMA (K=16) per k-block.
//   MMF_VAULT_FP32: fp32-exact.  x*2^8 = hi + lo (two fp16 planes, 22+ bits), and
//       q.v * 2^16 = qh.vh + ql.vh + qh.vl  (+ ql.vl, < 2^-22 relative, dropped)
//     -> 12 UMMA per k-block (N = 128), all into ONE fp32 TMEM accumulator; scores = D*2^-16.
//     qh is the smem-resident A operand; ql (another 128 KB) lives in TENSOR MEMORY
//     (256 columns, written once per strip with tcgen05.st) and feeds the TS-form MMA.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected
// lane), warps 2-5 = epilogue.  Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA) and a
// double-buffered TMEM accumulator (tmem_full/tmem_empty, MMA <-> epilogue), so the epilogue of
// tile i overlaps the MMAs of tile i+1.  TMEM: 2 x N accumulator columns (+ 256 for ql) = 512.
//
// Epilogue = streaming top-k.  TMEM lane == query, so each epilogue thread owns one query:
// it reads 32 accumulator columns at a time (tcgen05.ld 32x32b.x32), compares them with its
// private threshold (a lower bound of its k-th best) and appends the rare survivors to its
// candidate list in global memory (L2-resident); a full list is compacted to the exact top-k
// by the whole warp (topk.cuh).  A block works on "strips" (one query tile x a run of vault
// tiles) so that state stays in registers; a merge kernel selects the final top-k per query
// from the strips' lists.
#include "common.cuh"
#include "topk.cuh"

#include <cuda.h>
#include <algorithm>
#include <new>

namespace mmf {

constexpr int TILE_M = 128;          // queries per tile   (UMMA M)
constexpr int KBLK = 64;             // elements per k-block: 128 B rows, one 128B-swizzle atom
constexpr int TILE_BYTES = 128 * KBLK * 2;   // 16 KB: [128 rows][64 elements]
constexpr int Q_RESIDENT_BYTES = (MMF_DIM / KBLK) * TILE_BYTES;   // 128 KB
constexpr int MMA_STAGES = 3;
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // 32 KB: bf16 [256][64], or fp16 hi [128][64] + lo [128][64]
__host__ __device__ constexpr int tile_n(bool split) { return split ? 128 : 256; }
constexpr int NUM_KBLK = MMF_DIM / KBLK;     // 8
constexpr int MMA_THREADS = 192;

struct MmaParams {
  int n_queries;           // valid queries
  int q_pad;               // padded to TILE_M
  long long n_rows;        // vault rows in this shard
  u32 row_base;            // global id of row 0
  int top_k;
  int q_tiles, v_tiles;
  long long units;         // q_tiles * v_tiles
  u64* cand;               // [strip][TILE_M][C]
  int* cand_cnt;           // [strip][TILE_M]
  float inv_scale;         // accumulator -> score
  const uint4* q_lo;       // fp32-exact mode: the lo plane of the query operand ([q_pad][512] fp16)
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, u64* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, u64* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, bf16/fp16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_f16(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same with A read from tensor memory (lane = row, two 16-bit K elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(u32 taddr, const u32 (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrives on the mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(u64* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(u32 taddr, u32 (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO = 1
// (ignored for swizzled K-major), descriptor version 1 (sm_100).
__device__ __forceinline__ u64 umma_smem_desc(u32 saddr) {
  return (u64)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: fp32 accumulator, A/B format fmt (0 = fp16, 1 = bf16), both K-major
__host__ __device__ constexpr u32 umma_idesc(u32 fmt, u32 m, u32 n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- query operand prep --------------------------------------------------------------------
// One warp per padded query row: q / ||q|| (misinfo_forensics.py:439), then the MMA operand
// planes: bf16 (1 plane) or fp16 hi/lo of q*2^8 (2 planes, plane p at row p*q_pad + i).
__global__ void __launch_bounds__(256) mma_query_prep_kernel(const float* __restrict__ q, int n_queries, int q_pad,
                                                             int split, void* __restrict__ planes) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= q_pad) return;
  float v[MMF_DIM / 32], ss = 0.f;
#pragma unroll
  for (int j = 0; j < MMF_DIM / 32; ++j) {
    v[j] = (w < n_queries) ? q[(long long)w * MMF_DIM + j * 32 + lane] : 0.f;
    ss = fmaf(v[j], v[j], ss);
  }
  const float norm = sqrtf(warp_sum(ss));
#pragma unroll
  for (int j = 0; j < MMF_DIM / 32; ++j) {
    const float x = (w < n_queries) ? v[j] / norm : 0.f;
    const long long o = (long long)w * MMF_DIM + j * 32 + lane;
    if (split) {
      __half* hi = reinterpret_cast<__half*>(planes);
      __half* lo = hi + (long long)q_pad * MMF_DIM;
      const float y = x * MMF_SPLIT_SCALE;
      const __half h = __float2half_rn(y);
      hi[o] = h;
      lo[o] = __float2half_rn(y - __half2float(h));
    } else {
      reinterpret_cast<__nv_bfloat16*>(planes)[o] = __float2bfloat16_rn(x);
    }
  }
}

// ---- the search kernel ---------------------------------------------------------------------
template <bool SPLIT, int KPL>
__global__ void __launch_bounds__(MMA_THREADS, 1)
vault_mma_topk_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const MmaParams p) {
  constexpr int TILE_N = tile_n(SPLIT);
  constexpr int STAGES = MMA_STAGES;
  constexpr int C = 32 * KPL;
  constexpr u32 IDESC = umma_idesc(SPLIT ? 0u : 1u, TILE_M, TILE_N);
  constexpr u32 QL_COL = 2 * TILE_N;                 // fp32-exact: TMEM columns [256,512) hold ql

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_smem = smem;                                   // resident query tile (qh / bf16)
  unsigned char* stage_smem = smem + Q_RESIDENT_BYTES;
  u64* full_bar = reinterpret_cast<u64*>(stage_smem + STAGES * STAGE_BYTES);
  u64* empty_bar = full_bar + STAGES;
  u64* tmem_full = empty_bar + STAGES;
  u64* tmem_empty = tmem_full + 2;
  u64* q_full = tmem_empty + 2;       // producer -> MMA: resident query tile landed
  u64* q_empty = q_full + 1;          // MMA -> producer: every MMA of the previous strip retired
  u64* ql_full = q_empty + 1;         // epilogue -> MMA: ql written to TMEM
  u32* tmem_slot = reinterpret_cast<u32*>(ql_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u0 = (long long)blockIdx.x * p.units / gridDim.x;
  const long long u1 = (long long)(blockIdx.x + 1) * p.units / gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(ql_full, 4);
    fence_barrier_init();
  }
  if (warp == 1) {   // the whole tensor memory: accumulators (+ ql)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const u32 tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      u32 it = 0, strip = 0;
      int cur_qt = -1;
      for (long long u = u0; u < u1; ++u) {
        const int qt = (int)(u / p.v_tiles);
        const int vt = (int)(u % p.v_tiles);
        if (qt != cur_qt) {                           // new strip: (re)load the resident query tile
          cur_qt = qt;
          mbar_wait(q_empty, (strip & 1) ^ 1);
          mbar_expect_tx(q_full, Q_RESIDENT_BYTES);
          for (int kb = 0; kb < NUM_KBLK; ++kb)
            tma_load_2d(q_smem + kb * TILE_BYTES, &tm_a, q_full, kb * KBLK, qt * TILE_M);
          ++strip;
        }
        for (int kb = 0; kb < NUM_KBLK; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar + s, ((it / STAGES) & 1) ^ 1);
          unsigned char* st = stage_smem + s * STAGE_BYTES;
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
          if (SPLIT) {
            tma_load_3d(st, &tm_b, full_bar + s, kb * KBLK, 0, vt * TILE_N);
            tma_load_3d(st + TILE_BYTES, &tm_b, full_bar + s, kb * KBLK, 1, vt * TILE_N);
          } else {
            tma_load_2d(st, &tm_b, full_bar + s, kb * KBLK, vt * TILE_N);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      u32 it = 0, tile = 0, strip = 0;
      int cur_qt = -1;
      const u32 q_addr = smem_u32(q_smem);
      for (long long u = u0; u < u1; ++u, ++tile) {
        const int qt = (int)(u / p.v_tiles);
        if (qt != cur_qt) {
          cur_qt = qt;
          mbar_wait(q_full, strip & 1);
          if (SPLIT) mbar_wait(ql_full, strip & 1);
          ++strip;
        }
        const u32 acc = tile & 1;
        mbar_wait(tmem_empty + acc, ((tile >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const u32 d_tmem = tmem_base + acc * TILE_N;
        for (int kb = 0; kb < NUM_KBLK; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar + s, (it / STAGES) & 1);
          tcgen05_fence_after();
          const u32 qa = q_addr + kb * TILE_BYTES;
          const u32 sb = smem_u32(stage_smem + s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < KBLK / 16; ++k)          // qh.vh  (bf16 mode: q.v)
            umma_f16(d_tmem, umma_smem_desc(qa + k * 32), umma_smem_desc(sb + k * 32), IDESC, (kb | k) != 0);
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < KBLK / 16; ++k)        // ql.vh, ql from tensor memory
              umma_f16_ts(d_tmem, tmem_base + QL_COL + kb * (KBLK / 2) + k * 8, umma_smem_desc(sb + k * 32), IDESC, 1);
#pragma unroll
            for (int k = 0; k < KBLK / 16; ++k)        // qh.vl
              umma_f16(d_tmem, umma_smem_desc(qa + k * 32), umma_smem_desc(sb + TILE_BYTES + k * 32), IDESC, 1);
          }
          umma_commit(empty_bar + s);                 // smem slot free once these MMAs retire
        }
        umma_commit(tmem_full + acc);                 // accumulator complete -> epilogue
        if (u + 1 == u1 || (int)((u + 1) / p.v_tiles) != qt) umma_commit(q_empty);   // strip done
      }
    }
  } else {
    // ===== epilogue: thread == query (TMEM lane), streaming top-k =====
    const int quarter = warp & 3;                     // TMEM lanes a warp may touch: 32*(warp%4)..+31
    const int m = quarter * 32 + lane;
    const u32 lane_base = tmem_base + ((u32)(quarter * 32) << 16);
    const int k = p.top_k;
    const float acc_scale = 1.0f / p.inv_scale;
    float tau_acc = -INFINITY;                        // threshold in accumulator units
    int cnt = 0;
    int cur_qt = -1;
    u64* buf = nullptr;
    bool valid_q = false;
    u32 tile = 0;
    for (long long u = u0; u < u1; ++u, ++tile) {
      const int qt = (int)(u / p.v_tiles);
      const int vt = (int)(u % p.v_tiles);
      if (qt != cur_qt) {                             // new strip: flush the old one, reset state
        if (cur_qt >= 0) p.cand_cnt[(long long)(blockIdx.x + cur_qt) * TILE_M + m] = cnt;
        cur_qt = qt;
        cnt = 0;
        tau_acc = -INFINITY;
        buf = p.cand + ((long long)(blockIdx.x + qt) * TILE_M + m) * C;
        valid_q = (qt * TILE_M + m) < p.n_queries;
        if (SPLIT) {
          // this thread's query row of the lo plane -> its TMEM lane, 2 fp16 per column.  Every
          // MMA of the previous strip has retired (its last accumulator was consumed above).
          const uint4* src = p.q_lo + (long long)(qt * TILE_M + m) * (MMF_DIM * 2 / 16);
#pragma unroll 1
          for (int c = 0; c < MMF_DIM / 64; ++c) {
            u32 w[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 x = __ldg(src + c * 8 + i);
              w[4 * i] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
            }
            tmem_st32(lane_base + QL_COL + c * 32, w);
          }
          tmem_wait_st();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ql_full);
        }
      }
      const u32 acc = tile & 1;
      mbar_wait(tmem_full + acc, (tile >> 1) & 1);
      tcgen05_fence_after();
      const long long row0 = (long long)vt * TILE_N;
      const int n_cols = (int)min((long long)TILE_N, p.n_rows - row0);   // valid columns of this tile
#pragma unroll 1
      for (int c = 0; c < TILE_N / 32; ++c) {
        u32 v[32];
        tmem_ld32(lane_base + acc * TILE_N + c * 32, v);
        tmem_wait_ld();
        // fast path (almost always): is the chunk maximum below the threshold?  A max tree keeps the
        // dependent chain at 5 instead of 32 -- one lone warp per scheduler cannot hide latency.
        // (fmaxf drops NaN next to a number; vaults with NaN rows never reach this kernel.)
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          mx[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                        fmaxf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
        const float m8 = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        if (!(m8 < tau_acc) && valid_q) {
          u32 mask = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) mask |= (!(__uint_as_float(v[j]) < tau_acc) ? 1u : 0u) << j;
          const int lim = n_cols - c * 32;            // columns >= lim are TMA zero fill, not vault rows
          if (lim < 32) mask &= (lim <= 0) ? 0u : ((1u << lim) - 1u);
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            // v[j] without dynamic register indexing: 5-level select tree
            u32 t16[16], t8[8], t4[4], t2[2];
#pragma unroll
            for (int i = 0; i < 16; ++i) t16[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = (j & 2) ? t16[2 * i + 1] : t16[2 * i];
#pragma unroll
            for (int i = 0; i < 4; ++i) t4[i] = (j & 4) ? t8[2 * i + 1] : t8[2 * i];
#pragma unroll
            for (int i = 0; i < 2; ++i) t2[i] = (j & 8) ? t4[2 * i + 1] : t4[2 * i];
            const float a = __uint_as_float((j & 16) ? t2[1] : t2[0]);
            buf[cnt++] = pack_key(a * p.inv_scale, p.row_base + (u32)(row0 + c * 32 + j));
          }
        }
        // keep room for the next 32 columns; compaction is warp-cooperative, one query at a time
        u32 need = __ballot_sync(FULL, cnt > C - 32);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          u64* b = reinterpret_cast<u64*>(__shfl_sync(FULL, reinterpret_cast<u64>(buf), src));
          const int n = __shfl_sync(FULL, cnt, src);
          __syncwarp();
          float t = 0.f;
          const int kept = warp_compact<KPL>(b, n, k, &t);
          if (lane == src) {
            cnt = kept;
            tau_acc = fmaxf(tau_acc, t * acc_scale);
          }
          __syncwarp();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
    }
    if (cur_qt >= 0) p.cand_cnt[(long long)(blockIdx.x + cur_qt) * TILE_M + m] = cnt;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// One block per query: gather the strips that cover its query tile, select + sort the top-k.
template <int KPL>
__global__ void __launch_bounds__(256) mma_merge_kernel(const MmaParams p, int grid_main, double threshold,
                                                        float* out_scores, long long* out_rows, u64* out_packed,
                                                        float* out_disc) {
  constexpr int C = 32 * KPL;
  __shared__ SelectSmem sel;
  __shared__ int s_first, s_count;
  const int qg = blockIdx.x;
  const int qt = qg / TILE_M, m = qg % TILE_M;
  if (threadIdx.x == 0) {
    int first = -1, count = 0;
    const long long lo = (long long)qt * p.v_tiles, hi = lo + p.v_tiles;
    for (int c = 0; c < grid_main; ++c) {
      const long long a = (long long)c * p.units / grid_main, b = (long long)(c + 1) * p.units / grid_main;
      if (a < hi && b > lo && b > a) {
        if (first < 0) first = c;
        ++count;
      }
    }
    s_first = first;
    s_count = count;
  }
  __syncthreads();
  CandidateLists src;
  src.lists = p.cand + ((long long)(s_first + qt) * TILE_M + m) * C;
  src.counts = p.cand_cnt + (long long)(s_first + qt) * TILE_M + m;
  src.n_lists = s_count;
  src.k_in = C;
  src.list_stride = (long long)TILE_M * C;
  src.count_stride = TILE_M;
  block_select_topk(src, p.top_k, sel, out_scores ? out_scores + (long long)qg * p.top_k : nullptr,
                    out_rows ? out_rows + (long long)qg * p.top_k : nullptr,
                    out_packed ? out_packed + (long long)qg * p.top_k : nullptr, out_disc ? out_disc + qg : nullptr,
                    threshold);
}

}  // namespace mmf

using namespace mmf;

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MmaState {
  EncodeTiledFn encode = nullptr;
  CUtensorMap tm_vault;
  bool vault_map_ok = false;
  bool attrs_set = false;
};

static MmaState* state_of(mmf_handle* h) {
  if (!h->mma_state) {
    MmaState* s = new (std::nothrow) MmaState();
    if (!s) return nullptr;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      s->encode = (EncodeTiledFn)fn;
    else
      cudaGetLastError();
    h->mma_state = s;
  }
  return (MmaState*)h->mma_state;
}

static bool encode_map(MmaState* s, CUtensorMap* map, CUtensorMapDataType dt, int rank, void* base,
                       const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  cuuint32_t estr[3] = {1, 1, 1};
  return s->encode(map, dt, (cuuint32_t)rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int mmf_mma_vault_changed(mmf_handle* h) {
  MmaState* s = state_of(h);
  if (!s) return MMF_OK;
  s->vault_map_ok = false;
  if (!s->encode || !h->vault || h->vault_rows <= 0) return MMF_OK;
  if (h->vault_mode == MMF_VAULT_BF16) {
    const cuuint64_t dims[2] = {MMF_DIM, (cuuint64_t)h->vault_rows};
    const cuuint64_t strides[1] = {MMF_DIM * 2};
    const cuuint32_t box[2] = {KBLK, (cuuint32_t)tile_n(false)};
    s->vault_map_ok = encode_map(s, &s->tm_vault, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->vault, dims, strides, box);
  } else {
    // [row][plane][512] fp16 viewed as (k, plane, row)
    const cuuint64_t dims[3] = {MMF_DIM, 2, (cuuint64_t)h->vault_rows};
    const cuuint64_t strides[2] = {MMF_DIM * 2, MMF_DIM * 4};
    const cuuint32_t box[3] = {KBLK, 1, (cuuint32_t)tile_n(true)};
    s->vault_map_ok = encode_map(s, &s->tm_vault, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, h->vault, dims, strides, box);
  }
  return MMF_OK;
}

int mmf_mma_supported(const mmf_handle* h, int64_t n_queries, int top_k) {
  const MmaState* s = (const MmaState*)h->mma_state;
  // NaN rows must rank first (np.argsort semantics); only the streaming kernel's compare keeps them
  return s && s->encode && s->vault_map_ok && h->vault_rows > 0 && h->vault_nan_rows == 0 && top_k >= 1 &&
         top_k <= MMF_MAX_TOP_K && n_queries > 0;
}

void mmf_mma_destroy(mmf_handle* h) {
  delete (MmaState*)h->mma_state;
  h->mma_state = nullptr;
}

template <bool SPLIT, int KPL>
static int launch_mma(mmf_handle* h, MmaState* s, const CUtensorMap& tm_q, const MmaParams& p, int grid, double threshold,
                      float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc, cudaStream_t st) {
  const int smem = Q_RESIDENT_BYTES + MMA_STAGES * STAGE_BYTES + 256 + 1024;
  auto kern = vault_mma_topk_kernel<SPLIT, KPL>;
  MMF_CUDA_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, MMA_THREADS, smem, st>>>(tm_q, s->tm_vault, p);
  MMF_LAUNCH_OK(h);
  mma_merge_kernel<KPL><<<p.n_queries, 256, 0, st>>>(p, grid, threshold, out_scores, (long long*)out_rows,
                                                      (u64*)out_packed, out_disc);
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

int mmf_mma_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                   float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc, cudaStream_t st) {
  MmaState* s = state_of(h);
  if (!s || !mmf_mma_supported(h, n_queries, top_k))
    return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "tcgen05 path unavailable (TMA descriptor encode failed?)");
  const bool split = h->vault_mode == MMF_VAULT_FP32;
  const int npl = split ? 2 : 1;
  const int kpl = top_k <= 32 ? 4 : top_k <= 128 ? 8 : 16;
  const int C = 32 * kpl;

  MmaParams p;
  p.n_queries = (int)n_queries;
  p.q_tiles = (int)((n_queries + TILE_M - 1) / TILE_M);
  p.q_pad = p.q_tiles * TILE_M;
  p.n_rows = h->vault_rows;
  p.row_base = (u32)h->vault_row_offset;
  p.top_k = top_k;
  p.v_tiles = (int)((h->vault_rows + tile_n(split) - 1) / tile_n(split));
  p.units = (long long)p.q_tiles * p.v_tiles;
  p.inv_scale = split ? (MMF_SPLIT_INV_SCALE * MMF_SPLIT_INV_SCALE) : 1.0f;
  const int grid = (int)std::min<long long>(h->sm_count, p.units);
  const long long strips = (long long)grid + p.q_tiles;

  // scratch: [64 KB counters (stream kernel) | query planes | cand_cnt | cand]
  auto al = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  const size_t off_q = 65536;
  const size_t off_cnt = off_q + al((size_t)npl * p.q_pad * MMF_DIM * 2);
  const size_t off_cand = off_cnt + al((size_t)strips * TILE_M * 4);
  const size_t total = off_cand + (size_t)strips * TILE_M * C * 8;
  int rc = mmf_ensure_scratch(h, total, st);
  if (rc != MMF_OK) return rc;
  char* sc = (char*)h->scratch;
  void* planes = sc + off_q;
  p.cand_cnt = (int*)(sc + off_cnt);
  p.cand = (u64*)(sc + off_cand);
  p.q_lo = reinterpret_cast<const uint4*>((const char*)planes + (size_t)p.q_pad * MMF_DIM * 2);

  mma_query_prep_kernel<<<(p.q_pad + 7) / 8, 256, 0, st>>>(queries, p.n_queries, p.q_pad, split ? 1 : 0, planes);
  MMF_LAUNCH_OK(h);

  CUtensorMap tm_q;
  const cuuint64_t dims[2] = {MMF_DIM, (cuuint64_t)npl * p.q_pad};
  const cuuint64_t strides[1] = {MMF_DIM * 2};
  const cuuint32_t box[2] = {KBLK, TILE_M};
  if (!encode_map(s, &tm_q, split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, planes, dims,
                  strides, box))
    return mmf_set_error(h, MMF_ERR_CUDA, "cuTensorMapEncodeTiled failed for the query operand");

#define MMF_MMA_CASE(SPLIT_, KPL_)                                                                                  \
  return launch_mma<SPLIT_, KPL_>(h, s, tm_q, p, grid, threshold, out_scores, out_rows, out_packed, out_disc, st)
  if (split) {
    if (kpl == 4) MMF_MMA_CASE(true, 4);
    if (kpl == 8) MMF_MMA_CASE(true, 8);
    MMF_MMA_CASE(true, 16);
  } else {
    if (kpl == 4) MMF_MMA_CASE(false, 4);
    if (kpl == 8) MMF_MMA_CASE(false, 8);
    MMF_MMA_CASE(false, 16);
  }
#undef MMF_MMA_CASE
}
