// K3: batched Truth-Vault search on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), with the
// top-k fused into the epilogue so the (queries x rows) score matrix never leaves the SM.
// Replaces the batched form of misinfo_forensics.py:446 (Vn @ q) and :449-450 (argsort top-k).
//
// Shape of the kernel (every choice below was measured; see DESIGN.md section 7 and tools/umma_micro.cu):
//   * thread-block PAIRS (tcgen05 cta_group::2): one UMMA covers M = 256 queries (two query tiles,
//     one per CTA) x N = 128 vault rows, and each CTA stores only HALF of every vault tile.  With one
//     CTA per MMA the TMA fill plus the tensor core's operand fetch saturate the 128 B/clk shared-
//     memory port; the pair halves that traffic.  (A single query tile falls back to CG = 1.)
//   * A = queries, normalised by the prep kernel, RESIDENT on the SM for a whole strip:
//       plane 0 (bf16 q, or fp16 qh) lives in TENSOR MEMORY: 256 columns, written once per strip by
//       the epilogue threads with tcgen05.st (lane = query, 2 elements per column) -> TS-form UMMA;
//       fp32-exact mode also keeps plane 1 (ql, 128 KB) in shared memory, loaded by TMA -> SS form.
//   * B = vault rows, K-major 128B-swizzled [rows][64] tiles streamed by TMA through an mbarrier ring
//       (bf16: 8 stages x 2 k-blocks x 8 KB per CTA; fp32-exact: 6 x (vh + vl = 16 KB) next to ql).
//   MMF_VAULT_BF16: D += q.v, 4 TS-UMMA (K=16) per 64-wide k-block.
//   MMF_VAULT_FP32: fp32-exact.  x*2^8 = hi + lo (two fp16 planes, 22+ bits).
//     top_k <= 16 (the reference's use: k = 5 / 10): SCREENED search (VAR_SCREEN).  One pass
//       qh.vh over the hi planes -- the bf16-shaped pipeline on fp16 data, TMA fetches only the first 1 KB of each
//       2 KB row -- is within a proven eps of the exact score; every element within 2*eps of the running k-th best
//       survives, mma_rerank_kernel re-scores the survivors exactly (fp32, from hi+lo, the streaming kernel's
//       arithmetic) and selects the top-k among them.  A third of the MMA work and half of the DRAM bytes of the
//       3-pass form; results bit-identical to the streaming kernel.  Bands that do not fit set a flag and the
//       guarded 3-pass kernel (VAR_GUARD, launched after every screened search, returns at once otherwise) redoes
//       the batch and merges its own lists behind a grid barrier: 4 launches per search (prep, screening, rerank,
//       guarded no-op).
//     top_k > 16: three passes,  q.v * 2^16 = qh.vh + qh.vl + ql.vh  (+ ql.vl, < 2^-22 relative, dropped)
//       -> 8 TS + 4 SS UMMA per k-block, all into ONE fp32 TMEM accumulator; scores = D * 2^-16.
//   * L2-aware schedule (pair_schedule): the pairs working on different query-tile groups sweep the
//     same vault segment together, so a vault tile is fetched from DRAM by one and hit in L2 by the rest;
//     their producers pace each other (lock-step, option "lockstep") so that they stay within an L2's worth.
//
// Roles (320 threads): warps 0-7 = epilogue (warp % 4 = the TMEM lane quarter it may touch), warp 8 =
// TMA producer, warp 9 = TMEM owner + MMA issuer (the whole warp runs the loop in the uniform datapath,
// one elected lane issues; leader CTA of the pair only).  Pipelines: smem ring (full/empty mbarriers,
// TMA <-> MMA) and a double-buffered TMEM accumulator (tmem_full/tmem_empty, MMA <-> epilogue), so the
// epilogue of tile i overlaps the MMAs of tile i+1.  TMEM: 2 x 128 accumulator columns + 256 operand
// columns = 512.  With pairs, `full`, `tmem_empty`, `q_full`, `qa_full` live in the leader (the peer
// signals them remotely), `empty`, `tmem_full`, `q_empty` exist in both CTAs (multicast commits).
//
// Epilogue = streaming top-k.  TMEM lane == query, so each epilogue thread owns one query: it reads
// 32 accumulator columns at a time (tcgen05.ld 32x32b.x32), compares their maximum with its private
// threshold (a lower bound of its k-th best) and appends the rare survivors to its candidate list in
// global memory (L2-resident).  The chunks of a tile are software-pipelined over two register buffers and the
// accumulator is handed back as soon as its last column is in registers.  Thresholds are grid-wide lower bounds:
// top_k <= 16 through a pool of exactly top_k hashed bucket maxima (pool_bucket: top_k distinct rows, so the minimum
// over the buckets bounds the k-th best from below), SEEDED at the start of a strip with every thread's tile (or
// chunk) maximum before anything is filtered; top_k > 16 through a per-query HISTOGRAM of candidate scores (the
// lower edge of the bin holding the k-th best counted candidate; the bucket minimum sits near rank k*H(k), the
// histogram edge near rank 1.4*k), refreshed every few tiles and when a list is compacted (whole warp, topk.cuh).
// A candidate event is a group mask + one indexed jump + a single event body.  A block works on "strips"
// (one group of query tiles x a run of vault tiles) so that state stays in registers; a tail kernel
// (mma_merge_kernel / mma_rerank_kernel: parallel slot gather, one staging sweep, rank-by-counting or radix
// select) produces the final top-k per query from the strips' lists.
//
// Switches: none are read from the environment on the search path.  mmf_set_option (include/mmf_b200.h) has "screen"
// (0 = 3-pass kernel instead of the screened search, A/B + triage), "debug" (bit 0: skip the filter, 1: skip the vault
// TMA, 2: skip the MMA warp's waits, 3: print in-kernel cycle counts; the bits only exist in a -DMMF_MMA_TRIAGE=1
// build), "force_cg" (1 or 2 CTAs per MMA), "flat_schedule" (plain flattened schedule instead of the L2-aware one),
// "epi_parity" and "lockstep".  Variants that lost their A/B on a B200 in round 2 and were deleted: a 12-stage ring
// and an L2 prefetch for the screening pass (both slower: the loader is not the limit, tools/tma_stream_micro.cu), the
// bucket-pool bound for top_k > 16 (histogram: 1.5x faster), sorted per-thread registers for top_k <= 16, the
// warm-up "storm" (one pair in eight filtering without a bound), a half-tile accumulator hand-off (N = 64 MMAs run
// the tensor pipe at ~60 %), programmatic dependent launch between the kernels of a search (no gain), the first form
// of the merge / rerank tails.
#include "common.cuh"
#include "topk.cuh"

#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

// Triage build (-DMMF_MMA_TRIAGE=1): the in-kernel cycle counters and the "debug" option bits are compiled in.  They are
// OFF in the shipped library: the epilogue is latency-bound (one warp per scheduler retires an instruction every ~8 clk),
// and the conditional clock reads + option tests were ~12 of its ~150 per-tile instructions.
#ifndef MMF_MMA_TRIAGE
#define MMF_MMA_TRIAGE 0
#endif

namespace mmf {

constexpr bool TRIAGE = MMF_MMA_TRIAGE != 0;
constexpr int TILE_M = 128;          // queries per tile   (UMMA M)
constexpr int KBLK = 64;             // elements per k-block: 128 B rows, one 128B-swizzle atom
constexpr int TILE_BYTES = 128 * KBLK * 2;   // 16 KB: [128 rows][64 elements]
constexpr int TILE_N = 128;          // vault rows per tile (UMMA N); 2 accumulator buffers = 256 TMEM columns
constexpr int Q_RESIDENT_BYTES = (MMF_DIM / KBLK) * TILE_BYTES;   // 128 KB
// A stage holds this CTA's share of one 128-row x 64-element vault k-block: 128/CG rows of 128 B per
// plane (CG = CTAs per MMA: with cta_group::2 each CTA of the pair stores HALF of the B tile).
// smem: bf16 = 12 x 16 KB (CG 1) / 16 x 8 KB (CG 2); fp32-exact = resident ql (128 KB) + 3 x 32 KB / 6 x 16 KB
// bf16 mode puts TWO k-blocks in a stage: 8 MMAs per barrier hand-off instead of 4 (a try_wait costs ~90 clk
// even when the data is there, against 256 clk of MMA work per k-block)
__host__ __device__ constexpr int kblk_per_stage(bool split) { return split ? 1 : 2; }
__host__ __device__ constexpr int stage_bytes(bool split, int cg) { return (split ? 2 : 1) * TILE_BYTES / cg * kblk_per_stage(split); }
__host__ __device__ constexpr int mma_stages(bool split, int cg) { return split ? 3 * cg : (cg == 1 ? 6 : 8); }
__host__ __device__ constexpr int mma_smem_bytes(bool split, int cg) {
  return (split ? Q_RESIDENT_BYTES : 0) + mma_stages(split, cg) * stage_bytes(split, cg);
}
constexpr int NUM_KBLK = MMF_DIM / KBLK;     // 8
constexpr int EPI_WARPS = 8;          // 2 per TMEM lane quarter: a lone warp per scheduler cannot hide its own latency
constexpr int PRODUCER_WARP = EPI_WARPS;      // warps 0-7: epilogue (warp%4 = TMEM lane quarter), 8: TMA producer,
constexpr int MMA_WARP = EPI_WARPS + 1;       // 9: MMA issuer -- the highest warp id on its scheduler, so it wins arbitration
constexpr int MMA_THREADS = 64 + 32 * EPI_WARPS;
constexpr int MERGE_MAX_SLOTS_BYTES = 4 * 160 * 4;     // (MERGE_MAX_SLOTS ints, see the tail kernels)

struct MmaParams {
  int n_queries;           // valid queries
  int q_pad;               // padded to TILE_M
  long long n_rows;        // vault rows in this shard
  u32 row_base;            // global id of row 0
  int top_k;
  int q_tiles, v_tiles;    // q_tiles is padded to a multiple of CG
  // L2-aware schedule (see pair_schedule): qtp = q_tiles / CG groups of query tiles, n_aligned = seg * qtp
  // pairs sweep [0, v_aligned) in lock-step, the other pairs share [v_aligned, v_tiles) in qtp-major order
  int qtp, seg, n_aligned, v_aligned;
  u64* cand;               // [strip][2 column halves][CG][TILE_M][C]
  int* cand_cnt;           // [strip][2][CG][TILE_M]
  u32* pool;               // [q_pad][256] bucket maxima: slot (row % top_k) = best score key seen among those rows by
                           // any block; top_k distinct rows, so min over the slots <= the k-th best (a grid-wide bound)
  u32* g_tau;              // [q_pad] best known lower bound of each query's k-th best (score key), shared grid-wide
  float inv_scale;         // accumulator -> score
  int debug;               // perf triage only (env MMF_MMA_DEBUG): 1 = epilogue skips the filter, 2 = no vault TMA
  const uint4* q_plane0;   // plane 0 of the query operand ([q_pad][512] bf16, or fp16 hi): goes to tensor memory
  // experimental variants only (VAR_SCREEN / VAR_GUARD), appended so that the fields above keep their offsets
  int* ovf;                // screened search: set to 1 when a candidate band overflowed -> the guarded exact pass runs
  float margin;            // screened search: width of the candidate band in score units (2 x error bound)
  const float* qn;         // screened search: [q_pad][512] fp32 normalised queries (exact re-scoring)
  // lock-step producers (qtp > 1): progress[pair] = vault tiles the pair's producer has requested; a producer does not run
  // more than lockstep_w tiles ahead of the slowest pair of its segment (0 = off)
  u32* progress;
  int lockstep_w;
  // guarded exact pass (VAR_GUARD) only: it merges its own candidate lists behind a grid barrier, so that the usual case
  // (no band overflowed) costs ONE launch that returns at once instead of two
  u32* grid_sync;          // zeroed by the preparation kernel
  float* g_scores; long long* g_rows; u64* g_packed; float* g_disc; double g_threshold;
};

// Which tiles a pair works on.  A vault tile is wanted by every group of query tiles (qtp of them), and
// the vault is far bigger than L2, so WHEN the groups read it decides the DRAM traffic: with a plain
// flattened split (the first version) pairs of different groups were at unrelated vault positions and
// ncu showed the 10 GB vault read ~15x per search.  Here pairs are arranged as `seg` segments x qtp
// groups: the qtp pairs of a segment start at the same vault tile and advance at the same rate (same
// work per tile), so one of them misses in L2 and the others hit.  Pairs that do not fit this grid
// (n_pairs - seg*qtp of them) share the tail [v_aligned, v_tiles) of the vault for all groups.
// Every pair gets the same number of tiles (+-1).
struct PairSchedule {
  int n_tiles;   // tiles this pair processes
  int tp0, vt0;  // first tile: query-tile group, vault tile
  int v_lo, v_hi;  // vault tile range it cycles through when moving to the next group
  int sid_base;  // strip id of (this pair, group tp) = sid_base + tp   (unique over the grid)
};
__host__ __device__ inline PairSchedule pair_schedule(const MmaParams& p, int pair, int n_pairs) {
  PairSchedule s;
  if (pair < p.n_aligned) {
    const int t = pair % p.qtp, g = pair / p.qtp;
    const int a = (int)((long long)g * p.v_aligned / p.seg), b = (int)((long long)(g + 1) * p.v_aligned / p.seg);
    s.n_tiles = b - a; s.tp0 = t; s.vt0 = a; s.v_lo = a; s.v_hi = 0x7fffffff; s.sid_base = pair - t;
  } else {
    const int j = pair - p.n_aligned, l = n_pairs - p.n_aligned, vl = p.v_tiles - p.v_aligned;
    const long long ul = (long long)p.qtp * vl;
    const long long a = j * ul / l, b = (j + 1) * ul / l;
    s.n_tiles = (int)(b - a); s.tp0 = vl ? (int)(a / vl) : 0; s.vt0 = p.v_aligned + (vl ? (int)(a % vl) : 0);
    s.v_lo = p.v_aligned; s.v_hi = p.v_tiles; s.sid_base = p.n_aligned + j;
  }
  return s;
}

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// for long waits (epilogue waiting for a whole tile of MMAs): back off between polls -- every poll is a
// shared-memory transaction, and shared-memory bandwidth is what the tensor core's operand fetch needs
__device__ __forceinline__ void mbar_wait_relaxed(u64* bar, u32 parity) {
  u32 done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(256);
  }
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

// ---- thread-block pair (cta_group::2) ------------------------------------------------------
__device__ __forceinline__ u32 cluster_ctarank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ u32 mapa(u32 local, u32 rank) {
  u32 r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
// relaxed on purpose: the barriers signalled this way only order TENSOR-MEMORY accesses (done with
// tcgen05.fence); a release here would drain this thread's outstanding global stores (candidate appends)
// at cluster scope on every tile -- measured at ~25% of all epilogue stall samples.
__device__ __forceinline__ void mbar_arrive_cluster(u32 cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on a barrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* map, u32 bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2(void* dst, const CUtensorMap* map, u32 bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// 2-CTA UMMA (M = 256: 128 rows from each CTA's A operand, each CTA holds half of B), issued by the leader
__device__ __forceinline__ void umma_f16_cg2(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_cg2(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// commit of the leader's MMAs, arriving on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(u64* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}

// one lane of a converged warp (always the same one, so tcgen05.commit sees that lane's MMAs)
__device__ __forceinline__ bool elect_one() {
  u32 pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_ld8(u32 taddr, u32 (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
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

// ---- score histogram (HIST variants) ---------------------------------------------------------
// Bin of a score given its fp32 bits: sign/exponent/top-5-mantissa bits, i.e. 32 bins per octave over
// [2^-7, 2) -> bins 0..255 (MMF_MAX_TOP_K words per query, the bucket pool's memory).  Monotone in the
// score; everything below 2^-7 (and every negative score) falls into bin 0, which carries no bound;
// anything >= 2 - 2^-5 (and NaN, which ranks first) falls into bin 255.
static_assert(MMF_MAX_TOP_K == 256, "the score histogram reuses the 256-word bucket pool");
constexpr int HIST_BIN0 = 120 * 32;      // (bits >> 18) of 2^-7
__host__ __device__ __forceinline__ int hist_bin(u32 score_bits) {
  const int b = ((int)score_bits >> 18) - HIST_BIN0;
  return b < 0 ? 0 : (b > 255 ? 255 : b);
}
__host__ __device__ __forceinline__ float hist_edge(int bin) {   // lower edge of bin >= 1
  const u32 u = (u32)(bin + HIST_BIN0) << 18;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; u32 u; } c; c.u = u; return c.f;
#endif
}

// ---- bucket pool (top_k <= 16) ---------------------------------------------------------------
// bucket of a vault row: a hash of the row id onto [0, top_k).  Every bucket keeps the best score key seen among
// ITS rows (atomicMax), the buckets partition the rows, so the minimum over the top_k bucket maxima is attained
// by top_k distinct rows: a lower bound of the top_k-th best.  With exactly top_k buckets it sits near rank
// top_k * H(top_k) of the rows seen grid-wide (29 for top_k = 10, 11 for top_k = 5; the 16 buckets of round 1: 54).
// Hashed, not row % top_k: the rows of a tile are consecutive, and the order statistics argument wants the
// buckets of a tile's rows to be independent of the tile.
__host__ __device__ __forceinline__ u32 pool_bucket(u32 row, u32 top_k) {
  return (((row * 0x9E3779B1u) >> 16) * top_k) >> 16;
}

// ---- query operand prep --------------------------------------------------------------------
// One warp per padded query row: q / ||q|| (misinfo_forensics.py:439), then the MMA operand
// planes: bf16 (1 plane) or fp16 hi/lo of q*2^8 (2 planes, plane p at row p*q_pad + i).
__device__ __forceinline__ void prep_query_row(const float* __restrict__ q, int n_queries, int q_pad, int split,
                                               void* __restrict__ planes, u32* __restrict__ g_tau, u32* __restrict__ pool,
                                               int top_k, int hist, float* __restrict__ qn) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= q_pad) return;
  if (lane == 0) g_tau[w] = 0;
  const int pool_n = top_k;                       // bucket pool (top_k <= 16): exactly top_k buckets, the rest never win the min
#pragma unroll
  for (int j = 0; j < MMF_MAX_TOP_K / 32; ++j)   // unused buckets never win the min
    pool[(long long)w * MMF_MAX_TOP_K + j * 32 + lane] = (hist || j * 32 + lane < pool_n) ? 0u : 0xFFFFFFFFu;
  float v[MMF_DIM / 32], ss = 0.f;
#pragma unroll
  for (int j = 0; j < MMF_DIM / 32; ++j) {
    v[j] = (w < n_queries) ? q[(long long)w * MMF_DIM + j * 32 + lane] : 0.f;
    ss = fmaf(v[j], v[j], ss);
  }
  const float norm = sqrtf(warp_sum(ss));
  // q / |q| is NaN for an all-zero (or NaN) embedding: every similarity is NaN, np.argsort keeps the row order and the
  // reference returns the top_k HIGHEST row ids with NaN similarity (misinfo_forensics.py:439-450) -- what the
  // streaming kernel's arithmetic yields by itself.  Here such a query is marked (g_tau = the NaN key), gets zero
  // operands, collects no candidates, and the tail kernels write that answer (nan_query_outputs).
  const bool nan_query = w < n_queries && !(norm > 0.f) ;
  if (lane == 0 && nan_query) g_tau[w] = 0xFFFFFFFFu;
#pragma unroll
  for (int j = 0; j < MMF_DIM / 32; ++j) {
    const float x = (w < n_queries && !nan_query) ? v[j] / norm : 0.f;
    const long long o = (long long)w * MMF_DIM + j * 32 + lane;
    if (qn) qn[o] = x;                              // screened search: fp32 copy for the exact re-scoring
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
// The ONE preparation launch of a search: normalised query planes, cleared bounds, cleared candidate counters
// (zero[0..zero_n): the counters of this search and, for the screened search, the overflow flag and the counters
// of the guarded pass) and -- screened search only -- a second set of bounds for the guarded exact pass (bucket
// pool, top_k <= 16), so that a screened search is 4 launches and no memset.
__global__ void __launch_bounds__(256) mma_query_prep_kernel(const float* __restrict__ q, int n_queries, int q_pad,
                                                             int split, void* __restrict__ planes,
                                                             u32* __restrict__ g_tau, u32* __restrict__ pool,
                                                             int top_k, int hist, float* __restrict__ qn,
                                                             int* __restrict__ zero, long long zero_n,
                                                             u32* __restrict__ g_tau2, u32* __restrict__ pool2) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  for (long long i = gtid; i < zero_n; i += nthr) zero[i] = 0;
  const int lane = threadIdx.x & 31;
  const int w = (int)(gtid >> 5);
  if (g_tau2 && w < q_pad) {
    if (lane == 0) g_tau2[w] = 0;
    const int pool_n = top_k;
#pragma unroll
    for (int j = 0; j < MMF_MAX_TOP_K / 32; ++j)
      pool2[(long long)w * MMF_MAX_TOP_K + j * 32 + lane] = (j * 32 + lane < pool_n) ? 0u : 0xFFFFFFFFu;
  }
  prep_query_row(q, n_queries, q_pad, split, planes, g_tau, pool, top_k, hist, qn);
  if (g_tau2 && w < q_pad && lane == 0 && g_tau[w] == 0xFFFFFFFFu) g_tau2[w] = 0xFFFFFFFFu;   // (same lane wrote it)
}

// ---- the search kernel ---------------------------------------------------------------------
// CG = 1: one CTA per MMA (M = 128).  CG = 2: a thread-block PAIR per MMA (tcgen05 cta_group::2,
// M = 256 = two query tiles, one per CTA); each CTA streams and stores only half of every vault
// tile, which halves the shared-memory traffic per flop -- measured, shared-memory bandwidth (TMA
// fill + tensor-core operand fetch = 128 B/clk) is what bounds the CG = 1 kernel.
// KR > 0 (top_k <= KR = 16): grid-wide bound from the BUCKET POOL (pool_bucket above): a thread appends an element
// when it is not below the minimum over its query's top_k bucket maxima.  (Round 1 also kept the KR best values
// of every list sorted in registers; measured in round 2, a list's own k-th best -- of 1/148 of the rows -- never
// beats the pool after the first chunk of a strip, while the sorted insert was half of a candidate event's ~70
// instructions and the events were 70 % of all epilogue instructions.  Removed: 0.308 -> 0.293 ms on C2.)
// KR == 0 bounds the k-th best grid-wide with a per-query HISTOGRAM of candidate scores: the lower edge of the bin
// holding the k-th best counted candidate (bins = 1/32 of an octave: rank <~ 1.4 * top_k; the minimum over top_k
// bucket maxima sits near rank top_k * H(top_k), ~700 for top_k = 100, and let ~7x more elements through).
//
// VAR_SCREEN (fp32-exact vaults, top_k <= 16): ONE f16 pass over the hi planes only (qh.vh: a third of the tensor
// work and half of the HBM bytes of the 3-pass kernel) gives scores within a PROVEN error bound eps of the exact
// ones; every element within 2*eps of the running k-th best is kept, and mma_rerank_kernel re-scores the survivors
// exactly in fp32 from the hi+lo planes and selects the top-k among them -- the result is the exact top-k, not an
// approximation.  If a band does not fit a candidate list, *p.ovf is set and the guarded (VAR_GUARD)
// 3-pass kernel redoes the batch.
// VAR_PARITY (KR > 0 only): how the 8 epilogue warps share the accumulators, see PARITY below.
constexpr int VAR_SCREEN = 2, VAR_GUARD = 4, VAR_PARITY = 8;
struct SelectSmem;
template <int KPL, int CG>
__device__ void merge_query(const MmaParams& p, int n_pairs, int qg, double threshold, float* out_scores, long long* out_rows,
                            u64* out_packed, float* out_disc, SelectSmem& sel, u64* staging, int* slots, int* n_slots);
template <bool SPLIT, int KPL, int CG, int KR, int VAR = 0>
__global__ void __launch_bounds__(MMA_THREADS, 1)
vault_mma_topk_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const MmaParams p) {
  constexpr int STAGES = mma_stages(SPLIT, CG);
  constexpr int STAGE_BYTES = stage_bytes(SPLIT, CG);
  constexpr int KBS = kblk_per_stage(SPLIT);         // k-blocks per stage
  constexpr int PLANE_BYTES = TILE_BYTES / CG;       // one plane of this CTA's share of a B k-block
  constexpr int KB_BYTES = STAGE_BYTES / KBS;        // one k-block (all planes) inside a stage
  constexpr int B_ROWS = TILE_N / CG;
  constexpr int C = 32 * KPL;
  constexpr bool HIST = KR == 0, SCREEN = (VAR & VAR_SCREEN) != 0, GUARD = (VAR & VAR_GUARD) != 0;
  static_assert(!(SCREEN && (SPLIT || KR == 0)), "the screening pass is the 1-plane pipeline with the bucket-pool bound");
  constexpr u32 IDESC = umma_idesc((SPLIT || SCREEN) ? 0u : 1u, TILE_M * CG, TILE_N);   // operands: fp16 planes / bf16
  constexpr u32 QA_COL = 2 * TILE_N;                 // TMEM columns [256,512): plane 0 of the query tile
  // How the 8 epilogue warps share the accumulators.  PARITY: warps 0-3 own buffer 0 (even tiles), warps 4-7
  // buffer 1 -- two tile periods per tile hide the hand-off latencies; right when a tile's MMAs take much longer
  // than its filtering (the 3-pass fp32-exact kernel: 6144 clk of MMA per tile).  Otherwise both warps of a lane
  // quarter work on EVERY tile, alternate 32-column chunks each, so an accumulator is released after half the
  // filtering time -- the MMAs of tile i+2 wait for exactly that.  Measured on the screening pass (2048 clk of MMA
  // per tile): one warp set needs ~2900 clk per tile (a lone warp per scheduler retires an instruction every
  // 4-5 clk), so with PARITY the tensor pipe waited 1400 clk per tile for its accumulator.
  constexpr bool PARITY = KR > 0 && (VAR & VAR_PARITY) != 0;
  constexpr int EMPTY_ARRIVALS = (PARITY ? EPI_WARPS / 2 : EPI_WARPS) * CG;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_smem = smem;                                   // fp32-exact: resident ql tile
  unsigned char* stage_smem = smem + (SPLIT ? Q_RESIDENT_BYTES : 0);
  u64* full_bar = reinterpret_cast<u64*>(stage_smem + STAGES * STAGE_BYTES);
  u64* empty_bar = full_bar + STAGES;
  u64* tmem_full = empty_bar + STAGES;
  u64* tmem_empty = tmem_full + 2;
  u64* q_full = tmem_empty + 2;       // producer -> MMA: resident ql tile landed (fp32-exact)
  u64* q_empty = q_full + 1;          // MMA -> producer: every MMA of the previous strip retired
  u64* qa_full = q_empty + 1;         // epilogue -> MMA: plane 0 written to tensor memory
  u32* tmem_slot = reinterpret_cast<u32*>(qa_full + 1);
  // With CG = 2 the MMA issuer lives in the leader CTA (rank 0): full / tmem_empty / q_full / qa_full
  // are used in the leader only (the peer signals them remotely); empty / tmem_full / q_empty exist
  // in both CTAs and receive the leader's multicast commits.

  if (GUARD && *reinterpret_cast<volatile int*>(p.ovf) == 0) return;   // grid-uniform: the screened search needed no redo

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int pair = blockIdx.x / CG, n_pairs = gridDim.x / CG;
  const PairSchedule sch = pair_schedule(p, pair, n_pairs);

  if (warp == PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, EMPTY_ARRIVALS); }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(qa_full, EPI_WARPS * CG);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {   // the whole tensor memory: accumulators + query operand
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const u32 tmem_base = *tmem_slot;

  if (warp == PRODUCER_WARP) {
    // ===== TMA producer (both CTAs of a pair: each loads its own share) =====
    if (lane == 0) {
      u32 it = 0, strip = 0;
      int cur_tp = -1;
      int tp = sch.tp0, vt = sch.vt0;      // tp: index of the group of CG query tiles
      const u32 q_full_l = (CG == 2) ? mapa(smem_u32(q_full), 0) : smem_u32(q_full);
      // Lock-step with the other pairs of the segment (they sweep the same vault tiles, each for its own group of query
      // tiles): whoever fetches a tile first brings it into L2 for the others -- as long as they stay within an L2's worth
      // of each other.  Candidate events make pairs drift apart (measured on 10 M rows: the vault was read 12x from
      // DRAM), so the leader's producer publishes its position every 16 tiles and waits while it is more than
      // lockstep_w tiles ahead of the slowest pair of its segment.  The slowest never waits, a finished pair publishes
      // "infinity", and the wait is bounded, so a pair that is not running yet cannot hang the others.
      const bool lockstep = leader && p.lockstep_w > 0 && p.qtp > 1 && pair < p.n_aligned;
      volatile u32* seg_progress = p.progress + (pair / max(p.qtp, 1)) * p.qtp;
      for (int u = 0; u < sch.n_tiles; ++u, ++vt) {
        if (lockstep && (u & 15) == 0) {
          p.progress[pair] = (u32)u;
          if (u >= p.lockstep_w) {
            const long long t_end = clock64() + 400000;       // ~0.25 ms: far beyond any healthy drift
            for (;;) {
              u32 mn = 0xFFFFFFFFu;
              for (int t = 0; t < p.qtp; ++t) mn = min(mn, seg_progress[t]);
              if ((u32)u <= mn + (u32)p.lockstep_w || clock64() > t_end) break;
              __nanosleep(500);
            }
          }
        }
        if (vt == sch.v_hi) { vt = sch.v_lo; ++tp; }
        if (SPLIT && tp != cur_tp) {                  // new strip: (re)load this CTA's resident ql tile (plane 1)
          cur_tp = tp;
          mbar_wait(q_empty, (strip & 1) ^ 1);
          if (leader) mbar_expect_tx(q_full, Q_RESIDENT_BYTES * CG);
          const int qrow = p.q_pad + (tp * CG + (int)rank) * TILE_M;
          for (int kb = 0; kb < NUM_KBLK; ++kb) {
            if (CG == 2) tma_load_2d_cg2(q_smem + kb * TILE_BYTES, &tm_a, q_full_l, kb * KBLK, qrow);
            else tma_load_2d(q_smem + kb * TILE_BYTES, &tm_a, q_full, kb * KBLK, qrow);
          }
          ++strip;
        }
        const int brow = vt * TILE_N + (int)rank * B_ROWS;
        for (int kb = 0; kb < NUM_KBLK; kb += KBS, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar + s, ((it / STAGES) & 1) ^ 1);   // (compile-time STAGES: mul-shift, no division)
          if (TRIAGE && (p.debug & 2)) { if (leader) mbar_arrive(full_bar + s); continue; }
          if (leader) mbar_expect_tx(full_bar + s, STAGE_BYTES * CG);
          const u32 fb = (CG == 2) ? mapa(smem_u32(full_bar + s), 0) : 0u;
#pragma unroll
          for (int j = 0; j < KBS; ++j) {
            unsigned char* st = stage_smem + s * STAGE_BYTES + j * KB_BYTES;
            const int kcol = (kb + j) * KBLK;
            if (CG == 2) {
              if (SPLIT) {
                tma_load_3d_cg2(st, &tm_b, fb, kcol, 0, brow);
                tma_load_3d_cg2(st + PLANE_BYTES, &tm_b, fb, kcol, 1, brow);
              } else if (SCREEN) {
                tma_load_3d_cg2(st, &tm_b, fb, kcol, 0, brow);        // hi plane only
              } else {
                tma_load_2d_cg2(st, &tm_b, fb, kcol, brow);
              }
            } else if (SPLIT) {
              tma_load_3d(st, &tm_b, full_bar + s, kcol, 0, brow);
              tma_load_3d(st + PLANE_BYTES, &tm_b, full_bar + s, kcol, 1, brow);
            } else if (SCREEN) {
              tma_load_3d(st, &tm_b, full_bar + s, kcol, 0, brow);
            } else {
              tma_load_2d(st, &tm_b, full_bar + s, kcol, brow);
            }
          }
        }
      }
      if (lockstep) p.progress[pair] = 0x7FFFFFFFu;
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer (leader CTA only) =====
    // The WHOLE warp runs the loop so that addresses and descriptors are computed once, in the
    // uniform datapath; only the tcgen05 instructions themselves are issued by the elected lane.
    if (leader) {
      u32 it = 0, tile = 0, strip = 0;
      int cur_tp = -1;
      const u32 q_addr = smem_u32(q_smem);
      const u32 st_addr = smem_u32(stage_smem);
      const u64 desc_hi = (u64)((1ull << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61));
      const bool dbg = TRIAGE && (p.debug & 8) != 0;
      const long long t_begin = dbg ? clock64() : 0;
      unsigned long long ns_begin = 0;
      if (dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
      long long dbg_empty = 0, dbg_full = 0;
      int tp = sch.tp0, vt = sch.vt0;
      for (int u = 0; u < sch.n_tiles; ++u, ++tile, ++vt) {
        if (vt == sch.v_hi) { vt = sch.v_lo; ++tp; }
        if (tp != cur_tp) {
          cur_tp = tp;
          if (SPLIT) mbar_wait(q_full, strip & 1);
          mbar_wait(qa_full, strip & 1);
          ++strip;
        }
        const u32 acc = tile & 1;
        const long long t_e0 = dbg ? clock64() : 0;
        if (!(TRIAGE && (p.debug & 4))) mbar_wait(tmem_empty + acc, ((tile >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        if (dbg) dbg_empty += clock64() - t_e0;
        const u32 d_tmem = tmem_base + acc * TILE_N;
#pragma unroll 1
        for (int kb0 = 0; kb0 < NUM_KBLK; kb0 += KBS, ++it) {
          const int s = it % STAGES;
          const long long t_f0 = dbg ? clock64() : 0;
          if (!(TRIAGE && (p.debug & 4))) mbar_wait(full_bar + s, (it / STAGES) & 1);
          tcgen05_fence_after();
          if (dbg) dbg_full += clock64() - t_f0;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < KBS; ++j) {
              const int kb = kb0 + j;
              const u64 ql = desc_hi | (u64)(((q_addr + kb * TILE_BYTES) >> 4) & 0x3FFF);
              const u64 vb = desc_hi | (u64)(((st_addr + s * STAGE_BYTES + j * KB_BYTES) >> 4) & 0x3FFF);
              const u32 qa = tmem_base + QA_COL + kb * (KBLK / 2);     // 2 elements per column
#pragma unroll
              for (int k = 0; k < KBLK / 16; ++k) {      // qh.vh  (bf16 mode: q.v); A from tensor memory
                if (CG == 2) umma_f16_ts_cg2(d_tmem, qa + k * 8, vb + 2 * k, IDESC, (kb | k) != 0);
                else umma_f16_ts(d_tmem, qa + k * 8, vb + 2 * k, IDESC, (kb | k) != 0);
              }
              if (SPLIT) {
#pragma unroll
                for (int k = 0; k < KBLK / 16; ++k) {    // qh.vl
                  if (CG == 2) umma_f16_ts_cg2(d_tmem, qa + k * 8, vb + (PLANE_BYTES >> 4) + 2 * k, IDESC, 1);
                  else umma_f16_ts(d_tmem, qa + k * 8, vb + (PLANE_BYTES >> 4) + 2 * k, IDESC, 1);
                }
#pragma unroll
                for (int k = 0; k < KBLK / 16; ++k) {    // ql.vh, ql from shared memory (+2 = 32 B = 16 elements)
                  if (CG == 2) umma_f16_cg2(d_tmem, ql + 2 * k, vb + 2 * k, IDESC, 1);
                  else umma_f16(d_tmem, ql + 2 * k, vb + 2 * k, IDESC, 1);
                }
              }
            }
            if (CG == 2) umma_commit_cg2(empty_bar + s); else umma_commit(empty_bar + s);   // smem slot free once these MMAs retire
            if (kb0 + KBS == NUM_KBLK) {
              if (CG == 2) umma_commit_cg2(tmem_full + acc); else umma_commit(tmem_full + acc);   // accumulator complete
              if (SPLIT && (u + 1 == sch.n_tiles || vt + 1 == sch.v_hi)) {                         // strip done
                if (CG == 2) umma_commit_cg2(q_empty); else umma_commit(q_empty);
              }
            }
          }
          __syncwarp();
        }
      }
      if (dbg && blockIdx.x == 0 && lane == 0 && tile > 0) {
        mbar_wait(tmem_full + ((tile - 1) & 1), ((tile - 1) >> 1) & 1);     // last accumulator complete
        const long long dt = clock64() - t_begin;
        unsigned long long ns_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
        printf("[mmf debug] block 0: %u tiles, %u k-blocks, %lld clk in the MMA loop -> %.1f clk per k-block "
               "(waiting: accumulator free %.1f, operands landed %.1f); %.1f us -> SM clock %.0f MHz\n", tile, it * KBS, dt,
               (double)dt / (it * KBS), (double)dbg_empty / (it * KBS), (double)dbg_full / (it * KBS),
               (double)(ns_end - ns_begin) * 1e-3, (double)dt / ((double)(ns_end - ns_begin) * 1e-3));
      }
    }
  } else {
    // ===== epilogue: thread == query (TMEM lane), streaming top-k =====
    // 8 warps: warp w may touch TMEM lanes 32*(w%4)..+31.  Warps 0-3 serve accumulator buffer 0 (even
    // tiles), warps 4-7 buffer 1 (odd tiles): each set has two tile periods to drain its tile, which
    // absorbs the barrier hand-off latencies and compaction bursts.  Each thread keeps its own
    // candidate list + threshold for its (query, tile parity).
    const int quarter = warp & 3;
    const int half = warp >> 2;                       // tile parity == accumulator buffer served
    const int m = quarter * 32 + lane;
    const u32 lane_base = tmem_base + ((u32)(quarter * 32) << 16);
    const u32 tmem_empty_l = (CG == 2) ? mapa(smem_u32(tmem_empty), 0) : smem_u32(tmem_empty);
    const u32 qa_full_l = (CG == 2) ? mapa(smem_u32(qa_full), 0) : smem_u32(qa_full);
    const int k = p.top_k;
    const float acc_scale = 1.0f / p.inv_scale;
    const float margin_acc = SCREEN ? p.margin * acc_scale : 0.f;   // candidate band in accumulator units
#define MMF_TAU_F (SCREEN ? tau_acc - margin_acc : tau_acc)       /* what the filter compares against */
    float tau_acc = -INFINITY;                        // threshold in accumulator units
    int cnt = 0;
    int cur_tp = -1;
    u64* buf = nullptr;
    int* cnt_out = nullptr;
    u32* g_tau = nullptr;
    u32* pool = nullptr;
    uint4 pool_prev[4];
    bool valid_q = false;
    u32 tile = 0, g_prev = 0;
    int strip_u0 = 0;                                 // first tile of the current strip (warm-up: see the chunk loop)
    const bool dbg = TRIAGE && (p.debug & 8) != 0;
    long long dbg_wait = 0, dbg_filter = 0, dbg_compact = 0;
    int dbg_ncompact = 0;
    int tp = sch.tp0, vt = sch.vt0;
    const int last_cols = (int)(p.n_rows - (long long)(p.v_tiles - 1) * TILE_N);   // valid columns of the LAST vault tile
    for (int u = 0; u < sch.n_tiles; ++u, ++tile, ++vt) {
      if (vt == sch.v_hi) { vt = sch.v_lo; ++tp; }
      if (tp != cur_tp) {                             // new strip: flush the old one, reset state
        if (cur_tp >= 0) {
          *cnt_out = cnt;
          // both warp sets must have drained the old strip (=> all its MMAs retired) before anyone
          // overwrites the query operand in tensor memory
          asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        }
        cur_tp = tp;
        strip_u0 = u;
        cnt = 0;
        g_prev = 0;
        tau_acc = -INFINITY;
        const int qt = tp * CG + (int)rank;           // this CTA's query tile
        const long long list = (((long long)(sch.sid_base + tp) * 2 + half) * CG + rank) * TILE_M + m;
        buf = p.cand + list * C;
        cnt_out = p.cand_cnt + list;
        g_tau = p.g_tau + qt * TILE_M + m;
        pool = p.pool + (long long)(qt * TILE_M + m) * MMF_MAX_TOP_K;
        valid_q = (qt * TILE_M + m) < p.n_queries && *reinterpret_cast<volatile u32*>(g_tau) != 0xFFFFFFFFu;   // (NaN query: see prep)
        if (KR > 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) pool_prev[i] = make_uint4(0, 0, 0, 0);
        }
        {
          // this thread's query row of plane 0 -> its TMEM lane, 2 elements per column (each warp of
          // a quarter writes half of the 256 columns)
          const uint4* src = p.q_plane0 + (long long)(qt * TILE_M + m) * (MMF_DIM * 2 / 16);
#pragma unroll 1
          for (int c = half * 4; c < half * 4 + 4; c += 2) {      // 16 independent 128-bit loads in flight per thread
            u32 w0[32], w1[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 x = __ldg(src + c * 8 + i);
              w0[4 * i] = x.x; w0[4 * i + 1] = x.y; w0[4 * i + 2] = x.z; w0[4 * i + 3] = x.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 x = __ldg(src + (c + 1) * 8 + i);
              w1[4 * i] = x.x; w1[4 * i + 1] = x.y; w1[4 * i + 2] = x.z; w1[4 * i + 3] = x.w;
            }
            tmem_st32(lane_base + QA_COL + c * 32, w0);
            tmem_st32(lane_base + QA_COL + (c + 1) * 32, w1);
          }
          tmem_wait_st();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_cluster(qa_full_l); else mbar_arrive(qa_full); }
        }
      }
      if (PARITY && (int)(tile & 1) != half) continue;   // the other warp set owns this tile
      // a tighter bound found by any other block / warp for this query (valid for every list of it).
      // Software-pipelined: the value loaded during the previous tile is applied now and the next
      // load is issued, so the L2 round trip never sits on the tile's critical path.
      if (g_prev) tau_acc = fmaxf(tau_acc, okey_inv(g_prev) * acc_scale);
      g_prev = *reinterpret_cast<volatile u32*>(g_tau);
      if (KR > 0) {
        // grid-wide bound: every bucket (row % top_k) holds the best score any block has seen among
        // its rows -- top_k distinct rows, so the minimum over the buckets is <= the global k-th best
        u32 mn = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 4; ++i) mn = min(min(mn, pool_prev[i].x), min(min(pool_prev[i].y, pool_prev[i].z), pool_prev[i].w));
        if (mn != 0u && mn != 0xFFFFFFFFu) tau_acc = fmaxf(tau_acc, okey_inv(mn) * acc_scale);
#pragma unroll
        for (int i = 0; i < 4; ++i) pool_prev[i] = __ldcv(reinterpret_cast<const uint4*>(pool) + i);
      }
      const u32 acc = tile & 1;
      const long long t_w0 = dbg ? clock64() : 0;
      mbar_wait(tmem_full + acc, (tile >> 1) & 1);
      tcgen05_fence_after();
      const long long t_w1 = dbg ? clock64() : 0;
      if (dbg) dbg_wait += t_w1 - t_w0;
      const bool partial = vt == p.v_tiles - 1 && last_cols < TILE_N;     // beyond n_cols the tile is TMA zero fill
      const int n_cols = partial ? last_cols : TILE_N;                    // valid columns of this tile
      const u32 row_id0 = p.row_base + (u32)vt * (u32)TILE_N;
      if (KR > 0 && u - strip_u0 < 2 && __any_sync(0xFFFFFFFFu, tau_acc == -INFINITY && valid_q)) {    // (warp-uniform)
        // Warm-up of a strip: no bound yet.  Filtering without one makes every element a candidate event (a store and
        // an atomic, 32 distinct lines each per warp instruction) in every block at once -- measured on the screening
        // pass, 50k of 324k clk.  Instead the grid SEEDS the bucket pool first: each thread publishes the maximum of
        // its part of this tile (one atomic; the tiles of all blocks are distinct rows, and a row always goes to its
        // own bucket), waits (bounded) until every bucket of its query has a value -- all blocks seed at the same
        // time, one maximum per 64 or 128 rows -- and only then filters the tile, which is still in tensor memory,
        // against a bound that is already the ~25th best of the first ~10k rows.
        auto pool_min = [&]() {
          u32 mn = 0xFFFFFFFFu;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 x = __ldcv(reinterpret_cast<const uint4*>(pool) + i);
            mn = min(min(mn, x.x), min(min(x.y, x.z), x.w));
          }
          return mn;
        };
        u32 mn = pool_min();
        // (a vault of a few tiles cannot fill top_k buckets with one seed per tile: it takes its few candidate events instead of the wait)
        if (p.v_tiles >= 8 * k && __any_sync(0xFFFFFFFFu, mn == 0u && valid_q) && !(TRIAGE && (p.debug & 64))) {
          // one seed per thread when the whole grid works on this query (C2: 74 pairs), one per 32-row chunk when only a
          // few pairs do (1 000 queries = 4 groups of 18 pairs: 18 seeds cannot tell much about 10 buckets)
          const bool seed_chunks = n_pairs < 64 * p.qtp;
          float best = -INFINITY;
          int bj = 0;
          auto publish = [&]() {
            if (valid_q && best > -INFINITY) {
              const u32 ub = __float_as_uint((SPLIT || SCREEN) ? best * p.inv_scale : best);
              const u32 row = row_id0 + bj;
              atomicMax(pool + pool_bucket(row, (u32)k), ub ^ ((u32)((int)ub >> 31) | 0x80000000u));
            }
          };
#pragma unroll 1
          for (int c = PARITY ? 0 : half; c < TILE_N / 32; c += PARITY ? 1 : 2) {
            u32 v[32];
            tmem_ld32(lane_base + acc * TILE_N + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = (partial && c * 32 + j >= n_cols) ? -INFINITY : __uint_as_float(v[j]);
              if (a > best) { best = a; bj = c * 32 + j; }
            }
            if (seed_chunks) { publish(); best = -INFINITY; }
          }
          if (!seed_chunks) publish();
          const long long t_end = clock64() + 20000;
          for (;;) {
            mn = pool_min();
            if (mn != 0u || !valid_q || clock64() > t_end) break;
            __nanosleep(200);
          }
          __syncwarp();
        }
        if (mn != 0u && mn != 0xFFFFFFFFu) tau_acc = fmaxf(tau_acc, okey_inv(mn) * acc_scale);
      }
      // The chunks of a tile are software-pipelined over two register buffers: the tcgen05.ld of the next chunk is in
      // flight while the current one is filtered (tcgen05.wait::ld waits for every outstanding load of the thread, so
      // the next load is issued right after the wait for the current one).
      auto refresh_bound = [&](int c) {
        if (KR > 0 && u - strip_u0 < (PARITY ? 4 : 2) && (u != strip_u0 || c >= 2)) {
          // warm-up of a strip: a list's own threshold is still loose (the 10th best of a few dozen rows) while the
          // whole grid has already seen thousands, so re-read the bucket pool before EVERY chunk of the first two
          // tiles instead of once per tile -- 22 + 0.2 candidate events per thread in tile 0 instead of 22 + 7, and
          // a candidate event costs ~70 instructions for the whole warp.  (The loads overlap the tcgen05.ld.)
          u32 mn = 0xFFFFFFFFu;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 x = __ldcv(reinterpret_cast<const uint4*>(pool) + i);
            mn = min(min(mn, x.x), min(min(x.y, x.z), x.w));
          }
          if (mn != 0u && mn != 0xFFFFFFFFu) tau_acc = fmaxf(tau_acc, okey_inv(mn) * acc_scale);
        }
      };
      auto filter_chunk = [&](int c, const u32 (&v)[32]) {
        // fast path (almost always): chunk maximum below the threshold.  A max tree keeps the
        // dependent chain short.  (fmaxf drops a NaN next to a number; vaults with NaN rows never
        // reach this kernel, see mmf_mma_supported.)
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          mx[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                        fmaxf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
        const float m8 = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        if (TRIAGE && (p.debug & 16)) { if (m8 == 12345.678f) cnt = 1; return; }      // triage: tcgen05.ld + max tree only
        if (!(m8 < MMF_TAU_F) && valid_q && !(TRIAGE && (p.debug & 32) && u - strip_u0 >= 2)) {   // triage bit 5: no events after the warm-up tiles
          // Candidate events: rare per chunk, but a lone warp pays ~15 clk for every branch it takes, and the obvious
          // form -- 8 group tests, 4 element tests per hit group, the event body unrolled 32x -- was a dozen branches
          // and ~300 clk per event.  Here: a bit mask of the groups of 4 that hold a candidate (no
          // branches), one indexed jump per hit group to fetch its 4 values, and ONE copy of the event body, fed with
          // the group's maximum (a candidate for sure); the group's other elements are looked at only when more than
          // one of them passes.
          auto emit = [&](float a, int col) {
            if (!partial || col < n_cols) {
              // the key is the order-preserving transform without okey()'s NaN / -0.0 canonicalisation: tensor-core
              // NaNs are positive (they still rank first) and -0.0 only matters for tie order
              const u32 ub = __float_as_uint((SPLIT || SCREEN) ? a * p.inv_scale : a);
              const u32 key = ub ^ ((u32)((int)ub >> 31) | 0x80000000u);
              const u32 row = row_id0 + col;
              buf[cnt++] = ((u64)key << 32) | row;
              if (HIST) {
                const int bin = hist_bin(ub);
                if (bin > 0) atomicAdd(pool + bin, 1u);     // bin 0 (score < 2^-7) carries no bound
              } else {
                atomicMax(pool + pool_bucket(row, (u32)k), key);
              }
            }
          };
          const float tau_f = MMF_TAU_F;
          u32 gm = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) gm |= (!(mx[i] < tau_f) ? 1u : 0u) << i;
#pragma unroll 1
          while (gm) {
            const int i = __ffs(gm) - 1;
            gm &= gm - 1;
            float a0, a1, a2, a3;
            switch (i) {
#define MMF_GROUP(I_) case I_: a0 = __uint_as_float(v[4 * I_]); a1 = __uint_as_float(v[4 * I_ + 1]); \
                               a2 = __uint_as_float(v[4 * I_ + 2]); a3 = __uint_as_float(v[4 * I_ + 3]); break;
              MMF_GROUP(0) MMF_GROUP(1) MMF_GROUP(2) MMF_GROUP(3) MMF_GROUP(4) MMF_GROUP(5) MMF_GROUP(6)
              default: a0 = __uint_as_float(v[28]); a1 = __uint_as_float(v[29]); a2 = __uint_as_float(v[30]); a3 = __uint_as_float(v[31]); break;
#undef MMF_GROUP
            }
            const int cb = c * 32 + 4 * i;
            const bool h0 = !(a0 < tau_f), h1 = !(a1 < tau_f), h2 = !(a2 < tau_f), h3 = !(a3 < tau_f);
            const int e = h0 ? 0 : h1 ? 1 : h2 ? 2 : 3;                       // first candidate of the group (there is one)
            emit(h0 ? a0 : h1 ? a1 : h2 ? a2 : a3, cb + e);
            if ((int)h0 + (int)h1 + (int)h2 + (int)h3 > 1) {                  // rare: the others
              if (h1 && e < 1) emit(a1, cb + 1);
              if (h2 && e < 2) emit(a2, cb + 2);
              if (h3 && e < 3) emit(a3, cb + 3);
            }
          }
        }
            };
      auto release_acc = [&]() {
        // hand the accumulator back as soon as its last column is in registers: the filtering of what has been read
        // overlaps the next MMAs into this buffer
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tmem_empty_l + acc * 8); else mbar_arrive(tmem_empty + acc); }
      };
      if (TRIAGE && (p.debug & 1)) {
        release_acc();
      } else {
        constexpr int C_STEP = PARITY ? 1 : 2, N_CHUNKS = TILE_N / 32;
        const int c_first = PARITY ? 0 : half;
        const u32 t_acc = lane_base + acc * TILE_N;
        u32 va[32], vb[32];
        const bool early = TRIAGE && (p.debug & 128);      // triage (results invalid): hand the accumulator back BEFORE reading it
        if (early) release_acc();
        tmem_ld32(t_acc + c_first * 32, va);
        tmem_ld32(t_acc + (c_first + C_STEP) * 32, vb);
#pragma unroll 1
        for (int c = c_first; c < N_CHUNKS; c += 2 * C_STEP) {
          const bool last = c + 2 * C_STEP >= N_CHUNKS;
          refresh_bound(c);
          tmem_wait_ld();
          if (last && !early) release_acc();
          filter_chunk(c, va);
          if (!last) tmem_ld32(t_acc + (c + 2 * C_STEP) * 32, va);
          refresh_bound(c + C_STEP);
          filter_chunk(c + C_STEP, vb);
          if (!last) tmem_ld32(t_acc + (c + 3 * C_STEP) * 32, vb);
        }
      }
      const long long t_f1 = dbg ? clock64() : 0;
      if (dbg) dbg_filter += t_f1 - t_w1;
      if (KR == 0 && valid_q && (tile < 96 ? (tile & 3) == 3 : (tile & 31) == 31)) {
        // large top_k: refresh the grid-wide bound from the bucket pool -- every 4th tile while the
        // thresholds still move fast (half of all candidate events happen in the first few thousand rows),
        // every 32nd afterwards (off the critical path: the accumulator has been handed back)
        if (HIST) {
          // walk the histogram from the top, 32 bins (8 independent 128-bit loads) at a time, until top_k
          // candidates have been counted: every counted candidate is a distinct vault row with a score >=
          // the lower edge of its bin, so that edge is a lower bound of this query's k-th best
          const uint4* h4 = reinterpret_cast<const uint4*>(pool);
          u32 seen = 0;
          int edge_bin = 0;
#pragma unroll 1
          for (int g = MMF_MAX_TOP_K / 4 - 8; g >= 0 && edge_bin == 0; g -= 8) {
            uint4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __ldcv(h4 + g + i);
#pragma unroll
            for (int i = 7; i >= 0; --i) {
              const u32 c4[4] = {x[i].w, x[i].z, x[i].y, x[i].x};      // bins (g+i)*4+3 ... (g+i)*4
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                seen += c4[e];
                if (edge_bin == 0 && seen >= (u32)k) edge_bin = (g + i) * 4 + 3 - e;
              }
            }
          }
          if (edge_bin > 0) {
            const float edge = hist_edge(edge_bin);
            if (edge * acc_scale > tau_acc) {
              tau_acc = edge * acc_scale;
              atomicMax(g_tau, okey(edge));
            }
          }
        }
      }
      // keep room for one more tile (ROOM appends per thread); warp-cooperative, one list at a time
      constexpr int ROOM = PARITY ? TILE_N : TILE_N / 2;
      static_assert(C - ROOM >= 32, "candidate capacity too small for a tile");
      u32 need = __ballot_sync(FULL, cnt > C - ROOM);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        u64* b = reinterpret_cast<u64*>(__shfl_sync(FULL, reinterpret_cast<u64>(buf), src));
        const int n = __shfl_sync(FULL, cnt, src);
        __syncwarp();
        float t = 0.f;
        bool band_ovf = false;
        const int kept = SCREEN ? warp_compact_band<KPL>(b, n, k, p.margin, C - ROOM, &t, &band_ovf)
                                : warp_compact<KPL>(b, n, k, &t);
        if (SCREEN && band_ovf && lane == 0) *p.ovf = 1;
        if (lane == src) {
          cnt = kept;
          tau_acc = fmaxf(tau_acc, t * acc_scale);
          if (t == t) atomicMax(g_tau, okey(t));
        }
        __syncwarp();
        if (dbg) ++dbg_ncompact;
      }
      if (dbg) dbg_compact += clock64() - t_f1;
    }
#undef MMF_TAU_F
    if (cur_tp >= 0) *cnt_out = cnt;
    if (dbg && blockIdx.x == 0 && lane == 0)
      printf("[mmf debug] epilogue warp %d: %u tiles; clk per OWN tile: wait %.0f, filter %.0f, arrive+compact %.0f; %d compactions, cnt %d\n",
             warp, tile, 2.0 * dbg_wait / tile, 2.0 * dbg_filter / tile, 2.0 * dbg_compact / tile, dbg_ncompact, cnt);
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer's smem / TMEM must outlive the leader's MMAs
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
  if constexpr (GUARD) {
    // The rare path finishes inside this launch: once every block of the grid has written its lists (one block per SM,
    // all resident: a counter in global memory is a grid barrier), the blocks share the queries and merge them, with the
    // staging areas of the merge carved from the now idle stage ring.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(p.grid_sync, 1u);
      while (*reinterpret_cast<volatile u32*>(p.grid_sync) < gridDim.x) __nanosleep(200);
      __threadfence();
    }
    __syncthreads();
    static_assert(sizeof(SelectSmem) <= 16384 && mma_smem_bytes(SPLIT, CG) >= 16384 + 4096 * 8 + MERGE_MAX_SLOTS_BYTES + 16,
                  "the merge's shared memory must fit the stage ring");
    SelectSmem& sel = *reinterpret_cast<SelectSmem*>(smem);
    u64* staging = reinterpret_cast<u64*>(smem + 16384);
    int* slots = reinterpret_cast<int*>(smem + 16384 + 4096 * 8);
    int* n_slots = reinterpret_cast<int*>(smem + 16384 + 4096 * 8 + MERGE_MAX_SLOTS_BYTES);
    for (int qg = blockIdx.x; qg < p.n_queries; qg += gridDim.x) {
      merge_query<KPL, CG>(p, n_pairs, qg, p.g_threshold, p.g_scores, p.g_rows, p.g_packed, p.g_disc, sel, staging, slots, n_slots);
      __syncthreads();
    }
  }
}

// Slots (strip * 2 + column half) of the candidate lists that cover query-tile group `tp`: every pair whose
// schedule touches the group contributes its strip.  One pair per thread; *n_slots must be 0 (and a barrier
// passed) on entry; may end up > MERGE_MAX_SLOTS, only the first MERGE_MAX_SLOTS entries are written.  The order of
// the slots depends on the schedule of the atomics; the selection that follows does not.
constexpr int MERGE_MAX_SLOTS = 4 * 160;
static_assert(MERGE_MAX_SLOTS * 4 == MERGE_MAX_SLOTS_BYTES, "keep in sync");
__device__ __forceinline__ void gather_slots_par(const MmaParams& p, int n_pairs, int tp, int* slots, int* n_slots) {
  for (int c = threadIdx.x; c < n_pairs; c += blockDim.x) {
    const PairSchedule sc = pair_schedule(p, c, n_pairs);
    if (sc.n_tiles <= 0) continue;
    bool touches;
    if (c < p.n_aligned) {
      touches = sc.tp0 == tp;
    } else {
      const int vl = sc.v_hi - sc.v_lo;
      const long long first = (long long)sc.tp0 * vl + (sc.vt0 - sc.v_lo), last = first + sc.n_tiles - 1;
      touches = vl > 0 && first / vl <= tp && tp <= last / vl;
    }
    if (touches) {
      const int pos = atomicAdd(n_slots, 2);
      if (pos + 2 <= MERGE_MAX_SLOTS) {
        slots[pos] = (sc.sid_base + tp) * 2;
        slots[pos + 1] = (sc.sid_base + tp) * 2 + 1;
      }
    }
  }
}

// Shared front half of the tail kernels: the candidate lists of query `qg` as a CandidateLists view.
// slots / n_slots: shared memory of the calling block.  All threads must call it.
template <int KPL, int CG>
__device__ __forceinline__ CandidateLists lists_of_query(const MmaParams& p, int n_pairs, int qg, int* slots, int* n_slots) {
  constexpr int C = 32 * KPL;
  const int qt = qg / TILE_M, m = qg % TILE_M;
  const int tp = qt / CG, r = qt % CG;
  if (threadIdx.x == 0) *n_slots = 0;
  __syncthreads();
  gather_slots_par(p, n_pairs, tp, slots, n_slots);
  __syncthreads();
  // lists of this query: [strip][half][r][m][C]; slot = strip*2 + half selects a block of CG*TILE_M lists
  CandidateLists src;
  const long long base = (long long)r * TILE_M + m;
  src.lists = p.cand + base * C;
  src.counts = p.cand_cnt + base;
  src.n_lists = min(*n_slots, MERGE_MAX_SLOTS);
  src.k_in = C;
  src.list_stride = (long long)CG * TILE_M * C;
  src.count_stride = CG * TILE_M;
  src.slots = slots;
  return src;
}

// The answer for a NaN query (see prep_query_row): the top_k highest row ids of the shard with NaN similarity,
// discrepancy 0; slots beyond the shard size are empty.  All threads of the block call it.
__device__ __forceinline__ void nan_query_outputs(const MmaParams& p, int qg, float* out_scores, long long* out_rows,
                                                  u64* out_packed, float* out_disc) {
  for (int i = threadIdx.x; i < p.top_k; i += blockDim.x) {
    const bool have = i < p.n_rows;
    const u32 row = p.row_base + (u32)(p.n_rows - 1 - i);
    const long long o = (long long)qg * p.top_k + i;
    if (out_packed) out_packed[o] = have ? ((0xFFFFFFFFull << 32) | row) : 0ull;
    if (out_scores) out_scores[o] = __int_as_float(0x7FC00000);
    if (out_rows) out_rows[o] = have ? (long long)row : -1ll;
  }
  if (threadIdx.x == 0 && out_disc) out_disc[qg] = 0.f;
}

// Best published lower bound (score key) of query qg's top_k-th best: g_tau, and for top_k <= 16 (bucket-pool kernels)
// the minimum over the 16 bucket maxima -- 16 distinct rows score at least that.  0 = no bound.
__device__ __forceinline__ u32 published_bound(const MmaParams& p, int qg) {
  u32 g = p.g_tau[qg];
  if (p.top_k <= 16) {
    u32 mn = 0xFFFFFFFFu;
    for (int j = 0; j < 16; ++j) mn = min(mn, p.pool[(long long)qg * MMF_MAX_TOP_K + j]);   // broadcast loads
    g = max(g, mn);                                                                          // (an empty bucket is 0)
  }
  return g;
}

// sm.win[0..top_k) := the top_k best of the staged keys (n of them; in staging[] if n <= cap, else the general path
// re-reads the lists).  Rank-by-counting when few were staged (the usual case), radix select otherwise.
__device__ __forceinline__ void select_staged(const CandidateLists& src, u32 n, u64* staging, int cap, u64 min_key, int top_k,
                                              SelectSmem& sel) {
  if (n <= (u32)RANK_SELECT_MAX) {
    block_rank_select(staging, (int)n, top_k, sel);
  } else if (n <= (u32)cap) {         // staged: radix select over the staged keys, read in place
    CandidateLists ex;
    ex.lists = staging; ex.counts = nullptr; ex.n_lists = 1; ex.k_in = (int)n; ex.list_stride = 0; ex.count_stride = 0;
    block_select_topk(ex, top_k, sel, staging, 0, 0ull, nullptr, nullptr, nullptr, nullptr, 0.0);
  } else {                            // does not fit: the general path re-reads global memory
    block_select_topk(src, top_k, sel, staging, cap, min_key, nullptr, nullptr, nullptr, nullptr, 0.0);
  }
}

// One block per query: gather the strips that cover its query tile, select + sort the top-k.
// GUARD: runs only when the screened search flagged an overflow (*p.ovf != 0).
// PUSH: row-sharded search over peer memory (exchange.cu): instead of writing outputs, the top_k winners go straight
// from shared memory into slot [rank] of EVERY rank's gather buffer (the own one included) -- the NVLink transfer of
// query i overlaps the selection of the other queries -- and the last block publishes this rank's epoch flag.
// One query's candidate lists -> its sorted top_k in the caller's outputs (all threads of the block call it; any block size).
template <int KPL, int CG>
__device__ void merge_query(const MmaParams& p, int n_pairs, int qg, double threshold, float* out_scores, long long* out_rows,
                            u64* out_packed, float* out_disc, SelectSmem& sel, u64* staging, int* slots, int* n_slots) {
  if (p.g_tau[qg] == 0xFFFFFFFFu) {                  // NaN query (block-uniform): the top_k highest row ids, NaN keys
    nan_query_outputs(p, qg, out_scores, out_rows, out_packed, out_disc);
    return;
  }
  const CandidateLists src = lists_of_query<KPL, CG>(p, n_pairs, qg, slots, n_slots);
  const u64 min_key = (u64)published_bound(p, qg) << 32;
  const u32 n = stage_candidates(src, sel, staging, 4096, min_key);
  select_staged(src, n, staging, 4096, min_key, p.top_k, sel);
  write_topk_outputs(sel, p.top_k, out_scores ? out_scores + (long long)qg * p.top_k : nullptr,
                     out_rows ? out_rows + (long long)qg * p.top_k : nullptr,
                     out_packed ? out_packed + (long long)qg * p.top_k : nullptr, out_disc ? out_disc + qg : nullptr, threshold);
}

template <int KPL, int CG, bool PUSH>
__global__ void __launch_bounds__(256) mma_merge_kernel(const MmaParams p, int n_pairs, double threshold,
                                                        float* out_scores, long long* out_rows, u64* out_packed,
                                                        float* out_disc, const mmf_push_ctx px) {
  __shared__ SelectSmem sel;
  __shared__ u64 staging[4096];
  __shared__ int slots[MERGE_MAX_SLOTS];  // (strip, half) slots holding lists of this query's tile group
  __shared__ int n_slots;
  __shared__ bool last;
  const int qg = blockIdx.x;
  if constexpr (!PUSH) {
    merge_query<KPL, CG>(p, n_pairs, qg, threshold, out_scores, out_rows, out_packed, out_disc, sel, staging, slots, &n_slots);
  } else {
    if (p.g_tau[qg] == 0xFFFFFFFFu) {                // NaN query (block-uniform): the top_k highest row ids, NaN keys
      for (int i = threadIdx.x; i < p.top_k; i += blockDim.x)
        sel.win[i] = i < p.n_rows ? ((0xFFFFFFFFull << 32) | (p.row_base + (u32)(p.n_rows - 1 - i))) : 0ull;
      __syncthreads();
    } else {
      const CandidateLists src = lists_of_query<KPL, CG>(p, n_pairs, qg, slots, &n_slots);
      const u64 min_key = (u64)published_bound(p, qg) << 32;
      const u32 n = stage_candidates(src, sel, staging, 4096, min_key);
      select_staged(src, n, staging, 4096, min_key, p.top_k, sel);
    }
    // sel.win[0..top_k): this query's winners, sorted (0 = empty).  Push them to every rank.
    for (int i = threadIdx.x; i < p.top_k * px.world; i += blockDim.x) {
      const int rr = i / p.top_k, j = i - rr * p.top_k;
      const int peer = (px.rank + rr) % px.world;                  // own copy first, then round the ring
      reinterpret_cast<u64*>(px.base[peer] + px.slot_off)[(long long)qg * p.top_k + j] = sel.win[j];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const u32 ticket = atomicAdd(px.done, 1u);
      last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *px.done = 0;
    __threadfence_system();
    if ((int)threadIdx.x < px.world)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<u32*>(px.base[threadIdx.x]) +
                                                              px.parity * MMF_XCHG_MAX_WORLD + px.rank), "r"(px.epoch) : "memory");
  }
}

// Screened search, second half (VAR_SCREEN): one block per query.
//   1. stage every candidate of the query whose APPROXIMATE score lies within the band below the grid-wide
//      bound, and select the approximate top-k among them: its k-th entry T is the exact k-th best approximate
//      score of the whole shard;
//   2. every staged candidate with score >= T - margin is re-scored EXACTLY: fp32 dot of the fp32-normalised
//      query with hi + lo of the vault row -- the arithmetic of the streaming kernel (vault_stream.cu), so the
//      reported scores are bit-identical to a batch-1 search; the others are dropped;
//   3. the exact keys are selected + sorted like any other candidate list.
// Why this is the exact top-k: |approx - exact| <= eps = margin / 2 for every row.  k rows have approx >= T,
// hence exact >= T - eps, so the exact k-th best is >= T - eps, and a row of the exact top-k has
// approx >= T - 2 eps -- it is among the re-scored ones.
// The staging bound is the better of g_tau (the best per-list k-th best) and the minimum over the bucket pool (a
// grid-wide bound near rank 54 for top_k = 10): with g_tau alone ~700 candidates per query were staged and the
// selection ran 8 radix passes; with the pool a few dozen are, and rank-by-counting does it in two barriers.
// ROWS: vault rows in flight per warp while re-scoring -- 4 when the grid is at most two waves of blocks (latency is what
// counts: C2, 256 queries), 2 otherwise (64 registers instead of 109: twice the blocks per SM; C1, 1 000 queries).
template <int KPL, int CG, int ROWS>
__global__ void __launch_bounds__(256, ROWS == 2 ? 4 : 2) mma_rerank_kernel(const MmaParams p, int n_pairs, double threshold,
                                                         const uint4* __restrict__ vault, float* out_scores,
                                                         long long* out_rows, u64* out_packed, float* out_disc) {
  constexpr int STAGING = 4096;
  __shared__ SelectSmem sel;
  __shared__ u64 staging[STAGING];
  __shared__ int slots[MERGE_MAX_SLOTS];
  __shared__ int n_slots;
  __shared__ u32 warp_cnt[8];
  const int qg = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p.g_tau[qg] == 0xFFFFFFFFu) { nan_query_outputs(p, qg, out_scores, out_rows, out_packed, out_disc); return; }
  // this lane's 16 elements of the normalised query: 8*lane..+7 and 256+8*lane..+7 (as vault_stream.cu); loaded first,
  // the latency hides behind the staging of the candidates
  float q[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) q[e] = p.qn[(long long)qg * MMF_DIM + (e >> 3) * 256 + lane * 8 + (e & 7)];
  // bounds on the k-th best APPROXIMATE score; candidates down to margin below it may matter
  const u32 g = published_bound(p, qg);
  const u64 min_key = g ? (u64)okey(okey_inv(g) - p.margin) << 32 : 0ull;
  const CandidateLists src = lists_of_query<KPL, CG>(p, n_pairs, qg, slots, &n_slots);
  const u32 n_staged = stage_candidates(src, sel, staging, STAGING, min_key);
  if (n_staged > (u32)STAGING || n_slots >= MERGE_MAX_SLOTS) {   // band too wide to stage: exact redo of the batch
    if (tid == 0) *p.ovf = 1;
    return;
  }
  select_staged(src, n_staged, staging, STAGING, min_key, p.top_k, sel);      // approximate top-k of the staged candidates
  const u64 kth = sel.win[p.top_k - 1];                          // 0: fewer than top_k candidates -> keep all
  const u64 cut = kth ? (u64)okey(okey_inv((u32)(kth >> 32)) - p.margin) << 32 : 0ull;
  __syncthreads();

  if (TRIAGE && (p.debug & 8) && tid == 0 && (qg & 127) == 0)
    printf("[mmf debug] rerank query %d: %d list slots, %u candidates staged, bound %.4f, approximate k-th best %.4f\n", qg, n_slots,
           n_staged, g ? okey_inv(g) : 0.f, kth ? okey_inv((u32)(kth >> 32)) : 0.f);
  // Compact the candidates inside the band to the front of the staging array (in place, 256 at a time: a round's
  // reads are done before its writes, which land below the round's end), so that the 8 warps share the rows to
  // re-score evenly -- each row is a 2 KB read from HBM, latency-bound.
  u32 n_band = 0;
  for (u32 base = 0; base < n_staged; base += 256) {
    const u32 i = base + tid;
    const u64 key = i < n_staged ? staging[i] : 0ull;
    const bool in = key != 0ull && key >= cut;
    const u32 bal = __ballot_sync(FULL, in);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    u32 off = n_band, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const u32 c = warp_cnt[w];
      if (w < warp) off += c;
      tot += c;
    }
    if (in) staging[off + __popc(bal & ((1u << lane) - 1u))] = key;
    n_band += tot;
    __syncthreads();
  }

  // exact score of a vault row against this query: same element order and reduction tree as vault_stream.cu
  auto load_row = [&](u32 row, uint4 (&ld)[4]) {
    const uint4* rp = vault + (long long)(row - p.row_base) * 128;   // [hi 64 x uint4 | lo 64 x uint4]
#pragma unroll
    for (int c = 0; c < 4; ++c) ld[c] = __ldg(rp + c * 32 + lane);
  };
  auto exact_key = [&](u32 row, const uint4 (&ld)[4]) {
    float v[16];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const u32 hx[4] = {ld[c].x, ld[c].y, ld[c].z, ld[c].w};
      const u32 lx[4] = {ld[c + 2].x, ld[c + 2].y, ld[c + 2].z, ld[c + 2].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hx[j]));
        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lx[j]));
        v[c * 8 + 2 * j] = hf.x + lf.x;                        // exact: hi + lo fits 24 bits
        v[c * 8 + 2 * j + 1] = hf.y + lf.y;
      }
    }
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) a = fmaf(v[e], q[e], a);
    a = warp_sum(a);
    return pack_key(a * MMF_SPLIT_INV_SCALE, row);
  };
  for (u32 base = warp; base < n_band; base += 8 * ROWS) {       // up to ROWS rows in flight per warp: with 4, one HBM round
    uint4 ld[ROWS][4];                                           // trip for the usual band of 2-3 dozen rows
    u32 row[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const u32 i = base + 8 * j;                                // (warp-uniform)
      row[j] = i < n_band ? (u32)staging[i] : 0u;
      if (i < n_band) load_row(row[j], ld[j]);
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const u32 i = base + 8 * j;
      if (i < n_band) {
        const u64 e = exact_key(row[j], ld[j]);
        if (lane == 0) staging[i] = e;
      }
    }
  }
  __syncthreads();
  // select + sort the exact keys: one list in shared memory, read in place
  if (n_band <= (u32)RANK_SELECT_MAX) {
    block_rank_select(staging, (int)n_band, p.top_k, sel);
  } else {
    CandidateLists ex;
    ex.lists = staging; ex.counts = nullptr; ex.n_lists = 1; ex.k_in = (int)n_band; ex.list_stride = 0; ex.count_stride = 0;
    block_select_topk(ex, p.top_k, sel, staging, 0, 0ull, nullptr, nullptr, nullptr, nullptr, 0.0);
  }
  write_topk_outputs(sel, p.top_k, out_scores ? out_scores + (long long)qg * p.top_k : nullptr,
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
  CUtensorMap tm_vault[2];        // [0]: box of 128 rows (CG = 1), [1]: box of 64 rows (CG = 2)
  bool vault_map_ok = false;
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
  bool ok = true;
  for (int cg = 1; cg <= 2; ++cg) {
    if (h->vault_mode == MMF_VAULT_BF16) {
      const cuuint64_t dims[2] = {MMF_DIM, (cuuint64_t)h->vault_rows};
      const cuuint64_t strides[1] = {MMF_DIM * 2};
      const cuuint32_t box[2] = {KBLK, (cuuint32_t)(TILE_N / cg)};
      ok = ok && encode_map(s, &s->tm_vault[cg - 1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->vault, dims, strides, box);
    } else {
      // [row][plane][512] fp16 viewed as (k, plane, row)
      const cuuint64_t dims[3] = {MMF_DIM, 2, (cuuint64_t)h->vault_rows};
      const cuuint64_t strides[2] = {MMF_DIM * 2, MMF_DIM * 4};
      const cuuint32_t box[3] = {KBLK, 1, (cuuint32_t)(TILE_N / cg)};
      ok = ok && encode_map(s, &s->tm_vault[cg - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, h->vault, dims, strides, box);
    }
  }
  s->vault_map_ok = ok;
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

// Work decomposition of one search: CTAs per MMA, padded query tiles, vault tiles, schedule.
static void mma_plan(int64_t n_queries, int64_t n_rows, int sm_count, MmaParams& p, int& cg, int& n_pairs,
                     int force_cg = 0, bool flat = false) {
  // a thread-block pair per MMA as soon as there are two query tiles to pair up
  cg = n_queries > TILE_M ? 2 : 1;
  if (force_cg == 1 || force_cg == 2) cg = force_cg;
  p.n_queries = (int)n_queries;
  p.q_tiles = (int)((n_queries + TILE_M - 1) / TILE_M);
  p.q_tiles = (p.q_tiles + cg - 1) / cg * cg;
  p.q_pad = p.q_tiles * TILE_M;
  p.n_rows = n_rows;
  p.v_tiles = (int)((n_rows + TILE_N - 1) / TILE_N);
  p.qtp = p.q_tiles / cg;
  const long long units = (long long)p.qtp * p.v_tiles;
  n_pairs = (int)std::max<long long>(1, std::min<long long>(sm_count / cg, units));
  // L2-aware schedule: seg x qtp pairs sweep the vault in lock-step, the rest share its tail
  p.seg = n_pairs / p.qtp;
  p.n_aligned = p.seg * p.qtp;
  p.v_aligned = p.n_aligned == n_pairs ? p.v_tiles
                                       : (int)(((long long)p.n_aligned * p.v_tiles + n_pairs / 2) / n_pairs);
  if (p.seg == 0) { p.seg = 1; p.n_aligned = 0; p.v_aligned = 0; }
  if (flat) { p.seg = 1; p.n_aligned = 0; p.v_aligned = 0; }
}

// Host-only self check of the decomposition (no GPU needed): every (query-tile group, vault tile) unit must
// be scheduled exactly once, strip ids must be unique and inside the candidate-list allocation.
extern "C" int mmf_mma_plan_check(int64_t n_queries, int64_t n_rows, int sm_count, int64_t* out_units,
                                  int* out_pairs, int* out_cg) {
  if (n_queries <= 0 || n_rows <= 0 || sm_count <= 0) return MMF_ERR_BAD_ARG;
  MmaParams p;
  int cg, n_pairs;
  mma_plan(n_queries, n_rows, sm_count, p, cg, n_pairs);
  const long long units = (long long)p.qtp * p.v_tiles;
  if (out_units) *out_units = units;
  if (out_pairs) *out_pairs = n_pairs;
  if (out_cg) *out_cg = cg;
  if (units > (1ll << 26)) return MMF_ERR_UNSUPPORTED;          // keep the check itself cheap
  std::vector<unsigned char> seen((size_t)units, 0);
  std::vector<int> strip_owner((size_t)(n_pairs + p.qtp), -1);
  int lo_tiles = 0x7fffffff, hi_tiles = 0;
  for (int c = 0; c < n_pairs; ++c) {
    const PairSchedule s = pair_schedule(p, c, n_pairs);
    if (s.n_tiles < 0) return MMF_ERR_CUDA;
    lo_tiles = std::min(lo_tiles, s.n_tiles);
    hi_tiles = std::max(hi_tiles, s.n_tiles);
    int tp = s.tp0, vt = s.vt0, last_tp = -1;
    for (int u = 0; u < s.n_tiles; ++u, ++vt) {
      if (vt == s.v_hi) { vt = s.v_lo; ++tp; }
      if (tp < 0 || tp >= p.qtp || vt < 0 || vt >= p.v_tiles) return MMF_ERR_CUDA;
      unsigned char& cell = seen[(size_t)tp * p.v_tiles + vt];
      if (cell) return MMF_ERR_CUDA;                              // scheduled twice
      cell = 1;
      if (tp != last_tp) {
        const int sid = s.sid_base + tp;
        if (sid < 0 || sid >= n_pairs + p.qtp || (strip_owner[sid] != -1 && strip_owner[sid] != c)) return MMF_ERR_CUDA;
        strip_owner[sid] = c;
        last_tp = tp;
      }
    }
  }
  for (unsigned char v : seen)
    if (!v) return MMF_ERR_CUDA;                                  // a unit nobody processes
  if (hi_tiles - lo_tiles > 1 + p.qtp) return MMF_ERR_UNSUPPORTED;   // badly balanced
  return MMF_OK;
}

// Host-only model of the HIST bound (same hist_bin / hist_edge as the kernel): the lower bound of the
// top_k-th best of `scores` that the histogram yields, -inf when it yields none.  Used by the CPU test-suite
// to check validity (bound <= k-th best) and tightness (rank of the bound) without a GPU.
extern "C" int mmf_mma_hist_bound(const float* scores, int64_t n, int top_k, float* out_bound) {
  if (!scores || n < 0 || top_k < 1 || top_k > MMF_MAX_TOP_K || !out_bound) return MMF_ERR_BAD_ARG;
  std::vector<u32> hist(MMF_MAX_TOP_K, 0u);
  for (int64_t i = 0; i < n; ++i) {
    union { float f; u32 u; } c;
    c.f = scores[i];
    const int bin = hist_bin(c.u);
    if (bin > 0) hist[bin]++;
  }
  u32 seen = 0;
  int edge_bin = 0;
  for (int b = MMF_MAX_TOP_K - 1; b >= 0 && edge_bin == 0; --b) {
    seen += hist[b];
    if (seen >= (u32)top_k) edge_bin = b;
  }
  *out_bound = edge_bin > 0 ? hist_edge(edge_bin) : -INFINITY;
  return MMF_OK;
}

// The error bound the screened search uses (score units), for the CPU test of the argument in DESIGN.md section 9.
constexpr float SCREEN_EPS = 1.05e-3f;
extern "C" double mmf_mma_screen_eps(void) { return (double)SCREEN_EPS; }

template <bool SPLIT, int KPL, int CG, int KR, int VAR = 0>
static int launch_mma(mmf_handle* h, MmaState* s, const CUtensorMap& tm_q, const MmaParams& p, int n_pairs,
                      double threshold, float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc,
                      cudaStream_t st) {
  const int smem = mma_smem_bytes(SPLIT, CG) + 256 + 1024;
  auto kern = vault_mma_topk_kernel<SPLIT, KPL, CG, KR, VAR>;
  static unsigned long long attr_set = 0;     // per instantiation and device: the attribute sticks to the function
  if (h->device >= 64 || !((attr_set >> h->device) & 1ull)) {
    MMF_CUDA_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (h->device < 64) attr_set |= 1ull << h->device;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_pairs * CG));
  cfg.blockDim = dim3(MMA_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MmaParams pk = p;
  if constexpr ((VAR & VAR_GUARD) != 0) {       // the guarded kernel writes the search's outputs itself
    pk.g_scores = out_scores; pk.g_rows = (long long*)out_rows; pk.g_packed = (u64*)out_packed; pk.g_disc = out_disc;
    pk.g_threshold = threshold;
  }
  MMF_CUDA_OK(h, cudaLaunchKernelEx(&cfg, kern, tm_q, s->tm_vault[CG - 1], pk));
  h->launches++;
  mmf_push_ctx no_push = {};
  if constexpr ((VAR & VAR_SCREEN) != 0) {
    if (p.n_queries <= 4 * h->sm_count)
      mma_rerank_kernel<KPL, CG, 4><<<p.n_queries, 256, 0, st>>>(p, n_pairs, threshold, (const uint4*)h->vault, out_scores,
                                                                  (long long*)out_rows, (u64*)out_packed, out_disc);
    else
      mma_rerank_kernel<KPL, CG, 2><<<p.n_queries, 256, 0, st>>>(p, n_pairs, threshold, (const uint4*)h->vault, out_scores,
                                                                  (long long*)out_rows, (u64*)out_packed, out_disc);
  } else if constexpr ((VAR & VAR_GUARD) != 0) {
    h->launches--;       // (counted below) the guarded kernel merges its own lists: no second launch
  } else if (h->push_ctx && out_packed && !out_scores && !out_rows && !out_disc) {
    // row-sharded search over peer memory: the winners go to every rank's gather buffer from this kernel
    mma_merge_kernel<KPL, CG, true><<<p.n_queries, 256, 0, st>>>(p, n_pairs, threshold, nullptr, nullptr, nullptr,
                                                                         nullptr, *h->push_ctx);
    h->push_fused = true;
  } else {
    mma_merge_kernel<KPL, CG, false><<<p.n_queries, 256, 0, st>>>(p, n_pairs, threshold, out_scores, (long long*)out_rows,
                                                                          (u64*)out_packed, out_disc, no_push);
  }
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

// Error bound of the screening pass (VAR_SCREEN), in score units.  With x*2^8 = hi + lo + r (fp16 hi/lo split,
// |lo| <= 2^-11 |hi|, |r| <= 2^-11 |lo|), unit rows and unit queries:
//   |qh.vh * 2^-16 - q.v| <= (|ql+rq| |vh| + |qh| |vl+rv| + |ql+rq| |vl+rv|) * 2^-16   (Cauchy-Schwarz)
//                         <= 2 * 2^-11 * (1 + 2^-10) + 2^-22 < 9.78e-4,
// plus the tensor core's fp32 accumulation error over 32 K-steps (measured 2.5e-6, budgeted 2e-5) and the fp32
// rounding of the exact re-scoring (~2e-7).  Measured worst case on random and clustered vaults: 1.0e-4.
// (SCREEN_EPS itself is defined above launch_mma.)

int mmf_mma_search(mmf_handle* h, const float* queries, int64_t n_queries, int top_k, double threshold,
                   float* out_scores, int64_t* out_rows, uint64_t* out_packed, float* out_disc, cudaStream_t st) {
  MmaState* s = state_of(h);
  if (!s || !mmf_mma_supported(h, n_queries, top_k))
    return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "tcgen05 path unavailable (TMA descriptor encode failed?)");
  const bool split = h->vault_mode == MMF_VAULT_FP32;
  const int npl = split ? 2 : 1;
  // a list keeps room for a whole tile (128 appends) on top of its top_k, and compaction should be
  // rare: C = 256 up to k = 64, else 512 (k = 100: 284 appends between compactions)
  const int kpl = top_k <= 64 ? 8 : 16;
  const int C = 32 * kpl;

  MmaParams p;
  p.debug = h->opt.debug;
  int cg, n_pairs;
  mma_plan(n_queries, h->vault_rows, h->sm_count, p, cg, n_pairs, h->opt.force_cg, h->opt.flat_schedule != 0);
  p.row_base = (u32)h->vault_row_offset;
  p.top_k = top_k;
  p.inv_scale = split ? (MMF_SPLIT_INV_SCALE * MMF_SPLIT_INV_SCALE) : 1.0f;
  const long long strips = (long long)n_pairs + p.qtp;
  const long long lists = strips * 2 * cg * TILE_M;
  // fp32-exact vaults, top_k <= 16: screened search (VAR_SCREEN); option "screen" = 0 selects the 3-pass kernel
  const bool screen = split && top_k <= 16 && h->opt.screen != 0;
  const bool hist = top_k > 16;                   // KR == 0 variants: score histogram instead of the bucket pool

  // scratch: [64 KB counters (stream kernel) | query planes | g_tau | pool | cand_cnt | (screened: overflow flag 1 KB |
  //           cand_cnt of the guarded pass | its g_tau | its pool | fp32 queries) | cand]
  auto al = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  const size_t off_q = 65536;
  const size_t off_tau = off_q + al((size_t)npl * p.q_pad * MMF_DIM * 2);
  const size_t off_pool = off_tau + al((size_t)p.q_pad * 4);
  const size_t off_cnt = off_pool + al((size_t)p.q_pad * MMF_MAX_TOP_K * 4);
  const size_t off_prog = off_cnt + al((size_t)lists * 4);                 // producers' progress words (lock-step), cleared with
  const size_t off_flag = off_prog + 1024;                                 // cand_cnt: both are contiguous with it
  const size_t off_cnt2 = off_flag + (screen ? 1024 : 0);
  const size_t off_tau2 = off_cnt2 + (screen ? al((size_t)lists * 4) : 0);
  const size_t off_pool2 = off_tau2 + (screen ? al((size_t)p.q_pad * 4) : 0);
  const size_t off_qn = off_pool2 + (screen ? al((size_t)p.q_pad * MMF_MAX_TOP_K * 4) : 0);
  const size_t off_cand = off_qn + (screen ? al((size_t)p.q_pad * MMF_DIM * 4) : 0);
  const size_t total = off_cand + (size_t)lists * C * 8;
  int rc = mmf_ensure_scratch(h, total, st);
  if (rc != MMF_OK) return rc;
  char* sc = (char*)h->scratch();
  void* planes = sc + off_q;
  p.cand_cnt = (int*)(sc + off_cnt);
  p.cand = (u64*)(sc + off_cand);
  p.g_tau = (u32*)(sc + off_tau);
  p.pool = (u32*)(sc + off_pool);
  p.q_plane0 = reinterpret_cast<const uint4*>(planes);
  p.ovf = screen ? (int*)(sc + off_flag) : nullptr;
  p.progress = (u32*)(sc + off_prog);
  p.grid_sync = p.progress + 255;                // (same cleared kilobyte: at most 148 progress words are in use)
  p.g_scores = nullptr; p.g_rows = nullptr; p.g_packed = nullptr; p.g_disc = nullptr; p.g_threshold = 0.0;
  // window: a quarter of L2 per segment-and-then-some (a bf16 / fp16-hi tile is 128 KB); option "lockstep" = 0 turns it off
  p.lockstep_w = h->opt.lockstep ? 64 : 0;
  p.margin = screen ? 2.0f * SCREEN_EPS : 0.f;
  p.qn = screen ? (const float*)(sc + off_qn) : nullptr;

  // ONE preparation launch: query planes, bounds, cleared counters (pairs without tiles never write theirs)
  mma_query_prep_kernel<<<(p.q_pad + 7) / 8, 256, 0, st>>>(
      queries, p.n_queries, p.q_pad, split ? 1 : 0, planes, p.g_tau, p.pool, top_k, hist ? 1 : 0,
      screen ? (float*)(sc + off_qn) : nullptr, p.cand_cnt, (long long)((screen ? off_tau2 : off_flag) - off_cnt) / 4,
      screen ? (u32*)(sc + off_tau2) : nullptr, screen ? (u32*)(sc + off_pool2) : nullptr);
  MMF_LAUNCH_OK(h);

  CUtensorMap tm_q;
  const cuuint64_t dims[2] = {MMF_DIM, (cuuint64_t)npl * p.q_pad};
  const cuuint64_t strides[1] = {MMF_DIM * 2};
  const cuuint32_t box[2] = {KBLK, TILE_M};
  if (!encode_map(s, &tm_q, split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, planes, dims,
                  strides, box))
    return mmf_set_error(h, MMF_ERR_CUDA, "cuTensorMapEncodeTiled failed for the query operand");

#define MMF_MMA_CASE(SPLIT_, KPL_, KR_, VAR_)                                                                       \
  return cg == 2 ? launch_mma<SPLIT_, KPL_, 2, KR_, VAR_>(h, s, tm_q, p, n_pairs, threshold, out_scores, out_rows,       \
                                                          out_packed, out_disc, st)                                    \
                 : launch_mma<SPLIT_, KPL_, 1, KR_, VAR_>(h, s, tm_q, p, n_pairs, threshold, out_scores, out_rows,       \
                                                          out_packed, out_disc, st)
  // 1-plane kernels with the bucket pool: PARITY (warp sets alternate tiles) wins on short strips, where the warm-up
  // of a strip weighs (C2: 105 tiles per strip, 0.293 vs 0.344 ms), all warps on every tile on long ones (4096
  // queries x 1.25 M rows, top-10: 3.47 vs 4.15 ms); option "epi_parity": -1 = by strip length, 0 / 1 = forced
  const long long tiles_per_strip = p.n_aligned ? p.v_aligned / p.seg : p.v_tiles;
  const bool parity = h->opt.epi_parity < 0 ? tiles_per_strip < 1024 : h->opt.epi_parity != 0;
  if (screen) {
    // 1 pass over the hi planes + exact re-scoring of the survivors; then the guarded 3-pass search, which
    // returns at once unless a candidate band overflowed (its bounds restart from scratch -- the screening
    // pass published bounds on APPROXIMATE scores -- and were prepared by the same launch as the first set)
#define MMF_SCREEN_LAUNCH(CG_, VAR_) \
  launch_mma<false, 8, CG_, 16, VAR_>(h, s, tm_q, p, n_pairs, threshold, out_scores, out_rows, out_packed, out_disc, st)
    rc = cg == 2 ? (parity ? MMF_SCREEN_LAUNCH(2, VAR_SCREEN | VAR_PARITY) : MMF_SCREEN_LAUNCH(2, VAR_SCREEN))
                 : (parity ? MMF_SCREEN_LAUNCH(1, VAR_SCREEN | VAR_PARITY) : MMF_SCREEN_LAUNCH(1, VAR_SCREEN));
#undef MMF_SCREEN_LAUNCH
    if (rc != MMF_OK) return rc;
    p.cand_cnt = (int*)(sc + off_cnt2);
    p.g_tau = (u32*)(sc + off_tau2);
    p.pool = (u32*)(sc + off_pool2);
    MMF_MMA_CASE(true, 8, 16, VAR_GUARD | VAR_PARITY);
  }
  if (split) {
    if (top_k <= 16) MMF_MMA_CASE(true, 8, 16, VAR_PARITY);       // 3-pass: 6144 clk of MMA per tile, PARITY hides the hand-offs
    if (kpl == 8) MMF_MMA_CASE(true, 8, 0, 0);
    MMF_MMA_CASE(true, 16, 0, 0);
  } else {
    if (top_k <= 16) { if (parity) MMF_MMA_CASE(false, 8, 16, VAR_PARITY); MMF_MMA_CASE(false, 8, 16, 0); }
    if (kpl == 8) MMF_MMA_CASE(false, 8, 0, 0);
    MMF_MMA_CASE(false, 16, 0, 0);
  }
#undef MMF_MMA_CASE
}
