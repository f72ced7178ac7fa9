// Shared device helpers: order-preserving score keys, warp reductions, handle layout.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

#include "../../include/mmf_b200.h"

#define MMF_DIM 512                 // CLIP ViT-B/32 projection dim (misinfo_forensics.py:78-79)
#define MMF_SPLIT_SCALE 256.0f      // fp32-exact mode stores v*2^8 as fp16 hi + fp16 lo
#define MMF_SPLIT_INV_SCALE (1.0f / 256.0f)

namespace mmf {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u32 FULL = 0xffffffffu;

// ---- total order on scores -------------------------------------------------------------
// np.argsort(similarities)[-k:][::-1] (misinfo_forensics.py:449): descending score, NaN
// ranks above everything (argsort puts NaN last), equal scores -> higher row id first
// (stable ascending sort, reversed).  key64 = okey(score) << 32 | row; larger == better;
// 0 is reserved for "empty" (okey 0 would be a negative NaN, which is canonicalised away).
__host__ __device__ __forceinline__ u32 okey(float s) {
  if (s != s) return 0xFFFFFFFFu;
  s = s + 0.0f;                                  // -0.0 -> +0.0 (numpy compares them equal)
#ifdef __CUDA_ARCH__
  u32 u = __float_as_uint(s);
#else
  union { float f; u32 u; } c; c.f = s; u32 u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float okey_inv(u32 k) {
  u32 u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; u32 u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ u64 pack_key(float s, u32 row) { return ((u64)okey(s) << 32) | row; }

// the reference evaluates `float(s) > 0.85` in Python double on an fp32 value
// (misinfo_forensics.py:463-464); NaN fails.
__device__ __forceinline__ float discrepancy_rule(float top, double threshold) {
  return ((double)top > threshold) ? top : 0.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {   // read-once data: keep it out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

}  // namespace mmf

// ---- fused candidate push (exchange.cu <-> vault_mma.cu) ----------------------------------
// When set on the handle during a search, the merge tail of the tcgen05 search writes its per-query winners
// straight into slot [rank] of every rank's gather buffer and the last block publishes the epoch flags -- the
// separate push kernel of exchange.cu is then not launched (push_fused reports that it happened).
#define MMF_XCHG_MAX_WORLD 16
struct mmf_push_ctx {
  unsigned char* base[MMF_XCHG_MAX_WORLD];   // rank r's exchange buffer as mapped in this process
  int rank, world, parity;
  unsigned epoch;
  size_t slot_off;                           // byte offset of slot [rank] of gather[parity] in every buffer
  unsigned* done;                            // ticket counter (device, self-resetting)
};

// ---- handle ----------------------------------------------------------------------------
// Switches of a handle.  Read ONCE from the environment at mmf_create (MMF_OPT_<NAME>, upper case) and settable
// with mmf_set_option; never consulted through getenv on the search path.
struct mmf_options {
  int screen = 1;          // fp32-exact vaults, top_k <= 16: screened search (1) or the 3-pass kernel (0, A/B + triage)
  int fused_push = 1;      // peer-memory exchange: the search's merge tail pushes the winners itself where it can
  int debug = 0;           // tcgen05 search triage bits: 1 skip the filter, 2 skip the vault TMA, 4 skip the MMA waits, 8 print clocks
  int force_cg = 0;        // tcgen05 search: force 1 or 2 CTAs per MMA (0 = automatic)
  int flat_schedule = 0;   // tcgen05 search: plain flattened schedule instead of the L2-aware one
  int stream_tma = 1;      // streaming (batch-1) search: TMA-staged kernel (1) or the register-staged one (0)
  int lockstep = 1;        // tcgen05 search with several query-tile groups: producers of a segment stay within 64 tiles of each other
  int epi_parity = -1;     // tcgen05 search, 1-plane kernels with top_k <= 16: epilogue warp sets alternate tiles (1), all warps
                           // work on every tile (0), or chosen by strip length (-1, default)
};

// One in-flight batch of the host-buffer entry points (api.cu): device I/O buffers, pinned staging for the results
struct mmf_host_slot {
  void* io = nullptr;      size_t io_bytes = 0;        // device: inputs, then one contiguous block of outputs
  void* pinned = nullptr;  size_t pinned_bytes = 0;    // host mirror of the output block
  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
  bool busy = false;
  int64_t n = 0;
  int top_k = 0;
};

struct mmf_handle {
  int device = -1;
  int sm_count = 148;
  std::string last_error;
  int64_t launches = 0;
  int64_t collectives = 0;
  mmf_options opt;
  // vault shard
  void* vault = nullptr;            // FP32 mode: [n][2][512] fp16 (hi row, lo row); BF16: [n][512] bf16
  bool vault_loaded = false;
  int64_t vault_rows = 0;
  int64_t vault_nan_rows = 0;      // zero-norm rows (NaN after normalisation): handled by the streaming kernel only
  unsigned long long* vault_nan_rows_dev = nullptr;
  int64_t vault_row_offset = 0;
  int vault_mode = 0;
  size_t vault_bytes = 0;
  // fusion judge
  float* fusion_params = nullptr;   // device, MMF_FUSION_PARAMS floats, re-laid out (see fusion.cu)
  bool fusion_loaded = false;
  // scratch: arena 0 serves the asynchronous entry points (caller's stream), arena 1 the *_host / submit entry
  // points (the handle's own stream), so that the two families never share counters or candidate lists
  void* scratch_arena[2] = {nullptr, nullptr};
  size_t scratch_arena_bytes[2] = {0, 0};
  int scratch_sel = 0;
  void* scratch() const { return scratch_arena[scratch_sel]; }
  mmf_host_slot slot[2];
  cudaStream_t own_stream = nullptr;       // compute of the host-buffer entry points
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  // TMA descriptors for the tcgen05 path live in vault_mma.cu's state
  void* mma_state = nullptr;
  // row-sharded search with the library's own NCCL communicator (shard.cu), null until mmf_shard_init
  void* shard_state = nullptr;
  // peer-memory candidate exchange (exchange.cu), null until mmf_exchange_attach
  void* xchg_state = nullptr;
  const mmf_push_ctx* push_ctx = nullptr;   // non-null only inside mmf_vault_search_exchange (fused push requested)
  bool push_fused = false;
};

int mmf_set_error(mmf_handle* h, int status, const char* fmt, ...);
int mmf_ensure_scratch(mmf_handle* h, size_t bytes, cudaStream_t stream);

#define MMF_CUDA_OK(h, expr)                                                                     \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return mmf_set_error((h), MMF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                                  \
  } while (0)

#define MMF_LAUNCH_OK(h)                                                                         \
  do {                                                                                           \
    (h)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess)                                                                      \
      return mmf_set_error((h), MMF_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                                  \
  } while (0)
