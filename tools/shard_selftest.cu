// Torch-free multi-GPU self test of the ROW-SHARDED Truth-Vault search through the C ABI (include/mmf_b200.h):
// one host thread per GPU, each with its own handle and its own row shard (mmf_vault_load with row_offset).
//   1. mmf_vault_search_sharded: local search + ncclAllGather (the library's own communicator, NCCL bound at run time)
//      + merge, one call per rank;
//   2. mmf_vault_search_exchange: the same exchange over NVLink peer memory (csrc/exchange.cu) -- every rank stores
//      its candidates into the peers' buffers (cudaDeviceEnablePeerAccess: plain device pointers in one process), the
//      merge kernel waits on per-rank flags; with and without the push fused into the search's merge tail.
// Every rank's result must equal, bit for bit, an UNSHARDED search of the whole vault done by rank 0.
// Default shape = BASELINE.json configs[3]: 4096 queries x 10 M bf16 rows, top-100.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/shard_selftest tools/shard_selftest.cu \
//        -Lmulti-modal-misinformation-detection-with-explanation-generation_b200 -lmmf_b200 -lpthread ...
//   tools/shard_selftest [n_gpus (default: all)] [rows_total (10000000)] [n_queries (4096)] [top_k (100)] [fp32|bf16]
#include <cuda_runtime.h>
#include <pthread.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/mmf_b200.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)
#define MM(h, x)                                                                                \
  do {                                                                                          \
    int rc_ = (x);                                                                              \
    if (rc_ != MMF_OK) { printf("mmf error %d (%s) at %s:%d: %s\n", rc_, mmf_status_string(rc_), __FILE__, __LINE__, mmf_last_error(h)); exit(3); } \
  } while (0)

__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
// element i of GLOBAL row (row0 + i / 512): the same value whichever rank generates it
__global__ void fill_rows(float* out, long long n_rows, long long row0, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * 512) return;
  const uint64_t gi = (uint64_t)(row0 * 512 + i);
  const uint32_t a = hash32(seed * 0x9E3779B97F4A7C15ull + gi * 2), b = hash32(seed * 0x9E3779B97F4A7C15ull + gi * 2 + 1);
  const float u = ((a & 0xFFFF) + (a >> 16) + (b & 0xFFFF) + (b >> 16)) * (1.0f / 65536.0f) - 2.0f;
  const float scale = 0.25f + (hash32(seed + (uint64_t)(gi / 512) * 7919) & 1023) * (1.0f / 256.0f);
  out[i] = u * scale;
}

struct Shared {
  int world, nq, k, mode;
  long long rows_total;
  unsigned char id[MMF_SHARD_ID_BYTES];
  uint64_t peer_ptr[16];
  int64_t xchg_bytes;
  std::vector<float> ref_scores, ref_disc;
  std::vector<int64_t> ref_rows;
  float ms[4][16];
  int fails[16];
  int can[16];
  pthread_barrier_t bar;
};
static Shared S;

struct Buffers { float* scores; int64_t* rows; float* disc; };

static int compare(int rank, const char* what, const Buffers& b) {
  std::vector<float> sc((size_t)S.nq * S.k), di(S.nq);
  std::vector<int64_t> ro((size_t)S.nq * S.k);
  CK(cudaMemcpy(sc.data(), b.scores, sc.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ro.data(), b.rows, ro.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(di.data(), b.disc, di.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (size_t i = 0; i < sc.size(); ++i) bad += memcmp(&sc[i], &S.ref_scores[i], 4) != 0 || ro[i] != S.ref_rows[i];
  for (int i = 0; i < S.nq; ++i) bad += memcmp(&di[i], &S.ref_disc[i], 4) != 0;
  if (bad) printf("  rank %d %-34s %lld differences vs the unsharded search -> MISMATCH\n", rank, what, bad);
  return bad != 0;
}

static void* rank_main(void* arg) {
  const int rank = (int)(intptr_t)arg, world = S.world;
  CK(cudaSetDevice(rank));
  mmf_handle* h = nullptr;
  { int rc = mmf_create(rank, &h); if (rc != MMF_OK) { printf("mmf_create(%d) failed: %s\n", rank, mmf_status_string(rc)); exit(1); } }
  const long long per = (S.rows_total + world - 1) / world, lo = rank * per < S.rows_total ? rank * per : S.rows_total;
  const long long hi = lo + per < S.rows_total ? lo + per : S.rows_total, n_local = hi - lo;
  float *d_rows = nullptr, *d_q = nullptr;
  CK(cudaMalloc(&d_q, (size_t)S.nq * 512 * 4));
  fill_rows<<<(unsigned)(((long long)S.nq * 512 + 255) / 256), 256>>>(d_q, S.nq, 0, 22);
  Buffers out;
  CK(cudaMalloc(&out.scores, (size_t)S.nq * S.k * 4)); CK(cudaMalloc(&out.rows, (size_t)S.nq * S.k * 8)); CK(cudaMalloc(&out.disc, S.nq * 4));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));

  if (rank == 0) {          // the reference: an unsharded search of the whole vault (slab by slab into one handle is not
                            // possible -- a handle holds one shard -- so rank 0 generates all rows once)
    mmf_handle* full = nullptr;
    MM(full, mmf_create(0, &full));
    float* all = nullptr;
    CK(cudaMalloc(&all, (size_t)S.rows_total * 512 * 4));
    fill_rows<<<(unsigned)((S.rows_total * 512 + 255) / 256), 256>>>(all, S.rows_total, 0, 21);
    CK(cudaDeviceSynchronize());
    MM(full, mmf_vault_load(full, all, 1, S.rows_total, 512, MMF_F32, S.mode, 0));
    CK(cudaFree(all));
    MM(full, mmf_vault_search(full, d_q, S.nq, S.k, 0.85, MMF_ALGO_AUTO, out.scores, out.rows, out.disc, st));
    CK(cudaStreamSynchronize(st));
    S.ref_scores.resize((size_t)S.nq * S.k); S.ref_rows.resize((size_t)S.nq * S.k); S.ref_disc.resize(S.nq);
    CK(cudaMemcpy(S.ref_scores.data(), out.scores, S.ref_scores.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(S.ref_rows.data(), out.rows, S.ref_rows.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(S.ref_disc.data(), out.disc, S.ref_disc.size() * 4, cudaMemcpyDeviceToHost));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < 3; ++i) MM(full, mmf_vault_search(full, d_q, S.nq, S.k, 0.85, MMF_ALGO_AUTO, out.scores, out.rows, out.disc, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&S.ms[0][0], e0, e1));
    S.ms[0][0] /= 3;
    mmf_destroy(full);
    MM(h, mmf_shard_unique_id(S.id));
  }
  CK(cudaMalloc(&d_rows, (size_t)(n_local > 0 ? n_local : 1) * 512 * 4));
  if (n_local > 0) fill_rows<<<(unsigned)((n_local * 512 + 255) / 256), 256>>>(d_rows, n_local, lo, 21);
  CK(cudaDeviceSynchronize());
  MM(h, mmf_vault_load(h, d_rows, 1, n_local, 512, MMF_F32, S.mode, lo));
  CK(cudaFree(d_rows));
  pthread_barrier_wait(&S.bar);                       // reference + unique id are there

  auto timed = [&](int slot, auto&& call) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) call();
    CK(cudaStreamSynchronize(st));
    pthread_barrier_wait(&S.bar);
    CK(cudaEventRecord(e0, st));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) call();
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    S.ms[slot][rank] = ms / reps;
    pthread_barrier_wait(&S.bar);
  };

  // ---- 1. library-owned NCCL
  MM(h, mmf_shard_init(h, rank, world, S.id));
  timed(1, [&] { MM(h, mmf_vault_search_sharded(h, d_q, S.nq, S.k, 0.85, MMF_ALGO_AUTO, out.scores, out.rows, out.disc, st)); });
  S.fails[rank] += compare(rank, "mmf_vault_search_sharded (NCCL)", out);
  MM(h, mmf_shard_finalize(h));

  // ---- 2. peer-memory exchange
  int can = 1;
  for (int r = 0; r < world && can; ++r)
    if (r != rank) { int ok = 0; CK(cudaDeviceCanAccessPeer(&ok, rank, r)); can = ok; }
  for (int r = 0; r < world && can; ++r)
    if (r != rank) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(r, 0);
      cudaGetLastError();
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        printf("  rank %d: cudaDeviceEnablePeerAccess(%d) failed: %s\n", rank, r, cudaGetErrorString(e));
        can = 0;
      }
    }
  S.can[rank] = can;
  pthread_barrier_wait(&S.bar);
  for (int r = 0; r < world; ++r) can = can && S.can[r];      // all ranks or none
  if (can) {
    void* buf = nullptr;
    CK(cudaMalloc(&buf, (size_t)S.xchg_bytes));
    CK(cudaMemset(buf, 0, (size_t)S.xchg_bytes));
    S.peer_ptr[rank] = (uint64_t)(uintptr_t)buf;
    pthread_barrier_wait(&S.bar);
    MM(h, mmf_exchange_attach(h, rank, world, S.peer_ptr, S.xchg_bytes));
    pthread_barrier_wait(&S.bar);                     // every rank has cleared its flags before anyone pushes
    const int k_local = (long long)S.k < per ? S.k : (int)per;
    for (int fused = 1; fused >= 0; --fused) {
      MM(h, mmf_set_option(h, "fused_push", fused));
      timed(2 + (1 - fused), [&] { MM(h, mmf_vault_search_exchange(h, d_q, S.nq, S.k, k_local, 0.85, MMF_ALGO_AUTO, out.scores, out.rows, out.disc, st)); });
      S.fails[rank] += compare(rank, fused ? "peer-memory exchange (fused push)" : "peer-memory exchange (push kernel)", out);
    }
    pthread_barrier_wait(&S.bar);
    MM(h, mmf_exchange_detach(h));
    CK(cudaFree(buf));
  } else if (rank == 0) {
    printf("  (no peer access between the GPUs: peer-memory exchange skipped)\n");
  }
  mmf_destroy(h);
  return nullptr;
}

int main(int argc, char** argv) {
  int n_dev = 0;
  CK(cudaGetDeviceCount(&n_dev));
  S.world = argc > 1 ? atoi(argv[1]) : n_dev;
  if (S.world < 1 || S.world > n_dev || S.world > 16) { printf("need 1..%d GPUs\n", n_dev < 16 ? n_dev : 16); return 1; }
  S.rows_total = argc > 2 ? atoll(argv[2]) : 10000000;
  S.nq = argc > 3 ? atoi(argv[3]) : 4096;
  S.k = argc > 4 ? atoi(argv[4]) : 100;
  S.mode = (argc > 5 && !strcmp(argv[5], "fp32")) ? MMF_VAULT_FP32 : MMF_VAULT_BF16;
  setvbuf(stdout, nullptr, _IOLBF, 0);
  const long long per = (S.rows_total + S.world - 1) / S.world;
  mmf_exchange_layout(S.world, S.nq, (long long)S.k < per ? S.k : (int)per, nullptr, &S.xchg_bytes);
  printf("%s: %d GPUs, %lld rows (%s), %d queries, top-%d\n", mmf_version(), S.world, S.rows_total,
         S.mode == MMF_VAULT_BF16 ? "bf16" : "fp32-exact", S.nq, S.k);
  pthread_barrier_init(&S.bar, nullptr, S.world);
  std::vector<pthread_t> th(S.world);
  for (int r = 0; r < S.world; ++r) pthread_create(&th[r], nullptr, rank_main, (void*)(intptr_t)r);
  for (int r = 0; r < S.world; ++r) pthread_join(th[r], nullptr);
  int fails = 0;
  const char* names[4] = {"unsharded on one GPU", "sharded, NCCL all-gather (library)", "sharded, peer memory, fused push", "sharded, peer memory, push kernel"};
  for (int s = 0; s < 4; ++s) {
    float mx = 0;
    for (int r = 0; r < (s == 0 ? 1 : S.world); ++r) mx = S.ms[s][r] > mx ? S.ms[s][r] : mx;
    if (mx > 0) printf("  %-38s %8.3f ms per batch (max over ranks) = %9.0f queries/s, %7.1f TFLOP/s per rank\n", names[s], mx, S.nq / mx * 1e3,
                       2.0 * S.nq * (s == 0 ? S.rows_total : per) * 512 / mx * 1e-9);
  }
  for (int r = 0; r < S.world; ++r) fails += S.fails[r];
  printf("%s\n", fails ? "SHARD SELFTEST FAILED" : "shard selftest ok: every rank's result == the unsharded search, bit for bit");
  return fails ? 1 : 0;
}
