// Torch-free self test of libmmf_b200.so through its C ABI (include/mmf_b200.h): the search kernels against each
// other (tcgen05 screened / 3-pass / histogram-bound vs the HBM-streaming kernel), the host-buffer entry points
// against the device ones, the peer-memory exchange protocol on one device, with timings.
// Starts in a second (no Python), so it fits the shortest GPU slot:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/cabi_selftest tools/cabi_selftest.cu \
//        -Lmulti-modal-misinformation-detection-with-explanation-generation_b200 -lmmf_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../multi-modal-misinformation-detection-with-explanation-generation_b200'
//   timeout 120 tools/cabi_selftest [rows_fp32 [rows_bf16 [rows_bf16_b]]]
//   ncu --set full --clock-control none --import-source on -k regex:vault_mma_topk -c 3 -o x tools/cabi_selftest profile
//        (one C2-shaped screened search + its guarded no-op + one C4-shard-shaped histogram search)
//
// Everything it checks is also covered by tests/ (pytest -m gpu); this is the quick look.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "../include/mmf_b200.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)
#define MM(x)                                                                                   \
  do {                                                                                          \
    int rc_ = (x);                                                                              \
    if (rc_ != MMF_OK) { printf("mmf error %d (%s) at %s:%d: %s\n", rc_, mmf_status_string(rc_), __FILE__, __LINE__, mmf_last_error(H)); exit(3); } \
  } while (0)

static mmf_handle* H = nullptr;

// Section / arm selection, so that a crash in an arm that has never run on a GPU cannot hide the others:
//   SELFTEST_ONLY=a,b   run only these
//   SELFTEST_SKIP=a,b   run everything but these
// tokens: core bf16 sweep host exchange
static bool in_list(const char* list, const char* tok) {
  if (!list) return false;
  const size_t n = strlen(tok);
  for (const char* p = list; (p = strstr(p, tok)) != nullptr; p += n)
    if ((p == list || p[-1] == ',') && (p[n] == 0 || p[n] == ',')) return true;
  return false;
}
static bool want(const char* tok) {
  const char* only = getenv("SELFTEST_ONLY");
  if (only && *only) return in_list(only, tok);
  return !in_list(getenv("SELFTEST_SKIP"), tok);
}

__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
// approximately normal: sum of 4 uniforms, centred (variance 1/3), times a per-row scale
__global__ void fill_rows(float* out, long long n_rows, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * 512) return;
  const uint32_t a = hash32(seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 2), b = hash32(seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 2 + 1);
  const float u = ((a & 0xFFFF) + (a >> 16) + (b & 0xFFFF) + (b >> 16)) * (1.0f / 65536.0f) - 2.0f;
  const float scale = 0.25f + (hash32(seed + (uint64_t)(i / 512) * 7919) & 1023) * (1.0f / 256.0f);
  out[i] = u * scale;
}
// query j (j < n_plant) = vault row (j * stride) + w * noise; identical copies of row `dup_src` over [dup_lo, dup_hi)
__global__ void plant_queries(float* q, const float* vault, int n_plant, long long stride, float w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_plant * 512) return;
  const int j = i / 512, d = i % 512;
  q[i] = vault[(long long)j * stride * 512 + d] * 3.0f + w * q[i];
}
__global__ void duplicate_rows(float* vault, long long src, long long lo, long long hi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (hi - lo) * 512) return;
  vault[lo * 512 + i] = vault[src * 512 + i % 512];
}

struct Result {
  std::vector<float> scores, disc;
  std::vector<int64_t> rows;
};

static float* d_q = nullptr;
static float* d_scores = nullptr;
static int64_t* d_rows = nullptr;
static float* d_disc = nullptr;

static Result search(int nq, int k, int algo, float* ms_out = nullptr, int reps = 0) {
  MM(mmf_vault_search(H, d_q, nq, k, 0.85, algo, d_scores, d_rows, d_disc, nullptr));
  CK(cudaDeviceSynchronize());
  Result r;
  r.scores.resize((size_t)nq * k); r.rows.resize((size_t)nq * k); r.disc.resize(nq);
  CK(cudaMemcpy(r.scores.data(), d_scores, r.scores.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r.rows.data(), d_rows, r.rows.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r.disc.data(), d_disc, r.disc.size() * 4, cudaMemcpyDeviceToHost));
  if (ms_out && reps > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) MM(mmf_vault_search(H, d_q, nq, k, 0.85, algo, d_scores, d_rows, d_disc, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) MM(mmf_vault_search(H, d_q, nq, k, 0.85, algo, d_scores, d_rows, d_disc, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(ms_out, e0, e1));
    *ms_out /= reps;
  }
  return r;
}

// bit-exact comparison; prints the first difference
static bool same(const Result& a, const Result& b, int nq, int k, const char* what) {
  long long bad_rows = 0, bad_scores = 0, bad_disc = 0, first = -1;
  for (size_t i = 0; i < a.rows.size(); ++i) {
    if (a.rows[i] != b.rows[i]) { ++bad_rows; if (first < 0) first = (long long)i; }
    if (memcmp(&a.scores[i], &b.scores[i], 4) != 0) ++bad_scores;
  }
  for (int i = 0; i < nq; ++i) bad_disc += memcmp(&a.disc[i], &b.disc[i], 4) != 0;
  printf("  %-46s rows differ %lld, scores differ %lld, discrepancy differ %lld of %d x %d -> %s\n", what, bad_rows,
         bad_scores, bad_disc, nq, k, (bad_rows | bad_scores | bad_disc) ? "MISMATCH" : "identical");
  if (first >= 0)
    printf("    first: query %lld rank %lld: row %lld (%.9g) vs %lld (%.9g)\n", first / k, first % k,
           (long long)a.rows[first], a.scores[first], (long long)b.rows[first], b.scores[first]);
  fflush(stdout);
  return !(bad_rows | bad_scores | bad_disc);
}

// tolerance comparison (two different kernels): scores within tol rank by rank, rows equal outside near-ties
static bool near_equal(const Result& a, const Result& b, int nq, int k, float tol, const char* what) {
  long long bad = 0, row_diff = 0;
  float worst = 0.f;
  for (int q = 0; q < nq; ++q)
    for (int j = 0; j < k; ++j) {
      const size_t i = (size_t)q * k + j;
      const float d = fabsf(a.scores[i] - b.scores[i]);
      if (!(d <= tol) && !(std::isnan(a.scores[i]) && std::isnan(b.scores[i]))) ++bad;
      if (d > worst) worst = d;
      if (a.rows[i] != b.rows[i]) {
        const float g1 = j > 0 ? fabsf(b.scores[i] - b.scores[i - 1]) : 1.f, g2 = j + 1 < k ? fabsf(b.scores[i] - b.scores[i + 1]) : 0.f;
        if (fminf(g1, g2) > 2 * tol) ++row_diff;
      }
    }
  printf("  %-46s max |score diff| %.3g (tol %.0e), %lld scores off, %lld row mismatches outside near-ties -> %s\n", what,
         worst, tol, bad, row_diff, (bad | row_diff) ? "MISMATCH" : "ok");
  fflush(stdout);
  return !(bad | row_diff);
}

static void opt(const char* name, int v) {
  if (mmf_set_option(H, name, v) != MMF_OK) { printf("mmf_set_option(%s) failed: %s\n", name, mmf_last_error(H)); exit(3); }
}

int main(int argc, char** argv) {
  const long long rows_fp32 = (argc > 1 && strcmp(argv[1], "profile") && strcmp(argv[1], "tune")) ? atoll(argv[1]) : 1000000;
  const long long rows_bf16_a = argc > 2 ? atoll(argv[2]) : 1250000;
  const long long rows_bf16_b = argc > 3 ? atoll(argv[3]) : 0;      // optional second bf16 size (e.g. 10000000)
  int fails = 0;
  setvbuf(stdout, nullptr, _IOLBF, 0);     // a timeout must not eat the lines already printed
  printf("%s\n", mmf_version());
  { int rc = mmf_create(0, &H); if (rc != MMF_OK) { printf("mmf_create failed: %s\n", mmf_status_string(rc)); return 1; } }
  const int NQ = 4096;
  CK(cudaMalloc(&d_q, (size_t)NQ * 512 * 4));
  CK(cudaMalloc(&d_scores, (size_t)NQ * 256 * 4));
  CK(cudaMalloc(&d_rows, (size_t)NQ * 256 * 8));
  CK(cudaMalloc(&d_disc, (size_t)NQ * 4));
  float* d_vault = nullptr;
  long long rows_max = rows_fp32 > rows_bf16_a ? rows_fp32 : rows_bf16_a;
  if (rows_bf16_b > rows_max) rows_max = rows_bf16_b;
  if (rows_max < 1300000) rows_max = 1300000;      // the fixed-size sections (sweep, host entry, exchange, profile) need this much
  CK(cudaMalloc(&d_vault, (size_t)rows_max * 512 * 4));

  if (argc > 1 && !strcmp(argv[1], "profile")) {     // one launch of each flagship kernel (library defaults), for ncu
    fill_rows<<<(unsigned)((1000000ll * 512 + 255) / 256), 256>>>(d_vault, 1000000, 11);
    fill_rows<<<(4096 * 512 + 255) / 256, 256>>>(d_q, 4096, 12);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, 1000000, 512, MMF_F32, MMF_VAULT_FP32, 0));
    if (rows_bf16_a > 1250000) {     // `profile <rows>`: the one-GPU C4 shape, lock-step producers on, then off (for ncu's DRAM counters)
      fill_rows<<<(unsigned)((rows_bf16_a * 512 + 255) / 256), 256>>>(d_vault, rows_bf16_a, 21);
      CK(cudaDeviceSynchronize());
      MM(mmf_vault_load(H, d_vault, 1, rows_bf16_a, 512, MMF_F32, MMF_VAULT_BF16, 0));
      opt("lockstep", 1); search(4096, 100, MMF_ALGO_MMA);
      opt("lockstep", 0); search(4096, 100, MMF_ALGO_MMA);
      printf("profile mode: 2 searches of 4096 x %lld bf16 rows done\n", rows_bf16_a);
      mmf_destroy(H);
      return 0;
    }
    if (getenv("PROFILE_C1")) {      // the C1 shape instead: 1 000 queries x 100 000 rows
      MM(mmf_vault_load(H, d_vault, 1, 100000, 512, MMF_F32, MMF_VAULT_FP32, 0));
      search(1000, 10, MMF_ALGO_MMA);
      printf("profile mode: C1 search done\n");
      mmf_destroy(H);
      return 0;
    }
    search(256, 10, MMF_ALGO_MMA);
    search(1, 10, MMF_ALGO_STREAM);
    fill_rows<<<(unsigned)((1250000ll * 512 + 255) / 256), 256>>>(d_vault, 1250000, 21);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, 1250000, 512, MMF_F32, MMF_VAULT_BF16, 0));
    search(4096, 100, MMF_ALGO_MMA);
    printf("profile mode: 3 searches done\n");
    mmf_destroy(H);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "tune")) {        // interleaved A/B of the epilogue options on the C2 shape
    fill_rows<<<(unsigned)((1000000ll * 512 + 255) / 256), 256>>>(d_vault, 1000000, 11);
    fill_rows<<<(256 * 512 + 255) / 256, 256>>>(d_q, 256, 12);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, 1000000, 512, MMF_F32, MMF_VAULT_FP32, 0));
    struct Arm { const char* name; int parity, debug; };
    const Arm arms[] = {{"all warps on every tile", 0, 0}, {"parity (warp sets alternate tiles)", 1, 0},
                        {"library default (by strip length)", -1, 0}};
    const int n_arms = sizeof arms / sizeof arms[0], rounds = 7;
    std::vector<std::vector<float>> t(n_arms);
    Result first;
    for (int r = 0; r < rounds; ++r)
      for (int a = 0; a < n_arms; ++a) {
        opt("epi_parity", arms[a].parity); opt("debug", arms[a].debug);
        float ms = 0;
        Result res = search(256, 10, MMF_ALGO_MMA, &ms, 40);
        if (r == 0 && a == 0) first = res;
        else if (r == 0) fails += !same(res, first, 256, 10, arms[a].name);
        t[a].push_back(ms);
      }
    for (int a = 0; a < n_arms; ++a) {
      std::vector<float> v = t[a];
      for (size_t i = 0; i < v.size(); ++i) for (size_t j = i + 1; j < v.size(); ++j) if (v[j] < v[i]) { float x = v[i]; v[i] = v[j]; v[j] = x; }
      printf("  %-36s min %.4f  median %.4f  max %.4f ms\n", arms[a].name, v[0], v[v.size() / 2], v.back());
    }
    // the long-strip shape (bf16 vault, 4096 queries, top-10)
    fill_rows<<<(unsigned)((1250000ll * 512 + 255) / 256), 256>>>(d_vault, 1250000, 21);
    fill_rows<<<(4096 * 512 + 255) / 256, 256>>>(d_q, 4096, 12);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, 1250000, 512, MMF_F32, MMF_VAULT_BF16, 0));
    for (int a = 0; a < n_arms; ++a) t[a].clear();
    for (int r = 0; r < 3; ++r)
      for (int a = 0; a < n_arms; ++a) {
        opt("epi_parity", arms[a].parity);
        float ms = 0;
        Result res = search(4096, 10, MMF_ALGO_MMA, &ms, 5);
        if (r == 0 && a == 0) first = res;
        else if (r == 0) fails += !same(res, first, 4096, 10, arms[a].name);
        t[a].push_back(ms);
      }
    for (int a = 0; a < n_arms; ++a) {
      std::vector<float> v = t[a];
      for (size_t i = 0; i < v.size(); ++i) for (size_t j = i + 1; j < v.size(); ++j) if (v[j] < v[i]) { float x = v[i]; v[i] = v[j]; v[j] = x; }
      printf("  bf16 4096 x 1.25M top-10: %-36s min %.4f  median %.4f ms\n", arms[a].name, v[0], v[v.size() / 2]);
    }
    if (rows_bf16_a > 1250000) {      // `tune <rows>`: the one-GPU C4 shape (4 096 queries, top-100), lock-step producers on / off
      opt("epi_parity", -1);
      fill_rows<<<(unsigned)((rows_bf16_a * 512 + 255) / 256), 256>>>(d_vault, rows_bf16_a, 21);
      CK(cudaDeviceSynchronize());
      MM(mmf_vault_load(H, d_vault, 1, rows_bf16_a, 512, MMF_F32, MMF_VAULT_BF16, 0));
      std::vector<float> t_on, t_off;
      Result r_on, r_off;
      for (int r = 0; r < 3; ++r) {
        float ms = 0;
        opt("lockstep", 1); r_on = search(4096, 100, MMF_ALGO_MMA, &ms, 2); t_on.push_back(ms);
        opt("lockstep", 0); r_off = search(4096, 100, MMF_ALGO_MMA, &ms, 2); t_off.push_back(ms);
      }
      opt("lockstep", 1);
      fails += !same(r_on, r_off, 4096, 100, "lock-step producers on vs off");
      printf("  bf16 4096 x %lld top-100: lock-step producers on  %.3f %.3f %.3f ms\n", rows_bf16_a, t_on[0], t_on[1], t_on[2]);
      printf("  bf16 4096 x %lld top-100: lock-step producers off %.3f %.3f %.3f ms\n", rows_bf16_a, t_off[0], t_off[1], t_off[2]);
    }
    printf("tune: %d mismatches\n", fails);
    mmf_destroy(H);
    return 0;
  }
  // ---------------- fp32-exact vault: stream vs 3-pass tcgen05 vs screened search (top-10, 256 queries)
  if (want("core")) {
    const int nq = 256, k = 10;
    fill_rows<<<(unsigned)((rows_fp32 * 512 + 255) / 256), 256>>>(d_vault, rows_fp32, 11);
    fill_rows<<<(nq * 512 + 255) / 256, 256>>>(d_q, nq, 12);
    plant_queries<<<(32 * 512 + 255) / 256, 256>>>(d_q, d_vault, 32, rows_fp32 / 32, 1.5f);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, rows_fp32, 512, MMF_F32, MMF_VAULT_FP32, 0));
    printf("[fp32-exact] %lld rows, %d queries, top-%d\n", rows_fp32, nq, k);
    float ms_mma = 0, ms_screen = 0, ms_parity = 0, ms_stream = 0, ms_one = 0;
    Result stream = search(nq, k, MMF_ALGO_STREAM, &ms_stream, 2);
    opt("screen", 0);
    Result mma = search(nq, k, MMF_ALGO_MMA, &ms_mma, 20);
    opt("screen", 1);
    fails += !near_equal(mma, stream, nq, k, 1e-5f, "3-pass tcgen05 vs streaming kernel");
    Result screen = search(nq, k, MMF_ALGO_MMA, &ms_screen, 50);
    fails += !same(screen, stream, nq, k, "screened search vs streaming kernel");
    opt("epi_parity", 0);
    Result parity = search(nq, k, MMF_ALGO_MMA, &ms_parity, 50);
    opt("epi_parity", -1);
    fails += !same(parity, stream, nq, k, "screened search, epi_parity=0 vs streaming kernel");
    opt("epi_parity", -1);
    Result one = search(1, k, MMF_ALGO_STREAM, &ms_one, 50);
    float ms_one_reg = 0, ms8 = 0, ms8_reg = 0;
    Result eight = search(8, k, MMF_ALGO_STREAM, &ms8, 20);
    opt("stream_tma", 0);
    Result one_reg = search(1, k, MMF_ALGO_STREAM, &ms_one_reg, 50);
    Result eight_reg = search(8, k, MMF_ALGO_STREAM, &ms8_reg, 20);
    Result stream_reg = search(nq, k, MMF_ALGO_STREAM);
    opt("stream_tma", 1);
    fails += !same(one, one_reg, 1, k, "batch-1: TMA-staged vs register-staged streaming kernel");
    fails += !same(eight, eight_reg, 8, k, "batch-8: TMA-staged vs register-staged streaming kernel");
    fails += !same(stream, stream_reg, nq, k, "256 queries: TMA-staged vs register-staged streaming kernel");
    printf("  streaming kernel: batch-1 TMA-staged %.4f ms (%.0f GB/s), register-staged %.4f ms (%.0f GB/s); batch-8 %.4f vs %.4f ms\n",
           ms_one, rows_fp32 * 2048.0 / ms_one * 1e-6, ms_one_reg, rows_fp32 * 2048.0 / ms_one_reg * 1e-6, ms8, ms8_reg);
    printf("  search time: 3-pass %.3f ms (%.0f GB/s algorithmic), screened %.4f ms (%.0f GB/s algorithmic, %.0f GB/s of hi planes), "
           "screened with all warps on every tile %.4f ms, streaming x256 %.2f ms, batch-1 %.4f ms (%.0f GB/s)\n", ms_mma,
           rows_fp32 * 2048.0 / ms_mma * 1e-6, ms_screen, rows_fp32 * 2048.0 / ms_screen * 1e-6, rows_fp32 * 1024.0 / ms_screen * 1e-6,
           ms_parity, ms_stream, ms_one, rows_fp32 * 2048.0 / ms_one * 1e-6);
    // band overflow: 3000 identical rows -> guarded 3-pass redo, ties by row id
    // (positions adapt to small vaults, e.g. under compute-sanitizer)
    const long long dup_n = rows_fp32 >= 80000 ? 3000 : rows_fp32 / 4, dup_lo = rows_fp32 >= 80000 ? 70000 : rows_fp32 / 2;
    const long long dup_hi = dup_lo + dup_n;
    duplicate_rows<<<(unsigned)((dup_n * 512 + 255) / 256), 256>>>(d_vault, 5, dup_lo, dup_hi);
    CK(cudaMemcpy(d_q, d_vault + 5 * 512, 512 * 4, cudaMemcpyDeviceToDevice));   // query 0 = the duplicated row
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, rows_fp32, 512, MMF_F32, MMF_VAULT_FP32, 0));
    opt("screen", 0);
    Result mma2 = search(nq, k, MMF_ALGO_MMA);
    opt("screen", 1);
    Result screen2 = search(nq, k, MMF_ALGO_MMA);
    Result screen3 = search(nq, k, MMF_ALGO_MMA);       // twice: the flag and both counter sets must reset
    fails += !same(screen2, mma2, nq, k, "screened search, overflowing band vs 3-pass");
    fails += !same(screen3, mma2, nq, k, "screened search, overflowing band, second call");
    printf("  query 0 top rows: %lld %lld %lld (expect %lld %lld %lld)\n", (long long)screen2.rows[0],
           (long long)screen2.rows[1], (long long)screen2.rows[2], dup_hi - 1, dup_hi - 2, dup_hi - 3);
    fails += screen2.rows[0] != dup_hi - 1;
  }

  // ---------------- bf16 vault, 4096 queries, top-100 (the C4 shard shape) and top-10: run-to-run determinism,
  // planted rows, timing.  (Oracle parity of this shape: tests/test_gpu_parity.py::test_c4_shape_vs_oracle.)
  for (int pass = 0; pass < 2 && want("bf16"); ++pass) {
    const int nq = 4096;
    const long long rows_bf16 = pass == 0 ? rows_bf16_a : rows_bf16_b;
    if (rows_bf16 <= 0) continue;
    fill_rows<<<(unsigned)((rows_bf16 * 512 + 255) / 256), 256>>>(d_vault, rows_bf16, 21);
    fill_rows<<<(nq * 512 + 255) / 256, 256>>>(d_q, nq, 22);
    plant_queries<<<(64 * 512 + 255) / 256, 256>>>(d_q, d_vault, 64, rows_bf16 / 64, 1.5f);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, rows_bf16, 512, MMF_F32, MMF_VAULT_BF16, 0));
    printf("[bf16] %lld rows, %d queries\n", rows_bf16, nq);
    float ms100 = 0, ms10 = 0, ms10p = 0;
    Result a = search(nq, 100, MMF_ALGO_MMA, &ms100, 10);
    Result b = search(nq, 100, MMF_ALGO_MMA);
    fails += !same(b, a, nq, 100, "top-100, second run vs first");
    long long planted_ok = 0;
    for (int j = 0; j < 64; ++j) planted_ok += a.rows[(size_t)j * 100] == (long long)j * (rows_bf16 / 64);
    printf("  planted rows found first: %lld of 64\n", planted_ok);
    fails += planted_ok < 60;        // (a few planted queries carry more noise than signal: see plant_queries)
    Result s64 = search(64, 100, MMF_ALGO_STREAM);
    Result m64 = a;
    m64.scores.resize(64 * 100); m64.rows.resize(64 * 100); m64.disc.resize(64);
    fails += !near_equal(m64, s64, 64, 100, 1e-2f, "top-100 tcgen05 vs streaming kernel (first 64 queries, bf16 band)");
    Result c = search(nq, 10, MMF_ALGO_MMA, &ms10, 10);
    opt("epi_parity", 1);
    Result d = search(nq, 10, MMF_ALGO_MMA, &ms10p, 10);
    opt("epi_parity", -1);
    fails += !same(d, c, nq, 10, "top-10, epi_parity=1 vs default");
    {
      float t1 = 0, t0 = 0;
      Result s1 = search(1, 10, MMF_ALGO_STREAM, &t1, 30);
      opt("stream_tma", 0);
      Result s0 = search(1, 10, MMF_ALGO_STREAM, &t0, 30);
      opt("stream_tma", 1);
      fails += !same(s1, s0, 1, 10, "bf16 batch-1: TMA-staged vs register-staged");
      printf("  bf16 batch-1: TMA-staged %.4f ms (%.0f GB/s), register-staged %.4f ms (%.0f GB/s)\n", t1, rows_bf16 * 1024.0 / t1 * 1e-6, t0,
             rows_bf16 * 1024.0 / t0 * 1e-6);
    }
    const double fl = 2.0 * nq * rows_bf16 * 512;
    printf("  search time: top-100 %.3f ms (%.0f TFLOP/s), top-10 %.3f ms (%.0f TFLOP/s), top-10 + epi_parity %.3f ms (%.0f TFLOP/s)\n",
           ms100, fl / ms100 * 1e-9, ms10, fl / ms10 * 1e-9, ms10p, fl / ms10p * 1e-9);
  }
  // ---------------- sweep of small / ragged shapes (those of tests/test_gpu_parity.py): every tcgen05 variant
  // against the HBM-streaming kernel (an independent implementation: CUDA cores, fp32 FMAs)
  if (want("sweep")) {
    struct Shape { long long n; int nq, k; long long off; };
    const Shape shapes[] = {{1, 1, 1, 0}, {31, 3, 5, 0}, {128, 1, 1, 0}, {129, 130, 5, 0}, {150, 3, 12, 7}, {150, 3, 200, 0},
                            {1000, 1, 10, 0}, {2000, 40, 256, 0}, {3000, 64, 256, 100}, {4099, 2, 7, 0}, {5000, 16, 10, 0},
                            {20000, 160, 5, 0}, {30011, 19, 10, 15005}, {33333, 300, 100, 0}, {40000, 257, 10, 0},
                            {65536, 200, 16, 0}, {70001, 129, 17, 3}, {100000, 33, 10, 0}, {200000, 128, 32, 0},
                            {300000, 513, 100, 1000000}, {262144, 1024, 10, 0}, {50000, 4096, 5, 0}};
    printf("[sweep] tcgen05 search (screened / 3-pass / histogram bound) vs the streaming kernel\n");
    for (const Shape& sh : shapes) {
      fill_rows<<<(unsigned)((sh.n * 512 + 255) / 256), 256>>>(d_vault, sh.n, 100 + sh.n);
      fill_rows<<<(sh.nq * 512 + 255) / 256, 256>>>(d_q, sh.nq, 200 + sh.nq);
      const int n_plant = sh.nq < 8 ? 1 : sh.nq / 8;
      plant_queries<<<(n_plant * 512 + 255) / 256, 256>>>(d_q, d_vault, n_plant, sh.n / n_plant > 0 ? sh.n / n_plant : 0, 1.0f);
      CK(cudaDeviceSynchronize());
      char what[96];
      for (int mode = 0; mode < 2; ++mode) {
        MM(mmf_vault_load(H, d_vault, 1, sh.n, 512, MMF_F32, mode == 0 ? MMF_VAULT_FP32 : MMF_VAULT_BF16, sh.off));
        snprintf(what, sizeof what, "%s N=%lld Q=%d k=%d off=%lld", mode ? "bf16" : "fp32", sh.n, sh.nq, sh.k, sh.off);
        Result stream = search(sh.nq, sh.k, MMF_ALGO_STREAM);
        opt("stream_tma", 0);
        Result stream_reg = search(sh.nq, sh.k, MMF_ALGO_STREAM);
        opt("stream_tma", 1);
        fails += !same(stream, stream_reg, sh.nq, sh.k, "   streaming kernel: TMA-staged vs register-staged");
        Result def = search(sh.nq, sh.k, MMF_ALGO_MMA);
        if (mode == 0 && sh.k <= 16) {
          fails += !same(def, stream, sh.nq, sh.k, what);                     // screened search: bit-identical
          for (int pa = 0; pa < 2; ++pa) {
            opt("epi_parity", pa);
            Result par = search(sh.nq, sh.k, MMF_ALGO_MMA);
            fails += !same(par, stream, sh.nq, sh.k, pa ? "   + epi_parity=1" : "   + epi_parity=0");
          }
          opt("epi_parity", -1);
          opt("screen", 0);
          Result base = search(sh.nq, sh.k, MMF_ALGO_MMA);
          opt("screen", 1);
          fails += !near_equal(base, stream, sh.nq, sh.k, 1e-5f, "   (3-pass vs streaming kernel)");
        } else {
          fails += !near_equal(def, stream, sh.nq, sh.k, mode ? 1e-2f : 1e-5f, what);
          if (sh.k <= 16) {
            for (int pa = 0; pa < 2; ++pa) {
              opt("epi_parity", pa);
              Result par = search(sh.nq, sh.k, MMF_ALGO_MMA);
              fails += !same(par, def, sh.nq, sh.k, pa ? "   + epi_parity=1 vs default" : "   + epi_parity=0 vs default");
            }
            opt("epi_parity", -1);
          }
        }
      }
    }
  }
  // ---------------- mmf_score_batch_host (one call, host buffers) vs the device entry points
  if (want("host")) {
    printf("[score_batch_host] vs cosine + search + fusion through the device entry points\n");
    const int nq = 256, k = 10;
    const long long n = 300000;
    std::vector<float> w(MMF_FUSION_PARAMS);
    for (int i = 0; i < MMF_FUSION_PARAMS; ++i) w[i] = 0.3f * sinf(0.37f * i + 1.0f);
    MM(mmf_fusion_load(H, w.data()));
    fill_rows<<<(unsigned)((n * 512 + 255) / 256), 256>>>(d_vault, n, 41);
    fill_rows<<<(nq * 512 + 255) / 256, 256>>>(d_q, nq, 42);
    plant_queries<<<(40 * 512 + 255) / 256, 256>>>(d_q, d_vault, 40, n / 40, 0.6f);
    float* d_text = nullptr;
    CK(cudaMalloc(&d_text, (size_t)nq * 512 * 4));
    fill_rows<<<(nq * 512 + 255) / 256, 256>>>(d_text, nq, 43);
    CK(cudaDeviceSynchronize());
    MM(mmf_vault_load(H, d_vault, 1, n, 512, MMF_F32, MMF_VAULT_FP32, 0));
    std::vector<float> text((size_t)nq * 512), img((size_t)nq * 512), head((size_t)nq * 3);
    std::vector<uint8_t> mod(nq);
    CK(cudaMemcpy(text.data(), d_text, text.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(img.data(), d_q, img.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nq * 3; ++i) head[i] = 0.5f + 0.5f * sinf(1.3f * i);
    for (int i = 0; i < nq; ++i) mod[i] = (uint8_t)(i % 4);
    std::vector<float> sim(nq), disc(nq), vs((size_t)nq * k), x5((size_t)nq * 5), probs((size_t)nq * 2), conf(nq);
    std::vector<int64_t> vr((size_t)nq * k);
    std::vector<int32_t> verdict(nq);
    // (a) no modality mask: every piece must equal the device entry point it is made of
    MM(mmf_score_batch_host(H, text.data(), img.data(), head.data(), nullptr, nq, k, 0.85, MMF_ALGO_AUTO, sim.data(), disc.data(),
                            vs.data(), vr.data(), x5.data(), probs.data(), verdict.data(), conf.data()));
    Result ref = search(nq, k, MMF_ALGO_AUTO);
    Result got;
    got.scores = vs; got.rows = vr; got.disc = disc;
    fails += !same(got, ref, nq, k, "vault part");
    float* d_sim = nullptr; float* d_x = nullptr; float* d_p = nullptr;
    CK(cudaMalloc(&d_sim, nq * 4)); CK(cudaMalloc(&d_x, nq * 5 * 4)); CK(cudaMalloc(&d_p, nq * 2 * 4));
    MM(mmf_cosine_pairs(H, d_text, d_q, nq, 512, 0.0, d_sim, nullptr, nullptr));
    std::vector<float> sim_ref(nq), probs_ref((size_t)nq * 2);
    CK(cudaMemcpy(sim_ref.data(), d_sim, nq * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < nq; ++i) {
      bad += memcmp(&sim[i], &sim_ref[i], 4) != 0;
      const float want[5] = {head[i * 3], head[i * 3 + 1], head[i * 3 + 2], sim_ref[i], ref.disc[i]};
      bad += memcmp(&x5[(size_t)i * 5], want, 20) != 0;
    }
    CK(cudaMemcpy(d_x, x5.data(), x5.size() * 4, cudaMemcpyHostToDevice));
    MM(mmf_fusion_forward(H, d_x, nq, d_p, nullptr, nullptr, nullptr));
    CK(cudaMemcpy(probs_ref.data(), d_p, probs_ref.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nq * 2; ++i) bad += memcmp(&probs[i], &probs_ref[i], 4) != 0;
    for (int i = 0; i < nq; ++i) bad += verdict[i] != (probs[i * 2 + 1] > 0.5f) || conf[i] != (verdict[i] ? probs[i * 2 + 1] : probs[i * 2]);
    printf("  cosine / assembled scores / fusion probabilities / verdicts: %lld differences -> %s\n", bad, bad ? "MISMATCH" : "identical");
    fails += bad != 0;
    // (b) modality mask: skipped modalities are zeroed, fallback verdict rule (misinfo_forensics.py:884-899)
    MM(mmf_score_batch_host(H, text.data(), img.data(), head.data(), mod.data(), nq, k, 0.85, MMF_ALGO_AUTO, sim.data(), disc.data(),
                            vs.data(), vr.data(), x5.data(), probs.data(), verdict.data(), conf.data()));
    bad = 0;
    for (int i = 0; i < nq; ++i) {
      const bool t = mod[i] & 1, v = mod[i] & 2;
      bad += sim[i] != ((t && v) ? sim_ref[i] : 0.f);
      bad += disc[i] != (v ? ref.disc[i] : 0.f);
      float fake;
      if (mod[i] == 3) fake = probs_ref[i * 2 + 1];
      else fake = fminf(1.f, fmaxf(0.f, mod[i] == 1 ? head[i * 3 + 1] : mod[i] == 2 ? fmaxf(head[i * 3 + 2], disc[i]) : 0.5f));
      bad += probs[i * 2 + 1] != fake;
    }
    printf("  with a modality mask (text only / visual only / neither / both): %lld differences -> %s\n", bad, bad ? "MISMATCH" : "identical");
    fails += bad != 0;
    // time it: pageable host buffers here, so this is an upper bound for pinned ones
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float* pin = nullptr;
    CK(cudaMallocHost(&pin, ((size_t)nq * 512 * 2 + nq * 3) * 4));
    memcpy(pin, text.data(), text.size() * 4); memcpy(pin + (size_t)nq * 512, img.data(), img.size() * 4);
    memcpy(pin + (size_t)nq * 1024, head.data(), head.size() * 4);
    for (int rep = 0; rep < 3; ++rep)
      MM(mmf_score_batch_host(H, pin, pin + (size_t)nq * 512, pin + (size_t)nq * 1024, nullptr, nq, k, 0.85, MMF_ALGO_AUTO, sim.data(),
                              disc.data(), vs.data(), vr.data(), x5.data(), probs.data(), verdict.data(), conf.data()));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int rep = 0; rep < 20; ++rep)
      MM(mmf_score_batch_host(H, pin, pin + (size_t)nq * 512, pin + (size_t)nq * 1024, nullptr, nq, k, 0.85, MMF_ALGO_AUTO, sim.data(),
                              disc.data(), vs.data(), vr.data(), x5.data(), probs.data(), verdict.data(), conf.data()));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms = ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6) / 20;
    printf("  end to end, pinned host buffers, %d queries x %lld rows: %.3f ms per call (%.0f queries/s)\n", nq, n, ms, nq / ms * 1e3);
    // (c) two batches in flight (submit / collect): same results, copies overlap the kernels
    {
      std::vector<float> probs2((size_t)nq * 2), vs2((size_t)nq * k);
      std::vector<int64_t> vr2((size_t)nq * k);
      const float* t_in = pin; const float* i_in = pin + (size_t)nq * 512; const float* h_in = pin + (size_t)nq * 1024;
      MM(mmf_score_batch_host(H, t_in, i_in, h_in, nullptr, nq, k, 0.85, MMF_ALGO_AUTO, nullptr, nullptr, vs.data(), vr.data(),
                              nullptr, probs.data(), nullptr, nullptr));
      MM(mmf_score_batch_submit(H, 0, t_in, i_in, h_in, nullptr, nq, k, 0.85, MMF_ALGO_AUTO));
      MM(mmf_score_batch_submit(H, 1, t_in, i_in, h_in, nullptr, nq, k, 0.85, MMF_ALGO_AUTO));
      MM(mmf_score_batch_collect(H, 0, nullptr, nullptr, vs2.data(), vr2.data(), nullptr, probs2.data(), nullptr, nullptr));
      long long bad2 = memcmp(probs.data(), probs2.data(), probs.size() * 4) != 0 || memcmp(vr.data(), vr2.data(), vr.size() * 8) != 0 ||
                       memcmp(vs.data(), vs2.data(), vs.size() * 4) != 0;
      MM(mmf_score_batch_collect(H, 1, nullptr, nullptr, vs2.data(), vr2.data(), nullptr, probs2.data(), nullptr, nullptr));
      bad2 += memcmp(probs.data(), probs2.data(), probs.size() * 4) != 0 || memcmp(vr.data(), vr2.data(), vr.size() * 8) != 0;
      bad2 += mmf_score_batch_collect(H, 1, nullptr, nullptr, nullptr, nullptr, nullptr, probs2.data(), nullptr, nullptr) == MMF_OK;   // nothing pending
      printf("  submit / collect, 2 slots in flight: %s\n", bad2 ? "MISMATCH" : "identical to the synchronous call");
      fails += bad2 != 0;
      const int reps = 40;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      MM(mmf_score_batch_submit(H, 0, t_in, i_in, h_in, nullptr, nq, k, 0.85, MMF_ALGO_AUTO));
      for (int rep = 1; rep < reps; ++rep) {
        MM(mmf_score_batch_submit(H, rep & 1, t_in, i_in, h_in, nullptr, nq, k, 0.85, MMF_ALGO_AUTO));
        MM(mmf_score_batch_collect(H, (rep - 1) & 1, nullptr, nullptr, vs2.data(), vr2.data(), nullptr, probs2.data(), nullptr, nullptr));
      }
      MM(mmf_score_batch_collect(H, (reps - 1) & 1, nullptr, nullptr, vs2.data(), vr2.data(), nullptr, probs2.data(), nullptr, nullptr));
      clock_gettime(CLOCK_MONOTONIC, &t1);
      const double ms2 = ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6) / reps;
      printf("  end to end, 2 batches in flight: %.3f ms per batch (%.0f queries/s)\n", ms2, nq / ms2 * 1e3);
    }
    cudaFreeHost(pin); cudaFree(d_text); cudaFree(d_sim); cudaFree(d_x); cudaFree(d_p);
  }
  // ---------------- peer-memory candidate exchange (csrc/exchange.cu): the protocol of 2 ranks on ONE device.
  // Kernels that wait for one another must never share a GPU (nothing guarantees that the one waited for runs;
  // B200_PROFILING.md: Xid 109), so both "ranks" use ONE stream and every exchange is enqueued as
  // push(rank 0), push(rank 1), merge(rank 0), merge(rank 1): when a merge kernel starts, all flags are already
  // set and nothing ever spins.  What this checks: slot layout, flags, epochs / parities, the merge itself.
  if (want("exchange")) {
    printf("[exchange] 2 ranks' protocol, phases in sequence on one stream, vs the unsharded search\n");
    mmf_handle* R[2] = {nullptr, nullptr};
    cudaStream_t S;
    CK(cudaStreamCreateWithFlags(&S, cudaStreamNonBlocking));
    for (int r = 0; r < 2; ++r)
      if (mmf_create(0, &R[r]) != MMF_OK) { printf("mmf_create failed\n"); return 1; }
    const long long n = 200001;
    const int nq = 24;
    int64_t need = 0;
    mmf_exchange_layout(2, nq, 100, nullptr, &need);
    void* B[2];
    uint64_t ptrs[2];
    for (int r = 0; r < 2; ++r) { CK(cudaMalloc(&B[r], (size_t)need)); ptrs[r] = (uint64_t)(uintptr_t)B[r]; }
    float* sc[2]; int64_t* ro[2]; float* di[2];
    for (int r = 0; r < 2; ++r) {
      CK(cudaMalloc(&sc[r], (size_t)nq * 100 * 4)); CK(cudaMalloc(&ro[r], (size_t)nq * 100 * 8)); CK(cudaMalloc(&di[r], nq * 4));
    }
    fill_rows<<<(unsigned)((n * 512 + 255) / 256), 256>>>(d_vault, n, 31);
    fill_rows<<<(nq * 512 + 255) / 256, 256>>>(d_q, nq, 32);
    plant_queries<<<(6 * 512 + 255) / 256, 256>>>(d_q, d_vault, 6, n / 6, 1.0f);
    CK(cudaDeviceSynchronize());
    const long long half = (n + 1) / 2;
    for (int mode = 0; mode < 2; ++mode) {
      const int vm = mode == 0 ? MMF_VAULT_FP32 : MMF_VAULT_BF16;
      MM(mmf_vault_load(H, d_vault, 1, n, 512, MMF_F32, vm, 0));
      for (int r = 0; r < 2; ++r) {
        int rc = mmf_vault_load(R[r], d_vault + (r ? half * 512 : 0), 1, r ? n - half : half, 512, MMF_F32, vm, r ? half : 0);
        if (rc == MMF_OK) rc = mmf_exchange_attach(R[r], r, 2, ptrs, need);
        if (rc != MMF_OK) { printf("rank %d setup failed: %s\n", r, mmf_last_error(R[r])); return 1; }
      }
      const int ks[3] = {10, 100, 5};
      for (int t = 0; t < 3; ++t) {
        const int k = ks[t];
        for (int algo = MMF_ALGO_STREAM; algo <= MMF_ALGO_MMA; ++algo) {
          Result full = search(nq, k, algo);
          for (int fused = 0; fused < 2; ++fused) {   // separate push kernel / push fused into the merge tail
            for (int r = 0; r < 2; ++r) mmf_set_option(R[r], "fused_push", fused);
            for (int rep = 0; rep < 3; ++rep) {       // 3 exchanges in a row: both parities + buffer reuse
              for (int r = 0; r < 2; ++r) {
                int rc = mmf_vault_search_push(R[r], d_q, nq, k, algo, S);
                if (rc != MMF_OK) { printf("search_push failed on rank %d: %s\n", r, mmf_last_error(R[r])); return 1; }
              }
              for (int r = 0; r < 2; ++r) {
                int rc = mmf_vault_exchange_merge(R[r], k, 0.85, sc[r], ro[r], di[r], S);
                if (rc != MMF_OK) { printf("exchange_merge failed on rank %d: %s\n", r, mmf_last_error(R[r])); return 1; }
              }
            }
            CK(cudaDeviceSynchronize());
            for (int r = 0; r < 2; ++r) {
              Result got;
              got.scores.resize((size_t)nq * k); got.rows.resize((size_t)nq * k); got.disc.resize(nq);
              CK(cudaMemcpy(got.scores.data(), sc[r], got.scores.size() * 4, cudaMemcpyDeviceToHost));
              CK(cudaMemcpy(got.rows.data(), ro[r], got.rows.size() * 8, cudaMemcpyDeviceToHost));
              CK(cudaMemcpy(got.disc.data(), di[r], got.disc.size() * 4, cudaMemcpyDeviceToHost));
              char what[112];
              snprintf(what, sizeof what, "%s k=%d %s rank %d%s", mode ? "bf16" : "fp32", k,
                       algo == MMF_ALGO_MMA ? "tcgen05" : "stream", r, fused ? " (fused push requested)" : "");
              fails += !same(got, full, nq, k, what);
            }
          }
        }
      }
    }
    for (int r = 0; r < 2; ++r) { mmf_destroy(R[r]); cudaFree(B[r]); cudaFree(sc[r]); cudaFree(ro[r]); cudaFree(di[r]); }
    cudaStreamDestroy(S);
  }
  printf("launches: %lld; %s\n", (long long)mmf_launch_count(H), fails ? "SELFTEST FAILED" : "selftest ok");
  mmf_destroy(H);
  return fails ? 1 : 0;
}
