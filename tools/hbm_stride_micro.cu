// Does reading only the first 1 KB of every 2 KB (the fp16 hi planes of the fp32-exact vault layout
// [row][hi 1 KB | lo 1 KB], which is what the screened search streams) cost HBM efficiency against reading the
// same number of bytes contiguously?  Decides whether a plane-separated vault layout is worth having
// (DESIGN.md section 8).  Plain 128-bit loads, a warp per row, 4 rows in flight per warp like vault_stream.cu.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_stride_micro tools/hbm_stride_micro.cu
//   tools/hbm_stride_micro [rows (default 4M)]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// every warp reads `piece` bytes (1024 or 2048) at the start of each `stride`-byte row, 4 rows in flight
template <int PIECE_U4>
__global__ void __launch_bounds__(256) read_rows(const uint4* __restrict__ base, long long rows, long long stride_u4, unsigned* sink) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned acc = 0;
  for (long long r0 = warp * 4; r0 < rows; r0 += warps * 4) {
    uint4 v[4][PIECE_U4 / 32];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < PIECE_U4 / 32; ++c)
        v[u][c] = (r0 + u < rows) ? ldg_stream(base + (r0 + u) * stride_u4 + c * 32 + lane) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < PIECE_U4 / 32; ++c) acc ^= v[u][c].x ^ v[u][c].y ^ v[u][c].z ^ v[u][c].w;
  }
  if (acc == 0x12345678u) *sink = acc;     // keeps the loads alive
}

template <int PIECE_U4>
static double run(const uint4* buf, long long rows, long long stride_u4, unsigned* sink, int sms) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    read_rows<PIECE_U4><<<sms * 8, 256>>>(buf, rows, stride_u4, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  return (double)rows * PIECE_U4 * 16 / (best * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? atoll(argv[1]) : 4000000;      // 8 GB of 2 KB rows: far beyond L2
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  uint4* buf; unsigned* sink;
  CK(cudaMalloc(&buf, (size_t)rows * 2048)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, (size_t)rows * 2048));
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, %lld rows of 2 KB\n", prop.name, sms, rows);
  printf("  whole rows, contiguous (2 KB of every 2 KB)        : %8.0f GB/s\n", run<128>(buf, rows, 128, sink, sms));
  printf("  hi planes only, interleaved (1 KB of every 2 KB)   : %8.0f GB/s of useful bytes\n", run<64>(buf, rows, 128, sink, sms));
  printf("  hi planes only, plane-separated (1 KB rows, dense) : %8.0f GB/s\n", run<64>(buf, rows, 64, sink, sms));
  printf("  (ratio of the last two lines = what a plane-separated vault layout could gain for the screened search)\n");
  return 0;
}
