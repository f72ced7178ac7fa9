// How fast can ONE producer thread per SM stream the vault through shared memory with TMA, and does the shape of
// the requests matter?  Decides the loader of the screened tcgen05 search (hi planes only) and
// of the TMA-staged batch-1 kernel.  148 CTAs, each sweeps its own contiguous slab of rows through an mbarrier
// ring; a consumer warp waits for every stage, reads 16 bytes of it and hands the slot back (no arithmetic: this
// measures the load pipeline alone).  Vault layout as in the library: [row][hi 1 KB | lo 1 KB] fp16.
//
//   mode 0  bulk1d      cp.async.bulk 1-D copies of whole rows (hi+lo, contiguous): STAGE bytes per copy
//   mode 1  box_kblk    3-D tensor map, box {64 el, 1 plane, 64 rows} = 8 KB: one 128 B slice of 64 rows per request
//                       (what vault_mma_topk_kernel<VAR_SCREEN> does today: 2 boxes per 16 KB stage)
//   mode 2  box_rows    tensor map (64 el, 8 k-blocks, rows), box {64, 8, 8 rows} = 8 KB: 8 whole 1 KB hi planes per
//                       request (UMMA-compatible: an 8-row swizzle atom per k-block, SBO = 8 KB)
//   mode 3  bulk1d_hi   cp.async.bulk 1-D copies of single 1 KB hi planes (one request per row)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_micro tools/tma_stream_micro.cu
//   tools/tma_stream_micro [rows (default 2M)]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(u64* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, u64* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// STAGE_BYTES per ring slot, STAGES slots; every CTA sweeps rows [blockIdx.x * rows_per_cta, ...)
template <int MODE, int STAGE_BYTES, int STAGES>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const unsigned char* vault,
                                                       long long rows, long long rows_per_cta, unsigned* sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  u64* full = reinterpret_cast<u64*>(smem + (size_t)STAGES * STAGE_BYTES);
  u64* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long r_begin = blockIdx.x * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
  // rows per stage: modes 0: STAGE/2048 whole rows; 1-3: STAGE/1024 hi planes
  constexpr int ROWS_PER_STAGE = (MODE == 0) ? STAGE_BYTES / 2048 : STAGE_BYTES / 1024;
  const long long n_stages = (r_end - r_begin + ROWS_PER_STAGE - 1) / ROWS_PER_STAGE;
  if (warp == 0) {
    if (lane == 0) {
      for (long long it = 0; it < n_stages; ++it) {
        const int s = (int)(it % STAGES);
        mbar_wait(empty + s, (u32)(((it / STAGES) & 1) ^ 1));
        mbar_expect_tx(full + s, STAGE_BYTES);
        unsigned char* dst = smem + (size_t)s * STAGE_BYTES;
        const long long row = r_begin + it * ROWS_PER_STAGE;      // (the last stage may run past r_end: the buffer is padded)
        if (MODE == 0) {
          bulk_load(dst, vault + row * 2048, STAGE_BYTES, full + s);
        } else if (MODE == 1) {
          // the search kernel's order: per 64-row group, k-blocks 0..7, 8 KB each; a stage holds STAGE/8 KB of them
          constexpr int BOXES = STAGE_BYTES / 8192;
#pragma unroll
          for (int b = 0; b < BOXES; ++b) {
            const long long box = it * BOXES + b;                 // global box index of this CTA
            const long long grp = box / 8; const int kb = (int)(box % 8);
            tma_load_3d(dst + b * 8192, &tm, full + s, kb * 64, 0, (int)(r_begin + grp * 64));
          }
        } else if (MODE == 2) {
          constexpr int BOXES = STAGE_BYTES / 8192;
#pragma unroll
          for (int b = 0; b < BOXES; ++b) tma_load_3d(dst + b * 8192, &tm, full + s, 0, 0, (int)(row + b * 8));
        } else {
#pragma unroll 4
          for (int r = 0; r < ROWS_PER_STAGE; ++r) bulk_load(dst + r * 1024, vault + (row + r) * 2048, 1024, full + s);
        }
      }
    }
  } else {
    unsigned acc = 0;
    for (long long it = 0; it < n_stages; ++it) {
      const int s = (int)(it % STAGES);
      mbar_wait(full + s, (u32)((it / STAGES) & 1));
      acc ^= *reinterpret_cast<const unsigned*>(smem + (size_t)s * STAGE_BYTES + lane * 4);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    if (acc == 0x12345678u) *sink = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int STAGE_BYTES, int STAGES>
static void run(const char* what, const CUtensorMap& tm, const unsigned char* vault, long long rows, unsigned* sink, int sms) {
  const int smem = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;
  auto kern = stream_kernel<MODE, STAGE_BYTES, STAGES>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_per_cta = ((rows + sms - 1) / sms + 63) / 64 * 64;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<sms, 64, smem>>>(tm, vault, rows, rows_per_cta, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = (double)rows * (MODE == 0 ? 2048 : 1024);
  printf("  %-58s %2d x %3d KB ring : %7.0f GB/s  (%.3f ms)\n", what, STAGES, STAGE_BYTES / 1024, bytes / (best * 1e-3) / 1e9, best);
}

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? atoll(argv[1]) : 2000000;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  unsigned char* vault; unsigned* sink;
  const size_t bytes = (size_t)(rows + 64 * 1024) * 2048;          // padding: the last stage of a CTA may run past its slab
  CK(cudaMalloc(&vault, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(vault, 1, bytes));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const cuuint32_t estr[3] = {1, 1, 1};
  const cuuint64_t n = (cuuint64_t)rows + 64 * 1024;
  CUtensorMap tm_kblk, tm_rows;
  {   // [row][plane][512] viewed as (k, plane, row): the library's map
    const cuuint64_t dims[3] = {512, 2, n}; const cuuint64_t strides[2] = {1024, 2048}; const cuuint32_t box[3] = {64, 1, 64};
    if (encode(&tm_kblk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, vault, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 1 failed\n"); return 2; }
  }
  {   // hi plane viewed as (64 el, 8 k-blocks, row): box = 8 whole hi planes
    const cuuint64_t dims[3] = {64, 8, n}; const cuuint64_t strides[2] = {128, 2048}; const cuuint32_t box[3] = {64, 8, 8};
    if (encode(&tm_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, vault, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 2 failed\n"); return 2; }
  }
  printf("%s, %d SMs, %lld rows of 2 KB (hi 1 KB | lo 1 KB)\n", prop.name, sms, rows);
  run<0, 16384, 8>("bulk1d: whole rows, 16 KB copies", tm_kblk, vault, rows, sink, sms);
  run<0, 32768, 6>("bulk1d: whole rows, 32 KB copies", tm_kblk, vault, rows, sink, sms);
  run<0, 16384, 12>("bulk1d: whole rows, 16 KB copies", tm_kblk, vault, rows, sink, sms);
  run<0, 8192, 24>("bulk1d: whole rows, 8 KB copies", tm_kblk, vault, rows, sink, sms);
  run<1, 16384, 8>("box_kblk: hi planes, 128 B slices of 64 rows (as today)", tm_kblk, vault, rows, sink, sms);
  run<1, 16384, 12>("box_kblk: hi planes, 128 B slices of 64 rows", tm_kblk, vault, rows, sink, sms);
  run<1, 65536, 3>("box_kblk: hi planes, whole 64-row tile share per stage", tm_kblk, vault, rows, sink, sms);
  run<2, 16384, 8>("box_rows: hi planes, 8 whole rows per request", tm_rows, vault, rows, sink, sms);
  run<2, 16384, 12>("box_rows: hi planes, 8 whole rows per request", tm_rows, vault, rows, sink, sms);
  run<2, 65536, 3>("box_rows: hi planes, 64-row tile share per stage", tm_rows, vault, rows, sink, sms);
  run<3, 16384, 8>("bulk1d_hi: hi planes, one 1 KB copy per row", tm_kblk, vault, rows, sink, sms);
  run<3, 16384, 12>("bulk1d_hi: hi planes, one 1 KB copy per row", tm_kblk, vault, rows, sink, sms);
  return 0;
}
