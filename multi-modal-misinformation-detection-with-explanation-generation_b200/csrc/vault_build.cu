// Vault load: the whole-vault renormalisation the reference repeats on EVERY query
// (misinfo_forensics.py:443-445, ~88% of its per-query time) is hoisted here and done once:
// rows are L2-normalised in fp32 (v / ||v||, no eps -> a zero row becomes NaN exactly like
// NumPy's 0/0) and written in the resident layout of the chosen mode (mmf_b200.h).
#include "common.cuh"

#include <algorithm>

namespace mmf {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<double>(double v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// One warp per row; lane owns elements {32*j + lane}.  HBM-bound, run once.
template <typename T>
__global__ void __launch_bounds__(256) vault_normalise_kernel(const T* __restrict__ src, long long n_rows, int mode,
                                                              void* __restrict__ dst, long long dst_row0,
                                                              unsigned long long* __restrict__ nan_rows) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; r < n_rows; r += warps) {
    float v[MMF_DIM / 32];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < MMF_DIM / 32; ++j) {
      v[j] = to_f32<T>(src[r * MMF_DIM + j * 32 + lane]);
      ss = fmaf(v[j], v[j], ss);
    }
    const float norm = sqrtf(warp_sum(ss));
    if (lane == 0 && !(norm > 0.f && norm < INFINITY)) atomicAdd(nan_rows, 1ull);   // row becomes NaN (0/0) like NumPy
    if (mode == MMF_VAULT_BF16) {
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst) + (dst_row0 + r) * MMF_DIM;
#pragma unroll
      for (int j = 0; j < MMF_DIM / 32; ++j) out[j * 32 + lane] = __float2bfloat16_rn(v[j] / norm);
    } else {
      __half* hi = reinterpret_cast<__half*>(dst) + (dst_row0 + r) * (2 * MMF_DIM);
      __half* lo = hi + MMF_DIM;
#pragma unroll
      for (int j = 0; j < MMF_DIM / 32; ++j) {
        const float y = (v[j] / norm) * MMF_SPLIT_SCALE;     // |y| <= 256, exact scaling
        const __half h = __float2half_rn(y);
        hi[j * 32 + lane] = h;
        lo[j * 32 + lane] = __float2half_rn(y - __half2float(h));   // residual is exact in fp32
      }
    }
  }
}

}  // namespace mmf

static size_t src_elem_size(int dt) {
  switch (dt) {
    case MMF_F32: return 4;
    case MMF_F16: return 2;
    case MMF_BF16: return 2;
    case MMF_F64: return 8;
  }
  return 0;
}

static int launch_normalise(mmf_handle* h, const void* dev_src, int dt, long long n, int mode, void* dst,
                            long long dst_row0, cudaStream_t st) {
  unsigned long long* nan_rows = h->vault_nan_rows_dev;
  const long long want = (n + 7) / 8;
  const int grid = (int)std::min<long long>(want, (long long)h->sm_count * 16);
  switch (dt) {
    case MMF_F32: mmf::vault_normalise_kernel<float><<<grid, 256, 0, st>>>((const float*)dev_src, n, mode, dst, dst_row0, nan_rows); break;
    case MMF_F16: mmf::vault_normalise_kernel<__half><<<grid, 256, 0, st>>>((const __half*)dev_src, n, mode, dst, dst_row0, nan_rows); break;
    case MMF_BF16: mmf::vault_normalise_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dev_src, n, mode, dst, dst_row0, nan_rows); break;
    case MMF_F64: mmf::vault_normalise_kernel<double><<<grid, 256, 0, st>>>((const double*)dev_src, n, mode, dst, dst_row0, nan_rows); break;
  }
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}

int mmf_mma_vault_changed(mmf_handle* h);   // vault_mma.cu: rebuilds TMA descriptors

extern "C" int mmf_vault_unload(mmf_handle* h) {
  if (!h) return MMF_ERR_BAD_ARG;
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  if (h->vault) {
    MMF_CUDA_OK(h, cudaDeviceSynchronize());
    MMF_CUDA_OK(h, cudaFree(h->vault));
  }
  h->vault = nullptr;
  h->vault_loaded = false;
  h->vault_rows = 0;
  h->vault_nan_rows = 0;
  h->vault_bytes = 0;
  h->vault_row_offset = 0;
  return mmf_mma_vault_changed(h);
}

extern "C" int mmf_vault_load(mmf_handle* h, const void* rows, int rows_on_device, int64_t n_rows, int dim,
                              int src_dtype, int vault_mode, int64_t row_offset) {
  if (!h) return MMF_ERR_BAD_ARG;
  const size_t es = src_elem_size(src_dtype);
  if (n_rows < 0 || (n_rows > 0 && !rows) || es == 0 || row_offset < 0 ||
      (vault_mode != MMF_VAULT_FP32 && vault_mode != MMF_VAULT_BF16))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "vault_load: bad argument (n_rows=%lld dtype=%d mode=%d)",
                         (long long)n_rows, src_dtype, vault_mode);
  if (dim != MMF_DIM)
    return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "vault_load: dim %d unsupported (CLIP ViT-B/32 projection dim %d only)", dim, MMF_DIM);
  if ((unsigned long long)row_offset + (unsigned long long)n_rows > 0xFFFFFFFFull)
    return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "vault_load: global row ids must fit 32 bits");
  MMF_CUDA_OK(h, cudaSetDevice(h->device));
  MMF_CUDA_OK(h, cudaDeviceSynchronize());   // a device-resident source must be complete
  int rc = mmf_vault_unload(h);
  if (rc != MMF_OK) return rc;
  if (n_rows == 0) {                      // an empty vault is "loaded" (the reference would search 0 rows)
    h->vault_mode = vault_mode;
    h->vault_row_offset = row_offset;
    h->vault_loaded = true;
    return MMF_OK;
  }
  const size_t row_bytes = (vault_mode == MMF_VAULT_BF16) ? MMF_DIM * 2 : MMF_DIM * 4;
  const size_t bytes = (size_t)n_rows * row_bytes;
  if (cudaMalloc(&h->vault, bytes) != cudaSuccess) {
    cudaGetLastError();
    h->vault = nullptr;
    return mmf_set_error(h, MMF_ERR_NOMEM, "vault_load: cannot allocate %zu bytes of HBM", bytes);
  }
  cudaStream_t st = h->own_stream;
  if (!h->vault_nan_rows_dev) MMF_CUDA_OK(h, cudaMalloc(&h->vault_nan_rows_dev, sizeof(unsigned long long)));
  MMF_CUDA_OK(h, cudaMemsetAsync(h->vault_nan_rows_dev, 0, sizeof(unsigned long long), st));
  if (rows_on_device) {
    rc = launch_normalise(h, rows, src_dtype, n_rows, vault_mode, h->vault, 0, st);
    if (rc != MMF_OK) return rc;
  } else {
    const long long chunk = std::min<long long>(n_rows, 32768);      // <= 128 MB of fp64 staging
    void* stage = nullptr;
    MMF_CUDA_OK(h, cudaMalloc(&stage, (size_t)chunk * MMF_DIM * es));
    for (long long r0 = 0; r0 < n_rows; r0 += chunk) {
      const long long n = std::min<long long>(chunk, n_rows - r0);
      cudaError_t e = cudaMemcpyAsync(stage, (const char*)rows + (size_t)r0 * MMF_DIM * es, (size_t)n * MMF_DIM * es,
                                      cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) {
        rc = launch_normalise(h, stage, src_dtype, n, vault_mode, h->vault, r0, st);
        if (rc == MMF_OK) e = cudaStreamSynchronize(st);
      }
      if (e != cudaSuccess || rc != MMF_OK) {
        cudaFree(stage);
        if (rc != MMF_OK) return rc;
        return mmf_set_error(h, MMF_ERR_CUDA, "vault_load: upload failed: %s", cudaGetErrorString(e));
      }
    }
    MMF_CUDA_OK(h, cudaFree(stage));
  }
  unsigned long long nan_rows = 0;
  MMF_CUDA_OK(h, cudaMemcpyAsync(&nan_rows, h->vault_nan_rows_dev, sizeof nan_rows, cudaMemcpyDeviceToHost, st));
  MMF_CUDA_OK(h, cudaStreamSynchronize(st));
  h->vault_nan_rows = (int64_t)nan_rows;
  h->vault_loaded = true;
  h->vault_rows = n_rows;
  h->vault_bytes = bytes;
  h->vault_mode = vault_mode;
  h->vault_row_offset = row_offset;
  return mmf_mma_vault_changed(h);
}

extern "C" int mmf_vault_info(const mmf_handle* h, int64_t* n_rows, int* dim, int* vault_mode, int64_t* row_offset) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_rows) *n_rows = h->vault_rows;
  if (dim) *dim = MMF_DIM;
  if (vault_mode) *vault_mode = h->vault_mode;
  if (row_offset) *row_offset = h->vault_row_offset;
  return h->vault_loaded ? MMF_OK : MMF_ERR_NOT_LOADED;
}
