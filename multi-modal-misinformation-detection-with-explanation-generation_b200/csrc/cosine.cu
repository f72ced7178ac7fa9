// K1: caption <-> image cosine (normalise + dot), one warp per pair.
// Replaces misinfo_forensics.py:399-404, :481-484 and clip_similarity_engine.py:103-111.
//
// HBM-bound: 2*dim*4 B in, 4 (+1) B out per pair (4 100 B at dim 512); each lane issues
// dim/64 independent 128-bit loads before any arithmetic, three running sums (a.b, a.a,
// b.b) in one pass, xor-shuffle reductions, cos = a.b / (|a| |b|).  No eps, like the
// reference: a zero embedding gives NaN.
#include "common.cuh"

namespace mmf {

template <int NCH>   // float4 chunks per lane (dim = 128*NCH); 0 = generic dim
__global__ void __launch_bounds__(256) cosine_pairs_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           long long n_pairs, int dim, double match_threshold,
                                                           float* __restrict__ out_sim,
                                                           unsigned char* __restrict__ out_match) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; p < n_pairs; p += warps) {
    float ab = 0.f, aa = 0.f, bb = 0.f;
    if (NCH > 0) {
      const uint4* pa = reinterpret_cast<const uint4*>(a + p * dim);
      const uint4* pb = reinterpret_cast<const uint4*>(b + p * dim);
      uint4 va[NCH > 0 ? NCH : 1], vb[NCH > 0 ? NCH : 1];
#pragma unroll
      for (int c = 0; c < NCH; ++c) { va[c] = ldg_stream(pa + c * 32 + lane); vb[c] = ldg_stream(pb + c * 32 + lane); }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float x[4] = {__uint_as_float(va[c].x), __uint_as_float(va[c].y), __uint_as_float(va[c].z), __uint_as_float(va[c].w)};
        const float y[4] = {__uint_as_float(vb[c].x), __uint_as_float(vb[c].y), __uint_as_float(vb[c].z), __uint_as_float(vb[c].w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) { ab = fmaf(x[e], y[e], ab); aa = fmaf(x[e], x[e], aa); bb = fmaf(y[e], y[e], bb); }
      }
    } else {
      for (int i = lane; i < dim; i += 32) {
        const float x = a[p * dim + i], y = b[p * dim + i];
        ab = fmaf(x, y, ab); aa = fmaf(x, x, aa); bb = fmaf(y, y, bb);
      }
    }
    ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
    if (lane == 0) {
      const float sim = ab / (sqrtf(aa) * sqrtf(bb));
      out_sim[p] = sim;
      if (out_match) out_match[p] = ((double)sim >= match_threshold) ? 1 : 0;
    }
  }
}

}  // namespace mmf

extern "C" int mmf_cosine_pairs(mmf_handle* h, const float* a, const float* b, int64_t n_pairs, int dim,
                                double match_threshold, float* out_sim, uint8_t* out_match, mmf_stream_t stream) {
  if (!h) return MMF_ERR_BAD_ARG;
  if (n_pairs < 0 || dim <= 0 || (n_pairs > 0 && (!a || !b || !out_sim)))
    return mmf_set_error(h, MMF_ERR_BAD_ARG, "cosine_pairs: bad argument (n_pairs=%lld dim=%d)", (long long)n_pairs, dim);
  if (n_pairs == 0) return MMF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long want = (n_pairs + 7) / 8;
  const int grid = (int)(want < (long long)h->sm_count * 8 ? want : (long long)h->sm_count * 8);
  const bool vec = (dim % 128 == 0) && ((((uintptr_t)a | (uintptr_t)b) & 15) == 0);
  if (vec && dim == 512)
    mmf::cosine_pairs_kernel<4><<<grid, 256, 0, st>>>(a, b, n_pairs, dim, match_threshold, out_sim, out_match);
  else if (vec && dim == 768)
    mmf::cosine_pairs_kernel<6><<<grid, 256, 0, st>>>(a, b, n_pairs, dim, match_threshold, out_sim, out_match);
  else if (vec && dim == 1024)
    mmf::cosine_pairs_kernel<8><<<grid, 256, 0, st>>>(a, b, n_pairs, dim, match_threshold, out_sim, out_match);
  else
    mmf::cosine_pairs_kernel<0><<<grid, 256, 0, st>>>(a, b, n_pairs, dim, match_threshold, out_sim, out_match);
  MMF_LAUNCH_OK(h);
  return MMF_OK;
}
