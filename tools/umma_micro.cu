// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) issued back to back by
// one thread per SM, for N in {64,128,256}, A from shared memory (SS) or tensor memory (TS), and for
// 1 or 2 alternating accumulators.  Operands are zeros; only timing matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_micro tools/umma_micro.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u64 desc(u32 saddr) {
  return (u64)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr u32 idesc(u32 fmt, u32 m, u32 n) { return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }
__device__ __forceinline__ void mma_ss(u32 d, u64 a, u64 b, u32 id, u32 acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(u32 d, u32 a, u64 b, u32 id, u32 acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ u64 bar; __shared__ u32 slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((u32*)smem)[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const u32 tm = slot;
  if (threadIdx.x == 0) {
    const u64 a = desc(smem_u32(smem)), b = desc(smem_u32(smem + 16384));
    constexpr u32 ID = idesc(1, 128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const u32 d = tm + ((NACC == 2 && (i & 1)) ? 256 : 0);
        if (TS) mma_ts(d, tm + 256 + (NACC == 2 ? 128 : 0) + kk * 8, b + 2 * kk, ID, 1);   // A columns (garbage data)
        else mma_ss(d, a + 2 * kk, b + 2 * kk, ID, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm)); }
}
// knobs: COMMIT = tcgen05.commit after every 4 MMAs; STAGES = cycle B over that many 16 KB tiles;
// SPIN = number of extra warps spinning on an mbarrier; RANDOM = non-zero operand data
template <bool TS, bool COMMIT, int STAGES, int SPIN, bool RANDOM, int BUSY = 0>
__global__ void __launch_bounds__(32 + 32 * SPIN + 32, 1) k2(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ u64 bar, spin_bar, sink_bar; __shared__ u32 slot;
  const int nwords = (16384 + STAGES * 16384) / 4;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
    u32 h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    // two bf16 values in [-1,1): sign/exponent chosen so they are normal numbers
    ((u32*)smem)[i] = RANDOM ? (((h & 0x807F807Fu) | 0x3F003F00u)) : 0u;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&spin_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&sink_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const u32 tm = slot;
  if (threadIdx.x == 0) {
    const u64 a = desc(smem_u32(smem));
    constexpr u32 ID = idesc(1, 128, 128);
    long long t0 = clock64();
    int st = 0;
    for (int i = 0; i < iters; ++i) {
      const u64 b = desc(smem_u32(smem + 16384 + st * 16384));
      st = (st + 1 == STAGES) ? 0 : st + 1;
      const u32 d = tm + ((i >> 3) & 1) * 128;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (TS) mma_ts(d, tm + 256 + (i & 7) * 32 + kk * 8, b + 2 * kk, ID, 1);
        else mma_ss(d, a + 2 * kk, b + 2 * kk, ID, 1);
      }
      if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sink_bar)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&spin_bar)) : "memory");
  } else if (threadIdx.x >= 64 && BUSY) {
    // ALU-busy helper warps (BUSY=1: fmax/compare chains like the epilogue filter; BUSY=2: also TMEM reads)
    float a = threadIdx.x, b = 1.0f, c = 0.5f;
    u32 done = 0;
    while (!done) {
#pragma unroll
      for (int i = 0; i < 64; ++i) { a = fmaxf(a * 1.0001f, b); b = fmaxf(b + c, a * 0.5f); c = fmaf(c, 0.999f, 1e-3f); }
      if (BUSY == 2) {
        u32 v0, v1, v2, v3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tm + (((threadIdx.x >> 5) & 3) * 32 << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        a += __uint_as_float(v0 ^ v1 ^ v2 ^ v3) * 0.f;
      }
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&spin_bar)) : "memory");
    }
    if (a + b + c == 12345.678f) out[1] = 1;
  } else if (threadIdx.x >= 64) {
    asm volatile("{\n\t.reg .pred p;\n\tW3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D3;\n\tbra W3;\n\tD3:\n\t}" ::"r"(smem_u32(&spin_bar)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm)); }
}

// k3: how much tensor-memory READ bandwidth is left for the epilogue while the tensor pipe runs?
// NRD reader warps issue tcgen05.ld.32x32b.x32 (4 KB per warp instruction: 32 lanes x 32 accumulator columns) +
// tcgen05.wait::ld back to back on the accumulator that the MMAs are NOT writing; thread 0 issues MMA = 0 (none),
// 1 (SS: A from shared memory) or 2 (TS: A from tensor memory) MMAs of M=128 N=128 K=16 back to back, alternating
// two accumulators every 32 MMAs like the search kernel.  Reports clk / MMA and bytes / clk read by the epilogue.
template <int MMA, int NRD>
__global__ void __launch_bounds__(64 + 32 * NRD, 1) k3(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ u64 bar, stop_bar; __shared__ u32 slot; __shared__ unsigned long long n_loads;
  for (int i = threadIdx.x; i < (16384 + 4 * 16384) / 4; i += blockDim.x) ((u32*)smem)[i] = ((i * 2654435761u) & 0x807F807Fu) | 0x3F003F00u;
  if (threadIdx.x == 0) {
    n_loads = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&stop_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const u32 tm = slot;
  if (threadIdx.x == 0) {
    const u64 a = desc(smem_u32(smem));
    constexpr u32 ID = idesc(1, 128, 128);
    const long long t0 = clock64();
    if (MMA) {
      int st = 0;
      for (int i = 0; i < iters; ++i) {
        const u64 b = desc(smem_u32(smem + 16384 + st * 16384));
        st = (st + 1) & 3;
        const u32 d = tm + ((i >> 3) & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (MMA == 2) mma_ts(d, tm + 256 + (i & 7) * 32 + kk * 8, b + 2 * kk, ID, 1);
          else mma_ss(d, a + 2 * kk, b + 2 * kk, ID, 1);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tW4:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D4;\n\tbra W4;\n\tD4:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    } else {
      while (clock64() - t0 < (long long)iters * 4 * 64) {}       // as long as the MMAs would take at their floor
    }
    const long long t1 = clock64();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&stop_bar)) : "memory");
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x >= 64) {
    const int w = (threadIdx.x >> 5) - 2;                         // reader warp: TMEM lane quarter w % 4
    const u32 lane_base = tm + ((u32)(((threadIdx.x >> 5) & 3) * 32) << 16);   // a warp may only touch lanes 32 * (warp id % 4) ..
    u32 done = 0, acc = 0;
    unsigned long long mine = 0;
    int c = w >> 2;
    while (!done) {
      u32 v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(lane_base + (u32)(c & 7) * 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
      c += 2;
      ++mine;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&stop_bar)) : "memory");
    }
    if ((threadIdx.x & 31) == 0) atomicAdd(&n_loads, mine);
    if (acc == 0x12345678u) out[3] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (long long)n_loads;
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm)); }
}
template <int MMA, int NRD> void run3(const char* name) {
  long long* out; cudaMalloc(&out, 32); cudaMemset(out, 0, 32);
  const int smem = 16384 + 4 * 16384 + 1024, iters = 4000;
  cudaFuncSetAttribute(k3<MMA, NRD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) { k3<MMA, NRD><<<148, 64 + 32 * NRD, smem>>>(iters, out); cudaDeviceSynchronize(); }
  cudaError_t e = cudaGetLastError();
  long long c[2] = {0, 0}; cudaMemcpy(c, out, 16, cudaMemcpyDeviceToHost);
  printf("%-58s %7.1f clk / MMA, epilogue reads %6.1f B/clk  %s\n", name, (double)c[0] / (iters * 4.0), (double)c[1] * 4096.0 / (double)c[0],
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
template <bool TS, bool COMMIT, int STAGES, int SPIN, bool RANDOM, int BUSY = 0> void run2(const char* name) {
  long long* out; cudaMalloc(&out, 16);
  const int smem = 16384 + STAGES * 16384 + 1024, iters = 4000;
  cudaFuncSetAttribute(k2<TS, COMMIT, STAGES, SPIN, RANDOM, BUSY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) { k2<TS, COMMIT, STAGES, SPIN, RANDOM, BUSY><<<148, 64 + 32 * SPIN, smem>>>(iters, out); cudaDeviceSynchronize(); }
  cudaError_t e = cudaGetLastError();
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  printf("%-60s %7.1f clk / MMA  %s\n", name, (double)c / (iters * 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
template <int N, bool TS, int NACC> void run(const char* name, int grid) {
  long long* out; cudaMalloc(&out, 8);
  const int smem = 16384 + 32768 + 1024, iters = 2000;
  cudaFuncSetAttribute(k<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS, NACC><<<grid, 128, smem>>>(iters, out); cudaDeviceSynchronize();
  k<N, TS, NACC><<<grid, 128, smem>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  printf("%-28s grid %3d: %7.1f clk / MMA   (floor %d)  %s\n", name, grid, (double)c / (iters * 4.0), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
int main(int argc, char** argv) {
  printf("--- tensor-memory read bandwidth left for the epilogue (tcgen05.ld.32x32b.x32 + wait, back to back) ---\n");
  run3<0, 4>("no MMA, 4 reader warps");
  run3<0, 8>("no MMA, 8 reader warps");
  run3<1, 4>("SS MMAs (A from shared memory), 4 reader warps");
  run3<1, 8>("SS MMAs, 8 reader warps");
  run3<2, 4>("TS MMAs (A from tensor memory), 4 reader warps");
  run3<2, 8>("TS MMAs, 8 reader warps");
  if (argc > 1) return 0;
  for (int grid : {1, 148}) {
    run<64, false, 1>("SS N=64  1 acc", grid);  run<128, false, 1>("SS N=128 1 acc", grid); run<256, false, 1>("SS N=256 1 acc", grid);
    run<128, false, 2>("SS N=128 2 acc alternating", grid);
    run<64, true, 1>("TS N=64  1 acc", grid);   run<128, true, 1>("TS N=128 1 acc", grid);  run<256, true, 1>("TS N=256 1 acc", grid);
    run<128, true, 2>("TS N=128 2 acc alternating", grid);
  }
  printf("--- N=128, grid 148, knobs ---\n");
  run2<false, false, 1, 0, false>("SS base (zeros, 1 B tile, no commit, no spinners)");
  run2<false, true, 1, 0, false>("SS + commit per 4 MMAs");
  run2<false, false, 12, 0, false>("SS + B cycling over 12 tiles");
  run2<false, false, 1, 8, false>("SS + 8 spinning warps");
  run2<false, false, 1, 0, true>("SS + random data");
  run2<false, true, 12, 8, true>("SS + all");
  run2<true, false, 1, 0, false>("TS base");
  run2<true, true, 1, 0, false>("TS + commit per 4 MMAs");
  run2<true, false, 12, 0, false>("TS + B cycling over 12 tiles");
  run2<true, false, 1, 8, false>("TS + 8 spinning warps");
  run2<true, false, 1, 0, true>("TS + random data");
  run2<true, true, 12, 8, true>("TS + all");
  run2<true, true, 12, 8, true, 1>("TS + all + 8 ALU-busy warps");
  run2<true, true, 12, 8, true, 2>("TS + all + 8 ALU-busy warps reading TMEM");
  run2<false, true, 12, 8, true, 1>("SS + all + 8 ALU-busy warps");
  return 0;
}
