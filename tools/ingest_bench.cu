// Microbenchmark: how fast can one SM (and the chip) ingest global memory into shared memory with cp.async.bulk rings?
// Steers the streamed decode kernel (ring depth, slot size, CTAs used).  nvcc -arch=sm_100a -O3 -o ingest_bench ingest_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(160, 1) ingest(const char* src, size_t region, size_t per_cta, int slot_bytes, int nslots, int shared_src,
                                                unsigned long long* sink, int nprod) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm;
  uint64_t* empty = full + 32;
  unsigned char* ring = sm + 512;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + i)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(empty + i)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t n = per_cta / slot_bytes;
  const size_t base = shared_src ? 0 : ((size_t)blockIdx.x * per_cta) % region;
  if (threadIdx.x >= 32 && (threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) - 1 < nprod) {
    for (size_t i = (threadIdx.x >> 5) - 1; i < n; i += nprod) {
      const int idx = i % nslots; const uint32_t par = ((i / nslots) & 1) ^ 1;
      uint32_t ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(empty + idx)), "r"(par) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + idx)), "r"(slot_bytes) : "memory");
      const char* g = src + (base + i * (size_t)slot_bytes) % region;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)idx * slot_bytes)), "l"(g), "r"(slot_bytes), "r"(s32(full + idx)) : "memory");
    }
  } else if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (size_t i = 0; i < n; ++i) {
      const int idx = i % nslots; const uint32_t par = (i / nslots) & 1;
      uint32_t ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(full + idx)), "r"(par) : "memory");
      acc += *(volatile unsigned long long*)(ring + (size_t)idx * slot_bytes);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + idx)) : "memory");
    }
    if (acc == 0x1234567) *sink = acc;
  }
}
int main() {
  const size_t region_big = (size_t)4 << 30, region_l2 = (size_t)32 << 20;
  char* buf; unsigned long long* sink;
  cudaMalloc(&buf, region_big); cudaMemset(buf, 1, region_big); cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ctas[] = {1, 104, 148};
  const int cfgs[][3] = {{16384, 8, 1}, {16384, 8, 2}, {16384, 8, 4}, {32768, 6, 1}, {32768, 6, 2}, {65536, 3, 1}, {4096, 32, 4}};
  for (int src_kind = 0; src_kind < 2; ++src_kind) {         // 0: HBM stream (distinct per CTA), 1: L2-resident 32 MB distinct offsets, 2: all CTAs read the same 32 MB
    for (auto& cf : cfgs) {
      for (int nc : ctas) {
        const size_t per_cta = src_kind == 0 ? ((size_t)16 << 20) : ((size_t)32 << 20);
        const size_t region = src_kind == 0 ? region_big : region_l2;
        const size_t smem = 512 + (size_t)cf[0] * cf[1];
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          ingest<<<nc, 160, smem>>>(buf, region, per_cta, cf[0], cf[1], src_kind == 2, sink, cf[2]);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          if (rep == 1) {
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double gbs = (double)per_cta * nc / (ms * 1e-3) / 1e9;
            printf("src=%s nprod=%d slot=%5d x%2d ctas=%3d : %8.1f GB/s total, %6.1f GB/s per SM\n", src_kind == 0 ? "hbm " : (src_kind == 1 ? "l2  " : "l2sh"), cf[2], cf[0], cf[1], nc, gbs, gbs / nc);
            fflush(stdout);
          }
        }
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
