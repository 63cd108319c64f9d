// L2 -> shared memory ingest rate of cp.async.bulk (1-D, no tensor map) per SM, as a function of the copy size and
// the number of copies in flight: one thread per CTA keeps `slots` copies of `bytes` each outstanding from an
// L2-resident source (every CTA streams the same 1 MB region, like a weight stream).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_probe bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void k(const uint8_t* src, size_t region, int bytes, int slots, int copies, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[64];
  if (threadIdx.x == 0) {
    for (int i = 0; i < slots; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long t0 = clock64();
    size_t off = (size_t(blockIdx.x) * 4096) % region;
    for (int c = 0; c < copies + slots; ++c) {
      const int s = c % slots;
      if (c >= slots) mbar_wait(&bars[s], ((c / slots) - 1) & 1);
      if (c < copies) {
        mbar_expect(&bars[s], bytes);
        bulk_g2s(smem + size_t(s) * bytes, src + off, bytes, &bars[s]);
        off += bytes;
        if (off + bytes > region) off = 0;
      }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const size_t region = 1 << 20;
  uint8_t* src;
  long long* cyc;
  cudaMalloc(&src, region);
  cudaMemset(src, 1, region);
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int cfg[][2] = {{32768, 3}, {32768, 6}, {16384, 6}, {16384, 12}, {8192, 12}, {8192, 24}, {4096, 24}, {4096, 48},
                        {2048, 48}, {1024, 48}, {65536, 3}};
  for (int grid : {148, 16}) {
    for (auto& c : cfg) {
      const int bytes = c[0], slots = c[1];
      const int copies = (64 << 20) / bytes;
      k<<<grid, 32, size_t(bytes) * slots, 0>>>(src, region, bytes, slots, copies, cyc);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
      long long h[148];
      cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < grid; ++i) mean += double(h[i]);
      mean /= grid;
      printf("grid %3d  copy %6d B x %2d in flight (%3d KB): %6.1f B/clk/SM   (%.0f clk per copy issued)\n", grid, bytes, slots,
             bytes * slots / 1024, double(copies) * bytes / mean, mean / copies);
    }
  }
  return 0;
}
