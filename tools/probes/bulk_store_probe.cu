// Shared memory -> global bulk store (cp.async.bulk.global.shared::cta): cycles from issue until the source has been
// read (wait_group.read 0) and until the store has completed (wait_group 0), per copy size, on an otherwise idle SM.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_store_probe bulk_store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(uint8_t* dst, int bytes, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t_read = 0, t_done = 0;
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + size_t(blockIdx.x) * 65536),
                   "r"(smem_u32(smem)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      const long long t1 = clock64();
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      const long long t2 = clock64();
      t_read += t1 - t0;
      t_done += t2 - t0;
    }
    out[blockIdx.x * 2] = t_read / iters;
    out[blockIdx.x * 2 + 1] = t_done / iters;
  }
}
int main() {
  uint8_t* dst;
  long long* out;
  cudaMalloc(&dst, 148 * 65536);
  cudaMalloc(&out, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int grid : {1, 148})
    for (int bytes : {1024, 4096, 8192, 16384, 32768, 65536}) {
      k<<<grid, 128, 65536>>>(dst, bytes, 200, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
      long long h[2];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("grid %3d  store %6d B: source read after %5lld clk (%.1f B/clk), complete after %5lld clk\n", grid, bytes, h[0],
             double(bytes) / h[0], h[1]);
    }
  return 0;
}
