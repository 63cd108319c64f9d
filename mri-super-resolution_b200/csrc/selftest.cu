// selftest.cu -- single-CTA tcgen05 GEMM used to pin the shared-memory / instruction descriptor conventions
// of umma.cuh on real hardware (there is no way to check them without a B200).
//   mode 0: K-major operands  a[128,K], b[N,K]  -> D = a * b^T         (forward / dgrad operand form)
//   mode 1: MN-major operands a[K,128], b[K,N]  -> D = a^T * b         (weight-gradient operand form)
#include <stdio.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

struct SelftestParams {
  int mode;
  const __nv_bfloat16* a;
  const __nv_bfloat16* b;
  float* d;
  int N, K;
  int lbo_a, sbo_a, lbo_b, sbo_b;
};

__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(const SelftestParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int N = p.N, K = p.K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* a_smem = smem;
  uint32_t a_bytes, a_blk, b_blk;
  if (p.mode == 0) {
    a_blk = 128 * 128;  // [128 rows][64 k]
    a_bytes = (K / 64) * a_blk;
    b_blk = N * 128;  // [N rows][64 k]
  } else {
    a_blk = K * 128;  // [K rows][64 m]
    a_bytes = 2 * a_blk;
    b_blk = K * 128;  // [K rows][64 n]
  }
  uint8_t* b_smem = smem + a_bytes;

  // generic-proxy fill of the swizzled blocks
  if (p.mode == 0) {
    for (int i = tid; i < 128 * K / 8; i += 128) {  // one 16-byte chunk per iteration
      const int row = i / (K / 8), c8 = i % (K / 8);
      const uint4 v = *reinterpret_cast<const uint4*>(p.a + size_t(row) * K + c8 * 8);
      *reinterpret_cast<uint4*>(a_smem + (c8 / 8) * a_blk + sw128_chunk_off(row, c8 % 8)) = v;
    }
    for (int i = tid; i < N * K / 8; i += 128) {
      const int row = i / (K / 8), c8 = i % (K / 8);
      const uint4 v = *reinterpret_cast<const uint4*>(p.b + size_t(row) * K + c8 * 8);
      *reinterpret_cast<uint4*>(b_smem + (c8 / 8) * b_blk + sw128_chunk_off(row, c8 % 8)) = v;
    }
  } else {
    for (int i = tid; i < K * 128 / 8; i += 128) {
      const int row = i / 16, c8 = i % 16;  // row = k, c8 = 8-wide group along m
      const uint4 v = *reinterpret_cast<const uint4*>(p.a + size_t(row) * 128 + c8 * 8);
      *reinterpret_cast<uint4*>(a_smem + (c8 / 8) * a_blk + sw128_chunk_off(row, c8 % 8)) = v;
    }
    for (int i = tid; i < K * N / 8; i += 128) {
      const int row = i / (N / 8), c8 = i % (N / 8);
      const uint4 v = *reinterpret_cast<const uint4*>(p.b + size_t(row) * N + c8 * 8);
      *reinterpret_cast<uint4*>(b_smem + (c8 / 8) * b_blk + sw128_chunk_off(row, c8 % 8)) = v;
    }
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (tid == 0) {
    const uint32_t a_base = smem_u32(a_smem), b_base = smem_u32(b_smem);
    if (p.mode == 0) {
      const uint64_t hia = smem_desc_hi_sw128(p.lbo_a >= 0 ? p.lbo_a : 0, p.sbo_a >= 0 ? p.sbo_a : 1024);
      const uint64_t hib = smem_desc_hi_sw128(p.lbo_b >= 0 ? p.lbo_b : 0, p.sbo_b >= 0 ? p.sbo_b : 1024);
      const uint32_t idesc = idesc_bf16(128, N, false, false);
      for (int kb = 0; kb < K / 64; ++kb)
        for (int k4 = 0; k4 < 4; ++k4)
          umma_bf16_ss(tmem_d, smem_desc(a_base + kb * a_blk + k4 * 32, hia),
                       smem_desc(b_base + kb * b_blk + k4 * 32, hib), idesc, (kb | k4) != 0);
    } else {
      const uint64_t hia = smem_desc_hi_sw128(p.lbo_a >= 0 ? p.lbo_a : a_blk, p.sbo_a >= 0 ? p.sbo_a : 1024);
      const uint64_t hib = smem_desc_hi_sw128(p.lbo_b >= 0 ? p.lbo_b : b_blk, p.sbo_b >= 0 ? p.sbo_b : 1024);
      const uint32_t idesc = idesc_bf16(128, N, true, true);
      for (int ks = 0; ks < K / 16; ++ks)
        umma_bf16_ss(tmem_d, smem_desc(a_base + ks * 2048, hia), smem_desc(b_base + ks * 2048, hib), idesc, ks != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_d + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d[size_t(row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_d);
}

int launch_selftest_umma(int mode, const void* a, const void* b, float* d, int N, int K, int lbo_a, int sbo_a,
                         int lbo_b, int sbo_b, cudaStream_t stream) {
  if (N % 16 != 0 || N < 16 || N > 256) return B200INR_ERR_BAD_SHAPE;
  if (mode == 0 && (K % 64 != 0 || K < 64 || K > 256)) return B200INR_ERR_BAD_SHAPE;
  if (mode == 1 && (K % 16 != 0 || K < 16 || K > 128 || N % 64 != 0)) return B200INR_ERR_BAD_SHAPE;
  SelftestParams p{mode, reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), d, N, K,
                   lbo_a,  sbo_a, lbo_b, sbo_b};
  const int smem = 200 * 1024;
  if (cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  selftest_umma_kernel<<<1, 128, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
