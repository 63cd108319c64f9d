// umma.cuh -- thin inline-PTX layer for sm_100a: mbarrier, bulk (TMA) copies, TMEM, tcgen05.mma.
//
// Everything here is a one-to-one wrapper around a PTX instruction; the kernels in this directory
// compose them. Layout conventions used throughout (all operands bf16, accumulators fp32 in TMEM):
//
//   "tile block" = [rows][64 elements] with 128-byte rows, 16-byte chunks XOR-swizzled with (row & 7)
//                  (the canonical SWIZZLE_128B atom: 8 rows x 128 B = 1024 B, atoms stacked along rows).
//   The same physical block serves as
//     * a K-major operand  (MN = rows, K = the 64 elements)   -> SBO = 1024 B, K-step of 16 = +32 B
//     * an MN-major operand (MN = the 64 elements, K = rows)  -> LBO = block stride, SBO = 1024 B,
//                                                                K-step of 16 rows = +2048 B
//   which is what lets the activation stash written by the forward kernel feed the weight-gradient
//   GEMM without any transposition.
#pragma once
#include <cuda_bf16.h>
#include <stdio.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200inr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ explicit shared-space accesses
// (pointers derived from the aligned dynamic-shared base are generic for the compiler: it would emit LD.E / ST.E)
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a mis-programmed pipeline traps instead of hanging the GPU box.  The report is out of line so that the
// ~40 wait sites of a warp-specialised kernel do not drag printf argument set-up through the instruction cache.
static __device__ __noinline__ void mbar_timeout(const void* bar, uint32_t parity) {
  printf("b200inr: mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) mbar_timeout(bar, parity);
  }
}

// Same with a nanosleep back-off between polls: for helper threads that run AHEAD of the critical path (loaders,
// converters), so that their polling does not take issue slots and barrier-unit bandwidth from the warps that compute.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (++spins > (1u << 24)) {
      printf("b200inr: mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x, (void*)bar,
             parity);
      __trap();
    }
  }
}
// Same, with a suspend-time hint: the hardware parks the thread until the phase completes or ~`kNs` ns have passed, so a
// warp that waits takes (almost) no issue slots away from the warps that compute.  Bounded (~4 s), then traps.
template <uint32_t kNs = 20000>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kNs)
        : "memory");
    if (ok) return;
    if (++spins > (4000000000u / kNs)) mbar_timeout(bar, parity);
  }
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ bulk (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// The same copy delivered to the same shared-memory offset of every CTA in `cta_mask` of the cluster; each destination
// CTA's mbarrier (same offset) receives the complete_tx for the bytes written there.
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                            uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// Pull `bytes` (multiple of 16) of global memory into L2 ahead of the loads that will need them.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// ------------------------------------------------------------------ TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor (sm_100 format): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout [61,64) (2 = SWIZZLE_128B).
__host__ __device__ constexpr uint64_t smem_desc_hi_sw128(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint64_t hi_bits) {
  return hi_bits | uint64_t((smem_addr >> 4) & 0x3FFF);
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                       // D = fp32
         | (1u << 7)                     // A = bf16
         | (1u << 10)                    // B = bf16
         | (uint32_t(a_mn_major) << 15)  // A major-ness
         | (uint32_t(b_mn_major) << 16)  // B major-ness
         | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- warp-converged issue: the WHOLE warp executes these, one elected lane issues.  Operands that are warp-uniform
// then live in uniform registers; issuing from inside an `if (lane == 0)` region instead makes ptxas wrap every
// tcgen05.mma in an ELECT / R2UR.BROADCAST x4 loop (~100 cycles per instruction, which dominates N <= 64 MMAs).
__device__ __forceinline__ void umma_bf16_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (lane = M row, two bf16 per 32-bit column along K).
__device__ __forceinline__ void umma_bf16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Four TS-mode MMAs of one K block from ONE elected lane: A at a_tmem + 8 k columns, B descriptor + k * kDescStep (in
// 16-byte units), k = 0..3.  One election and one predicate for the batch instead of one per instruction (the per-MMA
// elect / vote / register-to-uniform moves made a 16-MMA chain step issue-bound: ~62 cycles per 32-cycle MMA).
template <int kDescStep>
__device__ __forceinline__ void umma_bf16_ts_w4(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 d1, d2, d3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
      "add.u64 d1, %2, %5;\n\tadd.u64 d2, %2, %6;\n\tadd.u64 d3, %2, %7;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], d1, %3, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], d2, %3, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], d3, %3, 1;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate_first), "n"(kDescStep), "n"(2 * kDescStep), "n"(3 * kDescStep)
      : "memory");
}
// Sixteen TS-mode MMAs (four K blocks of four K steps) from ONE elected lane: A at a_tmem + 8 j columns, B descriptor +
// (kb * kBlkStep + ks * kStep) 16-byte units, j = 4 kb + ks.  The first one overwrites the accumulator.
template <int kStep, int kBlkStep>
__device__ __forceinline__ void umma_bf16_ts_w16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b32 a;\n\t.reg .b64 d;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 0;\n\t"
      "add.u32 a, %1, 8;\n\tadd.u64 d, %2, %4;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 16;\n\tadd.u64 d, %2, %5;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 24;\n\tadd.u64 d, %2, %6;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 32;\n\tadd.u64 d, %2, %7;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 40;\n\tadd.u64 d, %2, %8;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 48;\n\tadd.u64 d, %2, %9;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 56;\n\tadd.u64 d, %2, %10;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 64;\n\tadd.u64 d, %2, %11;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 72;\n\tadd.u64 d, %2, %12;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 80;\n\tadd.u64 d, %2, %13;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 88;\n\tadd.u64 d, %2, %14;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 96;\n\tadd.u64 d, %2, %15;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 104;\n\tadd.u64 d, %2, %16;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 112;\n\tadd.u64 d, %2, %17;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "add.u32 a, %1, 120;\n\tadd.u64 d, %2, %18;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], d, %3, 1;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(0 * kBlkStep + 1 * kStep), "n"(0 * kBlkStep + 2 * kStep), "n"(0 * kBlkStep + 3 * kStep), "n"(1 * kBlkStep + 0 * kStep), "n"(1 * kBlkStep + 1 * kStep), "n"(1 * kBlkStep + 2 * kStep), "n"(1 * kBlkStep + 3 * kStep), "n"(2 * kBlkStep + 0 * kStep), "n"(2 * kBlkStep + 1 * kStep), "n"(2 * kBlkStep + 2 * kStep), "n"(2 * kBlkStep + 3 * kStep), "n"(3 * kBlkStep + 0 * kStep), "n"(3 * kBlkStep + 1 * kStep), "n"(3 * kBlkStep + 2 * kStep), "n"(3 * kBlkStep + 3 * kStep)
      : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// commit of a 1-CTA MMA stream that arrives on the mbarrier at the same offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mc_w(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// Arrive (release at cluster scope) on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// Plain remote arrive (default semantics: release at CTA scope).  This is the hand-off used after a
// fence.proxy.async by the warps that fill an operand tile read by the peer-issued tcgen05.mma: the writes are local
// shared-memory stores, complete before the fence returns, and the arrive is issued after it; a cluster-scope release
// would add a MEMBAR.ALL.GPU that waits for every outstanding GLOBAL store of the warp.
__device__ __forceinline__ void mbar_arrive_peer(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// The same without the release: for hand-backs of a buffer whose contents this thread has finished READING (its loads
// have returned); a cluster-scope release costs ~1-2 k cycles under load.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("b200inr: cluster mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x,
             (void*)bar, parity);
      __trap();
    }
  }
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by ONE thread of the leader.
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// Arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs have completed.
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}

// Warp-converged variants for the CTA pair (see umma_bf16_ss_w).
__device__ __forceinline__ void umma_bf16_ss_2cta_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(
          d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_w(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(
          smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}

// ------------------------------------------------------------------ swizzled tile-block addressing
// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][64 x bf16] block.
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// ------------------------------------------------------------------ packed fp32 pairs (sm_100: FADD2 / FFMA2)
// Two fp32 lanes per instruction: the fused epilogues are ISSUE-bound (a polynomial sine that moved work from the MUFU
// to the FMA pipe made the forward slower, not faster), so every pair of element-wise adds / fmas is one instruction.
__device__ __forceinline__ unsigned long long f32x2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long u32x2(uint32_t a, uint32_t b) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ void add_f32x2(float& a0, float& a1, float b0, float b1) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f32x2(a0, a1)), "l"(f32x2(b0, b1)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(d));
}
// (a0, a1) * (m, m) + (c, c)
__device__ __forceinline__ void fma_f32x2(float a0, float a1, float m, float c, float& d0, float& d1) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f32x2(a0, a1)), "l"(f32x2(m, m)), "l"(f32x2(c, c)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// MUFU.TANH (rel. error ~2^-11): activations that are rounded to bf16 anyway
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace b200inr
