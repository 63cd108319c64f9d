// mlp_bwdp.cu -- SIREN backward as ONE layer-pipelined kernel for sm_100a: dgrad chain, weight and bias gradients.
//
// Replaces loss.backward() through nn.Sequential(SineLayer x (L+1), nn.Linear) (autograd of the reference's
// INR/SRDWI.py:58-59,87-91; math in SURVEY.md App. B.1):
//     dTheta_L   = (dOut W_f) .* cos(theta_L)                 dW_f = dOut^T Y_L               db_f = colsum(dOut)
//     dTheta_l-1 = (dTheta_l W'_l) .* cos(theta_l-1)          dW_l = w_l dTheta_l^T Y_l-1     db_l = w_l colsum(dTheta_l)
//     dW_0 = w_0 dTheta_0^T X
// The staged path (mlp_bwd.cu + wgrad.cu) writes every dTheta_l to HBM and reads it back together with the stashed
// sin outputs: 11 GB per cfg2 step, HBM bound.  Here the only HBM stream is the 16-bit phase stash of the forward
// (2.8 GB): sin AND cos are recomputed from the phase, dTheta tiles travel from layer to layer through small ring
// buffers that live in L2, and the weight gradients never leave tensor memory until the end of the kernel.
//
// Work split: the SMs form P = floor(#SM / 2(L+1)) independent pipelines of 2(L+1) CTAs; pipeline p walks its share of
// the 64-row tiles.  Every CTA is WEIGHT-STATIONARY for its whole life:
//   stage CTA (l, h), l = L..1, h = 0/1 (feature half):                                   [the 256x256 layers]
//       holds rows [128h, 128h+128) of W'_l^T (64 KB, A operand) and the [256 x 128] block dW_l[:, 128h..] in TMEM;
//       per tile:  receive dTheta_l (64 x 256 bf16, 32 KB)                                <- ring l
//                  D^T[in-half, rows]   = W'_l^T[in-half, :] dTheta_l^T                   (tcgen05 128 x 64 x 256)
//                  y = sin(phase_l-1), c = cos(phase_l-1)      for its 128 features       (phases by bulk copy from HBM)
//                  dTheta_l-1[:, half]  = D .* c  -> bf16                                  -> ring l-1
//                  dW_l[:, half]       += dTheta_l^T y                                     (tcgen05 2 x 128 x 128 x 64)
//                  db_l-1[half]        += colsum(dTheta_l-1)                               (registers: lane = feature)
//   edge CTA E(h):                                                                        [both ends of the chain]
//       top:     dOut tile -> bf16, D^T = W_f^T[half] dOut^T, y = sin(phase_L), dTheta_L = D .* cos(phase_L) -> ring L,
//                dW_f^T[half] += y^T dOut, db_f, db_L
//       bottom:  receive dTheta_0[:, half]                                                <- ring 0
//                dW_0[half]   += dTheta_0^T [x_hi | x_lo]    (coordinates from the voxel index, bf16 hi + lo split)
// The chain is computed TRANSPOSED (features on the 128 TMEM lanes, tile rows on the columns), so a tile may have any
// row count (64 here: operand slots + weights fit the 227 KB of shared memory) at full tensor-core rate and the bias
// gradient is a per-thread running sum.  A dTheta tile is stored FEATURE-major ([64 features][64 rows] swizzled
// blocks: each thread writes its feature's 16 rows with two 16-byte stores); read with MN-major descriptors it is the
// B operand of the chain step (N = rows, K = features), with K-major descriptors the A operand of the
// weight-gradient step (M = features, K = rows).
//
// Rings: per pipeline and layer boundary kRing slots of one tile; producers bulk-store their half and release a
// counter, consumers acquire, bulk-load and return a credit counter (all bounded spins: a mis-programmed pipeline
// traps instead of hanging).  All 2(L+1)P CTAs must be co-resident: the grid never exceeds the SM count (1 CTA/SM).
//
// Warp roles (640 threads): 0 = ring/weight loader (edge: phase loader), 1 = MMA issuer + TMEM owner, 2 = ring store,
// 3 = phase loader (edge: the whole bottom half), 4..19 = epilogue (TMEM lane quadrant = warp & 3).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

constexpr int kPThreads = 640;
constexpr int kPEpiWarps = 16;
constexpr int kPFirstEpiWarp = 4;
constexpr int kPEpiThreads = kPEpiWarps * 32;  // 512
constexpr int kPDzSlots = 3;                    // incoming dTheta tiles
constexpr int kPPhSlots = 2;                    // phase tiles
constexpr int kPDobSlots = 3;                   // edge: bf16 dOut blocks (converted two tiles ahead)
constexpr int kPRawSlots = 3;                   // edge: raw fp32 dOut tiles (bulk-copied ahead of the conversion)
constexpr int kPZSlots = 3;                     // edge bottom: dTheta_0 slots (2 tiles of look-ahead)
constexpr int kPStoreDepth = 2;                 // bulk stores in flight per ring-store thread
constexpr int kPBlk = kPipeTileRows * 128;      // one [64][64] bf16 block: 8 KB
constexpr int kPHalf = 2 * kPBlk;               // 128 features of a tile: 16 KB
constexpr int kPTile = 4 * kPBlk;               // 256 features of a tile: 32 KB
constexpr int kPPhChunk = kPipeTileRows * 16 + 16;  // phases of 8 features x 64 rows, padded (bank spread)
constexpr int kPPhSlot = 16 * kPPhChunk;            // 16 chunks = 128 features
constexpr int kPRawSlot = kPipeTileRows * kOutPad * 4;  // fp32 dOut tile, C <= 32

// shared memory map (bytes); operand blocks are 1024-byte aligned
struct PSmem {
  // stage CTA
  static constexpr int kWt = 0;                              // W'^T half: 4 x [128][64]      64 KB
  static constexpr int kDz = kWt + 4 * 128 * 128;            // kPDzSlots x tile               96 KB
  static constexpr int kY = kDz + kPDzSlots * kPTile;        // sin outputs, 2 blocks          16 KB
  static constexpr int kStg = kY + kPHalf;                   // outgoing dTheta half           16 KB
  static constexpr int kPh = kStg + kPHalf;                  // kPPhSlots phase slots          32.5 KB
  static constexpr int kBar = kPh + kPPhSlots * kPPhSlot;
  static constexpr int kBytes = kBar + 512;
  // edge CTA (re-uses kY, kStg, kPh, kBar)
  static constexpr int kWf = 0;                              // W_f^T half [128][64]           16 KB
  static constexpr int kDob = kWf + 128 * 128;               // kPDobSlots x dOut block        24 KB
  static constexpr int kZ = kDob + kPDobSlots * kPBlk;       // kPZSlots x dTheta_0 half       48 KB
  static constexpr int kXb = kZ + kPZSlots * kPHalf;         // kPZSlots x coordinate block    24 KB
  static constexpr int kRaw = kXb + kPZSlots * kPBlk;        // kPRawSlots x fp32 dOut tile    24 KB
  static_assert(kRaw + kPRawSlots * kPRawSlot <= kY, "edge layout overlaps");
  static_assert(kBytes + 1024 <= 232448, "shared memory budget");
};

// barrier indices
enum PBar : int {
  kBW = 0,          // static weights landed
  kBDzFull = 1,     // [3]
  kBDzEmpty = 4,    // [3]
  kBPhFull = 7,     // [2]
  kBPhEmpty = 9,    // [2]
  kBAccFull = 11,   // [2]
  kBAccEmpty = 13,  // [2]
  kBYFull = 15,
  kBYEmpty = 16,
  kBStgFull = 17,
  kBStgEmpty = 18,
  kBFin = 19,
  kBDobFull = 20,   // [3] edge
  kBDobEmpty = 23,  // [3] edge
  kBZFull = 26,     // [3] edge bottom
  kBZEmpty = 29,    // [3] edge bottom
  kBFinB = 32,
  kBRawFull = 33,   // [3] edge
  kBRawEmpty = 36,  // [3] edge
  kBCount = 39
};

struct PipeParams {
  const uint8_t* packed;
  PackLayout pl;
  const float* grad_out;  // [rows, C]
  long long rows;
  int fwd_tiles;          // 128-row tiles of the forward stash
  int L, C, d;
  const uint8_t* ph;      // phase stash: (L+1) x fwd_tiles x [32 chunks][128 rows][8] u16
  size_t layer_stride;
  const uint4* xa;        // coordinate stash: per row bf16 {hi x4, lo x4} (x = hi + lo), written by the forward
  uint8_t* ring;          // [pipeline][edge 0..L][kPipeRing][32 KB]
  uint32_t* flags;        // [pipeline][edge][4] counters, 128 bytes apart: produced by half 0/1, consumed by half 0/1
  float* grads;
  long long off[2 * (kMaxSineLayers + 2)];
  float omega0, omegah;
  int pipelines;
  unsigned long long* prof;  // nullptr, or [grid][kPipeProfSlots] stall-cycle counters (B200INR_BWDP_PROF=1)
};

// Stall accounting for pipeline tuning: PW(k, stmt) runs stmt and, when profiling, adds its duration to counter k of
// the calling thread (each recording thread owns a private range of the CTA's kPipeProfSlots counters).
#define PW(k, stmt)                                   \
  do {                                                \
    if (prof_on) {                                    \
      const long long t0_ = clock64();                \
      stmt;                                           \
      pw[k] += (unsigned long long)(clock64() - t0_); \
    } else {                                          \
      stmt;                                           \
    }                                                 \
  } while (0)
#define PW_FLUSH(base, cnt)                                                                                  \
  do {                                                                                                       \
    if (prof_on)                                                                                             \
      for (int k_ = 0; k_ < (cnt); ++k_) p.prof[size_t(blockIdx.x) * kPipeProfSlots + (base) + k_] = pw[k_]; \
  } while (0)

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// counter[0..3] += src[0..3] through the bulk-copy (async proxy) engine: issued after the bulk stores it publishes
// have completed, so the producer needs no generic-proxy fence (a gpu-scope release costs ~1-2.5 k cycles here).
__device__ __forceinline__ void bulk_red_add_u32x4(uint32_t* gmem_dst, const void* smem_src) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], 16;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src))
               : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}

// Wait until both counters reach `target`; `k0`, `k1` cache the last values seen (the counters only grow).
// Bounded: a mis-programmed pipeline traps instead of hanging.
__device__ __forceinline__ void poll2_ge(const uint32_t* f0, const uint32_t* f1, uint32_t target, uint32_t& k0,
                                         uint32_t& k1) {
  uint32_t spins = 0;
  while (k0 < target || k1 < target) {
    const uint32_t a = ld_acquire_gpu(f0), b = ld_acquire_gpu(f1);
    k0 = a;
    k1 = b;
    if (k0 >= target && k1 >= target) break;
    __nanosleep(32);
    if (++spins > (1u << 23)) {
      printf("b200inr: pipeline flag timeout block %d thread %d flags %p %p target %u (%u, %u)\n", blockIdx.x,
             threadIdx.x, (const void*)f0, (const void*)f1, target, a, b);
      __trap();
    }
  }
}

__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.b16 %0, [%1];\n" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// phase (u16, 65536 = one turn) -> radians: float(2^23 + ph) is built with one PRMT / LOP, then one FFMA
constexpr float kPhToRad = 9.587379924285257e-05f;   // 2*pi / 65536
constexpr float kPhBias = -804.247719318987f;        // -(2^23) * 2*pi / 65536
__device__ __forceinline__ float rad_lo16(uint32_t w) {
  return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), kPhToRad, kPhBias);
}
__device__ __forceinline__ float rad_hi16(uint32_t w) {
  return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)), kPhToRad, kPhBias);
}

// ---- tile epilogue, shared by stage and edge CTAs ----------------------------------------------------------------
// TMEM lane = feature f, columns = rows [16 cg, 16 cg + 16).  From ONE phase read per element:
//   y^T[f][row]      = sin(phase)                 (operand of the weight-gradient MMA)
//   dTheta^T[f][row] = D^T[f][row] * cos(phase)   (sent on to the next layer)
// both packed feature-major (16 rows = two 16-byte chunks of the feature's 128-byte row).  Returns the fp32 sum of
// dTheta over the 16 rows (bias gradient partial).
__device__ __forceinline__ float sincos_tile(uint32_t acc_t, uint32_t ph_s, int f, int cg, uint32_t (&ys)[8],
                                             uint32_t (&ds)[8]) {
  uint32_t v[16];
  tmem_ld16(acc_t + cg * 16, v);
  const uint32_t ph_f = ph_s + (f >> 3) * kPPhChunk + (f & 7) * 2 + cg * 16 * 16;
  uint32_t ph[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) ph[j] = lds16(ph_f + j * 16);
  tmem_ld_wait();
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float r0 = rad_lo16(ph[2 * j]), r1 = rad_lo16(ph[2 * j + 1]);
    const float d0 = __uint_as_float(v[2 * j]) * __cosf(r0);
    const float d1 = __uint_as_float(v[2 * j + 1]) * __cosf(r1);
    sum += d0 + d1;
    ds[j] = pack_bf16x2(d0, d1);
    ys[j] = pack_bf16x2(__sinf(r0), __sinf(r1));
  }
  return sum;
}
__device__ __forceinline__ void store_feature_rows(uint32_t half_s, int f, int cg, const uint32_t (&o)[8]) {
  const uint32_t blk = half_s + (f >> 6) * kPBlk;
  sts128(blk + sw128_chunk_off(f & 63, 2 * cg), make_uint4(o[0], o[1], o[2], o[3]));
  sts128(blk + sw128_chunk_off(f & 63, 2 * cg + 1), make_uint4(o[4], o[5], o[6], o[7]));
}

__global__ void __launch_bounds__(kPThreads, 1) siren_bwdp_kernel(const PipeParams p) {
  using S = PSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBCount);
  uint32_t* one_s = reinterpret_cast<uint32_t*>(smem + S::kBar + 320);  // 16-byte aligned constant {1, 0, 0, 0}
  static_assert(kBCount * 8 + 4 <= 320 && 320 + 16 <= 512, "barrier area layout");
  const uint32_t sbase = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int L = p.L;
  const int S2 = 2 * (L + 1);
  const int pipe = int(blockIdx.x) / S2;
  const int role = int(blockIdx.x) % S2;
  const int stage = role >> 1;  // 0 = edge, s >= 1 = layer L + 1 - s
  const int h = role & 1;
  const bool edge = stage == 0;
  const int layer = L + 1 - stage;            // stage CTA: the layer whose weights it holds
  const int in_edge = edge ? 0 : layer;       // ring this CTA consumes
  const int out_edge = edge ? L : layer - 1;  // ring this CTA produces
  const int ph_layer = edge ? L : layer - 1;  // phases needed by the epilogue

  // tiles of this pipeline: forward tiles T = pipe, pipe + P, ...; two 64-row tiles each
  const int my_fwd = pipe < p.fwd_tiles ? (p.fwd_tiles - pipe + p.pipelines - 1) / p.pipelines : 0;
  const int n = 2 * my_fwd;

  uint8_t* ring_in = p.ring + size_t(pipe * (L + 1) + in_edge) * kPipeRing * kPTile;
  uint8_t* ring_out = p.ring + size_t(pipe * (L + 1) + out_edge) * kPipeRing * kPTile;
  uint32_t* fl_in = p.flags + size_t(pipe * (L + 1) + in_edge) * 4 * 32;
  uint32_t* fl_out = p.flags + size_t(pipe * (L + 1) + out_edge) * 4 * 32;

  if (threadIdx.x == 0) {
    one_s[0] = 1u;  // {1, 0, 0, 0}: source of the bulk add that publishes a ring slot
    one_s[1] = one_s[2] = one_s[3] = 0u;
    mbar_init(&bars[kBW], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kBPhFull + i], 1);
      mbar_init(&bars[kBPhEmpty + i], kPEpiWarps);
      mbar_init(&bars[kBAccFull + i], 1);
      mbar_init(&bars[kBAccEmpty + i], kPEpiWarps);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars[kBDzFull + i], 1);
      mbar_init(&bars[kBDzEmpty + i], 1);
      mbar_init(&bars[kBDobFull + i], kPEpiWarps);
      mbar_init(&bars[kBDobEmpty + i], 1);
      mbar_init(&bars[kBZFull + i], 1);
      mbar_init(&bars[kBZEmpty + i], 1);
      mbar_init(&bars[kBRawFull + i], 1);
      mbar_init(&bars[kBRawEmpty + i], kPEpiWarps);
    }
    mbar_init(&bars[kBYFull], kPEpiWarps);
    mbar_init(&bars[kBYEmpty], 1);
    mbar_init(&bars[kBStgFull], kPEpiWarps);
    mbar_init(&bars[kBStgEmpty], 1);
    mbar_init(&bars[kBFin], 1);
    mbar_init(&bars[kBFinB], 1);
    fence_mbar_init();
    fence_proxy_async_smem();  // one_s is read by the bulk-copy engine
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns.  stage: dW block [0,256) (two M halves x 128), chain accumulators 256 + 64 j.
  //               edge : chain accumulators 0 + 64 j, dW_f^T [128,192), dW_0 [192,256).
  const uint32_t t_acc = edge ? tmem : tmem + 256;
  const uint32_t t_w = edge ? tmem + 128 : tmem;
  const uint32_t t_w0 = tmem + 192;

  const uint64_t hiK = smem_desc_hi_sw128(0, 1024);       // K-major blocks
  const uint64_t hiMN = smem_desc_hi_sw128(kPBlk, 1024);  // MN-major: 64-wide MN blocks kPBlk apart

  const bool prof_on = p.prof != nullptr;
  const long long t_begin = prof_on ? clock64() : 0;

  if (n > 0) {
    if (warp == 0) {
      // =============================== loader ===============================
      if (lane == 0) {
        if (!edge) {
          unsigned long long pw[2] = {0, 0};
          const uint8_t* src = p.packed + p.pl.wht + size_t(layer - 1) * 256 * 256 * 2 + size_t(h) * 128 * 128;
          mbar_arrive_expect_tx(&bars[kBW], 4 * 128 * 128);
          for (int kb = 0; kb < 4; ++kb)
            bulk_g2s(smem + S::kWt + kb * 128 * 128, src + size_t(kb) * 256 * 128, 128 * 128, &bars[kBW]);
          uint32_t k0 = 0, k1 = 0;
          for (int i = 0; i < n; ++i) {
            const int slot = i % kPDzSlots, round = i / kPDzSlots;
            if (round > 0) PW(0, mbar_wait(&bars[kBDzEmpty + slot], (round - 1) & 1));
            PW(1, poll2_ge(fl_in + 0 * 32, fl_in + 1 * 32, uint32_t(i + 1), k0, k1));
            // (no proxy fence: the tile was written AND published through the async proxy, and is read through it)
            mbar_arrive_expect_tx(&bars[kBDzFull + slot], kPTile);
            bulk_g2s(smem + S::kDz + slot * kPTile, ring_in + size_t(i % kPipeRing) * kPTile, kPTile,
                     &bars[kBDzFull + slot]);
          }
          PW_FLUSH(1, 2);
        } else {
          mbar_arrive_expect_tx(&bars[kBW], 128 * 128);
          bulk_g2s(smem + S::kWf, p.packed + p.pl.wft + size_t(h) * 128 * 128, 128 * 128, &bars[kBW]);
        }
      }
    }
    if ((edge && warp == 0) || (!edge && warp == 3)) {
      // =============================== phase loader (edge: + raw dOut tiles) ===============================
      if (lane == 0) {
        unsigned long long pw[1] = {0};
        int next_raw = 0;
        for (int i = 0; i < n; ++i) {
          if (edge) {  // fp32 dOut tiles run two tiles ahead of the phases (they are converted two tiles ahead)
            const int lim = (i + 3 < n) ? i + 3 : n;
            for (; next_raw < lim; ++next_raw) {
              const int j = next_raw, rs = j % kPRawSlots;
              if (j >= kPRawSlots) mbar_wait(&bars[kBRawEmpty + rs], ((j / kPRawSlots) - 1) & 1);
              const int Tj = pipe + (j >> 1) * p.pipelines;
              const long long row0 = (long long)Tj * 128 + (j & 1) * kPipeTileRows;
              if (row0 + kPipeTileRows <= p.rows) {
                const uint32_t bytes = uint32_t(kPipeTileRows) * p.C * 4;
                mbar_arrive_expect_tx(&bars[kBRawFull + rs], bytes);
                bulk_g2s(smem + S::kRaw + rs * kPRawSlot, p.grad_out + row0 * p.C, bytes, &bars[kBRawFull + rs]);
              } else {
                mbar_arrive(&bars[kBRawFull + rs]);  // ragged last tile: the epilogue reads it from global memory
              }
            }
          }
          const int slot = i % kPPhSlots, round = i / kPPhSlots;
          if (round > 0) PW(0, mbar_wait(&bars[kBPhEmpty + slot], (round - 1) & 1));
          const int T = pipe + (i >> 1) * p.pipelines;
          const uint8_t* src = p.ph + size_t(ph_layer) * p.layer_stride + size_t(T) * (128 * 256 * 2) +
                               size_t(i & 1) * (kPipeTileRows * 16) + size_t(h) * 16 * (128 * 16);
          mbar_arrive_expect_tx(&bars[kBPhFull + slot], 16 * kPipeTileRows * 16);
          for (int c = 0; c < 16; ++c)
            bulk_g2s(smem + S::kPh + slot * kPPhSlot + c * kPPhChunk, src + size_t(c) * (128 * 16), kPipeTileRows * 16,
                     &bars[kBPhFull + slot]);
        }
        PW_FLUSH(3, 1);
      }
    } else if (warp == 1) {
      // =============================== MMA issuer ===============================
      // The chain MMA of tile i + 1 is issued BEFORE waiting for the epilogue of tile i, so its accumulator is ready
      // when the epilogue warps get there; the weight-gradient MMA of tile i follows the epilogue of tile i.
      if (lane == 0) {
        unsigned long long pw[4] = {0, 0, 0, 0};
        mbar_wait(&bars[kBW], 0);
        tc_fence_after();
        if (!edge) {
          const uint32_t idesc_d = idesc_bf16(128, kPipeTileRows, false, true);  // A = W'^T (K-major), B = dTheta (MN-major)
          const uint32_t idesc_w = idesc_bf16(128, 128, false, false);           // A = dTheta, B = y: both K-major
          auto chain = [&](int i) {  // D^T[in-half][rows] = sum over the 256 outputs: 4 feature blocks x 4 K steps
            const int ds = i % kPDzSlots, as = i & 1;
            PW(0, mbar_wait(&bars[kBDzFull + ds], (i / kPDzSlots) & 1));
            PW(3, st_relaxed_gpu(fl_in + (2 + h) * 32, uint32_t(i + 1)));  // credit: the ring slot has been read out
            if (i >= 2) PW(1, mbar_wait(&bars[kBAccEmpty + as], ((i >> 1) - 1) & 1));
            tc_fence_after();
            const uint32_t dz = sbase + S::kDz + ds * kPTile;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16_ss(t_acc + as * 64, smem_desc(sbase + S::kWt + kb * 128 * 128 + ks * 32, hiK),
                             smem_desc(dz + kb * kPBlk + ks * 2048, hiMN), idesc_d, (kb | ks) != 0);
            umma_commit(&bars[kBAccFull + as]);
          };
          chain(0);
          for (int i = 0; i < n; ++i) {
            if (i + 1 < n) chain(i + 1);
            const uint32_t dz = sbase + S::kDz + (i % kPDzSlots) * kPTile;
            PW(2, mbar_wait(&bars[kBYFull], i & 1));
            tc_fence_after();
            // dW[out][in-half] += sum over the 64 rows: M halves of 128 outputs, 4 K steps of 16 rows
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)
#pragma unroll
              for (int ks = 0; ks < kPipeTileRows / 16; ++ks)
                umma_bf16_ss(t_w + mh * 128, smem_desc(dz + 2 * mh * kPBlk + ks * 32, hiK),
                             smem_desc(sbase + S::kY + ks * 32, hiK), idesc_w, (i | ks) != 0);
            umma_commit(&bars[kBYEmpty]);
            umma_commit(&bars[kBDzEmpty + i % kPDzSlots]);
          }
        } else {
          const uint32_t idesc_d = idesc_bf16(128, kPipeTileRows, false, false);  // A = W_f^T, B = dOut: both K-major
          const uint32_t idesc_w = idesc_bf16(128, 64, false, true);              // A = y (K-major), B = dOut (MN-major)
          auto chain = [&](int i) {
            const int as = i & 1, bs = i % kPDobSlots;
            PW(0, mbar_wait(&bars[kBDobFull + bs], (i / kPDobSlots) & 1));
            if (i >= 2) PW(1, mbar_wait(&bars[kBAccEmpty + as], ((i >> 1) - 1) & 1));
            tc_fence_after();
            const uint32_t dob = sbase + S::kDob + bs * kPBlk;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(t_acc + as * 64, smem_desc(sbase + S::kWf + k4 * 32, hiK), smem_desc(dob + k4 * 32, hiK),
                           idesc_d, k4 != 0);
            umma_commit(&bars[kBAccFull + as]);
          };
          chain(0);
          for (int i = 0; i < n; ++i) {
            if (i + 1 < n) chain(i + 1);
            const int bs = i % kPDobSlots;
            const uint32_t dob = sbase + S::kDob + bs * kPBlk;
            PW(2, mbar_wait(&bars[kBYFull], i & 1));
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kPipeTileRows / 16; ++ks)
              umma_bf16_ss(t_w, smem_desc(sbase + S::kY + ks * 32, hiK), smem_desc(dob + ks * 2048, hiMN), idesc_w,
                           (i | ks) != 0);
            umma_commit(&bars[kBYEmpty]);
            umma_commit(&bars[kBDobEmpty + bs]);
          }
        }
        umma_commit(&bars[kBFin]);
        PW_FLUSH(4, 3);
        if (prof_on) p.prof[size_t(blockIdx.x) * kPipeProfSlots + 25] = pw[3];
      }
    } else if (warp == 2) {
      // =============================== ring store ===============================
      // Up to kPStoreDepth bulk stores stay in flight: the counter that publishes tile i is bumped (by a bulk add) once
      // the group of tile i has completed, kPStoreDepth iterations later (or at the end).
      if (lane == 0) {
        unsigned long long pw[4] = {0, 0, 0, 0};
        uint32_t c0 = 0, c1 = 0;
        // ring 0 is consumed per half (by the edge CTA of the same half); the other rings by both stage CTAs
        const uint32_t* fa = fl_out + ((out_edge != 0 || h == 0) ? 2 : 3) * 32;
        const uint32_t* fb = fl_out + ((out_edge != 0 || h == 1) ? 3 : 2) * 32;
        for (int i = 0; i < n; ++i) {
          PW(0, mbar_wait(&bars[kBStgFull], i & 1));
          if (i >= kPipeRing) PW(1, poll2_ge(fa, fb, uint32_t(i - kPipeRing + 1), c0, c1));  // slot read out
          bulk_s2g(ring_out + size_t(i % kPipeRing) * kPTile + size_t(h) * kPHalf, smem + S::kStg, kPHalf);
          bulk_commit();
          PW(2, bulk_wait_read0());
          mbar_arrive(&bars[kBStgEmpty]);
          if (i >= kPStoreDepth) {  // the store of tile i - kPStoreDepth is complete: publish it (joins the next group)
            PW(3, bulk_wait_group<kPStoreDepth>());
            bulk_red_add_u32x4(fl_out + h * 32, one_s);
          }
        }
        PW(3, bulk_wait0());
        for (int i = (n > kPStoreDepth ? n - kPStoreDepth : 0); i < n; ++i) bulk_red_add_u32x4(fl_out + h * 32, one_s);
        bulk_commit();
        bulk_wait0();
        PW_FLUSH(7, 4);
      }
    } else if (warp == 3) {
      // =============================== edge bottom: dW_0 ===============================
      // (reached only by edge CTAs: stage CTAs took the phase-loader branch above)
      unsigned long long pw[5] = {0, 0, 0, 0, 0};
      const uint32_t idesc_w = idesc_bf16(128, 64, false, true);  // A = dTheta_0 (K-major), B = coordinates (MN-major)
      for (int i = lane; i < kPZSlots * kPBlk / 16; i += 32) sts128(sbase + S::kXb + i * 16, make_uint4(0u, 0u, 0u, 0u));
      __syncwarp();
      uint32_t k0 = 0;
      auto xa_row = [&](int i, int rr) {  // coordinate record of row lane + 32 rr of tile i
        const int T = pipe + (i >> 1) * p.pipelines;
        return p.xa + (size_t(T) * 128 + size_t(i & 1) * kPipeTileRows + lane + 32 * rr);
      };
      uint4 xv0 = __ldg(xa_row(0, 0)), xv1 = __ldg(xa_row(0, 1));
      // issue(i): bulk-load dTheta_0 tile i (this half) and stage its coordinate block (fetched one call ahead)
      auto issue = [&](int i) {
        const int zs = i % kPZSlots;
        if (i >= kPZSlots) PW(0, mbar_wait(&bars[kBZEmpty + zs], ((i / kPZSlots) - 1) & 1));
        if (lane == 0) {
          uint32_t dummy = 0xffffffffu;
          PW(1, poll2_ge(fl_in + h * 32, fl_in + h * 32, uint32_t(i + 1), k0, dummy));
          mbar_arrive_expect_tx(&bars[kBZFull + zs], kPHalf);
          bulk_g2s(smem + S::kZ + zs * kPHalf, ring_in + size_t(i % kPipeRing) * kPTile + size_t(h) * kPHalf, kPHalf,
                   &bars[kBZFull + zs]);
        }
        sts128(sbase + S::kXb + zs * kPBlk + sw128_chunk_off(lane, 0), xv0);
        sts128(sbase + S::kXb + zs * kPBlk + sw128_chunk_off(lane + 32, 0), xv1);
        if (i + 1 < n) {
          xv0 = __ldg(xa_row(i + 1, 0));
          xv1 = __ldg(xa_row(i + 1, 1));
        }
        fence_proxy_async_smem();
        __syncwarp();
      };
      for (int i = 0; i < kPZSlots - 1 && i < n; ++i) PW(3, issue(i));
      for (int i = 0; i < n; ++i) {
        const int zs = i % kPZSlots;
        if (lane == 0) {
          PW(2, mbar_wait(&bars[kBZFull + zs], (i / kPZSlots) & 1));
          const long long tm0 = prof_on ? clock64() : 0;
          st_relaxed_gpu(fl_in + (2 + h) * 32, uint32_t(i + 1));
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kPipeTileRows / 16; ++ks)
            umma_bf16_ss(t_w0, smem_desc(sbase + S::kZ + zs * kPHalf + ks * 32, hiK),
                         smem_desc(sbase + S::kXb + zs * kPBlk + ks * 2048, hiMN), idesc_w, (i | ks) != 0);
          umma_commit(&bars[kBZEmpty + zs]);
          if (prof_on) pw[4] += (unsigned long long)(clock64() - tm0);
        }
        __syncwarp();
        if (i + kPZSlots - 1 < n) PW(3, issue(i + kPZSlots - 1));
      }
      if (lane == 0) {
        umma_commit(&bars[kBFinB]);
        PW_FLUSH(18, 3);
        if (prof_on) {
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 26] = pw[3];
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 27] = pw[4];
        }
      }
    } else if (warp >= kPFirstEpiWarp) {
      // =============================== epilogue warps ===============================
      // Per tile: sin and cos of this CTA's 64 x 128 phases (thread = feature x 16 rows), dTheta = D .* cos -> staging
      // -> ring, y = sin -> operand of the weight-gradient MMA; edge CTAs also turn the dOut tile two tiles ahead into
      // its bf16 block.  One proxy fence and one round of barrier arrivals per tile.
      unsigned long long pw[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const long long te0 = prof_on ? clock64() : 0;
      const int et = threadIdx.x - kPFirstEpiWarp * 32;  // 0..511
      const int ew = warp - kPFirstEpiWarp;
      const int q = ew & 3;         // TMEM lane quadrant (== warp & 3)
      const int cg = ew >> 2;       // 16-row column group
      const int f = q * 32 + lane;  // feature inside this CTA's half
      const uint32_t t_lane = uint32_t(q * 32) << 16;
      const int C = p.C;
      float dbsum = 0.f;
      float dbf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      // edge: dOut tile i (fp32, bulk-copied to shared memory; a ragged last tile comes straight from global memory)
      // -> bf16 [64 rows][64] block: this thread converts row et & 63, columns 8 (et >> 6) .. +8 (zero beyond C / rows)
      auto convert_dout = [&](int i) {
        const int bs = i % kPDobSlots, rs = i % kPRawSlots;
        PW(0, mbar_wait(&bars[kBRawFull + rs], (i / kPRawSlots) & 1));
        if (i >= kPDobSlots) PW(0, mbar_wait(&bars[kBDobEmpty + bs], ((i / kPDobSlots) - 1) & 1));
        const int T = pipe + (i >> 1) * p.pipelines;
        const long long row0 = (long long)T * 128 + (i & 1) * kPipeTileRows;
        const int r = et & 63, c0 = (et >> 6) * 8;
        float gv[8];
        if (row0 + kPipeTileRows <= p.rows) {
          const uint32_t src = sbase + S::kRaw + rs * kPRawSlot + uint32_t(r * C + c0) * 4;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) gv[jj] = (c0 + jj < C) ? __uint_as_float(lds32(src + jj * 4)) : 0.f;
        } else {
          const bool valid = row0 + r < p.rows;
          const float* g = p.grad_out + (row0 + r) * C + c0;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) gv[jj] = (valid && c0 + jj < C) ? __ldg(g + jj) : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) dbf[jj] += gv[jj];
        sts128(sbase + S::kDob + bs * kPBlk + sw128_chunk_off(r, et >> 6),
               make_uint4(pack_bf16x2(gv[0], gv[1]), pack_bf16x2(gv[2], gv[3]), pack_bf16x2(gv[4], gv[5]),
                          pack_bf16x2(gv[6], gv[7])));
      };
      if (edge) {  // prologue: dOut blocks 0 and 1
        convert_dout(0);
        if (n > 1) convert_dout(1);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[kBDobFull + 0]);
          mbar_arrive(&bars[kBRawEmpty + 0]);
          if (n > 1) {
            mbar_arrive(&bars[kBDobFull + 1]);
            mbar_arrive(&bars[kBRawEmpty + 1]);
          }
        }
      }
      for (int i = 0; i < n; ++i) {
        const int as = i & 1, ps = i % kPPhSlots;
        PW(1, mbar_wait(&bars[kBPhFull + ps], (i / kPPhSlots) & 1));
        PW(3, mbar_wait(&bars[kBAccFull + as], (i >> 1) & 1));
        tc_fence_after();
        uint32_t ys[8], ds[8];
        PW(6, dbsum += sincos_tile(t_acc + t_lane + as * 64, sbase + S::kPh + ps * kPPhSlot, f, cg, ys, ds));
        tc_fence_before();
        if (i >= 1) {  // the previous tile's y has been consumed by its MMA, its dTheta half read out by the store
          PW(2, mbar_wait(&bars[kBYEmpty], (i - 1) & 1));
          PW(4, mbar_wait(&bars[kBStgEmpty], (i - 1) & 1));
        }
        store_feature_rows(sbase + S::kY, f, cg, ys);
        store_feature_rows(sbase + S::kStg, f, cg, ds);
        if (edge && i + 2 < n) {
          const long long td0 = prof_on ? clock64() : 0;
          convert_dout(i + 2);
          if (prof_on) pw[7] += (unsigned long long)(clock64() - td0);
        }
        PW(8, (fence_proxy_async_smem(), __syncwarp()));
        if (lane == 0) {
          mbar_arrive(&bars[kBYFull]);
          mbar_arrive(&bars[kBStgFull]);
          mbar_arrive(&bars[kBAccEmpty + as]);
          mbar_arrive(&bars[kBPhEmpty + ps]);
          if (edge && i + 2 < n) {
            mbar_arrive(&bars[kBDobFull + (i + 2) % kPDobSlots]);
            mbar_arrive(&bars[kBRawEmpty + (i + 2) % kPRawSlots]);
          }
        }
      }
      if (et == 0) {
        PW_FLUSH(11, 7);
        if (prof_on) {
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 28] = pw[7];
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 29] = (unsigned long long)(clock64() - te0);
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 30] = pw[8];
        }
      }

      // ---- flush: bias gradient of the layer whose dTheta this CTA produced
      {
        const float sc = (ph_layer == 0) ? p.omega0 : p.omegah;
        atomicAdd(p.grads + p.off[2 * ph_layer + 1] + h * 128 + f, sc * dbsum);
      }
      mbar_wait(&bars[kBFin], 0);
      tc_fence_after();
      if (!edge) {
        // dW_l[out][128 h + c] += omega_h * D[mh][out % 128][c]; this warp: M half cg >> 1, columns 64 (cg & 1) ..
        const int mh = cg >> 1;
        float* dst = p.grads + p.off[2 * layer] + (long long)(mh * 128 + f) * 256 + h * 128 + (cg & 1) * 64;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_w + t_lane + mh * 128 + (cg & 1) * 64 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4f(dst + c0 + j, p.omegah * __uint_as_float(v[j]), p.omegah * __uint_as_float(v[j + 1]),
                        p.omegah * __uint_as_float(v[j + 2]), p.omegah * __uint_as_float(v[j + 3]));
        }
      } else {
        // dW_f[c][128 h + f] += D[f][c]
        {
          uint32_t v[16];
          tmem_ld16(t_w + t_lane + cg * 16, v);
          tmem_ld_wait();
          float* dst = p.grads + p.off[2 * (L + 1)] + h * 128 + f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = cg * 16 + j;
            if (c < C) atomicAdd(dst + (long long)c * 256, __uint_as_float(v[j]));
          }
        }
        // db_f: column sums of dOut (one edge CTA per pipeline)
        if (h == 0) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            float s = dbf[jj];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const int c = (et >> 6) * 8 + jj;
            if (lane == 0 && c < C) atomicAdd(p.grads + p.off[2 * (L + 1) + 1] + c, s);
          }
        }
        // dW_0[128 h + f][j] += omega_0 * (D[f][j] + D[f][4 + j])   (x = hi + lo)
        mbar_wait(&bars[kBFinB], 0);
        tc_fence_after();
        if (cg == 0) {
          uint32_t v[16];
          tmem_ld16(t_w0 + t_lane, v);
          tmem_ld_wait();
          float* dst = p.grads + p.off[0] + (long long)(h * 128 + f) * p.d;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < p.d) atomicAdd(dst + j, p.omega0 * (__uint_as_float(v[j]) + __uint_as_float(v[4 + j])));
        }
      }
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
  if (prof_on && threadIdx.x == 0) {
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 0] = (unsigned long long)(clock64() - t_begin);
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 21] = (unsigned long long)n;
  }
}

int launch_siren_bwdp(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                      const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params, int num_sms,
                      cudaStream_t stream) {
  constexpr int H = 256;
  const int L = net->hidden_layers;
  const PipeStashLayout sl = make_pipe_stash_layout(H, L, rows);
  PipeParams p{};
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.pl = make_pack_layout(H, L);
  (void)coords;
  (void)grid;  // the coordinates come from the stash (written by the training forward)
  p.grad_out = grad_out;
  p.rows = rows;
  p.fwd_tiles = int(sl.tiles);
  p.L = L;
  p.C = net->out_features;
  p.d = net->in_features;
  uint8_t* st = reinterpret_cast<uint8_t*>(stash);
  p.ph = st + sl.ph;
  p.layer_stride = sl.layer_stride;
  p.xa = reinterpret_cast<const uint4*>(st + sl.xa);
  p.ring = st + sl.ring;
  p.flags = reinterpret_cast<uint32_t*>(st + sl.flags);
  p.grads = grad_params;
  int64_t off[2 * (kMaxSineLayers + 2)];
  param_offsets(p.d, H, L, p.C, off);
  for (int i = 0; i < 2 * (L + 2); ++i) p.off[i] = off[i];
  p.omega0 = net->first_omega_0;
  p.omegah = net->hidden_omega_0;
  const int S2 = 2 * (L + 1);
  int P = num_sms / S2;
  if (P > p.fwd_tiles) P = p.fwd_tiles;
  if (P < 1 || P * (L + 1) > kPipeMaxEdges) return B200INR_ERR_BAD_SHAPE;
  p.pipelines = P;
  const char* env_prof = getenv("B200INR_BWDP_PROF");
  if (env_prof != nullptr && env_prof[0] == '1' && P * S2 <= kPipeProfCtas)
    p.prof = reinterpret_cast<unsigned long long*>(st + sl.prof);
  if (cudaMemsetAsync(p.flags, 0, sl.flags_bytes, stream) != cudaSuccess) return B200INR_ERR_CUDA;
  const int smem = PSmem::kBytes + 1024;
  if (cudaFuncSetAttribute(siren_bwdp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  siren_bwdp_kernel<<<P * S2, kPThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
