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
// the 64-row tiles.  Every CTA is WEIGHT-STATIONARY for its whole life, with its weights IN TENSOR MEMORY:
//   stage CTA (l, h), l = L..1, h = 0/1 (feature half):                                   [the 256x256 layers]
//       TMEM: rows [128h, 128h+128) of W'_l^T (A operand, 128 columns), the block dW_l[:, 128h..]^T (256 columns),
//             one chain accumulator (64 columns) and the sin outputs y^T of the two epilogue groups (2 x 32
//             columns, A operand of the weight-gradient MMA) = all 512 columns  [kPYTmem; without it: two
//             accumulators and y^T in shared memory];
//       per tile:  receive dTheta_l (64 rows x 256, bf16, 32 KB)                          <- ring l
//                  D^T[in-half, rows]   = W'_l^T[in-half, :] dTheta_l^T                   (tcgen05 128 x 64 x 256, A in TMEM)
//                  s, c = sin, cos(phase_l-1) for its 128 features x 64 rows              (phases: bulk copy from HBM)
//                  dTheta_l-1[:, half]  = D .* c  -> bf16                                  -> ring l-1
//                  y^T = s -> bf16 -> tensor memory (tcgen05.st: thread = feature = lane)
//                  dW_l[:, half]^T     += y^T dTheta_l                                     (tcgen05 128 x 256 x 64)
//                  db_l-1[half]        += colsum(dTheta_l-1)                               (registers: lane = feature)
//   edge CTA E(h):                                                                        [both ends of the chain]
//       top:     dOut tile -> bf16, D^T = W_f^T[half] dOut^T, s, c of phase_L, dTheta_L = D .* c -> ring L,
//                dW_f^T[half] += y^T dOut, db_f, db_L
//       bottom:  receive dTheta_0[:, half]                                                <- ring 0
//                dW_0[half]   += dTheta_0^T [x_hi | x_lo]    (coordinate records stashed by the forward)
// The chain is computed TRANSPOSED (features on the 128 TMEM lanes, tile rows on the columns): a tile may have any row
// count (64 here) at full tensor-core rate, the bias gradient is a per-thread running sum, one thread owns one feature
// for 32 rows, so sin outputs and dTheta go to shared memory with 16-byte stores.  A dTheta tile is stored
// FEATURE-major ([64 features][64 rows] swizzled blocks): read with MN-major descriptors it is the B operand of the
// chain step (N = rows, K = features), with K-major descriptors the B operand of the weight-gradient step
// (N = features, K = rows).  Two groups of 8 epilogue warps alternate tiles, so one group's barrier / fence / TMEM
// latency hides behind the other's sin / cos work.  The hot loops are kept small on purpose (see the kInstr note).
//
// Rings: per pipeline and layer boundary kPipeRing slots of one tile; producers bulk-store their half and publish it
// with a bulk add on a counter (no generic-proxy fence), consumers poll (acquire), bulk-load and return a credit
// counter (all bounded spins: a mis-programmed pipeline traps instead of hanging).  All 2(L+1)P CTAs must be
// co-resident: the grid never exceeds the SM count (1 CTA/SM).
//
// Warp roles (736 threads): 0 = ring loader (edge: phase + raw dOut loader), 1 = chain MMA issuer + TMEM owner,
// 2 = ring store, 3 = phase loader (edge: the whole bottom half), 4..19 = epilogue (TMEM lane quadrant = warp & 3),
// 20..21 = edge: dOut fp32 -> bf16 converters (idle in stage CTAs), 22 = weight-gradient MMA issuer.
//
// Tuning builds (tools/build_variant.sh <name> mlp_bwdp.cu -D...; tools/run_variants.sh times them on the box):
// B200INR_PTRACE=1 compiles the event trace into the lean kernel, B200INR_PKO=<mask> knocks one resource out at a
// time, B200INR_PMC=1 multicasts ring tiles across a stage's CTA pair, B200INR_PYT / PDZ / PPH / PEPH / PSTG / PRAW /
// PSD / PCVT / PPF choose buffer placement and depths.  The defaults are the measured best.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "umma.cuh"

#ifndef B200INR_TUNING
#define B200INR_TUNING 0  // 1: the launcher honours the B200INR_BWDP_* environment hooks (trace pointer, stall counters)
#endif

namespace b200inr {

#ifndef B200INR_PCVT
#define B200INR_PCVT 2
#endif
constexpr int kPCvtWarps = B200INR_PCVT;                      // 2 (one pair) or 4 (two pairs alternating tiles)
constexpr int kPThreads = 672 + 32 * kPCvtWarps;              // 736 / 800
constexpr int kPEpiWarps = 16;
constexpr int kPFirstEpiWarp = 4;
constexpr int kPCvtPair = 2;                                  // warps converting one tile (64 rows)
constexpr int kPFirstCvtWarp = kPFirstEpiWarp + kPEpiWarps;  // 20
constexpr int kPWgradWarp = kPFirstCvtWarp + kPCvtWarps;      // 22
#ifndef B200INR_PYT
#define B200INR_PYT 1
#endif
#ifndef B200INR_PKO
#define B200INR_PKO 0
#endif
// tuning aid, never set in production builds: knock out one resource at a time (results are garbage) to see what the
// tile period is made of.  1 = no sin / cos, 2 = one chain MMA instead of 16, 4 = one weight-gradient MMA instead of 4,
// 8 = no phase loads, 16 = ring loads of half a tile, 32 = no dOut conversion.
constexpr int kPKo = B200INR_PKO;
#ifndef B200INR_PGPH
#define B200INR_PGPH 1
#endif
#ifndef B200INR_PW16
#define B200INR_PW16 1  // all 16 chain MMAs of a tile behind one election (0: four batches of four)
#endif
// kPGroupPh: each epilogue group owns two of four phase slots (and their barriers), so a warp no longer has to pass --
// wait + arrive, ~100 cycles each on the busy shared-memory pipe -- the barriers of the OTHER group's tiles.  Needs four
// phase slots in both CTA kinds, which only fit without the 1 KB alignment reserve of the dynamic shared memory (its
// base is 1024-aligned on this architecture; the kernel traps if it is not).
constexpr bool kPGroupPh = B200INR_PGPH != 0;
#ifndef B200INR_PTRACE
#define B200INR_PTRACE 0
#endif
constexpr bool kPTraceOnly = B200INR_PTRACE != 0;  // tuning build: event trace compiled into the lean kernel
#ifndef B200INR_PEND
#define B200INR_PEND 0
#endif
constexpr bool kPEndOnly = B200INR_PEND != 0;  // tuning build: the lean kernel records every CTA's total cycle count
#ifndef B200INR_PPF
#define B200INR_PPF 0
#endif
constexpr int kPPhPrefetch = B200INR_PPF;  // forward tiles of phases pulled into L2 ahead of the bulk copies (0 = off)
#ifndef B200INR_PADAPT
#define B200INR_PADAPT 1
#endif
// kPAdapt: the pipelines' tile shares follow the speed each pipeline showed in the PREVIOUS launch on the same stash.
// Pipelines are independent chains with static shares, and they do not run equally fast: with equal shares the
// per-pipeline finish times of a cfg2 launch spread over 1251 .. 1365 us (the same pipelines slow in every launch -- it
// follows the SMs a pipeline sits on), so the kernel ran at the pace of its slowest pipeline, 5 % behind the mean.
// Every pipeline records (tiles, cycles) at its end; the next launch splits the tiles in proportion to tiles / cycles
// (clamped, validated: anything stale or foreign gives equal shares).  Records live in the unused tail of the stash's
// profiling area, two banks by launch parity, so a launch never reads what it writes.
constexpr bool kPAdapt = B200INR_PADAPT != 0;
constexpr uint32_t kCalMagic = 0xB2001A7Eu;
constexpr int kCalBankWords = 4 * 32;  // per bank: 32 pipelines x {tiles, cycles, epoch, magic ^ P}
constexpr int kCalRow = kPipeCalRow;   // first row of the profiling area used (CTAs write rows < grid <= 148)
#ifndef B200INR_PMC
#define B200INR_PMC 0
#endif
// kPMc: the two CTAs of a stage (feature halves h = 0 / 1) form a thread-block cluster; each bulk-copies HALF of the
// incoming dTheta tile and multicasts it to both, so a ring tile crosses the L2 fabric once instead of twice (ring
// loads are 2/3 of this kernel's L2 -> SM traffic, which runs near the fabric's limit).
constexpr bool kPMc = B200INR_PMC != 0;
#if B200INR_PMC
#define B200INR_PCLUSTER __cluster_dims__(2, 1, 1)
#else
#define B200INR_PCLUSTER
#endif
#ifndef B200INR_PDZ
#define B200INR_PDZ (B200INR_PYT ? 4 : 3)
#endif
#ifndef B200INR_PPH
#define B200INR_PPH (B200INR_PGPH ? 4 : 3)
#endif
// kPYTmem: the sin outputs (A operand of the weight-gradient MMA) are written to TENSOR memory instead of shared memory
// (thread = feature = TMEM lane, its 32 rows = 16 packed columns): 32 KB less shared-memory traffic per tile and 32 KB
// of shared memory freed.  The columns come from the second chain accumulator: ONE accumulator, handed back right
// after the epilogue's tcgen05.ld.
constexpr bool kPYTmem = B200INR_PYT != 0;
constexpr int kPDzSlots = B200INR_PDZ;          // incoming dTheta tiles
constexpr int kPPhSlots = B200INR_PPH;          // phase tiles (stage CTAs)
#ifndef B200INR_PEPH
#define B200INR_PEPH (B200INR_PGPH ? 4 : 3)
#endif
constexpr int kPPhBars = B200INR_PPH > B200INR_PEPH ? B200INR_PPH : B200INR_PEPH;  // phase slots of the edge CTAs: B200INR_PEPH
#ifndef B200INR_PSTG
#define B200INR_PSTG 2
#endif
constexpr int kPStgSlots = B200INR_PSTG;        // outgoing dTheta halves (even: a slot always belongs to one epilogue group)
constexpr int kPDobSlots = 4;                   // edge: bf16 dOut blocks
#ifndef B200INR_PRAW
#define B200INR_PRAW (B200INR_PGPH ? 4 : (B200INR_PYT ? 6 : 2))
#endif
constexpr int kPRawSlots = B200INR_PRAW;        // edge: raw fp32 dOut tiles, bulk-copied ahead of the conversion (the
                                                // copies come from HBM: two slots leave the converters latency-bound)
constexpr int kPZSlots = 2;                     // edge bottom: dTheta_0 slots
#ifndef B200INR_PSD
#define B200INR_PSD 2
#endif
constexpr int kPStoreDepth = B200INR_PSD;                 // bulk stores in flight per ring-store thread (0 / 1 / 2 / 3:
                                                          // 1.30 / 1.28 / 1.255 / 1.31 ms for the cfg2 backward)
constexpr int kPBlk = kPipeTileRows * 128;      // one [64][64] bf16 block: 8 KB
constexpr int kPHalf = 2 * kPBlk;               // 128 features of a tile: 16 KB
constexpr int kPTile = 4 * kPBlk;               // 256 features of a tile: 32 KB
constexpr int kPPhChunk = kPipePhChunk;              // phases of 8 features x 64 rows, padded (bank spread)
constexpr int kPPhSlot = 16 * kPPhChunk;            // 16 chunks = 128 features
constexpr int kPRawSlot = kPipeTileRows * kOutPad * 4;  // fp32 dOut tile, C <= 32

// shared memory maps (bytes); operand blocks are 1024-byte aligned
struct PSmemE {  // edge CTA
  static constexpr int kWf = 0;                              // W_f^T half [128][64]           16 KB
  static constexpr int kDob = kWf + 128 * 128;               // kPDobSlots x dOut block        32 KB
  static constexpr int kZ = kDob + kPDobSlots * kPBlk;       // kPZSlots x dTheta_0 half       32 KB
  static constexpr int kXb = kZ + kPZSlots * kPHalf;         // kPZSlots x coordinate block    16 KB
  static constexpr int kRaw = kXb + kPZSlots * kPBlk;        // kPRawSlots x fp32 dOut tile    8 KB each
  static constexpr int kY = kRaw + kPRawSlots * kPRawSlot;   // 2 x sin outputs                32 KB
  static constexpr int kStg = kY + (kPYTmem ? 0 : 2 * kPHalf);               // kPStgSlots x dTheta half       32 KB
  static constexpr int kPh = kStg + kPStgSlots * kPHalf;     // 3 phase slots                  48.75 KB
  static constexpr int kPhSlots = B200INR_PEPH;
  static constexpr int kEnd = kPh + kPhSlots * kPPhSlot;
};
struct PSmem {  // stage CTA
  static constexpr int kDz = 0;                              // kPDzSlots x tile               96 KB
  static constexpr int kY = kDz + kPDzSlots * kPTile;        // 2 x sin outputs (128 features)  32 KB
  static constexpr int kStg = kY + (kPYTmem ? 0 : 2 * kPHalf);               // kPStgSlots x dTheta half       32 KB
  static constexpr int kPh = kStg + kPStgSlots * kPHalf;     // kPPhSlots phase slots          65 KB
  static constexpr int kEnd = kPh + kPPhSlots * kPPhSlot;
  static constexpr int kBar = kEnd > PSmemE::kEnd ? kEnd : PSmemE::kEnd;  // barrier area, shared by both CTA kinds
  static constexpr int kBytes = kBar + 512;
  static constexpr int kSlack = kPGroupPh ? 0 : 1024;  // manual 1024-byte alignment reserve
  static_assert(kBytes + kSlack <= 232448, "shared memory budget");
};

// barrier indices
enum PBar : int {
  // NOTE: a parity wait may be at most one phase behind its barrier, so every barrier is waited on, phase after
  // phase, by the same thread(s): barriers used by the epilogue come in per-group (tile parity) pairs.
  kBW = 0,                                // static weights in place
  kBDzFull = kBW + 1,                     // [kPDzSlots]
  kBDzEmpty = kBDzFull + kPDzSlots,       // [kPDzSlots]
  kBPhFull = kBDzEmpty + kPDzSlots,       // [kPPhBars]
  kBPhEmpty = kBPhFull + kPPhBars,        // [kPPhBars]
  kBAccFull = kBPhEmpty + kPPhBars,      // [2] by tile parity (ONE accumulator: the barriers alternate, not the storage)
  kBAccEmpty = kBAccFull + 2,             // [2]
  kBYFull = kBAccEmpty + 2,               // [2]
  kBYEmpty = kBYFull + 2,                 // [2]
  kBStgFull = kBYEmpty + 2,               // [kPStgSlots]
  kBStgEmpty = kBStgFull + kPStgSlots,    // [kPStgSlots]
  kBFin = kBStgEmpty + kPStgSlots,
  kBDobFull = kBFin + 1,                  // [kPDobSlots] edge
  kBDobEmpty = kBDobFull + kPDobSlots,    // [kPDobSlots] edge
  kBZFull = kBDobEmpty + kPDobSlots,      // [kPZSlots] edge bottom
  kBZEmpty = kBZFull + kPZSlots,          // [kPZSlots] edge bottom
  kBFinB = kBZEmpty + kPZSlots,
  kBRawFull = kBFinB + 1,                 // [kPRawSlots] edge
  kBRawEmpty = kBRawFull + kPRawSlots,    // [kPRawSlots] edge
  kBCount = kBRawEmpty + kPRawSlots
};
static_assert(kPStgSlots % 2 == 0 && kPZSlots == 2, "per-parity buffers");

struct PipeParams {
  const uint8_t* packed;
  PackLayout pl;
  const float* grad_out;  // [rows, C]
  long long rows;
  int fwd_tiles;          // 128-row tiles of the forward stash
  int L, C, d;
  int Hr;                    // parameter width (<= kSirenWidth)
  const uint8_t* ph;      // phase stash: (L+1) x fwd_tiles x kPipePhTile bytes (common.cuh)
  size_t layer_stride;
  const uint4* xa;        // coordinate stash: per row bf16 {hi x4, lo x4} (x = hi + lo), written by the forward
  uint8_t* ring;          // [pipeline][edge 0..L][kPipeRing][32 KB]
  uint32_t* flags;        // [pipeline][edge][4] counters, 128 bytes apart: produced by half 0/1, consumed by half 0/1
  float* grads;
  long long off[2 * (kMaxSineLayers + 2)];
  float omega0, omegah;
  int relu_tail;             // B200INR_NET_RELU_TAIL: activated layer L is Linear + ReLU (its stash slot holds bf16 y)
  int pipelines;
  int skew_ns;               // start-up delay of the second epilogue group
  int dbg;                   // tuning switches (B200INR_BWDP_DBG), 0 in production
  uint32_t* trace;           // nullptr, or the event trace buffer (see TR)
  unsigned long long* prof;  // nullptr, or [grid][kPipeProfSlots] stall-cycle counters (B200INR_BWDP_PROF=1)
  int skip_ph0;              // kPipeSkipPh0 and L >= 1: layer-0 phases are recomputed from fp32 coordinate records
  uint32_t* calib;           // speed records of the previous launch on this stash (see kPAdapt): [epoch, pad x15, 2 banks]
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
// Event trace (tuning aid, B200INR_BWDP_TRACE_PTR = device pointer): pipeline 0 records the time of event `ev` of tile
// `tile` as trace[(role * kTraceTiles + tile) * kTraceEvents + ev] (cycles since kernel start, 32 bit).
constexpr int kTraceTiles = 256;
constexpr int kTraceEvents = 24;
#define TR(ev, tile)                                                                                          \
  do {                                                                                                        \
    if (trace_on && (tile) < kTraceTiles)                                                                     \
      p.trace[(size_t(role) * kTraceTiles + (tile)) * kTraceEvents + (ev)] = uint32_t(clock64() - t_begin);  \
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

static __device__ __noinline__ void flag_timeout(const uint32_t* f0, const uint32_t* f1, uint32_t target, uint32_t a,
                                          uint32_t b) {
  printf("b200inr: pipeline flag timeout block %d thread %d flags %p %p target %u (%u, %u)\n", blockIdx.x, threadIdx.x,
         (const void*)f0, (const void*)f1, target, a, b);
  __trap();
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
    if (++spins > (1u << 23)) flag_timeout(f0, f1, target, a, b);
  }
}

__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.b16 %0, [%1];\n" : "=h"(v) : "r"(addr));
  return v;
}

// ---- tensor-memory operand helpers
// 32 lanes x 8 / 16 consecutive 32-bit columns: thread i of the warp writes lane (base_lane + i).
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// D[tmem] (+)= A[tmem: lane = M row, 2 bf16 per column along K] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// phase (u16, 65536 = one turn) -> radians: float(2^23 + ph) is built with one PRMT, then one FFMA
constexpr float kPhToRad = 9.587379924285257e-05f;   // 2*pi / 65536
constexpr float kPhBias = -804.247719318987f;        // -(2^23) * 2*pi / 65536
__device__ __forceinline__ float rad_lo16(uint32_t w) {  // w: zero-extended 16-bit phase
  return fmaf(__uint_as_float(w | 0x4B000000u), kPhToRad, kPhBias);
}

// kInstr = true compiles the stall counters / event trace in (tuning runs only): they double every wait statement and
// push the hot loops out of the instruction cache (no_instruction stalls 1.8 per issue with them).
// kRelu: the ReLU-tail network (B200INR_NET_RELU_TAIL).  A TEMPLATE switch, not a run-time one: the hot loops of this
// kernel live on the edge of the instruction cache, and the extra path cost the plain SIREN 13 % (cycles 1.93 M ->
// 2.19 M, no_instruction stalls 0.21 -> 0.36 per issue) when it was a branch.
template <bool kInstr, bool kRelu = false>
__global__ void B200INR_PCLUSTER __launch_bounds__(kPThreads, 1) siren_bwdp_kernel(const PipeParams p) {
  using S = PSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (PSmem::kSlack == 0 && smem != smem_raw) __trap();  // no reserve: the base must already be aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBCount);
  uint32_t* one_s = reinterpret_cast<uint32_t*>(smem + S::kBar + 448);  // 16-byte aligned constant {1, 0, 0, 0}
  static_assert(kBCount * 8 + 4 <= 448 && 448 + 16 <= 512, "barrier area layout");
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

  uint32_t* share_s = reinterpret_cast<uint32_t*>(smem + S::kBar + 464);  // {first forward tile, count, epoch}
  static_assert(464 + 12 <= 512, "barrier area layout");
  const long long t_launch = clock64();

  uint8_t* ring_in = p.ring + size_t(pipe * (L + 1) + in_edge) * kPipeRing * kPTile;
  uint8_t* ring_out = p.ring + size_t(pipe * (L + 1) + out_edge) * kPipeRing * kPTile;
  uint32_t* fl_in = p.flags + size_t(pipe * (L + 1) + in_edge) * 4 * 32;
  uint32_t* fl_out = p.flags + size_t(pipe * (L + 1) + out_edge) * 4 * 32;

  if (threadIdx.x == 0) {
    one_s[0] = 1u;  // {1, 0, 0, 0}: source of the bulk add that publishes a ring slot
    one_s[1] = one_s[2] = one_s[3] = 0u;
    mbar_init(&bars[kBW], edge ? 1 : 4);
    for (int i = 0; i < kPDzSlots; ++i) {
      mbar_init(&bars[kBDzFull + i], 1);
      mbar_init(&bars[kBDzEmpty + i], kPMc ? 2 : 1);  // multicast loads: released by both CTAs of the pair
    }
    for (int i = 0; i < kPPhBars; ++i) {
      mbar_init(&bars[kBPhFull + i], 1);
      mbar_init(&bars[kBPhEmpty + i], kPGroupPh ? kPEpiWarps / 2 : kPEpiWarps);
    }
    for (int i = 0; i < kPDobSlots; ++i) {
      mbar_init(&bars[kBDobFull + i], kPCvtPair);
      mbar_init(&bars[kBDobEmpty + i], 1);
    }
    for (int i = 0; i < kPRawSlots; ++i) {
      mbar_init(&bars[kBRawFull + i], 1);
      mbar_init(&bars[kBRawEmpty + i], kPMc ? 2 * kPCvtPair : kPCvtPair);  // multicast: both CTAs' converters
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kBAccFull + i], 1);
      mbar_init(&bars[kBAccEmpty + i], kPEpiWarps / 2);
      mbar_init(&bars[kBYFull + i], kPEpiWarps / 2);
      mbar_init(&bars[kBYEmpty + i], 1);
      mbar_init(&bars[kBZFull + i], 1);
      mbar_init(&bars[kBZEmpty + i], 1);
    }
    for (int i = 0; i < kPStgSlots; ++i) {
      mbar_init(&bars[kBStgFull + i], kPEpiWarps / 2);
      mbar_init(&bars[kBStgEmpty + i], 1);
    }
    mbar_init(&bars[kBFin], 1);
    mbar_init(&bars[kBFinB], 1);
    fence_mbar_init();
    fence_proxy_async_smem();  // one_s is read by the bulk-copy engine
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (warp == 2) {
    // ---- tile shares: pipeline q walks the forward tiles [b_q, b_q+1), b = rounded prefix sums of the pipelines'
    //      rates.  Every CTA of the grid computes the same boundaries from the same records (same operations in the
    //      same order), one record per lane.
    const int P = p.pipelines, N = p.fwd_tiles;
    uint32_t epoch = 0;
    bool ok = false;
    float r = 0.f;
    if (kPAdapt && p.calib != nullptr && P <= 32) {
      epoch = __ldcg(p.calib);
      uint4 e = make_uint4(1u, 1u, epoch, kCalMagic ^ uint32_t(P));
      if (lane < P) e = __ldcg(reinterpret_cast<const uint4*>(p.calib + 16 + (epoch & 1u) * kCalBankWords) + lane);
      ok = __all_sync(0xffffffffu, e.w == (kCalMagic ^ uint32_t(P)) && e.z == epoch && e.x > 0u && e.y > 0u);
      r = float(e.x) / float(e.y);
    }
    if (!ok) r = 1.f;
    if (lane >= P) r = 0.f;
    auto warp_sum = [](float v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      return v;
    };
    const float mean = warp_sum(r) / float(P);
    if (lane < P) r = fminf(fmaxf(r, 0.8f * mean), 1.25f * mean);  // (a pipeline is never starved or flooded)
    const float tot = warp_sum(r);
    float inc = r;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int bnd = (lane >= P - 1) ? N : int(floorf(float(N) * inc / tot + 0.5f));
    bnd = bnd > N ? N : bnd;
    int prev = __shfl_up_sync(0xffffffffu, bnd, 1);
    if (lane == 0) prev = 0;
    if (P > 32) {  // (more pipelines than lanes: equal contiguous shares)
      prev = int((long long)N * pipe / P);
      bnd = int((long long)N * (pipe + 1) / P);
    }
    if (lane == (P > 32 ? 0 : pipe)) {
      share_s[0] = uint32_t(prev);
      share_s[1] = uint32_t(bnd > prev ? bnd - prev : 0);
      share_s[2] = epoch;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPMc) cluster_sync_all();  // the peer's barriers exist before any multicast copy / remote commit reaches them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // did the forward leave coordinate records instead of layer-0 phases?  (its word in the stash; see common.cuh)
  const bool skip0 = p.skip_ph0 != 0 && p.calib != nullptr && __ldcg(p.calib + kPipeSkipWord) != 0u;
  // tiles of this pipeline: forward tiles t_first .. t_first + my_fwd - 1; two 64-row tiles each
  const int t_first = int(share_s[0]);
  const int my_fwd = int(share_s[1]);
  const int n = 2 * my_fwd;
  // TMEM columns.  stage: dW^T block [0,256), W'^T half [256,384), chain accumulators 384 + 64 j.
  //               edge : chain accumulators 0 + 64 j, dW_f^T [128,192), dW_0 [192,256).
  const uint32_t t_w = edge ? tmem + 128 : tmem;
  const uint32_t t_wt = tmem + 256;
  const uint32_t t_acc = edge ? tmem : tmem + 384;
  const uint32_t t_y = t_acc + 64;  // kPYTmem: y^T of the two epilogue groups, 32 columns each (64 rows, 2 bf16 per column)
  const uint32_t t_w0 = tmem + 192;

  using E = PSmemE;
  const uint32_t oY = sbase + (edge ? E::kY : S::kY);        // 2 x y operand (feature-major [128][64 rows])
  const uint32_t oStg = edge ? E::kStg : S::kStg;            // staging halves (byte offset from smem)
  const uint32_t oPh = edge ? E::kPh : S::kPh;               // phase slots (byte offset from smem)
  const int nph = edge ? E::kPhSlots : kPPhSlots;

  const uint64_t hiK = smem_desc_hi_sw128(0, 1024);       // K-major blocks
  const uint64_t hiMN = smem_desc_hi_sw128(kPBlk, 1024);  // MN-major: 64-wide MN blocks kPBlk apart

  const bool prof_on = kInstr && p.prof != nullptr;
  const bool trace_on = (kInstr || kPTraceOnly) && p.trace != nullptr && pipe == 0;
  const bool end_on = kPEndOnly && p.prof != nullptr;
  const long long t_begin = (prof_on || trace_on || end_on) ? clock64() : 0;
  if (end_on && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 2] = gt;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 4] = smid;
  }

  if (n > 0) {
    if (warp == 0) {
      // =============================== loader ===============================
      if (lane == 0) {
        if (!edge) {
          unsigned long long pw[2] = {0, 0};
          uint32_t k0 = 0, k1 = 0;
          for (int i = 0; i < n; ++i) {
            const int slot = i % kPDzSlots, round = i / kPDzSlots;
            if (round > 0) PW(0, mbar_wait(&bars[kBDzEmpty + slot], (round - 1) & 1));
            TR(0, i);
            if (kPMc) {
              // this CTA fetches the half its namesake produced (blocks 2h, 2h + 1) for both CTAs of the pair; the slot
              // is free in BOTH (kBDzEmpty counts both weight-gradient streams) and the barrier expects the whole tile
              PW(1, poll2_ge(fl_in + h * 32, fl_in + h * 32, uint32_t(i + 1), k0, k1));
              TR(1, i);
              mbar_arrive_expect_tx(&bars[kBDzFull + slot], kPTile);
              bulk_g2s_mc(smem + S::kDz + slot * kPTile + h * kPHalf,
                          ring_in + size_t(i % kPipeRing) * kPTile + size_t(h) * kPHalf, kPHalf, &bars[kBDzFull + slot],
                          uint16_t(3));
            } else {
              PW(1, poll2_ge(fl_in + 0 * 32, fl_in + 1 * 32, uint32_t(i + 1), k0, k1));
              TR(1, i);
              // (no proxy fence: the tile was written AND published through the async proxy, and is read through it)
              mbar_arrive_expect_tx(&bars[kBDzFull + slot], kPTile);
              bulk_g2s(smem + S::kDz + slot * kPTile, ring_in + size_t(i % kPipeRing) * kPTile, kPTile,
                       &bars[kBDzFull + slot]);
            }
          }
          PW_FLUSH(1, 2);
        } else {
          mbar_arrive_expect_tx(&bars[kBW], 128 * 128);
          bulk_g2s(smem + E::kWf, p.packed + p.pl.wft + size_t(h) * 128 * 128, 128 * 128, &bars[kBW]);
        }
      }
    }
    if ((edge && warp == 0) || (!edge && warp == 3)) {
      // =============================== phase loader ===============================
      if (lane == 0) {
        unsigned long long pw[1] = {0};
        for (int i = 0; i < n; ++i) {
          if (edge) {
            // the raw fp32 dOut tile of tile i for the converters (kPRawSlots ahead of them; issued from here because
            // the converter loop is the longest serial loop of the edge CTA)
            const int rs = i % kPRawSlots;
            if (i >= kPRawSlots) mbar_wait(&bars[kBRawEmpty + rs], ((i / kPRawSlots) - 1) & 1);
            const long long row0 = (long long)(t_first + (i >> 1)) * 128 + (i & 1) * kPipeTileRows;
            if (row0 + kPipeTileRows <= p.rows) {
              const uint32_t bytes = uint32_t(kPipeTileRows) * p.C * 4;
              mbar_arrive_expect_tx(&bars[kBRawFull + rs], bytes);
              if (kPMc)  // both edge CTAs convert the same dOut tile: each fetches 32 rows of it for the pair
                bulk_g2s_mc(smem + E::kRaw + rs * kPRawSlot + h * (bytes / 2), p.grad_out + row0 * p.C + h * (bytes / 8),
                            bytes / 2, &bars[kBRawFull + rs], uint16_t(3));
              else
                bulk_g2s(smem + E::kRaw + rs * kPRawSlot, p.grad_out + row0 * p.C, bytes, &bars[kBRawFull + rs]);
            } else {
              mbar_arrive(&bars[kBRawFull + rs]);  // ragged last tile: the converters read it from global memory
            }
          }
          // shared slots: tile i -> slot i mod nph; per-group slots: group i & 1 owns slots {2g, 2g + 1}
          const int slot = kPGroupPh ? ((i & 1) * 2 + ((i >> 1) & 1)) : i % nph;
          const int round = kPGroupPh ? (i >> 2) : i / nph;
          if (round > 0) PW(0, mbar_wait(&bars[kBPhEmpty + slot], (round - 1) & 1));
          TR(18, i);
          const int T = t_first + (i >> 1);
          if (kPPhPrefetch > 0 && (i & 1) == 0 && (i >> 1) + kPPhPrefetch < my_fwd)
            bulk_prefetch_l2(p.ph + size_t(ph_layer) * p.layer_stride + size_t(T + kPPhPrefetch) * kPipePhTile,
                             kPipePhTile);
          const uint8_t* src = p.ph + size_t(ph_layer) * p.layer_stride + size_t(T) * kPipePhTile +
                               size_t(i & 1) * kPipePhHalf + size_t(h) * kPPhSlot;
          if (kPKo & 8) {
            mbar_arrive(&bars[kBPhFull + slot]);
            continue;
          }
          if (!edge && ph_layer == 0 && skip0) {
            // layer 0 is not stashed: the tile's 64 coordinate records (fp32 x4, 1 KB) instead of 16.5 KB of phases
            mbar_arrive_expect_tx(&bars[kBPhFull + slot], kPipeTileRows * 16);
            bulk_g2s(smem + oPh + slot * kPPhSlot, p.ph + (size_t(T) * 128 + size_t(i & 1) * kPipeTileRows) * 16,
                     kPipeTileRows * 16, &bars[kBPhFull + slot]);
            TR(17, i);
            continue;
          }
          // one copy: the forward stores the 16 chunks of a feature half contiguously, padding included
          mbar_arrive_expect_tx(&bars[kBPhFull + slot], kPPhSlot);
          bulk_g2s(smem + oPh + slot * kPPhSlot, src, kPPhSlot, &bars[kBPhFull + slot]);
          TR(17, i);
        }
        PW_FLUSH(3, 1);
      }
    } else if (warp == 1) {
      // =============================== chain MMA issuer ===============================
      // Chain step of tile i + 1 as soon as its operand tile is here and the (single) accumulator has been read out by
      // the epilogue group of tile i.  The whole warp runs the loop converged, one elected lane issues (umma_*_w).
      {
        unsigned long long pw[4] = {0, 0, 0, 0};
        mbar_wait(&bars[kBW], 0);
        tc_fence_after();
        if (!edge) {
          const uint32_t idesc_d = idesc_bf16(128, kPipeTileRows, false, true);  // A = W'^T (TMEM), B = dTheta (MN-major)
          for (int i = 0; i < n; ++i) {  // D^T[in-half][rows] = sum over the 256 outputs: 16 K steps of 16 features
            const int ds = i % kPDzSlots;
            PW(0, mbar_wait(&bars[kBDzFull + ds], (i / kPDzSlots) & 1));
            if (lane == 0) TR(2, i);
            if (lane == 0) st_relaxed_gpu(fl_in + (2 + h) * 32, uint32_t(i + 1));  // credit: ring slot read out
            if (kPYTmem) {  // one accumulator: read out by the group of tile i - 1
              if (i >= 1) PW(1, mbar_wait(&bars[kBAccEmpty + ((i - 1) & 1)], ((i - 1) >> 1) & 1));
            } else {
              if (i >= 2) PW(1, mbar_wait(&bars[kBAccEmpty + (i & 1)], ((i >> 1) - 1) & 1));
            }
            if (lane == 0) TR(3, i);
            tc_fence_after();
            const uint32_t dz = sbase + S::kDz + ds * kPTile;
#if B200INR_PW16
            // all 16 MMAs of the step (4 K blocks of kPBlk bytes x 4 K steps of 2048 bytes) behind one election
            umma_bf16_ts_w16<2048 / 16, kPBlk / 16>(t_acc + (kPYTmem ? 0 : (i & 1) * 64), t_wt, smem_desc(dz, hiMN), idesc_d);
#else
#pragma unroll 1  // (compact loops: this warp shares an instruction cache with four epilogue warps)
            for (int kb = 0; kb < ((kPKo & 2) ? 1 : 4); ++kb)  // four K steps (2048 bytes apart) per batched issue
              umma_bf16_ts_w4<2048 / 16>(t_acc + (kPYTmem ? 0 : (i & 1) * 64), t_wt + kb * 32,
                                         smem_desc(dz + kb * kPBlk, hiMN), idesc_d, kb != 0);
#endif
            umma_commit_w(&bars[kBAccFull + (i & 1)]);
            if (lane == 0) TR(4, i);
          }
        } else {
          const uint32_t idesc_d = idesc_bf16(128, kPipeTileRows, false, false);  // A = W_f^T, B = dOut: both K-major
          for (int i = 0; i < n; ++i) {
            const int bs = i % kPDobSlots;
            PW(0, mbar_wait(&bars[kBDobFull + bs], (i / kPDobSlots) & 1));
            if (lane == 0) TR(2, i);
            if (kPYTmem) {
              if (i >= 1) PW(1, mbar_wait(&bars[kBAccEmpty + ((i - 1) & 1)], ((i - 1) >> 1) & 1));
            } else {
              if (i >= 2) PW(1, mbar_wait(&bars[kBAccEmpty + (i & 1)], ((i >> 1) - 1) & 1));
            }
            if (lane == 0) TR(3, i);
            tc_fence_after();
            const uint32_t dob = sbase + E::kDob + bs * kPBlk;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss_w(t_acc + (kPYTmem ? 0 : (i & 1) * 64), smem_desc(sbase + E::kWf + k4 * 32, hiK), smem_desc(dob + k4 * 32, hiK), idesc_d,
                             k4 != 0);
            umma_commit_w(&bars[kBAccFull + (i & 1)]);
            if (lane == 0) TR(4, i);
          }
        }
        if (lane == 0) PW_FLUSH(4, 2);
      }
    } else if (warp == kPWgradWarp) {
      // =============================== weight-gradient MMA issuer ===============================
      // Its own warp: the chain step of the next tile must not queue behind the wait for this tile's epilogue.
      {
        unsigned long long pw[1] = {0};
        if (!edge) {
          const uint32_t idesc_w = idesc_bf16(128, 256, false, false);  // A = y^T, B = dTheta: both K-major
          for (int i = 0; i < n; ++i) {
            const int yb = i & 1, ds = i % kPDzSlots;
            const uint32_t dz = sbase + S::kDz + ds * kPTile;
            mbar_wait(&bars[kBDzFull + ds], (i / kPDzSlots) & 1);
            PW(0, mbar_wait(&bars[kBYFull + yb], (i >> 1) & 1));
            if (lane == 0) TR(5, i);
            tc_fence_after();
            // dW^T[in-half][out] += sum over the 64 rows: 4 K steps of 16 rows, N = all 256 outputs
#pragma unroll
            for (int ks = 0; ks < ((kPKo & 4) ? 1 : kPipeTileRows / 16); ++ks)
              if (kPYTmem)
                umma_bf16_ts_w(t_w, t_y + yb * 32 + ks * 8, smem_desc(dz + ks * 32, hiK), idesc_w, (i | ks) != 0);
              else
                umma_bf16_ss_w(t_w, smem_desc(oY + yb * kPHalf + ks * 32, hiK), smem_desc(dz + ks * 32, hiK), idesc_w,
                               (i | ks) != 0);
            umma_commit_w(&bars[kBYEmpty + yb]);
            if (kPMc)
              umma_commit_mc_w(&bars[kBDzEmpty + ds], uint16_t(3));
            else
              umma_commit_w(&bars[kBDzEmpty + ds]);
            if (lane == 0) TR(6, i);
          }
        } else {
          const uint32_t idesc_w = idesc_bf16(128, 64, false, true);  // A = y^T (K-major), B = dOut (MN-major)
          for (int i = 0; i < n; ++i) {
            const int bs = i % kPDobSlots, yb = i & 1;
            const uint32_t dob = sbase + E::kDob + bs * kPBlk;
            mbar_wait(&bars[kBDobFull + bs], (i / kPDobSlots) & 1);
            PW(0, mbar_wait(&bars[kBYFull + yb], (i >> 1) & 1));
            if (lane == 0) TR(5, i);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kPipeTileRows / 16; ++ks)
              if (kPYTmem)
                umma_bf16_ts_w(t_w, t_y + yb * 32 + ks * 8, smem_desc(dob + ks * 2048, hiMN), idesc_w, (i | ks) != 0);
              else
                umma_bf16_ss_w(t_w, smem_desc(oY + yb * kPHalf + ks * 32, hiK), smem_desc(dob + ks * 2048, hiMN), idesc_w,
                               (i | ks) != 0);
            umma_commit_w(&bars[kBYEmpty + yb]);
            umma_commit_w(&bars[kBDobEmpty + bs]);
            if (lane == 0) TR(6, i);
          }
        }
        umma_commit_w(&bars[kBFin]);
        if (lane == 0) PW_FLUSH(6, 1);
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
          const int ss = i % kPStgSlots;
          // the credit poll (an L2 round trip almost every tile: the consumers run just in time) goes first, under the
          // wait for the epilogue; the previous tile is published before this tile's read-out wait
          if (i >= kPipeRing) PW(1, poll2_ge(fa, fb, uint32_t(i - kPipeRing + 1), c0, c1));  // slot read out
          TR(8, i);
          PW(0, mbar_wait(&bars[kBStgFull + ss], (i / kPStgSlots) & 1));
          TR(7, i);
          bulk_s2g(ring_out + size_t(i % kPipeRing) * kPTile + size_t(h) * kPHalf, smem + oStg + ss * kPHalf,
                   kPHalf);
          bulk_commit();
          if (i >= kPStoreDepth) {  // the store of tile i - kPStoreDepth is complete: publish it (joins the next group)
            PW(3, bulk_wait_group<kPStoreDepth>());
            bulk_red_add_u32x4(fl_out + h * 32, one_s);
            TR(10, i - kPStoreDepth);
          }
          PW(2, bulk_wait_read0());
          mbar_arrive(&bars[kBStgEmpty + ss]);
          TR(9, i);
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
      for (int i = lane; i < kPZSlots * kPBlk / 16; i += 32) sts128(sbase + E::kXb + i * 16, make_uint4(0u, 0u, 0u, 0u));
      __syncwarp();
      uint32_t k0 = 0;
      auto xa_row = [&](int i, int rr) {  // coordinate record of row lane + 32 rr of tile i
        const int T = t_first + (i >> 1);
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
          bulk_g2s(smem + E::kZ + zs * kPHalf, ring_in + size_t(i % kPipeRing) * kPTile + size_t(h) * kPHalf, kPHalf,
                   &bars[kBZFull + zs]);
        }
        sts128(sbase + E::kXb + zs * kPBlk + sw128_chunk_off(lane, 0), xv0);
        sts128(sbase + E::kXb + zs * kPBlk + sw128_chunk_off(lane + 32, 0), xv1);
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
        {
          PW(2, mbar_wait(&bars[kBZFull + zs], (i / kPZSlots) & 1));
          const long long tm0 = prof_on ? clock64() : 0;
          if (lane == 0) st_relaxed_gpu(fl_in + (2 + h) * 32, uint32_t(i + 1));
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kPipeTileRows / 16; ++ks)
            umma_bf16_ss_w(t_w0, smem_desc(sbase + E::kZ + zs * kPHalf + ks * 32, hiK),
                           smem_desc(sbase + E::kXb + zs * kPBlk + ks * 2048, hiMN), idesc_w, (i | ks) != 0);
          umma_commit_w(&bars[kBZEmpty + zs]);
          if (prof_on) pw[4] += (unsigned long long)(clock64() - tm0);
        }
        __syncwarp();
        if (i + kPZSlots - 1 < n) PW(3, issue(i + kPZSlots - 1));
      }
      umma_commit_w(&bars[kBFinB]);
      if (lane == 0) {
        PW_FLUSH(18, 3);
        if (prof_on) {
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 26] = pw[3];
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 27] = pw[4];
        }
      }
    } else if (warp >= kPFirstCvtWarp) {
      // =============================== edge: dOut fp32 -> bf16 converters ===============================
      // Thread t owns row t of every tile: [64 rows][C] fp32 (bulk-copied to shared memory; a ragged last tile comes
      // straight from global memory) -> bf16 [64 rows][64] block, zero beyond C / rows; running column sums = db_f.
      if (edge) {
        const int t = (threadIdx.x - kPFirstCvtWarp * 32) & 63;  // row of the tile
        // with four converter warps two pairs alternate tiles (one pair was the slowest loop of the edge CTA: ~2 k
        // cycles per tile); slot counts are even, so a slot and its barriers always belong to the same pair
        constexpr int kPairs = kPCvtWarps / kPCvtPair;
        static_assert(kPRawSlots % kPairs == 0 && kPDobSlots % kPairs == 0, "slot ownership");
        const int cp = (threadIdx.x - kPFirstCvtWarp * 32) >> 6;
        const int C = p.C;
        float dbf[kOutPad];
#pragma unroll
        for (int c = 0; c < kOutPad; ++c) dbf[c] = 0.f;
        for (int sidx = cp; sidx < kPDobSlots; sidx += kPairs)  // columns 32..63 stay zero for the whole kernel
#pragma unroll
          for (int ch = 4; ch < 8; ++ch)
            sts128(sbase + E::kDob + sidx * kPBlk + sw128_chunk_off(t, ch), make_uint4(0u, 0u, 0u, 0u));
        for (int j = cp; j < n; j += kPairs) {
          const int bs = j % kPDobSlots, rs = j % kPRawSlots;
          if (t == 0) TR(19, j);
          mbar_wait(&bars[kBRawFull + rs], (j / kPRawSlots) & 1);
          if (t == 0) TR(20, j);
          if (j >= kPDobSlots) mbar_wait(&bars[kBDobEmpty + bs], ((j / kPDobSlots) - 1) & 1);
          if (t == 0) TR(21, j);
          const int T = t_first + (j >> 1);
          const long long row0 = (long long)T * 128 + (j & 1) * kPipeTileRows;
          float gv[kOutPad];
          if (kPKo & 32) {  // knock-out: no conversion work (the barriers still cycle)
#pragma unroll
            for (int c = 0; c < kOutPad; ++c) gv[c] = 0.f;
          } else if (row0 + kPipeTileRows <= p.rows) {
            const uint32_t src = sbase + E::kRaw + rs * kPRawSlot + uint32_t(t * C) * 4;
#pragma unroll
            for (int c = 0; c < kOutPad; ++c) gv[c] = (c < C) ? __uint_as_float(lds32(src + c * 4)) : 0.f;
          } else {
            const bool valid = row0 + t < p.rows;
            const float* g = p.grad_out + (row0 + t) * C;
#pragma unroll
            for (int c = 0; c < kOutPad; ++c) gv[c] = (valid && c < C) ? __ldg(g + c) : 0.f;
          }
#pragma unroll
          for (int c = 0; c < kOutPad; ++c) dbf[c] += gv[c];
#pragma unroll
          for (int ch = 0; ch < ((kPKo & 32) ? 0 : 4); ++ch)
            sts128(sbase + E::kDob + bs * kPBlk + sw128_chunk_off(t, ch),
                   make_uint4(pack_bf16x2(gv[8 * ch], gv[8 * ch + 1]), pack_bf16x2(gv[8 * ch + 2], gv[8 * ch + 3]),
                              pack_bf16x2(gv[8 * ch + 4], gv[8 * ch + 5]), pack_bf16x2(gv[8 * ch + 6], gv[8 * ch + 7])));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bars[kBDobFull + bs]);
            mbar_arrive(&bars[kBRawEmpty + rs]);
            if (kPMc) mbar_arrive_remote_relaxed(&bars[kBRawEmpty + rs], uint32_t(h ^ 1));
          }
          if (t == 0) TR(22, j);
          __syncwarp();
        }
        if (h == 0) {  // db_f: column sums of dOut (one edge CTA per pipeline)
#pragma unroll
          for (int c = 0; c < kOutPad; ++c) {
            float sum = dbf[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0 && c < C) atomicAdd(p.grads + p.off[2 * (L + 1) + 1] + c, sum);
          }
        }
      }
    } else if (warp >= kPFirstEpiWarp) {
      // =============================== epilogue warps ===============================
      // Two groups of 8 warps alternate tiles (group g takes tiles i = g mod 2), so the barrier / fence / TMEM latency
      // of one group's tile hides behind the other group's sin / cos work (the MUFU pipe is the floor of this kernel).
      // Thread = feature f (TMEM lane) x rows [32 rh, 32 rh + 32) of its tiles, in two batches of 16 rows.  From ONE
      // phase read per element:
      //   dTheta^T[f][row] = D^T[f][row] * cos(phase) -> staging (feature-major, 16-byte stores) -> ring
      //   y^T[f][row]      = sin(phase)               -> TMEM operand of the weight-gradient MMA
      // The chain accumulator and the phase slot are handed back as soon as both batches are in registers.
      unsigned long long pw[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const long long te0 = prof_on ? clock64() : 0;
      const int ew = warp - kPFirstEpiWarp;
      const int q = ew & 3;          // TMEM lane quadrant (== warp & 3)
      const int cg = ew >> 2;        // flush: 64-output column group
      const int rh = (ew >> 2) & 1;  // row half inside a tile
      const int grp = ew >> 3;       // tile parity this warp works on
      const int f = q * 32 + lane;   // feature inside this CTA's half
      const uint32_t t_lane = uint32_t(q * 32) << 16;
      const int C = p.C;
      float dbsum = 0.f;
      // ReLU-tail network: the top activated layer (whose stash the edge CTAs read) is Linear + ReLU
      const bool relu_top = kRelu && edge;
      // layer 0 not stashed (kPipeSkipPh0): the CTAs of layer 1 recompute theta_0[f][row] = w' x' + b' from the row's
      // coordinate record.  w', b' = bf16 hi + bf16 lo of omega_0 W_0 / omega_0 b_0, the operand pack.cu builds for the
      // forward's first-layer MMA, so that the angle is the forward's (to fp32 rounding).
      const bool l0x = !edge && ph_layer == 0 && skip0;
      float w0x = 0.f, w0y = 0.f, w0z = 0.f, w0w = 0.f, b0x = 0.f;
      if (l0x) {
        auto hilo = [](float v) {
          const float hi = __bfloat162float(__float2bfloat16_rn(v));
          return hi + __bfloat162float(__float2bfloat16_rn(v - hi));
        };
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.packed + p.pl.w0) + (h * 128 + f));
        w0x = hilo(w.x);
        w0y = hilo(w.y);
        w0z = hilo(w.z);
        w0w = hilo(w.w);
        b0x = hilo(__ldg(reinterpret_cast<const float*>(p.packed + p.pl.bias) + (h * 128 + f)));
      }
      // inside an aligned batch of 16 grid rows only the LAST coordinate changes (the forward's condition for skipping)
      const float w0l = p.d == 1 ? w0x : (p.d == 2 ? w0y : (p.d == 3 ? w0z : w0w));
      const uint32_t lastoff = uint32_t(p.d - 1) * 4;

      if (!edge) {
        // ---- W'^T half -> TMEM (A operand of the chain MMA): lane = input feature 128 h + f, 2 bf16 per column along K
        if (cg == 0) {
          const uint8_t* src = p.packed + p.pl.wht + size_t(layer - 1) * 256 * 256 * 2 + size_t(h * 128 + f) * 128;
          // (two K blocks = 16 independent 16-byte loads per round trip to L2: the pipeline cannot start before this
          //  operand is in place, and one block at a time was eight serialised L2 latencies)
#pragma unroll 1
          for (int kb2 = 0; kb2 < 4; kb2 += 2) {
            uint4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[j] = __ldg(reinterpret_cast<const uint4*>(src + size_t(kb2 + (j >> 3)) * 256 * 128 +
                                                          (((j & 7) ^ (f & 7)) << 4)));
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {  // q4 = (K block, half) of this round
              uint32_t w[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                w[4 * c] = v[4 * q4 + c].x;
                w[4 * c + 1] = v[4 * q4 + c].y;
                w[4 * c + 2] = v[4 * q4 + c].z;
                w[4 * c + 3] = v[4 * q4 + c].w;
              }
              tmem_st16(t_wt + t_lane + (kb2 + (q4 >> 1)) * 32 + (q4 & 1) * 16, w);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[kBW]);
        }
      }

      if (grp == 1 && p.skew_ns > 0) __nanosleep(p.skew_ns);  // start the two groups half a cycle apart
      // The tile loop exists twice in the binary -- with and without the recomputed layer-0 angle -- and a CTA runs one
      // of them: as a uniform branch INSIDE the one loop the extra 2.5 KB of code cost every CTA 6 % (instruction-cache
      // stalls 0.22 -> 0.36 per issue; the hot loops of this kernel live on the edge of the instruction cache).
      auto tile_loop = [&](auto l0c) {
      constexpr bool kL0 = decltype(l0c)::value;
      for (int i = 0; i < n; ++i) {
        if (kPGroupPh && (i & 1) != grp) continue;  // per-group phase slots: the other group's tiles are not our business
        const int ps = kPGroupPh ? ((i & 1) * 2 + ((i >> 1) & 1)) : i % nph;
        // shared slots: every warp passes every phase of the phase-slot barriers (a parity wait may not skip a phase)
        PW(1, mbar_wait(&bars[kBPhFull + ps], (kPGroupPh ? (i >> 2) : (i / nph)) & 1));
        if ((i & 1) != grp) {  // the other group's tile: just let the slot go (it is refilled once ALL warps passed)
          if ((ew & 7) == 0 && lane == 0) TR(15, i);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[kBPhEmpty + ps]);
          continue;
        }
        const int yb = grp, u = i >> 1;  // u: use count of this group's buffers
        const int sb = i % kPStgSlots, su = i / kPStgSlots;  // staging slot and its use count
        const bool tr_me = (ew & 7) == 0 && lane == 0;
        if (tr_me) TR(11, i);
        PW(3, mbar_wait(&bars[kBAccFull + grp], u & 1));
        if (u >= 1) {  // this group's previous tile: its y consumed by the MMA, its dTheta half read out by the store
          PW(2, mbar_wait(&bars[kBYEmpty + yb], (u - 1) & 1));
          if (su >= 1) PW(4, mbar_wait(&bars[kBStgEmpty + sb], (su - 1) & 1));
        }
        if (tr_me) TR(12, i);
        tc_fence_after();
        const long long tc0 = prof_on ? clock64() : 0;
        const uint32_t ph_f = sbase + oPh + ps * kPPhSlot + (f >> 3) * kPPhChunk + (f & 7) * 2 + rh * 32 * 16;
        const uint32_t acc = t_acc + t_lane + (kPYTmem ? 0 : grp * 64) + rh * 32;
        const uint32_t yblk = oY + yb * kPHalf + (f >> 6) * kPBlk;
        const uint32_t dblk = sbase + oStg + sb * kPHalf + (f >> 6) * kPBlk;
        // the whole 32-row accumulator slice goes to registers first and is handed back at once: the chain MMA of
        // this group's NEXT tile (same accumulator) then runs under the sin / cos work below
        uint32_t v[32];
        tmem_ld32(acc, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[kBAccEmpty + grp]);
        if (tr_me) TR(13, i);
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          uint32_t ph[16], ys[8], ds[8];
          if (kL0) {  // ph[j] = the bits of the recomputed angle (radians) of row 32 rh + 16 b2 + j:
                      // base (the batch's first row without its last-axis term) + w_last x_last(row)
            const uint32_t xr = sbase + oPh + ps * kPPhSlot + (rh * 32 + 16 * b2) * 16;
            const uint4 x0 = lds128(xr);
            float base = fmaf(w0w, __uint_as_float(x0.w),
                              fmaf(w0z, __uint_as_float(x0.z),
                                   fmaf(w0y, __uint_as_float(x0.y), fmaf(w0x, __uint_as_float(x0.x), b0x))));
            base = fmaf(-w0l, __uint_as_float(lds32(xr + lastoff)), base);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              ph[j] = __float_as_uint(fmaf(w0l, __uint_as_float(lds32(xr + j * 16 + lastoff)), base));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ph[j] = lds16(ph_f + (16 * b2 + j) * 16);
          }
          if (b2 == 1) {  // all phases are in registers: hand the slot back
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[kBPhEmpty + ps]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float r0 = kL0 ? __uint_as_float(ph[2 * j]) : rad_lo16(ph[2 * j]);
            const float r1 = kL0 ? __uint_as_float(ph[2 * j + 1]) : rad_lo16(ph[2 * j + 1]);
            if (kPKo & 1) {
              const float d0 = __uint_as_float(v[16 * b2 + 2 * j]) * r0, d1 = __uint_as_float(v[16 * b2 + 2 * j + 1]) * r1;
              dbsum += d0 + d1;
              ds[j] = pack_bf16x2(d0, d1);
              ys[j] = pack_bf16x2(r0, r1);
              continue;
            }
            if (relu_top) {  // stash slot = bf16 y = relu(theta): dTheta = D where y > 0, the operand of dW_f is y itself
              const uint32_t y0 = ph[2 * j], y1 = ph[2 * j + 1];
              const float d0 = (y0 & 0x7FFFu) ? __uint_as_float(v[16 * b2 + 2 * j]) : 0.f;  // (y >= 0: no sign to test)
              const float d1 = (y1 & 0x7FFFu) ? __uint_as_float(v[16 * b2 + 2 * j + 1]) : 0.f;
              dbsum += d0 + d1;
              ds[j] = pack_bf16x2(d0, d1);
              ys[j] = y0 | (y1 << 16);
              continue;
            }
            const float d0 = __uint_as_float(v[16 * b2 + 2 * j]) * __cosf(r0);
            const float d1 = __uint_as_float(v[16 * b2 + 2 * j + 1]) * __cosf(r1);
            dbsum += d0 + d1;
            ds[j] = pack_bf16x2(d0, d1);
            ys[j] = pack_bf16x2(__sinf(r0), __sinf(r1));
          }
          const uint32_t ch = 4 * rh + 2 * b2;
          if (kPYTmem) {
            tmem_st8(t_y + t_lane + yb * 32 + rh * 16 + b2 * 8, ys);
          } else {
            sts128(yblk + sw128_chunk_off(f & 63, ch), make_uint4(ys[0], ys[1], ys[2], ys[3]));
            sts128(yblk + sw128_chunk_off(f & 63, ch + 1), make_uint4(ys[4], ys[5], ys[6], ys[7]));
          }
          sts128(dblk + sw128_chunk_off(f & 63, ch), make_uint4(ds[0], ds[1], ds[2], ds[3]));
          sts128(dblk + sw128_chunk_off(f & 63, ch + 1), make_uint4(ds[4], ds[5], ds[6], ds[7]));
        }
        if (tr_me) TR(14, i);
        if (prof_on) pw[6] += (unsigned long long)(clock64() - tc0);
        const long long ts0 = prof_on ? clock64() : 0;
        if (kPYTmem) {
          tmem_st_wait();
          tc_fence_before();
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[kBYFull + yb]);
          mbar_arrive(&bars[kBStgFull + sb]);
        }
        if (tr_me) TR(16, i);
        if (prof_on) pw[8] += (unsigned long long)(clock64() - ts0);
      }
      };
#ifndef B200INR_PL0OFF
#define B200INR_PL0OFF 0
#endif
      if (l0x && !B200INR_PL0OFF)  // (tuning: PL0OFF = 1 times the kernel with the stashed-phase arithmetic everywhere)
        tile_loop(std::true_type{});
      else
        tile_loop(std::false_type{});
      if (end_on && threadIdx.x == kPFirstEpiWarp * 32)  // tile loop done, flush not started
        p.prof[size_t(blockIdx.x) * kPipeProfSlots + 1] = (unsigned long long)(clock64() - t_begin);
      if (threadIdx.x == kPFirstEpiWarp * 32) {
        PW_FLUSH(11, 7);
        if (prof_on) {
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 29] = (unsigned long long)(clock64() - te0);
          p.prof[size_t(blockIdx.x) * kPipeProfSlots + 30] = pw[8];
        }
      }

      // ---- flush.  The operands are kSirenWidth wide; a narrower network (Hr < 256) has zero rows / columns there and
      // its gradient buffers are Hr wide, so the padded features are skipped (after the warp-collective TMEM loads).
      const int Hr = p.Hr;
      const int fi = h * 128 + f;  // feature owned by this thread
      // bias gradient of the layer whose dTheta this CTA produced
      {
        const float sc = (ph_layer == 0) ? p.omega0 : (relu_top ? 1.0f : p.omegah);
        if (fi < Hr) atomicAdd(p.grads + p.off[2 * ph_layer + 1] + fi, sc * dbsum);
      }
      mbar_wait(&bars[kBFin], 0);
      tc_fence_after();
      if (!edge) {
        // dW_l[out][128 h + f] += omega_h * D[f][out]; this warp: outputs 64 cg .. 64 cg + 64 (lanes = consecutive inputs)
        float* dst = p.grads + p.off[2 * layer] + fi;
        const float wsc = (kRelu && layer == L) ? 1.0f : p.omegah;  // (the Linear + ReLU layer has no omega)
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_w + t_lane + cg * 64 + c0, v);
          tmem_ld_wait();
          if (fi < Hr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cg * 64 + c0 + j < Hr)
                atomicAdd(dst + (long long)(cg * 64 + c0 + j) * Hr, wsc * __uint_as_float(v[j]));
          }
        }
      } else {
        // dW_f[c][128 h + f] += D[f][c]
        {
          uint32_t v[16];
          tmem_ld16(t_w + t_lane + cg * 16, v);
          tmem_ld_wait();
          float* dst = p.grads + p.off[2 * (L + 1)] + fi;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = cg * 16 + j;
            if (c < C && fi < Hr) atomicAdd(dst + (long long)c * Hr, __uint_as_float(v[j]));
          }
        }
        // dW_0[128 h + f][j] += omega_0 * (D[f][j] + D[f][4 + j])   (x = hi + lo)
        mbar_wait(&bars[kBFinB], 0);
        tc_fence_after();
        if (cg == 0) {
          uint32_t v[16];
          tmem_ld16(t_w0 + t_lane, v);
          tmem_ld_wait();
          float* dst = p.grads + p.off[0] + (long long)fi * p.d;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < p.d && fi < Hr) atomicAdd(dst + j, p.omega0 * (__uint_as_float(v[j]) + __uint_as_float(v[4 + j])));
        }
      }
      tc_fence_before();
    }
  }

  __syncthreads();
  if (kPMc) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 1) tmem_dealloc<512>(tmem);
  if (kPAdapt && p.calib != nullptr && role == 0 && threadIdx.x == 0 && p.pipelines <= 32) {
    // this pipeline's record for the next launch (the edge CTA holds both ends of the chain: it is the last to finish)
    const uint32_t epoch = share_s[2] + 1u;
    const long long cyc = clock64() - t_launch;
    __stcg(reinterpret_cast<uint4*>(p.calib + 16 + (epoch & 1u) * kCalBankWords) + pipe,
           make_uint4(uint32_t(my_fwd), uint32_t(cyc > 0xffffffffLL ? 0xffffffffLL : cyc), epoch,
                      kCalMagic ^ uint32_t(p.pipelines)));
    if (pipe == 0) __stcg(p.calib, epoch);
  }
  if (end_on && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 3] = gt;
  }
  if ((prof_on || end_on) && threadIdx.x == 0) {
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 0] = (unsigned long long)(clock64() - t_begin);
    p.prof[size_t(blockIdx.x) * kPipeProfSlots + 21] = (unsigned long long)n;
  }
}

int launch_siren_bwdp(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                      const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params, int num_sms,
                      cudaStream_t stream) {
  constexpr int H = kSirenWidth;  // operand width; the parameters are net->hidden_features wide (zero-padded operands)
  const int L = net->hidden_layers;
  const PipeStashLayout sl = make_pipe_stash_layout(H, L, rows);
  PipeParams p{};
  p.Hr = net->hidden_features;
  p.relu_tail = (net->flags & B200INR_NET_RELU_TAIL) != 0;
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
  p.calib = reinterpret_cast<uint32_t*>(st + sl.prof + size_t(kCalRow) * kPipeProfSlots * 8);
  static_assert((kCalRow + 16) <= kPipeProfCtas && (16 + 2 * kCalBankWords) * 4 <= 16 * kPipeProfSlots * 8, "calibration area");
  p.skip_ph0 = kPipeSkipPh0 ? 1 : 0;  // (and the forward's word in the stash says whether it did skip them)
  p.grads = grad_params;
  int64_t off[2 * (kMaxSineLayers + 2)];
  param_offsets(p.d, p.Hr, L, p.C, off);
  for (int i = 0; i < 2 * (L + 2); ++i) p.off[i] = off[i];
  p.omega0 = net->first_omega_0;
  p.omegah = net->hidden_omega_0;
  const int S2 = 2 * (L + 1);
  const int smem = PSmem::kBytes + PSmem::kSlack;
  // Every CTA of the grid waits on flags written by other CTAs of its pipeline, so the whole grid must be resident at
  // once.  That is not assumed: the launch is COOPERATIVE (the driver refuses it -- a clean error code, no hang and no
  // trap -- when the grid does not fit, e.g. with SMs taken by another context, MPS partition or debugger), and the
  // number of pipelines comes from the occupancy the driver reports for this kernel on this device (cached).
  static int resident_ctas[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B200INR_ERR_CUDA;
  const bool instr = false;
  auto kern = p.relu_tail ? siren_bwdp_kernel<false, true> : siren_bwdp_kernel<false, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  if (dev < 0 || dev >= 64 || resident_ctas[dev] == 0) {
    int per_sm = 0;
    if (kPMc) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(unsigned(num_sms & ~1));
      cfg.blockDim = dim3(kPThreads);
      cfg.dynamicSmemBytes = size_t(smem);
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = 2;
      at.val.clusterDim.y = 1;
      at.val.clusterDim.z = 1;
      cfg.attrs = &at;
      cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) return B200INR_ERR_CUDA;
      per_sm = -2 * nc;  // (negative: an absolute CTA count)
    } else if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPThreads, size_t(smem)) != cudaSuccess) {
      return B200INR_ERR_CUDA;
    }
    const int total = per_sm < 0 ? -per_sm : per_sm * num_sms;
    if (total < S2) return B200INR_ERR_CUDA;
    if (dev >= 0 && dev < 64) resident_ctas[dev] = total;
  }
  const int resident = (dev >= 0 && dev < 64) ? resident_ctas[dev] : num_sms;
  int P = (resident < num_sms ? resident : num_sms) / S2;  // one CTA per SM: the kernel is sized for a whole SM
  if (P > p.fwd_tiles) P = p.fwd_tiles;
  if (P < 1 || P * (L + 1) > kPipeMaxEdges) return B200INR_ERR_BAD_SHAPE;
  p.pipelines = P;
#if B200INR_TUNING
  // tuning builds only (tools/build_variant.sh ... -DB200INR_TUNING=1): event trace / stall counters / start skew
  const char* env_skew = getenv("B200INR_BWDP_SKEW_NS");
  p.skew_ns = env_skew != nullptr ? atoi(env_skew) : 0;
  const char* env_dbg = getenv("B200INR_BWDP_DBG");
  p.dbg = env_dbg != nullptr ? atoi(env_dbg) : 0;
  const char* env_trace = getenv("B200INR_BWDP_TRACE_PTR");
  p.trace = env_trace != nullptr ? reinterpret_cast<uint32_t*>(strtoull(env_trace, nullptr, 0)) : nullptr;
  const char* env_prof = getenv("B200INR_BWDP_PROF");
  if (env_prof != nullptr && env_prof[0] == '1' && P * S2 <= kPipeProfCtas)
    p.prof = reinterpret_cast<unsigned long long*>(st + sl.prof);
#endif
  if (cudaMemsetAsync(p.flags, 0, sl.flags_bytes, stream) != cudaSuccess) return B200INR_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(P * S2));
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = size_t(smem);
  cfg.stream = stream;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeCooperative;
  at.val.cooperative = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  cudaError_t e;
#if B200INR_TUNING
  if ((p.prof != nullptr && !kPEndOnly) || (p.trace != nullptr && !kPTraceOnly) || p.dbg != 0) {
    if (cudaFuncSetAttribute(siren_bwdp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B200INR_ERR_CUDA;
    e = cudaLaunchKernelEx(&cfg, siren_bwdp_kernel<true>, p);
  } else
#endif
  {
    (void)instr;
    e = cudaLaunchKernelEx(&cfg, kern, p);
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return B200INR_ERR_CUDA;
  }
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
