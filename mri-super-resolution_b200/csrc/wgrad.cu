// wgrad.cu -- weight and bias gradients of the coordinate MLPs, contracted over the row (coordinate) dimension.
//
// Replaces the dW / db half of loss.backward() (autograd of nn.Linear inside INR/SRDWI.py:58-59):
//     dW_l = omega_l * dTheta_l^T Y_{l-1}      db_l = omega_l * colsum(dTheta_l)        l = 1 .. L
//     dW_f = dOut^T Y_L                        db_f = colsum(dOut)
//     dW_0 = omega_0 * dTheta_0^T X            db_0 = omega_0 * colsum(dTheta_0)
// The stash tiles written by the forward / dgrad kernels are [rows][64] SWIZZLE_128B blocks; read with MN-major
// descriptors the same bytes are a K = rows operand, so no transposition happens anywhere.  A work item is one
// 256 x 256 block of one layer's dW (the whole matrix at H = 256, a quarter at H = 512); items are split over the
// row range across CTAs in proportion to the bytes they stream; every CTA keeps its partial block in TMEM for its
// whole row range (2 x [128 lanes x N] fp32 = all 512 columns) and flushes once with vector red.global.add.
// Items of one layer walk the rows in step, so the tiles two items share are served from L2 the second time.
//
// SIREN on raw coordinates (K = d <= 4 inputs): the B operand of the first layer is the [rows][64] block written by
// the training forward whose first eight columns hold the coordinates split into bf16 hi and lo parts
// (x = hi + lo to 2^-17), so dW_0 = D[:, j] + D[:, 4 + j].
//
// Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = bias column sums on
// CUDA cores (fp32) while the tiles stream through, then the TMEM flush.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

constexpr int kWgThreads = 192;
constexpr int kWgAuxThreads = 128;
constexpr int kWgStages = 6;
constexpr int kWgStageRows = 32;                       // K rows per stage
constexpr int kWgBlkBytes = kWgStageRows * 128;        // 4 KB: 32 rows of one [128][64] block
constexpr int kWgStageBytes = 8 * kWgBlkBytes;         // 4 A blocks + up to 4 B blocks
constexpr int kWgStagesPerTile = kTileRows / kWgStageRows;
constexpr int kWgMaxItems = 24;

enum WgOut : int {
  kWgOutBlock = 0,  // gw[(row_off + m) * ldw + col_off + n] += D[m][n], N = 64 * nb       (activated layers)
  kWgOutFinal = 1,  // gw[n * ldw + row_off + m] += D[m][n] for n < C, N = 64               (final linear, transposed)
  kWgOutCoord = 2   // gw[m * d + j] += D[m][j] + D[m][4 + j], N = 64                       (SIREN first layer)
};

struct WgItem {
  int out;                 // WgOut
  int cta_begin, cta_count;  // blocks cta_begin + k * cta_stride, k < cta_count, work on this item (row split k)
  int cta_stride;            // the items of a group are interleaved: blocks that stream the same tiles are neighbours
  const uint8_t* a_src;    // M operand tiles: a_tile_bytes per tile, 4 blocks used starting at block a_blk0
  const uint8_t* b_src;    // N operand tiles: b_tile_bytes per tile, nb blocks used starting at block b_blk0
  uint32_t a_tile_bytes, a_blk0, b_tile_bytes, b_blk0;
  int nb;                  // B blocks per stage (1..4); the MMA N is 64 * nb
  float* gw;
  int ldw, row_off, col_off;
  float* gb;               // bias gradient of the summed features, or nullptr
  int sum_b;               // 0: column sums over the A blocks (dTheta), 1: over the B block (dOut)
  int gb_count;            // number of valid bias entries (H features or C channels)
  int m_valid, n_valid;    // valid M rows / N columns of the block when the network is narrower than the operands
                           // (0 = all)
  float scale;
};

struct WgParams {
  WgItem items[kWgMaxItems];
  int num_items;
  int num_tiles;
  int d, C;
};

struct WgSmem {
  static constexpr int kOffStage = 0;
  static constexpr int kOffBar = kWgStages * kWgStageBytes;
  static constexpr int kBytes = kOffBar + 256;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const WgParams p) {
  using S = WgSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* full = bars;                // [kWgStages]
  uint64_t* empty = bars + kWgStages;   // [kWgStages]
  uint64_t* d_full = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which item, which tile range
  int it = 0;
  for (int i = 0; i < p.num_items; ++i) {
    const int rel = int(blockIdx.x) - p.items[i].cta_begin;
    if (rel >= 0 && rel % p.items[i].cta_stride == 0 && rel / p.items[i].cta_stride < p.items[i].cta_count) it = i;
  }
  const WgItem item = p.items[it];
  const int split = (int(blockIdx.x) - item.cta_begin) / item.cta_stride;
  const int tile_begin = int((long long)p.num_tiles * split / item.cta_count);
  const int tile_end = int((long long)p.num_tiles * (split + 1) / item.cta_count);
  const int num_stages = (tile_end - tile_begin) * kWgStagesPerTile;
  const int nb = item.nb;          // B blocks per stage
  const int ncols = 64 * item.nb;  // N of the MMA == TMEM columns per M half

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1 + kWgAuxThreads / 32);
    }
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (num_stages > 0) {
    if (warp == 0) {
      // =============================== producer ===============================
      if (lane == 0) {
        const uint32_t bytes = uint32_t(4 + nb) * kWgBlkBytes;
        for (int s = 0; s < num_stages; ++s) {
          const int slot = s % kWgStages;
          const int round = s / kWgStages;
          const int tile = tile_begin + s / kWgStagesPerTile;
          const int sub = s % kWgStagesPerTile;
          if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
          mbar_arrive_expect_tx(&full[slot], bytes);
          uint8_t* dst = smem + S::kOffStage + slot * kWgStageBytes;
          const uint8_t* a = item.a_src + size_t(tile) * item.a_tile_bytes + size_t(item.a_blk0) * (kTileRows * 128) +
                             size_t(sub) * kWgBlkBytes;
          for (int b = 0; b < 4; ++b)
            bulk_g2s(dst + b * kWgBlkBytes, a + size_t(b) * (kTileRows * 128), kWgBlkBytes, &full[slot]);
          const uint8_t* bs = item.b_src + size_t(tile) * item.b_tile_bytes + size_t(item.b_blk0) * (kTileRows * 128) +
                              size_t(sub) * kWgBlkBytes;
          for (int b = 0; b < nb; ++b)
            bulk_g2s(dst + (4 + b) * kWgBlkBytes, bs + size_t(b) * (kTileRows * 128), kWgBlkBytes, &full[slot]);
        }
      }
    } else if (warp == 1) {
      // =============================== MMA issuer ===============================
      {  // whole warp converged, one elected lane issues (umma_*_w)
        // MN-major operands: LBO = stride between 64-wide feature blocks, SBO = 8-row groups along K.
        const uint64_t hi = smem_desc_hi_sw128(kWgBlkBytes, 1024);
        const uint32_t idesc = idesc_bf16(128, ncols, true, true);
        for (int s = 0; s < num_stages; ++s) {
          const int slot = s % kWgStages;
          mbar_wait(&full[slot], (s / kWgStages) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + S::kOffStage + slot * kWgStageBytes);
#pragma unroll
          for (int ks = 0; ks < kWgStageRows / 16; ++ks) {
#pragma unroll
            for (int mh = 0; mh < 2; ++mh) {
              const uint64_t da = smem_desc(base + (2 * mh) * kWgBlkBytes + ks * 2048, hi);
              const uint64_t db = smem_desc(base + 4 * kWgBlkBytes + ks * 2048, hi);
              umma_bf16_ss_w(tmem_d + mh * ncols, da, db, idesc, (s | ks) != 0);
            }
          }
          umma_commit_w(&empty[slot]);
        }
        umma_commit_w(d_full);
      }
    } else {
      // =============================== column sums + flush ===============================
      const int at = threadIdx.x - 64;  // 0..127
      const int q = warp & 3;           // TMEM lane quadrant of this warp
      // column pair owned by this thread inside the summed operand: features 2*at, 2*at + 1
      const int sum_blocks = item.sum_b ? 1 : 4;
      const uint32_t sum_off = item.sum_b ? 4 * kWgBlkBytes : 0;
      const bool sums = item.gb != nullptr && (2 * at) < sum_blocks * 64;
      const int sblk = (2 * at) >> 6, sch = ((2 * at) & 63) >> 3, sel = (2 * at) & 7;
      float s0 = 0.f, s1 = 0.f;
      for (int s = 0; s < num_stages; ++s) {
        const int slot = s % kWgStages;
        mbar_wait(&full[slot], (s / kWgStages) & 1);
        if (sums) {
          const uint32_t blk = smem_u32(smem + S::kOffStage + slot * kWgStageBytes) + sum_off + sblk * kWgBlkBytes;
#pragma unroll 8
          for (int rr = 0; rr < kWgStageRows; ++rr) {
            const uint32_t v = lds32(blk + sw128_chunk_off(rr, sch) + sel * 2);
            s0 += bf16lo(v);
            s1 += bf16hi(v);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
      // ---- bias gradients
      if (sums) {
        const int c0 = 2 * at;
        if (c0 < item.gb_count) atomicAdd(item.gb + c0, item.scale * s0);
        if (c0 + 1 < item.gb_count) atomicAdd(item.gb + c0 + 1, item.scale * s1);
      }
      // ---- flush the TMEM partial products
      {
        mbar_wait(d_full, 0);
        tc_fence_after();
        const uint32_t t_lane = uint32_t(q * 32) << 16;
        const int m_valid = item.m_valid ? item.m_valid : 0x7fffffff;
        const int n_valid = item.n_valid ? item.n_valid : 0x7fffffff;
#pragma unroll 1
        for (int mh = 0; mh < 2; ++mh) {
          const int m = mh * 128 + q * 32 + lane;  // M index inside the item
#pragma unroll 1
          for (int c0 = 0; c0 < ncols; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_d + t_lane + mh * ncols + c0, v);
            tmem_ld_wait();
            if (m >= m_valid) {
              // padded feature row: nothing to add (the loads above are warp-collective, so no early exit)
            } else if (item.out == kWgOutBlock) {
              float* dst = item.gw + (long long)(item.row_off + m) * item.ldw + item.col_off + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (c0 + j < n_valid)
                  red_add_v4(dst + j, item.scale * __uint_as_float(v[j]), item.scale * __uint_as_float(v[j + 1]),
                             item.scale * __uint_as_float(v[j + 2]), item.scale * __uint_as_float(v[j + 3]));
            } else if (item.out == kWgOutFinal) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;  // N index: output channel
                if (c < p.C)
                  atomicAdd(item.gw + (long long)c * item.ldw + item.row_off + m, item.scale * __uint_as_float(v[j]));
              }
            } else if (c0 == 0) {  // coordinates: columns 0..3 = sum dTheta*x_hi, 4..7 = sum dTheta*x_lo
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < p.d)
                  atomicAdd(item.gw + (long long)m * p.d + j,
                            item.scale * (__uint_as_float(v[j]) + __uint_as_float(v[4 + j])));
            }
          }
        }
        tc_fence_before();
      }
    }
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_d);
}

// Distribute CTAs over the items in proportion to `weight` (bytes streamed per tile) and launch.
// group[i]: items that stream the SAME stash tiles (the blocks of one layer's dW) carry the same group id; they get the
// same CTA count, hence the same row splits, and their blocks are interleaved (neighbouring blocks stream the same
// tiles): cfg4 reads 13.8 GB from DRAM instead of 15.8 GB for 9.8 GB of unique tiles.  Measured and NOT kept: holding a
// group's producers within a few tiles of each other through global progress counters brings DRAM reads down to the
// unique 9.8 GB, but every check costs the producer ~4 us under load (the L2 fabric is the busy resource: 18.5 GB
// cross it either way) and the kernel got slower (2.6 -> 3.7-7.4 ms).  Also measured and NOT kept: clusters of four CTAs
// (one row split of a layer's four blocks) in which every CTA bulk-copies half of each shared operand and multicasts
// it to its partner, consumers releasing a slot on the `empty` barriers of all CTAs that copy into it -- bit-identical
// results, but 5.0 ms on cfg4 (B pairs only: 3.6 ms; the cluster launch alone, without sharing: 3.1 ms, fewer
// co-resident CTAs): with 32-row stages the cross-CTA release -> copy -> full round trip exceeds the six-stage ring.
#ifndef B200INR_WG_WAVES
#define B200INR_WG_WAVES 3
#endif
static int launch_items(WgParams& p, const double* weight, const int* group, int num_sms, cudaStream_t stream) {
  // kWaves > 1: the row range of every item is cut kWaves times finer than the SMs need, and the hardware's block
  // scheduler hands the next block to whichever SM finishes first (one CTA fits per SM: all 512 TMEM columns).
  // With one block per SM the SMs were idle for 18 % of the kernel (ncu: active / elapsed cycles 0.82 -- blocks of
  // different items, and of the same item on different SMs, do not run equally fast); 1 / 2 / 3 / 4 waves:
  // cfg4 2.61 / 2.50 / 2.37 / 2.38 ms, WIRE 1.36 / 1.25 / 1.19 / 1.20 ms (at 3 the kernel reads its 13.8 GB at 89 % of
  // the HBM peak).
  constexpr int kWaves = B200INR_WG_WAVES;
  num_sms *= kWaves;
  double total = 0.0;
  for (int i = 0; i < p.num_items; ++i) total += weight[i];
  int count[kWgMaxItems];
  int cta = 0;
  for (int i = 0; i < p.num_items; ++i) {
    int n = int(num_sms * weight[i] / total);
    count[i] = n < 1 ? 1 : n;
    cta += count[i];
  }
  // leftovers: one more CTA for every member of a group, heaviest groups first, while whole groups still fit
  for (bool progress = true; progress && cta < num_sms;) {
    progress = false;
    for (int i = 0; i < p.num_items && cta < num_sms; ++i) {
      if (i > 0 && group[i] == group[i - 1]) continue;  // group leader only
      int members = 0;
      for (int j = i; j < p.num_items && group[j] == group[i]; ++j) ++members;
      if (weight[i] < 100.0 || cta + members > num_sms) continue;
      for (int j = i; j < i + members; ++j) count[j] += 1;
      cta += members;
      progress = true;
    }
  }
  if (cta > num_sms) {  // more items than SMs can take one CTA each: the grid still launches (CTAs queue)
  }
  int begin = 0;
  for (int i = 0; i < p.num_items;) {  // one group at a time: its members' blocks interleaved
    int members = 1;
    while (i + members < p.num_items && group[i + members] == group[i]) ++members;
    for (int j = 0; j < members; ++j) {
      p.items[i + j].cta_begin = begin + j;
      p.items[i + j].cta_stride = members;
      p.items[i + j].cta_count = count[i];
    }
    begin += members * count[i];
    i += members;
  }
  const int smem = WgSmem::kBytes + 1024;
  if (cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  wgrad_kernel<<<begin, kWgThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

int launch_siren_wgrad(const b200inr_net* net, void* stash, const float* coords, const b200inr_grid* grid,
                       int64_t rows, float* grad_params, int num_sms, cudaStream_t stream) {
  (void)coords;
  (void)grid;  // the coordinate operand of the first layer comes from the stash (written by the training forward)
  constexpr int H = kSirenWidth;           // operand width
  const int Hr = net->hidden_features;     // parameter width (<= H; the extra operand features are zero)
  const int L = net->hidden_layers, d = net->in_features, C = net->out_features;
  const StashLayout sl = make_stash_layout(H, L, rows);
  const uint8_t* st = reinterpret_cast<const uint8_t*>(stash);
  int64_t off[2 * (kMaxSineLayers + 1)];
  param_offsets(d, Hr, L, C, off);

  WgParams p{};
  p.num_tiles = int(sl.tiles);
  p.d = d;
  p.C = C;
  double weight[kWgMaxItems];
  int group[kWgMaxItems];
  int ni = 0;
  const uint32_t tile_h = uint32_t(sl.tile_bytes), tile_s = kTileRows * 128;
  for (int l = 1; l <= L; ++l) {
    group[ni] = ni;
    WgItem& w = p.items[ni];
    w = WgItem{};
    w.out = kWgOutBlock;
    w.nb = 4;
    w.a_src = st + sl.dz + size_t(l) * sl.layer_stride;
    w.a_tile_bytes = tile_h;
    w.b_src = st + sl.y + size_t(l - 1) * sl.layer_stride;
    w.b_tile_bytes = tile_h;
    w.gw = grad_params + off[2 * l];
    w.ldw = Hr;
    w.m_valid = w.n_valid = Hr;
    w.gb = grad_params + off[2 * l + 1];
    w.gb_count = Hr;
    w.scale = net->hidden_omega_0;
    weight[ni++] = 128.0;
  }
  {
    WgItem& w = p.items[ni];
    w = WgItem{};
    w.out = kWgOutFinal;
    w.nb = 1;
    w.a_src = st + sl.y + size_t(L) * sl.layer_stride;
    w.a_tile_bytes = tile_h;
    w.b_src = st + sl.dzo;
    w.b_tile_bytes = tile_s;
    w.gw = grad_params + off[2 * (L + 1)];
    w.ldw = Hr;
    w.m_valid = Hr;
    w.gb = grad_params + off[2 * (L + 1) + 1];
    w.sum_b = 1;
    w.gb_count = C;
    w.scale = 1.0f;
    group[ni] = ni;
    weight[ni++] = 80.0;
  }
  {
    group[ni] = ni;
    WgItem& w = p.items[ni];
    w = WgItem{};
    w.out = kWgOutCoord;
    w.nb = 1;
    w.a_src = st + sl.dz;
    w.a_tile_bytes = tile_h;
    w.b_src = st + sl.xa;
    w.b_tile_bytes = tile_s;
    w.gw = grad_params + off[0];
    w.m_valid = Hr;
    w.gb = grad_params + off[1];
    w.gb_count = Hr;
    w.scale = net->first_omega_0;
    weight[ni++] = 80.0;
  }
  p.num_items = ni;
  return launch_items(p, weight, group, num_sms, stream);
}

int launch_gen_wgrad(const b200inr_net* net, void* stash, int64_t rows, float* grad_params, int num_sms,
                     cudaStream_t stream) {
  const GenDims g = make_gen_dims(net);
  const GenStashLayout sl = make_gen_stash_layout(g, rows);
  const uint8_t* st = reinterpret_cast<const uint8_t*>(stash);
  int64_t off[2 * (kMaxSineLayers + 2) + 1];
  gen_param_offsets(g, off);

  WgParams p{};
  p.num_tiles = int(sl.tiles);
  p.d = 0;
  p.C = g.C;
  double weight[kWgMaxItems];
  int group[kWgMaxItems];
  int ni = 0;
  const uint32_t tile_s = kTileRows * 128;
  for (int l = 0; l <= g.L; ++l) {
    const int K = (l == 0) ? g.K0 : g.H;
    for (int mh = 0; mh < g.H / 256; ++mh) {
      for (int b0 = 0; b0 < K / 64; b0 += 4) {
        if (ni >= kWgMaxItems - 2) return B200INR_ERR_BAD_SHAPE;
        group[ni] = l;
        WgItem& w = p.items[ni];
        w = WgItem{};
        w.out = kWgOutBlock;
        w.nb = (K / 64 - b0) < 4 ? (K / 64 - b0) : 4;
        w.a_src = st + sl.dz + size_t(l) * sl.layer_stride;
        w.a_tile_bytes = uint32_t(sl.tile_h);
        w.a_blk0 = 4 * mh;
        w.b_src = (l == 0) ? st + sl.ain : st + sl.y + size_t(l - 1) * sl.layer_stride;
        w.b_tile_bytes = (l == 0) ? uint32_t(sl.tile_in) : uint32_t(sl.tile_h);
        w.b_blk0 = b0;
        w.gw = grad_params + off[2 * l];
        w.ldw = K;
        w.row_off = 256 * mh;
        w.col_off = 64 * b0;
        w.gb = (b0 == 0) ? grad_params + off[2 * l + 1] + 256 * mh : nullptr;
        w.gb_count = 256;
        w.scale = (l == 0) ? g.omega0 : g.omegah;
        // (one weight per layer, so that the whole group gets equal CTA counts: a partial last column group streams
        // fewer bytes but keeps the layer's row splits)
        weight[ni++] = 128.0;
      }
    }
  }
  for (int mh = 0; mh < g.H / 256; ++mh) {
    group[ni] = g.L + 1;
    WgItem& w = p.items[ni];
    w = WgItem{};
    w.out = kWgOutFinal;
    w.nb = 1;
    w.a_src = st + sl.y + size_t(g.L) * sl.layer_stride;
    w.a_tile_bytes = uint32_t(sl.tile_h);
    w.a_blk0 = 4 * mh;
    w.b_src = st + sl.dzo;
    w.b_tile_bytes = tile_s;
    w.gw = grad_params + off[2 * (g.L + 1)];
    w.ldw = g.H;
    w.row_off = 256 * mh;
    w.gb = (mh == 0) ? grad_params + off[2 * (g.L + 1) + 1] : nullptr;
    w.sum_b = 1;
    w.gb_count = g.C;
    w.scale = 1.0f;
    weight[ni++] = 80.0;
  }
  p.num_items = ni;
  return launch_items(p, weight, group, num_sms, stream);
}

// WIRE: the contraction produces the gradient of the real-block matrices into the fp32 scratch of the stash
// (zeroed here); wire.cu's combine kernel folds them into the complex parameters afterwards.
int launch_wire_wgrad(const b200inr_net* net, void* stash, int64_t rows, int num_sms, cudaStream_t stream) {
  const WireDims w = make_wire_dims(net);
  const WireStashLayout sl = make_wire_stash_layout(w, rows);
  uint8_t* st = reinterpret_cast<uint8_t*>(stash);
  float* g = reinterpret_cast<float*>(st + sl.gblk);
  if (cudaMemsetAsync(g, 0, sl.gblk_bytes, stream) != cudaSuccess) return B200INR_ERR_CUDA;

  WgParams p{};
  p.num_tiles = int(sl.tiles);
  p.d = w.d;
  p.C = w.C;
  double weight[kWgMaxItems];
  int group[kWgMaxItems];
  int ni = 0;
  const uint32_t tile_s = kTileRows * 128;
  for (int l = 0; l <= w.L; ++l) {
    for (int mh = 0; mh < 2; ++mh) {
      // feature-fed first layer: K0 columns in groups of up to 4 blocks of the stashed network input
      const int ngroups = (l == 0 && w.K0) ? (w.kb0() + 3) / 4 : 1;
      for (int gq = 0; gq < ngroups; ++gq) {
        if (ni >= kWgMaxItems - 1) return B200INR_ERR_BAD_SHAPE;
        group[ni] = l;
        WgItem& it = p.items[ni];
        it = WgItem{};
        it.a_src = st + sl.dz + size_t(l) * sl.stride_z;
        it.a_tile_bytes = uint32_t(sl.tile_z);
        it.a_blk0 = 4 * mh;
        it.gb = (gq == 0) ? g + wire_gblk_b(w, l) + 256 * mh : nullptr;
        it.gb_count = 256;
        it.scale = 1.0f;
        if (l == 0 && w.K0) {
          it.out = kWgOutBlock;
          it.nb = (w.kb0() - 4 * gq) < 4 ? (w.kb0() - 4 * gq) : 4;
          it.b_src = st + sl.ain;
          it.b_tile_bytes = uint32_t(sl.tile_in);
          it.b_blk0 = 4 * gq;
          it.gw = g + wire_gblk_w(w, 0);
          it.ldw = w.K0;
          it.row_off = 256 * mh;
          it.col_off = 256 * gq;
          weight[ni++] = 128.0;
        } else if (l == 0) {
          it.out = kWgOutCoord;
          it.nb = 1;
          it.b_src = st + sl.xa;
          it.b_tile_bytes = tile_s;
          it.gw = g + wire_gblk_w(w, 0) + size_t(256) * mh * w.d;
          weight[ni++] = 80.0;
        } else {
          it.out = kWgOutBlock;
          it.nb = 4;
          it.b_src = st + sl.y + size_t(l - 1) * sl.stride_y;
          it.b_tile_bytes = uint32_t(sl.tile_y);
          it.gw = g + wire_gblk_w(w, l);
          it.ldw = 2 * w.H;
          it.row_off = 256 * mh;
          weight[ni++] = 128.0;
        }
      }
    }
  }
  {
    group[ni] = w.L + 1;
    WgItem& it = p.items[ni];
    it = WgItem{};
    it.out = kWgOutFinal;
    it.nb = 1;
    it.a_src = st + sl.y + size_t(w.L) * sl.stride_y;
    it.a_tile_bytes = uint32_t(sl.tile_y);
    it.b_src = st + sl.dzo;
    it.b_tile_bytes = tile_s;
    it.gw = g + wire_gblk_wf(w);
    it.ldw = 2 * w.H;
    it.gb = g + wire_gblk_bf(w);
    it.sum_b = 1;
    it.gb_count = w.C;
    it.scale = 1.0f;
    weight[ni++] = 80.0;
  }
  p.num_items = ni;
  return launch_items(p, weight, group, num_sms, stream);
}

}  // namespace b200inr
