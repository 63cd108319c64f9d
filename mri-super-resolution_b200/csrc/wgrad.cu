// wgrad.cu -- weight and bias gradients of the SIREN, contracted over the row (coordinate) dimension.
//
// Replaces the dW / db half of loss.backward() (autograd of nn.Linear inside INR/SRDWI.py:58-59):
//     dW_l = omega_l * dTheta_l^T Y_{l-1}      db_l = omega_l * colsum(dTheta_l)        l = 1 .. L
//     dW_f = dOut^T Y_L                        db_f = colsum(dOut)
//     dW_0 = omega_0 * dTheta_0^T X            db_0 = omega_0 * colsum(dTheta_0)
// The stash tiles written by mlp_fwd.cu / mlp_bwd.cu are [rows][64] SWIZZLE_128B blocks; read with MN-major
// descriptors the same bytes are a K = rows operand, so no transposition happens anywhere.  Work items
// (one per layer) are split over the row range across CTAs; every CTA keeps its partial dW in TMEM for its whole
// row range (2 x [128 lanes x N] fp32) and flushes once with vector red.global.add.
//
// The first layer (K = d <= 4 inputs) uses the same machinery: its B operand is the [rows][64] block written by the
// training forward whose first eight columns hold the coordinates split into bf16 hi and lo parts
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
constexpr uint32_t kWgAuxBarId = 1;
constexpr int kWgStages = 6;
constexpr int kWgStageRows = 32;                       // K rows per stage
constexpr int kWgBlkBytes = kWgStageRows * 128;        // 4 KB: 32 rows of one [128][64] block
constexpr int kWgStageBytes = 8 * kWgBlkBytes;         // 4 A blocks + 4 B blocks
constexpr int kWgStagesPerTile = kTileRows / kWgStageRows;

enum WgKind : int { kWgFirst = 0, kWgHidden = 1, kWgFinal = 2 };

struct WgItem {
  int kind;
  int cta_begin, cta_count;  // CTAs [cta_begin, cta_begin + cta_count) share this item
  const uint8_t* a_src;      // tile stride kTileBytes, 4 blocks          (kFinal: Y_L, else dTheta_l)
  const uint8_t* b_src;      // kHidden: Y_{l-1} (4 blocks); kFinal / kFirst: dOut / coordinate tiles (1 block, 16 KB)
  float* gw;                 // gradient of the weight (reference layout [out, in])
  float* gb;                 // gradient of the bias
  float scale;               // omega of the layer (1 for the final linear)
};

struct WgParams {
  WgItem items[kMaxSineLayers + 2];
  int num_items;
  int num_tiles;
  long long rows;
  int d, C;
  const float* coords;
  GridDesc grid;
};

template <int H>
struct WgSmem {
  static constexpr int kOffStage = 0;
  static constexpr int kOffBar = kWgStages * kWgStageBytes;
  static constexpr int kBytes = kOffBar + 256;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int H>
__global__ void __launch_bounds__(kWgThreads, 1) siren_wgrad_kernel(const WgParams p) {
  static_assert(H == 256, "tile blocking below assumes 4 blocks of 64 features");
  using S = WgSmem<H>;
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
  for (int i = 0; i < p.num_items; ++i)
    if (int(blockIdx.x) >= p.items[i].cta_begin && int(blockIdx.x) < p.items[i].cta_begin + p.items[i].cta_count) it = i;
  const WgItem item = p.items[it];
  const int split = int(blockIdx.x) - item.cta_begin;
  const int tile_begin = int((long long)p.num_tiles * split / item.cta_count);
  const int tile_end = int((long long)p.num_tiles * (split + 1) / item.cta_count);
  const int num_stages = (tile_end - tile_begin) * kWgStagesPerTile;
  const bool first = item.kind == kWgFirst;
  constexpr size_t kTileBytes = size_t(kTileRows) * H * 2;

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
        const int na = 4;
        const int nb = item.kind == kWgHidden ? 4 : 1;
        const uint32_t bytes = uint32_t(na + nb) * kWgBlkBytes;
        for (int s = 0; s < num_stages; ++s) {
          const int slot = s % kWgStages;
          const int round = s / kWgStages;
          const int tile = tile_begin + s / kWgStagesPerTile;
          const int sub = s % kWgStagesPerTile;
          if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
          mbar_arrive_expect_tx(&full[slot], bytes);
          uint8_t* dst = smem + S::kOffStage + slot * kWgStageBytes;
          const uint8_t* a = item.a_src + size_t(tile) * kTileBytes + size_t(sub) * kWgBlkBytes;
          for (int b = 0; b < na; ++b)
            bulk_g2s(dst + b * kWgBlkBytes, a + size_t(b) * (kTileRows * 128), kWgBlkBytes, &full[slot]);
          if (item.kind == kWgHidden) {
            const uint8_t* bsrc = item.b_src + size_t(tile) * kTileBytes + size_t(sub) * kWgBlkBytes;
            for (int b = 0; b < 4; ++b)
              bulk_g2s(dst + (4 + b) * kWgBlkBytes, bsrc + size_t(b) * (kTileRows * 128), kWgBlkBytes, &full[slot]);
          } else {
            const uint8_t* bsrc = item.b_src + size_t(tile) * (kTileRows * 128) + size_t(sub) * kWgBlkBytes;
            bulk_g2s(dst + 4 * kWgBlkBytes, bsrc, kWgBlkBytes, &full[slot]);
          }
        }
      }
    } else if (warp == 1) {
      // =============================== MMA issuer ===============================
      if (lane == 0) {
        // MN-major operands: LBO = stride between 64-wide feature blocks, SBO = 8-row groups along K.
        const uint64_t hi = smem_desc_hi_sw128(kWgBlkBytes, 1024);
        const int ncols = item.kind == kWgHidden ? H : kDzoPad;  // N of the MMA == TMEM columns per M half
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
              umma_bf16_ss(tmem_d + mh * ncols, da, db, idesc, (s | ks) != 0);
            }
          }
          umma_commit(&empty[slot]);
        }
        umma_commit(d_full);
      }
    } else {
      // =============================== column sums + flush ===============================
      const int at = threadIdx.x - 64;  // 0..127
      const int q = warp & 3;           // TMEM lane quadrant of this warp
      // column pair owned by this thread inside the summed operand: features 2*at, 2*at + 1
      const int sum_blocks = item.kind == kWgFinal ? 1 : 4;
      const int sum_off = item.kind == kWgFinal ? 4 * kWgBlkBytes : 0;  // dOut (final) or dTheta (sine layers)
      const bool sums = (2 * at) < sum_blocks * 64;
      const int sblk = (2 * at) >> 6, sch = ((2 * at) & 63) >> 3, sel = (2 * at) & 7;
      float s0 = 0.f, s1 = 0.f;
      for (int s = 0; s < num_stages; ++s) {
        const int slot = s % kWgStages;
        mbar_wait(&full[slot], (s / kWgStages) & 1);
        if (sums) {
          const uint8_t* blk = smem + S::kOffStage + slot * kWgStageBytes + sum_off + sblk * kWgBlkBytes;
#pragma unroll 8
          for (int rr = 0; rr < kWgStageRows; ++rr) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(blk + sw128_chunk_off(rr, sch) + sel * 2);
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
        const int nfeat = item.kind == kWgFinal ? p.C : H;
        if (c0 < nfeat) atomicAdd(item.gb + c0, item.scale * s0);
        if (c0 + 1 < nfeat) atomicAdd(item.gb + c0 + 1, item.scale * s1);
      }
      // ---- flush the TMEM partial products
      {
        mbar_wait(d_full, 0);
        tc_fence_after();
        const uint32_t t_lane = uint32_t(q * 32) << 16;
        const int ncols = item.kind == kWgHidden ? H : kDzoPad;
#pragma unroll 1
        for (int mh = 0; mh < 2; ++mh) {
          const int feat = mh * 128 + q * 32 + lane;  // M index: out feature (hidden) / in feature (final)
#pragma unroll 1
          for (int c0 = 0; c0 < ncols; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_d + t_lane + mh * ncols + c0, v);
            tmem_ld_wait();
            if (item.kind == kWgHidden) {
              float* dst = item.gw + (long long)feat * H + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                red_add_v4(dst + j, item.scale * __uint_as_float(v[j]), item.scale * __uint_as_float(v[j + 1]),
                           item.scale * __uint_as_float(v[j + 2]), item.scale * __uint_as_float(v[j + 3]));
            } else if (item.kind == kWgFinal) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;  // N index: output channel
                if (c < p.C) atomicAdd(item.gw + (long long)c * H + feat, item.scale * __uint_as_float(v[j]));
              }
            } else if (c0 == 0) {  // first layer: columns 0..3 = sum dTheta*x_hi, 4..7 = sum dTheta*x_lo
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < p.d)
                  atomicAdd(item.gw + (long long)feat * p.d + j,
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

int launch_siren_wgrad(const b200inr_net* net, void* stash, const float* coords, const b200inr_grid* grid,
                       int64_t rows, float* grad_params, int num_sms, cudaStream_t stream) {
  constexpr int H = 256;
  const int L = net->hidden_layers, d = net->in_features, C = net->out_features;
  const StashLayout sl = make_stash_layout(H, L, rows);
  const uint8_t* st = reinterpret_cast<const uint8_t*>(stash);
  int64_t off[2 * (kMaxSineLayers + 1)];
  param_offsets(d, H, L, C, off);

  WgParams p{};
  p.num_tiles = int(sl.tiles);
  p.rows = rows;
  p.d = d;
  p.C = C;
  p.coords = coords;
  if (grid) {
    p.grid.ndim = grid->ndim;
    long long tot = 1;
    for (int j = 0; j < 4; ++j) {
      p.grid.shape[j] = (j < grid->ndim) ? grid->shape[j] : 1;
      tot *= p.grid.shape[j];
    }
    p.grid.row_begin = grid->row_begin;
    p.grid.total = tot;
  }
  // CTA shares proportional to the bytes each item streams per tile (the kernel is HBM-bound).
  const double w_first = 80.0, w_hidden = 128.0, w_final = 80.0;
  const double w_total = w_first + L * w_hidden + w_final;
  int n_first = int(num_sms * w_first / w_total + 0.5);
  int n_final = int(num_sms * w_final / w_total + 0.5);
  if (n_first < 1) n_first = 1;
  if (n_final < 1) n_final = 1;
  int n_hidden = L > 0 ? (num_sms - n_first - n_final) / L : 0;
  if (L > 0 && n_hidden < 1) n_hidden = 1;
  int cta = 0, ni = 0;
  int mask = 7;  // development hook: B200INR_WGRAD_ITEMS = bitmask {1: hidden, 2: final, 4: first}
  if (const char* e = getenv("B200INR_WGRAD_ITEMS")) mask = atoi(e);
  auto add = [&](int kind, int count, const uint8_t* a, const uint8_t* b, float* gw, float* gb, float scale) {
    if (!(mask & (kind == kWgHidden ? 1 : (kind == kWgFinal ? 2 : 4)))) return;
    WgItem& w = p.items[ni++];
    w.kind = kind;
    w.cta_begin = cta;
    w.cta_count = count;
    w.a_src = a;
    w.b_src = b;
    w.gw = gw;
    w.gb = gb;
    w.scale = scale;
    cta += count;
  };
  for (int l = 1; l <= L; ++l)
    add(kWgHidden, n_hidden, st + sl.dz + size_t(l) * sl.layer_stride, st + sl.y + size_t(l - 1) * sl.layer_stride,
        grad_params + off[2 * l], grad_params + off[2 * l + 1], net->hidden_omega_0);
  add(kWgFinal, n_final, st + sl.y + size_t(L) * sl.layer_stride, st + sl.dzo, grad_params + off[2 * (L + 1)],
      grad_params + off[2 * (L + 1) + 1], 1.0f);
  add(kWgFirst, n_first, st + sl.dz, st + sl.xa, grad_params + off[0], grad_params + off[1], net->first_omega_0);
  p.num_items = ni;

  const int smem = WgSmem<H>::kBytes + 1024;
  if (cudaFuncSetAttribute(siren_wgrad_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  siren_wgrad_kernel<H><<<cta, kWgThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
