// mlp_bwd.cu -- fused SIREN backward, activation-gradient chain (dgrad) for sm_100a.
//
// Replaces loss.backward() through nn.Sequential(SineLayer x (L+1), nn.Linear) (autograd of the reference's
// INR/SRDWI.py:58-59,87-91; math in SURVEY.md App. B.1).  With theta_l = omega_l z_l and the omega-folded bf16
// weights W'_l = omega_l W_l this kernel computes, per 128-row tile and entirely on chip,
//     dTheta_L   = (dOut  W_f ) .* cos(theta_L)
//     dTheta_l-1 = (dTheta_l W'_l) .* cos(theta_l-1)          l = L .. 1
// (tcgen05.mma 128x256x{64,256}, fp32 accumulate in TMEM; cos from the 16-bit phases stashed by the forward)
// and writes every dTheta_l and the bf16 dOut tile to the stash in UMMA tile layout, where wgrad.cu contracts them
// with the stashed activations over the row dimension.  No input gradient: SRDWI.Siren detaches its coordinates
// (INR/SRDWI.py:88).
//
// Warp roles: warp 0 = bulk-copy producer (W'^T chunks), warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue.
#include <stdio.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

constexpr int kBwdThreads = 320;
constexpr int kBwdEpiThreads = 256;
constexpr uint32_t kBwdEpiBarId = 1;
constexpr int kBwdSlots = 4;

struct BwdParams {
  const uint8_t* packed;
  PackLayout pl;
  long long rows;
  int num_tiles;
  int L, C;
  const float* grad_out;  // [rows, C]
  const uint8_t* stash_ph;
  uint8_t* stash_dz;
  uint8_t* stash_dzo;
  size_t layer_stride;
};

template <int H>
struct BwdSmem {
  static constexpr int kKB = H / 64;
  static constexpr int kABlock = kTileRows * 128;
  static constexpr int kABytes = kKB * kABlock;  // dTheta tile
  static constexpr int kSlotBytes = H * 128;     // [H rows (N = in)][64 (K = out chunk)]
  static constexpr int kOffA = 0;
  static constexpr int kOffW = kABytes;
  static constexpr int kOffDzo = kOffW + kBwdSlots * kSlotBytes;
  static constexpr int kOffBar = kOffDzo + kTileRows * 128;
  static constexpr int kBytes = kOffBar + 128;
};

constexpr float kPhaseToRad = 9.587379924285257e-05f;  // 2*pi / 65536

__device__ __forceinline__ float cos_from_phase(uint32_t ph16) {
  const float f = __uint_as_float(0x4B000000u | ph16) - 8388608.0f;  // exact u16 -> float without I2F
  return __cosf(f * kPhaseToRad);
}

template <int H>
__global__ void __launch_bounds__(kBwdThreads, 1) siren_bwd_kernel(const BwdParams p) {
  using S = BwdSmem<H>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem + S::kOffA;
  uint8_t* w_smem = smem + S::kOffW;
  uint8_t* dzo_smem = smem + S::kOffDzo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* w_full = bars;                   // [kBwdSlots]
  uint64_t* w_empty = bars + kBwdSlots;      // [kBwdSlots]
  uint64_t* a_ready = bars + 2 * kBwdSlots;
  uint64_t* d_full = bars + 2 * kBwdSlots + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBwdSlots + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int L = p.L;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdSlots; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(a_ready, kBwdEpiThreads);
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int my_tiles = (p.num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int chunks_per_tile = 1 + L * S::kKB;  // W_f^T, then kKB chunks of each W'_l^T, l = L .. 1

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int j = 0; j < chunks_per_tile; ++j, ++c) {
          const uint32_t slot = c % kBwdSlots;
          const uint32_t round = c / kBwdSlots;
          const uint8_t* src;
          if (j == 0) {
            src = p.packed + p.pl.wft;
          } else {
            const int l = L - (j - 1) / S::kKB;  // hidden layer whose weights are used (L .. 1)
            const int kb = (j - 1) % S::kKB;
            src = p.packed + p.pl.wht + size_t(l - 1) * H * H * 2 + size_t(kb) * S::kSlotBytes;
          }
          if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
          mbar_arrive_expect_tx(&w_full[slot], S::kSlotBytes);
          bulk_g2s(w_smem + slot * S::kSlotBytes, src, S::kSlotBytes, &w_full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint64_t hi = smem_desc_hi_sw128(0, 1024);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t dzo_base = smem_u32(dzo_smem);
      const uint32_t idesc = idesc_bf16(128, H, false, false);
      uint32_t c = 0, n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int u = 0; u <= L; ++u) {  // u = 0: dOut W_f ; u >= 1: dTheta_{L-u+1} W'_{L-u+1}
          mbar_wait(a_ready, n & 1);
          ++n;
          tc_fence_after();
          const int nkb = (u == 0) ? 1 : S::kKB;
          for (int kb = 0; kb < nkb; ++kb, ++c) {
            const uint32_t slot = c % kBwdSlots;
            mbar_wait(&w_full[slot], (c / kBwdSlots) & 1);
            tc_fence_after();
            const uint32_t a_blk = (u == 0) ? dzo_base : a_base + kb * S::kABlock;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              umma_bf16_ss(tmem_d, smem_desc(a_blk + k4 * 32, hi),
                           smem_desc(w_base + slot * S::kSlotBytes + k4 * 32, hi), idesc, (kb | k4) != 0);
            }
            umma_commit(&w_empty[slot]);
          }
          umma_commit(d_full);
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = uint32_t(q * 32) << 16;
    const int C = p.C;
    uint32_t n = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = int(blockIdx.x) + t * int(gridDim.x);
      const long long row0 = (long long)tile * kTileRows;
      const uint8_t* ph_tile = p.stash_ph + size_t(tile) * S::kABytes;
      uint8_t* dz_tile = p.stash_dz + size_t(tile) * S::kABytes;

      // ---- dOut tile -> bf16 [128][64] block (columns >= C and rows >= rows are zero)
      if (et == 0) bulk_wait_read0();  // previous tile's stores out of dzo_smem / a_smem have been read
      named_bar_sync(kBwdEpiBarId, kBwdEpiThreads);
      {
        const bool valid = (row0 + r) < p.rows;
        const float* g = p.grad_out + (row0 + r) * C;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int ch = h * 4 + cc;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = ch * 8 + j;
            v[j] = (valid && col < C) ? g[col] : 0.f;
          }
          *reinterpret_cast<uint4*>(dzo_smem + sw128_chunk_off(r, ch)) =
              make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                         pack_bf16x2(v[6], v[7]));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(kBwdEpiBarId, kBwdEpiThreads);
        if (et == 0) {
          bulk_s2g(p.stash_dzo + size_t(tile) * (kTileRows * 128), dzo_smem, kTileRows * 128);
          bulk_commit();
        }
        mbar_arrive(a_ready);
      }

      // ---- dTheta_l = dY_l .* cos(theta_l), l = L .. 0
      for (int l = L; l >= 0; --l) {
        const uint8_t* ph_l = ph_tile + size_t(l) * p.layer_stride;
        mbar_wait(d_full, n & 1);
        ++n;
        tc_fence_after();
        if (et == 0) bulk_wait_read0();
        named_bar_sync(kBwdEpiBarId, kBwdEpiThreads);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int col0 = h * 128 + cc * 32;
          uint4 ph[4];
#pragma unroll
          for (int g = 0; g < 4; ++g)
            ph[g] = *reinterpret_cast<const uint4*>(ph_l + (size_t((col0 >> 3) + g) * kTileRows + r) * 16);
          uint32_t v[32];
          tmem_ld32(tmem_d + t_lane + col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t pw[4] = {ph[g].x, ph[g].y, ph[g].z, ph[g].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float d0 = __uint_as_float(v[g * 8 + 2 * j]) * cos_from_phase(pw[j] & 0xFFFFu);
              const float d1 = __uint_as_float(v[g * 8 + 2 * j + 1]) * cos_from_phase(pw[j] >> 16);
              o[j] = pack_bf16x2(d0, d1);
            }
            const int col = col0 + g * 8;
            *reinterpret_cast<uint4*>(a_smem + (col >> 6) * S::kABlock + sw128_chunk_off(r, (col & 63) >> 3)) =
                make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(kBwdEpiBarId, kBwdEpiThreads);
        if (et == 0) {
          bulk_s2g(dz_tile + size_t(l) * p.layer_stride, a_smem, S::kABytes);
          bulk_commit();
        }
        if (l > 0) mbar_arrive(a_ready);
      }
    }
    if (et == 0) bulk_wait0();
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_d);
}

int launch_siren_bwd(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                     int num_sms, cudaStream_t stream) {
  constexpr int H = 256;
  BwdParams p{};
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.pl = make_pack_layout(H, net->hidden_layers);
  p.rows = rows;
  p.num_tiles = int((rows + kTileRows - 1) / kTileRows);
  p.L = net->hidden_layers;
  p.C = net->out_features;
  p.grad_out = grad_out;
  const StashLayout sl = make_stash_layout(H, net->hidden_layers, rows);
  p.stash_ph = reinterpret_cast<const uint8_t*>(stash) + sl.ph;
  p.stash_dz = reinterpret_cast<uint8_t*>(stash) + sl.dz;
  p.stash_dzo = reinterpret_cast<uint8_t*>(stash) + sl.dzo;
  p.layer_stride = sl.layer_stride;
  const int smem = BwdSmem<H>::kBytes + 1024;
  const int grid_x = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (cudaFuncSetAttribute(siren_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  siren_bwd_kernel<H><<<grid_x, kBwdThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
