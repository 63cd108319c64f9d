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
// Same pipeline as mlp_fwd.cu: the epilogue emits dTheta_l one 64-wide K block at a time (one mbarrier per block),
// the accumulator is double buffered in TMEM, so the MMAs of the next chain step run underneath the epilogue.
//
// Warp roles: warp 0 = weight producer (W'^T chunks), warp 1 = MMA issuer + TMEM owner, warp 2 = stash store,
//             warps 3..18 = epilogue (TMEM lane quadrant = warp & 3, 16-column slice = (warp - 3) >> 2).
#include <stdio.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

constexpr int kBwdEpiWarps = 16;
constexpr int kBwdFirstEpiWarp = 3;
constexpr int kBwdThreads = (kBwdFirstEpiWarp + kBwdEpiWarps) * 32;  // 608
constexpr int kBwdEpiThreads = kBwdEpiWarps * 32;
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
  static constexpr int kBytes = kOffBar + 256;
};

constexpr float kPhaseToRad = 9.587379924285257e-05f;  // 2*pi / 65536

__device__ __forceinline__ float cos_from_phase(uint32_t ph16) {
  const float f = __uint_as_float(0x4B000000u | ph16) - 8388608.0f;  // exact u16 -> float without I2F
  return __cosf(f * kPhaseToRad);
}

template <int H>
__global__ void __launch_bounds__(kBwdThreads, 1) siren_bwd_kernel(const BwdParams p) {
  using S = BwdSmem<H>;
  static_assert(S::kKB == 4, "epilogue slicing assumes 4 K blocks");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem + S::kOffA;
  uint8_t* w_smem = smem + S::kOffW;
  uint8_t* dzo_smem = smem + S::kOffDzo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* w_full = bars;                     // [kBwdSlots]
  uint64_t* w_empty = bars + kBwdSlots;        // [kBwdSlots]
  uint64_t* a_ready = bars + 2 * kBwdSlots;    // [4] K block kb of dTheta_l is in shared memory
  uint64_t* dzo_ready = bars + 2 * kBwdSlots + 4;
  uint64_t* d_full = bars + 2 * kBwdSlots + 5;
  uint64_t* a_free = bars + 2 * kBwdSlots + 6;     // [4] stash store of dTheta block kb has been read out
  uint64_t* dzo_free = bars + 2 * kBwdSlots + 10;  // stash store out of dzo_smem has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBwdSlots + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int L = p.L;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdSlots; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], kBwdEpiWarps);
    mbar_init(dzo_ready, kBwdEpiWarps);
    mbar_init(d_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&a_free[i], 1);
    mbar_init(dzo_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
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
        const int tile = int(blockIdx.x) + t * int(gridDim.x);
        for (int j = 0; j < chunks_per_tile; ++j, ++c) {
          const uint32_t slot = c % kBwdSlots;
          const uint32_t round = c / kBwdSlots;
          const uint8_t* src;
          if (j == 0 || (j - 1) % S::kKB == 0) {  // first chunk of chain step u: its epilogue needs the phases of layer L - u
            const int u = (j == 0) ? 0 : 1 + (j - 1) / S::kKB;
            bulk_prefetch_l2(p.stash_ph + size_t(L - u) * p.layer_stride + size_t(tile) * S::kABytes, S::kABytes);
          }
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
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        // a_ready completes L + 1 times per tile (layers L .. 0); instance u - 1 feeds chain step u
        const uint32_t inst0 = uint32_t(t) * uint32_t(L + 1);
        for (int u = 0; u <= L; ++u) {  // u = 0: dOut W_f ; u >= 1: dTheta_{L-u+1} W'_{L-u+1}
          const uint32_t d_addr = tmem_d + uint32_t(u & 1) * 256;
          const int nkb = (u == 0) ? 1 : S::kKB;
          for (int kb = 0; kb < nkb; ++kb, ++c) {
            const uint32_t slot = c % kBwdSlots;
            if (u == 0)
              mbar_wait(dzo_ready, t & 1);
            else
              mbar_wait(&a_ready[kb], (inst0 + u - 1) & 1);
            mbar_wait(&w_full[slot], (c / kBwdSlots) & 1);
            tc_fence_after();
            const uint32_t a_blk = (u == 0) ? dzo_base : a_base + kb * S::kABlock;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              umma_bf16_ss(d_addr, smem_desc(a_blk + k4 * 32, hi),
                           smem_desc(w_base + slot * S::kSlotBytes + k4 * 32, hi), idesc, (kb | k4) != 0);
            }
            umma_commit(&w_empty[slot]);
          }
          umma_commit(d_full);
        }
      }
    }
  } else if (warp == 2) {
    // =============================== stash store ===============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = int(blockIdx.x) + t * int(gridDim.x);
        uint8_t* dz_tile = p.stash_dz + size_t(tile) * S::kABytes;
        mbar_wait(dzo_ready, t & 1);
        bulk_s2g(p.stash_dzo + size_t(tile) * (kTileRows * 128), dzo_smem, kTileRows * 128);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(dzo_free);
        for (int l = L; l >= 0; --l, ++n) {
          for (int kb = 0; kb < S::kKB; ++kb) {
            mbar_wait(&a_ready[kb], n & 1);
            bulk_s2g(dz_tile + size_t(l) * p.layer_stride + size_t(kb) * S::kABlock, a_smem + kb * S::kABlock,
                     S::kABlock);
            bulk_commit();
            if (kb > 0) {
              bulk_wait_read1();
              mbar_arrive(&a_free[kb - 1]);
            }
          }
          bulk_wait_read0();
          mbar_arrive(&a_free[S::kKB - 1]);
        }
      }
      bulk_wait0();
    }
  } else if (warp >= kBwdFirstEpiWarp) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;
    const int s = (warp - kBwdFirstEpiWarp) >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = uint32_t(q * 32) << 16;
    const uint32_t a_addr = smem_u32(a_smem);
    const uint32_t dzo_addr = smem_u32(dzo_smem);
    const int C = p.C;
    uint32_t n = 0, nf = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = int(blockIdx.x) + t * int(gridDim.x);
      const long long row0 = (long long)tile * kTileRows;
      const uint8_t* ph_row = p.stash_ph + size_t(tile) * S::kABytes + size_t(r) * 16;

      // ---- dOut tile -> bf16 [128][64] block (columns >= C and rows >= rows are zero); slice s = chunks 2s, 2s+1
      if (t > 0) mbar_wait(dzo_free, (t - 1) & 1);  // the previous tile's dOut block has been stored and multiplied
      {
        const bool valid = (row0 + r) < p.rows;
        const float* g = p.grad_out + (row0 + r) * C;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ch = 2 * s + cc;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = ch * 8 + j;
            v[j] = (valid && col < C) ? g[col] : 0.f;
          }
          sts128(dzo_addr + sw128_chunk_off(r, ch),
                 make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                            pack_bf16x2(v[6], v[7])));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dzo_ready);
      }

      // ---- dTheta_l = dY_l .* cos(theta_l), l = L .. 0
      for (int l = L; l >= 0; --l) {
        const uint8_t* ph_l = ph_row + size_t(l) * p.layer_stride + size_t(2 * s) * (kTileRows * 16);
        // phases of the first K block are fetched while the MMAs are still running; the whole phase tile was pulled
        // into L2 by the producer thread one layer ahead
        uint4 ph[2], phn[2];
        phn[0] = *reinterpret_cast<const uint4*>(ph_l);
        phn[1] = *reinterpret_cast<const uint4*>(ph_l + kTileRows * 16);
        mbar_wait(d_full, n & 1);
        ++n;
        tc_fence_after();
        const uint32_t d_addr = tmem_d + t_lane + uint32_t((L - l) & 1) * 256 + s * 16;
        uint32_t v[16], vn[16];
        tmem_ld16(d_addr, vn);
#pragma unroll
        for (int kb = 0; kb < S::kKB; ++kb) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = vn[j];
          ph[0] = phn[0];
          ph[1] = phn[1];
          if (kb + 1 < S::kKB) {
            tmem_ld16(d_addr + (kb + 1) * 64, vn);
            phn[0] = *reinterpret_cast<const uint4*>(ph_l + size_t((kb + 1) * 8) * (kTileRows * 16));
            phn[1] = *reinterpret_cast<const uint4*>(ph_l + size_t((kb + 1) * 8 + 1) * (kTileRows * 16));
          }
          if (nf > 0) mbar_wait(&a_free[kb], (nf - 1) & 1);  // the previous dTheta block kb has been stored
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t pw[4] = {ph[c].x, ph[c].y, ph[c].z, ph[c].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float d0 = __uint_as_float(v[c * 8 + 2 * j]) * cos_from_phase(pw[j] & 0xFFFFu);
              const float d1 = __uint_as_float(v[c * 8 + 2 * j + 1]) * cos_from_phase(pw[j] >> 16);
              o[j] = pack_bf16x2(d0, d1);
            }
            sts128(a_addr + kb * S::kABlock + sw128_chunk_off(r, 2 * s + c), make_uint4(o[0], o[1], o[2], o[3]));
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[kb]);
        }
        ++nf;
      }
    }
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_d);
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
