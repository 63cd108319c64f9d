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
// Same two-tile ping-pong as mlp_fwd.cu (see the comment above the kernel).
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
constexpr int kBwdSlots = 3;

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
  static constexpr int kABytes = kKB * kABlock;  // dTheta tile; block 0 doubles as the dOut block of chain step 0
  static constexpr int kSlotBytes = H * 128;     // [H rows (N = in)][64 (K = out chunk)]
  static constexpr int kOffA = 0;                // two tiles
  static constexpr int kOffW = 2 * kABytes;
  static constexpr int kOffBar = kOffW + kBwdSlots * kSlotBytes;
  static constexpr int kBytes = kOffBar + 256;
};

constexpr float kPhaseToRad = 9.587379924285257e-05f;  // 2*pi / 65536

__device__ __forceinline__ float cos_from_phase(uint32_t ph16) {
  const float f = __uint_as_float(0x4B000000u | ph16) - 8388608.0f;  // exact u16 -> float without I2F
  return __cosf(f * kPhaseToRad);
}

// Two 128-row tiles (X, Y) ping-pong exactly as in mlp_fwd.cu: one dTheta tile in shared memory and one 256-column
// TMEM accumulator each; the epilogue warps alternate X, Y chain step by chain step, so every MMA runs underneath the
// other tile's epilogue.  Stage 0 of a tile converts dOut to bf16 into block 0 of its A tile (K = 64 operand of the
// first chain step), stages 1 .. L+1 produce dTheta_L .. dTheta_0.
template <int H>
__global__ void __launch_bounds__(kBwdThreads, 1) siren_bwd_kernel(const BwdParams p) {
  using S = BwdSmem<H>;
  static_assert(S::kKB == 4, "epilogue slicing assumes 4 K blocks");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem + S::kOffA;
  uint8_t* w_smem = smem + S::kOffW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* w_full = bars;                      // [kBwdSlots]
  uint64_t* w_empty = bars + kBwdSlots;         // [kBwdSlots]
  uint64_t* a_ready = bars + 2 * kBwdSlots;     // [2] operand tile j complete in shared memory
  uint64_t* d_full = bars + 2 * kBwdSlots + 2;  // [2] accumulator j complete
  uint64_t* a_free = bars + 2 * kBwdSlots + 4;  // [2] stash store out of tile j has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBwdSlots + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int L = p.L;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdSlots; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(&a_ready[j], kBwdEpiWarps);
      mbar_init(&d_full[j], 1);
      mbar_init(&a_free[j], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int my_tiles = (p.num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int num_pairs = (my_tiles + 1) / 2;

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t c = 0;
      for (int pr = 0; pr < num_pairs; ++pr) {
        const int nt = (my_tiles - 2 * pr) < 2 ? (my_tiles - 2 * pr) : 2;
        for (int u = 0; u <= L; ++u) {  // u = 0: W_f^T (one chunk); u >= 1: the kKB chunks of W'^T of layer L - u + 1
          const int nchunks = (u == 0) ? 1 : S::kKB;
          const uint8_t* src = (u == 0) ? p.packed + p.pl.wft : p.packed + p.pl.wht + size_t(L - u) * H * H * 2;
          for (int j = 0; j < nt; ++j) {
            const int tile = int(blockIdx.x) + (2 * pr + j) * int(gridDim.x);
            // the epilogue after chain step u needs the phases of layer L - u: pull them into L2 now
            bulk_prefetch_l2(p.stash_ph + size_t(L - u) * p.layer_stride + size_t(tile) * S::kABytes, S::kABytes);
            for (int kb = 0; kb < nchunks; ++kb, ++c) {
              const uint32_t slot = c % kBwdSlots, round = c / kBwdSlots;
              if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
              mbar_arrive_expect_tx(&w_full[slot], S::kSlotBytes);
              bulk_g2s(w_smem + slot * S::kSlotBytes, src + size_t(kb) * S::kSlotBytes, S::kSlotBytes, &w_full[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // the whole warp runs the loop converged, one elected lane issues (umma_*_w: no per-instruction R2UR loop)
    {
      const uint64_t hi = smem_desc_hi_sw128(0, 1024);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t idesc = idesc_bf16(128, H, false, false);
      uint32_t c = 0, na[2] = {0, 0};
      for (int pr = 0; pr < num_pairs; ++pr) {
        const int nt = (my_tiles - 2 * pr) < 2 ? (my_tiles - 2 * pr) : 2;
        for (int u = 0; u <= L; ++u) {
          const int nkb = (u == 0) ? 1 : S::kKB;
          for (int j = 0; j < nt; ++j) {
            mbar_wait(&a_ready[j], na[j] & 1);
            ++na[j];
            tc_fence_after();
            for (int kb = 0; kb < nkb; ++kb, ++c) {
              const uint32_t slot = c % kBwdSlots;
              mbar_wait(&w_full[slot], (c / kBwdSlots) & 1);
              tc_fence_after();
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                umma_bf16_ss_w(tmem_d + j * 256, smem_desc(a_base + j * S::kABytes + kb * S::kABlock + k4 * 32, hi),
                             smem_desc(w_base + slot * S::kSlotBytes + k4 * 32, hi), idesc, (kb | k4) != 0);
              }
              umma_commit_w(&w_empty[slot]);
            }
            umma_commit_w(&d_full[j]);
          }
        }
        // the dTheta_0 tiles feed no MMA, but their a_ready phase must still be observed: a parity wait may only
        // ever be one phase behind the barrier
        for (int j = 0; j < nt; ++j) {
          mbar_wait(&a_ready[j], na[j] & 1);
          ++na[j];
        }
      }
    }
  } else if (warp == 2) {
    // =============================== stash store ===============================
    if (lane == 0) {
      uint32_t na[2] = {0, 0};
      for (int pr = 0; pr < num_pairs; ++pr) {
        const int nt = (my_tiles - 2 * pr) < 2 ? (my_tiles - 2 * pr) : 2;
        for (int st = 0; st <= L + 1; ++st) {  // stage 0: dOut block; stage st >= 1: dTheta of layer L + 1 - st
          for (int j = 0; j < nt; ++j) {
            const int tile = int(blockIdx.x) + (2 * pr + j) * int(gridDim.x);
            mbar_wait(&a_ready[j], na[j] & 1);
            ++na[j];
            if (st == 0)
              bulk_s2g(p.stash_dzo + size_t(tile) * (kTileRows * 128), a_smem + j * S::kABytes, kTileRows * 128);
            else
              bulk_s2g(p.stash_dz + size_t(L + 1 - st) * p.layer_stride + size_t(tile) * S::kABytes,
                       a_smem + j * S::kABytes, S::kABytes);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(&a_free[j]);
          }
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
    const int C = p.C;
    uint32_t nd[2] = {0, 0}, nf[2] = {0, 0};
    bool first_store[2] = {true, true};
    for (int pr = 0; pr < num_pairs; ++pr) {
      const int nt = (my_tiles - 2 * pr) < 2 ? (my_tiles - 2 * pr) : 2;

      // ---- stage 0: dOut tile -> bf16 [128][64] block in block 0 of the tile (columns >= C, rows >= rows zero)
      for (int j = 0; j < nt; ++j) {
        const int tile = int(blockIdx.x) + (2 * pr + j) * int(gridDim.x);
        const long long row0 = (long long)tile * kTileRows;
        const uint32_t a_addr = smem_u32(a_smem) + j * S::kABytes;
        if (!first_store[j]) {  // the previous pair's dTheta_0 store out of this tile has been read
          mbar_wait(&a_free[j], nf[j] & 1);
          ++nf[j];
        }
        first_store[j] = false;
        const bool valid = (row0 + r) < p.rows;
        const float* g = p.grad_out + (row0 + r) * C;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ch = 2 * s + cc;
          float v[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int col = ch * 8 + jj;
            v[jj] = (valid && col < C) ? g[col] : 0.f;
          }
          sts128(a_addr + sw128_chunk_off(r, ch),
                 make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                            pack_bf16x2(v[6], v[7])));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[j]);
      }

      // ---- dTheta_l = dY_l .* cos(theta_l), l = L .. 0, alternating X, Y
      for (int l = L; l >= 0; --l) {
        for (int j = 0; j < nt; ++j) {
          const int tile = int(blockIdx.x) + (2 * pr + j) * int(gridDim.x);
          const uint32_t a_addr = smem_u32(a_smem) + j * S::kABytes;
          const uint8_t* ph_l = p.stash_ph + size_t(l) * p.layer_stride + size_t(tile) * S::kABytes + size_t(r) * 16 +
                                size_t(2 * s) * (kTileRows * 16);
          uint4 ph[2], phn[2];
          phn[0] = *reinterpret_cast<const uint4*>(ph_l);
          phn[1] = *reinterpret_cast<const uint4*>(ph_l + kTileRows * 16);
          mbar_wait(&d_full[j], nd[j] & 1);
          ++nd[j];
          mbar_wait(&a_free[j], nf[j] & 1);  // the previous store out of this tile (dOut block or dTheta_{l+1})
          ++nf[j];
          tc_fence_after();
          const uint32_t d_addr = tmem_d + t_lane + uint32_t(j) * 256 + s * 16;
          uint32_t v[16], vn[16];
          tmem_ld16(d_addr, vn);
#pragma unroll
          for (int kb = 0; kb < S::kKB; ++kb) {
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] = vn[jj];
            ph[0] = phn[0];
            ph[1] = phn[1];
            if (kb + 1 < S::kKB) {
              tmem_ld16(d_addr + (kb + 1) * 64, vn);
              phn[0] = *reinterpret_cast<const uint4*>(ph_l + size_t((kb + 1) * 8) * (kTileRows * 16));
              phn[1] = *reinterpret_cast<const uint4*>(ph_l + size_t((kb + 1) * 8 + 1) * (kTileRows * 16));
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint32_t pw[4] = {ph[c].x, ph[c].y, ph[c].z, ph[c].w};
              uint32_t o[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float d0 = __uint_as_float(v[c * 8 + 2 * jj]) * cos_from_phase(pw[jj] & 0xFFFFu);
                const float d1 = __uint_as_float(v[c * 8 + 2 * jj + 1]) * cos_from_phase(pw[jj] >> 16);
                o[jj] = pack_bf16x2(d0, d1);
              }
              sts128(a_addr + kb * S::kABlock + sw128_chunk_off(r, 2 * s + c), make_uint4(o[0], o[1], o[2], o[3]));
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[j]);
        }
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
  const int pairs = (p.num_tiles + 1) / 2;
  const int grid_x = pairs < num_sms ? pairs : num_sms;
  if (cudaFuncSetAttribute(siren_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  siren_bwd_kernel<H><<<grid_x, kBwdThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
