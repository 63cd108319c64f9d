// gen_bwd.cu -- activation-gradient chain (dgrad) of the generic coordinate-MLP family (see gen_fwd.cu) for sm_100a.
//
// Replaces loss.backward() through the activated layers (autograd of INR/SRDWI.py:58-59 for sine; of nn.ReLU for the
// Fourier-feature MLP of BASELINE config 4):
//     dTheta_L   = (dOut  W_f ) .* act'(theta_L)
//     dTheta_l-1 = (dTheta_l W'_l) .* act'(theta_l-1)          l = L .. 1
// act' = cos(theta) from the stashed 16-bit phase (sine) or 1[y > 0] from the stashed bf16 output (ReLU).  The chain
// normally stops at layer 0.  With grad_in != nullptr (explicit feature rows whose producer needs a gradient: the
// reference's INRmodel.Siren does not detach its input, INR/INRmodel.py:147-149, and the PerturbNet phase trains
// through it, INR/inrDWI.py:141-147) one more step runs on the tensor cores,
//     dX = dTheta_0 (omega_0 W_0)                  [rows, K0] fp32, written to grad_in
// Every dTheta_l tile and the bf16 dOut tile go to the stash for wgrad.cu.
//
// Warp roles as in gen_fwd.cu.
#include <stdio.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

constexpr int kGenBwdEpiWarps = 16;
constexpr int kGenBwdFirstEpiWarp = 3;
constexpr int kGenBwdThreads = (kGenBwdFirstEpiWarp + kGenBwdEpiWarps) * 32;

struct GenBwdParams {
  const uint8_t* packed;
  GenDims g;
  GenPackLayout pl;
  long long rows;
  int num_tiles;
  const float* grad_out;  // [rows, C]
  const uint8_t* stash_y;
  const uint8_t* stash_ph;
  uint8_t* stash_dz;
  uint8_t* stash_dzo;
  size_t layer_stride;
  float* grad_in;  // nullptr, or [rows, K0] fp32: dL/d(network input)
  // IN_FOURIER: dL/d(coordinates) [rows, d] instead -- the adjoint of input_mapping (SURVEY.md App. B.2:
  // dp = dphi_sin cos p - dphi_cos sin p, dx = 2 pi dp B) applied to the input gradient while it is still in TMEM, so
  // the [rows, 2m] feature gradient never exists (the PerturbNet phase, INR/inrDWI.py:141-147)
  float* grad_coords;
  const float* coords;  // explicit coordinates [rows, d] or nullptr (grid)
  GridDesc grid;
  const float* out_y;   // B200INR_NET_TANH_OUT: the forward's output y = s tanh(.), [rows, C]; dOut *= s - y^2 / s
  float out_tanh;
  int lean;             // B200INR_NET_DGRAD_ONLY: no dL/dtheta tiles are stored (there is no weight-gradient pass)
};

template <int H>
struct GenBwdSmem {
  static constexpr int kKB = H / 64;
  static constexpr int kABlock = kTileRows * 128;
  static constexpr int kABytes = kKB * kABlock;
  // CTA pairs (cta_group::2, as in gen_fwd.cu): every CTA stages HALF of a weight chunk (128 of its 256 rows), so the
  // 64 KB the two whole chunks took at H = 512 now hold a four-deep ring
  static constexpr int kSlots = (H == 512) ? 4 : 6;
  static constexpr int kSlotBytes = kGenChunkBytes / 2;
  static constexpr int kOffA = 0;
  static constexpr int kOffW = kABytes;
  static constexpr int kOffDzo = kOffW + kSlots * kSlotBytes;
  static constexpr int kOffBar = kOffDzo + kTileRows * 128;
  static constexpr int kOffRed = kOffBar + 256;                   // [4 slices][128 rows] float4: coordinate-gradient partials
  static constexpr int kBytes = kOffRed + 4 * kTileRows * 16;
};

constexpr float kGenPhaseToRad = 9.587379924285257e-05f;  // 2*pi / 65536

__device__ __forceinline__ float gen_cos_from_phase(uint32_t ph16) {
  const float f = __uint_as_float(0x4B000000u | ph16) - 8388608.0f;
  return __cosf(f * kGenPhaseToRad);
}

template <int H, int ACT>
__global__ void __launch_bounds__(kGenBwdThreads, 1) gen_bwd_kernel(const GenBwdParams p) {
  using S = GenBwdSmem<H>;
  constexpr int NH = H / 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem + S::kOffA;
  uint8_t* w_smem = smem + S::kOffW;
  uint8_t* dzo_smem = smem + S::kOffDzo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* w_full = bars;               // leader: own half landed + the peer's relay; peer: own half
  uint64_t* w_empty = bars + S::kSlots;  // multicast commit
  // the dTheta tile (A operand of the next chain step) is handed over in two halves of K blocks, as in gen_fwd.cu: at
  // H = 512 the next step's MMAs into D[0:256) start while the epilogue still works on D[256:512).
  // a_half / dzo_ready: on the LEADER, count the epilogue warps of both CTAs (the MMA issuer's view);
  // a_loc / dzo_loc: this CTA's own warps (its stash-store thread's view)
  uint64_t* a_half = bars + 2 * S::kSlots;  // [2]
  uint64_t* dzo_ready = bars + 2 * S::kSlots + 2;
  uint64_t* d_full = bars + 2 * S::kSlots + 3;  // multicast commit
  uint64_t* a_free = bars + 2 * S::kSlots + 4;
  uint64_t* dzo_free = bars + 2 * S::kSlots + 5;
  uint64_t* a_loc = bars + 2 * S::kSlots + 6;  // [2]
  uint64_t* dzo_loc = bars + 2 * S::kSlots + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::kSlots + 9);
  static_assert((2 * S::kSlots + 9) * 8 + 4 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const GenDims g = p.g;
  const int L = g.L;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S::kSlots; ++i) {
      mbar_init(&w_full[i], leader ? 2 : 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(&a_half[0], 2 * kGenBwdEpiWarps);
    mbar_init(&a_half[1], 2 * kGenBwdEpiWarps);
    mbar_init(dzo_ready, 2 * kGenBwdEpiWarps);
    mbar_init(&a_loc[0], kGenBwdEpiWarps);
    mbar_init(&a_loc[1], kGenBwdEpiWarps);
    mbar_init(dzo_loc, kGenBwdEpiWarps);
    mbar_init(d_full, 1);
    mbar_init(a_free, 1);
    mbar_init(dzo_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // the pair walks tile pairs (tile = 2 * (pair + t * pairs) + rank); a peer past the end recomputes the last tile
  const int num_pairs_grid = int(gridDim.x) / 2;
  const int tile_pairs = (p.num_tiles + 1) / 2;
  const int my_tiles = (tile_pairs - int(blockIdx.x) / 2 + num_pairs_grid - 1) / num_pairs_grid;
  auto tile_of = [&](int t) {
    const int tile = 2 * (int(blockIdx.x) / 2 + t * num_pairs_grid) + int(rank);
    return tile < p.num_tiles ? tile : p.num_tiles - 1;
  };
  const int dx = (p.grad_in != nullptr || p.grad_coords != nullptr) ? 1 : 0;  // extra chain step: input gradient
  const int NX = (g.K0 + 255) / 256;            // its N, in 256-column halves

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        // u = 0: W_f^T (NH chunks); u = 1..L: W'^T of layer L - u + 1 (NH * kKB chunks); u = L + 1 (input gradient):
        // (omega_0 W_0)^T, NX * kKB chunks
        for (int u = 0; u <= L + dx; ++u) {
          const int nchunks = (u == 0) ? NH : (u <= L ? NH : NX) * S::kKB;
          const uint8_t* src = (u == 0)   ? p.packed + p.pl.wft
                               : (u <= L) ? p.packed + p.pl.wt_layer(g, L - u + 1)
                                          : p.packed + p.pl.wt0;
          for (int j = 0; j < nchunks; ++j, ++c) {
            // this CTA's half of the chunk: rows [128 rank, 128 rank + 128) (accumulator column == input feature)
            const uint32_t slot = c % S::kSlots, round = c / S::kSlots;
            if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
            mbar_arrive_expect_tx(&w_full[slot], S::kSlotBytes);
            bulk_g2s(w_smem + slot * S::kSlotBytes, src + size_t(j) * kGenChunkBytes + size_t(rank) * S::kSlotBytes,
                     S::kSlotBytes, &w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // =============================== MMA issuer (pair leader) ===============================
      // whole warp converged, one elected lane issues (umma_*_w); M = 256: both CTAs' tiles in one instruction
      const uint64_t hi = smem_desc_hi_sw128(0, 1024);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t dzo_base = smem_u32(dzo_smem);
      const uint32_t idesc = idesc_bf16(256, 256, false, false);
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t inst0 = uint32_t(t) * uint32_t(L + 1);  // a_ready completes L + 1 times per tile
        for (int u = 0; u <= L + dx; ++u) {
          bool second = (u == 0);  // second half of the dTheta tile awaited (step 0 reads the dOut block instead)
          if (u == 0)
            mbar_wait(dzo_ready, t & 1);
          else
            mbar_wait(&a_half[0], (inst0 + u - 1) & 1);
          tc_fence_after();
          const int kbn = (u == 0) ? 1 : S::kKB;
          const int nhn = (u <= L) ? NH : NX;
          for (int nh = 0; nh < nhn; ++nh) {
            for (int kb = 0; kb < kbn; ++kb, ++c) {
              if (!second && (nh > 0 || kb >= S::kKB / 2)) {
                mbar_wait(&a_half[1], (inst0 + u - 1) & 1);
                second = true;
              }
              const uint32_t slot = c % S::kSlots;
              mbar_wait(&w_full[slot], (c / S::kSlots) & 1);
              tc_fence_after();
              const uint32_t a_blk = (u == 0) ? dzo_base : a_base + kb * S::kABlock;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                umma_bf16_ss_2cta_w(tmem_d + nh * 256, smem_desc(a_blk + k4 * 32, hi),
                                    smem_desc(w_base + slot * S::kSlotBytes + k4 * 32, hi), idesc, (kb | k4) != 0);
              }
              umma_commit_2cta_w(&w_empty[slot]);
            }
          }
          if (!second) mbar_wait(&a_half[1], (inst0 + u - 1) & 1);
          umma_commit_2cta_w(d_full);
        }
        // without the input-gradient step nobody reads the last tile (dTheta_0) through these barriers' final phase;
        // consume it so that the next tile's waits stay one phase behind at most
        if (!dx) {
          mbar_wait(&a_half[0], (inst0 + L) & 1);
          mbar_wait(&a_half[1], (inst0 + L) & 1);
        }
      }
    } else if (lane == 0) {
      // =============================== weight relay (pair peer) ===============================
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int u = 0; u <= L + dx; ++u) {
          const int nchunks = (u == 0) ? NH : (u <= L ? NH : NX) * S::kKB;
          for (int j = 0; j < nchunks; ++j, ++c) {
            const uint32_t slot = c % S::kSlots;
            mbar_wait(&w_full[slot], (c / S::kSlots) & 1);
            mbar_arrive_peer(&w_full[slot], 0);
          }
        }
      }
    }
  } else if (warp == 2) {
    // =============================== stash store ===============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = tile_of(t);
        mbar_wait(dzo_loc, t & 1);
        if (!p.lean) {
          bulk_s2g(p.stash_dzo + size_t(tile) * (kTileRows * 128), dzo_smem, kTileRows * 128);
          bulk_commit();
          bulk_wait_read0();
        }
        mbar_arrive(dzo_free);
        for (int l = L; l >= 0; --l, ++n) {
          mbar_wait(&a_loc[0], n & 1);
          mbar_wait(&a_loc[1], n & 1);
          if (!p.lean) {
            bulk_s2g(p.stash_dz + size_t(l) * p.layer_stride + size_t(tile) * S::kABytes, a_smem, S::kABytes);
            bulk_commit();
            bulk_wait_read0();
          }
          mbar_arrive(a_free);
        }
      }
      bulk_wait0();
    }
  } else if (warp >= kGenBwdFirstEpiWarp) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;
    const int s = (warp - kGenBwdFirstEpiWarp) >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = uint32_t(q * 32) << 16;
    const uint32_t a_addr = smem_u32(a_smem);
    const uint32_t dzo_addr = smem_u32(dzo_smem);
    const int C = g.C;
    uint32_t n = 0, nf = 0;
    // hand-offs: the leader's MMA issuer counts both CTAs' warps, the local barrier feeds this CTA's stash thread
    auto arrive2 = [&](uint64_t* pair_bar, uint64_t* local_bar) {
      mbar_arrive(local_bar);
      if (leader)
        mbar_arrive(pair_bar);
      else
        mbar_arrive_peer(pair_bar, 0);
    };
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = tile_of(t);
      const long long row0 = (long long)tile * kTileRows;

      // ---- dOut tile -> bf16 [128][64] block
      if (t > 0) mbar_wait(dzo_free, (t - 1) & 1);
      {
        const bool valid = (row0 + r) < p.rows;
        const float* gp = p.grad_out + (row0 + r) * C;
        const float* yp = p.out_y ? p.out_y + (row0 + r) * C : nullptr;
        const float inv_s = p.out_y ? 1.0f / p.out_tanh : 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ch = 2 * s + cc;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = ch * 8 + j;
            v[j] = (valid && col < C) ? gp[col] : 0.f;
            if (yp != nullptr && valid && col < C) {  // d(s tanh z)/dz = s (1 - tanh^2) = s - y^2 / s
              const float y = yp[col];
              v[j] *= p.out_tanh - y * y * inv_s;
            }
          }
          sts128(dzo_addr + sw128_chunk_off(r, ch),
                 make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                            pack_bf16x2(v[6], v[7])));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive2(dzo_ready, dzo_loc);
      }

      // ---- dTheta_l = dY_l .* act'(theta_l), l = L .. 0
      for (int l = L; l >= 0; --l) {
        // derivative source of this thread's row: 16-bit phases (sine) or the bf16 outputs themselves (ReLU)
        const uint8_t* src_l = (ACT == B200INR_ACT_SINE)
                                   ? p.stash_ph + size_t(l) * p.layer_stride + size_t(tile) * S::kABytes +
                                         size_t(r) * 16 + size_t(2 * s) * (kTileRows * 16)
                                   : p.stash_y + size_t(l) * p.layer_stride + size_t(tile) * S::kABytes;
        auto load_src = [&](int kb, uint4 (&dst)[2]) {
          if (ACT == B200INR_ACT_SINE) {
            dst[0] = *reinterpret_cast<const uint4*>(src_l + size_t(kb * 8) * (kTileRows * 16));
            dst[1] = *reinterpret_cast<const uint4*>(src_l + size_t(kb * 8 + 1) * (kTileRows * 16));
          } else {
            dst[0] = *reinterpret_cast<const uint4*>(src_l + size_t(kb) * S::kABlock + sw128_chunk_off(r, 2 * s));
            dst[1] = *reinterpret_cast<const uint4*>(src_l + size_t(kb) * S::kABlock + sw128_chunk_off(r, 2 * s + 1));
          }
        };
        uint4 sv[2], svn[2];
        load_src(0, svn);
        mbar_wait(d_full, n & 1);
        ++n;
        if (nf > 0) mbar_wait(a_free, (nf - 1) & 1);
        ++nf;
        tc_fence_after();
        const uint32_t d_addr = tmem_d + t_lane + s * 16;
        uint32_t v[16], vn[16];
        tmem_ld16(d_addr, vn);
#pragma unroll 2
        for (int kb = 0; kb < S::kKB; ++kb) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = vn[j];
          sv[0] = svn[0];
          sv[1] = svn[1];
          if (kb + 1 < S::kKB) {
            tmem_ld16(d_addr + (kb + 1) * 64, vn);
            load_src(kb + 1, svn);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t pw[4] = {sv[c].x, sv[c].y, sv[c].z, sv[c].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float d0 = __uint_as_float(v[c * 8 + 2 * j]);
              float d1 = __uint_as_float(v[c * 8 + 2 * j + 1]);
              if (ACT == B200INR_ACT_SINE) {
                d0 *= gen_cos_from_phase(pw[j] & 0xFFFFu);
                d1 *= gen_cos_from_phase(pw[j] >> 16);
              } else if (ACT == B200INR_ACT_TANH) {  // 1 - y^2 from the stashed bf16 outputs
                const float y0 = bf16lo(pw[j]), y1 = bf16hi(pw[j]);
                d0 *= fmaf(-y0, y0, 1.0f);
                d1 *= fmaf(-y1, y1, 1.0f);
              } else {  // bf16 y > 0  <=>  sign bit clear and magnitude non-zero
                d0 = ((pw[j] & 0x8000u) == 0u && (pw[j] & 0x7FFFu) != 0u) ? d0 : 0.f;
                d1 = ((pw[j] & 0x80000000u) == 0u && (pw[j] & 0x7FFF0000u) != 0u) ? d1 : 0.f;
              }
              o[j] = pack_bf16x2(d0, d1);
            }
            sts128(a_addr + kb * S::kABlock + sw128_chunk_off(r, 2 * s + c), make_uint4(o[0], o[1], o[2], o[3]));
          }
          if (NH == 2 && kb == S::kKB / 2 - 1) {  // blocks 0..kKB/2-1 written, D[0:256) read out
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive2(&a_half[0], &a_loc[0]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NH != 2) arrive2(&a_half[0], &a_loc[0]);
          arrive2(&a_half[1], &a_loc[1]);
        }
      }

      // ---- input gradient: D = dTheta_0 (omega_0 W_0), fp32, straight to global memory (row r, 16 columns per block)
      if (dx && p.grad_coords != nullptr) {
        // ---- adjoint of input_mapping on the fly: this thread owns row r and 16 of every 64 feature columns
        mbar_wait(d_full, n & 1);
        ++n;
        tc_fence_after();
        long long row = row0 + r;
        if (row >= p.rows) row = p.rows - 1;
        float x[4] = {0.f, 0.f, 0.f, 0.f}, xs[4];
        if (p.coords != nullptr) {
          for (int j = 0; j < g.d; ++j) x[j] = p.coords[row * g.d + j];
        } else {
          grid_coords(p.grid, row0 + r, x);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) xs[j] = __fmul_rn(6.283185307179586f, x[j]);
        const float4* bmat_g = reinterpret_cast<const float4*>(p.packed + p.pl.bmat);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int kb0 = g.K0 / 64;
#pragma unroll 1
        for (int kb = 0; kb < kb0; ++kb) {
          uint32_t v[16];
          tmem_ld16(tmem_d + t_lane + kb * 64 + s * 16, v);
          tmem_ld_wait();
          const int col0 = kb * 64 + s * 16;
          const bool is_sin = col0 < g.m;
          const int k0 = is_sin ? col0 : col0 - g.m;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bk = __ldg(bmat_g + k0 + j);
            float pr = xs[0] * bk.x;
            pr = fmaf(xs[1], bk.y, pr);
            pr = fmaf(xs[2], bk.z, pr);
            pr = fmaf(xs[3], bk.w, pr);
            float sn, cs;
            __sincosf(pr, &sn, &cs);
            const float dp = __uint_as_float(v[j]) * (is_sin ? cs : -sn);
            acc[0] = fmaf(dp, bk.x, acc[0]);
            acc[1] = fmaf(dp, bk.y, acc[1]);
            acc[2] = fmaf(dp, bk.z, acc[2]);
            acc[3] = fmaf(dp, bk.w, acc[3]);
          }
        }
        tc_fence_before();
        float4* red = reinterpret_cast<float4*>(smem + S::kOffRed);
        red[s * kTileRows + r] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        named_bar_sync(1, kGenBwdEpiWarps * 32);
        if (s == 0 && (row0 + r) < p.rows) {
          const float4 a0 = red[r], a1 = red[kTileRows + r], a2 = red[2 * kTileRows + r], a3 = red[3 * kTileRows + r];
          const float tot[4] = {(a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                                (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w)};
          for (int j = 0; j < g.d; ++j) p.grad_coords[(row0 + r) * g.d + j] = 6.283185307179586f * tot[j];
        }
      } else if (dx) {
        mbar_wait(d_full, n & 1);
        ++n;
        tc_fence_after();
        const bool valid = (row0 + r) < p.rows;
        float* gi = p.grad_in + (row0 + r) * (long long)g.K0;
        const int kb0 = g.K0 / 64;
#pragma unroll 1
        for (int kb = 0; kb < kb0; ++kb) {
          uint32_t v[16];
          tmem_ld16(tmem_d + t_lane + kb * 64 + s * 16, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(gi + kb * 64 + s * 16 + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          }
        }
        tc_fence_before();
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees tensor memory) while its peer may still address it
  if (warp == 1) tmem_dealloc_2cta<512>(tmem_d);
}

template <int H, int ACT>
static int launch_gen_bwd_t(const GenBwdParams& p, int grid_x, cudaStream_t stream) {
  const int smem = GenBwdSmem<H>::kBytes + 1024;
  if (cudaFuncSetAttribute(gen_bwd_kernel<H, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid_x));
  cfg.blockDim = dim3(kGenBwdThreads);
  cfg.dynamicSmemBytes = size_t(smem);
  cfg.stream = stream;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = 2;
  at.val.clusterDim.y = 1;
  at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, gen_bwd_kernel<H, ACT>, p) != cudaSuccess) return B200INR_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

int launch_gen_bwd(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                   float* grad_in, int num_sms, cudaStream_t stream, float* grad_coords, const float* coords,
                   const b200inr_grid* grid, const float* out_y) {
  GenBwdParams p{};
  p.grad_in = grad_in;
  p.grad_coords = grad_coords;
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
  p.out_y = (net->flags & B200INR_NET_TANH_OUT) ? out_y : nullptr;
  p.out_tanh = net->scale_0;
  p.lean = (net->flags & B200INR_NET_DGRAD_ONLY) ? 1 : 0;
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.g = make_gen_dims(net);
  p.pl = make_gen_pack_layout(p.g);
  p.rows = rows;
  p.num_tiles = int((rows + kTileRows - 1) / kTileRows);
  p.grad_out = grad_out;
  const GenStashLayout sl = make_gen_stash_layout(p.g, rows);
  uint8_t* st = reinterpret_cast<uint8_t*>(stash);
  p.stash_y = st + sl.y;
  p.stash_ph = st + sl.ph;
  p.stash_dz = st + sl.dz;
  p.stash_dzo = st + sl.dzo;
  p.layer_stride = sl.layer_stride;
  // persistent CTA pairs (clusters of 2) walk tile pairs
  const int tile_pairs_h = (p.num_tiles + 1) / 2;
  int grid_x = 2 * (tile_pairs_h < num_sms / 2 ? tile_pairs_h : num_sms / 2);
  if (grid_x < 2) grid_x = 2;
  const bool sine = net->activation == B200INR_ACT_SINE;
  if (net->activation == B200INR_ACT_TANH) return launch_gen_bwd_t<256, B200INR_ACT_TANH>(p, grid_x, stream);
  if (p.g.H == 256)
    return sine ? launch_gen_bwd_t<256, B200INR_ACT_SINE>(p, grid_x, stream)
                : launch_gen_bwd_t<256, B200INR_ACT_RELU>(p, grid_x, stream);
  return sine ? launch_gen_bwd_t<512, B200INR_ACT_SINE>(p, grid_x, stream)
              : launch_gen_bwd_t<512, B200INR_ACT_RELU>(p, grid_x, stream);
}

}  // namespace b200inr
