// wire.cu -- WIRE (complex Gabor wavelet) coordinate network for sm_100a: operand packing and fused forward.
//
// Replaces the WIRE network of the reference (INR/wiretest.ipynb cell 2: Sequential(ComplexGaborLayer2D x (L+1),
// complex nn.Linear), `.real` returned; layer math INR/INRmodel.py:109-120):
//     lin = h W1^T + b1,  orth = h W2^T + b2               (complex; real in the first layer)
//     h'  = exp(1j w0 lin) * exp(-s0^2 (|lin|^2 + |orth|^2))
// as a REAL block GEMM chain: activations are [h_r | h_i] (K = 2H reals), the weights of a
// hidden layer form one [4H x 2H] real matrix whose output column n = 4u + {0,1,2,3} is (a, b, c, d) =
// (Re lin, Im lin, Re orth, Im orth) of unit u, so one tcgen05 accumulator row holds everything a unit needs:
//     E = exp(-w0 b - s0^2 (a^2 + b^2 + c^2 + d^2)),   h'_r = E cos(w0 a),   h'_i = E sin(w0 a).
// H = 128 complex units: K = 256, N = 512 fp32 columns = all of TMEM.  First layer (d <= 4 real inputs) on CUDA
// cores; final layer out = h_r Re(W_f)^T - h_i Im(W_f)^T + Re(b_f) as a 128 x 32 x 256 MMA.
//
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer + TMEM owner, warp 2 = stash store (training),
//             warps 3..18 = epilogue (TMEM lane quadrant = warp & 3, 16-column slice = (warp - 3) >> 2).
#include <stdio.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

// ------------------------------------------------------------------ packing
struct WirePackParams {
  const float* params;
  uint8_t* packed;
  WireDims w;
  WirePackLayout pl;
  long long off[4 * kMaxSineLayers + 3];
};

__device__ __forceinline__ void wire_put_bf16(uint8_t* base, uint32_t row, uint32_t k, float v) {
  *reinterpret_cast<__nv_bfloat16*>(base + sw128_chunk_off(row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256) wire_pack_kernel(const WirePackParams p) {
  const int H = p.w.H, L = p.w.L, C = p.w.C, d = p.w.d;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  // first layer (real)
  const int K0 = p.w.K0;
  for (long long i = tid; i < 2 * H; i += nthreads) {
    const int u = int(i >> 1), which = int(i & 1);  // 0: lin, 1: orth
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (K0 == 0)
      for (int j = 0; j < d; ++j) wv[j] = p.params[p.off[2 * which] + (long long)u * d + j];
    reinterpret_cast<float4*>(p.packed + p.pl.w0)[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    reinterpret_cast<float*>(p.packed + p.pl.b0)[i] = p.params[p.off[2 * which + 1] + u];
  }
  if (K0) {
    // feature-fed first layer on the tensor cores: forward operand N = 2u + {lin, orth}, K = feature;
    // input-gradient operand N = feature (whole 256-row chunks, zero padded), K = 4u + {a, b, c, d} (b, d rows zero)
    for (long long i = tid; i < (long long)2 * H * K0; i += nthreads) {
      const int n = int(i / K0), k = int(i % K0);
      const float v = p.params[p.off[2 * (n & 1)] + (long long)(n >> 1) * K0 + k];
      wire_put_bf16(p.packed + p.pl.w0f + size_t(k >> 6) * kGenChunkBytes, n, k & 63, v);
    }
    const int K0p = ((K0 + 255) / 256) * 256;
    for (long long i = tid; i < (long long)4 * H * K0p; i += nthreads) {
      const int n = int(i / K0p), k = int(i % K0p);  // n = 4u + comp
      const int comp = n & 3;
      float v = 0.f;
      if ((comp & 1) == 0 && k < K0) v = p.params[p.off[comp] + (long long)(n >> 2) * K0 + k];  // comp 0: lin, 2: orth
      wire_put_bf16(p.packed + p.pl.wt0 + size_t((k >> 8) * 8 + (n >> 6)) * kGenChunkBytes, k & 255, n & 63, v);
    }
    for (long long i = tid; i < p.w.m; i += nthreads) {
      float wv[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < d; ++j) wv[j] = p.params[p.off[4 * (L + 1) + 2] + i * d + j];
      reinterpret_cast<float4*>(p.packed + p.pl.bmat)[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    }
  }
  // hidden layers: biases in column order, real-block weights in both orientations
  float* bias = reinterpret_cast<float*>(p.packed + p.pl.bias);
  for (long long i = tid; i < (long long)L * 4 * H + 32; i += nthreads) {
    float v = 0.f;
    if (i < (long long)L * 4 * H) {
      const int l = int(i / (4 * H)) + 1, n = int(i % (4 * H));
      const int u = n >> 2, comp = n & 3;
      v = p.params[p.off[4 * l + 2 * (comp >> 1) + 1] + 2 * u + (comp & 1)];
    } else {
      const int c = int(i - (long long)L * 4 * H);
      if (c < C) v = p.params[p.off[4 * (L + 1) + 1] + 2 * c];  // Re b_f
    }
    bias[i] = v;
  }
  const long long per_layer = (long long)4 * H * 2 * H;
  for (long long i = tid; i < (long long)L * per_layer; i += nthreads) {
    const int l = int(i / per_layer) + 1;
    const int n = int((i % per_layer) / (2 * H));  // output column 4u + comp
    const int k = int(i % (2 * H));                // input real: k < H: h_r[k], else h_i[k - H]
    const int u = n >> 2, comp = n & 3;
    const float* W = p.params + p.off[4 * l + 2 * (comp >> 1)];  // lin (comp 0,1) or orth (comp 2,3) weights
    const int kk = k < H ? k : k - H;
    const float wr = W[((long long)u * H + kk) * 2], wi = W[((long long)u * H + kk) * 2 + 1];
    // Re out = h_r wr - h_i wi ; Im out = h_r wi + h_i wr
    const float v = (comp & 1) == 0 ? (k < H ? wr : -wi) : (k < H ? wi : wr);
    uint8_t* wf = p.packed + p.pl.w + size_t(l - 1) * 8 * kGenChunkBytes;
    uint8_t* wt = p.packed + p.pl.wt + size_t(l - 1) * 8 * kGenChunkBytes;
    wire_put_bf16(wf + size_t((n >> 8) * 4 + (k >> 6)) * kGenChunkBytes, n & 255, k & 63, v);  // N = n, K = k
    wire_put_bf16(wt + size_t(n >> 6) * kGenChunkBytes, k, n & 63, v);                          // N = k, K = n
  }
  // final linear (real part only)
  for (long long i = tid; i < (long long)kOutPad * 2 * H; i += nthreads) {
    const int c = int(i / (2 * H)), k = int(i % (2 * H));
    float v = 0.f;
    if (c < C) {
      const int kk = k < H ? k : k - H;
      const float* W = p.params + p.off[4 * (L + 1)];
      v = k < H ? W[((long long)c * H + kk) * 2] : -W[((long long)c * H + kk) * 2 + 1];
    }
    wire_put_bf16(p.packed + p.pl.wf + size_t(k >> 6) * kOutPad * 128, c, k & 63, v);
    if (c < kDzoPad) wire_put_bf16(p.packed + p.pl.wft, k, c, v);
  }
  for (long long i = tid; i < (long long)2 * H * (kDzoPad - kOutPad); i += nthreads) {  // zero the padded channels
    const int k = int(i / (kDzoPad - kOutPad)), c = kOutPad + int(i % (kDzoPad - kOutPad));
    wire_put_bf16(p.packed + p.pl.wft, k, c, 0.f);
  }
}

int launch_wire_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream) {
  WirePackParams p{};
  p.params = params;
  p.packed = reinterpret_cast<uint8_t*>(packed);
  p.w = make_wire_dims(net);
  p.pl = make_wire_pack_layout(p.w);
  int64_t off[4 * kMaxSineLayers + 3] = {0};
  wire_param_offsets(p.w, off);
  for (int i = 0; i < 4 * (p.w.L + 1) + 3; ++i) p.off[i] = off[i];
  wire_pack_kernel<<<592, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

// ------------------------------------------------------------------ forward
constexpr int kWireEpiWarps = 16;
constexpr int kWireFirstEpiWarp = 3;
constexpr int kWireThreads = (kWireFirstEpiWarp + kWireEpiWarps) * 32;
constexpr int kWireEpiThreads = kWireEpiWarps * 32;
constexpr uint32_t kWireEpiBarId = 1;
constexpr int kWireH = 128;                 // complex units
constexpr int kWireKB = 2 * kWireH / 64;    // 4 K blocks of the activation tile
constexpr int kWireABlock = kTileRows * 128;
constexpr int kWireABytes = kWireKB * kWireABlock;  // 64 KB
constexpr int kWireSlots = 4;

struct WireFwdParams {
  const uint8_t* packed;
  WireDims w;
  WirePackLayout pl;
  const float* coords;
  GridDesc grid;
  long long rows;
  int num_tiles;
  float* out;
  int clamp;
  float clamp_min;
  uint8_t* stash_y;  // nullptr => inference
  uint8_t* stash_z;
  uint8_t* stash_xa;
  uint8_t* stash_ain;  // feature-fed networks: the network input tiles (bf16)
  size_t stride_y, stride_z, tile_z, tile_in;
};

struct WireSmem {
  static constexpr int kOffA = 0;
  static constexpr int kOffW = kWireABytes;
  static constexpr int kOffXa = kOffW + kWireSlots * kGenChunkBytes;
  static constexpr int kOffBar = kOffXa + kTileRows * 128;
  static constexpr int kBytes = kOffBar + 256;
};

constexpr float kLog2e = 1.4426950408889634f;

// Gabor wavelet of one unit: (a, b, c, d) -> (h_r, h_i)
__device__ __forceinline__ void gabor(float a, float b, float c, float d, float omega, float s2, float& hr,
                                      float& hi) {
  const float q = fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d)));
  const float e = exp2f(kLog2e * (-omega * b - s2 * q));
  float sn, cs;
  __sincosf(omega * a, &sn, &cs);
  hr = e * cs;
  hi = e * sn;
}

// kFeat: the first layer reads K0 (<= 512) explicit or in-kernel Fourier features per row and runs on the tensor cores
// (N = 256 columns n = 2u + {lin, orth}); the A tile holds K = 256 columns, so wider inputs are fed in passes of four
// 64-column blocks that accumulate into the same TMEM columns.
template <bool kStash, bool kFeat>
__global__ void __launch_bounds__(kWireThreads, 1) wire_fwd_kernel(const WireFwdParams p) {
  using S = WireSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem + S::kOffA;
  uint8_t* w_smem = smem + S::kOffW;
  uint8_t* xa_smem = smem + S::kOffXa;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + kWireSlots;
  uint64_t* a_ready = bars + 2 * kWireSlots;
  uint64_t* d_full = bars + 2 * kWireSlots + 1;
  uint64_t* a_free = bars + 2 * kWireSlots + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWireSlots + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const WireDims w = p.w;
  const int L = w.L;
  constexpr int H = kWireH;

  if (kStash) {
    for (int i = threadIdx.x; i < kTileRows * 8; i += blockDim.x)
      reinterpret_cast<uint4*>(xa_smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWireSlots; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(a_ready, kWireEpiWarps);
    mbar_init(d_full, 1);
    mbar_init(a_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int my_tiles = (p.num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int KB0 = kFeat ? w.K0 / 64 : 0;
  const int npass = (KB0 + kWireKB - 1) / kWireKB;

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int j = 0; j < KB0; ++j, ++c) {  // feature-fed first layer
          const uint32_t slot = c % kWireSlots, round = c / kWireSlots;
          if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
          mbar_arrive_expect_tx(&w_full[slot], kGenChunkBytes);
          bulk_g2s(w_smem + slot * kGenChunkBytes, p.packed + p.pl.w0f + size_t(j) * kGenChunkBytes, kGenChunkBytes,
                   &w_full[slot]);
        }
        for (int l = 1; l <= L + 1; ++l) {
          const bool hidden = (l <= L);
          const int nchunks = hidden ? 8 : kWireKB;
          const uint8_t* src = hidden ? p.packed + p.pl.w + size_t(l - 1) * 8 * kGenChunkBytes : p.packed + p.pl.wf;
          const uint32_t bytes = hidden ? uint32_t(kGenChunkBytes) : uint32_t(kOutPad * 128);
          for (int j = 0; j < nchunks; ++j, ++c) {
            const uint32_t slot = c % kWireSlots, round = c / kWireSlots;
            if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
            mbar_arrive_expect_tx(&w_full[slot], bytes);
            bulk_g2s(w_smem + slot * kGenChunkBytes, src + size_t(j) * bytes, bytes, &w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    {  // whole warp converged, one elected lane issues (umma_*_w)
      const uint64_t hi = smem_desc_hi_sw128(0, 1024);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t idesc_h = idesc_bf16(128, 256, false, false);
      const uint32_t idesc_f = idesc_bf16(128, kOutPad, false, false);
      uint32_t c = 0, n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int ps = 0; ps < npass; ++ps, ++n) {  // feature-fed first layer, four K blocks per pass
          mbar_wait(a_ready, n & 1);
          tc_fence_after();
          const int nblk = (KB0 - ps * kWireKB) < kWireKB ? (KB0 - ps * kWireKB) : kWireKB;
          for (int kb = 0; kb < nblk; ++kb, ++c) {
            const uint32_t slot = c % kWireSlots;
            mbar_wait(&w_full[slot], (c / kWireSlots) & 1);
            tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t da = smem_desc(a_base + kb * kWireABlock + k4 * 32, hi);
              const uint64_t db = smem_desc(w_base + slot * kGenChunkBytes + k4 * 32, hi);
              umma_bf16_ss_w(tmem_d, da, db, idesc_h, (ps | kb | k4) != 0);
            }
            umma_commit_w(&w_empty[slot]);
          }
          umma_commit_w(d_full);
        }
        for (int l = 1; l <= L + 1; ++l, ++n) {
          mbar_wait(a_ready, n & 1);
          tc_fence_after();
          const bool hidden = (l <= L);
          for (int nh = 0; nh < (hidden ? 2 : 1); ++nh) {
            for (int kb = 0; kb < kWireKB; ++kb, ++c) {
              const uint32_t slot = c % kWireSlots;
              mbar_wait(&w_full[slot], (c / kWireSlots) & 1);
              tc_fence_after();
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t da = smem_desc(a_base + kb * kWireABlock + k4 * 32, hi);
                const uint64_t db = smem_desc(w_base + slot * kGenChunkBytes + k4 * 32, hi);
                umma_bf16_ss_w(tmem_d + nh * 256, da, db, hidden ? idesc_h : idesc_f, (kb | k4) != 0);
              }
              umma_commit_w(&w_empty[slot]);
            }
          }
          umma_commit_w(d_full);
        }
      }
    }
  } else if (warp == 2) {
    // =============================== stash store (training) ===============================
    if (kStash && lane == 0) {
      uint32_t n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = int(blockIdx.x) + t * int(gridDim.x);
        for (int ps = 0; ps < npass; ++ps, ++n) {  // the network input, one pass of feature blocks at a time
          mbar_wait(a_ready, n & 1);
          const int nblk = (KB0 - ps * kWireKB) < kWireKB ? (KB0 - ps * kWireKB) : kWireKB;
          bulk_s2g(p.stash_ain + size_t(tile) * p.tile_in + size_t(ps) * kWireABytes, a_smem,
                   uint32_t(nblk) * kWireABlock);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(a_free);
        }
        for (int l = 0; l <= L; ++l, ++n) {
          mbar_wait(a_ready, n & 1);
          bulk_s2g(p.stash_y + size_t(l) * p.stride_y + size_t(tile) * kWireABytes, a_smem, kWireABytes);
          if (l == 0 && !kFeat) bulk_s2g(p.stash_xa + size_t(tile) * (kTileRows * 128), xa_smem, kTileRows * 128);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(a_free);
        }
      }
      bulk_wait0();
    }
  } else if (warp >= kWireFirstEpiWarp) {
    // =============================== epilogue warps ===============================
    const int et = threadIdx.x - kWireFirstEpiWarp * 32;
    const int q = warp & 3;
    const int s = (warp - kWireFirstEpiWarp) >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = uint32_t(q * 32) << 16;
    const uint32_t a_addr = smem_u32(a_smem);
    const float4* w0_g = reinterpret_cast<const float4*>(p.packed + p.pl.w0);
    const float2* b0_g = reinterpret_cast<const float2*>(p.packed + p.pl.b0);
    const float* bias_g = reinterpret_cast<const float*>(p.packed + p.pl.bias);
    const float s2 = w.s0 * w.s0;
    uint32_t n = 0, nf = 0;
    // store h_r (column u) and h_i (column H + u) of 4 consecutive units u0 .. u0+3 (u0 % 4 == 0) of row r
    auto store_units = [&](int u0, const float (&hr)[4], const float (&hi)[4]) {
      const uint32_t off_r = uint32_t(u0 >> 6) * kWireABlock + sw128_chunk_off(r, (u0 & 63) >> 3) + (u0 & 7) * 2;
      const uint32_t off_i = uint32_t((H + u0) >> 6) * kWireABlock + sw128_chunk_off(r, ((H + u0) & 63) >> 3) +
                             ((H + u0) & 7) * 2;
      sts32(a_addr + off_r, pack_bf16x2(hr[0], hr[1]));
      sts32(a_addr + off_r + 4, pack_bf16x2(hr[2], hr[3]));
      sts32(a_addr + off_i, pack_bf16x2(hi[0], hi[1]));
      sts32(a_addr + off_i + 4, pack_bf16x2(hi[2], hi[3]));
    };
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = int(blockIdx.x) + t * int(gridDim.x);
      const long long row0 = (long long)tile * kTileRows;
      uint8_t* z_row = kStash ? p.stash_z + size_t(tile) * p.tile_z + size_t(r) * 16 : nullptr;

      if (kFeat) {
        // ---- feature-fed first Gabor layer: features -> A operand (passes of four blocks), MMA, then the epilogue
        long long row = row0 + r;
        if (row >= p.rows) row = p.rows - 1;
        float xs[4] = {0.f, 0.f, 0.f, 0.f};
        if (w.in_mode == B200INR_IN_FOURIER) {
          float x[4] = {0.f, 0.f, 0.f, 0.f};
          if (p.coords != nullptr) {
            for (int j = 0; j < w.d; ++j) x[j] = p.coords[row * w.d + j];
          } else {
            grid_coords(p.grid, row0 + r, x);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) xs[j] = __fmul_rn(6.283185307179586f, x[j]);  // (2 pi x) first, like the reference
        }
        const float4* bmat_g = reinterpret_cast<const float4*>(p.packed + p.pl.bmat);
        for (int ps = 0; ps < npass; ++ps) {
          if (ps > 0) {  // the previous pass has been consumed by its MMAs (and stored, when training)
            mbar_wait(d_full, n & 1);
            ++n;
            if (kStash) {
              mbar_wait(a_free, nf & 1);
              ++nf;
            }
          }
          const int nblk = (KB0 - ps * kWireKB) < kWireKB ? (KB0 - ps * kWireKB) : kWireKB;
          for (int kb = 0; kb < nblk; ++kb) {
            const int col0 = (ps * kWireKB + kb) * 64 + s * 16;
            float v[16];
            if (w.in_mode == B200INR_IN_FOURIER) {
              const bool is_sin = col0 < w.m;  // m is a multiple of 16: a 16-column slice never straddles sin | cos
              const int k0 = is_sin ? col0 : col0 - w.m;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 bk = __ldg(bmat_g + k0 + j);
                float pr = xs[0] * bk.x;
                pr = fmaf(xs[1], bk.y, pr);
                pr = fmaf(xs[2], bk.z, pr);
                pr = fmaf(xs[3], bk.w, pr);
                v[j] = is_sin ? __sinf(pr) : __cosf(pr);
              }
            } else {
              const float4* f = reinterpret_cast<const float4*>(p.coords + row * w.K0 + col0);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 x4 = __ldg(f + j4);
                v[j4 * 4 + 0] = x4.x; v[j4 * 4 + 1] = x4.y; v[j4 * 4 + 2] = x4.z; v[j4 * 4 + 3] = x4.w;
              }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c)
              sts128(a_addr + kb * kWireABlock + sw128_chunk_off(r, 2 * s + c),
                     make_uint4(pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]), pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                                pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7])));
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready);
        }
        // epilogue of the first layer: D[:, 0:256), column n = 2u + {lin, orth}
        mbar_wait(d_full, n & 1);
        ++n;
        if (kStash) {
          mbar_wait(a_free, nf & 1);
          ++nf;
        }
        tc_fence_after();
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {  // 64-column groups: columns 64 g + 16 s .. + 15 = units 32 g + 8 s .. + 7
          uint32_t v[16];
          tmem_ld16(tmem_d + t_lane + g * 64 + s * 16, v);
          tmem_ld_wait();
          const int u0 = 32 * g + 8 * s;
          float av[8], cv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 bb = __ldg(b0_g + u0 + j);
            av[j] = __uint_as_float(v[2 * j]) + bb.x;
            cv[j] = __uint_as_float(v[2 * j + 1]) + bb.y;
          }
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            float hr[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) gabor(av[4 * h4 + j], 0.f, cv[4 * h4 + j], 0.f, w.omega0, s2, hr[j], hi[j]);
            store_units(u0 + 4 * h4, hr, hi);
          }
          if (kStash) {  // pre-activations (a, 0, c, 0), 8 columns (2 units) per 16-byte chunk
#pragma unroll
            for (int h2 = 0; h2 < 4; ++h2)
              *reinterpret_cast<uint4*>(z_row + size_t(u0 / 2 + h2) * (kTileRows * 16)) =
                  make_uint4(pack_bf16x2(av[2 * h2], 0.f), pack_bf16x2(cv[2 * h2], 0.f),
                             pack_bf16x2(av[2 * h2 + 1], 0.f), pack_bf16x2(cv[2 * h2 + 1], 0.f));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      } else
      // ---- first Gabor layer (real inputs) on CUDA cores: this thread owns units 32 s .. 32 s + 31 of its row
      {
        float x[4];
        if (p.coords != nullptr) {
          long long row = row0 + r;
          if (row >= p.rows) row = p.rows - 1;
          x[0] = x[1] = x[2] = x[3] = 0.0f;
          for (int j = 0; j < w.d; ++j) x[j] = p.coords[row * w.d + j];
        } else {
          grid_coords(p.grid, row0 + r, x);
        }
        if (kStash && s == 0) {
          float hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            hi[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
            lo[j] = x[j] - hi[j];
          }
          sts128(smem_u32(xa_smem) + sw128_chunk_off(r, 0),
                 make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(lo[0], lo[1]),
                            pack_bf16x2(lo[2], lo[3])));
        }
#pragma unroll 1
        for (int g4 = 0; g4 < 8; ++g4) {
          const int u0 = 32 * s + 4 * g4;
          float hr[4], hi[4], av[4], cv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 wl = __ldg(w0_g + 2 * (u0 + j)), wo = __ldg(w0_g + 2 * (u0 + j) + 1);
            const float2 bb = __ldg(b0_g + u0 + j);
            const float a = fmaf(x[0], wl.x, fmaf(x[1], wl.y, fmaf(x[2], wl.z, fmaf(x[3], wl.w, bb.x))));
            const float c = fmaf(x[0], wo.x, fmaf(x[1], wo.y, fmaf(x[2], wo.z, fmaf(x[3], wo.w, bb.y))));
            av[j] = a;
            cv[j] = c;
            gabor(a, 0.f, c, 0.f, w.omega0, s2, hr[j], hi[j]);
          }
          store_units(u0, hr, hi);
          if (kStash) {  // pre-activations (a, 0, c, 0), 8 columns (2 units) per 16-byte chunk
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
              *reinterpret_cast<uint4*>(z_row + size_t((4 * u0) / 8 + h2) * (kTileRows * 16)) =
                  make_uint4(pack_bf16x2(av[2 * h2], 0.f), pack_bf16x2(cv[2 * h2], 0.f),
                             pack_bf16x2(av[2 * h2 + 1], 0.f), pack_bf16x2(cv[2 * h2 + 1], 0.f));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }

      // ---- hidden Gabor layers
      for (int l = 1; l <= L; ++l) {
        mbar_wait(d_full, n & 1);
        ++n;
        if (kStash) {
          mbar_wait(a_free, nf & 1);
          ++nf;
        }
        tc_fence_after();
        const float* bl = bias_g + size_t(l - 1) * 4 * H;
        const uint32_t d_addr = tmem_d + t_lane + s * 16;
        uint8_t* z_l = kStash ? z_row + size_t(l) * p.stride_z : nullptr;
        uint32_t v[16], vn[16];
        tmem_ld16(d_addr, vn);
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {  // 64-column groups: columns 64 g + 16 s .. + 15 = units 16 g + 4 s .. + 3
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = vn[j];
          if (g + 1 < 8) tmem_ld16(d_addr + (g + 1) * 64, vn);
          const int col0 = g * 64 + s * 16;
          float z[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bl + col0 + j4 * 4));
            z[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + b.x;
            z[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + b.y;
            z[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + b.z;
            z[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + b.w;
          }
          float hr[4], hi[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) gabor(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3], w.omegah, s2, hr[j], hi[j]);
          store_units(col0 >> 2, hr, hi);
          if (kStash) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
              *reinterpret_cast<uint4*>(z_l + size_t(col0 / 8 + h2) * (kTileRows * 16)) =
                  make_uint4(pack_bf16x2(z[8 * h2], z[8 * h2 + 1]), pack_bf16x2(z[8 * h2 + 2], z[8 * h2 + 3]),
                             pack_bf16x2(z[8 * h2 + 4], z[8 * h2 + 5]), pack_bf16x2(z[8 * h2 + 6], z[8 * h2 + 7]));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }

      // ---- final linear (real part): D[:, 0:32) + Re b_f -> out
      {
        mbar_wait(d_full, n & 1);
        ++n;
        if (kStash) {
          mbar_wait(a_free, nf & 1);
          ++nf;
        }
        tc_fence_after();
        const int C = w.C;
        if (s == 0) {
          uint32_t v[32];
          tmem_ld32(tmem_d + t_lane, v);
          tmem_ld_wait();
          const float* bf = bias_g + size_t(L) * 4 * H;
#pragma unroll
          for (int c = 0; c < kOutPad; ++c) {
            if (c < C) {
              float o = __uint_as_float(v[c]) + __ldg(bf + c);
              if (p.clamp) o = fmaxf(o, p.clamp_min);
              sts32(a_addr + uint32_t(r * C + c) * 4, __float_as_uint(o));
            }
          }
        }
        tc_fence_before();
        named_bar_sync(kWireEpiBarId, kWireEpiThreads);
        long long valid = p.rows - row0;
        if (valid > kTileRows) valid = kTileRows;
        const int nout = int(valid) * C;
        float* dst = p.out + row0 * C;
        for (int i = et; i < nout; i += kWireEpiThreads) dst[i] = __uint_as_float(lds32(a_addr + uint32_t(i) * 4));
        named_bar_sync(kWireEpiBarId, kWireEpiThreads);
      }
    }
  }

  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_d);
}

int launch_wire_fwd(const b200inr_net* net, const void* packed, const float* coords, const b200inr_grid* grid,
                    int64_t rows, float* out, int clamp, float clamp_min, void* stash, int num_sms,
                    cudaStream_t stream) {
  WireFwdParams p{};
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.w = make_wire_dims(net);
  p.pl = make_wire_pack_layout(p.w);
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
  p.rows = rows;
  p.num_tiles = int((rows + kTileRows - 1) / kTileRows);
  p.out = out;
  p.clamp = clamp;
  p.clamp_min = clamp_min;
  if (stash) {
    const WireStashLayout sl = make_wire_stash_layout(p.w, rows);
    uint8_t* st = reinterpret_cast<uint8_t*>(stash);
    p.stash_y = st + sl.y;
    p.stash_z = st + sl.z;
    p.stash_xa = st + sl.xa;
    p.stash_ain = st + sl.ain;
    p.stride_y = sl.stride_y;
    p.stride_z = sl.stride_z;
    p.tile_z = sl.tile_z;
    p.tile_in = sl.tile_in;
  }
  const int smem = WireSmem::kBytes + 1024;
  const int grid_x = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  void (*kern)(const WireFwdParams) =
      p.w.K0 ? (stash ? wire_fwd_kernel<true, true> : wire_fwd_kernel<false, true>)
             : (stash ? wire_fwd_kernel<true, false> : wire_fwd_kernel<false, false>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  kern<<<grid_x, kWireThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr

// =====================================================================================================================
// Backward: activation-gradient chain of the WIRE network.
// With real upstream gradients (G_r, G_i) of a unit's output (h_r, h_i) and S = G_r h_r + G_i h_i:
//     da = -2 s0^2 a S + w0 (G_i h_r - G_r h_i)      db = (-w0 - 2 s0^2 b) S
//     dc = -2 s0^2 c S                               dd = -2 s0^2 d S
// (first layer: b = d = 0 and their gradients are dropped).  Chain:  dH_L = dOut [Re W_f | -Im W_f],
// dH_{l-1} = dZ_l Wblk_l (128 x 256 x 512 MMA), dZ_l from dH_l and the stashed pre-activations (a, b, c, d).
// Every dZ_l tile (4H wide) and the bf16 dOut tile go to the stash for the wgrad contraction.
namespace b200inr {

constexpr int kWireZBlocks = 4 * kWireH / 64;                // 8 blocks of the dZ tile
constexpr int kWireZBytes = kWireZBlocks * kWireABlock;      // 128 KB
constexpr int kWireBwdSlots = 2;

struct WireBwdParams {
  const uint8_t* packed;
  WireDims w;
  WirePackLayout pl;
  long long rows;
  int num_tiles;
  const float* grad_out;
  const uint8_t* stash_z;
  uint8_t* stash_dz;
  uint8_t* stash_dzo;
  size_t stride_z, tile_z;
  float* grad_in;  // nullptr, or [rows, K0] fp32: dL/d(feature rows) of a feature-fed network (one more chain step,
                   // dX = dZ_0 [W_lin ; W_orth] on the tensor cores -- the PerturbNet phase of wiretest.ipynb cell 10)
};

struct WireBwdSmem {
  static constexpr int kOffA = 0;
  static constexpr int kOffW = kWireZBytes;
  static constexpr int kOffDzo = kOffW + kWireBwdSlots * kGenChunkBytes;
  static constexpr int kOffBar = kOffDzo + kTileRows * 128;
  static constexpr int kBytes = kOffBar + 256;
};

__global__ void __launch_bounds__(kWireThreads, 1) wire_bwd_kernel(const WireBwdParams p) {
  using S = WireBwdSmem;
  constexpr int H = kWireH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem + S::kOffA;
  uint8_t* w_smem = smem + S::kOffW;
  uint8_t* dzo_smem = smem + S::kOffDzo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + kWireBwdSlots;
  uint64_t* a_ready = bars + 2 * kWireBwdSlots;
  uint64_t* dzo_ready = bars + 2 * kWireBwdSlots + 1;
  uint64_t* d_full = bars + 2 * kWireBwdSlots + 2;
  uint64_t* a_free = bars + 2 * kWireBwdSlots + 3;
  uint64_t* dzo_free = bars + 2 * kWireBwdSlots + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWireBwdSlots + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const WireDims w = p.w;
  const int L = w.L;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWireBwdSlots; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(a_ready, kWireEpiWarps);
    mbar_init(dzo_ready, kWireEpiWarps);
    mbar_init(d_full, 1);
    mbar_init(a_free, 1);
    mbar_init(dzo_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int my_tiles = (p.num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int dx = p.grad_in != nullptr ? 1 : 0;  // extra chain step: gradient of the feature rows
  const int NX = (w.K0 + 255) / 256;            // its N, in 256-column halves

  if (warp == 0) {
    if (lane == 0) {  // weight producer: W_f^T (1 chunk), the 8 chunks of Wblk_l^T for l = L .. 1, [first layer^T]
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int u = 0; u <= L + dx; ++u) {
          const int nchunks = (u == 0) ? 1 : (u <= L ? 1 : NX) * kWireZBlocks;
          const uint8_t* src = (u == 0)   ? p.packed + p.pl.wft
                               : (u <= L) ? p.packed + p.pl.wt + size_t(L - u) * 8 * kGenChunkBytes
                                          : p.packed + p.pl.wt0;
          for (int j = 0; j < nchunks; ++j, ++c) {
            const uint32_t slot = c % kWireBwdSlots, round = c / kWireBwdSlots;
            if (round > 0) mbar_wait(&w_empty[slot], (round - 1) & 1);
            mbar_arrive_expect_tx(&w_full[slot], kGenChunkBytes);
            bulk_g2s(w_smem + slot * kGenChunkBytes, src + size_t(j) * kGenChunkBytes, kGenChunkBytes, &w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // MMA issuer: whole warp converged, one elected lane issues (umma_*_w)
      const uint64_t hi = smem_desc_hi_sw128(0, 1024);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t dzo_base = smem_u32(dzo_smem);
      const uint32_t idesc = idesc_bf16(128, 256, false, false);
      uint32_t c = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t inst0 = uint32_t(t) * uint32_t(L + 1);
        for (int u = 0; u <= L + dx; ++u) {
          if (u == 0)
            mbar_wait(dzo_ready, t & 1);
          else
            mbar_wait(a_ready, (inst0 + u - 1) & 1);
          tc_fence_after();
          const int kbn = (u == 0) ? 1 : kWireZBlocks;
          for (int nh = 0; nh < (u <= L ? 1 : NX); ++nh) {
            for (int kb = 0; kb < kbn; ++kb, ++c) {
              const uint32_t slot = c % kWireBwdSlots;
              mbar_wait(&w_full[slot], (c / kWireBwdSlots) & 1);
              tc_fence_after();
              const uint32_t a_blk = (u == 0) ? dzo_base : a_base + kb * kWireABlock;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_bf16_ss_w(tmem_d + nh * 256, smem_desc(a_blk + k4 * 32, hi),
                               smem_desc(w_base + slot * kGenChunkBytes + k4 * 32, hi), idesc, (kb | k4) != 0);
              umma_commit_w(&w_empty[slot]);
            }
          }
          umma_commit_w(d_full);
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {  // stash store
      uint32_t n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = int(blockIdx.x) + t * int(gridDim.x);
        mbar_wait(dzo_ready, t & 1);
        bulk_s2g(p.stash_dzo + size_t(tile) * (kTileRows * 128), dzo_smem, kTileRows * 128);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(dzo_free);
        for (int l = L; l >= 0; --l, ++n) {
          mbar_wait(a_ready, n & 1);
          bulk_s2g(p.stash_dz + size_t(l) * p.stride_z + size_t(tile) * kWireZBytes, a_smem, kWireZBytes);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(a_free);
        }
      }
      bulk_wait0();
    }
  } else if (warp >= kWireFirstEpiWarp) {
    const int q = warp & 3;
    const int s = (warp - kWireFirstEpiWarp) >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = uint32_t(q * 32) << 16;
    const uint32_t a_addr = smem_u32(a_smem);
    const uint32_t dzo_addr = smem_u32(dzo_smem);
    const int C = w.C;
    const float s2 = w.s0 * w.s0;
    uint32_t n = 0, nf = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = int(blockIdx.x) + t * int(gridDim.x);
      const long long row0 = (long long)tile * kTileRows;
      if (t > 0) mbar_wait(dzo_free, (t - 1) & 1);
      {
        const bool valid = (row0 + r) < p.rows;
        const float* gp = p.grad_out + (row0 + r) * C;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ch = 2 * s + cc;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = ch * 8 + j;
            v[j] = (valid && col < C) ? gp[col] : 0.f;
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

      for (int l = L; l >= 0; --l) {
        const float omega = (l == 0) ? w.omega0 : w.omegah;
        // this thread owns units 32 s .. 32 s + 31 of its row: z chunks (2 units each) 16 s .. 16 s + 15
        const uint8_t* z_l = p.stash_z + size_t(l) * p.stride_z + size_t(tile) * p.tile_z + size_t(r) * 16 +
                             size_t(16 * s) * (kTileRows * 16);
        mbar_wait(d_full, n & 1);
        ++n;
        if (nf > 0) mbar_wait(a_free, (nf - 1) & 1);
        ++nf;
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // 16 units per pass
          const int u0 = 32 * s + 16 * half;
          uint32_t gr[16], gi[16];
          tmem_ld16(tmem_d + t_lane + u0, gr);
          tmem_ld16(tmem_d + t_lane + H + u0, gi);
          uint4 zc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            zc[j] = *reinterpret_cast<const uint4*>(z_l + size_t(8 * half + j) * (kTileRows * 16));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // chunk j: units u0 + 2j, u0 + 2j + 1
            const uint32_t zw[4] = {zc[j].x, zc[j].y, zc[j].z, zc[j].w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float a = bf16lo(zw[2 * e]), b = bf16hi(zw[2 * e]);
              const float c = bf16lo(zw[2 * e + 1]), d = bf16hi(zw[2 * e + 1]);
              float hr, hi_;
              gabor(a, b, c, d, omega, s2, hr, hi_);
              const float Gr = __uint_as_float(gr[2 * j + e]), Gi = __uint_as_float(gi[2 * j + e]);
              const float Ssum = fmaf(Gr, hr, Gi * hi_);
              const float m2 = -2.f * s2 * Ssum;
              float da = fmaf(m2, a, omega * (Gi * hr - Gr * hi_));
              float db = fmaf(m2, b, -omega * Ssum);
              float dc = m2 * c;
              float dd = m2 * d;
              if (l == 0) {  // real first layer: no imaginary pre-activations
                db = 0.f;
                dd = 0.f;
              }
              o[2 * e] = pack_bf16x2(da, db);
              o[2 * e + 1] = pack_bf16x2(dc, dd);
            }
            const int ncol = 4 * (u0 + 2 * j);  // first dZ column of this unit pair
            sts128(a_addr + uint32_t(ncol >> 6) * kWireABlock + sw128_chunk_off(r, (ncol & 63) >> 3),
                   make_uint4(o[0], o[1], o[2], o[3]));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }

      if (dx) {  // D[:, 0:K0) = dL/d(feature rows) -> grad_in
        mbar_wait(d_full, n & 1);
        ++n;
        tc_fence_after();
        const bool valid = (row0 + r) < p.rows;
        float* gi = p.grad_in + (row0 + r) * (long long)w.K0;
#pragma unroll 1
        for (int kb = 0; kb < w.K0 / 64; ++kb) {
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

  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_d);
}

int launch_wire_bwd(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                    float* grad_in, int num_sms, cudaStream_t stream) {
  WireBwdParams p{};
  p.grad_in = grad_in;
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.w = make_wire_dims(net);
  p.pl = make_wire_pack_layout(p.w);
  p.rows = rows;
  p.num_tiles = int((rows + kTileRows - 1) / kTileRows);
  p.grad_out = grad_out;
  const WireStashLayout sl = make_wire_stash_layout(p.w, rows);
  uint8_t* st = reinterpret_cast<uint8_t*>(stash);
  p.stash_z = st + sl.z;
  p.stash_dz = st + sl.dz;
  p.stash_dzo = st + sl.dzo;
  p.stride_z = sl.stride_z;
  p.tile_z = sl.tile_z;
  const int smem = WireBwdSmem::kBytes + 1024;
  const int grid_x = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (cudaFuncSetAttribute(wire_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return B200INR_ERR_CUDA;
  wire_bwd_kernel<<<grid_x, kWireThreads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

// ------------------------------------------------------------------ real-block gradients -> complex parameters
// wgrad.cu leaves, per layer, the gradient of the real-block matrix G[n = 4u + comp][k] (and the column sums bs[n])
// in the stash scratch; the complex weight W = W_r + i W_i occupies two positions of the block matrix
// (a = h_r W_r - h_i W_i, b = h_r W_i + h_i W_r), so
//     dW_r[u, k] = G[4u + 0][k] + G[4u + 1][H + k]        dW_i[u, k] = G[4u + 1][k] - G[4u + 0][H + k]
// (same with rows 4u + 2, 4u + 3 for the orth weights); the result is ACCUMULATED into the flat parameter gradient.
struct WireCombineParams {
  const float* g;   // scratch
  float* grad;      // flat parameter gradient
  WireDims w;
  long long off[4 * kMaxSineLayers + 3];
};

__global__ void __launch_bounds__(256) wire_combine_kernel(const WireCombineParams p) {
  const int H = p.w.H, L = p.w.L, C = p.w.C, d = p.w.kin();  // d: columns of the first layer's (real) weights
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  // first layer (real): scratch rows 4u (lin) and 4u + 2 (orth), d columns
  {
    const float* G = p.g + wire_gblk_w(p.w, 0);
    const float* bs = p.g + wire_gblk_b(p.w, 0);
    for (long long i = tid; i < (long long)2 * H * (d + 1); i += nthreads) {
      const int which = int(i / ((long long)H * (d + 1)));
      const int rem = int(i % ((long long)H * (d + 1)));
      const int u = rem / (d + 1), j = rem % (d + 1);
      const int row = 4 * u + 2 * which;
      if (j < d)
        p.grad[p.off[2 * which] + (long long)u * d + j] += G[(long long)row * d + j];
      else
        p.grad[p.off[2 * which + 1] + u] += bs[row];
    }
  }
  for (int l = 1; l <= L; ++l) {
    const float* G = p.g + wire_gblk_w(p.w, l);
    const float* bs = p.g + wire_gblk_b(p.w, l);
    const long long nw = (long long)2 * H * H;  // (which, u, k)
    for (long long i = tid; i < nw; i += nthreads) {
      const int which = int(i / ((long long)H * H));
      const int u = int((i / H) % H), k = int(i % H);
      const float* g0 = G + (long long)(4 * u + 2 * which) * (2 * H);      // Re row
      const float* g1 = G + (long long)(4 * u + 2 * which + 1) * (2 * H);  // Im row
      float* dst = p.grad + p.off[4 * l + 2 * which] + ((long long)u * H + k) * 2;
      dst[0] += g0[k] + g1[H + k];
      dst[1] += g1[k] - g0[H + k];
    }
    for (long long i = tid; i < 4 * H; i += nthreads) {
      const int u = int(i >> 2), comp = int(i & 3);
      p.grad[p.off[4 * l + 2 * (comp >> 1) + 1] + 2 * u + (comp & 1)] += bs[i];
    }
  }
  {  // final linear: out = h_r Re W - h_i Im W + Re b
    const float* G = p.g + wire_gblk_wf(p.w);
    const float* bs = p.g + wire_gblk_bf(p.w);
    for (long long i = tid; i < (long long)C * H; i += nthreads) {
      const int c = int(i / H), k = int(i % H);
      float* dst = p.grad + p.off[4 * (L + 1)] + ((long long)c * H + k) * 2;
      dst[0] += G[(long long)c * 2 * H + k];
      dst[1] -= G[(long long)c * 2 * H + H + k];  // out = ... - h_i Im W
    }
    for (long long i = tid; i < C; i += nthreads) p.grad[p.off[4 * (L + 1) + 1] + 2 * i] += bs[i];
  }
}

int launch_wire_combine(const b200inr_net* net, void* stash, int64_t rows, float* grad_params, cudaStream_t stream) {
  WireCombineParams p{};
  p.w = make_wire_dims(net);
  const WireStashLayout sl = make_wire_stash_layout(p.w, rows);
  p.g = reinterpret_cast<const float*>(reinterpret_cast<uint8_t*>(stash) + sl.gblk);
  p.grad = grad_params;
  int64_t off[4 * kMaxSineLayers + 3] = {0};
  wire_param_offsets(p.w, off);
  for (int i = 0; i < 4 * (p.w.L + 1) + 3; ++i) p.off[i] = off[i];
  wire_combine_kernel<<<296, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
