// common.cuh -- byte layouts shared by the kernels and the C-ABI launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200inr.h"

namespace b200inr {

constexpr int kTileRows = 128;     // coordinate rows per tile == UMMA M
constexpr int kSirenWidth = 256;   // width the raw-coordinate SIREN kernels are built for (narrower nets are padded)
constexpr int kMaxSineLayers = 8;  // L + 1 <= 8
constexpr int kOutPad = 32;        // final layer N (C <= 32)
constexpr int kDzoPad = 64;        // dL/dout tile width (one 128-byte swizzle row)

// Byte offsets inside the packed (bf16, omega-folded) weight buffer of a width-H SIREN.
struct PackLayout {
  size_t w0;     // float4[H]            omega0 * W0 rows, zero padded to 4 inputs
  size_t bias;   // float[(L+1)*H + 32 + H]  omega * b of the sine layers, the final bias (padded to 32), then H zeros
                 //                      (what the forward adds in the first layer, whose bias is part of the GEMM)
  size_t wh;     // L x [H/64][H][64]    forward B operand of hidden layer l (N = out, K = in), omega_h * W_l
  size_t wht;    // L x [H/64][H][64]    dgrad   B operand of hidden layer l (N = in, K = out), omega_h * W_l^T
  size_t wf;     // [H/64][32][64]       forward B operand of the final linear (N = c, K = in)
  size_t wft;    // [1][H][64]           dgrad   B operand of the final linear (N = in, K = c padded to 64)
  size_t w0p;    // [1][H][64]           forward B operand of the FIRST layer on tensor cores (N = out, K = 32 used):
                 //                      omega0 * W0 and omega0 * b0 split into bf16 hi + lo parts, see pack.cu
  size_t bias16; // __half[(L+1)*H + 32 + H]  the bias table once more in fp16 (same indices): what the forward's
                 //                      epilogue adds -- every thread adds the same 64 biases per layer-tile to its row,
                 //                      and as fp32 those broadcast loads were as many bytes into registers as the
                 //                      accumulator itself (mlp_fwd.cu).  |omega b| < 4: half an fp16 ulp is < 1e-3 rad,
                 //                      a third of what the bf16 operands already put into every angle.
  size_t total;
};

__host__ __device__ inline PackLayout make_pack_layout(int H, int L) {
  PackLayout p;
  size_t o = 0;
  p.w0 = o;
  o += size_t(H) * 16;
  p.bias = o;
  o += (size_t(L + 1) * H + 32 + H) * 4;
  o = (o + 1023) & ~size_t(1023);
  p.wh = o;
  o += size_t(L) * H * H * 2;
  p.wht = o;
  o += size_t(L) * H * H * 2;
  p.wf = o;
  o += size_t(H / 64) * kOutPad * 128;
  p.wft = o;
  o += size_t(H) * 128;
  p.w0p = o;
  o += size_t(H) * 128;
  p.bias16 = o;
  o += (size_t(L + 1) * H + 32 + H) * 2;
  o = (o + 1023) & ~size_t(1023);
  p.total = o;
  return p;
}

// Activation stash for `rows` coordinates (T = ceil(rows/128) tiles), width-H SIREN with L+1 sine layers.
struct StashLayout {
  size_t y;             // (L+1) x T x [H/64][128][64] bf16   sin outputs, UMMA tile layout
  size_t ph;            // (L+1) x T x [H/8][128][8]   u16    phase = round(theta * 65536 / 2pi) mod 65536
  size_t dz;            // (L+1) x T x [H/64][128][64] bf16   dL/dtheta, UMMA tile layout
  size_t dzo;           // T x [128][64] bf16                  dL/dout, one block per tile
  size_t xa;            // T x [128][64] bf16                  coordinates as bf16 hi (cols 0..3) + lo (cols 4..7)
  size_t layer_stride;  // bytes per layer inside y / ph / dz
  size_t tile_bytes;    // 128 * H * 2
  size_t total;
  int64_t tiles;
};

__host__ __device__ inline StashLayout make_stash_layout(int H, int L, int64_t rows) {
  StashLayout s;
  s.tiles = (rows + kTileRows - 1) / kTileRows;
  s.tile_bytes = size_t(kTileRows) * H * 2;
  s.layer_stride = size_t(s.tiles) * s.tile_bytes;
  size_t o = 0;
  s.y = o;
  o += size_t(L + 1) * s.layer_stride;
  s.ph = o;
  o += size_t(L + 1) * s.layer_stride;
  s.dz = o;
  o += size_t(L + 1) * s.layer_stride;
  s.dzo = o;
  o += size_t(s.tiles) * kTileRows * kDzoPad * 2;
  s.xa = o;
  o += size_t(s.tiles) * kTileRows * kDzoPad * 2;
  s.total = o;
  return s;
}

// Stash of the pipelined SIREN training path (mlp_bwdp.cu): only the 16-bit phases go to HBM; the rest is the
// L2-resident ring storage of the layer pipelines and their flag counters.
constexpr int kPipeTileRows = 64;    // rows per pipeline tile (half a forward tile)
// Phase stash of the pipelined path: per (layer, forward tile) two 64-row halves of [H/8 chunks][64 rows][8] u16, every
// chunk padded by 32 bytes -- the bank spread mlp_bwdp.cu reads with; 32 keeps the forward's 512-byte warp stores on
// whole sectors (a 16-byte pad cost the forward 15 %) -- so that the 128 features x 64 rows a backward CTA needs are
// ONE contiguous bulk copy (sixteen 1 KB copies cost the loader thread ~1.6 k cycles per tile).
constexpr int kPipePhChunk = kPipeTileRows * 16 + 32;                 // 1056 bytes
constexpr int kPipePhHalf = (kSirenWidth / 8) * kPipePhChunk;        // 64 rows x 256 features: 33 792 bytes
constexpr int kPipePhTile = 2 * kPipePhHalf;                         // one 128-row forward tile: 67 584 bytes
#ifndef B200INR_PRING
#define B200INR_PRING 10
#endif
// ring depth, in tiles, per layer boundary: 4 / 6 / 8 / 10 / 12 / 16 slots give a cfg2 backward of 1.40 / 1.33 / 1.29 /
// 1.27 / 1.27 / 1.36 ms (a ring has to absorb the jitter of its two ends; 16 costs more L2 than it buys)
constexpr int kPipeRing = B200INR_PRING;
constexpr int kPipeMaxEdges = 96;    // pipelines x (L+1) <= #SM / 2
// Networks with at least one hidden layer, evaluated on a coordinate GRID whose last axis and first row are multiples of
// 16, do not stash the phases of layer 0: inside an aligned batch of 16 rows only the last coordinate changes, so
// theta_0 = base + w_last x_last is ONE FMA and one 4-byte shared-memory read per element, from the fp32 coordinate
// records (16 bytes per row) the forward leaves at the start of the layer-0 phase region.  (A per-row dot product from
// the whole record was measured first: the same bytes broadcast to 32 lanes cost four times the shared-memory return
// bandwidth of a 2-byte phase, the layer-1 CTAs fell behind and every pipeline with them: +7 %.)  A fifth of the stash
// traffic of both directions.  The forward leaves what it did in word kPipeSkipWord of the calibration area.
#ifndef B200INR_PSKIP0
#define B200INR_PSKIP0 1
#endif
constexpr bool kPipeSkipPh0 = B200INR_PSKIP0 != 0;
constexpr int kPipeCalRow = 176;     // first row of the profiling area used for launch-to-launch state (rows < grid <= 148
                                     // belong to the profiling builds): [epoch, skip word, pad x14, 2 banks of speed records]
constexpr int kPipeSkipWord = 1;
struct PipeStashLayout {
  size_t ph;            // (L+1) x T x kPipePhTile bytes (see above)
  size_t layer_stride;  // bytes per layer inside ph
  size_t xa;            // T x 128 rows x 16 bytes: coordinates as bf16 {hi x4, lo x4} (operand of dW_0)
  size_t ring;          // kPipeMaxEdges x kPipeRing x [64 rows x H bf16]
  size_t flags;         // kPipeMaxEdges x 4 counters, 128 bytes apart
  size_t flags_bytes;
  size_t prof;          // kPipeProfCtas x kPipeProfSlots stall-cycle counters (written only when profiling is on);
                        // always the LAST kPipeProfCtas * kPipeProfSlots * 8 bytes of the stash
  size_t total;
  int64_t tiles;        // 128-row forward tiles
};
constexpr int kPipeProfCtas = 192;
constexpr int kPipeProfSlots = 32;

__host__ __device__ inline PipeStashLayout make_pipe_stash_layout(int H, int L, int64_t rows) {
  PipeStashLayout s;
  s.tiles = (rows + kTileRows - 1) / kTileRows;
  s.layer_stride = size_t(s.tiles) * kPipePhTile;
  size_t o = 0;
  s.ph = o;
  o += size_t(L + 1) * s.layer_stride;
  o = (o + 1023) & ~size_t(1023);
  s.xa = o;
  o += size_t(s.tiles) * kTileRows * 16;
  s.ring = o;
  o += size_t(kPipeMaxEdges) * kPipeRing * kPipeTileRows * H * 2;
  s.flags = o;
  s.flags_bytes = size_t(kPipeMaxEdges) * 4 * 128;
  o += s.flags_bytes;
  s.prof = o;
  o += size_t(kPipeProfCtas) * kPipeProfSlots * 8;
  s.total = o;
  return s;
}

// Flat fp32 parameter offsets (in floats): W_i at off[2i], b_i at off[2i+1], i = 0..L+1.
__host__ __device__ inline int64_t param_offsets(int d, int H, int L, int C, int64_t* off) {
  int64_t o = 0;
  auto seg = [&](int64_t n) {
    int64_t at = o;
    o += (n + 3) & ~int64_t(3);
    return at;
  };
  int64_t w, b;
  w = seg(int64_t(H) * d);
  b = seg(H);
  if (off) { off[0] = w; off[1] = b; }
  for (int l = 1; l <= L; ++l) {
    w = seg(int64_t(H) * H);
    b = seg(H);
    if (off) { off[2 * l] = w; off[2 * l + 1] = b; }
  }
  w = seg(int64_t(C) * H);
  b = seg(C);
  if (off) { off[2 * (L + 1)] = w; off[2 * (L + 1) + 1] = b; }
  return o;
}

// ------------------------------------------------------------------ generic family (gen_*.cu)
// Networks whose first layer is already a tensor-core layer: input = Fourier features computed in-kernel
// (B200INR_IN_FOURIER) or explicit feature rows (B200INR_IN_FEATURES); H = 256 or 512; sine or ReLU.
struct GenDims {
  int d;    // raw coordinate dimension (IN_FOURIER) or 0
  int m;    // mapping size (IN_FOURIER) or 0
  int K0;   // width of the network input (2m or in_features), multiple of 64, <= H
  int H, L, C;
  int act, in_mode;
  float omega0, omegah;
  int lean;  // B200INR_NET_DGRAD_ONLY: the stash only holds what the activation-gradient chain reads
  __host__ __device__ int kb0() const { return K0 / 64; }
  __host__ __device__ int kbh() const { return H / 64; }
  __host__ __device__ int nh() const { return H / 256; }
};

__host__ inline GenDims make_gen_dims(const b200inr_net* n) {
  GenDims g{};
  g.in_mode = n->input_mode;
  g.d = n->input_mode == B200INR_IN_FOURIER ? n->in_features : 0;
  g.m = n->input_mode == B200INR_IN_FOURIER ? n->mapping_size : 0;
  g.K0 = n->input_mode == B200INR_IN_FOURIER ? 2 * n->mapping_size : n->in_features;
  g.H = n->hidden_features;
  g.L = n->hidden_layers;
  g.C = n->out_features;
  g.act = n->activation;
  g.omega0 = n->activation == B200INR_ACT_SINE ? n->first_omega_0 : 1.0f;
  g.omegah = n->activation == B200INR_ACT_SINE ? n->hidden_omega_0 : 1.0f;
  g.lean = (n->flags & B200INR_NET_DGRAD_ONLY) ? 1 : 0;
  return g;
}

// Flat fp32 parameter offsets of the generic family: W_i at off[2i], b_i at off[2i+1] (i = 0..L+1), B at off[2(L+2)].
__host__ __device__ inline int64_t gen_param_offsets(const GenDims& g, int64_t* off) {
  int64_t o = 0;
  auto seg = [&](int64_t n) {
    int64_t at = o;
    o += (n + 3) & ~int64_t(3);
    return at;
  };
  for (int l = 0; l <= g.L; ++l) {
    const int64_t w = seg(int64_t(g.H) * (l == 0 ? g.K0 : g.H));
    const int64_t b = seg(g.H);
    if (off) { off[2 * l] = w; off[2 * l + 1] = b; }
  }
  const int64_t w = seg(int64_t(g.C) * g.H);
  const int64_t b = seg(g.C);
  if (off) { off[2 * (g.L + 1)] = w; off[2 * (g.L + 1) + 1] = b; }
  if (g.in_mode == B200INR_IN_FOURIER) {
    const int64_t bm = seg(int64_t(g.m) * g.d);
    if (off) off[2 * (g.L + 2)] = bm;
  }
  return o;
}

constexpr int kGenChunkBytes = 256 * 128;  // weight chunk: [256 rows (N)][64 (K)] bf16, SWIZZLE_128B

// Packed operand buffer of the generic family.  Forward chunks of layer l are ordered (n-half, k-block);
// dgrad chunks of layer l >= 1 are ordered (n-half over inputs, k-block over outputs).
struct GenPackLayout {
  size_t bias;   // float [(L+1)*H + 32]  omega-folded biases, then the final bias padded to 32
  size_t bmat;   // float4 [m]            frequency matrix rows (zero padded to 4 coordinates)
  size_t w;      // layer 0: nh*kb0 chunks; layers 1..L: nh*kbh chunks each
  size_t wt;     // layers 1..L: nh*kbh chunks each
  size_t wf;     // [kbh][32][64]
  size_t wft;    // nh chunks [256][64]
  size_t wt0;    // layer 0 transposed (input-gradient step): ceil(K0/256) * kbh chunks, rows >= K0 zero
  size_t total;
  __host__ __device__ size_t w_layer(const GenDims& g, int l) const {
    return l == 0 ? w : w + size_t(g.nh()) * g.kb0() * kGenChunkBytes +
                            size_t(l - 1) * g.nh() * g.kbh() * kGenChunkBytes;
  }
  __host__ __device__ size_t wt_layer(const GenDims& g, int l) const {  // l >= 1
    return wt + size_t(l - 1) * g.nh() * g.kbh() * kGenChunkBytes;
  }
};

__host__ __device__ inline GenPackLayout make_gen_pack_layout(const GenDims& g) {
  GenPackLayout p;
  size_t o = 0;
  p.bias = o;
  o += (size_t(g.L + 1) * g.H + 32) * 4;
  o = (o + 15) & ~size_t(15);
  p.bmat = o;
  o += size_t(g.m) * 16;
  o = (o + 1023) & ~size_t(1023);
  p.w = o;
  o += (size_t(g.nh()) * g.kb0() + size_t(g.L) * g.nh() * g.kbh()) * kGenChunkBytes;
  p.wt = o;
  o += size_t(g.L) * g.nh() * g.kbh() * kGenChunkBytes;
  p.wf = o;
  o += size_t(g.kbh()) * kOutPad * 128;
  p.wft = o;
  o += size_t(g.nh()) * kGenChunkBytes;
  p.wt0 = o;
  o += size_t((g.K0 + 255) / 256) * g.kbh() * kGenChunkBytes;
  p.total = o;
  return p;
}

// Activation stash of the generic family (T = ceil(rows/128) tiles).
struct GenStashLayout {
  size_t ain;  // T x [K0/64][128][64] bf16          network input (A operand of layer 0, B operand of dW_0)
  size_t y;    // (L+1) x T x [H/64][128][64] bf16   layer outputs
  size_t ph;   // (L+1) x T x [H/8][128][8] u16      phases (sine only, else empty)
  size_t dz;   // (L+1) x T x [H/64][128][64] bf16   dL/dtheta
  size_t dzo;  // T x [128][64] bf16                  dL/dout
  size_t tile_in, tile_h, layer_stride;
  size_t total;
  int64_t tiles;
};

__host__ __device__ inline GenStashLayout make_gen_stash_layout(const GenDims& g, int64_t rows) {
  GenStashLayout s;
  s.tiles = (rows + kTileRows - 1) / kTileRows;
  s.tile_in = size_t(kTileRows) * g.K0 * 2;
  s.tile_h = size_t(kTileRows) * g.H * 2;
  s.layer_stride = size_t(s.tiles) * s.tile_h;
  size_t o = 0;
  s.ain = o;
  if (!g.lean) o += size_t(s.tiles) * s.tile_in;
  s.y = o;
  if (!g.lean || g.act != B200INR_ACT_SINE) o += size_t(g.L + 1) * s.layer_stride;
  s.ph = o;
  if (g.act == B200INR_ACT_SINE) o += size_t(g.L + 1) * s.layer_stride;
  s.dz = o;
  if (!g.lean) o += size_t(g.L + 1) * s.layer_stride;
  s.dzo = o;
  if (!g.lean) o += size_t(s.tiles) * kTileRows * kDzoPad * 2;
  s.total = o > 0 ? o : 1024;
  return s;
}

// ------------------------------------------------------------------ WIRE family (wire.cu)
// Complex Gabor network of INR/wiretest.ipynb cell 2 as a real-block GEMM chain: H complex units,
// activations [h_r | h_i] (K = 2H reals), hidden layers produce the four reals (a, b, c, d) = (Re lin, Im lin, Re orth,
// Im orth) of every unit, interleaved as column n = 4u + component (N = 4H).
struct WireDims {
  int d, H, L, C;
  float omega0, omegah, s0;
  int in_mode;  // B200INR_IN_COORDS: first layer on raw coordinates (CUDA cores); IN_FEATURES / IN_FOURIER: explicit or
                // in-kernel Fourier feature rows (wiretest.ipynb cell 7), first layer on the tensor cores
  int m;        // mapping size (IN_FOURIER) or 0
  int K0;       // feature width (2m or in_features; multiple of 64, <= 512), 0 for raw coordinates
  __host__ __device__ int kin() const { return K0 ? K0 : d; }  // columns of the first layer's weight matrices
  __host__ __device__ int kb0() const { return K0 / 64; }
};

__host__ inline WireDims make_wire_dims(const b200inr_net* n) {
  WireDims w{};
  w.d = n->in_features;
  w.H = n->hidden_features;
  w.L = n->hidden_layers;
  w.C = n->out_features;
  w.omega0 = n->first_omega_0;
  w.omegah = n->hidden_omega_0;
  w.s0 = n->scale_0;
  w.in_mode = n->input_mode;
  w.m = n->input_mode == B200INR_IN_FOURIER ? n->mapping_size : 0;
  w.K0 = n->input_mode == B200INR_IN_FOURIER ? 2 * n->mapping_size
         : n->input_mode == B200INR_IN_FEATURES ? n->in_features : 0;
  if (n->input_mode == B200INR_IN_FEATURES) w.d = 0;
  return w;
}

// Flat fp32 parameters: layer l: W_lin b_lin W_orth b_orth at off[4l .. 4l+3] (l = 0 real, l >= 1 complex as (re, im)
// pairs), final W_f b_f (complex) at off[4(L+1)], off[4(L+1)+1]; IN_FOURIER: the frozen matrix B [m, d] at off[4(L+1)+2].
__host__ __device__ inline int64_t wire_param_offsets(const WireDims& w, int64_t* off) {
  int64_t o = 0;
  auto seg = [&](int64_t n) {
    int64_t at = o;
    o += (n + 3) & ~int64_t(3);
    return at;
  };
  for (int l = 0; l <= w.L; ++l) {
    const int64_t nw = (l == 0) ? int64_t(w.H) * w.kin() : int64_t(w.H) * w.H * 2;
    const int64_t nbias = (l == 0) ? w.H : 2 * w.H;
    for (int j = 0; j < 2; ++j) {
      const int64_t a = seg(nw), b = seg(nbias);
      if (off) { off[4 * l + 2 * j] = a; off[4 * l + 2 * j + 1] = b; }
    }
  }
  const int64_t a = seg(int64_t(w.C) * w.H * 2), b = seg(2 * w.C);
  if (off) { off[4 * (w.L + 1)] = a; off[4 * (w.L + 1) + 1] = b; }
  if (w.in_mode == B200INR_IN_FOURIER) {
    const int64_t bm = seg(int64_t(w.m) * w.d);
    if (off) off[4 * (w.L + 1) + 2] = bm;
  }
  return o;
}

struct WirePackLayout {
  size_t w0;     // float4 [2H]: unit u -> lin weights at 2u, orth weights at 2u+1 (zero padded to 4 coordinates)
  size_t b0;     // float2 [H]: (b_lin, b_orth) of the first layer
  size_t bias;   // float [L*4H + 32]: hidden-layer biases in column order n = 4u + component, then Re(b_f) padded
  size_t w;      // L x 8 chunks [256 rows (n)][64 (k)]   forward operand, order (n-half, k-block)
  size_t wt;     // L x 8 chunks [256 rows (k)][64 (n)]   dgrad operand, order k-block over n
  size_t wf;     // [4][32][64]     final linear: rows c, k < H: Re W_f, k >= H: -Im W_f
  size_t wft;    // [256 rows (k)][64 (c)]
  size_t w0f;    // feature-fed first layer: K0/64 chunks [256 rows (n = 2u + {lin, orth})][64 (k)]
  size_t wt0;    // feature-fed first layer, input-gradient operand: (K0/256 n-halves) x 8 chunks [256 rows (input k)][64 (n = 4u + comp)]
  size_t bmat;   // float4 [m]: rows of the Fourier matrix B (zero padded to 4 coordinates)
  size_t total;
};

__host__ __device__ inline WirePackLayout make_wire_pack_layout(const WireDims& w) {
  WirePackLayout p;
  size_t o = 0;
  p.w0 = o;
  o += size_t(2 * w.H) * 16;
  p.b0 = o;
  o += size_t(w.H) * 8;
  p.bias = o;
  o += (size_t(w.L) * 4 * w.H + 32) * 4;
  o = (o + 1023) & ~size_t(1023);
  p.w = o;
  o += size_t(w.L) * 8 * kGenChunkBytes;
  p.wt = o;
  o += size_t(w.L) * 8 * kGenChunkBytes;
  p.wf = o;
  o += size_t(4) * kOutPad * 128;
  p.wft = o;
  o += kGenChunkBytes;
  p.w0f = o;
  o += size_t(w.kb0()) * kGenChunkBytes;
  p.wt0 = o;
  o += size_t((w.K0 + 255) / 256) * 8 * kGenChunkBytes;
  p.bmat = o;
  o += size_t(w.m) * 16;
  p.total = o;
  return p;
}

// fp32 offsets (in floats) inside the wgrad scratch of the WIRE family: per layer the real-block gradient [4H][K] and
// the column sums [4H] (K = 2H; the first layer of a feature-fed network has K = K0 columns, so its slot is sized for
// the wider of the two), then the final layer's [32][2H] + [32].
__host__ __device__ inline size_t wire_gblk_w(const WireDims& w, int l) {
  const size_t k_first = size_t(w.K0 > 2 * w.H ? w.K0 : 2 * w.H);
  const size_t first = size_t(4 * w.H) * k_first + 4 * w.H;
  return l == 0 ? 0 : first + size_t(l - 1) * (size_t(4 * w.H) * 2 * w.H + 4 * w.H);
}
__host__ __device__ inline size_t wire_gblk_b(const WireDims& w, int l) {
  const size_t k = (l == 0) ? size_t(w.K0 > 2 * w.H ? w.K0 : 2 * w.H) : size_t(2 * w.H);
  return wire_gblk_w(w, l) + size_t(4 * w.H) * k;
}
__host__ __device__ inline size_t wire_gblk_wf(const WireDims& w) { return wire_gblk_w(w, w.L + 1); }
__host__ __device__ inline size_t wire_gblk_bf(const WireDims& w) { return wire_gblk_wf(w) + size_t(kOutPad) * 2 * w.H; }

// Stash of the WIRE training forward: per Gabor layer the activations [h_r | h_i] (bf16 tile, 2H wide) and the
// pre-activations (a, b, c, d) (bf16 tile, 4H wide; layer 0 stores (a, 0, c, 0)); dz = dL/d(a,b,c,d) written by dgrad.
struct WireStashLayout {
  size_t y, z, dz, dzo, xa;
  size_t ain;   // feature-fed networks: T x [K0/64][128][64] bf16, the network input (A operand of layer 0, B operand of dW_0)
  size_t tile_in;
  size_t gblk;  // fp32 scratch of wgrad: per layer the real-block gradient [4H][K] + column sums [4H]; final [32][2H] + [32]
  size_t gblk_bytes;
  size_t tile_y, tile_z, stride_y, stride_z;
  size_t total;
  int64_t tiles;
};

__host__ __device__ inline WireStashLayout make_wire_stash_layout(const WireDims& w, int64_t rows) {
  WireStashLayout s;
  s.tiles = (rows + kTileRows - 1) / kTileRows;
  s.tile_y = size_t(kTileRows) * 2 * w.H * 2;
  s.tile_z = size_t(kTileRows) * 4 * w.H * 2;
  s.stride_y = size_t(s.tiles) * s.tile_y;
  s.stride_z = size_t(s.tiles) * s.tile_z;
  size_t o = 0;
  s.y = o;
  o += size_t(w.L + 1) * s.stride_y;
  s.z = o;
  o += size_t(w.L + 1) * s.stride_z;
  s.dz = o;
  o += size_t(w.L + 1) * s.stride_z;
  s.dzo = o;
  o += size_t(s.tiles) * kTileRows * kDzoPad * 2;
  s.xa = o;
  o += size_t(s.tiles) * kTileRows * kDzoPad * 2;
  s.ain = o;
  s.tile_in = size_t(kTileRows) * w.K0 * 2;
  o += size_t(s.tiles) * s.tile_in;
  s.gblk = o;
  s.gblk_bytes = (wire_gblk_w(w, w.L + 1) + size_t(kOutPad) * 2 * w.H + kOutPad) * 4;
  o += (s.gblk_bytes + 1023) & ~size_t(1023);
  s.total = o;
  return s;
}

struct GridDesc {
  int ndim;
  int shape[4];
  long long row_begin;
  long long total;  // product of shape
};

// torch.linspace(-1, 1, n)[i] in fp32, bit-for-bit (ATen RangeFactories: symmetric two-sided evaluation,
// step = (end - start) / (n - 1), multiply and add rounded separately).
__host__ __device__ inline float linspace_m1p1(int i, int n) {
  if (n <= 1) return -1.0f;
  const float step = 2.0f / float(n - 1);
#ifdef __CUDA_ARCH__
  if (i < n / 2) return __fadd_rn(-1.0f, __fmul_rn(step, float(i)));
  return __fsub_rn(1.0f, __fmul_rn(step, float(n - 1 - i)));
#else
  volatile float prod;
  if (i < n / 2) {
    prod = step * float(i);
    return -1.0f + prod;
  }
  prod = step * float(n - 1 - i);
  return 1.0f - prod;
#endif
}

__device__ inline void grid_coords(const GridDesc& g, long long row, float (&x)[4]) {
  long long idx = g.row_begin + row;
  if (idx >= g.total) idx = g.total - 1;
  x[0] = x[1] = x[2] = x[3] = 0.0f;
  if (g.total <= 0x7fffffffLL) {  // 32-bit index arithmetic (every BASELINE grid, 2^26 voxels at most)
    unsigned int i32 = (unsigned int)idx;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      if (j < g.ndim) {
        const unsigned int n = (unsigned int)g.shape[j];
        const unsigned int q = i32 / n;
        x[j] = linspace_m1p1(int(i32 - q * n), int(n));
        i32 = q;
      }
    }
    return;
  }
#pragma unroll
  for (int j = 3; j >= 0; --j) {
    if (j < g.ndim) {
      const int n = g.shape[j];
      const int i = int(idx % n);
      idx /= n;
      x[j] = linspace_m1p1(i, n);
    }
  }
}

}  // namespace b200inr
