// pack.cu -- fp32 reference-layout parameters -> bf16 tensor-core operand buffer (omega folded in).
//
// The reference keeps nn.Linear weights as fp32 [out, in] (INR/SRDWI.py:47) and multiplies by omega_0 after the
// linear (INR/SRDWI.py:59).  The fused kernels consume theta = x (omega W)^T + omega b directly, so this kernel
// folds omega into W and b while it converts to bf16 and lays every matrix out as SWIZZLE_128B tile blocks
// (see umma.cuh).  It runs once per optimiser step (0.5 MB of output for BASELINE config 2).
#include "common.cuh"
#include "umma.cuh"

namespace b200inr {

struct PackParams {
  const float* params;
  uint8_t* packed;
  PackLayout pl;
  long long off[2 * (kMaxSineLayers + 1)];
  int d, H, L, C;
  int Hr;  // real hidden width of the network (parameters); H is the padded operand width
  float omega0, omegah;
};

__device__ __forceinline__ void put_bf16(uint8_t* base, uint32_t row, uint32_t k, float v) {
  // element k (0..63) of row `row` inside a [rows][64] swizzled block
  *reinterpret_cast<__nv_bfloat16*>(base + sw128_chunk_off(row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
  const int H = p.H, L = p.L, C = p.C, d = p.d, Hr = p.Hr;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;

  // first layer: float4 rows, omega0 folded
  for (long long i = tid; i < H; i += nthreads) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < Hr)
      for (int j = 0; j < d; ++j) w[j] = p.omega0 * p.params[p.off[0] + i * d + j];
    reinterpret_cast<float4*>(p.packed + p.pl.w0)[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  // first layer as a tensor-core operand: theta_0 = [x_hi x_hi x_lo x_lo 1 1] . [W_hi W_lo W_hi W_lo b_hi b_lo]
  // (x = x_hi + x_lo and omega0*W = W_hi + W_lo, omega0*b = b_hi + b_lo in bf16: exact to 2^-17, fp32 accumulation).
  // Row = output feature; columns 4g + j (g = 0..3, j = coordinate), 16 = b_hi, 17 = b_lo, the rest zero.
  for (long long i = tid; i < (long long)H * 64; i += nthreads) {
    const int o = int(i >> 6), k = int(i & 63);
    float v = 0.f;
    if (k < 16) {
      const int g = k >> 2, j = k & 3;
      if (j < d && o < Hr) {
        const float w = p.omega0 * p.params[p.off[0] + (long long)o * d + j];
        const float hi = __bfloat162float(__float2bfloat16_rn(w));
        v = (g & 1) ? (w - hi) : hi;
      }
    } else if (k < 18 && o < Hr) {
      const float b = p.omega0 * p.params[p.off[1] + o];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      v = (k == 17) ? (b - hi) : hi;
    }
    put_bf16(p.packed + p.pl.w0p, o, k, v);
  }
  // biases
  float* bias = reinterpret_cast<float*>(p.packed + p.pl.bias);
  for (long long i = tid; i < (long long)(L + 1) * H + 32; i += nthreads) {
    float v = 0.f;
    if (i < (long long)(L + 1) * H) {
      const int l = int(i / H), h = int(i % H);
      if (h < Hr) v = (l == 0 ? p.omega0 : p.omegah) * p.params[p.off[2 * l + 1] + h];
    } else {
      const int c = int(i - (long long)(L + 1) * H);
      if (c < C) v = p.params[p.off[2 * (L + 1) + 1] + c];
    }
    bias[i] = v;
  }
  // hidden layers, both orientations
  const long long per_layer = (long long)H * H;
  for (long long i = tid; i < (long long)L * per_layer; i += nthreads) {
    const int l = int(i / per_layer);
    const int o = int((i % per_layer) / H);  // out feature
    const int k = int(i % H);                // in feature
    const float v = (o < Hr && k < Hr) ? p.omegah * p.params[p.off[2 * (l + 1)] + (long long)o * Hr + k] : 0.f;
    uint8_t* wh = p.packed + p.pl.wh + size_t(l) * per_layer * 2;
    uint8_t* wht = p.packed + p.pl.wht + size_t(l) * per_layer * 2;
    put_bf16(wh + size_t(k >> 6) * H * 128, o, k & 63, v);   // forward: N = out, K = in
    put_bf16(wht + size_t(o >> 6) * H * 128, k, o & 63, v);  // dgrad:   N = in,  K = out
  }
  // final linear: forward operand [H/64][32][64], rows >= C zero
  for (long long i = tid; i < (long long)kOutPad * H; i += nthreads) {
    const int c = int(i / H), k = int(i % H);
    const float v = (c < C && k < Hr) ? p.params[p.off[2 * (L + 1)] + (long long)c * Hr + k] : 0.f;
    put_bf16(p.packed + p.pl.wf + size_t(k >> 6) * kOutPad * 128, c, k & 63, v);
  }
  // final linear: dgrad operand [H][64] (N = in, K = c padded to 64)
  for (long long i = tid; i < (long long)H * kDzoPad; i += nthreads) {
    const int n = int(i / kDzoPad), c = int(i % kDzoPad);
    const float v = (c < C && n < Hr) ? p.params[p.off[2 * (L + 1)] + (long long)c * Hr + n] : 0.f;
    put_bf16(p.packed + p.pl.wft, n, c, v);
  }
}

// ------------------------------------------------------------------ generic family (see common.cuh)
struct GenPackParams {
  const float* params;
  uint8_t* packed;
  GenDims g;
  GenPackLayout pl;
  long long off[2 * (kMaxSineLayers + 2) + 1];
};

__global__ void __launch_bounds__(256) gen_pack_kernel(const GenPackParams p) {
  const GenDims g = p.g;
  const int H = g.H, L = g.L, C = g.C;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;

  float* bias = reinterpret_cast<float*>(p.packed + p.pl.bias);
  for (long long i = tid; i < (long long)(L + 1) * H + 32; i += nthreads) {
    float v = 0.f;
    if (i < (long long)(L + 1) * H) {
      const int l = int(i / H), h = int(i % H);
      v = (l == 0 ? g.omega0 : g.omegah) * p.params[p.off[2 * l + 1] + h];
    } else {
      const int c = int(i - (long long)(L + 1) * H);
      if (c < C) v = p.params[p.off[2 * (L + 1) + 1] + c];
    }
    bias[i] = v;
  }
  for (long long i = tid; i < g.m; i += nthreads) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < g.d; ++j) w[j] = p.params[p.off[2 * (L + 2)] + i * g.d + j];
    reinterpret_cast<float4*>(p.packed + p.pl.bmat)[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  // activated layers
  for (int l = 0; l <= L; ++l) {
    const int K = (l == 0) ? g.K0 : H;
    const int kbn = K / 64;
    const float om = (l == 0) ? g.omega0 : g.omegah;
    uint8_t* wf = p.packed + p.pl.w_layer(g, l);
    uint8_t* wt = (l >= 1) ? p.packed + p.pl.wt_layer(g, l) : nullptr;
    for (long long i = tid; i < (long long)H * K; i += nthreads) {
      const int o = int(i / K), k = int(i % K);
      const float v = om * p.params[p.off[2 * l] + i];
      put_bf16(wf + size_t((o >> 8) * kbn + (k >> 6)) * kGenChunkBytes, o & 255, k & 63, v);
      if (wt) put_bf16(wt + size_t((k >> 8) * (H / 64) + (o >> 6)) * kGenChunkBytes, k & 255, o & 63, v);
    }
  }
  // layer 0 transposed, operand of the input-gradient step dX = dTheta_0 (omega_0 W_0): N = input k (padded to whole
  // 256-row chunks with zeros), K = output o; chunk order (n-half over inputs, k-block over outputs) like wt
  {
    const int K0p = ((g.K0 + 255) / 256) * 256;
    uint8_t* wt0 = p.packed + p.pl.wt0;
    for (long long i = tid; i < (long long)H * K0p; i += nthreads) {
      const int o = int(i / K0p), k = int(i % K0p);
      const float v = (k < g.K0) ? g.omega0 * p.params[p.off[0] + (long long)o * g.K0 + k] : 0.f;
      put_bf16(wt0 + size_t((k >> 8) * (H / 64) + (o >> 6)) * kGenChunkBytes, k & 255, o & 63, v);
    }
  }
  // final linear
  for (long long i = tid; i < (long long)kOutPad * H; i += nthreads) {
    const int c = int(i / H), k = int(i % H);
    const float v = (c < C) ? p.params[p.off[2 * (L + 1)] + (long long)c * H + k] : 0.f;
    put_bf16(p.packed + p.pl.wf + size_t(k >> 6) * kOutPad * 128, c, k & 63, v);
  }
  for (long long i = tid; i < (long long)H * kDzoPad; i += nthreads) {
    const int n = int(i / kDzoPad), c = int(i % kDzoPad);
    const float v = (c < C) ? p.params[p.off[2 * (L + 1)] + (long long)c * H + n] : 0.f;
    put_bf16(p.packed + p.pl.wft + size_t(n >> 8) * kGenChunkBytes, n & 255, c, v);
  }
}

int launch_gen_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream) {
  GenPackParams p{};
  p.params = params;
  p.packed = reinterpret_cast<uint8_t*>(packed);
  p.g = make_gen_dims(net);
  p.pl = make_gen_pack_layout(p.g);
  int64_t off[2 * (kMaxSineLayers + 2) + 1] = {0};
  gen_param_offsets(p.g, off);
  for (int i = 0; i < 2 * (p.g.L + 2) + 1; ++i) p.off[i] = off[i];
  gen_pack_kernel<<<592, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

int launch_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream) {
  if (net->input_mode != B200INR_IN_COORDS) return launch_gen_pack(net, params, packed, stream);
  PackParams p{};
  p.params = params;
  p.packed = reinterpret_cast<uint8_t*>(packed);
  p.d = net->in_features;
  p.H = kSirenWidth;
  p.Hr = net->hidden_features;
  p.L = net->hidden_layers;
  p.C = net->out_features;
  p.omega0 = net->first_omega_0;
  p.omegah = net->hidden_omega_0;
  p.pl = make_pack_layout(p.H, p.L);
  int64_t off[2 * (kMaxSineLayers + 1)];
  param_offsets(p.d, p.Hr, p.L, p.C, off);
  for (int i = 0; i < 2 * (p.L + 2); ++i) p.off[i] = off[i];
  pack_kernel<<<296, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
