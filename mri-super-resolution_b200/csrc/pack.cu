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
  float omega0, omegah;
};

__device__ __forceinline__ void put_bf16(uint8_t* base, uint32_t row, uint32_t k, float v) {
  // element k (0..63) of row `row` inside a [rows][64] swizzled block
  *reinterpret_cast<__nv_bfloat16*>(base + sw128_chunk_off(row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
  const int H = p.H, L = p.L, C = p.C, d = p.d;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;

  // first layer: float4 rows, omega0 folded
  for (long long i = tid; i < H; i += nthreads) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < d; ++j) w[j] = p.omega0 * p.params[p.off[0] + i * d + j];
    reinterpret_cast<float4*>(p.packed + p.pl.w0)[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  // biases
  float* bias = reinterpret_cast<float*>(p.packed + p.pl.bias);
  for (long long i = tid; i < (long long)(L + 1) * H + 32; i += nthreads) {
    float v = 0.f;
    if (i < (long long)(L + 1) * H) {
      const int l = int(i / H), h = int(i % H);
      v = (l == 0 ? p.omega0 : p.omegah) * p.params[p.off[2 * l + 1] + h];
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
    const float v = p.omegah * p.params[p.off[2 * (l + 1)] + (long long)o * H + k];
    uint8_t* wh = p.packed + p.pl.wh + size_t(l) * per_layer * 2;
    uint8_t* wht = p.packed + p.pl.wht + size_t(l) * per_layer * 2;
    put_bf16(wh + size_t(k >> 6) * H * 128, o, k & 63, v);   // forward: N = out, K = in
    put_bf16(wht + size_t(o >> 6) * H * 128, k, o & 63, v);  // dgrad:   N = in,  K = out
  }
  // final linear: forward operand [H/64][32][64], rows >= C zero
  for (long long i = tid; i < (long long)kOutPad * H; i += nthreads) {
    const int c = int(i / H), k = int(i % H);
    const float v = (c < C) ? p.params[p.off[2 * (L + 1)] + (long long)c * H + k] : 0.f;
    put_bf16(p.packed + p.pl.wf + size_t(k >> 6) * kOutPad * 128, c, k & 63, v);
  }
  // final linear: dgrad operand [H][64] (N = in, K = c padded to 64)
  for (long long i = tid; i < (long long)H * kDzoPad; i += nthreads) {
    const int n = int(i / kDzoPad), c = int(i % kDzoPad);
    const float v = (c < C) ? p.params[p.off[2 * (L + 1)] + (long long)c * H + n] : 0.f;
    put_bf16(p.packed + p.pl.wft, n, c, v);
  }
}

int launch_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream) {
  PackParams p{};
  p.params = params;
  p.packed = reinterpret_cast<uint8_t*>(packed);
  p.d = net->in_features;
  p.H = net->hidden_features;
  p.L = net->hidden_layers;
  p.C = net->out_features;
  p.omega0 = net->first_omega_0;
  p.omegah = net->hidden_omega_0;
  p.pl = make_pack_layout(p.H, p.L);
  int64_t off[2 * (kMaxSineLayers + 1)];
  param_offsets(p.d, p.H, p.L, p.C, off);
  for (int i = 0; i < 2 * (p.L + 2); ++i) p.off[i] = off[i];
  pack_kernel<<<296, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

}  // namespace b200inr
