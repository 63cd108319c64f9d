// pack.cu -- fp32 reference-layout parameters -> bf16 tensor-core operand buffer (omega folded in).
//
// The reference keeps nn.Linear weights as fp32 [out, in] (INR/SRDWI.py:47) and multiplies by omega_0 after the
// linear (INR/SRDWI.py:59).  The fused kernels consume theta = x (omega W)^T + omega b directly, so this kernel
// folds omega into W and b while it converts to bf16 and lays every matrix out as SWIZZLE_128B tile blocks
// (see umma.cuh).  It runs once per optimiser step (0.5 MB of output for BASELINE config 2).
#include <cuda_fp16.h>

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
  int relu_tail;  // B200INR_NET_RELU_TAIL: the last hidden layer is Linear + ReLU (omega = 1)
  // omega folded into the operands of activated layer l (0 = first)
  __host__ __device__ float omega(int l) const { return l == 0 ? omega0 : ((relu_tail && l == L) ? 1.0f : omegah); }
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
  for (long long i = tid; i < (long long)(L + 1) * H + 32 + H; i += nthreads) {
    float v = 0.f;
    if (i < (long long)(L + 1) * H) {
      const int l = int(i / H), h = int(i % H);
      if (h < Hr) v = p.omega(l) * p.params[p.off[2 * l + 1] + h];
    } else if (i < (long long)(L + 1) * H + 32) {
      const int c = int(i - (long long)(L + 1) * H);
      if (c < C) v = p.params[p.off[2 * (L + 1) + 1] + c];
    }  // else: the zero block (written here once; the fused optimiser step never touches it)
    bias[i] = v;
    reinterpret_cast<__half*>(p.packed + p.pl.bias16)[i] = __float2half_rn(v);
  }
  // hidden layers, both orientations
  const long long per_layer = (long long)H * H;
  for (long long i = tid; i < (long long)L * per_layer; i += nthreads) {
    const int l = int(i / per_layer);
    const int o = int((i % per_layer) / H);  // out feature
    const int k = int(i % H);                // in feature
    const float v = (o < Hr && k < Hr) ? p.omega(l + 1) * p.params[p.off[2 * (l + 1)] + (long long)o * Hr + k] : 0.f;
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


// ------------------------------------------------------------------ fused optimiser step (raw-coordinate SIREN)
// ONE launch for what the reference does with opt.step() + opt.zero_grad() (INR/superresDWI.py:136-138) plus the
// bf16 re-staging of the operands: the thread that would pack parameter i first applies Adam to it (same arithmetic
// as adam_kernel in elementwise.cu, torch.optim.Adam defaults), clears its gradient for the next step, and writes
// the updated value into every operand buffer it belongs to.  Each parameter is visited by exactly one thread.
// The step counter is advanced by the last block to finish (ticket), so the launch is graph-capturable and there is
// no separate zero / tick / pack kernel (4 launches -> 1: the launch-latency regime of cfg1 and of strong-scaled
// shards).  The loss accumulator that rides behind the gradients is moved to loss_out and cleared as well.
struct AdamArgs {
  float* params;
  float* grads;     // single GPU: read and cleared.  Multi GPU: the buffer to clear for the NEXT step (other parity)
  float* m;
  float* v;
  float* state;     // {step, ticket (u32), -, -}
  float* loss_out;  // nullptr, or receives grads[n] (the step's loss) before it is cleared
  long long n;
  float lr, beta1, beta2, eps;
  // multi GPU (peer_grads != nullptr): every rank's [grad | loss] buffer of this step in peer-mapped memory, summed here in rank
  // order (the same order on every rank: replicated weights stay bit-identical), and the flag words of the start barrier
  const float* const* peer_grads;  // device array [world]
  uint32_t* const* peer_flags;     // device array [world]: flags[p][r] = last epoch rank r has announced to rank p
  int world, rank;
};

__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Start barrier of the multi-GPU optimiser step: rank r announces epoch = step + 1 to every peer, and every block
// waits until all peers have announced it -- i.e. until every rank has ENTERED this kernel, which in stream order means
// its backward of this step is complete (its gradients are final) and its previous optimiser step is over (nobody reads
// the other-parity buffer any more, so it may be cleared).  One kernel per GPU, all resident at the same time: the
// ranks are different devices.  Bounded spin: a missing rank traps instead of hanging the box.
__device__ __forceinline__ void peer_barrier(const AdamArgs& a) {
  if (a.peer_grads == nullptr) return;
  const uint32_t epoch = uint32_t(a.state[0]) + 1u;
  if (blockIdx.x == 0 && int(threadIdx.x) < a.world) st_release_sys_u32(a.peer_flags[threadIdx.x] + a.rank, epoch);
  if (int(threadIdx.x) < a.world) {
    const uint32_t* f = a.peer_flags[a.rank] + threadIdx.x;
    uint32_t spins = 0;
    while (ld_acquire_sys_u32(f) < epoch) {
      __nanosleep(64);
      if (++spins > (1u << 24)) {
        printf("b200inr: optimiser-step barrier timeout: rank %d waits for rank %d (epoch %u)\n", a.rank, threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

struct AdamCoef {
  float step_size, bc2_sqrt, omb1, omb2, beta2, eps;
};

__device__ __forceinline__ AdamCoef adam_coefficients(const AdamArgs& a, float* s_tmp /* 2 shared floats */) {
  if (threadIdx.x == 0) {
    const double step = double(a.state[0]) + 1.0;
    const double bc1 = 1.0 - pow(double(a.beta1), step);
    const double bc2 = 1.0 - pow(double(a.beta2), step);
    s_tmp[0] = float(double(a.lr) / bc1);
    s_tmp[1] = float(sqrt(bc2));
  }
  __syncthreads();
  AdamCoef c;
  c.step_size = s_tmp[0];
  c.bc2_sqrt = s_tmp[1];
  c.omb1 = 1.f - a.beta1;
  c.omb2 = 1.f - a.beta2;
  c.beta2 = a.beta2;
  c.eps = a.eps;
  return c;
}

__device__ __forceinline__ float adam_apply(const AdamArgs& a, const AdamCoef& c, long long i) {
  float gi;
  if (a.peer_grads != nullptr) {
    // all peer loads are issued before the first add (a load-add-load-add chain would pay one NVLink round trip per
    // rank); the sum is taken in rank order on every rank
    gi = 0.f;
    for (int r0 = 0; r0 < a.world; r0 += 8) {
      float part[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) part[r] = (r0 + r < a.world) ? ld_relaxed_sys_f32(a.peer_grads[r0 + r] + i) : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) gi += part[r];
    }
  } else {
    gi = a.grads[i];
  }
  float mi = a.m[i], vi = a.v[i];
  mi = fmaf(gi - mi, c.omb1, mi);
  vi = fmaf(c.omb2 * gi, gi, c.beta2 * vi);
  const float denom = __fsqrt_rn(vi) / c.bc2_sqrt + c.eps;
  const float pi = a.params[i] - c.step_size * (mi / denom);
  a.m[i] = mi;
  a.v[i] = vi;
  a.params[i] = pi;
  a.grads[i] = 0.f;
  return pi;
}

// last block out: advance the step counter, move the loss, reset the ticket
__device__ __forceinline__ void adam_finish(const AdamArgs& a) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(a.state + 1);
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    a.state[0] += 1.f;
    *reinterpret_cast<unsigned int*>(a.state + 1) = 0u;
    float loss;
    if (a.peer_grads != nullptr) {
      loss = 0.f;
      for (int r0 = 0; r0 < a.world; r0 += 8) {  // (loads first, then the sum: see adam_apply)
        float part[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) part[r] = (r0 + r < a.world) ? ld_relaxed_sys_f32(a.peer_grads[r0 + r] + a.n) : 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) loss += part[r];
      }
    } else {
      loss = a.grads[a.n];
    }
    if (a.loss_out != nullptr) a.loss_out[0] = loss;
    a.grads[a.n] = 0.f;
  }
}

// Work items (one per thread, so that every thread's load -> update -> store chain is a single short one and the
// whole kernel is one wave of independent chains): [0, 32 H) first layer, one warp per output feature (lanes = the 32
// live columns of the hi/lo operand row; lanes j < d own W0[o][j], lane 16 owns b0[o], the values travel by shuffle);
// then L H^2 hidden weights, 64 H final-linear entries, L H + 32 remaining biases.  H = kSirenWidth (operand width).
constexpr int kAdamPackThreads = 256;
__host__ __device__ inline long long siren_adam_pack_items(int L) {
  constexpr long long H = kSirenWidth;
  return 32 * H + (long long)L * H * H + kDzoPad * H + (long long)L * H + 32;
}

__global__ void __launch_bounds__(kAdamPackThreads) siren_adam_pack_kernel(const PackParams p, const AdamArgs a) {
  __shared__ float s_coef[2];
  peer_barrier(a);
  const AdamCoef c = adam_coefficients(a, s_coef);
  constexpr int H = kSirenWidth;
  const int L = p.L, C = p.C, d = p.d, Hr = p.Hr;
  float* bias = reinterpret_cast<float*>(p.packed + p.pl.bias);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;

  if (i < 32 * H) {
    // ---- first layer: row o of the float4 table, of the hi/lo tensor-core operand and of the bias table
    const int o = int(i >> 5), k = int(i & 31);
    float mine = 0.f;
    if (o < Hr) {
      if (k < d) mine = p.omega0 * adam_apply(a, c, p.off[0] + (long long)o * d + k);
      if (k == 16) mine = p.omega0 * adam_apply(a, c, p.off[1] + o);
    }
    const float wj = __shfl_sync(0xffffffffu, mine, k & 3);
    const float b = __shfl_sync(0xffffffffu, mine, 16);
    const float w0 = __shfl_sync(0xffffffffu, mine, 0), w1 = __shfl_sync(0xffffffffu, mine, 1);
    const float w2 = __shfl_sync(0xffffffffu, mine, 2), w3 = __shfl_sync(0xffffffffu, mine, 3);
    if (k == 0) {
      reinterpret_cast<float4*>(p.packed + p.pl.w0)[o] = make_float4(w0, w1, w2, w3);
      bias[o] = b;
      reinterpret_cast<__half*>(p.packed + p.pl.bias16)[o] = __float2half_rn(b);
    }
    float v = 0.f;  // columns 4g + j (g = 0..3: hi lo hi lo), 16 = b_hi, 17 = b_lo, the rest zero (32..63 stay zero)
    if (k < 16) {
      const float hi = __bfloat162float(__float2bfloat16_rn(wj));
      v = ((k >> 2) & 1) ? (wj - hi) : hi;
    } else if (k < 18) {
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      v = (k == 17) ? (b - hi) : hi;
    }
    put_bf16(p.packed + p.pl.w0p, uint32_t(o), uint32_t(k), v);
  } else if ((i -= 32 * H) < (long long)L * H * H) {
    // ---- hidden layers, both orientations
    const int l = int(i >> 16), o = int(i >> 8) & (H - 1), k = int(i) & (H - 1);
    static_assert(H == 256, "index arithmetic above");
    const float v = (o < Hr && k < Hr) ? p.omega(l + 1) * adam_apply(a, c, p.off[2 * (l + 1)] + (long long)o * Hr + k) : 0.f;
    uint8_t* wh = p.packed + p.pl.wh + size_t(l) * H * H * 2;
    uint8_t* wht = p.packed + p.pl.wht + size_t(l) * H * H * 2;
    put_bf16(wh + size_t(k >> 6) * H * 128, o, k & 63, v);   // forward: N = out, K = in
    put_bf16(wht + size_t(o >> 6) * H * 128, k, o & 63, v);  // dgrad:   N = in,  K = out
  } else if ((i -= (long long)L * H * H) < kDzoPad * H) {
    // ---- final linear: forward operand [H/64][32][64] (rows >= C zero) and dgrad operand [H][64] (N = in, K = c)
    const int cc = int(i >> 8), k = int(i) & (H - 1);
    const float v = (cc < C && k < Hr) ? adam_apply(a, c, p.off[2 * (L + 1)] + (long long)cc * Hr + k) : 0.f;
    if (cc < kOutPad) put_bf16(p.packed + p.pl.wf + size_t(k >> 6) * kOutPad * 128, cc, k & 63, v);
    put_bf16(p.packed + p.pl.wft, k, cc, v);
  } else if ((i -= kDzoPad * H) < (long long)L * H + 32) {
    // ---- biases of the hidden sine layers, then the final bias (padded to 32)
    float v = 0.f;
    if (i < (long long)L * H) {
      const int l = int(i >> 8) + 1, h = int(i) & (H - 1);
      if (h < Hr) v = p.omega(l) * adam_apply(a, c, p.off[2 * l + 1] + h);
    } else {
      const int cc = int(i - (long long)L * H);
      if (cc < C) v = adam_apply(a, c, p.off[2 * (L + 1) + 1] + cc);
    }
    bias[H + i] = v;
    reinterpret_cast<__half*>(p.packed + p.pl.bias16)[H + i] = __float2half_rn(v);
  }
  adam_finish(a);
}

// The other families: Adam + gradient clearing + step counter in one launch (the packing stays a second one).
__global__ void __launch_bounds__(256) adam_zero_kernel(const AdamArgs a) {
  __shared__ float s_coef[2];
  peer_barrier(a);
  const AdamCoef c = adam_coefficients(a, s_coef);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += stride) adam_apply(a, c, i);
  adam_finish(a);
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
  p.relu_tail = (net->flags & B200INR_NET_RELU_TAIL) != 0;
  p.pl = make_pack_layout(p.H, p.L);
  int64_t off[2 * (kMaxSineLayers + 1)];
  param_offsets(p.d, p.Hr, p.L, p.C, off);
  for (int i = 0; i < 2 * (p.L + 2); ++i) p.off[i] = off[i];
  pack_kernel<<<296, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
}

int launch_wire_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream);  // wire.cu

int launch_optimizer_step(const b200inr_net* net, float* params, float* grads, float* m, float* v, int64_t n, float lr,
                          float beta1, float beta2, float eps, float* state, void* packed, float* loss_out,
                          const float* const* peer_grads, uint32_t* const* peer_flags, int world, int rank,
                          cudaStream_t stream) {
  AdamArgs a{params, grads, m, v, state, loss_out, (long long)n, lr, beta1, beta2, eps, peer_grads, peer_flags, world, rank};
  if (net->input_mode == B200INR_IN_COORDS && net->activation == B200INR_ACT_SINE) {
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
    p.relu_tail = (net->flags & B200INR_NET_RELU_TAIL) != 0;
    p.pl = make_pack_layout(p.H, p.L);
    int64_t off[2 * (kMaxSineLayers + 1)];
    param_offsets(p.d, p.Hr, p.L, p.C, off);
    for (int i = 0; i < 2 * (p.L + 2); ++i) p.off[i] = off[i];
    const long long items = siren_adam_pack_items(p.L);
    siren_adam_pack_kernel<<<int((items + kAdamPackThreads - 1) / kAdamPackThreads), kAdamPackThreads, 0, stream>>>(p, a);
    return cudaGetLastError() == cudaSuccess ? B200INR_OK : B200INR_ERR_CUDA;
  }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_zero_kernel<<<int(blocks), 256, 0, stream>>>(a);
  if (cudaGetLastError() != cudaSuccess) return B200INR_ERR_CUDA;
  if (net->activation == B200INR_ACT_GABOR) return launch_wire_pack(net, params, packed, stream);
  return launch_pack(net, params, packed, stream);
}


}  // namespace b200inr
