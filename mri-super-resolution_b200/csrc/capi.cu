// capi.cu -- the extern "C" boundary declared in include/b200inr.h: argument checks, launch, error codes.
// No state is kept between calls except the cached SM count of the current device.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b200inr {
int launch_selftest_umma(int mode, const void* a, const void* b, float* d, int N, int K, int lbo_a, int sbo_a,
                         int lbo_b, int sbo_b, cudaStream_t stream);
int launch_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream);
int launch_siren_fwd(const b200inr_net* net, const void* packed, const float* coords, const b200inr_grid* grid,
                     int64_t rows, float* out, int clamp, float clamp_min, void* stash, int num_sms,
                     cudaStream_t stream);
bool fwd_pool_loss_supported(const b200inr_net* net, const b200inr_grid* grid, int64_t rows);
int launch_siren_fwd_pool_loss(const b200inr_net* net, const void* packed, const b200inr_grid* grid, int64_t rows,
                               const float* target_lr, double count, float* grad_out, float* loss_accum, void* stash,
                               int num_sms, cudaStream_t stream);
int launch_siren_bwd(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                     int num_sms, cudaStream_t stream);
int launch_siren_bwdp(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                      const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params, int num_sms,
                      cudaStream_t stream);
int launch_siren_wgrad(const b200inr_net* net, void* stash, const float* coords, const b200inr_grid* grid,
                       int64_t rows, float* grad_params, int num_sms, cudaStream_t stream);
int launch_gen_fwd(const b200inr_net* net, const void* packed, const float* coords, const b200inr_grid* grid,
                   int64_t rows, float* out, int clamp, float clamp_min, void* stash, int num_sms,
                   cudaStream_t stream);
int launch_gen_bwd(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                   float* grad_in, int num_sms, cudaStream_t stream, float* grad_coords = nullptr,
                   const float* coords = nullptr, const b200inr_grid* grid = nullptr, const float* out_y = nullptr);
int launch_pn_effective(const float* master, int64_t n_net, int64_t bias_off, int H, float acq, float* eff,
                        float* clear, cudaStream_t stream);
int launch_pn_fold_grad(float* grads, int64_t n_net, int64_t bias_off, int H, float acq, cudaStream_t stream);
int launch_gen_wgrad(const b200inr_net* net, void* stash, int64_t rows, float* grad_params, int num_sms,
                     cudaStream_t stream);
int launch_wire_pack(const b200inr_net* net, const float* params, void* packed, cudaStream_t stream);
int launch_wire_fwd(const b200inr_net* net, const void* packed, const float* coords, const b200inr_grid* grid,
                    int64_t rows, float* out, int clamp, float clamp_min, void* stash, int num_sms,
                    cudaStream_t stream);
int launch_wire_bwd(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                    float* grad_in, int num_sms, cudaStream_t stream);
int launch_wire_wgrad(const b200inr_net* net, void* stash, int64_t rows, int num_sms, cudaStream_t stream);
int launch_wire_combine(const b200inr_net* net, void* stash, int64_t rows, float* grad_params, cudaStream_t stream);
int launch_mse(const float* pred, const float* target, const float* weight, int64_t n, double count, float* grad,
               float* loss_accum, cudaStream_t stream, int relu_out = 0);
int launch_soft_erd(const float* signal, const float* b0, int64_t voxels, int n, double noise_level, double mul,
                    double slope, float* weights, float* soft_mean, cudaStream_t stream);
int launch_pool_mse(const float* pred, const float* target, int X, int Y, int64_t ZC, double count, float* grad,
                    float* loss_accum, cudaStream_t stream);
int launch_taps(const float* in, float* out, int in_y, int out_x, int out_y, int64_t ZC, const b200inr_axis_taps* tx,
                const b200inr_axis_taps* ty, cudaStream_t stream);
int launch_blurpool_mse(const float* pred, const float* target, int X, int Y, int64_t ZC, double count,
                        const float* bx6, const float* by6, const float* ax3, const float* ay3, float* resid,
                        float* grad, float* loss_accum, int x_begin, int x_end, cudaStream_t stream);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float* state, cudaStream_t stream);
int launch_optimizer_step(const b200inr_net* net, float* params, float* grads, float* m, float* v, int64_t n, float lr,
                          float beta1, float beta2, float eps, float* state, void* packed, float* loss_out,
                          const float* const* peer_grads, uint32_t* const* peer_flags, int world, int rank,
                          cudaStream_t stream);
int launch_sine_pre(const float* x, const float* W, const float* b, int64_t rows, int d, int H, float omega, float* out,
                    cudaStream_t stream);
int launch_mgrid(const GridDesc& g, int64_t rows, float* coords, cudaStream_t stream);
int launch_ffm(const float* x, const float* B, int64_t rows, int d, int m, float* out, cudaStream_t stream);
int launch_combinations(const float* b0, const float* b1, const float* b2, const float* b3, int64_t voxels, int n1, int n2,
                        int n3, float* out, cudaStream_t stream);
int launch_adc(const float* signal, const float* bvalues_host, int64_t voxels, int nb, float* adc, cudaStream_t stream);
int launch_ffm_bwd(const float* x, const float* B, const float* grad_out, int64_t rows, int d, int m, float* grad_x,
                   cudaStream_t stream);

static bool is_gen(const b200inr_net* net) { return net->input_mode != B200INR_IN_COORDS; }
static bool is_wire(const b200inr_net* net) { return net->activation == B200INR_ACT_GABOR; }
// SIREN on raw coordinates trained through the layer-pipelined backward (mlp_bwdp.cu) unless the staged path is asked for
static bool is_piped(const b200inr_net* net) {
  return !is_gen(net) && !is_wire(net) && (net->flags & B200INR_NET_STAGED_BWD) == 0;
}

static int check_net(const b200inr_net* net) {
  if (!net) return B200INR_ERR_NULL;
  if (net->hidden_layers < 0 || net->hidden_layers + 1 > kMaxSineLayers) return B200INR_ERR_BAD_SHAPE;
  if (net->out_features < 1 || net->out_features > kOutPad) return B200INR_ERR_BAD_SHAPE;
  const int H = net->hidden_features;
  if (net->activation == B200INR_ACT_GABOR) {  // WIRE: 128 complex units
    if (H != 128) return B200INR_ERR_BAD_SHAPE;
    switch (net->input_mode) {
      case B200INR_IN_COORDS:  // raw coordinates (first layer on CUDA cores)
        if (net->in_features < 1 || net->in_features > 4 || net->mapping_size != 0) return B200INR_ERR_BAD_SHAPE;
        return B200INR_OK;
      case B200INR_IN_FOURIER:  // input_mapping fused into the first layer (wiretest.ipynb cells 6-9)
        if (net->in_features < 1 || net->in_features > 4) return B200INR_ERR_BAD_SHAPE;
        if (net->mapping_size < 32 || net->mapping_size % 32 != 0 || net->mapping_size > 256) return B200INR_ERR_BAD_SHAPE;
        return B200INR_OK;
      case B200INR_IN_FEATURES:  // explicit feature rows: Siren(in_features=2*mapping_size, ...) of wiretest.ipynb cell 7
        if (net->in_features < 64 || net->in_features % 64 != 0 || net->in_features > 512 || net->mapping_size != 0)
          return B200INR_ERR_BAD_SHAPE;
        return B200INR_OK;
      default:
        return B200INR_ERR_BAD_SHAPE;
    }
  }
  if (net->activation != B200INR_ACT_SINE && net->activation != B200INR_ACT_RELU &&
      net->activation != B200INR_ACT_TANH)
    return B200INR_ERR_BAD_SHAPE;
  if ((net->flags & B200INR_NET_RELU_TAIL) && net->input_mode != B200INR_IN_COORDS) return B200INR_ERR_BAD_SHAPE;
  // tanh networks (the perturbation network), tanh output and the dgrad-only stash belong to the generic family
  if (net->input_mode == B200INR_IN_COORDS &&
      (net->activation == B200INR_ACT_TANH || (net->flags & (B200INR_NET_DGRAD_ONLY | B200INR_NET_TANH_OUT))))
    return B200INR_ERR_BAD_SHAPE;
  if (net->activation == B200INR_ACT_TANH && H != 256) return B200INR_ERR_BAD_SHAPE;
  if ((net->flags & B200INR_NET_TANH_OUT) && !(net->scale_0 != 0.f)) return B200INR_ERR_BAD_SHAPE;
  switch (net->input_mode) {
    case B200INR_IN_COORDS:  // SIREN on raw coordinates: any width up to 256 (multiple of 8) runs on the 256-wide
                             // kernels with zero-padded operands; parameters and gradients keep the real width
      if (H < 8 || H > kSirenWidth || H % 8 != 0 || net->activation != B200INR_ACT_SINE) return B200INR_ERR_BAD_SHAPE;
      if (net->in_features < 1 || net->in_features > 4 || net->mapping_size != 0) return B200INR_ERR_BAD_SHAPE;
      if (net->flags & B200INR_NET_RELU_TAIL)  // ReLU-tail SIREN: pipelined backward only, at least the ReLU layer
        if ((net->flags & B200INR_NET_STAGED_BWD) || net->hidden_layers < 1) return B200INR_ERR_BAD_SHAPE;
      return B200INR_OK;
    case B200INR_IN_FOURIER: {
      if (H != 256 && H != 512) return B200INR_ERR_BAD_SHAPE;
      if (net->in_features < 1 || net->in_features > 4) return B200INR_ERR_BAD_SHAPE;
      const int k0 = 2 * net->mapping_size;
      if (net->mapping_size < 32 || k0 % 64 != 0 || k0 > H) return B200INR_ERR_BAD_SHAPE;
      return B200INR_OK;
    }
    case B200INR_IN_FEATURES:
      if (H != 256 && H != 512) return B200INR_ERR_BAD_SHAPE;
      if (net->in_features < 64 || net->in_features % 64 != 0 || net->in_features > H || net->mapping_size != 0)
        return B200INR_ERR_BAD_SHAPE;
      return B200INR_OK;
    default:
      return B200INR_ERR_BAD_SHAPE;
  }
}

static int check_grid(const b200inr_net* net, const b200inr_grid* grid, int64_t rows) {
  if (net->input_mode == B200INR_IN_FEATURES) return B200INR_ERR_BAD_SHAPE;  // explicit features have no grid form
  if (grid->ndim != net->in_features) return B200INR_ERR_BAD_SHAPE;
  long long tot = 1;
  for (int j = 0; j < grid->ndim; ++j) {
    if (grid->shape[j] < 1) return B200INR_ERR_BAD_SHAPE;
    tot *= grid->shape[j];
  }
  if (grid->row_begin < 0 || grid->row_begin + rows > tot) return B200INR_ERR_BAD_SHAPE;
  return B200INR_OK;
}

// sm_100 only; SM count of the current device (cached per device ordinal).
static int device_sms(int* sms) {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B200INR_ERR_CUDA;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) {
    *sms = cached[dev];
    return B200INR_OK;
  }
  int major = 0, n = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return B200INR_ERR_CUDA;
  if (major != 10) return B200INR_ERR_UNSUPPORTED_ARCH;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return B200INR_ERR_CUDA;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  *sms = n;
  return B200INR_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b200inr

using namespace b200inr;

extern "C" {

const char* b200inr_version(void) { return "b200inr 0.1 (sm_100a)"; }

const char* b200inr_error_string(int code) {
  switch (code) {
    case B200INR_OK: return "ok";
    case B200INR_ERR_BAD_SHAPE: return "unsupported or inconsistent shape";
    case B200INR_ERR_BAD_ALIGN: return "pointer not 16-byte aligned";
    case B200INR_ERR_UNSUPPORTED_ARCH: return "device is not sm_100 (B200)";
    case B200INR_ERR_CUDA: return "CUDA runtime error";
    case B200INR_ERR_NULL: return "null pointer argument";
    default: return "unknown error";
  }
}

int b200inr_param_count(const b200inr_net* net, int64_t* n_floats) {
  int e = check_net(net);
  if (e) return e;
  if (!n_floats) return B200INR_ERR_NULL;
  if (is_wire(net))
    *n_floats = wire_param_offsets(make_wire_dims(net), nullptr);
  else if (is_gen(net))
    *n_floats = gen_param_offsets(make_gen_dims(net), nullptr);
  else
    *n_floats = param_offsets(net->in_features, net->hidden_features, net->hidden_layers, net->out_features, nullptr);
  return B200INR_OK;
}

int b200inr_param_offsets(const b200inr_net* net, int64_t* offsets) {
  int e = check_net(net);
  if (e) return e;
  if (!offsets) return B200INR_ERR_NULL;
  if (is_wire(net))
    wire_param_offsets(make_wire_dims(net), offsets);
  else if (is_gen(net))
    gen_param_offsets(make_gen_dims(net), offsets);
  else
    param_offsets(net->in_features, net->hidden_features, net->hidden_layers, net->out_features, offsets);
  return B200INR_OK;
}

int b200inr_packed_bytes(const b200inr_net* net, size_t* bytes) {
  int e = check_net(net);
  if (e) return e;
  if (!bytes) return B200INR_ERR_NULL;
  *bytes = is_wire(net) ? make_wire_pack_layout(make_wire_dims(net)).total
           : is_gen(net) ? make_gen_pack_layout(make_gen_dims(net)).total
                         : make_pack_layout(kSirenWidth, net->hidden_layers).total;
  return B200INR_OK;
}

int b200inr_pack_weights(const b200inr_net* net, const float* params, void* packed, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!params || !packed) return B200INR_ERR_NULL;
  if (!aligned16(params) || (reinterpret_cast<uintptr_t>(packed) & 1023)) return B200INR_ERR_BAD_ALIGN;
  if (is_wire(net)) return launch_wire_pack(net, params, packed, static_cast<cudaStream_t>(stream));
  return launch_pack(net, params, packed, static_cast<cudaStream_t>(stream));
}

int b200inr_stash_bytes(const b200inr_net* net, int64_t rows, size_t* bytes) {
  int e = check_net(net);
  if (e) return e;
  if (!bytes) return B200INR_ERR_NULL;
  if (rows < 0) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) {
    *bytes = 0;
    return B200INR_OK;
  }
  *bytes = is_wire(net) ? make_wire_stash_layout(make_wire_dims(net), rows).total
           : is_gen(net) ? make_gen_stash_layout(make_gen_dims(net), rows).total
           : is_piped(net) ? make_pipe_stash_layout(kSirenWidth, net->hidden_layers, rows).total
                           : make_stash_layout(kSirenWidth, net->hidden_layers, rows).total;
  return B200INR_OK;
}

int b200inr_siren_forward(const b200inr_net* net, const void* packed, const float* coords, const b200inr_grid* grid,
                          int64_t rows, float* out, int clamp, float clamp_min, void* stash, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !out) return B200INR_ERR_NULL;
  if ((coords == nullptr) == (grid == nullptr)) return B200INR_ERR_NULL;
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if (grid && (e = check_grid(net, grid, rows))) return e;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (stash && (reinterpret_cast<uintptr_t>(stash) & 1023)))
    return B200INR_ERR_BAD_ALIGN;
  if (net->input_mode == B200INR_IN_FEATURES && !aligned16(coords)) return B200INR_ERR_BAD_ALIGN;  // float4 row loads
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  if (is_wire(net))
    return launch_wire_fwd(net, packed, coords, grid, rows, out, clamp, clamp_min, stash, sms,
                           static_cast<cudaStream_t>(stream));
  if (is_gen(net))
    return launch_gen_fwd(net, packed, coords, grid, rows, out, clamp, clamp_min, stash, sms,
                          static_cast<cudaStream_t>(stream));
  return launch_siren_fwd(net, packed, coords, grid, rows, out, clamp, clamp_min, stash, sms,
                          static_cast<cudaStream_t>(stream));
}

int b200inr_siren_forward_pool_loss(const b200inr_net* net, const void* packed, const b200inr_grid* grid, int64_t rows,
                                    const float* target_lr, double count, float* grad_hr, float* loss_accum,
                                    void* stash, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !grid || !target_lr || !grad_hr || !loss_accum || !stash) return B200INR_ERR_NULL;
  if (rows < 0 || rows > (int64_t(1) << 37) || !(count > 0)) return B200INR_ERR_BAD_SHAPE;
  if ((e = check_grid(net, grid, rows))) return e;
  if (!fwd_pool_loss_supported(net, grid, rows)) return B200INR_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (reinterpret_cast<uintptr_t>(stash) & 1023))
    return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  return launch_siren_fwd_pool_loss(net, packed, grid, rows, target_lr, count, grad_hr, loss_accum, stash, sms,
                                    static_cast<cudaStream_t>(stream));
}

int b200inr_siren_backward(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                           const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params,
                           void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !stash || !grad_out || !grad_params) return B200INR_ERR_NULL;
  if ((coords == nullptr) == (grid == nullptr)) return B200INR_ERR_NULL;
  // tanh output: the derivative needs the forward's output (b200inr_siren_backward_tanh_out); dgrad-only stash: there
  // is nothing to contract weight gradients from (b200inr_siren_backward_coords)
  if (net->flags & (B200INR_NET_TANH_OUT | B200INR_NET_DGRAD_ONLY)) return B200INR_ERR_BAD_SHAPE;
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if (grid && (e = check_grid(net, grid, rows))) return e;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (reinterpret_cast<uintptr_t>(stash) & 1023) ||
      !aligned16(grad_params))
    return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_wire(net)) {
    if ((e = launch_wire_bwd(net, packed, stash, rows, grad_out, nullptr, sms, s))) return e;
    if ((e = launch_wire_wgrad(net, stash, rows, sms, s))) return e;
    return launch_wire_combine(net, stash, rows, grad_params, s);
  }
  if (is_gen(net)) {
    if ((e = launch_gen_bwd(net, packed, stash, rows, grad_out, nullptr, sms, s))) return e;
    return launch_gen_wgrad(net, stash, rows, grad_params, sms, s);
  }
  if (is_piped(net) && !aligned16(grad_out)) return B200INR_ERR_BAD_ALIGN;  // dOut tiles are bulk-copied
  if (is_piped(net)) return launch_siren_bwdp(net, packed, stash, coords, grid, rows, grad_out, grad_params, sms, s);
  if ((e = launch_siren_bwd(net, packed, stash, rows, grad_out, sms, s))) return e;
  return launch_siren_wgrad(net, stash, coords, grid, rows, grad_params, sms, s);
}

int b200inr_siren_backward_input(const b200inr_net* net, const void* packed, void* stash, int64_t rows,
                                 const float* grad_out, float* grad_params, float* grad_input, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !stash || !grad_out || !grad_params || !grad_input) return B200INR_ERR_NULL;
  if (net->input_mode != B200INR_IN_FEATURES) return B200INR_ERR_BAD_SHAPE;  // explicit feature rows only
  if (net->flags & (B200INR_NET_TANH_OUT | B200INR_NET_DGRAD_ONLY)) return B200INR_ERR_BAD_SHAPE;
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (reinterpret_cast<uintptr_t>(stash) & 1023) ||
      !aligned16(grad_params) || !aligned16(grad_input))
    return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_wire(net)) {
    if ((e = launch_wire_bwd(net, packed, stash, rows, grad_out, grad_input, sms, s))) return e;
    if ((e = launch_wire_wgrad(net, stash, rows, sms, s))) return e;
    return launch_wire_combine(net, stash, rows, grad_params, s);
  }
  if ((e = launch_gen_bwd(net, packed, stash, rows, grad_out, grad_input, sms, s))) return e;
  return launch_gen_wgrad(net, stash, rows, grad_params, sms, s);
}

int b200inr_siren_backward_coords(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                                  const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params,
                                  float* grad_coords, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !stash || !grad_out || !grad_coords) return B200INR_ERR_NULL;
  if ((coords == nullptr) == (grid == nullptr)) return B200INR_ERR_NULL;
  if (is_wire(net) || net->input_mode != B200INR_IN_FOURIER) return B200INR_ERR_BAD_SHAPE;
  if ((net->flags & B200INR_NET_TANH_OUT)) return B200INR_ERR_BAD_SHAPE;
  if (((net->flags & B200INR_NET_DGRAD_ONLY) != 0) != (grad_params == nullptr)) return B200INR_ERR_BAD_SHAPE;
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if (grid && (e = check_grid(net, grid, rows))) return e;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (reinterpret_cast<uintptr_t>(stash) & 1023) ||
      (grad_params && !aligned16(grad_params)))
    return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((e = launch_gen_bwd(net, packed, stash, rows, grad_out, nullptr, sms, s, grad_coords, coords, grid))) return e;
  if (grad_params) return launch_gen_wgrad(net, stash, rows, grad_params, sms, s);
  return B200INR_OK;
}

int b200inr_siren_backward_tanh_out(const b200inr_net* net, const void* packed, void* stash, int64_t rows,
                                    const float* out, const float* grad_out, float* grad_params, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !stash || !out || !grad_out || !grad_params) return B200INR_ERR_NULL;
  if (is_wire(net) || !is_gen(net) || !(net->flags & B200INR_NET_TANH_OUT) || (net->flags & B200INR_NET_DGRAD_ONLY))
    return B200INR_ERR_BAD_SHAPE;
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (reinterpret_cast<uintptr_t>(stash) & 1023) ||
      !aligned16(grad_params))
    return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((e = launch_gen_bwd(net, packed, stash, rows, grad_out, nullptr, sms, s, nullptr, nullptr, nullptr, out))) return e;
  return launch_gen_wgrad(net, stash, rows, grad_params, sms, s);
}

int b200inr_pn_effective_params(const float* master, int64_t n_net, int64_t bias_off, int32_t H, float acq, float* eff,
                                float* clear_grads, void* stream) {
  if (!master || !eff) return B200INR_ERR_NULL;
  if (n_net < 1 || bias_off < 0 || H < 1 || bias_off + H > n_net) return B200INR_ERR_BAD_SHAPE;
  return launch_pn_effective(master, n_net, bias_off, H, acq, eff, clear_grads, static_cast<cudaStream_t>(stream));
}

int b200inr_pn_fold_grad(float* grads, int64_t n_net, int64_t bias_off, int32_t H, float acq, void* stream) {
  if (!grads) return B200INR_ERR_NULL;
  if (n_net < 1 || bias_off < 0 || H < 1 || bias_off + H > n_net) return B200INR_ERR_BAD_SHAPE;
  return launch_pn_fold_grad(grads, n_net, bias_off, H, acq, static_cast<cudaStream_t>(stream));
}

int b200inr_siren_dgrad(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                        void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!packed || !stash || !grad_out) return B200INR_ERR_NULL;
  if (is_piped(net)) return B200INR_ERR_BAD_SHAPE;  // one-kernel backward: use b200inr_siren_backward
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if ((reinterpret_cast<uintptr_t>(packed) & 1023) || (reinterpret_cast<uintptr_t>(stash) & 1023))
    return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  if (is_wire(net)) return launch_wire_bwd(net, packed, stash, rows, grad_out, nullptr, sms, static_cast<cudaStream_t>(stream));
  if (is_gen(net))
    return launch_gen_bwd(net, packed, stash, rows, grad_out, nullptr, sms, static_cast<cudaStream_t>(stream));
  return launch_siren_bwd(net, packed, stash, rows, grad_out, sms, static_cast<cudaStream_t>(stream));
}

int b200inr_siren_wgrad(const b200inr_net* net, void* stash, const float* coords, const b200inr_grid* grid,
                        int64_t rows, float* grad_params, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!stash || !grad_params) return B200INR_ERR_NULL;
  if (is_piped(net)) return B200INR_ERR_BAD_SHAPE;  // one-kernel backward: use b200inr_siren_backward
  if ((coords == nullptr) == (grid == nullptr)) return B200INR_ERR_NULL;
  if (rows < 0 || rows > (int64_t(1) << 37)) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  if (grid && (e = check_grid(net, grid, rows))) return e;
  if ((reinterpret_cast<uintptr_t>(stash) & 1023) || !aligned16(grad_params)) return B200INR_ERR_BAD_ALIGN;
  int sms = 0;
  if ((e = device_sms(&sms))) return e;
  if (is_wire(net)) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((e = launch_wire_wgrad(net, stash, rows, sms, s))) return e;
    return launch_wire_combine(net, stash, rows, grad_params, s);
  }
  if (is_gen(net)) return launch_gen_wgrad(net, stash, rows, grad_params, sms, static_cast<cudaStream_t>(stream));
  return launch_siren_wgrad(net, stash, coords, grid, rows, grad_params, sms, static_cast<cudaStream_t>(stream));
}

int b200inr_mse_loss(const float* pred, const float* target, const float* weight, int64_t n, double count,
                     float* grad, float* loss_accum, void* stream) {
  if (!pred || !target) return B200INR_ERR_NULL;
  if (n < 0 || !(count > 0)) return B200INR_ERR_BAD_SHAPE;
  if (n == 0) return B200INR_OK;
  return launch_mse(pred, target, weight, n, count, grad, loss_accum, static_cast<cudaStream_t>(stream));
}

int b200inr_mse_loss_relu_out(const float* pred, const float* target, const float* weight, int64_t n, double count,
                              float* grad, float* loss_accum, void* stream) {
  if (!pred || !target) return B200INR_ERR_NULL;
  if (n < 0 || !(count > 0)) return B200INR_ERR_BAD_SHAPE;
  if (n == 0) return B200INR_OK;
  return launch_mse(pred, target, weight, n, count, grad, loss_accum, static_cast<cudaStream_t>(stream), 1);
}

int b200inr_soft_erd(const float* signal, const float* b0, int64_t voxels, int32_t n, double noise_level, double mul,
                     double slope, float* weights, float* soft_mean, void* stream) {
  if (!signal || !b0) return B200INR_ERR_NULL;
  if (voxels < 0 || n < 2 || n > 64) return B200INR_ERR_BAD_SHAPE;
  if (voxels == 0) return B200INR_OK;
  return launch_soft_erd(signal, b0, voxels, n, noise_level, mul, slope, weights, soft_mean,
                         static_cast<cudaStream_t>(stream));
}

int b200inr_degrade_build_axis_host(int32_t n_hr, int blur, b200inr_axis_taps* fwd_host, b200inr_axis_taps* adj_host) {
  if (!fwd_host || !adj_host) return B200INR_ERR_NULL;
  if (n_hr < 2 || (n_hr & 1)) return B200INR_ERR_BAD_SHAPE;
  const int n_lr = n_hr / 2;
  // Gaussian sigma = 0.5, truncate 4.0 -> radius 2 (scipy.ndimage.gaussian_filter), normalised in double.
  double g[5];
  double gs = 0.0;
  for (int t = -2; t <= 2; ++t) {
    g[t + 2] = exp(-0.5 * double(t * t) / 0.25);
    gs += g[t + 2];
  }
  for (int t = 0; t < 5; ++t) g[t] /= gs;
  auto mirror = [n_hr](int i) {  // scipy 'mirror' == numpy 'reflect': d c b | a b c d | c b a
    if (n_hr == 1) return 0;
    const int period = 2 * (n_hr - 1);
    i %= period;
    if (i < 0) i += period;
    return i < n_hr ? i : period - i;
  };
  memset(fwd_host, 0, sizeof(b200inr_axis_taps) * size_t(n_lr));
  memset(adj_host, 0, sizeof(b200inr_axis_taps) * size_t(n_hr));
  for (int i = 0; i < n_lr; ++i) {
    double acc_w[B200INR_DEGRADE_MAX_TAPS];
    int acc_i[B200INR_DEGRADE_MAX_TAPS];
    int cnt = 0;
    auto push = [&](int idx, double w) {
      for (int k = 0; k < cnt; ++k)
        if (acc_i[k] == idx) {
          acc_w[k] += w;
          return true;
        }
      if (cnt == B200INR_DEGRADE_MAX_TAPS) return false;
      acc_i[cnt] = idx;
      acc_w[cnt] = w;
      ++cnt;
      return true;
    };
    for (int s = 0; s < 2; ++s) {
      const int x = 2 * i + s;
      if (blur) {
        for (int t = -2; t <= 2; ++t)
          if (!push(mirror(x + t), 0.5 * g[t + 2])) return B200INR_ERR_BAD_SHAPE;
      } else {
        if (!push(x, 0.5)) return B200INR_ERR_BAD_SHAPE;
      }
    }
    for (int k = 0; k < cnt; ++k) {
      fwd_host[i].idx[k] = acc_i[k];
      fwd_host[i].w[k] = float(acc_w[k]);
      // transpose entry
      b200inr_axis_taps& a = adj_host[acc_i[k]];
      int slot = -1;
      for (int q = 0; q < B200INR_DEGRADE_MAX_TAPS; ++q)
        if (a.w[q] == 0.f) {
          slot = q;
          break;
        }
      if (slot < 0) return B200INR_ERR_BAD_SHAPE;
      a.idx[slot] = i;
      a.w[slot] = float(acc_w[k]);
    }
  }
  return B200INR_OK;
}

int b200inr_degrade_build_band_host(int32_t n_hr, int blur, float* fwd6_host, float* adj3_host) {
  if (!fwd6_host || !adj3_host) return B200INR_ERR_NULL;
  if (n_hr < 2 || (n_hr & 1) || n_hr > (1 << 20)) return B200INR_ERR_BAD_SHAPE;
  const int n_lr = n_hr / 2;
  b200inr_axis_taps* fwd = new b200inr_axis_taps[n_lr];
  b200inr_axis_taps* adj = new b200inr_axis_taps[n_hr];
  int e = b200inr_degrade_build_axis_host(n_hr, blur, fwd, adj);
  if (e == B200INR_OK) {
    memset(fwd6_host, 0, sizeof(float) * size_t(n_lr) * 6);
    memset(adj3_host, 0, sizeof(float) * size_t(n_hr) * 3);
    for (int i = 0; i < n_lr && e == B200INR_OK; ++i)
      for (int k = 0; k < B200INR_DEGRADE_MAX_TAPS; ++k) {
        if (fwd[i].w[k] == 0.f) continue;
        const int x = fwd[i].idx[k];
        const int a = x - (2 * i - 2);    // position inside LR row i's band of HR rows
        const int t = i - ((x - 2) >> 1);  // position inside HR row x's band of LR rows
        if (a < 0 || a >= 6 || t < 0 || t >= 3) {
          e = B200INR_ERR_BAD_SHAPE;
          break;
        }
        fwd6_host[i * 6 + a] += fwd[i].w[k];
        adj3_host[x * 3 + t] += fwd[i].w[k];
      }
  }
  delete[] fwd;
  delete[] adj;
  return e;
}

int b200inr_blurpool_mse(const float* pred_hr, const float* target_lr, int32_t X, int32_t Y, int64_t ZC, double count,
                         const float* bx6, const float* by6, const float* ax3, const float* ay3, float* resid_lr,
                         float* grad_hr, float* loss_accum, void* stream) {
  if (!pred_hr || !target_lr || !bx6 || !by6 || !ax3 || !ay3 || !resid_lr) return B200INR_ERR_NULL;
  if (!(count > 0)) return B200INR_ERR_BAD_SHAPE;
  return launch_blurpool_mse(pred_hr, target_lr, X, Y, ZC, count, bx6, by6, ax3, ay3, resid_lr, grad_hr, loss_accum,
                             0, X, static_cast<cudaStream_t>(stream));
}

int b200inr_blurpool_mse_slab(const float* pred_ext, const float* target_ext, int32_t X, int32_t Y, int64_t ZC,
                              double count, const float* bx6, const float* by6, const float* ax3, const float* ay3,
                              int32_t x_begin, int32_t x_end, float* resid_ext, float* grad_hr, float* loss_accum,
                              void* stream) {
  if (!pred_ext || !target_ext || !bx6 || !by6 || !ax3 || !ay3 || !resid_ext) return B200INR_ERR_NULL;
  if (!(count > 0)) return B200INR_ERR_BAD_SHAPE;
  return launch_blurpool_mse(pred_ext, target_ext, X, Y, ZC, count, bx6, by6, ax3, ay3, resid_ext, grad_hr, loss_accum,
                             x_begin, x_end, static_cast<cudaStream_t>(stream));
}

int b200inr_degrade_forward(const float* hr, float* lr, int32_t X, int32_t Y, int64_t ZC, const b200inr_axis_taps* tx,
                            const b200inr_axis_taps* ty, void* stream) {
  if (!hr || !lr || !tx || !ty) return B200INR_ERR_NULL;
  if (X < 2 || Y < 2 || (X & 1) || (Y & 1) || ZC < 1) return B200INR_ERR_BAD_SHAPE;
  return launch_taps(hr, lr, Y, X / 2, Y / 2, ZC, tx, ty, static_cast<cudaStream_t>(stream));
}

int b200inr_degrade_adjoint(const float* lr, float* hr, int32_t X, int32_t Y, int64_t ZC, const b200inr_axis_taps* ax,
                            const b200inr_axis_taps* ay, void* stream) {
  if (!hr || !lr || !ax || !ay) return B200INR_ERR_NULL;
  if (X < 2 || Y < 2 || (X & 1) || (Y & 1) || ZC < 1) return B200INR_ERR_BAD_SHAPE;
  return launch_taps(lr, hr, Y / 2, X, Y, ZC, ax, ay, static_cast<cudaStream_t>(stream));
}

int b200inr_pool_mse(const float* pred_hr, const float* target_lr, int32_t X, int32_t Y, int64_t ZC, double count,
                     float* grad_hr, float* loss_accum, void* stream) {
  if (!pred_hr || !target_lr) return B200INR_ERR_NULL;
  if (!(count > 0)) return B200INR_ERR_BAD_SHAPE;
  return launch_pool_mse(pred_hr, target_lr, X, Y, ZC, count, grad_hr, loss_accum, static_cast<cudaStream_t>(stream));
}

int b200inr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float* state, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !state) return B200INR_ERR_NULL;
  if (n < 0) return B200INR_ERR_BAD_SHAPE;
  if (n == 0) return B200INR_OK;
  return launch_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, state,
                     static_cast<cudaStream_t>(stream));
}

int b200inr_optimizer_step(const b200inr_net* net, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                           float lr, float beta1, float beta2, float eps, float* state, void* packed, float* loss_out,
                           void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !state || !packed) return B200INR_ERR_NULL;
  if (!aligned16(params) || (reinterpret_cast<uintptr_t>(packed) & 1023)) return B200INR_ERR_BAD_ALIGN;
  int64_t n = 0;
  if ((e = b200inr_param_count(net, &n))) return e;
  return launch_optimizer_step(net, params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, state, packed,
                               loss_out, nullptr, nullptr, 1, 0, static_cast<cudaStream_t>(stream));
}

int b200inr_optimizer_step_peers(const b200inr_net* net, float* params, float* grads_next, const float* const* peer_grads,
                                 uint32_t* const* peer_flags, int32_t world, int32_t rank, float* exp_avg,
                                 float* exp_avg_sq, float lr, float beta1, float beta2, float eps, float* state,
                                 void* packed, float* loss_out, void* stream) {
  int e = check_net(net);
  if (e) return e;
  if (!params || !grads_next || !peer_grads || !peer_flags || !exp_avg || !exp_avg_sq || !state || !packed)
    return B200INR_ERR_NULL;
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return B200INR_ERR_BAD_SHAPE;
  if (!aligned16(params) || (reinterpret_cast<uintptr_t>(packed) & 1023)) return B200INR_ERR_BAD_ALIGN;
  int64_t n = 0;
  if ((e = b200inr_param_count(net, &n))) return e;
  // (world == 1 runs the same code path -- a sum over one "peer", this rank's own buffer -- and is how one GPU tests it)
  return launch_optimizer_step(net, params, grads_next, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, state, packed,
                               loss_out, peer_grads, peer_flags, world, rank,
                               static_cast<cudaStream_t>(stream));
}

size_t b200inr_net_size(void) { return sizeof(b200inr_net); }

int b200inr_param_offset_count(const b200inr_net* net, int32_t* count) {
  int e = check_net(net);
  if (e) return e;
  if (!count) return B200INR_ERR_NULL;
  if (is_wire(net))
    *count = 4 * (net->hidden_layers + 1) + 2 + (net->input_mode == B200INR_IN_FOURIER ? 1 : 0);
  else
    *count = 2 * (net->hidden_layers + 2) + (net->input_mode == B200INR_IN_FOURIER ? 1 : 0);
  return B200INR_OK;
}

int b200inr_get_mgrid(const b200inr_grid* grid, int64_t rows, float* coords, void* stream) {
  if (!grid || !coords) return B200INR_ERR_NULL;
  if (grid->ndim < 1 || grid->ndim > 4 || rows < 0) return B200INR_ERR_BAD_SHAPE;
  GridDesc g{};
  g.ndim = grid->ndim;
  long long tot = 1;
  for (int j = 0; j < 4; ++j) {
    g.shape[j] = (j < grid->ndim) ? grid->shape[j] : 1;
    if (g.shape[j] < 1) return B200INR_ERR_BAD_SHAPE;
    tot *= g.shape[j];
  }
  g.row_begin = grid->row_begin;
  g.total = tot;
  if (grid->row_begin < 0 || grid->row_begin + rows > tot) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  return launch_mgrid(g, rows, coords, static_cast<cudaStream_t>(stream));
}

int b200inr_input_mapping(const float* x, const float* B, int64_t rows, int32_t d, int32_t m, float* out,
                          void* stream) {
  if (!x || !B || !out) return B200INR_ERR_NULL;
  if (rows < 0 || d < 1 || m < 1) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  return launch_ffm(x, B, rows, d, m, out, static_cast<cudaStream_t>(stream));
}

int b200inr_sine_layer_pre(const float* x, const float* W, const float* b, int64_t rows, int32_t d, int32_t H,
                           float omega, float* out, void* stream) {
  if (!x || !W || !b || !out) return B200INR_ERR_NULL;
  if (rows < 0 || d < 1 || d > 4096 || H < 1) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  return launch_sine_pre(x, W, b, rows, d, H, omega, out, static_cast<cudaStream_t>(stream));
}

int b200inr_combinations(const float* b0, const float* b1, const float* b2, const float* b3, int64_t voxels, int32_t n1,
                         int32_t n2, int32_t n3, float* out, void* stream) {
  if (!b0 || !b1 || !b2 || !b3 || !out) return B200INR_ERR_NULL;
  if (voxels < 0 || n1 < 1 || n2 < 1 || n3 < 1 || int64_t(n1) * n2 * n3 > (1 << 20)) return B200INR_ERR_BAD_SHAPE;
  if (voxels == 0) return B200INR_OK;
  return launch_combinations(b0, b1, b2, b3, voxels, n1, n2, n3, out, static_cast<cudaStream_t>(stream));
}

int b200inr_adc_fit(const float* signal, const float* bvalues_host, int64_t voxels, int32_t nb, float* adc,
                    void* stream) {
  if (!signal || !bvalues_host || !adc) return B200INR_ERR_NULL;
  if (voxels < 0) return B200INR_ERR_BAD_SHAPE;
  if (voxels == 0) return (nb >= 2 && nb <= 64) ? B200INR_OK : B200INR_ERR_BAD_SHAPE;
  return launch_adc(signal, bvalues_host, voxels, nb, adc, static_cast<cudaStream_t>(stream));
}

int b200inr_input_mapping_backward(const float* x, const float* B, const float* grad_out, int64_t rows, int32_t d,
                                   int32_t m, float* grad_x, void* stream) {
  if (!x || !B || !grad_out || !grad_x) return B200INR_ERR_NULL;
  if (rows < 0 || d < 1 || d > 8 || m < 1) return B200INR_ERR_BAD_SHAPE;
  if (rows == 0) return B200INR_OK;
  return launch_ffm_bwd(x, B, grad_out, rows, d, m, grad_x, static_cast<cudaStream_t>(stream));
}

int b200inr_selftest_umma(int mode, const void* a_bf16, const void* b_bf16, float* d, int32_t N, int32_t K,
                          int32_t lbo_a, int32_t sbo_a, int32_t lbo_b, int32_t sbo_b, void* stream) {
  if (!a_bf16 || !b_bf16 || !d) return B200INR_ERR_NULL;
  if (mode != 0 && mode != 1) return B200INR_ERR_BAD_SHAPE;
  int sms = 0;
  int e = device_sms(&sms);
  if (e) return e;
  return launch_selftest_umma(mode, a_bf16, b_bf16, d, N, K, lbo_a, sbo_a, lbo_b, sbo_b,
                              static_cast<cudaStream_t>(stream));
}

}  // extern "C"
