/* b200inr.h -- C ABI of the B200-native INR fit/query hot path.
 *
 * The reference (MRIRC/MRI-super-resolution) has no FFI of its own: its hot path is PyTorch calls made from
 * Python.  Each entry point below names the reference expression it replaces
 * (paths relative to the reference checkout, INR/ = implicit-neural-representations/).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless the name ends in _host.  The caller owns every buffer,
 *     including workspaces; the library allocates nothing and keeps no state between calls.
 *   - `stream` is a cudaStream_t passed as void*.  All work is stream-ordered; no call synchronises the host.
 *   - Return value: 0 on success, a negative B200INR_ERR_* otherwise.  Nothing throws across the ABI.
 *   - Row order everywhere is the reference's: C-order flatten, last axis fastest (INR/SRDWI.py:12-18),
 *     tensors are [rows, features] row-major fp32.
 *   - The library targets sm_100a only.  There is no CPU fallback.
 */
#ifndef B200INR_H_
#define B200INR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200INR_OK 0
#define B200INR_ERR_BAD_SHAPE (-1)
#define B200INR_ERR_BAD_ALIGN (-2)
#define B200INR_ERR_UNSUPPORTED_ARCH (-3)
#define B200INR_ERR_CUDA (-4)
#define B200INR_ERR_NULL (-5)

#define B200INR_ACT_SINE 0 /* sin(omega * z)   (SineLayer, INR/SRDWI.py:59)                                  */
#define B200INR_ACT_RELU 1 /* max(z, 0)         (Fourier-feature ReLU MLP of BASELINE config 4; omegas ignored) */
#define B200INR_ACT_GABOR 2 /* complex Gabor wavelet exp(i w0 lin) exp(-s0^2 (|lin|^2 + |orth|^2))
                               (ComplexGaborLayer2D, INR/INRmodel.py:109-120; WIRE network INR/wiretest.ipynb cell 2):
                               hidden_features = H complex units (128), IN_COORDS only */

#define B200INR_ACT_TANH 3 /* tanh(z): the perturbation network PN (INR/INRmodel.py:153-169) as a generic-family
                              network (H = 256 operands; a 128-wide PN is zero padded); omegas ignored            */

#define B200INR_IN_COORDS 0   /* network input = the d <= 4 raw coordinates; H <= 256, multiple of 8 (zero-padded) */
#define B200INR_IN_FOURIER 1  /* network input = input_mapping(coords, B) (INR/SRDWI.py:111-116) computed in-kernel
                                 from the d raw coordinates; in_features = d, first layer K = 2 * mapping_size      */
#define B200INR_IN_FEATURES 2 /* network input = explicit fp32 feature rows [rows, in_features] (what the reference
                                 scripts feed: pre-computed Fourier features, INR/superresDWI.py:121-122)          */

/* Network description == ctor arguments of Siren (INR/SRDWI.py:68-71, INR/INRmodel.py:123-125). */
typedef struct b200inr_net {
  int32_t in_features;     /* IN_COORDS / IN_FOURIER: d = 1..4; IN_FEATURES: K0, a multiple of 64, <= H        */
  int32_t hidden_features; /* H: IN_COORDS 8..256 (multiple of 8); IN_FOURIER / IN_FEATURES 256 or 512      */
  int32_t hidden_layers;   /* L, hidden->hidden layers; there are L+1 activated layers + 1 final linear      */
  int32_t out_features;    /* C, 1..32                                                                       */
  float first_omega_0;     /* omega of the first SineLayer (INR/SRDWI.py:78)                                  */
  float hidden_omega_0;    /* omega of the hidden SineLayers (INR/SRDWI.py:81)                                */
  int32_t activation;      /* B200INR_ACT_*                                                                  */
  int32_t input_mode;      /* B200INR_IN_*                                                                   */
  int32_t mapping_size;    /* m of input_mapping (IN_FOURIER only; 2m a multiple of 64, <= H), else 0         */
  float scale_0;           /* s0 of the Gabor window (ACT_GABOR only), else 0                                */
  int32_t flags;           /* B200INR_NET_* bits, 0 by default                                               */
} b200inr_net;

/* SIREN on raw coordinates (IN_COORDS) trains through ONE layer-pipelined backward kernel whose only HBM stream is
 * the 16-bit phase stash (b200inr_siren_backward).  This bit selects the older staged path instead: forward stashes
 * sin outputs + phases, b200inr_siren_dgrad writes every dL/dtheta, b200inr_siren_wgrad contracts them (the two
 * calls b200inr_siren_backward then makes).  The stash layout differs, so forward, backward and
 * b200inr_stash_bytes must see the same flag.  Ignored by the other network families (always staged). */
#define B200INR_NET_STAGED_BWD 1
/* Generic family (IN_FOURIER / IN_FEATURES): the forward's activations are only needed for the activation-gradient
 * chain (a FROZEN network whose input gradient is wanted: the INR inside the PerturbNet phase, INR/inrDWI.py:141-147,
 * where the reference's autograd computes INR weight gradients that nobody uses).  The training forward then stashes
 * only what the chain reads (16-bit phases for sine, bf16 outputs for ReLU / tanh), the backward stores no dL/dtheta
 * tiles, and parameter gradients are unavailable (b200inr_siren_backward_coords with grad_params = NULL). */
#define B200INR_NET_DGRAD_ONLY 4
/* Generic family: the network output passes through scale_0 * tanh(.) (PN's `eps * tanh(...)`, INR/INRmodel.py:167).
 * The backward needs the forward's output to form the derivative: b200inr_siren_backward_tanh_out. */
#define B200INR_NET_TANH_OUT 8
/* The ReLU-tail SIREN of INR/INR_ERD.py:28-67 (SIREN on raw coordinates, pipelined backward only): the LAST of the
 * hidden_layers hidden layers is nn.Linear + nn.ReLU (no omega) instead of a SineLayer, and the network output passes
 * through a ReLU:  net = SineLayer(first), (hidden_layers - 1) x SineLayer, Linear + ReLU, final Linear, ReLU.
 * hidden_layers >= 1.  b200inr_siren_forward then always returns relu(out); the loss gradient handed to
 * b200inr_siren_backward must already carry the output ReLU's mask (b200inr_mse_loss_relu_out). */
#define B200INR_NET_RELU_TAIL 2

/* Dense coordinate grid == get_mgrid(shape) (INR/SRDWI.py:12-18) restricted to linear rows
 * [row_begin, row_begin + rows).  Coordinates are never materialised; kernels derive them from the index. */
typedef struct b200inr_grid {
  int32_t ndim;      /* must equal net.in_features */
  int32_t shape[4];  /* global grid shape; unused trailing entries = 1 */
  int64_t row_begin; /* first global linear index covered by this call (rank shard offset) */
} b200inr_grid;

/* Per-axis sparse operator of the separable LR degradation (2x average pooling, optionally preceded by the
 * 5-tap Gaussian sigma=0.5 blur with mirror boundary).  Built on the host by
 * b200inr_degrade_build_axis_host and uploaded by the caller. */
#define B200INR_DEGRADE_MAX_TAPS 8
typedef struct b200inr_axis_taps {
  int32_t idx[B200INR_DEGRADE_MAX_TAPS];
  float w[B200INR_DEGRADE_MAX_TAPS];
} b200inr_axis_taps;

const char* b200inr_version(void);
const char* b200inr_error_string(int code);

/* ---- parameter layout -------------------------------------------------------------------------------
 * Flat fp32 parameter vector in reference units (W, not omega*W), canonical order
 *   W0[H,d] b0[H]  W1[H,H] b1[H] ... WL[H,H] bL[H]  Wf[C,H] bf[C]
 * each segment starting at a multiple of 4 floats.  offsets[2*i], offsets[2*i+1] = start of W_i, b_i;
 * offsets has b200inr_param_offset_count entries: 2*(L+2) here, one more (the frequency matrix B) for
 * B200INR_IN_FOURIER, and for B200INR_ACT_GABOR 4*(L+1)+2: per Gabor layer W_lin b_lin W_orth b_orth (complex values as
 * (re, im) pairs), then W_f b_f.  (nn.Linear weight layout [out,in], INR/SRDWI.py:47.) */
int b200inr_param_count(const b200inr_net* net, int64_t* n_floats);
int b200inr_param_offsets(const b200inr_net* net, int64_t* offsets);

/* bf16 tensor-core staging of the weights (omega folded in, UMMA shared-memory layout, both orientations).
 * Replaces nothing in the reference; it is the operand format of the fused kernels. */
int b200inr_packed_bytes(const b200inr_net* net, size_t* bytes);
int b200inr_pack_weights(const b200inr_net* net, const float* params, void* packed, void* stream);

/* Activation stash written by a training forward and consumed by backward (bf16 sin outputs in UMMA tile
 * layout, 16-bit phases, bf16 pre-activation gradients). */
int b200inr_stash_bytes(const b200inr_net* net, int64_t rows, size_t* bytes);

/* ---- fused MLP ---------------------------------------------------------------------------------------
 * Siren.forward (INR/SRDWI.py:87-91 == nn.Sequential of SineLayer.forward :58-59 and the final nn.Linear).
 * Exactly one of coords ([rows,d] fp32) / grid must be non-NULL.
 * out: [rows,C] fp32.  clamp != 0 applies torch.clamp(min=clamp_min) (INR/superresDWI.py:161).
 * stash == NULL: inference/query; otherwise the stash is filled for b200inr_siren_backward. */
int b200inr_siren_forward(const b200inr_net* net, const void* packed, const float* coords,
                          const b200inr_grid* grid, int64_t rows, float* out, int clamp, float clamp_min,
                          void* stash, void* stream);

/* The training forward of the fused fit with the 2x2x1 pooled LR-consistency loss taken in the kernel's final
 * epilogue: replaces b200inr_siren_forward (stash != NULL) followed by b200inr_pool_mse -- the [rows, C] prediction
 * never goes to HBM.  target_lr: this slab's LR volume [X/2, Y/2, Z, C]; grad_hr [rows, C] fp32 receives dL/dpred;
 * loss_accum[0] += this slab's share of the loss; count = global number of LR elements.  Same arithmetic as
 * b200inr_pool_mse.  Supported for SIREN on raw coordinates with the pipelined backward, a 3-D grid with
 * 128 % (2 Z) == 0, Y even, (Y Z) % 128 == 0, and a slab of whole x-plane pairs (rows, row_begin multiples of 2 Y Z);
 * B200INR_ERR_BAD_SHAPE otherwise (call the two-kernel form). */
int b200inr_siren_forward_pool_loss(const b200inr_net* net, const void* packed, const b200inr_grid* grid, int64_t rows,
                                    const float* target_lr, double count, float* grad_hr, float* loss_accum,
                                    void* stash, void* stream);

/* loss.backward() through Siren (autograd of INR/SRDWI.py:58-59,87-91): given dL/dout [rows,C] fp32,
 * ACCUMULATES dL/dparams into grad_params (flat layout above, reference units).  No input gradient
 * (SRDWI.Siren detaches its coords, INR/SRDWI.py:88). */
int b200inr_siren_backward(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                           const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params,
                           void* stream);

/* The same for a network fed with explicit feature rows (B200INR_IN_FEATURES), plus the gradient of those rows:
 * grad_input [rows, in_features] fp32 (overwritten) = dL/d(features) = dTheta_0 (omega_0 W_0).  Replaces the part of
 * loss.backward() that reaches the network input when it is not detached -- INRmodel.Siren.forward
 * (INR/INRmodel.py:147-149), which the PerturbNet phase trains through (INR/inrDWI.py:141-147); the SRDWI variant
 * detaches its input (INR/SRDWI.py:88) and never needs it.  One more tcgen05 chain step inside the dgrad kernel.
 * B200INR_ERR_BAD_SHAPE for the other input modes. */
int b200inr_siren_backward_input(const b200inr_net* net, const void* packed, void* stash, int64_t rows,
                                 const float* grad_out, float* grad_params, float* grad_input, void* stream);

/* Generic-family network on in-kernel Fourier features (B200INR_IN_FOURIER): backward that also returns
 * dL/d(coordinates) [rows, d] -- the input gradient chained through the adjoint of input_mapping (INR/SRDWI.py:111-116)
 * inside the kernel, so neither the [rows, 2m] features nor their gradient touch HBM.  This is the INR's share of the
 * reference's PerturbNet phase (INR/inrDWI.py:141-147: loss.backward() through INR.forward(input_mapping(PN(...), B))).
 * grad_params = NULL with B200INR_NET_DGRAD_ONLY set (frozen network, lean stash); otherwise parameter gradients are
 * accumulated as in b200inr_siren_backward.  coords / grid: the same rows the forward was given. */
int b200inr_siren_backward_coords(const b200inr_net* net, const void* packed, void* stash, const float* coords,
                                  const b200inr_grid* grid, int64_t rows, const float* grad_out, float* grad_params,
                                  float* grad_coords, void* stream);
/* Backward of a generic-family network with B200INR_NET_TANH_OUT (PN: out = scale_0 * tanh(final linear)): `out` is the
 * forward's output [rows, C], grad_out the gradient with respect to it. */
int b200inr_siren_backward_tanh_out(const b200inr_net* net, const void* packed, void* stash, int64_t rows,
                                    const float* out, const float* grad_out, float* grad_params, void* stream);
/* PN (INR/INRmodel.py:153-169) feeds its first layer cat(features, acq) with the constant acq = sample / 10: the same as
 * a bias b + acq * w_last.  master = [generic-family parameters (n_net floats) | w_last[H]].
 * effective_params: eff[0:n_net] = master[0:n_net] with eff[bias_off + h] += acq * w_last[h] (input of
 * b200inr_pack_weights), and clear_grads[0 : n_net + H] = 0 when not NULL (the step's zero_grad);
 * fold_grad: grads[n_net + h] = acq * grads[bias_off + h] after the weight-gradient pass. */
int b200inr_pn_effective_params(const float* master, int64_t n_net, int64_t bias_off, int32_t H, float acq, float* eff,
                                float* clear_grads, void* stream);
int b200inr_pn_fold_grad(float* grads, int64_t n_net, int64_t bias_off, int32_t H, float acq, void* stream);

/* The two kernels of the STAGED b200inr_siren_backward as separate calls (same arguments; backward == dgrad then
 * wgrad).  B200INR_ERR_BAD_SHAPE for a pipelined SIREN (no B200INR_NET_STAGED_BWD): its backward is one kernel.
 * dgrad: activation-gradient chain, fills the stash with dL/dtheta of every sine layer and the bf16 dL/dout tile;
 * wgrad: contracts the stash over the rows and ACCUMULATES into grad_params. */
int b200inr_siren_dgrad(const b200inr_net* net, const void* packed, void* stash, int64_t rows, const float* grad_out,
                        void* stream);
int b200inr_siren_wgrad(const b200inr_net* net, void* stash, const float* coords, const b200inr_grid* grid,
                        int64_t rows, float* grad_params, void* stream);

/* ---- loss and LR degradation -------------------------------------------------------------------------
 * ((out - gt)**2).mean() and its gradient (INR/superresDWI.py:135; weighted form INR/INR_ERD.py:265).
 * loss_accum[0] += sum(w*(pred-target)^2)/count ; grad = 2*w*(pred-target)/count.  weight may be NULL. */
int b200inr_mse_loss(const float* pred, const float* target, const float* weight, int64_t n, double count,
                     float* grad, float* loss_accum, void* stream);

/* The same loss when `pred` is the output of a ReLU-tail network (relu(raw)): grad = 2 w (pred - target) / count where
 * pred > 0 and 0 elsewhere, i.e. dL/d(raw) (INR/INR_ERD.py:65-66 followed by :203-204 / :264-266). */
int b200inr_mse_loss_relu_out(const float* pred, const float* target, const float* weight, int64_t n, double count,
                              float* grad, float* loss_accum, void* stream);

/* Soft-ERD loss weights (INR/INR_ERD.py:222-235) and the soft-ERD image (:126-160), one thread per voxel:
 * x = signal[v, 0:n] (the acquisitions of one b-value), b0[v]; if mean(x) > 2 noise_level:
 *   T = max(mul * exp(-slope * mean(x) / b0), 2),  weights[v, :] = exp(x / T),  soft_mean[v] = sum(softmax(x / T) * x)
 *   (one-hot on overflow, like the reference's RuntimeWarning branch);
 * else weights[v, :] = 1 / n and soft_mean[v] = mean(x).  Double arithmetic inside (the reference is NumPy float64),
 * fp32 in and out.  weights / soft_mean may be NULL.  2 <= n <= 64. */
int b200inr_soft_erd(const float* signal, const float* b0, int64_t voxels, int32_t n, double noise_level, double mul,
                     double slope, float* weights, float* soft_mean, void* stream);

/* Host helper: per-axis taps of D (LR row i <- HR columns) and of its transpose (HR column x <- LR rows).
 * n_hr must be even; blur = 0 -> 2-tap box mean; blur = 1 -> Gaussian sigma 0.5 (5 taps, mirror) then box. */
int b200inr_degrade_build_axis_host(int32_t n_hr, int blur, b200inr_axis_taps* fwd_host /*[n_hr/2]*/,
                                    b200inr_axis_taps* adj_host /*[n_hr]*/);

/* D: hr [X,Y,ZC] -> lr [X/2,Y/2,ZC] (in-plane only, mirrors the reference's [::2, ::2] decimation axes,
 * INR/superresDWI.py:94).  tx/ty: device copies of the fwd taps of the x/y axes. */
int b200inr_degrade_forward(const float* hr, float* lr, int32_t X, int32_t Y, int64_t ZC,
                            const b200inr_axis_taps* tx, const b200inr_axis_taps* ty, void* stream);
/* D^T: lr [X/2,Y/2,ZC] -> hr [X,Y,ZC] (adjoint taps). */
int b200inr_degrade_adjoint(const float* lr, float* hr, int32_t X, int32_t Y, int64_t ZC,
                            const b200inr_axis_taps* ax, const b200inr_axis_taps* ay, void* stream);
/* The blurred degradation as a banded operator (host helper): fwd6_host [n_hr/2][6] = D[i][2i-2 .. 2i+3] (mirror-merged
 * weights, zero where a tap leaves the volume), adj3_host [n_hr][3] = D[((x-2)>>1) + t][x], t = 0..2. */
int b200inr_degrade_build_band_host(int32_t n_hr, int blur, float* fwd6_host, float* adj3_host);
/* Fused blur + pool consistency loss (dwi_inr.ipynb#c6:L8's rescale(.5, anti_aliasing=True) as the degradation; SURVEY.md
 * section 8c) in two streaming passes: resid_lr [X/2, Y/2, ZC] = D pred - target (workspace, also an output),
 * loss_accum[0] += sum(resid^2)/count, grad_hr = D^T (2 resid / count) (skipped when NULL).  bx6/by6/ax3/ay3: DEVICE copies
 * of the band tables of the x and y axes.  Replaces b200inr_degrade_forward + b200inr_mse_loss + b200inr_degrade_adjoint. */
int b200inr_blurpool_mse(const float* pred_hr, const float* target_lr, int32_t X, int32_t Y, int64_t ZC, double count,
                         const float* bx6, const float* by6, const float* ax3, const float* ay3, float* resid_lr,
                         float* grad_hr, float* loss_accum, void* stream);
/* The same loss for ONE RANK'S SLAB of the volume (multi-GPU fit): the rank owns the HR planes
 * [x_begin, x_end) (even bounds) of the X-plane volume.  pred_ext holds the planes [max(x_begin-4, 0), min(x_end+4, X))
 * (own planes plus the neighbours' four halo planes), target_ext and resid_ext the LR rows
 * [max(x_begin/2-1, 0), min(x_end/2+1, X/2)) (one halo row on either side: what the adjoint of the own planes reads);
 * loss_accum[0] += sum over the OWN LR rows of resid^2 / count, grad_hr [x_end-x_begin, Y, ZC] = own planes of
 * D^T (2 resid / count).  count is the global LR element count.  x_begin = 0, x_end = X is b200inr_blurpool_mse. */
int b200inr_blurpool_mse_slab(const float* pred_ext, const float* target_ext, int32_t X, int32_t Y, int64_t ZC,
                              double count, const float* bx6, const float* by6, const float* ax3, const float* ay3,
                              int32_t x_begin, int32_t x_end, float* resid_ext, float* grad_hr, float* loss_accum,
                              void* stream);
/* Fused 2x2x1 average-pool consistency loss: loss_accum[0] += sum((pool(pred)-target_lr)^2)/count and
 * grad_hr = pool^T(2*(pool(pred)-target_lr)/count), one pass (the fast path of BASELINE config 2). */
int b200inr_pool_mse(const float* pred_hr, const float* target_lr, int32_t X, int32_t Y, int64_t ZC,
                     double count, float* grad_hr, float* loss_accum, void* stream);

/* ---- optimiser ----------------------------------------------------------------------------------------
 * torch.optim.Adam.step with defaults amsgrad=False, weight_decay=0 (INR/superresDWI.py:115-116,138).
 * state: 4 floats on device {step, bias_correction1, bias_correction2, unused}; zero-initialised by the
 * caller, advanced by the kernel (so the call can be captured in a CUDA graph). */
int b200inr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, float* state, void* stream);

/* opt.step() + opt.zero_grad() of the reference loop (INR/superresDWI.py:136-138) plus the bf16 re-staging of the
 * operands, for the fused fit: Adam (same arithmetic as b200inr_adam_step) over the n = b200inr_param_count floats of
 * `net`, every consumed gradient cleared (grads is all-zero on return, so the next step needs no memset), the step
 * counter advanced on the device, and `packed` rewritten from the updated parameters.  grads has n + 4 floats: the
 * loss accumulator of the step rides at grads[n] (see b200inr_mse_loss / b200inr_pool_mse); it is moved to
 * loss_out[0] (may be NULL) and cleared.  state: 4 floats, zero-initialised by the caller ({step, ticket, -, -}).
 * ONE launch for SIREN on raw coordinates (every parameter is updated by the thread that packs it), two for the
 * other families.  Graph-capturable. */
int b200inr_optimizer_step(const b200inr_net* net, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                           float lr, float beta1, float beta2, float eps, float* state, void* packed,
                           float* loss_out, void* stream);

/* The same step for data-parallel ranks (one process per GPU) with the gradient exchange INSIDE
 * the kernel: instead of ncclAllReduce + b200inr_optimizer_step, every rank's [grad | loss] buffer of this step lives in
 * peer-mapped (symmetric) memory, the kernel waits until every rank has entered it (flag words written over NVLink),
 * sums the `world` buffers in rank order -- identical arithmetic on every rank, so the replicated weights stay
 * bit-identical -- applies Adam, re-stages the operands and clears grads_next, the buffer the NEXT step accumulates
 * into (the two buffers alternate, so no rank clears what a peer may still be reading and no end barrier is needed).
 * peer_grads / peer_flags: DEVICE arrays of `world` pointers (peer_flags[p] = rank p's flag words, >= world uint32,
 * zero-initialised once).  loss_out[0] = sum over ranks of the loss accumulators (at index n of each buffer). */
int b200inr_optimizer_step_peers(const b200inr_net* net, float* params, float* grads_next,
                                 const float* const* peer_grads, uint32_t* const* peer_flags, int32_t world,
                                 int32_t rank, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                                 float eps, float* state, void* packed, float* loss_out, void* stream);

/* sizeof(b200inr_net) as this library was compiled: a binding whose struct definition has drifted (fewer fields)
 * must refuse to call rather than pass a short struct. */
size_t b200inr_net_size(void);
/* Number of entries b200inr_param_offsets writes for `net` (2(L+2) for SIREN on raw coordinates / explicit features,
 * one more for B200INR_IN_FOURIER -- the frequency matrix B --, 4(L+1)+2 for B200INR_ACT_GABOR). */
int b200inr_param_offset_count(const b200inr_net* net, int32_t* count);

/* ---- coordinate helpers (API parity; the fused kernels do not need them) ------------------------------ */
/* get_mgrid (INR/SRDWI.py:12-18): coords [rows, ndim] fp32. */
int b200inr_get_mgrid(const b200inr_grid* grid, int64_t rows, float* coords, void* stream);
/* input_mapping (INR/SRDWI.py:111-116): out [rows, 2m] = cat(sin(2*pi*x@B.T), cos(2*pi*x@B.T)). */
int b200inr_input_mapping(const float* x, const float* B, int64_t rows, int32_t d, int32_t m, float* out,
                          void* stream);
/* Its adjoint with respect to x (the PerturbNet phase differentiates through the feature map, INR/inrDWI.py:142-147):
 * grad_x [rows, d] = 2 pi (grad_out[:, :m] .* cos p - grad_out[:, m:] .* sin p) B,  p = 2 pi x B^T.  d <= 8. */
int b200inr_input_mapping_backward(const float* x, const float* B, const float* grad_out, int64_t rows, int32_t d,
                                   int32_t m, float* grad_x, void* stream);

/* The pre-activation returned by SineLayer.forward_with_intermediate (INR/SRDWI.py:61-64), a probing helper ("for
 * visualization of activation distributions"): out [rows, H] = omega * (x [rows, d] W[H, d]^T + b[H]), fp32 on CUDA
 * cores, any input width d <= 4096. */
int b200inr_sine_layer_pre(const float* x, const float* W, const float* b, int64_t rows, int32_t d, int32_t H,
                           float omega, float* out, void* stream);

/* calculate_ADC (INR/SRDWI.py:118-130), the step right after the query: per voxel the least-squares line through
 * (b_k / 1000, log(signal_k + 1e-7)); adc[v] = -slope clamped to [-10, 3].  signal [voxels, nb] fp32 on the device
 * (e.g. the [rows, C] output of b200inr_siren_forward, nb = C), bvalues_host [nb] on the HOST (2 <= nb <= 64, not all
 * equal), adc [voxels] fp32 on the device. */
int b200inr_adc_fit(const float* signal, const float* bvalues_host, int64_t voxels, int32_t nb, float* adc,
                    void* stream);

/* calculate_combinations (INR/SRDWI.py:143-152) for every voxel at once, the step right before the fit
 * (INR/superresDWI.py:57-76): b0 [voxels], b1 [voxels, n1], b2 [voxels, n2], b3 [voxels, n3] (one b-value each, one
 * echo time) -> out [voxels, 4, n1*n2*n3]: column c = (i1, i2, i3) in C order holds (b0, b1[i1], b2[i2], b3[i3]),
 * i.e. np.asarray(list(itertools.product(...))).T per voxel.  All device pointers, fp32. */
int b200inr_combinations(const float* b0, const float* b1, const float* b2, const float* b3, int64_t voxels, int32_t n1,
                         int32_t n2, int32_t n3, float* out, void* stream);

/* ---- self test of the tensor-core plumbing ------------------------------------------------------------
 * One CTA computes D[128,N] = A * B^T with tcgen05.mma from swizzled shared memory.
 * mode 0: K-major operands, a[128,K], b[N,K] row-major bf16.  mode 1: MN-major operands, a[K,128], b[K,N].
 * lbo/sbo < 0 select the library's own descriptor strides (the values used by the production kernels). */
int b200inr_selftest_umma(int mode, const void* a_bf16, const void* b_bf16, float* d, int32_t N, int32_t K,
                          int32_t lbo_a, int32_t sbo_a, int32_t lbo_b, int32_t sbo_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200INR_H_ */
