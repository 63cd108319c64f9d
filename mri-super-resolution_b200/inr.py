"""Host-side mirror of the reference's INR model module on top of the B200 C ABI (include/b200inr.h).

Same names, constructor arguments, state-dict keys, RNG consumption and call protocol as the reference's
INR/SRDWI.py (get_mgrid :12-18, ImageFitting_set :20-39, SineLayer :41-64, Siren :67-91, input_mapping :111-116) and
INR/INRmodel.py (Siren :122-151), so the reference's scripts and notebooks run unchanged against it.  Every tensor
operation on the hot path is a hand-written sm_100a kernel reached through ctypes; torch supplies device memory,
streams, autograd bookkeeping and (multi-GPU) the NCCL process group.  There is no CPU path: modules must live on a
CUDA device before forward / fit / query are called, and a missing libb200inr.so raises at the first call.

Besides the nn.Module protocol (`forward` + autograd, as used at INR/superresDWI.py:134-138) two fused entry points
bypass autograd entirely:

    Siren.fit(...)    the in-lined training loop of INR/superresDWI.py:132-138 (full batch, fixed order, Adam), with
                      the loss taken either point-wise or through the LR degradation operator
    Siren.query(...)  torch.clamp(INR.forward(get_mgrid(shape)), min=0) of INR/superresDWI.py:161 without ever
                      materialising the coordinate grid
"""
import ctypes
import os
import math

import numpy as np
import torch
from torch import nn
from torch.utils.data import Dataset

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"b200inr: {what} must be a CUDA tensor (the B200 kernels are the only implementation)")


def _aligned_bytes(nbytes, device, align=1024, zero=True):
    """Byte buffer whose data pointer is `align`-aligned.  zero=False skips the fill: the activation stash is written
    tile by tile before anything reads it (a 13 GB memset per forward otherwise costs 2 ms at cfg4-like sizes).
    B200INR_POISON_STASH=1 fills such buffers with 0xFF (bf16 NaN) instead -- the tests' proof of that claim."""
    if zero:
        raw = torch.zeros(int(nbytes) + align, dtype=torch.uint8, device=device)
    else:
        raw = torch.empty(int(nbytes) + align, dtype=torch.uint8, device=device)
        if os.environ.get("B200INR_POISON_STASH", "0") == "1":
            raw.fill_(255)
    off = (-raw.data_ptr()) % align
    return raw[off:off + int(nbytes)]


# ------------------------------------------------------------------------------------------------ coordinates
def get_mgrid(shape, device=None):
    """Reference INR/SRDWI.py:12-18: [prod(shape), len(shape)] coordinates in [-1, 1], C order, last axis fastest.

    With device=None the result is a CPU tensor exactly like the reference's (built with the same torch calls, it is
    data preparation, not the hot path); with a CUDA device the grid is written by the b200inr_get_mgrid kernel.
    """
    shape = tuple(int(s) for s in shape)
    if device is None or torch.device(device).type == "cpu":
        axes = tuple(torch.linspace(-1, 1, steps=n) for n in shape)
        return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, len(shape))
    if not 1 <= len(shape) <= 4:
        raise RuntimeError("b200inr: get_mgrid supports 1 to 4 dimensions on the device")
    rows = int(np.prod(shape))
    out = torch.empty((rows, len(shape)), dtype=torch.float32, device=device)
    grid = _lib.make_grid(shape)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().b200inr_get_mgrid(ctypes.byref(grid), rows, _ptr(out), _stream()), "get_mgrid")
    return out


class ImageFitting_set(Dataset):
    """Reference INR/SRDWI.py:20-39: flattens N-D volumes to pixels [n, N, 1] and coords [n, N, d]; __getitem__
    ignores its index and returns everything."""

    def __init__(self, img_dataset):
        super().__init__()
        shape = img_dataset[0].shape
        self.shape = shape
        n = int(np.prod(shape))
        self.pixels = torch.empty((len(img_dataset), n, 1))
        self.coords = torch.empty((len(img_dataset), n, len(shape)))
        grid = get_mgrid(shape)
        for i, img in enumerate(img_dataset):
            self.pixels[i] = torch.from_numpy(np.ascontiguousarray(img)).float().reshape(-1, 1)
            self.coords[i] = grid

    def __len__(self):
        return len(self.pixels)

    def __getitem__(self, idx):
        return self.coords, self.pixels


class _InputMappingFunction(torch.autograd.Function):
    """input_mapping as one kernel; its adjoint with respect to x as another (B carries no gradient: it is a fixed
    random matrix in the reference, INR/superresDWI.py:105-106)."""

    @staticmethod
    def forward(ctx, x, B):
        rows, d = x.shape
        m = B.shape[0]
        out = torch.empty((rows, 2 * m), dtype=torch.float32, device=x.device)
        if rows:
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().b200inr_input_mapping(_ptr(x), _ptr(B), rows, d, m, _ptr(out), _stream()),
                           "input_mapping")
        ctx.save_for_backward(x, B)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, B = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        rows, d = x.shape
        grad_x = torch.empty_like(x)
        if rows:
            grad_out = grad_out.contiguous().float()
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().b200inr_input_mapping_backward(_ptr(x), _ptr(B), _ptr(grad_out), rows, d,
                                                                      B.shape[0], _ptr(grad_x), _stream()),
                           "input_mapping_backward")
        return grad_x, None


def input_mapping(x, B):
    """Reference INR/SRDWI.py:111-116: cat([sin(2 pi x B^T), cos(2 pi x B^T)], -1); identity when B is None.
    Differentiable with respect to x (the PerturbNet phase trains through it, INR/inrDWI.py:142-147)."""
    if B is None:
        return x
    _require_cuda(x, "input_mapping input")
    _require_cuda(B, "input_mapping B")
    B = B.detach().contiguous().float()
    if x.dim() != 2 or B.dim() != 2 or B.shape[1] != x.shape[1]:
        raise RuntimeError("b200inr: input_mapping expects x [N, d] and B [m, d]")
    if x.requires_grad and torch.is_grad_enabled():
        if x.shape[1] > 8:
            raise RuntimeError("b200inr: input_mapping is differentiable for d <= 8 coordinates")
        return _InputMappingFunction.apply(x.contiguous().float(), B)
    with torch.no_grad():
        return _InputMappingFunction.apply(x.detach().contiguous().float(), B)


def all_combinations(hybrid_raw_norm, te=0, device=None):
    """Every voxel's calculate_combinations table in one kernel: what INR/superresDWI.py:57-76 assembles with a
    32-process pool and a Python loop.  hybrid_raw_norm[b][te]: b = 0 a volume [X, Y, Z], b = 1..3 volumes
    [X, Y, Z, n_b] (NumPy arrays or tensors).  Returns a CUDA fp32 tensor [X, Y, Z, 4, n1*n2*n3] (`acquisitions`)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    vols = [torch.as_tensor(np.asarray(hybrid_raw_norm[b][te]) if not torch.is_tensor(hybrid_raw_norm[b][te])
                            else hybrid_raw_norm[b][te]).to(dev, torch.float32).contiguous() for b in range(4)]
    shape = tuple(vols[0].shape)
    if any(tuple(v.shape[:-1]) != shape for v in vols[1:]):
        raise RuntimeError("b200inr: all_combinations expects b0 [X, Y, Z] and b1..b3 [X, Y, Z, n]")
    n1, n2, n3 = (int(v.shape[-1]) for v in vols[1:])
    voxels = int(np.prod(shape))
    out = torch.empty(shape + (4, n1 * n2 * n3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200inr_combinations(_ptr(vols[0]), _ptr(vols[1]), _ptr(vols[2]), _ptr(vols[3]), voxels,
                                                    n1, n2, n3, _ptr(out), _stream()), "combinations")
    return out


def calculate_combinations(voxel, hybrid_raw_norm):
    """Reference INR/SRDWI.py:143-152, same signature and result (one voxel, echo time 0): a [4, n1*n2*n3] float64
    array.  One voxel is a 4 x N gather -- done on the host; the whole-volume form is all_combinations()."""
    i, j, k = voxel
    b0 = np.asarray(hybrid_raw_norm[0][0])[i, j, k]
    b1, b2, b3 = (np.asarray(hybrid_raw_norm[b][0])[i, j, k, :] for b in (1, 2, 3))
    g1, g2, g3 = np.meshgrid(b1, b2, b3, indexing="ij")
    return np.stack([np.full(g1.size, b0), g1.ravel(), g2.ravel(), g3.ravel()]).astype(np.float64)


def resize_array(arr, new_size=128, kind='cubic'):
    """Reference INR/SRDWI.py:132-141 (= INR/INRmodel.py:192-201): resample the third axis of a [X, Y, Z] array to
    `new_size` samples with SciPy's 1-D interpolator (cubic spline by default), both axes spanning [0, 1] inclusive.
    Host-side data preparation exactly like the reference (same SciPy call, evaluated for all new positions at once
    instead of one plane per Python iteration); float64 [X, Y, new_size]."""
    from scipy.interpolate import interp1d
    arr = np.asarray(arr)
    if arr.ndim != 3:
        raise ValueError("resize_array expects a 3-D array [X, Y, Z]")
    x_old = np.linspace(0, 1, arr.shape[2])
    x_new = np.linspace(0, 1, int(new_size))
    return np.asarray(interp1d(x_old, arr, kind=kind, axis=2)(x_new), dtype=np.float64)


def calculate_ADC(bvalues, slicedata):
    """Reference INR/SRDWI.py:118-130 (= INR/INRmodel.py): per-voxel mono-exponential fit, ADC = -slope of the
    least-squares line through (b / 1000, log(signal + 1e-7)), clamped to [-10, 3] -- one kernel instead of a Python
    double loop over np.polyfit.

    slicedata [..., nb]: a CUDA tensor (e.g. query(...).view(X, Y, Z, C): stays on the device, returns a CUDA fp32
    tensor [...]) or a NumPy array as in the reference (uploaded to the current CUDA device; returns a float64 array
    of shape slicedata.shape[:-1], the reference's return type).  bvalues: nb values (array / list / tensor).
    """
    b = np.asarray(bvalues.detach().cpu() if torch.is_tensor(bvalues) else bvalues, dtype=np.float32).reshape(-1)
    as_numpy = not torch.is_tensor(slicedata)
    if as_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("b200inr: calculate_ADC needs a CUDA device; there is no CPU path")
        slicedata = torch.from_numpy(np.ascontiguousarray(slicedata, dtype=np.float32)).cuda()
    _require_cuda(slicedata, "calculate_ADC input")
    if slicedata.shape[-1] != b.size:
        raise RuntimeError("b200inr: calculate_ADC expects one signal per b-value along the last axis")
    sig = slicedata.detach().contiguous().float()
    voxels = sig.numel() // b.size if b.size else 0
    adc = torch.empty(sig.shape[:-1], dtype=torch.float32, device=sig.device)
    bh = np.ascontiguousarray(b)
    with torch.cuda.device(sig.device):
        _lib.check(_lib.load().b200inr_adc_fit(_ptr(sig), bh.ctypes.data, voxels, int(b.size), _ptr(adc), _stream()),
                   "adc_fit")
    return adc.cpu().numpy().astype(np.float64) if as_numpy else adc


# ------------------------------------------------------------------------------------------------ layers
class PN(nn.Module):
    """Perturbation network of the reference (INR/INRmodel.py:153-169): a (features + acquisition index) -> hidden ->
    `dimension` tanh MLP whose output, scaled by eps, replaces the coordinates fed to input_mapping in the PerturbNet
    phase (INR/inrDWI.py:141-142).  Same constructor, parameter names (state-dict keys
    perturb_linear.*, perturb_linear2.*) and RNG order.  Two tiny dense layers: plain PyTorch; the expensive part of
    the phase -- the gradient through input_mapping and the INR -- runs in the fused kernels."""

    def __init__(self, in_features, hidden_features, dimension):
        super().__init__()
        self.tanh = nn.Tanh()
        self.perturb_linear = nn.Linear(in_features + 1, hidden_features)
        self.perturb_linear2 = nn.Linear(hidden_features, dimension)

    def forward(self, coords, sample=0, eps=0):
        x = coords.detach()  # the reference detaches its input (INR/INRmodel.py:161)
        # Linear(cat(x, acq)) with a constant acq column = x W[:, :-1]^T + (b + acq W[:, -1]): the [N, in + 1] copy of
        # the features (1 GB per step at the script's sizes) is never made
        w = self.perturb_linear.weight
        hidden = self.tanh(nn.functional.linear(x, w[:, :-1], self.perturb_linear.bias + (sample / 10.) * w[:, -1]))
        return eps * self.tanh(self.perturb_linear2(hidden))


class SineLayer(nn.Module):
    """Reference INR/SRDWI.py:41-64.  Parameter container with the reference's initialisation; the arithmetic runs
    fused inside Siren (one kernel for the whole network), so a stand-alone forward is a 0-hidden-layer fused call."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega_0=30):
        super().__init__()
        self.omega_0 = omega_0
        self.is_first = is_first
        self.in_features = in_features
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.init_weights()

    def init_weights(self):
        with torch.no_grad():
            if self.is_first:
                bound = 1 / self.in_features
            else:
                bound = np.sqrt(6 / self.in_features) / self.omega_0
            self.linear.weight.uniform_(-bound, bound)

    def _fused(self, input):
        """sin(omega_0 (x W^T + b)) through the fused kernels as a 0-hidden-layer network whose (unused) final linear
        is zero: the layer's activations are read back from the training stash (bf16 sin outputs, UMMA tile layout).
        Inference only -- training goes through the enclosing Siren, whose backward is fused as well."""
        _require_cuda(input, "SineLayer input")
        lin = self.linear
        k, h = lin.in_features, lin.out_features
        if input.dim() != 2 or input.shape[1] != k:
            raise RuntimeError(f"b200inr: expected an input of shape [N, {k}]")
        dev = input.device
        if k <= 4:    # raw coordinates: the 256-wide SIREN kernels (any width 8..256), staged stash keeps sin outputs
            desc = _lib.make_net(k, h, 0, 1, self.omega_0, self.omega_0, flags=_lib.NET_STAGED_BWD)
            width, y_off = 256, 0
        else:         # explicit feature rows: the generic family (K a multiple of 64; other widths run zero padded)
            width = _generic_width(h, k)
            desc = _lib.make_net(k, width, 0, 1, self.omega_0, self.omega_0, input_mode=_lib.IN_FEATURES)
        rows = input.shape[0]
        if rows == 0:
            return torch.empty((0, h), dtype=torch.float32, device=dev)
        tiles = (rows + 127) // 128
        if k > 4:
            y_off = tiles * 128 * k * 2  # the network input (A operand of layer 0) is stashed first
        with torch.no_grad(), torch.cuda.device(dev):
            lib = _lib.load()
            off = _lib.param_offsets(desc)  # raises for unsupported widths
            flat = torch.zeros(_lib.param_count(desc), dtype=torch.float32, device=dev)
            hw = width if k > 4 else h  # rows of the (padded) weight segment
            flat[off[0]:off[0] + hw * k].view(hw, k)[:h].copy_(lin.weight.detach())
            if lin.bias is not None:
                flat[off[1]:off[1] + h].copy_(lin.bias.detach())
            packed = _aligned_bytes(_lib.packed_bytes(desc), dev)
            _lib.check(lib.b200inr_pack_weights(ctypes.byref(desc), _ptr(flat), _ptr(packed), _stream()), "pack_weights")
            x = input.detach().contiguous().float()
            out = torch.empty((rows, 1), dtype=torch.float32, device=dev)
            stash = _aligned_bytes(_lib.stash_bytes(desc, rows), dev, zero=False)
            _lib.check(lib.b200inr_siren_forward(ctypes.byref(desc), _ptr(packed), _ptr(x), None, rows, _ptr(out), 0,
                                                 0.0, _ptr(stash), _stream()), "siren_forward")
            y = stash[y_off:y_off + tiles * 128 * width * 2].view(torch.bfloat16).view(tiles, width // 64, 128, 8, 8)
            r = torch.arange(128, device=dev)
            phys = torch.arange(8, device=dev)[None, :] ^ (r[:, None] & 7)  # logical 16-byte chunk -> SWIZZLE_128B slot
            y = y[:, :, r[:, None], phys]                                    # [tiles, kb, 128, 8, 8]
            return y.permute(0, 2, 1, 3, 4).reshape(tiles * 128, width)[:rows, :h].float()

    def forward(self, input):
        """Reference INR/SRDWI.py:58-59 for a stand-alone layer call (activation probing): one fused kernel, bf16
        activations (rel-err <= 2e-2 against the fp32 expression).  No autograd: train through Siren."""
        return self._fused(input)

    def forward_with_intermediate(self, input):
        """Reference INR/SRDWI.py:61-64: (sin(i), i) with i = omega_0 * linear(input), "for visualization of activation
        distributions".  The pre-activation does not survive the fused kernel (it stashes 16-bit phases, i mod 2 pi),
        so it is evaluated in fp32 by b200inr_sine_layer_pre (a probing helper on CUDA cores, any input width)."""
        lin = self.linear
        _require_cuda(input, "SineLayer input")
        x = input.detach().contiguous().float()
        rows, h = x.shape[0], lin.out_features
        pre = torch.empty((rows, h), dtype=torch.float32, device=x.device)
        bias = lin.bias.detach().contiguous() if lin.bias is not None else torch.zeros(h, device=x.device)
        if rows:
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().b200inr_sine_layer_pre(_ptr(x), _ptr(lin.weight.detach().contiguous()), _ptr(bias),
                                                              rows, lin.in_features, h, float(self.omega_0), _ptr(pre),
                                                              _stream()), "sine_layer_pre")
        return self._fused(input), pre


def _generic_width(hidden_features, k0):
    """Operand width the generic-family kernels run a (hidden_features, input width k0) network at: 256 or 512."""
    need = max(int(hidden_features), int(k0))
    if need <= 256:
        return 256
    if need <= 512:
        return 512
    raise RuntimeError("b200inr: hidden / input widths above 512 are not supported")


def _mlp_padded_shape(hp, hidden_layers, idx, p):
    """Flat-segment shape of canonical parameter idx (W0 b0 ... WL bL Wf bf [B]) of an MLP whose hidden width is zero
    padded to hp (None: not padded)."""
    if hp is None or idx >= 2 * (hidden_layers + 2):
        return None
    layer, is_w = idx // 2, idx % 2 == 0
    if layer == 0:
        return (hp, p.shape[1]) if is_w else (hp,)
    if layer <= hidden_layers:
        return (hp, hp) if is_w else (hp,)
    return (p.shape[0], hp) if is_w else None


class _SirenFunction(torch.autograd.Function):
    """Siren.forward as one fused kernel, loss.backward() as the fused dgrad + wgrad kernels."""

    @staticmethod
    def forward(ctx, coords, module, *params):
        needs_grad = any(ctx.needs_input_grad[2:]) or ctx.needs_input_grad[0]
        out, stash = module._forward_rows(coords=coords, grid=None, rows=coords.shape[0], train=needs_grad)
        ctx.module = module
        ctx.stash = stash
        ctx.coords = coords
        ctx.key = module._engine["key"]  # identity + version of every parameter the stash was computed with
        ctx.relu_out = out if (module._desc.flags & _lib.NET_RELU_TAIL) else None  # output ReLU: mask of the backward
        return out

    @staticmethod
    def backward(ctx, grad_out):
        module = ctx.module
        if ctx.stash is None:
            raise RuntimeError("b200inr: backward called on a forward that did not record activations")
        if module._engine is None or module._engine["key"] != ctx.key:
            # forward A, optimizer.step(), forward B, A.backward(): the bf16 operands no longer match A's stash
            raise RuntimeError("b200inr: one of the parameters needed for gradient computation has been modified by an "
                               "inplace operation since this forward (the operand buffer was re-staged)")
        if ctx.relu_out is not None:  # d relu(raw) / d raw (INR/INR_ERD.py:65-66)
            grad_out = grad_out * (ctx.relu_out > 0)
        grad_in = torch.empty_like(ctx.coords) if ctx.needs_input_grad[0] else None
        flat_grad = module._backward_rows(ctx.stash, ctx.coords, None, ctx.coords.shape[0], grad_out, grad_in=grad_in)
        ctx.stash = None
        grads = []
        for i, (o, p) in enumerate(zip(module._offsets_canonical(), module._canonical())):
            v = module._param_view(flat_grad, o, p, i)
            # torch's convention for complex parameters: grad = dL/dRe + 1j dL/dIm
            grads.append(torch.view_as_complex(v) if p.is_complex() else v)
        return (grad_in, None, *grads)


class _FusedMLP(nn.Module):
    """Engine plumbing shared by the reference-facing modules: flat fp32 parameter vector + bf16 operand buffer on the
    device, the fused forward / backward calls, and the fused query / fit entry points."""

    def _init_engine(self, desc, grid_dim):
        self._desc = desc
        self._grid_dim = grid_dim  # rank of the coordinate grid accepted by query()/fit(); None: explicit features only
        self._engine = None        # device-side staging (flat fp32 params, packed bf16 operands)
        self._optim = None         # Adam state of fit()

    def _frozen(self):
        """Non-trainable tensors stored in the flat vector after the parameters (the Fourier matrix B)."""
        return []

    def invalidate(self):
        """Drop the device-side staging (flat fp32 copy, bf16 operands): the next call rebuilds it from the
        nn.Parameters.  Needed after writes that bypass autograd's version counter (`p.data.copy_(...)`)."""
        self._engine = None

    def __getstate__(self):
        # copy.deepcopy / torch.save(module): the staging buffers are derived state (and the operand buffer is an
        # aligned VIEW whose alignment a copy does not preserve); the copy rebuilds them on first use
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_optim"] = None
        return state

    # ---------------------------------------------------------------- parameter plumbing
    def _offsets_canonical(self):
        return self._engine_state()["offsets"]

    def _padded_shape(self, idx, p):
        """Shape of the flat-vector segment that holds canonical parameter `idx` when the engine runs a zero-padded
        (wider) network than the module describes; None: the segment has the parameter's own shape."""
        return None

    def _param_view(self, flat, o, p, idx):
        """The (strided) view of `flat` where parameter p lives: real parameters in their own shape, complex ones as
        interleaved (re, im) pairs [..., 2]; a padded engine keeps the parameter in the leading corner of its segment."""
        shp = tuple(p.shape) + ((2,) if p.is_complex() else ())
        pad = self._padded_shape(idx, p)
        if pad is None:
            n = int(np.prod(shp)) if shp else 1
            return flat[o:o + n].view(shp)
        n = int(np.prod(pad))
        return flat[o:o + n].view(pad)[tuple(slice(0, d) for d in shp)]

    def _engine_state(self):
        dev = self._canonical()[0].device
        if dev.type != "cuda":
            raise RuntimeError("b200inr: the module must be on a CUDA device (call .cuda()); there is no CPU path")
        eng = self._engine
        if eng is None or eng["device"] != dev:
            _lib.load()
            n = _lib.param_count(self._desc)
            eng = {
                "device": dev,
                "offsets": _lib.param_offsets(self._desc),
                "flat": torch.zeros(n, dtype=torch.float32, device=dev),
                "packed": _aligned_bytes(_lib.packed_bytes(self._desc), dev),
                "key": None,
            }
            self._engine = eng
        return eng

    def _sync_params(self):
        """Refresh the flat fp32 copy and the bf16 operand buffer when any parameter changed since the last call."""
        eng = self._engine_state()
        ps = self._canonical() + self._frozen()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key != eng["key"]:
            flat = eng["flat"]
            with torch.no_grad():
                for i, (o, p) in enumerate(zip(eng["offsets"], ps)):
                    src = torch.view_as_real(p) if p.is_complex() else p  # complex64 -> interleaved (re, im)
                    self._param_view(flat, o, p, i).copy_(src)
            self._pack(eng)
            eng["key"] = key
        return eng

    def _pack(self, eng):
        with torch.cuda.device(eng["device"]):
            _lib.check(_lib.load().b200inr_pack_weights(ctypes.byref(self._desc), _ptr(eng["flat"]),
                                                        _ptr(eng["packed"]), _stream()), "pack_weights")

    def _writeback_params(self, eng):
        """Copy the flat fp32 master (updated by fit) back into the nn.Parameters."""
        ps = self._canonical()
        with torch.no_grad():
            for i, (o, p) in enumerate(zip(eng["offsets"], ps)):
                v = self._param_view(eng["flat"], o, p, i)
                p.copy_(torch.view_as_complex(v.contiguous()) if p.is_complex() else v)
        eng["key"] = tuple((p.data_ptr(), p._version) for p in ps + self._frozen())

    # ---------------------------------------------------------------- kernel calls
    def _forward_rows(self, coords, grid, rows, train, clamp=None, out=None, eng=None):
        eng = eng or self._sync_params()
        dev = eng["device"]
        if out is None:
            out = torch.empty((rows, self.out_features), dtype=torch.float32, device=dev)
        if rows == 0:  # empty input: nothing to launch (an empty tensor has no device pointer)
            return out, (torch.empty(0, dtype=torch.uint8, device=dev) if train else None)
        stash = _aligned_bytes(_lib.stash_bytes(self._desc, rows), dev, zero=False) if train else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b200inr_siren_forward(
                ctypes.byref(self._desc), _ptr(eng["packed"]), _ptr(coords), ctypes.byref(grid) if grid else None,
                rows, _ptr(out), int(clamp is not None), float(clamp or 0.0), _ptr(stash), _stream()), "siren_forward")
        return out, stash

    def _backward_rows(self, stash, coords, grid, rows, grad_out, flat_grad=None, eng=None, grad_in=None):
        """grad_in: None, or a [rows, in_features] fp32 tensor that receives dL/d(input rows) (explicit-feature
        networks only: b200inr_siren_backward_input)."""
        eng = eng or self._engine_state()
        dev = eng["device"]
        grad_out = grad_out.contiguous().float()
        if grad_out.data_ptr() % 16:  # the pipelined backward bulk-copies dL/dout tiles (16-byte aligned source)
            grad_out = grad_out.clone()
        if flat_grad is None:
            flat_grad = torch.zeros_like(eng["flat"])
        if rows == 0:
            return flat_grad
        if grad_in is not None:
            with torch.cuda.device(dev):
                _lib.check(_lib.load().b200inr_siren_backward_input(
                    ctypes.byref(self._desc), _ptr(eng["packed"]), _ptr(stash), rows, _ptr(grad_out), _ptr(flat_grad),
                    _ptr(grad_in), _stream()), "siren_backward_input")
            return flat_grad
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b200inr_siren_backward(
                ctypes.byref(self._desc), _ptr(eng["packed"]), _ptr(stash), _ptr(coords),
                ctypes.byref(grid) if grid else None, rows, _ptr(grad_out), _ptr(flat_grad), _stream()),
                "siren_backward")
        return flat_grad

    # ---------------------------------------------------------------- nn.Module protocol
    def forward(self, coords):
        """out = INR.forward(x) (INR/superresDWI.py:134).  SRDWI.Siren detaches its input (INR/SRDWI.py:88), so does
        this; INRmodel.Siren does not (INR/INRmodel.py:147-149): when that variant is fed explicit feature rows that
        require grad (the PerturbNet phase, INR/inrDWI.py:141-147), dL/d(features) comes out of the backward kernel."""
        _require_cuda(coords, "forward input")
        if coords.dim() != 2 or coords.shape[1] != self.in_features:
            raise RuntimeError(f"b200inr: expected an input of shape [N, {self.in_features}]")
        keep = (getattr(self, "variant", "SRDWI") == "INRmodel" and coords.requires_grad and torch.is_grad_enabled())
        if keep and self._desc.input_mode != _lib.IN_FEATURES and self._desc.activation == _lib.ACT_GABOR:
            keep = False  # WIRE on raw coordinates: the coordinate grid is data, no producer needs its gradient
        if keep and self._desc.input_mode != _lib.IN_FEATURES:
            raise RuntimeError("b200inr: the input gradient is implemented for explicit feature rows (in_features >= 64), "
                               "the way the reference's PerturbNet phase feeds the network; raw coordinates carry none")
        coords = coords.contiguous().float() if keep else coords.detach().contiguous().float()
        return _SirenFunction.apply(coords, self, *self._canonical())

    # ---------------------------------------------------------------- fused query
    def query(self, shape, clamp_min=0.0, out=None, row_range=None):
        """torch.clamp(INR.forward(get_mgrid(shape)), min=0) of INR/superresDWI.py:161, coordinates derived in-kernel.

        row_range=(begin, end) restricts the call to a contiguous range of the flattened grid (a rank's shard).
        Returns [rows, out_features] fp32 on the module's device; pass clamp_min=None for the raw output.
        """
        shape = tuple(int(s) for s in shape)
        if self._grid_dim is None:
            raise RuntimeError("b200inr: a network fed with explicit features has no grid form; call forward(features)")
        if len(shape) != self._grid_dim:
            raise RuntimeError("b200inr: query grid rank must equal the coordinate dimension")
        total = int(np.prod(shape))
        begin, end = (0, total) if row_range is None else (int(row_range[0]), int(row_range[1]))
        grid = _lib.make_grid(shape, begin)
        with torch.no_grad():
            res, _ = self._forward_rows(None, grid, end - begin, train=False, clamp=clamp_min, out=out)
        return res

    # ---------------------------------------------------------------- fused fit
    def fit(self, target, shape, steps, lr=1e-4, degrade=None, betas=(0.9, 0.999), eps=1e-8, row_range=None,
            global_count=None, process_group=None, reset_optimizer=False, graph=None, weight=None):
        """The reference training loop (INR/superresDWI.py:132-138) on a dense coordinate grid, without autograd:
        per step  fused forward -> loss (+ LR degradation) -> fused backward -> [all-reduce] -> Adam -> re-stage bf16.

        target   CUDA fp32.  degrade=None: [rows, C] values at the grid points.  degrade='pool': the LR volume
                 [X/2, Y/2, Z, C] (flattened or not) that the 2x2x1 in-plane average of the prediction must match.
        shape    the (HR) coordinate grid; rows = prod(shape) unless row_range=(begin, end) selects this rank's slab.
        process_group / global_count: multi-GPU data parallel -- gradients are summed with one all-reduce per step and
                 the loss is normalised by the global element count.
        weight   optional [rows, C] loss weights: (w * (out - gt)**2).mean() of INR/INR_ERD.py:265 (degrade=None only).
        Returns the per-step loss as a CUDA tensor [steps] (this rank's share of the global mean).
        """
        session = FitSession(self, target, shape, lr=lr, degrade=degrade, betas=betas, eps=eps, row_range=row_range,
                             global_count=global_count, process_group=process_group, reset_optimizer=reset_optimizer,
                             weight=weight)
        losses = torch.zeros(steps, dtype=torch.float32, device=session.device)
        if graph is None:  # replay the step as one CUDA graph: the launch gaps are a third of a cfg1 step and still
            graph = process_group is None and steps >= 16  # 1.7 % of a cfg2 step (2.10 -> 2.06 ms)
        done = 0
        if graph:
            losses[0:1].copy_(session.step())  # one eager step (lazy initialisation, allocator warm-up)
            done = 1
            session.capture()
        for it in range(done, steps):
            losses[it:it + 1].copy_(session.step())
        session.finish()
        return losses


class Siren(_FusedMLP):
    """Reference INR/SRDWI.py:67-91 (default) or INR/INRmodel.py:122-151 (`variant='INRmodel'`).

    Same constructor arguments, registration order (final_linear is both an attribute and the last element of `net`,
    SURVEY.md App. A-1) and RNG consumption order as the reference class, so `torch.manual_seed(s)` gives identical
    initial weights and reference state_dicts load unchanged.
    """

    def __init__(self, in_features, hidden_features, hidden_layers, out_features, first_omega_0=30.,
                 hidden_omega_0=30., variant="SRDWI"):
        super().__init__()
        if variant not in ("SRDWI", "INRmodel"):
            raise ValueError("variant must be 'SRDWI' or 'INRmodel'")
        self.variant = variant
        self.in_features, self.hidden_features = int(in_features), int(hidden_features)
        self.hidden_layers, self.out_features = int(hidden_layers), int(out_features)
        bound = np.sqrt(6 / hidden_features) / hidden_omega_0

        def make_final():
            fin = nn.Linear(hidden_features, out_features)
            with torch.no_grad():
                fin.weight.uniform_(-bound, bound)
            return fin

        net = []
        if variant == "SRDWI":  # final linear first (INR/SRDWI.py:75-77)
            self.final_linear = make_final()
            net.append(SineLayer(in_features, hidden_features, is_first=True, omega_0=first_omega_0))
            for _ in range(hidden_layers):
                net.append(SineLayer(hidden_features, hidden_features, is_first=False, omega_0=hidden_omega_0))
        else:  # sine layers first with the class-default omega 30, final linear last (INR/INRmodel.py:133-143)
            first_omega_0 = 30.
            net.append(SineLayer(in_features, hidden_features, is_first=True))
            for _ in range(hidden_layers):
                net.append(SineLayer(hidden_features, hidden_features, is_first=False))
            self.final_linear = make_final()
        net.append(self.final_linear)
        self.net = nn.Sequential(*net)
        self.first_omega_0 = float(first_omega_0)
        self.hidden_omega_0 = float(net[1].omega_0) if hidden_layers > 0 else float(hidden_omega_0)
        # in_features <= 4: raw coordinates (first layer on CUDA cores, grid-mode fit/query available);
        # wider inputs are explicit feature rows, e.g. pre-computed Fourier features (INR/superresDWI.py:108-122).
        # The generic family's kernels are built for 256- and 512-wide operands: other widths (the reference's
        # Siren(in_features=256, hidden_features=128, ...) of SR3D.ipynb cell 4 / INR/automate_INR.py:23-27, or an input
        # wider than the hidden layers) run zero padded -- sin(0) = 0 units with zero outgoing weights change nothing
        # and receive zero gradients --, while parameters, gradients and state dicts keep the module's own shapes.
        mode = _lib.IN_COORDS if self.in_features <= 4 else _lib.IN_FEATURES
        self._hp = None
        h_eng = self.hidden_features
        if mode == _lib.IN_FEATURES:
            h_eng = _generic_width(self.hidden_features, self.in_features)
            self._hp = h_eng if h_eng != self.hidden_features else None
        self._init_engine(_lib.make_net(in_features, h_eng, hidden_layers, out_features, self.first_omega_0,
                                        self.hidden_omega_0, input_mode=mode),
                          self.in_features if mode == _lib.IN_COORDS else None)

    def _canonical(self):
        """Parameters in the flat-layout order of include/b200inr.h: W0 b0 ... WL bL Wf bf."""
        ps = []
        for i in range(self.hidden_layers + 1):
            ps += [self.net[i].linear.weight, self.net[i].linear.bias]
        ps += [self.final_linear.weight, self.final_linear.bias]
        return ps

    def _padded_shape(self, idx, p):
        return _mlp_padded_shape(self._hp, self.hidden_layers, idx, p)


class SirenERD(_FusedMLP):
    """The ReLU-tail SIREN of the reference's soft-ERD script (INR/INR_ERD.py:28-67, there called `Siren`):
    SineLayer(first), hidden_layers x SineLayer, nn.Linear(H, H) + ReLU, final nn.Linear, ReLU on the output.
    Same constructor arguments, module registration order, RNG consumption (the perturbation head's two linears are
    created and initialised last, exactly like the reference) and state-dict keys.  forward = one fused kernel
    (b200inr_net.flags = B200INR_NET_RELU_TAIL), backward = the layer-pipelined kernel; fit(weight=...) is the
    weighted loss of :264-266.  perturb=True is the notebook-only path that reads a global `model_input` (SURVEY.md
    App. A-14) and is not provided."""

    def __init__(self, in_features, hidden_features, hidden_layers, out_features, first_omega_0=30.,
                 hidden_omega_0=30., perturb=False):
        super().__init__()
        if perturb:
            raise NotImplementedError("b200inr: INR_ERD.Siren(perturb=True) depends on a notebook global in the reference")
        self.in_features, self.hidden_features = int(in_features), int(hidden_features)
        self.hidden_layers, self.out_features = int(hidden_layers), int(out_features)
        net = [SineLayer(in_features, hidden_features, is_first=True, omega_0=first_omega_0)]
        self.relu = nn.ReLU()
        for _ in range(hidden_layers):
            net.append(SineLayer(hidden_features, hidden_features, is_first=False, omega_0=hidden_omega_0))
        net.append(nn.Linear(hidden_features, hidden_features))
        net.append(nn.ReLU())
        self.final_linear = nn.Linear(hidden_features, out_features)
        bound = np.sqrt(6 / hidden_features) / hidden_omega_0
        with torch.no_grad():
            self.final_linear.weight.uniform_(-bound, bound)
        self.net = nn.Sequential(*net)
        self.perturb_linear = nn.Linear(3, hidden_features)
        self.perturb_linear2 = nn.Linear(hidden_features, out_features)
        with torch.no_grad():
            self.perturb_linear.weight.uniform_(-bound, bound)
            self.perturb_linear2.weight.uniform_(-bound, bound)
        self.tanh = nn.Tanh()
        self.perturb = perturb
        self.first_omega_0, self.hidden_omega_0 = float(first_omega_0), float(hidden_omega_0)
        # engine view: hidden_layers + 1 hidden layers, the last of them Linear + ReLU
        self._init_engine(_lib.make_net(in_features, hidden_features, hidden_layers + 1, out_features,
                                        self.first_omega_0, self.hidden_omega_0, flags=_lib.NET_RELU_TAIL),
                          self.in_features)

    def _canonical(self):
        ps = []
        for i in range(self.hidden_layers + 1):
            ps += [self.net[i].linear.weight, self.net[i].linear.bias]
        lin = self.net[self.hidden_layers + 1]
        ps += [lin.weight, lin.bias, self.final_linear.weight, self.final_linear.bias]
        return ps

    def forward(self, coords, sample=0, eps=0):
        return super().forward(coords)

    def query(self, shape, clamp_min=0.0, out=None, row_range=None):
        # (the network output is already >= 0; clamp_min=None cannot undo the output ReLU)
        return super().query(shape, clamp_min=clamp_min, out=out, row_range=row_range)


def soft_erd(signal, b0, noise_level, mul=1000.0, slope=20.0):
    """Soft-ERD over the last axis of `signal` (the acquisitions of one b-value), one kernel for the whole volume:
    returns (weights, soft_mean) -- the loss weights `accept` of INR/INR_ERD.py:222-235 (exp(x / T), or 1/n below the
    noise floor) and the softmax-weighted image of calc_adc_erd_single2 (:126-160).  CUDA fp32 tensors in and out."""
    _require_cuda(signal, "soft_erd signal")
    sig = signal.detach().contiguous().float()
    n = sig.shape[-1]
    b0t = b0.detach().to(sig.device).contiguous().float()
    if tuple(b0t.shape) != tuple(sig.shape[:-1]):
        raise RuntimeError("b200inr: soft_erd expects b0 of shape signal.shape[:-1]")
    weights = torch.empty_like(sig)
    mean = torch.empty(sig.shape[:-1], dtype=torch.float32, device=sig.device)
    voxels = mean.numel()
    with torch.cuda.device(sig.device):
        _lib.check(_lib.load().b200inr_soft_erd(_ptr(sig), _ptr(b0t), voxels, int(n), float(noise_level), float(mul),
                                               float(slope), _ptr(weights), _ptr(mean), _stream()), "soft_erd")
    return weights, mean


class FourierMLP(_FusedMLP):
    """Fourier features + MLP with the feature map fused into the first layer: the [N, 2m] matrix that input_mapping
    (INR/SRDWI.py:111-116) materialises, and that the reference re-reads from HBM every step, never exists.

    activation="relu": BASELINE config 4, nn.Sequential(Linear, ReLU, ..., Linear) with torch default init
                       (state-dict keys net.0.*, net.2.*, ...);
    activation="sine": the reference scripts' own combination, Siren(in_features=2m, ...) fed with
                       input_mapping(x, B) (INR/superresDWI.py:105-113): same construction order, init and keys.
    forward(coords [N, d]) == net(input_mapping(coords, B)); query / fit take the coordinate grid.
    """

    def __init__(self, in_features, mapping_size, hidden_features, hidden_layers, out_features, B, activation="relu",
                 first_omega_0=30., hidden_omega_0=30.):
        super().__init__()
        if activation not in ("relu", "sine"):
            raise ValueError("activation must be relu or sine")
        self.in_features, self.mapping_size = int(in_features), int(mapping_size)
        self.hidden_features, self.hidden_layers = int(hidden_features), int(hidden_layers)
        self.out_features, self.activation = int(out_features), activation
        B = torch.as_tensor(B, dtype=torch.float32)
        if tuple(B.shape) != (self.mapping_size, self.in_features):
            raise ValueError("B must have shape [mapping_size, in_features]")
        self.register_buffer("B", B.clone())
        k0 = 2 * self.mapping_size
        if activation == "sine":  # built exactly like Siren(in_features=2m, ...) (INR/SRDWI.py:75-85)
            bound = np.sqrt(6 / hidden_features) / hidden_omega_0
            self.final_linear = nn.Linear(hidden_features, out_features)
            with torch.no_grad():
                self.final_linear.weight.uniform_(-bound, bound)
            layers = [SineLayer(k0, hidden_features, is_first=True, omega_0=first_omega_0)]
            layers += [SineLayer(hidden_features, hidden_features, is_first=False, omega_0=hidden_omega_0)
                       for _ in range(hidden_layers)]
            self.net = nn.Sequential(*layers, self.final_linear)
            self._linears = [layer.linear for layer in layers] + [self.final_linear]
            act = _lib.ACT_SINE
        else:
            mods = [nn.Linear(k0, hidden_features), nn.ReLU()]
            for _ in range(hidden_layers):
                mods += [nn.Linear(hidden_features, hidden_features), nn.ReLU()]
            mods.append(nn.Linear(hidden_features, out_features))
            self.net = nn.Sequential(*mods)
            self._linears = [m for m in mods if isinstance(m, nn.Linear)]
            act = _lib.ACT_RELU
        h_eng = _generic_width(self.hidden_features, k0)  # other widths run zero padded (see Siren)
        self._hp = h_eng if h_eng != self.hidden_features else None
        self._init_engine(_lib.make_net(in_features, h_eng, hidden_layers, out_features, first_omega_0,
                                        hidden_omega_0, activation=act, input_mode=_lib.IN_FOURIER,
                                        mapping_size=mapping_size), self.in_features)

    def _canonical(self):
        ps = []
        for lin in self._linears:
            ps += [lin.weight, lin.bias]
        return ps

    def _padded_shape(self, idx, p):
        return _mlp_padded_shape(self._hp, self.hidden_layers, idx, p)

    def _frozen(self):
        return [self.B]


class ComplexGaborLayer2D(nn.Module):
    """Reference INR/INRmodel.py:66-120 (== wiretest.ipynb cell 1): parameter container of one complex Gabor layer.
    omega_0 / scale_0 are frozen nn.Parameters as in the reference (they appear in state_dict and parameters());
    the nested init_weights of the reference is dead code, so weights keep torch's default (complex) init.
    The arithmetic runs fused inside Wire."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega0=10.0, sigma0=10.0,
                 trainable=False):
        super().__init__()
        self.is_first = is_first
        self.in_features = in_features
        dtype = torch.float if is_first else torch.cfloat
        self.omega_0 = nn.Parameter(omega0 * torch.ones(1), trainable)
        self.scale_0 = nn.Parameter(sigma0 * torch.ones(1), trainable)
        self.linear = nn.Linear(in_features, out_features, bias=bias, dtype=dtype)
        self.scale_orth = nn.Linear(in_features, out_features, bias=bias, dtype=dtype)

    def forward(self, input):
        """Reference INR/INRmodel.py:109-120 for a stand-alone call of a FIRST layer (real coordinates in, complex64
        out): the layer runs as a 0-hidden-layer WIRE network whose final linear is zero and its activations
        [h_r | h_i] are read back from the training stash (bf16).  Hidden layers take complex inputs, which only exist
        on chip inside Wire.forward / fit / query."""
        if not self.is_first:
            raise RuntimeError("b200inr: a hidden ComplexGaborLayer2D runs fused inside Wire (its complex input never "
                               "leaves the chip); only first layers can be called stand-alone")
        _require_cuda(input, "ComplexGaborLayer2D input")
        lin, orth = self.linear, self.scale_orth
        k, h = lin.in_features, lin.out_features
        dev, rows = input.device, input.shape[0]
        if rows == 0:
            return torch.empty((0, h), dtype=torch.complex64, device=dev)
        desc = _lib.make_net(k, h, 0, 1, float(self.omega_0), float(self.omega_0), activation=_lib.ACT_GABOR,
                             scale_0=float(self.scale_0))
        tiles = (rows + 127) // 128
        with torch.no_grad(), torch.cuda.device(dev):
            lib = _lib.load()
            off = _lib.param_offsets(desc)  # W_lin b_lin W_orth b_orth | W_f b_f
            flat = torch.zeros(_lib.param_count(desc), dtype=torch.float32, device=dev)
            for o, t in zip(off[:4], (lin.weight, lin.bias, orth.weight, orth.bias)):
                if t is not None:
                    flat[o:o + t.numel()].copy_(t.detach().reshape(-1))
            packed = _aligned_bytes(_lib.packed_bytes(desc), dev)
            _lib.check(lib.b200inr_pack_weights(ctypes.byref(desc), _ptr(flat), _ptr(packed), _stream()), "pack_weights")
            x = input.detach().contiguous().float()
            out = torch.empty((rows, 1), dtype=torch.float32, device=dev)
            stash = _aligned_bytes(_lib.stash_bytes(desc, rows), dev, zero=False)
            _lib.check(lib.b200inr_siren_forward(ctypes.byref(desc), _ptr(packed), _ptr(x), None, rows, _ptr(out), 0,
                                                 0.0, _ptr(stash), _stream()), "siren_forward")
            y = stash[:tiles * 128 * 2 * h * 2].view(torch.bfloat16).view(tiles, 2 * h // 64, 128, 8, 8)
            r = torch.arange(128, device=dev)
            phys = torch.arange(8, device=dev)[None, :] ^ (r[:, None] & 7)
            y = y[:, :, r[:, None], phys].permute(0, 2, 1, 3, 4).reshape(tiles * 128, 2 * h)[:rows].float()
            return torch.complex(y[:, :h].contiguous(), y[:, h:].contiguous())


class Wire(_FusedMLP):
    """The WIRE network of the reference (INR/wiretest.ipynb cell 2, there also called `Siren`):
    Sequential(ComplexGaborLayer2D(first) , hidden_layers x ComplexGaborLayer2D, complex Linear), real part returned.
    Same constructor arguments, construction order (RNG) and state-dict keys.  hidden_features = complex units (128).

    in_features <= 4: raw coordinates (grid-mode fit / query available).
    in_features  > 4: explicit feature rows, the notebook's own configuration (cell 7:
        Siren(in_features=2*mapping_size, hidden_features=128, hidden_layers=3, out_features=1, first_omega_0=1.2,
        hidden_omega_0=1.2, scale=1.2) fed with input_mapping(coords, B)): the first Gabor layer is a tensor-core layer
        (K = in_features, a multiple of 64 up to 512), and forward() hands back dL/d(features) when its input requires
        grad (the PerturbNet phase of cell 10).
    B=[m, in_features] (extension): input_mapping fused into the first layer -- forward(coords) ==
        net(input_mapping(coords, B)) without the [N, 2m] matrix, and query / fit take the coordinate grid."""

    def __init__(self, in_features, hidden_features, hidden_layers, out_features, first_omega_0=10,
                 hidden_omega_0=30., scale=10.0, B=None):
        super().__init__()
        self.in_features, self.hidden_features = int(in_features), int(hidden_features)
        self.hidden_layers, self.out_features = int(hidden_layers), int(out_features)
        self.variant = "INRmodel"  # the notebook's class does not detach its input
        mode, mapping, k_in = _lib.IN_COORDS, 0, self.in_features
        if B is not None:
            B = torch.as_tensor(B, dtype=torch.float32)
            if B.dim() != 2 or B.shape[1] != self.in_features:
                raise ValueError("B must have shape [mapping_size, in_features]")
            self.register_buffer("B", B.clone())
            mode, mapping, k_in = _lib.IN_FOURIER, int(B.shape[0]), 2 * int(B.shape[0])
        elif self.in_features > 4:
            mode = _lib.IN_FEATURES
        net = [ComplexGaborLayer2D(k_in, hidden_features, omega0=first_omega_0, sigma0=scale, is_first=True,
                                   trainable=False)]
        for _ in range(hidden_layers):
            net.append(ComplexGaborLayer2D(hidden_features, hidden_features, is_first=False, omega0=hidden_omega_0,
                                           sigma0=scale))
        self.final_linear = nn.Linear(hidden_features, out_features, dtype=torch.cfloat)
        net.append(self.final_linear)
        self.net = nn.Sequential(*net)
        self._init_engine(_lib.make_net(in_features, hidden_features, hidden_layers, out_features, first_omega_0,
                                        hidden_omega_0, activation=_lib.ACT_GABOR, scale_0=scale, input_mode=mode,
                                        mapping_size=mapping),
                          None if mode == _lib.IN_FEATURES else self.in_features)

    def _canonical(self):
        ps = []
        for i in range(self.hidden_layers + 1):
            layer = self.net[i]
            ps += [layer.linear.weight, layer.linear.bias, layer.scale_orth.weight, layer.scale_orth.bias]
        ps += [self.final_linear.weight, self.final_linear.bias]
        return ps

    def _frozen(self):
        return [self.B] if "B" in self._buffers else []


class FitSession:
    """Device-resident state of one fused fit (buffers, Adam moments, activation stash) and its step function.

    step() issues, on the current stream and without host synchronisation:
        fused forward (stash) -> loss / degradation + dL/dpred -> dgrad -> wgrad
        -> [all-reduce of the flat [grad | loss] buffer] -> optimiser step (Adam + gradient clearing + step counter +
        bf16 re-staging of the weights: b200inr_optimizer_step, one launch for raw-coordinate SIRENs)
    `marks`, when given, receives a CUDA event after every stage (used by bench.py for per-kernel timing).
    """

    STAGES = ("forward", "loss", "dgrad", "wgrad", "allreduce", "optimizer")

    def __init__(self, module, target, shape, lr=1e-4, degrade=None, betas=(0.9, 0.999), eps=1e-8, row_range=None,
                 global_count=None, process_group=None, reset_optimizer=False, weight=None):
        _require_cuda(target, "fit target")
        self.module = module
        eng = self.eng = module._sync_params()
        dev = self.device = eng["device"]
        self.lib = _lib.load()
        shape = tuple(int(s) for s in shape)
        total = int(np.prod(shape))
        begin, end = (0, total) if row_range is None else (int(row_range[0]), int(row_range[1]))
        rows = self.rows = end - begin
        C = self.C = module.out_features
        if module._grid_dim is None or len(shape) != module._grid_dim:
            raise RuntimeError("b200inr: fit grid rank must equal the coordinate dimension")
        self.grid = _lib.make_grid(shape, begin)
        self.degrade = degrade
        target = target.detach().contiguous().float().reshape(-1)
        if degrade is None:
            if target.numel() != rows * C:
                raise RuntimeError("b200inr: fit target must have rows*C elements")
            self.count = float(global_count if global_count is not None else rows * C)
        elif degrade == "pool":
            if len(shape) != 3:
                raise RuntimeError("b200inr: degrade='pool' needs a 3-D grid (X, Y, Z)")
            plane = shape[1] * shape[2]
            if rows % (2 * plane) or begin % (2 * plane) or shape[1] % 2:
                raise RuntimeError("b200inr: pooled fit needs whole pairs of x-planes and an even Y")
            self.x_local, self.Y, self.ZC = rows // plane, shape[1], shape[2] * C
            if target.numel() != rows * C // 4:
                raise RuntimeError("b200inr: pooled fit target must have rows*C/4 elements")
            self.count = float(global_count if global_count is not None else rows * C // 4)
        elif degrade == "blur_pool":
            # Gaussian sigma = 0.5 (mirror boundary) then 2x2x1 average: skimage rescale(0.5, anti_aliasing=True)
            # (dwi_inr.ipynb#c6:L8).  The blur crosses slab borders: a rank that owns the x-planes
            # [xa, xb) receives its neighbours' four edge planes of the prediction after the forward (one grouped
            # send/recv) and evaluates the residual on one extra LR row per side (b200inr_blurpool_mse_slab); its
            # `target` then holds the LR rows [max(xa/2-1, 0), min(xb/2+1, X/2)) (parallel.lr_slab(..., halo=1)).
            if len(shape) != 3:
                raise RuntimeError("b200inr: degrade='blur_pool' needs a 3-D grid (X, Y, Z)")
            if shape[0] % 2 or shape[1] % 2:
                raise RuntimeError("b200inr: blurred pooling needs even X and Y")
            plane = shape[1] * shape[2]
            if rows % (2 * plane) or begin % (2 * plane) or rows == 0:
                raise RuntimeError("b200inr: a blurred-pooling slab is a non-empty range of whole pairs of x-planes")
            self.X, self.Y, self.ZC = shape[0], shape[1], shape[2] * C
            xa, xb = begin // plane, end // plane
            self.slab = (xa, xb)
            px0, px1 = max(xa - 4, 0), min(xb + 4, shape[0])
            ie0, ie1 = max(xa // 2 - 1, 0), min(xb // 2 + 1, shape[0] // 2)
            lr_row = (shape[1] // 2) * shape[2] * C
            if target.numel() != (ie1 - ie0) * lr_row:
                raise RuntimeError("b200inr: blurred-pooling target must hold the slab's LR rows plus one halo row per "
                                   "interior side (parallel.lr_slab(target, shape, row_range, halo=1))")
            self.count = float(global_count if global_count is not None else total * C // 4)
            self.halo = None
            if px0 < xa or px1 > xb:  # interior borders: neighbours' planes are needed
                if process_group is None:
                    raise RuntimeError("b200inr: a partial blur_pool slab needs the process group of its neighbours")
                if xb - xa < 4:
                    raise RuntimeError("b200inr: blur_pool slabs must hold at least four x-planes (the halo width)")
                import torch.distributed as dist
                rk = dist.get_rank(process_group)
                self.halo = {"left": dist.get_global_rank(process_group, rk - 1) if px0 < xa else None,
                             "right": dist.get_global_rank(process_group, rk + 1) if px1 > xb else None,
                             "nl": (xa - px0) * plane, "nr": (px1 - xb) * plane, "edge": 4 * plane}
            # banded form of D along x and y (6 taps per LR row, 3 per HR row): operands of b200inr_blurpool_mse
            self.bands = []
            for n_hr in (shape[0], shape[1]):
                fwd6, adj3 = _lib.build_band_tables(n_hr, True)
                self.bands.append((torch.from_numpy(fwd6).to(dev), torch.from_numpy(adj3).to(dev)))
            self.lr_resid = torch.empty((ie1 - ie0) * lr_row, dtype=torch.float32, device=dev)
            self.pred_ext = torch.empty(((px1 - px0) * plane, C), dtype=torch.float32, device=dev)
        else:
            raise ValueError("degrade must be None, 'pool' or 'blur_pool'")
        if degrade is not None and (module._desc.flags & _lib.NET_RELU_TAIL):
            raise RuntimeError("b200inr: the ReLU-tail network is fitted with the point-wise (weighted) loss of "
                               "INR/INR_ERD.py; the degradation losses have no output-ReLU mask")
        self.target = target
        # per-element loss weights: (w * (out - gt)**2).mean() of INR/INR_ERD.py:265 (point-wise loss only)
        self.weight = None
        if weight is not None:
            if degrade is not None:
                raise RuntimeError("b200inr: loss weights go with the point-wise loss (degrade=None)")
            _require_cuda(weight, "fit weight")
            self.weight = weight.detach().contiguous().float().reshape(-1)
            if self.weight.numel() != rows * C:
                raise RuntimeError("b200inr: fit weight must have rows*C elements")
        if module._optim is None or reset_optimizer or module._optim["m"].device != dev:
            module._optim = {"m": torch.zeros_like(eng["flat"]), "v": torch.zeros_like(eng["flat"]),
                             "state": torch.zeros(4, dtype=torch.float32, device=dev)}
        self.opt = module._optim
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.process_group = process_group
        self.n_flat = eng["flat"].numel()
        # [grad ..., loss accumulator, pad]: cleared once here, then by every optimiser step (no per-step memset).
        # Multi-GPU: two alternating buffers in peer-mapped memory, summed across ranks INSIDE the optimiser-step
        # kernel (b200inr_optimizer_step_peers); without peer access (or with B200INR_PEER_ALLREDUCE=0) one NCCL
        # all-reduce of the buffer precedes the single-GPU optimiser step.
        self.peer = None
        self._parity = 0
        if process_group is not None and torch.distributed.get_world_size(process_group) > 1 and \
                os.environ.get("B200INR_PEER_ALLREDUCE", "1") == "1":
            from . import parallel
            try:
                self.peer = parallel.PeerGradients(self.n_flat + 4, dev, process_group)
            except Exception as exc:  # no symmetric memory on this topology: keep the NCCL path
                if os.environ.get("B200INR_PEER_ALLREDUCE_STRICT", "0") == "1":
                    raise
                import warnings
                warnings.warn(f"b200inr: peer-mapped gradient exchange unavailable ({exc}); using ncclAllReduce")
        if self.peer is not None:
            self.grads = self.peer.grads[0]
        else:
            self.grads = torch.zeros(self.n_flat + 4, dtype=torch.float32, device=dev)
        self.loss_acc = self.grads[self.n_flat:self.n_flat + 1]
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)  # the last finished step's loss
        if degrade == "blur_pool":  # the forward writes the own planes into the middle of the halo-extended buffer
            lo = (self.slab[0] - max(self.slab[0] - 4, 0)) * shape[1] * shape[2]
            self.pred = self.pred_ext[lo:lo + rows]
        else:
            self.pred = torch.empty((rows, C), dtype=torch.float32, device=dev)
        self.dpred = torch.empty((rows, C), dtype=torch.float32, device=dev)
        self.stash = _aligned_bytes(_lib.stash_bytes(module._desc, rows), dev)
        d = module._desc
        # raw-coordinate SIREN without the staged flag: backward is one kernel (the marks 'dgrad' and 'wgrad' then
        # bracket that kernel and nothing, respectively)
        self.piped = (d.activation == _lib.ACT_SINE and d.input_mode == _lib.IN_COORDS
                      and not (d.flags & _lib.NET_STAGED_BWD))
        # Pooled loss fused into the forward's final epilogue (b200inr_siren_forward_pool_loss): possible when a 128-row
        # tile holds whole y pairs and its pooling partner is a whole tile away.  It removes one launch and 260 MB of
        # HBM traffic per cfg2 step, but measured on B200 it is a tie at full size (0.754 ms vs 0.706 + 0.049 ms: the
        # epilogue warps are the forward's critical resource and, unlike the plain copy-out, the loss arithmetic does
        # not hide under their waits for the final MMAs), so it is the default only for small slabs (strong-scaled
        # shards, cfg1-like sizes), where the saved launch counts.  B200INR_FUSED_LOSS=0/1 overrides.
        self.fused_loss = False
        want = os.environ.get("B200INR_FUSED_LOSS", "1" if rows <= (1 << 18) else "0") == "1"
        if degrade == "pool" and self.piped and want:
            Y, Z = shape[1], shape[2]
            self.fused_loss = (128 % (2 * Z) == 0 and (Y * Z) % 128 == 0 and rows % (2 * Y * Z) == 0
                               and begin % (2 * Y * Z) == 0)
        # forward, loss (2 kernels for blur_pool), backward (dgrad + wgrad when staged), optimiser step (+ pack kernel
        # for the families whose operands are re-staged by a second launch)
        self.kernel_launches_per_step = (1 + (2 if degrade == "blur_pool" else (0 if self.fused_loss else 1)) +
                                         (1 if self.piped else 2) +
                                         (1 if (d.activation == _lib.ACT_SINE and d.input_mode == _lib.IN_COORDS) else 2))
        self._graph = None
        self._copy_stream = None

    def capture(self):
        """Capture one step (every kernel is stream-ordered, Adam's step counter lives on the device) into CUDA graphs;
        later step() calls replay them.  One graph per (gradient-buffer parity, target buffer): the multi-GPU step
        alternates two gradient buffers, and stage_target() / commit_target() alternate two target buffers, so a new
        combination is captured the first time step() meets it.  Not available when the gradient exchange is an NCCL
        call or the step exchanges halo planes (both stay eager)."""
        if self.process_group is not None and self.peer is None:
            raise RuntimeError("b200inr: graph capture of a multi-GPU step needs the in-kernel gradient exchange "
                               "(the NCCL all-reduce stays eager)")
        if getattr(self, "halo", None) is not None:
            raise RuntimeError("b200inr: a sharded blur_pool step exchanges halo planes with NCCL send/recv and stays "
                               "eager")
        self._graph = {}
        self._capture_current()

    def _capture_current(self):
        tkey = self.target.data_ptr()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.peer is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_eager(None)
                self._graph[(0, tkey)] = g
            else:
                # the two gradient buffers alternate: one graph per parity (every rank captures and replays the same
                # sequence, so the kernels' start barriers keep meeting; the epoch counter lives on the device)
                start = self._parity
                for _ in range(2):
                    par = self._parity
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._step_eager(None)  # (toggles self._parity on the host)
                    self._graph[(par, tkey)] = g
                self._parity = start
                self.grads = self.peer.grads[start]
                self.loss_acc = self.grads[self.n_flat:self.n_flat + 1]

    def set_target(self, target):
        """Replace the target values (same size), e.g. from pinned host memory."""
        self.target.copy_(target.reshape(-1), non_blocking=True)

    def stage_target(self, host_target):
        """Start the host-to-device copy of the NEXT step's target (pinned host memory) on a side stream into the back
        buffer, so that it travels while the current step computes; commit_target() makes it the current target."""
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
                self._target_back = torch.empty_like(self.target)
                self._copy_done = torch.cuda.Event()
                self._steps_done = torch.cuda.Event()
            # the back buffer was the target of the step before last: every step issued so far must be done with it
            self._steps_done.record(torch.cuda.current_stream())
            self._copy_stream.wait_event(self._steps_done)
            with torch.cuda.stream(self._copy_stream):
                self._target_back.copy_(host_target.reshape(-1), non_blocking=True)
                self._copy_done.record(self._copy_stream)

    def commit_target(self):
        """Make the staged target current (the compute stream waits for its copy)."""
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().wait_event(self._copy_done)
            self.target, self._target_back = self._target_back, self.target

    def step(self, marks=None):
        if self._graph is not None and marks is None:
            key = (self._parity if self.peer is not None else 0, self.target.data_ptr())
            if key not in self._graph:  # a target buffer this session has not stepped on yet
                self._capture_current()
            self._graph[key].replay()
            if self.peer is not None:  # multi-GPU: the gradient buffers alternate
                self._parity ^= 1
                self.grads = self.peer.grads[self._parity]
                self.loss_acc = self.grads[self.n_flat:self.n_flat + 1]
            return self.loss
        return self._step_eager(marks)

    def _step_eager(self, marks=None):
        m, eng, lib = self.module, self.eng, self.lib
        net, gref = ctypes.byref(m._desc), ctypes.byref(self.grid)
        rows, C = self.rows, self.C

        def mark():
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)

        with torch.cuda.device(self.device), torch.no_grad():
            s = _stream()
            mark()
            if rows == 0:
                # a rank without rows (more ranks than x-plane pairs): it only takes part in the gradient exchange
                for _ in range(4):
                    mark()
            else:
                self._issue_compute(net, gref, s, mark)
            if self.process_group is not None and self.peer is None:
                torch.distributed.all_reduce(self.grads, group=self.process_group)
            mark()
            if self.peer is not None:  # gradient exchange inside the kernel; the other buffer is cleared for the next step
                pg, par = self.peer, self._parity
                _lib.check(lib.b200inr_optimizer_step_peers(
                    net, _ptr(eng["flat"]), _ptr(pg.grads[par ^ 1]), _ptr(pg.peer_grads[par]), _ptr(pg.peer_flags),
                    pg.world, pg.rank, _ptr(self.opt["m"]), _ptr(self.opt["v"]), self.lr, self.betas[0], self.betas[1],
                    self.eps, _ptr(self.opt["state"]), _ptr(eng["packed"]), _ptr(self.loss), s), "optimizer_step_peers")
                self._parity = par ^ 1
                self.grads = pg.grads[self._parity]
                self.loss_acc = self.grads[self.n_flat:self.n_flat + 1]
            else:
                _lib.check(lib.b200inr_optimizer_step(net, _ptr(eng["flat"]), _ptr(self.grads), _ptr(self.opt["m"]),
                                                      _ptr(self.opt["v"]), self.lr, self.betas[0], self.betas[1],
                                                      self.eps, _ptr(self.opt["state"]), _ptr(eng["packed"]),
                                                      _ptr(self.loss), s), "optimizer_step")
            mark()
        return self.loss

    def _issue_compute(self, net, gref, s, mark):
        """forward -> loss (+ degradation) -> backward of this rank's rows; `mark` records the stage boundaries."""
        m, eng, lib = self.module, self.eng, self.lib
        rows, C = self.rows, self.C
        if self.fused_loss:  # forward + pooled loss + dL/dpred in one kernel: the prediction stays on chip
            _lib.check(lib.b200inr_siren_forward_pool_loss(net, _ptr(eng["packed"]), gref, rows, _ptr(self.target),
                                                           self.count, _ptr(self.dpred), _ptr(self.loss_acc),
                                                           _ptr(self.stash), s), "siren_forward_pool_loss")
        else:
            _lib.check(lib.b200inr_siren_forward(net, _ptr(eng["packed"]), None, gref, rows, _ptr(self.pred), 0,
                                                 0.0, _ptr(self.stash), s), "siren_forward")
        mark()
        if self.fused_loss:
            pass
        elif self.degrade is None:
            mse = lib.b200inr_mse_loss_relu_out if (m._desc.flags & _lib.NET_RELU_TAIL) else lib.b200inr_mse_loss
            _lib.check(mse(_ptr(self.pred), _ptr(self.target), _ptr(self.weight), rows * C, self.count,
                           _ptr(self.dpred), _ptr(self.loss_acc), s), "mse_loss")
        elif self.degrade == "pool":
            _lib.check(lib.b200inr_pool_mse(_ptr(self.pred), _ptr(self.target), self.x_local, self.Y, self.ZC,
                                            self.count, _ptr(self.dpred), _ptr(self.loss_acc), s), "pool_mse")
        else:  # r = D pred - target (+ loss), then dL/dpred = D^T 2 r / count: two streaming passes
            if self.halo is not None:
                self._exchange_halo()
            (bx6, ax3), (by6, ay3) = self.bands
            _lib.check(lib.b200inr_blurpool_mse_slab(_ptr(self.pred_ext), _ptr(self.target), self.X, self.Y, self.ZC,
                                                     self.count, _ptr(bx6), _ptr(by6), _ptr(ax3), _ptr(ay3),
                                                     self.slab[0], self.slab[1], _ptr(self.lr_resid), _ptr(self.dpred),
                                                     _ptr(self.loss_acc), s), "blurpool_mse_slab")
        mark()
        if self.piped:  # one layer-pipelined kernel: dgrad chain + every weight / bias gradient
            _lib.check(lib.b200inr_siren_backward(net, _ptr(eng["packed"]), _ptr(self.stash), None, gref, rows,
                                                  _ptr(self.dpred), _ptr(self.grads), s), "siren_backward")
            mark()
        else:
            _lib.check(lib.b200inr_siren_dgrad(net, _ptr(eng["packed"]), _ptr(self.stash), rows,
                                               _ptr(self.dpred), s), "siren_dgrad")
            mark()
            _lib.check(lib.b200inr_siren_wgrad(net, _ptr(self.stash), None, gref, rows, _ptr(self.grads), s),
                       "siren_wgrad")
        mark()

    def _exchange_halo(self):
        """Blurred pooling across slab borders: the four edge planes of this rank's prediction go to each neighbour and
        theirs arrive in the halo planes of pred_ext -- one grouped NCCL send/recv per step, stream-ordered (the only
        data-path exchange of the fit besides the gradient sum)."""
        import torch.distributed as dist
        h, pg = self.halo, self.process_group
        n = self.pred_ext.shape[0]
        ops = []
        if h["left"] is not None:
            ops.append(dist.P2POp(dist.isend, self.pred[:h["edge"]], h["left"], group=pg))
            ops.append(dist.P2POp(dist.irecv, self.pred_ext[:h["nl"]], h["left"], group=pg))
        if h["right"] is not None:
            ops.append(dist.P2POp(dist.isend, self.pred[self.rows - h["edge"]:], h["right"], group=pg))
            ops.append(dist.P2POp(dist.irecv, self.pred_ext[n - h["nr"]:], h["right"], group=pg))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def finish(self):
        """Write the fp32 master weights back into the module's nn.Parameters."""
        with torch.cuda.device(self.device):
            self.module._writeback_params(self.eng)
