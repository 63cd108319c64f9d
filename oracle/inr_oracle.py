"""CPU oracle of the INR fit / query hot path of MRIRC/MRI-super-resolution.

TEST INFRASTRUCTURE ONLY.  Nothing under mri-super-resolution_b200/ may import this file: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or the
timed CPU baseline -- never as the product path.

It restates, in plain NumPy (explicit forward AND hand-derived backward, fp32) and in plain CPU PyTorch (autograd,
used for multi-step trajectories and for timing the CPU baseline exactly the way the reference runs), the arithmetic
of the reference (paths relative to the reference checkout, INR/ = implicit-neural-representations/):

    get_mgrid            INR/SRDWI.py:12-18
    input_mapping        INR/SRDWI.py:111-116
    SineLayer            INR/SRDWI.py:41-64
    Siren                INR/SRDWI.py:67-91 and INR/INRmodel.py:122-151
    ComplexGaborLayer2D  INR/INRmodel.py:66-120, WIRE network INR/wiretest.ipynb cells 1-2
    fit loop             INR/superresDWI.py:132-138
    query                INR/superresDWI.py:161
    Adam                 torch.optim.Adam as called at INR/superresDWI.py:115-116 (PyTorch, un-vendored dependency:
                         INR/requirements.txt:81 pins torch==2.0.0; restated from its documented update rule)

Pinning: tests/test_oracle_golden.py checks every function here against tests/golden/*.npz, which
tools/make_golden.py generated in the build container by importing the UNMODIFIED reference modules
(SRDWI.py, INRmodel.py, wiretest.ipynb cells 1-2) under fixed seeds.  The LR degradation operator has no in-loop
counterpart in the reference: its oracle here is defined by the survey
(avg-pool 2x2x1, optional Gaussian sigma=0.5 mirror pre-blur == skimage rescale(0.5, anti_aliasing=True)) and is
pinned against scipy.ndimage / torch.nn.functional.avg_pool3d instead -- parity unpinned against the reference for
that one operator, and DESIGN.md says so.
"""
import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------- coordinates
def get_mgrid(shape):
    """INR/SRDWI.py:12-18 -- torch.linspace(-1, 1, n) per axis, meshgrid 'ij', last axis fastest.

    torch.linspace evaluates symmetrically from both ends in fp32: i < n//2 -> -1 + step*i, else 1 - step*(n-1-i),
    with step = 2/(n-1) rounded to fp32 (ATen RangeFactories).
    """
    axes = []
    for n in shape:
        n = int(n)
        if n == 1:
            axes.append(np.array([-1.0], dtype=F32))
            continue
        step = F32(2.0) / F32(n - 1)
        i = np.arange(n)
        lo = F32(-1.0) + (step * i.astype(F32)).astype(F32)
        hi = F32(1.0) - (step * (n - 1 - i).astype(F32)).astype(F32)
        axes.append(np.where(i < n // 2, lo, hi).astype(F32))
    mesh = np.meshgrid(*axes, indexing="ij")
    return np.stack(mesh, axis=-1).reshape(-1, len(shape)).astype(F32)


def input_mapping(x, B):
    """INR/SRDWI.py:111-116 -- [sin(2 pi x B^T), cos(2 pi x B^T)], sin block first; B None -> identity."""
    if B is None:
        return x
    proj = (F32(2.0 * np.pi) * x.astype(F32)) @ B.astype(F32).T
    return np.concatenate([np.sin(proj), np.cos(proj)], axis=-1).astype(F32)


# --------------------------------------------------------------------------------------------- SIREN, explicit math
def siren_forward(weights, biases, x, first_omega_0=30.0, hidden_omega_0=30.0, return_intermediates=False):
    """Siren.forward (INR/SRDWI.py:87-91): sine layers sin(omega * (h W^T + b)) (INR/SRDWI.py:59) then a plain linear.

    weights/biases: lists of L+2 arrays in nn.Linear layout [out, in] / [out]; the last pair is final_linear.
    Returns out [N, C] (and the list of sine-layer outputs plus pre-activations omega*z when asked).
    """
    h = x.astype(F32)
    acts, pre = [], []
    n_sine = len(weights) - 1
    for l in range(n_sine):
        omega = F32(first_omega_0 if l == 0 else hidden_omega_0)
        theta = omega * (h @ weights[l].T.astype(F32) + biases[l].astype(F32))
        h = np.sin(theta).astype(F32)
        pre.append(theta.astype(F32))
        acts.append(h)
    out = (h @ weights[-1].T.astype(F32) + biases[-1].astype(F32)).astype(F32)
    if return_intermediates:
        return out, acts, pre
    return out


def siren_backward(weights, biases, x, grad_out, first_omega_0=30.0, hidden_omega_0=30.0):
    """Hand-derived gradients of siren_forward, float64 accumulation.

    dz = omega * cos(omega z) * g ; dW = dz^T h_prev ; db = sum_rows dz ; g_prev = dz W.
    Returns (list dW, list db) in the layout of `weights` / `biases`.
    """
    _, acts, pre = siren_forward(weights, biases, x, first_omega_0, hidden_omega_0, True)
    n_sine = len(weights) - 1
    dW = [None] * len(weights)
    db = [None] * len(weights)
    g = grad_out.astype(np.float64)
    dW[-1] = g.T @ acts[-1].astype(np.float64)
    db[-1] = g.sum(0)
    g = g @ weights[-1].astype(np.float64)
    for l in range(n_sine - 1, -1, -1):
        omega = float(first_omega_0 if l == 0 else hidden_omega_0)
        dz = omega * np.cos(pre[l].astype(np.float64)) * g
        h_prev = (acts[l - 1] if l > 0 else x).astype(np.float64)
        dW[l] = dz.T @ h_prev
        db[l] = dz.sum(0)
        if l > 0:
            g = dz @ weights[l].astype(np.float64)
    return [a.astype(F32) for a in dW], [a.astype(F32) for a in db]


# --------------------------------------------------------------------------------------------- WIRE, explicit math
def gabor_layer_forward(W1, b1, W2, b2, x, omega_0, scale_0):
    """ComplexGaborLayer2D.forward (INR/INRmodel.py:109-120): exp(1j w0 lin) * exp(-s0^2 (|lin|^2 + |orth|^2))."""
    lin = x @ W1.T + b1
    orth = x @ W2.T + b2
    freq = np.exp(1j * omega_0 * lin)
    gauss = np.exp(-(scale_0 ** 2) * (np.abs(lin) ** 2 + np.abs(orth) ** 2))
    return (freq * gauss).astype(np.complex64)


def wire_forward(layers, final_W, final_b, x, omega_0, scale_0):
    """WIRE network of INR/wiretest.ipynb cell 2 (L1-33): Gabor layers, complex final linear, real part returned.

    layers: list of (W1, b1, W2, b2); the first layer's are real, the rest complex64.
    """
    h = x.astype(F32)
    for (W1, b1, W2, b2) in layers:
        h = gabor_layer_forward(W1, b1, W2, b2, h, omega_0, scale_0)
    return (h @ final_W.T + final_b).real.astype(F32)


# --------------------------------------------------------------------------------------------- loss / degradation
def mse_loss(pred, target, weight=None):
    """((out - gt)**2).mean() (INR/superresDWI.py:135); weighted form (w*(out-gt)**2).mean() (INR/INR_ERD.py:265).
    Returns (loss, dloss/dpred)."""
    r = pred.astype(np.float64) - target.astype(np.float64)
    w = 1.0 if weight is None else weight.astype(np.float64)
    n = r.size
    return float((w * r * r).sum() / n), (2.0 * w * r / n).astype(F32)


GAUSS_SIGMA_HALF = None


def gaussian_taps_sigma_half():
    """scipy.ndimage.gaussian_filter(sigma=0.5, truncate=4.0) kernel: radius 2, normalised."""
    t = np.arange(-2, 3, dtype=np.float64)
    g = np.exp(-0.5 * t * t / 0.25)
    return g / g.sum()


def degrade_axis_matrix(n_hr, blur):
    """Dense [n_hr/2, n_hr] matrix of the 1-D operator: optional 5-tap Gaussian (mirror boundary) then 2-tap box mean."""
    n_lr = n_hr // 2
    Bm = np.eye(n_hr)
    if blur:
        g = gaussian_taps_sigma_half()
        Bm = np.zeros((n_hr, n_hr))
        period = 2 * (n_hr - 1)
        for x in range(n_hr):
            for t in range(-2, 3):
                i = (x + t) % period
                i = i if i < n_hr else period - i
                Bm[x, i] += g[t + 2]
    P = np.zeros((n_lr, n_hr))
    for i in range(n_lr):
        P[i, 2 * i] = 0.5
        P[i, 2 * i + 1] = 0.5
    return P @ Bm


def degrade_forward(vol, blur=False):
    """D: [X, Y, ...] -> [X/2, Y/2, ...] in-plane only (the axes the reference decimates, INR/superresDWI.py:94)."""
    X, Y = vol.shape[:2]
    Dx, Dy = degrade_axis_matrix(X, blur), degrade_axis_matrix(Y, blur)
    return np.einsum("ax,by,xy...->ab...", Dx, Dy, vol.astype(np.float64)).astype(F32)


def degrade_adjoint(vol_lr, blur=False):
    """D^T: [X/2, Y/2, ...] -> [X, Y, ...]."""
    X, Y = vol_lr.shape[0] * 2, vol_lr.shape[1] * 2
    Dx, Dy = degrade_axis_matrix(X, blur), degrade_axis_matrix(Y, blur)
    return np.einsum("ax,by,ab...->xy...", Dx, Dy, vol_lr.astype(np.float64)).astype(F32)


def degraded_mse(pred_hr, target_lr, blur=False):
    """L = mean((D pred - target_lr)^2) and dL/dpred = D^T 2 (D pred - target_lr) / n_lr."""
    r = degrade_forward(pred_hr, blur).astype(np.float64) - target_lr.astype(np.float64)
    n = r.size
    return float((r * r).sum() / n), degrade_adjoint((2.0 * r / n).astype(F32), blur)


# --------------------------------------------------------------------------------------------- Adam
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (defaults: no amsgrad, no weight decay), fp32 state, `step` counted from 1.
    m <- lerp(m, g, 1-b1); v <- b2 v + (1-b2) g^2; p <- p - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)."""
    p, g, m, v = (a.astype(F32) for a in (p, g, m, v))
    m = (m + (g - m) * F32(1.0 - beta1)).astype(F32)
    v = (F32(beta2) * v + F32(1.0 - beta2) * g * g).astype(F32)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (np.sqrt(v) / F32(math.sqrt(bc2)) + F32(eps)).astype(F32)
    p = (p - F32(lr / bc1) * (m / denom)).astype(F32)
    return p, m, v


# --------------------------------------------------------------------------------------------- metrics
def psnr(a, b, data_range=1.0):
    """10 log10(R^2 / MSE) (skimage.metrics.peak_signal_noise_ratio semantics, INR/inr_toy.py:16)."""
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10.0 * math.log10(data_range ** 2 / max(mse, 1e-30))


def ssim2d(a, b, data_range=1.0):
    """skimage.metrics.structural_similarity defaults (INR/superresDWI.py:186): 7x7 uniform window, K1=0.01, K2=0.03,
    sample covariance, mean over the image cropped by 3 px."""
    from scipy.ndimage import uniform_filter
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    win, npx = 7, 49
    cov_norm = npx / (npx - 1.0)
    ux, uy = uniform_filter(a, win), uniform_filter(b, win)
    uxx, uyy, uxy = uniform_filter(a * a, win), uniform_filter(b * b, win), uniform_filter(a * b, win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())


def ssim_volume(a, b, data_range=1.0):
    """Mean of ssim2d over the slices of [X, Y, Z] or [X, Y, Z, C] volumes (per-slice use, INR/superresDWI.py:179-187)."""
    a = a.reshape(a.shape[0], a.shape[1], -1)
    b = b.reshape(b.shape[0], b.shape[1], -1)
    return float(np.mean([ssim2d(a[:, :, k], b[:, :, k], data_range) for k in range(a.shape[2])]))


# --------------------------------------------------------------------------------------------- torch restatement
def torch_siren(in_features, hidden_features, hidden_layers, out_features, first_omega_0=30.0, hidden_omega_0=30.0,
                order="SRDWI"):
    """A CPU PyTorch module with the arithmetic AND the RNG consumption of the reference Siren.

    order='SRDWI'   : final linear constructed first (INR/SRDWI.py:75-81)
    order='INRmodel': sine layers first, final linear last (INR/INRmodel.py:133-143; first omega fixed at 30)
    Parameter order of .parameters(): final linear first for 'SRDWI' (SURVEY.md App. A-1).
    """
    import torch
    from torch import nn

    class _Sine(nn.Module):
        def __init__(self, fan_in, fan_out, first, omega):
            super().__init__()
            self.omega_0 = omega
            self.linear = nn.Linear(fan_in, fan_out)
            bound = 1.0 / fan_in if first else math.sqrt(6.0 / fan_in) / omega
            with torch.no_grad():
                self.linear.weight.uniform_(-bound, bound)

        def forward(self, h):
            return torch.sin(self.omega_0 * self.linear(h))

    class _Net(nn.Module):
        def __init__(self):
            super().__init__()
            bound = math.sqrt(6.0 / hidden_features) / hidden_omega_0

            def make_final():
                fin = nn.Linear(hidden_features, out_features)
                with torch.no_grad():
                    fin.weight.uniform_(-bound, bound)
                return fin

            if order == "SRDWI":
                self.final_linear = make_final()
            w0 = first_omega_0 if order == "SRDWI" else 30.0
            wh = hidden_omega_0 if order == "SRDWI" else 30.0
            layers = [_Sine(in_features, hidden_features, True, w0)]
            layers += [_Sine(hidden_features, hidden_features, False, wh) for _ in range(hidden_layers)]
            if order != "SRDWI":
                self.final_linear = make_final()
            self.net = nn.Sequential(*layers, self.final_linear)

        def forward(self, coords):
            return self.net(coords)

    return _Net()


def torch_siren_erd(in_features, hidden_features, hidden_layers, out_features, first_omega_0=30.0, hidden_omega_0=30.0):
    """The ReLU-tail network of INR/INR_ERD.py:28-67 (perturb=False): SineLayer(first), hidden_layers x SineLayer,
    Linear(H, H) + ReLU, final Linear (U(+-sqrt(6/H)/omega_h) weights), ReLU on the output -- with the reference's
    registration order (final_linear before net) and RNG consumption (sine layers, Linear(H, H), final linear, then
    the two linears of the unused perturbation head, :47-51)."""
    import torch
    from torch import nn

    class _Sine(nn.Module):
        def __init__(self, fan_in, fan_out, first, omega):
            super().__init__()
            self.omega_0 = omega
            self.linear = nn.Linear(fan_in, fan_out)
            bound = 1.0 / fan_in if first else math.sqrt(6.0 / fan_in) / omega
            with torch.no_grad():
                self.linear.weight.uniform_(-bound, bound)

        def forward(self, h):
            return torch.sin(self.omega_0 * self.linear(h))

    class _Net(nn.Module):
        def __init__(self):
            super().__init__()
            layers = [_Sine(in_features, hidden_features, True, first_omega_0)]
            self.relu = nn.ReLU()
            layers += [_Sine(hidden_features, hidden_features, False, hidden_omega_0) for _ in range(hidden_layers)]
            layers += [nn.Linear(hidden_features, hidden_features), nn.ReLU()]
            self.final_linear = nn.Linear(hidden_features, out_features)
            bound = math.sqrt(6.0 / hidden_features) / hidden_omega_0
            with torch.no_grad():
                self.final_linear.weight.uniform_(-bound, bound)
            self.net = nn.Sequential(*layers)
            self.perturb_linear = nn.Linear(3, hidden_features)
            self.perturb_linear2 = nn.Linear(hidden_features, out_features)
            with torch.no_grad():
                self.perturb_linear.weight.uniform_(-bound, bound)
                self.perturb_linear2.weight.uniform_(-bound, bound)

        def forward(self, coords):
            return self.relu(self.final_linear(self.net(coords)))

    return _Net()


def soft_erd(signal, b0, noise_level, mul=1000.0, slope=20.0):
    """Soft-ERD of INR/INR_ERD.py in float64, vectorised over voxels: signal [..., n], b0 [...].
    weights [..., n] = the `accept` loop of :222-235 (exp(x / T) with T = max(mul exp(-slope mean(x)/b0), 2) where
    mean(x) > 2 noise_level, one-hot at the arg-max when exp overflows -- the script turns RuntimeWarning into that
    branch --, 1/n below the noise floor); soft_mean [...] = calc_adc_erd_single2's image (:143-156): sum(softmax(x/T) x),
    or mean(x) below the noise floor."""
    x = np.asarray(signal, dtype=np.float64)
    b0 = np.asarray(b0, dtype=np.float64)
    n = x.shape[-1]
    mean = x.mean(-1)
    strong = mean > 2.0 * noise_level
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        T = np.maximum(mul * np.exp(-slope * (mean / b0)), 2.0)
        w = np.exp(x / T[..., None])
        over = np.isinf(w).any(-1)
        onehot = (np.arange(n) == np.argmax(x, -1)[..., None]).astype(np.float64)
        soft = (w * x).sum(-1) / w.sum(-1)
    weights = np.where(strong[..., None], np.where(over[..., None], onehot, w), 1.0 / n)
    soft_mean = np.where(strong, np.where(over, x.max(-1), soft), mean)
    return weights, soft_mean


def calculate_adc(bvalues, data):
    """calculate_ADC (INR/SRDWI.py:118-130) vectorised: np.polyfit(b / 1000, log(y + 1e-7), 1) per voxel in closed form
    (float64), ADC = -slope clamped to [-10, 3].  data [..., nb] -> [...]."""
    x = np.asarray(bvalues, dtype=np.float64).reshape(-1) / 1000.0
    y = np.log(np.asarray(data) + 1e-7).astype(np.float64)  # log in the input's own precision, as the reference
    xc = x - x.mean()
    slope = (y * xc).sum(-1) / (xc * xc).sum()
    return np.clip(-slope, -10.0, 3.0)


def torch_input_mapping(x, B):
    """input_mapping (INR/SRDWI.py:111-116) as a differentiable CPU torch expression: cat(sin p, cos p), p = 2 pi x B^T."""
    import torch
    p = (2.0 * math.pi * x) @ B.T
    return torch.cat([torch.sin(p), torch.cos(p)], dim=-1)


def torch_pn(in_features, hidden_features, dimension):
    """CPU restatement of the reference's perturbation network PN (INR/INRmodel.py:153-169): detach the input, append
    the acquisition index sample / 10 as one more column, Linear -> tanh -> Linear -> eps * tanh.  Same parameter names
    and construction order (the reference builds its index column with .cuda(); this one stays on the input's device)."""
    import torch
    from torch import nn

    class _PN(nn.Module):
        def __init__(self):
            super().__init__()
            self.perturb_linear = nn.Linear(in_features + 1, hidden_features)
            self.perturb_linear2 = nn.Linear(hidden_features, dimension)

        def forward(self, coords, sample=0, eps=0):
            x = coords.detach()
            col = torch.full((x.shape[0], 1), sample / 10.0, dtype=x.dtype)
            h = torch.tanh(self.perturb_linear(torch.cat([x, col], dim=-1)))
            return eps * torch.tanh(self.perturb_linear2(h))

    return _PN()


def torch_wire(in_features, hidden_features, hidden_layers, out_features, first_omega_0=10.0, hidden_omega_0=30.0,
               scale=10.0):
    """CPU PyTorch restatement of the WIRE network of INR/wiretest.ipynb cells 1-2 (layer: INR/INRmodel.py:66-120) with
    the reference's parameter names, registration and RNG order: per Gabor layer the frozen omega_0 / scale_0, then
    `linear` and `scale_orth` (real in the first layer, complex64 after it, torch default init -- the reference's
    nested init_weights is dead code); complex final linear; the real part is returned.
        out = exp(1j * omega_0 * lin) * exp(-scale_0^2 * (|lin|^2 + |orth|^2))"""
    import torch
    from torch import nn

    class _Gabor(nn.Module):
        def __init__(self, fan_in, fan_out, first, omega, sigma):
            super().__init__()
            self.omega_0 = nn.Parameter(omega * torch.ones(1), False)
            self.scale_0 = nn.Parameter(sigma * torch.ones(1), False)
            dt = torch.float if first else torch.cfloat
            self.linear = nn.Linear(fan_in, fan_out, dtype=dt)
            self.scale_orth = nn.Linear(fan_in, fan_out, dtype=dt)

        def forward(self, h):
            lin, orth = self.linear(h), self.scale_orth(h)
            env = torch.exp(-self.scale_0 * self.scale_0 * (lin.abs().square() + orth.abs().square()))
            return torch.exp(1j * self.omega_0 * lin) * env

    class _Net(nn.Module):
        def __init__(self):
            super().__init__()
            layers = [_Gabor(in_features, hidden_features, True, first_omega_0, scale)]
            layers += [_Gabor(hidden_features, hidden_features, False, hidden_omega_0, scale)
                       for _ in range(hidden_layers)]
            self.final_linear = nn.Linear(hidden_features, out_features, dtype=torch.cfloat)
            self.net = nn.Sequential(*layers, self.final_linear)

        def forward(self, coords):
            return self.net(coords).real

    return _Net()


def torch_perturb_loop(inr, pn, B, coords, mean_target, targets, number_of_epochs, pertubation_epochs, lr_inr=5e-5,
                       lr_pn=1e-6, eps=1 / 128.0):
    """The alternating INR / PerturbNet loop of INR/inrDWI.py:122-148 around CPU torch modules (torch_siren with
    order='INRmodel', torch_pn): INR steps on input_mapping(coords, B) against mean_target while
    ctr < number_of_epochs - pertubation_epochs or ctr is odd, otherwise one PerturbNet step per acquisition through
    INR(input_mapping(PN(model_input, sample, eps), B)).  Returns (inr_losses, pn_losses) as python lists."""
    import torch
    inr_optim = torch.optim.Adam(lr=lr_inr, params=list(inr.parameters()))
    pn_optim = torch.optim.Adam(lr=lr_pn, params=list(pn.parameters()))
    model_input = torch_input_mapping(coords, B)
    inr_losses, pn_losses = [], []
    for ctr in range(number_of_epochs):
        if ctr < number_of_epochs - pertubation_epochs or ctr % 2:
            loss = ((inr(model_input) - mean_target) ** 2).mean()
            inr_optim.zero_grad()
            loss.backward()
            inr_optim.step()
            inr_losses.append(float(loss.detach()))
        else:
            for sample, gt in enumerate(targets):
                feats = torch_input_mapping(pn(model_input, sample, eps), B)
                loss = ((inr(feats) - gt) ** 2).mean()
                pn_optim.zero_grad()
                loss.backward()
                pn_optim.step()
                pn_losses.append(float(loss.detach()))
    return inr_losses, pn_losses


def torch_relu_mlp(in_dim, hidden_features, hidden_layers, out_features):
    """BASELINE config 4's network: Linear(in, H) + ReLU, `hidden_layers` x (Linear(H, H) + ReLU), Linear(H, C) with
    torch's default initialisation, fed with input_mapping(coords, B) (the reference only ever
    feeds Fourier features into a SIREN, so the ReLU variant is a plain nn.Sequential)."""
    from torch import nn
    mods = [nn.Linear(in_dim, hidden_features), nn.ReLU()]
    for _ in range(hidden_layers):
        mods += [nn.Linear(hidden_features, hidden_features), nn.ReLU()]
    mods.append(nn.Linear(hidden_features, out_features))
    return nn.Sequential(*mods)


def torch_fit(model, coords, target, steps, lr, degrade=None, hr_shape=None, weight=None):
    """The reference's in-lined loop (INR/superresDWI.py:132-138): full batch, fixed order, Adam defaults.

    degrade None        : loss = ((out - target)**2).mean()
    degrade 'pool'      : out reshaped to hr_shape + (C,), 2x2x1 average pooled in-plane, then the same MSE against the
                          LR target.
    degrade 'blur_pool' : Gaussian sigma 0.5 (mirror) pre-blur, then the pooling (degrade_axis_matrix).
    Returns the list of per-step losses (python floats).
    """
    import torch
    import torch.nn.functional as Fn
    opt = torch.optim.Adam(lr=lr, params=list(model.parameters()))
    losses = []
    for _ in range(steps):
        out = model.forward(coords)
        if degrade == "pool":
            X, Y, Z = hr_shape
            vol = out.reshape(X, Y, Z, -1).permute(3, 0, 1, 2).unsqueeze(0)
            pooled = Fn.avg_pool3d(vol, kernel_size=(2, 2, 1), stride=(2, 2, 1))
            out = pooled.squeeze(0).permute(1, 2, 3, 0).reshape(-1, vol.shape[1])
        elif degrade == "blur_pool":
            X, Y, Z = hr_shape
            Dx = torch.from_numpy(degrade_axis_matrix(X, True)).float()
            Dy = torch.from_numpy(degrade_axis_matrix(Y, True)).float()
            vol = out.reshape(X, Y, Z, -1)
            out = torch.einsum("ax,by,xyzc->abzc", Dx, Dy, vol).reshape(-1, vol.shape[-1])
        # weight: the per-element loss weights of INR/INR_ERD.py:265, (w * (out - gt)**2).mean()
        loss = ((out - target) ** 2).mean() if weight is None else (weight * (out - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses
