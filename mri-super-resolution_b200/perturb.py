"""The PerturbNet phase of the reference as one fused loop.

Reference: INR/inrDWI.py:122-148 (same loop in INR/superresDWI.py:127-157, INR/forbagci.py, wiretest.ipynb cell 10):

    for ctr in range(number_of_epochs):
        if ctr < number_of_epochs - pertubation_epochs or ctr % 2:      # INR step on the mean image
            loss = ((INR.forward(model_input) - LR_ground_truth) ** 2).mean(); inr_optim step
        else:                                                            # PerturbNet steps, one per acquisition
            for sample in range(len(dataset)):
                perturbed_input = input_mapping(PerturbNet.forward(model_input, sample, 1/128.), B)
                loss = ((INR.forward(perturbed_input) - dataset.pixels[sample]) ** 2).mean(); perturb_optim step

with model_input = input_mapping(get_mgrid(shape), B), INR = INRmodel.Siren(2m, H, L, C), PerturbNet = PN(2m, 128, d).

What runs here per PerturbNet step (no autograd, no [N, 2m] matrix in HBM, no cuBLAS):
  * PN is a generic-family network (B200INR_ACT_TANH, B200INR_NET_TANH_OUT, scale_0 = eps) on in-kernel Fourier
    features of the grid: its 128 hidden units are zero-padded to the 256-wide operands, its constant acquisition
    column `sample / 10` is folded into the first bias (b200inr_pn_effective_params / b200inr_pn_fold_grad);
  * the INR runs on the perturbation as explicit coordinates with input_mapping fused into its first layer, with the
    dgrad-only stash (B200INR_NET_DGRAD_ONLY: the reference's autograd also computes INR weight gradients in this phase
    and throws them away -- perturb_optim only steps PN); its backward returns dL/d(perturbation) directly, the
    adjoint of input_mapping being applied to the input gradient while it is still in tensor memory
    (b200inr_siren_backward_coords);
  * PN's backward + weight gradients (b200inr_siren_backward_tanh_out) and one flat Adam over [PN | w_last].
INR steps are the ordinary fused fit (FitSession) on the same in-kernel features.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .inr import FitSession, FourierMLP, _aligned_bytes, _ptr, _require_cuda, _stream

PN_WIDTH = 256  # operand width of the generic family; a narrower PN is zero padded


class PerturbSession:
    """Device state of the alternating loop for one (INR, PerturbNet, B, grid) quadruple.

    inr      INRmodel.Siren(in_features=2m, H, L, C) (explicit-feature variant, H in {256, 512}) or a sine FourierMLP
    pn       PN(in_features=2m, hidden_features <= 256, dimension=d)
    B        [m, d] frequency matrix of input_mapping
    shape    the d-dimensional coordinate grid (rows = prod(shape)); row_range selects a contiguous slab
    inr_step(target) / perturb_step(target, sample) issue one optimiser step each and return the loss (CUDA scalar);
    finish() writes the trained weights back into `inr` and `pn`.
    """

    def __init__(self, inr, pn, B, shape, lr_inr=5e-5, lr_pn=1e-6, eps=1 / 128., betas=(0.9, 0.999), adam_eps=1e-8,
                 row_range=None):
        self.lib = _lib.load()
        self.inr, self.pn = inr, pn
        shape = tuple(int(s) for s in shape)
        dev = self.device = next(pn.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("b200inr: PerturbSession needs CUDA modules; there is no CPU path")
        B = torch.as_tensor(B, dtype=torch.float32, device=dev)
        m, d = int(B.shape[0]), int(B.shape[1])
        if len(shape) != d:
            raise RuntimeError("b200inr: the grid rank must equal B.shape[1]")
        total = int(np.prod(shape))
        begin, end = (0, total) if row_range is None else (int(row_range[0]), int(row_range[1]))
        self.rows, self.shape, self.d, self.m = end - begin, shape, d, m
        self.grid = _lib.make_grid(shape, begin)
        self.eps, self.lr_pn, self.betas, self.adam_eps = float(eps), float(lr_pn), betas, float(adam_eps)
        self.lr_inr = float(lr_inr)

        # ---- INR on in-kernel Fourier features: a sine FourierMLP sharing the caller's weights
        if isinstance(inr, FourierMLP):
            if inr.activation != "sine":
                raise RuntimeError("b200inr: the PerturbNet loop is the reference's SIREN pipeline")
            self.fm = inr
        else:
            if inr.in_features != 2 * m:
                raise RuntimeError("b200inr: INR.in_features must be 2 * mapping_size")
            self.fm = FourierMLP(d, m, inr.hidden_features, inr.hidden_layers, inr.out_features, B.cpu(),
                                 activation="sine", first_omega_0=inr.first_omega_0,
                                 hidden_omega_0=inr.hidden_omega_0).to(dev)
            sd = {k: v for k, v in inr.state_dict().items() if k.startswith("net.")}
            sd["B"] = B
            self.fm.load_state_dict(sd, strict=False)
        self.C = self.fm.out_features
        dsc = self.fm._desc
        self.inr_lean = _lib.make_net(dsc.in_features, dsc.hidden_features, dsc.hidden_layers, dsc.out_features,
                                      dsc.first_omega_0, dsc.hidden_omega_0, activation=_lib.ACT_SINE,
                                      input_mode=_lib.IN_FOURIER, mapping_size=m, flags=_lib.NET_DGRAD_ONLY)
        self._inr_sess = None
        self.inr_stash = _aligned_bytes(_lib.stash_bytes(self.inr_lean, self.rows), dev, zero=False)

        # ---- PN as a generic-family tanh network, master vector [net parameters | w_last]
        hp = pn.perturb_linear.out_features
        if pn.perturb_linear.in_features != 2 * m + 1 or pn.perturb_linear2.out_features != d or hp > PN_WIDTH:
            raise RuntimeError("b200inr: PN must be PN(in_features=2 * mapping_size, hidden_features <= 256, dimension=d)")
        self.hp = hp
        self.pn_net = _lib.make_net(d, PN_WIDTH, 0, d, activation=_lib.ACT_TANH, input_mode=_lib.IN_FOURIER,
                                    mapping_size=m, scale_0=self.eps, flags=_lib.NET_TANH_OUT)
        self.pn_off = _lib.param_offsets(self.pn_net)  # W1 b1 Wf bf B
        self.n_net = _lib.param_count(self.pn_net)
        n = self.n_net + PN_WIDTH
        self.master = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            k0 = 2 * m
            w1 = pn.perturb_linear.weight.detach()
            self.master[self.pn_off[0]:self.pn_off[0] + PN_WIDTH * k0].view(PN_WIDTH, k0)[:hp].copy_(w1[:, :k0])
            self.master[self.pn_off[1]:self.pn_off[1] + hp].copy_(pn.perturb_linear.bias.detach())
            self.master[self.pn_off[2]:self.pn_off[2] + d * PN_WIDTH].view(d, PN_WIDTH)[:, :hp].copy_(
                pn.perturb_linear2.weight.detach())
            self.master[self.pn_off[3]:self.pn_off[3] + d].copy_(pn.perturb_linear2.bias.detach())
            self.master[self.pn_off[4]:self.pn_off[4] + m * d].copy_(B.reshape(-1))
            self.master[self.n_net:self.n_net + hp].copy_(w1[:, k0])
        self.pn_eff = torch.zeros(self.n_net, dtype=torch.float32, device=dev)
        self.pn_grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.pn_m, self.pn_v = torch.zeros_like(self.master), torch.zeros_like(self.master)
        self.pn_state = torch.zeros(4, dtype=torch.float32, device=dev)
        self.pn_packed = _aligned_bytes(_lib.packed_bytes(self.pn_net), dev)
        self.pn_stash = _aligned_bytes(_lib.stash_bytes(self.pn_net, self.rows), dev, zero=False)
        self.pert = torch.empty((self.rows, d), dtype=torch.float32, device=dev)
        self.dpert = torch.empty((self.rows, d), dtype=torch.float32, device=dev)
        self.pred = torch.empty((self.rows, self.C), dtype=torch.float32, device=dev)
        self.dpred = torch.empty((self.rows, self.C), dtype=torch.float32, device=dev)
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=dev)
        self.kernel_launches_per_perturb_step = 12

    # ---------------------------------------------------------------- INR step (INR/inrDWI.py:124-128, :132-136)
    def inr_step(self, target):
        """One Adam step of the INR on the grid's own features against `target` ([rows, C], e.g. the mean image)."""
        if self._inr_sess is None:
            self._inr_sess = FitSession(self.fm, target, self.shape, lr=self.lr_inr, betas=self.betas,
                                        eps=self.adam_eps,
                                        row_range=(self.grid.row_begin, self.grid.row_begin + self.rows))
            self._inr_target = target
        elif target is not self._inr_target:
            self._inr_sess.set_target(target)
            self._inr_target = target
        return self._inr_sess.step()

    # ---------------------------------------------------------------- PerturbNet step (INR/inrDWI.py:138-147)
    def perturb_step(self, target, sample):
        """One Adam step of PN for acquisition `sample` against `target` [rows, C]; the INR is frozen."""
        _require_cuda(target, "perturb target")
        lib, rows = self.lib, self.rows
        acq = float(sample) / 10.0
        tgt = target.detach().contiguous().float().reshape(-1)
        if tgt.numel() != rows * self.C:
            raise RuntimeError("b200inr: perturb target must have rows*C elements")
        eng = self.fm._sync_params()
        pn_net, inr_net, gref = ctypes.byref(self.pn_net), ctypes.byref(self.inr_lean), ctypes.byref(self.grid)
        with torch.cuda.device(self.device), torch.no_grad():
            s = _stream()
            ck = _lib.check
            ck(lib.b200inr_pn_effective_params(_ptr(self.master), self.n_net, self.pn_off[1], PN_WIDTH, acq,
                                               _ptr(self.pn_eff), _ptr(self.pn_grads), s), "pn_effective_params")
            ck(lib.b200inr_pack_weights(pn_net, _ptr(self.pn_eff), _ptr(self.pn_packed), s), "pack_weights")
            self.loss_acc.zero_()
            # perturbation = eps * tanh(W2 tanh(W1 [features | acq] + b1) + b2)          (INR/INRmodel.py:160-167)
            ck(lib.b200inr_siren_forward(pn_net, _ptr(self.pn_packed), None, gref, rows, _ptr(self.pert), 0, 0.0,
                                         _ptr(self.pn_stash), s), "pn forward")
            # INR.forward(input_mapping(perturbation, B))                                (INR/inrDWI.py:142-143)
            ck(lib.b200inr_siren_forward(inr_net, _ptr(eng["packed"]), _ptr(self.pert), None, rows, _ptr(self.pred), 0,
                                         0.0, _ptr(self.inr_stash), s), "inr forward")
            ck(lib.b200inr_mse_loss(_ptr(self.pred), _ptr(tgt), None, rows * self.C, float(rows * self.C),
                                    _ptr(self.dpred), _ptr(self.loss_acc), s), "mse_loss")
            # loss.backward(): INR input gradient -> adjoint of input_mapping -> PN                 (INR/inrDWI.py:146)
            ck(lib.b200inr_siren_backward_coords(inr_net, _ptr(eng["packed"]), _ptr(self.inr_stash), _ptr(self.pert),
                                                 None, rows, _ptr(self.dpred), None, _ptr(self.dpert), s),
               "siren_backward_coords")
            ck(lib.b200inr_siren_backward_tanh_out(pn_net, _ptr(self.pn_packed), _ptr(self.pn_stash), rows,
                                                   _ptr(self.pert), _ptr(self.dpert), _ptr(self.pn_grads), s),
               "siren_backward_tanh_out")
            ck(lib.b200inr_pn_fold_grad(_ptr(self.pn_grads), self.n_net, self.pn_off[1], PN_WIDTH, acq, s),
               "pn_fold_grad")
            # perturb_optim.step()                                                                   (INR/inrDWI.py:147)
            ck(lib.b200inr_adam_step(_ptr(self.master), _ptr(self.pn_grads), _ptr(self.pn_m), _ptr(self.pn_v),
                                     self.master.numel(), self.lr_pn, self.betas[0], self.betas[1], self.adam_eps,
                                     _ptr(self.pn_state), s), "adam_step")
        return self.loss_acc

    def perturbation(self, sample):
        """PN.forward(model_input, sample, eps) of the current weights on the grid: [rows, d]."""
        acq = float(sample) / 10.0
        with torch.cuda.device(self.device), torch.no_grad():
            s = _stream()
            _lib.check(self.lib.b200inr_pn_effective_params(_ptr(self.master), self.n_net, self.pn_off[1], PN_WIDTH, acq,
                                                            _ptr(self.pn_eff), None, s), "pn_effective_params")
            _lib.check(self.lib.b200inr_pack_weights(ctypes.byref(self.pn_net), _ptr(self.pn_eff), _ptr(self.pn_packed),
                                                     s), "pack_weights")
            out = torch.empty_like(self.pert)
            _lib.check(self.lib.b200inr_siren_forward(ctypes.byref(self.pn_net), _ptr(self.pn_packed), None,
                                                      ctypes.byref(self.grid), self.rows, _ptr(out), 0, 0.0, None, s),
                       "pn forward")
        return out

    def finish(self):
        """Write the trained weights back into the caller's modules."""
        hp, m, d, k0 = self.hp, self.m, self.d, 2 * self.m
        if self._inr_sess is not None:
            self._inr_sess.finish()
        with torch.no_grad():
            if self.fm is not self.inr:
                self.inr.load_state_dict({k: v for k, v in self.fm.state_dict().items() if k != "B"}, strict=False)
            w1 = self.master[self.pn_off[0]:self.pn_off[0] + PN_WIDTH * k0].view(PN_WIDTH, k0)[:hp]
            self.pn.perturb_linear.weight[:, :k0].copy_(w1)
            self.pn.perturb_linear.weight[:, k0].copy_(self.master[self.n_net:self.n_net + hp])
            self.pn.perturb_linear.bias.copy_(self.master[self.pn_off[1]:self.pn_off[1] + hp])
            self.pn.perturb_linear2.weight.copy_(
                self.master[self.pn_off[2]:self.pn_off[2] + d * PN_WIDTH].view(d, PN_WIDTH)[:, :hp])
            self.pn.perturb_linear2.bias.copy_(self.master[self.pn_off[3]:self.pn_off[3] + d])


def perturb_fit(inr, pn, B, shape, mean_target, targets, number_of_epochs, pertubation_epochs, lr_inr=5e-5, lr_pn=1e-6,
                eps=1 / 128.):
    """The alternating training loop of INR/inrDWI.py:122-148, fused (see the module docstring).

    mean_target [rows, C]: LR_ground_truth of the INR steps; targets: sequence of per-acquisition [rows, C] tensors
    (dataset.pixels).  Returns (inr_losses, pn_losses): CUDA tensors with one entry per INR step and per PerturbNet
    step, in issue order.  The trained weights are written back into `inr` and `pn`.
    """
    sess = PerturbSession(inr, pn, B, shape, lr_inr=lr_inr, lr_pn=lr_pn, eps=eps)
    inr_losses, pn_losses = [], []
    for ctr in range(number_of_epochs):
        if ctr < number_of_epochs - pertubation_epochs or ctr % 2:
            inr_losses.append(sess.inr_step(mean_target).clone())
        else:
            for sample in range(len(targets)):
                pn_losses.append(sess.perturb_step(targets[sample], sample).clone())
    sess.finish()
    dev = sess.device
    cat = lambda xs: torch.cat(xs) if xs else torch.zeros(0, device=dev)  # noqa: E731
    return cat(inr_losses), cat(pn_losses)
