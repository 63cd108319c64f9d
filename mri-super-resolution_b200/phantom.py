"""Deterministic synthetic DWI volume: closed-form, so every rank can regenerate its shard.

S0 = ellipsoid mask x (0.6 + 0.4 x Gaussian blobs); ADC map D in [0.5, 3.0]e-3 from a second blob set; channel 0 = b0,
channels 1..n_dirs = S0 exp(-1000 D (1 + 0.3 (g_k . n)^2)) with g_k the k-th Fibonacci-sphere direction and n a fixed
unit-vector field; Rician noise; each channel divided by its maximum (the reference max-normalises per (b, TE),
INR/superresDWI.py:50-55).
"""
import numpy as np


def _blobs(X, Y, Z, centres, widths, amps):
    out = np.zeros(X.shape, dtype=np.float64)
    for (cx, cy, cz), w, a in zip(centres, widths, amps):
        out += a * np.exp(-((X - cx) ** 2 + (Y - cy) ** 2 + (Z - cz) ** 2) / (2 * w * w))
    return out


def fibonacci_sphere(n):
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], -1)


def dwi_phantom(shape=(128, 128, 64), n_dirs=30, noise=0.01, seed=0, x_range=None):
    """float32 [X, Y, Z, 1 + n_dirs] in [0, 1].  x_range=(x0, x1) returns only that slab of x-planes."""
    nx, ny, nz = shape
    x0, x1 = (0, nx) if x_range is None else x_range
    xs = np.linspace(-1, 1, nx)[x0:x1]
    X, Y, Z = np.meshgrid(xs, np.linspace(-1, 1, ny), np.linspace(-1, 1, nz), indexing="ij")
    mask = ((X / 0.85) ** 2 + (Y / 0.75) ** 2 + (Z / 0.9) ** 2) < 1.0
    c1 = [(-0.3, -0.2, 0.1), (0.35, 0.25, -0.2), (0.0, 0.4, 0.3), (-0.4, 0.3, -0.4), (0.2, -0.45, 0.0), (0.0, 0.0, 0.0)]
    w1 = [0.25, 0.2, 0.3, 0.15, 0.22, 0.5]
    a1 = [1.0, 0.8, 0.6, 0.9, 0.7, 0.5]
    s0 = mask * (0.6 + 0.4 * np.clip(_blobs(X, Y, Z, c1, w1, a1) / 1.6, 0, 1))
    c2 = [(0.3, -0.3, 0.2), (-0.35, 0.2, -0.1), (0.1, 0.35, -0.35), (-0.1, -0.4, 0.4)]
    w2 = [0.3, 0.25, 0.2, 0.35]
    a2 = [1.0, 0.9, 0.8, 0.6]
    adc = (0.5 + 2.5 * np.clip(_blobs(X, Y, Z, c2, w2, a2) / 1.4, 0, 1)) * 1e-3
    nvec = np.stack([np.cos(2.0 * Y + 0.5), np.sin(2.0 * X - 0.3) * np.cos(1.5 * Z), np.sin(1.5 * Z + 1.0)], -1)
    nvec /= np.linalg.norm(nvec, axis=-1, keepdims=True) + 1e-12
    g = fibonacci_sphere(n_dirs)
    vol = np.empty(X.shape + (1 + n_dirs,), dtype=np.float64)
    vol[..., 0] = s0
    for k in range(n_dirs):
        proj = (nvec * g[k]).sum(-1)
        vol[..., 1 + k] = s0 * np.exp(-1000.0 * adc * (1 + 0.3 * proj ** 2))
    if noise > 0:
        rng = np.random.default_rng(seed + 7919 * x0)
        re = vol + noise * rng.standard_normal(vol.shape)
        im = noise * rng.standard_normal(vol.shape)
        vol = np.sqrt(re * re + im * im)
    # analytic per-channel maxima are <= 1 (b0) and <= exp(-0.5) (dwi): normalise by fixed constants so slabs agree
    scale = np.ones(1 + n_dirs)
    scale[1:] = np.exp(-0.5)
    vol = vol / scale
    return np.clip(vol, 0, None).astype(np.float32)


def avg_pool_inplane(vol):
    """LR target of BASELINE config 2: 2x2x1 average pooling over (x, y) -- the degradation operator D."""
    X, Y = vol.shape[:2]
    return vol.reshape(X // 2, 2, Y // 2, 2, *vol.shape[2:]).mean(axis=(1, 3)).astype(np.float32)
