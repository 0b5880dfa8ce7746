"""Synthetic lesion-like patches (SURVEY.md section 8 d).  Used by tests and bench.py; there is no
network for real ISIC data.  ``make_patches`` is the NumPy generator (parity seed 0, perf seed
1234); ``make_patches_torch`` mirrors it with torch ops so bench.py can build 100k+ patches on
the device in seconds."""
from __future__ import annotations

import math

import numpy as np


def _blur_kernel(sigma=1.5, radius=4):
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def make_patches(n, H=64, W=None, seed=0, coverage=None, dtype=np.uint8, vmax=255):
    """Returns ``(images [n,H,W] dtype, masks [n,H,W] uint8 in {0,255})``."""
    from scipy.ndimage import convolve1d

    W = W or H
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    cy = H / 2 + rng.uniform(-H / 8, H / 8, n)
    cx = W / 2 + rng.uniform(-W / 8, W / 8, n)
    if coverage is None:
        ay = rng.uniform(0.20, 0.45, n) * H
        ax = rng.uniform(0.20, 0.45, n) * W
    else:  # target mask coverage (fraction of the patch), for the load-imbalance config
        cov = rng.uniform(coverage[0], coverage[1], n)
        r = np.sqrt(cov * H * W / math.pi)
        ecc = rng.uniform(0.7, 1.4, n)
        ay, ax = r * ecc, r / ecc
    th = rng.uniform(0, math.pi, n)
    amp = rng.uniform(0, 0.10 / 3, (n, 3))
    ph = rng.uniform(0, 2 * math.pi, (n, 3))
    imgs = np.empty((n, H, W), dtype=dtype)
    masks = np.empty((n, H, W), dtype=np.uint8)
    k = _blur_kernel()
    scale = vmax / 255.0
    for i in range(n):
        dy, dx = yy - cy[i], xx - cx[i]
        u = (dx * math.cos(th[i]) + dy * math.sin(th[i])) / ax[i]
        v = (-dx * math.sin(th[i]) + dy * math.cos(th[i])) / ay[i]
        rr = np.sqrt(u * u + v * v)
        ang = np.arctan2(v, u)
        pert = 1.0 + sum(amp[i, h] * np.cos((h + 2) * ang + ph[i, h]) for h in range(3))
        masks[i] = np.where(rr <= pert, 255, 0)
        soft = 1.0 / (1.0 + np.exp((rr - 1.0) * min(ax[i], ay[i]) / 2.0))  # ~2 px soft edge
        bg = rng.normal(185, 8, (H, W))
        les = rng.normal(95, 22, (H, W))
        tex = convolve1d(convolve1d(rng.normal(0, 1, (H, W)), k, axis=0, mode="reflect"), k, axis=1,
                         mode="reflect") * 18.0 / 0.19  # blurred unit noise has std ~0.19
        img = bg * (1 - soft) + les * soft + tex
        imgs[i] = np.clip(np.rint(img * scale), 0, vmax).astype(dtype)
    return imgs, masks


def make_patches_torch(n, H=64, W=None, seed=1234, device="cuda", chunk=16384):
    """Same recipe with torch ops on ``device`` (uint8 images, {0,255} masks)."""
    import torch
    import torch.nn.functional as F

    W = W or H
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    imgs = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    masks = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    yy, xx = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                            torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    k = torch.tensor(_blur_kernel(), dtype=torch.float32, device=device)
    kr = len(k) // 2

    def U(lo, hi, *shape):
        return torch.rand(shape, generator=g, device=device) * (hi - lo) + lo

    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        cy = (H / 2 + U(-H / 8, H / 8, m)).view(m, 1, 1)
        cx = (W / 2 + U(-W / 8, W / 8, m)).view(m, 1, 1)
        ay = (U(0.20, 0.45, m) * H).view(m, 1, 1)
        ax = (U(0.20, 0.45, m) * W).view(m, 1, 1)
        th = U(0, math.pi, m).view(m, 1, 1)
        amp = U(0, 0.10 / 3, m, 3)
        ph = U(0, 2 * math.pi, m, 3)
        dy, dx = yy - cy, xx - cx
        u = (dx * torch.cos(th) + dy * torch.sin(th)) / ax
        v = (-dx * torch.sin(th) + dy * torch.cos(th)) / ay
        rr = torch.sqrt(u * u + v * v)
        ang = torch.atan2(v, u)
        pert = torch.ones_like(rr)
        for h in range(3):
            pert = pert + amp[:, h].view(m, 1, 1) * torch.cos((h + 2) * ang + ph[:, h].view(m, 1, 1))
        masks[s:s + m] = torch.where(rr <= pert, 255, 0).to(torch.uint8)
        soft = torch.sigmoid(-(rr - 1.0) * torch.minimum(ax, ay) / 2.0)
        bg = torch.randn((m, H, W), generator=g, device=device) * 8 + 185
        les = torch.randn((m, H, W), generator=g, device=device) * 22 + 95
        noise = torch.randn((m, 1, H, W), generator=g, device=device)
        noise = F.conv2d(F.pad(noise, (0, 0, kr, kr), mode="reflect"), k.view(1, 1, -1, 1))
        noise = F.conv2d(F.pad(noise, (kr, kr, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
        img = bg * (1 - soft) + les * soft + noise[:, 0] * (18.0 / 0.19)
        imgs[s:s + m] = img.round().clamp(0, 255).to(torch.uint8)
    return imgs, masks
