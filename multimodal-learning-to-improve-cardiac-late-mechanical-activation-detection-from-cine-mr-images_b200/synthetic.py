"""Synthetic cine stacks and initial velocities.

There is no network for datasets, and the reference's dataset is a private
path (/root/reference/configs/config.json:7), so benchmarks and tests use
annulus-shaped myocardium masks (the reference's "images" ARE binary myocardium
masks, /root/reference/README.md:21) and smooth random initial velocities that
stand in for the missing registration network.  Pure torch, device-agnostic;
this is data synthesis, not the measured path.
"""
from __future__ import annotations

import math

import torch


def synthetic_masks(B: int, T: int, H: int, W: int, seed: int = 2434, device="cpu") -> torch.Tensor:
    """(B,1,T,H,W) fp32 binary annuli that contract and relax over the T frames.  The per-slice random draws always
    come from a CPU generator; the pixel arithmetic runs on ``device`` (large benchmark batches are built on the GPU;
    tests build on the CPU and copy, so that oracle and CUDA path see the same bits)."""
    g = torch.Generator().manual_seed(seed)
    dr = (torch.rand(B, generator=g) * 8 - 4).view(B, 1, 1, 1).to(device)
    dc = (torch.rand(B, generator=g) * 8 - 4).view(B, 1, 1, 1).to(device)
    phi0 = (torch.rand(B, generator=g) * 2 * math.pi).view(B, 1, 1, 1).to(device)
    t = torch.arange(T, dtype=torch.float32, device=device).view(1, T, 1, 1)
    ph = torch.sin(math.pi * t / max(T - 1, 1))
    r_in = 0.18 * H * (1 - 0.25 * ph)
    r_out = 0.30 * H * (1 - 0.10 * ph)
    rr = torch.arange(H, dtype=torch.float32, device=device).view(1, 1, H, 1) - (H / 2 + dr)
    cc = torch.arange(W, dtype=torch.float32, device=device).view(1, 1, 1, W) - (W / 2 + dc)
    rad = torch.sqrt(rr * rr + cc * cc)
    theta = torch.atan2(rr, cc)
    mod = 1 + 0.05 * torch.cos(3 * theta + phi0)
    mask = ((rad >= r_in * mod) & (rad <= r_out * mod)).to(torch.float32)
    return mask.unsqueeze(1)


def synthetic_v0(P: int, H: int, W: int, seed: int = 7, max_disp: float = 3.0,
                 params=(1.0, 0.1, 0.05), device="cpu") -> torch.Tensor:
    """(P,2,H,W) smooth velocities: sharp(noise) rescaled so max |v0| = ``max_disp`` px per pair."""
    alpha, beta, gamma = params
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(P, 2, H, W, generator=g)
    k0 = torch.arange(H, dtype=torch.float64).view(H, 1)
    k1 = torch.arange(W // 2 + 1, dtype=torch.float64).view(1, -1)
    c0, s0 = 2 * (1 - torch.cos(2 * math.pi * k0 / H)), torch.sin(2 * math.pi * k0 / H)
    c1, s1 = 2 * (1 - torch.cos(2 * math.pi * k1 / W)), torch.sin(2 * math.pi * k1 / W)
    lam = gamma + alpha * (c0 + c1)
    L00, L11, L01 = lam + beta * c0, lam + beta * c1, beta * s0 * s1
    det = L00 * L11 - L01 * L01
    F = torch.fft.rfft2(eps.to(torch.float64), norm="ortho")
    G0 = (L11 * F[:, 0] - L01 * F[:, 1]) / det
    G1 = (L00 * F[:, 1] - L01 * F[:, 0]) / det
    v = torch.fft.irfft2(torch.stack([G0, G1], 1), s=(H, W), norm="ortho").to(torch.float32)
    mag = v.pow(2).sum(1).sqrt().amax(dim=(1, 2)).clamp(min=1e-12).view(P, 1, 1, 1)
    return (v * (max_disp / mag)).to(device)
