"""Conv VAE of GPPVAE on stock torch -- the encoder / decoder around the GP term.

BASELINE.json's north_star keeps this part on stock torch ("reported but not optimised"): it is here so that the
epoch of train_gppvae.py can be driven end to end with the B200 GP term (gppvae_b200/epoch.py).  The module tree
and parameter names follow /root/reference/pysrc/faceplace/vae.py:23-125 so that `state_dict`s interchange with
checkpoints written by the reference's train_vae.py (`econv.{i}.conv{1,2}`, `dconv.{i}.conv{1,2}`, `dense_zm`,
`dense_zs`, `dense_dec`, `vy`).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

_ACTS = {"elu": F.elu, "relu": F.relu, "linear": lambda t: t}


class _Down(nn.Module):
    """3x3 conv at full resolution, then a stride-2 3x3 conv (vae.py:23-33)."""

    def __init__(self, ni: int, no: int, act: str):
        super().__init__()
        self.act = _ACTS[act]
        self.conv1 = nn.Conv2d(ni, no, 3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(no, no, 3, stride=2, padding=1)

    def forward(self, x):
        return self.act(self.conv2(self.act(self.conv1(x))))


class _Up(nn.Module):
    """Nearest-neighbour x2 upsampling, then two 3x3 convs (vae.py:36-49)."""

    def __init__(self, ni: int, no: int, act1: str, act2: str):
        super().__init__()
        self.act1, self.act2 = _ACTS[act1], _ACTS[act2]
        self.conv1 = nn.Conv2d(ni, no, 3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(no, no, 3, stride=1, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        return self.act2(self.conv2(self.act1(self.conv1(x))))


class FaceVAE(nn.Module):
    def __init__(self, img_size: int = 128, nf: int = 32, zdim: int = 256, steps: int = 5, colors: int = 3,
                 act: str = "elu", vy: float = 1e-3):
        super().__init__()
        self.red_img_size = img_size // (2 ** steps)
        self.nf = nf
        self.size_flat = self.red_img_size ** 2 * nf
        self.K = img_size ** 2 * colors                                   # pixels per image (vae.py:63)
        self.vy = nn.Parameter(torch.tensor([vy]), requires_grad=False)   # fixed observation variance (vae.py:67)
        self.econv = nn.ModuleList([_Down(colors if i == 0 else nf, nf, act) for i in range(steps)])
        self.dconv = nn.ModuleList([_Up(nf, nf, act, act) for _ in range(steps - 1)] + [_Up(nf, colors, act, "linear")])
        self.dense_zm = nn.Linear(self.size_flat, zdim)
        self.dense_zs = nn.Linear(self.size_flat, zdim)
        self.dense_dec = nn.Linear(zdim, self.size_flat)

    def encode(self, x):
        """Posterior mean and (softplus) scale of the latent code (vae.py:90-96)."""
        for cell in self.econv:
            x = cell(x)
        x = x.reshape(-1, self.size_flat)
        return self.dense_zm(x), F.softplus(self.dense_zs(x))

    def sample(self, x, eps):
        zm, zs = self.encode(x)
        return zm + eps * zs

    def decode(self, z):
        x = self.dense_dec(z).reshape(-1, self.nf, self.red_img_size, self.red_img_size)
        for cell in self.dconv:
            x = cell(x)
        return x

    def nll(self, x, xr):
        """Gaussian reconstruction term per image and its mean squared error (vae.py:110-114)."""
        mse = ((xr - x) ** 2).reshape(x.shape[0], self.K).mean(1, keepdim=True)
        return mse / (2 * self.vy) + 0.5 * torch.log(self.vy), mse

    def forward(self, x, eps):
        """Plain-VAE ELBO terms (vae.py:116-125; used by train_vae.py, not by the GPPVAE epoch)."""
        zm, zs = self.encode(x)
        xr = self.decode(zm + eps * zs)
        nll, mse = self.nll(x, xr)
        kld = -0.5 * (1 + 2 * torch.log(zs) - zm ** 2 - zs ** 2).sum(1, keepdim=True) / self.K
        return nll + kld, mse, nll, kld
