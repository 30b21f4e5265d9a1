"""Drop-in `Vmodel` / `normalize_rows`: the reference's vmod.py API on top of the sm_100a kernels.

Same constructor `Vmodel(P, Q, p, q)`, parameters `x0 (P x p)` and `v0 (Q x q)`, methods `x()`, `v()`,
`forward(d, w)` and initialisation as /root/reference/pysrc/faceplace/vmod.py:10-40.  (In the reference
`Q` is the number of views; everywhere else in this package Q is the rank p*q.)
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .ops import normalize_rows  # noqa: F401  (re-exported: vmod.py:10-12)


class Vmodel(nn.Module):
    def __init__(self, P: int, Q: int, p: int, q: int):
        super().__init__()
        self.x0 = nn.Parameter(torch.randn(P, p))     # vmod.py:18
        self.v0 = nn.Parameter(torch.randn(Q, q))     # vmod.py:19
        self._init_params()

    def x(self) -> torch.Tensor:
        """Row-normalised object table (vmod.py:22-23)."""
        return ops.normalize_rows(self.x0)

    def v(self) -> torch.Tensor:
        """Row-normalised view table (vmod.py:25-26)."""
        return ops.normalize_rows(self.v0)

    def forward(self, d: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """V[i, j*q + k] = x()[d_i, j] * v()[w_i, k]  (vmod.py:28-35), differentiable w.r.t. x0 and v0."""
        return ops._KhatriRao.apply(self.x(), self.v(), d, w)

    def _init_params(self) -> None:
        """vmod.py:37-40: objects start at e_0 (+1e-3 noise), views at the identity (+1e-3 noise)."""
        with torch.no_grad():
            self.x0[:, 0] = 1.0
            self.x0[:, 1:] = 1e-3 * torch.randn(self.x0.shape[0], self.x0.shape[1] - 1)
            self.v0.copy_(torch.eye(*self.v0.shape) + 1e-3 * torch.randn(*self.v0.shape))
