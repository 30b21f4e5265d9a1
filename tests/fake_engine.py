"""TEST-ONLY stand-in for gppvae_b200.ops on CPU tensors, built from the oracle's Q-space model.

Lets the `not gpu` suite exercise the HOST logic of the drop-in classes (row sharding, all-reduce packing,
n_total bookkeeping, factorisation cache, shape/padding plumbing) without a GPU.  The product never imports
this: tests monkeypatch it over `gppvae_b200.ops` attributes.
"""
import torch

from gppvae_b200 import ops as real_ops
from gppvae_b200._lib import NSCAL, S_LOGDETB, S_QUAD, S_ROWCONST, S_TRBINV, S_V0, S_VN, S_WNORM2, S_XB2


def as_matrix(t, name):
    t = t.detach().to(torch.float32)
    n, c = t.shape
    if c % 4 == 0 and t.is_contiguous():
        return t, c
    buf = torch.zeros(n, real_ops.round4(c))
    buf[:, :c] = t
    return buf, buf.shape[1]


def gram_vtz(V, ldv, X, ldx, n, Q, L):
    V = V.double()
    parts = [V.t() @ V] + ([V.t() @ X.double()] if L else [])
    return torch.cat(parts, 1).to(torch.float32)


def atb(A, lda, B, ldb, n, ka, kb):
    return (A.double().t() @ B.double()).to(torch.float32)


def factor(G, ldg, Q, vs, want_binv):
    vs = vs.detach().double()
    B = torch.eye(Q, dtype=torch.float64) + (vs[0] / vs[-1]) * G[:, :Q].double()
    Lc = torch.linalg.cholesky(B)
    Binv = torch.cholesky_inverse(Lc)
    scal = torch.zeros(NSCAL, dtype=torch.float64)
    scal[S_V0], scal[S_VN] = vs[0], vs[-1]
    scal[S_LOGDETB] = 2 * Lc.diagonal().log().sum()
    scal[S_TRBINV] = Binv.diagonal().sum()
    f = real_ops.Factorisation(Q, Binv, scal, Binv.to(torch.float32) if want_binv else None)
    return f


def solve_w(f, C, ldc, L, L_true, n_total):
    scal = f.scal.clone()
    W = (scal[S_V0] / scal[S_VN]) * (f.state @ C.double())
    scal[S_WNORM2] = (W * W).sum()
    scal[S_ROWCONST] = 0.5 * L_true * (scal[S_VN].log() + scal[S_LOGDETB] / n_total)
    return W.to(torch.float32), scal


def xb_nll(V, ldv, X, ldx, W, n, Q, L, scal):
    Xb = (X.double() - V.double() @ W.double()) / scal[S_VN]
    quad = (X.double() * Xb).sum(1, keepdim=True)
    scal[S_XB2] = (Xb * Xb).sum()
    scal[S_QUAD] = quad.sum()
    return Xb.to(torch.float32), (0.5 * quad + scal[S_ROWCONST]).to(torch.float32)


def vbs_from_scal(scal, n_total, Q, L):
    v0, vn, trb = scal[S_V0], scal[S_VN], scal[S_TRBINV]
    return torch.stack([-0.5 * scal[S_WNORM2] / (v0 * v0) + 0.5 * L * (Q - trb) / v0,
                        -0.5 * scal[S_XB2] + 0.5 * L * (n_total - Q + trb) / vn]).to(torch.float32)


def vb(V, ldv, Xb, Binv, W, scal, n, Q, L, L_true):
    r = scal[S_V0] / scal[S_VN]
    return (r * L_true * V.double() @ Binv.double() - Xb.double() @ W.double().t()).to(torch.float32)


def install(monkeypatch):
    for name in ("as_matrix", "gram_vtz", "atb", "factor", "solve_w", "xb_nll", "vbs_from_scal", "vb"):
        monkeypatch.setattr(real_ops, name, globals()[name])
