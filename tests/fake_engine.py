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


# ---- structured route (csrc/structured.cu): the same contracts as the C entries, in float64 on CPU ----------------
def require_cuda_f32(t, name, ndim=2):
    return None


def _check_index(t, name, device):
    return t.contiguous()


def khatri_rao_fwd(xn, wn, d, w):
    return (xn.double()[d].unsqueeze(2) * wn.double()[w].unsqueeze(1)).reshape(d.shape[0], -1).to(torch.float32)


def kr_slot_sums(X, ldx, order, slot_start, xn, nviews, L, with_x):
    P, p = xn.shape
    counts = slot_start[1:] - slot_start[:-1]
    slot_of_sorted = torch.repeat_interleave(torch.arange(P * nviews), counts)
    rows = order[: slot_of_sorted.numel()]
    Zs = torch.zeros(P * nviews, L, dtype=torch.float64)
    Zs.index_add_(0, slot_of_sorted, X.double()[rows][:, :L])
    Zs = Zs.view(P, nviews * L)
    if not with_x:
        return Zs.to(torch.float32)
    Xn = (counts.view(P, nviews, 1).double() * xn.double().view(P, 1, p)).reshape(P, nviews * p)
    return torch.cat([Xn, Zs], 1).to(torch.float32)


def kr_assemble_gc(ST, wn, p, L, with_g):
    nv, q = wn.shape
    wd, ST = wn.double(), ST.double()
    off = nv * p if with_g else 0
    T = ST[:, off:].view(p, nv, L)
    C = torch.einsum("vk,jvl->jkl", wd, T).reshape(p * q, L)
    if not with_g:
        return C.to(torch.float32)
    S = ST[:, :off].view(p, nv, p)
    G = torch.einsum("vk,vm,jvi->jkim", wd, wd, S).reshape(p * q, p * q)
    return torch.cat([G, C], 1).to(torch.float32)


def kr_assemble_m(W, wn, p, L):
    nv, q = wn.shape
    return torch.einsum("vk,jkl->jvl", wn.double(), W.double().view(p, q, L)).reshape(p, nv * L).to(torch.float32)


def am(A, lda, M, ldm, n, k, m, alpha=1.0):
    return (alpha * (A.double()[:, :k] @ M.double())).to(torch.float32)


def kr_xb_nll(X, ldx, Y, d, w, P, nviews, L, scal):
    Yg = Y.double().view(P, nviews, L)[d, w]
    Xb = (X.double() - Yg) / scal[S_VN]
    quad = (X.double() * Xb).sum(1, keepdim=True)
    scal[S_XB2] = (Xb * Xb).sum()
    scal[S_QUAD] = quad.sum()
    return Xb.to(torch.float32), (0.5 * quad + scal[S_ROWCONST]).to(torch.float32)


# ---- differentiable ops (the product implements these as autograd Functions around CUDA kernels; here plain torch
# expressions of the same formulas -- vmod.py:10-12, 28-35, gp.py:127-133 -- with torch's own autograd)
def normalize_rows(x):
    return x / (x * x).sum(1, keepdim=True).sqrt()


class _KhatriRao:
    @staticmethod
    def apply(xn, wn, d, w):
        return (xn[d].unsqueeze(2) * wn[w].unsqueeze(1)).reshape(d.shape[0], -1)


class _TaylorExpansion:
    @staticmethod
    def apply(X, V, lvs, Xb, Vb, vbs):
        out = (Xb * X).sum(1, keepdim=True)
        if V is not None:
            out = out + (Vb * V).sum(1, keepdim=True)
        return out + (vbs * torch.softmax(lvs, 0)).sum() / float(X.shape[0])


def planes_supported(n, Q, L):
    return False     # the operand-plane kernels are tensor-core code: the stand-in always takes the fp32 entries


def install(monkeypatch):
    for name in ("normalize_rows", "_KhatriRao", "_TaylorExpansion"):
        monkeypatch.setattr(real_ops, name, globals()[name])
    for name in ("as_matrix", "gram_vtz", "atb", "factor", "solve_w", "xb_nll", "vbs_from_scal", "vb",
                 "require_cuda_f32", "_check_index", "khatri_rao_fwd", "kr_slot_sums", "kr_assemble_gc", "kr_assemble_m",
                 "am", "kr_xb_nll", "planes_supported"):
        monkeypatch.setattr(real_ops, name, globals()[name])
