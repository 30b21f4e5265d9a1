"""Seeded synthetic Face-Place-shaped inputs; shared by tests and bench.py.

q views, P = ceil(N/q) objects, every object seen in every view (d_i = i // q, w_i = i % q) under one
fixed row permutation; tables either `init` (Vmodel._init_params, vmod.py:37-40: the ill-conditioned
end) or `trained` (x0 ~ N(0,1), v0 = I + 0.5 N(0,1)); Z = 0.5 randn + V randn so the solve matters.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class Problem:
    x0: torch.Tensor     # (P, p)
    v0: torch.Tensor     # (q, q)
    d: torch.Tensor      # (N,) int64
    w: torch.Tensor      # (N,) int64
    Z: torch.Tensor      # (N, L)
    lvs: torch.Tensor    # (2,)
    meta: dict


def make_tables(P: int, p: int, q: int, kind: str, gen: torch.Generator, device="cpu"):
    if kind == "init":
        x0 = torch.empty(P, p, device=device)
        x0[:, 0] = 1.0
        x0[:, 1:] = 1e-3 * torch.randn(P, p - 1, generator=gen, device=device)
        v0 = torch.eye(q, device=device) + 1e-3 * torch.randn(q, q, generator=gen, device=device)
    elif kind == "trained":
        x0 = torch.randn(P, p, generator=gen, device=device)
        v0 = torch.eye(q, device=device) + 0.5 * torch.randn(q, q, generator=gen, device=device)
    else:
        raise ValueError(kind)
    return x0, v0


def make_problem(N: int, p: int, q: int, L: int, kind: str = "trained", lvs=(0.0, 0.0), seed: int = 0,
                 device="cpu", row_offset: int = 0, n_rows: int | None = None) -> Problem:
    """Build the problem (or, with row_offset / n_rows, one rank's contiguous row shard of it).

    Z needs V, which is formed here with plain torch ops -- this is input generation, not the path
    under test."""
    gen = torch.Generator(device=device).manual_seed(seed)
    P = -(-N // q)
    x0, v0 = make_tables(P, p, q, kind, gen, device)
    perm = torch.randperm(N, generator=gen, device=device)
    n_rows = N - row_offset if n_rows is None else n_rows
    rows = perm[row_offset:row_offset + n_rows]
    d = (rows // q).to(torch.int64)
    w = (rows % q).to(torch.int64)
    mix = torch.randn(p * q, L, generator=gen, device=device)
    xn = x0 / x0.norm(dim=1, keepdim=True)
    wn = v0 / v0.norm(dim=1, keepdim=True)
    # The noise is drawn in fixed chunks of global rows, each from its own seeded stream, so a row shard sees exactly
    # the rows the unsharded problem has (multi-GPU runs are comparable with the 1-GPU run).
    CH = 1 << 16
    Z = torch.empty(n_rows, L, device=device)
    for c in range(row_offset // CH, (row_offset + n_rows + CH - 1) // CH if n_rows else 0):
        g0, g1 = max(c * CH, row_offset), min((c + 1) * CH, row_offset + n_rows, N)
        if g1 <= g0:
            continue
        gen_c = torch.Generator(device=device).manual_seed(seed * 1000003 + 17 + c)
        noise = torch.randn(min(CH, N - c * CH), L, generator=gen_c, device=device)[g0 - c * CH:g1 - c * CH]
        s, e = g0 - row_offset, g1 - row_offset
        Vc = (xn[d[s:e]].unsqueeze(2) * wn[w[s:e]].unsqueeze(1)).reshape(e - s, -1)
        Z[s:e] = 0.5 * noise + Vc @ mix
    return Problem(x0, v0, d, w, Z, torch.tensor(lvs, dtype=torch.float32, device=device),
                   dict(N=N, p=p, q=q, Q=p * q, L=L, kind=kind, lvs=list(lvs), seed=seed,
                        row_offset=row_offset, n_rows=n_rows))


# the configurations of BASELINE.json
CONFIGS = {
    "c1": dict(N=4005, p=64, q=9, L=256),       # Face-Place shaped, Q = 576
    "c2": dict(N=100_000, p=64, q=16, L=256),   # Q = 1024 (the metric's 1-GPU configuration)
    "c3": dict(N=1_000_000, p=256, q=16, L=256),  # Q = 4096, row-sharded over 1/2/4/8 GPUs
}
