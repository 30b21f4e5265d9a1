"""Coherent (same-sign) accumulation bias of the tensor-core Gram / cross product: mean signed relative error of
V^T [V | Z] against float64 on all-positive data of several distributions.  Background for kDiagComp (gemm_tc.cu); see also smoke_diag.py."""
import os, sys
sys.path.insert(0, ".")
lib = os.environ.get("GPPVAE_LIB")
if lib:
    import gppvae_b200._lib as L
    L.LIB_PATH = lib
import torch
from gppvae_b200 import ops
dev = torch.device("cuda:0")
Q, Lz = 512, 256
for n in (20000, 150000):
    for name in ("absnormal*colscale", "uniform(0.5,1.5)", "lognormal", "exp(-u*8)", "mixed-sign normal"):
        torch.manual_seed(1)
        if name == "absnormal*colscale":
            V = (torch.randn(n, Q, device=dev) * torch.rand(1, Q, device=dev)).abs(); X = torch.randn(n, Lz, device=dev).abs()
        elif name == "uniform(0.5,1.5)":
            V = torch.rand(n, Q, device=dev) + 0.5; X = torch.rand(n, Lz, device=dev) + 0.5
        elif name == "lognormal":
            V = torch.randn(n, Q, device=dev).exp(); X = torch.randn(n, Lz, device=dev).exp()
        elif name == "exp(-u*8)":
            V = (-8 * torch.rand(n, Q, device=dev)).exp(); X = (-8 * torch.rand(n, Lz, device=dev)).exp()
        else:
            V = torch.randn(n, Q, device=dev); X = torch.randn(n, Lz, device=dev)
        ref = V.double().t() @ torch.cat([V.double(), X.double()], 1)
        GC = ops.gram_vtz(V, Q, X, Lz, n, Q, Lz).double()
        if name.startswith("mixed"):
            d = torch.diagonal(GC[:, :Q]); dr = torch.diagonal(ref[:, :Q])
            print(f"n={n:6d} {name:20s}: diag mean signed rel {((d - dr) / dr).mean().item():+.2e}   max-rel err {((GC - ref).abs().max() / ref.abs().max()).item():.2e}")
        else:
            rel = (GC - ref) / ref
            print(f"n={n:6d} {name:20s}: mean signed rel G {rel[:, :Q].mean().item():+.2e}  C {rel[:, Q:].mean().item():+.2e}  "
                  f"rms {rel.pow(2).mean().sqrt().item():.2e}  max {rel.abs().max().item():.2e}")
