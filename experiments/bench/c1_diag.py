"""Stage-by-stage error of the ill-conditioned c1 case (trained tables, lvs = (2, -4)) against float64, for a build variant:
GPPVAE_LIB=... python experiments/bench/c1_diag.py"""
import os, sys
sys.path.insert(0, ".")
lib = os.environ.get("GPPVAE_LIB")
if lib:
    import gppvae_b200._lib as L
    L.LIB_PATH = lib
import torch
import gppvae_b200
from gppvae_b200 import ops
from gppvae_b200.synth import make_problem
dev = torch.device("cuda:0")
def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
for lvs in ((0.0, 0.0), (2.0, -4.0)):
    pr = make_problem(4005, 64, 9, 256, kind="trained", lvs=lvs, seed=11)
    n, Q, L = 4005, 576, 256
    xn = ops.normalize_rows_fwd(pr.x0.to(dev)); wn = ops.normalize_rows_fwd(pr.v0.to(dev))
    V = ops.khatri_rao_fwd(xn, wn, pr.d.to(dev), pr.w.to(dev))
    Z = pr.Z.to(dev)
    V64, Z64 = V.double(), Z.double()
    vs = torch.exp(pr.lvs.double()).to(dev); vs = vs / vs.sum() if False else vs
    import oracle.gp_oracle as O
    o64 = O.taylor_coeff(pr.Z.double(), [V64.cpu()], pr.lvs.double())
    GC = ops.gram_vtz(V, Q, Z, L, n, Q, L)
    ref = V64.t() @ torch.cat([V64, Z64], 1)
    print(f"lvs={lvs}: GC err {rel(GC, ref):.2e} (G {rel(GC[:, :Q], ref[:, :Q]):.2e}, C {rel(GC[:, Q:], ref[:, Q:]):.2e})")
    gp = gppvae_b200.GP().to(dev)
    with torch.no_grad():
        gp.lvs.copy_(pr.lvs.to(dev))
    Xb, Vbs, vbs, nll = gp.taylor_coeff(Z, [V])
    print(f"   Xb err vs fp64 {rel(Xb.cpu(), o64[0]):.2e}   nll {abs(nll.double().sum().item() - o64[3].sum().item()) / abs(o64[3].sum().item()):.2e}")
    # exact-GC path: feed the fp64 Gram (rounded to fp32) to the Q-space stage and pass 2
    vsf = gp.get_vs().detach()
    for tag, gc in (("device GC", GC), ("fp64 GC  ", ref.float().contiguous())):
        fac = ops.factor(gc, Q + L, Q, vsf, False)
        W, scal = ops.solve_w(fac, gc[:, Q:], Q + L, L, L, n)
        r = (vsf[0] / vsf[1]).double()
        B = torch.eye(Q, device=dev, dtype=torch.float64) + r * ref[:, :Q]
        W64 = r * torch.linalg.solve(B, ref[:, Q:])
        Xb2, _ = ops.xb_nll(V, Q, Z, L, W, n, Q, L, scal)
        Xb64 = (Z64 - V64 @ W64) / vsf[1].double()
        XbW = (Z64 - V64 @ W.double()) / vsf[1].double()
        print(f"   [{tag}] W err {rel(W, W64):.2e}  Xb err {rel(Xb2, Xb64):.2e}  (pass 2 alone, given W: {rel(Xb2, XbW):.2e})  cond(B) {torch.linalg.cond(B).item():.1e}")
