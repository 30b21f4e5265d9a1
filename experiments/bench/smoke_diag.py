"""NLL / Xb error of the smoke-sized problem (N=1536, Q=128, L=64) and neighbours against the float64 oracle, for a build
variant: GPPVAE_LIB=... python experiments/bench/smoke_diag.py"""
import os, sys
sys.path.insert(0, ".")
lib = os.environ.get("GPPVAE_LIB")
if lib:
    import gppvae_b200._lib as L
    L.LIB_PATH = lib
import torch
import gppvae_b200
from gppvae_b200.synth import make_problem
from oracle import gp_oracle as O
dev = torch.device("cuda:0")
tag = os.path.basename(lib or "default")
out = []
for (n, p, q, Lz, lvs, seed) in [(1536, 16, 8, 64, (0.4, -0.6), 1), (1536, 16, 8, 64, (0.4, -0.6), 2), (1536, 16, 8, 64, (0.4, -0.6), 3),
                                 (4000, 32, 8, 64, (0.0, 0.0), 4), (4000, 32, 8, 128, (1.0, -2.0), 5), (20000, 64, 8, 256, (0.4, -0.6), 6)]:
    pr = make_problem(n, p, q, Lz, kind="trained", lvs=lvs, seed=seed)
    V64 = O.feature_map(pr.x0.double(), pr.v0.double(), pr.d, pr.w)
    oXb, _, _, onll = O.taylor_coeff(pr.Z.double(), [V64], pr.lvs.double())
    vm = gppvae_b200.Vmodel(pr.x0.shape[0], q, p, q).to(dev)
    gp = gppvae_b200.GP().to(dev)
    with torch.no_grad():
        vm.x0.copy_(pr.x0.to(dev)); vm.v0.copy_(pr.v0.to(dev)); gp.lvs.copy_(pr.lvs.to(dev))
    d, w, Z = pr.d.to(dev), pr.w.to(dev), pr.Z.to(dev)
    V = vm(d, w).detach()
    Xb, _, _, nll = gp.taylor_coeff(Z, [V], need_vb=False)
    e = (nll.double().sum().item() - onll.sum().item()) / abs(onll.sum().item())
    exb = float((Xb.double().cpu() - oXb).abs().max() / oXb.abs().max())
    amp = float((pr.Z.double() ** 2).sum() / (oXb * pr.Z.double()).sum().abs() )
    out.append(f"n={n} Q={p*q} L={Lz} lvs={lvs}: nll {e:+.2e} Xb {exb:.1e} (|Z|^2/quad*vn~{amp:.0f})")
print(tag + "\n  " + "\n  ".join(out))
