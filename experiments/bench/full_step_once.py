"""One full taylor_coeff (Xb, Vb, vbs, nll) at c3 through the public API: the target of ncu captures."""
import sys

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200.synth import CONFIGS, make_problem  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0")
pr = make_problem(cfg["N"], cfg["p"], cfg["q"], cfg["L"], kind="trained", lvs=(0.0, 0.0), seed=0, device=dev)
vm = gppvae_b200.Vmodel(pr.x0.shape[0], cfg["q"], cfg["p"], cfg["q"]).to(dev)
gp = gppvae_b200.GP().to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0); gp.lvs.copy_(pr.lvs)
    V = vm(pr.d, pr.w)
    Xb, Vbs, vbs, nll = gp.taylor_coeff(pr.Z, [V], need_vb=True)
torch.cuda.synchronize()
print("nll mean", float(nll.mean()), "Vb absmax", float(Vbs[0].abs().max()))
