"""A few structured-route evaluations at c3 (for the ncu launch list of that route)."""
import sys
sys.path.insert(0, ".")
import torch
import gppvae_b200
from gppvae_b200.synth import CONFIGS, make_problem
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0")
pr = make_problem(cfg["N"], cfg["p"], cfg["q"], cfg["L"], seed=0, device=dev)
vm = gppvae_b200.Vmodel(pr.x0.shape[0], cfg["q"], cfg["p"], cfg["q"]).to(dev)
gp = gppvae_b200.GP().to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0)
    for _ in range(3):
        gp._cache = type(gp._cache)()
        out = gp.taylor_coeff(pr.Z, [vm.lazy(pr.d, pr.w)], need_vb=False)
torch.cuda.synchronize()
print("ok", float(out[3].sum()))
