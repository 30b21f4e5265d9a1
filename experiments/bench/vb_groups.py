"""Time of the full taylor_coeff (Xb, Vb, vbs, nll) at c3 minus the evaluation without Vb: the Vb product + B^-1, for the
column-group bound in the environment (GPP_PL_COL_GROUP_MB; 1000 = one group).  The column-group experiment itself was
dropped from the library after this script measured it (profiles/r02_vb_wave_alignment.txt): the variable is no longer read,
the script remains as the timing harness of the Vb product (GPP_TC_WAVE_SYNC=0 switches the producers' alignment off)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200.synth import CONFIGS, make_problem  # noqa: E402

cfg = CONFIGS["c3"]
dev = torch.device("cuda:0")
pr = make_problem(cfg["N"], cfg["p"], cfg["q"], cfg["L"], kind="trained", lvs=(0.0, 0.0), seed=0, device=dev)
vm = gppvae_b200.Vmodel(pr.x0.shape[0], cfg["q"], cfg["p"], cfg["q"]).to(dev)
gp = gppvae_b200.GP().to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0); gp.lvs.copy_(pr.lvs)


def run(need_vb):
    with torch.no_grad():
        V = vm(pr.d, pr.w)
        return gp.taylor_coeff(pr.Z, [V], need_vb=need_vb)


res = {}
for need_vb in (False, True):
    for _ in range(2):
        out = run(need_vb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        out = run(need_vb)
    e1.record()
    torch.cuda.synchronize()
    res[need_vb] = e0.elapsed_time(e1) / 4
    if need_vb:
        chk = float(out[1][0].double().abs().sum())
print(f"GPP_PL_COL_GROUP_MB={os.environ.get('GPP_PL_COL_GROUP_MB', 'default')}: without Vb {res[False]:.2f} ms, with Vb {res[True]:.2f} ms, "
      f"Vb + Binv {res[True] - res[False]:.2f} ms, sum|Vb| {chk:.6e}")
