"""Where the end-to-end step's margin over the device-resident step goes: the pipelined leg of bench.py re-enacted at c3
with CUDA events at every boundary of the three streams (main: compute, s_in: host -> device, s_out: device -> host)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200.synth import CONFIGS, make_problem  # noqa: E402

cfg = CONFIGS["c3"]
dev = torch.device("cuda:0")
pr = make_problem(cfg["N"], cfg["p"], cfg["q"], cfg["L"], kind="trained", lvs=(0.0, 0.0), seed=0, device=dev)
n, L = cfg["N"], cfg["L"]
vm = gppvae_b200.Vmodel(pr.x0.shape[0], cfg["q"], cfg["p"], cfg["q"]).to(dev)
gp = gppvae_b200.GP().to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0); gp.lvs.copy_(pr.lvs)
hd, hw, hZ = (t.cpu().pin_memory() for t in (pr.d, pr.w, pr.Z))
h_out = [(torch.empty(n, 1).pin_memory(), torch.empty(n, L).pin_memory(), torch.empty(2).pin_memory()) for _ in range(2)]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
Zd = [torch.empty_like(pr.Z) for _ in range(2)]
main = torch.cuda.current_stream()
mode = sys.argv[1] if len(sys.argv) > 1 else "split"


def ev(stream=None):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream if stream is not None else main)
    return e


def device_step():
    with torch.no_grad():
        V = vm(pr.d, pr.w)
        return gp.taylor_coeff(pr.Z, [V], need_vb=False)


for _ in range(3):
    device_step()
torch.cuda.synchronize()
a = ev()
for _ in range(5):
    device_step()
b = ev()
torch.cuda.synchronize()
print(f"device-resident step: {a.elapsed_time(b) / 5:.2f} ms")

out_done = [None, None]
z_free = [None, None]
log = []


def step(i, d2h=True, h2d=True):
    k = i & 1
    marks = {}
    marks["t0"] = ev()
    if h2d:
        with torch.cuda.stream(s_in):
            if z_free[k] is not None:
                s_in.wait_event(z_free[k])
            marks["h2d0"] = ev(s_in)
            Zd[k].copy_(hZ, non_blocking=True)
            marks["h2d1"] = ev(s_in)
    d = hd.to(dev, non_blocking=True); w = hw.to(dev, non_blocking=True)
    marks["dw"] = ev()
    with torch.no_grad():
        V = vm(d, w)
        marks["kr"] = ev()
        if mode == "split":
            gp.U_UBi_Shb([V], gp.get_vs())
            marks["gram_factor"] = ev()
        if h2d:
            main.wait_event(marks["h2d1"])
        marks["zwait"] = ev()
        Xb, _, vbs, nll = gp.taylor_coeff(Zd[k], [V], need_vb=False)
    marks["done"] = ev()
    z_free[k] = marks["done"]
    if d2h:
        if out_done[k] is not None:
            out_done[k].synchronize()
        with torch.cuda.stream(s_out):
            s_out.wait_event(marks["done"])
            for t in (Xb, vbs, nll):
                t.record_stream(s_out)
            marks["d2h0"] = ev(s_out)
            h_out[k][0].copy_(nll, non_blocking=True); h_out[k][1].copy_(Xb, non_blocking=True)
            h_out[k][2].copy_(vbs, non_blocking=True)
            marks["d2h1"] = ev(s_out)
        out_done[k] = marks["d2h1"]
    log.append(marks)


for variant, kw in (("both copies", {}), ("no device->host", {"d2h": False}), ("no host->device", {"h2d": False}),
                    ("no copies", {"d2h": False, "h2d": False})):
    log.clear()
    out_done[:] = [None, None]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(8):
        step(i, **kw)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / 8
    base = log[3]["t0"]
    print(f"== {mode}, {variant}: wall {wall:.2f} ms/step; steps 3..5 relative to step 3's start (ms):")
    for m in log[3:6]:
        print("   " + "  ".join(f"{k}={base.elapsed_time(v):7.2f}" for k, v in m.items()))
