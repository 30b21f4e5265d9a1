"""Timeline of gpp_gp_term_host (the C entry with host buffers) at c3, each call waiting for its results: every memcpy
and the large kernels with their start / end on the device (CUPTI through torch.profiler)."""
import ctypes
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from gppvae_b200 import _lib  # noqa: E402
from gppvae_b200.synth import CONFIGS, make_problem  # noqa: E402

cfg = CONFIGS["c3"]
dev = torch.device("cuda:0")
pr = make_problem(cfg["N"], cfg["p"], cfg["q"], cfg["L"], kind="trained", lvs=(0.0, 0.0), seed=0, device=dev)
n, L, p, q = cfg["N"], cfg["L"], cfg["p"], cfg["q"]
hd, hw, hZ = (t.cpu().pin_memory() for t in (pr.d, pr.w, pr.Z))
hx0, hv0, hl = (t.detach().cpu().contiguous().pin_memory() for t in (pr.x0, pr.v0, pr.lvs))
h_out = [(torch.empty(n, 1).pin_memory(), torch.empty(n, L).pin_memory(), torch.empty(2).pin_memory()) for _ in range(2)]
del pr
torch.cuda.empty_cache()
lib = _lib.load()
ctx = ctypes.c_void_p()
_lib.check(lib.gpp_host_ctx_create(ctypes.byref(ctx)), "host_ctx_create")


def call(i):
    _lib.check(lib.gpp_gp_term_host(ctx, hx0.data_ptr(), hx0.shape[0], p, hv0.data_ptr(), q, q, hd.data_ptr(), hw.data_ptr(),
                                    hZ.data_ptr(), n, L, hl.data_ptr(), h_out[i][0].data_ptr(), h_out[i][1].data_ptr(),
                                    h_out[i][2].data_ptr()), "gp_term_host")


for i in range(3):
    call(i & 1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        call(i & 1)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    dur = (e.time_range.end - e.time_range.start) / 1e3
    if dur > 0.25 or "emcpy" in e.name:
        print(f"{(e.time_range.start - t0) / 1e3:9.2f} ms  +{dur:8.2f} ms  {e.name[:90]}")
lib.gpp_host_ctx_destroy(ctx)
