"""CUDA-graph capture of one GP-term evaluation, for shapes whose step is bound by launch latency rather than by the
kernels (BASELINE.json configs[0] / [1]: ~55 launches of 5-30 us each -- DESIGN 5.3).

    step = CapturedGPTerm(vm, gp, d, w, Z)         # warm-up + capture
    Xb, _, vbs, nll = step()                       # replay with the captured inputs
    Xb, _, vbs, nll = step(Z=Z_new)                # copy new inputs into the static buffers, replay

The captured sequence is exactly `V = vm(d, w); gp.taylor_coeff(Z, [V], need_vb=False)` (train_gppvae.py:161,166): the
library is capture-safe (no host synchronisation and no allocation of its own on the path; tensor maps travel as kernel
parameters), the parameters `x0`, `v0`, `lvs` are read from their tensors at replay time, so optimiser steps between
replays are seen.  Returned tensors are the graph's static outputs: copy them if they must survive the next replay.
"""
from __future__ import annotations

from typing import Optional

import torch


class CapturedGPTerm:
    def __init__(self, vm, gp, d: torch.Tensor, w: torch.Tensor, Z: torch.Tensor, warmup: int = 2):
        if getattr(gp, "_sharded", False):
            raise NotImplementedError("CapturedGPTerm captures the single-GPU evaluation")
        self.vm, self.gp = vm, gp
        self.d, self.w, self.Z = d.clone(), w.clone(), Z.detach().clone()
        side = torch.cuda.Stream(device=Z.device)
        side.wait_stream(torch.cuda.current_stream(Z.device))
        with torch.cuda.stream(side):               # lazy initialisation (attributes, occupancy queries) outside the capture
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(Z.device).wait_stream(side)
        gp.invalidate_cache()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        with torch.no_grad():
            V = self.vm(self.d, self.w)
            return self.gp.taylor_coeff(self.Z, [V], need_vb=False)

    def __call__(self, d: Optional[torch.Tensor] = None, w: Optional[torch.Tensor] = None,
                 Z: Optional[torch.Tensor] = None):
        if d is not None:
            self.d.copy_(d)
        if w is not None:
            self.w.copy_(w)
        if Z is not None:
            self.Z.copy_(Z)
        self.graph.replay()
        return self.out
