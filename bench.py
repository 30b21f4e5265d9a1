#!/usr/bin/env python
"""Benchmark of the GP prior term (Woodbury NLL + dNLL/dZ) -- BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one evaluation of the GP term over the whole latent matrix, as train_gppvae.py:161-167 does
once per epoch: normalise tables -> Khatri-Rao map V -> pass 1 (V^T[V|Z]) -> all-reduce (N>1) ->
Cholesky/solve -> pass 2 + epilogue, producing nll (N x 1), Xb = dNLL/dZ (N x L) and vbs.  The workload is
BASELINE.json configs[2] (N=1M, L=256, Q=4096: the configuration the north-star target is quoted on; it
fits one B200), row-sharded over the ranks (strong scaling).  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's CPU algorithm for the same path (oracle/gp_oracle.py, an
op-for-op restatement with the same torch/MKL calls; the reference is pure Python and cannot travel to
the GPU box) on the host cores, on a bounded row sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gppvae_b200.synth import CONFIGS, make_problem  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of tc_pass1_kernel per launch, from profiles/ (ncu --set full, 1 GPU)
TRAFFIC = {("c3", 1): 42.348001e9 + 0.451665e9, ("c2", 1): 562.136576e6 + 65.065472e6}

METRIC = "gp_term_samples_per_s"
UNIT = "samples/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"],
                    bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_timing(cfg, budget_s: float = 20.0, rows: int | None = None):
    """Time the reference algorithm (oracle port, torch CPU, all host threads) on a bounded row sample.

    The reference's cost is a*N + b with b the N-independent Q x Q work (svd + inverse, gp.py:33-35).
    b is measured at 256 rows, a on the sample, and the throughput is extrapolated linearly to the full
    N of the workload: value = N / (a*N + b)."""
    from oracle import gp_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    N, p, q, L = cfg["N"], cfg["p"], cfg["q"], cfg["L"]
    Q = p * q

    def run(n):
        pr = make_problem(N, p, q, L, kind="trained", lvs=(0.0, 0.0), seed=0, n_rows=n)
        t0 = time.perf_counter()
        with torch.no_grad():
            V = O.feature_map(pr.x0, pr.v0, pr.d, pr.w)
            nll, Xb = O.nll_and_grad(pr.Z, [V], pr.lvs)
        dt = time.perf_counter() - t0
        return dt, float(nll.sum())

    run(min(256, N))                       # thread-pool / MKL warm-up
    tb, _ = run(min(256, N))               # ~ fixed cost b
    if rows is None:
        # calibrate the per-row cost on a small sample, then size the timed sample to ~budget_s of CPU work
        r_cal = min(N, 4096)
        tc, _ = run(r_cal)
        per_row = max(tc - tb, 1e-6) / max(r_cal - min(256, N), 1)
        rows = int(max(1024, min(N, (budget_s - min(tb, budget_s * 0.5)) / per_row)))
        rows = min(rows, 131072, N)   # bounds host memory: the reference keeps ~6 N x Q fp32 matrices alive
    ts, _ = run(rows)
    a = max(ts - tb, 1e-9) / max(rows - min(256, N), 1)
    full = a * N + tb
    return dict(value=N / full, unit=UNIT, cores=threads, kind="port",
                sample=f"{rows} of {N} rows (Q={Q}, L={L}) in {ts:.2f}s; Q-space cost b={tb:.2f}s measured at "
                       f"{min(256, N)} rows; extrapolated linearly in N: N/(a*N+b)",
                seconds_full_extrapolated=full, torch=torch.__version__)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.workload]
    vals = []
    total = args.warmup + args.steps
    t_start = time.time()
    last = None
    for i in range(total):
        if time.time() - t_start > 240 and i >= args.warmup + 1:   # keep the whole run within minutes
            break
        last = cpu_reference_timing(cfg, budget_s=8.0)
        if i >= args.warmup:
            vals.append(last["value"])
    v = statistics.median(vals) if vals else last["value"]
    N = cfg["N"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * N / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus),
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(name, gpus):
    c = CONFIGS[name]
    return {"workload": f"{name}: GP term N={c['N']} L={c['L']} Q={c['p'] * c['q']} (p={c['p']}, q={c['q']}), "
                        f"NLL + dNLL/dZ incl. Khatri-Rao map, rows sharded over {gpus} GPU(s)",
            "N": c["N"], "L": c["L"], "Q": c["p"] * c["q"], "p": c["p"], "q": c["q"], "tables": "trained-like",
            "lvs": [0.0, 0.0], "seed": 0, "parallelism": f"rows/{gpus}",
            "l2_policy": "inputs larger than L2 (V alone is N*Q*4 bytes per rank; no explicit flush)"}


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons = index, period, [], set()
        self._stop_evt = threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    import gppvae_b200
    from gppvae_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gppvae_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    cfg = CONFIGS[args.workload]
    N, p, q, L = cfg["N"], cfg["p"], cfg["q"], cfg["L"]
    Q = p * q
    per = -(-N // world)
    off = rank * per
    n = max(0, min(per, N - off))
    pr = make_problem(N, p, q, L, kind="trained", lvs=(0.0, 0.0), seed=0, device=dev, row_offset=off, n_rows=n)

    vm = gppvae_b200.Vmodel(pr.x0.shape[0], q, p, q).to(dev)
    gp = gppvae_b200.GP().to(dev)
    with torch.no_grad():
        vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0); gp.lvs.copy_(pr.lvs)
    if world > 1:
        gp.shard_rows()

    events = []

    def hook(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        events.append((name, ev))

    def step(d, w, Z):
        with torch.no_grad():
            if gp.stage_hook is not None:
                gp.stage_hook("kr:start")
            V = vm(d, w)
            return gp.taylor_coeff(Z, [V], need_vb=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_count() - l0

    # ---- device-resident throughput (inputs already in HBM)
    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        step(pr.d, pr.w, pr.Z)
    barrier()
    sampler.start()
    gp.stage_hook = hook
    ms, launches = timed(lambda: step(pr.d, pr.w, pr.Z), args.steps, 0)
    gp.stage_hook = None
    clocks = sampler.stop()
    stages = {}
    for (na, ea), (nb, eb) in zip(events[:-1], events[1:]):
        if nb.endswith(":end"):
            stages.setdefault(nb[:-4], []).append(ea.elapsed_time(eb))
        elif na == "kr:start" and nb == "pass1:start":
            stages.setdefault("khatri_rao", []).append(ea.elapsed_time(eb))
    stage_ms = {k: sum(v) / len(v) for k, v in stages.items()}

    # ---- full taylor_coeff (adds B^-1, Vb): reported, not the headline
    def full_step():
        with torch.no_grad():
            V = vm(pr.d, pr.w)
            return gp.taylor_coeff(pr.Z, [V], need_vb=True)
    ms_full = None
    if not args.skip_full:
        ms_full, _ = timed(full_step, max(1, min(args.steps, 3)), 1)
        gp._cache = type(gp._cache)()      # drop the cached V / Binv before the e2e leg
        torch.cuda.empty_cache()

    # ---- structured route: same inputs, same outputs, V never materialised.  Reported beside
    #      the dense path, which stays the headline (it is the path BASELINE.json's north_star names).
    structured = None
    if not args.skip_structured:
        def kr_step():
            with torch.no_grad():
                if gp.stage_hook is not None:
                    gp.stage_hook("kr:start")
                return gp.taylor_coeff(pr.Z, [vm.lazy(pr.d, pr.w)], need_vb=False)
        gp._cache = type(gp._cache)()
        torch.cuda.empty_cache()
        for _ in range(3):
            kr_step()
        events.clear()
        gp.stage_hook = hook
        ms_kr, launches_kr = timed(kr_step, args.steps, 0)
        gp.stage_hook = None
        st_kr = {}
        for (na, ea), (nb, eb) in zip(events[:-1], events[1:]):
            if nb.endswith(":end"):
                st_kr.setdefault(nb[:-4], []).append(ea.elapsed_time(eb))
            elif na == "kr:start" and nb == "pass1:start":
                st_kr.setdefault("index_and_tables", []).append(ea.elapsed_time(eb))
        Xb_s, _, vbs_s, nll_s = kr_step()
        Xb_d, _, vbs_d, nll_d = step(pr.d, pr.w, pr.Z)
        structured = {"ms_per_step": ms_kr, "value": N / (ms_kr * 1e-3), "gpu_launches": launches_kr,
                      "stage_ms": {k: sum(v) / len(v) for k, v in st_kr.items()},
                      "vs_dense_route": {"xb_max_rel": float((Xb_s - Xb_d).abs().max() / Xb_d.abs().max()),
                                         "nll_sum_rel": float(((nll_s.double().sum() - nll_d.double().sum()) /
                                                               nll_d.double().sum()).abs())},
                      "api": "GP.taylor_coeff(Z, [Vmodel.lazy(d, w)], need_vb=False)"}
        del Xb_s, Xb_d, nll_s, nll_d
        gp._cache = type(gp._cache)()
        torch.cuda.empty_cache()

    # ---- end to end: inputs start in pinned host memory, results end in host memory
    hd, hw, hZ = (t.cpu().pin_memory() for t in (pr.d, pr.w, pr.Z))
    h_nll = torch.empty(n, 1).pin_memory()
    h_Xb = torch.empty(n, L).pin_memory()
    h_vbs = torch.empty(2).pin_memory()

    def e2e_step():
        d = hd.to(dev, non_blocking=True); w = hw.to(dev, non_blocking=True); Z = hZ.to(dev, non_blocking=True)
        Xb, _, vbs, nll = step(d, w, Z)
        h_nll.copy_(nll, non_blocking=True); h_Xb.copy_(Xb, non_blocking=True); h_vbs.copy_(vbs, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e, _ = timed(e2e_step, args.steps, 1)
    h2d = hd.numel() * 8 + hw.numel() * 8 + hZ.numel() * 4
    d2h = h_nll.numel() * 4 + h_Xb.numel() * 4 + 8
    # checksum of the result over ALL ranks (comparable across --gpus: the synthetic rows do not depend on sharding)
    chk = torch.stack([h_nll.double().sum(), h_Xb.double().pow(2).sum()]).to(dev)
    if world > 1:
        dist.all_reduce(chk)
    nll_mean = float(chk[0]) / N
    xb_sq = float(chk[1])
    vbs_host = [float(v) for v in h_vbs]

    # ---- pure C entry with host buffers (1 GPU only): the same pipeline without Python between the calls
    ms_c = None
    if world == 1 and not args.skip_c_entry:
        import ctypes
        lib = _lib.load()
        ctx = ctypes.c_void_p()
        _lib.check(lib.gpp_host_ctx_create(ctypes.byref(ctx)), "host_ctx_create")
        hx0, hv0, hl = (t.detach().cpu().contiguous().pin_memory() for t in (pr.x0, pr.v0, pr.lvs))
        torch.cuda.synchronize()
        gp._cache = type(gp._cache)()
        torch.cuda.empty_cache()

        def c_step():
            _lib.check(lib.gpp_gp_term_host(ctx, hx0.data_ptr(), hx0.shape[0], p, hv0.data_ptr(), q, q, hd.data_ptr(),
                                            hw.data_ptr(), hZ.data_ptr(), n, L, hl.data_ptr(), h_nll.data_ptr(),
                                            h_Xb.data_ptr(), h_vbs.data_ptr()), "gp_term_host")
        c_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c_step()
        ms_c = 1e3 * (time.perf_counter() - t0) / args.steps
        lib.gpp_host_ctx_destroy(ctx)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    tf32_peak = pk["bf16_sustained"] / 2.0     # dense TF32 = half the dense bf16 rate; kernel timed inside a long step
    flops_pass1 = float(n) * Q * (Q + 1) + 2.0 * n * Q * L          # algorithmic: symmetric Gram + V^T Z
    t_pass1 = stage_ms.get("pass1", float("nan"))
    achieved = flops_pass1 / (t_pass1 * 1e-3) / 1e12 if t_pass1 == t_pass1 and t_pass1 > 0 else None
    flops_total = float(n) * (Q * (Q + 1) + 4.0 * Q * L + Q + 3 * L)
    bytes_total = float(n) * (12 * Q + 12 * L + 20)
    t_roof = max(flops_total / (tf32_peak * 1e12), bytes_total / (pk["hbm_gbs"] * 1e9)) * 1e3
    # executed tensor work: every product is issued as three fp16 MMAs (hi.hi, hi.lo, lo.hi; K = 16), each half the
    # tensor-pipe time of a TF32 one: 1.5 TF32-pass equivalents
    passes = 1.5
    # operand stream of pass 1: every 256 x 256 output tile streams its two 256-column operand slabs from L2 into shared
    # memory (TMA); measured ceiling of that stream with nothing consuming it: experiments/tc/exp2_tma_stream.cu
    t_q = -(-Q // 256)
    stream_bytes = float(n) * 512 * 4 * (t_q * (t_q + 1) / 2 + t_q * -(-L // 256))
    flops_gemm = float(n) * (Q * (Q + 1) + 4.0 * Q * L)
    t_roof_k3 = max(passes * flops_gemm / (tf32_peak * 1e12), bytes_total / (pk["hbm_gbs"] * 1e9)) * 1e3
    # DRAM traffic of the pass-1 kernel per launch from the committed `ncu --set full` captures (profiles/):
    # dram__bytes_read.sum + dram__bytes_write.sum at 1 GPU; None where no capture exists for the shape
    traffic = TRAFFIC.get((args.workload, world))
    line = {
        "metric": METRIC, "value": N / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, world),
        "clocks": clocks, "gpu_launches": launches, "engine": _lib.gemm_engine(),
        "e2e": {"value": N / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "api": "Vmodel.forward + GP.taylor_coeff(need_vb=False) from pinned host "
                                              "tensors; nll, Xb, vbs copied back to pinned host memory",
                "c_entry_ms_per_step": ms_c},
        "roofline": {"bound": "tensor", "kernel": "pass 1: V^T[V|Z] (tc_pass1_kernel + tc_reduce_kernel + tc_mirror_kernel)",
                     "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": (achieved / tf32_peak) if achieved else None, "traffic": traffic,
                     "traffic_source": "ncu --set full, profiles/r01_pass1_c3_f16_ncu_full.txt / r01_c2_final_ncu_full.txt"
                                       if traffic else None,
                     "algorithmic_bytes_per_launch": float(n) * (4 * Q + 4 * L) + 4.0 * Q * (Q + L),
                     "peak_source": f"{pk['source']} bf16_tflops_sustained / 2 (dense TF32), k=1 algorithmic flops",
                     "algorithmic_flops_per_launch": flops_pass1, "ms_per_launch": t_pass1,
                     "split_terms": "hi.hi, hi.lo and lo.hi as kind::f16 (K=16) with a common power-of-two scale: 1.5 "
                                    "TF32-pass equivalents",
                     "operand_stream": {"l2_to_smem_bytes_per_launch": stream_bytes,
                                        "achieved_tbs": (stream_bytes / (t_pass1 * 1e-3) / 1e12) if achieved else None,
                                        "ceiling_tbs": 6.45,
                                        "ceiling_source": "pure TMA stream of the same boxes at 1.9 GHz, "
                                                          "experiments/tc/exp2_tma_stream.cu (DESIGN 5.1)"},
                     "tf32_pass_equivalents": passes,
                     "executed_tflops_tf32_equivalent": passes * achieved if achieved else None,
                     "executed_frac": (passes * achieved / tf32_peak) if achieved else None,
                     "whole_step_frac_of_roofline": t_roof / ms, "whole_step_roofline_ms": t_roof,
                     "whole_step_frac_of_roofline_at_executed_passes": t_roof_k3 / ms,
                     "whole_step_roofline_ms_at_executed_passes": t_roof_k3},
        "stage_ms": stage_ms, "nll_mean": nll_mean, "xb_sumsq": xb_sq, "vbs": vbs_host,
        "full_taylor_coeff": None if ms_full is None else {"ms_per_step": ms_full, "value": N / (ms_full * 1e-3)},
        "structured_route": structured,
    }
    if world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpu_reference_timing(cfg, budget_s=20.0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--skip-full", action="store_true", help="omit the full taylor_coeff (Vb) leg")
    ap.add_argument("--skip-c-entry", action="store_true", help="omit the pure-C host-buffer leg")
    ap.add_argument("--skip-structured", action="store_true", help="omit the structured-route leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
