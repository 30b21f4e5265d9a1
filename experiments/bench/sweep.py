"""Rank / size sweep (BASELINE.json configs[4]): GP-term time (NLL + dNLL/dZ) for the dense and the structured route over
Q in {256 ... 16384} and N in {10k ... 4M} on one GPU.  q = 16 views, p = Q / 16, L = 256.  Every point: 2 warm-up + >= 10
timed iterations (CUDA events), the stage split of the dense step averaged over the timed iterations, the fraction of the
governing roofline and which bound governs.  The dense route needs V and its operand planes resident (8 N Q bytes): points
that do not fit one B200 run the structured route only.

    python experiments/bench/sweep.py [--quick]      # prints a markdown table, writes gpurun_out/sweep.json
"""
import json
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import gppvae_b200  # noqa: E402
from gppvae_b200.synth import make_problem  # noqa: E402

dev = torch.device("cuda:0")
quick = "--quick" in sys.argv
Qs = [256, 1024, 4096] if quick else [256, 512, 1024, 2048, 4096, 8192, 16384]
Ns = [10_000, 1_000_000] if quick else [10_000, 100_000, 1_000_000, 4_000_000]
points = None
if os.environ.get("SWEEP_POINTS"):      # e.g. SWEEP_POINTS=4000000x4096,1000000x16384: only these (N x Q) points
    points = {tuple(int(v) for v in pt.split("x")) for pt in os.environ["SWEEP_POINTS"].split(",")}
L, q = 256, 16
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
P_TF32 = peaks["bf16_tflops_sustained"] / 2 * 1e12
HBM = peaks["hbm_gbs"] * 1e9
rows = []


def timed(fn, gp=None, steps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = []
    if gp is not None:
        def hook(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))
        gp.stage_hook = hook
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if gp is not None:
            gp.stage_hook("kr:start")
        fn()
    e1.record()
    torch.cuda.synchronize()
    st = {}
    if gp is not None:
        gp.stage_hook = None
        for a, b in zip(ev[:-1], ev[1:]):
            if b[0].endswith(":end"):
                st.setdefault(b[0][:-4], []).append(a[1].elapsed_time(b[1]))
            elif a[0] == "kr:start" and b[0] == "pass1:start":
                st.setdefault("khatri_rao", []).append(a[1].elapsed_time(b[1]))
    return e0.elapsed_time(e1) / steps, {k: sum(v) / len(v) for k, v in st.items()}


for Q in Qs:
    p = Q // q
    for N in Ns:
        if points is not None and (N, Q) not in points:
            continue
        need_dense = 8.0 * N * Q + 12.0 * N * L + 4.0 * 4 * Q * Q + (1 << 30)
        free, total = torch.cuda.mem_get_info()
        row = dict(N=N, Q=Q, p=p)
        pr = gp = vm = None
        try:
            pr = make_problem(N, p, q, L, seed=0, device=dev)
            torch.cuda.empty_cache()          # the generator's chunk buffers: give them back before the big allocations
            free, total = torch.cuda.mem_get_info()
            vm = gppvae_b200.Vmodel(pr.x0.shape[0], q, p, q).to(dev)
            gp = gppvae_b200.GP().to(dev)
            with torch.no_grad():
                vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0)

            def dense():
                with torch.no_grad():
                    return gp.taylor_coeff(pr.Z, [vm(pr.d, pr.w)], need_vb=False)

            def structured():
                with torch.no_grad():
                    return gp.taylor_coeff(pr.Z, [vm.lazy(pr.d, pr.w)], need_vb=False)

            xd = None
            fits = need_dense < 0.9 * free
            if fits:
                try:
                    ms_d, st = timed(dense, gp)
                    xd = dense()[0]
                except torch.cuda.OutOfMemoryError:
                    fits = False
                    gp.stage_hook = None
                    gp.invalidate_cache()
                    torch.cuda.empty_cache()
            if fits:
                flops = N * (Q * (Q + 1.0) + 4.0 * Q * L) + Q ** 3 / 3.0 + 2.0 * Q * Q * L
                nbytes = N * (12.0 * Q + 12 * L + 20)
                t_tc, t_hbm = flops / P_TF32 * 1e3, nbytes / HBM * 1e3
                qspace = st.get("factor", 0) + st.get("solve", 0)
                bound = "q-space latency (Cholesky chain)" if qspace > 0.5 * ms_d else ("tensor" if t_tc > t_hbm else "hbm")
                row.update(dense_ms=ms_d, dense_msamples_s=N / ms_d / 1e3, stage_ms=st, roofline_ms=max(t_tc, t_hbm),
                           t_tensor_ms=t_tc, t_hbm_ms=t_hbm, frac_of_roofline=max(t_tc, t_hbm) / ms_d, bound=bound)
            elif need_dense < 0.9 * free:
                row.update(dense_ms=None, note=f"dense route ran out of memory ({need_dense / 1e9:.0f} GB for V + operand planes "
                                               "plus the generator's and the factorisation's buffers): structured only")
            else:
                row.update(dense_ms=None, note=f"dense route needs {need_dense / 1e9:.0f} GB (V + operand planes): structured only")
            gp.invalidate_cache()
            torch.cuda.empty_cache()
            ms_s, st_s = timed(structured, gp)
            sb = N * (12.0 * L + 36)
            row.update(structured_ms=ms_s, structured_msamples_s=N / ms_s / 1e3, structured_stage_ms=st_s,
                       structured_hbm_floor_ms=sb / HBM * 1e3)
            if xd is not None:
                xs = structured()[0]
                row["xb_struct_vs_dense"] = float((xs - xd).abs().max() / xd.abs().max())
        except torch.cuda.OutOfMemoryError:
            row["note"] = "out of memory"
        finally:
            rows.append(row)
            print(row, flush=True)
            del pr
            gp = vm = xd = None
            torch.cuda.empty_cache()
        json.dump(rows, open(os.environ.get("SWEEP_OUT", "gpurun_out/sweep.json"), "w"), indent=1)

print("\n| N | Q | dense ms (M samples/s) | KR / pass 1 / factor / solve / pass 2 ms | roofline ms (bound) | frac | structured ms (M samples/s) | Xb structured vs dense |")
print("|---|---|---|---|---|---|---|---|")
for r in rows:
    st = r.get("stage_ms") or {}
    d = "—" if r.get("dense_ms") is None else f"{r['dense_ms']:.2f} ({r['dense_msamples_s']:.1f})"
    s = "—" if r.get("structured_ms") is None else f"{r['structured_ms']:.2f} ({r['structured_msamples_s']:.1f})"
    stg = " / ".join(f"{st.get(k, float('nan')):.2f}" for k in ("khatri_rao", "pass1", "factor", "solve", "pass2")) if st else "—"
    rf = "—" if r.get("dense_ms") is None else f"{r['roofline_ms']:.2f} ({r['bound']})"
    fr = "—" if r.get("dense_ms") is None else f"{r['frac_of_roofline']:.2f}"
    e = "—" if r.get("xb_struct_vs_dense") is None else f"{r['xb_struct_vs_dense']:.1e}"
    print(f"| {r['N']:,} | {r['Q']} | {d} | {stg} | {rf} | {fr} | {s} | {e} |")
