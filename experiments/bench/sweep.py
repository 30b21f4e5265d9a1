"""Rank / size sweep (BASELINE.json configs[4]): GP-term time (NLL + dNLL/dZ) for the dense and the structured route
over Q and N on one GPU, with the stage split of the dense step.  q = 16 views, p = Q / 16, L = 256.

    python experiments/bench/sweep.py [--full]      # prints a markdown table, writes gpurun_out/sweep.json
"""
import json, sys
sys.path.insert(0, ".")
import torch
import gppvae_b200
from gppvae_b200.synth import make_problem

dev = torch.device("cuda:0")
full = "--full" in sys.argv
Qs = [256, 1024, 4096, 8192] if not full else [256, 512, 1024, 2048, 4096, 8192, 16384]
Ns = [10_000, 100_000, 1_000_000] if not full else [10_000, 100_000, 1_000_000, 4_000_000]
L, q = 256, 16
rows = []


def timed(fn, steps=3):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for Q in Qs:
    p = Q // q
    for N in Ns:
        dense_bytes = N * Q * 4
        free, _ = torch.cuda.mem_get_info()
        pr = None
        try:
            pr = make_problem(N, p, q, L, seed=0, device=dev)
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

            ms_d = None
            if dense_bytes * 1.3 + 3 * Q * Q * 4 * 2 < free * 0.8:
                ev = []
                gp.stage_hook = lambda name: (ev.append((name, torch.cuda.Event(enable_timing=True))), ev[-1][1].record())
                dense(); torch.cuda.synchronize()
                st = {b[0][:-4]: a[1].elapsed_time(b[1]) for a, b in zip(ev[:-1], ev[1:]) if b[0].endswith(":end")}
                gp.stage_hook = None
                ms_d = timed(dense)
                xd = dense()[0]
            else:
                st, xd = {}, None
            gp._cache = type(gp._cache)()
            ms_s = timed(structured)
            err = None
            if xd is not None:
                xs = structured()[0]
                err = float((xs - xd).abs().max() / xd.abs().max())
            rows.append(dict(N=N, Q=Q, p=p, dense_ms=ms_d, structured_ms=ms_s, stage_ms=st, xb_struct_vs_dense=err))
            print(rows[-1], flush=True)
        except torch.cuda.OutOfMemoryError:
            rows.append(dict(N=N, Q=Q, p=p, dense_ms=None, structured_ms=None, note="out of memory"))
            print(rows[-1], flush=True)
        finally:
            del pr
            gp = vm = None
            torch.cuda.empty_cache()

json.dump(rows, open("gpurun_out/sweep.json", "w"), indent=1)
print("\n| N | Q | dense ms (M samples/s) | pass 1 / factor / pass 2 ms | structured ms (M samples/s) | Xb structured vs dense |")
print("|---|---|---|---|---|---|")
for r in rows:
    d, s, st = r.get("dense_ms"), r.get("structured_ms"), r.get("stage_ms") or {}
    f = lambda ms: "—" if ms is None else f"{ms:.2f} ({r['N'] / ms / 1e3:.1f})"
    split = " / ".join(f"{st.get(k, float('nan')):.2f}" for k in ("pass1", "factor", "pass2")) if st else "—"
    e = r.get("xb_struct_vs_dense")
    print(f"| {r['N']:,} | {r['Q']} | {f(d)} | {split} | {f(s)} | {'—' if e is None else f'{e:.1e}'} |")
