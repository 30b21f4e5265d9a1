#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.csv> "<title>"
    python profiles/summarize.py full <report.ncu-rep> <out.txt>
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# {title}", "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
           "kernel,launches,total_us,share"]
    out += [f"{k},{c},{v:.1f},{v / tot:.3f}" for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        out.append("kernel: " + row[hdr.index("Kernel Name")][:100])
        for k in hdr:
            if any(k == key or k.startswith(key) for key in KEYS) or "tensor" in k and "pct" in k:
                i = hdr.index(k)
                out.append(f"  {k} = {row[i]} {units[i]}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
