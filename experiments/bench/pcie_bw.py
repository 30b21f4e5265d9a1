"""Host <-> device copy bandwidth per rank when ALL ranks of the box copy at once (torchrun), from pinned buffers
allocated (a) wherever the process happens to run and (b) after binding the process to the CPUs NVML reports as local to
its GPU.  Answers why the end-to-end leg moves ~14 GB/s per GPU at 8 ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        experiments/bench/pcie_bw.py
"""
import json
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = 256
dbuf_in = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
dbuf_out = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def measure(tag):
    h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0); s2.wait_event(e0)
            for _ in range(4):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        dbuf_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(dbuf_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        nbytes = 4 * (MB << 20) * (2 if mode == "both" else 1)
        res[mode] = nbytes / ms / 1e6
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], device=dev, dtype=torch.float64)
    lo, hi = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"placement": tag, "ranks": world, "GBs_per_rank_min": dict(zip(("h2d", "d2h", "both_sum"), lo.tolist())),
                          "GBs_per_rank_max": dict(zip(("h2d", "d2h", "both_sum"), hi.tolist())),
                          "cpus_allowed": len(os.sched_getaffinity(0))}), flush=True)


measure("default")
from bench import _pin_to_local_numa  # noqa: E402
_pin_to_local_numa(local)
measure("bound to the GPU's local CPUs (NVML affinity)")
if world > 1:
    dist.destroy_process_group()
