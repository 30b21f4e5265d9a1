"""Row-sharded path on the REAL collective (NCCL over NVLink): two ranks launched with torchrun through bench.py,
whose `run_check` leg fails the run unless every rank's all-reduced G, C and W are bit-identical and sum(nll), the
rank's rows of Xb and vbs agree with the unsharded evaluation (train_gppvae.py has no
multi-GPU path of its own: the 1-GPU result is the reference).  Skipped where fewer than two GPUs are visible; the
host logic of the same path runs on CPU under gloo in tests/test_host_logic.py and tests/test_epoch.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload", ["c1", "c2"])
def test_two_rank_nccl_parity_through_bench_check(workload):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", workload, "--steps", "2",
           "--warmup", "3", "--skip-cpu", "--skip-full", "--skip-c-entry", "--skip-structured", "--check"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, res.stdout[-2000:]
    line = json.loads(lines[-1])
    chk = line["multi_gpu_check"]
    assert line["n_gpus"] == 2 and chk["ok"], chk
    for key in ("G_bit_identical_across_ranks", "C_bit_identical_across_ranks", "W_bit_identical_across_ranks"):
        assert chk[key] is True, chk
    assert chk["nll_sum_rel_vs_1gpu"] <= 1e-6 and chk["xb_max_rel_vs_1gpu"] <= 1e-5 and chk["vbs_max_rel_vs_1gpu"] <= 1e-5, chk
