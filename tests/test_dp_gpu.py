"""Multi-GPU (needs >= 2 CUDA devices): two ranks over NCCL, each running the fused step on its half of the global
batch, against the oracle on the concatenated batch (rank-ordered feature gather, labels rank*n+i, bucketed
gradient averaging).  Skipped on single-GPU boxes; the same logic is covered on CPU with gloo in test_host_cpu.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_ranks_equal_global_batch_oracle(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py"), precision]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[-1])
    tol = 1e-5 if precision == "fp32" else 2e-2
    assert abs(r["loss"] - r["oracle_loss"]) <= tol * abs(r["oracle_loss"]), r
    assert not r["failed"], r
    assert r["buckets"] >= 3
    assert r["replicas_bit_identical_after_4_steps"], r
