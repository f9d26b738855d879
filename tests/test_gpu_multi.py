"""GPU, >= 2 devices: one process per GPU over NCCL -- striped transform, optional NCCL gather
and the fused peer-store gather (dist.PeerImage) on a small image, checked against the oracle
inside benchmarks/stripes.py.  Skipped on single-GPU boxes (the CPU suite covers the host
logic with gloo)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_two_rank_striped_transform_and_fused_gather(dtype):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577" if dtype == "u8" else "29578",
           os.path.join(ROOT, "benchmarks", "stripes.py"), "--rows", "2048", "--cols", "4096", "--dtype", dtype,
           "--steps", "5", "--warmup", "2", "--gather", "--fused-gather"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["parity_vs_oracle"] is True
    assert d["fused_gather_parity"] is True
    assert d["collective_on_data_path"] == "none"
