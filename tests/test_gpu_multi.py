"""GPU, >= 2 devices: batch sharding through the PRODUCT API over NCCL — `pipe(...)` under torchrun must return, on every
rank, the same latents a single process computes for the whole batch (skipped on a 1-GPU box; the host protocol is covered
on CPU by tests/test_parallel_gloo.py)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("batch", [5, 1])
def test_pipeline_shards_over_nccl(lib, tmp_path, batch):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "verdict.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tools" / "check_sharded_pipeline.py"), str(out), str(batch)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0,1"))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    v = json.loads(out.read_text())
    assert v["world"] == 2 and v["shape"][0] == batch and v["finite"]
    assert v["all_ranks_agree"] and v["equals_single_rank"] and v["generator_path_equals_single_rank"], v
