"""Row-sharded N-rank run == single-rank run, through the CUDA kernels (peer exchange over NVLink inside the kernels, the fused
loop kernel, NCCL fallback for MOG / OMGP): spawns tools/check_multi_gpu.py under torchrun when the box has at least two GPUs.
(The host-side logic of the sharding is covered on CPU by tests/test_sharding_gloo.py with world_size 2.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_rank():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (this box has %d)" % n)
    world = 2 if n < 4 else 4 if n < 8 else 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "FAIL" not in res.stdout
