"""GPU parity with more than one GPU: N y-slabs with NCCL halo exchange equal the one-block run and
the oracle bitwise (decomposition invariance).  Skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", [(150, 203, 40), (77, 64, 25)])
def test_slabs_over_nccl_bitwise(swlib, cuda_device, shape):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "run_multi_gpu.py"), *map(str, shape)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bitwise equal") == 3, r.stdout
    assert r.stdout.count("sync_test") == 2 and "FAILED" not in r.stdout, r.stdout
