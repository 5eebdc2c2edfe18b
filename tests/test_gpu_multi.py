"""GPU parity with more than one GPU: N y-slabs (N = every GPU of the box, up to 8) with NCCL or peer-memory
halo exchange equal the one-block run and the oracle bitwise in the bitwise arithmetic, and reproduce the
one-block run bitwise (and the oracle within 1e-12) in tolerance mode (decomposition invariance).
Skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", [(150, 203, 40), (77, 64, 25), (150, 203, 40, "balance"), (150, 203, 40, "peer"),
                                   (300, 1030, 30)],
                         ids=["150x203", "77x64", "150x203-balanced-slabs", "150x203-peer-memory-halos-tracers", "300x1030"])
def test_slabs_over_nccl_bitwise(swlib, cuda_device, shape):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "run_multi_gpu.py"), *map(str, shape)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bitwise equal") == 3, r.stdout
    assert r.stdout.count("identical to 1 GPU") == 3, r.stdout      # tolerance mode: fused push, tracers, NCCL
    assert r.stdout.count("sync_test") == 2 and "FAILED" not in r.stdout, r.stdout


@pytest.mark.parametrize("layout", [(1, 2), (2, 2)])
def test_one_process_many_gpus_linked_blocks(swlib, cuda_device, layout):
    """The reference's _GPU_MULTI_ mode: ONE process drives blocks on several GPUs; halos are pulled over
    peer copies (swcu_link across devices).  Bitwise equal to the oracle in both modes."""
    import numpy as np
    import torch

    import basins
    from ocean_model_arch_b200 import model
    from ocean_model_arch_b200._lib import MODE_FUSED, MODE_REFERENCE
    from oracle_lib import OracleModel, make_config
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nx, ny = 150, 203
    mask = basins.island_mask(nx, ny)
    o = OracleModel(make_config(nx, ny, keep_mu=1, use_tracers=1), mask)
    o.step(30)
    for mode in (MODE_REFERENCE, MODE_FUSED):
        m = model.BlockGridModel(model.BasinPar(nx=nx, ny=ny), model.SwPar(use_tracers=1), bnx=layout[0], bny=layout[1],
                                 mask=mask, mode=mode, keep_mu=True, devices=tuple(range(min(n, 4))))
        m.step(30)
        assert m.synchronize() == 0
        for f in ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp", "ff1", "ff1p"):
            assert np.array_equal(m.get(f)[2:-2, 2:-2], o.get(f)[2:-2, 2:-2]), (f, mode, layout)
        m.close()
