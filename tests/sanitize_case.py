"""Small all-modes run for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tests/sanitize_case.py
Covers the tiled kernel with ragged tiles, the two-launch path, REFERENCE mode and tracers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import basins  # noqa: E402
from ocean_model_arch_b200 import model  # noqa: E402
from ocean_model_arch_b200._lib import MODE_FUSED, MODE_REFERENCE  # noqa: E402

nx, ny = 77, 45
mask = basins.island_mask(nx, ny, ndisc=2)
ref = None
for mode, tiled, tracers in ((MODE_FUSED, 1, 0), (MODE_FUSED, 0, 0), (MODE_REFERENCE, 0, 0), (MODE_FUSED, 1, 1)):
    m = model.ShallowWaterModel(model.BasinPar(nx=nx, ny=ny), model.SwPar(use_tracers=tracers), mask=mask, mode=mode,
                                keep_mu=True)
    if mode == MODE_FUSED:
        m.block.set_option("tiled", tiled)
    m.step(5)
    assert m.block.synchronize() == 0
    ssh = m.get("ssh")
    hhu = m.get("hhu")
    if ref is None:
        ref = ssh
    assert np.array_equal(ssh, ref)
    m.block.close()
print("sanitize_case: ok")
