"""The C ABI from a plain-C host: examples/sw_driver.c is compiled with the system C compiler against
include/swcuda.h + libswcuda.so (no Python, no torch in that process) and its output is checked
against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import OracleModel, make_config

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ocean_model_arch_b200")


def fnv1a(a):
    h = 14695981039346656037
    for byte in np.ascontiguousarray(a).tobytes():
        h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def driver(swlib, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("drv") / "sw_driver")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "cc"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "sw_driver.c"), "-L", PKG, "-lswcuda", "-lm",
                    f"-Wl,-rpath,{PKG}", "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("args", [("fused",), ("reference",), ("fused", "2", "3"), ("reference", "3", "2")])
def test_c_driver_matches_oracle(driver, cuda_device, args):
    nx, ny, steps = 97, 75, 40
    r = subprocess.run([driver, str(nx), str(ny), str(steps), *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    o = OracleModel(make_config(nx, ny))
    o.step(steps)
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())
    for f in ("ssh", "ubrtr", "vbrtr"):
        want = o.get(f)[2:-2, 2:-2]
        got = dict(kv.split("=") for kv in lines[f].split())
        assert int(got["fnv1a"], 16) == fnv1a(want), (f, args)
        assert float(got["max_abs"]) == float(np.abs(want).max())
    assert "launches=" in r.stdout
