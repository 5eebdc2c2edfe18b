"""CPU tests of the drop-in boundary: libswcuda.so loads, exports every symbol include/swcuda.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "swcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sw[ch]u?_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(swlib):
    from ocean_model_arch_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    # and the Python binding covers each of them
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "swcuda.h"\nint main(void){swcu_dims d; (void)d; return SWCU_OK;}\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback(swlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ocean_model_arch_b200 import _lib
    assert swlib.swcu_device_count() == 0
    d = _lib.SwcuDims(3, 10, 3, 10, 1, 12, 1, 12)
    p = _lib.SwcuParams(1, 1, 1, 0.5, 0, _lib.MODE_FUSED)
    h = C.c_void_p()
    rc = swlib.swcu_create(C.byref(h), C.byref(d), C.byref(p), 0)
    assert rc == _lib.SWCU_ERR_CUDA
    assert b"CUDA" in swlib.swcu_last_error()
    with pytest.raises(_lib.SwcuError):
        _lib.check(rc)


def test_bad_dims_rejected(swlib):
    from ocean_model_arch_b200 import _lib
    d = _lib.SwcuDims(3, 10, 3, 10, 2, 12, 1, 12)  # border narrower than 2
    rc = swlib.swcu_sw_update_ssh_kernel(C.byref(d), 1.0, *([None] * 11), None)
    assert rc == _lib.SWCU_ERR_ARG


def test_product_does_not_touch_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "ocean_model_arch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sw_oracle" not in text and "oracle_lib" not in text and "swo_" not in text, f
    from ocean_model_arch_b200 import _lib
    out = subprocess.check_output(["nm", "-D", _lib.LIB_PATH], text=True)
    assert "swo_" not in out
