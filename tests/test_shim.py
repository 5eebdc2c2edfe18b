"""CPU checks of the Fortran side of the boundary (no Fortran compiler exists in this image, so the shim is
checked as text): the generated iso_c_binding module declares EVERY entry point of include/swcuda.h with the
header's argument count, agrees with the ctypes prototypes the GPU tests call through, and the hand-written
shim (fortran/sw_interface_cuda.f90) only calls what is declared, with the right number of arguments."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fortran"))
import gen_bindings  # noqa: E402

BIND = os.path.join(ROOT, "fortran", "swcuda_c_binding.f90")
SHIM = os.path.join(ROOT, "fortran", "sw_interface_cuda.f90")


def fortran_interfaces():
    """{name: [dummy arguments]} of the generated module."""
    text = open(BIND).read().replace("&\n", " ")
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)", text):
        assert m.group(1) == m.group(3)
        out[m.group(1)] = [a.strip() for a in m.group(2).split(",") if a.strip()]
    return out


def test_generated_module_is_current():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "fortran", "gen_bindings.py"), "--check"]) == 0, \
        "run python fortran/gen_bindings.py"


def test_every_header_symbol_is_bound_with_the_headers_arguments():
    protos = {name: args for _, name, args in gen_bindings.prototypes(open(gen_bindings.HEADER).read())}
    f = fortran_interfaces()
    assert sorted(f) == sorted(protos)
    for name, args in protos.items():
        assert f[name] == [n for _, n in args], name
    # the same list the ctypes binding (and test_abi.py) works from
    from ocean_model_arch_b200 import _lib
    assert sorted(f) == sorted(_lib.EXPORTED_SYMBOLS)
    for name, sig in _lib._SIGNATURES.items():
        assert len(sig) == len(f[name]), name


def test_constants_match_the_header():
    text = open(BIND).read()
    hdr = open(gen_bindings.HEADER).read()
    consts = dict(re.findall(r"parameter :: (\w+) = (-?\d+)", text))
    from ocean_model_arch_b200 import _lib
    for name, fid in _lib.FIELD_ID.items():
        key = [k for k in consts if k.startswith("SWCU_F_") and k[7:].lower() == name.lower()]
        assert key and int(consts[key[0]]) == fid, name
    assert int(consts["SWCU_PEER_BLOB_BYTES"]) == int(re.search(r"#define SWCU_PEER_BLOB_BYTES (\d+)", hdr).group(1))
    assert int(consts["SWCU_K_TRACER_NEXT_STEP"]) == 14 and int(consts["SWCU_MODE_FUSED"]) == 1


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_shim_calls_only_declared_entry_points_with_matching_arity():
    f = fortran_interfaces()
    text = re.sub(r"!.*", "", open(SHIM).read()).replace("&\n", " ")
    assert "use swcuda_c_binding" in text and "module swcuda_c_binding" not in text
    calls = 0
    for m in re.finditer(r"\b(sw[ch]u?_[a-z0-9_]+)\s*\(", text):
        name = m.group(1)
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        args = split_args(text[m.end():i - 1])
        assert name in f, f"{name} is not part of the C ABI"
        assert len(args) == len(f[name]), (name, args, f[name])
        calls += 1
    assert calls >= 30
    # the shim picks its GPU by the node-local rank, and reports the library's own error text
    assert "mpi_comm_rank(node_comm, node_rank" in text and "mod(node_rank, ndev)" in text
    assert "swcu_last_error()" in text and "swcu_widen_halos" in text
