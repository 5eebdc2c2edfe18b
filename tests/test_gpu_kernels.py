"""GPU parity, Level A: every 1:1 CUDA kernel against the oracle's restatement of the same
reference kernel, on identical inputs, through the C ABI.  Bar: BITWISE equality (the CUDA side
is compiled with -fmad=false, the oracle with -ffp-contract=off)."""
import ctypes as C

import numpy as np
import pytest

import basins
from oracle_lib import OracleModel, call_kernel, make_config

pytestmark = pytest.mark.gpu

M4 = ["lu", "luu", "luh", "lcu", "lcv", "llu", "llv"]
G4 = ["dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb"]

# kernel -> (C-ABI symbol, oracle symbol, argument names in reference order, outputs)
# scalars are given as ("tau",) / ("ts",) / ("ffs",) / ("one",) placeholders
KERNELS = {
    "sw_update_ssh": ("swcu_sw_update_ssh_kernel", "sw_update_ssh_kernel",
                      ["$tau", "lu", "dx", "dy", "dxh", "dyh", "hhu", "hhv", "sshn", "sshp", "ubrtr", "vbrtr"],
                      ["sshn"]),
    "sw_update_uv": ("swcu_sw_update_uv", "sw_update_uv",
                     ["$tau", "lcu", "lcv", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "hhu", "hhu_n", "hhu_p",
                      "hhv", "hhv_n", "hhv_p", "hhh", "ssh", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp",
                      "r_diss", "rlh_s", "RHSx", "RHSy", "RHSx_adv", "RHSy_adv", "RHSx_dif", "RHSy_dif"],
                     ["ubrtrn", "vbrtrn"]),
    "sw_next_step": ("swcu_sw_next_step", "sw_next_step",
                     ["$ts", "lu", "lcu", "lcv", "ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp",
                      "vbrtr", "vbrtrn", "vbrtrp"],
                     ["ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp"]),
    "uv_trans_vort": ("swcu_uv_trans_vort_kernel", "uv_trans_vort_kernel",
                      ["luu", "dxt", "dyt", "dxb", "dyb", "ubrtr", "vbrtr", "vort"], ["vort"]),
    "uv_trans": ("swcu_uv_trans_kernel", "uv_trans_kernel",
                 ["lcu", "lcv", "luu", "dxh", "dyh", "ubrtr", "vbrtr", "vort", "hhq", "hhu", "hhv", "hhh",
                  "RHSx_adv", "RHSy_adv"], ["RHSx_adv", "RHSy_adv"]),
    "uv_diff2": ("swcu_uv_diff2_kernel", "uv_diff2_kernel",
                 ["lcu", "lcv"] + G4 + ["mu", "str_t", "str_s", "hhq", "hhu", "hhv", "hhh", "RHSx_dif", "RHSy_dif"],
                 ["RHSx_dif", "RHSy_dif"]),
    "stress_components": ("swcu_stress_components_kernel", "stress_components_kernel",
                          ["lu", "luu"] + G4 + ["ubrtrp", "vbrtrp", "str_t", "str_s"], ["str_t", "str_s"]),
    "hh_init": ("swcu_hh_init_kernel", "hh_init_kernel",
                ["$ffs", "lu", "llu", "llv", "luh"] + G4 +
                ["hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n", "hhh", "hhh_p", "hhh_n",
                 "ssh", "sshp", "hhq_rest"],
                ["hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n", "hhh", "hhh_p", "hhh_n"]),
    "hh_update": ("swcu_hh_update_kernel", "hh_update_kernel",
                  ["lu", "llu", "llv", "luh"] + G4 + ["hhq_n", "hhu_n", "hhv_n", "hhh_n", "ssh", "hhq_rest"],
                  ["hhq_n", "hhu_n", "hhv_n", "hhh_n"]),
    "hh_shift": ("swcu_hh_shift_kernel", "hh_shift_kernel",
                 ["$ts", "lu", "llu", "llv", "luh", "hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n",
                  "hhv", "hhv_p", "hhv_n", "hhh", "hhh_p", "hhh_n"],
                 ["hhq", "hhq_p", "hhu", "hhu_p", "hhv", "hhv_p", "hhh", "hhh_p"]),
    "tran_diff_fluxes": ("swcu_tran_diff_fluxes_kernel", "tran_diff_fluxes_kernel",
                         ["lcu", "lcv", "dxt", "dyt", "dxh", "dyh", "hhu", "hhv", "ff1", "ff1p", "ubrtr", "vbrtr", "mu",
                          "$one", "flux_x", "flux_y"], ["flux_x", "flux_y"]),
    "tran_diff_tracer": ("swcu_tran_diff_tracer_kernel", "tran_diff_tracer_kernel",
                         ["lu", "dx", "dy", "$tau", "hhq_n", "hhq_p", "flux_x", "flux_y", "ff1p", "ff1n"], ["ff1n"]),
    "tracer_next_step": ("swcu_tracer_next_step_kernel", "tracer_next_step_kernel",
                         ["$ts", "lu", "ff1n", "ff1p", "ff1"], ["ff1p", "ff1"]),
}
SCALARS = {"$tau": 1.5, "$ts": 0.5, "$one": 1.0, "$ffs": 1}


def realistic_state(nx, ny, mask, steps=5, seed=1):
    """All fields of the model after a few oracle steps, then every array a kernel only READS is
    perturbed so that no term of any formula is identically zero (mu, r_diss, RHSx/y, vort, ...)."""
    o = OracleModel(make_config(nx, ny, use_tracers=1, keep_mu=1, r_diss=5e-6), mask)
    o.step(steps)
    names8 = ["ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp", "RHSx", "RHSy",
              "RHSx_adv", "RHSy_adv", "RHSx_dif", "RHSy_dif", "mu", "str_t", "str_s", "vort", "hhq_rest", "hhq",
              "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n", "hhh", "hhh_p", "hhh_n",
              "flux_x", "flux_y", "ff1", "ff1n", "ff1p"]
    st = {n: o.get(n) for n in names8 + M4 + G4 + ["rlh_s", "r_diss"]}
    rng = np.random.default_rng(seed)
    sea = st["lu"] > 0.5
    for n in ("ubrtr", "ubrtrp", "ubrtrn", "vbrtr", "vbrtrp", "vbrtrn"):
        st[n] = st[n] + 1e-3 * rng.standard_normal(st[n].shape) * (st["lcu" if n[0] == "u" else "lcv"] > 0.5)
    st["RHSx"] = 1e-2 * rng.standard_normal(sea.shape)
    st["RHSy"] = 1e-2 * rng.standard_normal(sea.shape)
    st["mu"] = st["mu"] * (1.0 + 0.3 * rng.random(sea.shape))
    st["r_diss"] = (st["r_diss"] * (1.0 + rng.random(sea.shape))).astype(np.float32)
    st["hhq_rest"] = st["hhq_rest"] + 20.0 * rng.random(sea.shape)
    for n in ("vort", "str_t", "str_s", "flux_x", "flux_y"):
        st[n] = st[n] + 1e-4 * rng.standard_normal(sea.shape)
    return st, tuple(o.block_dims(0))


@pytest.fixture(scope="module")
def state():
    nx, ny = 100, 77   # odd sizes: rows are not vector-aligned, tiles are ragged
    return realistic_state(nx, ny, basins.island_mask(nx, ny))


@pytest.mark.parametrize("kname", sorted(KERNELS))
def test_level_a_kernel_bitwise(swlib, cuda_device, state, kname):
    import torch
    from ocean_model_arch_b200._lib import SwcuDims, check
    st, dims = state
    sym, osym, args, outs = KERNELS[kname]
    # oracle on host copies
    host = {a: st[a].copy() for a in args if not a.startswith("$")}
    oargs = [SCALARS[a] if a.startswith("$") else host[a] for a in args]
    call_kernel(osym, dims, *oargs)
    # CUDA on device copies, through the C ABI
    dev = {a: torch.from_numpy(st[a].copy()).to(cuda_device) for a in args if not a.startswith("$")}
    cargs = []
    for a in args:
        if a.startswith("$"):
            v = SCALARS[a]
            cargs.append(C.c_int(v) if isinstance(v, int) else C.c_double(v))
        else:
            cargs.append(C.c_void_p(dev[a].data_ptr()))
    d = SwcuDims(*dims)
    check(getattr(swlib, sym)(C.byref(d), *cargs, None))
    torch.cuda.synchronize()
    for o in outs:
        got = dev[o].cpu().numpy()
        assert np.array_equal(got, host[o]), (kname, o, float(np.abs(got - host[o]).max()))
    assert any(not np.array_equal(host[o], st[o]) for o in outs), (kname, "kernel changed nothing")
    # arrays that are not outputs must be untouched
    for a in dev:
        if a not in outs:
            assert np.array_equal(dev[a].cpu().numpy(), st[a]), (kname, a)


def test_check_ssh_err_counts(swlib, cuda_device, state):
    import torch
    from ocean_model_arch_b200._lib import SwcuDims, check
    st, dims = state
    ssh = st["ssh"].copy()
    sea = np.argwhere((st["lu"] > 0.5)[2:-2, 2:-2]) + 2
    ssh[tuple(sea[0])] = np.nan
    ssh[tuple(sea[5])] = 2.0e4
    ssh[tuple(sea[9])] = -1.0e4
    ssh[0, 0] = np.inf          # land / frame cell: ignored
    want = call_kernel("check_ssh_err_kernel", dims, st["lu"], ssh)
    assert want == 3
    lu = torch.from_numpy(st["lu"]).to(cuda_device)
    s = torch.from_numpy(ssh).to(cuda_device)
    bad = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    d = SwcuDims(*dims)
    check(swlib.swcu_check_ssh_err_kernel(C.byref(d), C.c_void_p(lu.data_ptr()), C.c_void_p(s.data_ptr()),
                                          C.c_void_p(bad.data_ptr()), None))
    torch.cuda.synchronize()
    assert int(bad.item()) == want


@pytest.mark.parametrize("seed", [1, 2024, 987654321])
def test_mdiv_is_ieee_division(swlib, cuda_device, seed):
    """The exact division of the fused kernels (q0 = a*y, two FMA residual corrections with y = RN(1/b),
    sign bit from q0) against the hardware IEEE division, bitwise, on 2^26 random operand pairs per
    seed: dividends over 2^-200..2^200 of both signs incl. +-0, divisors both promoted real(4) values and
    arbitrary doubles."""
    bad = C.c_long(-1)
    from ocean_model_arch_b200._lib import check
    check(swlib.swcu_selftest_mdiv(1 << 26, seed, C.byref(bad)))
    assert bad.value == 0
