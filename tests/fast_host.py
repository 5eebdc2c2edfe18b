"""ctypes binding of tests/fast_host.cpp: the product's tolerance-mode formulas (csrc/sw_fast.cuh)
compiled as plain C++ and run over whole arrays on the CPU.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fast_host.cpp")
HDR = os.path.join(HERE, "..", "ocean_model_arch_b200", "csrc", "sw_fast.cuh")
LIB = os.path.join(HERE, "_build", "libfast_host.so")

MASK_BITS = (("lu", 1), ("lcu", 2), ("lcv", 4), ("luu", 8), ("luh", 16), ("llu", 32), ("llv", 64))
METRICS = ("dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s")
STATE = ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp")


def lib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        # -ffp-contract=off: only the explicit fma() calls of sw_fast.cuh fuse, as on the device (-fmad=false)
        subprocess.check_call(["g++", "-O2", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-std=c++17", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    L.swf_host_steps.restype = C.c_long
    return L


class FastHostModel:
    """Takes its inputs (masks, metrics, bathymetry, viscosity, state) from an OracleModel and steps them
    with the tolerance-mode formulas."""

    def __init__(self, oracle, cfg):
        self.L = lib()
        self.nx, self.ny = oracle.nx, oracle.ny
        self.cfg = cfg
        bits = np.zeros((self.ny, self.nx), np.uint8)
        for name, bit in MASK_BITS:
            bits |= np.where(oracle.get(name) > 0.5, bit, 0).astype(np.uint8)
        self.bits = np.ascontiguousarray(bits)
        self.metrics = [np.ascontiguousarray(oracle.get(n)) for n in METRICS]
        self.state = {n: np.ascontiguousarray(oracle.get(n)) for n in STATE}
        self.h_r = np.ascontiguousarray(oracle.get("hhq_rest"))
        self.mu = np.ascontiguousarray(oracle.get("mu"))
        rd = np.ascontiguousarray(oracle.get("r_diss"))
        self.rdis = rd if np.any(rd != 0) else None
        self.bad = 0
        if cfg.use_tracers > 0:
            self.state["ff1"] = np.ascontiguousarray(oracle.get("ff1"))
            self.state["ff1p"] = np.ascontiguousarray(oracle.get("ff1p"))

    def step(self, n=1):
        P = C.c_void_p
        ptr = lambda a: a.ctypes.data_as(P) if a is not None else P(None)
        mets = (P * 9)(*[m.ctypes.data for m in self.metrics])
        cfg = self.cfg
        s = self.state
        self.bad += self.L.swf_host_steps(
            C.c_int(self.nx), C.c_int(self.ny), C.c_int(n), C.c_double(float(cfg.time_step)), C.c_double(cfg.time_smooth),
            C.c_int(cfg.full_free_surface), C.c_int(cfg.trans_terms), C.c_int(cfg.ksw_lat), ptr(self.bits), mets,
            ptr(s["ssh"]), ptr(s["sshp"]), ptr(s["ubrtr"]), ptr(s["ubrtrp"]), ptr(s["vbrtr"]), ptr(s["vbrtrp"]),
            ptr(self.h_r), ptr(self.mu), P(None), P(None), ptr(self.rdis), ptr(s.get("ff1")), ptr(s.get("ff1p")))
        return self.bad

    def get(self, name):
        return self.state[name]
