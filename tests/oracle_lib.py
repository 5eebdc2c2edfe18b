"""ctypes binding of the CPU oracle (oracle/libsw_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class SwoConfig(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int),
        ("dxst", C.c_double), ("dyst", C.c_double), ("rlon", C.c_double), ("rlat", C.c_double),
        ("curve_grid", C.c_int),
        ("rotation_on_lon", C.c_double), ("rotation_on_lat", C.c_double),
        ("full_free_surface", C.c_int), ("trans_terms", C.c_int), ("ksw_lat", C.c_int),
        ("time_smooth", C.c_double), ("lvisc_2", C.c_double),
        ("use_tracers", C.c_int), ("tracer_num", C.c_int),
        ("time_step", C.c_float),
        ("bnx", C.c_int), ("bny", C.c_int),
        ("keep_mu", C.c_int), ("r_diss", C.c_float), ("hhq_rest", C.c_double),
        ("nthreads", C.c_int),
    ]


F8 = ["ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp",
      "RHSx", "RHSy", "RHSx_adv", "RHSy_adv", "RHSx_dif", "RHSy_dif", "mu", "str_t", "str_s", "vort",
      "hhq_rest", "hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n",
      "hhh", "hhh_p", "hhh_n", "flux_x", "flux_y", "ff1", "ff1n", "ff1p"]
F4 = ["lu", "lu1", "luu", "luh", "lcu", "lcv", "llu", "llv",
      "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s", "r_diss"]

_lib = {}


def build(fast=False):
    target = "libsw_oracle_fast.so" if fast else "libsw_oracle.so"
    path = os.path.join(ORACLE_DIR, target)
    src = os.path.join(ORACLE_DIR, "sw_oracle.c")
    if (not os.path.exists(path)) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, target], stdout=subprocess.DEVNULL)
    return path


def lib(fast=False):
    if fast not in _lib:
        L = C.CDLL(build(fast))
        L.swo_create.restype = C.c_void_p
        L.swo_create.argtypes = [C.POINTER(SwoConfig), C.c_void_p]
        L.swo_destroy.argtypes = [C.c_void_p]
        L.swo_step.restype = C.c_long
        L.swo_step.argtypes = [C.c_void_p, C.c_int]
        L.swo_get_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.swo_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.swo_block_count.argtypes = [C.c_void_p]
        L.swo_block_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int * 8)]
        L.swo_block_field.restype = C.c_void_p
        L.swo_block_field.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int)]
        _lib[fast] = L
    return _lib[fast]


def make_config(nx, ny, *, dxst=0.00312, dyst=0.00225, rlon=34.75156, rlat=44.801125, curve_grid=1,
                full_free_surface=1, trans_terms=1, ksw_lat=1, time_smooth=0.5, lvisc_2=1.0e3,
                use_tracers=0, tracer_num=1, time_step=1.0, bnx=1, bny=1, keep_mu=0, r_diss=0.0,
                hhq_rest=100.0, nthreads=0):
    """Defaults = the shipped basin.par / sw.par / ocean_run.par values."""
    return SwoConfig(nx, ny, dxst, dyst, rlon, rlat, curve_grid, 0.0, 0.0,
                     full_free_surface, trans_terms, ksw_lat, time_smooth, lvisc_2,
                     use_tracers, tracer_num, time_step, bnx, bny, keep_mu, r_diss, hhq_rest, nthreads)


class OracleModel:
    def __init__(self, cfg, mask=None, fast=False):
        self.L = lib(fast)
        self.cfg = cfg
        self.nx, self.ny = cfg.nx, cfg.ny
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
            assert mask.shape == (self.ny, self.nx)
            mp = mask.ctypes.data_as(C.c_void_p)
        self.h = self.L.swo_create(C.byref(cfg), mp)

    def step(self, n=1):
        return self.L.swo_step(self.h, n)

    def get(self, name):
        if name in F8:
            out = np.empty((self.ny, self.nx), dtype=np.float64)
            rc = self.L.swo_get_field(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), None)
        else:
            out = np.empty((self.ny, self.nx), dtype=np.float32)
            rc = self.L.swo_get_field(self.h, name.encode(), None, out.ctypes.data_as(C.c_void_p))
        if rc:
            raise KeyError(name)
        return out

    def set(self, name, arr):
        if name in F8:
            a = np.ascontiguousarray(arr, dtype=np.float64)
            rc = self.L.swo_set_field(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), None)
        else:
            a = np.ascontiguousarray(arr, dtype=np.float32)
            rc = self.L.swo_set_field(self.h, name.encode(), None, a.ctypes.data_as(C.c_void_p))
        if rc:
            raise KeyError(name)

    def block_dims(self, k):
        d = (C.c_int * 8)()
        self.L.swo_block_dims(self.h, k, C.byref(d))
        return list(d)

    def close(self):
        if self.h:
            self.L.swo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def call_kernel(name, dims, *args, fast=False):
    """Calls oracle kernel `swo_<name>(8 ints, ...)`: numpy arrays go as pointers, Python floats as
    double, ints as int.  `dims` is an 8-tuple (nx_start, nx_end, ny_start, ny_end, bnd_x1..bnd_y2)."""
    fn = getattr(lib(fast), "swo_" + name)
    cargs = [C.c_int(int(v)) for v in dims]
    for a in args:
        if isinstance(a, np.ndarray):
            assert a.flags["C_CONTIGUOUS"]
            cargs.append(a.ctypes.data_as(C.c_void_p))
        elif isinstance(a, float):
            cargs.append(C.c_double(a))
        elif isinstance(a, (int, np.integer)):
            cargs.append(C.c_int(int(a)))
        else:
            raise TypeError(type(a))
    fn.restype = C.c_int
    return fn(*cargs)


def redo_hh_init(o):
    """The init-time envoke(hh_init) (control/init_data.f90:60-63) again, after a test overwrote hhq_rest or
    ssh of a one-block oracle model: refreshes the twelve depth arrays."""
    f4 = {n: o.get(n) for n in ("lu", "llu", "llv", "luh", "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb")}
    outs = {n: o.get(n) for n in ("hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n",
                                  "hhh", "hhh_p", "hhh_n")}
    call_kernel("hh_init_kernel", o.block_dims(0), int(o.cfg.full_free_surface), *[f4[n] for n in f4],
                *[outs[n] for n in outs], o.get("ssh"), o.get("sshp"), o.get("hhq_rest"))
    for n, arr in outs.items():
        o.set(n, arr)
