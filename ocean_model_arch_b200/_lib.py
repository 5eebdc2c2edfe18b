"""ctypes binding of libswcuda.so (include/swcuda.h).

The library is the product; there is no Python or CPU fallback.  If the shared object is missing
the import raises, and every compute entry point fails with SWCU_ERR_CUDA when no GPU is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libswcuda.so")

SWCU_OK, SWCU_ERR_CUDA, SWCU_ERR_ARG, SWCU_ERR_NCCL, SWCU_ERR_STATE, SWCU_ERR_BLOWUP = range(6)
MODE_REFERENCE, MODE_FUSED = 0, 1

F8_NAMES = ["ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp",
            "RHSx", "RHSy", "RHSx_adv", "RHSy_adv", "RHSx_dif", "RHSy_dif", "mu", "str_t", "str_s", "vort",
            "hhq_rest", "hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n",
            "hhh", "hhh_p", "hhh_n", "flux_x", "flux_y", "ff1", "ff1n", "ff1p"]
F4_NAMES = ["lu", "luu", "luh", "lcu", "lcv", "llu", "llv",
            "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s", "r_diss"]
FIELD_ID = {n: i for i, n in enumerate(F8_NAMES)}
FIELD_ID.update({n: 100 + i for i, n in enumerate(F4_NAMES)})
MASK_NAMES = F4_NAMES[:7]
KERNEL_ID = {n: i + 1 for i, n in enumerate([
    "sw_update_ssh", "hh_update", "uv_trans_vort", "uv_trans", "stress_components", "uv_diff2", "sw_update_uv",
    "sw_next_step", "hh_shift", "hh_init", "check_ssh_err", "tran_diff_fluxes", "tran_diff_tracer", "tracer_next_step"])}


class SwcuDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("nx_start", "nx_end", "ny_start", "ny_end", "bnd_x1", "bnd_x2", "bnd_y1", "bnd_y2")]

    @property
    def shape(self):
        return (self.bnd_y2 - self.bnd_y1 + 1, self.bnd_x2 - self.bnd_x1 + 1)

    def as_tuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class SwcuParams(C.Structure):
    _fields_ = [("full_free_surface", C.c_int), ("trans_terms", C.c_int), ("ksw_lat", C.c_int),
                ("time_smooth", C.c_double), ("use_tracers", C.c_int), ("mode", C.c_int)]


class SwhBasin(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int),
                ("dxst", C.c_double), ("dyst", C.c_double), ("rlon", C.c_double), ("rlat", C.c_double),
                ("curve_grid", C.c_int), ("rotation_on_lon", C.c_double), ("rotation_on_lat", C.c_double)]


class SwcuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"swcuda error {code}: {msg}")
        self.code = code


_lib = None

_P, _D, _I, _V = C.c_void_p, C.c_double, C.c_int, C.c_void_p
_DIMS = C.POINTER(SwcuDims)

# name -> argtypes (restype is always int unless listed in _RESTYPES)
PEER_BLOB_BYTES = 2048   # SWCU_PEER_BLOB_BYTES

_SIGNATURES = {
    "swcu_sw_update_ssh_kernel": [_DIMS, _D] + [_P] * 11 + [_V],
    "swcu_sw_update_uv": [_DIMS, _D] + [_P] * 30 + [_V],
    "swcu_sw_next_step": [_DIMS, _D] + [_P] * 12 + [_V],
    "swcu_uv_trans_vort_kernel": [_DIMS] + [_P] * 8 + [_V],
    "swcu_uv_trans_kernel": [_DIMS] + [_P] * 14 + [_V],
    "swcu_uv_diff2_kernel": [_DIMS] + [_P] * 19 + [_V],
    "swcu_stress_components_kernel": [_DIMS] + [_P] * 14 + [_V],
    "swcu_hh_init_kernel": [_DIMS, _I] + [_P] * 27 + [_V],
    "swcu_hh_update_kernel": [_DIMS] + [_P] * 18 + [_V],
    "swcu_hh_shift_kernel": [_DIMS, _D] + [_P] * 16 + [_V],
    "swcu_check_ssh_err_kernel": [_DIMS, _P, _P, _P, _V],
    "swcu_tran_diff_fluxes_kernel": [_DIMS] + [_P] * 13 + [_D, _P, _P, _V],
    "swcu_tran_diff_tracer_kernel": [_DIMS, _P, _P, _P, _D] + [_P] * 6 + [_V],
    "swcu_tracer_next_step_kernel": [_DIMS, _D] + [_P] * 4 + [_V],
    "swcu_create": [C.POINTER(_P), _DIMS, C.POINTER(SwcuParams), _I],
    "swcu_destroy": [_P],
    "swcu_upload": [_P, _I, _P],
    "swcu_upload_rows": [_P, _I, _P, _I, _I],
    "swcu_download": [_P, _I, _P],
    "swcu_output_record": [_P, _I, _P],
    "swcu_upload_from_device": [_P, _I, _P],
    "swcu_download_to_device": [_P, _I, _P],
    "swcu_set_option": [_P, C.c_char_p, _I],
    "swcu_uses_metric_tables": [_P],
    "swcu_envoke_hh_init": [_P],
    "swcu_envoke_kernel": [_P, _I, _D],
    "swcu_envoke_sync": [_P, _I],
    "swcu_step": [_P, _D, _I],
    "swcu_synchronize": [_P, C.POINTER(C.c_long)],
    "swcu_timer_start": [_P],
    "swcu_timer_stop": [_P, C.POINTER(C.c_float)],
    "swcu_profile_steps": [_P, _D, _I, C.POINTER(C.c_float), C.POINTER(C.c_long), C.POINTER(C.c_float), C.POINTER(C.c_long)],
    "swcu_selftest_mdiv": [C.c_long, C.c_ulonglong, C.POINTER(C.c_long)],
    "swcu_launch_count": [_P],
    "swcu_device_bytes": [_P],
    "swcu_stream": [_P],
    "swcu_comm_unique_id": [_P],
    "swcu_comm_init": [_P, _I, _I, _P],
    "swcu_comm_destroy": [_P],
    "swcu_halo_plan": [_DIMS, _I, _I, C.POINTER(_I), C.POINTER(_I)],
    "swcu_halo_exchange": [_P, _I],
    "swh_balanced_slabs": [_I, _I, _P, _I, _I, _I, _D, _P, _P],
    "swh_block_weights": [_I, _I, _I, _I, _P, _P],
    "swh_hilbert_d2xy": [_I, _I, C.POINTER(_I), C.POINTER(_I)],
    "swh_hilbert_partition": [_I, _P, _I, _P, _P],
    "swh_uniform_partition": [_I, _I, _I, _I, _P, _P],
    "swcu_init_grid": [_P, _P, _P],
    "swcu_fill": [_P, _I, _D],
    "swcu_copy_field": [_P, _I, _I],
    "swcu_peer_export": [_P, _P],
    "swcu_peer_attach": [_P, _I, _P],
    "swcu_peer_detach": [_P],
    "swcu_link": [_P, _P],
    "swcu_unlink": [_P],
    "swcu_step_group": [C.POINTER(_P), _I, _D, _I],
    "swcu_last_error": [],
    "swcu_version": [],
    "swcu_device_count": [],
    "swcu_widen_halos": [_P],
    "swcu_march_band_rows": [_I, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I)],
    "swh_masks": [C.POINTER(SwhBasin), _DIMS] + [_P] * 8,
    "swh_metrics": [C.POINTER(SwhBasin), _DIMS] + [_P] * 9,
    "swh_gaussian": [_DIMS, _P, _P, _D, _I, _I],
    "swh_uniform_split": [_I, _I, _I, C.POINTER(_I), C.POINTER(_I)],
}
_RESTYPES = {"swcu_last_error": C.c_char_p, "swcu_launch_count": C.c_long, "swcu_device_bytes": C.c_long,
             "swcu_stream": C.c_void_p}
EXPORTED_SYMBOLS = sorted(_SIGNATURES)


def lib():
    """Loads libswcuda.so; raises if it has not been built (python -m ocean_model_arch_b200.build)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SWCU_LIB", LIB_PATH)   # tuning builds (ocean_model_arch_b200.build.build(lib=...))
        if not os.path.exists(path):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m ocean_model_arch_b200.build` "
                              "(there is no CPU fallback)")
        L = C.CDLL(path)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(rc):
    if rc != SWCU_OK:
        raise SwcuError(rc, lib().swcu_last_error().decode())
    return rc
