"""Host-side mirror of the reference's algorithm layer for the shallow-water path.

Names follow the reference: the four .par files (configs/*.f90), `domain` geometry
(core/decomposition.f90), `init_grid_data` / `init_ocean_data` (control/init_data.f90) and
`expl_shallow_water(tau)` (control/shallow_water/shallow_water.f90:22).  All numerics happen in
libswcuda.so (CUDA kernels; C++ host init); this file only moves arrays and calls the C ABI.
"""
import ctypes as C
import dataclasses
import os

import numpy as np

from . import _lib
from ._lib import FIELD_ID, F4_NAMES, F8_NAMES, MODE_FUSED, MODE_REFERENCE, SwcuDims, SwcuParams, SwhBasin, check


# ----------------------------------------------------------------------------- configs (.par)
def read_par(path):
    """legacy/service/read_write_parameters.f90:7-42 (readpar): one value per line, text after the
    first ':' is a comment."""
    vals = []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            vals.append(line.split(":", 1)[0].strip())
    return vals


def _f(s):
    return float(s.lower().replace("d", "e"))


@dataclasses.dataclass
class BasinPar:  # configs/basinpar.f90:64-83 ; defaults = shipped basin.par
    nx: int = 1525
    ny: int = 1115
    dxst: float = 0.00312
    dyst: float = 0.00225
    rlon: float = 34.751560
    rlat: float = 44.801125
    curve_grid: int = 1
    rotation_on_lon: float = 0.0
    rotation_on_lat: float = 0.0
    mask_file_name: str = "none"
    bottom_topography_file_name: str = "none"

    @classmethod
    def from_file(cls, path):
        v = read_par(path)
        if int(v[9]) != 0 or int(v[10]) != 0:
            raise ValueError("only regular grids (xgr_type = ygr_type = 0) are supported")
        return cls(nx=int(v[0]), ny=int(v[1]), dxst=_f(v[5]), dyst=_f(v[6]), rlon=_f(v[7]), rlat=_f(v[8]),
                   curve_grid=int(v[11]), rotation_on_lon=_f(v[12]), rotation_on_lat=_f(v[13]),
                   mask_file_name=v[18].split()[0], bottom_topography_file_name=v[19].split()[0])

    def basin(self):
        return SwhBasin(self.nx, self.ny, self.dxst, self.dyst, self.rlon, self.rlat, self.curve_grid,
                        self.rotation_on_lon, self.rotation_on_lat)


@dataclasses.dataclass
class SwPar:  # configs/sw.f90:34-41 ; defaults = shipped sw.par
    full_free_surface: int = 1
    trans_terms: int = 1
    ksw_lat: int = 1
    time_smooth: float = 0.5
    lvisc_2: float = 1.0e3
    use_tracers: int = 0
    tracer_num: int = 1
    ssh_init_file_name: str = "none"

    @classmethod
    def from_file(cls, path):
        v = read_par(path)
        return cls(int(v[0]), int(v[1]), int(v[2]), _f(v[3]), _f(v[4]), int(v[5]), int(v[6]), v[7].split()[0])


@dataclasses.dataclass
class ParallelPar:  # configs/parallel.f90:34-42
    mod_decomposition: int = 0
    bppnx: int = 1
    bppny: int = 1

    @classmethod
    def from_file(cls, path):
        v = read_par(path)
        return cls(int(v[0]), int(v[2]), int(v[3]))


@dataclasses.dataclass
class RunPar:  # tools/time_manager.f90:139-175 ; time_step and run_duration are real(4) there
    time_step: float = 1.0
    run_duration: float = 0.007

    @classmethod
    def from_file(cls, path):
        v = read_par(path)
        return cls(_f(v[1]), _f(v[2]))

    @property
    def tau(self):
        return float(np.float32(self.time_step))  # tools/time_manager.f90:270

    @property
    def num_step_max(self):
        nstep_per_day = np.rint(np.float32(86400.0) / np.float32(self.time_step))  # :226
        return int(np.float32(self.run_duration) * np.float32(nstep_per_day))       # :266


def read_mask_file(path, nx, ny):
    """tools/io.f90:61-71: a comment line, then ny rows of nx digits, first row = n = ny."""
    with open(path) as f:
        f.readline()
        rows = [f.readline().rstrip("\n") for _ in range(ny)]
    m = np.empty((ny, nx), dtype=np.int32)
    for i, row in enumerate(rows):
        if len(row) < nx:
            raise ValueError("mask row too short")
        m[ny - 1 - i, :] = np.frombuffer(row[:nx].encode(), dtype=np.uint8) - ord("0")
    return m


# ----------------------------------------------------------------------------- decomposition
def block_dims(nx, ny, bnx, bny, bm, bn):
    """block_uniform_decomposition (core/decomposition.f90:427-503) for block (bm, bn), 0-based."""
    L = _lib.lib()
    xs, xn, ys, yn = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(L.swh_uniform_split(nx - 4, bnx, bm, C.byref(xs), C.byref(xn)))
    check(L.swh_uniform_split(ny - 4, bny, bn, C.byref(ys), C.byref(yn)))
    x0, y0 = 3 + xs.value, 3 + ys.value
    return SwcuDims(x0, x0 + xn.value - 1, y0, y0 + yn.value - 1,
                    x0 - 2, x0 + xn.value + 1, y0 - 2, y0 + yn.value + 1)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def balanced_slab_dims(nx, ny, world, rank, mask=None, band_rows=8, tile_cols=32, land_cost=0.03):
    """Block `rank` of `world` y-slabs cut for equal work (swh_balanced_slabs): slabs with much land get more
    rows.  Deterministic, so every rank computes the same cut."""
    ys = np.zeros(world, dtype=np.int32)
    yn = np.zeros(world, dtype=np.int32)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        assert mask.shape == (ny, nx)
    check(_lib.lib().swh_balanced_slabs(nx, ny, _ptr(mask) if mask is not None else None, world, band_rows, tile_cols,
                                        float(land_cost), _ptr(ys), _ptr(yn)))
    y0 = 3 + int(ys[rank])
    return SwcuDims(3, nx - 2, y0, y0 + int(yn[rank]) - 1, 1, nx, y0 - 2, y0 + int(yn[rank]) + 1)


def block_weights(nx, ny, bnx, bny, mask=None):
    """bglob_weight(bm, bn): sea cells per block (core/decomposition.f90:505-520), shape (bny, bnx)."""
    w = np.zeros((bny, bnx), dtype=np.float64)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        assert mask.shape == (ny, nx)
    check(_lib.lib().swh_block_weights(nx, ny, bnx, bny, _ptr(mask) if mask is not None else None, _ptr(w)))
    return w


def hilbert_partition(weights, nranks, powers=None):
    """create_hilbert_curve_decomposition (core/decomposition.f90:532-612): owner rank per block, shape like
    `weights` (bny, bnx) with bnx = bny = 2^M; -1 = land-only block."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    assert w.ndim == 2 and w.shape[0] == w.shape[1]
    owner = np.zeros(w.shape, dtype=np.int32)
    pw = None if powers is None else np.ascontiguousarray(powers, dtype=np.float64)
    check(_lib.lib().swh_hilbert_partition(w.shape[0], _ptr(w), int(nranks), _ptr(pw) if pw is not None else None,
                                           _ptr(owner)))
    return owner


def uniform_partition(weights, px, py):
    """create_uniform_decomposition (core/decomposition.f90:614-670) on a px x py process grid."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    owner = np.zeros(w.shape, dtype=np.int32)
    check(_lib.lib().swh_uniform_partition(w.shape[1], w.shape[0], int(px), int(py), _ptr(w), _ptr(owner)))
    return owner


# ----------------------------------------------------------------------------- host inputs
class BlockInputs:
    """What init_grid_data + init_ocean_data leave in grid_data / ocean_data for one block
    (control/init_data.f90:29-125), as host arrays in the reference layout."""

    def __init__(self, basin: BasinPar, sw: SwPar, dims: SwcuDims, mask=None, *, hhq_rest=100.0,
                 keep_mu=False, r_diss=0.0):
        L = _lib.lib()
        self.dims = dims
        shape = dims.shape
        b = basin.basin()
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
            assert mask.shape == (basin.ny, basin.nx)
        self.f = {}
        for n in F4_NAMES:
            self.f[n] = np.zeros(shape, dtype=np.float32)
        f = self.f
        check(L.swh_masks(C.byref(b), C.byref(dims), _ptr(mask) if mask is not None else None,
                          *[_ptr(f[n]) for n in ("lu", "luu", "luh", "lcu", "lcv", "llu", "llv")]))
        check(L.swh_metrics(C.byref(b), C.byref(dims),
                            *[_ptr(f[n]) for n in ("dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s")]))
        if r_diss:
            f["r_diss"][:] = np.float32(r_diss)
        f["hhq_rest"] = np.full(shape, float(hhq_rest), dtype=np.float64)   # init_data.f90:112-114
        # Gaussian bump on the sea cells of the global interior that this block's array covers
        # (own interior + the halo cells the reference's following sync fills), vel_ssh.f90:28-36
        wide = SwcuDims(max(dims.bnd_x1, 3), min(dims.bnd_x2, basin.nx - 2),
                        max(dims.bnd_y1, 3), min(dims.bnd_y2, basin.ny - 2),
                        dims.bnd_x1, dims.bnd_x2, dims.bnd_y1, dims.bnd_y2)
        ssh = np.zeros(shape, dtype=np.float64)
        check(L.swh_gaussian(C.byref(wide), _ptr(f["lu"]), _ptr(ssh), 1.0, basin.nx // 2, basin.ny // 2))
        f["ssh"] = ssh
        f["sshp"] = ssh.copy()
        f["sshn"] = ssh.copy()
        for n in ("ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp"):
            f[n] = np.zeros(shape, dtype=np.float64)
        # init_data.f90:76-77: mu is filled with lvisc_2 and immediately zeroed (keep_mu skips the zeroing)
        f["mu"] = np.full(shape, sw.lvisc_2 if keep_mu else 0.0, dtype=np.float64)
        if sw.use_tracers > 0:  # init_data.f90:80-90
            ff = np.zeros(shape, dtype=np.float64)
            check(L.swh_gaussian(C.byref(wide), _ptr(f["lu"]), _ptr(ff), 0.5, basin.nx // 2, basin.ny // 2))
            f["ff1"], f["ff1n"], f["ff1p"] = ff, ff.copy(), ff.copy()


# ----------------------------------------------------------------------------- device context
class DeviceBlock:
    """One block resident on one GPU (Level B of include/swcuda.h)."""

    def __init__(self, dims: SwcuDims, sw: SwPar, device=0, mode=MODE_FUSED, exact=None):
        """exact: True = arithmetic bitwise equal to the reference's CPU path (k_step), False = the same scheme
        re-associated for speed (k_march; rel. L2 <= 1e-12 after 1000 steps), None = the library's default
        (tolerance mode unless the environment says SWCU_EXACT=1)."""
        self.L = _lib.lib()
        self.dims = dims
        self.sw = sw
        self.mode = mode
        self.params = SwcuParams(sw.full_free_surface, sw.trans_terms, sw.ksw_lat, sw.time_smooth,
                                 1 if sw.use_tracers > 0 else 0, mode)
        h = C.c_void_p()
        check(self.L.swcu_create(C.byref(h), C.byref(dims), C.byref(self.params), device))
        self.h = h
        if exact is not None and mode == MODE_FUSED:
            self.set_option("exact", 1 if exact else 0)
        if sw.use_tracers > 0 and sw.tracer_num > 1:       # core/ocean.f90:91-94: ff1(tracer_num)
            self.set_option("tracer_num", int(sw.tracer_num))

    def select_tracer(self, k):
        """Tracer k (0-based) becomes the one the names ff1 / ff1n / ff1p address."""
        self.set_option("tracer_select", int(k))

    def upload(self, name, arr):
        want = np.float64 if name in F8_NAMES else np.float32
        a = np.ascontiguousarray(arr, dtype=want)
        assert a.shape == self.dims.shape, (name, a.shape, self.dims.shape)
        check(self.L.swcu_upload(self.h, FIELD_ID[name], _ptr(a)))
        check(self.L.swcu_synchronize(self.h, None))

    def upload_ptr(self, name, host_ptr):
        """Asynchronous upload from a (pinned) host pointer; caller keeps the buffer alive."""
        check(self.L.swcu_upload(self.h, FIELD_ID[name], C.c_void_p(host_ptr)))

    def download(self, name, out=None):
        want = np.float64 if name in F8_NAMES else np.float32
        if out is None:
            out = np.empty(self.dims.shape, dtype=want)
        check(self.L.swcu_download(self.h, FIELD_ID[name], _ptr(out)))
        return out

    def output_record(self, name):
        """fp32 interior record with undef on land, as local_output writes it (control/output.f90:101-135)."""
        d = self.dims
        out = np.empty((d.ny_end - d.ny_start + 1, d.nx_end - d.nx_start + 1), dtype=np.float32)
        check(self.L.swcu_output_record(self.h, FIELD_ID[name], _ptr(out)))
        return out

    def download_ptr(self, name, host_ptr):
        check(self.L.swcu_download(self.h, FIELD_ID[name], C.c_void_p(host_ptr)))

    def upload_rows(self, name, arr, first_row):
        want = np.float64 if name in F8_NAMES else np.float32
        a = np.ascontiguousarray(arr, dtype=want)
        assert a.shape[1] == self.dims.shape[1]
        check(self.L.swcu_upload_rows(self.h, FIELD_ID[name], _ptr(a), int(first_row), int(a.shape[0])))
        check(self.L.swcu_synchronize(self.h, None))

    def upload_inputs_striped(self, basin, sw, mask=None, stripe_rows=512, **kw):
        """Same result as upload_inputs(BlockInputs(...)) but built and uploaded `stripe_rows` array rows
        at a time, so host memory stays bounded for blocks far larger than host RAM (16384^2 per GPU).
        Arrays that are identically zero are not uploaded (device fields start zero-filled)."""
        d = self.dims
        h = d.shape[0]
        for a in range(0, h, stripe_rows):
            b = min(a + stripe_rows, h) - 1                      # upload array rows a..b
            extra = 1 if b < h - 1 else 0                        # derived masks of row b need lu of row b+1
            n1, n2 = d.bnd_y1 + a, d.bnd_y1 + b + extra
            sd = SwcuDims(d.nx_start, d.nx_end, n1, n2, d.bnd_x1, d.bnd_x2, n1, n2)
            inp = BlockInputs(basin, sw, sd, mask, **kw)
            for name, arr in inp.f.items():
                part = arr[:b - a + 1]
                if name in ("sshn", "ubrtrn", "vbrtrn", "ff1n") or not part.any():
                    continue
                self.upload_rows(name, part, a)
            del inp
        self.hh_init()

    def init_on_device(self, basin, sw, mask=None, *, hhq_rest=100.0, keep_mu=False, r_diss=0.0, stripe_rows=1024):
        """init_grid_data + init_ocean_data (control/init_data.f90:29-125) with the static inputs built ON
        THE DEVICE (swcu_init_grid / swcu_fill / swcu_copy_field).  Only the Gaussian initial state is
        evaluated on the host -- libm's exp, `stripe_rows` rows at a time -- and uploaded once; the
        other time levels are device copies.  Same resident state as upload_inputs(BlockInputs(...))."""
        L, d = self.L, self.dims
        b = basin.basin()
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
            assert mask.shape == (basin.ny, basin.nx)
        check(L.swcu_init_grid(self.h, C.byref(b), _ptr(mask) if mask is not None else None))
        check(L.swcu_fill(self.h, FIELD_ID["hhq_rest"], float(hhq_rest)))            # init_data.f90:112-114
        if keep_mu:
            check(L.swcu_fill(self.h, FIELD_ID["mu"], float(sw.lvisc_2)))            # :76 without :77
        if r_diss:
            check(L.swcu_fill(self.h, FIELD_ID["r_diss"], float(np.float32(r_diss))))
        ics = [("ssh", 1.0, ("sshp", "sshn"))]
        if sw.use_tracers > 0:
            ics.append(("ff1", 0.5, ("ff1p", "ff1n")))                             # :80-90
        h, w = d.shape
        for a in range(0, h, stripe_rows):
            rows = min(stripe_rows, h - a)
            n1, n2 = d.bnd_y1 + a, d.bnd_y1 + a + rows - 1
            sd = SwcuDims(d.nx_start, d.nx_end, n1, n2, d.bnd_x1, d.bnd_x2, n1, n2)
            lu = np.zeros((rows, w), dtype=np.float32)
            check(L.swh_masks(C.byref(b), C.byref(sd), _ptr(mask) if mask is not None else None, _ptr(lu),
                              None, None, None, None, None, None))
            wide = SwcuDims(max(d.bnd_x1, 3), min(d.bnd_x2, basin.nx - 2), max(n1, 3), min(n2, basin.ny - 2),
                            d.bnd_x1, d.bnd_x2, n1, n2)
            for name, sigma, _ in ics:
                f = np.zeros((rows, w), dtype=np.float64)
                if wide.ny_start <= wide.ny_end:
                    check(L.swh_gaussian(C.byref(wide), _ptr(lu), _ptr(f), sigma, basin.nx // 2, basin.ny // 2))
                if f.any():
                    self.upload_rows(name, f, a)
        for name, _, copies in ics:
            for c in copies:
                check(L.swcu_copy_field(self.h, FIELD_ID[c], FIELD_ID[name]))
        self.hh_init()
        check(L.swcu_synchronize(self.h, None))

    def upload_inputs(self, inp: BlockInputs):
        for name, arr in inp.f.items():
            if name == "r_diss" and not arr.any():
                continue  # the reference never assigns r_diss (core/ocean.f90:32)
            self.upload(name, arr)
        for k in range(1, self.sw.tracer_num if self.sw.use_tracers > 0 else 1):   # init_data.f90:82-86: every tracer
            self.select_tracer(k)                                                 # starts from the same Gaussian
            for name in ("ff1", "ff1n", "ff1p"):
                self.upload(name, inp.f[name])
        if self.sw.use_tracers > 0 and self.sw.tracer_num > 1:
            self.select_tracer(0)
        self.hh_init()

    def hh_init(self):
        """envoke(hh_init) of init_ocean_data (control/init_data.f90:60-63); no-op in FUSED mode."""
        check(self.L.swcu_envoke_hh_init(self.h))

    def step(self, tau, nsteps=1):
        check(self.L.swcu_step(self.h, float(tau), int(nsteps)))

    def envoke(self, kernel, tau=0.0):
        """envoke(sub_kernel, sub_sync, kernel_parameters) for one (kernel, sync) pair on the resident
        arrays (core/kernel_interface.f90:48-119), REFERENCE mode."""
        kid = _lib.KERNEL_ID[kernel]
        check(self.L.swcu_envoke_kernel(self.h, kid, float(tau)))
        check(self.L.swcu_envoke_sync(self.h, kid))

    def envoke_kernel(self, kernel, tau=0.0):
        check(self.L.swcu_envoke_kernel(self.h, _lib.KERNEL_ID[kernel], float(tau)))

    def envoke_sync(self, kernel):
        check(self.L.swcu_envoke_sync(self.h, _lib.KERNEL_ID[kernel]))

    def set_option(self, name, value):
        check(self.L.swcu_set_option(self.h, name.encode(), int(value)))

    @property
    def uses_metric_tables(self):
        return bool(self.L.swcu_uses_metric_tables(self.h))

    def synchronize(self):
        bad = C.c_long()
        check(self.L.swcu_synchronize(self.h, C.byref(bad)))
        return bad.value

    def timer_start(self):
        check(self.L.swcu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.L.swcu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    @property
    def launches(self):
        return self.L.swcu_launch_count(self.h)

    @property
    def device_bytes(self):
        return self.L.swcu_device_bytes(self.h)

    def comm_init(self, nranks, rank, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        check(self.L.swcu_comm_init(self.h, nranks, rank, buf))

    def peer_export(self):
        """IPC handles of this block's prognostic buffers for the neighbouring ranks (swcu_peer_export)."""
        buf = C.create_string_buffer(_lib.PEER_BLOB_BYTES)
        check(self.L.swcu_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, side, blob: bytes):
        """side 0: the block below (rank-1), 1: the block above (rank+1)."""
        check(self.L.swcu_peer_attach(self.h, int(side), C.create_string_buffer(blob, _lib.PEER_BLOB_BYTES)))

    def halo_exchange(self, name):
        check(self.L.swcu_halo_exchange(self.h, FIELD_ID[name]))

    def widen_halos(self):
        """Two halo layers of every resident input from the neighbouring blocks (FUSED-mode precondition when
        the uploaded arrays are valid one layer out only, like the reference's: include/swcuda.h)."""
        check(self.L.swcu_widen_halos(self.h))

    def link(self, other):
        """Ties this block to a neighbouring block of the same process (side or corner; swcu_link)."""
        check(self.L.swcu_link(self.h, other.h))

    def close(self):
        if self.h:
            self.L.swcu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def step_group(blocks, tau, nsteps=1):
    """One expl_shallow_water call over all blocks of this process (swcu_step_group)."""
    arr = (C.c_void_p * len(blocks))(*[b.h for b in blocks])
    check(_lib.lib().swcu_step_group(arr, len(blocks), float(tau), int(nsteps)))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(_lib.lib().swcu_comm_unique_id(buf))
    return buf.raw


# ----------------------------------------------------------------------------- the model
class ShallowWaterModel:
    """`program model` restricted to the shallow-water path (model.f90:47-200): one block per GPU,
    y-slabs (parallel.par bppnx = 1, bppny = world size)."""

    def __init__(self, basin: BasinPar = None, sw: SwPar = None, run: RunPar = None, *, mask=None,
                 device=0, mode=MODE_FUSED, rank=0, world=1, hhq_rest=100.0, keep_mu=False, r_diss=0.0,
                 stripe_rows=None, device_init=False, balance=False, exact=None):
        self.basin = basin or BasinPar()
        self.sw = sw or SwPar()
        self.run = run or RunPar()
        self.rank, self.world = rank, world
        if mask is None and self.basin.mask_file_name != "none":
            mask = read_mask_file(self.basin.mask_file_name, self.basin.nx, self.basin.ny)
        if balance and world > 1:   # slabs of equal work (sea-weight balancing) instead of equal height
            self.dims = balanced_slab_dims(self.basin.nx, self.basin.ny, world, rank, mask)
        else:
            self.dims = block_dims(self.basin.nx, self.basin.ny, 1, world, 0, rank)
        self.block = DeviceBlock(self.dims, self.sw, device=device, mode=mode, exact=exact)
        if device_init:
            self.inputs = None
            self.block.init_on_device(self.basin, self.sw, mask, hhq_rest=hhq_rest, keep_mu=keep_mu, r_diss=r_diss,
                                      stripe_rows=stripe_rows or 1024)
        elif stripe_rows:
            self.inputs = None
            self.block.upload_inputs_striped(self.basin, self.sw, mask, stripe_rows, hhq_rest=hhq_rest, keep_mu=keep_mu,
                                             r_diss=r_diss)
        else:
            self.inputs = BlockInputs(self.basin, self.sw, self.dims, mask, hhq_rest=hhq_rest, keep_mu=keep_mu,
                                      r_diss=r_diss)
            self.block.upload_inputs(self.inputs)
        self.tau = self.run.tau
        self.num_step = 0

    @classmethod
    def from_par_dir(cls, path, **kw):
        return cls(BasinPar.from_file(os.path.join(path, "basin.par")), SwPar.from_file(os.path.join(path, "sw.par")),
                   RunPar.from_file(os.path.join(path, "ocean_run.par")), **kw)

    def attach_comm(self, unique_id):
        self.block.comm_init(self.world, self.rank, unique_id)

    def attach_peers(self, all_gather_object):
        """Halo exchange over peer memory instead of NCCL (ranks on one node).  `all_gather_object(list, obj)`
        is torch.distributed's (or any equivalent): the blobs travel once, at set-up."""
        if self.world == 1:
            return
        blobs = [None] * self.world
        try:
            mine = self.block.peer_export()
        except _lib.SwcuError as e:
            mine = e
        all_gather_object(blobs, mine)
        ok, err = True, None
        try:
            for b in blobs:
                if isinstance(b, Exception):
                    raise b
            if self.rank > 0:
                self.block.peer_attach(0, blobs[self.rank - 1])
            if self.rank + 1 < self.world:
                self.block.peer_attach(1, blobs[self.rank + 1])
        except _lib.SwcuError as e:
            ok, err = False, e
        # every rank must end up on the same path: all attached, or none
        oks = [None] * self.world
        all_gather_object(oks, ok)
        if not all(oks):
            check(self.block.L.swcu_peer_detach(self.block.h))
            raise _lib.SwcuError(4, f"peer-memory halo path unavailable on some rank ({err})")

    def expl_shallow_water(self, nsteps=1):
        self.block.step(self.tau, nsteps)
        self.num_step += nsteps

    step = expl_shallow_water

    def expl_shallow_water_envokes(self):
        """The reference's own algorithm layer, statement by statement (control/shallow_water/
        shallow_water.f90:22-94 then control/tracer.f90:33-62): every (kernel, sync) pair is envoked
        through the C ABI on the resident arrays.  REFERENCE mode only."""
        sw, e, tau = self.sw, self.block.envoke, self.tau
        e("sw_update_ssh", tau)
        if sw.full_free_surface > 0:
            e("hh_update")
        if sw.trans_terms > 0:
            e("uv_trans_vort")
            e("uv_trans")
        if sw.ksw_lat > 0:
            e("stress_components")
            e("uv_diff2")
        e("sw_update_uv", tau)
        e("sw_next_step")
        if sw.full_free_surface > 0:
            e("hh_shift")
            e("hh_init")
        e("check_ssh_err")
        if sw.use_tracers > 0:
            e("tran_diff_fluxes")
            e("tran_diff_tracer", tau)
            e("tracer_next_step")
        self.num_step += 1

    def get(self, name):
        return self.block.download(name)

    @property
    def cells_per_step(self):
        d = self.dims
        return (d.nx_end - d.nx_start + 1) * (d.ny_end - d.ny_start + 1)


class BlockGridModel:
    """The same program with the domain cut into bnx x bny blocks that all live in THIS process
    (parallel.par bppnx x bppny blocks per process; core/decomposition.f90:427-503), on one GPU or
    dealt round-robin over `devices` (the reference's _GPU_MULTI_ mode).  Blocks are linked to their
    eight neighbours and step in lockstep through swcu_step_group."""

    def __init__(self, basin: BasinPar = None, sw: SwPar = None, run: RunPar = None, *, bnx=1, bny=1, mask=None,
                 devices=(0,), mode=MODE_FUSED, hhq_rest=100.0, keep_mu=False, r_diss=0.0, skip_land_blocks=True,
                 device_init=False, decomposition="round_robin", exact=None):
        self.basin = basin or BasinPar()
        self.sw = sw or SwPar()
        self.run = run or RunPar()
        self.bnx, self.bny = bnx, bny
        if mask is None and self.basin.mask_file_name != "none":
            mask = read_mask_file(self.basin.mask_file_name, self.basin.nx, self.basin.ny)
        # which GPU holds which block: parallel.par mod_decomposition 0 (uniform rectangles) / 1 (pieces of
        # the Hilbert walk balanced by sea cells), with the devices in the role of the ranks
        self.owner = None
        if decomposition in ("hilbert", "uniform"):
            wts = block_weights(self.basin.nx, self.basin.ny, bnx, bny, mask)
            self.owner = hilbert_partition(wts, len(devices)) if decomposition == "hilbert" else \
                uniform_partition(wts, 1, len(devices))
        elif decomposition != "round_robin":
            raise ValueError(decomposition)
        self.grid = {}
        self.land_blocks = []
        for bn in range(bny):
            for bm in range(bnx):
                d = block_dims(self.basin.nx, self.basin.ny, bnx, bny, bm, bn)
                # Blocks without a single sea cell get no context, like bglob_proc = -1 in the reference
                # (core/decomposition.f90:515-521,576-580): their cells never change and their neighbours'
                # halo cells towards them keep the uploaded (land) values.
                if skip_land_blocks and mask is not None and \
                        np.all(np.asarray(mask)[d.ny_start - 1:d.ny_end, d.nx_start - 1:d.nx_end] != 0):
                    self.land_blocks.append((bm, bn))
                    continue
                dev = devices[len(self.grid) % len(devices)] if self.owner is None else devices[max(self.owner[bn, bm], 0)]
                blk = DeviceBlock(d, self.sw, device=dev, mode=mode, exact=exact)
                if device_init:
                    blk.init_on_device(self.basin, self.sw, mask, hhq_rest=hhq_rest, keep_mu=keep_mu, r_diss=r_diss)
                else:
                    blk.upload_inputs(BlockInputs(self.basin, self.sw, d, mask, hhq_rest=hhq_rest, keep_mu=keep_mu,
                                                  r_diss=r_diss))
                self.grid[(bm, bn)] = blk
        for (bm, bn), blk in self.grid.items():   # each pair once: E, N, NE, NW of every block
            for dm, dn in ((1, 0), (0, 1), (1, 1), (-1, 1)):
                other = self.grid.get((bm + dm, bn + dn))
                if other is not None:
                    blk.link(other)
        self.blocks = list(self.grid.values())
        self.tau = self.run.tau
        self.num_step = 0

    def expl_shallow_water(self, nsteps=1):
        step_group(self.blocks, self.tau, nsteps)
        self.num_step += nsteps

    step = expl_shallow_water

    def synchronize(self):
        return sum(b.synchronize() for b in self.blocks)

    def get(self, name):
        """The global (ny, nx) array assembled from the block interiors (frame cells stay zero)."""
        want = np.float64 if name in F8_NAMES else np.float32
        out = np.zeros((self.basin.ny, self.basin.nx), dtype=want)
        for blk in self.blocks:
            d = blk.dims
            a = blk.download(name)
            out[d.ny_start - 1:d.ny_end, d.nx_start - 1:d.nx_end] = \
                a[d.ny_start - d.bnd_y1:d.ny_end - d.bnd_y1 + 1, d.nx_start - d.bnd_x1:d.nx_end - d.bnd_x1 + 1]
        return out

    def close(self):
        for b in self.blocks:
            b.close()
