"""CPU tests of the oracle (oracle/sw_oracle.c): golden pins and the invariants SURVEY.md 8c lists.
The reference ships no golden vectors for this path -- parity is unpinned; these tests pin the
restatement against itself (committed digests) and against properties the scheme must have."""
import hashlib
import json
import os

import numpy as np
import pytest

import basins
from golden.make_fixtures import CASES, case_mask
from oracle_lib import OracleModel, call_kernel, make_config

GOLD = json.load(open(os.path.join(basins.GOLDEN, "oracle_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_digests(name):
    c = CASES[name]
    m = OracleModel(make_config(c["nx"], c["ny"], **c.get("cfg", {})), case_mask(c))
    done = 0
    for s in c["steps"]:
        assert m.step(s - done) == 0
        done = s
        for f in ("ssh", "ubrtr", "vbrtr"):
            assert sha(m.get(f)) == GOLD[f"{name}/{f}/{s}"]["sha256"], (name, f, s)


def test_bs_mask_fixture():
    m = basins.bs_mask()
    assert m.shape == (163, 289)
    assert int((m == 0).sum()) == 25547
    assert m[:2].all() and m[-2:].all() and m[:, :2].all() and m[:, -2:].all()


@pytest.mark.parametrize("blocks", [(2, 2), (1, 4), (3, 2), (5, 1)])
def test_decomposition_invariance_bitwise(blocks):
    """No reductions, halo exchange is a pure copy: any block grid gives the same bits."""
    nx, ny = 68, 52
    mask = basins.island_mask(nx, ny)
    a = OracleModel(make_config(nx, ny), mask)
    b = OracleModel(make_config(nx, ny, bnx=blocks[0], bny=blocks[1]), mask)
    a.step(60); b.step(60)
    for f in ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp"):
        assert np.array_equal(a.get(f), b.get(f)), f


def test_land_cells_never_change_and_ssh_bounded():
    nx, ny = 68, 52
    mask = basins.island_mask(nx, ny)
    m = OracleModel(make_config(nx, ny), mask)
    assert m.step(200) == 0
    for f in ("ssh", "sshp", "sshn"):
        assert not m.get(f)[mask == 1].any()
    lcu, lcv = m.get("lcu"), m.get("lcv")
    assert not m.get("ubrtr")[lcu < 0.5].any()
    assert not m.get("vbrtr")[lcv < 0.5].any()
    assert np.abs(m.get("ssh")).max() < 1.0


def test_mass_conservation_closed_basin():
    """sum(ssh*dx*dy) over sea cells is conserved to round-off (flux-form K1, closed boundaries)."""
    nx, ny = 68, 52
    m = OracleModel(make_config(nx, ny), basins.island_mask(nx, ny))
    # K1 divides by the real(4) product dx*dy, so that is the cell area the scheme conserves with
    area = (m.get("dx") * m.get("dy")).astype(np.float64) * m.get("lu")
    v0 = (m.get("ssh") * area).sum()
    m.step(300)
    v1 = (m.get("ssh") * area).sum()
    assert abs(v1 - v0) <= 1e-12 * abs(v0)


def test_k2_redundancy_facts():
    """after a step hhq_n..hhh_n computed by K2 in the NEXT step equal hhq..hhh
    bitwise, which is what lets the fused path drop K2/K9."""
    nx, ny = 44, 36
    m = OracleModel(make_config(nx, ny), basins.island_mask(nx, ny, ndisc=3))
    m.step(7)
    hhu, hhv, hhh, hhq = (m.get(f).copy() for f in ("hhu", "hhv", "hhh", "hhq"))
    d = m.block_dims(0)
    shape = (ny, nx)
    f4 = {n: m.get(n) for n in ("lu", "llu", "llv", "luh", "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb")}
    hqn, hun, hvn, hhn = (np.zeros(shape) for _ in range(4))
    call_kernel("hh_update_kernel", d, *[f4[n] for n in f4], hqn, hun, hvn, hhn, m.get("ssh"), m.get("hhq_rest"))
    assert np.array_equal(hqn, hhq)
    assert np.array_equal(hun, hhu) and np.array_equal(hvn, hhv) and np.array_equal(hhn, hhh)


def test_fast_build_stays_within_tolerance():
    """The -O3 -march=native timing build (FMA contraction allowed, like the reference's own -Ofast)
    stays within the north-star tolerance of the strict build: rel. L2 <= 1e-12 after 1000 steps."""
    nx, ny = 68, 52
    mask = basins.island_mask(nx, ny)
    a = OracleModel(make_config(nx, ny), mask)
    b = OracleModel(make_config(nx, ny), mask, fast=True)
    a.step(1000); b.step(1000)
    for f in ("ssh", "ubrtr", "vbrtr"):
        x, y = a.get(f), b.get(f)
        assert np.linalg.norm(x - y) <= 1e-12 * np.linalg.norm(x), f


def test_linear_gravity_mode_known_answer():
    """A known-answer test that does NOT come from reading the reference's code: with transport,
    viscosity, Coriolis, the free-surface depth correction and the time filter switched off, the scheme
    is the classical leapfrog C-grid discretisation of the linear shallow-water equations.  For a
    standing eigenmode cos(p pi x/Lx) cos(q pi y/Ly) of the closed flat basin, eliminating u and v
    between two consecutive steps gives the exact three-level identity
        ssh(n+2) - 2 ssh(n) + ssh(n-2) = -4 tau^2 g H Lambda ssh(n),
        Lambda = (2 sin(p pi/(2 Nx))/dx)^2 + (2 sin(q pi/(2 Ny))/dy)^2,
    to rounding -- except that K1 divides by the REAL(4) product dx*dy (kernel/shallow_water/vel_ssh.f90:100),
    so Lambda carries the factor dx*dy / real4(dx*dy); the test resolves that 5e-8 effect.
    It pins the divergence and cell-area factor of K1, the pressure gradient and the
    bp/bp0 bookkeeping of K7, the level rotation of K8 and the wall masks."""
    nx, ny, p, q = 68, 52, 2, 1
    Nx, Ny = nx - 4, ny - 4
    tau = 5.0
    m = OracleModel(make_config(nx, ny, curve_grid=0, dxst=0.01, dyst=0.01, full_free_surface=0, trans_terms=0,
                                ksw_lat=0, time_smooth=0.0, time_step=tau), None)
    m.set("rlh_s", np.zeros((ny, nx), np.float32))
    dx = float(m.get("dx")[10, 10]); dy = float(m.get("dy")[10, 10])
    jj, ii = np.mgrid[1:ny + 1, 1:nx + 1]
    phi = np.cos(p * np.pi * (ii - 2.5) / Nx) * np.cos(q * np.pi * (jj - 2.5) / Ny)   # cell centres
    phi[m.get("lu") < 0.5] = 0.0
    for f in ("ssh", "sshp", "sshn"):
        m.set(f, 0.1 * phi)
    a = []
    sea = m.get("lu") > 0.5
    for _ in range(400):
        a.append(float((m.get("ssh")[sea] * phi[sea]).sum() / (phi[sea] ** 2).sum()))
        # the state stays in the one-dimensional eigenspace (up to rounding)
        m.step(1)
    a = np.array(a)
    resid = m.get("ssh") - a[-1] * phi
    g = float(np.float32(9.8)); H = 100.0
    lam = (2 * np.sin(p * np.pi / (2 * Nx)) / dx) ** 2 + (2 * np.sin(q * np.pi / (2 * Ny)) / dy) ** 2
    area4 = float(np.float32(dx) * np.float32(dy))                # the real(4) product K1 divides by
    want = -4 * tau ** 2 * g * H * lam * (dx * dy / area4)
    assert abs(dx * dy / area4 - 1) > 1e-9                         # (and the test can tell the difference)
    n = np.arange(2, len(a) - 2)
    big = np.abs(a[n]) > 0.02
    got = (a[n + 2] - 2 * a[n] + a[n - 2])[big] / a[n][big]
    assert np.abs(got - want).max() < 1e-9 * abs(want) + 1e-11, (got.min(), got.max(), want)
    assert 0.05 < np.abs(a).max() < 0.1 * 1.001                  # neutral: no growth (the start excites a tiny computational mode)
    assert np.sign(a).min() < 0 < np.sign(a).max()               # it really oscillates


# ---- the C oracle against an independent NumPy restatement written from the Fortran sources -------------
import np_restatement as npr  # noqa: E402


def _numpy_twin(o, cfg, rng=None):
    """NumpyModel holding a copy of every array of the (single-block) oracle model."""
    fields = {n: o.get(n) for n in npr.F8 + npr.F4}
    if cfg.use_tracers:
        fields.update({n: o.get(n) for n in npr.F8_TRACER})
    return npr.NumpyModel(fields, full_free_surface=cfg.full_free_surface, trans_terms=cfg.trans_terms,
                          ksw_lat=cfg.ksw_lat, time_smooth=cfg.time_smooth, use_tracers=cfg.use_tracers)


@pytest.mark.parametrize("variant", ["shipped", "viscous_friction_tracer", "rigid_lid", "no_adv_no_visc", "cartesian",
                                     "rough_bottom_forced"])
def test_c_oracle_equals_numpy_restatement_bitwise(variant):
    nx, ny = 61, 47
    kw = dict(shipped={}, viscous_friction_tracer=dict(keep_mu=1, r_diss=5e-6, use_tracers=1),
              rigid_lid=dict(full_free_surface=0, keep_mu=1), no_adv_no_visc=dict(trans_terms=0, ksw_lat=0),
              cartesian=dict(curve_grid=0, keep_mu=1, use_tracers=1, time_step=0.75),
              rough_bottom_forced=dict(keep_mu=1, r_diss=1e-5, use_tracers=1))[variant]
    cfg = make_config(nx, ny, **kw)
    o = OracleModel(cfg, basins.island_mask(nx, ny))
    if variant == "rough_bottom_forced":
        # random bathymetry, viscosity and external forcing (arrays the shipped set-up leaves constant / zero)
        rng = np.random.default_rng(11)
        o.set("hhq_rest", 50.0 + 100.0 * rng.random((ny, nx)))
        o.set("mu", 2000.0 * rng.random((ny, nx)))
        o.set("RHSx", 1e-2 * rng.standard_normal((ny, nx)))
        o.set("RHSy", 1e-2 * rng.standard_normal((ny, nx)))
        o.set("ubrtr", 0.1 * rng.standard_normal((ny, nx)) * o.get("lcu"))
        o.set("ubrtrp", o.get("ubrtr"))
        o.set("vbrtr", 0.1 * rng.standard_normal((ny, nx)) * o.get("lcv"))
        o.set("vbrtrp", o.get("vbrtr"))
        o.step(1)   # lets the oracle's own hh_init pick the new bathymetry up before the twin is taken
    twin = _numpy_twin(o, cfg)
    tau = float(np.float32(cfg.time_step))
    names = list(npr.F8) + (list(npr.F8_TRACER) if cfg.use_tracers else [])
    done = 0
    for upto in (1, 2, 25):
        o.step(upto - done); twin.step(tau, upto - done)
        done = upto
        for n in names:
            assert np.array_equal(twin.f[n], o.get(n), equal_nan=True), (variant, n, upto)
    assert twin.bad == 0
    assert np.abs(o.get("ssh")).max() > 1e-3          # the comparison is not of zeros
    if cfg.use_tracers:
        assert np.abs(o.get("ff1")).max() > 1e-3
