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


def test_inertial_and_viscous_known_answers():
    """Two more known answers from the PHYSICS, not from the reference's code (carthesian grid, flat
    bottom, rigid-lid depths, flat ssh, so every other term vanishes at cells away from the walls):

    * f-plane: a uniform current (u0, v0) turns clockwise in the northern hemisphere,
      du/dt = +f v, dv/dt = -f u.  One leapfrog step from equal time levels must give
      u = u0 + 2 tau f v0, v = v0 - 2 tau f u0 with f = rlh_s = 2 Omega sin(45 deg) (real(4)).
      Pins the sign, the 4-point averaging and the hhh*dxb*dyb / (hhu*dxt*dyh) bookkeeping of K7's
      Coriolis term, and that uniform flow feels no advection (K3, K4).
    * a shear flow u(y) under lateral viscosity nu diffuses, du/dt = nu d2u/dy2, evaluated at the lagged
      level like every leapfrog scheme must: u = u0 + 2 tau nu (u0(n+1) - 2 u0(n) + u0(n-1)) / dy^2.
      Pins K5's shearing stress, K6's dxb^2 * mu * hh * str_s flux difference and the sign of RHS_dif."""
    nx, ny, tau = 60, 56, 2.0
    jj, ii = np.mgrid[1:ny + 1, 1:nx + 1]
    core = (slice(20, 36), slice(20, 40))                      # cells far from the walls

    def model(**kw):
        m = OracleModel(make_config(nx, ny, curve_grid=0, dxst=0.01, dyst=0.02, full_free_surface=0, time_step=tau,
                                    time_smooth=0.0, **kw), None)
        for f in ("ssh", "sshp", "sshn"):
            m.set(f, np.zeros((ny, nx)))
        return m

    # ---- inertial turning
    m = model()
    u0, v0 = 0.3, -0.2
    for f, val, mask in (("ubrtr", u0, "lcu"), ("ubrtrp", u0, "lcu"), ("vbrtr", v0, "lcv"), ("vbrtrp", v0, "lcv")):
        m.set(f, val * m.get(mask).astype(np.float64))
    f_cor = float(m.get("rlh_s")[30, 30])
    omega = 7.2921159e-5
    assert abs(f_cor - 2 * omega * np.sqrt(0.5)) < 1e-10 and f_cor > 0      # northern hemisphere
    m.step(1)
    assert np.abs(m.get("ubrtr")[core] - (u0 + 2 * tau * f_cor * v0)).max() < 1e-15
    assert np.abs(m.get("vbrtr")[core] - (v0 - 2 * tau * f_cor * u0)).max() < 1e-15
    assert not m.get("ssh")[core].any()                                       # uniform flow: no divergence there

    # ---- viscous decay of a shear flow
    nu = 500.0
    m = model(keep_mu=1, lvisc_2=nu)
    m.set("rlh_s", np.zeros((ny, nx), np.float32))
    dy = float(m.get("dy")[30, 30])
    prof = 0.1 * np.sin(2 * np.pi * jj / 17.0) + 0.05 * np.cos(2 * np.pi * jj / 9.0)
    for f in ("ubrtr", "ubrtrp"):
        m.set(f, prof * m.get("lcu"))
    assert np.array_equal(m.get("mu")[core], np.full((16, 20), nu))
    m.step(1)
    lap = (np.roll(prof, -1, 0) - 2 * prof + np.roll(prof, 1, 0)) / dy ** 2
    dx = float(m.get("dx")[30, 30])
    sq4 = float(np.float32(dx) * np.float32(dx)) / dx ** 2     # K6 squares dxb in REAL(4) (vel_ssh.f90:432)
    assert abs(sq4 - 1) > 1e-9                                   # ... and the test resolves that
    want = prof + 2 * tau * nu * lap * sq4
    got = m.get("ubrtr")
    assert np.abs(got[core] - want[core]).max() < 1e-9 * np.abs(want[core] - prof[core]).max()
    assert np.abs(got[core] - prof[core]).max() > 1e-9                        # viscosity did act
    assert np.abs(m.get("vbrtr")[core]).max() < 1e-18                         # only rounding residue of the metric terms


def test_free_surface_mass_flux_known_answer():
    """full_free_surface = 1: the volume flux through a face is u * (H + ssh averaged onto the face), so in a
    uniform current u0 the surface elevation is advected, ssh - 2 tau u0 (ssh(m+1) - ssh(m-1)) / (2 dx)
    (continuity with the depth correction of K10 feeding K1).  The model's own Gaussian bump is the profile."""
    nx, ny, tau, u0 = 64, 48, 2.0, 0.4
    m = OracleModel(make_config(nx, ny, curve_grid=0, dxst=0.01, dyst=0.01, time_step=tau), None)
    m.set("rlh_s", np.zeros((ny, nx), np.float32))
    for f in ("ubrtr", "ubrtrp"):
        m.set(f, u0 * m.get("lcu").astype(np.float64))
    ssh0 = m.get("ssh")
    dx = float(m.get("dx")[24, 30])
    core = (slice(12, 36), slice(16, 48))
    assert np.abs(m.get("hhu")[core] - (100.0 + (ssh0 + np.roll(ssh0, -1, 1)) / 2)[core]).max() < 1e-12
    m.step(1)
    inc = -2 * tau * u0 * (np.roll(ssh0, -1, 1) - np.roll(ssh0, 1, 1)) / (2 * dx)
    got = m.get("ssh") - ssh0
    assert np.abs(got[core] - inc[core]).max() < 1e-6 * np.abs(inc[core]).max()   # real(4) dx*dy in K1: 6e-8
    assert np.abs(inc[core]).max() > 1e-6


def test_advection_filter_and_tracer_known_answers():
    """Known answers for the remaining terms, again from the equations (carthesian, flat, rigid-lid depths,
    no Coriolis; cells far from the walls, one step from equal time levels):

    * momentum self-advection in flux form, u_t = -(u u)_x: u - 2 tau ([u]_e^2 - [u]_w^2)/dx with face
      averages [u]_e = (u(m) + u(m+1))/2 (K4);
    * the Robert-Asselin filter: the lagged level becomes u0 + (time_smooth/2)(u_new - 2 u0 + u0) (K8);
    * a tracer in a uniform current with diffusivity K = mu: centred advection plus diffusion,
      f - 2 tau u0 (f(m+1) - f(m-1))/(2 dx) + 2 tau K (f(m+1) - 2 f(m) + f(m-1))/dx^2 (tracer kernels)."""
    nx, ny, tau, ts = 64, 48, 2.0, 0.5
    jj, ii = np.mgrid[1:ny + 1, 1:nx + 1]
    core = (slice(16, 32), slice(20, 44))
    m = OracleModel(make_config(nx, ny, curve_grid=0, dxst=0.01, dyst=0.01, full_free_surface=0, time_step=tau,
                                time_smooth=ts, ksw_lat=0), None)
    m.set("rlh_s", np.zeros((ny, nx), np.float32))
    for f in ("ssh", "sshp", "sshn"):
        m.set(f, np.zeros((ny, nx)))
    dx = float(m.get("dx")[24, 30])
    prof = 0.2 + 0.05 * np.sin(2 * np.pi * ii / 13.0)
    for f in ("ubrtr", "ubrtrp"):
        m.set(f, prof * m.get("lcu"))
    m.step(1)
    ue, uw = (prof + np.roll(prof, -1, 1)) / 2, (prof + np.roll(prof, 1, 1)) / 2
    want = prof - 2 * tau * (ue ** 2 - uw ** 2) / dx
    got = m.get("ubrtr")
    assert np.abs(got[core] - want[core]).max() < 1e-9 * np.abs(want[core] - prof[core]).max()
    assert np.abs(want[core] - prof[core]).max() > 1e-7
    want_p = prof + ts * (want - 2 * prof + prof) / 2                   # Robert-Asselin, coefficient time_smooth/2
    assert np.abs(m.get("ubrtrp")[core] - want_p[core]).max() < 1e-9 * np.abs(want_p[core] - prof[core]).max()
    assert np.abs(m.get("vbrtr")[core]).max() < 1e-18

    K, u0 = 300.0, 0.25
    m = OracleModel(make_config(nx, ny, curve_grid=0, dxst=0.01, dyst=0.01, full_free_surface=0, time_step=tau,
                                time_smooth=ts, use_tracers=1, keep_mu=1, lvisc_2=K), None)
    m.set("rlh_s", np.zeros((ny, nx), np.float32))
    for f in ("ssh", "sshp", "sshn"):
        m.set(f, np.zeros((ny, nx)))
    for f in ("ubrtr", "ubrtrp"):
        m.set(f, u0 * m.get("lcu").astype(np.float64))
    tr = 1.0 + 0.3 * np.cos(2 * np.pi * ii / 11.0) * m.get("lu")
    for f in ("ff1", "ff1p", "ff1n"):
        m.set(f, tr)
    m.step(1)
    assert np.abs(m.get("ubrtr")[core] - u0).max() < 1e-15               # uniform current: nothing acts on it
    adv = -u0 * (np.roll(tr, -1, 1) - np.roll(tr, 1, 1)) / (2 * dx)
    dif = K * (np.roll(tr, -1, 1) - 2 * tr + np.roll(tr, 1, 1)) / dx ** 2
    want = tr + 2 * tau * (adv + dif)
    got = m.get("ff1")
    assert np.abs(got[core] - want[core]).max() < 1e-9 * np.abs(want[core] - tr[core]).max()
    assert np.abs(2 * tau * adv[core]).max() > 1e-6 and np.abs(2 * tau * dif[core]).max() > 1e-6


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


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c_kernels_equal_numpy_kernels_on_random_fields_and_masks(seed):
    """Per kernel, with every mask an INDEPENDENT random 0/1 field (so masked stores, the slu divisors and
    every stencil arm are exercised in combinations a real coastline never produces) and random positive
    metrics that vary in both directions."""
    rng = np.random.default_rng(seed)
    nx, ny = 37, 29
    b = npr.Block(nx, ny)
    dims = (3, nx - 2, 3, ny - 2, 1, nx, 1, ny)
    M = {n: (rng.random((ny, nx)) < 0.7).astype(np.float32) for n in ("lu", "luu", "luh", "lcu", "lcv", "llu", "llv")}
    G = {n: (1000.0 + 500.0 * rng.random((ny, nx))).astype(np.float32)
         for n in ("dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb")}
    G["rlh_s"] = (1e-4 * rng.standard_normal((ny, nx))).astype(np.float32)
    G["r_diss"] = (1e-5 * rng.random((ny, nx))).astype(np.float32)

    def fld(scale=1.0, positive=False):
        a = rng.random((ny, nx)) + 0.5 if positive else rng.standard_normal((ny, nx))
        return np.ascontiguousarray(scale * a)
    metrics8 = [G[n] for n in ("dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb")]
    tau, ts = 0.75, 0.5

    def both(cname, c_args, np_fn, np_args, outs):
        """outs: indices (into c_args, np_args) of the arrays the kernel writes."""
        ca = [a.copy() if isinstance(a, np.ndarray) else a for a in c_args]
        na = [a.copy() if isinstance(a, np.ndarray) else a for a in np_args]
        call_kernel(cname, dims, *ca)
        np_fn(b, *na)
        for ic, inp in outs:
            assert np.array_equal(ca[ic], na[inp], equal_nan=True), (cname, ic)
            assert not np.array_equal(ca[ic], c_args[ic]), (cname, ic, "kernel wrote nothing")

    u, v, up, vp, ssh, sshp = fld(0.3), fld(0.3), fld(0.3), fld(0.3), fld(0.5), fld(0.5)
    un, vn, sshn = fld(0.3), fld(0.3), fld(0.5)
    hq, hu, hv, hh = (fld(100.0, True) for _ in range(4))
    dep = [fld(100.0, True) for _ in range(12)]
    h_r, mu, vort, st, ss = fld(100.0, True), fld(1000.0, True), fld(1e-3), fld(1e-4), fld(1e-4)
    rx, ry, rxa, rya, rxd, ryd = (fld(1e-2) for _ in range(6))
    ff, ffp, ffn, fx, fy = fld(1.0), fld(1.0), fld(1.0), fld(1.0), fld(1.0)
    lu, luu, luh, lcu, lcv, llu, llv = (M[n] for n in ("lu", "luu", "luh", "lcu", "lcv", "llu", "llv"))

    a = [tau, lu, G["dx"], G["dy"], G["dxh"], G["dyh"], hu, hv, sshn, sshp, u, v]
    both("sw_update_ssh_kernel", a, npr.sw_update_ssh, a, [(8, 8)])
    a = [lu, llu, llv, luh, *metrics8, dep[2], dep[5], dep[8], dep[11], ssh, h_r]
    both("hh_update_kernel", a, npr.hh_update, a, [(12, 12), (13, 13), (14, 14), (15, 15)])
    a = [luu, G["dxt"], G["dyt"], G["dxb"], G["dyb"], u, v, vort]
    both("uv_trans_vort_kernel", a, npr.uv_trans_vort, a, [(7, 7)])
    c = [lcu, lcv, luu, G["dxh"], G["dyh"], u, v, vort, hq, hu, hv, hh, rxa, rya]
    n = [lcu, lcv, luu, G["dxh"], G["dyh"], u, v, vort, hu, hv, hh, rxa, rya]
    both("uv_trans_kernel", c, npr.uv_trans, n, [(12, 11), (13, 12)])
    a = [lu, luu, *metrics8, up, vp, st, ss]
    both("stress_components_kernel", a, npr.stress_components, a, [(12, 12), (13, 13)])
    c = [lcu, lcv, *metrics8, mu, st, ss, hq, hu, hv, hh, rxd, ryd]
    n = [lcu, lcv, *metrics8, mu, st, ss, hq, hh, rxd, ryd]
    both("uv_diff2_kernel", c, npr.uv_diff2, n, [(17, 15), (18, 16)])
    a = [tau, lcu, lcv, G["dxt"], G["dyt"], G["dxh"], G["dyh"], G["dxb"], G["dyb"], dep[3], dep[5], dep[4], dep[6], dep[8],
         dep[7], hh, ssh, u, un, up, v, vn, vp, G["r_diss"], G["rlh_s"], rx, ry, rxa, rya, rxd, ryd]
    both("sw_update_uv", a, npr.sw_update_uv, a, [(18, 18), (21, 21)])
    a = [ts, lu, lcu, lcv, ssh, sshn, sshp, u, un, up, v, vn, vp]
    both("sw_next_step", a, npr.sw_next_step, a, [(4, 4), (6, 6), (7, 7), (9, 9), (10, 10), (12, 12)])
    a = [ts, lu, llu, llv, luh, *dep]
    both("hh_shift_kernel", a, npr.hh_shift, a, [(5 + i, 5 + i) for i in (0, 1, 3, 4, 6, 7, 9, 10)])
    a = [1, lu, llu, llv, luh, *metrics8, *dep, ssh, sshp, h_r]
    both("hh_init_kernel", a, npr.hh_init, a, [(13 + i, 13 + i) for i in range(12)])
    c = [lcu, lcv, G["dxt"], G["dyt"], G["dxh"], G["dyh"], hu, hv, ff, ffp, u, v, mu, 1.0, fx, fy]
    n = [lcu, lcv, G["dxt"], G["dyt"], G["dxh"], G["dyh"], hu, hv, ff, u, v, mu, 1.0, fx, fy]
    both("tran_diff_fluxes_kernel", c, npr.tran_diff_fluxes, n, [(14, 13), (15, 14)])
    a = [lu, G["dx"], G["dy"], tau, dep[2], dep[1], fx, fy, ffp, ffn]
    both("tran_diff_tracer_kernel", a, npr.tran_diff_tracer, a, [(9, 9)])
    a = [ts, lu, ffn, ffp, ff]
    both("tracer_next_step_kernel", a, npr.tracer_next_step, a, [(3, 3), (4, 4)])
    bad = ssh.copy()
    bad[5, 7], bad[9, 9], bad[11, 4] = 2e4, np.nan, -1e4
    want = int(sum(lu[j, i] > 0.5 for j, i in ((5, 7), (9, 9), (11, 4))))
    assert call_kernel("check_ssh_err_kernel", dims, lu, bad) == npr.check_ssh_err(b, lu, bad) == want
