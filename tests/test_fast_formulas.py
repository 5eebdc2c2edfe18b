"""CPU check of the TOLERANCE-mode arithmetic (ocean_model_arch_b200/csrc/sw_fast.cuh, the formulas the
device kernel k_march evaluates): the header is compiled as plain C++ (tests/fast_host.cpp) and stepped
over whole arrays against the strict oracle.  North-star bar: relative L2 <= 1e-12 on ssh / u / v after
1000 steps, masks / land cells bit-exact.  Measured: 1e-15 .. 4e-15."""
import numpy as np
import pytest

import basins
from fast_host import FastHostModel
from oracle_lib import OracleModel, make_config, redo_hh_init

TOL = 1e-12   # BASELINE.json north_star: "ssh/u/v must match within relative L2 <= 1e-12" after 1000 steps
STATE = ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp")


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-300))


def run_case(nx, ny, mask, steps, **kw):
    cfg = make_config(nx, ny, **kw)
    o = OracleModel(cfg, mask)
    f = FastHostModel(o, cfg)
    o.step(steps)
    assert f.step(steps) == 0
    lu = o.get("lu")
    for n in STATE:
        a, b = o.get(n), f.get(n)
        assert rel(a, b) <= TOL, (n, rel(a, b))
    for n in ("ssh", "sshp"):                      # land: bit-exact (never assigned)
        assert np.array_equal(o.get(n)[lu < 0.5], f.get(n)[lu < 0.5])
    for n, m in (("ubrtr", "lcu"), ("vbrtr", "lcv")):
        off = o.get(m) < 0.5
        assert np.array_equal(o.get(n)[off], f.get(n)[off])
    return max(rel(o.get(n), f.get(n)) for n in ("ssh", "ubrtr", "vbrtr"))


def test_black_sea_mask_1000_steps():
    """BASELINE config 1's mask (data/BS/mask_bs4km.txt) with the shipped parameters, 1000 steps."""
    m = basins.bs_mask()
    ny, nx = m.shape
    assert run_case(nx, ny, m, 1000) < 1e-13


def test_islands_with_viscosity_and_friction_1000_steps():
    """config 3's mask generator with config 4's physics (mu = lvisc_2, r_diss = 5e-6)."""
    nx, ny = 100, 77
    assert run_case(nx, ny, basins.island_mask(nx, ny), 1000, keep_mu=1, r_diss=5e-6) < 1e-13


@pytest.mark.parametrize("kw", [
    dict(trans_terms=0), dict(ksw_lat=0), dict(full_free_surface=0), dict(trans_terms=0, ksw_lat=0, full_free_surface=0),
    dict(curve_grid=0, dxst=0.01, dyst=0.01, keep_mu=1), dict(time_step=0.75, keep_mu=1), dict(time_smooth=0.1, keep_mu=1),
])
def test_flag_combinations(kw):
    nx, ny = 68, 52
    assert run_case(nx, ny, basins.island_mask(nx, ny), 200, **kw) < 1e-13


def test_random_bathymetry_and_viscosity():
    """Every term active with non-trivial coefficients: random bathymetry, viscosity and velocities."""
    nx, ny = 90, 70
    cfg = make_config(nx, ny, keep_mu=1, r_diss=1e-5)
    o = OracleModel(cfg, basins.island_mask(nx, ny))
    rng = np.random.default_rng(7)
    lu = o.get("lu")
    o.set("hhq_rest", 50.0 + 100.0 * rng.random((ny, nx)))
    o.set("mu", 500.0 + 1000.0 * rng.random((ny, nx)))
    for n, m in (("ubrtr", "lcu"), ("ubrtrp", "lcu"), ("vbrtr", "lcv"), ("vbrtrp", "lcv")):
        o.set(n, 0.05 * (rng.random((ny, nx)) - 0.5) * (o.get(m) > 0.5))
    redo_hh_init(o)   # the oracle stores its depth fields; they derive from hhq_rest
    f = FastHostModel(o, cfg)
    for steps in (1, 30):
        o.step(steps)
        assert f.step(steps) == 0
        for n in STATE:
            assert rel(o.get(n), f.get(n)) <= 1e-13, (n, steps)


def test_tracer_transport_1000_steps():
    """expl_tracer in tolerance arithmetic (tracer_flux / tracer_update of sw_fast.cuh) with viscosity on."""
    nx, ny = 100, 77
    cfg = make_config(nx, ny, keep_mu=1, use_tracers=1)
    o = OracleModel(cfg, basins.island_mask(nx, ny))
    f = FastHostModel(o, cfg)
    o.step(1000)
    assert f.step(1000) == 0
    lu = o.get("lu")
    for n in STATE + ("ff1", "ff1p"):
        assert rel(o.get(n), f.get(n)) <= 1e-13, (n, rel(o.get(n), f.get(n)))
    assert np.array_equal(o.get("ff1")[lu < 0.5], f.get("ff1")[lu < 0.5])


def test_mass_conservation_and_bounds_in_tolerance_arithmetic():
    """Size-independent properties of the re-associated scheme (the ones the GPU tests use at the BASELINE sizes):
    K1 stays in flux form -- the SAME volume flux u*dyh*hhu leaves one cell and enters its neighbour -- so the
    volume sum(ssh * real4(dx*dy)) over the sea is conserved to rounding; land cells never change; |ssh| bounded."""
    nx, ny = 120, 90
    mask = basins.island_mask(nx, ny)
    cfg = make_config(nx, ny, keep_mu=1)
    o = OracleModel(cfg, mask)
    f = FastHostModel(o, cfg)
    lu = o.get("lu")
    area = (o.get("dx") * o.get("dy")).astype(np.float64) * lu
    v0 = float((f.get("ssh") * area).sum())
    ssh0 = f.get("ssh").copy()
    assert f.step(500) == 0
    v1 = float((f.get("ssh") * area).sum())
    assert abs(v1 - v0) <= 1e-12 * abs(v0)
    assert np.array_equal(f.get("ssh")[lu < 0.5], ssh0[lu < 0.5])
    assert np.isfinite(f.get("ssh")).all() and np.abs(f.get("ssh")).max() <= np.abs(ssh0).max() * 1.0000001
    assert np.abs(f.get("ubrtr")).max() > 0
