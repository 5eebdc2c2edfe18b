"""GPU parity of the TOLERANCE mode (swcu_set_option "exact" = 0, the library default): k_march, one launch
per step, the reference's scheme with re-associated arithmetic (csrc/sw_fast.cuh).  Bar (BASELINE.json
north_star): ssh / u / v within relative L2 <= 1e-12 of the reference's CPU path after 1000 steps; masks and
land / masked-out cells bit-exact.  The bitwise mode (exact = 1) is covered by tests/test_gpu_step.py."""
import numpy as np
import pytest

import basins
from fast_host import FastHostModel
from ocean_model_arch_b200 import model
from ocean_model_arch_b200._lib import MODE_FUSED, SwcuError
from oracle_lib import OracleModel, make_config, redo_hh_init

pytestmark = pytest.mark.gpu
TOL = 1e-12
STATE = ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp")


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-300))


def check_against(o, m, tol=TOL):
    lu = o.get("lu")
    worst = 0.0
    for f in STATE:
        a, b = o.get(f), m.get(f)
        r = rel(a, b)
        assert r <= tol, (f, r)
        worst = max(worst, r)
    for f in ("ssh", "sshp"):   # cells the reference never assigns keep their bits
        assert np.array_equal(o.get(f)[lu < 0.5], m.get(f)[lu < 0.5]), f
    for f, mk in (("ubrtr", "lcu"), ("ubrtrp", "lcu"), ("vbrtr", "lcv"), ("vbrtrp", "lcv")):
        off = o.get(mk) < 0.5
        assert np.array_equal(o.get(f)[off], m.get(f)[off]), f
    return worst


def fast_model(nx, ny, mask=None, sw=None, **kw):
    bkw = {k: kw.pop(k) for k in ("curve_grid", "dxst", "dyst") if k in kw}
    return model.ShallowWaterModel(model.BasinPar(nx=nx, ny=ny, **bkw), sw, mask=mask, mode=MODE_FUSED, exact=False, **kw)


def test_black_sea_1000_steps(swlib, cuda_device):
    """BASELINE config 1's mask (data/BS/mask_bs4km.txt), shipped parameters, 1000 steps."""
    mask = basins.bs_mask()
    ny, nx = mask.shape
    o = OracleModel(make_config(nx, ny, nthreads=8, bny=4), mask)
    m = fast_model(nx, ny, mask)
    for f in ("lu", "luu", "luh", "lcu", "lcv", "llu", "llv"):
        assert np.array_equal(m.get(f), o.get(f)), f
    l0 = m.block.launches
    o.step(1000); m.step(1000)
    assert m.block.synchronize() == 0
    assert m.block.launches - l0 == 1000 + 1     # one k_march launch per step (+ the coefficient table once)
    assert check_against(o, m) < 1e-13


def test_islands_1000_steps_viscosity_friction(swlib, cuda_device):
    """config 3's mask generator at 516^2 with config 4's physics (mu = lvisc_2, r_diss = 5e-6), 1000 steps."""
    nx = ny = 516
    mask = basins.island_mask(nx, ny)
    o = OracleModel(make_config(nx, ny, keep_mu=1, r_diss=5e-6, nthreads=8, bny=8), mask)
    m = fast_model(nx, ny, mask, keep_mu=True, r_diss=5e-6)
    o.step(1000); m.step(1000)
    assert m.block.synchronize() == 0
    assert check_against(o, m) < 1e-13


def test_config2_size_2048_against_oracle(swlib, cuda_device):
    """BASELINE config 2 (2048^2 cells, flat bottom) against the oracle itself at full size, 20 steps."""
    n = 2052
    o = OracleModel(make_config(n, n, nthreads=16, bny=16), None)
    m = fast_model(n, n)
    o.step(20); m.step(20)
    assert m.block.synchronize() == 0
    assert check_against(o, m, 1e-13) < 1e-14


@pytest.mark.parametrize("kw", [dict(trans_terms=0), dict(ksw_lat=0), dict(full_free_surface=0),
                                dict(trans_terms=0, ksw_lat=0, full_free_surface=0)])
def test_flag_combinations_one_launch(swlib, cuda_device, kw):
    """Every sw.par flag combination runs as ONE launch per step (shallow_water.f90:36-92 gates K2/K9/K10,
    K3/K4 and K5/K6 independently)."""
    nx, ny = 133, 91
    mask = basins.island_mask(nx, ny)
    o = OracleModel(make_config(nx, ny, keep_mu=1, **kw), mask)
    m = fast_model(nx, ny, mask, model.SwPar(**kw), keep_mu=True)
    l0 = m.block.launches
    o.step(100); m.step(100)
    assert m.block.synchronize() == 0
    assert m.block.launches - l0 == 101
    check_against(o, m, 1e-13)


@pytest.mark.parametrize("shape", [(5, 5), (6, 9), (31, 7), (32, 33), (33, 64), (61, 47), (150, 203)])
def test_ragged_and_tiny_basins(swlib, cuda_device, shape):
    """Warp columns of 28 cells and row bands that do not divide the basin; basins down to one cell."""
    nx, ny = shape
    mask = basins.island_mask(nx, ny) if min(nx, ny) > 20 else None
    o = OracleModel(make_config(nx, ny, keep_mu=1), mask)
    m = fast_model(nx, ny, mask, keep_mu=True)
    o.step(25); m.step(25)
    assert m.block.synchronize() == 0
    check_against(o, m, 1e-13)


def test_device_kernel_equals_host_evaluation_of_the_same_formulas(swlib, cuda_device):
    """k_march wires registers, shuffles and the shared-memory ring around the functions of sw_fast.cuh; the
    host harness runs the same functions over whole arrays.  They may differ only through the reciprocal
    (MUFU seed + Newton on the device, IEEE division on the host: <= 1 ulp), so after a few steps the two
    agree to ~1e-16 -- far tighter than either agrees with the reference."""
    nx, ny = 150, 119
    mask = basins.island_mask(nx, ny)
    cfg = make_config(nx, ny, keep_mu=1, r_diss=5e-6)
    o = OracleModel(cfg, mask)
    rng = np.random.default_rng(3)
    hrest = 50.0 + 100.0 * rng.random((ny, nx))
    o.set("hhq_rest", hrest)
    redo_hh_init(o)
    h = FastHostModel(o, cfg)
    m = fast_model(nx, ny, mask, keep_mu=True, r_diss=5e-6)
    m.block.upload("hhq_rest", hrest)
    m.step(5); h.step(5)
    assert m.block.synchronize() == 0
    for f in STATE:
        assert rel(h.get(f), m.get(f)) <= 5e-16, (f, rel(h.get(f), m.get(f)))


def test_cartesian_grid_and_time_step_change(swlib, cuda_device):
    """The coefficient table holds tau: changing the step rebuilds it."""
    nx, ny = 132, 100
    mask = basins.island_mask(nx, ny)
    kw = dict(curve_grid=0, dxst=0.01, dyst=0.01)
    for tau in (1.0, 0.75):
        o = OracleModel(make_config(nx, ny, keep_mu=1, time_step=tau, **kw), mask)
        m = fast_model(nx, ny, mask, keep_mu=True, **kw)
        m.tau = tau
        o.step(60); m.step(60)
        check_against(o, m, 1e-13)
    m.tau = 0.5   # same context, new tau
    o2 = OracleModel(make_config(nx, ny, keep_mu=1, time_step=0.5, **kw), mask)
    for f in STATE:
        o2.set(f, m.get(f))
    redo_hh_init(o2)
    o2.step(10); m.step(10)
    check_against(o2, m, 1e-13)


def test_block_grid_is_bitwise_decomposition_invariant(swlib, cuda_device):
    """Tolerance mode is still deterministic: a cell's value does not depend on which warp, band or block
    computes it, so any block grid reproduces the one-block run BITWISE (the reference's sync_test idea)."""
    nx, ny = 133, 91
    mask = basins.island_mask(nx, ny)
    one = fast_model(nx, ny, mask, keep_mu=True)
    one.step(40)
    for layout in ((2, 1), (1, 2), (3, 2)):
        g = model.BlockGridModel(model.BasinPar(nx=nx, ny=ny), bnx=layout[0], bny=layout[1], mask=mask, keep_mu=True,
                                 exact=False)
        g.step(40)
        assert g.synchronize() == 0
        for f in STATE:
            assert np.array_equal(g.get(f)[2:-2, 2:-2], one.get(f)[2:-2, 2:-2]), (f, layout)
        g.close()


def test_exact_option_switches_kernels(swlib, cuda_device):
    """exact = 1 is bitwise equal to the oracle; exact = 0 is within tolerance but (on a real basin) not bitwise;
    switching back and forth on one context works."""
    nx, ny = 133, 91
    mask = basins.island_mask(nx, ny)
    o = OracleModel(make_config(nx, ny), mask)
    m = model.ShallowWaterModel(model.BasinPar(nx=nx, ny=ny), mask=mask, exact=True)
    o.step(30); m.step(30)
    assert all(np.array_equal(o.get(f), m.get(f)) for f in STATE)
    m.block.set_option("exact", 0)
    o.step(30); m.step(30)
    assert not all(np.array_equal(o.get(f), m.get(f)) for f in STATE)
    check_against(o, m, 1e-14)
    m.block.set_option("exact", 1)
    o.step(5); m.step(5)
    check_against(o, m, 1e-14)


def test_blowup_flag_in_tolerance_mode(swlib, cuda_device):
    nx, ny = 64, 48
    m = fast_model(nx, ny)
    bad = m.get("ssh").copy()
    bad[ny // 2, nx // 2] = 3.0e4
    m.block.upload("ssh", bad)
    m.block.upload("sshp", bad)
    m.step(1)
    with pytest.raises(SwcuError):
        m.block.synchronize()


def test_mass_conservation_and_land_at_8192_masked(swlib, cuda_device):
    """BASELINE config 3 size in tolerance mode: all-land bands skipped, size-independent properties hold,
    and the exact mode on the same basin agrees to rounding."""
    n = 8196
    bp = model.BasinPar(nx=n, ny=n, curve_grid=0)
    mask = basins.island_mask(n, n, ndisc=12)
    b = model.ShallowWaterModel(bp, mask=mask, mode=MODE_FUSED, device_init=True, exact=False)
    ssh0 = b.get("ssh")
    b.step(10)
    assert b.block.synchronize() == 0
    got = {f: b.get(f) for f in ("ssh", "ubrtr", "vbrtr")}
    lu = b.get("lu")
    area = (b.get("dx") * b.get("dy")).astype(np.float64) * lu
    b.block.close()
    e = model.ShallowWaterModel(bp, mask=mask, mode=MODE_FUSED, device_init=True, exact=True)
    e.step(10)
    assert e.block.synchronize() == 0
    for f in got:
        assert rel(e.get(f), got[f]) <= 1e-14, f
        assert np.array_equal(e.get(f)[lu < 0.5], got[f][lu < 0.5])
    e.block.close()
    v0, v1 = (ssh0 * area).sum(), (got["ssh"] * area).sum()
    assert abs(v1 - v0) <= 1e-12 * abs(v0)
    assert not got["ssh"][lu < 0.5].any() and not got["ubrtr"][lu < 0.5].any()


def test_config3_masked_basin_4096_against_oracle(swlib, cuda_device):
    """BASELINE config 3's basin at 4096^2 cells against the oracle in tolerance mode, 10 steps."""
    n = 4100
    mask = basins.island_mask(n, n, ndisc=12)
    o = OracleModel(make_config(n, n, curve_grid=0, bnx=1, bny=16, nthreads=16), mask)
    m = fast_model(n, n, mask, curve_grid=0, device_init=True)
    o.step(10); m.step(10)
    assert m.block.synchronize() == 0
    check_against(o, m, 1e-13)


def test_config4_and_5_physics_1024_against_oracle(swlib, cuda_device):
    """BASELINE configs 4 / 5 physics (mu = lvisc_2, r_diss = 5e-6, tracers) at 1024^2 in tolerance mode."""
    n = 1028
    o = OracleModel(make_config(n, n, curve_grid=0, keep_mu=1, r_diss=5e-6, use_tracers=1, bnx=1, bny=16, nthreads=16), None)
    m = fast_model(n, n, None, model.SwPar(use_tracers=1), curve_grid=0, keep_mu=True, r_diss=5e-6)
    o.step(20); m.step(20)
    assert m.block.synchronize() == 0
    check_against(o, m, 1e-13)
    for f in ("ff1", "ff1p"):
        assert rel(o.get(f), m.get(f)) <= 1e-13, f


@pytest.mark.parametrize("shape", [(133, 91), (61, 47), (33, 64)])
def test_tracer_transport_tolerance_mode(swlib, cuda_device, shape):
    """k_tracer_march (expl_tracer in one launch after k_march) against the oracle: 300 steps, viscosity on."""
    nx, ny = shape
    mask = basins.island_mask(nx, ny)
    o = OracleModel(make_config(nx, ny, keep_mu=1, use_tracers=1), mask)
    m = fast_model(nx, ny, mask, model.SwPar(use_tracers=1), keep_mu=True)
    l0 = m.block.launches
    o.step(300); m.step(300)
    assert m.block.synchronize() == 0
    assert m.block.launches - l0 == 2 * 300 + 1
    check_against(o, m, 1e-13)
    lu = o.get("lu")
    for f in ("ff1", "ff1p"):
        assert rel(o.get(f), m.get(f)) <= 1e-13, f
        assert np.array_equal(o.get(f)[lu < 0.5], m.get(f)[lu < 0.5]), f
