"""Device-side input construction (swcu_init_grid / swcu_fill / swcu_copy_field): the
resident state it leaves must equal -- bit for bit -- what building every array on the host
(swh_* = the reference's init_grid_data / init_ocean_data) and uploading it leaves."""
import numpy as np
import pytest

import basins
from ocean_model_arch_b200 import model
from ocean_model_arch_b200._lib import MODE_FUSED, MODE_REFERENCE
from oracle_lib import OracleModel, make_config

pytestmark = pytest.mark.gpu
STATIC = ("lu", "luu", "luh", "lcu", "lcv", "llu", "llv", "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s",
          "hhq_rest", "mu")
STATE = ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp")


@pytest.mark.parametrize("mode", [MODE_REFERENCE, MODE_FUSED])
@pytest.mark.parametrize("case", ["spherical_none", "spherical_islands", "cartesian_islands", "rotated_pole",
                                  "friction_viscosity_tracer"])
def test_device_init_equals_host_init(swlib, cuda_device, case, mode):
    nx, ny = 141, 95
    bp = model.BasinPar(nx=nx, ny=ny, curve_grid=0 if case.startswith("cartesian") else 1,
                        rotation_on_lat=12.5 if case == "rotated_pole" else 0.0,
                        rotation_on_lon=3.0 if case == "rotated_pole" else 0.0)
    mask = None if case == "spherical_none" else basins.island_mask(nx, ny)
    extra = dict(keep_mu=True, r_diss=5e-6) if case == "friction_viscosity_tracer" else {}
    sw = model.SwPar(use_tracers=1 if case == "friction_viscosity_tracer" else 0)
    host = model.ShallowWaterModel(bp, sw, mask=mask, mode=mode, **extra)
    dev = model.ShallowWaterModel(bp, sw, mask=mask, mode=mode, device_init=True, stripe_rows=17, **extra)
    names = STATIC + STATE + (("r_diss", "ff1", "ff1p") if extra else ())
    for f in names:
        assert np.array_equal(dev.get(f), host.get(f)), (f, case, mode)
    host.step(30); dev.step(30)
    assert dev.block.synchronize() == 0
    # the per-row metric tables are usable exactly when the metrics do not vary along x
    assert dev.block.uses_metric_tables == host.block.uses_metric_tables == (case != "rotated_pole")
    for f in STATE + (("ff1",) if extra else ()):
        assert np.array_equal(dev.get(f), host.get(f)), (f, case, mode)
    if case != "rotated_pole":       # the oracle's grid has no rotated pole
        kw = dict(curve_grid=bp.curve_grid)
        if extra:
            kw.update(keep_mu=1, r_diss=5e-6, use_tracers=1)
        o = OracleModel(make_config(nx, ny, **kw), mask)
        o.step(30)
        for f in STATE:
            assert np.array_equal(dev.get(f), o.get(f)), (f, case, mode)


@pytest.mark.parametrize("layout", [(2, 3), (3, 1)])
def test_device_init_on_every_block_of_a_grid(swlib, cuda_device, layout):
    """Blocks that touch the basin frame on some sides only: the [2..nx-1] x [2..ny-1] rule of the metric
    arrays and the derived-mask range differ per block."""
    nx, ny = 90, 77
    bp, sw = model.BasinPar(nx=nx, ny=ny), model.SwPar()
    mask = basins.island_mask(nx, ny)
    for bn in range(layout[1]):
        for bm in range(layout[0]):
            d = model.block_dims(nx, ny, layout[0], layout[1], bm, bn)
            for mode in (MODE_REFERENCE, MODE_FUSED):
                a = model.DeviceBlock(d, sw, mode=mode)
                a.upload_inputs(model.BlockInputs(bp, sw, d, mask))
                b = model.DeviceBlock(d, sw, mode=mode)
                b.init_on_device(bp, sw, mask, stripe_rows=11)
                for f in STATIC + ("ssh", "sshp"):
                    assert np.array_equal(a.download(f), b.download(f)), (f, bm, bn, mode)
                a.close(); b.close()
