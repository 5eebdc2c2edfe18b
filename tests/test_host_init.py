"""CPU tests of the product's host layer: C++ input construction (csrc/sw_host.cpp) and the Python
mirror of the .par configs / decomposition, checked bitwise against the oracle's independent
restatement of the same reference code."""
import os

import numpy as np
import pytest

import basins
from ocean_model_arch_b200 import model
from oracle_lib import OracleModel, make_config

F4 = ["lu", "luu", "luh", "lcu", "lcv", "llu", "llv", "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s"]


@pytest.mark.parametrize("curve_grid", [0, 1])
@pytest.mark.parametrize("masked", [False, True])
def test_single_block_inputs_match_oracle_bitwise(swlib, curve_grid, masked):
    nx, ny = 68, 52
    mask = basins.island_mask(nx, ny) if masked else None
    bp = model.BasinPar(nx=nx, ny=ny, curve_grid=curve_grid)
    inp = model.BlockInputs(bp, model.SwPar(), model.block_dims(nx, ny, 1, 1, 0, 0), mask)
    o = OracleModel(make_config(nx, ny, curve_grid=curve_grid), mask)
    for f in F4 + ["ssh", "sshp", "hhq_rest", "mu", "ubrtr"]:
        assert np.array_equal(inp.f[f], o.get(f)), f


def test_bs4km_inputs_match_oracle_bitwise(swlib):
    mask = basins.bs_mask()
    bp = model.BasinPar(nx=289, ny=163, dxst=0.05, dyst=0.04, rlon=27.525, rlat=40.94)
    inp = model.BlockInputs(bp, model.SwPar(), model.block_dims(289, 163, 1, 1, 0, 0), mask)
    o = OracleModel(make_config(289, 163, dxst=0.05, dyst=0.04, rlon=27.525, rlat=40.94), mask)
    for f in F4 + ["ssh"]:
        assert np.array_equal(inp.f[f], o.get(f)), f


def test_slab_blocks_match_oracle_blocks(swlib):
    """y-slab decomposition (bppnx=1, bppny=G): interior + width-1 halo of every block equals the
    oracle's block arrays (which restate core/decomposition.f90 + the halo syncs)."""
    nx, ny, G = 52, 70, 3
    mask = basins.island_mask(nx, ny)
    o = OracleModel(make_config(nx, ny, bnx=1, bny=G), mask)
    bp = model.BasinPar(nx=nx, ny=ny)
    for r in range(G):
        d = model.block_dims(nx, ny, 1, G, 0, r)
        assert list(d.as_tuple()) == o.block_dims(r)
        inp = model.BlockInputs(bp, model.SwPar(), d, mask)
        glob = {f: o.get(f) for f in F4 + ["ssh"]}
        for f in F4 + ["ssh"]:
            # rows/cols the reference defines for this block: start-1 .. end+1
            a = inp.f[f][1:-1, 1:-1]
            b = glob[f][d.bnd_y1:d.bnd_y2 - 1, d.bnd_x1:d.bnd_x2 - 1]
            assert np.array_equal(a, b), (r, f)


def test_par_files_and_time_manager(tmp_path):
    (tmp_path / "basin.par").write_text("\n".join([
        "1525 : nx", "1115 : ny", "1 : nz", "0 : px", "0 : py", "0.00312d0 : dx", "0.00225d0 : dy",
        "34.751560d0 : rlon", "44.801125d0 : rlat", "0 : xgr", "0 : ygr", "1 : curve", "0.0d0 : rot", "0.0d0 :",
        "90.0d0 :", "60.0d0 :", "90.0d0 :", "-90.0d0 :", "none : mask", "none : topo"]))
    (tmp_path / "sw.par").write_text("\n".join(["1 : ffs", "1 : trans", "1 : ksw", "0.5d0 : ts", "1.0d+03 : lvisc",
                                                "0 : tracers", "1 : n", "none : ssh file"]))
    (tmp_path / "ocean_run.par").write_text("0 : start\n1.0d0 : step\n0.007 : days\n0 : n\n")
    (tmp_path / "parallel.par").write_text("0 : mod\nnone : file\n1 : bx\n1 : by\n")
    b = model.BasinPar.from_file(tmp_path / "basin.par")
    assert (b.nx, b.ny, b.curve_grid, b.mask_file_name) == (1525, 1115, 1, "none")
    assert b == model.BasinPar()                      # defaults are the shipped basin.par
    s = model.SwPar.from_file(tmp_path / "sw.par")
    assert s == model.SwPar()
    r = model.RunPar.from_file(tmp_path / "ocean_run.par")
    assert r.tau == 1.0 and r.num_step_max == 604
    p = model.ParallelPar.from_file(tmp_path / "parallel.par")
    assert (p.mod_decomposition, p.bppnx, p.bppny) == (0, 1, 1)


def test_mask_file_reader(tmp_path):
    nx, ny = 12, 9
    m = basins.island_mask(nx, ny, ndisc=1, coast=False)
    rows = ["comment"] + ["".join(str(v) for v in m[n]) for n in range(ny - 1, -1, -1)]
    p = tmp_path / "mask.txt"
    p.write_text("\n".join(rows) + "\n")
    assert np.array_equal(model.read_mask_file(p, nx, ny), m)


def test_uniform_split_matches_reference_rule(swlib):
    # core/decomposition.f90:441-458: floor((N - done)/(blocks left)), last block takes the rest
    for n, nb in [(100, 3), (8192, 8), (1111, 7), (5, 5)]:
        tot = 0
        for i in range(nb):
            d = model.block_dims(n + 4, 9, nb, 1, i, 0)
            assert d.nx_start == 3 + tot
            size = d.nx_end - d.nx_start + 1
            expect = n - tot if i == nb - 1 else (n - tot) // (nb - i)
            assert size == expect
            tot += size
        assert tot == n


# ---- block -> rank assignment (core/decomposition.f90:505-670, shared/mpp/hilbert_curve.f90) ---------------
def _ref_d2xy(m, d):
    """Test-side restatement of hilbert_d2xy / hilbert_rot as the reference writes them."""
    n, x, y, t, s = 2 ** m, 0, 0, d, 1
    while s < n:
        rx = (t // 2) % 2
        ry = t % 2 if rx == 0 else (t ^ rx) % 2
        if ry == 0:
            if rx == 1:
                x, y = s - 1 - x, s - 1 - y
            x, y = y, x
        x, y, t, s = x + s * rx, y + s * ry, t // 4, s * 2
    return x, y


def test_hilbert_curve(swlib):
    import ctypes as C
    for order in range(0, 6):
        n = 1 << order
        seen, prev = set(), None
        for d in range(n * n):
            x, y = C.c_int(), C.c_int()
            assert swlib.swh_hilbert_d2xy(order, d, C.byref(x), C.byref(y)) == 0
            assert (x.value, y.value) == _ref_d2xy(order, d)
            assert 0 <= x.value < n and 0 <= y.value < n
            if prev is not None:
                assert abs(x.value - prev[0]) + abs(y.value - prev[1]) == 1     # a walk over neighbours
            prev = (x.value, y.value)
            seen.add(prev)
        assert len(seen) == n * n


def test_block_weights_and_partitions(swlib):
    nx, ny, nb = 260, 196, 8
    mask = basins.island_mask(nx, ny)
    mask[:, :70] = 1                                                   # a continent: land-only blocks
    w = model.block_weights(nx, ny, nb, nb, mask)
    assert w.sum() == (mask[2:-2, 2:-2] == 0).sum()
    d = model.block_dims(nx, ny, nb, nb, 5, 2)
    assert w[2, 5] == (mask[d.ny_start - 1:d.ny_end, d.nx_start - 1:d.nx_end] == 0).sum()
    assert model.block_weights(nx, ny, nb, nb, None).sum() == (nx - 4) * (ny - 4)
    for nranks in (1, 3, 4, 7):
        own = model.hilbert_partition(w, nranks)
        assert np.array_equal(own == -1, w == 0) and (w == 0).sum() >= 8
        walk = []
        for k in range(nb * nb):
            x, y = _ref_d2xy(3, k)
            if own[y, x] >= 0:
                walk.append(own[y, x])
        assert walk[0] == 0 and all(b - a in (0, 1) for a, b in zip(walk, walk[1:]))   # consecutive pieces
        assert walk[-1] == nranks - 1                                                   # every rank got a piece
        load = np.array([w[own == r].sum() for r in range(nranks)])
        assert load.max() <= 1.35 * w.sum() / nranks, load                              # balanced by sea cells
    heavy = model.hilbert_partition(w, 2, powers=[3.0, 1.0])                            # compute_powers
    assert w[heavy == 0].sum() > 2.0 * w[heavy == 1].sum()
    uni = model.uniform_partition(w, 2, 4)
    assert uni[0, 7] == 1 * 4 + 0 and uni[7, 7] == 1 * 4 + 3 and uni[3, 1] == -1 and uni[3, 2] == 0 * 4 + 1
    assert set(np.unique(uni)) <= set(range(-1, 8))
    with pytest.raises(Exception):
        model.hilbert_partition(np.ones((6, 6)), 2)                                     # needs 2^M blocks a side


def test_balanced_slabs(swlib):
    nx, ny = 260, 404
    mask = basins.island_mask(nx, ny)
    mask[:150, 40:] = 1                                   # the southern third is mostly land
    for world in (1, 2, 3, 8):
        ds = [model.balanced_slab_dims(nx, ny, world, r, mask) for r in range(world)]
        assert ds[0].ny_start == 3 and ds[-1].ny_end == ny - 2
        for a, b in zip(ds, ds[1:]):
            assert b.ny_start == a.ny_end + 1 and (b.ny_start - 3) % 8 == 0     # contiguous, cut on tile rows
        assert all(d.ny_end - d.ny_start + 1 >= 8 for d in ds[:-1])
        assert all((d.bnd_y1, d.bnd_y2) == (d.ny_start - 2, d.ny_end + 2) for d in ds)

        def work(d):   # tiles with any sea cell
            sea = mask[d.ny_start - 1:d.ny_end, 2:-2] == 0
            return sum(sea[j:j + 8, i:i + 32].any() for j in range(0, sea.shape[0], 8) for i in range(0, sea.shape[1], 32))
        if world > 1:
            bal = [work(d) for d in ds]
            uni = [work(model.block_dims(nx, ny, 1, world, 0, r)) for r in range(world)]
            assert max(bal) <= max(uni)
            if world in (2, 3):
                assert max(bal) < 0.8 * max(uni), (bal, uni)          # the land-heavy slab got more rows
    ds = [model.balanced_slab_dims(nx, ny, 5, r, None) for r in range(5)]       # no mask: near-uniform
    assert max(d.ny_end - d.ny_start for d in ds) - min(d.ny_end - d.ny_start for d in ds) <= 8


def test_march_bands_tile_the_rows_exactly_once(swlib):
    """k_march's row bookkeeping (csrc/sw_fused.h march_band_rows): for any row range, band count and number /
    size of shortened late bands, the bands are consecutive, non-overlapping, cover [n0 .. n1] exactly, and the
    late bands are late_cut rows shorter than the others (to within the rounding of the division)."""
    import ctypes as C
    import random
    rnd = random.Random(20241018)
    cases = [(3, 2050, 16, 1, 24), (3, 2050, 16, 2, 24), (3, 2050, 16, 0, 0), (3, 4, 1, 0, 0), (5, 5, 3, 1, 0),
             (3, 8194, 64, 0, 0), (7, 1030, 9, 2, 10)]
    for _ in range(300):
        n0 = rnd.randint(1, 50); rows = rnd.randint(1, 5000); nb = rnd.randint(1, 80)
        late = rnd.randint(0, min(nb, 3)); cut = rnd.choice([0, 1, 8, 24, 40])
        if late and rows // nb <= 2 * cut:
            cut = 0        # the library only shortens bands that are comfortably longer than the cut
        cases.append((n0, n0 + rows - 1, nb, late, cut))
    for n0, n1, nb, late, cut in cases:
        nxt = n0
        lens = []
        for b in range(nb):
            f, l = C.c_int(), C.c_int()
            assert swlib.swcu_march_band_rows(n0, n1, nb, late, cut, b, C.byref(f), C.byref(l)) == 0
            assert f.value == nxt and l.value >= f.value - 1, (n0, n1, nb, late, cut, b, f.value, l.value)
            nxt = l.value + 1
            lens.append(l.value - f.value + 1)
        assert nxt == n1 + 1, (n0, n1, nb, late, cut)
        if late and cut and nb > late:
            normal, short = lens[:nb - late], lens[nb - late:]
            assert max(short) <= min(normal) - cut + 2 and min(short) >= max(normal) - cut - 2, (n0, n1, nb, late, cut, lens)
    f, l = C.c_int(), C.c_int()
    assert swlib.swcu_march_band_rows(1, 10, 4, 5, 0, 0, C.byref(f), C.byref(l)) != 0     # more late bands than bands
