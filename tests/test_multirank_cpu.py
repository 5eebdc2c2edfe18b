"""CPU (gloo, world_size 2 and 3) tests of the N>1 host logic: y-slab decomposition, per-rank input
construction and the halo-row bookkeeping the CUDA library uses for its NCCL exchange
(swcu_halo_plan).  Ranks exchange rows with gloo exactly where NCCL would and must reproduce the
global single-block fields."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, ny, nrows, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import basins
        from ocean_model_arch_b200 import _lib, model
        L = _lib.lib()
        mask = basins.island_mask(nx, ny)
        bp = model.BasinPar(nx=nx, ny=ny)
        d = model.block_dims(nx, ny, 1, world, 0, rank)
        inp = model.BlockInputs(bp, model.SwPar(), d, mask)
        glob = model.BlockInputs(bp, model.SwPar(), model.block_dims(nx, ny, 1, 1, 0, 0), mask)
        # a field with distinct values everywhere (sync_test's i*j pattern, syncborder_block2D_gen_test.fi:10-97)
        jj, ii = np.mgrid[1:ny + 1, 1:nx + 1]
        pattern = (ii * jj).astype(np.float64)
        mine = np.zeros(d.shape)
        r0, r1 = d.ny_start - d.bnd_y1, d.ny_end - d.bnd_y1
        mine[r0:r1 + 1] = pattern[d.ny_start - 1:d.ny_end]          # interior rows only; halos unknown
        reqs = []
        recv = {}
        for side, peer in ((0, rank - 1), (1, rank + 1)):
            if peer < 0 or peer >= world:
                continue
            s, r = C.c_int(), C.c_int()
            assert L.swcu_halo_plan(C.byref(d), nrows, side, C.byref(s), C.byref(r)) == 0
            send = torch.from_numpy(mine[s.value:s.value + nrows].copy())
            buf = torch.empty_like(send)
            recv[r.value] = buf
            reqs.append(dist.isend(send, peer))
            reqs.append(dist.irecv(buf, peer))
        for q in reqs:
            q.wait()
        for row, buf in recv.items():
            mine[row:row + nrows] = buf.numpy()
        ok = True
        # after the exchange every halo row that a neighbour owns equals the global field
        lo = (r0 - nrows) if rank > 0 else r0
        hi = (r1 + nrows) if rank < world - 1 else r1
        ok &= np.array_equal(mine[lo:hi + 1], pattern[d.bnd_y1 - 1 + lo:d.bnd_y1 + hi])
        # per-rank inputs equal the matching rows of the one-block inputs wherever the reference defines them
        for f in ("lu", "lcu", "lcv", "llu", "llv", "luu", "luh", "dx", "dyh", "rlh_s", "ssh", "hhq_rest"):
            a = inp.f[f][1:-1, 1:-1]
            b = glob.f[f][d.bnd_y1:d.bnd_y2 - 1, 1:-1]
            ok &= np.array_equal(a, b)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nrows", [(2, 2), (3, 1), (2, 1)])
def test_slab_halo_plan_gloo(swlib, world, nrows):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + world * 10 + nrows
    procs = [ctx.Process(target=_worker, args=(r, world, port, 60, 47, nrows, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_halo_plan_values(swlib):
    from ocean_model_arch_b200 import _lib
    d = _lib.SwcuDims(3, 50, 13, 22, 1, 52, 11, 24)   # rows 13..22 interior, array rows 11..24
    s, r = C.c_int(), C.c_int()
    assert swlib.swcu_halo_plan(C.byref(d), 2, 0, C.byref(s), C.byref(r)) == 0 and (s.value, r.value) == (2, 0)
    assert swlib.swcu_halo_plan(C.byref(d), 2, 1, C.byref(s), C.byref(r)) == 0 and (s.value, r.value) == (10, 12)
    assert swlib.swcu_halo_plan(C.byref(d), 1, 0, C.byref(s), C.byref(r)) == 0 and (s.value, r.value) == (2, 1)
    assert swlib.swcu_halo_plan(C.byref(d), 1, 1, C.byref(s), C.byref(r)) == 0 and (s.value, r.value) == (11, 12)
    assert swlib.swcu_halo_plan(C.byref(d), 3, 1, C.byref(s), C.byref(r)) == _lib.SWCU_ERR_ARG


# ---- set-up logic of the peer-memory halo path and of the balanced slab cut, with gloo ----------------------
class _StubBlock:
    """Stands in for DeviceBlock (no GPU here): records what attach_peers hands to which side."""

    def __init__(self, rank, fail_on_rank):
        self.rank, self.fail, self.attached, self.detached = rank, rank == fail_on_rank, {}, False
        self.h = None

        class _L:
            @staticmethod
            def swcu_peer_detach(h):
                self.detached = True
                return 0
        self.L = _L()

    def peer_export(self):
        return b"blob-of-rank-%d" % self.rank

    def peer_attach(self, side, blob):
        from ocean_model_arch_b200._lib import SwcuError
        if self.fail:
            raise SwcuError(1, "cudaIpcOpenMemHandle refused (simulated)")
        self.attached[side] = blob


def _peer_worker(rank, world, port, fail_on_rank, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import basins
        from ocean_model_arch_b200 import model
        from ocean_model_arch_b200._lib import SwcuError
        m = model.ShallowWaterModel.__new__(model.ShallowWaterModel)   # host logic only, no device
        m.rank, m.world, m.block = rank, world, _StubBlock(rank, fail_on_rank)
        try:
            m.attach_peers(dist.all_gather_object)
            raised = False
        except SwcuError:
            raised = True
        b = m.block
        want = {}
        if rank > 0:
            want[0] = b"blob-of-rank-%d" % (rank - 1)
        if rank + 1 < world:
            want[1] = b"blob-of-rank-%d" % (rank + 1)
        if fail_on_rank < 0:
            ok = (not raised) and b.attached == want and not b.detached
        else:                                   # one rank cannot map its neighbours: EVERY rank backs out
            ok = raised and b.detached
        # the balanced slab cut is a pure function of the mask: all ranks agree, slabs tile the basin
        nx, ny = 120, 203
        mask = basins.island_mask(nx, ny)
        mask[:90, 30:] = 1
        d = model.balanced_slab_dims(nx, ny, world, rank, mask)
        spans = [None] * world
        dist.all_gather_object(spans, (d.ny_start, d.ny_end))
        ok &= spans[0][0] == 3 and spans[-1][1] == ny - 2
        ok &= all(b2[0] == a2[1] + 1 for a2, b2 in zip(spans, spans[1:]))
        ok &= spans[rank] == (d.ny_start, d.ny_end)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,fail_on_rank", [(2, -1), (3, -1), (3, 1)])
def test_peer_attach_and_balanced_cut_gloo(swlib, world, fail_on_rank):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29660 + world * 3 + (fail_on_rank + 1)
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, fail_on_rank, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
