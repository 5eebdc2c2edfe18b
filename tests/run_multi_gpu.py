"""Multi-GPU decomposition-invariance check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/run_multi_gpu.py [nx ny steps]

Every rank owns one y-slab of the global basin and exchanges halo rows every step (NCCL or peer memory).
Rank 0 also runs the same basin as ONE block on its GPU and (for small basins) on the CPU oracle.
Bitwise arithmetic (exact = 1): ssh/sshp/u/up/v/vp of the N-slab run must equal both BITWISE.
Tolerance arithmetic (exact = 0): the N-slab run must equal the one-block run of the same arithmetic BITWISE
and the oracle within 1e-12."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STATE = ("ssh", "sshp", "ubrtr", "ubrtrp", "vbrtr", "vbrtrp")


def main():
    import torch
    import torch.distributed as dist
    import basins
    from ocean_model_arch_b200 import model
    from ocean_model_arch_b200._lib import MODE_FUSED, MODE_REFERENCE

    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 203
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    balance = len(sys.argv) > 4 and sys.argv[4] == "balance"   # slabs of equal work instead of equal height
    peer = len(sys.argv) > 4 and sys.argv[4] == "peer"         # halo rows over peer memory instead of NCCL
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mask = basins.island_mask(nx, ny)
    bp = model.BasinPar(nx=nx, ny=ny)
    ok = True
    # the reference's sync_test (shared/mpp/syncborder_block2D_gen_test.fi:10-97): interiors hold i*j, after
    # the halo exchange every halo row a neighbour owns must hold i*j too
    for mode, nrows in ((MODE_FUSED, 2), (MODE_REFERENCE, 1)):
        m = model.ShallowWaterModel(bp, mask=mask, device=local, mode=mode, rank=rank, world=world, balance=balance)
        ids = [model.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        m.attach_comm(ids[0])
        d = m.dims
        jj, ii = np.mgrid[d.bnd_y1:d.bnd_y2 + 1, d.bnd_x1:d.bnd_x2 + 1]
        pat = (ii * jj).astype(np.float64)
        mine = np.zeros_like(pat)
        r0, r1 = d.ny_start - d.bnd_y1, d.ny_end - d.bnd_y1
        mine[r0:r1 + 1] = pat[r0:r1 + 1]
        m.block.upload("mu", mine)
        m.block.halo_exchange("mu")
        got = m.block.download("mu")
        lo = r0 - nrows if rank > 0 else r0
        hi = r1 + nrows if rank < world - 1 else r1
        good = np.array_equal(got[lo:hi + 1], pat[lo:hi + 1]) and not got[:lo].any() and not got[hi + 1:].any()
        flag = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(flag.item())
        if rank == 0:
            print(f"sync_test mode={mode} halo={nrows}: {'ok' if flag.item() else 'FAILED'}", flush=True)
        m.block.close()
        dist.barrier()
    for mode, tiled in ((MODE_FUSED, 1), (MODE_FUSED, 0), (MODE_REFERENCE, 0)):
        sw = model.SwPar(use_tracers=1 if peer else 0)
        fields = STATE + (("ff1", "ff1p") if peer else ())
        m = model.ShallowWaterModel(bp, sw, mask=mask, device=local, mode=mode, rank=rank, world=world, keep_mu=True,
                                    balance=balance, device_init=balance, exact=True)   # the bitwise arithmetic
        if mode == MODE_FUSED:
            m.block.set_option("tiled", tiled)
        if peer and mode == MODE_FUSED:
            m.attach_peers(dist.all_gather_object)
        else:
            ids = [model.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            m.attach_comm(ids[0])
        m.step(steps)
        assert m.block.synchronize() == 0
        d = m.dims
        rows = slice(d.ny_start - d.bnd_y1, d.ny_end - d.bnd_y1 + 1)
        parts = {}
        for f in fields:
            mine = torch.from_numpy(m.get(f)[rows].copy()).cuda()
            sizes = [None] * world
            dist.all_gather_object(sizes, mine.shape[0])
            bufs = [torch.empty((s, mine.shape[1]), dtype=mine.dtype, device="cuda") for s in sizes]
            dist.all_gather(bufs, mine)
            parts[f] = torch.cat(bufs, 0).cpu().numpy()
        if rank == 0:
            one = model.ShallowWaterModel(bp, sw, mask=mask, device=local, mode=MODE_FUSED, keep_mu=True, exact=True)
            one.step(steps)
            for f in fields:
                ref = one.get(f)[2:-2]
                same = np.array_equal(parts[f], ref)
                ok &= same
                if not same:
                    print(f"MISMATCH vs 1-GPU mode={mode} tiled={tiled} {f}: max|d|={np.abs(parts[f] - ref).max()}")
            if nx * ny <= 400 * 400:
                from oracle_lib import OracleModel, make_config
                o = OracleModel(make_config(nx, ny, keep_mu=1, use_tracers=1 if peer else 0), mask)
                o.step(steps)
                for f in fields:
                    same = np.array_equal(parts[f], o.get(f)[2:-2])
                    ok &= same
                    if not same:
                        print(f"MISMATCH vs oracle mode={mode} tiled={tiled} {f}")
            print(f"mode={mode} tiled={tiled} world={world}: {'bitwise equal' if ok else 'FAILED'} "
                  f"(launches {m.block.launches})", flush=True)
        m.block.close()
        dist.barrier()
    # ---- tolerance mode (exact = 0, k_march): deterministic, so N slabs reproduce the one-block run of the same
    # mode BITWISE; both stay within 1e-12 of the oracle.  Halos: peer memory with the push fused into the strip
    # launch of k_march, peer memory + tracers (k_tracer pushes separately), and NCCL (k_march strips).
    for halo, tracers in (("peer", 0), ("peer", 1), ("nccl", 0)):
        sw = model.SwPar(use_tracers=tracers)
        fields = STATE + (("ff1", "ff1p") if tracers else ())
        m = model.ShallowWaterModel(bp, sw, mask=mask, device=local, mode=MODE_FUSED, rank=rank, world=world, keep_mu=True,
                                    balance=balance, device_init=balance, exact=False)
        if halo == "peer":
            m.attach_peers(dist.all_gather_object)
        else:
            ids = [model.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            m.attach_comm(ids[0])
        l0 = m.block.launches
        m.step(steps)
        assert m.block.synchronize() == 0
        per_step = (m.block.launches - l0 - 1) / steps     # - 1: the coefficient table
        d = m.dims
        rows = slice(d.ny_start - d.bnd_y1, d.ny_end - d.bnd_y1 + 1)
        parts = {}
        for f in fields:
            mine = torch.from_numpy(m.get(f)[rows].copy()).cuda()
            sizes = [None] * world
            dist.all_gather_object(sizes, mine.shape[0])
            bufs = [torch.empty((s, mine.shape[1]), dtype=mine.dtype, device="cuda") for s in sizes]
            dist.all_gather(bufs, mine)
            parts[f] = torch.cat(bufs, 0).cpu().numpy()
        if rank == 0:
            good = True
            one = model.ShallowWaterModel(bp, sw, mask=mask, device=local, mode=MODE_FUSED, keep_mu=True, exact=False)
            one.step(steps)
            for f in fields:
                same = np.array_equal(parts[f], one.get(f)[2:-2])
                good &= same
                if not same:
                    print(f"MISMATCH tolerance mode vs 1-GPU halo={halo} {f}: max|d|={np.abs(parts[f] - one.get(f)[2:-2]).max()}")
            if nx * ny <= 400 * 400:
                from oracle_lib import OracleModel, make_config
                o = OracleModel(make_config(nx, ny, keep_mu=1, use_tracers=tracers), mask)
                o.step(steps)
                for f in fields:
                    ref = o.get(f)[2:-2]
                    r = np.linalg.norm(parts[f] - ref) / max(np.linalg.norm(ref), 1e-300)
                    if r > 1e-12:
                        good = False
                        print(f"tolerance mode vs oracle halo={halo} {f}: rel L2 {r}")
            if halo == "peer" and not tracers and d.ny_end - d.ny_start + 1 >= 4 and per_step != 2.0:
                good = False      # the lean interior launch + the concurrent strip / push launch, no push kernel
                print(f"fused halo push: expected two launches per step, counted {per_step}")
            ok &= good
            print(f"tolerance mode halo={halo} tracers={tracers} world={world}: "
                  f"{'identical to 1 GPU' if good else 'FAILED'} ({per_step:.2f} launches per step)", flush=True)
        m.block.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
