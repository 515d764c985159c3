"""world_size-2 (gloo, CPU) test of the N>1 host path: partition + all-gather."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sbdart_b200.sharding import bin_partition, gather_outputs


def test_partition_covers_and_respects_wavelength_groups():
    group = np.repeat(np.arange(40), 3)[:-1]           # 119 bins, 3 k-terms per wavelength
    for world in (1, 2, 3, 4, 8):
        parts = bin_partition(len(group), world, group)
        assert parts[0][0] == 0 and parts[-1][1] == len(group)
        for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
            assert b == c
            if 0 < b < len(group):
                assert group[b] != group[b - 1]
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 5


def _worker(rank, world, port, nbins):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = bin_partition(nbins, world)
    a, b = parts[rank]
    idx = torch.arange(a, b, dtype=torch.float64)
    local = torch.stack([idx, idx * idx, -idx], dim=1)           # stand-in for [bin][6] fluxes
    full = gather_outputs(local, parts, dist)
    ref = torch.arange(nbins, dtype=torch.float64)
    assert full.shape == (nbins, 3)
    assert torch.equal(full[:, 0], ref) and torch.equal(full[:, 1], ref * ref)
    dist.destroy_process_group()


def test_two_rank_gather_restores_bin_order():
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, 1343), nprocs=2, join=True)


def _run_worker(rank, world, port, nl, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sbdart_b200.frontend import Sbdart
    from solvers import solve_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    txt = Sbdart(nl).run_sharded(lambda b: solve_oracle(b, nthreads=2), dist)
    q.put((rank, txt))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_prints_the_same_records():
    """A whole SBDART run split over two ranks (bins solved by the CPU checker here, by the GPU
    on the box): the records are byte-identical to the single-process run on every rank."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sbdart_b200.frontend import Sbdart
    from solvers import solve_oracle
    for nl in ("&INPUT idatm=4, wlinf=.3, wlsup=.9, wlinc=.02, iout=1 /",
               "&INPUT idatm=2, nstr=4, wlinf=.6, wlsup=.64, wlinc=.02, sza=30, iout=20, uzen=20,120, phi=0,90 /"):
        want = Sbdart(nl).run(solve_oracle)
        ctx = mp.get_context("spawn")
        q = ctx.SimpleQueue()
        port = 31000 + os.getpid() % 2000
        procs = [ctx.Process(target=_run_worker, args=(r, 2, port, nl, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = dict(q.get() for _ in range(2))
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert got[0] == want and got[1] == want


def _failing_worker(rank, world, port, nl, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sbdart_b200.frontend import Sbdart
    from solvers import solve_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def solve(b):
        if rank == 1:
            raise RuntimeError("boom on rank 1")
        return solve_oracle(b, nthreads=2)
    try:
        Sbdart(nl).run_sharded(solve, dist)
        q.put((rank, "returned"))
    except RuntimeError as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_raises_on_every_rank_when_one_rank_fails():
    """A rank whose solve raises must not leave the other rank blocked in the all-gather."""
    nl = "&INPUT idatm=4, wlinf=.3, wlsup=.5, wlinc=.02, iout=1 /"
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_failing_worker, args=(r, 2, port, nl, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, "a rank hung or crashed"
    got = dict(q.get() for _ in range(2))
    assert got[1] == "boom on rank 1"
    assert "another rank" in got[0]
