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
