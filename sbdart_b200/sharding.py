"""Sharding of a run's bins over GPUs.

Bins (wavelength x k-term, and columns for retrieval batches) are independent
inside the solve, so each rank takes one contiguous block and the only
exchange is a single all-gather of the per-bin outputs at the end; the host
then performs the ordered accumulation of drt.f:977-982 on the gathered
spectrum, which keeps the text output identical for 1/2/4/8 GPUs.
"""
from __future__ import annotations

import numpy as np


def bin_partition(nbins: int, world: int, group=None):
    """Contiguous, balanced [start, stop) blocks, one per rank.

    `group[b]` (optional, non-decreasing) is the wavelength index of bin b: a
    block never splits the k-terms of one wavelength (drt.f:529-560 sums them
    in order)."""
    cuts = [round(r * nbins / world) for r in range(world + 1)]
    if group is not None:
        g = np.asarray(group)
        for r in range(1, world):
            c = cuts[r]
            while 0 < c < nbins and g[c] == g[c - 1]:
                c += 1
            cuts[r] = c
        cuts = sorted(cuts)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_outputs(local, partition, dist=None):
    """All-gather per-bin outputs (torch tensor [n_local, ...]) into the full
    [nbins, ...] tensor, in original bin order, on every rank."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [b - a for a, b in partition]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)
