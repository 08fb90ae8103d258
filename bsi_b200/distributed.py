"""Batch sharding of sample / elbo over the GPUs of one box (one process per GPU).

Sampling and ELBO evaluation are embarrassingly parallel over samples / data points:
each rank works on a contiguous slice, noise is keyed by the GLOBAL sample index so the result
does not depend on the world size, and the only communication is one final all_gather.
The split rule is the reference's per-rank batch split (bsi/data/h5image.py:309-312):
n // W items per rank, the first n % W ranks take one more.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """(start, count) of rank's contiguous slice of n items."""
    base, extra = divmod(n, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of unevenly sharded rows (dim 0) into the full [n_total, ...] tensor on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_range(n_total, r, world)[1] for r in range(world)]
    width = max(counts)
    padded = local.new_zeros((width, *local.shape[1:]))
    padded[: local.shape[0]] = local
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(pieces, counts)], dim=0)


def sharded_sample(bsi, n_total: int, seed: int, *, t=None, gather: bool = True, group=None) -> torch.Tensor:
    """BSI.sample(n_total) split over the ranks; identical to the single-GPU result for the same seed."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, count = shard_range(n_total, rank, world)
    local = bsi.sample(count, t=t, seed=seed, sample_offset=start) if count > 0 else torch.empty((0, *bsi.data_shape), **bsi.tensor_args)
    return gather_rows(local, n_total, group) if gather else local


def sharded_elbo(bsi, x_full: torch.Tensor, n_recon: int, n_measure: int, seed: int, *, estimate_var: bool = False, group=None):
    """``bsi.elbo(x_full, n_recon, n_measure, Generator(seed))`` with the data points split over the ranks.

    Every rank seeds the same generator, so the Philox keys, the low-discrepancy offset and the permutation over all
    ``n_measure * B`` grid points (bsi/bsi.py:422-440) are drawn ONCE for the whole batch and each rank keeps its columns; the
    noise of row (replica, data point) is keyed by its index in the full problem.  The gathered result therefore equals the
    single-GPU call entry by entry for any world size; the only communication is the final all_gather of ``[n, B]`` losses.
    Returns (elbo[B], bpd[B], extra) like ``BSI.elbo``."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B = len(x_full)
    start, count = shard_range(B, rank, world)
    gen = torch.Generator(device=x_full.device).manual_seed(seed)
    x_local = x_full[start : start + count]
    l_recon = bsi.reconstruction_loss(x_local, n_recon, gen, _shard=(start, B))
    l_measure = bsi.inf_measurement_loss(x_local, n_measure, gen, _shard=(start, B))
    # [n, B_local] -> gather along the data axis
    l_recon = gather_rows(l_recon.t().contiguous(), B, group).t().contiguous()
    l_measure = gather_rows(l_measure.t().contiguous(), B, group).t().contiguous()
    return bsi._assemble_elbo(l_recon, l_measure, estimate_var)
