"""Multi-GPU sharding of the projection path.

Every projection is independent, so the path shards trivially: rank r owns a contiguous slice of the
counter-based seed stream (no seed exchange — seed i depends only on (rng_seed, i)) and projects it with
the same kernel.  The ONE exchange step is collecting the converged states for the planner's batched
sampler: an all-gather of the per-rank converged counts followed by an all-gather of the per-rank
compacted (densely packed by the kernel epilogue) converged states, padded to a fixed capacity so that
no host synchronisation is needed to size the collective.  torch.distributed is the plumbing (NCCL over
NVLink/NVSwitch on GPUs; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of [0, total) into `world` slices whose sizes differ by at most one.
    Returns (first_index, count) of `rank`."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_capacity(count: int, ok_fraction_bound: float = 0.40) -> int:
    """Rows reserved per rank in the gathered pool.  Uniform seeds converge inside the joint limits on
    ~21-23 % of samples, so 40 % is a safe static bound; overflow is detected, not silent."""
    return max(1, min(count, int(count * ok_fraction_bound) + 64))


def gather_converged(compact, n_ok, capacity: int, group=None, pool=None, counts=None):
    """All-gather the converged states of every rank.

    compact : (>=capacity, n) float64 tensor, this rank's converged states packed at the front
    n_ok    : int64[1] tensor, how many rows of `compact` are valid
    returns (pool, counts): pool (world, capacity, n) with rank r's states in pool[r, :counts[r]];
    counts int64[world].  Asynchronous on the current stream (NCCL); no host sync.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = compact.shape[1]
    if pool is None:
        pool = torch.empty((world, capacity, n), dtype=compact.dtype, device=compact.device)
    if counts is None:
        counts = torch.empty(world, dtype=torch.int64, device=compact.device)
    dist.all_gather_into_tensor(counts, n_ok.view(1), group=group)
    dist.all_gather_into_tensor(pool.view(world * capacity, n), compact[:capacity], group=group)
    return pool, counts


def unpack_pool(pool, counts):
    """Drop the padding: (sum(counts), n) tensor of all ranks' converged states, rank-major.
    Raises if a rank overflowed its capacity (its count exceeds the rows that were gathered)."""
    import torch

    cap = pool.shape[1]
    cl = [int(v) for v in counts.tolist()]
    if any(v > cap for v in cl):
        raise OverflowError(f"a rank converged {max(cl)} states but the gather capacity is {cap}")
    return torch.cat([pool[r, : cl[r]] for r in range(pool.shape[0])], dim=0)


class ShardedSampleProjector:
    """sample -> project -> compact on this rank's slice of the seed stream, then gather across ranks.

    Used by the batched sampler when a process group exists; with world size 1 it degrades to the local
    kernel call with no collective.
    """

    def __init__(self, constraint, group=None):
        import torch.distributed as dist

        self.c = constraint
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    def sample_project(self, rng_seed: int, first_index: int, total: int, mode: int = 0, distance: float = 0.0,
                       near=None, wrap_bounds: bool = False):
        """Projects seeds [first_index, first_index+total) of stream `rng_seed`, sharded over the ranks.
        Returns (states, counts): all ranks' converged states (sum(counts), n) and the per-rank counts."""
        import ctypes as C

        import numpy as np
        import torch

        from . import _capi

        c = self.c
        first, count = shard_range(total, self.rank, self.world)
        dev = torch.device("cuda", c.device)
        n = c.getAmbientDimension()
        cap = gather_capacity(max(shard_range(total, 0, self.world)[1], 1))
        compact = torch.empty((max(count, cap), n), dtype=torch.float64, device=dev)
        n_ok = torch.zeros(1, dtype=torch.int64, device=dev)
        near_arr = None if near is None else np.ascontiguousarray(near, dtype=np.float64)
        args = _capi.SamplerArgs(rng_seed=rng_seed, first_index=first_index + first, mode=mode,
                                 wrap_bounds=1 if wrap_bounds else 0, distance=distance,
                                 near_host=None if near_arr is None else near_arr.ctypes.data_as(C.POINTER(C.c_double)))
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = c._lib.ccp_sample_project_batch(c._h, C.byref(args), count, _capi.CCP_LAYOUT_AOS, None, None, None,
                                             compact.data_ptr(), n_ok.data_ptr(), stream)
        if rc != 0:
            raise _capi.CcpError(c._lib.ccp_last_error(c._h).decode())
        if self.world == 1:
            k = int(n_ok.item())
            return compact[:k], torch.tensor([k], dtype=torch.int64, device=dev)
        pool, counts = gather_converged(compact, n_ok, cap, self.group)
        return unpack_pool(pool, counts), counts
