"""Multi-GPU sharding of the projection path.

Every projection is independent, so the path shards trivially: rank r owns a contiguous slice of the
counter-based seed stream (no seed exchange — seed i depends only on (rng_seed, i)) and projects it with
the same kernel.  The ONE exchange step is collecting the converged states for the planner's batched
sampler.  Two implementations:

  PeerPool (GPUs)     the all-gather is FUSED into the projection kernel: every rank's pool lives in peer-mapped
                      (symmetric) device memory and the kernel's epilogue stores each converged state straight into
                      its rows of every rank's pool over NVLink (ccp_set_gather_peers); only the 8-byte counts are
                      exchanged afterwards (one tiny NCCL all-gather, which also orders the stores for the readers).
  gather_converged    NCCL all-gather of the counts and of the compacted states padded to a fixed capacity — the
                      fallback when peer memory cannot be mapped, and what the gloo CPU tests exercise.

torch.distributed is the plumbing (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Tuple


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


class PeerPool:
    """Every rank's pool of converged states in symmetric (peer-mapped) device memory + the engine hook that makes
    the projection kernel write into all of them.

    pool[r, :counts[r]] holds rank r's converged states once `exchange_counts` has run after the projection
    launch(es) on the same stream.  A pool must not be written by a new projection before every rank has finished
    reading it (any later collective of the group, e.g. the next exchange_counts, is such a point).
    """

    def __init__(self, constraint, capacity: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.c = constraint
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerPool spans one NVLink domain of at most 8 GPUs")
        self.capacity = int(capacity)
        self.n = constraint.getAmbientDimension()
        dev = torch.device("cuda", constraint.device)
        self.pool = symm.empty((self.world, self.capacity, self.n), dtype=torch.float64, device=dev)
        self.handle = symm.rendezvous(self.pool, self.group)
        # buffer_ptrs are the ranks' allocation bases: add the tensor's offset inside its allocation (0 unless torch
        # carved it out of a symmetric-memory pool)
        off = int(self.pool.data_ptr()) - int(self.handle.buffer_ptrs[self.rank])
        self.ptrs = [int(p) + off for p in self.handle.buffer_ptrs]
        # NVLS: a multicast mapping of the pool, when the fabric offers one (0 otherwise): one store reaches every rank
        import os

        self.multicast_ptr = 0
        if os.environ.get("CCP_GATHER_MULTICAST", "1") != "0":
            try:
                mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
                self.multicast_ptr = mc + off if mc else 0
            except Exception:  # noqa: BLE001 - older torch: no multicast support
                self.multicast_ptr = 0
        # the per-rank counts live in symmetric memory too: every rank stores its own count into all of them
        self.counts = symm.empty((self.world,), dtype=torch.int64, device=dev)
        self.counts.zero_()
        self.counts_handle = symm.rendezvous(self.counts, self.group)
        coff = int(self.counts.data_ptr()) - int(self.counts_handle.buffer_ptrs[self.rank])
        self.count_ptrs = [int(p) + coff for p in self.counts_handle.buffer_ptrs]
        self.use_nccl_counts = False

    def attach(self):
        """Point the engine's fused gather at this pool (host-side state only; applies to later launches)."""
        import ctypes as C

        arr = (C.c_uint64 * self.world)(*self.ptrs)
        rc = self.c._lib.ccp_set_gather_peers(self.c._h, self.world, self.rank, arr, self.capacity)
        if rc == 0 and self.multicast_ptr:
            rc = self.c._lib.ccp_set_gather_multicast(self.c._h, self.multicast_ptr)
        if rc != 0:
            raise RuntimeError(self.c._lib.ccp_last_error(self.c._h).decode())

    def detach(self):
        self.c._lib.ccp_set_gather_peers(self.c._h, 0, 0, None, 0)

    def exchange_counts(self, n_ok):
        """Exchange of the 8-byte converged counts (async on the current stream): a one-warp kernel stores this
        rank's count into every rank's count array, then the symmetric-memory group's device-side barrier — ordered
        after the projection kernels of this rank on the stream, so once it has passed, counts[r] and rank r's rows
        are both in place.  No collective-library call.  (use_nccl_counts = True: an NCCL all-gather instead.)"""
        import ctypes as C

        import torch
        import torch.distributed as dist

        if self.use_nccl_counts:
            dist.all_gather_into_tensor(self.counts, n_ok.view(1), group=self.group)
            return self.counts
        arr = (C.c_uint64 * self.world)(*self.count_ptrs)
        stream = torch.cuda.current_stream(self.pool.device).cuda_stream
        rc = self.c._lib.ccp_publish_count(self.c._h, n_ok.data_ptr(), self.world, self.rank, arr, stream)
        if rc != 0:
            raise RuntimeError(self.c._lib.ccp_last_error(self.c._h).decode())
        self.counts_handle.barrier(channel=0)
        return self.counts


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

    def __init__(self, constraint, group=None, fused: bool = True):
        import torch.distributed as dist

        self.c = constraint
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.fused = fused and 1 < self.world <= 8
        self._peer = None  # PeerPool, (re)allocated when the capacity grows
        self.fused_error = None

    def _peer_pool(self, cap):
        """The symmetric pool for the fused gather, or None when peer memory cannot be mapped (collective: every
        rank takes the same branch because the allocation either works on the whole node or on none of it)."""
        if not self.fused:
            return None
        if self._peer is None or self._peer.capacity < cap:
            try:
                self._peer = PeerPool(self.c, cap, self.group)
            except Exception as e:  # no P2P / symmetric memory on this box: NCCL gather instead
                self.fused = False
                self.fused_error = repr(e)
                self._peer = None
        return self._peer

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
        peer = self._peer_pool(cap) if self.world > 1 else None
        if peer is not None:
            # fused: the kernel stores the converged states into every rank's pool; no local compact buffer needed
            import torch.distributed as dist

            dist.barrier(group=self.group)  # nobody is still reading the pool of the previous call
            peer.attach()
            try:
                rc = c._lib.ccp_sample_project_batch(c._h, C.byref(args), count, _capi.CCP_LAYOUT_AOS, None, None, None,
                                                     None, n_ok.data_ptr(), stream)
            finally:
                peer.detach()
            if rc != 0:
                raise _capi.CcpError(c._lib.ccp_last_error(c._h).decode())
            counts = peer.exchange_counts(n_ok)
            return unpack_pool(peer.pool[:, :cap], counts), counts.clone()
        rc = c._lib.ccp_sample_project_batch(c._h, C.byref(args), count, _capi.CCP_LAYOUT_AOS, None, None, None,
                                             compact.data_ptr(), n_ok.data_ptr(), stream)
        if rc != 0:
            raise _capi.CcpError(c._lib.ccp_last_error(c._h).decode())
        if self.world == 1:
            k = int(n_ok.item())
            return compact[:k], torch.tensor([k], dtype=torch.int64, device=dev)
        pool, counts = gather_converged(compact, n_ok, cap, self.group)
        return unpack_pool(pool, counts), counts
