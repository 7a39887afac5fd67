"""
Batched sweeps over independent (cosmology, redshift, mass-function) models, sharded across the GPUs of one box.

Every model of a sweep is independent (model.power_spectrum reads only its arguments, pyhalomodel.py:361-513), so
the batch axis is partitioned contiguously across ranks -- one process per GPU, launched with torchrun -- and
each rank builds and evaluates its own samples; nothing crosses the interconnect on the data path.  The only
collective is the gather of the result spectra [B_local, S, 3, nk] (NCCL all-gather over NVLink; gloo in the
CPU tests).  Shards may be uneven: they are padded to the largest shard for the collective and trimmed after.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int):
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    bounds, lo = [], 0
    for r in range(world):
        hi = lo+base+(1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_range(n_items: int, rank: int, world: int):
    return shard_bounds(n_items, world)[rank]


def gather_spectra(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-rank results along dim 0 into global sample order.

    local: [n_local, ...] on this rank, where n_local = len(shard_range(n_items, rank, world)).
    Returns [n_items, ...] on every rank.  Works on CUDA tensors (nccl) and CPU tensors (gloo)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_items, world)
    lo, hi = bounds[rank]
    if local.shape[0] != hi-lo:
        raise ValueError('rank %d holds %d items, expected %d' % (rank, local.shape[0], hi-lo))
    nmax = max(b-a for a, b in bounds)
    send = local
    if hi-lo < nmax:        # pad to the common shard size required by all_gather_into_tensor
        pad = torch.zeros((nmax-(hi-lo),)+tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send = torch.cat([local, pad], dim=0)
    recv = torch.empty((world*nmax,)+tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(b-a == nmax for a, b in bounds):
        return recv
    return torch.cat([recv[r*nmax:r*nmax+(b-a)] for r, (a, b) in enumerate(bounds)], dim=0)


class PeerGatherBuffer:
    """Gather fused into the kernels: a symmetric (peer-mapped over NVLink/NVSwitch) results buffer
    [world, slots, *block] lives on every rank; `slot(i)` returns THIS rank's block of slot i inside the ROOT
    rank's copy, so a kernel given that tensor as its output (PowerSpectrumPlan.run(out=...)) stores its spectra
    straight into the root's memory with peer stores -- no collective kernel, nothing to overlap, and the SMs
    stay with the quadrature.  After every rank has synchronised its stream and passed a barrier, the root's
    `local` tensor holds everybody's results.  The NCCL path (gather_spectra) is the baseline it is checked against."""

    def __init__(self, block_shape, slots, dtype=torch.float64, root=0, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.root = dist.get_world_size(group), dist.get_rank(group), root
        self.shape = (self.world, slots)+tuple(block_shape)
        dev = torch.device('cuda', torch.cuda.current_device())
        self.local = symm_mem.empty(self.shape, dtype=dtype, device=dev)
        self.handle = symm_mem.rendezvous(self.local, group=self.group)
        self.on_root = self.handle.get_buffer(root, self.shape, dtype)      # the root's buffer, mapped here
        self.slots = slots

    def slot(self, i):
        return self.on_root[self.rank, i % self.slots]

    def finish(self):
        """Make every rank's peer stores visible on the root: stream sync + barrier."""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        return self.local if self.rank == self.root else None


class OverlappedGather:
    """Gather equal-sized per-rank result blocks on a side stream so the collective of step n overlaps the
    kernels of step n+1 (the spectra are tiny next to the grids: the cost to hide is NCCL launch latency).

    submit(out) snapshots `out` into one of two staging buffers on the compute stream and enqueues the all-gather
    on the communication stream; `latest()` synchronises and returns the most recent gathered tensor."""

    def __init__(self, like: torch.Tensor, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.comm = torch.cuda.Stream(device=like.device)
        self.stage = [torch.empty_like(like) for _ in range(2)]
        self.recv = [torch.empty((self.world*like.shape[0],)+tuple(like.shape[1:]), dtype=like.dtype,
                                 device=like.device) for _ in range(2)]
        self.done = [None, None]
        self.n = 0

    def submit(self, out: torch.Tensor):
        i = self.n & 1
        compute = torch.cuda.current_stream(out.device)
        if self.done[i] is not None:
            compute.wait_event(self.done[i])          # the gather that last used this slot has finished
        self.stage[i].copy_(out, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            dist.all_gather_into_tensor(self.recv[i], self.stage[i], group=self.group)
            self.done[i] = torch.cuda.Event()
            self.done[i].record(self.comm)
        self.n += 1

    def latest(self):
        if self.n == 0:
            return None
        i = (self.n-1) & 1
        torch.cuda.current_stream().wait_event(self.done[i])
        return self.recv[i]


class ShardedSweep:
    """Owns this rank's shard of a sweep: `build(lo, hi)` must return an object with `.plan()` (e.g. a
    workloads.Batch restricted to samples lo..hi-1).  run() evaluates the shard and returns the gathered
    [n_samples*F, S, 3, nk] spectra on every rank."""

    def __init__(self, n_samples: int, build, models_per_sample: int = 1, group=None):
        self.n_samples, self.F, self.group = n_samples, models_per_sample, group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(n_samples, rank, world)
        self.batch = build(self.lo, self.hi) if self.hi > self.lo else None
        self.plan = self.batch.plan() if self.batch is not None else None

    def run_local(self):
        if self.plan is None:
            return None
        out, _ = self.plan.run()
        return out

    def run_to_root(self, root=0):
        """Like run(), but with the gather fused into the epilogue kernel: every rank's spectra are stored straight
        into the root's memory over NVLink (PeerGatherBuffer), no collective kernel runs, and only the root gets
        the [n_samples*F, S, 3, nk] result (the other ranks get None).  CUDA only; uneven shards are padded."""
        if self.plan is None:
            raise RuntimeError('empty shard: more ranks than samples')
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.run_local()
        rank = dist.get_rank(self.group)
        bounds = shard_bounds(self.n_samples, world)
        nmax = max(b-a for a, b in bounds)
        tail = tuple(self.plan.out.shape[1:])
        if getattr(self, '_peer', None) is None or self._peer.root != root:
            self._peer = PeerGatherBuffer((nmax*self.F,)+tail, 1, root=root, group=self.group)
        n_mine = (self.hi-self.lo)*self.F
        self.plan.run(out=self._peer.slot(0)[:n_mine], peer=True)
        local = self._peer.finish()
        if rank != root:
            return None
        return torch.cat([local[r, 0, :(b-a)*self.F] for r, (a, b) in enumerate(bounds)], dim=0)

    def run(self):
        out = self.run_local()
        if out is None:
            raise RuntimeError('empty shard: more ranks than samples')
        # models are sample-major, so gathering per-sample blocks keeps the global model order
        S, nk = out.shape[1], out.shape[3]
        per_sample = out.reshape(self.hi-self.lo, self.F, S, 3, nk)
        return gather_spectra(per_sample, self.n_samples, self.group).reshape(self.n_samples*self.F, S, 3, nk)
