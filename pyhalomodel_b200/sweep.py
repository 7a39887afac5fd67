"""
Batched sweeps over independent (cosmology, redshift, mass-function) models, sharded across the GPUs of one box.

Every model of a sweep is independent (model.power_spectrum reads only its arguments, pyhalomodel.py:361-513), so
the batch axis is partitioned contiguously across ranks -- one process per GPU, launched with torchrun -- and
each rank builds and evaluates its own samples; nothing crosses the interconnect on the data path.  The only
collective is the gather of the result spectra [B_local, S, 3, nk] (NCCL all-gather over NVLink; gloo in the
CPU tests).  Shards may be uneven: they are padded to the largest shard for the collective and trimmed after.
"""
from __future__ import annotations

import numpy as np
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


class HostSweep:
    """Batched front end for sweeps whose inputs live on the HOST: per (cosmology, redshift) sample the vectors a user
    of the reference computes before calling power_spectrum -- P_lin(k), sigma(M), the halo radii and concentrations,
    the HOD occupation numbers and variances (tests/test_consistency.py:62-88) -- go in, and the matter / galaxy
    spectra of every sample under every requested mass function come back as one numpy array.  Only those vectors
    cross PCIe: the (k, M) Fourier windows (NFW matter profile; isothermal or NFW galaxies), the galaxy normalisation
    <N> per mass function and the spectra are computed on the device.

    The sweep is cut into chunks of `chunk` samples that alternate between two sets of device buffers: the upload of
    chunk i+1 (copy stream), the kernels of chunk i (compute stream) and the download of chunk i-1's spectra (second
    copy stream) overlap.  Uploads are asynchronous when the input arrays are page-locked (e.g. numpy views of
    torch.empty(...).pin_memory()); pageable arrays work too, with the driver's staged copies.

    Tracers: 'm' = pyhalomodel.matter_profile (:1065-1077), 'g' = profile.Fourier(U, amplitude=N, normalisation=<N>,
    variance=V, discrete_tracer=True).  Output pairs: m-m, m-g, g-g."""

    def __init__(self, k, M, families, chunk=512, galaxy_window='isothermal', flags=None):
        from . import _lib, engine, halomodel
        self.engine, self.halomodel, self._lib = engine, halomodel, _lib
        self.kh, self.Mh = np.asarray(k, dtype=float), np.asarray(M, dtype=float)
        self.k, self.M = engine.to_dev(self.kh), engine.to_dev(self.Mh)
        self.families, self.chunk, self.F = tuple(families), int(chunk), len(families)
        if galaxy_window not in ('isothermal', 'NFW'):
            raise ValueError('galaxy_window must be isothermal or NFW')
        self.galaxy_window = galaxy_window
        self.nk, self.nM = len(self.kh), len(self.Mh)
        dev, f64 = self.k.device, torch.float64
        C, B, nk, nM = self.chunk, self.chunk*self.F, self.nk, self.nM
        self.slots = []
        for _ in range(2):
            sl = dict(Pk=torch.ones((C, nk), dtype=f64, device=dev), sigma=torch.ones((C, nM), dtype=f64, device=dev),
                      rv=torch.ones((C, nM), dtype=f64, device=dev), c=torch.ones((C, nM), dtype=f64, device=dev),
                      N=torch.ones((C, nM), dtype=f64, device=dev), V=torch.ones((C, nM), dtype=f64, device=dev),
                      models=torch.zeros(B*halomodel.MODEL_DESC_DTYPE.itemsize, dtype=torch.uint8, device=dev),
                      hmodels=torch.zeros(B*halomodel.MODEL_DESC_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
                      Um=torch.empty((C, nk, nM), dtype=f64, device=dev), Ug=torch.empty((C, nk, nM), dtype=f64, device=dev),
                      norm_m=torch.ones(B, dtype=f64, device=dev), norm_g=torch.ones(B, dtype=f64, device=dev),
                      out=torch.empty((B, 3, 3, nk), dtype=f64, device=dev),
                      up=torch.cuda.Event(), done=torch.cuda.Event(), down=torch.cuda.Event())
            Mamp = self.M[None, :].expand(C, -1).contiguous()
            tracers = [engine.Tracer(sl['Um'], Mamp, sl['norm_m'], mode=_lib.GRID_UK, mass_tracer=True),
                       engine.Tracer(sl['Ug'], sl['N'], sl['norm_g'], variance=sl['V'], mode=_lib.GRID_UK, discrete_tracer=True)]
            sl['plan'] = engine.PowerSpectrumPlan(self.k, sl['Pk'], self.M, sl['sigma'], tracers, sl['models'],
                                                  models_per_sample=self.F, **(flags or {}))
            self.slots.append(sl)
        self.s_up, self.s_run, self.s_down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self._result = None

    def _result_buffer(self, n_models):
        if self._result is None or self._result.shape[0] < n_models:
            self._result = torch.empty((n_models, 3, 3, self.nk), dtype=torch.float64).pin_memory()
        return self._result[:n_models]

    def run(self, z, Om_m, Dv, dc, Pk_lin, sigmaM, rv, c, N, V):
        """z, Om_m, Dv, dc: [G]; Pk_lin: [G, nk]; sigmaM, rv, c, N, V: [G, nM] (numpy, float64, C-contiguous).
        Returns numpy [G*F, 3 pairs (m-m, m-g, g-g), 3 terms (2h, 1h, hm), nk], models sample-major.  The array is a
        view of a page-locked buffer that the next run() of this object overwrites."""
        engine, hm = self.engine, self.halomodel
        arrays = dict(Pk=Pk_lin, sigma=sigmaM, rv=rv, c=c, N=N, V=V)
        for key, a in arrays.items():
            a = np.asarray(a)
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.ndim != 2:
                raise ValueError('%s must be a C-contiguous float64 [G, n] array' % key)
            arrays[key] = a
        G = len(z)
        result = self._result_buffer(G*self.F)
        main = torch.cuda.current_stream()
        for st in (self.s_up, self.s_run, self.s_down):
            st.wait_stream(main)
        for i, lo in enumerate(range(0, G, self.chunk)):
            hi = min(G, lo+self.chunk)
            n = hi-lo
            sl = self.slots[i & 1]
            desc = hm.model_descriptors(z[lo:hi], Om_m[lo:hi], self.families, Dv[lo:hi], dc[lo:hi])
            sl['up'].synchronize()                       # the previous upload out of this slot's staging has finished
            sl['hmodels'][:desc.nbytes].numpy()[:] = desc.view(np.uint8).reshape(-1)
            with torch.cuda.stream(self.s_up):
                self.s_up.wait_event(sl['done'])         # the kernels that read this slot's inputs have finished
                for key, a in arrays.items():
                    sl[key][:n].copy_(torch.from_numpy(a[lo:hi]), non_blocking=True)
                sl['models'][:desc.nbytes].copy_(sl['hmodels'][:desc.nbytes], non_blocking=True)
                sl['up'].record(self.s_up)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(sl['up'])
                self.s_run.wait_event(sl['down'])        # the previous download of this slot's spectra has finished
                if self.galaxy_window == 'NFW':
                    engine.window_nfw(self.k, sl['rv'], sl['c'], out=sl['Um'])
                    sl['Ug'].copy_(sl['Um'])
                else:
                    engine.window_nfw_isothermal(self.k, sl['rv'], sl['c'], out_nfw=sl['Um'], out_iso=sl['Ug'])
                rec = sl['models'].view(torch.float64).view(-1, 15)       # hm_model_desc = 15 doubles; [2] is rhom
                sl['norm_m'].copy_(rec[:, 2])
                engine.average_batch(sl['models'], self.M, sl['sigma'], sl['N'], models_per_sample=self.F, out=sl['norm_g'])
                sl['plan'].run(out=sl['out'])
                sl['done'].record(self.s_run)
            with torch.cuda.stream(self.s_down):
                self.s_down.wait_event(sl['done'])
                result[lo*self.F:hi*self.F].copy_(sl['out'][:n*self.F], non_blocking=True)
                sl['down'].record(self.s_down)
        self.s_down.synchronize()
        main.wait_stream(self.s_run)
        return result.numpy()

    def bytes_per_sample(self):
        """(host-to-device, device-to-host) bytes per sample."""
        return 8*(self.nk+5*self.nM)+self.F*self.halomodel.MODEL_DESC_DTYPE.itemsize, 8*self.F*9*self.nk


class ShardedSweep:
    """Owns this rank's shard of a sweep: `build(lo, hi)` must return an object with `.plan()` (e.g. a
    workloads.Batch restricted to samples lo..hi-1).  run() evaluates the shard and returns the gathered
    [n_samples*F, S, 3, nk] spectra on every rank.  A rank may own no samples (more ranks than samples): it still
    takes part in every collective with a zero-length block, so nobody waits on it."""

    def __init__(self, n_samples: int, build, models_per_sample: int = 1, group=None, tail_shape=None):
        self.n_samples, self.F, self.group = n_samples, models_per_sample, group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(n_samples, rank, world)
        self.batch = build(self.lo, self.hi) if self.hi > self.lo else None
        self.plan = self.batch.plan() if self.batch is not None else None
        self.tail_shape = tuple(tail_shape) if tail_shape is not None else None     # (S, 3, nk): needed by empty shards
        self._peer, self._calls = None, 0

    def _tail(self):
        """(S, 3, nk) of one model's block.  Ranks without samples have no plan to read it from: they use the
        `tail_shape` given at construction, or ask the first rank (which always owns a sample)."""
        if self.plan is not None:
            return tuple(self.plan.out.shape[1:])
        if self.tail_shape is None:
            raise ValueError('a rank without samples needs tail_shape=(S, 3, nk)')
        return self.tail_shape

    def run_local(self):
        if self.plan is None:
            return None
        out, _ = self.plan.run()
        return out

    def run_to_root(self, root=0):
        """Like run(), but with the gather fused into the epilogue kernel: every rank's spectra are stored straight
        into the root's memory over NVLink (PeerGatherBuffer), no collective kernel runs, and only the root gets
        the [n_samples*F, S, 3, nk] result (the other ranks get None).  CUDA only; uneven shards are padded.
        Consecutive calls alternate between two slots of the peer buffer and every call ends with the root's copy
        out of its slot followed by a barrier, so no rank can store into a slot the root is still reading."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.run_local()
        rank = dist.get_rank(self.group)
        bounds = shard_bounds(self.n_samples, world)
        nmax = max(b-a for a, b in bounds)
        tail = self._tail()
        if self._peer is None or self._peer.root != root:
            self._peer = PeerGatherBuffer((nmax*self.F,)+tail, 2, root=root, group=self.group)
            self._calls = 0
        slot = self._calls % 2
        self._calls += 1
        n_mine = (self.hi-self.lo)*self.F
        if self.plan is not None:
            self.plan.run(out=self._peer.slot(slot)[:n_mine], peer=True)
        local = self._peer.finish()                      # stream sync + barrier: every rank's stores have landed
        res = None
        if rank == root:
            res = torch.cat([local[r, slot, :(b-a)*self.F] for r, (a, b) in enumerate(bounds)], dim=0)
            torch.cuda.current_stream().synchronize()    # the slot has been read before anybody may reuse it
        dist.barrier(group=self.group)
        return res

    def run(self):
        out = self.run_local()
        S, three, nk = self._tail()
        if out is None:     # no samples on this rank: a zero-length block keeps the collective aligned
            dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
            per_sample = torch.zeros((0, self.F, S, three, nk), dtype=torch.float64, device=dev)
        else:
            # models are sample-major, so gathering per-sample blocks keeps the global model order
            per_sample = out.reshape(self.hi-self.lo, self.F, S, three, nk)
        return gather_spectra(per_sample, self.n_samples, self.group).reshape(self.n_samples*self.F, S, three, nk)
