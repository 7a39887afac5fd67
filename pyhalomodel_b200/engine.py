"""
Device-side engine: thin, batch-capable wrappers over the C ABI (include/halob200.h) working on torch float64
CUDA tensors.  The reference-shaped API in `halomodel.py` and the sweep driver in `sweep.py` both sit on this.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libhalob200.so.  Nothing in this
module can run without a CUDA device and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GramProblem, HaloB200Error, ModelDesc, Problem, check, lib

F64 = torch.float64


_HAVE_CUDA = None


def device():
    global _HAVE_CUDA
    if _HAVE_CUDA is None:
        _HAVE_CUDA = bool(torch.cuda.is_available())
    if not _HAVE_CUDA:
        raise HaloB200Error('no CUDA device visible: pyhalomodel_b200 runs on B200 only and has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


_CUDART = None
DEFAULT_STACK_LIMIT = 1024


def stack_limit(nbytes=None):
    """Get (and with `nbytes`, set) cudaLimitStackSize of the current device.  NCCL raises it (1024 -> 1872 bytes on this
    stack) when its first communicator is created, and the driver raises it whenever a kernel needs more; a raised
    limit costs the HBM-bound quadrature kernels 2-3.5 % (tools/dev_stack_probe.py) and it is
    never lowered again on its own.  Sweeps that mix collectives with the quadrature call
    `stack_limit(DEFAULT_STACK_LIMIT)` after the collective; a later kernel that needs more stack gets it again."""
    global _CUDART
    if _CUDART is None:
        for name in ('libcudart.so.12', 'libcudart.so'):
            try:
                _CUDART = C.CDLL(name)
                break
            except OSError:
                continue
        if _CUDART is None:
            raise HaloB200Error('libcudart is not loadable: cannot query cudaLimitStackSize')
    old = C.c_size_t(0)
    if _CUDART.cudaDeviceGetLimit(C.byref(old), 0) != 0:
        raise HaloB200Error('cudaDeviceGetLimit(cudaLimitStackSize) failed')
    if nbytes is not None and int(nbytes) != old.value:
        torch.cuda.synchronize()
        if _CUDART.cudaDeviceSetLimit(0, C.c_size_t(int(nbytes))) != 0:
            raise HaloB200Error('cudaDeviceSetLimit(cudaLimitStackSize, %d) failed' % int(nbytes))
    return old.value


def is_host(x):
    """True if `x` lives on the host (numpy array, scalar, list, CPU tensor)."""
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def to_dev(x, shape=None):
    """numpy / list / scalar / torch tensor  ->  contiguous float64 CUDA tensor (a copy only when needed)."""
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == F64 and x.is_contiguous():
        if shape is not None and tuple(x.shape) != tuple(shape):
            raise ValueError('expected shape %s, got %s' % (tuple(shape), tuple(x.shape)))
        return x
    dev = device()
    if isinstance(x, torch.Tensor):
        t = x.to(device=dev, dtype=F64, non_blocking=True)
    else:
        a = np.asarray(x, dtype=np.float64)
        if a.ndim and not a.flags.c_contiguous:      # (np.ascontiguousarray would promote scalars to 1-d)
            a = np.ascontiguousarray(a)
        t = torch.from_numpy(a).to(dev, non_blocking=True)
    t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError('expected shape %s, got %s' % (tuple(shape), tuple(t.shape)))
    return t


def to_dev_many(arrays):
    """Upload several small host arrays with ONE host-to-device copy (latency matters more than bytes for the
    per-call vectors k, P_lin, M, sigma(M)); entries that already live on the device pass through."""
    host = [(i, np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1)) for i, a in enumerate(arrays)
            if not (isinstance(a, torch.Tensor) and a.is_cuda)]
    out = [to_dev(a) if (isinstance(a, torch.Tensor) and a.is_cuda) else None for a in arrays]
    if host:
        flat = torch.from_numpy(np.concatenate([h for _, h in host])).to(device(), non_blocking=True)
        off = 0
        for i, h in host:
            out[i] = flat[off:off+h.size].reshape(np.shape(arrays[i]))
            off += h.size
    return out


def to_host(t):
    return t.detach().cpu().numpy()


_RAW_STREAM = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream():
    # (torch.cuda.current_stream() builds a Stream object through several Python layers: 20 us per kernel launch)
    if _RAW_STREAM is not None:
        return C.c_void_p(_RAW_STREAM(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def pack_models(descs):
    """list of ModelDesc, or numpy record array of hm_model_desc (halomodel.model_descriptors) -> uint8 CUDA tensor
    holding the packed hm_model_desc array."""
    if isinstance(descs, np.ndarray):
        if descs.dtype.itemsize != C.sizeof(ModelDesc):
            raise ValueError('record array does not have the layout of hm_model_desc')
        host = torch.from_numpy(np.ascontiguousarray(descs).view(np.uint8).reshape(-1))
    else:
        arr = (ModelDesc*len(descs))(*descs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device())


# ---------------------------------------------------------------------------------------------------------------
def massfn_nu(desc: ModelDesc, nu, want_g=True, want_b=True):
    nu = to_dev(nu)
    g = torch.empty_like(nu) if want_g else None
    b = torch.empty_like(nu) if want_b else None
    check(lib().hm_massfn_nu(C.byref(desc), _ptr(nu), nu.numel(), _ptr(g), _ptr(b), _stream()))
    return g, b


def low_mass_correction(descs, nu0):
    nu0 = to_dev(nu0).reshape(-1)
    models = pack_models(descs)
    A = torch.empty_like(nu0)
    check(lib().hm_low_mass_correction(_ptr(models), _ptr(nu0), nu0.numel(), _ptr(A), _stream()))
    return A


def average(desc: ModelDesc, M, sigma, funcs):
    """funcs: [nf, nM] -> [nf] halo averages (model.average, pyhalomodel.py:335-359)."""
    M, sigma, funcs = to_dev_many([M, sigma, funcs])      # (host arrays go up in one copy)
    if funcs.dim() == 1:
        funcs = funcs[None, :]
    nf, nM = funcs.shape
    out = torch.empty(nf, dtype=F64, device=M.device)
    check(lib().hm_average(C.byref(desc), _ptr(M), _ptr(sigma), nM, _ptr(funcs), nf, _ptr(out), _stream()))
    return out


def average_batch(models_dev, M, sigma, funcs, models_per_sample=1, out=None):
    """sigma, funcs: [G, nM]; models_dev: packed descriptors of the G*F models -> [G*F] halo averages."""
    M, sigma, funcs = to_dev(M), to_dev(sigma), to_dev(funcs)
    G, nM = sigma.shape
    if out is None:
        out = torch.empty(G*models_per_sample, dtype=F64, device=M.device)
    check(lib().hm_average_batch(_ptr(models_dev), _ptr(M), _ptr(sigma), nM, G, models_per_sample, _ptr(funcs),
                                 _ptr(out), _stream()))
    return out


def fp64_peaks():
    """(DFMA, DMMA) TFLOP/s of the current device, measured now (roofline denominators of the FP64-bound kernels)."""
    device()
    f, m = C.c_double(0.), C.c_double(0.)
    check(lib().hm_measure_fp64_peaks(C.byref(f), C.byref(m), _stream()))
    return f.value, m.value


def multiplicity_function(desc: ModelDesc, M, sigma, want_n=False):
    """M [nM], sigma [nM] -> F = M^2 n(M)/rhobar, or n(M) itself (pyhalomodel.py:286-315)."""
    M, sigma = to_dev(M), to_dev(sigma)
    out = torch.empty_like(M)
    check(lib().hm_multiplicity_function(C.byref(desc), _ptr(M), _ptr(sigma), M.shape[0], None if want_n else _ptr(out),
                                         _ptr(out) if want_n else None, _stream()))
    return out


def cross_halo_spectrum(desc: ModelDesc, k, Pk_lin, M, sigma, tracers, interp_weights, beta=None, simple_twohalo=False,
                        k_trunc=None):
    """tracers: engine.Tracer with [nk, nM] grids and scalar normalisations -> (out [P, 3, nk], [A, nu_h, b_h])."""
    k, Pk_lin, M, sigma, L = to_dev(k), to_dev(Pk_lin), to_dev(M), to_dev(sigma), to_dev(interp_weights)
    nk, nM, P = k.shape[0], M.shape[0], len(tracers)
    ct = (_lib.Tracer*P)()
    norms = (C.c_double*P)()
    keep = []
    for p, t in enumerate(tracers):
        grid = to_dev(t.grid, (nk, nM))
        keep.append(grid)
        ct[p].grid_dev = grid.data_ptr()
        if t.mode == _lib.GRID_SEPARATE:
            wk = to_dev(t.wk, (nk, nM))
            keep.append(wk)
            ct[p].wk_dev = wk.data_ptr()
        if t.amplitude is not None:
            amp = to_dev(t.amplitude, (nM,))
            keep.append(amp)
            ct[p].amplitude_dev = amp.data_ptr()
        ct[p].grid_mode, ct[p].mass_tracer, ct[p].discrete_tracer = int(t.mode), int(t.mass_tracer), int(t.discrete_tracer)
        norms[p] = float(t.normalisation)
    beta = to_dev(beta, (nk, nM)) if beta is not None else None
    out = torch.empty((P, 3, nk), dtype=F64, device=k.device)
    scal = torch.empty(3, dtype=F64, device=k.device)
    nbytes = lib().hm_cross_halo_workspace(nM)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=k.device)
    check(lib().hm_cross_halo_spectrum(C.byref(desc), _ptr(k), _ptr(Pk_lin), _ptr(M), _ptr(sigma), nk, nM, ct, norms, P,
                                       _ptr(L), _ptr(beta), int(bool(simple_twohalo)),
                                       float(k_trunc) if k_trunc is not None else 0., _ptr(out), _ptr(scal), _ptr(ws),
                                       nbytes, _stream()))
    return out, scal


def weighted_rowsum(grid, w):
    """grid [nk, nM], w [nw, nM] -> [nw, nk] with out[j, k] = sum_m w[j, m] grid[k, m]."""
    grid, w = to_dev(grid), to_dev(w)
    if w.dim() == 1:
        w = w[None, :]
    nk, nM = grid.shape
    out = torch.empty((w.shape[0], nk), dtype=F64, device=grid.device)
    check(lib().hm_weighted_rowsum(_ptr(grid), _ptr(w), nk, nM, w.shape[0], _ptr(out), _stream()))
    return out


def sici(x):
    x = to_dev(x)
    si, ci = torch.empty_like(x), torch.empty_like(x)
    check(lib().hm_sici(_ptr(x), x.numel(), _ptr(si), _ptr(ci), _stream()))
    return si, ci


def window_nfw(k, rv, c, out=None):
    """k [nk]; rv, c [nM] or [G, nM]  ->  U [nk, nM] or [G, nk, nM] (pyhalomodel.py:1040-1055)."""
    k, rv, c = to_dev(k), to_dev(rv), to_dev(c)
    batched = rv.dim() == 2
    G, nM = (rv.shape if batched else (1, rv.shape[0]))
    if out is None:
        out = torch.empty((G, k.shape[0], nM), dtype=F64, device=k.device)
    check(lib().hm_window_nfw(_ptr(k), _ptr(rv), _ptr(c), k.shape[0], nM, G, _ptr(out), _stream()))
    return out if batched else out[0]


def window_nfw_isothermal(k, rv, c, out_nfw=None, out_iso=None):
    """Both windows of the same radii in one launch: (window_nfw(k, rv, c), window_isothermal(k, rv)), the matter + galaxy
    pair of a sweep (the isothermal element shares the loop and the sine of k rv with the NFW one)."""
    k, rv, c = to_dev(k), to_dev(rv), to_dev(c)
    batched = rv.dim() == 2
    G, nM = (rv.shape if batched else (1, rv.shape[0]))
    if out_nfw is None:
        out_nfw = torch.empty((G, k.shape[0], nM), dtype=F64, device=k.device)
    if out_iso is None:
        out_iso = torch.empty((G, k.shape[0], nM), dtype=F64, device=k.device)
    check(lib().hm_window_nfw_isothermal(_ptr(k), _ptr(rv), _ptr(c), k.shape[0], nM, G, _ptr(out_nfw), _ptr(out_iso), _stream()))
    return (out_nfw, out_iso) if batched else (out_nfw[0], out_iso[0])


def window_isothermal(k, rv, out=None):
    k, rv = to_dev(k), to_dev(rv)
    batched = rv.dim() == 2
    G, nM = (rv.shape if batched else (1, rv.shape[0]))
    if out is None:
        out = torch.empty((G, k.shape[0], nM), dtype=F64, device=k.device)
    check(lib().hm_window_isothermal(_ptr(k), _ptr(rv), k.shape[0], nM, G, _ptr(out), _stream()))
    return out if batched else out[0]


def window_configuration(k, rv, samples, weights):
    """k [nk], rv [nM], samples [nM, nr] = P(R_j; M) on R = linspace(0, rv, nr), weights [nr] -> W [nk, nM]."""
    k, rv, samples, weights = to_dev(k), to_dev(rv), to_dev(samples), to_dev(weights)
    nM, nr = samples.shape
    out = torch.empty((k.shape[0], nM), dtype=F64, device=k.device)
    check(lib().hm_window_configuration(_ptr(k), _ptr(rv), _ptr(samples), _ptr(weights), k.shape[0], nM, nr, _ptr(out),
                                        _stream()))
    return out


def sigma_tophat(kgrid, Pk, dlnk, R):
    """kgrid [nkg]; Pk [nkg] or [G, nkg]; R [nR] or [G, nR] -> sigma with R's shape (cosmology.py:89-106)."""
    kgrid, Pk, R = to_dev(kgrid), to_dev(Pk), to_dev(R)
    batched = R.dim() == 2
    R2 = R if batched else R[None, :]
    P2 = Pk if Pk.dim() == 2 else Pk[None, :].expand(R2.shape[0], -1).contiguous()
    G, nR = R2.shape
    out = torch.empty_like(R2)
    for g0 in range(0, G, 65535):
        g1 = min(G, g0+65535)
        check(lib().hm_sigma_tophat(_ptr(kgrid), _ptr(P2[g0:g1]), float(dlnk), kgrid.shape[0], _ptr(R2[g0:g1]), nR,
                                    g1-g0, _ptr(out[g0:g1]), _stream()))
    return out if batched else out[0]


# ---------------------------------------------------------------------------------------------------------------
class Tracer:
    """Device-side description of one halo profile for a batch (mirrors hm_tracer)."""
    __slots__ = ('grid', 'wk', 'amplitude', 'variance', 'normalisation', 'mode', 'mass_tracer', 'discrete_tracer')

    def __init__(self, grid, amplitude, normalisation, variance=None, wk=None, mode=_lib.GRID_UK,
                 mass_tracer=False, discrete_tracer=False):
        self.grid, self.wk, self.amplitude, self.variance = grid, wk, amplitude, variance
        self.normalisation, self.mode = normalisation, mode
        self.mass_tracer, self.discrete_tracer = bool(mass_tracer), bool(discrete_tracer)


class PowerSpectrumPlan:
    """A validated hm_problem plus its workspace and output buffers, reusable across calls (no allocation and
    no host work in the steady state; CUDA-graph friendly)."""

    def __init__(self, k, Pk_lin, M, sigma, tracers, models, models_per_sample=1, simple_twohalo=False,
                 subtract_shotnoise=True, correct_discrete=True, k_trunc=None, beta=None, do_I11=True,
                 do_I12_I21=True):
        P = len(tracers)
        # per-call vectors and scalar normalisations travel in one host-to-device copy
        scalar_norms = [t.normalisation for t in tracers if np.isscalar(t.normalisation)]
        self.k, self.M, self.Pk_lin, self.sigma, norms_dev = to_dev_many(
            [k, M, Pk_lin, sigma, np.asarray(scalar_norms, dtype=np.float64)])
        nk, nM = self.k.shape[0], self.M.shape[0]
        if self.Pk_lin.dim() == 1:
            self.Pk_lin = self.Pk_lin[None, :]
        if self.sigma.dim() == 1:
            self.sigma = self.sigma[None, :]
        G = self.sigma.shape[0]
        if tuple(self.Pk_lin.shape) != (G, nk) or self.sigma.shape[1] != nM:
            raise ValueError('Pk_lin must be [G, nk] and sigma [G, nM]')
        n_scalar = 0
        self.beta = None
        if beta is not None:    # beta_NL(k, M1, M2) per sample (pyhalomodel.py:372, :546-571)
            self.beta = to_dev(beta)
            if self.beta.dim() == 3:
                self.beta = self.beta[None]
            if tuple(self.beta.shape) != (G, nk, nM, nM):
                raise ValueError('beta must be [G, nk, nM, nM]')
            if P > 6:
                raise NotImplementedError('beta_NL is supported for up to 6 tracers per call')
        if not 1 <= P <= _lib.HM_MAX_TRACERS:
            raise ValueError('between 1 and %d tracers are supported by the quadrature path' % _lib.HM_MAX_TRACERS)
        B = G*models_per_sample
        self.tracers = tracers
        self.G, self.B, self.P, self.S, self.nk, self.nM = G, B, P, P*(P+1)//2, nk, nM
        self.keep = []          # tensors the C struct borrows
        pr = Problem()
        pr.nk, pr.nM, pr.n_tracers, pr.n_samples, pr.models_per_sample = nk, nM, P, G, models_per_sample
        pr.simple_twohalo, pr.subtract_shotnoise = int(bool(simple_twohalo)), int(bool(subtract_shotnoise))
        pr.correct_discrete = int(bool(correct_discrete))
        pr.k_trunc = float(k_trunc) if k_trunc is not None else 0.
        pr.k_dev, pr.M_dev = self.k.data_ptr(), self.M.data_ptr()
        pr.sigma_dev, pr.Pk_lin_dev = self.sigma.data_ptr(), self.Pk_lin.data_ptr()
        if isinstance(models, torch.Tensor):
            self.models_dev = models
            pr.models_dev = models.data_ptr()
        elif len(models) == 1 and B == 1:
            self.models_dev = None
            pr.models_dev = None
            pr.model_host = models[0]
        else:
            if len(models) != B:
                raise ValueError('need one model descriptor per model (%d), got %d' % (B, len(models)))
            self.models_dev = pack_models(models)
            pr.models_dev = self.models_dev.data_ptr()
        for p, t in enumerate(tracers):
            grid = to_dev(t.grid)
            grid = grid if grid.dim() == 3 else grid[None]
            if tuple(grid.shape) != (G, nk, nM):
                raise ValueError('tracer %d: grid must be [G, nk, nM]' % p)
            ct = pr.tracers[p]
            ct.grid_dev = grid.data_ptr()
            self.keep.append(grid)
            if t.mode == _lib.GRID_SEPARATE:
                wk = to_dev(t.wk)
                wk = wk if wk.dim() == 3 else wk[None]
                ct.wk_dev = wk.data_ptr()
                self.keep.append(wk)
            if t.amplitude is not None:
                amp = to_dev(t.amplitude).reshape(G, nM)
                ct.amplitude_dev = amp.data_ptr()
                self.keep.append(amp)
            elif t.mode != _lib.GRID_WK:
                raise ValueError('tracer %d: amplitude is required' % p)
            if t.variance is not None:
                var = to_dev(t.variance).reshape(G, nM)
                ct.variance_dev = var.data_ptr()
                self.keep.append(var)
            norm = t.normalisation
            if np.isscalar(norm):
                norm = norms_dev[n_scalar:n_scalar+1]
                n_scalar += 1
                norm = norm if B == 1 else norm.expand(B).contiguous()
            else:
                norm = to_dev(norm).reshape(B)
            ct.normalisation_dev = norm.data_ptr()
            self.keep.append(norm)
            ct.grid_mode, ct.mass_tracer, ct.discrete_tracer = int(t.mode), int(t.mass_tracer), int(t.discrete_tracer)
        if self.beta is not None:
            pr.beta_dev = self.beta.data_ptr()
        pr.beta_do_I11, pr.beta_do_I12_I21 = int(bool(do_I11)), int(bool(do_I12_I21))
        self.problem = pr
        nbytes = lib().hm_power_spectrum_workspace(C.byref(pr))
        dev = self.k.device
        self.workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        self.workspace_bytes = nbytes
        self.out = torch.empty((B, self.S, 3, nk), dtype=F64, device=dev)
        self.lowmass = torch.empty(B, dtype=F64, device=dev)

    def weights(self):
        check(lib().hm_halo_weights(C.byref(self.problem), _ptr(self.lowmass), _ptr(self.workspace),
                                    self.workspace_bytes, _stream()))

    def quadrature(self, stages=3):
        """stages: bit 0 = streaming kernel, bit 1 = epilogue kernel (both by default)."""
        check(lib().hm_mass_quadrature_stages(C.byref(self.problem), _ptr(self.out), _ptr(self.workspace),
                                              self.workspace_bytes, _stream(), int(stages)))

    def run(self, out=None, peer=False):
        """Enqueue both stages on the current stream; returns (out [B,S,3,nk], A [B]) device tensors.  `out` may be
        a caller-owned contiguous float64 CUDA tensor of that shape (e.g. a slot of a sweep's results buffer);
        peer=True says it lives on another GPU (sweep.PeerGatherBuffer), so the epilogue fences system-wide."""
        dst = self.out if out is None else out
        if out is not None and (tuple(out.shape) != tuple(self.out.shape) or out.dtype != F64 or not out.is_contiguous()):
            raise ValueError('out must be a contiguous float64 tensor of shape %s' % (tuple(self.out.shape),))
        self.problem.output_is_peer = int(bool(peer))
        check(lib().hm_power_spectrum(C.byref(self.problem), _ptr(dst), _ptr(self.lowmass), _ptr(self.workspace),
                                      self.workspace_bytes, _stream()))
        return dst, self.lowmass

    def weight_vectors(self):
        """Per-mass weight vectors the quadrature kernel reads: P + S per model, or -- sweeps of 2..5 mass functions on
        one to three tracers, where the weights are kept factored (hm_quad.cu) -- 2 F + 2 P per sample."""
        F = self.B//self.G
        rows = (2 <= F <= 5 and self.P <= 3 and self.nM % 2 == 0 and self.beta is None
                and all(t.mode != _lib.GRID_SEPARATE for t in self.tracers))
        return self.G*(2*F+2*self.P) if rows else self.B*(self.P+self.S)

    def stream_kernel_bytes(self):
        """Algorithmic bytes of the streaming quadrature kernel alone: every distinct grid once plus its weight
        vectors once (its partial-sum writes are an implementation artefact and not counted)."""
        ngrids = sum(2 if t.mode == _lib.GRID_SEPARATE else 1 for t in self.tracers)
        return 8*(self.G*ngrids*self.nk*self.nM + self.weight_vectors()*self.nM)

    def algorithmic_bytes(self):
        """Bytes the quadrature kernel must move per launch: every distinct grid once, the weight vectors once
        per model, P_lin once per sample, three output vectors per unique pair."""
        ngrids = sum(2 if t.mode == _lib.GRID_SEPARATE else 1 for t in self.tracers)
        return 8*(self.G*ngrids*self.nk*self.nM + self.weight_vectors()*self.nM + self.G*self.nk
                  + self.B*3*self.S*self.nk)


class PipelinedPlans:
    """Software pipelining of consecutive steps of a sweep: two plans (double-buffered workspaces and outputs)
    alternate between two CUDA streams, so the small weights and epilogue kernels of one step run beside the
    persistent streaming kernel of the neighbouring step (their CTAs fit in the registers and the shared memory the
    streaming CTAs leave free on every SM).  submit() enqueues one step; drain() joins the streams.  Every submit
    orders its step behind whatever the caller has enqueued on the current stream so far (refreshed inputs, a
    consumer of the previous result); the object drains itself when it is dropped or leaves a `with` block."""

    def __init__(self, make_plan):
        self.plans = [make_plan(), make_plan()]
        self.streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        self.n = 0
        self.forked = False

    def submit(self, out=None, peer=False, then=None):
        """Enqueue one step; `then(out, lowmass)`, if given, is called with the step's stream current (e.g. to
        enqueue a copy of the spectra behind the epilogue kernel)."""
        main = torch.cuda.current_stream()
        i = self.n & 1
        self.streams[i].wait_stream(main)
        self.forked = True
        with torch.cuda.stream(self.streams[i]):
            res = self.plans[i].run(out=out, peer=peer)
            if then is not None:
                then(*res)
        self.n += 1
        return res

    def drain(self):
        if not self.forked:
            return
        main = torch.cuda.current_stream()
        for st in self.streams:
            main.wait_stream(st)
        self.forked = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.drain()
        return False

    def __del__(self):
        try:        # the plans' buffers were allocated on the caller's stream: join before the allocator recycles them
            self.drain()
        except Exception:
            pass


class GramPlan:
    """Many-tracer counterpart of PowerSpectrumPlan (2..64 single-grid tracers): the cross one-halo terms run as a
    per-wavenumber Gram contraction on the FP64 tensor cores (csrc/hm_gram.cu).  Same output layout."""

    def __init__(self, k, Pk_lin, M, sigma, tracers, models, models_per_sample=1, simple_twohalo=False,
                 subtract_shotnoise=True, correct_discrete=True, k_trunc=None):
        self.k, self.M = to_dev(k), to_dev(M)
        nk, nM = self.k.shape[0], self.M.shape[0]
        self.Pk_lin, self.sigma = to_dev(Pk_lin), to_dev(sigma)
        if self.Pk_lin.dim() == 1:
            self.Pk_lin = self.Pk_lin[None, :]
        if self.sigma.dim() == 1:
            self.sigma = self.sigma[None, :]
        G = self.sigma.shape[0]
        P = len(tracers)
        if not 2 <= P <= _lib.HM_GRAM_MAX_TRACERS:
            raise ValueError('the Gram path takes between 2 and %d tracers' % _lib.HM_GRAM_MAX_TRACERS)
        B = G*models_per_sample
        self.tracers = tracers
        self.G, self.B, self.P, self.S, self.nk, self.nM = G, B, P, P*(P+1)//2, nk, nM
        self.keep = []
        pr = GramProblem()
        pr.nk, pr.nM, pr.n_tracers, pr.n_samples, pr.models_per_sample = nk, nM, P, G, models_per_sample
        pr.simple_twohalo, pr.subtract_shotnoise = int(bool(simple_twohalo)), int(bool(subtract_shotnoise))
        pr.correct_discrete = int(bool(correct_discrete))
        pr.k_trunc = float(k_trunc) if k_trunc is not None else 0.
        pr.k_dev, pr.M_dev = self.k.data_ptr(), self.M.data_ptr()
        pr.sigma_dev, pr.Pk_lin_dev = self.sigma.data_ptr(), self.Pk_lin.data_ptr()
        if isinstance(models, torch.Tensor):
            self.models_dev = models
            pr.models_dev = models.data_ptr()
        elif len(models) == 1 and B == 1:
            self.models_dev, pr.models_dev, pr.model_host = None, None, models[0]
        else:
            self.models_dev = pack_models(models)
            pr.models_dev = self.models_dev.data_ptr()
        self.ctracers = (_lib.Tracer*P)()
        for p, t in enumerate(tracers):
            if t.mode not in (_lib.GRID_UK, _lib.GRID_WK):
                raise ValueError('tracer %d: the Gram path needs profiles built by profile.Fourier' % p)
            grid = to_dev(t.grid)
            grid = grid if grid.dim() == 3 else grid[None]
            if tuple(grid.shape) != (G, nk, nM):
                raise ValueError('tracer %d: grid must be [G, nk, nM]' % p)
            ct = self.ctracers[p]
            ct.grid_dev = grid.data_ptr()
            self.keep.append(grid)
            if t.amplitude is not None:
                amp = to_dev(t.amplitude).reshape(G, nM)
                ct.amplitude_dev = amp.data_ptr()
                self.keep.append(amp)
            if t.variance is not None:
                var = to_dev(t.variance).reshape(G, nM)
                ct.variance_dev = var.data_ptr()
                self.keep.append(var)
            norm = t.normalisation
            norm = to_dev(np.full(B, float(norm))) if np.isscalar(norm) else to_dev(norm).reshape(B)
            ct.normalisation_dev = norm.data_ptr()
            self.keep.append(norm)
            ct.grid_mode, ct.mass_tracer, ct.discrete_tracer = int(t.mode), int(t.mass_tracer), int(t.discrete_tracer)
        pr.tracers = self.ctracers
        self.problem = pr
        nbytes = lib().hm_gram_workspace(C.byref(pr))
        dev = self.k.device
        self.workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        self.workspace_bytes = nbytes
        self.out = torch.empty((B, self.S, 3, nk), dtype=F64, device=dev)
        self.lowmass = torch.empty(B, dtype=F64, device=dev)

    def run(self):
        check(lib().hm_gram_power_spectrum(C.byref(self.problem), _ptr(self.out), _ptr(self.lowmass),
                                           _ptr(self.workspace), self.workspace_bytes, _stream()))
        return self.out, self.lowmass

    def gram_only(self):
        """The Gram (DMMA) kernel alone, on the weights of the last run() (for timing the contraction)."""
        check(lib().hm_gram_power_spectrum_stages(C.byref(self.problem), _ptr(self.out), _ptr(self.lowmass),
                                                  _ptr(self.workspace), self.workspace_bytes, _stream(), 2))

    def algorithmic_bytes(self):
        """Every grid once, four per-(tracer, mass) vectors once per model, three output vectors per pair."""
        return 8*(self.G*self.P*self.nk*self.nM + self.B*(2+2*self.P)*self.nM + self.G*self.nk
                  + self.B*3*self.S*self.nk)

    def gram_flops(self):
        """Useful flops of the cross one-halo contraction: 2 per (unique pair incl. diagonal blocks, mass, k)."""
        return 2.*self.B*self.S*self.nM*self.nk
