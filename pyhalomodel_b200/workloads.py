"""
Builders for the BASELINE.json configurations on synthetic inputs: they produce the device
tensors a `engine.PowerSpectrumPlan` consumes, using the package's own kernels for sigma(M) and the Fourier
windows, and can hand the same arrays back as numpy for the oracle / CPU baseline.

Shared by bench.py and the tests; not part of the reference's API surface.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, cosmology, engine, synthetic as syn
from .halomodel import HOD_mean, HOD_variance, concentration, model

TINKER10 = 'Tinker et al. (2010)'


class Batch:
    """A batch of G samples x F mass functions for P tracers on shared (k, M) grids, resident in HBM."""

    def __init__(self):
        self.k = self.M = None
        self.Pk_lin = self.sigma = None            # [G, nk], [G, nM]
        self.rv = self.c = None                    # [G, nM]
        self.tracers = []                          # engine.Tracer
        self.names = []
        self.models = []                           # halomodel.model, length B = G*F
        self.F = 1
        self.meta = []                             # per-sample dict (cosmology, hod)

    @property
    def G(self):
        return self.sigma.shape[0]

    @property
    def B(self):
        return len(self.models)

    def gram_plan(self, **flags):
        return engine.GramPlan(self.k, self.Pk_lin, self.M, self.sigma, self.tracers,
                               [m.descriptor() for m in self.models], models_per_sample=self.F, **flags)

    def plan(self, **flags):
        return engine.PowerSpectrumPlan(self.k, self.Pk_lin, self.M, self.sigma, self.tracers,
                                        [m.descriptor() for m in self.models], models_per_sample=self.F, **flags)


def _cosmo_inputs(cosmos, k, M):
    """Per-sample P_lin(k), sigma(M), dc, Dv from the analytic spectrum; sigma(M) on the device kernel."""
    kg, dlnk = syn.sigma_k_grid()
    Pk, Pg, R, dcDv = [], [], [], []
    for cs in cosmos:
        P0 = syn.LinearSpectrum(cs['Om_m'], cs['h'], cs['ns'], cs['sigma8'])
        Pk.append(P0(k, cs['z']))
        Pg.append(P0(kg, cs['z']))
        R.append(cosmology.Lagrangian_radius(M, cs['Om_m']))
        Omz = syn.Om_m_of_z(cs['z'], cs['Om_m'])
        dcDv.append((cosmology.dc_NakamuraSuto(Omz), cosmology.Dv_BryanNorman(Omz)))
    sigma = engine.sigma_tophat(kg, np.stack(Pg), dlnk, np.stack(R))
    return np.stack(Pk), sigma, dcDv


def build(G, nk, nM, tracers=('m', 'g', 'p'), families=(TINKER10,), seed=1234, fiducial_first=True):
    """Config 2 is build(G, 512, 2048); config 3 is build(G, 256, 1024, ('m','g'), all five families);
    config 1 is build(1, 128, 256, ('m',), ('Sheth & Tormen (1999)',))."""
    rng = np.random.default_rng(seed)
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    cosmos, hods = [], []
    for g in range(G):
        if g == 0 and fiducial_first:
            cosmos.append(dict(Om_m=0.3, sigma8=0.8, ns=0.96, h=0.7, z=0.))
            hods.append(syn.zheng_defaults())
        else:
            cosmos.append(syn.draw_cosmology(rng))
            hods.append(syn.draw_hod(rng))
    Pk, sigma, dcDv = _cosmo_inputs(cosmos, k, M)
    b = Batch()
    b.k, b.M, b.F = engine.to_dev(k), engine.to_dev(M), len(families)
    b.Pk_lin, b.sigma = engine.to_dev(Pk), sigma
    b.meta = [dict(cosmo=c, hod=h, dc=d[0], Dv=d[1]) for c, h, d in zip(cosmos, hods, dcDv)]
    for cs, (dc, Dv) in zip(cosmos, dcDv):
        for fam in families:
            b.models.append(model(cs['z'], cs['Om_m'], name=fam, Dv=Dv, dc=dc))
    Mh = M
    rv = np.stack([b.models[g*b.F].virial_radius(Mh) for g in range(G)])
    c = np.stack([concentration(Mh, cs['z'], halo_definition='Mvir') for cs in cosmos])
    b.rv, b.c = engine.to_dev(rv), engine.to_dev(c)
    Mdev = b.M[None, :].expand(G, -1).contiguous()
    for name in tracers:
        if name == 'm':        # matter_profile (pyhalomodel.py:1065-1077)
            norm = np.array([m.rhom for m in b.models])
            b.tracers.append(engine.Tracer(engine.window_nfw(b.k, b.rv, b.c), Mdev, norm, mode=_lib.GRID_UK,
                                           mass_tracer=True))
        elif name == 'g':      # Zheng et al. (2005) HOD galaxies, isothermal, Bernoulli+Poisson variance
            N, V = [], []
            for h in hods:
                Nc, Ns = HOD_mean(Mh, method='Zheng et al. (2005)', **h)
                Vc, Vs, _ = HOD_variance(Nc, Ns)
                N.append(Nc+Ns)
                V.append(Vc+Vs)
            N, V = engine.to_dev(np.stack(N)), engine.to_dev(np.stack(V))
            # <N> under every model of the batch in one launch (the same kernel arithmetic as model.average, bit for bit)
            norm = engine.average_batch(engine.pack_models([m.descriptor() for m in b.models]), b.M, b.sigma, N,
                                        models_per_sample=b.F)
            b.tracers.append(engine.Tracer(engine.window_isothermal(b.k, b.rv), N, norm, variance=V,
                                           mode=_lib.GRID_UK, discrete_tracer=True))
        elif name == 'p':      # synthetic dimensionful pressure profile, amplitude=None (pyhalomodel.py:874-878)
            kv = b.k[None, :, None]*b.rv[:, None, :]
            grid = 1e-20*Mdev[:, None, :]**(5./3.)*(1.+(kv/2.)**2)**(-1.5)
            b.tracers.append(engine.Tracer(grid.contiguous(), None, np.ones(len(b.models)), mode=_lib.GRID_WK))
        else:
            raise ValueError('unknown tracer '+name)
        b.names.append(name)
    return b


def build_many_tracer(n_tracers=64, nk=256, nM=2048, seed=4, family=TINKER10):
    """BASELINE config 4: `n_tracers` Zheng05 HOD galaxy samples (NFW profile, Bernoulli+Poisson variance, discrete)
    on one cosmology; returns a Batch whose .gram_plan() evaluates all n(n+1)/2 unique pairs."""
    rng = np.random.default_rng(seed)
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    cs = dict(Om_m=0.3, sigma8=0.8, ns=0.96, h=0.7, z=0.)
    Pk, sigma, dcDv = _cosmo_inputs([cs], k, M)
    b = Batch()
    b.k, b.M, b.F = engine.to_dev(k), engine.to_dev(M), 1
    b.Pk_lin, b.sigma = engine.to_dev(Pk), sigma
    hmod = model(cs['z'], cs['Om_m'], name=family, Dv=dcDv[0][1], dc=dcDv[0][0])
    b.models = [hmod]
    b.meta = [dict(cosmo=cs, dc=dcDv[0][0], Dv=dcDv[0][1], hods=[])]
    b.rv = engine.to_dev(hmod.virial_radius(M))[None]
    b.c = engine.to_dev(concentration(M, cs['z'], halo_definition='Mvir'))[None]
    U = engine.window_nfw(b.k, b.rv, b.c)
    for i in range(n_tracers):
        hod = syn.draw_hod(rng)
        Nc, Ns = HOD_mean(M, method='Zheng et al. (2005)', **hod)
        Vc, Vs, _ = HOD_variance(Nc, Ns)
        N, V = engine.to_dev(Nc+Ns)[None], engine.to_dev(Vc+Vs)[None]
        norm = engine.average(hmod.descriptor(), b.M, b.sigma[0], N[0])
        # every tracer owns its grid, as separate profile objects do in the reference API
        b.tracers.append(engine.Tracer(U.clone(), N, norm, variance=V, mode=_lib.GRID_UK, discrete_tracer=True))
        b.names.append('g%d' % i)
        b.meta[0]['hods'].append(hod)
    return b


def host_profiles(batch: Batch, g):
    """Sample g of a batch as host (numpy) inputs for a reference-shaped API (the oracle module
    or pyhalomodel_b200.halomodel).  Returns (k, Pk_lin, M, sigmaM, grids) with grids[name] = dict of numpy arrays."""
    k, M = engine.to_host(batch.k), engine.to_host(batch.M)
    out = {}
    for name, tr in zip(batch.names, batch.tracers):
        d = dict(grid=engine.to_host(tr.grid[g]))
        if tr.amplitude is not None:
            d['amplitude'] = engine.to_host(tr.amplitude[g])
        if tr.variance is not None:
            d['variance'] = engine.to_host(tr.variance[g])
        out[name] = d
    return k, engine.to_host(batch.Pk_lin[g]), M, engine.to_host(batch.sigma[g]), out
