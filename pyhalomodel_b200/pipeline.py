"""
End-to-end evaluation from a linear spectrum (BASELINE config 5): for every sample of a chain
    P_lin(k) -> sigma(M) [top-hat quadrature] -> nu, g(nu), b(nu), A -> NFW / isothermal Fourier profiles [Si/Ci]
             -> HOD amplitudes and <N> normalisation -> all auto and cross spectra,
entirely on the device, in chunks sized for HBM.  Only a handful of scalars per sample cross PCIe (cosmology and HOD
parameters in; spectra out), or nothing at all when the caller keeps the results on the device.

Every stage is a libhalob200 kernel (hm_sigma_tophat, hm_window_nfw_isothermal, hm_average_batch,
hm_power_spectrum) except the element-wise vector preparations (radii, Duffy concentrations, Zheng05 occupation
numbers, BBKS transfer function), which are a few torch ops on [chunk, nM] / [chunk, nkg] tensors.

The reference's recipe for the same chain is tests/test_consistency.py:62-88 with cosmology.sigmaR(...,'brute')
(cosmology.py:89-106) in place of CAMB's sigma(R).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, constants, engine, halomodel, synthetic as syn

F64 = torch.float64


def _bbks_device(k, Om_m, h, ns, growth):
    """Unnormalised BBKS spectrum k^ns T(k)^2 D(z)^2 for a chunk: k [n], parameters [G] -> [G, n]."""
    q = k[None, :]/(Om_m*h)[:, None]
    T = torch.log(1.+2.34*q)/(2.34*q)*(1.+3.89*q+(16.1*q)**2+(5.46*q)**3+(6.71*q)**4)**(-0.25)
    return k[None, :]**ns[:, None]*T**2*(growth**2)[:, None]


class EndToEnd:
    """Chunked device pipeline for matter + galaxy spectra of a list of samples.

    samples: list of dict(Om_m, h, ns, sigma8, z, hod=dict(Mmin, sigma, M0, M1, alpha))
    families: mass functions evaluated per sample (they share the sample's grids)."""

    def __init__(self, k, M, families=('Tinker et al. (2010)',), chunk=256):
        self.kh, self.Mh = np.asarray(k, dtype=float), np.asarray(M, dtype=float)
        self.k, self.M = engine.to_dev(self.kh), engine.to_dev(self.Mh)
        self.families, self.chunk = tuple(families), int(chunk)
        kg, self.dlnk = syn.sigma_k_grid()
        self.kg = engine.to_dev(kg)
        self.S = 3

    def _chunk(self, samples):
        G, F = len(samples), len(self.families)
        dev = self.k.device
        col = lambda key: torch.tensor([s[key] for s in samples], dtype=F64, device=dev)
        Om_m, h, ns, s8, z = col('Om_m'), col('h'), col('ns'), col('sigma8'), col('z')
        growth = torch.tensor([syn.growth_cpt(s['z'], s['Om_m']) for s in samples], dtype=F64, device=dev)
        rhom = constants.rho_critical*Om_m
        # 1. linear spectrum on the sigma grid and on k (unnormalised), 2. sigma at R_L(M) and at 8 Mpc/h
        Pg = _bbks_device(self.kg, Om_m, h, ns, growth)
        Pk = _bbks_device(self.k, Om_m, h, ns, growth)
        RL = torch.pow(3.*self.M[None, :]/(4.*np.pi*rhom[:, None]), 1./3.)
        RL = RL-(RL*RL*RL-3.*self.M[None, :]/(4.*np.pi*rhom[:, None]))/(3.*RL*RL)       # one Newton step: cbrt
        R = torch.cat([RL, torch.full((G, 1), 8., dtype=F64, device=dev)], dim=1)
        sig_all = engine.sigma_tophat(self.kg, Pg, self.dlnk, R.contiguous())
        # sigma_8 normalisation at z=0: sigma scales with sqrt(amplitude) and with the growth factor
        fac = s8/(sig_all[:, -1]/growth)
        sigma = (sig_all[:, :-1]*fac[:, None]).contiguous()
        Pk = (Pk*(fac**2)[:, None]).contiguous()
        # 3. spherical-collapse fits, radii, concentrations (cosmology.py:169-188, pyhalomodel.py:713-719, :1104-1124)
        Omz = Om_m*(1.+z)**3/(Om_m*(1.+z)**3+(1.-Om_m))
        dc = ((3./20.)*(12.*np.pi)**(2./3.))*(1.+0.012299*torch.log10(Omz))
        x = Omz-1.
        Dv = (18.*np.pi**2+82.*x-39.*x**2)/Omz
        cb = torch.pow(Dv, 1./3.)
        cb = cb-(cb*cb*cb-Dv)/(3.*cb*cb)
        rv = (RL/cb[:, None]).contiguous()
        conc = (7.85*(self.M[None, :]/2e12)**(-0.081)*((1.+z)**(-0.71))[:, None]).contiguous()
        # 4. Zheng et al. (2005) occupation numbers and Bernoulli + Poisson variance (pyhalomodel.py:1177-1221)
        hod = lambda key: torch.tensor([s['hod'][key] for s in samples], dtype=F64, device=dev)[:, None]
        Mmin, sg, M0, M1, alpha = hod('Mmin'), hod('sigma'), hod('M0'), hod('M1'), hod('alpha')
        Mrow = self.M[None, :]
        Nc = 0.5*(1.+torch.erf(torch.log10(Mrow/Mmin)/sg))
        Ns = (((Mrow-M0) >= 0.).to(F64)*(Mrow-M0)/M1)**alpha
        N, V = (Nc+Ns).contiguous(), (Nc*(1.-Nc)+Ns).contiguous()
        # 5. model descriptors (host, array arithmetic: Tinker table interpolation etc., pyhalomodel.py:56-192)
        desc = halomodel.model_descriptors(z.cpu().numpy(), Om_m.cpu().numpy(), self.families, Dv.cpu().numpy(),
                                           dc.cpu().numpy())
        mdev = engine.pack_models(desc)
        rho_g = engine.average_batch(mdev, self.M, sigma, N, models_per_sample=F)
        rho_m = torch.as_tensor(np.ascontiguousarray(desc['rhom']), device=dev)
        # 6. Fourier profiles and spectra
        Mamp = self.M[None, :].expand(G, -1).contiguous()
        Um, Ug = engine.window_nfw_isothermal(self.k, rv, conc)       # both windows of the sample's radii in one launch
        tracers = [engine.Tracer(Um, Mamp, rho_m, mode=_lib.GRID_UK, mass_tracer=True),
                   engine.Tracer(Ug, N, rho_g, variance=V, mode=_lib.GRID_UK,
                                 discrete_tracer=True)]
        plan = engine.PowerSpectrumPlan(self.k, Pk, self.M, sigma, tracers, mdev, models_per_sample=F)
        out, lowmass = plan.run()
        return out, dict(sigma=sigma, Pk_lin=Pk, dc=dc, Dv=Dv, rv=rv, c=conc, N=N, V=V, rho_g=rho_g, models=desc)

    def run(self, samples, keep_on_device=True, return_inputs=False):
        """Spectra [len(samples)*F, 3 pairs (m-m, m-g, g-g), 3 terms, nk] for all samples."""
        outs, extras = [], []
        for i in range(0, len(samples), self.chunk):
            out, ex = self._chunk(samples[i:i+self.chunk])
            outs.append(out.clone() if keep_on_device else out.cpu())
            if return_inputs:
                extras.append(ex)
        res = torch.cat(outs, dim=0)
        return (res, extras) if return_inputs else res


def draw_samples(n, seed=4321):
    """Priors of BASELINE config 5: seed 4321, same priors as the sweep."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        cs = syn.draw_cosmology(rng)
        cs['hod'] = syn.draw_hod(rng)
        out.append(cs)
    return out



class EndToEndShard:
    """Adapter that lets sweep.ShardedSweep partition a chain over the ranks of one box: build(lo, hi) returns one
    of these for samples[lo:hi]; .plan().run() evaluates the rank's samples and returns their spectra."""

    def __init__(self, pipe: EndToEnd, samples):
        self.pipe, self.samples = pipe, samples
        F = len(pipe.families)
        self.out = torch.empty((len(samples)*F, 3, 3, len(pipe.kh)), dtype=F64, device=pipe.k.device)

    def plan(self):
        return self

    def run(self, out=None, peer=False):
        dst = self.out if out is None else out
        F = len(self.pipe.families)
        for i in range(0, len(self.samples), self.pipe.chunk):
            part, _ = self.pipe._chunk(self.samples[i:i+self.pipe.chunk])
            dst[i*F:i*F+part.shape[0]].copy_(part)
        return dst, None
