"""
Synthetic, analytic inputs for the halo-model hot path (host side, numpy, float64).

There is no CAMB in this environment (the reference's tests build their inputs with it,
tests/test_consistency.py:41,72), so the named BASELINE.json configurations are driven by an analytic linear
spectrum instead: BBKS transfer function, Carroll-Press-Turner growth, normalised to sigma_8 through the
top-hat sigma(R) of cosmology.py:89-106.  Grids follow the reference's tests: k = logspace(-3, 1, nk)
(tests/make_benchmarks.py:25-27) and M = logspace(9, 17, nM) (tests/test_consistency.py:24-26).

Nothing here is on the timed path; these are input generators shared by tests/, bench.py and
tests/golden/make_golden.py so that the CUDA path, the oracle and the live reference all see identical arrays.
"""
from __future__ import annotations

import numpy as np

FAMILY_NAMES = ('Press & Schecter (1974)', 'Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)',
                'Tinker et al. (2010)', 'Despali et al. (2016)')


def k_grid(nk, kmin=1e-3, kmax=1e1):
    return np.logspace(np.log10(kmin), np.log10(kmax), nk)


def M_grid(nM, Mmin=1e9, Mmax=1e17):
    return np.logspace(np.log10(Mmin), np.log10(Mmax), nM)


def bbks_transfer(k, Om_m, h):
    """Bardeen, Bond, Kaiser & Szalay (1986) CDM transfer function with shape Gamma = Om_m h (k in h/Mpc)."""
    q = np.asarray(k, dtype=float)/(Om_m*h)
    return np.log(1.+2.34*q)/(2.34*q)*(1.+3.89*q+(16.1*q)**2+(5.46*q)**3+(6.71*q)**4)**(-0.25)


def growth_cpt(z, Om_m):
    """Carroll, Press & Turner (1992) growth factor for flat LCDM, normalised to D(0)=1."""
    def g(Om, OL):
        return 2.5*Om/(Om**(4./7.)-OL+(1.+Om/2.)*(1.+OL/70.))
    Ez2 = Om_m*(1.+z)**3+(1.-Om_m)
    Omz, OLz = Om_m*(1.+z)**3/Ez2, (1.-Om_m)/Ez2
    return g(Omz, OLz)/(1.+z)/g(Om_m, 1.-Om_m)


def Om_m_of_z(z, Om_m):
    """tests/test_consistency.py:65"""
    return Om_m*(1.+z)**3/(Om_m*(1.+z)**3.+(1.-Om_m))


def tophat(x):
    """T(x) with the reference's Taylor switch at |x|<1e-5 (cosmology.py:41-48)."""
    x = np.asarray(x, dtype=float)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(np.abs(x) < 1e-5, 1.-x**2/10., (3./x**3)*(np.sin(x)-x*np.cos(x)))


def sigma_k_grid(kmin=1e-5, kmax=1e5, nk=100000):
    """Integration abscissae exactly as the reference builds them (utility.py:43, cosmology.py:99-100)."""
    k = np.logspace(np.log10(kmin), np.log10(kmax), nk)
    return k, np.log(k[1]/k[0])


def sigmaR_numpy(R, Pk_on_grid, kgrid, dlnk, chunk=64):
    """Vectorised host evaluation of the rectangle-rule sum of cosmology.py:103 (numpy pairwise sum)."""
    R = np.atleast_1d(np.asarray(R, dtype=float))
    w = dlnk*kgrid*(Pk_on_grid*kgrid**2)
    out = np.empty_like(R)
    for i in range(0, len(R), chunk):
        T = tophat(np.outer(R[i:i+chunk], kgrid))
        out[i:i+chunk] = np.sqrt(np.sum(w*T**2, axis=1)/(2.*np.pi**2))
    return out


class LinearSpectrum:
    """Callable P_lin(k) at z=0 (BBKS, sigma_8 normalised) plus its growth-scaled value at z."""

    def __init__(self, Om_m=0.3, h=0.7, ns=0.96, sigma8=0.8):
        self.Om_m, self.h, self.ns, self.sigma8 = Om_m, h, ns, sigma8
        self.amp = 1.
        kg, dlnk = sigma_k_grid()
        s8 = sigmaR_numpy(np.array([8.]), self(kg), kg, dlnk)[0]
        self.amp = (sigma8/s8)**2

    def __call__(self, k, z=0.):
        k = np.asarray(k, dtype=float)
        return self.amp*k**self.ns*bbks_transfer(k, self.Om_m, self.h)**2*growth_cpt(z, self.Om_m)**2


def zheng_defaults():
    return dict(Mmin=1e12, sigma=0.15, M0=1e12, M1=1e13, alpha=1.)


def draw_hod(rng):
    """Zheng et al. (2005) HOD parameter priors of SURVEY.md section 8(d)."""
    return dict(Mmin=10**rng.uniform(11.5, 13.5), sigma=rng.uniform(0.1, 0.6), M0=10**rng.uniform(11.5, 13.),
                M1=10**rng.uniform(12.5, 14.5), alpha=rng.uniform(0.8, 1.2))


def draw_cosmology(rng):
    """Priors of the batched sweep: Om_m, sigma8, ns, h, z."""
    return dict(Om_m=rng.uniform(.25, .35), sigma8=rng.uniform(.7, .9), ns=rng.uniform(.92, 1.),
                h=rng.uniform(.6, .8), z=rng.uniform(0., 2.))


def pressure_profile(k, M, rv):
    """Synthetic dimensionful 'electron pressure' Fourier profile for config 2:
    1e-20 M^(5/3) [1+(k rv/2)^2]^(-3/2), to be wrapped with profile.Fourier(amplitude=None)."""
    kv = np.outer(k, rv)
    return 1e-20*M**(5./3.)*(1.+(kv/2.)**2)**(-1.5)
