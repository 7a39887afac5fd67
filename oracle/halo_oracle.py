"""
oracle/halo_oracle.py -- CPU restatement of the pyhalomodel hot path (numpy/scipy, float64).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker (or as the
timed CPU baseline).  Nothing under `pyhalomodel_b200/` imports this module; the product path fails loudly
if the CUDA library is missing.

Parity pinning: this restatement is pinned against the *live* reference (imported from /root/reference in
the authoring container) by `tests/golden/make_golden.py`, which writes both the reference outputs and the
inputs into `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them through this file.  The
reference's own `.dat` goldens (tests/benchmarks/*.dat) cannot be replayed: they need CAMB, which is absent.

Every function cites the reference lines it restates (paths relative to /root/reference/).  The arithmetic
*order* of the reference is kept wherever it is observable in the last bits (e.g. `(amplitude*Uk)/norm`,
`amp*(amp-1.) + variance`), and the loop structure of `power_spectrum` is kept too (one scipy trapezoid per
(pair, k), mass function and bias re-evaluated inside every integral) because this file is also the timed
CPU baseline: it must cost what the reference costs.

Third-party arithmetic used exactly as the reference uses it (installed scipy 1.18.1 / numpy 2.3.5;
reference pins scipy ^1.10, numpy ^1.24 in pyproject.toml:16-18):
  scipy.integrate.trapezoid (pyhalomodel.py:27), scipy.integrate.quad (pyhalomodel.py:413),
  scipy.special.sici (pyhalomodel.py:1033,1044), scipy.special.gamma (pyhalomodel.py:98),
  scipy.special.erf (pyhalomodel.py:1181), scipy.interpolate.interp1d (pyhalomodel.py:152).
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass

import numpy as np
from scipy import integrate as _integrate
from scipy import special as _special
from scipy.interpolate import interp1d as _interp1d

# ---------------------------------------------------------------------------------------------
# constants.py:4-13 and cosmology.py:11-15,28-34
# ---------------------------------------------------------------------------------------------
_G_SI = 6.6743e-11
_MSUN_KG = 1.9884e30
_MPC_M = 3.0857e16*1e6
_H0 = 100.
_G_COSMO = _G_SI*_MSUN_KG/(_MPC_M*1e3**2)            # constants.py:12
RHO_CRITICAL = 3.*_H0**2/(8.*math.pi*_G_COSMO)        # constants.py:13
DV_EDS = 18.*np.pi**2                                 # cosmology.py:11
DC_EDS = (3./20.)*(12.*np.pi)**(2./3.)                # cosmology.py:13
TOPHAT_XMIN = 1e-5                                    # cosmology.py:15

PS74, ST99, SMT01, TINKER10, DESPALI16 = (
    'Press & Schecter (1974)', 'Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)',
    'Tinker et al. (2010)', 'Despali et al. (2016)')
FAMILIES = (PS74, ST99, SMT01, TINKER10, DESPALI16)


def matter_density(Om_m):
    """cosmology.py:28-34"""
    return RHO_CRITICAL*Om_m


def lagrangian_radius(M, Om_m):
    """cosmology.py:148-155"""
    return np.cbrt(3.*M/(4.*np.pi*matter_density(Om_m)))


def virial_radius(M, Om_m, Dv):
    """pyhalomodel.py:713-719"""
    return lagrangian_radius(M, Om_m)/np.cbrt(Dv)


def dc_nakamura_suto(Om_mz):
    """cosmology.py:169-175"""
    return DC_EDS*(1.+0.012299*np.log10(Om_mz))


def Dv_bryan_norman(Om_mz):
    """cosmology.py:178-188"""
    x = Om_mz-1.
    Dv = DV_EDS+82.*x-39.*x**2
    return Dv/Om_mz


# ---------------------------------------------------------------------------------------------
# model: family constants, g(nu), b(nu)                      pyhalomodel.py:56-192, 217-284
# ---------------------------------------------------------------------------------------------
_TINKER_DV = np.array([200., 300., 400., 600., 800., 1200., 1600, 2400., 3200.])
_TINKER_TAB = {  # pyhalomodel.py:142-151 (Tinker et al. 2010, table 4)
    'alpha': np.array([0.368, 0.363, 0.385, 0.389, 0.393, 0.365, 0.379, 0.355, 0.327]),
    'beta': np.array([0.589, 0.585, 0.544, 0.543, 0.564, 0.623, 0.637, 0.673, 0.702]),
    'gamma': np.array([0.864, 0.922, 0.987, 1.09, 1.20, 1.34, 1.50, 1.68, 1.81]),
    'phi': np.array([-0.729, -0.789, -0.910, -1.05, -1.20, -1.26, -1.45, -1.50, -1.49]),
    'eta': np.array([-0.243, -0.261, -0.261, -0.273, -0.278, -0.301, -0.301, -0.319, -0.336]),
}


@dataclass
class HaloModel:
    """Scalar state of one reference `model` instance (pyhalomodel.py:73-79,101-184)."""
    z: float
    Om_m: float
    name: str
    Dv: float
    dc: float
    rhom: float
    par: dict
    tinker_pbs: bool = False  # module switch Tinker_PBS, pyhalomodel.py:22


def make_model(z, Om_m, name=TINKER10, Dv=200., dc=1.686, tinker_z_dep=False, tinker_pbs=False) -> HaloModel:
    """pyhalomodel.py:56-192"""
    par = {}
    if name == PS74:
        pass
    elif name in (ST99, SMT01):
        p, q = 0.3, 0.707
        par.update(p=p, q=q, A=np.sqrt(2.*q)/(np.sqrt(np.pi)+_special.gamma(0.5-p)/2**p))  # :99-104
        if name == SMT01:
            par.update(a=0.707, b=0.5, c=0.6)                                              # :109-111
    elif name == DESPALI16:
        p, q = 0.2536, 0.7689
        par.update(p=p, q=q, A=0.3295*np.sqrt(2.*q/np.pi))                                 # :119-122
    elif name == TINKER10:
        lnDv, lnDv_tab = np.log(Dv), np.log(_TINKER_DV)
        for key, tab in _TINKER_TAB.items():                                               # :152-161
            par[key] = _interp1d(lnDv_tab, tab, fill_value='extrapolate')(lnDv)
        if tinker_z_dep:                                                                   # :163-167
            par['beta'] = par['beta']*(1.+z)**0.20
            par['gamma'] = par['gamma']*(1.+z)**-0.01
            par['phi'] = par['phi']*(1.+z)**-0.08
            par['eta'] = par['eta']*(1.+z)**0.27
        y = np.log10(Dv)                                                                   # :177-184
        e = np.exp(-(4./y)**4)
        par.update(bA=1.+0.24*y*e, ba=0.44*y-0.88, bB=0.183, bb=1.5, bC=0.019+0.107*y+0.19*e, bc=2.4)
    else:
        raise ValueError('Halo model not recognised')                                       # :190
    return HaloModel(z=z, Om_m=Om_m, name=name, Dv=Dv, dc=dc, rhom=matter_density(Om_m), par=par,
                     tinker_pbs=tinker_pbs)


def mass_function_nu(hm: HaloModel, nu):
    """g(nu), pyhalomodel.py:217-240"""
    P = hm.par
    if hm.name == PS74:
        return np.sqrt(2./np.pi)*np.exp(-(nu**2)/2.)
    if hm.name in (ST99, SMT01, DESPALI16):
        A, q, p = P['A'], P['q'], P['p']
        return A*(1.+((q*nu**2)**(-p)))*np.exp(-q*nu**2/2.)
    if hm.name == TINKER10:
        f1 = 1.+(P['beta']*nu)**(-2.*P['phi'])
        f2 = nu**(2.*P['eta'])
        f3 = np.exp(-P['gamma']*nu**2/2.)
        return P['alpha']*f1*f2*f3
    raise ValueError('Halo model not recognised in mass_function')


def linear_bias_nu(hm: HaloModel, nu):
    """b(nu), pyhalomodel.py:242-284"""
    P, dc = hm.par, hm.dc
    if hm.name == PS74:
        return 1.+(nu**2-1.)/dc
    if hm.name in (ST99, DESPALI16):
        p, q = P['p'], P['q']
        return 1.+(q*(nu**2)-1.+2.*p/(1.+(q*nu**2)**p))/dc
    if hm.name == SMT01:
        a, b, c = P['a'], P['b'], P['c']
        anu2 = a*nu**2
        f1 = np.sqrt(a)*anu2
        f2 = np.sqrt(a)*b*anu2**(1.-c)
        f3 = anu2**c
        f4 = anu2**c+b*(1.-c)*(1.-c/2.)
        return 1.+(f1+f2-f3/f4)/(dc*np.sqrt(a))
    if hm.name == TINKER10:
        if hm.tinker_pbs:                                                                  # :264-271
            f1 = (P['gamma']*nu**2-(1.+2.*P['eta']))/dc
            f2 = (2.*P['phi']/dc)/(1.+(P['beta']*nu)**(2.*P['phi']))
            return 1.+f1+f2
        fA = P['bA']*nu**P['ba']/(nu**P['ba']+dc**P['ba'])                                  # :273-282
        fB = P['bB']*nu**P['bb']
        fC = P['bC']*nu**P['bc']
        return 1.-fA+fB+fC
    raise ValueError('Halo model not recognised in linear_bias')


def peak_height(hm: HaloModel, sigmaM):
    """pyhalomodel.py:706-711"""
    return hm.dc/sigmaM


def low_mass_correction(hm: HaloModel, nu0):
    """A = 1 - int_{nu0}^{inf} g b dnu with QUADPACK, pyhalomodel.py:413-414"""
    return 1.-_integrate.quad(lambda x: mass_function_nu(hm, x)*linear_bias_nu(hm, x), nu0, np.inf)[0]


def derivative_from_samples(x, xs, fs):
    """Derivative at x of the quadratic through the three samples around it, utility.py:15-35 (scipy's `lagrange`
    polynomial in the monomial basis, differentiated: the reference's own arithmetic)."""
    from scipy.interpolate import lagrange
    ix = int(np.abs(xs-x).argmin())               # utility.find_closest_index_value, utility.py:50-56
    if ix == 0:
        imin, imax = (0, 1) if x < xs[0] else (0, 2)
    elif ix == len(xs)-1:
        nx = len(xs)
        imin, imax = (nx-2, nx-1) if x > xs[-1] else (nx-3, nx-1)
    else:
        imin, imax = ix-1, ix+1
    return lagrange(xs[imin:imax+1], fs[imin:imax+1]).deriv()(x)


def multiplicity_function(hm: HaloModel, M, sigmaM):
    """M^2 n(M)/rhobar = g(nu) dnu/dlnM, pyhalomodel.py:298-315"""
    nu = peak_height(hm, sigmaM)
    logR, logsig = np.log(lagrangian_radius(M, hm.Om_m)), np.log(sigmaM)
    deriv = np.array([2.*derivative_from_samples(x, logR, logsig) for x in logR])
    return mass_function_nu(hm, nu)*(-(nu/6.)*deriv)


def mass_function(hm: HaloModel, M, sigmaM):
    """n(M) = F rhom/M^2, pyhalomodel.py:286-296"""
    return multiplicity_function(hm, M, sigmaM)*hm.rhom/M**2


def is_increasing(x):
    """utility.py:63-67"""
    return np.all(np.diff(x) > 0.)


def average(hm: HaloModel, M, sigmaM, func):
    """<f> = rhom trapz((f/M) g, nu), pyhalomodel.py:335-359"""
    if not is_increasing(M):
        raise ValueError('Halo mass array must be increasing monotonically')
    nu = peak_height(hm, sigmaM)
    return _integrate.trapezoid((func/M)*mass_function_nu(hm, nu), nu)*hm.rhom


# ---------------------------------------------------------------------------------------------
# profile container                                           pyhalomodel.py:807-884
# ---------------------------------------------------------------------------------------------
class Profile:
    """Same fields and copy semantics as the reference `profile` (pyhalomodel.py:812-819)."""

    def __init__(self, k, M, Uk, Wk, amplitude=None, normalisation=1., variance=None,
                 mass_tracer=False, discrete_tracer=False):
        self.k, self.M = np.array(k, dtype=float), np.array(M, dtype=float)
        self.Uk, self.Wk = np.array(Uk, dtype=float), np.array(Wk, dtype=float)
        self.amplitude = None if amplitude is None else np.array(amplitude, dtype=float)
        self.normalisation = normalisation
        self.variance = None if variance is None else np.array(variance, dtype=float)
        self.mass_tracer, self.discrete_tracer = mass_tracer, discrete_tracer

    @classmethod
    def Fourier(cls, k, M, Uk, amplitude=None, normalisation=1., variance=None,
                mass_tracer=False, discrete_tracer=False):
        """pyhalomodel.py:853-884"""
        if Uk.shape != (k.shape[0], M.shape[0]):
            raise ValueError('k, M, array shapes do not match')
        if amplitude is None:                       # dimensionful input: amplitude is the lowest-k row
            amplitude = Uk[0, :]
            Wk = Uk.copy()
            U = Uk/amplitude
        else:
            U = Uk.copy()
            Wk = (amplitude*Uk)/normalisation
        return cls(k, M, U, Wk, amplitude=amplitude, normalisation=normalisation, variance=variance,
                   mass_tracer=mass_tracer, discrete_tracer=discrete_tracer)


WIN_INTEGRATION = 'romb'      # pyhalomodel.py:30-38 (sample-based choices only)
NR_WIN_INTEGRATION = 1+2**8    # pyhalomodel.py:38


def halo_window(k, M, rv, c, diff_profile):
    """pyhalomodel.py:930-961, sample-based branches"""
    R = np.linspace(0., rv, NR_WIN_INTEGRATION)
    integrand = _special.spherical_jn(0, k*R)*diff_profile(R, M, rv, c)
    if WIN_INTEGRATION == 'trapezoid':
        return _integrate.trapezoid(integrand, R)
    if WIN_INTEGRATION == 'simps':
        return _integrate.simpson(integrand, x=R)
    if WIN_INTEGRATION == 'romb':
        return _integrate.romb(integrand, R[1]-R[0])
    raise ValueError('Halo window function integration method not recognised')


def configuration_profile(k, M, rv, c, diff_profile, amplitude=None, normalisation=1., variance=None,
                          mass_tracer=False, discrete_tracer=False):
    """profile.configuration, pyhalomodel.py:886-925"""
    amp0 = np.array([halo_window(0., _M, _rv, _c, diff_profile) for _M, _rv, _c in zip(M, rv, c)])
    Uk = np.zeros((len(k), len(M)))
    for ik, _k in enumerate(k):
        for iM, (_M, _rv, _c) in enumerate(zip(M, rv, c)):
            Uk[ik, iM] = halo_window(_k, _M, _rv, _c, diff_profile)/amp0[iM]
    amp = amp0 if amplitude is None else np.array(amplitude, dtype=float)
    return Profile.Fourier(k, M, Uk, amplitude=amp, normalisation=normalisation, variance=variance,
                           mass_tracer=mass_tracer, discrete_tracer=discrete_tracer)


# ---------------------------------------------------------------------------------------------
# Fourier windows, concentration, HOD                          pyhalomodel.py:1006-1221
# ---------------------------------------------------------------------------------------------
def nfw_factor(c):
    """pyhalomodel.py:1058-1062"""
    return np.log(1.+c)-c/(1.+c)


def window_nfw(k, rv, c):
    """pyhalomodel.py:1040-1055"""
    rs = rv/c
    kv, ks = np.outer(k, rv), np.outer(k, rs)
    Si_sv, Ci_sv = _special.sici(ks+kv)
    Si_s, Ci_s = _special.sici(ks)
    t1 = np.cos(ks)*(Ci_sv-Ci_s)
    t2 = np.sin(ks)*(Si_sv-Si_s)
    t3 = np.sin(kv)/(ks+kv)
    return (t1+t2-t3)/nfw_factor(c)


def window_isothermal(k, rv):
    """pyhalomodel.py:1029-1037"""
    kv = np.outer(k, rv)
    Si, _ = _special.sici(kv)
    return Si/kv


def window_function(k, rv, *args, profile=None):
    """pyhalomodel.py:1006-1026"""
    if profile == 'delta':
        return np.ones((len(k), len(rv)))
    if profile == 'isothermal':
        return window_isothermal(k, rv)
    if profile == 'NFW':
        return window_nfw(k, rv, *args)
    raise ValueError('Halo profile not recognised')


def matter_profile(k, M, rv, c, Om_m):
    """pyhalomodel.py:1065-1077"""
    return Profile.Fourier(k, M, window_nfw(k, rv, c), amplitude=M, normalisation=matter_density(Om_m),
                           mass_tracer=True)


_DUFFY = {('M200', '200', '200b'): (10.14, -0.081, -1.01),
          ('vir', 'virial', 'Mvir'): (7.85, -0.081, -0.71),
          ('M200c', '200c'): (5.71, -0.084, -0.47)}


def concentration(M, z, method='Duffy et al. (2008)', halo_definition='M200'):
    """pyhalomodel.py:1084-1124"""
    if method != 'Duffy et al. (2008)':
        raise ValueError('Halo concentration not recognised')
    for names, (A, B, C) in _DUFFY.items():
        if halo_definition in names:
            return A*(M/2e12)**B*(1.+z)**C
    raise ValueError('Halo definition not recognised')


def hod_mean(M, method='Zheng et al. (2005)', **kw):
    """pyhalomodel.py:1131-1201"""
    if method == 'simple':
        Mmin, Msat, alpha = kw.get('Mmin', 1e12), kw.get('Msat', 1e13), kw.get('alpha', 1.)
        Nc = np.heaviside(M-Mmin, 1.)
        return Nc, Nc*(M/Msat)**alpha
    if method == 'Zehavi et al. (2004)':
        Mmin, M1, alpha = kw.get('Mmin', 1e12), kw.get('M1', 1e13), kw.get('alpha', 1.)
        return np.heaviside(M-Mmin, 1.), (M/M1)**alpha
    if method == 'Zheng et al. (2005)':
        Mmin, sigma, M0 = kw.get('Mmin', 1e12), kw.get('sigma', 0.15), kw.get('M0', 1e12)
        M1, alpha = kw.get('M1', 1e13), kw.get('alpha', 1.)
        Nc = np.heaviside(M-Mmin, 1.) if sigma == 0. else 0.5*(1.+_special.erf(np.log10(M/Mmin)/sigma))
        return Nc, (np.heaviside(M-M0, 1.)*(M-M0)/M1)**alpha
    if method == 'Zhai et al. (2017)':
        Mmin, sigma = kw.get('Mmin', 10**13.68), kw.get('sigma', 0.82)
        Msat, alpha, Mcut = kw.get('Msat', 10**14.87), kw.get('alpha', 0.41), kw.get('Mcut', 10**12.32)
        Nc = np.heaviside(M-Mmin, 1.) if sigma == 0. else 0.5*(1.+_special.erf(np.log10(M/Mmin)/sigma))
        return Nc, ((M/Msat)**alpha)*np.exp(-Mcut/M)
    raise ValueError('HOD method not recognised')


def hod_variance(p, lam, central_condition=False):
    """pyhalomodel.py:1204-1221"""
    vcc = p*(1.-p)
    if central_condition:
        return vcc, p*lam*(1.+lam*(1.-p)), lam*(1.-p)
    return vcc, lam, 0.


# ---------------------------------------------------------------------------------------------
# sigma(R)                                                     cosmology.py:41-48, 61-106
# ---------------------------------------------------------------------------------------------
def tophat_k(x):
    """cosmology.py:41-48"""
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(np.abs(x) < TOPHAT_XMIN, 1.-x**2/10., (3./x**3)*(np.sin(x)-x*np.cos(x)))


def sigma_grid(kmin=1e-5, kmax=1e5, nk=int(1e5)):
    """The reference's integration grid: utility.py:39-43, cosmology.py:99-100"""
    k = np.logspace(np.log10(kmin), np.log10(kmax), nk)
    return k, np.log(k[1]/k[0])


def sigmaR_brute(R, Pk, kmin=1e-5, kmax=1e5, nk=int(1e5), sequential=True):
    """cosmology.py:89-106.  `Pk` is a callable, as in the reference.  `sequential=True` keeps Python's
    left-to-right `sum` (cosmology.py:103); False uses numpy's pairwise sum (faster, ~1e-14 different)."""
    k, dlnk = sigma_grid(kmin, kmax, nk)
    Pk_k = Pk(k)
    out = np.empty(np.shape(R), dtype=float)
    flat, Rflat = out.reshape(-1), np.asarray(R, dtype=float).reshape(-1)
    for i, r in enumerate(Rflat):
        terms = dlnk*k*(Pk_k*(k**2)*tophat_k(k*r)**2)        # cosmology.py:86,103
        total = sum(terms) if sequential else np.sum(terms)
        flat[i] = np.sqrt(total/(2.*np.pi**2))
    return out


# ---------------------------------------------------------------------------------------------
# power_spectrum                                               pyhalomodel.py:361-544
# ---------------------------------------------------------------------------------------------
def _one_halo(hm, M, nu, WuWv):
    """pyhalomodel.py:526-533"""
    return _integrate.trapezoid(WuWv*mass_function_nu(hm, nu)/M, nu)*hm.rhom


def _two_halo_integral(hm, M, nu, W, mass, A):
    """pyhalomodel.py:535-544"""
    I = _integrate.trapezoid(W*linear_bias_nu(hm, nu)*mass_function_nu(hm, nu)/M, nu)
    if mass:
        I += A*W[0]/M[0]
    return I*hm.rhom


DO_I11 = True        # module switches of pyhalomodel.py:45-46
DO_I12_I21 = True


def _beta_integral(hm, beta, M, nu, Wu, Wv, mass_u, mass_v, A):
    """Non-linear halo bias double integral, pyhalomodel.py:546-571 (np.trapz is np.trapezoid in numpy >= 2)."""
    g, b = mass_function_nu(hm, nu), linear_bias_nu(hm, nu)
    integrand = beta*np.outer(Wu, Wv)*np.outer(g, g)*np.outer(b, b)/np.outer(M, M)
    integral = np.trapezoid(np.trapezoid(integrand, nu), nu)             # utility.trapz2d, utility.py:46-52
    if DO_I11 and mass_u and mass_v:
        integral += beta[0, 0]*(A**2)*Wu[0]*Wv[0]/M[0]**2
    if DO_I12_I21 and mass_u:
        integral += (A*Wu[0]/M[0])*np.trapezoid(beta[0, :]*Wv*g*b/M, nu)
    if DO_I12_I21 and mass_v:
        integral += (A*Wv[0]/M[0])*np.trapezoid(beta[:, 0]*Wu*g*b/M, nu)
    return integral*hm.rhom**2


def _two_halo(hm, Pk_lin, M, nu, Wu, Wv, mass_u, mass_v, A, beta=None):
    """pyhalomodel.py:515-524"""
    I_NL = 0. if beta is None else _beta_integral(hm, beta, M, nu, Wu, Wv, mass_u, mass_v, A)
    Iu = _two_halo_integral(hm, M, nu, Wu, mass_u, A)
    Iv = _two_halo_integral(hm, M, nu, Wv, mass_v, A)
    return (Iu*Iv+I_NL)*Pk_lin


def cross_halo_spectrum(hm: HaloModel, k, Pk_lin, M, sigmaM, profiles: dict, M_halo, beta=None, simple_twohalo=False,
                        k_trunc=None):
    """pyhalomodel.py:573-704: each tracer crossed with haloes of mass M_halo."""
    if simple_twohalo and (beta is not None):
        raise ValueError('A simple two-halo term is not compatible with non-linear halo bias')
    if not is_increasing(M):
        raise ValueError('Halo mass array must be increasing monotonically')
    nu = peak_height(hm, sigmaM)
    g, b = mass_function_nu(hm, nu), linear_bias_nu(hm, nu)
    A = 1.-_integrate.trapezoid(g*b, nu)                                              # :624-625
    nu_halo = _interp1d(np.log(M), nu, kind='cubic')(np.log(M_halo))                   # :631-632
    b_halo = linear_bias_nu(hm, nu_halo)
    P2, P1, PH = {}, {}, {}
    for name, prof in profiles.items():
        W_halo = np.array([_interp1d(np.log(M), prof.Wk[ik, :], kind='cubic')(np.log(M_halo)) for ik in range(len(k))])
        p2h, p1h = np.zeros_like(k), np.zeros_like(k)
        for ik in range(len(k)):
            Wrow = prof.amplitude/prof.normalisation if simple_twohalo else prof.Wk[ik, :]
            Iu = _two_halo_integral(hm, M, nu, Wrow, prof.mass_tracer, A)              # :688
            I_NL = 0.
            if beta is not None:                                                       # :691-704
                integral = np.trapezoid(beta[ik, :]*prof.Wk[ik, :]*g*b/M, nu)
                if prof.mass_tracer:
                    integral += A*beta[ik, 0]*prof.Wk[ik, 0]/M[0]
                I_NL = b_halo*integral*hm.rhom
            p2h[ik] = (b_halo*Iu+I_NL)*Pk_lin[ik]                                      # :689
            p1h[ik] = W_halo[ik]                                                       # :662
        if k_trunc is not None:
            p1h *= 1.-np.exp(-(k/k_trunc)**2)
        P2[name], P1[name] = p2h.copy(), p1h.copy()
        PH[name] = P2[name]+P1[name]
    return P2, P1, PH


def power_spectrum(hm: HaloModel, k, Pk_lin, M, sigmaM, profiles: dict, beta=None, simple_twohalo=False,
                   subtract_shotnoise=True, correct_discrete=True, k_trunc=None):
    """pyhalomodel.py:361-513.  Returns (Pk_2h, Pk_1h, Pk_hm) dictionaries keyed 'u-v'
    for every ordered pair; the lower-triangle keys alias the upper-triangle arrays (:503-507)."""
    if simple_twohalo and (beta is not None):
        raise ValueError('A simple two-halo term is not compatible with non-linear halo bias')
    if not is_increasing(M):
        raise ValueError('Halo mass array must be increasing monotonically')
    if not isinstance(profiles, dict):
        raise TypeError('profiles must be a dictionary')
    for prof in profiles.values():
        if (k != prof.k).all():
            raise ValueError('k arrays must all be identical to those in profiles')
        if (M != prof.M).all():
            raise ValueError('Mass arrays must be identical to those in profiles')
    nu = peak_height(hm, sigmaM)
    A = low_mass_correction(hm, nu[0])
    if A < 0.:
        warnings.warn('Warning: Mass function/bias correction is negative!', RuntimeWarning)
    shot = {name: (_one_halo(hm, M, nu, prof.amplitude/prof.normalisation**2) if prof.discrete_tracer else 0.)
            for name, prof in profiles.items()}                                            # :423-426
    P2, P1, PH = {}, {}, {}
    p2h, p1h = np.zeros_like(k), np.zeros_like(k)
    items = list(profiles.items())
    for iu, (nu_name, pu) in enumerate(items):
        for iv, (nv_name, pv) in enumerate(items):
            key, rkey = nu_name+'-'+nv_name, nv_name+'-'+nu_name
            if simple_twohalo:                                                              # :442-446
                p2h = _two_halo(hm, Pk_lin, M, nu, pu.amplitude/pu.normalisation, pv.amplitude/pv.normalisation,
                                pu.mass_tracer, pv.mass_tracer, A)
            if iu > iv:                                                                     # :503-507
                P2[key], P1[key], PH[key] = P2[rkey], P1[rkey], PH[rkey]
                continue
            for ik in range(len(Pk_lin)):                                                   # :453-480
                if not simple_twohalo:
                    p2h[ik] = _two_halo(hm, Pk_lin[ik], M, nu, pu.Wk[ik, :], pv.Wk[ik, :],
                                        pu.mass_tracer, pv.mass_tracer, A,
                                        beta=None if beta is None else beta[ik, :, :])
                if nu_name == nv_name:
                    if correct_discrete and pu.discrete_tracer:
                        fac = pu.amplitude*(pu.amplitude-1.)
                    else:
                        fac = pu.amplitude**2
                    if pu.variance is not None:
                        fac = fac+pu.variance
                    prod = fac*(pu.Uk[ik, :]/pu.normalisation)**2
                else:
                    prod = pu.Wk[ik, :]*pv.Wk[ik, :]
                p1h[ik] = _one_halo(hm, M, nu, prod)
            if (nu_name == nv_name) and pu.discrete_tracer:                                 # :484-491
                if correct_discrete and (not subtract_shotnoise):
                    p1h += shot[nu_name]
                elif (not correct_discrete) and subtract_shotnoise:
                    warnings.warn('Warning: Subtracting shot noise while not treating discreteness properly is bad',
                                  RuntimeWarning)
                    p1h[:] -= shot[nu_name]
            if k_trunc is not None:                                                         # :494-495
                p1h *= 1.-np.exp(-(k/k_trunc)**2)
            P2[key], P1[key] = p2h.copy(), p1h.copy()
            PH[key] = P2[key]+P1[key]
    return P2, P1, PH
