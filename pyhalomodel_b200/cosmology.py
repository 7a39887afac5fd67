"""
Cosmology helpers of the hot path with the reference's names (pyhalomodel/cosmology.py): background density,
Lagrangian radii, spherical-collapse fits, and the top-hat sigma(R) whose 1e5-point log-k quadrature runs in the
sm_100a kernel `hm_sigma_tophat` (csrc/hm_sigma.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from . import constants as const
from . import engine
from . import utility as util

Dv0 = 18.*np.pi**2                          # EdS virial overdensity ~178 (cosmology.py:11)
dc0 = (3./20.)*(12.*np.pi)**(2./3.)         # EdS collapse threshold ~1.686 (cosmology.py:13)
xmin_Tk = 1e-5                              # Taylor switch of the top-hat window (cosmology.py:15); fixed in the kernel


def redshift_from_scalefactor(a):
    return -1.+1./a


def scalefactor_from_redshift(z):
    return 1./(1.+z)


def comoving_matter_density(Om_m: float) -> float:
    '''Comoving matter density [Msun/h / (Mpc/h)^3] (cosmology.py:28-34)'''
    return const.rho_critical*Om_m


def _cbrt(x):
    return _torch_cbrt(x) if isinstance(x, torch.Tensor) else np.cbrt(x)


def _torch_cbrt(x):
    # torch has no cbrt: Newton-polish pow(x, 1/3) once so the result is correctly rounded like np.cbrt in practice
    y = torch.pow(x, 1./3.)
    return y-(y*y*y-x)/(3.*y*y)


def Lagrangian_radius(M, Om_m: float):
    '''Radius [Mpc/h] of a sphere containing mass M in a homogeneous universe (cosmology.py:148-155)'''
    return _cbrt(3.*M/(4.*np.pi*comoving_matter_density(Om_m)))


def mass(R, Om_m: float):
    '''Mass [Msun/h] within a sphere of radius R [Mpc/h] (cosmology.py:158-162)'''
    return (4./3.)*np.pi*R**3*comoving_matter_density(Om_m)


def dc_NakamuraSuto(Om_mz: float) -> float:
    '''Nakamura & Suto (1997) fit for the linear collapse threshold (cosmology.py:169-175)'''
    return dc0*(1.+0.012299*np.log10(Om_mz))


def Dv_BryanNorman(Om_mz: float) -> float:
    '''Bryan & Norman (1998) virial overdensity relative to the mean matter density (cosmology.py:178-188)'''
    x = Om_mz-1.
    Dv = Dv0+82.*x-39.*x**2
    return Dv/Om_mz


def _Tophat_k(x):
    '''Fourier transform of a top hat with the reference's Taylor switch (cosmology.py:41-48); host helper.'''
    x = np.asarray(x, dtype=float)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(np.abs(x) < xmin_Tk, 1.-x**2/10., (3./x**3)*(np.sin(x)-x*np.cos(x)))


def sigmaR(Rs, Pk_lin: callable, kmin=1e-5, kmax=1e5, nk=int(1e5), integration_type='brute'):
    '''
    sigma(R) of the linear density field (cosmology.py:61-106).  `Pk_lin` is a callable evaluated ONCE on the
    reference's integration grid (built on the host exactly as utility.py:43 / cosmology.py:99-100 do, so the
    abscissae are bit-identical); the nR x nk top-hat sum runs on the device (csrc/hm_sigma.cu).

    Two stated departures from the reference's arithmetic for T(x) = 3(sin x - x cos x)/x^3 (cosmology.py:41-48):
    below |x| = 2 the Maclaurin series is used instead of the closed form, which cancels catastrophically there (so
    radii whose whole integrand sits at small x agree with the reference only as well as the reference's own
    rounding allows: 1e-7 at R = 1e-9); and terms with |x| = k R > 5e4 are dropped (T^2 < 4e-19 there).  For the
    spectra this package is used with (k^3 P(k) growing no faster than logarithmically at high k) the second changes
    sigma by less than 2e-13; a user-supplied P(k) that is much bluer than that at k R > 5e4 would lose those terms.
    '''
    if integration_type != 'brute':
        raise NotImplementedError("only integration_type='brute' is on the B200 path (adaptive quad is out of scope)")
    on_dev = isinstance(Rs, torch.Tensor) and Rs.is_cuda
    kgrid = util.logspace(kmin, kmax, nk)
    dlnk = np.log(kgrid[1]/kgrid[0])
    Pk = np.asarray(Pk_lin(kgrid), dtype=float)
    R = engine.to_dev(Rs)
    shape = R.shape
    out = engine.sigma_tophat(kgrid, Pk, dlnk, R.reshape(-1)).reshape(shape)
    if on_dev:
        return out
    out = engine.to_host(out)
    return out if out.ndim else float(out)
