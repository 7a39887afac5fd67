"""
Reference-shaped API of the B200 halo-model engine: `model`, `profile` and the module functions of
pyhalomodel/pyhalomodel.py, with the same names, arguments, dictionary conventions, exceptions and warnings, and
the arithmetic done by libhalob200.so (hand-written sm_100a CUDA, fp64) on torch CUDA tensors.

    import pyhalomodel_b200 as halo
    hmod = halo.model(z, Om_m, name='Tinker et al. (2010)', Dv=330., dc=1.686)
    Pk_2h, Pk_1h, Pk_hm = hmod.power_spectrum(k, Pk_lin, M, sigmaM, {'m': matter, 'g': galaxies})

Array arguments may be numpy arrays or torch tensors.  Results come back as numpy arrays unless an input of the
call was a CUDA tensor, in which case they stay on the device.  There is no CPU fallback.

Module-level switches are read at call time, as in the reference (pyhalomodel.py:16-46).
"""
from __future__ import annotations

import warnings
from time import time

import numpy as np
import torch

from . import _lib, constants, cosmology, engine
from . import utility as util
from ._lib import ModelDesc

# Parameters (same names as pyhalomodel.py:16-46)
dc_rel_tol = 1e-3
Dv_rel_tol = 1e-3
Tinker_z_dep = False    # redshift-dependent Tinker et al. (2010) parameters (pyhalomodel.py:20, :163-167)
Tinker_PBS = False      # peak-background-split bias for Tinker et al. (2010) (pyhalomodel.py:22, :264-271)
halo_integration = 'trapezoid'    # mass integrator (pyhalomodel.py:27): the trapezoid rule is the one built on the device
# W(k) integration scheme in r for configuration-space profiles (pyhalomodel.py:30-39).  The reference's default,
# adaptive scipy quad with epsrel=1e-3, cannot be reproduced bit for bit (and its code path is broken under
# scipy >= 1.14); the three sample-based schemes it offers are: 'trapezoid', 'simps', 'romb'.
win_integration = 'romb'
nr_win_integration = 1+2**8   # number of points in r; romb needs 2^m+1

# beta_NL (pyhalomodel.py:44-46)
do_I11 = True      # add the low-M1 and low-M2 portion of the beta_NL integral
do_I12_I21 = True  # add the low-M1 or low-M2 portions

_FAMILY = {'Press & Schecter (1974)': _lib.PS74, 'Sheth & Tormen (1999)': _lib.ST99,
           'Sheth, Mo & Tormen (2001)': _lib.SMT01, 'Tinker et al. (2010)': _lib.TINKER10,
           'Despali et al. (2016)': _lib.DESPALI16}

_TINKER_DV = np.array([200., 300., 400., 600., 800., 1200., 1600, 2400., 3200.])
_TINKER_TABLE4 = dict(   # Tinker et al. (2010) table 4, as tabulated at pyhalomodel.py:142-151
    alpha=[0.368, 0.363, 0.385, 0.389, 0.393, 0.365, 0.379, 0.355, 0.327],
    beta=[0.589, 0.585, 0.544, 0.543, 0.564, 0.623, 0.637, 0.673, 0.702],
    gamma=[0.864, 0.922, 0.987, 1.09, 1.20, 1.34, 1.50, 1.68, 1.81],
    phi=[-0.729, -0.789, -0.910, -1.05, -1.20, -1.26, -1.45, -1.50, -1.49],
    eta=[-0.243, -0.261, -0.261, -0.273, -0.278, -0.301, -0.301, -0.319, -0.336])


def _lin_interp_extrap(x, xs, ys):
    """Piecewise-linear interpolation with linear extrapolation from the end segments, in the two-weight form
    the installed scipy interp1d(fill_value='extrapolate') evaluates (so the Tinker parameters of
    pyhalomodel.py:152-161 are reproduced to the bit): w_hi*y_hi + w_lo*y_lo on the bracketing/end segment."""
    xs, ys = np.asarray(xs, dtype=float), np.asarray(ys, dtype=float)
    hi = int(np.clip(np.searchsorted(xs, x), 1, len(xs)-1))
    lo = hi-1
    return ((x-xs[lo])/(xs[hi]-xs[lo]))*ys[hi]+((xs[hi]-x)/(xs[hi]-xs[lo]))*ys[lo]


def _check_halo_integration():
    """The reference binds `halo_integration` to a scipy.integrate function and reads it at call time
    (pyhalomodel.py:27, :359, :531, :540); only the trapezoid rule exists on the device, anything else is refused."""
    hi = halo_integration
    if hi == 'trapezoid' or getattr(hi, '__name__', None) in ('trapezoid', 'trapz'):
        return
    raise NotImplementedError("halo_integration: only the trapezoid rule is implemented on the B200 path")


def _wants_device(*arrays):
    return any(isinstance(a, torch.Tensor) and a.is_cuda for a in arrays)


def _own(x):
    """Device copy of a profile input that the profile object owns (the reference copies every array it is given,
    pyhalomodel.py:814-818).  Host arrays are uploaded -- the upload IS the copy; only a caller-owned CUDA tensor
    needs a second buffer."""
    t = engine.to_dev(x)
    return t.clone() if (isinstance(x, torch.Tensor) and x.is_cuda and t.data_ptr() == x.data_ptr()) else t


def _give_back(t, on_device):
    return t if on_device else engine.to_host(t)


class model():
    '''
    Halo model: mass function g(nu), linear bias b(nu) and halo-model power spectra (pyhalomodel.py:51-719)
    '''

    def __init__(self, z: float, Om_m: float, name='Tinker et al. (2010)', Dv=200., dc=1.686, verbose=False):
        '''
        Args (identical to the reference, pyhalomodel.py:56-71):
            z: Redshift
            Om_m: Cosmological matter density
            name: 'Press & Schecter (1974)' | 'Sheth & Tormen (1999)' | 'Sheth, Mo & Tormen (2001)' |
                  'Tinker et al. (2010)' | 'Despali et al. (2016)'
            Dv: Halo overdensity definition
            dc: Halo linear collapse threshold
        '''
        self.z = z
        self.a = cosmology.scalefactor_from_redshift(z)
        self.Om_m = Om_m
        self.name = name
        self.dc = dc
        self.Dv = Dv
        self.rhom = cosmology.comoving_matter_density(Om_m)
        if verbose:
            print('Initialising halo model')
            print('Halo model:', self.name)
            print('Redshift: %1.3f' % (self.z))
            print('Scale factor: %1.3f' % (self.a))
            print('Omega_m(z=0): %1.3f' % (self.Om_m))
            print('delta_c: %1.4f' % (self.dc))
            print('Delta_v: %4.1f' % (self.Dv))
        if name not in _FAMILY:
            raise ValueError('Halo model not recognised')
        if name in ('Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)'):
            from scipy.special import gamma as Gamma
            self.p_ST, self.q_ST = 0.3, 0.707
            self.A_ST = np.sqrt(2.*self.q_ST)/(np.sqrt(np.pi)+Gamma(0.5-self.p_ST)/2**self.p_ST)   # ~0.2161
            if name == 'Sheth, Mo & Tormen (2001)':
                self.a_SMT, self.b_SMT, self.c_SMT = 0.707, 0.5, 0.6
        elif name == 'Despali et al. (2016)':
            self.p_ST, self.q_ST = 0.2536, 0.7689
            self.A_ST = 0.3295*np.sqrt(2.*self.q_ST/np.pi)
        elif name == 'Tinker et al. (2010)':
            lnDv, lnDv_tab = np.log(self.Dv), np.log(_TINKER_DV)
            par = {key: _lin_interp_extrap(lnDv, lnDv_tab, tab) for key, tab in _TINKER_TABLE4.items()}
            if Tinker_z_dep:    # equations (9-12) of Tinker et al. (2010); only calibrated for M200
                par['beta'] *= (1.+z)**0.20
                par['gamma'] *= (1.+z)**-0.01
                par['phi'] *= (1.+z)**-0.08
                par['eta'] *= (1.+z)**0.27
            self.alpha_Tinker, self.beta_Tinker, self.gamma_Tinker = par['alpha'], par['beta'], par['gamma']
            self.phi_Tinker, self.eta_Tinker = par['phi'], par['eta']
            y = np.log10(self.Dv)       # calibrated bias, table 2
            e = np.exp(-(4./y)**4)
            self.A_Tinker, self.a_Tinker = 1.+0.24*y*e, 0.44*y-0.88
            self.B_Tinker, self.b_Tinker = 0.183, 1.5
            self.C_Tinker, self.c_Tinker = 0.019+0.107*y+0.19*e, 2.4
        if verbose:
            self._print_parameters()
            print()

    def _print_parameters(self):
        if self.name in ['Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)', 'Despali et al. (2016)']:
            print('p: %1.3f; q: %1.3f; A: %1.4f' % (self.p_ST, self.q_ST, self.A_ST))
            if self.name in ['Sheth, Mo & Tormen (2001)']:
                print('a: %1.3f; b: %1.3f; c: %1.3f' % (self.a_SMT, self.b_SMT, self.c_SMT))
        elif self.name in ['Tinker et al. (2010)']:
            print('alpha: %1.3f; beta: %1.3f; gamma: %1.3f; phi: %1.3f; eta: %1.3f'
                  % (self.alpha_Tinker, self.beta_Tinker, self.gamma_Tinker, self.phi_Tinker, self.eta_Tinker))
            print('A: %1.3f; a: %1.3f' % (self.A_Tinker, self.a_Tinker))
            print('B: %1.3f; b: %1.3f' % (self.B_Tinker, self.b_Tinker))
            print('C: %1.3f; c: %1.3f' % (self.C_Tinker, self.c_Tinker))

    def __str__(self):
        print('Halo model:', self.name)
        print('Scale factor: %1.3f' % (self.a))
        print('Redshift: %1.3f' % (self.z))
        print('Omega_m(z=0): %1.3f' % (self.Om_m))
        print('delta_c: %1.4f' % (self.dc))
        print('Delta_v: %4.1f' % (self.Dv))
        if self.name != 'Press & Schecter (1974)':
            print('Parameters:')
        self._print_parameters()
        return ''

    # -- C-ABI descriptor -----------------------------------------------------------------------------------
    def descriptor(self) -> ModelDesc:
        '''hm_model_desc of this model (include/halob200.h); Tinker_PBS is read now, as the reference does.'''
        d = ModelDesc()
        d.family, d.dc, d.rhom = _FAMILY[self.name], float(self.dc), float(self.rhom)
        if d.family in (_lib.ST99, _lib.SMT01, _lib.DESPALI16):
            d.par[0], d.par[1], d.par[2] = self.A_ST, self.q_ST, self.p_ST
            if d.family == _lib.SMT01:
                d.par[3], d.par[4], d.par[5] = self.a_SMT, self.b_SMT, self.c_SMT
        elif d.family == _lib.TINKER10:
            vals = (self.alpha_Tinker, self.beta_Tinker, self.gamma_Tinker, self.phi_Tinker, self.eta_Tinker,
                    self.A_Tinker, self.a_Tinker, self.B_Tinker, self.b_Tinker, self.C_Tinker, self.c_Tinker)
            for i, v in enumerate(vals):
                d.par[i] = float(v)
            if Tinker_PBS:
                d.family = _lib.TINKER10_PBS
        return d

    # -- functions of nu ------------------------------------------------------------------------------------
    def _mass_function_nu(self, nu):
        '''g(nu) with nu=delta_c/sigma(M); integrates to unity (pyhalomodel.py:217-240)'''
        g, _ = engine.massfn_nu(self.descriptor(), nu, want_b=False)
        return _give_back(g, _wants_device(nu))

    def _linear_bias_nu(self, nu):
        '''b(nu) (pyhalomodel.py:242-284)'''
        _, b = engine.massfn_nu(self.descriptor(), nu, want_g=False)
        return _give_back(b, _wants_device(nu))

    def _peak_height(self, M, sigmaM):
        '''nu = delta_c/sigma(M) (pyhalomodel.py:706-711)'''
        return self.dc/sigmaM

    def linear_bias(self, M, sigmaM):
        '''Linear halo bias as a function of halo mass (pyhalomodel.py:317-325)'''
        return self._linear_bias_nu(self._peak_height(M, sigmaM))

    def multiplicity_function(self, M, sigmaM):
        '''M^2 n(M)/rhobar (pyhalomodel.py:298-315): g(nu) dnu/dlnM with the reference's three-point Lagrange
        derivative of ln sigma against ln R (utility.py:15-35); one element-wise kernel (hm_multiplicity_function).'''
        return _give_back(engine.multiplicity_function(self.descriptor(), M, sigmaM), _wants_device(M, sigmaM))

    def mass_function(self, M, sigmaM):
        '''n(M) = F rhom/M^2 (pyhalomodel.py:286-296)'''
        return _give_back(engine.multiplicity_function(self.descriptor(), M, sigmaM, want_n=True), _wants_device(M, sigmaM))

    def Lagrangian_radius(self, M):
        '''Radius [Mpc/h] of a sphere containing mass M in a homogeneous universe (pyhalomodel.py:327-333)'''
        return cosmology.Lagrangian_radius(M, self.Om_m)

    def virial_radius(self, M):
        '''Halo virial radius from the halo mass and overdensity (pyhalomodel.py:713-719)'''
        R = self.Lagrangian_radius(M)
        return R/np.cbrt(self.Dv)

    def average(self, M, sigmaM, func):
        '''<f> = int f(M) n(M) dM = rhom trapz((f/M) g(nu), nu) (pyhalomodel.py:335-359)'''
        Mh = engine.to_host(M) if isinstance(M, torch.Tensor) else np.asarray(M)
        if not util.is_array_monotonic(Mh):
            raise ValueError('Halo mass array must be increasing monotonically')
        _check_halo_integration()
        out = engine.average(self.descriptor(), M, sigmaM, func)
        if np.ndim(func) > 1:       # func [nf, nM] broadcasts in the reference and gives one average per row
            return out if _wants_device(M, sigmaM, func) else engine.to_host(out)
        return out[0] if _wants_device(M, sigmaM, func) else float(out[0].item())

    def cross_halo_spectrum(self, k, Pk_lin, M, sigmaM, profiles: dict, M_halo: float, beta=None, simple_twohalo=False,
                            k_trunc=None, verbose=False) -> tuple:
        '''
        Power spectrum of each tracer crossed with haloes of one mass M_halo (pyhalomodel.py:573-677): dictionaries
        keyed by the profile names.  beta is beta_NL(k, M) at M_halo.  The mass integrals and the cubic
        interpolation in ln M are weighted row sums of the (k, M) grids on the device.
        '''
        return _cross_halo_spectrum(self, k, Pk_lin, M, sigmaM, profiles, M_halo, beta=beta,
                                    simple_twohalo=simple_twohalo, k_trunc=k_trunc, verbose=verbose)

    def _many_tracer_spectra(self, k, Pk_lin, M, sigmaM, tracers, beta, flags, on_dev):
        """More than 8 profiles.  The Gram path (FP64 DMMA, up to 64 single-grid Fourier profiles, even nM, 16-byte
        aligned grids, no beta) when its preconditions hold; otherwise -- profile(k, M, Uk, Wk) objects with separate
        grids, odd nM, beta_NL, more than 64 profiles -- the quadrature path on unions of two blocks of tracers, so that
        every pair the reference computes (pyhalomodel.py:433-507) is still produced.  Returns ([S, 3, nk], A)."""
        P, nM = len(tracers), np.shape(M)[0]
        gram_ok = (beta is None and P <= _lib.HM_GRAM_MAX_TRACERS and nM % 2 == 0
                   and all(t.mode in (_lib.GRID_UK, _lib.GRID_WK) and t.grid.data_ptr() % 16 == 0 for t in tracers))
        if gram_ok:
            out, lowmass = engine.GramPlan(k, Pk_lin, M, sigmaM, tracers, [self.descriptor()], **flags).run()
            if on_dev:
                return out[0], float(lowmass[0].item())
            packed = torch.cat([out.reshape(-1), lowmass]).cpu().numpy()
            return packed[:-1].reshape(P*(P+1)//2, 3, -1), float(packed[-1])
        blk = (6 if beta is not None else _lib.HM_MAX_TRACERS)//2
        blocks = [list(range(i, min(i+blk, P))) for i in range(0, P, blk)]
        pair_index = {}
        s = 0
        for u in range(P):
            for v in range(u, P):
                pair_index[(u, v)] = s
                s += 1
        res, A = None, 0.
        for bi, rows in enumerate(blocks):
            for cols in blocks[bi:]:
                if rows is cols and len(blocks) > 1:
                    continue            # the pairs inside a block come with its unions with the other blocks
                ids = rows if rows is cols else rows+cols
                plan = engine.PowerSpectrumPlan(k, Pk_lin, M, sigmaM, [tracers[i] for i in ids], [self.descriptor()],
                                                beta=beta, do_I11=do_I11, do_I12_I21=do_I12_I21, **flags)
                out, lowmass = plan.run()
                if res is None:
                    res = torch.empty((s, 3, plan.nk), dtype=torch.float64, device=out.device)
                    A = float(lowmass[0].item())
                t = 0
                for a_ in range(len(ids)):
                    for b_ in range(a_, len(ids)):
                        res[pair_index[(ids[a_], ids[b_])]] = out[0, t]
                        t += 1
        return (res if on_dev else res.cpu().numpy()), A

    # -- the hot path ---------------------------------------------------------------------------------------
    def power_spectrum(self, k, Pk_lin, M, sigmaM, profiles: dict, beta=None, simple_twohalo=False,
                       subtract_shotnoise=True, correct_discrete=True, k_trunc=None, verbose=False) -> tuple:
        '''
        Halo-model power spectra; returns the two-halo, one-halo and total dictionaries keyed 'u-v' for every
        ordered pair of profiles (pyhalomodel.py:361-513).  Arguments as in the reference, including `beta`
        (beta_NL(k, M1, M2), the non-linear halo bias of pyhalomodel.py:546-571; up to 6 profiles per call).
        '''
        if verbose:
            t_start = time()
        if simple_twohalo and (beta is not None):
            raise ValueError('A simple two-halo term is not compatible with non-linear halo bias')
        Mh = engine.to_host(M) if isinstance(M, torch.Tensor) else np.asarray(M, dtype=float)
        kh = engine.to_host(k) if isinstance(k, torch.Tensor) else np.asarray(k, dtype=float)
        if not util.is_array_monotonic(Mh):
            raise ValueError('Halo mass array must be increasing monotonically')
        if not isinstance(profiles, dict):
            raise TypeError('profiles must be a dictionary')
        for prof in profiles.values():
            if np.shape(kh) != np.shape(prof.k) or (kh != prof.k).all():
                raise ValueError('k arrays must all be identical to those in profiles')
            if np.shape(Mh) != np.shape(prof.M) or (Mh != prof.M).all():
                raise ValueError('Mass arrays must be identical to those in profiles')
        on_dev = _wants_device(k, Pk_lin, M, sigmaM, beta)
        names = list(profiles.keys())
        tracers = [prof._tracer() for prof in profiles.values()]
        flags = dict(simple_twohalo=simple_twohalo, subtract_shotnoise=subtract_shotnoise,
                     correct_discrete=correct_discrete, k_trunc=k_trunc)
        _check_halo_integration()
        if len(names) > _lib.HM_MAX_TRACERS:    # many tracers: Gram contraction on the FP64 tensor cores
            res, A = self._many_tracer_spectra(k, Pk_lin, M, sigmaM, tracers, beta, flags, on_dev)
        else:
            plan = engine.PowerSpectrumPlan(k, Pk_lin, M, sigmaM, tracers, [self.descriptor()], beta=beta,
                                            do_I11=do_I11, do_I12_I21=do_I12_I21, **flags)
            out, lowmass = plan.run()
            if on_dev:
                res, A = out[0], float(lowmass[0].item())
            else:   # one device->host copy for everything the caller sees
                packed = torch.cat([out.reshape(-1), lowmass]).cpu().numpy()
                res, A = packed[:-1].reshape(plan.S, 3, plan.nk), float(packed[-1])
        if verbose:
            sig = engine.to_host(engine.to_dev(sigmaM))
            print('Redshift:', self.z)
            print('Halo mass range [log10(Msun/h)]:', np.log10(Mh[0]), np.log10(Mh[-1]))
            print('Peak height range:', self.dc/sig[0], self.dc/sig[-1])
            print('Number of integration points:', len(Mh))
            print('Missing halo-bias-mass from the low-mass end of the two-halo integrand:', A)
        if A < 0.:
            warnings.warn('Warning: Mass function/bias correction is negative!', RuntimeWarning)
        if (not correct_discrete) and subtract_shotnoise and any(p.discrete_tracer for p in profiles.values()):
            warnings.warn('Warning: Subtracting shot noise while not treating discreteness properly is bad',
                          RuntimeWarning)
        Pk_2h_dict, Pk_1h_dict, Pk_hm_dict = {}, {}, {}
        s = 0
        for iu, name_u in enumerate(names):     # unique pairs first, in the kernel's (u<=v) order ...
            for iv in range(iu, len(names)):
                key = name_u+'-'+names[iv]
                Pk_2h_dict[key], Pk_1h_dict[key], Pk_hm_dict[key] = _split3(res[s], on_dev)
                s += 1
        ordered = ({}, {}, {})
        for name_u in names:                    # ... then every ordered pair in dict insertion order, with the
            for name_v in names:                # lower triangle aliasing the upper one (pyhalomodel.py:503-507)
                key, rkey = name_u+'-'+name_v, name_v+'-'+name_u
                for dst, src in zip(ordered, (Pk_2h_dict, Pk_1h_dict, Pk_hm_dict)):
                    dst[key] = src[key] if key in src else src[rkey]
        if verbose:
            print('Halomodel calculation time [s]:', time()-t_start, '\n')
        return ordered


MODEL_DESC_DTYPE = np.dtype([('family', '<i4'), ('reserved', '<i4'), ('dc', '<f8'), ('rhom', '<f8'), ('par', '<f8', (12,))])


def model_descriptors(z, Om_m, names, Dv, dc):
    """The hm_model_desc records of len(z) x len(names) models, sample-major, built with array arithmetic: the same
    expressions as model.__init__ / model.descriptor() element by element (pyhalomodel.py:73-184), so every record is
    bit-identical to model(z[i], Om_m[i], name=names[f], Dv=Dv[i], dc=dc[i]).descriptor().  For sweeps, where building
    the models one Python object at a time would cost more than the kernels.  Module switches (Tinker_z_dep,
    Tinker_PBS) are read now, as the reference does."""
    z, Om_m, Dv, dc = (np.atleast_1d(np.asarray(x, dtype=float)) for x in (z, Om_m, Dv, dc))
    G, F = len(z), len(names)
    out = np.zeros((G, F), dtype=MODEL_DESC_DTYPE)
    rhom = cosmology.comoving_matter_density(Om_m)
    for f, name in enumerate(names):
        if name not in _FAMILY:
            raise ValueError('Halo model not recognised')
        rec = out[:, f]
        rec['family'], rec['dc'], rec['rhom'] = _FAMILY[name], dc, rhom
        par = np.zeros((G, 12))
        if name in ('Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)'):
            from scipy.special import gamma as Gamma
            p_ST, q_ST = 0.3, 0.707
            par[:, 0], par[:, 1], par[:, 2] = np.sqrt(2.*q_ST)/(np.sqrt(np.pi)+Gamma(0.5-p_ST)/2**p_ST), q_ST, p_ST
            if name == 'Sheth, Mo & Tormen (2001)':
                par[:, 3], par[:, 4], par[:, 5] = 0.707, 0.5, 0.6
        elif name == 'Despali et al. (2016)':
            p_ST, q_ST = 0.2536, 0.7689
            par[:, 0], par[:, 1], par[:, 2] = 0.3295*np.sqrt(2.*q_ST/np.pi), q_ST, p_ST
        elif name == 'Tinker et al. (2010)':
            lnDv, lnDv_tab = np.log(Dv), np.log(_TINKER_DV)
            hi = np.clip(np.searchsorted(lnDv_tab, lnDv), 1, len(lnDv_tab)-1)
            lo = hi-1
            tp = {}
            for key, tab in _TINKER_TABLE4.items():
                ys = np.asarray(tab, dtype=float)
                tp[key] = ((lnDv-lnDv_tab[lo])/(lnDv_tab[hi]-lnDv_tab[lo]))*ys[hi] \
                    + ((lnDv_tab[hi]-lnDv)/(lnDv_tab[hi]-lnDv_tab[lo]))*ys[lo]
            if Tinker_z_dep:    # scalar pow per sample: numpy's array pow may differ from libm's in the last bit
                for key, ex in (('beta', 0.20), ('gamma', -0.01), ('phi', -0.08), ('eta', 0.27)):
                    tp[key] = tp[key]*np.array([(1.+zi)**ex for zi in z])
            y = np.log10(Dv)
            e = np.exp(-(4./y)**4)
            cols = (tp['alpha'], tp['beta'], tp['gamma'], tp['phi'], tp['eta'], 1.+0.24*y*e, 0.44*y-0.88,
                    np.full(G, 0.183), np.full(G, 1.5), 0.019+0.107*y+0.19*e, np.full(G, 2.4))
            for i, col in enumerate(cols):
                par[:, i] = col
            if Tinker_PBS:
                rec['family'] = _lib.TINKER10_PBS
        rec['par'] = par
    return out.reshape(G*F)


def _cubic_interp_weights(x, x0):
    """Weights L_m with interp1d(x, y, kind='cubic')(x0) = sum_m L_m y_m: scipy's cubic interp1d is the not-a-knot
    interpolating spline (make_interp_spline, k=3), which is linear in y, so one solve against the identity gives
    the weights for every row of a (k, M) grid at once (pyhalomodel.py:631-639)."""
    from scipy.interpolate import make_interp_spline
    return np.asarray(make_interp_spline(x, np.eye(len(x)), k=3)(x0), dtype=float)


def _sample_rule_weights(scheme, nr):
    """Weights per unit spacing of the sample-based rules of scipy.integrate the reference selects with
    `win_integration` (pyhalomodel.py:938-948): every one is linear in the samples."""
    if scheme == 'trapezoid':
        w = np.ones(nr)
        w[0] = w[-1] = 0.5
        return w
    if scheme in ('simps', 'simpson'):
        if nr % 2 == 0:
            raise ValueError('the Simpson rule here needs an odd number of samples')
        w = np.ones(nr)
        w[1:-1:2], w[2:-1:2] = 4., 2.
        return w/3.
    if scheme == 'romb':
        n = nr-1
        if n < 1 or n & (n-1):
            raise ValueError('Romberg integration needs 2^m+1 samples')
        from scipy.integrate import romb
        eye = np.eye(nr)
        return np.array([romb(eye[i], dx=1.) for i in range(nr)])
    raise ValueError('Halo window function integration method not recognised')


def _cross_halo_spectrum(self, k, Pk_lin, M, sigmaM, profiles, M_halo, beta=None, simple_twohalo=False,
                         k_trunc=None, verbose=False):
    if simple_twohalo and (beta is not None):
        raise ValueError('A simple two-halo term is not compatible with non-linear halo bias')
    Mh = engine.to_host(M) if isinstance(M, torch.Tensor) else np.asarray(M, dtype=float)
    kh = engine.to_host(k) if isinstance(k, torch.Tensor) else np.asarray(k, dtype=float)
    if not util.is_array_monotonic(Mh):
        raise ValueError('Halo mass array must be increasing monotonically')
    if not isinstance(profiles, dict):
        raise TypeError('profiles must be a dictionary')
    for prof in profiles.values():
        if np.shape(kh) != np.shape(prof.k) or (kh != prof.k).all():
            raise ValueError('k arrays must all be identical to those in profiles')
        if np.shape(Mh) != np.shape(prof.M) or (Mh != prof.M).all():
            raise ValueError('Mass arrays must be identical to those in profiles')
    on_dev = _wants_device(k, Pk_lin, M, sigmaM)
    L = _cubic_interp_weights(np.log(Mh), np.log(M_halo))              # cubic interpolation in ln M (:631-639)
    names = list(profiles.keys())
    out2, out1, outh = {}, {}, {}
    for i0 in range(0, len(names), _lib.HM_MAX_TRACERS):               # the kernel takes up to 8 tracers per launch
        chunk = names[i0:i0+_lib.HM_MAX_TRACERS]
        res, _ = engine.cross_halo_spectrum(self.descriptor(), k, Pk_lin, M, sigmaM, [profiles[n]._tracer() for n in chunk],
                                            L, beta=beta, simple_twohalo=simple_twohalo, k_trunc=k_trunc)
        res = res if on_dev else engine.to_host(res)
        for p, name in enumerate(chunk):
            out2[name], out1[name], outh[name] = _split3(res[p], on_dev)
    return out2, out1, outh


def _split3(block, on_dev):
    if on_dev:
        return block[0].clone(), block[1].clone(), block[2].clone()
    return block[0].copy(), block[1].copy(), block[2].copy()


class profile():
    '''
    Halo profile container with the data semantics of the reference (pyhalomodel.py:807-884).  The (k,M) grids
    live in HBM; `Uk` and `Wk` are materialised lazily (as numpy arrays for host-constructed profiles).
    '''

    def __init__(self, k, M, Uk, Wk, amplitude=None, normalisation=1., variance=None,
                 mass_tracer=False, discrete_tracer=False):
        self._init_common(k, M, amplitude, normalisation, variance, mass_tracer, discrete_tracer,
                          _wants_device(Uk, Wk))
        self._mode = _lib.GRID_SEPARATE
        self._grid = _own(Uk)
        self._wk = _own(Wk)

    def _init_common(self, k, M, amplitude, normalisation, variance, mass_tracer, discrete_tracer, on_dev):
        self.k = np.array(engine.to_host(k) if isinstance(k, torch.Tensor) else k, dtype=float)
        self.M = np.array(engine.to_host(M) if isinstance(M, torch.Tensor) else M, dtype=float)
        self._on_dev = on_dev
        self._amplitude = _own(amplitude) if amplitude is not None else None
        self.normalisation = normalisation
        self._variance = _own(variance) if variance is not None else None
        self.mass_tracer, self.discrete_tracer = mass_tracer, discrete_tracer
        self._wk = None

    @classmethod
    def Fourier(cls, k, M, Uk, amplitude=None, normalisation=1., variance=None, mass_tracer=False,
                discrete_tracer=False):
        '''
        Fourier-space profile (pyhalomodel.py:853-884).  If amplitude is None, Uk is taken to be dimensionful:
        amplitude := Uk[0,:], Wk := Uk, Uk := Uk/amplitude; otherwise Wk := (amplitude*Uk)/normalisation.
        Only ONE grid is kept in HBM; the quadrature kernel derives the other on the fly.
        '''
        if tuple(Uk.shape) != (k.shape[0], M.shape[0]):
            raise ValueError('k, M, array shapes do not match')
        self = cls.__new__(cls)
        grid = _own(Uk)
        if amplitude is None:
            self._init_common(k, M, grid[0, :], normalisation, variance, mass_tracer, discrete_tracer,
                              _wants_device(Uk))
            self._mode = _lib.GRID_WK
        else:
            self._init_common(k, M, amplitude, normalisation, variance, mass_tracer, discrete_tracer,
                              _wants_device(Uk, amplitude))
            self._mode = _lib.GRID_UK
        self._grid = grid
        return self

    @classmethod
    def configuration(cls, k, M, rv, c, differential_profile, amplitude=None, normalisation=1., variance=None,
                      mass_tracer=False, discrete_tracer=False):
        '''
        Configuration-space halo profile (pyhalomodel.py:886-925): differential_profile(r, M, rv, c) is the halo
        profile times 4 pi r^2; its Fourier transform W(k,M) = int_0^rv j0(kr) P(r) dr is evaluated on the device with
        the reference's sample-based rules (module switch `win_integration`: 'trapezoid', 'simps' or 'romb' on
        `nr_win_integration` points; the callable is sampled on the host).  If amplitude is None the profile is taken
        to be correctly normalised, otherwise it is renormalised to `amplitude`.
        '''
        Mh = np.asarray(engine.to_host(M) if isinstance(M, torch.Tensor) else M, dtype=float)
        rvh = np.asarray(engine.to_host(rv) if isinstance(rv, torch.Tensor) else rv, dtype=float)
        ch = np.asarray(engine.to_host(c) if isinstance(c, torch.Tensor) else c, dtype=float)
        nr = int(nr_win_integration)
        wq = _sample_rule_weights(win_integration, nr)
        samples = np.empty((len(Mh), nr))
        for iM, (_M, _rv, _c) in enumerate(zip(Mh, rvh, ch)):      # the callable is user Python code (:939-940)
            samples[iM] = differential_profile(np.linspace(0., _rv, nr), _M, _rv, _c)
        kk = np.concatenate([[0.], np.asarray(engine.to_host(k) if isinstance(k, torch.Tensor) else k, dtype=float)])
        W = engine.window_configuration(kk, rvh, samples, wq)       # row 0 is k = 0: the profile amplitudes (:909-911)
        amp0, W = W[0], W[1:]
        Uk = W/amp0                                                 # :917, :919
        amp = amp0 if amplitude is None else engine.to_dev(amplitude)
        prof = cls.Fourier(engine.to_dev(k), engine.to_dev(M), Uk, amplitude=amp, normalisation=normalisation,
                           variance=variance, mass_tracer=mass_tracer, discrete_tracer=discrete_tracer)
        prof._on_dev = _wants_device(k, M, rv, c)
        return prof

    # lazily materialised views with the reference's field names
    @property
    def Uk(self):
        if self._mode == _lib.GRID_WK:
            return _give_back(self._grid/self._amplitude, self._on_dev)
        return _give_back(self._grid, self._on_dev)

    @property
    def Wk(self):
        if self._mode == _lib.GRID_UK:
            return _give_back((self._amplitude*self._grid)/self.normalisation, self._on_dev)
        if self._mode == _lib.GRID_WK:
            return _give_back(self._grid, self._on_dev)
        return _give_back(self._wk, self._on_dev)

    @property
    def amplitude(self):
        return None if self._amplitude is None else _give_back(self._amplitude, self._on_dev)

    @property
    def variance(self):
        return None if self._variance is None else _give_back(self._variance, self._on_dev)

    def _tracer(self) -> engine.Tracer:
        norm = self.normalisation
        if isinstance(norm, torch.Tensor):
            norm = float(norm.item())
        return engine.Tracer(self._grid, self._amplitude, float(norm), variance=self._variance, wk=self._wk,
                             mode=self._mode, mass_tracer=self.mass_tracer, discrete_tracer=self.discrete_tracer)

    def __str__(self):
        log_limit = 1e5
        amp = engine.to_host(self._amplitude)
        print('Mass tracer:', self.mass_tracer)
        print('Discrete tracer:', self.discrete_tracer)
        if self.normalisation != 1.:
            print('Field normalisation:', self.normalisation)
        print('Number of wavenumber bins:', len(self.k))
        print('Number of mass bins:', len(self.M))
        print('Wavenumber range [log10(h/Mpc)]: %1.3f %1.3f' % (np.log10(self.k[0]), np.log10(self.k[-1])))
        print('Halo mass range [log10(Msun/h)]: %1.3f %1.3f' % (np.log10(self.M[0]), np.log10(self.M[-1])))
        print('The following are at the low and high halo mass ends')
        if amp[0] > log_limit:
            print('Profile amplitude mean [log10]:', np.log10(amp[0]), np.log10(amp[-1]))
        else:
            print('Profile amplitude mean:', amp[0], amp[-1])
        if self._variance is not None:
            var = engine.to_host(self._variance)
            if var[0] > log_limit:
                print('Profile amplitude variance [log10]:', np.log10(var[0]), np.log10(var[-1]))
            else:
                print('Profile amplitude variance:', var[0], var[-1])
        U, W = engine.to_host(engine.to_dev(self.Uk)), engine.to_host(engine.to_dev(self.Wk))
        print('Dimensionless profiles at low k (should be ~1):', U[0, 0], U[0, -1])
        print('Dimensionful profiles at low k (should be amplitudes):', W[0, 0], W[0, -1])
        return ''


### Halo profiles in Fourier space ###

def window_function(k, rv, *args, profile=None):
    '''
    Fourier transforms of halo profiles: 'delta', 'isothermal', 'NFW' (pyhalomodel.py:1006-1026).
    Scalar rv / c are accepted for 'isothermal' and 'NFW' (shape (nk, 1)), as np.outer does in the reference.
    '''
    on_dev = _wants_device(k, rv, *args)
    if profile == 'delta':
        Wk = torch.ones((len(k), len(rv)), dtype=torch.float64, device=engine.device())
    elif profile == 'isothermal':
        Wk = engine.window_isothermal(k, _vector(rv))
    elif profile == 'NFW':
        (c,) = args
        rvv = _vector(rv)
        cv = _vector(c)
        if cv.shape[0] != rvv.shape[0]:
            cv = cv.expand(rvv.shape[0]).contiguous() if cv.shape[0] == 1 else cv
            rvv = rvv.expand(cv.shape[0]).contiguous() if rvv.shape[0] == 1 else rvv
        Wk = engine.window_nfw(k, rvv, cv)
    else:
        raise ValueError('Halo profile not recognised')
    return _give_back(Wk, on_dev)


def _vector(x):
    t = engine.to_dev(x)
    return t.reshape(-1)


def matter_profile(k, M, rv, c, Om_m: float) -> profile:
    '''Pre-configured matter NFW profile (pyhalomodel.py:1065-1077); the grid never leaves the device.'''
    rhom = cosmology.comoving_matter_density(Om_m)
    Uk = engine.window_nfw(k, _vector(rv), _vector(c))
    prof = profile.Fourier(engine.to_dev(k), engine.to_dev(M), Uk, amplitude=engine.to_dev(M), normalisation=rhom,
                           mass_tracer=True)
    prof._on_dev = _wants_device(k, M, rv, c)
    return prof


### Halo concentration ###

_DUFFY = ((('M200', '200', '200b'), (10.14, -0.081, -1.01)),       # Duffy et al. (2008) table 1, full samples
          (('vir', 'virial', 'Mvir'), (7.85, -0.081, -0.71)),
          (('M200c', '200c'), (5.71, -0.084, -0.47)))


def concentration(M, z: float, method='Duffy et al. (2008)', halo_definition='M200'):
    '''Halo concentration c(M, z) (pyhalomodel.py:1084-1124)'''
    if method != 'Duffy et al. (2008)':
        raise ValueError('Halo concentration not recognised')
    for names, (A, B, C) in _DUFFY:
        if halo_definition in names:
            return A*(M/2e12)**B*(1.+z)**C
    raise ValueError('Halo definition not recognised')


### Halo-occupation distribution (HOD) ###

def _xp(M):
    return torch if isinstance(M, torch.Tensor) else np


def _step(x):
    '''np.heaviside(x, 1.)'''
    if isinstance(x, torch.Tensor):
        return (x >= 0.).to(x.dtype)
    return np.heaviside(x, 1.)


def _erf(x):
    if isinstance(x, torch.Tensor):
        return torch.erf(x)
    from scipy.special import erf
    return erf(x)


_HOD_KEYS = {'simple': ('Mmin', 'Msat', 'alpha'), 'Zehavi et al. (2004)': ('Mmin', 'M1', 'alpha'),
             'Zheng et al. (2005)': ('Mmin', 'sigma', 'M0', 'M1', 'alpha'),
             'Zhai et al. (2017)': ('Mmin', 'sigma', 'Msat', 'alpha', 'Mcut')}


def HOD_mean(M, method='Zheng et al. (2005)', **kwargs) -> tuple:
    '''Mean numbers of central and satellite galaxies (pyhalomodel.py:1131-1201); same methods and keywords (an unknown
    keyword is a TypeError, as it is for the reference's per-method functions).'''
    xp = _xp(M)
    if method not in _HOD_KEYS:
        raise ValueError('HOD method not recognised')
    for key in kwargs:
        if key not in _HOD_KEYS[method]:
            raise TypeError("HOD_mean(method=%r) got an unexpected keyword argument %r" % (method, key))
    if method == 'simple':
        Mmin, Msat, alpha = kwargs.get('Mmin', 1e12), kwargs.get('Msat', 1e13), kwargs.get('alpha', 1.)
        Nc = _step(M-Mmin)
        return Nc, Nc*(M/Msat)**alpha
    if method == 'Zehavi et al. (2004)':
        Mmin, M1, alpha = kwargs.get('Mmin', 1e12), kwargs.get('M1', 1e13), kwargs.get('alpha', 1.)
        return _step(M-Mmin), (M/M1)**alpha
    if method == 'Zheng et al. (2005)':
        Mmin, sigma, M0 = kwargs.get('Mmin', 1e12), kwargs.get('sigma', 0.15), kwargs.get('M0', 1e12)
        M1, alpha = kwargs.get('M1', 1e13), kwargs.get('alpha', 1.)
        Nc = _step(M-Mmin) if sigma == 0. else 0.5*(1.+_erf(xp.log10(M/Mmin)/sigma))
        return Nc, (_step(M-M0)*(M-M0)/M1)**alpha
    Mmin, sigma = kwargs.get('Mmin', 10**13.68), kwargs.get('sigma', 0.82)
    Msat, alpha, Mcut = kwargs.get('Msat', 10**14.87), kwargs.get('alpha', 0.41), kwargs.get('Mcut', 10**12.32)
    Nc = _step(M-Mmin) if sigma == 0. else 0.5*(1.+_erf(xp.log10(M/Mmin)/sigma))
    return Nc, ((M/Msat)**alpha)*xp.exp(-Mcut/M)


def HOD_variance(p, lam, central_condition=False) -> tuple:
    '''Bernoulli centrals, Poisson satellites (pyhalomodel.py:1204-1221)'''
    vcc = p*(1.-p)
    if central_condition:
        return vcc, p*lam*(1.+lam*(1.-p)), lam*(1.-p)
    return vcc, lam, 0.
