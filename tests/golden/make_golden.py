#!/usr/bin/env python
"""
Generate the golden fixtures under tests/golden/ by running the LIVE reference (alexander-mead/pyhalomodel,
imported read-only from /root/reference) on synthetic inputs.  The reference cannot travel to the GPU box, so
inputs AND reference outputs are stored; tests replay them through oracle/halo_oracle.py (CPU) and through
the CUDA path (GPU).

Run from the repo root, in the authoring container only:
    python tests/golden/make_golden.py
Environment at generation time is recorded in each file ('meta').
"""
import os
import sys
import warnings

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import pyhalomodel.pyhalomodel as ref          # noqa: E402  (the live reference)
import pyhalomodel.cosmology as refcos         # noqa: E402
from pyhalomodel_b200 import synthetic as syn  # noqa: E402

META = np.array('reference pyhalomodel 1.0.2 @ /root/reference; numpy %s scipy %s' % (np.__version__, scipy.__version__))
SHORT = {'PS74': syn.FAMILY_NAMES[0], 'ST99': syn.FAMILY_NAMES[1], 'SMT01': syn.FAMILY_NAMES[2],
         'Tinker2010': syn.FAMILY_NAMES[3], 'Despali2016': syn.FAMILY_NAMES[4]}


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, meta=META, **arrays)
    print('wrote %-28s %8.1f KB' % (name, os.path.getsize(path)/1024.))


def dcDv(z, Om_m):
    Omz = syn.Om_m_of_z(z, Om_m)
    return refcos.dc_NakamuraSuto(Omz), refcos.Dv_BryanNorman(Omz)


def sigma_of_M(hmod, M, Plin0, z):
    R = hmod.Lagrangian_radius(M)
    return refcos.sigmaR(R, lambda k: Plin0(k, z)), R


# ------------------------------------------------------------------------------------------------
def golden_massfn():
    """g(nu), b(nu), A(nu0) for every family x (Dv, dc) and the two Tinker module switches."""
    nu = np.logspace(-2, 1.3, 96)
    nu0s = np.array([0.01, 0.03, 0.1, 0.2, 0.35, 0.6, 1., 2., 3.])
    DvDc = [(200., 1.686), (330., 1.686), (800., 1.6), (178., 1.675), (2500., 1.686), (4000., 1.7)]
    out = dict(nu=nu, nu0=nu0s, DvDc=np.array(DvDc))
    for short, name in SHORT.items():
        for i, (Dv, dc) in enumerate(DvDc):
            m = ref.model(0.5, 0.3, name=name, Dv=Dv, dc=dc)
            out['g_%s_%d' % (short, i)] = m._mass_function_nu(nu)
            out['b_%s_%d' % (short, i)] = m._linear_bias_nu(nu)
            A = []
            for x in nu0s:
                from scipy import integrate
                A.append(1.-integrate.quad(lambda t: m._mass_function_nu(t)*m._linear_bias_nu(t), x, np.inf)[0])
            out['A_%s_%d' % (short, i)] = np.array(A)
    # module switches (pyhalomodel.py:20-22)
    ref.Tinker_z_dep = True
    m = ref.model(1.3, 0.3, name=SHORT['Tinker2010'], Dv=200., dc=1.686)
    ref.Tinker_z_dep = False
    out['g_Tinker2010_zdep'] = m._mass_function_nu(nu)
    out['b_Tinker2010_zdep'] = m._linear_bias_nu(nu)
    ref.Tinker_PBS = True
    m = ref.model(0.5, 0.3, name=SHORT['Tinker2010'], Dv=330., dc=1.686)
    out['b_Tinker2010_pbs'] = m._linear_bias_nu(nu)
    from scipy import integrate
    out['A_Tinker2010_pbs'] = np.array([1.-integrate.quad(lambda t: m._mass_function_nu(t)*m._linear_bias_nu(t), x, np.inf)[0]
                                        for x in nu0s])
    ref.Tinker_PBS = False
    save('massfn.npz', **out)


def golden_mass_function():
    """multiplicity_function / mass_function (pyhalomodel.py:286-315, utility.py:15-35) for the five families at two
    redshifts on 128 masses, with the sigma(M) they were evaluated on."""
    nM = 128
    M = syn.M_grid(nM)
    Om_m = 0.3
    Plin0 = syn.LinearSpectrum(Om_m, 0.7, 0.96, 0.8)
    out = dict(M=M, Om_m=np.array(Om_m), zs=np.array([0., 1.]))
    for z in out['zs']:
        dc, Dv = dcDv(z, Om_m)
        out['dcDv_z%g' % z] = np.array([dc, Dv])
        sig = None
        for short, name in SHORT.items():
            hmod = ref.model(z, Om_m, name=name, Dv=Dv, dc=dc)
            if sig is None:
                sig, _ = sigma_of_M(hmod, M, Plin0, z)
                out['sigmaM_z%g' % z] = sig
            out['F_%s_z%g' % (short, z)] = hmod.multiplicity_function(M, sig)
            out['n_%s_z%g' % (short, z)] = hmod.mass_function(M, sig)
    save('mass_function.npz', **out)


def golden_windows():
    """U_NFW, U_iso, concentration, radii, HOD on small grids, plus a wide-argument sici stress grid."""
    k = syn.k_grid(48, 1e-3, 1e2)
    M = syn.M_grid(40)
    m = ref.model(0.7, 0.31, name=SHORT['Tinker2010'], Dv=330., dc=1.686)
    rv, RL = m.virial_radius(M), m.Lagrangian_radius(M)
    out = dict(k=k, M=M, z=np.array(0.7), Om_m=np.array(0.31), Dv=np.array(330.), rv=rv, RL=RL, rhom=np.array(m.rhom))
    for hd in ['M200', 'Mvir', 'M200c']:
        out['c_'+hd] = ref.concentration(M, 0.7, halo_definition=hd)
    c = out['c_Mvir']
    out['U_nfw'] = ref.window_function(k, rv, c, profile='NFW')
    out['U_iso'] = ref.window_function(k, rv, profile='isothermal')
    from scipy.special import sici
    x = np.concatenate([np.logspace(-12, 3.5, 400), np.linspace(0.5, 40., 400), [4., 8., 16., 1., 1e-300, 5e4, 1e6]])
    out['sici_x'] = x
    out['sici_si'], out['sici_ci'] = sici(x)
    for meth, kw in [('simple', {}), ('Zehavi et al. (2004)', {}), ('Zheng et al. (2005)', {}),
                     ('Zhai et al. (2017)', {}), ('Zheng et al. (2005)', dict(sigma=0., Mmin=3e12, M0=2e12, M1=4e13, alpha=0.9))]:
        Nc, Ns = ref.HOD_mean(M, method=meth, **kw)
        tag = meth.split()[0]+('_s0' if kw else '')
        out['Nc_'+tag], out['Ns_'+tag] = Nc, Ns
    Nc, Ns = out['Nc_Zheng'], out['Ns_Zheng']
    for cc in (False, True):
        v = ref.HOD_variance(Nc, Ns, central_condition=cc)
        out['var_%d' % cc] = np.array([np.broadcast_to(x, Nc.shape) for x in v])
    out['dcDv_Omz'] = np.array([0.2, 0.3, 0.5, 0.9, 1.0])
    out['dc_NS'] = refcos.dc_NakamuraSuto(out['dcDv_Omz'])
    out['Dv_BN'] = refcos.Dv_BryanNorman(out['dcDv_Omz'])
    save('windows.npz', **out)


def golden_sigma():
    """sigma(R) by the reference's brute log-k rectangle rule on 24 radii (cosmology.py:89-106)."""
    Plin0 = syn.LinearSpectrum(0.3, 0.7, 0.96, 0.8)
    R = np.concatenate([np.logspace(-2.2, 2.2, 22), [8., 1e-9]])
    kg, _ = syn.sigma_k_grid()
    sig = refcos.sigmaR(R, lambda k: Plin0(k))
    save('sigma.npz', R=R, sigma=sig, Pk_grid=Plin0(kg), amp=np.array(Plin0.amp),
         cosmo=np.array([0.3, 0.7, 0.96, 0.8]))


def build_profiles(hmod, k, M, sigmaM, z, hod=None, with_pressure=False, split_cs=False):
    """The recipe of tests/test_consistency.py:68-85 (+ a dimensionful 'pressure' profile with amplitude=None)."""
    rv = hmod.virial_radius(M)
    c = ref.concentration(M, z, halo_definition='Mvir')
    profs = {'m': ref.matter_profile(k, M, rv, c, hmod.Om_m)}
    extra = dict(rv=rv, c=c)
    if hod is not None:
        Nc, Ns = ref.HOD_mean(M, method='Zheng et al. (2005)', **hod)
        Vc, Vs, _ = ref.HOD_variance(Nc, Ns)
        Uiso = ref.window_function(k, rv, profile='isothermal')
        if split_cs:   # separate central / satellite tracers: exercises the exact-zero c-c one-halo term
            rho_g = hmod.average(M, sigmaM, Nc+Ns)
            Udel = ref.window_function(k, rv, profile='delta')
            profs['c'] = ref.profile.Fourier(k, M, Udel, amplitude=Nc, normalisation=rho_g, variance=Vc, discrete_tracer=True)
            profs['s'] = ref.profile.Fourier(k, M, Uiso, amplitude=Ns, normalisation=rho_g, variance=Vs, discrete_tracer=True)
            extra.update(Nc=Nc, Ns=Ns, Vc=Vc, Vs=Vs, rho_g=np.array(rho_g))
        else:
            Ng, Vg = Nc+Ns, Vc+Vs
            rho_g = hmod.average(M, sigmaM, Ng)
            profs['g'] = ref.profile.Fourier(k, M, Uiso, amplitude=Ng, normalisation=rho_g, variance=Vg, discrete_tracer=True)
            extra.update(Ng=Ng, Vg=Vg, rho_g=np.array(rho_g))
    if with_pressure:
        profs['p'] = ref.profile.Fourier(k, M, syn.pressure_profile(k, M, rv))
    return profs, extra


def flatten(prefix, P2, P1, PH, names):
    out = {}
    for u in names:
        for v in names:
            key = u+'-'+v
            out['%s_%s' % (prefix, key)] = np.array([P2[key], P1[key], PH[key]])
    return out


def golden_power_families():
    """m+g spectra (all pairs, three terms) for 5 families x z in {0, 1}; 64 k x 128 M."""
    nk, nM = 64, 128
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    Om_m = 0.3
    Plin0 = syn.LinearSpectrum(Om_m, 0.7, 0.96, 0.8)
    out = dict(k=k, M=M, Om_m=np.array(Om_m), zs=np.array([0., 1.]))
    for z in (0., 1.):
        dc, Dv = dcDv(z, Om_m)
        Pk = Plin0(k, z)
        out['Pk_lin_z%g' % z] = Pk
        out['dcDv_z%g' % z] = np.array([dc, Dv])
        for short, name in SHORT.items():
            hmod = ref.model(z, Om_m, name=name, Dv=Dv, dc=dc)
            if short == 'PS74':
                sig, _ = sigma_of_M(hmod, M, Plin0, z)
                out['sigmaM_z%g' % z] = sig
            sig = out['sigmaM_z%g' % z]
            profs, extra = build_profiles(hmod, k, M, sig, z, hod=syn.zheng_defaults())
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                P2, P1, PH = hmod.power_spectrum(k, Pk, M, sig, profs)
            out.update(flatten('%s_z%g' % (short, z), P2, P1, PH, ['m', 'g']))
            out['rho_g_%s_z%g' % (short, z)] = extra['rho_g']
    save('power_families.npz', **out)


def golden_power_c1():
    """BASELINE config 1: matter-only NFW, ST99, z=0, 128 k x 256 M (Dv=330, dc=1.686)."""
    nk, nM = 128, 256
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    Om_m, z = 0.3, 0.
    Plin0 = syn.LinearSpectrum(Om_m, 0.7, 0.96, 0.8)
    hmod = ref.model(z, Om_m, name=SHORT['ST99'], Dv=330., dc=1.686)
    sig, R = sigma_of_M(hmod, M, Plin0, z)
    Pk = Plin0(k, z)
    profs, extra = build_profiles(hmod, k, M, sig, z)
    P2, P1, PH = hmod.power_spectrum(k, Pk, M, sig, profs)
    save('power_c1.npz', k=k, M=M, Pk_lin=Pk, sigmaM=sig, z=np.array(z), Om_m=np.array(Om_m),
         dcDv=np.array([1.686, 330.]), rv=extra['rv'], c=extra['c'], **flatten('P', P2, P1, PH, ['m']))


def golden_power_flags():
    """Tinker10, 4 tracers (m, c, s, p) on 40 k x 96 M under every flag combination of power_spectrum
    (pyhalomodel.py:361-363): exercises N(N-1), the exactly-zero c-c term, shot noise, k_trunc, simple_twohalo
    and the amplitude=None profile."""
    nk, nM = 40, 96
    k, M = syn.k_grid(nk), syn.M_grid(nM, 1e10, 1e16)
    Om_m, z = 0.28, 0.4
    Plin0 = syn.LinearSpectrum(Om_m, 0.68, 0.97, 0.82)
    dc, Dv = dcDv(z, Om_m)
    hmod = ref.model(z, Om_m, name=SHORT['Tinker2010'], Dv=Dv, dc=dc)
    sig, R = sigma_of_M(hmod, M, Plin0, z)
    Pk = Plin0(k, z)
    hod = dict(Mmin=2e12, sigma=0.3, M0=1.5e12, M1=3e13, alpha=1.1)
    profs, extra = build_profiles(hmod, k, M, sig, z, hod=hod, with_pressure=True, split_cs=True)
    names = ['m', 'c', 's', 'p']
    out = dict(k=k, M=M, Pk_lin=Pk, sigmaM=sig, z=np.array(z), Om_m=np.array(Om_m), dcDv=np.array([dc, Dv]),
               hod=np.array([hod[x] for x in ('Mmin', 'sigma', 'M0', 'M1', 'alpha')]), rho_g=extra['rho_g'])
    variants = {
        'default': {},
        'noshot': dict(subtract_shotnoise=False),
        'nocorr': dict(correct_discrete=False, subtract_shotnoise=False),
        'nocorr_sub': dict(correct_discrete=False, subtract_shotnoise=True),
        'ktrunc': dict(k_trunc=0.1),
        'simple': dict(simple_twohalo=True),
        'simple_ktrunc_noshot': dict(simple_twohalo=True, k_trunc=0.05, subtract_shotnoise=False),
    }
    for tag, kw in variants.items():
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            P2, P1, PH = hmod.power_spectrum(k, Pk, M, sig, profs, **kw)
        out.update(flatten(tag, P2, P1, PH, names))
    # aliasing behaviour (pyhalomodel.py:503-507)
    assert PH['c-m'] is PH['m-c']
    save('power_flags.npz', **out)


def golden_power_c2_small():
    """BASELINE config 2 shape reduced to 96 k x 320 M: Tinker10, m + g + p, 6 unique pairs."""
    nk, nM = 96, 320
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    Om_m, z = 0.3, 0.
    Plin0 = syn.LinearSpectrum(Om_m, 0.7, 0.96, 0.8)
    dc, Dv = dcDv(z, Om_m)
    hmod = ref.model(z, Om_m, name=SHORT['Tinker2010'], Dv=Dv, dc=dc)
    sig, R = sigma_of_M(hmod, M, Plin0, z)
    Pk = Plin0(k, z)
    profs, extra = build_profiles(hmod, k, M, sig, z, hod=syn.zheng_defaults(), with_pressure=True)
    P2, P1, PH = hmod.power_spectrum(k, Pk, M, sig, profs)
    avg = {'avg_M': hmod.average(M, sig, M), 'avg_one': hmod.average(M, sig, np.ones_like(M)),
           'avg_N': hmod.average(M, sig, extra['Ng'])}
    save('power_c2_small.npz', k=k, M=M, Pk_lin=Pk, sigmaM=sig, z=np.array(z), Om_m=np.array(Om_m),
         dcDv=np.array([dc, Dv]), rho_g=extra['rho_g'], bias=hmod.linear_bias(M, sig),
         **{a: np.array(b) for a, b in avg.items()}, **flatten('P', P2, P1, PH, ['m', 'g', 'p']))


def synthetic_beta(k, M):
    """A smooth, deliberately asymmetric beta_NL(k, M1, M2) so that row/column mix-ups are visible."""
    lm = np.log10(M)
    x1, x2 = (lm[:, None]-13.)/2., (lm[None, :]-13.)/2.
    shape = 0.6*np.exp(-0.5*(x1-0.3)**2)*np.exp(-0.25*(x2+0.2)**2)+0.15*np.sin(1.3*x1)*np.cos(0.7*x2)-0.1*x1
    kk = k[:, None, None]
    return (kk/0.3)**0.8*np.exp(-(kk/1.5)**2)*shape[None, :, :]


def golden_power_beta():
    """Non-linear halo bias (beta_NL, pyhalomodel.py:546-571): Tinker10, m + g + p on 14 k x 56 M, with the two
    module switches do_I11 / do_I12_I21 in all four states."""
    nk, nM = 14, 56
    k, M = syn.k_grid(nk, 1e-2, 3.), syn.M_grid(nM, 1e10, 1e16)
    Om_m, z = 0.31, 0.25
    Plin0 = syn.LinearSpectrum(Om_m, 0.69, 0.965, 0.81)
    dc, Dv = dcDv(z, Om_m)
    hmod = ref.model(z, Om_m, name=SHORT['Tinker2010'], Dv=Dv, dc=dc)
    sig, R = sigma_of_M(hmod, M, Plin0, z)
    Pk = Plin0(k, z)
    profs, extra = build_profiles(hmod, k, M, sig, z, hod=syn.zheng_defaults(), with_pressure=True)
    beta = synthetic_beta(k, M)
    out = dict(k=k, M=M, Pk_lin=Pk, sigmaM=sig, z=np.array(z), Om_m=np.array(Om_m), dcDv=np.array([dc, Dv]),
               rho_g=extra['rho_g'], beta=beta)
    for i11 in (True, False):
        for i12 in (True, False):
            ref.do_I11, ref.do_I12_I21 = i11, i12
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                P2, P1, PH = hmod.power_spectrum(k, Pk, M, sig, profs, beta=beta)
            out.update(flatten('I11_%d_I12_%d' % (i11, i12), P2, P1, PH, ['m', 'g', 'p']))
    ref.do_I11, ref.do_I12_I21 = True, True
    save('power_beta.npz', **out)


def diff_profile_nfw(r, M, rv, c):
    """4 pi r^2 rho_NFW(r), truncated at rv, normalised to M (cf. the commented example at pyhalomodel.py:995-1000)."""
    rs = rv/c
    return M*(r/rs**2)/(1.+r/rs)**2/(np.log(1.+c)-c/(1.+c))


def diff_profile_gas(r, M, rv, c):
    """A dimensionful beta-model-like 'gas' profile times 4 pi r^2 (amplitude left free: amplitude=None)."""
    return 1e-3*M**0.8*4.*np.pi*r**2/(1.+(r/(0.3*rv))**2)**1.2


def golden_configuration():
    """profile.configuration (pyhalomodel.py:886-961) with the three sample-based win_integration schemes.  The
    installed scipy removed integrate.simps/quadrature/romberg, which the reference names at call time, so they are
    aliased here exactly as SURVEY.md section 8(c) describes."""
    from scipy import integrate
    integrate.simps = integrate.simpson
    integrate.quadrature = integrate.romberg = None
    k, M = syn.k_grid(20, 1e-2, 30.), syn.M_grid(12, 1e11, 1e15)
    hmod = ref.model(0.3, 0.3, name=SHORT['Tinker2010'], Dv=330., dc=1.686)
    rv, c = hmod.virial_radius(M), ref.concentration(M, 0.3, halo_definition='Mvir')
    out = dict(k=k, M=M, rv=rv, c=c, rhom=np.array(hmod.rhom))
    old = ref.win_integration
    try:
        for tag, scheme in (('trapezoid', integrate.trapezoid), ('simps', integrate.simps), ('romb', integrate.romb)):
            ref.win_integration = scheme
            pm = ref.profile.configuration(k, M, rv, c, diff_profile_nfw, amplitude=M, normalisation=hmod.rhom, mass_tracer=True)
            pg = ref.profile.configuration(k, M, rv, c, diff_profile_gas)
            for name, pr in (('nfw', pm), ('gas', pg)):
                out['%s_%s_Uk' % (tag, name)], out['%s_%s_Wk' % (tag, name)] = pr.Uk, pr.Wk
                out['%s_%s_amp' % (tag, name)] = pr.amplitude
    finally:
        ref.win_integration = old
    out['U_nfw_exact'] = ref.window_function(k, rv, c, profile='NFW')
    save('configuration.npz', **out)


def golden_cross_halo():
    """model.cross_halo_spectrum (pyhalomodel.py:573-704): tracer x halo-of-mass-M_h spectra for m + g + p, with and
    without beta_NL(k, M), k_trunc and the simple two-halo option."""
    nk, nM = 18, 64
    k, M = syn.k_grid(nk, 1e-2, 5.), syn.M_grid(nM, 1e10, 1e16)
    Om_m, z = 0.29, 0.35
    Plin0 = syn.LinearSpectrum(Om_m, 0.7, 0.96, 0.8)
    dc, Dv = dcDv(z, Om_m)
    hmod = ref.model(z, Om_m, name=SHORT['Tinker2010'], Dv=Dv, dc=dc)
    sig, R = sigma_of_M(hmod, M, Plin0, z)
    Pk = Plin0(k, z)
    profs, extra = build_profiles(hmod, k, M, sig, z, hod=syn.zheng_defaults(), with_pressure=True)
    beta = synthetic_beta(k, M)[:, :, nM//3]          # beta_NL(k, M) at the halo mass
    out = dict(k=k, M=M, Pk_lin=Pk, sigmaM=sig, z=np.array(z), Om_m=np.array(Om_m), dcDv=np.array([dc, Dv]),
               rho_g=extra['rho_g'], beta=beta, M_halo=np.array([3.3e12, 2.1e14]))
    for ih, Mh in enumerate(out['M_halo']):
        for tag, kw in (('plain', {}), ('beta', dict(beta=beta)), ('ktrunc', dict(k_trunc=0.2)), ('simple', dict(simple_twohalo=True))):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                P2, P1, PH = hmod.cross_halo_spectrum(k, Pk, M, sig, profs, Mh, **kw)
            for name in ('m', 'g', 'p'):
                out['h%d_%s_%s' % (ih, tag, name)] = np.array([P2[name], P1[name], PH[name]])
    save('cross_halo.npz', **out)


def golden_power_many():
    """Many tracers (BASELINE config 4 in small): matter + 12 HOD galaxy samples with NFW profiles, 91 unique
    pairs on 24 k x 96 M, Tinker10.  Pins the oracle's pair loop for a tracer count that takes the Gram path."""
    nk, nM, ngal = 24, 96, 12
    k, M = syn.k_grid(nk), syn.M_grid(nM, 1e10, 1e16)
    Om_m, z = 0.31, 0.25
    Plin0 = syn.LinearSpectrum(Om_m, 0.69, 0.965, 0.81)
    dc, Dv = dcDv(z, Om_m)
    hmod = ref.model(z, Om_m, name=SHORT['Tinker2010'], Dv=Dv, dc=dc)
    sig, R = sigma_of_M(hmod, M, Plin0, z)
    Pk = Plin0(k, z)
    rv = hmod.virial_radius(M)
    c = ref.concentration(M, z, halo_definition='Mvir')
    Unfw = ref.window_function(k, rv, c, profile='NFW')
    rng = np.random.default_rng(44)
    profs = {'m': ref.matter_profile(k, M, rv, c, Om_m)}
    hods, rho = [], []
    for i in range(ngal):
        hod = syn.draw_hod(rng)
        Nc, Ns = ref.HOD_mean(M, method='Zheng et al. (2005)', **hod)
        Vc, Vs, _ = ref.HOD_variance(Nc, Ns)
        rho_g = hmod.average(M, sig, Nc+Ns)
        profs['g%d' % i] = ref.profile.Fourier(k, M, Unfw, amplitude=Nc+Ns, normalisation=rho_g, variance=Vc+Vs,
                                               discrete_tracer=True)
        hods.append([hod[x] for x in ('Mmin', 'sigma', 'M0', 'M1', 'alpha')])
        rho.append(rho_g)
    names = list(profs.keys())
    out = dict(k=k, M=M, Pk_lin=Pk, sigmaM=sig, z=np.array(z), Om_m=np.array(Om_m), dcDv=np.array([dc, Dv]),
               hods=np.array(hods), rho_g=np.array(rho))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        P2, P1, PH = hmod.power_spectrum(k, Pk, M, sig, profs)
    for iu, u in enumerate(names):          # unique pairs only (the lower triangle aliases them)
        for v in names[iu:]:
            key = u+'-'+v
            out['P_'+key] = np.array([P2[key], P1[key], PH[key]])
    assert PH['g3-m'] is PH['m-g3']
    save('power_many.npz', **out)


if __name__ == '__main__':
    if len(sys.argv) > 1:                   # e.g. `make_golden.py power_many`: regenerate one fixture only
        for name in sys.argv[1:]:
            globals()['golden_'+name]()
        sys.exit(0)
    golden_massfn()
    golden_mass_function()
    golden_windows()
    golden_sigma()
    golden_power_c1()
    golden_power_families()
    golden_power_flags()
    golden_power_c2_small()
    golden_power_beta()
    golden_configuration()
    golden_cross_halo()
    golden_power_many()
