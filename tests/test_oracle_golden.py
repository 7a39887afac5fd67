"""CPU: pin oracle/halo_oracle.py against the golden vectors generated from the LIVE reference
(tests/golden/make_golden.py).  The oracle uses the same scipy/numpy calls in the same order, so the bar here
is bit-level (rtol 1e-14) -- much tighter than the 1e-10 the CUDA path is held to."""
import warnings

import numpy as np
import pytest

from conftest import assert_spectra_close, load_golden
from oracle import halo_oracle as orc
from pyhalomodel_b200 import synthetic as syn

SHORT = dict(zip(['PS74', 'ST99', 'SMT01', 'Tinker2010', 'Despali2016'], orc.FAMILIES))
TIGHT = 1e-14


def test_massfn_bias_and_low_mass_correction():
    g = load_golden('massfn.npz')
    nu, nu0 = g['nu'], g['nu0']
    for short, name in SHORT.items():
        for i, (Dv, dc) in enumerate(g['DvDc']):
            hm = orc.make_model(0.5, 0.3, name=name, Dv=Dv, dc=dc)
            np.testing.assert_allclose(orc.mass_function_nu(hm, nu), g['g_%s_%d' % (short, i)], rtol=TIGHT)
            np.testing.assert_allclose(orc.linear_bias_nu(hm, nu), g['b_%s_%d' % (short, i)], rtol=TIGHT)
            A = np.array([orc.low_mass_correction(hm, x) for x in nu0])
            np.testing.assert_allclose(A, g['A_%s_%d' % (short, i)], rtol=0, atol=1e-15)
    hm = orc.make_model(1.3, 0.3, name=orc.TINKER10, Dv=200., dc=1.686, tinker_z_dep=True)
    np.testing.assert_allclose(orc.mass_function_nu(hm, nu), g['g_Tinker2010_zdep'], rtol=TIGHT)
    np.testing.assert_allclose(orc.linear_bias_nu(hm, nu), g['b_Tinker2010_zdep'], rtol=TIGHT)
    hm = orc.make_model(0.5, 0.3, name=orc.TINKER10, Dv=330., dc=1.686, tinker_pbs=True)
    np.testing.assert_allclose(orc.linear_bias_nu(hm, nu), g['b_Tinker2010_pbs'], rtol=TIGHT)


def test_mass_function_golden():
    """multiplicity_function / mass_function (pyhalomodel.py:286-315 with utility.derivative_from_samples) for the
    five families at two redshifts: the oracle repeats the reference's scipy `lagrange` arithmetic bit for bit."""
    g = load_golden('mass_function.npz')
    M, Om_m = g['M'], float(g['Om_m'])
    for z in g['zs']:
        dc, Dv = g['dcDv_z%g' % z]
        sig = g['sigmaM_z%g' % z]
        for short, name in SHORT.items():
            hm = orc.make_model(z, Om_m, name=name, Dv=Dv, dc=dc)
            np.testing.assert_allclose(orc.multiplicity_function(hm, M, sig), g['F_%s_z%g' % (short, z)], rtol=TIGHT)
            np.testing.assert_allclose(orc.mass_function(hm, M, sig), g['n_%s_z%g' % (short, z)], rtol=TIGHT)


def test_windows_concentration_hod():
    g = load_golden('windows.npz')
    k, M, z, Om_m, Dv = g['k'], g['M'], float(g['z']), float(g['Om_m']), float(g['Dv'])
    np.testing.assert_allclose(orc.matter_density(Om_m), g['rhom'], rtol=0)
    np.testing.assert_allclose(orc.lagrangian_radius(M, Om_m), g['RL'], rtol=TIGHT)
    np.testing.assert_allclose(orc.virial_radius(M, Om_m, Dv), g['rv'], rtol=TIGHT)
    for hd in ['M200', 'Mvir', 'M200c']:
        np.testing.assert_allclose(orc.concentration(M, z, halo_definition=hd), g['c_'+hd], rtol=TIGHT)
    np.testing.assert_allclose(orc.window_nfw(k, g['rv'], g['c_Mvir']), g['U_nfw'], rtol=TIGHT)
    np.testing.assert_allclose(orc.window_isothermal(k, g['rv']), g['U_iso'], rtol=TIGHT)
    for meth, tag, kw in [('simple', 'simple', {}), ('Zehavi et al. (2004)', 'Zehavi', {}),
                          ('Zheng et al. (2005)', 'Zheng', {}), ('Zhai et al. (2017)', 'Zhai', {}),
                          ('Zheng et al. (2005)', 'Zheng_s0', dict(sigma=0., Mmin=3e12, M0=2e12, M1=4e13, alpha=0.9))]:
        Nc, Ns = orc.hod_mean(M, method=meth, **kw)
        np.testing.assert_allclose(Nc, g['Nc_'+tag], rtol=TIGHT)
        np.testing.assert_allclose(Ns, g['Ns_'+tag], rtol=TIGHT)
    for cc in (False, True):
        v = orc.hod_variance(g['Nc_Zheng'], g['Ns_Zheng'], central_condition=cc)
        for a, b in zip(v, g['var_%d' % cc]):
            np.testing.assert_allclose(np.broadcast_to(a, b.shape), b, rtol=TIGHT)
    np.testing.assert_allclose(orc.dc_nakamura_suto(g['dcDv_Omz']), g['dc_NS'], rtol=TIGHT)
    np.testing.assert_allclose(orc.Dv_bryan_norman(g['dcDv_Omz']), g['Dv_BN'], rtol=TIGHT)
    with pytest.raises(ValueError):
        orc.window_function(k, g['rv'], profile='bogus')
    with pytest.raises(ValueError):
        orc.concentration(M, z, halo_definition='bogus')
    with pytest.raises(ValueError):
        orc.hod_mean(M, method='bogus')


def test_sigmaR_brute():
    g = load_golden('sigma.npz')
    Om_m, h, ns, s8 = g['cosmo']
    Plin0 = syn.LinearSpectrum(Om_m, h, ns, s8)
    np.testing.assert_allclose(Plin0.amp, g['amp'], rtol=1e-13)
    kg, _ = orc.sigma_grid()
    np.testing.assert_allclose(Plin0(kg), g['Pk_grid'], rtol=1e-13)
    R = g['R'][:6]
    grid = g['Pk_grid']
    # sequential Python sum (cosmology.py:103) on the stored P(k) grid: bit-level
    sig = orc.sigmaR_brute(R, lambda k: grid)
    np.testing.assert_allclose(sig, g['sigma'][:6], rtol=TIGHT)
    # pairwise numpy sum: the re-association budget of SURVEY appendix B
    sig2 = orc.sigmaR_brute(g['R'], lambda k: grid, sequential=False)
    np.testing.assert_allclose(sig2, g['sigma'], rtol=1e-12)
    np.testing.assert_allclose(syn.sigmaR_numpy(g['R'], grid, *syn.sigma_k_grid()), g['sigma'], rtol=1e-12)
    assert abs(g['sigma'][-2]-s8) < 1e-12     # R = 8 Mpc/h normalisation


def _profiles(hm, g, hod=None, pressure=False, split=False):
    k, M, sig = g['k'], g['M'], g['sigmaM']
    rv = orc.virial_radius(M, hm.Om_m, hm.Dv)
    c = orc.concentration(M, hm.z, halo_definition='Mvir')
    profs = {'m': orc.matter_profile(k, M, rv, c, hm.Om_m)}
    if hod is not None:
        Nc, Ns = orc.hod_mean(M, **hod)
        Vc, Vs, _ = orc.hod_variance(Nc, Ns)
        Uiso = orc.window_isothermal(k, rv)
        if split:
            rho_g = orc.average(hm, M, sig, Nc+Ns)
            profs['c'] = orc.Profile.Fourier(k, M, orc.window_function(k, rv, profile='delta'), amplitude=Nc,
                                             normalisation=rho_g, variance=Vc, discrete_tracer=True)
            profs['s'] = orc.Profile.Fourier(k, M, Uiso, amplitude=Ns, normalisation=rho_g, variance=Vs,
                                             discrete_tracer=True)
        else:
            rho_g = orc.average(hm, M, sig, Nc+Ns)
            profs['g'] = orc.Profile.Fourier(k, M, Uiso, amplitude=Nc+Ns, normalisation=rho_g, variance=Vc+Vs,
                                             discrete_tracer=True)
        np.testing.assert_allclose(rho_g, g['rho_g'] if 'rho_g' in g else rho_g, rtol=TIGHT)
    if pressure:
        profs['p'] = orc.Profile.Fourier(k, M, syn.pressure_profile(k, M, rv))
    return profs


def _check(P2, P1, PH, g, prefix, names, rtol=TIGHT):
    for u in names:
        for v in names:
            want = g['%s_%s-%s' % (prefix, u, v)]
            for t, got in enumerate((P2, P1, PH)):
                assert_spectra_close(got[u+'-'+v], want[t], rtol=rtol, what='%s %s-%s term %d' % (prefix, u, v, t), zero_scale=0.)


def test_power_c1():
    g = load_golden('power_c1.npz')
    hm = orc.make_model(float(g['z']), float(g['Om_m']), name=orc.ST99, Dv=g['dcDv'][1], dc=g['dcDv'][0])
    profs = _profiles(hm, g)
    np.testing.assert_allclose(orc.virial_radius(g['M'], hm.Om_m, hm.Dv), g['rv'], rtol=TIGHT)
    out = orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs)
    _check(*out, g, 'P', ['m'])


def test_power_families():
    g = load_golden('power_families.npz')
    for z in g['zs']:
        dc, Dv = g['dcDv_z%g' % z]
        gg = dict(k=g['k'], M=g['M'], sigmaM=g['sigmaM_z%g' % z])
        for short, name in SHORT.items():
            hm = orc.make_model(z, float(g['Om_m']), name=name, Dv=Dv, dc=dc)
            gg['rho_g'] = g['rho_g_%s_z%g' % (short, z)]
            profs = _profiles(hm, gg, hod=syn.zheng_defaults())
            out = orc.power_spectrum(hm, g['k'], g['Pk_lin_z%g' % z], g['M'], gg['sigmaM'], profs)
            _check(*out, g, '%s_z%g' % (short, z), ['m', 'g'])


VARIANTS = {
    'default': {},
    'noshot': dict(subtract_shotnoise=False),
    'nocorr': dict(correct_discrete=False, subtract_shotnoise=False),
    'nocorr_sub': dict(correct_discrete=False, subtract_shotnoise=True),
    'ktrunc': dict(k_trunc=0.1),
    'simple': dict(simple_twohalo=True),
    'simple_ktrunc_noshot': dict(simple_twohalo=True, k_trunc=0.05, subtract_shotnoise=False),
}


def test_power_flags_and_aliasing():
    g = load_golden('power_flags.npz')
    dc, Dv = g['dcDv']
    hm = orc.make_model(float(g['z']), float(g['Om_m']), name=orc.TINKER10, Dv=Dv, dc=dc)
    hod = dict(zip(('Mmin', 'sigma', 'M0', 'M1', 'alpha'), g['hod']))
    profs = _profiles(hm, g, hod=hod, pressure=True, split=True)
    for tag, kw in VARIANTS.items():
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            P2, P1, PH = orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs, **kw)
        _check(P2, P1, PH, g, tag, ['m', 'c', 's', 'p'])
        assert PH['c-m'] is PH['m-c'] and P1['p-s'] is P1['s-p']
    # C16: central-central one-halo is identically zero under the default flags
    assert np.all(g['default_c-c'][1] == 0.)
    with pytest.warns(RuntimeWarning):
        orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs, correct_discrete=False)


def test_power_many_tracers():
    """Matter + 12 HOD samples (91 unique pairs): the oracle's pair loop at a tracer count that takes the Gram path on
    the GPU, pinned to the live reference."""
    g = load_golden('power_many.npz')
    dc, Dv = g['dcDv']
    k, M, sig = g['k'], g['M'], g['sigmaM']
    hm = orc.make_model(float(g['z']), float(g['Om_m']), name=orc.TINKER10, Dv=Dv, dc=dc)
    rv, c = orc.virial_radius(M, hm.Om_m, hm.Dv), orc.concentration(M, hm.z, halo_definition='Mvir')
    Unfw = orc.window_nfw(k, rv, c)
    profs = {'m': orc.matter_profile(k, M, rv, c, hm.Om_m)}
    for i, row in enumerate(g['hods']):
        hod = dict(zip(('Mmin', 'sigma', 'M0', 'M1', 'alpha'), row))
        Nc, Ns = orc.hod_mean(M, **hod)
        Vc, Vs, _ = orc.hod_variance(Nc, Ns)
        rho_g = orc.average(hm, M, sig, Nc+Ns)
        np.testing.assert_allclose(rho_g, g['rho_g'][i], rtol=TIGHT)
        profs['g%d' % i] = orc.Profile.Fourier(k, M, Unfw, amplitude=Nc+Ns, normalisation=rho_g, variance=Vc+Vs,
                                               discrete_tracer=True)
    P2, P1, PH = orc.power_spectrum(hm, k, g['Pk_lin'], M, sig, profs)
    names = list(profs.keys())
    assert len(names) == 13
    for iu, u in enumerate(names):
        for v in names[iu:]:
            key = u+'-'+v
            want = g['P_'+key]
            for got, ref_row in zip((P2[key], P1[key], PH[key]), want):
                np.testing.assert_allclose(got, ref_row, rtol=1e-13, atol=0., err_msg=key)
    assert PH['g3-m'] is PH['m-g3']


def test_power_c2_small_and_average():
    g = load_golden('power_c2_small.npz')
    dc, Dv = g['dcDv']
    hm = orc.make_model(float(g['z']), float(g['Om_m']), name=orc.TINKER10, Dv=Dv, dc=dc)
    profs = _profiles(hm, g, hod=syn.zheng_defaults(), pressure=True)
    out = orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs)
    _check(*out, g, 'P', ['m', 'g', 'p'])
    M, sig = g['M'], g['sigmaM']
    np.testing.assert_allclose(orc.average(hm, M, sig, M), g['avg_M'], rtol=TIGHT)
    np.testing.assert_allclose(orc.average(hm, M, sig, np.ones_like(M)), g['avg_one'], rtol=TIGHT)
    np.testing.assert_allclose(orc.linear_bias_nu(hm, orc.peak_height(hm, sig)), g['bias'], rtol=TIGHT)


def test_error_behaviour():
    g = load_golden('power_c1.npz')
    hm = orc.make_model(0., 0.3, name=orc.ST99, Dv=330.)
    profs = _profiles(hm, g)
    with pytest.raises(ValueError):
        orc.make_model(0., 0.3, name='nope')
    with pytest.raises(ValueError):
        orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'][::-1], g['sigmaM'], profs)
    with pytest.raises(TypeError):
        orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], list(profs.values()))
    with pytest.raises(ValueError):
        orc.power_spectrum(hm, g['k']*2., g['Pk_lin'], g['M'], g['sigmaM'], profs)
    with pytest.raises(ValueError):
        orc.Profile.Fourier(g['k'], g['M'], np.ones((3, 3)))


def test_power_beta_nl():
    """Non-linear halo bias (pyhalomodel.py:546-571) under the four states of do_I11 / do_I12_I21."""
    g = load_golden('power_beta.npz')
    dc, Dv = g['dcDv']
    hm = orc.make_model(float(g['z']), float(g['Om_m']), name=orc.TINKER10, Dv=Dv, dc=dc)
    profs = _profiles(hm, g, hod=syn.zheng_defaults(), pressure=True)
    try:
        for i11 in (True, False):
            for i12 in (True, False):
                orc.DO_I11, orc.DO_I12_I21 = i11, i12
                out = orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs, beta=g['beta'])
                _check(*out, g, 'I11_%d_I12_%d' % (i11, i12), ['m', 'g', 'p'], rtol=1e-13)
    finally:
        orc.DO_I11, orc.DO_I12_I21 = True, True
    with pytest.raises(ValueError):
        orc.power_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs, beta=g['beta'], simple_twohalo=True)


def _diff_profile_nfw(r, M, rv, c):
    rs = rv/c
    return M*(r/rs**2)/(1.+r/rs)**2/(np.log(1.+c)-c/(1.+c))


def _diff_profile_gas(r, M, rv, c):
    return 1e-3*M**0.8*4.*np.pi*r**2/(1.+(r/(0.3*rv))**2)**1.2


def test_configuration_profiles():
    """profile.configuration with the sample-based integrators (pyhalomodel.py:886-961)."""
    g = load_golden('configuration.npz')
    k, M, rv, c = g['k'], g['M'], g['rv'], g['c']
    try:
        for tag in ('trapezoid', 'simps', 'romb'):
            orc.WIN_INTEGRATION = tag
            pm = orc.configuration_profile(k, M, rv, c, _diff_profile_nfw, amplitude=M, normalisation=float(g['rhom']),
                                           mass_tracer=True)
            pg = orc.configuration_profile(k, M, rv, c, _diff_profile_gas)
            for name, pr in (('nfw', pm), ('gas', pg)):
                np.testing.assert_allclose(pr.Uk, g['%s_%s_Uk' % (tag, name)], rtol=1e-13)
                np.testing.assert_allclose(pr.Wk, g['%s_%s_Wk' % (tag, name)], rtol=1e-13)
                np.testing.assert_allclose(pr.amplitude, g['%s_%s_amp' % (tag, name)], rtol=1e-13)
    finally:
        orc.WIN_INTEGRATION = 'romb'


CROSS_VARIANTS = (('plain', {}), ('beta', 'beta'), ('ktrunc', dict(k_trunc=0.2)), ('simple', dict(simple_twohalo=True)))


def test_cross_halo_spectrum():
    """model.cross_halo_spectrum (pyhalomodel.py:573-704)."""
    g = load_golden('cross_halo.npz')
    dc, Dv = g['dcDv']
    hm = orc.make_model(float(g['z']), float(g['Om_m']), name=orc.TINKER10, Dv=Dv, dc=dc)
    profs = _profiles(hm, g, hod=syn.zheng_defaults(), pressure=True)
    for ih, Mh in enumerate(g['M_halo']):
        for tag, kw in CROSS_VARIANTS:
            kw = dict(beta=g['beta']) if kw == 'beta' else kw
            out = orc.cross_halo_spectrum(hm, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs, Mh, **kw)
            for name in ('m', 'g', 'p'):
                for t in range(3):
                    assert_spectra_close(out[t][name], g['h%d_%s_%s' % (ih, tag, name)][t], rtol=1e-12,
                                         what='cross %s %s %d' % (tag, name, t))
