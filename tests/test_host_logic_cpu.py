"""Host-side pieces of the reference-shaped API that need no GPU: the element-wise ingredient functions on numpy
inputs against the golden vectors from the live reference, and the quadrature / interpolation weight vectors
against scipy."""
import numpy as np
import pytest
from scipy import integrate, interpolate

from conftest import load_golden
import pyhalomodel_b200.pyhalomodel as halo
from pyhalomodel_b200 import cosmology, halomodel, utility

TIGHT = 1e-14


def test_ingredient_functions_on_numpy_match_the_reference():
    g = load_golden('windows.npz')
    M, z, Om_m, Dv = g['M'], float(g['z']), float(g['Om_m']), float(g['Dv'])
    hmod = halo.model(z, Om_m, name='Tinker et al. (2010)', Dv=Dv)
    assert hmod.rhom == float(g['rhom'])                                            # constants.py:4-13, bit for bit
    np.testing.assert_allclose(hmod.Lagrangian_radius(M), g['RL'], rtol=TIGHT)      # pyhalomodel.py:327
    np.testing.assert_allclose(hmod.virial_radius(M), g['rv'], rtol=TIGHT)          # :713
    for hd in ['M200', 'Mvir', 'M200c']:                                            # :1084-1124
        np.testing.assert_allclose(halo.concentration(M, z, halo_definition=hd), g['c_'+hd], rtol=TIGHT)
    for meth, tag, kw in [('simple', 'simple', {}), ('Zehavi et al. (2004)', 'Zehavi', {}),
                          ('Zheng et al. (2005)', 'Zheng', {}), ('Zhai et al. (2017)', 'Zhai', {}),
                          ('Zheng et al. (2005)', 'Zheng_s0', dict(sigma=0., Mmin=3e12, M0=2e12, M1=4e13, alpha=0.9))]:
        Nc, Ns = halo.HOD_mean(M, method=meth, **kw)                                # :1131-1196
        np.testing.assert_allclose(Nc, g['Nc_'+tag], rtol=TIGHT)
        np.testing.assert_allclose(Ns, g['Ns_'+tag], rtol=TIGHT)
    for cc in (False, True):                                                        # :1199-1221
        v = halo.HOD_variance(g['Nc_Zheng'], g['Ns_Zheng'], central_condition=cc)
        for a, b in zip(v, g['var_%d' % cc]):
            np.testing.assert_allclose(np.broadcast_to(a, b.shape), b, rtol=TIGHT)
    np.testing.assert_allclose([cosmology.dc_NakamuraSuto(x) for x in g['dcDv_Omz']], g['dc_NS'], rtol=TIGHT)
    np.testing.assert_allclose([cosmology.Dv_BryanNorman(x) for x in g['dcDv_Omz']], g['Dv_BN'], rtol=TIGHT)
    with pytest.raises(ValueError):
        halo.concentration(M, z, halo_definition='bogus')
    with pytest.raises(ValueError):
        halo.HOD_mean(M, method='bogus')


def test_cubic_interpolation_weights_are_scipys_interp1d():
    rng = np.random.default_rng(5)
    x = np.log(np.logspace(9., 16., 37))
    for x0 in (np.log(3e12), x[0], x[-1], 0.5*(x[10]+x[11])):
        L = halomodel._cubic_interp_weights(x, x0)
        for _ in range(3):
            y = rng.normal(size=x.size)
            want = float(interpolate.interp1d(x, y, kind='cubic')(x0))
            assert abs(L @ y-want) <= 1e-12*max(1., abs(want))
    assert abs(halomodel._cubic_interp_weights(x, x[7]).sum()-1.) < 1e-12          # reproduces constants


def test_sample_rule_weights_are_scipys_rules():
    rng = np.random.default_rng(6)
    nr, dx = 257, 0.0123
    y = rng.normal(size=nr)
    simpson = getattr(integrate, 'simpson', None) or integrate.simps
    for scheme, want in [('trapezoid', integrate.trapezoid(y, dx=dx)), ('simps', simpson(y, dx=dx)),
                         ('romb', integrate.romb(y, dx=dx))]:
        w = halomodel._sample_rule_weights(scheme, nr)
        assert abs(dx*(w @ y)-want) <= 1e-13*max(1., abs(want)), scheme
    with pytest.raises(ValueError):
        halomodel._sample_rule_weights('romb', 100)
    with pytest.raises(ValueError):
        halomodel._sample_rule_weights('simps', 100)
    with pytest.raises(ValueError):
        halomodel._sample_rule_weights('bogus', 257)


def test_utility_and_model_validation():
    assert utility.is_array_monotonic(np.array([1., 2., 3.])) and not utility.is_array_monotonic(np.array([1., 3., 2.]))
    np.testing.assert_allclose(utility.logspace(1e-3, 1e2, 6), np.logspace(-3., 2., 6), rtol=1e-15)
    with pytest.raises(ValueError):
        halo.model(0., 0.3, name='no such mass function')
    for name in ('Press & Schecter (1974)', 'Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)',
                 'Tinker et al. (2010)', 'Despali et al. (2016)'):
        d = halo.model(0.3, 0.3, name=name, Dv=330., dc=1.68).descriptor()
        assert d.dc == 1.68 and d.rhom > 0. and 0 <= d.family <= 5


def test_vectorised_model_descriptors_are_bit_identical_to_model_objects():
    """halomodel.model_descriptors (array arithmetic, for sweeps) against model(...).descriptor() per model, including
    the Tinker table extrapolation and both module switches (pyhalomodel.py:73-184, :20-22)."""
    from pyhalomodel_b200 import synthetic as syn
    rng = np.random.default_rng(0)
    G = 40
    z, Om = rng.uniform(0, 2, G), rng.uniform(.25, .35, G)
    Omz = syn.Om_m_of_z(z, Om)
    dc, Dv = cosmology.dc_NakamuraSuto(Omz), cosmology.Dv_BryanNorman(Omz)
    Dv[:5] = [150., 200., 3200., 5000., 800.]
    try:
        for zdep, pbs in [(False, False), (True, True)]:
            halo.Tinker_z_dep, halo.Tinker_PBS = zdep, pbs
            rec = halo.model_descriptors(z, Om, syn.FAMILY_NAMES, Dv, dc)
            assert rec.shape == (G*5,) and rec.dtype.itemsize == 120
            for i in range(G):
                for f, name in enumerate(syn.FAMILY_NAMES):
                    d = halo.model(z[i], Om[i], name=name, Dv=Dv[i], dc=dc[i]).descriptor()
                    assert bytes(d) == rec[i*5+f].tobytes(), (zdep, i, name)
    finally:
        halo.Tinker_z_dep, halo.Tinker_PBS = False, False
    with pytest.raises(ValueError):
        halo.model_descriptors(z, Om, ['nope'], Dv, dc)


def test_hod_keywords_and_halo_integration_switch():
    M = np.logspace(10, 15, 8)
    with pytest.raises(TypeError):
        halo.HOD_mean(M, Mmim=1e12)                       # the reference's per-method functions reject unknown keywords
    with pytest.raises(TypeError):
        halo.HOD_mean(M, method='simple', sigma=0.2)
    halo.halo_integration = 'simps'
    try:
        with pytest.raises(NotImplementedError):
            halomodel._check_halo_integration()
    finally:
        halo.halo_integration = 'trapezoid'
    halomodel._check_halo_integration()
