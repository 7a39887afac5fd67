"""GPU (-m gpu): the many-tracer path (more than 8 tracers -> Gram contraction on the FP64 tensor cores, BASELINE
config 4) against the oracle at sizes it finishes in seconds, and at config-4 size against the <=8-tracer
streaming path on tracer subsets plus size-independent properties."""
import warnings

import numpy as np
import pytest
import torch

from conftest import assert_spectra_close, load_golden
from oracle import halo_oracle as orc
from pyhalomodel_b200 import synthetic as syn
from test_gpu_parity import _synthetic_case

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope='module')
def halo():
    assert torch.cuda.is_available()
    import pyhalomodel_b200.pyhalomodel as h
    return h


def _tracers(api, model, g, hods, with_matter, with_pressure, nfw):
    """Build the same dictionary of profiles with the oracle API or the B200 API."""
    is_orc = api is orc
    k, M, sig = g['k'], g['M'], g['sigmaM']
    if is_orc:
        rv, c = orc.virial_radius(M, model.Om_m, model.Dv), orc.concentration(M, model.z, halo_definition='Mvir')
        Ug = orc.window_nfw(k, rv, c) if nfw else orc.window_isothermal(k, rv)
        Fourier, hod_mean, hod_var, avg = orc.Profile.Fourier, orc.hod_mean, orc.hod_variance, \
            (lambda N: orc.average(model, M, sig, N))
        matter = lambda: orc.matter_profile(k, M, rv, c, model.Om_m)
    else:
        rv, c = model.virial_radius(M), api.concentration(M, model.z, halo_definition='Mvir')
        Ug = api.window_function(k, rv, c, profile='NFW') if nfw else api.window_function(k, rv, profile='isothermal')
        Fourier, hod_mean, hod_var, avg = api.profile.Fourier, api.HOD_mean, api.HOD_variance, \
            (lambda N: model.average(M, sig, N))
        matter = lambda: api.matter_profile(k, M, rv, c, model.Om_m)
    profs = {}
    if with_matter:
        profs['m'] = matter()
    for i, hod in enumerate(hods):
        Nc, Ns = hod_mean(M, **hod)
        Vc, Vs, _ = hod_var(Nc, Ns)
        profs['g%d' % i] = Fourier(k, M, Ug, amplitude=Nc+Ns, normalisation=avg(Nc+Ns), variance=Vc+Vs,
                                   discrete_tracer=True)
    if with_pressure:
        profs['p'] = Fourier(k, M, syn.pressure_profile(k, M, rv))
    return profs


@pytest.mark.parametrize('ntr,nk,nM,name', [(9, 13, 130, orc.TINKER10), (12, 10, 96, orc.ST99), (17, 7, 64, orc.SMT01),
                                            (24, 5, 200, orc.PS74), (33, 4, 66, orc.DESPALI16)])
def test_gram_path_vs_oracle(halo, ntr, nk, nM, name):
    g = _synthetic_case(nk, nM, z=0.4)
    rng = np.random.default_rng(ntr)
    hods = [syn.draw_hod(rng) for _ in range(ntr-2)]
    om = orc.make_model(g['z'], g['Om_m'], name=name, Dv=g['Dv'], dc=g['dc'])
    hmod = halo.model(g['z'], g['Om_m'], name=name, Dv=g['Dv'], dc=g['dc'])
    op = _tracers(orc, om, g, hods, True, True, nfw=False)
    gp = _tracers(halo, hmod, g, hods, True, True, nfw=False)
    for kw in ({}, dict(subtract_shotnoise=False, k_trunc=0.3), dict(simple_twohalo=True, correct_discrete=False)):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            want = orc.power_spectrum(om, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], op, **kw)
            got = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], gp, **kw)
        assert list(got[0].keys()) == list(want[0].keys())
        for t in range(3):
            for key in want[t]:
                assert_spectra_close(got[t][key], want[t][key], rtol=RTOL, what='%s term %d %s' % (key, t, kw),
                                     zero_scale=np.max(np.abs(want[2][key])))
        assert got[1]['g0-m'] is got[1]['m-g0']


def test_many_tracers_without_gram_preconditions_vs_oracle(halo):
    """More than 8 profiles where the Gram path does not apply -- an odd number of masses, a profile(k, M, Uk, Wk)
    object with separate grids, beta_NL -- fall back to the quadrature path on unions of tracer blocks and still
    return every pair the reference computes (pyhalomodel.py:433-507)."""
    g = _synthetic_case(9, 65, z=0.3)
    rng = np.random.default_rng(5)
    hods = [syn.draw_hod(rng) for _ in range(8)]
    om = orc.make_model(g['z'], g['Om_m'], Dv=g['Dv'], dc=g['dc'])
    hmod = halo.model(g['z'], g['Om_m'], Dv=g['Dv'], dc=g['dc'])
    op = _tracers(orc, om, g, hods, True, True, nfw=False)          # 10 tracers, odd nM
    gp = _tracers(halo, hmod, g, hods, True, True, nfw=False)
    Uk = 1./(1.+np.outer(g['k'], np.linspace(0.1, 2., 65)))
    Wk = Uk*rng.uniform(0.5, 1.5, size=Uk.shape)
    amp = rng.uniform(0.5, 2., 65)
    op['x'] = orc.Profile(g['k'], g['M'], Uk, Wk, amplitude=amp, normalisation=2.2)
    gp['x'] = halo.profile(g['k'], g['M'], Uk, Wk, amplitude=amp, normalisation=2.2)
    beta = rng.normal(0., 0.3, size=(9, 65, 65))
    for kw in ({}, dict(beta=beta), dict(subtract_shotnoise=False, k_trunc=0.4)):
        want = orc.power_spectrum(om, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], op, **kw)
        got = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], gp, **kw)
        assert list(got[0].keys()) == list(want[0].keys()) and len(got[0]) == 121
        for t in range(3):
            for key in want[t]:
                assert_spectra_close(got[t][key], want[t][key], rtol=RTOL, what='blocked %s term %d' % (key, t),
                                     zero_scale=np.max(np.abs(want[2][key])))


def test_gram_64_tracers_vs_oracle_small_grid(halo):
    """All 2080 unique pairs of 64 HOD tracers (config 4's tracer count) on a grid the oracle finishes in seconds."""
    g = _synthetic_case(6, 128, z=0.2)
    rng = np.random.default_rng(64)
    hods = [syn.draw_hod(rng) for _ in range(64)]
    om = orc.make_model(g['z'], g['Om_m'], Dv=g['Dv'], dc=g['dc'])
    hmod = halo.model(g['z'], g['Om_m'], Dv=g['Dv'], dc=g['dc'])
    want = orc.power_spectrum(om, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], _tracers(orc, om, g, hods, False, False, True))
    got = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], _tracers(halo, hmod, g, hods, False, False, True))
    assert len(got[0]) == 64*64 and len({id(v) for v in got[2].values()}) == 2080
    for t in range(3):
        for key in want[t]:
            assert_spectra_close(got[t][key], want[t][key], rtol=RTOL, what='%s term %d' % (key, t))


def test_gram_replays_live_reference_golden_13_tracers(halo):
    """tests/golden/power_many.npz -- matter + 12 HOD samples, 91 unique pairs, written by the LIVE reference's pair
    loop (pyhalomodel.py:433-507) -- through the Gram path (13 tracers > 8)."""
    g = load_golden('power_many.npz')
    dc, Dv = g['dcDv']
    hmod = halo.model(float(g['z']), float(g['Om_m']), name=orc.TINKER10, Dv=Dv, dc=dc)
    hods = [dict(zip(('Mmin', 'sigma', 'M0', 'M1', 'alpha'), h)) for h in g['hods']]
    profs = _tracers(halo, hmod, g, hods, True, False, nfw=True)
    for i, rho in enumerate(g['rho_g']):
        np.testing.assert_allclose(profs['g%d' % i].normalisation, rho, rtol=1e-13)
    from pyhalomodel_b200 import _lib
    assert len(profs) == 13 > _lib.HM_MAX_TRACERS
    got = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs)
    names = list(profs.keys())
    assert list(got[0].keys()) == [u+'-'+v for u in names for v in names]
    for iu, u in enumerate(names):
        for v in names[iu:]:
            want = g['P_%s-%s' % (u, v)]
            for t in range(3):
                assert_spectra_close(got[t][u+'-'+v], want[t], rtol=RTOL, what='golden many %s-%s term %d' % (u, v, t),
                                     zero_scale=np.max(np.abs(want)))
            assert got[2][v+'-'+u] is got[2][u+'-'+v]


def test_gram_config4_size_vs_oracle_subsets_and_streaming_path(halo):
    """Config 4 at full size (64 HOD tracers, NFW, 256 k x 2048 M): every pair among 8-tracer subsets against the
    ORACLE on the same subset (rtol 1e-10) and against the <=8-tracer streaming-quadrature path (1e-12); the run is
    bitwise reproducible."""
    g = _synthetic_case(256, 2048, z=0.)
    rng = np.random.default_rng(4)
    hods = [syn.draw_hod(rng) for _ in range(64)]
    hmod = halo.model(g['z'], g['Om_m'], Dv=g['Dv'], dc=g['dc'])
    om = orc.make_model(g['z'], g['Om_m'], Dv=g['Dv'], dc=g['dc'])
    profs = _tracers(halo, hmod, g, hods, False, False, True)
    got = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs)
    again = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs)
    for key in ('g0-g0', 'g3-g60', 'g63-g63', 'g17-g18'):
        for t in range(3):
            assert np.array_equal(got[t][key], again[t][key])
    oprofs_all = _tracers(orc, om, g, hods, False, False, True)
    for subset in ([0, 9, 18, 27, 36, 45, 54, 63], [56, 57, 58, 59, 60, 61, 62, 63]):
        names = ['g%d' % i for i in subset]
        want = orc.power_spectrum(om, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], {n: oprofs_all[n] for n in names})
        for t in range(3):
            for u in names:
                for v in names:
                    assert_spectra_close(got[t][u+'-'+v], want[t][u+'-'+v], rtol=RTOL, what='C4 oracle %s-%s term %d' % (u, v, t))
    for subset in ([0, 1, 2, 3, 4, 5, 6, 7], [0, 9, 18, 27, 36, 45, 54, 63]):
        names = ['g%d' % i for i in subset]
        ref = hmod.power_spectrum(g['k'], g['Pk_lin'], g['M'], g['sigmaM'], {n: profs[n] for n in names})
        for t in range(3):
            for u in names:
                for v in names:
                    assert_spectra_close(got[t][u+'-'+v], ref[t][u+'-'+v], rtol=1e-12, what='%s-%s term %d' % (u, v, t))


def test_gram_batched_models_vs_single_calls(halo):
    """Several samples x mass functions in ONE Gram plan (the stream-K unit ranges of the Gram kernel then cross group AND
    model boundaries, and the epilogue adds partial slots of groups that belong to different models): every model of the
    batch against the single-model API call (which the tests above hold to the oracle), with ragged shapes."""
    from pyhalomodel_b200 import _lib, engine
    nk, nM, ntr = 23, 150, 11
    names = [orc.TINKER10, orc.ST99, orc.DESPALI16]
    cases = [_synthetic_case(nk, nM, z=z, Om_m=Om) for z, Om in [(0.1, 0.3), (0.9, 0.28)]]
    k, M = cases[0]['k'], cases[0]['M']
    rng = np.random.default_rng(42)
    hods = [syn.draw_hod(rng) for _ in range(ntr-1)]
    G, F = len(cases), len(names)
    singles, models = [], []
    grids = [[] for _ in range(ntr)]
    amps, vars_ = [[] for _ in range(ntr)], [[] for _ in range(ntr)]
    norms = [[] for _ in range(ntr)]
    for gi, cs in enumerate(cases):
        for name in names:
            hmod = halo.model(cs['z'], cs['Om_m'], name=name, Dv=cs['Dv'], dc=cs['dc'])
            profs = _tracers(halo, hmod, cs, hods, True, False, True)
            singles.append(hmod.power_spectrum(k, cs['Pk_lin'], M, cs['sigmaM'], profs))
            models.append(hmod.descriptor())
            for p, key in enumerate(profs):
                norms[p].append(profs[key].normalisation)
        for p, key in enumerate(profs):              # the grids, amplitudes and variances are per sample
            pr = profs[key]
            grids[p].append(engine.to_dev(pr.Uk))
            amps[p].append(engine.to_dev(pr.amplitude))
            vars_[p].append(engine.to_dev(pr.variance) if pr.variance is not None else None)
    keys = list(profs.keys())
    tr = []
    for p, key in enumerate(keys):
        var = torch.stack(vars_[p]) if vars_[p][0] is not None else None
        tr.append(engine.Tracer(torch.stack(grids[p]).contiguous(), torch.stack(amps[p]), np.array(norms[p], dtype=float),
                                variance=var, mode=_lib.GRID_UK, mass_tracer=profs[key].mass_tracer,
                                discrete_tracer=profs[key].discrete_tracer))
    plan = engine.GramPlan(k, np.stack([c['Pk_lin'] for c in cases]), M, np.stack([c['sigmaM'] for c in cases]), tr, models,
                           models_per_sample=F)
    out = plan.run()[0].cpu().numpy()
    s = 0
    for iu in range(ntr):
        for iv in range(iu, ntr):
            key = '%s-%s' % (keys[iu], keys[iv])
            for b, single in enumerate(singles):
                for t in range(3):
                    np.testing.assert_allclose(out[b, s, t], single[t][key], rtol=1e-12, atol=0., err_msg=str((b, key, t)))
            s += 1
