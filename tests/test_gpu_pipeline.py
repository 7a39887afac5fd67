"""GPU (-m gpu): the end-to-end device pipeline of BASELINE config 5 (P_lin -> sigma(M) -> mass function -> NFW /
isothermal profiles -> spectra) against the oracle's chain on the same samples."""
import numpy as np
import pytest
import torch

from conftest import assert_spectra_close
from oracle import halo_oracle as orc
from pyhalomodel_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _oracle_chain(s, k, M, family):
    P0 = syn.LinearSpectrum(s['Om_m'], s['h'], s['ns'], s['sigma8'])
    Omz = syn.Om_m_of_z(s['z'], s['Om_m'])
    dc, Dv = orc.dc_nakamura_suto(Omz), orc.Dv_bryan_norman(Omz)
    om = orc.make_model(s['z'], s['Om_m'], name=family, Dv=Dv, dc=dc)
    sig = orc.sigmaR_brute(orc.lagrangian_radius(M, s['Om_m']), lambda kk: P0(kk, s['z']), sequential=False)
    rv, c = orc.virial_radius(M, om.Om_m, om.Dv), orc.concentration(M, om.z, halo_definition='Mvir')
    Nc, Ns = orc.hod_mean(M, **s['hod'])
    Vc, Vs, _ = orc.hod_variance(Nc, Ns)
    profs = {'m': orc.matter_profile(k, M, rv, c, om.Om_m),
             'g': orc.Profile.Fourier(k, M, orc.window_isothermal(k, rv), amplitude=Nc+Ns,
                                      normalisation=orc.average(om, M, sig, Nc+Ns), variance=Vc+Vs, discrete_tracer=True)}
    return orc.power_spectrum(om, k, P0(k, s['z']), M, sig, profs), sig


def test_end_to_end_pipeline_vs_oracle():
    from pyhalomodel_b200 import pipeline
    nk, nM = 24, 64
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    samples = pipeline.draw_samples(5, seed=4321)
    fams = (orc.TINKER10, orc.ST99)
    e2e = pipeline.EndToEnd(k, M, families=fams, chunk=3)        # two chunks: 3 + 2 samples
    out, extras = e2e.run(samples, return_inputs=True)
    out = out.cpu().numpy()
    assert out.shape == (len(samples)*len(fams), 3, 3, nk)
    sig_dev = np.concatenate([e['sigma'].cpu().numpy() for e in extras])
    for i, s in enumerate(samples):
        for f, fam in enumerate(fams):
            want, sig = _oracle_chain(s, k, M, fam)
            np.testing.assert_allclose(sig_dev[i], sig, rtol=1e-12)
            for sp, key in enumerate(['m-m', 'm-g', 'g-g']):
                for t in range(3):
                    assert_spectra_close(out[i*len(fams)+f, sp, t], want[t][key], rtol=1e-10,
                                         what='sample %d %s %s term %d' % (i, fam, key, t))


def test_end_to_end_pipeline_full_shape_vs_oracle():
    """Config 5 at its full grid (256 k x 1024 M, Tinker10): two MCMC samples through the device chain against the
    oracle chain (sigma(R) on 1024 radii x 1e5 wavenumbers, NFW Si/Ci profiles, spectra)."""
    from pyhalomodel_b200 import pipeline
    nk, nM = 256, 1024
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    samples = pipeline.draw_samples(2, seed=4321)
    out, extras = pipeline.EndToEnd(k, M, chunk=2).run(samples, return_inputs=True)
    out = out.cpu().numpy()
    sig_dev = extras[0]['sigma'].cpu().numpy()
    for i, s in enumerate(samples):
        want, sig = _oracle_chain(s, k, M, orc.TINKER10)
        np.testing.assert_allclose(sig_dev[i], sig, rtol=1e-12)
        for sp, key in enumerate(['m-m', 'm-g', 'g-g']):
            for t in range(3):
                assert_spectra_close(out[i, sp, t], want[t][key], rtol=1e-10, what='C5 full sample %d %s term %d' % (i, key, t))
