"""Engine-level behaviour on the GPU that the benchmark relies on: software-pipelined steps over two streams give
the same bits as plain calls, the per-stage entry points compose to the full call, and the `then` hook runs on
the step's stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipelined_plans_equal_plain_runs():
    from pyhalomodel_b200 import engine, workloads
    batch = workloads.build(5, 70, 192, ('m', 'g', 'p'), seed=21)
    plan = batch.plan()
    want, lowmass = plan.run()
    want, lowmass = want.clone(), lowmass.clone()
    pipe = engine.PipelinedPlans(batch.plan)
    kept = []
    for _ in range(5):          # odd count: both plans and both streams are used, the last step on stream 0
        out, low = pipe.submit(then=lambda o, a: kept.append((o.clone(), a.clone())))
    pipe.drain()
    torch.cuda.synchronize()
    assert len(kept) == 5
    for o, a in kept:
        assert torch.equal(o, want) and torch.equal(a, lowmass)
    assert torch.equal(out, want)
    # an explicit output buffer is honoured
    mine = torch.empty_like(want)
    pipe.submit(out=mine)
    pipe.drain()
    torch.cuda.synchronize()
    assert torch.equal(mine, want)


def test_stage_entry_points_compose():
    from pyhalomodel_b200 import workloads
    batch = workloads.build(3, 41, 128, ('m', 'g'), seed=22)
    plan = batch.plan()
    want = plan.run()[0].clone()
    plan.out.zero_()
    plan.weights()
    plan.quadrature(1)          # streaming kernel only: partial sums
    plan.quadrature(2)          # epilogue only
    torch.cuda.synchronize()
    assert torch.equal(plan.out, want)
    assert plan.stream_kernel_bytes() == 8*(3*2*41*128+3*5*128) and plan.algorithmic_bytes() > plan.stream_kernel_bytes()


def test_host_sweep_front_end_vs_public_api_and_oracle():
    """sweep.HostSweep -- host vectors in, host spectra out, chunks double-buffered over three streams -- against the
    reference-shaped API model by model (same kernels, 1e-12) and against the oracle (1e-10), with a ragged last
    chunk, two consecutive runs (slot reuse) and pinned as well as pageable inputs."""
    import pyhalomodel_b200.pyhalomodel as halo
    from conftest import assert_spectra_close
    from oracle import halo_oracle as orc
    from pyhalomodel_b200 import cosmology, sweep, synthetic as syn
    nk, nM, G = 40, 128, 5
    k, M = syn.k_grid(nk), syn.M_grid(nM)
    rng = np.random.default_rng(8)
    cos = [syn.draw_cosmology(rng) for _ in range(G)]
    hods = [syn.draw_hod(rng) for _ in range(G)]
    fams = syn.FAMILY_NAMES
    kg, dlnk = syn.sigma_k_grid()
    z, Om = np.array([c['z'] for c in cos]), np.array([c['Om_m'] for c in cos])
    Omz = syn.Om_m_of_z(z, Om)
    dc, Dv = cosmology.dc_NakamuraSuto(Omz), cosmology.Dv_BryanNorman(Omz)
    Pk, sig, rv, cc, N, V = (np.empty((G, n)) for n in (nk, nM, nM, nM, nM, nM))
    for i, cs in enumerate(cos):
        P0 = syn.LinearSpectrum(cs['Om_m'], cs['h'], cs['ns'], cs['sigma8'])
        Pk[i] = P0(k, cs['z'])
        sig[i] = syn.sigmaR_numpy(cosmology.Lagrangian_radius(M, cs['Om_m']), P0(kg, cs['z']), kg, dlnk)
        rv[i] = cosmology.Lagrangian_radius(M, cs['Om_m'])/np.cbrt(Dv[i])
        cc[i] = halo.concentration(M, cs['z'], halo_definition='Mvir')
        Nc, Ns = halo.HOD_mean(M, **hods[i])
        Vc, Vs, _ = halo.HOD_variance(Nc, Ns)
        N[i], V[i] = Nc+Ns, Vc+Vs
    hs = sweep.HostSweep(k, M, fams, chunk=2)
    out = hs.run(z, Om, Dv, dc, Pk, sig, rv, cc, N, V).copy()
    assert out.shape == (G*5, 3, 3, nk)
    pinned = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (Pk, sig, rv, cc, N, V)]
    again = hs.run(z, Om, Dv, dc, *pinned)
    assert np.array_equal(out, again)
    for i in range(G):
        for f, name in enumerate(fams):
            hmod = halo.model(z[i], Om[i], name=name, Dv=Dv[i], dc=dc[i])
            profs = {'m': halo.matter_profile(k, M, rv[i], cc[i], Om[i]),
                     'g': halo.profile.Fourier(k, M, halo.window_function(k, rv[i], profile='isothermal'), amplitude=N[i],
                                               normalisation=hmod.average(M, sig[i], N[i]), variance=V[i], discrete_tracer=True)}
            api = hmod.power_spectrum(k, Pk[i], M, sig[i], profs)
            om = orc.make_model(z[i], Om[i], name=name, Dv=Dv[i], dc=dc[i])
            oprofs = {'m': orc.matter_profile(k, M, rv[i], cc[i], Om[i]),
                      'g': orc.Profile.Fourier(k, M, orc.window_isothermal(k, rv[i]), amplitude=N[i],
                                               normalisation=orc.average(om, M, sig[i], N[i]), variance=V[i], discrete_tracer=True)}
            want = orc.power_spectrum(om, k, Pk[i], M, sig[i], oprofs)
            for s, key in enumerate(['m-m', 'm-g', 'g-g']):
                for t in range(3):
                    np.testing.assert_allclose(out[i*5+f, s, t], api[t][key], rtol=1e-12, err_msg=str((i, name, key, t)))
                    assert_spectra_close(out[i*5+f, s, t], want[t][key], rtol=1e-10, what='%d %s %s %d' % (i, name, key, t))


def test_stack_limit_round_trip():
    """engine.stack_limit reads and sets cudaLimitStackSize (the knob NCCL raises at communicator creation, which costs
    the HBM-bound kernels 2-3.5 %: tools/dev_stack_probe.py); a plan still runs, bit for bit, at either setting."""
    from pyhalomodel_b200 import engine, workloads
    old = engine.stack_limit()
    plan = workloads.build(2, 40, 128, ('m', 'g')).plan()
    ref = plan.run()[0].clone()
    try:
        assert engine.stack_limit(2048) == old
        assert engine.stack_limit() == 2048
        assert torch.equal(plan.run()[0], ref)
    finally:
        engine.stack_limit(old)
    assert engine.stack_limit() == old
