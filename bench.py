#!/usr/bin/env python
"""
bench.py -- halo-model spectra per second on B200 (BASELINE.json metric), contract of the build driver.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 2 -- Tinker et al. (2010), three tracers (matter NFW + Zheng05 HOD
galaxies + synthetic electron pressure), all 6 unique spectra, 512 k x 2048 M -- as a batch of B independent
models (distinct cosmology / redshift / HOD draws, distinct (k,M) grids) per GPU per step, so that the inputs of
one step (B x 25 MB) exceed the 126 MB L2.  One "step" = model.power_spectrum for every model of the batch:
the weights kernel + the streaming mass-quadrature kernel.  One spectrum = one unique tracer pair of one model
(its 2h, 1h and total vectors).  Weak scaling: every rank owns its own batch; results are all-gathered over NCCL.

value   : spectra/s with all inputs resident in HBM (CUDA events around exactly K steps, max over ranks)
e2e     : spectra/s through the public Python API with HOST (pinned numpy) inputs: profile.Fourier() uploads the
          grids, power_spectrum() computes, results come back as numpy -- H2D and D2H inside the timed region
roofline: the quadrature kernel alone, algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference : the oracle port of the reference's numpy/scipy algorithm on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NK, NM = 512, 2048
TRACERS = ('m', 'g', 'p')
METRIC, UNIT = 'halo_model_spectra_per_sec', 'spectra/s'
WORKLOAD = ('C2: Tinker et al. (2010), 3 tracers (matter NFW + Zheng05 HOD galaxies + pressure), 6 unique spectra, '
            '512 k x 2048 M; batch of %d independent models per GPU per step')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=64, help='models per GPU per step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-models', type=int, default=16, help='models in the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm (oracle/halo_oracle.py)
# ---------------------------------------------------------------------------------------------------------------
def _oracle_params(seed, n):
    """The (cosmology, HOD) draws of n config-2 models: the same recipe as workloads.build."""
    from pyhalomodel_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        cs = dict(Om_m=0.3, sigma8=0.8, ns=0.96, h=0.7, z=0.) if i == 0 else syn.draw_cosmology(rng)
        hod = syn.zheng_defaults() if i == 0 else syn.draw_hod(rng)
        out.append((cs, hod))
    return out


def _oracle_case(param):
    """Host-only construction of the inputs of one config-2 model (no GPU needed); sigma(M) takes ~15 s of numpy, so
    the reference arm builds its cases in parallel, outside the timed region."""
    from oracle import halo_oracle as orc
    from pyhalomodel_b200 import synthetic as syn
    cs, hod = param
    k, M = syn.k_grid(NK), syn.M_grid(NM)
    kg, dlnk = syn.sigma_k_grid()
    P0 = syn.LinearSpectrum(cs['Om_m'], cs['h'], cs['ns'], cs['sigma8'])
    Omz = syn.Om_m_of_z(cs['z'], cs['Om_m'])
    dc, Dv = orc.dc_nakamura_suto(Omz), orc.Dv_bryan_norman(Omz)
    sig = syn.sigmaR_numpy(orc.lagrangian_radius(M, cs['Om_m']), P0(kg, cs['z']), kg, dlnk)
    return dict(k=k, M=M, Pk=P0(k, cs['z']), sig=sig, z=cs['z'], Om_m=cs['Om_m'], dc=dc, Dv=Dv, hod=hod)


def _oracle_one(case):
    """One config-2 model through the oracle: profiles prebuilt (not timed), power_spectrum timed."""
    from oracle import halo_oracle as orc
    from pyhalomodel_b200 import synthetic as syn
    k, M, sig = case['k'], case['M'], case['sig']
    om = orc.make_model(case['z'], case['Om_m'], name=orc.TINKER10, Dv=case['Dv'], dc=case['dc'])
    rv, c = orc.virial_radius(M, om.Om_m, om.Dv), orc.concentration(M, om.z, halo_definition='Mvir')
    Nc, Ns = orc.hod_mean(M, **case['hod'])
    Vc, Vs, _ = orc.hod_variance(Nc, Ns)
    profs = {'m': orc.matter_profile(k, M, rv, c, om.Om_m),
             'g': orc.Profile.Fourier(k, M, orc.window_isothermal(k, rv), amplitude=Nc+Ns,
                                      normalisation=orc.average(om, M, sig, Nc+Ns), variance=Vc+Vs, discrete_tracer=True),
             'p': orc.Profile.Fourier(k, M, syn.pressure_profile(k, M, rv))}
    t0 = time.perf_counter()
    out = orc.power_spectrum(om, k, case['Pk'], M, sig, profs)
    dt = time.perf_counter()-t0
    return dt, float(out[2]['m-m'][0])


def cases_from_batch(batch, n):
    """The first n models of a device batch as host inputs for the oracle (same arrays the GPU path consumed)."""
    from pyhalomodel_b200 import engine
    k, M = engine.to_host(batch.k), engine.to_host(batch.M)
    out = []
    for g in range(n):
        meta = batch.meta[g]
        out.append(dict(k=k, M=M, Pk=engine.to_host(batch.Pk_lin[g]), sig=engine.to_host(batch.sigma[g]),
                        z=meta['cosmo']['z'], Om_m=meta['cosmo']['Om_m'], dc=meta['dc'], Dv=meta['Dv'], hod=meta['hod']))
    return out


def cpu_baseline_serial(cases):
    n_models = len(cases)
    t = sum(_oracle_one(cs)[0] for cs in cases)
    return dict(value=6*n_models/t, unit=UNIT, cores=1, kind='port',
                sample='%d config-2 models (6 spectra each) through oracle.power_spectrum on 1 core, %.1f s; '
                       'profiles prebuilt outside the timed region' % (n_models, t))


def run_reference(args, rank):
    """--impl reference: the oracle port on every host core (one process per model), same metric and config."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = cores
    ctx = mp.get_context('fork')
    times = []
    with ctx.Pool(cores) as pool:
        cases = pool.map(_oracle_case, _oracle_params(1234, per_step))     # inputs: built in parallel, not timed
        for step in range(args.warmup+args.steps):
            t0 = time.perf_counter()
            pool.map(_oracle_one, cases)
            if step >= args.warmup:
                times.append(time.perf_counter()-t0)
    total = sum(times)
    value = 6*per_step*len(times)/total
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3*total/len(times), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic', impl='reference',
                config=dict(workload=WORKLOAD % per_step, nk=NK, nM=NM, tracers=list(TRACERS), family='Tinker et al. (2010)'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind='port',
                                  sample='%d config-2 models per step, one per process, oracle.power_spectrum '
                                         '(numpy/scipy port of pyhalomodel.py:361-544)' % per_step),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu='+self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def summary(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        sm = sorted(float(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[1+1+i].lower().startswith('active') for r in rows)]
        return dict(sm_mhz=sm[len(sm)//2], sm_max_mhz=float(rows[0][1]), reasons=reasons, samples=len(rows))


def pinned_numpy(arr):
    import torch
    t = torch.empty(arr.shape, dtype=torch.float64).pin_memory()
    a = t.numpy()
    a[...] = arr
    return a, t


_STDOUT_FD = None


def emit(line):
    """Write the one JSON line to the process's original stdout."""
    data = (json.dumps(line)+'\n').encode()
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, data)


def main():
    args = parse()
    # stdout carries the JSON line and nothing else: libraries that print to file descriptor 1 (NCCL's version
    # banner, for one) are sent to stderr, and the line is written to the saved descriptor at the end
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # gloo carries the control plane (barriers, max over ranks of the timings); the NCCL communicator is created
        # lazily by the first CUDA collective, which is the all-gather that checks the fused gather AFTER all the
        # timed regions.  Measured on this stack (tools/dev_multi_probe.py): once an NCCL communicator exists in the
        # process, the streaming kernel runs 6 % slower (262 -> 277 us), whatever NCCL's transport settings.
        dist.init_process_group('cpu:gloo,cuda:nccl')
    import pyhalomodel_b200.pyhalomodel as halo
    from pyhalomodel_b200 import engine, workloads

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    batch = workloads.build(B, NK, NM, TRACERS, seed=1234+rank, fiducial_first=(rank == 0))
    plan = batch.plan()
    S = plan.S
    from pyhalomodel_b200 import sweep
    # N > 1: the one exchange of the path -- gathering every rank's spectra on rank 0 -- is fused into the epilogue
    # kernel: its output pointer is this rank's slot of a symmetric buffer that lives on rank 0, so the spectra are
    # stored over NVLink as they are produced (sweep.PeerGatherBuffer).  The NCCL all-gather is the baseline it is
    # checked against after the timed region.
    peer = sweep.PeerGatherBuffer(plan.out.shape, min(max(K, W), 32)) if world > 1 else None     # ring of result slots
    step_no = [0]

    # consecutive steps are software-pipelined over two CUDA streams (double-buffered plans): the weights and
    # epilogue kernels of one step run beside the persistent streaming kernel of the next (engine.PipelinedPlans)
    pipe = engine.PipelinedPlans(batch.plan)

    def step():
        if world > 1:
            pipe.submit(out=peer.slot(step_no[0]), peer=True)
            step_no[0] += 1
        else:
            pipe.submit()

    token = torch.zeros(1, dtype=torch.float64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(token)           # CPU tensor: gloo
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- HBM-resident throughput ------------------------------------------------------------------------------
    for _ in range(W):
        step()
    pipe.drain()
    barrier()
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    step_no[0] = 0
    for _ in range(K):
        step()
    pipe.drain()                 # the timed stream waits for both pipeline streams
    e1.record()                  # (the barrier below completes the gather: all peer stores have landed on rank 0)
    barrier()
    t1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world*B*S*K/(ms_total*1e-3)
    gathered = peer.local[:, (K-1) % peer.slots].clone() if (world > 1 and rank == 0) else None

    # ---- roofline of the dominant kernel (the streaming mass-quadrature kernel alone), CUDA events on the
    # launching stream; the weights and epilogue kernels are timed the same way for the step breakdown ----------
    reps = max(K, 20)

    def time_stage(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1)/reps

    quad_ms = time_stage(lambda: plan.quadrature(1))
    epi_ms = time_stage(lambda: plan.quadrature(2))
    weights_ms = time_stage(plan.weights)
    clocks = sampler.summary(t0, t1)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'MEASURED_PEAKS.json hbm_gbs (burst copy)'
    else:
        peak, peak_src = 6650., 'fallback (B200_PROFILING.md)'
    alg_bytes = plan.stream_kernel_bytes()
    achieved = alg_bytes/(quad_ms*1e-3)/1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'quad_traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get('batch') == B:
            traffic = tj.get('dram_bytes_per_launch')
    roofline = dict(bound='hbm', kernel='hm_quad_stream_kernel<3,3,2>', achieved=achieved, peak=peak, unit='GB/s',
                    frac=achieved/peak, traffic=traffic, algorithmic_bytes_per_launch=alg_bytes,
                    kernel_ms=quad_ms, epilogue_kernel_ms=epi_ms, weights_kernel_ms=weights_ms,
                    step_algorithmic_bytes=plan.algorithmic_bytes(),
                    step_achieved=plan.algorithmic_bytes()/(ms_total/K*1e-3)/1e9, peak_source=peak_src)

    # ---- end to end through the public API with host buffers --------------------------------------------------
    e2e = None
    n_e2e = min(B, 8)
    hosts, h2d, d2h = [], 0, 0
    for g in range(n_e2e):
        k, Pk, M, sig, grids = workloads.host_profiles(batch, g)
        pin = lambda a: pinned_numpy(a)[0]
        item = dict(k=pin(k), Pk=pin(Pk), M=pin(M), sig=pin(sig), model=batch.models[g], grids={})
        for name, d in grids.items():
            item['grids'][name] = {key: pin(v) for key, v in d.items()}
        item['norm_g'] = float(batch.tracers[1].normalisation[g].item())
        hosts.append(item)

    def e2e_one(it):
        gm, gg, gp = it['grids']['m'], it['grids']['g'], it['grids']['p']
        k, M = it['k'], it['M']
        profs = {'m': halo.profile.Fourier(k, M, gm['grid'], amplitude=gm['amplitude'], normalisation=it['model'].rhom,
                                           mass_tracer=True),
                 'g': halo.profile.Fourier(k, M, gg['grid'], amplitude=gg['amplitude'], normalisation=it['norm_g'],
                                           variance=gg['variance'], discrete_tracer=True),
                 'p': halo.profile.Fourier(k, M, gp['grid'])}
        return it['model'].power_spectrum(k, it['Pk'], M, it['sig'], profs)

    it = hosts[0]
    h2d_model = sum(v.nbytes for d in it['grids'].values() for v in d.values())+it['k'].nbytes+it['Pk'].nbytes \
        + it['M'].nbytes+it['sig'].nbytes
    d2h_model = (S*3*NK+1)*8
    for _ in range(2):
        for it in hosts:
            e2e_one(it)
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tt0 = time.perf_counter()
    ee0.record()
    e2e_steps = max(3, min(K, 10))
    for _ in range(e2e_steps):
        for it in hosts:
            res = e2e_one(it)
    ee1.record()
    barrier()
    e2e_wall = time.perf_counter()-tt0
    e2e_ms = max_over_ranks(max(ee0.elapsed_time(ee1), 1e3*e2e_wall))
    e2e = dict(value=world*n_e2e*S*e2e_steps/(e2e_ms*1e-3), unit=UNIT,
               h2d_bytes_per_step=h2d_model*n_e2e, d2h_bytes_per_step=d2h_model*n_e2e,
               models_per_step=n_e2e, steps=e2e_steps,
               path='profile.Fourier(host numpy) x3 + model.power_spectrum -> numpy, per model')

    # ---- the fused gather against the NCCL baseline (first NCCL collective of the process, after all timing) ---
    gather_ok = None
    if world > 1:
        mine = plan.run()[0].clone()
        torch.cuda.synchronize()
        ref_all = sweep.gather_spectra(mine, world*B)
        if rank == 0:
            got = gathered.reshape(ref_all.shape)
            gather_ok = bool(torch.equal(got, ref_all))
            if not gather_ok:
                bad = (got != ref_all)
                print('gather mismatch: %d of %d elements; per rank %s' % (
                    int(bad.sum()), bad.numel(), [int(b.sum()) for b in bad.reshape(world, -1)]), file=sys.stderr)

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cases = cases_from_batch(batch, min(args.cpu_models, B))
        cpu = cpu_baseline_serial(cases)
        # parity spot check of what was just timed: GPU batch sample 0 vs the oracle on the same inputs
        _, ref0 = _oracle_one(cases[0])
        out0 = plan.run()[0]
        torch.cuda.synchronize()
        cpu['parity_mm_k0'] = abs(float(out0[0, 0, 2, 0].item())/ref0-1.)
        if not cpu['parity_mm_k0'] < 1e-10:
            raise SystemExit('bench: GPU spectrum differs from the oracle on the timed inputs (rel %.3e)' % cpu['parity_mm_k0'])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total/K,
                    higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                    config=dict(workload=WORKLOAD % B, nk=NK, nM=NM, tracers=list(TRACERS), family='Tinker et al. (2010)',
                                models_per_gpu_per_step=B, pipelining='consecutive steps alternate between two CUDA streams (double-buffered workspaces)', l2='inputs larger than L2: %.0f MB of grids per step per GPU'
                                % (B*len(TRACERS)*NK*NM*8/1e6), collective=('gather to rank 0 fused into the epilogue kernel (peer stores over NVLink into a symmetric buffer); '
                                            'verified against NCCL all_gather: %s' % gather_ok) if world > 1 else 'none'),
                    clocks=clocks, e2e=e2e, gpu_launches=3*K, roofline=roofline, cpu_baseline=cpu)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
