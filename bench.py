#!/usr/bin/env python
"""
bench.py -- halo-model spectra per second on B200 (BASELINE.json metric), contract of the build driver.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (config.workload, the same in both arms): BASELINE config 3, the batched sweep the metric's
"1/2/4/8 B200" refers to -- (cosmology, z) samples x 5 mass functions, matter + HOD galaxies (3 unique spectra per
model), 256 k x 1024 M.  One spectrum = one unique tracer pair of one model (its 2h, 1h and total vectors).

b200 arm
  value    spectra/s with the (k, M) grids resident in HBM.  A step is --sub sub-batches of --samples samples per GPU
           (each sub-batch: weights kernel -> row-per-lane sweep kernel, which finishes the spectra itself; 2.1 GB of
           grids, larger than L2; two distinct sub-batches alternate), so that K = 20 steps last about a second.  Weak scaling: every
           rank owns its samples; the spectra are gathered on rank 0 by a copy-engine push of each sub-batch's block
           into a symmetric buffer behind its epilogue (no collective kernel; verified against NCCL all_gather).
           At N = 8 one sub-batch per GPU is exactly config 3's 4096 samples.
  roofline the sweep kernel hm_quad_rows_kernel<2,5> alone, algorithmic bytes / CUDA-event time, against
           MEASURED_PEAKS.json hbm_gbs; `step_frac` is the same for whole sub-batch steps.
  e2e      the same metric through the public batched API with HOST inputs (sweep.HostSweep): per sample the vectors
           a user of the reference prepares -- P_lin(k), sigma(M), r_v, c, <N>, var N -- are uploaded from pinned
           memory, the Fourier windows, normalisations and spectra are computed on the device and the spectra are
           downloaded, all inside the timed region.  Beside it (e2e.legs): the reference-shaped per-model API with
           host vectors, and with host (k, M) grids (25 MB per model over PCIe).
  configs  measured entries for BASELINE configs 1, 2, 4, 5 (and 3 strong-scaled) with kernel times, algorithmic
           bytes / flops and roofline fractions; the FP64 peaks (DFMA, DMMA) are measured in the same run.
  cpu_baseline  the reference arm's step on a bounded sample, on rank 0 at N = 1.

reference arm (--impl reference): the UNMODIFIED reference staged under baseline/_ref (tools/stage_reference.sh), or
the oracle port when it is absent, on every host core: per sample window_function + profile.Fourier +
model.power_spectrum for the five mass functions (the call sequence of tests/test_consistency.py:62-88).
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

NK, NM = 256, 1024
METRIC, UNIT = 'halo_model_spectra_per_sec', 'spectra/s'
FAMILIES = ('Press & Schecter (1974)', 'Sheth & Tormen (1999)', 'Sheth, Mo & Tormen (2001)', 'Tinker et al. (2010)',
            'Despali et al. (2016)')
CONFIG = dict(workload='C3: batched sweep of (cosmology, z) samples x 5 mass functions, matter NFW + Zheng05 HOD galaxies, '
                       '3 unique spectra per model, 256 k x 1024 M',
              nk=NK, nM=NM, tracers=['m', 'g'], families=list(FAMILIES), spectra_per_sample=15)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--samples', type=int, default=512, help='samples per GPU per sub-batch')
    ap.add_argument('--sub', type=int, default=96, help='sub-batches per GPU per step')
    ap.add_argument('--e2e-samples', type=int, default=2048, help='samples per GPU per end-to-end step')
    ap.add_argument('--c5-samples', type=int, default=2048, help='config-5 samples per GPU (16384 over 8 GPUs)')
    ap.add_argument('--cpu-samples', type=int, default=0, help='samples of the bounded CPU baseline (0: one per core, at most 32)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-configs', action='store_true', help='skip the configs block (C1, C2, C4, C5)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# Host inputs of one sample (shared by the reference arm, the CPU baseline and the end-to-end legs)
# ---------------------------------------------------------------------------------------------------------------
def _draws(seed, n, fiducial_first=True):
    from pyhalomodel_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        if i == 0 and fiducial_first:
            out.append((dict(Om_m=0.3, sigma8=0.8, ns=0.96, h=0.7, z=0.), syn.zheng_defaults()))
        else:
            out.append((syn.draw_cosmology(rng), syn.draw_hod(rng)))
    return out


def host_case(param, sigma=None):
    """The vectors a user prepares for one sample before calling the reference: k, M, P_lin(k), sigma(M), dc, Dv, r_v,
    c, HOD numbers and variances.  numpy only (sigma(M) takes ~4 s unless it is supplied): built outside every timed
    region."""
    from oracle import halo_oracle as orc
    from pyhalomodel_b200 import synthetic as syn
    cs, hod = param
    k, M = syn.k_grid(NK), syn.M_grid(NM)
    kg, dlnk = syn.sigma_k_grid()
    P0 = syn.LinearSpectrum(cs['Om_m'], cs['h'], cs['ns'], cs['sigma8'])
    Omz = syn.Om_m_of_z(cs['z'], cs['Om_m'])
    dc, Dv = orc.dc_nakamura_suto(Omz), orc.Dv_bryan_norman(Omz)
    if sigma is None:
        sigma = syn.sigmaR_numpy(orc.lagrangian_radius(M, cs['Om_m']), P0(kg, cs['z']), kg, dlnk)
    sig = sigma
    rv, c = orc.virial_radius(M, cs['Om_m'], Dv), orc.concentration(M, cs['z'], halo_definition='Mvir')
    Nc, Ns = orc.hod_mean(M, **hod)
    Vc, Vs, _ = orc.hod_variance(Nc, Ns)
    return dict(k=k, M=M, Pk=P0(k, cs['z']), sig=sig, z=cs['z'], Om_m=cs['Om_m'], dc=dc, Dv=Dv, rv=rv, c=c, N=Nc+Ns, V=Vc+Vs)


def _reference_module():
    """(module, kind): the unmodified reference from baseline/_ref when it is staged, else the oracle port."""
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if os.path.isdir(os.path.join(ref_dir, 'pyhalomodel')) and not os.environ.get('HALOB200_BENCH_PORT'):
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        import pyhalomodel.pyhalomodel as ref
        if os.path.dirname(os.path.abspath(ref.__file__)).startswith(ref_dir):
            return ref, 'reference'
    from oracle import halo_oracle as orc
    return orc, 'port'


def reference_sample(case):
    """One sample through the reference's own API: the windows, then per mass function the model, <N>, the two
    profiles and power_spectrum.  Returns (seconds, [F, 3 pairs, 3 terms, nk])."""
    ref, kind = _reference_module()
    k, M, sig, Pk = case['k'], case['M'], case['sig'], case['Pk']
    out = np.empty((len(FAMILIES), 3, 3, len(k)))
    t0 = time.perf_counter()
    if kind == 'reference':
        Um = ref.window_function(k, case['rv'], case['c'], profile='NFW')
        Ug = ref.window_function(k, case['rv'], profile='isothermal')
        for f, name in enumerate(FAMILIES):
            hmod = ref.model(case['z'], case['Om_m'], name=name, Dv=case['Dv'], dc=case['dc'])
            profs = {'m': ref.profile.Fourier(k, M, Um, amplitude=M, normalisation=hmod.rhom, mass_tracer=True),
                     'g': ref.profile.Fourier(k, M, Ug, amplitude=case['N'], normalisation=hmod.average(M, sig, case['N']),
                                              variance=case['V'], discrete_tracer=True)}
            res = hmod.power_spectrum(k, Pk, M, sig, profs)
            for s, key in enumerate(('m-m', 'm-g', 'g-g')):
                for t in range(3):
                    out[f, s, t] = res[t][key]
    else:
        Um, Ug = ref.window_nfw(k, case['rv'], case['c']), ref.window_isothermal(k, case['rv'])
        for f, name in enumerate(FAMILIES):
            om = ref.make_model(case['z'], case['Om_m'], name=name, Dv=case['Dv'], dc=case['dc'])
            profs = {'m': ref.Profile.Fourier(k, M, Um, amplitude=M, normalisation=om.rhom, mass_tracer=True),
                     'g': ref.Profile.Fourier(k, M, Ug, amplitude=case['N'], normalisation=ref.average(om, M, sig, case['N']),
                                              variance=case['V'], discrete_tracer=True)}
            res = ref.power_spectrum(om, k, Pk, M, sig, profs)
            for s, key in enumerate(('m-m', 'm-g', 'g-g')):
                for t in range(3):
                    out[f, s, t] = res[t][key]
    return time.perf_counter()-t0, out


def _pool(cores):
    import multiprocessing as mp
    return mp.get_context('fork').Pool(cores)


def run_reference(args, rank):
    """--impl reference: every host core evaluates one sample per step (bounded sample of the workload)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cores
    _, kind = _reference_module()
    times = []
    with _pool(cores) as pool:
        cases = pool.map(host_case, _draws(1234, per_step))              # inputs: built in parallel, not timed
        for step in range(args.warmup+args.steps):
            t0 = time.perf_counter()
            pool.map(reference_sample, cases)
            if step >= args.warmup:
                times.append(time.perf_counter()-t0)
    total = sum(times)
    value = CONFIG['spectra_per_sample']*per_step*len(times)/total
    sample = ('%d samples per step, one per process on %d cores: window_function x2 + per mass function model, average, '
              'profile.Fourier x2, power_spectrum (%s)' % (per_step, cores, 'unmodified reference from baseline/_ref'
                                                           if kind == 'reference' else 'oracle port of pyhalomodel.py:361-544'))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3*total/len(times), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic', impl='reference', config=CONFIG,
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


def cpu_baseline(cases, n_cores):
    """The reference arm's step on a bounded sample, timed on this box's host cores; returns (dict, outputs)."""
    _, kind = _reference_module()
    t0 = time.perf_counter()
    if n_cores > 1:
        with _pool(n_cores) as pool:
            res = pool.map(reference_sample, cases)
    else:
        res = [reference_sample(c) for c in cases]
    wall = time.perf_counter()-t0
    value = CONFIG['spectra_per_sample']*len(cases)/wall
    one_core = CONFIG['spectra_per_sample']*len(cases)/sum(r[0] for r in res)
    return dict(value=value, unit=UNIT, cores=n_cores, kind=kind, one_core_value=one_core,
                sample='%d samples (%d spectra) on %d cores in %.1f s: window_function x2 + per mass function model, average, '
                       'profile.Fourier x2, power_spectrum' % (len(cases), CONFIG['spectra_per_sample']*len(cases), n_cores, wall)), \
        [r[1] for r in res]


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw'

    def __init__(self, index, period_ms=50):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu='+self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', str(period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def summary(self, t0, t1):
        """Clocks and throttle reasons of the samples taken INSIDE [t0, t1] (the timed region); none there is said so."""
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'], samples_in_window=0)
        self.proc.terminate()
        allrows = [r for _, r in self.rows if len(r) >= 6]
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=float(allrows[0][1]) if allrows else None, reasons=['no samples in the timed region'],
                        samples_in_window=0, samples_total=len(allrows))
        sm = sorted(float(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2+i].lower().startswith('active') for r in rows)]
        power = [float(r[6]) for r in rows if len(r) > 6 and r[6].replace('.', '', 1).isdigit()]
        return dict(sm_mhz=sm[len(sm)//2], sm_min_mhz=sm[0], sm_max_mhz=float(rows[0][1]), reasons=reasons,
                    samples_in_window=len(rows), samples_total=len(allrows), power_w_max=max(power) if power else None)


_STDOUT_FD = None


def emit(line):
    """Write the one JSON line to the process's original stdout."""
    data = (json.dumps(line)+'\n').encode()
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, data)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node (PCIe copies and pinned staging stay local)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu+63)//64)
        cpus = [64*w+b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------------------
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
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # gloo carries the control plane (barriers, max over ranks of the timings); the NCCL communicator is created
        # lazily by the first CUDA collective, which is the all-gather that checks the gather AFTER the timed regions
        # (an NCCL communicator in the process costs the streaming kernels a few per cent).
        dist.init_process_group('cpu:gloo,cuda:nccl')
    import pyhalomodel_b200.pyhalomodel as halo
    from pyhalomodel_b200 import engine, pipeline, sweep, synthetic as syn, workloads

    G, SUB, K, W = args.samples, args.sub, args.steps, max(args.warmup, 3)
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

    def event_time(fn, reps):
        """Mean milliseconds of fn() over `reps` back-to-back calls, CUDA events on the launching stream."""
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

    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'MEASURED_PEAKS.json hbm_gbs (burst copy)'
    else:
        hbm_peak, peak_src = 6650., 'fallback (B200_PROFILING.md)'

    # ---- the sweep: two distinct sub-batches per rank, double-buffered plans on two streams ------------------------
    t_build = time.perf_counter()
    batches = [workloads.build(G, NK, NM, ('m', 'g'), FAMILIES, seed=1234+17*rank+b, fiducial_first=(rank == 0 and b == 0))
               for b in range(2)]
    plans = iter([batches[0].plan(), batches[1].plan()])
    pipe = engine.PipelinedPlans(lambda: next(plans))
    plan0, plan1 = pipe.plans
    S, B = plan0.S, plan0.B
    log('rank %d: built 2 x %d samples in %.1f s' % (rank, G, time.perf_counter()-t_build))
    # N > 1: the one exchange of the path -- gathering every rank's spectra on rank 0.  A copy engine pushes each
    # sub-batch's block into rank 0's symmetric buffer behind its epilogue, beside the next sub-batch's kernels.
    peer = sweep.PeerGatherBuffer(plan0.out.shape, 2) if world > 1 else None
    sub_no = [0]

    def sub_batch():
        if world > 1:
            slot = peer.slot(sub_no[0])
            pipe.submit(then=lambda out, lowmass: slot.copy_(out, non_blocking=True))
        else:
            pipe.submit()
        sub_no[0] += 1

    def step():
        for _ in range(SUB):
            sub_batch()

    for _ in range(W):
        step()
    pipe.drain()
    barrier()
    sampler = ClockSampler(local_rank)
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    sub_no[0] = 0
    for _ in range(K):
        step()
    pipe.drain()                 # the timed stream waits for both pipeline streams
    e1.record()                  # (the barrier below completes the gather: all pushes have landed on rank 0)
    barrier()
    t1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.summary(t0, t1)
    value = world*B*S*SUB*K/(ms_total*1e-3)
    last = (K*SUB-1) % 2
    gathered = peer.local[:, last].clone() if (world > 1 and rank == 0) else None

    # ---- roofline of the dominant kernel (the sweep kernel alone) and the step breakdown ---------------------------
    flip = [0]

    def alternate(method, *a):
        def fn():
            getattr(pipe.plans[flip[0] & 1], method)(*a)
            flip[0] += 1
        return fn
    reps = 40
    quad_ms = event_time(alternate('quadrature', 1), reps)
    epi_ms = event_time(alternate('quadrature', 2), reps)
    weights_ms = event_time(alternate('weights'), reps)
    serial_ms = event_time(alternate('run'), reps)
    alg_bytes = plan0.stream_kernel_bytes()
    step_bytes = plan0.algorithmic_bytes()
    achieved = alg_bytes/(quad_ms*1e-3)/1e9
    sub_ms = ms_total/(K*SUB)
    roofline = dict(bound='hbm', kernel='hm_quad_rows_kernel<2,5>', achieved=achieved, peak=hbm_peak, unit='GB/s',
                    frac=achieved/hbm_peak, traffic=None, algorithmic_bytes_per_launch=alg_bytes, kernel_ms=quad_ms,
                    weights_kernel_ms=weights_ms, epilogue_kernel_ms=epi_ms,
                    epilogue_note='fused into the sweep kernel under default flags: the separate epilogue stage is a no-op here',
                    serial_sub_batch_ms=serial_ms,
                    pipelined_sub_batch_ms=sub_ms, step_algorithmic_bytes=step_bytes,
                    step_achieved=step_bytes/(sub_ms*1e-3)/1e9, step_frac=step_bytes/(sub_ms*1e-3)/1e9/hbm_peak,
                    peak_source=peak_src,
                    traffic_note='dram bytes of this kernel: profiles/r2_c3_rows_ncu.txt (ncu --set full); not re-measured in-run')

    # ---- the gather against the NCCL baseline (first NCCL collective of the process, after the timed region) ------
    gather_ok = None
    if world > 1:
        mine = pipe.plans[last].run()[0].clone()
        torch.cuda.synchronize()
        ref_all = sweep.gather_spectra(mine, world*B)
        if rank == 0:
            gather_ok = bool(torch.equal(gathered.reshape(ref_all.shape), ref_all))
            if not gather_ok:
                log('gather mismatch')
        # NCCL raised cudaLimitStackSize when it created its communicator, which alone costs the HBM-bound kernels
        # 2-3.5 % (tools/dev_stack_probe.py): put it back for the legs below
        stack_after_nccl = engine.stack_limit(engine.DEFAULT_STACK_LIMIT)
        log('cudaLimitStackSize after the NCCL gather: %d B, reset to %d' % (stack_after_nccl, engine.DEFAULT_STACK_LIMIT))

    # ---- configs block --------------------------------------------------------------------------------------------
    configs = {}
    dfma_tf = dmma_tf = None
    if not args.no_configs:
        dfma_tf, dmma_tf = engine.fp64_peaks()
        configs['fp64_peaks'] = dict(dfma_tflops=dfma_tf, dmma_tflops=dmma_tf, how='hm_measure_fp64_peaks: 8 independent FMA / '
                                     'mma.sync.m8n8k4.f64 chains per thread, one CTA per SM, best of 8/16/32 warps, CUDA events, this run')
        # C3 strong: config 3 itself (4096 samples x 5) split over the ranks
        n_sub = max(1, 4096//(G*world))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps3 = 12
        s0.record()
        for _ in range(reps3*n_sub):
            sub_batch()
        pipe.drain()
        s1.record()
        barrier()
        c3_ms = max_over_ranks(s0.elapsed_time(s1))/reps3
        configs['C3'] = dict(workload='4096 samples x 5 mass functions in total, sharded over %d GPU(s): %d sub-batch(es) of %d samples per GPU'
                             % (world, n_sub, G), scaling='strong', ms_per_sweep=c3_ms, value=4096*5*S/(c3_ms*1e-3), unit=UNIT,
                             kernel='hm_quad_rows_kernel<2,5>', kernel_ms=quad_ms, algorithmic_bytes=alg_bytes,
                             frac=achieved/hbm_peak, step_frac=roofline['step_frac'], bound='hbm')
    if not args.no_configs and rank == 0 and world == 1:
        configs.update(measure_other_configs(args, engine, workloads, pipeline, syn, halo, event_time, hbm_peak, dfma_tf, dmma_tf))
    if not args.no_configs:
        configs['C5'] = measure_c5(args, rank, world, pipeline, sweep, syn, barrier, max_over_ranks, dfma_tf,
                                   configs.get('C5'))

    # ---- end to end through the public API with host buffers ------------------------------------------------------
    e2e = measure_e2e(args, rank, world, sweep, syn, halo, engine, barrier, max_over_ranks, S)

    # ---- CPU baseline and the full-array parity gate (rank 0, N = 1 only) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_and_parity(args, batches[0], plan0, engine)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total/K,
                    higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic', config=CONFIG,
                    run=dict(samples_per_gpu_per_sub_batch=G, sub_batches_per_step=SUB, models_per_gpu_per_step=B*SUB,
                             timed_region_s=ms_total*1e-3, l2='inputs larger than L2: %.0f MB of grids per sub-batch, two distinct '
                             'sub-batches alternate' % (G*2*NK*NM*8/1e6),
                             pipelining='consecutive sub-batches alternate between two CUDA streams (double-buffered plans)',
                             collective=('gather to rank 0: copy-engine push of each sub-batch\'s spectra into a symmetric '
                                         'buffer over NVLink; verified against NCCL all_gather: %s' % gather_ok)
                             if world > 1 else 'none', numa_bound_cpus=numa_cpus),
                    clocks=clocks, e2e=e2e, gpu_launches=2*SUB*K, roofline=roofline, cpu_baseline=cpu, configs=configs)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
def measure_other_configs(args, engine, workloads, pipeline, syn, halo, event_time, hbm_peak, dfma_tf, dmma_tf):
    """BASELINE configs 1, 2 and 4 on one GPU: kernel times, algorithmic bytes / flops, roofline fractions."""
    import torch
    out = {}
    # C1: single model, ST99, matter only, 128 x 256: latency of one call as a captured CUDA graph (3 launches)
    b1 = workloads.build(1, 128, 256, ('m',), ('Sheth & Tormen (1999)',), seed=1)
    p1 = b1.plan()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        p1.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            p1.run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        c1_us = e0.elapsed_time(e1)/200*1e3
    eager_us = event_time(p1.run, 200)*1e3
    out['C1'] = dict(workload='ST99, matter NFW, 128 k x 256 M, single model', latency_us_graph=c1_us, latency_us_eager=eager_us,
                     value=1e6/c1_us, unit=UNIT, algorithmic_bytes=p1.algorithmic_bytes(), bound='latency (3 dependent launches)',
                     frac=p1.algorithmic_bytes()/(c1_us*1e-6)/1e9/hbm_peak)
    del b1, p1
    # C2: Tinker10, m + g + p, 512 x 2048, batch of 64 models per step
    b2 = workloads.build(64, 512, 2048, ('m', 'g', 'p'), seed=1234)
    plans = iter([b2.plan(), b2.plan()])
    pp = engine.PipelinedPlans(lambda: next(plans))
    p2 = pp.plans[0]
    q_ms = event_time(lambda: p2.quadrature(1), 40)
    w_ms = event_time(p2.weights, 40)
    ep_ms = event_time(lambda: p2.quadrature(2), 40)

    def piped():
        pp.submit()
    for _ in range(6):
        piped()
    pp.drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        piped()
    pp.drain()
    e1.record()
    torch.cuda.synchronize()
    st_ms = e0.elapsed_time(e1)/200
    ab = p2.stream_kernel_bytes()
    out['C2'] = dict(workload='Tinker10, matter + galaxies + pressure, 6 unique spectra, 512 k x 2048 M, 64 models per step',
                     value=64*p2.S/(st_ms*1e-3), unit=UNIT, ms_per_step=st_ms, kernel='hm_quad_stream_kernel<3,3,2>', kernel_ms=q_ms,
                     weights_kernel_ms=w_ms, epilogue_kernel_ms=ep_ms, algorithmic_bytes=ab, achieved_gbs=ab/(q_ms*1e-3)/1e9,
                     frac=ab/(q_ms*1e-3)/1e9/hbm_peak, step_frac=p2.algorithmic_bytes()/(st_ms*1e-3)/1e9/hbm_peak, bound='hbm')
    del b2, pp, p2
    torch.cuda.empty_cache()
    # the config-3 sweep on config 2's three tracers (not a BASELINE config: the widened sweep kernel, one row per lane)
    b3 = workloads.build(256, NK, NM, ('m', 'g', 'p'), tuple(syn.FAMILY_NAMES), seed=99)
    p3 = b3.plan()
    q_ms, st_ms = event_time(lambda: p3.quadrature(1), 20), event_time(p3.run, 20)
    ab = p3.stream_kernel_bytes()
    out['C3_three_tracers'] = dict(workload='256 samples x 5 mass functions, matter + galaxies + pressure, 6 unique spectra per model, '
                                   '256 k x 1024 M', value=p3.B*p3.S/(st_ms*1e-3), unit=UNIT, ms_per_step=st_ms,
                                   kernel='hm_quad_rows_kernel<3,5>', kernel_ms=q_ms, algorithmic_bytes=ab,
                                   frac=ab/(q_ms*1e-3)/1e9/hbm_peak, step_frac=p3.algorithmic_bytes()/(st_ms*1e-3)/1e9/hbm_peak,
                                   bound='hbm')
    del b3, p3
    torch.cuda.empty_cache()
    # C4:64 HOD tracers, 2080 unique spectra, 256 x 2048, Gram contraction on the FP64 tensor cores
    b4 = workloads.build_many_tracer(64, 256, 2048)
    p4 = b4.gram_plan()
    p4.run()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        p4.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            p4.run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        call_us = e0.elapsed_time(e1)/50*1e3
    gram_us = event_time(p4.gram_only, 50)*1e3 if hasattr(p4, 'gram_only') else None
    fl = p4.gram_flops()
    out['C4'] = dict(workload='64 Zheng05 HOD tracers (NFW), 2080 unique spectra, 256 k x 2048 M, one model', call_us=call_us,
                     value=p4.S/(call_us*1e-6), unit=UNIT, kernel='hm_gram_kernel', kernel_us=gram_us, flops=fl,
                     algorithmic_bytes=p4.algorithmic_bytes(), bound='fp64_dmma',
                     achieved_tflops=(fl/(gram_us*1e-6)/1e12) if gram_us else None,
                     frac=(fl/(gram_us*1e-6)/1e12/dmma_tf) if (gram_us and dmma_tf) else None,
                     hbm_frac=(p4.algorithmic_bytes()/(gram_us*1e-6)/1e9/hbm_peak) if gram_us else None)
    del b4, p4
    torch.cuda.empty_cache()
    # sigma(R) kernel alone (93 % of config 5 in round 1): terms per second against the DFMA peak
    kg, dlnk = syn.sigma_k_grid()
    Gs = 128
    kgd = engine.to_dev(kg)
    Pg = engine.to_dev(np.stack([syn.LinearSpectrum(0.3)(kg)]*Gs))
    R = engine.to_dev(np.stack([np.logspace(-1.3, 1.7, NM)]*Gs))
    sig_ms = event_time(lambda: engine.sigma_tophat(kgd, Pg, dlnk, R), 5)
    terms = Gs*NM*len(kg)
    out['C5'] = dict(sigma_kernel='hm_sigma_tophat_kernel', sigma_kernel_ms=sig_ms, sigma_terms_per_s=terms/(sig_ms*1e-3),
                     sigma_terms=terms)
    return out


def measure_c5(args, rank, world, pipeline, sweep, syn, barrier, max_over_ranks, dfma_tf, extra):
    """BASELINE config 5: end to end from P_lin (sigma(R) quadrature -> mass function -> NFW Si/Ci profiles -> spectra),
    --c5-samples MCMC samples per GPU (2048 per GPU x 8 GPUs = the 16384 of BASELINE.json), sharded by
    sweep.ShardedSweep; the spectra stay on the owning GPU (they are gathered in the tests, not in the timed loop)."""
    import torch
    n = args.c5_samples
    k, M = syn.k_grid(NK), syn.M_grid(NM)
    samples = pipeline.draw_samples(n*world, seed=4321)
    pipe5 = pipeline.EndToEnd(k, M, chunk=256)
    shard = sweep.ShardedSweep(n*world, lambda lo, hi: pipeline.EndToEndShard(pipe5, samples[lo:hi]), models_per_sample=1,
                               tail_shape=(3, 3, NK))
    pipe5._chunk(samples[:256])                      # warm-up (allocator, kernels)
    barrier()
    t0 = time.perf_counter()
    shard.run_local()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter()-t0)
    barrier()
    d = dict(extra or {})
    d.update(workload='end to end from P_lin, Tinker10, matter + galaxies, 256 k x 1024 M: %d samples per GPU on %d GPU(s) (%d in total; '
                      'BASELINE config 5 is 16384 on 8)' % (n, world, n*world), samples=n*world, seconds=dt,
             samples_per_s=n*world/dt, value=3*n*world/dt, unit=UNIT, projected_seconds_16384_on_8_gpus=16384/8/(n/dt),
             bound='fp64 (sigma(R) top-hat sum: sincos)')
    if d.get('sigma_terms_per_s') and dfma_tf:
        d['sigma_note'] = 'sigma kernel: %.3g top-hat terms/s; DFMA peak %.1f TFLOP/s => %.0f FP64 flop-slots per term' \
            % (d['sigma_terms_per_s'], dfma_tf, dfma_tf*1e12/d['sigma_terms_per_s'])
    return d


def measure_e2e(args, rank, world, sweep, syn, halo, engine, barrier, max_over_ranks, S):
    """Spectra/s through the public API with HOST inputs; H2D and D2H inside the timed region."""
    import torch
    from pyhalomodel_b200 import cosmology
    n = args.e2e_samples
    k, M = syn.k_grid(NK), syn.M_grid(NM)
    rng = np.random.default_rng(99+rank)
    # host vectors of n samples: 16 distinct ones (numpy sigma(M) is slow) tiled, with distinct HOD draws for all
    draws = _draws(500+rank, 16, fiducial_first=False)
    kg, dlnk = syn.sigma_k_grid()
    Pg = np.stack([syn.LinearSpectrum(c['Om_m'], c['h'], c['ns'], c['sigma8'])(kg, c['z']) for c, _ in draws])
    RL = np.stack([cosmology.Lagrangian_radius(M, c['Om_m']) for c, _ in draws])
    sigmas = engine.to_host(engine.sigma_tophat(kg, Pg, dlnk, RL))        # set-up only (sigma(M) is an INPUT of this workload)
    base = [host_case(p, sigma=sg) for p, sg in zip(draws, sigmas)]
    pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    z, Om, Dv, dc = (np.empty(n) for _ in range(4))
    Pk, sig, rv, cc, N, V = pin((n, NK)), pin((n, NM)), pin((n, NM)), pin((n, NM)), pin((n, NM)), pin((n, NM))
    for i in range(n):
        cs = base[i % 16]
        z[i], Om[i], Dv[i], dc[i] = cs['z'], cs['Om_m'], cs['Dv'], cs['dc']
        Pk[i], sig[i], rv[i], cc[i] = cs['Pk'], cs['sig'], cs['rv'], cs['c']
        Nc, Ns = halo.HOD_mean(M, **syn.draw_hod(rng))
        Vc, Vs, _ = halo.HOD_variance(Nc, Ns)
        N[i], V[i] = Nc+Ns, Vc+Vs
    hs = sweep.HostSweep(k, M, FAMILIES, chunk=args.samples)
    for _ in range(2):
        res = hs.run(z, Om, Dv, dc, Pk, sig, rv, cc, N, V)
    steps = 5
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = hs.run(z, Om, Dv, dc, Pk, sig, rv, cc, N, V)
    torch.cuda.synchronize()
    wall = max_over_ranks(time.perf_counter()-t0)
    barrier()
    h2d, d2h = hs.bytes_per_sample()
    e2e = dict(value=world*n*len(FAMILIES)*S*steps/wall, unit=UNIT, h2d_bytes_per_step=h2d*n, d2h_bytes_per_step=d2h*n,
               samples_per_step=n, steps=steps, finite=bool(np.isfinite(res).all()),
               path='sweep.HostSweep.run: pinned host vectors (P_lin, sigma, r_v, c, <N>, var N per sample) -> H2D -> NFW / isothermal '
                    'windows, <N> normalisations, 5 mass functions, spectra on the device -> D2H numpy; wall clock around %d runs' % steps)
    legs = {}
    if rank == 0:
        # reference-shaped per-model API (what a user of the reference types), numpy spectra out.  "from vectors": the
        # per-sample vectors r_v, c are uploaded as CUDA tensors, so the windows never leave the device; "from host
        # grids": the two (k, M) grids come from numpy and cross PCIe for every model.
        def api_model(cs, name, grids=None):
            hmod = halo.model(cs['z'], cs['Om_m'], name=name, Dv=cs['Dv'], dc=cs['dc'])
            if grids is None:
                rv_d, c_d = torch.as_tensor(cs['rv']).cuda(), torch.as_tensor(cs['c']).cuda()
                Um = halo.window_function(k, rv_d, c_d, profile='NFW')
                Ug = halo.window_function(k, rv_d, profile='isothermal')
            else:
                Um, Ug = grids
            profs = {'m': halo.profile.Fourier(k, M, Um, amplitude=M, normalisation=hmod.rhom, mass_tracer=True),
                     'g': halo.profile.Fourier(k, M, Ug, amplitude=cs['N'], normalisation=hmod.average(M, cs['sig'], cs['N']),
                                               variance=cs['V'], discrete_tracer=True)}
            return hmod.power_spectrum(k, cs['Pk'], M, cs['sig'], profs)
        cases = base[:8]
        for cs in cases:
            api_model(cs, FAMILIES[3])
        t0 = time.perf_counter()
        for cs in cases:
            for name in FAMILIES:
                res1 = api_model(cs, name)
        torch.cuda.synchronize()
        dt = time.perf_counter()-t0
        assert isinstance(res1[0]['m-m'], np.ndarray)
        legs['per_model_api_from_vectors'] = dict(value=len(cases)*len(FAMILIES)*S/dt, unit=UNIT, models=len(cases)*len(FAMILIES),
                                                  ms_per_model=1e3*dt/(len(cases)*len(FAMILIES)),
                                                  path='model() + window_function x2 (r_v, c as CUDA tensors) + profile.Fourier x2 + average + '
                                                       'power_spectrum per model, numpy spectra out')
        grids = [(halo.window_function(k, cs['rv'], cs['c'], profile='NFW'), halo.window_function(k, cs['rv'], profile='isothermal'))
                 for cs in cases]
        t0 = time.perf_counter()
        for cs, gr in zip(cases, grids):
            for name in FAMILIES:
                api_model(cs, name, gr)
        torch.cuda.synchronize()
        dt = time.perf_counter()-t0
        legs['per_model_api_from_host_grids'] = dict(value=len(cases)*len(FAMILIES)*S/dt, unit=UNIT,
                                                     ms_per_model=1e3*dt/(len(cases)*len(FAMILIES)), h2d_bytes_per_model=2*NK*NM*8,
                                                     path='as above with the two (k, M) grids uploaded from numpy per model')
    e2e['legs'] = legs
    return e2e


def cpu_and_parity(args, batch, plan, engine):
    """The reference arm's step on a bounded sample of the SAME inputs the GPU just timed, and the parity gate: every
    value of every spectrum of those samples (S x 3 x nk per model) against the CPU result, rtol 1e-10."""
    import torch
    cores = os.cpu_count() or 1
    n = args.cpu_samples or min(cores, 32)
    n = min(n, batch.G)
    k, M = engine.to_host(batch.k), engine.to_host(batch.M)
    rv, c = engine.to_host(batch.rv), engine.to_host(batch.c)
    N, V = engine.to_host(batch.tracers[1].amplitude), engine.to_host(batch.tracers[1].variance)
    Pk, sig = engine.to_host(batch.Pk_lin), engine.to_host(batch.sigma)
    cases = []
    for g in range(n):
        meta = batch.meta[g]
        cases.append(dict(k=k, M=M, Pk=Pk[g], sig=sig[g], z=meta['cosmo']['z'], Om_m=meta['cosmo']['Om_m'], dc=meta['dc'],
                          Dv=meta['Dv'], rv=rv[g], c=c[g], N=N[g], V=V[g]))
    cpu, outs = cpu_baseline(cases, min(cores, n))
    got = plan.run()[0]
    torch.cuda.synchronize()
    got = got.cpu().numpy().reshape(batch.G, len(FAMILIES), 3, 3, NK)[:n]
    want = np.stack(outs)
    nz = want != 0.
    err = float(np.max(np.abs(got[nz]/want[nz]-1.)))
    zero_ok = bool(np.all(np.abs(got[~nz]) <= 1e-13*np.max(np.abs(want)))) if (~nz).any() else True
    cpu['parity'] = dict(max_rel_err=err, values_compared=int(want.size), zeros_reproduced=zero_ok, rtol=1e-10,
                         what='all %d x 5 x 3 x 3 x %d values of the first %d samples of the timed batch vs the CPU arm' % (n, NK, n))
    if not (err < 1e-10 and zero_ok):
        raise SystemExit('bench: GPU spectra differ from the CPU reference on the timed inputs (max rel %.3e)' % err)
    return cpu


if __name__ == '__main__':
    main()
