#!/usr/bin/env python
"""
BASELINE config 3 as a sharded sweep: (cosmology, z) samples x 5 mass functions, matter + galaxies, 256 k x 1024 M,
a fixed number of samples per GPU per step (weak scaling), spectra gathered on rank 0 by the epilogue's peer
stores.  Same timing rules and JSON shape as bench.py (which stays on config 2, the configuration the metric is
quoted on); this script supplies the scaling evidence for the sweep itself.

  python tools/bench_sweep.py [--samples 512] [--steps 20] [--warmup 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sweep.py --gpus N
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

NK, NM = 256, 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--samples', type=int, default=512, help='samples per GPU per step (4096 over 8 GPUs)')
    ap.add_argument('--gather', default='dma', choices=['dma', 'stores'], help='how the spectra reach rank 0')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    out_fd = os.dup(1)
    os.dup2(2, 1)                                  # stdout carries the JSON line only
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('cpu:gloo,cuda:nccl')      # gloo control plane, NCCL only for the check (bench.py)
    from pyhalomodel_b200 import engine, sweep, synthetic as syn, workloads
    G, K, W = args.samples, args.steps, max(args.warmup, 3)
    batch = workloads.build(G, NK, NM, ('m', 'g'), syn.FAMILY_NAMES, seed=4321+rank, fiducial_first=(rank == 0))
    plan = batch.plan()
    S, B = plan.S, plan.B
    peer = sweep.PeerGatherBuffer(plan.out.shape, 2) if world > 1 else None
    pipe = engine.PipelinedPlans(batch.plan)
    step_no = [0]

    # The spectra of this configuration are large next to its compute (47 MB per 0.68 ms step and GPU): written by
    # the epilogue's peer stores they arrive on rank 0 at ~340 GB/s and the epilogue, which cannot share the SMs with
    # the persistent sweep kernel, holds up the next step (0.96 ms per step at N=8).  --gather dma keeps the epilogue's
    # stores local and lets a copy engine push the block into rank 0's buffer behind it, beside the next step.
    def step():
        if world > 1 and args.gather == 'stores':
            pipe.submit(out=peer.slot(step_no[0]), peer=True)
        elif world > 1:
            slot = peer.slot(step_no[0])
            pipe.submit(then=lambda out, lowmass: slot.copy_(out, non_blocking=True))
        else:
            pipe.submit()
        step_no[0] += 1

    token = torch.zeros(1, dtype=torch.float64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(token)
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    pipe.drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    step_no[0] = 0
    e0.record()
    for _ in range(K):
        step()
    pipe.drain()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    gathered = peer.local[:, (K-1) % peer.slots].clone() if (world > 1 and rank == 0) else None

    def time_stage(fn, reps=20):
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

    quad_ms, epi_ms, weights_ms = time_stage(lambda: plan.quadrature(1)), time_stage(lambda: plan.quadrature(2)), \
        time_stage(plan.weights)
    peaks = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak = json.load(open(peaks))['hbm_gbs'] if os.path.exists(peaks) else 6650.
    alg = plan.stream_kernel_bytes()
    gather_ok = None
    if world > 1:
        mine = plan.run()[0].clone()
        torch.cuda.synchronize()
        ref_all = sweep.gather_spectra(mine, world*B)
        if rank == 0:
            gather_ok = bool(torch.equal(gathered.reshape(ref_all.shape), ref_all))
    if rank == 0:
        line = dict(metric='halo_model_spectra_per_sec', value=world*B*S*K/(ms_total*1e-3), unit='spectra/s',
                    n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total/K, higher_is_better=True, scaling='weak',
                    dtype='f64', data='synthetic',
                    config=dict(workload='C3: sweep of (cosmology, z) samples x 5 mass functions, matter + galaxies, 3 '
                                         'unique spectra, 256 k x 1024 M; %d samples per GPU per step' % G,
                                samples_per_gpu_per_step=G, models_per_gpu_per_step=B,
                                l2='inputs larger than L2: %.0f MB of grids per step per GPU' % (G*2*NK*NM*8/1e6),
                                collective=('gather to rank 0: %s over NVLink into a symmetric buffer; '
                                            'verified against NCCL all_gather: %s'
                                            % ('copy-engine push of each step\'s spectra behind its epilogue'
                                               if args.gather == 'dma' else 'peer stores of the epilogue kernel',
                                               gather_ok)) if world > 1 else 'none'),
                    gpu_launches=3*K,
                    roofline=dict(bound='hbm', kernel='hm_quad_rows_kernel<2,5>', achieved=alg/(quad_ms*1e-3)/1e9,
                                  peak=peak, unit='GB/s', frac=alg/(quad_ms*1e-3)/1e9/peak,
                                  algorithmic_bytes_per_launch=alg, kernel_ms=quad_ms, epilogue_kernel_ms=epi_ms,
                                  weights_kernel_ms=weights_ms))
        os.write(out_fd, (json.dumps(line)+'\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
