#!/usr/bin/env python
"""Second root-cause probe for "an NCCL communicator in the process slows the HBM-bound kernels" (DESIGN.md section 6).
Two ranks under torchrun; rank 0 times the config-2 streaming kernel and the config-3 sweep kernel
  (a) before any process group, (b) after init_process_group('nccl') alone (no communicator yet: it is created
  lazily), (c) after the first all_reduce (communicator, channels, proxy thread, peer-mapped buffers exist),
  (d) after destroy_process_group(), and prints the device limits and the SM clock next to each timing.
Run: gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dev_nccl_probe2.py"""
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pyhalomodel_b200 import synthetic as syn, workloads

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
rt = ctypes.CDLL('libcudart.so.12')
p2 = workloads.build(64, 512, 2048, ('m', 'g', 'p'), seed=1).plan()
p3 = workloads.build(256, 256, 1024, ('m', 'g'), syn.FAMILY_NAMES, seed=2).plan()
p2.run(); p3.run()


def limits():
    out = {}
    for name, idx in (('stack', 0), ('malloc_heap', 2), ('persisting_l2', 6)):
        v = ctypes.c_size_t(0)
        out[name] = v.value if rt.cudaDeviceGetLimit(ctypes.byref(v), idx) == 0 else -1
    return out


def clock():
    try:
        q = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader', '-i', str(torch.cuda.current_device())],
                           capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:       # noqa: BLE001
        q = repr(e)
    return q


def t(label):
    res = []
    for plan in (p2, p3):
        for _ in range(5):
            plan.quadrature(1)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(30):
                plan.quadrature(1)
            b.record()
            torch.cuda.synchronize()
            best = min(best, 1e3*a.elapsed_time(b)/30)
        res.append(best)
    if rank == 0:
        print('%-44s C2 stream %.2f us   C3 sweep %.2f us   %s   [%s]' % (label, res[0], res[1], limits(), clock()), flush=True)


t('(a) no process group')
dist.init_process_group('nccl', rank=rank, world_size=world)
t('(b) init_process_group(nccl), no communicator')
x = torch.ones(1 << 20, device='cuda')
dist.all_reduce(x)
torch.cuda.synchronize()
t('(c) after the first all_reduce')
for _ in range(3):
    t('(c) again')
dist.barrier()
dist.destroy_process_group()
t('(d) after destroy_process_group')
