#!/usr/bin/env python
"""Third probe for the NCCL side effect: NCCL raises cudaLimitStackSize from 1024 to 1872 bytes when
its first communicator is created (tools/dev_nccl_probe2.py).  Here the limit alone is changed, without NCCL, and the
two HBM-bound kernels are timed at each setting."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import synthetic as syn, workloads

torch.cuda.set_device(0)
rt = ctypes.CDLL('libcudart.so.12')
p2 = workloads.build(64, 512, 2048, ('m', 'g', 'p'), seed=1).plan()
p3 = workloads.build(256, 256, 1024, ('m', 'g'), syn.FAMILY_NAMES, seed=2).plan()
p2.run(); p3.run()


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
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), 0)
    print('%-34s stack limit %5d B   C2 stream %.2f us   C3 sweep %.2f us' % (label, v.value, res[0], res[1]), flush=True)


t('untouched')
for size in (1872, 8192, 65536, 1024, 1872, 1024):
    torch.cuda.synchronize()
    rc = rt.cudaDeviceSetLimit(0, ctypes.c_size_t(size))
    t('cudaLimitStackSize -> %d (rc %d)' % (size, rc))
