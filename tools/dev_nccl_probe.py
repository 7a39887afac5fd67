#!/usr/bin/env python
"""Root-cause probe for "an NCCL communicator in the process slows the HBM-bound kernels" (DESIGN.md section 6).
Single process, two visible GPUs: times the config-2 streaming kernel and the config-3 sweep kernel on GPU 0
 (a) untouched, (b) after cudaDeviceEnablePeerAccess(0 -> 1) -- what NCCL's P2P transport does at communicator
creation -- (c) after the reverse direction as well, (d) after disabling both again, and prints the device limits
NCCL could have changed.  Run: gpurun --gpus 2 -- python tools/dev_nccl_probe.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import synthetic as syn, workloads

torch.cuda.set_device(0)
rt = ctypes.CDLL('libcudart.so.12') if os.path.exists('/usr/local/cuda/lib64/libcudart.so.12') else ctypes.CDLL('libcudart.so')
b2 = workloads.build(64, 512, 2048, ('m', 'g', 'p'), seed=1)
p2 = b2.plan()
p2.run()
b3 = workloads.build(256, 256, 1024, ('m', 'g'), syn.FAMILY_NAMES, seed=2)
p3 = b3.plan()
p3.run()


def limits():
    out = {}
    for name, idx in (('stack', 0), ('printf_fifo', 1), ('malloc_heap', 2), ('max_l2_fetch_granularity', 5), ('persisting_l2', 6)):
        v = ctypes.c_size_t(0)
        rc = rt.cudaDeviceGetLimit(ctypes.byref(v), idx)
        out[name] = v.value if rc == 0 else 'err %d' % rc
    return out


def t(label):
    res = []
    for plan in (p2, p3):
        for _ in range(3):
            plan.quadrature(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30):
            plan.quadrature(1)
        b.record()
        torch.cuda.synchronize()
        res.append(1e3*a.elapsed_time(b)/30)
    print('%-52s C2 stream %.2f us   C3 sweep %.2f us   %s' % (label, res[0], res[1], limits()), flush=True)


t('untouched')
if torch.cuda.device_count() < 2:
    print('only one GPU visible: peer-access steps skipped')
    sys.exit(0)
rc = rt.cudaDeviceEnablePeerAccess(1, 0)
t('after cudaDeviceEnablePeerAccess(0 -> 1), rc=%d' % rc)
torch.cuda.set_device(1)
x = torch.zeros(1, device='cuda:1')
rc = rt.cudaDeviceEnablePeerAccess(0, 0)
torch.cuda.set_device(0)
t('after cudaDeviceEnablePeerAccess(1 -> 0) too, rc=%d' % rc)
y = torch.zeros(1 << 20, device='cuda:1')
z = y.to('cuda:0')
torch.cuda.synchronize()
t('after a peer copy 1 -> 0')
rc = rt.cudaDeviceDisablePeerAccess(1)
torch.cuda.set_device(1)
rc2 = rt.cudaDeviceDisablePeerAccess(0)
torch.cuda.set_device(0)
t('after disabling both again, rc=%d,%d' % (rc, rc2))
