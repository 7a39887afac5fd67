"""Developer timing of the many-tracer (Gram/DMMA) path at BASELINE config 4 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nM = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
batch = workloads.build_many_tracer(P, nk, nM)
plan = batch.gram_plan()
plan.run(); torch.cuda.synchronize()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    plan.run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        plan.run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1)/20*1e3
print('P=%d nk=%d nM=%d: %.1f us per call (4 kernels) -> %.0f spectra/s, %.2f TFLOP/s useful, %.0f GB/s algorithmic'
      % (P, nk, nM, us, plan.S/us*1e6, plan.gram_flops()/us/1e6, plan.algorithmic_bytes()/us/1e3))
