"""Developer timing of the many-tracer (Gram/DMMA) path at BASELINE config 4 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nM = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
batch = workloads.build_many_tracer(P, nk, nM)
if len(sys.argv) > 4 and sys.argv[4] == 'shared':      # experiment: all tracers read ONE grid allocation
    for t in batch.tracers[1:]:
        t.grid = batch.tracers[0].grid
plan = batch.gram_plan()
plan.run(); torch.cuda.synchronize()
def t(fn, reps=30):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
us, gus = t(plan.run), t(plan.gram_only)
print('P=%d nk=%d nM=%d: %.1f us per call (3 kernels), Gram kernel alone %.1f us -> %.2f TFLOP/s useful, %.0f GB/s algorithmic'
      % (P, nk, nM, us, gus, plan.gram_flops()/gus/1e6, plan.algorithmic_bytes()/gus/1e3))
