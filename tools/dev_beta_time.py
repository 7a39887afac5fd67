"""Developer timing of the beta_NL two-halo kernel: one model, m+g, nk x nM^2 doubles of beta streamed from HBM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads, engine
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nM = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
batch = workloads.build(1, nk, nM, ('m', 'g'))
beta = torch.randn(1, nk, nM, nM, dtype=torch.float64, device='cuda')*0.1
plan = engine.PowerSpectrumPlan(batch.k, batch.Pk_lin, batch.M, batch.sigma, batch.tracers,
                                [m.descriptor() for m in batch.models], beta=beta)
base = batch.plan()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
tb, t0 = t(plan.run), t(base.run)
nbytes = beta.numel()*8
print('beta_NL %dx%d^2: %.1f us with beta, %.1f us without -> beta kernel %.1f us, %.0f GB/s of beta (%.2f GB)'
      % (nk, nM, tb, t0, tb-t0, nbytes/(tb-t0)/1e3, nbytes/1e9))
