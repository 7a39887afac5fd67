"""Developer timing of the Fourier-window kernels on config 3's shape (G samples x 256 k x 1024 M)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads, engine, synthetic as syn
G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = workloads.build(G, 256, 1024, ('m',), syn.FAMILY_NAMES[3:4])
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
n = G*256*1024
tn, ti = t(lambda: engine.window_nfw(b.k, b.rv, b.c)), t(lambda: engine.window_isothermal(b.k, b.rv))
tf = t(lambda: engine.window_nfw_isothermal(b.k, b.rv, b.c))
print('windows G=%d: NFW %.3f ms (%.1f G elements/s), isothermal %.3f ms (%.1f G elements/s), both in one launch %.3f ms'
      % (G, tn, n/tn/1e6, ti, n/ti/1e6, tf))
