"""Developer timing of the ingredient kernels at config-3/5 shapes: sigma(R), NFW / isothermal windows, halo averages."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyhalomodel_b200 import engine, synthetic as syn, workloads
G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nk, nM = 256, 1024
batch = workloads.build(G, nk, nM, ('m', 'g'), syn.FAMILY_NAMES)
kg, dlnk = syn.sigma_k_grid()
kgd = engine.to_dev(kg)
Pg = engine.to_dev(np.stack([syn.LinearSpectrum(0.3)(kg)]*G))
R = engine.to_dev(np.stack([np.logspace(-2, 1.5, nM)]*G))
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
ts = t(lambda: engine.sigma_tophat(kgd, Pg, dlnk, R))
tn = t(lambda: engine.window_nfw(batch.k, batch.rv, batch.c))
ti = t(lambda: engine.window_isothermal(batch.k, batch.rv))
N = batch.tracers[1].amplitude
mdev = engine.pack_models([m.descriptor() for m in batch.models])
ta = t(lambda: engine.average_batch(mdev, batch.M, batch.sigma, N, models_per_sample=5))
print('G=%d: sigma %.2f ms (%.3g terms/s), NFW window %.3f ms (%.3g elem/s), iso window %.3f ms, average_batch %.3f ms'
      % (G, ts, G*nM*1e5/ts*1e3, tn, G*nk*nM/tn*1e3, ti, ta))
