"""Developer timing of the host-buffer (e2e) path for one config-2 model: where do the 0.7 ms go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyhalomodel_b200.pyhalomodel as halo
from pyhalomodel_b200 import workloads, engine
batch = workloads.build(2, 512, 2048)
k, Pk, M, sig, grids = workloads.host_profiles(batch, 0)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.float64).pin_memory(); v = t.numpy(); v[...] = a; return v
k, Pk, M, sig = pin(k), pin(Pk), pin(M), pin(sig)
g = {n: {kk: pin(v) for kk, v in d.items()} for n, d in grids.items()}
hmod = batch.models[0]; norm_g = float(batch.tracers[1].normalisation[0].item())
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
x = g['m']['grid']
print('pinned?', torch.from_numpy(x).is_pinned(), ' grid MB', x.nbytes/1e6)
print('H2D one grid via to_dev   : %.3f ms -> %.1f GB/s' % (timeit(lambda: engine.to_dev(x)), 0), )
t = timeit(lambda: engine.to_dev(x)); print('   = %.1f GB/s' % (x.nbytes/t/1e6))
dst = torch.empty(x.shape, dtype=torch.float64, device='cuda'); src = torch.from_numpy(x)
t = timeit(lambda: dst.copy_(src, non_blocking=True)); print('raw copy_ pinned->dev    : %.3f ms = %.1f GB/s' % (t, x.nbytes/t/1e6))
def profs():
    return {'m': halo.profile.Fourier(k, M, g['m']['grid'], amplitude=g['m']['amplitude'], normalisation=hmod.rhom, mass_tracer=True),
            'g': halo.profile.Fourier(k, M, g['g']['grid'], amplitude=g['g']['amplitude'], normalisation=norm_g, variance=g['g']['variance'], discrete_tracer=True),
            'p': halo.profile.Fourier(k, M, g['p']['grid'])}
print('3 x profile.Fourier       : %.3f ms' % timeit(profs))
pr = profs()
print('power_spectrum (dev profs): %.3f ms' % timeit(lambda: hmod.power_spectrum(k, Pk, M, sig, pr)))
print('full e2e per model        : %.3f ms' % timeit(lambda: hmod.power_spectrum(k, Pk, M, sig, profs())))
import cProfile, pstats
prof = cProfile.Profile()
prof.enable()
for _ in range(50): hmod.power_spectrum(k, Pk, M, sig, pr)
prof.disable()
pstats.Stats(prof).sort_stats('cumulative').print_stats(22)
