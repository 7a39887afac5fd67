"""Small end-to-end exercise of the kernels for compute-sanitizer (memcheck / racecheck): three-tracer sweep, two-tracer
sweep with ragged shapes, beta_NL with several row chunks, the fused windows, a 13-tracer Gram call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyhalomodel_b200 import workloads, engine, synthetic as syn
for tracers, G, nk, nM, F in [(('m', 'g', 'p'), 3, 70, 94, 5), (('m', 'g'), 5, 65, 130, 4), (('m',), 2, 33, 64, 2)]:
    b = workloads.build(G, nk, nM, tracers, syn.FAMILY_NAMES[:F], seed=3, fiducial_first=False)
    for kw in ({}, dict(k_trunc=0.3)):
        out, A = b.plan(**kw).run()
        assert torch.isfinite(out).all()
b = workloads.build(1, 9, 200, ('m', 'g', 'p'))
beta = torch.randn(1, 9, 200, 200, dtype=torch.float64, device='cuda')*0.1
out, A = engine.PowerSpectrumPlan(b.k, b.Pk_lin, b.M, b.sigma, b.tracers, [m.descriptor() for m in b.models], beta=beta).run()
assert torch.isfinite(out).all()
Um, Ug = engine.window_nfw_isothermal(b.k, b.rv, b.c)
assert torch.isfinite(Um).all() and torch.isfinite(Ug).all()
g = workloads.build_many_tracer(13, 20, 96)
out, A = g.gram_plan().run()
assert torch.isfinite(out).all()
torch.cuda.synchronize()
print('sanitize run ok')
