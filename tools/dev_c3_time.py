"""Developer timing of BASELINE config 3 (sweep element: G samples x 5 mass functions, m+g, 256 k x 1024 M)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads, synthetic as syn
G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fams = syn.FAMILY_NAMES if nf == 5 else (syn.FAMILY_NAMES[3:4] if nf == 1 else syn.FAMILY_NAMES[5-nf:])
tracers = tuple(sys.argv[3]) if len(sys.argv) > 3 else ('m', 'g')          # e.g. "mgp": three tracers
batch = workloads.build(G, 256, 1024, tracers, fams)
plan = batch.plan()
plan.run(); torch.cuda.synchronize()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
tw, tq, te, ts = t(plan.weights), t(lambda: plan.quadrature(1)), t(lambda: plan.quadrature(2)), t(plan.run)
print('C3 G=%d F=%d: weights %.1f us, stream %.1f us, epilogue %.1f us, step %.1f us -> %.2f M spectra/s; grids %.0f MB; stream alg %.0f GB/s'
      % (G, len(fams), tw, tq, te, ts, plan.B*plan.S/ts, G*len(tracers)*256*1024*8/1e6, plan.stream_kernel_bytes()/tq/1e3))
if len(sys.argv) > 4:      # the same sweep as F single-family plans (per-model streaming kernel: grids read once per family)
    from pyhalomodel_b200 import engine
    F, tot = len(fams), 0.
    for f in range(F):
        trs = [engine.Tracer(x.grid, x.amplitude, engine.to_dev(x.normalisation)[f::F].contiguous(), variance=x.variance,
                             mode=x.mode, mass_tracer=x.mass_tracer, discrete_tracer=x.discrete_tracer) for x in batch.tracers]
        sp = engine.PowerSpectrumPlan(batch.k, batch.Pk_lin, batch.M, batch.sigma, trs, [m.descriptor() for m in batch.models[f::F]])
        tot += t(sp.run)
    print('  as %d single-family plans: %.1f us -> %.2f M spectra/s' % (F, tot, plan.B*plan.S/tot))
