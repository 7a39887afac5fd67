"""Developer timing: per-stage GPU times of the C2 batch with CUDA graphs (no host launch overhead)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
nM = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
tr = tuple(sys.argv[4]) if len(sys.argv) > 4 else ('m', 'g', 'p')
batch = workloads.build(B, nk, nM, tr)
plan = batch.plan()
plan.run(); torch.cuda.synchronize()

def graph_time(fn, reps=50, inner=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(inner): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps/inner*1e3

def eager_time(fn, reps=200):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3

ab = plan.algorithmic_bytes()
for name, fn in [('weights', plan.weights), ('quad stream only', lambda: plan.quadrature(1)), ('quad epilogue only', lambda: plan.quadrature(2)),
                 ('quadrature(stream+epi)', plan.quadrature), ('step', plan.run)]:
    tg, te = graph_time(fn), eager_time(fn)
    print('%-24s graph %8.2f us   eager %8.2f us' % (name, tg, te), ('  -> %.0f GB/s alg' % (ab/tg/1e3)) if 'quad' in name or name == 'step' else '')
t0 = time.perf_counter()
for _ in range(200): plan.quadrature()
print('host cost per quadrature() call: %.1f us' % ((time.perf_counter()-t0)/200*1e6)); torch.cuda.synchronize()
print('alg bytes', ab, 'B', B, 'spectra/step', B*plan.S)
