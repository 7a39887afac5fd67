"""Developer timing of BASELINE config 5 (end to end from P_lin) on one GPU: samples per second and stage split."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import pipeline, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k, M = syn.k_grid(256), syn.M_grid(1024)
samples = pipeline.draw_samples(n)
e2e = pipeline.EndToEnd(k, M, chunk=chunk)
e2e.run(samples[:chunk]); torch.cuda.synchronize()
t0 = time.perf_counter()
out = e2e.run(samples)
torch.cuda.synchronize()
dt = time.perf_counter()-t0
print('C5: %d samples in %.3f s -> %.0f samples/s, %.0f spectra/s (3 pairs) on one GPU; sigma terms/s %.3g'
      % (n, dt, n/dt, 3*n/dt, n*1025*1e5/dt))
