"""Developer timing and accuracy of the sigma(R) kernel: G samples x 1025 radii x 1e5 wavenumbers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyhalomodel_b200 import engine, synthetic as syn, cosmology
G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kg, dlnk = syn.sigma_k_grid()
P0 = syn.LinearSpectrum(0.3)
M = syn.M_grid(1024)
R = np.concatenate([cosmology.Lagrangian_radius(M, 0.3), [8.]])
kgd = engine.to_dev(kg)
Pg = engine.to_dev(np.stack([P0(kg)*(1.+0.01*g) for g in range(G)]))
Rd = engine.to_dev(np.stack([R*(1.+0.003*g) for g in range(G)]))
out = engine.sigma_tophat(kgd, Pg, dlnk, Rd); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = engine.sigma_tophat(kgd, Pg, dlnk, Rd)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
idx = [0, 100, 300, 512, 700, 900, 1023, 1024]
want = syn.sigmaR_numpy(R[idx], P0(kg), kg, dlnk)
got = out[0].cpu().numpy()[idx]
print('sigma: G=%d %.2f ms -> %.3g terms/s; max |gpu/numpy - 1| over 8 radii = %.2e' % (G, ms, G*len(R)*len(kg)/ms*1e3, np.max(np.abs(got/want-1.))))
