"""Probe for the end-to-end path at N ranks on one box (torchrun): concurrent pinned-memory DMA bandwidth in both directions
on every rank, then sweep.HostSweep on every rank at the same time.  Tells whether the per-GPU e2e rate at N = 8 is the box's
aggregate DMA rate or something in the code."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dist.init_process_group('gloo')
def bar(): dist.barrier()
def allmax(x):
    t = torch.tensor([x], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t[0])
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=8):
    fn(); torch.cuda.synchronize(); bar(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = time.perf_counter()-t0; return allmax(dt)/reps
def h2d(): d.copy_(h, non_blocking=True)
def d2h(): h.copy_(d, non_blocking=True)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t_up, t_dn, t_bi = timed(h2d), timed(d2h), timed(both)
if rank == 0:
    print('N=%d concurrent on all ranks, per rank: H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f + %.1f GB/s (aggregate %.0f GB/s)'
          % (world, n/t_up/1e9, n/t_dn/1e9, n/t_bi/1e9, n/t_bi/1e9, 2*world*n/t_bi/1e9))
# HostSweep on every rank
from pyhalomodel_b200 import sweep, synthetic as syn, cosmology
from pyhalomodel_b200.halomodel import HOD_mean, HOD_variance, concentration
G, nk, nM = 2048, 256, 1024
k, M = syn.k_grid(nk), syn.M_grid(nM)
rng = np.random.default_rng(rank)
def pin(shape): return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
z, Om, Dv, dc = (np.empty(G) for _ in range(4))
Pk, sig, rv, c, N, V = pin((G, nk)), pin((G, nM)), pin((G, nM)), pin((G, nM)), pin((G, nM)), pin((G, nM))
kg, dlnk = syn.sigma_k_grid()
P0 = syn.LinearSpectrum(0.3, 0.7, 0.96, 0.8)
sig0 = syn.sigmaR_numpy(cosmology.Lagrangian_radius(M, 0.3), P0(kg, 0.), kg, dlnk)
Nc, Ns = HOD_mean(M, method='Zheng et al. (2005)', **syn.zheng_defaults()); Vc, Vs, _ = HOD_variance(Nc, Ns)
for g in range(G):
    zz = rng.uniform(0., 2.); z[g], Om[g] = zz, 0.3
    Omz = syn.Om_m_of_z(zz, 0.3); dc[g], Dv[g] = cosmology.dc_NakamuraSuto(Omz), cosmology.Dv_BryanNorman(Omz)
    gr = 1./(1.+zz)
    Pk[g], sig[g] = P0(k, 0.)*gr*gr, sig0*gr
    rv[g] = (3.*M/(4.*np.pi*Dv[g]*2.775e11*0.3))**(1./3.)
    c[g] = concentration(M, zz, halo_definition='Mvir'); N[g], V[g] = Nc+Ns, Vc+Vs
hs = sweep.HostSweep(k, M, syn.FAMILY_NAMES, chunk=512)
run = lambda: hs.run(z, Om, Dv, dc, Pk, sig, rv, c, N, V)
t_all = timed(run, 5)
# the same with only rank 0 working (the others idle): is it the concurrency?
run(); torch.cuda.synchronize(); bar()
if rank == 0:
    t0 = time.perf_counter()
    for _ in range(5): run()
    t_one = (time.perf_counter()-t0)/5
bar()
if rank == 0:
    print('HostSweep %d samples per rank: all %d ranks at once %.2f M spectra/s per rank; rank 0 alone %.2f M spectra/s'
          % (G, world, G*15/t_all/1e6, G*15/t_one/1e6))
dist.destroy_process_group()
