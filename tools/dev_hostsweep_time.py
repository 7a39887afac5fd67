"""Developer timing of sweep.HostSweep (the bench's e2e path): samples from pinned host vectors to numpy spectra."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyhalomodel_b200 import sweep, synthetic as syn, cosmology
from pyhalomodel_b200.halomodel import HOD_mean, HOD_variance, concentration
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nk, nM = 256, 1024
k, M = syn.k_grid(nk), syn.M_grid(nM)
rng = np.random.default_rng(1)
def pin(shape):
    return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
z, Om, Dv, dc = (np.empty(G) for _ in range(4))
Pk, sig, rv, c, N, V = pin((G, nk)), pin((G, nM)), pin((G, nM)), pin((G, nM)), pin((G, nM)), pin((G, nM))
kg, dlnk = syn.sigma_k_grid()
cs0 = syn.draw_cosmology(rng); P0 = syn.LinearSpectrum(cs0['Om_m'], cs0['h'], cs0['ns'], cs0['sigma8'])
sig0 = syn.sigmaR_numpy(cosmology.Lagrangian_radius(M, cs0['Om_m']), P0(kg, 0.), kg, dlnk)
for g in range(G):          # cheap synthetic variation around one cosmology (the timing does not depend on the values)
    zz = rng.uniform(0., 2.); z[g], Om[g] = zz, cs0['Om_m']
    Omz = syn.Om_m_of_z(zz, cs0['Om_m']); dc[g], Dv[g] = cosmology.dc_NakamuraSuto(Omz), cosmology.Dv_BryanNorman(Omz)
    gr = 1./(1.+zz)
    Pk[g], sig[g] = P0(k, 0.)*gr*gr, sig0*gr
    rv[g] = (3.*M/(4.*np.pi*Dv[g]*2.775e11*cs0['Om_m']))**(1./3.)
    c[g] = concentration(M, zz, halo_definition='Mvir')
    Nc, Ns = HOD_mean(M, method='Zheng et al. (2005)', **syn.draw_hod(rng)); Vc, Vs, _ = HOD_variance(Nc, Ns)
    N[g], V[g] = Nc+Ns, Vc+Vs
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 512
hs = sweep.HostSweep(k, M, syn.FAMILY_NAMES, chunk=CH)
out = hs.run(z, Om, Dv, dc, Pk, sig, rv, c, N, V)
assert np.isfinite(out).all()
torch.cuda.synchronize(); t0 = time.perf_counter()
R = 5
for _ in range(R): out = hs.run(z, Om, Dv, dc, Pk, sig, rv, c, N, V)
dt = (time.perf_counter()-t0)/R
print('HostSweep %d samples x %d mass functions: %.2f ms per sweep -> %.2f M spectra/s (%.2f ms per 512-sample chunk)'
      % (G, len(syn.FAMILY_NAMES), dt*1e3, G*15/dt/1e6, dt*1e3/(G/512)))
