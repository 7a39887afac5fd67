"""Where does a single power_spectrum call through the reference-shaped API spend its host time?  (config-3 shape, m + g)"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyhalomodel_b200.pyhalomodel as halo
from pyhalomodel_b200 import synthetic as syn, engine
from oracle import halo_oracle as orc
nk, nM, z, Om_m = 256, 1024, 0.3, 0.3
k, M = syn.k_grid(nk), syn.M_grid(nM)
P0 = syn.LinearSpectrum(Om_m); kg, dlnk = syn.sigma_k_grid()
Pk = P0(k, z)
Omz = syn.Om_m_of_z(z, Om_m); dc, Dv = orc.dc_nakamura_suto(Omz), orc.Dv_bryan_norman(Omz)
hmod = halo.model(z, Om_m, name='Tinker et al. (2010)', Dv=Dv, dc=dc)
sig = syn.sigmaR_numpy(hmod.Lagrangian_radius(M), P0(kg, z), kg, dlnk)
rv, c = hmod.virial_radius(M), halo.concentration(M, z, halo_definition='Mvir')
Nc, Ns = halo.HOD_mean(M); Vc, Vs, _ = halo.HOD_variance(Nc, Ns)
rvd, cd = engine.to_dev(rv), engine.to_dev(c)
def call():
    hm = halo.model(z, Om_m, name='Tinker et al. (2010)', Dv=Dv, dc=dc)
    Um = halo.window_function(k, rvd, cd, profile='NFW'); Ug = halo.window_function(k, rvd, profile='isothermal')
    rho_g = hm.average(M, sig, Nc+Ns)
    profs = {'m': halo.profile.Fourier(k, M, Um, amplitude=M, normalisation=hm.rhom, mass_tracer=True),
             'g': halo.profile.Fourier(k, M, Ug, amplitude=Nc+Ns, normalisation=rho_g, variance=Vc+Vs, discrete_tracer=True)}
    return hm.power_spectrum(k, Pk, M, sig, profs)
for _ in range(5): call()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): call()
torch.cuda.synchronize(); print('per call %.3f ms' % ((time.perf_counter()-t0)/50*1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
