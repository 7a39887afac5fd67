"""Physical constants in the units of the reference (pyhalomodel/constants.py:4-13).  rho_critical is formed by
the same expression so that rhom = rho_critical*Om_m is bit-identical to the reference's."""
from math import pi

G = 6.6743e-11           # [m^3 kg^-1 s^-2]
Sun_mass = 1.9884e30     # [kg]
Mpc = 3.0857e16*1e6      # [m]
H0 = 100.                # [h km/s/Mpc]
G_cosmological = G*Sun_mass/(Mpc*1e3**2)          # [(Msun/h)^-1 (km/s)^2 (Mpc/h)]
rho_critical = 3.*H0**2/(8.*pi*G_cosmological)    # [(Msun/h) (Mpc/h)^-3] ~ 2.775e11
