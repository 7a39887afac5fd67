"""
pyhalomodel_b200 -- B200-native (sm_100a) engine for the hot path of alexander-mead/pyhalomodel.

Drop-in surface (pyhalomodel/__init__.py:2-3 exports exactly `model` and `profile`):
    import pyhalomodel_b200 as halo            # halo.model, halo.profile
    import pyhalomodel_b200.pyhalomodel as halo  # plus window_function, matter_profile, concentration, HOD_*, ...
    import pyhalomodel_b200.cosmology as cosmology

The arithmetic lives in libhalob200.so (csrc/*.cu, C ABI in include/halob200.h), loaded lazily with ctypes on the
first compute call; importing the package needs neither the library nor a GPU, computing needs both.
"""
import sys as _sys

from . import halomodel as pyhalomodel   # reference module name: `import pyhalomodel.pyhalomodel as halo`
from . import cosmology, constants, utility
from .halomodel import model, profile

_sys.modules[__name__+'.pyhalomodel'] = pyhalomodel

__all__ = ['model', 'profile', 'pyhalomodel', 'cosmology', 'constants', 'utility']
__version__ = '0.1.0'
