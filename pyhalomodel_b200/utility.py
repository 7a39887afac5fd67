"""Small host helpers mirroring the two utilities the hot path uses (pyhalomodel/utility.py:39-43, 63-67)."""
import numpy as np


def logspace(xmin, xmax, nx):
    """utility.py:39-43"""
    return np.logspace(np.log10(xmin), np.log10(xmax), nx)


def is_array_monotonic(x):
    """True iff strictly increasing (utility.py:63-67)."""
    return bool(np.all(np.diff(np.asarray(x)) > 0.))
