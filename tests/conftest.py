import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as f:
        return {key: f[key] for key in f.files}


@pytest.fixture(scope='session')
def golden():
    return load_golden


def assert_spectra_close(got, want, rtol=1e-10, what=''):
    """|got/want - 1| <= rtol, with an absolute floor tied to max|want| only where the reference is
    identically 0.0 (e.g. the central-central one-halo term, pyhalomodel.py:468-473)."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    zero = want == 0.
    if zero.any():
        scale = np.max(np.abs(want)) if np.any(~zero) else 1.
        assert np.all(np.abs(got[zero]) <= 1e-14*scale), (what, 'reference zeros not reproduced', np.max(np.abs(got[zero])))
    if (~zero).any():
        err = np.max(np.abs(got[~zero]/want[~zero]-1.))
        assert err <= rtol, (what, 'max rel err %.3e > %.1e' % (err, rtol))
