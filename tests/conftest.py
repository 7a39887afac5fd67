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


def assert_spectra_close(got, want, rtol=1e-10, what='', zero_scale=None):
    """|got/want - 1| <= rtol.  Where the reference is identically 0.0 a relative test is undefined: such rows are
    differences of equal integrals (e.g. N^2 + variance - shot noise for Bernoulli centrals), so they are held to
    1e-13 of `zero_scale` (the magnitude of the spectra of the same pair) -- the exactly-zero central-central
    one-halo term under default flags (pyhalomodel.py:468-473) is additionally asserted bitwise in the tests."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    zero = want == 0.
    if zero.any():
        scale = zero_scale if zero_scale is not None else (np.max(np.abs(want)) if np.any(~zero) else 0.)
        assert np.all(np.abs(got[zero]) <= 1e-13*scale), (what, 'reference zeros not reproduced', np.max(np.abs(got[zero])))
    if (~zero).any():
        err = np.max(np.abs(got[~zero]/want[~zero]-1.))
        assert err <= rtol, (what, 'max rel err %.3e > %.1e' % (err, rtol))
