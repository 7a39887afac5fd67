"""GPU (-m gpu): the C ABI consumed from plain C++ (tests/cabi/cabi_host.cpp: cudaMalloc'd buffers, no Python or
torch in the loop), compared with the oracle.  This is the binding a non-Python maintainer would write."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import assert_spectra_close
from oracle import halo_oracle as orc
from pyhalomodel_b200 import synthetic as syn
from test_gpu_parity import _oracle_profiles, _synthetic_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cabi_exe(tmp_path_factory):
    from pyhalomodel_b200 import _build
    lib = _build.build()
    exe = str(tmp_path_factory.mktemp('cabi')/'cabi_host')
    libdir = os.path.dirname(lib)
    subprocess.run(['nvcc', '-O2', '-std=c++17', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'cabi', 'cabi_host.cpp'),
                    '-o', exe, '-L', libdir, '-lhalob200', '-Xlinker', '-rpath', '-Xlinker', libdir], check=True)
    return exe


@pytest.mark.parametrize('name,kw', [(orc.TINKER10, {}), (orc.ST99, dict(subtract_shotnoise=False, k_trunc=0.2)),
                                     (orc.SMT01, dict(simple_twohalo=True))])
def test_power_spectrum_from_cpp_host(cabi_exe, tmp_path, name, kw):
    import pyhalomodel_b200 as halo     # host-side constants only (model descriptor); no GPU work from Python here
    g = _synthetic_case(48, 160, z=0.5)
    hods = [syn.zheng_defaults()]
    om = orc.make_model(g['z'], g['Om_m'], name=name, Dv=g['Dv'], dc=g['dc'])
    profs = _oracle_profiles(om, g, hods, True)          # m (NFW), g0 (HOD), p (dimensionful, amplitude=None)
    want = orc.power_spectrum(om, g['k'], g['Pk_lin'], g['M'], g['sigmaM'], profs, **kw)
    d = halo.model(g['z'], g['Om_m'], name=name, Dv=g['Dv'], dc=g['dc']).descriptor()
    blob = struct.pack('<4i', len(g['k']), len(g['M']), len(profs), d.family)
    blob += struct.pack('<2d', d.dc, d.rhom)+struct.pack('<12d', *list(d.par))+struct.pack('<d', kw.get('k_trunc') or 0.)
    blob += struct.pack('<3i', int(kw.get('simple_twohalo', False)), int(kw.get('subtract_shotnoise', True)),
                        int(kw.get('correct_discrete', True)))
    for arr in (g['k'], g['M'], g['sigmaM'], g['Pk_lin']):
        blob += np.ascontiguousarray(arr, dtype='<f8').tobytes()
    for key, pr in profs.items():
        # profiles built by Fourier(): pass the single grid the kernel wants (Uk with amplitude, or Wk if amplitude=None)
        mode = 2 if key == 'p' else 1
        grid = pr.Wk if mode == 2 else pr.Uk
        blob += struct.pack('<4i', mode, int(pr.mass_tracer), int(pr.discrete_tracer), int(pr.variance is not None))
        blob += struct.pack('<d', float(pr.normalisation))
        blob += np.ascontiguousarray(grid, dtype='<f8').tobytes()+np.ascontiguousarray(pr.amplitude, dtype='<f8').tobytes()
        if pr.variance is not None:
            blob += np.ascontiguousarray(pr.variance, dtype='<f8').tobytes()
    fin, fout = tmp_path/'in.bin', tmp_path/'out.bin'
    fin.write_bytes(blob)
    r = subprocess.run([cabi_exe, str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout+r.stderr
    res = np.fromfile(fout, dtype='<f8')
    nk, names = len(g['k']), list(profs)
    S = len(names)*(len(names)+1)//2
    out, A = res[:-1].reshape(S, 3, nk), res[-1]
    assert abs(A-orc.low_mass_correction(om, om.dc/g['sigmaM'][0])) < 3e-14
    s = 0
    for iu, u in enumerate(names):
        for v in names[iu:]:
            for t in range(3):
                assert_spectra_close(out[s, t], want[t][u+'-'+v], rtol=1e-10, what='%s-%s term %d' % (u, v, t),
                                     zero_scale=np.max(np.abs(want[2][u+'-'+v])))
            s += 1
