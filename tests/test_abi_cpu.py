"""CPU: the C-ABI library builds, loads, exports every symbol include/halob200.h declares, its ctypes mirror has
the C layout, and the product path refuses to run without a GPU (no silent CPU fallback).  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'halob200.h')


@pytest.fixture(scope='module')
def handle():
    import __graft_entry__ as ge
    ge.build()
    from pyhalomodel_b200 import _lib
    return _lib.lib()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(hm_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(handle):
    from pyhalomodel_b200 import _lib
    names = _declared()
    assert len(names) >= 14
    for name in names:
        assert hasattr(handle, name), name
    assert set(names) == set(_lib.EXPORTS)
    assert handle.hm_abi_version() == 1


def test_ctypes_layout_matches_c(tmp_path):
    from pyhalomodel_b200 import _lib
    prog = tmp_path/'layout.c'
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "halob200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(hm_model_desc),sizeof(hm_tracer),sizeof(hm_problem),offsetof(hm_problem,k_trunc),'
                    'offsetof(hm_problem,model_host),offsetof(hm_problem,tracers),offsetof(hm_tracer,grid_mode));return 0;}')
    exe = tmp_path/'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(prog), '-o', str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.ModelDesc), C.sizeof(_lib.Tracer), C.sizeof(_lib.Problem), _lib.Problem.k_trunc.offset,
            _lib.Problem.model_host.offset, _lib.Problem.tracers.offset, _lib.Tracer.grid_mode.offset]
    assert got == want


def test_header_compiles_as_plain_c(tmp_path):
    prog = tmp_path/'c.c'
    prog.write_text('#include "halob200.h"\nint main(void){return HM_OK;}\n')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), '-c', str(prog),
                    '-o', str(tmp_path/'c.o')], check=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback(handle):
    import pyhalomodel_b200 as halo
    from pyhalomodel_b200._lib import HaloB200Error
    assert handle.hm_device_count() == 0
    hmod = halo.model(0., 0.3)
    with pytest.raises(HaloB200Error):
        hmod._mass_function_nu(np.array([1., 2.]))
    with pytest.raises(HaloB200Error):
        halo.pyhalomodel.window_function(np.array([1.]), np.array([1.]), np.array([5.]), profile='NFW')
    # the C entry points themselves refuse too
    rc = handle.hm_sici(None, 0, None, None, None)
    assert rc == -4 and b'no CPU fallback' in handle.hm_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'pyhalomodel_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_model_constants_and_errors():
    """Host logic of model.__init__ (pyhalomodel.py:56-192) against the oracle (which uses scipy like the reference)."""
    import pyhalomodel_b200 as halo
    from oracle import halo_oracle as orc
    for name in orc.FAMILIES:
        for Dv in (200., 330., 177.7, 5000.):
            m, o = halo.model(0.3, 0.31, name=name, Dv=Dv, dc=1.68), orc.make_model(0.3, 0.31, name=name, Dv=Dv, dc=1.68)
            assert m.rhom == o.rhom
            d = m.descriptor()
            if name == orc.TINKER10:
                for i, key in enumerate(['alpha', 'beta', 'gamma', 'phi', 'eta', 'bA', 'ba', 'bB', 'bb', 'bC', 'bc']):
                    assert d.par[i] == float(o.par[key]), (name, Dv, key)
            elif name != orc.PS74:
                assert (d.par[0], d.par[1], d.par[2]) == (o.par['A'], o.par['q'], o.par['p'])
    with pytest.raises(ValueError):
        halo.model(0., 0.3, name='Press & Schechter (1974)')   # the reference spells it 'Schecter'
    halo.pyhalomodel.Tinker_z_dep = True
    try:
        m = halo.model(1.3, 0.3)
        o = orc.make_model(1.3, 0.3, tinker_z_dep=True)
        assert m.beta_Tinker == o.par['beta'] and m.eta_Tinker == o.par['eta']
    finally:
        halo.pyhalomodel.Tinker_z_dep = False
    from pyhalomodel_b200 import cosmology
    assert cosmology.dc_NakamuraSuto(0.3) == orc.dc_nakamura_suto(0.3)
    assert cosmology.Dv_BryanNorman(0.3) == orc.Dv_bryan_norman(0.3)
    M = np.logspace(9, 17, 33)
    assert np.array_equal(cosmology.Lagrangian_radius(M, 0.3), orc.lagrangian_radius(M, 0.3))
    assert np.array_equal(halo.pyhalomodel.concentration(M, 0.5, halo_definition='Mvir'),
                          orc.concentration(M, 0.5, halo_definition='Mvir'))
    for meth in ('simple', 'Zehavi et al. (2004)', 'Zheng et al. (2005)', 'Zhai et al. (2017)'):
        for a, b in zip(halo.pyhalomodel.HOD_mean(M, method=meth), orc.hod_mean(M, method=meth)):
            assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        halo.pyhalomodel.HOD_mean(M, method='nope')
    with pytest.raises(ValueError):
        halo.pyhalomodel.concentration(M, 0., halo_definition='nope')


def test_integration_stub_matches_the_abi():
    """The ctypes structures shown to a maintainer in INTEGRATION.md have the layout of the shipped binding."""
    import ctypes as C
    from pyhalomodel_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    code = text[text.index('class ModelDesc(C.Structure):'):text.index('lib.hm_power_spectrum_workspace.restype')]
    ns = {'C': C}
    exec(code, ns)
    for name in ('ModelDesc', 'Tracer', 'Problem'):
        stub, real = ns[name], getattr(_lib, name)
        assert C.sizeof(stub) == C.sizeof(real), name
        assert [(f[0], getattr(stub, f[0]).offset) for f in stub._fields_] == \
               [(f[0], getattr(real, f[0]).offset) for f in real._fields_], name
