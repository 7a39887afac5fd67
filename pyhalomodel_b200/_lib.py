"""ctypes binding of libhalob200.so (include/halob200.h).  There is no CPU fallback: if the library is missing
or no CUDA device is visible, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get('HALOB200_LIB') or os.path.join(HERE, 'libhalob200.so')    # (override: kernel A/B runs)
HM_MAX_TRACERS = 8

PS74, ST99, SMT01, TINKER10, DESPALI16, TINKER10_PBS = range(6)
GRID_SEPARATE, GRID_UK, GRID_WK = 0, 1, 2


class ModelDesc(C.Structure):
    _fields_ = [('family', C.c_int32), ('reserved', C.c_int32), ('dc', C.c_double), ('rhom', C.c_double),
                ('par', C.c_double*12)]


class Tracer(C.Structure):
    _fields_ = [('grid_dev', C.c_void_p), ('wk_dev', C.c_void_p), ('amplitude_dev', C.c_void_p),
                ('variance_dev', C.c_void_p), ('normalisation_dev', C.c_void_p), ('grid_mode', C.c_int32),
                ('mass_tracer', C.c_int32), ('discrete_tracer', C.c_int32), ('reserved', C.c_int32)]


class Problem(C.Structure):
    _fields_ = [('nk', C.c_int32), ('nM', C.c_int32), ('n_tracers', C.c_int32), ('n_samples', C.c_int32),
                ('models_per_sample', C.c_int32), ('simple_twohalo', C.c_int32), ('subtract_shotnoise', C.c_int32),
                ('correct_discrete', C.c_int32), ('k_trunc', C.c_double), ('k_dev', C.c_void_p),
                ('M_dev', C.c_void_p), ('sigma_dev', C.c_void_p), ('Pk_lin_dev', C.c_void_p),
                ('models_dev', C.c_void_p), ('model_host', ModelDesc), ('tracers', Tracer*HM_MAX_TRACERS),
                ('beta_dev', C.c_void_p), ('beta_do_I11', C.c_int32), ('beta_do_I12_I21', C.c_int32),
                ('output_is_peer', C.c_int32), ('reserved2', C.c_int32)]


class GramProblem(C.Structure):
    _fields_ = [('nk', C.c_int32), ('nM', C.c_int32), ('n_tracers', C.c_int32), ('n_samples', C.c_int32),
                ('models_per_sample', C.c_int32), ('simple_twohalo', C.c_int32), ('subtract_shotnoise', C.c_int32),
                ('correct_discrete', C.c_int32), ('k_trunc', C.c_double), ('k_dev', C.c_void_p),
                ('M_dev', C.c_void_p), ('sigma_dev', C.c_void_p), ('Pk_lin_dev', C.c_void_p),
                ('models_dev', C.c_void_p), ('model_host', ModelDesc), ('tracers', C.POINTER(Tracer)),
                ('output_is_peer', C.c_int32), ('reserved2', C.c_int32)]


HM_GRAM_MAX_TRACERS = 64


class HaloB200Error(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    'hm_abi_version': (C.c_int, []),
    'hm_last_error': (C.c_char_p, []),
    'hm_device_count': (C.c_int, []),
    'hm_massfn_nu': (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'hm_low_mass_correction': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'hm_average': (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                             C.c_void_p, C.c_void_p]),
    'hm_average_batch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    'hm_multiplicity_function': (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    'hm_cross_halo_workspace': (C.c_size_t, [C.c_int32]),
    'hm_cross_halo_spectrum': (C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int32, C.POINTER(Tracer), C.POINTER(C.c_double), C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    'hm_weighted_rowsum': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'hm_power_spectrum_workspace': (C.c_size_t, [C.POINTER(Problem)]),
    'hm_power_spectrum': (C.c_int, [C.POINTER(Problem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'hm_halo_weights': (C.c_int, [C.POINTER(Problem), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'hm_mass_quadrature': (C.c_int, [C.POINTER(Problem), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'hm_mass_quadrature_stages': (C.c_int, [C.POINTER(Problem), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                            C.c_int]),
    'hm_gram_workspace': (C.c_size_t, [C.POINTER(GramProblem)]),
    'hm_gram_power_spectrum': (C.c_int, [C.POINTER(GramProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    'hm_gram_power_spectrum_stages': (C.c_int, [C.POINTER(GramProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                C.c_void_p, C.c_int]),
    'hm_sici': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'hm_window_nfw': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p]),
    'hm_window_isothermal': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    'hm_window_nfw_isothermal': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    'hm_window_configuration': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p]),
    'hm_measure_fp64_peaks': (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
    'hm_sigma_tophat': (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p]),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise HaloB200Error('libhalob200.so is not built (run `python -m pyhalomodel_b200._build` or '
                                '`python -c "import __graft_entry__ as g; g.build()"`); there is no CPU fallback')
        handle = C.CDLL(LIBPATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.hm_abi_version() != 1:
            raise HaloB200Error('libhalob200.so ABI version mismatch')
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().hm_last_error().decode('utf-8', 'replace')
        raise HaloB200Error('libhalob200 error %d: %s' % (rc, msg))
