"""ctypes binding of liblda_b200.so (include/lda_b200.h).  There is no CPU fallback: if the CUDA library
is missing this raises, and every compute call fails with the library's error text when no GPU is present."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblda_b200.so")


class Schedule(C.Structure):
    _fields_ = [("S", C.c_double), ("Tau", C.c_double), ("Kappa", C.c_double)]


class Config(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("iterations", C.c_int32), ("batch_size", C.c_int32), ("burn_in_passes", C.c_int32),
        ("transformation_passes", C.c_int32), ("perplexity_evaluation_frequency", C.c_int32),
        ("change_evaluation_frequency", C.c_int32), ("processes", C.c_int32),
        ("perplexity_tolerance", C.c_double), ("mean_change_tolerance", C.c_double),
        ("alpha", C.c_double), ("eta", C.c_double), ("rho_phi", Schedule), ("rho_theta", Schedule),
    ]


class Counters(C.Structure):
    _fields_ = [("rho_phi_t", C.c_double), ("rho_theta_t", C.c_double), ("words_in_corpus", C.c_double)]


class Pcg(C.Structure):
    _fields_ = [("high", C.c_uint64), ("low", C.c_uint64)]


class Matrix(C.Structure):
    """lda_b200_matrix"""
    _fields_ = [("format", C.c_int32), ("rows", C.c_int64), ("cols", C.c_int64), ("ptr", C.c_void_p), ("ind", C.c_void_p),
                ("data", C.c_void_p)]


FORMAT_CSC, FORMAT_CSR, FORMAT_DENSE = 0, 1, 2

# every symbol include/lda_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "lda_b200_default_config": (None, [C.c_int32, C.POINTER(Config)]),
    "lda_b200_create": (C.c_int, [C.POINTER(Config), C.c_int32, C.POINTER(_P)]),
    "lda_b200_destroy": (None, [_P]),
    "lda_b200_set_config": (C.c_int, [_P, C.POINTER(Config)]),
    "lda_b200_last_error": (C.c_char_p, [_P]),
    "lda_b200_comm_unique_id": (C.c_int, [_P]),
    "lda_b200_comm_init": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "lda_b200_plan_step": (C.c_int, [C.POINTER(Schedule), C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]),
    "lda_b200_plan_shard": (C.c_int64, [C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, _P, C.c_int64]),
    "lda_b200_fit": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(Pcg), C.POINTER(Counters), _P, _P,
                               C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lda_b200_partial_fit": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg), C.POINTER(Counters),
                                       C.c_int32, _P]),
    "lda_b200_transform": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg), C.c_double, _P, _P]),
    "lda_b200_perplexity": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg), C.c_double, C.POINTER(C.c_double)]),
    "lda_b200_fit_matrix": (C.c_int, [_P, C.POINTER(Matrix), _P, _P, C.POINTER(Pcg), C.POINTER(Counters), _P, _P, C.c_int32,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lda_b200_partial_fit_matrix": (C.c_int, [_P, C.POINTER(Matrix), _P, C.POINTER(Pcg), C.POINTER(Counters), C.c_int32, _P]),
    "lda_b200_transform_matrix": (C.c_int, [_P, C.POINTER(Matrix), _P, C.POINTER(Pcg), C.c_double, _P, _P]),
    "lda_b200_perplexity_matrix": (C.c_int, [_P, C.POINTER(Matrix), _P, C.POINTER(Pcg), C.c_double, C.POINTER(C.c_double)]),
    "lda_b200_fit_flat": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(Pcg), C.POINTER(Counters),
                                    _P, _P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lda_b200_partial_fit_flat": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg),
                                            C.POINTER(Counters), C.c_int32, _P]),
    "lda_b200_transform_flat": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg), C.c_double, _P, _P]),
    "lda_b200_perplexity_flat": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg), C.c_double,
                                           C.POINTER(C.c_double)]),
    "lda_b200_save_size": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "lda_b200_save": (C.c_int, [_P, C.POINTER(Counters), _P, C.c_int64, C.POINTER(C.c_int64)]),
    "lda_b200_load": (C.c_int, [_P, _P, C.c_int64, C.POINTER(Counters)]),
    "lda_b200_components": (C.c_int, [_P, _P]),
    "lda_b200_get_dims": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "lda_b200_get_state": (C.c_int, [_P, _P, _P]),
    "lda_b200_set_state": (C.c_int, [_P, C.c_int64, _P, _P]),
    "lda_b200_pcg_seed": (None, [C.POINTER(Pcg), C.c_uint64]),
    "lda_b200_pcg_uint64": (C.c_uint64, [C.POINTER(Pcg)]),
    "lda_b200_pcg_fill": (None, [C.POINTER(Pcg), C.c_int64, C.c_int64, _P]),
    "lda_b200_pcg_advance": (None, [C.POINTER(Pcg), C.c_uint64]),
    "lda_b200_ingest_roundtrip": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_double)]),
    "lda_b200_ingest_roundtrip_matrix": (C.c_int, [_P, C.POINTER(Matrix), C.POINTER(C.c_int64), _P, _P, _P, _P,
                                                   C.POINTER(C.c_double)]),
    "lda_b200_csr_to_csc": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "lda_b200_fit_begin": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(Pcg), C.POINTER(Counters)]),
    "lda_b200_fit_iterate": (C.c_int, [_P, C.c_int32, C.POINTER(Counters), _P, C.c_int32, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "lda_b200_fit_end": (C.c_int, [_P, _P]),
    "lda_b200_set_deterministic": (C.c_int, [_P, C.c_int32]),
    "lda_b200_set_profiling": (C.c_int, [_P, C.c_int32]),
    "lda_b200_launch_count": (C.c_int64, [_P]),
    "lda_b200_last_kernel_time": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                            C.POINTER(C.c_int32)]),
    "lda_b200_comm_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lda_b200_last_allreduce_time": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "lda_b200_last_exchange_phases": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "lda_b200_probe_memory_mix": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                            C.POINTER(C.c_float)]),
}

# every symbol include/lda_b200_index.h declares
SYMBOLS.update({
    "lda_b200_index_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(_P)]),
    "lda_b200_index_destroy": (None, [_P]),
    "lda_b200_index_last_error": (C.c_char_p, [_P]),
    "lda_b200_index_add": (C.c_int, [_P, C.c_int32, C.c_int64, _P, C.c_int64, _P]),
    "lda_b200_transform_into_index": (C.c_int, [_P, C.POINTER(Matrix), _P, C.POINTER(Pcg), C.c_double, _P, _P]),
    "lda_b200_transform_into_index_flat": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(Pcg),
                                                     C.c_double, _P, _P]),
    "lda_b200_index_remove": (C.c_int, [_P, C.c_int64]),
    "lda_b200_index_len": (C.c_int64, [_P]),
    "lda_b200_index_search": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P, _P]),
    "lda_b200_index_last_search_time": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_int32)]),
})

_LIB = None


class LdaB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("nlp: lda_b200 status %d: %s" % (code, text))
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "nlp_b200: %s is missing -- build it with `python -m nlp_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(code, handle=None):
    if code != 0:
        text = lib().lda_b200_last_error(handle)
        raise LdaB200Error(code, text.decode() if text else "")


def check_index(code, index=None):
    if code != 0:
        text = lib().lda_b200_index_last_error(index)
        raise LdaB200Error(code, text.decode() if text else "")
