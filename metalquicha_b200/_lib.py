"""ctypes binding of libmqcb200.so -- the C ABI declared in include/mqcb200.h.

The library is built in-tree by ``__graft_entry__.build()`` (or
``make -C metalquicha_b200/csrc``).  There is no fallback: if the shared object is
missing, importing the engine raises with the build command to run.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmqcb200.so")

MQCB200_OK = 0
MQCB200_FAIL = 1
MQCB200_BAD_HANDLE = 2
NUM_TIMERS = 8
TIMER_NAMES = ("upload", "j_gamma", "j_accumulate", "k_half_transform", "k_accumulate",
               "finalize", "allreduce", "download")

_dp = POINTER(c_double)

# name -> (restype, argtypes); mirrors include/mqcb200.h one to one
SIGNATURES = {
    "mqcb200_create": (c_int, [c_int, POINTER(c_void_p)]),
    "mqcb200_destroy": (c_int, [c_void_p]),
    "mqcb200_last_error": (None, [c_int, c_char_p]),
    "mqcb200_version": (c_int, []),
    "mqcb200_get_stream": (c_int, [c_void_p, POINTER(c_void_p)]),
    "mqcb200_set_workspace_limit": (c_int, [c_void_p, c_size_t]),
    "mqcb200_set_fuse_threshold": (c_int, [c_void_p, c_size_t]),
    "mqcb200_set_overlap": (c_int, [c_void_p, c_int]),
    "mqcb200_tensor_shape": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mqcb200_last_set_tensor": (c_int, [c_void_p, _dp, _dp]),
    "mqcb200_set_tensor": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "mqcb200_set_tensor_shard": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mqcb200_set_tensor_from_3c": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mqcb200_metric_inverse_sqrt": (c_int, [c_void_p, c_int, c_void_p, c_double, c_void_p, POINTER(c_int)]),
    "mqcb200_build_df_tensor": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_void_p]),
    "mqcb200_whiten_begin": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mqcb200_whiten_push": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, ctypes.c_longlong]),
    "mqcb200_whiten_end": (c_int, [c_void_p, c_int]),
    "mqcb200_last_metric": (c_int, [c_void_p, _dp, POINTER(c_int)]),
    "mqcb200_synth_tensor": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_uint64, c_double]),
    "mqcb200_clear_tensor": (c_int, [c_void_p, c_int]),
    "mqcb200_tensor_bytes": (c_int, [c_void_p, c_int, POINTER(c_size_t)]),
    "mqcb200_build_fock": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_double, c_double, c_void_p]),
    "mqcb200_build_jk": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mqcb200_build_jk_uhf": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                     c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mqcb200_build_fock_uhf": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_void_p, c_int, c_int, c_double, c_void_p, c_void_p]),
    "mqcb200_build_g_two_factor": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                           c_void_p, c_int, c_int, c_double, c_double, c_void_p]),
    "mqcb200_response_operator": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                          c_double, c_void_p]),
    "mqcb200_fitted_potential_general": (c_int, [c_void_p, c_int, c_void_p, c_double, c_void_p]),
    "mqcb200_scf_fragment": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_int,
                                     c_double, _dp, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "mqcb200_scf": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_int,
                            c_double, _dp, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "mqcb200_scf_fragment_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double,
                                           c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mqcb200_set_scf_check_every": (c_int, [c_void_p, c_int]),
    "mqcb200_df_gradient_densities": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                              c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p]),
    "mqcb200_last_energy": (c_int, [c_void_p, POINTER(c_double)]),
    "mqcb200_build_fock_device": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                          c_double, c_double, c_void_p, c_int]),
    "mqcb200_build_fock_uhf_device": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                              c_void_p, c_int, c_double, c_void_p, c_void_p, c_int]),
    "mqcb200_comm_unique_id": (c_int, [c_char_p]),
    "mqcb200_comm_init": (c_int, [c_void_p, c_int, c_int, c_char_p]),
    "mqcb200_comm_destroy": (c_int, [c_void_p]),
    "mqcb200_queue_create": (c_int, [POINTER(c_int64), c_int64, POINTER(c_void_p)]),
    "mqcb200_queue_pop": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int)]),
    "mqcb200_queue_is_empty": (c_int, [c_void_p, POINTER(c_int)]),
    "mqcb200_queue_destroy": (c_int, [c_void_p]),
    "mqcb200_set_profiling": (c_int, [c_void_p, c_int]),
    "mqcb200_last_timings": (c_int, [c_void_p, _dp]),
    "mqcb200_last_launches": (c_int, [c_void_p, POINTER(c_int)]),
    "mqcb200_last_whiten": (c_int, [c_void_p, _dp, _dp]),
    "mqcb200_last_gamma_fused": (c_int, [c_void_p, POINTER(c_int)]),
}

_lib = None


class EngineLibraryMissing(ImportError):
    pass


def load() -> ctypes.CDLL:
    """Load libmqcb200.so and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineLibraryMissing(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C metalquicha_b200/csrc`. There is no CPU fallback for this engine.")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library drift
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(1024)
    load().mqcb200_last_error(len(buf), buf)
    return buf.value.decode("utf-8", "replace")
