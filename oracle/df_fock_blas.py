"""ctypes front of oracle/df_fock_blas.c -- TEST INFRASTRUCTURE ONLY (bench.py's CPU legs, tests/).

The C file restates the reference's ``build_fock_df`` (mqc_libcint_rhf.f90:1576-1646) with its
own loop structure on a vendor ``dgemm``; this module hands it the Fortran-ABI ``dgemm`` of the
OpenBLAS SciPy ships (``scipy.linalg.cython_blas``) and pins the BLAS thread count with
``threadpoolctl`` -- explicitly, because ``torch.distributed.run`` exports OMP_NUM_THREADS=1.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libdf_fock_blas.so")
_lib = None
_dgemm = None


def _load():
    global _lib, _dgemm
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        import subprocess
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = ctypes.CDLL(LIB_PATH)
    dp = ctypes.c_void_p
    lib.df_jk_blas.restype = ctypes.c_int
    lib.df_jk_blas.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                               ctypes.c_double, dp, dp, dp, dp]
    lib.df_assemble.restype = None
    lib.df_assemble.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp]
    import scipy.linalg.cython_blas as cb
    capsule = cb.__pyx_capi__["dgemm"]
    ctypes.pythonapi.PyCapsule_GetName.restype = ctypes.c_char_p
    ctypes.pythonapi.PyCapsule_GetName.argtypes = [ctypes.py_object]
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    _dgemm = ctypes.pythonapi.PyCapsule_GetPointer(capsule, ctypes.pythonapi.PyCapsule_GetName(capsule))
    _lib = lib
    return lib


def blas_threads(n_threads: int):
    """Context manager pinning every loaded BLAS to ``n_threads`` (1 = sequential)."""
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=int(n_threads), user_api="blas")


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(None)


def jk_df(b, density, coeff, n_occ, k_alpha=2.0, want_j=True, want_k=True):
    lib = _load()
    n = int(round(np.sqrt(b.shape[0])))
    naux = b.shape[1]
    b = np.asfortranarray(b, dtype=np.float64)
    density = np.asfortranarray(density, dtype=np.float64) if want_j else None
    coeff = np.asfortranarray(coeff, dtype=np.float64)
    j = np.empty((n, n), order="F") if want_j else None
    k = np.empty((n, n), order="F") if want_k else None
    cvec = np.empty(naux)
    work = np.empty(2 * n * max(n_occ, 1))
    lib.df_jk_blas(_dgemm, n, naux, _p(b), _p(density), _p(coeff), coeff.shape[0], int(n_occ), float(k_alpha),
                   _p(j), _p(k), _p(cvec), _p(work))
    return j, k, cvec


def build_fock_df(h, b, density, coeff, n_occ, k_scale=None, j_scale=None):
    """``build_fock_df`` (rhf.f90:1576-1646): F = H + jf*J - kf*K, kf = 0.5 (*k_scale), jf = 1 (or j_scale)."""
    lib = _load()
    j, k, _ = jk_df(b, density, coeff, n_occ, 2.0)
    n = j.shape[0]
    fock = np.empty((n, n), order="F")
    h = np.asfortranarray(h, dtype=np.float64)
    lib.df_assemble(n, _p(h), _p(j), _p(k), 1.0 if j_scale is None else float(j_scale),
                    0.5 * (1.0 if k_scale is None else float(k_scale)), _p(fock))
    return fock


def build_fock_df_uhf(h, b, d_alpha, d_beta, c_alpha, n_alpha, c_beta, n_beta, k_scale=None):
    """Two-spin twin: F_s = H + J[Da+Db] - k_scale*K[C_s], K_s without the factor 2."""
    lib = _load()
    kf = 1.0 if k_scale is None else float(k_scale)
    j, ka, _ = jk_df(b, np.asarray(d_alpha) + np.asarray(d_beta), c_alpha, n_alpha, 1.0)
    n = j.shape[0]
    h = np.asfortranarray(h, dtype=np.float64)
    fa, fb = np.empty((n, n), order="F"), np.empty((n, n), order="F")
    lib.df_assemble(n, _p(h), _p(j), _p(ka), 1.0, kf, _p(fa))
    kb = None
    if n_beta > 0:
        _, kb, _ = jk_df(b, None, c_beta, n_beta, 1.0, want_j=False)
    lib.df_assemble(n, _p(h), _p(j), _p(kb), 1.0, kf, _p(fb))
    return fa, fb
