"""Host-side mirror of the reference's Fock-build interface on top of the C ABI.

Method names, argument meaning and error behaviour follow the reference routines
they stand in for (paths relative to the reference tree):

* ``build_fock_df``       -- backends/libcint/mqc_libcint_rhf.f90:1576-1646
* ``assemble_fock``       -- DF branch, mqc_libcint_rhf.f90:1089-1106, energy :1207
* ``atomic_guess_fock``   -- mqc_libcint_rhf.f90:1382-1411 (pseudo-orbitals :1413-1462)
* ``set_tensor``          -- the ``bmat`` produced by build_df_tensor,
                             backends/libcint/mqc_libcint_integrals.F90:913-990
* ``WorkQueue``           -- src/fragmentation/common/mqc_work_queue.f90:10-57

The tensor argument ``b`` of the reference routines is replaced by a resident
device copy (set once per geometry, exactly as ``run_libcint_rhf`` builds ``bmat``
once at :486-498).  Every call goes through ``libmqcb200.so``; nothing here
computes J or K on the host, and a failure in the library raises ``B200Error``
with the library's message (the reference sets ``error_t`` and returns).
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_int64, c_size_t, c_void_p

import numpy as np

from . import _lib

SLOT_FULL_RANGE = 0
SLOT_ATTENUATED = 1

OCCUPATION_FLOOR = 1.0e-12   # mqc_libcint_rhf.f90:1438


class B200Error(RuntimeError):
    """A call into libmqcb200.so returned MQCB200_FAIL / MQCB200_BAD_HANDLE."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _check(rc: int) -> None:
    if rc != _lib.MQCB200_OK:
        raise B200Error(rc, _lib.last_error())


def _f64_colmajor(a, name: str, allow_ld: bool = False) -> np.ndarray:
    """Column-major float64 matrix.  Only the coefficient matrices carry a leading dimension
    through the C ABI (``allow_ld``): every other operand is read as contiguous n*n, so a
    strided view (``big[:n, :n]``) is copied."""
    arr = np.asarray(a)
    if arr.dtype != np.float64:
        raise TypeError(f"{name} must be float64 (the reference's real(dp)), got {arr.dtype}")
    if arr.ndim != 2:
        raise ValueError(f"{name} must be a matrix")
    if not arr.flags.f_contiguous:
        if (allow_ld and arr.strides[0] == arr.itemsize and arr.strides[1] >= arr.itemsize * arr.shape[0]
                and arr.strides[1] % arr.itemsize == 0):
            return arr                    # a column-major view with a leading dimension
        arr = np.asfortranarray(arr)
    return arr


def _ptr(arr) -> c_void_p:
    return c_void_p(arr.ctypes.data) if arr is not None else c_void_p(None)


class B200FockEngine:
    """One engine handle == one GPU (``mod(device_rank, device_count)``)."""

    def __init__(self, device_rank: int = 0):
        self._lib = _lib.load()
        self._h = c_void_p(None)
        _check(self._lib.mqcb200_create(int(device_rank), byref(self._h)))
        self.n = {}          # slot -> n
        self.naux = {}       # slot -> (naux_total, q_begin, q_count)

    # -- lifetime ---------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            rc = self._lib.mqcb200_destroy(self._h)
            self._h = c_void_p(None)
            _check(rc)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> c_void_p:
        return self._h

    def stream(self) -> int:
        s = c_void_p(None)
        _check(self._lib.mqcb200_get_stream(self._h, byref(s)))
        return s.value or 0

    def set_workspace_limit(self, n_bytes: int) -> None:
        _check(self._lib.mqcb200_set_workspace_limit(self._h, c_size_t(int(n_bytes))))

    def set_fuse_threshold(self, n_bytes: int) -> None:
        _check(self._lib.mqcb200_set_fuse_threshold(self._h, c_size_t(int(n_bytes))))

    def set_overlap(self, on: bool) -> None:
        """Coulomb kernels on a second stream, concurrent with the exchange kernels (default on)."""
        _check(self._lib.mqcb200_set_overlap(self._h, 1 if on else 0))

    def tensor_shape(self, slot: int = SLOT_FULL_RANGE):
        """``(n, naux_total, q_begin, q_count)`` of the resident tensor, zeros when the slot is empty."""
        n, nt, qb, qc = c_int(0), c_int(0), c_int(0), c_int(0)
        _check(self._lib.mqcb200_tensor_shape(self._h, slot, byref(n), byref(nt), byref(qb), byref(qc)))
        return n.value, nt.value, qb.value, qc.value

    def _check_operand(self, n: int, slot: int) -> None:
        """The build calls carry no size: refuse operands that do not match the resident tensor
        (a stale tensor of another fragment/geometry would be read out of bounds)."""
        n_res = self.tensor_shape(slot)[0]
        if n_res and n_res != n:
            raise ValueError(f"the resident tensor has n = {n_res}, the operands have n = {n}")

    def last_set_tensor(self):
        """``(ms, h2d_bytes)`` of the last host -> packed tensor upload."""
        ms, by = c_double(0.0), c_double(0.0)
        _check(self._lib.mqcb200_last_set_tensor(self._h, byref(ms), byref(by)))
        return ms.value, by.value

    # -- the fitted tensor ------------------------------------------------------------
    def set_tensor(self, b, n: int | None = None, slot: int = SLOT_FULL_RANGE) -> None:
        """``b`` is the reference's ``bmat(nao*nao, naux)`` (column-major)."""
        b = _f64_colmajor(b, "b")
        if not b.flags.f_contiguous:
            b = np.asfortranarray(b)
        nn, naux = b.shape
        if n is None:
            n = int(round(np.sqrt(nn)))
        if n * n != nn:
            raise ValueError("b must have nao*nao rows")
        _check(self._lib.mqcb200_set_tensor(self._h, slot, n, naux, _ptr(b)))
        self.n[slot] = n
        self.naux[slot] = (naux, 0, naux)

    def set_tensor_shard(self, b_shard, n: int, naux_total: int, q_begin: int,
                         slot: int = SLOT_FULL_RANGE) -> None:
        """Only auxiliary functions ``[q_begin, q_begin + b_shard.shape[1])`` live here."""
        b_shard = np.asfortranarray(_f64_colmajor(b_shard, "b_shard"))
        if b_shard.shape[0] != n * n:
            raise ValueError("b_shard must have nao*nao rows")
        q_count = b_shard.shape[1]
        _check(self._lib.mqcb200_set_tensor_shard(self._h, slot, n, naux_total, q_begin, q_count,
                                                  _ptr(b_shard)))
        self.n[slot] = n
        self.naux[slot] = (naux_total, q_begin, q_count)

    def set_tensor_from_3c(self, three, half, n: int, slot: int = SLOT_FULL_RANGE) -> None:
        three = np.asfortranarray(_f64_colmajor(three, "three"))
        half = np.asfortranarray(_f64_colmajor(half, "half"))
        naux = three.shape[1]
        _check(self._lib.mqcb200_set_tensor_from_3c(self._h, slot, n, naux, _ptr(three), _ptr(half)))
        self.n[slot] = n
        self.naux[slot] = (naux, 0, naux)

    def build_df_tensor(self, three, metric, n: int, slot: int = SLOT_FULL_RANGE) -> np.ndarray:
        """Last two stages of ``build_df_tensor`` (integrals.F90:981-987), both on the device:
        ``half = metric^(-1/2)`` (one-sided Jacobi eigensolver + GEMM), ``b = three . half`` slab by
        slab straight into the packed resident layout.  Returns ``half`` (the gradient needs it)."""
        three = np.asfortranarray(_f64_colmajor(three, "three"))
        metric = np.asfortranarray(_f64_colmajor(metric, "metric"))
        naux = three.shape[1]
        if three.shape[0] != n * n or metric.shape != (naux, naux):
            raise ValueError("three must be (nao*nao, naux) and metric (naux, naux)")
        half = np.empty((naux, naux), order="F")
        _check(self._lib.mqcb200_build_df_tensor(self._h, slot, n, naux, _ptr(three), _ptr(metric),
                                                 c_double(NULL_THRESHOLD), _ptr(half)))
        self.n[slot] = n
        self.naux[slot] = (naux, 0, naux)
        return half

    def metric_inverse_sqrt(self, metric) -> np.ndarray:
        """``metric_inverse_sqrt`` (mqc_libcint_integrals.F90:992-1038) on the device; modes at or
        below 1e-10 are zeroed, a metric with no surviving mode raises the reference's message."""
        metric = np.asfortranarray(_f64_colmajor(metric, "metric"))
        naux = metric.shape[0]
        if metric.shape != (naux, naux):
            raise ValueError("the metric is square")
        half = np.empty((naux, naux), order="F")
        kept = c_int(0)
        _check(self._lib.mqcb200_metric_inverse_sqrt(self._h, naux, _ptr(metric), c_double(NULL_THRESHOLD), _ptr(half),
                                                     byref(kept)))
        self.last_metric_kept = kept.value
        return half

    def whiten_begin(self, n: int, naux_total: int, half, q_begin: int = 0, q_count: int | None = None,
                     slot: int = SLOT_FULL_RANGE) -> None:
        """Start a slab-streamed ``b = three . half`` into this rank's auxiliary slab (see mqcb200.h)."""
        half = np.asfortranarray(_f64_colmajor(half, "half"))
        if q_count is None:
            q_count = naux_total - q_begin
        _check(self._lib.mqcb200_whiten_begin(self._h, slot, n, naux_total, q_begin, q_count, _ptr(half)))
        self._whiten = (slot, n, naux_total, q_begin, q_count)

    def whiten_push(self, nu_begin: int, three_block, slot: int = SLOT_FULL_RANGE) -> None:
        """``three_block`` is ``(nao*nu_count, naux_total)`` column-major: (mu nu|P) for nu in
        ``[nu_begin, nu_begin + nu_count)``, all mu, all P."""
        _, n, naux_total, _, _ = self._whiten
        blk = np.asfortranarray(_f64_colmajor(three_block, "three_block"))
        if blk.shape[1] != naux_total or blk.shape[0] % n:
            raise ValueError("three_block must be (nao*nu_count, naux_total)")
        _check(self._lib.mqcb200_whiten_push(self._h, slot, int(nu_begin), blk.shape[0] // n, _ptr(blk),
                                             ctypes.c_longlong(blk.shape[0])))

    def whiten_end(self, slot: int = SLOT_FULL_RANGE) -> None:
        _check(self._lib.mqcb200_whiten_end(self._h, slot))
        _, n, naux_total, q_begin, q_count = self._whiten
        self.n[slot] = n
        self.naux[slot] = (naux_total, q_begin, q_count)

    def last_metric(self):
        """``(ms, sweeps)`` of the last device metric^(-1/2)."""
        ms, sw = c_double(0.0), c_int(0)
        _check(self._lib.mqcb200_last_metric(self._h, byref(ms), byref(sw)))
        return ms.value, sw.value

    def synth_tensor(self, n: int, naux_total: int, seed: int, scale: float,
                     q_begin: int = 0, q_count: int | None = None,
                     slot: int = SLOT_FULL_RANGE) -> None:
        if q_count is None:
            q_count = naux_total - q_begin
        _check(self._lib.mqcb200_synth_tensor(self._h, slot, n, naux_total, q_begin, q_count,
                                              ctypes.c_uint64(seed), c_double(scale)))
        self.n[slot] = n
        self.naux[slot] = (naux_total, q_begin, q_count)

    def clear_tensor(self, slot: int = SLOT_FULL_RANGE) -> None:
        _check(self._lib.mqcb200_clear_tensor(self._h, slot))
        self.n.pop(slot, None)
        self.naux.pop(slot, None)

    def tensor_bytes(self, slot: int = SLOT_FULL_RANGE) -> int:
        out = c_size_t(0)
        _check(self._lib.mqcb200_tensor_bytes(self._h, slot, byref(out)))
        return out.value

    # -- the Fock build -----------------------------------------------------------------
    def build_fock_df(self, h, density, coeff, n_occ: int, k_scale=None, j_scale=None,
                      slot: int = SLOT_FULL_RANGE, out=None) -> np.ndarray:
        """``build_fock_df(h, b, density, coeff, n_occ, fock, k_scale, j_scale)`` with ``b`` resident."""
        h = _f64_colmajor(h, "h")
        density = _f64_colmajor(density, "density")
        n = h.shape[0]
        if h.shape != (n, n) or density.shape != (n, n):
            raise ValueError("h and density must both be n x n")
        self._check_operand(n, slot)
        coeff, ldc = self._coeff(coeff, n, n_occ)
        if out is not None and not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (n, n)
                                    and out.flags.f_contiguous and out.flags.writeable):
            raise ValueError("out must be a writeable column-major float64 n x n array (the library writes n*n doubles into it)")
        fock = out if out is not None else np.empty((n, n), dtype=np.float64, order="F")
        _check(self._lib.mqcb200_build_fock(
            self._h, slot, _ptr(h), _ptr(density), _ptr(coeff), ldc, int(n_occ),
            c_double(1.0 if k_scale is None else k_scale),
            c_double(1.0 if j_scale is None else j_scale), _ptr(fock)))
        return fock

    def build_jk(self, density, coeff, n_occ: int, want_j: bool = True, want_k: bool = True,
                 slot: int = SLOT_FULL_RANGE):
        """J and K of ``build_fock_df`` before scaling (K carries the RHF factor 2)."""
        n = self.n[slot]
        density = _f64_colmajor(density, "density") if density is not None else None
        if density is not None and density.shape != (n, n):
            raise ValueError(f"the resident tensor has n = {n}, the density is {density.shape}")
        coeff, ldc = self._coeff(coeff, n, n_occ) if want_k else (None, n)
        j = np.empty((n, n), dtype=np.float64, order="F") if want_j else None
        k = np.empty((n, n), dtype=np.float64, order="F") if want_k else None
        _check(self._lib.mqcb200_build_jk(self._h, slot, _ptr(density), _ptr(coeff), ldc, int(n_occ),
                                          _ptr(j), _ptr(k)))
        return j, k

    def build_jk_uhf(self, density_total, coeff_a, n_alpha: int, coeff_b, n_beta: int,
                     slot: int = SLOT_FULL_RANGE):
        """J[Da+Db], K_alpha, K_beta (no factor 2); K_beta is ``None`` when n_beta == 0."""
        n = self.n[slot]
        density_total = _f64_colmajor(density_total, "density_total")
        if density_total.shape != (n, n):
            raise ValueError(f"the resident tensor has n = {n}, the density is {density_total.shape}")
        ca, lda = self._coeff(coeff_a, n, n_alpha)
        cb, ldb = self._coeff(coeff_b, n, n_beta) if n_beta > 0 else (None, n)
        j = np.empty((n, n), dtype=np.float64, order="F")
        ka = np.empty((n, n), dtype=np.float64, order="F")
        kb = np.empty((n, n), dtype=np.float64, order="F") if n_beta > 0 else None
        _check(self._lib.mqcb200_build_jk_uhf(self._h, slot, _ptr(density_total), _ptr(ca), lda,
                                              int(n_alpha), _ptr(cb), ldb, int(n_beta),
                                              _ptr(j), _ptr(ka), _ptr(kb)))
        return j, ka, kb

    def build_fock_df_uhf(self, h, d_alpha, d_beta, coeff_a, n_alpha: int, coeff_b, n_beta: int,
                          k_scale=None, slot: int = SLOT_FULL_RANGE):
        """F_sigma = H + J[Da+Db] - k_scale*K[C_sigma]."""
        h = _f64_colmajor(h, "h")
        n = h.shape[0]
        self._check_operand(n, slot)
        dt = np.asfortranarray(np.asarray(d_alpha) + np.asarray(d_beta))
        if h.shape != (n, n) or dt.shape != (n, n):
            raise ValueError("h and the spin densities must all be n x n")
        ca, lda = self._coeff(coeff_a, n, n_alpha)
        cb, ldb = self._coeff(coeff_b, n, n_beta) if n_beta > 0 else (None, n)
        fa = np.empty((n, n), dtype=np.float64, order="F")
        fb = np.empty((n, n), dtype=np.float64, order="F")
        _check(self._lib.mqcb200_build_fock_uhf(self._h, slot, _ptr(h), _ptr(dt), _ptr(ca), lda, int(n_alpha),
                                                _ptr(cb), ldb, int(n_beta),
                                                c_double(1.0 if k_scale is None else k_scale),
                                                _ptr(fa), _ptr(fb)))
        return fa, fb

    def _g_two_factor(self, density, ca, n_a, cb, n_b, ka, kb, slot):
        n = self.n[slot]
        density = _f64_colmajor(density, "density")
        if density.shape != (n, n):
            raise ValueError(f"the resident tensor has n = {n}, the density is {density.shape}")
        ca, lda = self._coeff(ca, n, n_a)
        cb, ldb = self._coeff(cb, n, n_b)
        g = np.empty((n, n), dtype=np.float64, order="F")
        _check(self._lib.mqcb200_build_g_two_factor(self._h, slot, _ptr(density), _ptr(ca), lda, int(n_a),
                                                    _ptr(cb), ldb, int(n_b), c_double(ka), c_double(kb), _ptr(g)))
        return g

    def response_operator_df(self, x, c_occ, dtilde, k_scale=None, slot: int = SLOT_FULL_RANGE) -> np.ndarray:
        """``response_operator_df(b, x, c_occ, dtilde, g, k_scale)`` with ``b`` resident
        (backends/libcint/mqc_libcint_cphf.F90:499-566): ``g = J[dtilde] - kf/2 * sum_P
        [(B_P X)(B_P C)^T + h.c.]``, evaluated in that direct rank-2 form on the device (one
        half-transform of the stacked ``[X | C]``, SYR2K-form accumulation)."""
        n = self.n[slot]
        x, ldx = self._coeff(x, n, np.asarray(x).shape[1])
        n_occ = x.shape[1]
        c_occ, ldc = self._coeff(c_occ, n, n_occ)
        if c_occ.shape[1] != n_occ:
            raise ValueError("x and c_occ must both be (n_ao, n_occ)")
        dtilde = _f64_colmajor(dtilde, "dtilde")
        if dtilde.shape != (n, n):
            raise ValueError(f"the resident tensor has n = {n}, dtilde is {dtilde.shape}")
        g = np.empty((n, n), dtype=np.float64, order="F")
        _check(self._lib.mqcb200_response_operator(self._h, slot, _ptr(x), ldx, _ptr(c_occ), ldc, int(n_occ),
                                                   _ptr(dtilde), c_double(1.0 if k_scale is None else k_scale), _ptr(g)))
        return g

    def fitted_potential_general(self, dens, k_scale=None, slot: int = SLOT_FULL_RANGE) -> np.ndarray:
        """``fitted_potential_general(b, dens, g, k_scale)`` (mqc_libcint_cphf.F90:568-616):
        ``g = J[D] - kf/2 * sum_P B_P D B_P`` for a symmetric, otherwise arbitrary ``D``, on the
        device through the rank-2 kernels (``X = D``, ``C = 1``): no eigendecomposition."""
        n = self.n[slot]
        dens = _f64_colmajor(dens, "dens")
        if dens.shape != (n, n):
            raise ValueError(f"the resident tensor has n = {n}, dens is {dens.shape}")
        g = np.empty((n, n), dtype=np.float64, order="F")
        _check(self._lib.mqcb200_fitted_potential_general(self._h, slot, _ptr(dens),
                                                          c_double(1.0 if k_scale is None else k_scale), _ptr(g)))
        return g

    def run_scf(self, h, s, n_electrons: int, **kw) -> dict:
        """``run_libcint_rhf``'s loop on the GPU for any size (``mqcb200_scf``): the one-CTA fragment
        route when it fits, the general kernels + batched GEMM + one-sided Jacobi otherwise; also for a
        tensor sharded over GPUs.  Same arguments and result as :meth:`run_scf_fragment`."""
        return self.run_scf_fragment(h, s, n_electrons, _entry="mqcb200_scf", **kw)

    def run_scf_fragment(self, h, s, n_electrons: int, max_iter: int = 100, energy_tol: float = 1e-10,
                         density_tol: float = 1e-8, diis_vectors: int = 8, guess: str = "gwh", k_scale=None,
                         e_nuc: float = 0.0, slot: int = SLOT_FULL_RANGE, check_every: int = 1,
                         _entry: str = "mqcb200_scf_fragment") -> dict:
        """``run_libcint_rhf`` (mqc_libcint_rhf.f90:321-680) on the resident tensor with the SCF
        step on the GPU (fragment-sized problems, n <= 80).  Returns the fields of the reference's
        ``rhf_result_t`` (:117-141): energy, electronic, nuclear_repulsion, iterations, converged,
        n_occupied, orbitals, orbital_energies, density -- plus ``e_history``."""
        h = np.asfortranarray(_f64_colmajor(h, "h"))
        s = np.asfortranarray(_f64_colmajor(s, "s"))
        n = h.shape[0]
        if h.shape != (n, n) or s.shape != (n, n):
            raise ValueError("h and s must both be n x n")
        self._check_operand(n, slot)
        _check(self._lib.mqcb200_set_scf_check_every(self._h, int(check_every)))
        e_el, it, conv, n_mo = c_double(0.0), c_int(0), c_int(0), c_int(0)
        coeff = np.zeros((n, n), order="F")
        eps = np.zeros(n)
        density = np.empty((n, n), order="F")
        hist = np.zeros(max_iter)
        _check(getattr(self._lib, _entry)(
            self._h, slot, _ptr(h), _ptr(s), int(n_electrons), {"core": 0, "gwh": 1}[guess], int(max_iter),
            c_double(energy_tol), c_double(density_tol), int(diis_vectors),
            c_double(1.0 if k_scale is None else k_scale), byref(e_el), byref(it), byref(conv), byref(n_mo),
            _ptr(coeff), _ptr(eps), _ptr(density), _ptr(hist)))
        m = n_mo.value
        return {"energy": e_el.value + e_nuc, "electronic": e_el.value, "nuclear_repulsion": e_nuc,
                "iterations": it.value, "converged": bool(conv.value), "n_occupied": n_electrons // 2,
                "orbitals": np.asfortranarray(coeff.reshape(-1, order="F")[:n * m].reshape((n, m), order="F")),
                "orbital_energies": eps[:m].copy(), "density": density, "e_history": hist[:it.value].copy()}

    def df_gradient_densities(self, half, total_density, orbitals, n_occupied: int, orbitals_beta=None,
                              n_occupied_beta: int = 0, exx_fraction=None, with_coulomb: bool = True,
                              slot: int = SLOT_FULL_RANGE):
        """The contraction half of ``df_two_electron_gradient`` (mqc_libcint_gradient.f90:1545-1812)
        on the resident whitened tensor: returns ``(gamma(nao, nao, naux), omega(naux, naux))``.
        ``half`` is the ``metric_inverse_sqrt`` the tensor was whitened with; ``orbitals_beta``
        present means the unrestricted form (channel weight ``exx_fraction`` each instead of twice it)."""
        n = self.n[slot]
        naux = self.naux[slot][0]
        half = np.asfortranarray(_f64_colmajor(half, "half"))
        dens = np.asfortranarray(_f64_colmajor(total_density, "total_density"))
        if half.shape != (naux, naux) or dens.shape != (n, n):
            raise ValueError("half must be naux x naux and total_density nao x nao")
        ca, lda = self._coeff(orbitals, n, n_occupied)
        unrestricted = orbitals_beta is not None
        cb, ldb = self._coeff(orbitals_beta, n, n_occupied_beta) if unrestricted and n_occupied_beta > 0 else (None, n)
        gamma = np.empty((n, n, naux), dtype=np.float64, order="F")
        omega = np.empty((naux, naux), dtype=np.float64, order="F")
        _check(self._lib.mqcb200_df_gradient_densities(
            self._h, slot, _ptr(half), _ptr(dens), _ptr(ca), lda, int(n_occupied), _ptr(cb), ldb, int(n_occupied_beta),
            1 if unrestricted else 0, c_double(1.0 if exx_fraction is None else exx_fraction), 1 if with_coulomb else 0,
            _ptr(gamma), _ptr(omega)))
        return gamma, omega

    def run_scf_fragment_batch(self, h_all, s_all, n_electrons: int, max_iter: int = 100, energy_tol: float = 1e-10,
                               density_tol: float = 1e-8, diis_vectors: int = 8, guess: str = "gwh", k_scale=None,
                               slot: int = SLOT_FULL_RANGE, check_every: int = 1, want_matrices: bool = True) -> dict:
        """A batch of fragments of one kind in lock-step (``mqcb200_scf_fragment_batch``): ``h_all`` and
        ``s_all`` are ``(n_fragments, n, n)`` stacks (each matrix symmetric), the slot holds the
        fragments' tensors back to back (``set_tensor(np.hstack([b_0, b_1, ...]))``).  Returns arrays
        over the fragments: electronic, iterations, converged (1 / 0 / -1), n_mo, and -- with
        ``want_matrices`` -- density ``(n_fragments, n, n)``, orbitals, orbital_energies."""
        h_all = np.ascontiguousarray(np.asarray(h_all, dtype=np.float64))
        s_all = np.ascontiguousarray(np.asarray(s_all, dtype=np.float64))
        nf, n = h_all.shape[0], h_all.shape[1]
        if h_all.shape != (nf, n, n) or s_all.shape != (nf, n, n):
            raise ValueError("h_all and s_all must both be (n_fragments, n, n)")
        self._check_operand(n, slot)
        _check(self._lib.mqcb200_set_scf_check_every(self._h, int(check_every)))
        e = np.zeros(nf)
        it = np.zeros(nf, dtype=np.int32)
        conv = np.zeros(nf, dtype=np.int32)
        nmo = np.zeros(nf, dtype=np.int32)
        coeff = np.zeros((nf, n, n)) if want_matrices else None     # [f] holds the column-major (n, n_mo) image
        eps = np.zeros((nf, n)) if want_matrices else None
        dens = np.zeros((nf, n, n)) if want_matrices else None
        _check(self._lib.mqcb200_scf_fragment_batch(
            self._h, slot, nf, _ptr(h_all), _ptr(s_all), int(n_electrons), {"core": 0, "gwh": 1}[guess], int(max_iter),
            c_double(energy_tol), c_double(density_tol), int(diis_vectors), c_double(1.0 if k_scale is None else k_scale),
            _ptr(e), _ptr(it), _ptr(conv), _ptr(nmo), _ptr(coeff), _ptr(eps), _ptr(dens)))
        out = {"electronic": e, "iterations": it, "converged": conv, "n_mo": nmo}
        if want_matrices:
            out["density"] = dens                                    # symmetric: the row-major view equals the column-major image
            out["orbitals"] = [np.asfortranarray(coeff[f].reshape(-1)[:n * nmo[f]].reshape((n, nmo[f]), order="F")) for f in range(nf)]
            out["orbital_energies"] = [eps[f, :nmo[f]].copy() for f in range(nf)]
        return out

    def last_energy(self) -> float:
        """``electronic_energy(h, fock, density)`` of the last ``build_fock_df`` (rhf.f90:1691-1697)."""
        e = c_double(0.0)
        _check(self._lib.mqcb200_last_energy(self._h, byref(e)))
        return e.value

    def assemble_fock(self, h, density, coeff, n_occ: int, k_scale=None, rs_k_lr=None):
        """DF branch of ``assemble_fock``: returns ``(fock, e_elec)`` before any V_xc.

        With a tensor on the attenuated slot and ``rs_k_lr`` given, the second pass of a
        range-separated functional runs with ``h = 0``, ``k_scale = rs_k_lr``, ``j_scale = 0``
        and is added to the Fock matrix (rhf.f90:1094-1105).
        """
        h = _f64_colmajor(h, "h")
        density = _f64_colmajor(density, "density")
        fock = self.build_fock_df(h, density, coeff, n_occ, k_scale=k_scale)
        if rs_k_lr is not None:
            if SLOT_ATTENUATED not in self.n:
                raise B200Error(_lib.MQCB200_FAIL, "range-separated pass requested but no attenuated tensor is set")
            k_lr = self.build_fock_df(np.zeros_like(h), density, coeff, n_occ, k_scale=rs_k_lr,
                                      j_scale=0.0, slot=SLOT_ATTENUATED)
            fock = fock + k_lr
            e_elec = 0.5 * float(np.sum(density * (h + fock)))
        else:
            e_elec = self.last_energy()
        return fock, e_elec

    def atomic_guess_fock(self, h, density) -> np.ndarray:
        """One build from a guess density through its pseudo-orbitals (rhf.f90:1402-1405)."""
        pseudo, n_modes = density_pseudo_orbitals(density)
        return self.build_fock_df(h, density, pseudo, n_modes)

    # -- device-resident variant (torch CUDA tensors, column-major storage) --------------
    def build_fock_device(self, d_h, d_density, d_coeff, n_occ: int, d_fock, k_scale=1.0,
                          j_scale=1.0, slot: int = SLOT_FULL_RANGE, sync: bool = True) -> None:
        """All arguments are CUDA tensors holding column-major matrices (data_ptr is used)."""
        for t in (d_h, d_density, d_coeff, d_fock):
            if not t.is_cuda or str(t.dtype) != "torch.float64" or not t.is_contiguous():
                raise TypeError("device operands must be contiguous float64 CUDA tensors")
        _check(self._lib.mqcb200_build_fock_device(
            self._h, slot, c_void_p(d_h.data_ptr()), c_void_p(d_density.data_ptr()),
            c_void_p(d_coeff.data_ptr()), int(n_occ), c_double(k_scale), c_double(j_scale),
            c_void_p(d_fock.data_ptr()), 1 if sync else 0))

    def build_fock_uhf_device(self, d_h, d_density_total, d_coeff_a, n_alpha: int, d_coeff_b, n_beta: int,
                              d_fock_a, d_fock_b, k_scale=1.0, slot: int = SLOT_FULL_RANGE,
                              sync: bool = True) -> None:
        tensors = [d_h, d_density_total, d_coeff_a, d_fock_a, d_fock_b] + ([d_coeff_b] if n_beta > 0 else [])
        for t in tensors:
            if not t.is_cuda or str(t.dtype) != "torch.float64" or not t.is_contiguous():
                raise TypeError("device operands must be contiguous float64 CUDA tensors")
        _check(self._lib.mqcb200_build_fock_uhf_device(
            self._h, slot, c_void_p(d_h.data_ptr()), c_void_p(d_density_total.data_ptr()),
            c_void_p(d_coeff_a.data_ptr()), int(n_alpha),
            c_void_p(d_coeff_b.data_ptr() if n_beta > 0 else None), int(n_beta), c_double(k_scale),
            c_void_p(d_fock_a.data_ptr()), c_void_p(d_fock_b.data_ptr()), 1 if sync else 0))

    # -- multi-GPU --------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _check(_lib.load().mqcb200_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes) -> None:
        if len(unique_id) != 128:
            raise ValueError("the communicator id is 128 bytes")
        _check(self._lib.mqcb200_comm_init(self._h, n_ranks, rank, unique_id))

    def comm_destroy(self) -> None:
        _check(self._lib.mqcb200_comm_destroy(self._h))

    # -- instrumentation --------------------------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        _check(self._lib.mqcb200_set_profiling(self._h, 1 if on else 0))

    def last_timings(self) -> dict:
        ms = (c_double * _lib.NUM_TIMERS)()
        _check(self._lib.mqcb200_last_timings(self._h, ms))
        return dict(zip(_lib.TIMER_NAMES, list(ms)))

    def last_whiten(self):
        """``(ms, flops)`` of the last device whitening GEMM."""
        ms, fl = c_double(0.0), c_double(0.0)
        _check(self._lib.mqcb200_last_whiten(self._h, byref(ms), byref(fl)))
        return ms.value, fl.value

    def last_gamma_fused(self) -> bool:
        """True when the last build took its Coulomb vector from the half-transform."""
        f = c_int(0)
        _check(self._lib.mqcb200_last_gamma_fused(self._h, byref(f)))
        return bool(f.value)

    def last_launches(self) -> int:
        n = c_int(0)
        _check(self._lib.mqcb200_last_launches(self._h, byref(n)))
        return n.value

    # -- helpers ------------------------------------------------------------------------------
    @staticmethod
    def _coeff(coeff, n: int, n_occ: int):
        if n_occ == 0:
            return None, n
        coeff = _f64_colmajor(coeff, "coeff", allow_ld=True)
        if coeff.shape[0] != n or coeff.shape[1] < n_occ:
            raise ValueError("coeff must be (n, >= n_occ)")
        ldc = coeff.strides[1] // coeff.itemsize if coeff.shape[1] > 1 else max(n, coeff.shape[0])
        return coeff, int(ldc)


NULL_THRESHOLD = 1.0e-10     # mqc_libcint_integrals.F90:1002


def metric_inverse_sqrt(metric, engine: "B200FockEngine | None" = None):
    """J^(-1/2) = U s^(-1/2) U^T over the modes above 1e-10 (mqc_libcint_integrals.F90:992-1038),
    formed ON THE DEVICE (``B200FockEngine.metric_inverse_sqrt``); without an engine a temporary
    one is created on GPU 0.  No host eigensolver is involved."""
    if engine is not None:
        return engine.metric_inverse_sqrt(metric)
    with B200FockEngine(0) as tmp:
        return tmp.metric_inverse_sqrt(metric)


def density_pseudo_orbitals(density):
    """Columns c_i with D = 2 sum_i c_i c_i^T  (mqc_libcint_rhf.f90:1413-1462).

    Host-side, as in the reference: one dsyev on the guess density, modes with
    occupation above 1e-12 kept, ``c_i = v_i*sqrt(w_i/2)``.  Raises the reference's
    message when nothing is occupied.
    """
    density = np.asarray(density, dtype=np.float64)
    values, vectors = np.linalg.eigh(density, UPLO="U")
    keep = values > OCCUPATION_FLOOR
    n_modes = int(np.count_nonzero(keep))
    if n_modes == 0:
        raise B200Error(_lib.MQCB200_FAIL, "guess: the guess density carries no occupation")
    coeff = np.asfortranarray(vectors[:, keep] * np.sqrt(0.5 * values[keep])[None, :])
    return coeff, n_modes


class WorkQueue:
    """FIFO of int64 fragment ids with the semantics of the reference's ``queue_t``
    (``queue_init_from_list`` / ``queue_pop`` / ``queue_is_empty`` / ``queue_destroy``)."""

    def __init__(self, ids):
        self._lib = _lib.load()
        arr = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
        self._q = c_void_p(None)
        _check(self._lib.mqcb200_queue_create(arr.ctypes.data_as(ctypes.POINTER(c_int64)),
                                              c_int64(arr.size), byref(self._q)))

    def pop(self):
        """Returns ``(item_idx, has_item)``; ``(-1, False)`` once drained."""
        idx = c_int64(0)
        has = c_int(0)
        _check(self._lib.mqcb200_queue_pop(self._q, byref(idx), byref(has)))
        return idx.value, bool(has.value)

    def is_empty(self) -> bool:
        out = c_int(0)
        _check(self._lib.mqcb200_queue_is_empty(self._q, byref(out)))
        return bool(out.value)

    def destroy(self) -> None:
        if self._q is not None and self._q.value:
            rc = self._lib.mqcb200_queue_destroy(self._q)
            self._q = c_void_p(None)
            _check(rc)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
