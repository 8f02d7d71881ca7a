"""CPU oracle for the density-fitted J/K Fock build -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy float64 *restatement* of the reference's CPU hot path.  It
is the checker the CUDA engine is compared against; it is never the thing that
is shipped or measured as the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product package ``metalquicha_b200`` never
does (``tests/test_no_oracle_in_product.py`` enforces that).

PARITY PIN.  The reference (JorgeG94/metalquicha) is Fortran; this image has no Fortran
compiler, no libcint/libfint, no pic-blas and no basis-set bundle, so the reference cannot
be compiled or run here, and none of its tests asserts a J, K, F or c_P element (SURVEY.md
section 8c).  What it DOES hold are total energies, and this oracle is pinned to them
(tests/test_reference_golden_energies.py, with the integrals restated in oracle/gto_integrals.py):

* the reference's own DENSITY-FITTED validation energies -- water/6-31G* fitted with 6-31G*
  (-76.188111755038) and CH4/6-31G** fitted with 6-31G** (-40.381603512964),
  validation/validation_tests_cpu.json:899-910, tolerance 1e-9 Eh -- are reproduced to 2e-12 Eh
  through ``metric_inverse_sqrt`` -> ``whiten`` -> ``build_fock_df`` inside the restated SCF loop
  (oracle/scf_oracle.py), and so is the fitting error against the exact-integral energy
  (-0.178 Eh: nothing cancels by accident).  6-31G* is published data; the H and O numbers are
  cross-checked against the copy the reference tree holds in GAMESS form;
* H2 and H2O in STO-3G (validation/check_rhf.f90:54-153, basis written out inline there) to 3e-11 Eh
  through ``build_fock_eri`` AND through ``build_fock_df`` on a tensor that fits the four-index
  integrals exactly.

That pins the metric threshold and inverse square root, the whitening and its (mu + n*nu, P)
flattening, the Coulomb contraction, the factor 2 of K, the 1/2 of F and the energy expression to
reference-held numbers.  The CUDA engine runs the same cases (tests/test_gpu_reference_df_energies.py,
tests/test_gpu_golden_scf.py).  The energies that need the other named basis sets (cc-pVDZ with
JKFIT/RIFIT, def2) or the XC quadrature stay unasserted (tests/golden/reference_energies.json).
Beyond the pin the restatement follows the reference *loop for loop* with file:line citations, is
cross-checked against two independent C restatements (``oracle/df_fock_ref.c``,
``oracle/df_fock_blas.c``), and satisfies the reference's own algebraic identities (tests/test_oracle.py).

All arrays are float64.  Matrices use the reference's layout: ``b`` is
``(n*n, naux)`` with slab ``b[:, p]`` holding the symmetric ``n x n`` matrix of
auxiliary function ``p`` flattened column-major, i.e. element ``(mu, nu)`` at
flat index ``mu + n*nu``  (reference backends/libcint/mqc_libcint_integrals.F90:1425-1437).
Because every slab is symmetric the C/Fortran order of the ``reshape`` does not
change any value.
"""
from __future__ import annotations

import numpy as np

NULL_THRESHOLD = 1.0e-10       # mqc_libcint_integrals.F90:1002
OCCUPATION_FLOOR = 1.0e-12     # mqc_libcint_rhf.f90:1438


def _slab(b: np.ndarray, p: int, n: int) -> np.ndarray:
    """``reshape(b(:, p), [n, n])`` -- Fortran (column-major) reshape of slab p."""
    return b[:, p].reshape((n, n), order="F")


def coulomb_vector(b: np.ndarray, density: np.ndarray) -> np.ndarray:
    """c(p) = sum(reshape(b(:,p),[n,n]) * density)      mqc_libcint_rhf.f90:1620-1622"""
    n = density.shape[0]
    naux = b.shape[1]
    c = np.empty(naux)
    for p in range(naux):
        c[p] = np.sum(_slab(b, p, n) * density)
    return c


def jk_df(b: np.ndarray, density: np.ndarray, coeff: np.ndarray, n_occ: int):
    """The J and K of ``build_fock_df`` before scaling (mqc_libcint_rhf.f90:1614-1639).

    J = sum_p c(p) B_p;   K = 2 sum_p (B_p C_occ)(B_p C_occ)^T   (RHF convention,
    the factor 2 is the ``alpha=2.0_dp`` of the second ``pic_gemm`` at :1637).
    """
    n = density.shape[0]
    naux = b.shape[1]
    c_occ = np.ascontiguousarray(coeff[:, :n_occ])                 # :1618
    c = coulomb_vector(b, density)                                 # :1620-1622
    j = np.zeros((n, n))                                           # :1624
    for p in range(naux):                                          # :1625-1627
        j = j + c[p] * _slab(b, p, n)
    k = np.zeros((n, n))                                           # :1629
    for p in range(naux):                                          # :1630-1639
        b_p = _slab(b, p, n)
        w = b_p @ c_occ                                            # pic_gemm(b_p, c_occ, w)   :1636
        k = 2.0 * (w @ w.T) + k                                    # alpha=2, beta=1, transb=T :1637
    return j, k, c


def build_fock_df(h, b, density, coeff, n_occ, k_scale=None, j_scale=None):
    """F = H + jf*J - kf*K        mqc_libcint_rhf.f90:1576-1646

    ``kf = 0.5`` (times ``k_scale`` when present, :1641-1642); ``jf = 1`` (or
    ``j_scale`` when present, :1643-1644); ``fock = h + jf*j - kf*k`` (:1645).
    """
    j, k, _ = jk_df(b, density, coeff, n_occ)
    kf = 0.5
    if k_scale is not None:
        kf = 0.5 * k_scale
    jf = 1.0
    if j_scale is not None:
        jf = j_scale
    return h + jf * j - kf * k


def electronic_energy(h, fock, density) -> float:
    """E = 1/2 sum D (H + F)                 mqc_libcint_rhf.f90:1691-1697"""
    return 0.5 * float(np.sum(density * (h + fock)))


def uhf_electronic_energy(h, fock_a, fock_b, d_alpha, d_beta) -> float:
    """E = 1/2 [sum (Da+Db) H + sum Da Fa + sum Db Fb]     mqc_libcint_rhf.f90:1683-1689"""
    return 0.5 * float(np.sum((d_alpha + d_beta) * h) + np.sum(d_alpha * fock_a)
                       + np.sum(d_beta * fock_b))


def build_density_closed_shell(coeff, n_occ):
    """D = 2 C_occ C_occ^T                   src/scf/mqc_scf_common.f90:84-96"""
    n = coeff.shape[0]
    if n_occ <= 0:
        return np.zeros((n, n))
    c = coeff[:, :n_occ]
    return 2.0 * (c @ c.T)


def build_density_spin(coeff, n_occ):
    """D_sigma = C_occ C_occ^T               src/scf/mqc_scf_common.f90:98-109"""
    n = coeff.shape[0]
    if n_occ <= 0:
        return np.zeros((n, n))
    c = coeff[:, :n_occ]
    return c @ c.T


def jk_df_uhf(b, d_total, c_alpha, n_alpha, c_beta, n_beta):
    """Two-spin J/K (row a8 of SURVEY section 8).

    The reference's CPU path refuses DF-UHF (mqc_libcint_bridge.f90:605-612);
    the conventions come from its cuEST SCF: J from the TOTAL density, shared by
    both channels; K_sigma = sum_p (B_p C_sigma)(B_p C_sigma)^T with NO factor 2
    (backends/cuest/backend/mqc_cuest_scf.f90:48-57, :826-838).  An empty beta
    channel is skipped, not zeroed (mqc_cuest_integrals.f90:1694-1701): K_beta is
    returned as ``None`` in that case.
    """
    n = d_total.shape[0]
    naux = b.shape[1]
    c = coulomb_vector(b, d_total)
    j = np.zeros((n, n))
    for p in range(naux):
        j = j + c[p] * _slab(b, p, n)

    def _k(cmat, n_sigma):
        if n_sigma <= 0:
            return None
        cs = np.ascontiguousarray(cmat[:, :n_sigma])
        k = np.zeros((n, n))
        for p in range(naux):
            w = _slab(b, p, n) @ cs
            k = (w @ w.T) + k
        return k

    return j, _k(c_alpha, n_alpha), _k(c_beta, n_beta)


def build_fock_df_uhf(h, b, d_alpha, d_beta, c_alpha, n_alpha, c_beta, n_beta, k_scale=None):
    """F_sigma = H + J[Da+Db] - k_scale*K[D_sigma]

    Mirrors the shape of ``build_fock_uhf`` (mqc_libcint_rhf.f90:1648-1681, ``kf``
    defaults to one) with the fitted K of :func:`jk_df_uhf`.
    """
    kf = 1.0 if k_scale is None else k_scale
    j, ka, kb = jk_df_uhf(b, d_alpha + d_beta, c_alpha, n_alpha, c_beta, n_beta)
    fa = h + j - (kf * ka if ka is not None else 0.0)
    fb = h + j - (kf * kb if kb is not None else 0.0)
    return fa, fb


def fitted_exchange_general(b, density):
    """K = sum_P B_P D B_P  (general-density form, mqc_libcint_cphf.F90:606-613).

    Used only for identity (iii) of SURVEY 8c: equals jk_df's K for D = 2CC^T.
    """
    n = density.shape[0]
    k = np.zeros((n, n))
    for p in range(b.shape[1]):
        b_p = _slab(b, p, n)
        k = k + b_p @ density @ b_p
    return k


def build_fock_eri(h, eri, density, k_scale=None):
    """The reference's in-core build from a four-index tensor, ``build_fock``
    (mqc_libcint_rhf.f90:1491-1574):  J(a,b) = sum_cd D(c,d) (ab|cd)  (:1553-1559),
    K(a,c) = sum_bd (ab|cd) D(b,d)  (the gemv at :1561),  F = H + J - kf K with
    kf = 0.5 (* k_scale)  (:1533-1534, :1572).  Used only to tie the fitted oracle to the
    reference's exact-integral routine on the fitted integrals (ab|cd) = sum_P B_P(ab) B_P(cd)."""
    n = h.shape[0]
    kf = 0.5 if k_scale is None else 0.5 * k_scale
    j_mat = np.zeros((n, n))
    k_mat = np.zeros((n, n))
    for c in range(n):
        for d in range(n):
            j_mat = j_mat + density[c, d] * eri[:, :, c, d]
            k_mat[:, c] = eri[:, :, c, d] @ density[:, d] + k_mat[:, c]
    return h + j_mat - kf * k_mat


def response_operator_df(b, x, c_occ, dtilde, k_scale=None):
    """g = J[dtilde] - kf/2 * sum_P [(B_P X)(B_P C)^T + (B_P C)(B_P X)^T]
    backends/libcint/mqc_libcint_cphf.F90:499-566, loop for loop (:551-562)."""
    kf = 1.0 if k_scale is None else k_scale
    n = c_occ.shape[0]
    coul = np.zeros((n, n))
    exch = np.zeros((n, n))
    for p in range(b.shape[1]):
        b_p = _slab(b, p, n)
        c_p = np.sum(b_p * dtilde)
        coul = coul + c_p * b_p
        bx = b_p @ x
        bc = b_p @ c_occ
        exch = bx @ bc.T + exch
        exch = bc @ bx.T + exch
    return coul - 0.5 * kf * exch


def fitted_potential_general(b, dens, k_scale=None):
    """g = J[D] - kf/2 * sum_P B_P D B_P     mqc_libcint_cphf.F90:568-616 (:604-613)."""
    kf = 1.0 if k_scale is None else k_scale
    n = dens.shape[0]
    coul = np.zeros((n, n))
    exch = np.zeros((n, n))
    for p in range(b.shape[1]):
        b_p = _slab(b, p, n)
        c_p = np.sum(b_p * dens)
        coul = coul + c_p * b_p
        bd = b_p @ dens
        exch = bd @ b_p + exch
    return coul - 0.5 * kf * exch


def metric_inverse_sqrt(metric: np.ndarray) -> np.ndarray:
    """J^(-1/2) = U s^(-1/2) U^T over modes with eigenvalue > 1e-10.

    mqc_libcint_integrals.F90:992-1038: ``pic_syev`` (LAPACK dsyev, 'V','U');
    ``scaled(:,i) = vectors(:,i)/sqrt(values(i))`` if ``values(i) > NULL_THRESHOLD``
    else 0 (:1027-1033); ``half = scaled * vectors^T`` (:1036).  Raises
    ``ValueError`` with the reference's message when no mode survives (:1015-1019).
    """
    values, vectors = np.linalg.eigh(metric, UPLO="U")
    kept = int(np.count_nonzero(values > NULL_THRESHOLD))
    if kept == 0:
        raise ValueError("density fitting: the auxiliary metric is singular")
    scaled = np.zeros_like(vectors)
    for i in range(values.shape[0]):
        if values[i] > NULL_THRESHOLD:
            scaled[:, i] = vectors[:, i] / np.sqrt(values[i])
    return scaled @ vectors.T


def whiten(three: np.ndarray, metric: np.ndarray) -> np.ndarray:
    """b = three . metric^(-1/2)           mqc_libcint_integrals.F90:981-986"""
    return three @ metric_inverse_sqrt(metric)


def density_pseudo_orbitals(density: np.ndarray):
    """Columns c_i with D = 2 sum_i c_i c_i^T      mqc_libcint_rhf.f90:1413-1462

    dsyev('V','U'); modes with ``w_i > OCCUPATION_FLOOR`` kept in ascending
    eigenvalue order; ``c_i = v_i * sqrt(0.5*w_i)`` (:1460).
    """
    values, vectors = np.linalg.eigh(density, UPLO="U")
    keep = values > OCCUPATION_FLOOR
    n_modes = int(np.count_nonzero(keep))
    if n_modes == 0:
        raise ValueError("guess: the guess density carries no occupation")
    coeff = vectors[:, keep] * np.sqrt(0.5 * values[keep])[None, :]
    return np.ascontiguousarray(coeff), n_modes


def assemble_fock_df(h, bmat, density, coeff, n_occ, k_scale=None,
                     bmat_lr=None, rs_k_lr=None):
    """The DF branch of ``assemble_fock`` (mqc_libcint_rhf.f90:1089-1106, :1207).

    Returns ``(fock, e_elec)`` with ``e_elec = 1/2 sum D (H + F)`` taken before
    any V_xc is added.  When ``bmat_lr`` is given the range-separated second pass
    runs with ``h = 0``, ``k_scale = rs_k_lr`` and ``j_scale = 0`` and its result
    is added to the Fock matrix (:1094-1105).
    """
    fock = build_fock_df(h, bmat, density, coeff, n_occ, k_scale=k_scale)
    if bmat_lr is not None:
        k_lr = build_fock_df(np.zeros_like(h), bmat_lr, density, coeff, n_occ,
                             k_scale=rs_k_lr, j_scale=0.0)
        fock = fock + k_lr
    return fock, electronic_energy(h, fock, density)


# --------------------------------------------------------------------------
# Fast forms used only to keep the *tests* quick at larger shapes.  They are
# algebraically the loops above with the per-p Python loop replaced by BLAS
# calls on the whole tensor, and tests/test_oracle.py checks them against the
# loop-for-loop forms at small sizes.
# --------------------------------------------------------------------------
def jk_df_fast(b, density, coeff, n_occ, rhf_factor=2.0):
    n = density.shape[0]
    naux = b.shape[1]
    c = b.T @ density.reshape(n * n, order="F")
    j = (b @ c).reshape((n, n), order="F")
    c_occ = np.ascontiguousarray(coeff[:, :n_occ])
    k = np.zeros((n, n))
    step = max(1, min(naux, (1 << 27) // max(1, n * max(n_occ, 1))))
    for p0 in range(0, naux, step):
        p1 = min(naux, p0 + step)
        bs = b[:, p0:p1].T.reshape(p1 - p0, n, n)          # symmetric slabs
        w = bs @ c_occ                                      # (q, n, o)
        wf = w.transpose(1, 0, 2).reshape(n, -1)            # (n, q*o)
        k += rhf_factor * (wf @ wf.T)
    return j, k, c
