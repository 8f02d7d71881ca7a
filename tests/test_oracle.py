"""The oracle against itself: the NumPy restatement vs the plain-C restatement, the
reference's own algebraic identities (SURVEY 8c i-iv), and the committed fixtures."""
import ctypes
import os

import numpy as np
import pytest

from metalquicha_b200 import synth
from oracle import df_fock_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C_LIB = os.path.join(ROOT, "oracle", "_build", "libdf_fock_ref.so")


def _c_oracle():
    if not os.path.exists(C_LIB):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(C_LIB)
    dp = ctypes.c_void_p
    lib.df_ref_build_fock.argtypes = [ctypes.c_int] * 3 + [dp] * 4 + [ctypes.c_int, ctypes.c_double,
                                                                   ctypes.c_double, dp]
    lib.df_ref_jk.argtypes = [ctypes.c_int] * 3 + [dp] * 3 + [ctypes.c_int, ctypes.c_double, dp, dp, dp]
    return lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("n,n_occ,naux", [(24, 5, 116), (19, 5, 19), (7, 2, 5), (40, 13, 30)])
def test_numpy_and_c_restatements_agree(n, n_occ, naux):
    b, h, density, coeff = synth.synth_problem(1, n, n_occ, naux)
    lib = _c_oracle()
    fock_c = np.empty((n, n), order="F")
    assert lib.df_ref_build_fock(n, naux, n_occ, _p(h), _p(b), _p(density), _p(coeff), n, 0.7, 1.0,
                                 _p(fock_c)) == 0
    fock_np = oracle.build_fock_df(h, b, density, coeff, n_occ, k_scale=0.7)
    assert np.max(np.abs(fock_c - fock_np)) <= 1e-12
    j = np.empty((n, n), order="F"); k = np.empty((n, n), order="F"); c = np.empty(naux)
    assert lib.df_ref_jk(n, naux, n_occ, _p(b), _p(density), _p(coeff), n, 2.0, _p(j), _p(k), _p(c)) == 0
    j_np, k_np, c_np = oracle.jk_df(b, density, coeff, n_occ)
    assert np.max(np.abs(j - j_np)) <= 1e-12 and np.max(np.abs(k - k_np)) <= 1e-12
    assert np.max(np.abs(c - c_np)) <= 1e-12


def test_default_scales_are_one_and_half():
    """kf = 0.5 and jf = 1 when the optional arguments are absent (rhf.f90:1641-1644)."""
    b, h, density, coeff = synth.synth_problem(2, 12, 3, 9)
    j, k, _ = oracle.jk_df(b, density, coeff, 3)
    assert np.allclose(oracle.build_fock_df(h, b, density, coeff, 3), h + j - 0.5 * k, rtol=0, atol=1e-14)
    assert np.allclose(oracle.build_fock_df(h, b, density, coeff, 3, k_scale=0.2, j_scale=0.0),
                       h - 0.1 * k, rtol=0, atol=1e-14)


def test_identity_i_rhf_vs_two_spin():
    """K_rhf = 2 K_sigma and equal J when C_alpha = C_beta (CUEST.md:392)."""
    n, o, q = 20, 4, 15
    b, _, density, coeff = synth.synth_problem(3, n, o, q)
    j, k, _ = oracle.jk_df(b, density, coeff, o)
    ds = oracle.build_density_spin(coeff, o)
    j2, ka, kb = oracle.jk_df_uhf(b, ds + ds, coeff, o, coeff, o)
    assert np.max(np.abs(j - j2)) <= 1e-13
    assert np.max(np.abs(k - 2.0 * ka)) <= 1e-13 and np.max(np.abs(ka - kb)) == 0.0
    assert oracle.jk_df_uhf(b, ds, coeff, o, coeff, 0)[2] is None      # empty channel skipped


def test_identity_ii_pseudo_orbitals_rebuild_density_and_k():
    """D = 2 sum c c^T holds exactly for the pseudo-orbitals (rhf.f90:1421-1423;
    reference test pseudo_orbitals_rebuild_their_density)."""
    n, o, q = 18, 5, 12
    b, _, density, coeff = synth.synth_problem(4, n, o, q)
    pseudo, n_modes = oracle.density_pseudo_orbitals(density)
    assert n_modes == o
    assert np.max(np.abs(2.0 * pseudo @ pseudo.T - density)) <= 1e-13
    _, k, _ = oracle.jk_df(b, density, coeff, o)
    _, k2, _ = oracle.jk_df(b, density, pseudo, n_modes)
    assert np.max(np.abs(k - k2)) <= 1e-12
    with pytest.raises(ValueError, match="carries no occupation"):
        oracle.density_pseudo_orbitals(np.zeros((4, 4)))


def test_identity_iii_general_density_form():
    """K from build_fock_df equals sum_P B_P D B_P for D = 2CC^T (cphf.F90:606-613)."""
    n, o, q = 16, 6, 10
    b, _, density, coeff = synth.synth_problem(5, n, o, q)
    _, k, _ = oracle.jk_df(b, density, coeff, o)
    assert np.max(np.abs(k - oracle.fitted_exchange_general(b, density))) <= 1e-12


def test_identity_iv_whitening_reproduces_metric_inverse():
    """sum_P B_P (x) B_P = three . metric^-1 . three^T when no mode is dropped."""
    n, q = 6, 9
    three, metric = synth.synth_physical_like_tensor(6, n, q, n_null=0)
    b = oracle.whiten(three, metric)
    assert np.max(np.abs(b @ b.T - three @ np.linalg.inv(metric) @ three.T)) <= 1e-9


def test_identity_v_fitted_build_equals_exact_build_on_fitted_integrals():
    """build_fock_df(h, B, D, C) == build_fock(h, eri, D) for (ab|cd) = sum_P B_P(ab) B_P(cd):
    the reference's two Fock routines (rhf.f90:1576-1646 and :1491-1574) must agree on the
    integrals the fit defines -- the statement its docstring derives K from (:1586-1593)."""
    n, o, q = 9, 3, 14
    b, h, density, coeff = synth.synth_problem(12, n, o, q)
    slabs = b.T.reshape(q, n, n)                                  # symmetric slabs
    eri = np.einsum("pab,pcd->abcd", slabs, slabs)
    for ks in (None, 0.2):
        f_df = oracle.build_fock_df(h, b, density, coeff, o, k_scale=ks)
        f_eri = oracle.build_fock_eri(h, eri, density, k_scale=ks)
        assert np.max(np.abs(f_df - f_eri)) <= 1e-12


def test_identity_vi_fitted_response_operators_equal_the_exact_one_on_fitted_integrals():
    """validation/check_df_cphf.f90 of the reference: "form the fitted integrals in four-index form,
    (uv|ls)_RI = sum_P B(uv,P) B(ls,P), hand those to the ordinary exact-ERI [operator], and require the two
    responses to agree to rounding" (:11-16, algebra tolerance 1e-9 at :158).  Here at the operator level, for
    ``response_operator_df`` (the factorised form, cphf.F90:499-566) and ``fitted_potential_general`` (:568-616),
    on the REAL water / 6-31G* fitted with 6-31G* tensor (the case whose energy pins the oracle), with an
    asymmetric trial rotation so that a transposed factor cannot hide behind symmetry (:43)."""
    from oracle import gto_integrals as gto, scf_oracle as scf
    s, h, three, metric, e_nuc, n_electrons, _ = gto.df_case_integrals("h2o_631gs")
    b = oracle.whiten(three, metric)
    n, naux = h.shape[0], metric.shape[0]
    slabs = b.T.reshape(naux, n, n)
    eri_ri = np.einsum("pab,pcd->abcd", slabs, slabs)
    x_orth = scf.build_orthogonalizer(s)
    coeff, _ = scf.diagonalize(scf.guess_fock_gwh(s, h), x_orth)
    n_occ = n_electrons // 2
    c_occ, c_vir = coeff[:, :n_occ], coeff[:, n_occ:]
    u = np.array([[1.0 / (a + 2 * i + 3) for i in range(n_occ)] for a in range(n - n_occ)])    # the rhs pattern of :111
    x = c_vir @ u
    dtilde = x @ c_occ.T + c_occ @ x.T
    for ks in (None, 0.25, 0.0):
        exact = oracle.build_fock_eri(np.zeros((n, n)), eri_ri, dtilde, k_scale=ks)            # J[Dt] - kf/2 K[Dt]
        g = oracle.response_operator_df(b, x, c_occ, dtilde, k_scale=ks)
        assert np.max(np.abs(g - exact)) <= 1e-11 * max(1.0, float(np.max(np.abs(exact))))
        g2 = oracle.fitted_potential_general(b, dtilde, k_scale=ks)
        assert np.max(np.abs(g2 - exact)) <= 1e-11 * max(1.0, float(np.max(np.abs(exact))))
    # an unstructured symmetric, indefinite, traceless density (check_fitted_reference_gradient.f90:27-33)
    rng = np.random.default_rng(5)
    a = rng.standard_normal((n, n))
    d = a + a.T
    d -= np.trace(d) / n * np.eye(n)
    exact = oracle.build_fock_eri(np.zeros((n, n)), eri_ri, d, k_scale=0.5)
    assert np.max(np.abs(oracle.fitted_potential_general(b, d, k_scale=0.5) - exact)) <= 1e-11 * float(np.max(np.abs(exact)))


def test_metric_inverse_sqrt_drops_null_modes():
    """Eigenvalues <= 1e-10 are zeroed, not errored (integrals.F90:1002,1027-1033)."""
    _, metric = synth.synth_physical_like_tensor(7, 4, 12, n_null=3)
    half = oracle.metric_inverse_sqrt(metric)
    assert np.allclose(half, half.T, atol=1e-10)
    w = np.linalg.eigvalsh(half @ metric @ half)
    assert np.sum(np.abs(w) < 1e-6) == 3 and np.allclose(np.sort(w)[3:], 1.0, atol=1e-8)
    with pytest.raises(ValueError, match="singular"):
        oracle.metric_inverse_sqrt(np.zeros((3, 3)))


def test_fast_forms_match_loop_forms():
    n, o, q = 30, 7, 25
    b, _, density, coeff = synth.synth_problem(8, n, o, q)
    j, k, c = oracle.jk_df(b, density, coeff, o)
    jf, kf, cf = oracle.jk_df_fast(b, density, coeff, o)
    assert np.max(np.abs(j - jf)) <= 1e-12 and np.max(np.abs(k - kf)) <= 1e-12 and np.max(np.abs(c - cf)) <= 1e-12


def test_energy_expressions():
    n, o, q = 10, 3, 8
    b, h, density, coeff = synth.synth_problem(9, n, o, q)
    f = oracle.build_fock_df(h, b, density, coeff, o)
    assert abs(oracle.electronic_energy(h, f, density) - 0.5 * np.sum(density * (h + f))) <= 1e-13
    ds = oracle.build_density_spin(coeff, o)
    fa, fb = oracle.build_fock_df_uhf(h, b, ds, ds, coeff, o, coeff, o)
    # closed-shell limit of the two-spin energy
    assert abs(oracle.uhf_electronic_energy(h, fa, fb, ds, ds) - oracle.electronic_energy(h, f, density)) <= 1e-11
    assert np.max(np.abs(fa - f)) <= 1e-12


def test_asymmetric_density_only_its_symmetric_part_matters():
    n, o, q = 12, 3, 7
    b, _, density, coeff = synth.synth_problem(10, n, o, q)
    rng = np.random.default_rng(0)
    d = density + rng.standard_normal((n, n))
    j1, _, _ = oracle.jk_df(b, d, coeff, o)
    j2, _, _ = oracle.jk_df(b, 0.5 * (d + d.T), coeff, o)
    assert np.max(np.abs(j1 - j2)) <= 1e-12


def test_c_blas_restatement_agrees_with_the_numpy_restatement():
    """oracle/df_fock_blas.c (bench.py's CPU baseline: the reference's loops on a vendor dgemm)
    vs the loop-for-loop NumPy oracle, closed shell and two-spin, at 1 and at all BLAS threads."""
    from oracle import df_fock_blas as blas_port
    n, n_occ, naux = 37, 9, 23
    b, h, density, coeff = synth.synth_problem(61, n, n_occ, naux)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ, k_scale=0.2, j_scale=0.7)
    for threads in (1, blas_port.host_threads()):
        with blas_port.blas_threads(threads):
            got = blas_port.build_fock_df(h, b, density, coeff, n_occ, k_scale=0.2, j_scale=0.7)
        assert np.max(np.abs(got - ref)) <= 1e-12
    padded = np.zeros((n + 4, n), order="F"); padded[:n, :n_occ] = coeff       # extra rows/columns are ignored
    j, k, c = blas_port.jk_df(b, density, np.asfortranarray(padded[:n, :]), n_occ)
    j_ref, k_ref, c_ref = oracle.jk_df(b, density, coeff, n_occ)
    assert np.max(np.abs(j - j_ref)) <= 1e-12 and np.max(np.abs(k - k_ref)) <= 1e-12 and np.max(np.abs(c - c_ref)) <= 1e-12
    cb = synth.synth_orbitals(62, n, n_occ - 2)
    da, db = oracle.build_density_spin(coeff, n_occ), oracle.build_density_spin(cb, n_occ - 2)
    fa, fb = blas_port.build_fock_df_uhf(h, b, da, db, coeff, n_occ, cb, n_occ - 2, k_scale=0.5)
    fa_ref, fb_ref = oracle.build_fock_df_uhf(h, b, da, db, coeff, n_occ, cb, n_occ - 2, k_scale=0.5)
    assert np.max(np.abs(fa - fa_ref)) <= 1e-12 and np.max(np.abs(fb - fb_ref)) <= 1e-12
