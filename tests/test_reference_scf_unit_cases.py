"""The reference's OWN unit cases for the pieces of the SCF step, run against the oracle (CPU).

SURVEY 8f row 2 (the device-resident SCF step) is checked on the GPU against ``oracle/scf_oracle.py``;
this file pins that oracle to the cases the reference's test suite holds for the same routines:

* ``test/test_mqc_diis.f90`` (diis_state_t, src/methods/mqc_diis.f90): the six cases, with the test's own
  deterministic filler (its LCG, :314-332), sizes and seeds, and its comparison target -- the
  shift-the-history, rebuild-B algorithm of :256-312, restated here as ``reference_extrapolate``;
* ``test/test_mqc_scf_common.f90`` (build_orthogonalizer / build_density_closed_shell / build_density_spin,
  src/scf/mqc_scf_common.f90:42-109): its model overlaps and its assertions, same tolerances.

Nothing here touches the GPU; the CUDA step is tied to the same oracle by tests/test_gpu_device_scf.py.
"""
import numpy as np
import pytest

from oracle import df_fock_oracle as oracle
from oracle import scf_oracle as scf

REFERENCE_TOL = 1.0e-12       # test_mqc_diis.f90:10
TOL = 1.0e-10                 # test_mqc_scf_common.f90:13


def fill_pseudorandom(rows, cols, seed):
    """test_mqc_diis.f90:314-332: column-major fill from the LCG 1103515245 / 12345 mod 2^31, minus 0.5."""
    a = np.empty((rows, cols))
    state = int(seed)
    for j in range(cols):
        for i in range(rows):
            state = (1103515245 * state + 12345) % 2147483648
            a[i, j] = state / 2147483648.0 - 0.5
    return a


def reference_extrapolate(hist_f, hist_e, n_stored):
    """test_mqc_diis.f90:256-312, the pre-ring-buffer algorithm the reference keeps as its comparison
    target: B from direct overlaps (NOT rescaled), bordered with -1 / 0, pivoted elimination with the
    1e-14 pivot floor, back substitution, F = sum c_i F_i."""
    fock = np.zeros(hist_f.shape[0])
    if n_stored < 2:
        return fock, False
    n = n_stored + 1
    b = -np.ones((n, n))
    b[n - 1, n - 1] = 0.0
    for i in range(n_stored):
        for j in range(n_stored):
            b[i, j] = float(np.sum(hist_e[:, i] * hist_e[:, j]))
    aug = np.zeros((n, n + 1))
    aug[:, :n] = b
    aug[n - 1, n] = -1.0
    for i in range(n):
        pivot_row = i
        for j in range(i + 1, n):
            if abs(aug[j, i]) > abs(aug[pivot_row, i]):
                pivot_row = j
        if pivot_row != i:
            aug[[i, pivot_row], :] = aug[[pivot_row, i], :]
        pivot = aug[i, i]
        if abs(pivot) < 1.0e-14:
            return fock, False
        for j in range(i + 1, n):
            factor = aug[j, i] / pivot
            aug[j, i:] = aug[j, i:] - factor * aug[i, i:]
    coef = np.zeros(n)
    for i in range(n - 1, -1, -1):
        coef[i] = (aug[i, n] - float(np.sum(aug[i, i + 1:n] * coef[i + 1:n]))) / aug[i, i]
    for i in range(n_stored):
        fock = fock + coef[i] * hist_f[:, i]
    return fock, True


# ---- test_mqc_diis.f90 -------------------------------------------------------------------------------
def test_diis_no_extrapolation_below_two():
    """:30-54 -- an empty subspace and a single stored vector leave fock untouched."""
    diis = scf.Diis(4)
    fock = np.array([1.0, 2.0, 3.0, 4.0])
    out, ok = diis.extrapolate(fock)
    assert not ok and np.array_equal(out, fock)
    diis.push(fock, np.array([1.0, 0.0, 0.0, 0.0]))
    out, ok = diis.extrapolate(fock)
    assert not ok and np.array_equal(out, fock)


def test_diis_zero_error_reproduces_fock():
    """:56-78 -- with one (nearly) converged entry among others DIIS lands on it."""
    diis = scf.Diis(4)
    diis.push(np.array([1.0, 1.0, 1.0]), np.array([1.0, 0.0, 0.0]))
    diis.push(np.array([2.0, 2.0, 2.0]), np.array([0.0, 1.0, 0.0]))
    target = np.array([7.0, 8.0, 9.0])
    diis.push(target, np.array([1.0e-12, 0.0, 0.0]))
    out, ok = diis.extrapolate(np.zeros(3))
    assert ok and np.max(np.abs(out - target)) < 1.0e-6


def test_diis_coefficients_sum_to_one():
    """:80-109 -- a constant history extrapolates to that constant (sum c = 1)."""
    nf, npush, constant = 5, 4, 3.25
    e_in = fill_pseudorandom(nf, npush, 7)
    diis = scf.Diis(6)
    for i in range(npush):
        diis.push(np.full(nf, constant), e_in[:, i])
    out, ok = diis.extrapolate(np.zeros(nf))
    assert ok and np.max(np.abs(out - constant)) < 1.0e-10
    coef, ok = diis.coefficients()
    assert ok and abs(float(np.sum(coef)) - 1.0) < 1.0e-12


def test_diis_ring_evicts_oldest():
    """:111-155 -- a history wrapped many times equals one holding only the last max_vectors entries."""
    nf, nmax, npush = 6, 3, 10
    f_in, e_in = fill_pseudorandom(nf, npush, 17), fill_pseudorandom(nf, npush, 23)
    wrapped, fresh = scf.Diis(nmax), scf.Diis(nmax)
    for i in range(npush):
        wrapped.push(f_in[:, i], e_in[:, i])
    for i in range(npush - nmax, npush):
        fresh.push(f_in[:, i], e_in[:, i])
    assert wrapped.count() == nmax
    a, ok_a = wrapped.extrapolate(np.zeros(nf))
    b, ok_b = fresh.extrapolate(np.zeros(nf))
    assert ok_a and ok_b and np.array_equal(a, b)


def test_diis_matches_reference_algorithm():
    """:157-223 -- step by step against the shift-the-history algorithm: same solvability, extrapolated
    Fock within 1e-12 * max(1, |F|) (the oracle rescales B before the solve, as mqc_diis.f90:206-209 does;
    the comparison target does not: agreement to rounding is the point of the reference's test too)."""
    nf, ne, nmax, npush = 12, 12, 4, 9
    f_in, e_in = fill_pseudorandom(nf, npush, 11), fill_pseudorandom(ne, npush, 29)
    diis = scf.Diis(nmax)
    hist_f, hist_e = np.zeros((nf, nmax)), np.zeros((ne, nmax))
    n_stored = 0
    for step in range(npush):
        diis.push(f_in[:, step], e_in[:, step])
        if n_stored < nmax:
            n_stored += 1
        else:
            hist_f[:, :nmax - 1] = hist_f[:, 1:nmax].copy()
            hist_e[:, :nmax - 1] = hist_e[:, 1:nmax].copy()
        hist_f[:, n_stored - 1] = f_in[:, step]
        hist_e[:, n_stored - 1] = e_in[:, step]
        out, ok = diis.extrapolate(np.zeros(nf))
        ref, ok_ref = reference_extrapolate(hist_f, hist_e, n_stored)
        assert ok == ok_ref, f"solvability disagrees at step {step + 1}"
        if ok:
            scale = max(1.0, float(np.max(np.abs(ref))))
            assert float(np.max(np.abs(out - ref))) <= REFERENCE_TOL * scale, f"step {step + 1}"


def test_diis_overlaps_are_direct_dot_products():
    """:225-254 -- the B entries the solve sees are the plain dot products of the stored error vectors
    (oldest first), up to the common rescaling."""
    ne, nmax, npush = 7, 3, 8
    e_in, f_in = fill_pseudorandom(ne, npush, 5), fill_pseudorandom(ne, npush, 13)
    diis = scf.Diis(nmax)
    for step in range(npush):
        diis.push(f_in[:, step], e_in[:, step])
    assert diis.count() == nmax
    for i in range(nmax):
        for j in range(nmax):
            assert np.array_equal(diis.errors[i], e_in[:, npush - nmax + i])
            direct = float(np.sum(e_in[:, npush - nmax + i] * e_in[:, npush - nmax + j]))
            assert float(np.sum(diis.errors[i] * diis.errors[j])) == direct


# ---- test_mqc_scf_common.f90 --------------------------------------------------------------------------
def model_overlap(n, off_diagonal):
    """:39-56 -- unit diagonal, one constant everywhere else."""
    s = np.full((n, n), off_diagonal)
    np.fill_diagonal(s, 1.0)
    return s


def test_orthogonalizer_diagonalizes_the_overlap():
    """:58-82 -- X^T S X = 1."""
    s = model_overlap(4, 0.25)
    x = scf.build_orthogonalizer(s)
    assert np.max(np.abs(x.T @ s @ x - np.eye(x.shape[1]))) <= TOL


def test_orthogonalizer_keeps_every_mode_when_well_conditioned():
    """:84-99."""
    assert scf.build_orthogonalizer(model_overlap(5, 0.1)).shape == (5, 5)


def test_orthogonalizer_drops_a_null_mode():
    """:101-121 -- a rank-deficient overlap is survivable: two of three modes kept."""
    s = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 1.0], [0.0, 1.0, 1.0]])
    assert scf.build_orthogonalizer(s).shape == (3, 2)


def test_orthogonalizer_refuses_a_singular_overlap():
    """:123-135 -- with the reference's message (mqc_scf_common.f90:69)."""
    with pytest.raises(ValueError, match="SCF: overlap matrix is singular"):
        scf.build_orthogonalizer(np.zeros((3, 3)))


def test_closed_shell_density_is_idempotent():
    """:137-161 -- D S D = 2 D to 1e-9."""
    s = model_overlap(4, 0.2)
    x = scf.build_orthogonalizer(s)
    for build in (scf.build_density_closed_shell, oracle.build_density_closed_shell):
        d = build(x, 2)
        assert np.max(np.abs(d @ s @ d - 2.0 * d)) <= 1.0e-9


def test_closed_shell_density_traces_to_the_electron_count():
    """:163-185 -- tr(D S) = 6 for three occupied orbitals."""
    s = model_overlap(4, 0.15)
    x = scf.build_orthogonalizer(s)
    assert abs(float(np.trace(scf.build_density_closed_shell(x, 3) @ s)) - 6.0) <= 1.0e-9


def test_spin_density_is_half_the_closed_shell_one():
    """:187-210."""
    c = np.eye(3)
    assert np.max(np.abs(oracle.build_density_closed_shell(c, 2) - 2.0 * oracle.build_density_spin(c, 2))) <= TOL


def test_density_of_no_electrons_is_zero():
    """:212-223."""
    assert np.max(np.abs(oracle.build_density_closed_shell(np.ones((3, 3)), 0))) <= TOL
