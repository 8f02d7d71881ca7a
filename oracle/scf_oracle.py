"""CPU restatement of the reference's closed-shell SCF loop -- TEST INFRASTRUCTURE ONLY.

Follows ``run_libcint_rhf`` (backends/libcint/mqc_libcint_rhf.f90:321-680) statement for
statement around a pluggable Fock build, so that the SAME loop can be driven by

  * the oracle's exact-integral ``build_fock`` restatement (pins the loop + the J/K conventions
    to the reference-held STO-3G energies of validation/check_rhf.f90),
  * the oracle's ``build_fock_df`` on an exactly-fitting tensor,
  * the CUDA engine through the C ABI (tests/test_gpu_golden_scf.py), and
  * the device-resident SCF step of the engine, whose trajectory it checks iteration by iteration.

Pieces and their sources:
  build_orthogonalizer   src/scf/mqc_scf_common.f90:42-82      X = U s^-1/2 over eigenvalues > 1e-7
  guess_fock (GWH)       mqc_libcint_rhf.f90:1354-1380, GWH_K = 1.75 (mqc_scf_common.f90:33)
  diagonalize            mqc_libcint_rhf.f90:1464-1489         F' = X^T F X, dsyev, C = X C'
  build_density          src/scf/mqc_scf_common.f90:84-96      D = 2 C_occ C_occ^T
  commutator             mqc_libcint_rhf.f90:1326-1352         e = X^T (F D S - S D F) X
  DIIS                   src/methods/mqc_diis.f90:94-275       ring of 8, scaled B matrix, pivoted elimination
  loop / convergence     mqc_libcint_rhf.f90:566-636           iter > 1 and |dE| < e_tol and rms(dD) < d_tol
  final rebuild          mqc_libcint_rhf.f90:646-649
  run_uhf                mqc_libcint_rhf.f90:682-974           the two-spin loop: one DIIS subspace over both spins
"""
from __future__ import annotations

import numpy as np

LINEAR_DEPENDENCE_TOL = 1.0e-7      # mqc_scf_common.f90:27
GWH_K = 1.75                        # mqc_scf_common.f90:33
PIVOT_FLOOR = 1.0e-14               # mqc_diis.f90:32


def build_orthogonalizer(overlap):
    values, vectors = np.linalg.eigh(overlap, UPLO="U")
    keep = values > LINEAR_DEPENDENCE_TOL
    if not np.any(keep):
        raise ValueError("SCF: overlap matrix is singular")
    return vectors[:, keep] / np.sqrt(values[keep])[None, :]


def guess_fock_gwh(overlap, h):
    n = h.shape[0]
    hd = np.diag(h)
    f = 0.5 * GWH_K * overlap * (hd[:, None] + hd[None, :])
    f[np.arange(n), np.arange(n)] = hd
    return f


def diagonalize(fock, x):
    f_ortho = x.T @ (fock @ x)
    eigenvalues, c_prime = np.linalg.eigh(f_ortho, UPLO="U")
    return x @ c_prime, eigenvalues


def build_density_closed_shell(coeff, n_occ):
    c = coeff[:, :n_occ]
    return 2.0 * (c @ c.T)


def commutator(fock, density, overlap, x):
    fds = (fock @ density) @ overlap - (overlap @ density) @ fock
    return x.T @ (fds @ x)


def solve_diis(b_matrix):
    """Pivoted Gaussian elimination of mqc_diis.f90:228-273; returns (coefficients, ok)."""
    n = b_matrix.shape[0]
    aug = np.zeros((n, n + 1))
    aug[:, :n] = b_matrix
    aug[n - 1, n] = -1.0
    for i in range(n):
        pivot_row = i + int(np.argmax(np.abs(aug[i:, i])))
        if pivot_row != i:
            aug[[i, pivot_row], :] = aug[[pivot_row, i], :]
        pivot = aug[i, i]
        if abs(pivot) < PIVOT_FLOOR:
            return None, False
        for j in range(i + 1, n):
            factor = aug[j, i] / pivot
            aug[j, i:] = aug[j, i:] - factor * aug[i, i:]
    coef = np.zeros(n)
    for i in range(n - 1, -1, -1):
        coef[i] = (aug[i, n] - np.dot(aug[i, i + 1:n], coef[i + 1:n])) / aug[i, i]
    return coef, True


class Diis:
    """``diis_state_t``: a ring of Fock/error vectors, oldest first when extrapolating."""

    def __init__(self, max_vectors: int):
        self.max_vectors = max(1, max_vectors)
        self.focks, self.errors = [], []

    def push(self, fock, error):
        self.focks.append(np.array(fock, copy=True))
        self.errors.append(np.array(error, copy=True))
        if len(self.focks) > self.max_vectors:
            self.focks.pop(0)
            self.errors.pop(0)

    def count(self):
        return len(self.focks)

    def coefficients(self):
        n = len(self.focks)
        if n < 2:
            return None, False
        b = -np.ones((n + 1, n + 1))
        b[n, n] = 0.0
        for i in range(n):
            for j in range(n):
                b[i, j] = float(np.sum(self.errors[i] * self.errors[j]))
        scale = float(np.max(np.abs(b[:n, :n])))
        if scale > 0.0:
            b[:n, :n] /= scale
        coef, ok = solve_diis(b)
        return (coef[:n] if ok else None), ok

    def extrapolate(self, fock):
        coef, ok = self.coefficients()
        if not ok:
            return fock, False
        out = np.zeros_like(fock)
        for c, f in zip(coef, self.focks):
            out = out + c * f
        return out, True


def run_rhf(h, s, n_electrons, fock_builder, e_nuc=0.0, max_iter=100, energy_tol=1e-10, density_tol=1e-8,
            diis_vectors=8, guess="gwh", record=None, guess_density=None, guess_fock=None):
    """``run_libcint_rhf`` around ``fock_builder(h, density, coeff, n_occ) -> (fock, e_elec)``.
    Returns a dict with energy, electronic, iterations, converged, orbitals, density.
    ``guess_density`` + ``guess_fock(h, density) -> fock``: the atomic guesses of the reference enter the loop as
    ONE Fock build from the guess density (``atomic_guess_fock``, mqc_libcint_rhf.f90:1382-1411, call site :546);
    on the fitted path that build goes through the density's pseudo-orbitals."""
    n_occ = n_electrons // 2
    x = build_orthogonalizer(s)
    if guess_density is not None:
        fock = guess_fock(h, guess_density)
    else:
        fock = guess_fock_gwh(s, h) if guess == "gwh" else np.array(h, copy=True)
    coeff, eigenvalues = diagonalize(fock, x)
    density = build_density_closed_shell(coeff, n_occ)
    diis = Diis(diis_vectors)
    e_old, converged, iterations = 0.0, False, 0
    for it in range(1, max_iter + 1):
        density_old = density
        fock, e_elec = fock_builder(h, density, coeff, n_occ)
        err = commutator(fock, density, s, x)
        if diis_vectors > 0:
            diis.push(fock, err)
            fock, _ = diis.extrapolate(fock)
        coeff, eigenvalues = diagonalize(fock, x)
        density = build_density_closed_shell(coeff, n_occ)
        de = abs(e_elec - e_old)
        drms = float(np.sqrt(np.sum((density - density_old) ** 2) / density.size))
        if record is not None:
            record.append({"iter": it, "e_elec": e_elec, "de": de, "drms": drms, "diis": diis.count()})
        e_old, iterations = e_elec, it
        if it > 1 and de < energy_tol and drms < density_tol:
            converged = True
            break
    _, electronic = fock_builder(h, density, coeff, n_occ)           # final rebuild, :646-649
    return {"energy": electronic + e_nuc, "electronic": electronic, "nuclear_repulsion": e_nuc,
            "iterations": iterations, "converged": converged, "orbitals": coeff,
            "orbital_energies": eigenvalues, "density": density, "n_occupied": n_occ}


UHF_DIIS_START = 4                  # DEFAULT_UHF_DIIS_START, mqc_libcint_rhf.f90:51


def spin_contamination(c_alpha, c_beta, overlap, n_alpha, n_beta):
    """<S^2> = S_z(S_z+1) + n_beta - sum_ij |<phi_i^a|phi_j^b>|^2      src/scf/mqc_scf_common.f90:111-138"""
    sz = 0.5 * (n_alpha - n_beta)
    s2 = sz * (sz + 1.0) + n_beta
    if n_alpha == 0 or n_beta == 0:
        return s2
    mo = c_alpha[:, :n_alpha].T @ (overlap @ c_beta[:, :n_beta])
    return s2 - float(np.sum(mo ** 2))


def run_uhf(h, s, n_electrons, multiplicity, fock_builder, e_nuc=0.0, max_iter=100, energy_tol=1e-10,
            density_tol=1e-8, diis_vectors=8, guess="gwh"):
    """``run_libcint_uhf`` (mqc_libcint_rhf.f90:682-974) around
    ``fock_builder(h, d_alpha, d_beta, c_alpha, n_alpha, c_beta, n_beta) -> (fock_a, fock_b, e_elec)``:
    occupations from the multiplicity (:767-789), the same symmetric guess for both spins (:846-852), ONE DIIS
    subspace over both spins -- the two Fock matrices laid end to end, the two commutators likewise, extrapolating
    from cycle 4 (:835-838, :900-912) -- the convergence test on |dE| and the rms over both densities (:927-935),
    and the final rebuild (:939-943)."""
    if multiplicity < 1 or (n_electrons + multiplicity - 1) % 2:
        raise ValueError("UHF: an electron count and multiplicity that cannot be paired -- their parities disagree")
    n_alpha = (n_electrons + multiplicity - 1) // 2
    n_beta = n_electrons - n_alpha
    n = h.shape[0]
    x = build_orthogonalizer(s)
    fock_a = guess_fock_gwh(s, h) if guess == "gwh" else np.array(h, copy=True)
    fock_b = fock_a.copy()
    c_a, eig_a = diagonalize(fock_a, x)
    c_b, eig_b = diagonalize(fock_b, x)
    d_a = c_a[:, :n_alpha] @ c_a[:, :n_alpha].T
    d_b = c_b[:, :n_beta] @ c_b[:, :n_beta].T
    diis = Diis(diis_vectors)
    e_old, converged, iterations = 0.0, False, 0
    for it in range(1, max_iter + 1):
        d_a_old, d_b_old = d_a, d_b
        fock_a, fock_b, e_elec = fock_builder(h, d_a, d_b, c_a, n_alpha, c_b, n_beta)
        if diis_vectors > 0:
            err = np.concatenate([commutator(fock_a, d_a, s, x).reshape(-1, order="F"),
                                  commutator(fock_b, d_b, s, x).reshape(-1, order="F")])
            flat = np.concatenate([fock_a.reshape(-1, order="F"), fock_b.reshape(-1, order="F")])
            diis.push(flat, err)
            if it >= UHF_DIIS_START:
                flat, ok = diis.extrapolate(flat)
                if ok:
                    fock_a = flat[:n * n].reshape(n, n, order="F")
                    fock_b = flat[n * n:].reshape(n, n, order="F")
        c_a, eig_a = diagonalize(fock_a, x)
        c_b, eig_b = diagonalize(fock_b, x)
        d_a = c_a[:, :n_alpha] @ c_a[:, :n_alpha].T
        d_b = c_b[:, :n_beta] @ c_b[:, :n_beta].T
        de = abs(e_elec - e_old)
        drms = float(np.sqrt((np.sum((d_a - d_a_old) ** 2) + np.sum((d_b - d_b_old) ** 2)) / (2 * n * n)))
        e_old, iterations = e_elec, it
        if it > 1 and de < energy_tol and drms < density_tol:
            converged = True
            break
    _, _, electronic = fock_builder(h, d_a, d_b, c_a, n_alpha, c_b, n_beta)
    return {"energy": electronic + e_nuc, "electronic": electronic, "iterations": iterations, "converged": converged,
            "orbitals": c_a, "orbitals_beta": c_b, "density": d_a, "density_beta": d_b,
            "n_occupied": n_alpha, "n_occupied_beta": n_beta,
            "spin_squared": spin_contamination(c_a, c_b, s, n_alpha, n_beta)}


def exact_fit_tensor(eri, tol=1e-13):
    """A fitted tensor b(n*n, naux) that reproduces the four-index integrals exactly:
    (ab|cd) = sum_P b(ab,P) b(cd,P), from the eigendecomposition of the (n^2 x n^2) positive
    semi-definite ERI matrix (modes below ``tol`` of the largest dropped).  With it the DF build
    IS the exact build, so a DF-SCF on it must land on the reference's exact-integral energy."""
    n = eri.shape[0]
    m = eri.reshape(n * n, n * n, order="F")          # row index mu + n*nu, the reference's flattening
    m = 0.5 * (m + m.T)
    w, v = np.linalg.eigh(m)
    keep = w > tol * w[-1]
    return np.asfortranarray(v[:, keep] * np.sqrt(w[keep])[None, :])
