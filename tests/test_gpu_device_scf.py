"""The device-resident SCF step: a fragment's whole closed-shell SCF on the GPU
-- Fock build, commutator, DIIS, Jacobi eigensolver, density, convergence test -- against the
oracle's restatement of run_libcint_rhf driven by the oracle's build_fock_df, iteration by
iteration (electronic energies to 1e-9 Eh), and against the energies the reference itself holds."""
import numpy as np
import pytest

from metalquicha_b200 import B200Error, synth
from oracle import df_fock_oracle as oracle
from oracle import gto_integrals as gto
from oracle import scf_oracle as scf

pytestmark = pytest.mark.gpu
TOL_E = 1e-9


def _oracle_scf(h, s, b, n_electrons, **kw):
    record = []

    def fock_builder(h_, density, coeff, n_occ):
        f = oracle.build_fock_df(h_, b, density, coeff, n_occ, k_scale=kw.get("k_scale"))
        return f, oracle.electronic_energy(h_, f, density)
    args = {k: v for k, v in kw.items() if k != "k_scale"}
    res = scf.run_rhf(h, s, n_electrons, fock_builder, record=record, **args)
    return res, np.array([r["e_elec"] for r in record])


def _compare(dev, ref, hist_ref, tol=TOL_E):
    assert dev["converged"] == ref["converged"]
    assert dev["iterations"] == ref["iterations"]
    assert np.max(np.abs(dev["e_history"] - hist_ref)) <= tol
    assert abs(dev["electronic"] - ref["electronic"]) <= tol
    assert np.max(np.abs(dev["density"] - ref["density"])) <= 1e-7
    assert np.max(np.abs(dev["orbital_energies"] - ref["orbital_energies"])) <= 1e-7


@pytest.mark.parametrize("case", ["h2", "h2o"])
def test_reference_held_energies_entirely_on_the_device(engine, case):
    symbols, coords, n_electrons, e_ref = {"h2": gto.H2_STO3G, "h2o": gto.H2O_STO3G}[case]
    s, h, eri, e_nuc = gto.molecule_integrals(symbols, coords)
    b = scf.exact_fit_tensor(eri)
    engine.set_tensor(b)
    dev = engine.run_scf_fragment(h, s, n_electrons, e_nuc=e_nuc)
    assert dev["converged"] and abs(dev["energy"] - e_ref) < TOL_E          # validation/check_rhf.f90:87, :152
    ref, hist = _oracle_scf(h, s, b, n_electrons, e_nuc=e_nuc)
    _compare(dev, ref, hist)
    plain = engine.run_scf_fragment(h, s, n_electrons, e_nuc=e_nuc, diis_vectors=0, max_iter=200)
    assert plain["converged"] and abs(plain["energy"] - e_ref) < TOL_E
    if case == "h2o":
        assert dev["iterations"] < plain["iterations"]                          # check_rhf.f90:134-135


def _synthetic_fragment(seed, n, n_occ, naux, coupling=0.35, null_mode=False):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n)) * (0.15 / np.sqrt(n))
    s = np.eye(n) + a + a.T
    if null_mode:                                    # one overlap eigenvalue far below the 1e-7 threshold
        w, u = np.linalg.eigh(s)
        w[0] = 1e-10
        s = (u * w[None, :]) @ u.T
        s = 0.5 * (s + s.T)
    h = synth.synth_core_hamiltonian(seed, n) - 2.0 * np.diag(np.linspace(1.0, 0.0, n))
    b = coupling * synth.synth_tensor(seed, n, naux)
    return np.asfortranarray(s), np.asfortranarray(h), np.asfortranarray(b)


@pytest.mark.parametrize("n,n_occ,naux,kw", [
    (72, 15, 340, {}),                                   # the (H2O)3 trimer shape of BASELINE configs[2]
    (48, 10, 227, {"guess": "core"}),
    (24, 5, 113, {"diis_vectors": 3}),
    (80, 64, 60, {"k_scale": 0.2}),                      # the largest shape the fragment kernels take
    (33, 7, 50, {"diis_vectors": 0, "max_iter": 80, "coupling": 0.12}),    # odd n: the Jacobi ordering pads one index
])
def test_synthetic_fragment_trajectory_matches_the_oracle_scf(engine, n, n_occ, naux, kw):
    kw = dict(kw)
    s, h, b = _synthetic_fragment(900 + n, n, n_occ, naux, coupling=kw.pop("coupling", 0.35))
    engine.set_tensor(b)
    dev = engine.run_scf_fragment(h, s, 2 * n_occ, **kw)
    ref, hist = _oracle_scf(h, s, b, 2 * n_occ, **kw)
    _compare(dev, ref, hist)
    assert dev["orbitals"].shape == (n, n)
    # orbitals are S-orthonormal and diagonalise nothing worse than the reference's
    c = dev["orbitals"]
    assert np.max(np.abs(c.T @ s @ c - np.eye(n))) <= 1e-9
    # bit-reproducible, and queued iterations (one look at the flag every 4) change nothing
    again = engine.run_scf_fragment(h, s, 2 * n_occ, **kw)
    assert again["electronic"] == dev["electronic"] and np.array_equal(again["density"], dev["density"])
    queued = engine.run_scf_fragment(h, s, 2 * n_occ, check_every=4, **kw)
    assert queued["electronic"] == dev["electronic"] and queued["iterations"] == dev["iterations"]


def test_linearly_dependent_basis_drops_a_mode_like_the_reference(engine):
    n, n_occ, naux = 40, 8, 60
    s, h, b = _synthetic_fragment(77, n, n_occ, naux, null_mode=True)
    engine.set_tensor(b)
    dev = engine.run_scf_fragment(h, s, 2 * n_occ)
    ref, hist = _oracle_scf(h, s, b, 2 * n_occ)
    assert dev["orbitals"].shape == (n, n - 1) and ref["orbitals"].shape == (n, n - 1)
    _compare(dev, ref, hist, tol=1e-8)


def test_refusals(engine):
    n, n_occ, naux = 30, 4, 20
    s, h, b = _synthetic_fragment(5, n, n_occ, naux)
    engine.set_tensor(b)
    with pytest.raises(B200Error, match="singular"):
        engine.run_scf_fragment(h, np.zeros((n, n)), 2 * n_occ)
    with pytest.raises(B200Error, match="even"):
        engine.run_scf_fragment(h, s, 7)
    with pytest.raises(B200Error, match="more occupied"):
        engine.run_scf_fragment(h, s, 2 * (n + 1))
    big = synth.synth_tensor(1, 96, 10)
    engine.set_tensor(big)
    with pytest.raises(B200Error, match="fragment-sized"):
        engine.run_scf_fragment(np.eye(96), np.eye(96), 4)


@pytest.mark.parametrize("n,n_occ,naux,nf", [(72, 15, 340, 6), (24, 5, 113, 33), (40, 8, 60, 3)])
def test_batch_of_fragments_in_lock_step(engine, n, n_occ, naux, nf):
    """mqcb200_scf_fragment_batch: nf fragments of one kind, tensors back to back on the slot, every
    kernel launched once per iteration with the fragment index on a grid axis.  Each fragment must
    follow the oracle's SCF (its own iteration count included) and agree with the one-at-a-time path."""
    problems = [_synthetic_fragment(4000 + 17 * f + n, n, n_occ, naux, coupling=0.25 + 0.02 * (f % 5),
                                    null_mode=(n == 40 and f == 1)) for f in range(nf)]
    s_all = np.stack([p[0] for p in problems])
    h_all = np.stack([p[1] for p in problems])
    engine.set_tensor(np.asfortranarray(np.hstack([p[2] for p in problems])), n=n)
    out = engine.run_scf_fragment_batch(h_all, s_all, 2 * n_occ)
    check = range(nf) if nf <= 8 else (0, 7, nf - 1)
    for f in check:
        s, h, b = problems[f]
        ref, hist = _oracle_scf(h, s, b, 2 * n_occ)
        assert out["converged"][f] == 1 and ref["converged"]
        assert out["iterations"][f] == ref["iterations"]
        assert abs(out["electronic"][f] - ref["electronic"]) <= TOL_E
        assert np.max(np.abs(out["density"][f] - ref["density"])) <= 1e-7
        assert out["orbitals"][f].shape == ref["orbitals"].shape
    # one-at-a-time on the same handle gives the same energies (different partial-sum grouping: 1e-11)
    s, h, b = problems[0]
    engine.set_tensor(b)
    one = engine.run_scf_fragment(h, s, 2 * n_occ)
    assert abs(one["electronic"] - out["electronic"][0]) <= 1e-10 and one["iterations"] == out["iterations"][0]
    # repeat: bit-identical; queued iterations change nothing
    engine.set_tensor(np.asfortranarray(np.hstack([p[2] for p in problems])), n=n)
    again = engine.run_scf_fragment_batch(h_all, s_all, 2 * n_occ, check_every=3, want_matrices=False)
    assert np.array_equal(again["electronic"], out["electronic"]) and np.array_equal(again["iterations"], out["iterations"])


def test_batch_refusals(engine):
    n, n_occ, naux = 30, 4, 20
    s, h, b = _synthetic_fragment(5, n, n_occ, naux)
    engine.set_tensor(np.asfortranarray(np.hstack([b, b, b])), n=n)
    with pytest.raises(B200Error, match="back to back"):
        engine.run_scf_fragment_batch(np.stack([h] * 7), np.stack([s] * 7), 2 * n_occ)    # 60 slabs do not split in 7
    out = engine.run_scf_fragment_batch(np.stack([h] * 3), np.stack([s] * 3), 2 * n_occ)
    assert np.all(out["converged"] == 1) and out["electronic"][0] == out["electronic"][1] == out["electronic"][2]
    # a fragment whose overlap leaves fewer orbitals than occupied is reported (-1), the others still converge
    bad = np.zeros_like(s); bad[0, 0] = 1.0
    out = engine.run_scf_fragment_batch(np.stack([h, h, h]), np.stack([s, bad, s]), 2 * n_occ)
    assert list(out["converged"]) == [1, -1, 1]


# ---- the general-size route of mqcb200_scf (n > 80 or n_occ > 64): general Fock-build kernels, batched DMMA
# GEMM for the dense algebra, one-sided Jacobi on the shifted F' in place of dsyev
@pytest.mark.parametrize("n,n_occ,naux,kw", [
    (96, 20, 80, {}),
    (130, 70, 60, {"k_scale": 0.2}),                      # n_occ > 64
    (150, 33, 90, {"guess": "core", "diis_vectors": 4}),
    (97, 11, 50, {"diis_vectors": 0, "max_iter": 80, "coupling": 0.1}),   # odd n, plain iteration
])
def test_general_size_device_scf_matches_the_oracle_scf(engine, n, n_occ, naux, kw):
    kw = dict(kw)
    s, h, b = _synthetic_fragment(7000 + n, n, n_occ, naux, coupling=kw.pop("coupling", 0.3))
    engine.set_tensor(b)
    dev = engine.run_scf(h, s, 2 * n_occ, **kw)
    ref, hist = _oracle_scf(h, s, b, 2 * n_occ, **kw)
    _compare(dev, ref, hist)
    c = dev["orbitals"]
    assert np.max(np.abs(c.T @ s @ c - np.eye(c.shape[1]))) <= 1e-9
    again = engine.run_scf(h, s, 2 * n_occ, **kw)
    assert again["electronic"] == dev["electronic"] and np.array_equal(again["density"], dev["density"])


def test_general_size_route_on_real_integrals_and_dispatch(engine):
    """Water/STO-3G padded to n = 96 runs the general route to the reference-held energy; at n = 7 mqcb200_scf
    dispatches to the one-CTA route and gives what mqcb200_scf_fragment gives."""
    symbols, coords, n_electrons, e_ref = gto.H2O_STO3G
    s0, h0, eri, e_nuc = gto.molecule_integrals(symbols, coords)
    b0 = scf.exact_fit_tensor(eri)
    n0, n, naux = h0.shape[0], 96, b0.shape[1]
    s_p, h_p = np.eye(n), 1.0e3 * np.eye(n)
    s_p[:n0, :n0], h_p[:n0, :n0] = s0, h0
    b_p = np.zeros((n * n, naux), order="F")
    for p in range(naux):
        slab = np.zeros((n, n)); slab[:n0, :n0] = b0[:, p].reshape(n0, n0, order="F")
        b_p[:, p] = slab.reshape(n * n, order="F")
    engine.set_tensor(b_p)
    dev = engine.run_scf(h_p, s_p, n_electrons, e_nuc=e_nuc)
    assert dev["converged"] and abs(dev["energy"] - e_ref) < TOL_E          # validation/check_rhf.f90:152
    engine.set_tensor(b0)
    small = engine.run_scf(h0, s0, n_electrons, e_nuc=e_nuc)
    frag = engine.run_scf_fragment(h0, s0, n_electrons, e_nuc=e_nuc)
    assert small["electronic"] == frag["electronic"] and small["iterations"] == frag["iterations"]
    assert abs(small["energy"] - e_ref) < TOL_E


def test_general_size_linear_dependence(engine):
    n, n_occ, naux = 100, 12, 40
    s, h, b = _synthetic_fragment(88, n, n_occ, naux, null_mode=True)
    engine.set_tensor(b)
    dev = engine.run_scf(h, s, 2 * n_occ)
    ref, hist = _oracle_scf(h, s, b, 2 * n_occ)
    assert dev["orbitals"].shape == (n, n - 1) and ref["orbitals"].shape == (n, n - 1)
    _compare(dev, ref, hist, tol=1e-8)
