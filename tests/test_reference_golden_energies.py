"""The pin of the oracle to the REFERENCE's own golden numbers.

validation/check_rhf.f90 of the reference writes its STO-3G basis out inline and asserts
    H2  / STO-3G, R = 1.4 bohr      E = -1.1167143251    (:87-88, tolerance 1e-9)
    H2O / STO-3G, standard geometry E = -74.9658162796   (:152-153, tolerance 1e-9)
through ``run_libcint_rhf`` -> ``assemble_fock`` -> ``build_fock`` (the exact-integral twin of
``build_fock_df``).  With the integrals restated (oracle/gto_integrals.py -- libcint is absent)
the oracle's SCF loop must land on those numbers, and it must do so through EVERY Fock build the
oracle holds: the exact-integral restatement, and the density-fitted restatements (NumPy,
plain C, C on BLAS) on a tensor that fits the four-index integrals exactly.  That ties the
factor 2 of K, the 1/2 of the Fock matrix, the energy expression and the (mu,nu) flattening of
the fitted path to numbers the reference itself holds.

Second half of the file -- the DENSITY-FITTED path proper.  The reference's validation manifest
(validation/validation_tests_cpu.json, tolerance 1e-9 Eh: validation/run_validation.py:398) holds
    H2O / 6-31G*  fitted with 6-31G*    E = -76.188111755038   (:899-904)   [exact integrals: -76.010317945971, :719-723]
    CH4 / 6-31G** fitted with 6-31G**   E = -40.381603512964   (:905-910)
through ``build_df_tensor`` (three_centre -> metric_inverse_sqrt -> whitening GEMM) and ``build_fock_df``
inside ``run_libcint_rhf``.  6-31G* / 6-31G** are published data (oracle/gto_integrals.py cites the papers;
the H and O numbers are cross-checked below against the copy the reference tree holds in GAMESS form), so
the whole fitted path of the oracle -- (mu nu|P), (P|Q), metric^(-1/2) with its 1e-10 threshold, the
(mu + n*nu, P) flattening, J, K with its factor 2, the energy -- is run on real integrals and must land on
those numbers.  The fitting error here is 0.18 Eh: an agreement to 1e-9 leaves no room for a wrong convention.
"""
import numpy as np
import pytest

from oracle import df_fock_oracle as oracle
from oracle import gto_integrals as gto
from oracle import scf_oracle as scf

CASES = {"h2": gto.H2_STO3G, "h2o": gto.H2O_STO3G}
TOL_E = 1e-9      # the reference's own tolerance (check_rhf.f90:87, :152)


@pytest.fixture(scope="module", params=list(CASES))
def molecule(request):
    symbols, coords, n_electrons, e_ref = CASES[request.param]
    s, h, eri, e_nuc = gto.molecule_integrals(symbols, coords)
    return request.param, s, h, eri, e_nuc, n_electrons, e_ref


def test_shapes_and_nuclear_repulsion(molecule):
    name, s, h, eri, e_nuc, _, _ = molecule
    assert s.shape[0] == {"h2": 2, "h2o": 7}[name]                  # check_rhf.f90:71, :121
    if name == "h2":
        assert abs(e_nuc - 1.0 / 1.4) < 1e-12                       # check_rhf.f90:83-84
    assert np.allclose(np.diag(s), 1.0, atol=1e-12)
    assert np.allclose(eri, eri.transpose(1, 0, 2, 3)) and np.allclose(eri, eri.transpose(2, 3, 0, 1))


def test_exact_integral_scf_reproduces_the_reference_energy(molecule):
    name, s, h, eri, e_nuc, n_electrons, e_ref = molecule

    def fock_builder(h_, density, coeff, n_occ):
        f = oracle.build_fock_eri(h_, eri, density)                 # rhf.f90:1491-1574
        return f, oracle.electronic_energy(h_, f, density)
    res = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=e_nuc)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E
    plain = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=e_nuc, diis_vectors=0, max_iter=200)
    assert plain["converged"] and abs(plain["energy"] - res["energy"]) < TOL_E      # check_rhf.f90:130-131
    if name == "h2o":
        assert res["iterations"] < plain["iterations"]                               # check_rhf.f90:134-135


@pytest.mark.parametrize("builder", ["numpy_loops", "plain_c", "c_on_blas"])
def test_fitted_scf_on_an_exactly_fitting_tensor_reproduces_it_too(molecule, builder):
    name, s, h, eri, e_nuc, n_electrons, e_ref = molecule
    b = scf.exact_fit_tensor(eri)
    n = h.shape[0]
    fitted = (b @ b.T).reshape(n, n, n, n, order="F")
    assert np.max(np.abs(fitted - eri)) < 1e-12                     # the fit IS exact

    if builder == "numpy_loops":
        build = oracle.build_fock_df
    elif builder == "c_on_blas":
        from oracle import df_fock_blas
        build = df_fock_blas.build_fock_df
    else:
        import ctypes
        import os
        lib = ctypes.CDLL(os.path.join(os.path.dirname(oracle.__file__), "_build", "libdf_fock_ref.so"))
        dp = ctypes.c_void_p
        lib.df_ref_build_fock.restype = ctypes.c_int
        lib.df_ref_build_fock.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_double, dp]

        def build(h_, b_, density, coeff, n_occ):
            out = np.empty((n, n), order="F")
            hh, bb, dd, cc = (np.asfortranarray(x, dtype=np.float64) for x in (h_, b_, density, coeff))
            assert lib.df_ref_build_fock(n, b_.shape[1], n_occ, hh.ctypes.data, bb.ctypes.data, dd.ctypes.data,
                                         cc.ctypes.data, cc.shape[0], 1.0, 1.0, out.ctypes.data) == 0
            return out

    def fock_builder(h_, density, coeff, n_occ):
        f = build(h_, b, density, coeff, n_occ)
        return f, oracle.electronic_energy(h_, f, density)
    res = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=e_nuc)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E


# =====================================================================================================
# The density-fitted validation energies of the reference (see the module docstring)
# =====================================================================================================
@pytest.fixture(scope="module", params=list(gto.DF_CASES))
def df_case(request):
    s, h, three, metric, e_nuc, n_electrons, e_ref = gto.df_case_integrals(request.param)
    return request.param, s, h, three, metric, e_nuc, n_electrons, e_ref


def _plain_c_builder(n):
    import ctypes
    import os
    lib = ctypes.CDLL(os.path.join(os.path.dirname(oracle.__file__), "_build", "libdf_fock_ref.so"))
    dp = ctypes.c_void_p
    lib.df_ref_build_fock.restype = ctypes.c_int
    lib.df_ref_build_fock.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, dp]

    def build(h_, b_, density, coeff, n_occ):
        out = np.empty((n, n), order="F")
        hh, bb, dd, cc = (np.asfortranarray(x, dtype=np.float64) for x in (h_, b_, density, coeff))
        assert lib.df_ref_build_fock(n, b_.shape[1], n_occ, hh.ctypes.data, bb.ctypes.data, dd.ctypes.data,
                                     cc.ctypes.data, cc.shape[0], 1.0, 1.0, out.ctypes.data) == 0
        return out
    return build


@pytest.mark.parametrize("builder", ["numpy_loops", "plain_c", "c_on_blas"])
def test_density_fitted_scf_reproduces_the_reference_held_df_energy(df_case, builder):
    name, s, h, three, metric, e_nuc, n_electrons, e_ref = df_case
    n, naux = h.shape[0], metric.shape[0]
    assert (n, naux) == {"h2o_631gs": (19, 19), "ch4_631gss": (35, 35)}[name]     # test_mqc_libcint_cartesian.f90:265-274: B is (361, 19)
    b = oracle.whiten(three, metric)                                                 # integrals.F90:981-987
    assert b.shape == (n * n, naux)
    if builder == "numpy_loops":
        build = oracle.build_fock_df
    elif builder == "c_on_blas":
        from oracle import df_fock_blas
        build = df_fock_blas.build_fock_df
    else:
        build = _plain_c_builder(n)

    def fock_builder(h_, density, coeff, n_occ):
        f = build(h_, b, density, coeff, n_occ)
        return f, oracle.electronic_energy(h_, f, density)
    res = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10, max_iter=100)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E, res["energy"]
    if builder == "numpy_loops":
        plain = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10,
                            diis_vectors=0, max_iter=300)
        assert plain["converged"] and abs(plain["energy"] - e_ref) < TOL_E


def test_atomic_guess_through_pseudo_orbitals_lands_on_the_same_reference_energy():
    """test/test_mqc_libcint_guess.f90:49 "guess_choice_does_not_change_the_energy", on the fitted path: an atomic
    guess enters run_libcint_rhf as ONE build_fock_df from the guess density through ``density_pseudo_orbitals``
    (atomic_guess_fock, rhf.f90:1382-1411; the factorisation :1413-1462) -- SURVEY 8 row a3.  From a crude
    superposition-of-atoms density (block diagonal over the atoms, not idempotent, 10 electrons) the loop must
    converge onto the reference-held fitted energy, and the pseudo-orbitals must rebuild their density to 1e-11
    (test_mqc_libcint_guess.f90:332)."""
    s, h, three, metric, e_nuc, n_electrons, e_ref = gto.df_case_integrals("h2o_631gs")
    b = oracle.whiten(three, metric)
    n = h.shape[0]
    # free-atom-like occupations on the diagonal blocks: O 1s2 2s2 2p4 spread over its functions, H 1s1 each
    occ = np.zeros(n)
    occ[0], occ[1], occ[2:5], occ[5], occ[6:9] = 2.0, 1.4, 1.0, 0.6, 1.0 / 3.0        # oxygen s, sp inner, sp outer
    occ[15], occ[16], occ[17], occ[18] = 0.7, 0.3, 0.7, 0.3                              # the two hydrogens
    guess = np.diag(occ)
    guess *= n_electrons / float(np.sum(guess * s))                                       # tr(D S) = N
    pseudo, n_modes = oracle.density_pseudo_orbitals(guess)
    assert np.max(np.abs(2.0 * pseudo @ pseudo.T - guess)) < 1e-11
    assert n_modes > n_electrons // 2                                                     # not an idempotent density

    def fock_builder(h_, density, coeff, n_occ):
        f = oracle.build_fock_df(h_, b, density, coeff, n_occ)
        return f, oracle.electronic_energy(h_, f, density)

    def guess_fock(h_, density):
        c, m = oracle.density_pseudo_orbitals(density)
        return oracle.build_fock_df(h_, b, density, c, m)
    res = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10,
                      guess_density=guess, guess_fock=guess_fock)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E, res["energy"]


def test_the_fitting_error_itself_matches():
    """check_df.f90:60-63 asserts the gap between the fitted and the exact energy, "the sharper one": a fitted
    energy could come out right with a wrong metric if errors cancelled, the error itself could not.  Same here on
    the pair of numbers the manifest holds for water / 6-31G*: exact integrals through build_fock (rhf.f90:1491-1574)."""
    (symbols, coords, n_electrons, e_df), table = gto.DF_CASES["h2o_631gs"]
    basis = gto.build_basis(symbols, coords, table)
    charges = [gto.CHARGE[x] for x in symbols]
    s, t, v = gto.one_electron(basis, charges, coords)
    eri = gto.electron_repulsion(basis)
    h = t + v

    def fock_builder(h_, density, coeff, n_occ):
        f = oracle.build_fock_eri(h_, eri, density)
        return f, oracle.electronic_energy(h_, f, density)
    exact = scf.run_rhf(h, s, n_electrons, fock_builder, e_nuc=gto.nuclear_repulsion(charges, coords),
                        energy_tol=1e-12, density_tol=1e-10)
    assert exact["converged"] and abs(exact["energy"] - gto.H2O_631GS_EXACT_ENERGY) < TOL_E
    _, _, three, metric, e_nuc, _, _ = gto.df_case_integrals("h2o_631gs")
    b = oracle.whiten(three, metric)

    def df_builder(h_, density, coeff, n_occ):
        f = oracle.build_fock_df(h_, b, density, coeff, n_occ)
        return f, oracle.electronic_energy(h_, f, density)
    fitted = scf.run_rhf(h, s, n_electrons, df_builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10)
    gap_ref = e_df - gto.H2O_631GS_EXACT_ENERGY                    # -0.177793809067
    assert abs((fitted["energy"] - exact["energy"]) - gap_ref) < TOL_E
    assert abs(gap_ref) > 0.1                                      # a coarse fit on purpose: nothing cancels by accident


def test_basis_table_agrees_with_the_copy_the_reference_tree_holds():
    """tools/efp_validation/reference/water_6-31gs_cmo.efp:1711-1743 (PROJECTION BASIS SET, written by GAMESS from
    the same 6-31G* water): exponents, and contraction coefficients multiplied by the primitive norms, to eight
    digits.  The numbers below are that block; the table of oracle/gto_integrals.py must reproduce them."""
    import math
    held = {
        ("O", 0, 0): ([5484.6716600000, 825.2349460000, 188.0469580000, 52.9645000000, 16.8975704000, 5.7996353400],
                      [0.83172368, 1.53081556, 2.47714854, 3.25628110, 2.79289337, 0.95493768]),
        ("O", 1, 0): ([15.5396162453, 3.5999335863, 1.0137617501], [-0.61793396, -0.27572093, 0.81420760]),
        ("O", 2, 1): ([15.5396162453, 3.5999335863, 1.0137617501], [3.11694427, 2.40143753, 1.05436042]),
        ("O", 3, 0): ([0.2700058226], [0.26695616]),
        ("O", 4, 1): ([0.2700058226], [0.27743197]),
        ("O", 5, 2): ([0.8000000000], [1.11382493]),
        ("H", 0, 0): ([18.7311369600, 2.8253943648, 0.6401216923], [0.21493545, 0.36457120, 0.41505143]),
        ("H", 1, 0): ([0.1612777588], [0.18138065]),
    }

    def double_factorial(k):
        return 1 if k <= 0 else k * double_factorial(k - 2)
    for (sym, idx, l), (exps, gamess) in held.items():
        shell_l, t_exps, t_coefs = gto.POPLE_631GS[sym][idx]
        assert shell_l == l and len(t_exps) == len(exps)
        for a_held, a_table in zip(exps, t_exps):
            assert abs(a_held - a_table) <= 1e-9 * a_held + 5e-10            # the table carries BSE's ten digits
        # GAMESS prints coefficient * norm of the primitive x^l exp(-a r^2)
        for a, c, g in zip(t_exps, t_coefs, gamess):
            norm = (2.0 * a / math.pi) ** 0.75 * (4.0 * a) ** (l / 2.0) / math.sqrt(double_factorial(2 * l - 1))
            assert abs(c * norm - g) <= 2e-8 * max(1.0, abs(g)), (sym, idx, a)


# =====================================================================================================
# The two-spin build against a reference-held unrestricted energy
# =====================================================================================================
def test_two_spin_fitted_build_reproduces_the_reference_held_uhf_energy():
    """validation_tests_cpu.json:1844-1849: UHF OH / STO-3G, doublet, E = -74.362637545612 (1e-9), computed by
    run_libcint_uhf with build_fock_uhf (mqc_libcint_rhf.f90:682-974, :1648-1681).  The reference's CPU path has no
    fitted two-spin build; the engine's (and the oracle's jk_df_uhf) takes its conventions from the cuEST loop --
    J from the TOTAL density, K_sigma = sum_P (B_P C_sigma)(B_P C_sigma)^T with no factor 2, the energy
    1/2 [sum (Da+Db) H + sum Da Fa + sum Db Fb].  On a tensor that fits the four-index integrals exactly the
    fitted two-spin build IS build_fock_uhf, so the restated loop must land on the reference's number."""
    symbols, coords, n_electrons, multiplicity, e_ref = gto.OH_STO3G_UHF
    s, h, eri, e_nuc = gto.molecule_integrals(symbols, coords, gto.STO3G_BSE)
    b = scf.exact_fit_tensor(eri)
    n = h.shape[0]
    assert np.max(np.abs((b @ b.T).reshape(n, n, n, n, order="F") - eri)) < 1e-12

    def fock_builder(h_, d_a, d_b, c_a, n_a, c_b, n_b):
        f_a, f_b = oracle.build_fock_df_uhf(h_, b, d_a, d_b, c_a, n_a, c_b, n_b)
        return f_a, f_b, oracle.uhf_electronic_energy(h_, f_a, f_b, d_a, d_b)
    res = scf.run_uhf(h, s, n_electrons, multiplicity, fock_builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10)
    assert res["converged"] and (res["n_occupied"], res["n_occupied_beta"]) == (5, 4)
    assert abs(res["energy"] - e_ref) < TOL_E, res["energy"]
    assert 0.75 <= res["spin_squared"] < 0.76                       # a doublet with next to no contamination
    plain = scf.run_uhf(h, s, n_electrons, multiplicity, fock_builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10,
                        diis_vectors=0, max_iter=300)
    assert plain["converged"] and abs(plain["energy"] - e_ref) < TOL_E


def test_bse_sto3g_table_agrees_with_the_inline_one():
    """The ten-digit STO-3G table against the inline table of check_rhf.f90:148-178 (an older tabulation with
    seven to eight significant digits): the same basis set to 3e-7 relative in every exponent and coefficient."""
    for sym in ("H", "O"):
        for (l8, e8, c8), (l10, e10, c10) in zip(gto.STO3G[sym], gto.STO3G_BSE[sym]):
            assert l8 == l10
            for a, b_ in zip(list(e8) + list(c8), list(e10) + list(c10)):
                assert abs(a - b_) <= 3e-7 * abs(b_)
