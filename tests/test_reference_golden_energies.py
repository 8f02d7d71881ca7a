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
