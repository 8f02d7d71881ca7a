"""The CUDA engine inside the reference's SCF loop, on REAL integrals, against the energies the
reference itself holds (validation/check_rhf.f90:87-88, :152-153; tolerance 1e-9 Eh there and here).

The tensor fits the STO-3G four-index integrals exactly (oracle/scf_oracle.exact_fit_tensor), so the
density-fitted build IS the exact build and the converged energy must be the reference's number.
Two routes through the engine: the fragment-sized one-pass kernel (n = 7), and -- with the same
molecule embedded in a basis padded to n = 96 by far-away, never-occupied functions -- the general
J kernels and the TMA-fed DMMA exchange kernels.
"""
import numpy as np
import pytest

from oracle import gto_integrals as gto
from oracle import scf_oracle as scf

pytestmark = pytest.mark.gpu
TOL_E = 1e-9


@pytest.fixture(scope="module")
def water():
    symbols, coords, n_electrons, e_ref = gto.H2O_STO3G
    s, h, eri, e_nuc = gto.molecule_integrals(symbols, coords)
    return s, h, scf.exact_fit_tensor(eri), e_nuc, n_electrons, e_ref


def _engine_builder(engine):
    def fock_builder(h, density, coeff, n_occ):
        fock = engine.build_fock_df(np.asfortranarray(h), np.asfortranarray(density), np.asfortranarray(coeff), n_occ)
        return fock, engine.last_energy()
    return fock_builder


def test_water_sto3g_scf_through_the_fragment_kernel(engine, water):
    s, h, b, e_nuc, n_electrons, e_ref = water
    engine.set_tensor(b)
    record = []
    res = scf.run_rhf(h, s, n_electrons, _engine_builder(engine), e_nuc=e_nuc, record=record)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E
    assert engine.last_launches() > 0 and len(record) == res["iterations"]


def test_h2_sto3g_scf(engine):
    symbols, coords, n_electrons, e_ref = gto.H2_STO3G
    s, h, eri, e_nuc = gto.molecule_integrals(symbols, coords)
    engine.set_tensor(scf.exact_fit_tensor(eri))
    res = scf.run_rhf(h, s, n_electrons, _engine_builder(engine), e_nuc=e_nuc)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E


def test_water_sto3g_scf_through_the_general_kernels(engine, water):
    """Pad the basis to n = 96 (> the fragment kernel's 80) with orthonormal functions that carry
    a huge one-electron energy and no two-electron integrals: they are never occupied, the physics
    is unchanged, and every build goes through j_gamma/j_accumulate + the DMMA exchange kernels."""
    s, h, b, e_nuc, n_electrons, e_ref = water
    n0, n = h.shape[0], 96
    s_p, h_p = np.eye(n), 1.0e3 * np.eye(n)
    s_p[:n0, :n0], h_p[:n0, :n0] = s, h
    naux = b.shape[1]
    b_p = np.zeros((n * n, naux), order="F")
    for p in range(naux):
        slab = np.zeros((n, n))
        slab[:n0, :n0] = b[:, p].reshape(n0, n0, order="F")
        b_p[:, p] = slab.reshape(n * n, order="F")
    engine.set_tensor(b_p)
    engine.set_fuse_threshold(0)            # D = 2CC^T every iteration: take gamma from the half-transform too
    try:
        res = scf.run_rhf(h_p, s_p, n_electrons, _engine_builder(engine), e_nuc=e_nuc)
        fused = engine.last_gamma_fused()
    finally:
        engine.set_fuse_threshold(32 << 20)
    assert res["converged"] and abs(res["energy"] - e_ref) < TOL_E and fused
    res2 = scf.run_rhf(h_p, s_p, n_electrons, _engine_builder(engine), e_nuc=e_nuc)
    assert abs(res2["energy"] - e_ref) < TOL_E
