"""The DF gradient densities against the energy they differentiate, on real integrals (CPU).

``df_two_electron_gradient`` (mqc_libcint_gradient.f90:1545-1746) turns the geometry derivative of the fitted
two-electron energy -- at fixed density and orbitals, the SCF being stationary -- into two densities,

    dE_2e/dx = sum_{uvP} Gamma^P_uv d(uv|P)/dx + sum_PQ Omega_PQ d(P|Q)/dx,

and the engine forms Gamma and Omega on the device (tests/test_gpu_gradient_densities.py compares it with
``oracle/df_gradient_oracle.py``).  The reference checks this family of routines against finite differences
(validation/check_fitted_reference_gradient.f90, check_df_ref_gradient.f90: central differences, the fitted tensor
rebuilt at every displaced geometry, exchange fraction 1 and 0.5, and the sum over atoms of the analytic gradient
must vanish).  Same check here for the oracle's Gamma/Omega, on the reference's own water / 6-31G* fitted with
6-31G* case: the energy side is the oracle's ``jk_df`` on ``three(R) . metric(R)^-1/2`` -- the path that is pinned
to the reference-held energy -- and the integral derivatives are central differences of the restated integrals.
A wrong channel weight, a missing 1/2 on the metric term or a sign cannot survive it.
"""
import numpy as np
import pytest

from oracle import df_fock_oracle as oracle
from oracle import df_gradient_oracle as grad
from oracle import gto_integrals as gto
from oracle import scf_oracle as scf

STEP = 2.0e-4        # bohr; central differences on both sides, O(h^2) apart


def _integrals(coords):
    (symbols, _, _, _), table = gto.DF_CASES["h2o_631gs"]
    basis = gto.build_basis(symbols, coords, table)
    aux = gto.build_basis(symbols, coords, table)
    n, naux = len(basis), len(aux)
    three = np.asfortranarray(gto.three_centre(basis, aux).reshape(n * n, naux, order="F"))
    return three, gto.two_centre(aux)


def _two_electron_energy(three, metric, density, coeff, n_occ, kf):
    """E_2e = 1/2 sum D (J - kf/2 K) from the fitted tensor b = three . metric^-1/2 (rhf.f90:1620-1645)."""
    b = oracle.whiten(three, metric)
    j, k, _ = oracle.jk_df(b, density, coeff, n_occ)
    return 0.5 * float(np.sum(density * (j - 0.5 * kf * k)))


@pytest.fixture(scope="module")
def converged_water():
    (symbols, coords, n_electrons, e_ref), _ = gto.DF_CASES["h2o_631gs"]
    s, h, three, metric, e_nuc, _, _ = gto.df_case_integrals("h2o_631gs")
    b = oracle.whiten(three, metric)

    def builder(h_, density, coeff, n_occ):
        f = oracle.build_fock_df(h_, b, density, coeff, n_occ)
        return f, oracle.electronic_energy(h_, f, density)
    res = scf.run_rhf(h, s, n_electrons, builder, e_nuc=e_nuc, energy_tol=1e-12, density_tol=1e-10)
    assert res["converged"] and abs(res["energy"] - e_ref) < 1e-9
    coords = np.array(coords, dtype=float)
    plus_minus = {}
    for atom in range(3):
        for comp in range(3):
            moved = []
            for sign in (+1.0, -1.0):
                c = coords.copy()
                c[atom, comp] += sign * STEP
                moved.append(_integrals(c.tolist()))
            plus_minus[(atom, comp)] = moved
    return three, metric, res["density"], res["orbitals"], n_electrons // 2, plus_minus


@pytest.mark.parametrize("kf,with_coulomb", [(1.0, True), (0.5, True), (0.0, True), (1.0, False)])
def test_gamma_and_omega_are_the_derivative_of_the_fitted_energy(converged_water, kf, with_coulomb):
    three, metric, density, coeff, n_occ, plus_minus = converged_water
    n, naux = density.shape[0], metric.shape[0]
    gamma, omega, rho, g = grad.df_gradient_densities(three, metric, density, coeff, n_occ, exx_fraction=kf,
                                                     with_coulomb=with_coulomb)
    assert gamma.shape == (n, n, naux) and omega.shape == (naux, naux)            # gradient.f90:1654-1655
    g3 = gamma.reshape(n * n, naux, order="F")
    analytic = np.zeros((3, 3))
    numeric = np.zeros((3, 3))
    for (atom, comp), ((three_p, metric_p), (three_m, metric_m)) in plus_minus.items():
        d_three = (three_p - three_m) / (2.0 * STEP)
        d_metric = (metric_p - metric_m) / (2.0 * STEP)
        analytic[atom, comp] = float(np.sum(g3 * d_three) + np.sum(omega * d_metric))
        e_p = _two_electron_energy(three_p, metric_p, density, coeff, n_occ, kf)
        e_m = _two_electron_energy(three_m, metric_m, density, coeff, n_occ, kf)
        if not with_coulomb:                                                       # the exchange part alone
            e_p -= _two_electron_energy(three_p, metric_p, density, coeff, n_occ, 0.0)
            e_m -= _two_electron_energy(three_m, metric_m, density, coeff, n_occ, 0.0)
        numeric[atom, comp] = (e_p - e_m) / (2.0 * STEP)
    scale = float(np.max(np.abs(numeric)))
    assert scale > 1e-2                                                            # there is a force to get right
    assert float(np.max(np.abs(analytic - numeric))) <= 2e-8 * max(1.0, scale), (analytic, numeric)
    # no net force: the sum over atoms vanishes (check_fitted_reference_gradient.f90:168), to the accuracy of the differences
    assert float(np.max(np.abs(analytic.sum(axis=0)))) <= 1e-7 * max(1.0, scale)


def test_closed_shell_through_the_unrestricted_path_gives_the_same_densities(converged_water):
    """gradient.f90:1753-1758: the check the reference itself names for the channel weights."""
    three, metric, density, coeff, n_occ, _ = converged_water
    g_r, o_r, _, _ = grad.df_gradient_densities(three, metric, density, coeff, n_occ, exx_fraction=0.7)
    g_u, o_u, _, _ = grad.df_gradient_densities(three, metric, density, coeff, n_occ, orbitals_beta=coeff,
                                                n_occupied_beta=n_occ, exx_fraction=0.7)
    assert float(np.max(np.abs(g_r - g_u))) <= 1e-12 * max(1.0, float(np.max(np.abs(g_r))))
    assert float(np.max(np.abs(o_r - o_u))) <= 1e-12 * max(1.0, float(np.max(np.abs(o_r))))
