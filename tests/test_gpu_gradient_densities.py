"""The DF gradient densities built on the device from the resident whitened tensor
vs the oracle's restatement of df_two_electron_gradient / add_exchange_channel
(mqc_libcint_gradient.f90:1545-1812) on the un-whitened integrals.  Tolerance 1e-10 relative to the
largest element."""
import numpy as np
import pytest

from metalquicha_b200 import synth
from metalquicha_b200.engine import metric_inverse_sqrt
from oracle import df_fock_oracle as oracle
from oracle import df_gradient_oracle as grad

pytestmark = pytest.mark.gpu


def _problem(seed, n, naux, n_null=0):
    three, metric = synth.synth_physical_like_tensor(seed, n, naux, n_null=n_null)
    half = metric_inverse_sqrt(metric)
    return three, metric, half


def _close(a, b, tol=1e-10):
    return float(np.max(np.abs(a - b))) <= tol * max(1.0, float(np.max(np.abs(b))))


@pytest.mark.parametrize("n,naux,n_occ,n_null", [(24, 40, 5, 0), (37, 70, 9, 2), (70, 130, 17, 0), (90, 65, 33, 1)])
def test_closed_shell_gradient_densities(engine, n, naux, n_occ, n_null):
    three, metric, half = _problem(10 + n, n, naux, n_null)
    c = synth.synth_orbitals(n, n, n_occ)
    d = np.asfortranarray(2.0 * c @ c.T)
    engine.set_tensor_from_3c(three, half, n)
    for kf, coul in ((None, True), (0.2, True), (0.0, True), (1.0, False)):
        gamma, omega = engine.df_gradient_densities(half, d, c, n_occ, exx_fraction=kf, with_coulomb=coul)
        g_ref, o_ref, rho, gvec = grad.df_gradient_densities(three, metric, d, c, n_occ, exx_fraction=kf, with_coulomb=coul)
        assert _close(gamma, g_ref) and _close(omega, o_ref)
        assert np.array_equal(omega, omega.T) or _close(omega, omega.T, 1e-14)
    # the Coulomb energy these densities differentiate: 1/2 g^T J^-1 g = 1/2 tr(D J)
    _, _, rho, gvec = grad.df_gradient_densities(three, metric, d, c, n_occ)
    j, _ = engine.build_jk(d, c, n_occ, want_k=False)
    assert abs(0.5 * float(rho @ gvec) - 0.5 * float(np.sum(d * j))) <= 1e-9 * max(1.0, abs(float(rho @ gvec)))


def test_unrestricted_form_and_its_closed_shell_limit(engine):
    n, naux, na, nb = 44, 80, 8, 6
    three, metric, half = _problem(3, n, naux)
    ca, cb = synth.synth_orbitals(1, n, na), synth.synth_orbitals(2, n, nb)
    d = np.asfortranarray(ca @ ca.T + cb @ cb.T)
    engine.set_tensor_from_3c(three, half, n)
    gamma, omega = engine.df_gradient_densities(half, d, ca, na, orbitals_beta=cb, n_occupied_beta=nb, exx_fraction=0.5)
    g_ref, o_ref, _, _ = grad.df_gradient_densities(three, metric, d, ca, na, orbitals_beta=cb, n_occupied_beta=nb,
                                                    exx_fraction=0.5)
    assert _close(gamma, g_ref) and _close(omega, o_ref)
    # a closed shell fed through the unrestricted path gives the restricted densities (gradient.f90:1753-1758)
    d2 = np.asfortranarray(2.0 * ca @ ca.T)
    gu, ou = engine.df_gradient_densities(half, d2, ca, na, orbitals_beta=ca, n_occupied_beta=na)
    gr, orr = engine.df_gradient_densities(half, d2, ca, na)
    assert _close(gu, gr) and _close(ou, orr)
    # an empty beta channel is skipped
    g1, o1 = engine.df_gradient_densities(half, np.asfortranarray(ca @ ca.T), ca, na, orbitals_beta=cb, n_occupied_beta=0)
    g1_ref, o1_ref, _, _ = grad.df_gradient_densities(three, metric, ca @ ca.T, ca, na, orbitals_beta=cb, n_occupied_beta=0)
    assert _close(g1, g1_ref) and _close(o1, o1_ref)


def test_general_batched_gemm_shapes_through_the_gradient_path(engine):
    """Ragged everything: n, naux and n_occ off every tile edge of the batched GEMM (64) and its K step (16)."""
    n, naux, n_occ = 67, 129, 13
    three, metric, half = _problem(8, n, naux)
    c = synth.synth_orbitals(5, n, n_occ)
    d = np.asfortranarray(2.0 * c @ c.T)
    engine.set_tensor_from_3c(three, half, n)
    gamma, omega = engine.df_gradient_densities(half, d, c, n_occ)
    g_ref, o_ref, _, _ = grad.df_gradient_densities(three, metric, d, c, n_occ)
    assert _close(gamma, g_ref) and _close(omega, o_ref)
    pad = np.zeros((n + 5, n_occ + 2), order="F"); pad[:n, :n_occ] = c
    g2, o2 = engine.df_gradient_densities(half, d, pad[:n, :], n_occ)
    assert np.array_equal(g2, gamma) and np.array_equal(o2, omega)


def test_nothing_asked_nothing_built(engine):
    """exx_fraction = 0 with the Coulomb shapes dropped leaves both densities exactly zero
    (gradient.f90:1672, :1691-1695, :1701); a pure functional keeps only the Coulomb shapes."""
    n, naux, n_occ = 30, 44, 6
    three, metric, half = _problem(21, n, naux)
    c = synth.synth_orbitals(21, n, n_occ)
    d = np.asfortranarray(2.0 * c @ c.T)
    engine.set_tensor_from_3c(three, half, n)
    gamma, omega = engine.df_gradient_densities(half, d, c, n_occ, exx_fraction=0.0, with_coulomb=False)
    assert not gamma.any() and not omega.any()
    gamma, omega = engine.df_gradient_densities(half, d, c, n_occ, exx_fraction=0.0)
    g_ref, o_ref, rho, _ = grad.df_gradient_densities(three, metric, d, c, n_occ, exx_fraction=0.0)
    assert _close(gamma, g_ref) and _close(omega, o_ref)
    assert _close(gamma[:, :, 3], rho[3] * d)
