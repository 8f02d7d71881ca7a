"""CPU oracle for the contraction half of the density-fitted two-electron gradient --
TEST INFRASTRUCTURE ONLY (same rules as oracle/df_fock_oracle.py).

NumPy restatement, loop for loop, of the part of ``df_two_electron_gradient`` and
``add_exchange_channel`` (backends/libcint/mqc_libcint_gradient.f90:1545-1812) that lies between the
energy-side integrals and the derivative integrals: the two densities ``gamma(nao, nao, naux)`` and
``omega(naux, naux)`` every derivative integral is then contracted with.  Inputs are what the
reference has at that point: the UN-whitened ``three(nao*nao, naux)`` and the ``metric``.
The reference's tests hold no gamma/omega element; what pins this restatement is the check the reference
itself applies to this family of routines (validation/check_fitted_reference_gradient.f90, check_df_ref_gradient.f90):
finite differences of the energy it differentiates.  On the reference's own water / 6-31G* fitted with 6-31G* case
(whose fitted energy the oracle reproduces to 2e-12 Eh) sum Gamma d(uv|P)/dx + sum Omega d(P|Q)/dx equals the central
difference of E_2e = 1/2 sum D (J - k/2 K) to 4e-10 on forces of 2 Eh/bohr, for exchange fractions 1, 0.5, 0 and for
the exchange part alone, and the sum over atoms vanishes (tests/test_gradient_densities_finite_difference.py);
plus the reference's own identity that a closed-shell system through the unrestricted path gives the same
densities (:1753-1758).
"""
from __future__ import annotations

import numpy as np

from .df_fock_oracle import metric_inverse_sqrt


def add_exchange_channel(three, jinv, orbitals, n_occ, weight, gamma, omega):
    """mqc_libcint_gradient.f90:1748-1812, in place on ``gamma`` and ``omega``."""
    nao = orbitals.shape[0]
    naux = jinv.shape[0]
    if n_occ <= 0:                                                      # :1772
        return
    c_occ = np.ascontiguousarray(orbitals[:, :n_occ])                   # :1774
    e = np.zeros((n_occ, n_occ, naux))
    for ip in range(naux):                                              # :1779-1785  e^P = C^T (uv|P) C
        block = three[:, ip].reshape((nao, nao), order="F")
        tmp = block @ c_occ
        e[:, :, ip] = c_occ.T @ tmp
    f = np.zeros((n_occ, n_occ, naux))
    for ip in range(naux):                                              # :1788-1794  f = J^-1 e
        for iq in range(naux):
            if jinv[ip, iq] == 0.0:
                continue
            f[:, :, ip] = f[:, :, ip] + jinv[ip, iq] * e[:, :, iq]
    for ip in range(naux):                                              # :1797-1803  Z^P = C f^P C^T
        tmp = c_occ @ f[:, :, ip]
        zp = tmp @ c_occ.T
        gamma[:, :, ip] = gamma[:, :, ip] - weight * zp
    ff = f.reshape(n_occ * n_occ, naux)
    omega += 0.5 * weight * (ff.T @ ff)                                 # :1806-1811  sum_ij f^P_ij f^Q_ij


def df_gradient_densities(three, metric, total_density, orbitals, n_occupied, orbitals_beta=None,
                          n_occupied_beta=0, exx_fraction=None, with_coulomb=True):
    """``gamma`` and ``omega`` of df_two_electron_gradient (:1654-1719)."""
    nao = total_density.shape[0]
    naux = three.shape[1]
    unrestricted = orbitals_beta is not None                            # :1632
    half = metric_inverse_sqrt(metric)                                  # :1659
    jinv = half @ half                                                  # :1663-1665
    g = np.empty(naux)
    for ip in range(naux):                                              # :1668-1670
        g[ip] = np.sum(three[:, ip].reshape((nao, nao), order="F") * total_density)
    rho = jinv @ g                                                      # :1671
    if not with_coulomb:
        rho = np.zeros(naux)                                            # :1672
    gamma = np.empty((nao, nao, naux))
    for ip in range(naux):                                              # :1675-1677
        gamma[:, :, ip] = rho[ip] * total_density
    kf = 1.0 if exx_fraction is None else exx_fraction                  # :1685-1686
    omega = -0.5 * np.outer(rho, rho) if with_coulomb else np.zeros((naux, naux))     # :1691-1695
    if kf != 0.0:                                                       # :1701
        if unrestricted:
            add_exchange_channel(three, jinv, orbitals, n_occupied, kf, gamma, omega)
            add_exchange_channel(three, jinv, orbitals_beta, n_occupied_beta, kf, gamma, omega)
        else:
            add_exchange_channel(three, jinv, orbitals, n_occupied, 2.0 * kf, gamma, omega)
    return gamma, omega, rho, g
