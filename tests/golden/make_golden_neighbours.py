"""Generates tests/golden/df_neighbours_golden.npz from the NumPy oracle: the operators next to the
Fock build on the same fitted tensor -- the CPHF response operator and the general-density
potential (mqc_libcint_cphf.F90:499-616), metric^(-1/2) and the whitened tensor with dropped modes
(mqc_libcint_integrals.F90:981-1038), and the DF gradient densities (mqc_libcint_gradient.f90:1545-1812).

Like df_fock_golden.npz these are ORACLE-generated regression pins (the reference cannot run in this
image and holds no element-level vectors for these routines); inputs come back from seeds, only
outputs are stored.

    python tests/golden/make_golden_neighbours.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from metalquicha_b200 import synth                      # noqa: E402
from oracle import df_fock_oracle as oracle             # noqa: E402
from oracle import df_gradient_oracle as grad           # noqa: E402

RESPONSE = dict(seed=61, n=37, n_occ=7, naux=45, k_scale=0.25)
WHITEN = dict(seed=62, n=19, naux=31, n_null=2, n_occ=5)
GRADIENT = dict(seed=63, n=16, naux=24, n_null=1, n_occ=4, exx=0.2)


def response_inputs():
    c = RESPONSE
    b, _, _, coeff_all = synth.synth_problem(c["seed"], c["n"], c["n"], c["naux"])
    c_occ = np.asfortranarray(coeff_all[:, :c["n_occ"]])
    rng = np.random.default_rng(c["seed"])
    x = np.asfortranarray(coeff_all[:, c["n_occ"]:] @ rng.standard_normal((c["n"] - c["n_occ"], c["n_occ"])))
    dtilde = np.asfortranarray(x @ c_occ.T + c_occ @ x.T)
    a = rng.standard_normal((c["n"], c["n"]))
    general = np.asfortranarray(a + a.T)
    return b, x, c_occ, dtilde, general


def whiten_inputs():
    c = WHITEN
    three, metric = synth.synth_physical_like_tensor(c["seed"], c["n"], c["naux"], n_null=c["n_null"])
    _, h, d, coeff = synth.synth_problem(c["seed"], c["n"], c["n_occ"], c["naux"], with_tensor=False)
    return three, metric, h, d, coeff


def gradient_inputs():
    c = GRADIENT
    three, metric = synth.synth_physical_like_tensor(c["seed"], c["n"], c["naux"], n_null=c["n_null"])
    orb = synth.synth_orbitals(c["seed"], c["n"], c["n_occ"])
    d = np.asfortranarray(2.0 * orb @ orb.T)
    return three, metric, d, orb


def compute():
    out = {}
    b, x, c_occ, dtilde, general = response_inputs()
    out["response/g"] = oracle.response_operator_df(b, x, c_occ, dtilde, k_scale=RESPONSE["k_scale"])
    out["response/g_general"] = oracle.fitted_potential_general(b, general)
    three, metric, h, d, coeff = whiten_inputs()
    out["whiten/half"] = oracle.metric_inverse_sqrt(metric)
    bw = oracle.whiten(three, metric)
    out["whiten/b"] = bw
    out["whiten/F"] = oracle.build_fock_df(h, bw, d, coeff, WHITEN["n_occ"])
    three, metric, d, orb = gradient_inputs()
    gamma, omega, rho, gvec = grad.df_gradient_densities(three, metric, d, orb, GRADIENT["n_occ"], exx_fraction=GRADIENT["exx"])
    out["gradient/gamma"], out["gradient/omega"], out["gradient/rho"], out["gradient/g"] = gamma, omega, rho, gvec
    return out


def main():
    out = compute()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "df_neighbours_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
