"""Generates tests/golden/df_fock_golden.npz from the NumPy oracle.

The reference itself cannot run in this image (Fortran, no compiler, no integral
library, no basis data), and its own tests hold no J/K/F element-level vectors, so
these fixtures are ORACLE-generated: they pin the oracle (and through it the CUDA
engine) against regressions, they do not add an independent reference pin.
Inputs are regenerated from seeds by metalquicha_b200.synth; only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from metalquicha_b200 import synth                      # noqa: E402
from oracle import df_fock_oracle as oracle             # noqa: E402

# name -> (seed, n, n_occ, naux, k_scale, j_scale)
RHF_CASES = {
    "water_ccpvdz_jkfit": (11, 24, 5, 116, None, None),        # BASELINE configs[0] shape
    "water_631gs_shape_pin": (12, 19, 5, 19, None, None),      # test_mqc_libcint_cartesian.f90:265-274: B is (361, 19)
    "trimer_def2svp": (13, 72, 15, 340, None, None),           # (H2O)3 fragment of configs[2]
    "hybrid_k_scale": (14, 33, 7, 40, 0.2, None),              # B3LYP-like exact-exchange fraction (configs[3])
    "attenuated_pass": (15, 33, 7, 40, 0.37, 0.0),             # rhf.f90:1100 second pass
}
UHF_CASES = {"radical_two_spin": (21, 43, 6, 5, 60, 1.0)}     # seed, n, n_alpha, n_beta, naux, k_scale


def main():
    out = {}
    for name, (seed, n, o, q, ks, js) in RHF_CASES.items():
        b, h, d, c = synth.synth_problem(seed, n, o, q)
        hh = np.zeros_like(h) if js == 0.0 else h
        j, k, cvec = oracle.jk_df(b, d, c, o)
        f = oracle.build_fock_df(hh, b, d, c, o, k_scale=ks, j_scale=js)
        out[name + "/J"], out[name + "/K"], out[name + "/c"], out[name + "/F"] = j, k, cvec, f
        out[name + "/E"] = np.array(oracle.electronic_energy(hh, f, d))
    for name, (seed, n, na, nb, q, ks) in UHF_CASES.items():
        b = synth.synth_tensor(seed, n, q)
        h = synth.synth_core_hamiltonian(seed, n)
        ca, cb = synth.synth_orbitals(seed, n, na), synth.synth_orbitals(seed + 1, n, nb)
        da, db = oracle.build_density_spin(ca, na), oracle.build_density_spin(cb, nb)
        j, ka, kb = oracle.jk_df_uhf(b, da + db, ca, na, cb, nb)
        fa, fb = oracle.build_fock_df_uhf(h, b, da, db, ca, na, cb, nb, k_scale=ks)
        out[name + "/J"], out[name + "/Ka"], out[name + "/Kb"] = j, ka, kb
        out[name + "/Fa"], out[name + "/Fb"] = fa, fb
        out[name + "/E"] = np.array(oracle.uhf_electronic_energy(h, fa, fb, da, db))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "df_fock_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
