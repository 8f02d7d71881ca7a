"""Committed fixtures (tests/golden/df_fock_golden.npz, made by make_golden.py from the
oracle): the oracle must still reproduce them bit-for-tolerance on CPU, and the CUDA
engine must reproduce them through the C ABI on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg                                    # noqa: E402

from metalquicha_b200 import synth                          # noqa: E402
from oracle import df_fock_oracle as oracle                 # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "df_fock_golden.npz"))


@pytest.mark.parametrize("name", sorted(mg.RHF_CASES))
def test_oracle_reproduces_rhf_fixture(name):
    seed, n, o, q, ks, js = mg.RHF_CASES[name]
    b, h, d, c = synth.synth_problem(seed, n, o, q)
    hh = np.zeros_like(h) if js == 0.0 else h
    f = oracle.build_fock_df(hh, b, d, c, o, k_scale=ks, j_scale=js)
    assert np.max(np.abs(f - GOLD[name + "/F"])) <= 1e-12
    assert GOLD[name + "/F"].shape == (n, n) and b.shape == (n * n, q)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.RHF_CASES))
def test_engine_reproduces_rhf_fixture(engine, name):
    seed, n, o, q, ks, js = mg.RHF_CASES[name]
    b, h, d, c = synth.synth_problem(seed, n, o, q)
    hh = np.zeros_like(h) if js == 0.0 else h
    engine.set_tensor(b)
    f = engine.build_fock_df(hh, d, c, o, k_scale=ks, j_scale=js)
    assert np.max(np.abs(f - GOLD[name + "/F"])) <= 1e-10
    assert abs(engine.last_energy() - float(GOLD[name + "/E"])) <= 1e-9
    j, k = engine.build_jk(d, c, o)
    assert np.max(np.abs(j - GOLD[name + "/J"])) <= 1e-10 and np.max(np.abs(k - GOLD[name + "/K"])) <= 1e-10


@pytest.mark.gpu
def test_engine_reproduces_two_spin_fixture(engine):
    name = "radical_two_spin"
    seed, n, na, nb, q, ks = mg.UHF_CASES[name]
    b = synth.synth_tensor(seed, n, q)
    h = synth.synth_core_hamiltonian(seed, n)
    ca, cb = synth.synth_orbitals(seed, n, na), synth.synth_orbitals(seed + 1, n, nb)
    da, db = oracle.build_density_spin(ca, na), oracle.build_density_spin(cb, nb)
    engine.set_tensor(b)
    j, ka, kb = engine.build_jk_uhf(da + db, ca, na, cb, nb)
    for got, key in ((j, "J"), (ka, "Ka"), (kb, "Kb")):
        assert np.max(np.abs(got - GOLD[f"{name}/{key}"])) <= 1e-10
    fa, fb = engine.build_fock_df_uhf(h, da, db, ca, na, cb, nb, k_scale=ks)
    assert np.max(np.abs(fa - GOLD[name + "/Fa"])) <= 1e-10 and np.max(np.abs(fb - GOLD[name + "/Fb"])) <= 1e-10
