"""Committed fixtures for the operators next to the Fock build (tests/golden/df_neighbours_golden.npz,
made by make_golden_neighbours.py from the oracle): the oracle must still reproduce them on CPU, and
the CUDA engine must reproduce them through the C ABI on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_neighbours as mg                         # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "df_neighbours_golden.npz"))


def _close(a, ref, tol):
    return float(np.max(np.abs(np.asarray(a) - ref))) <= tol * max(1.0, float(np.max(np.abs(ref))))


def test_oracle_reproduces_the_neighbour_fixtures():
    now = mg.compute()
    assert sorted(now) == sorted(GOLD.files)
    for key in GOLD.files:
        assert now[key].shape == GOLD[key].shape, key
        assert _close(now[key], GOLD[key], 1e-12), key


def test_fixture_shapes_follow_the_reference_layouts():
    r, w, g = mg.RESPONSE, mg.WHITEN, mg.GRADIENT
    assert GOLD["response/g"].shape == (r["n"], r["n"])                       # cphf.F90:499-566: g(nao, nao)
    assert GOLD["whiten/b"].shape == (w["n"] ** 2, w["naux"])                 # integrals.F90:985-986: b(nao*nao, naux)
    assert GOLD["whiten/half"].shape == (w["naux"], w["naux"])
    assert GOLD["gradient/gamma"].shape == (g["n"], g["n"], g["naux"])        # gradient.f90:1654: gamma(nao, nao, naux)
    assert GOLD["gradient/omega"].shape == (g["naux"], g["naux"])
    # the dropped modes really are dropped: half . metric . half projects on naux - n_null modes
    _, metric, *_ = mg.whiten_inputs()
    proj = GOLD["whiten/half"] @ metric @ GOLD["whiten/half"]
    assert abs(np.trace(proj) - (w["naux"] - w["n_null"])) <= 1e-8


@pytest.mark.gpu
def test_engine_reproduces_the_response_fixtures(engine):
    b, x, c_occ, dtilde, general = mg.response_inputs()
    engine.set_tensor(b)
    g = engine.response_operator_df(x, c_occ, dtilde, k_scale=mg.RESPONSE["k_scale"])
    assert _close(g, GOLD["response/g"], 1e-10)
    assert _close(engine.fitted_potential_general(general), GOLD["response/g_general"], 1e-10)


@pytest.mark.gpu
def test_engine_reproduces_the_tensor_construction_fixtures(engine):
    three, metric, h, d, coeff = mg.whiten_inputs()
    n, n_occ = mg.WHITEN["n"], mg.WHITEN["n_occ"]
    assert _close(engine.metric_inverse_sqrt(metric), GOLD["whiten/half"], 1e-9)
    assert engine.last_metric_kept == mg.WHITEN["naux"] - mg.WHITEN["n_null"]
    half = engine.build_df_tensor(three, metric, n)                          # metric^(-1/2) and the whitening on the device
    assert _close(half, GOLD["whiten/half"], 1e-9)
    assert _close(engine.build_fock_df(h, d, coeff, n_occ), GOLD["whiten/F"], 1e-10)
    # the same tensor set from the fixture's whitened b gives the same F
    engine.set_tensor(np.asfortranarray(GOLD["whiten/b"]))
    assert _close(engine.build_fock_df(h, d, coeff, n_occ), GOLD["whiten/F"], 1e-10)


@pytest.mark.gpu
def test_engine_reproduces_the_gradient_density_fixtures(engine):
    three, metric, d, orb = mg.gradient_inputs()
    c = mg.GRADIENT
    half = engine.build_df_tensor(three, metric, c["n"])
    gamma, omega = engine.df_gradient_densities(half, d, orb, c["n_occ"], exx_fraction=c["exx"])
    assert _close(gamma, GOLD["gradient/gamma"], 1e-10)
    assert _close(omega, GOLD["gradient/omega"], 1e-10)
