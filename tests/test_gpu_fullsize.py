"""Full-size checks at BASELINE.json's single-GPU configuration, (n, n_occ, naux) =
(688, 80, 1800), where a complete oracle build is too slow for a test: size-independent
properties of the contraction plus oracle parity on an auxiliary sub-range."""
import numpy as np
import pytest

from metalquicha_b200 import synth
from oracle import df_fock_oracle as oracle

pytestmark = pytest.mark.gpu

CFG = synth.CONFIGS["c2"]
N, NOCC, NAUX = CFG["n"], CFG["n_occ"], CFG["naux"]
SEED = 4242


@pytest.fixture(scope="module")
def c2(engine):
    scale = synth.default_scale(N, NAUX)
    engine.synth_tensor(N, NAUX, SEED, scale)
    _, h, d, c = synth.synth_problem(SEED, N, NOCC, NAUX, with_tensor=False)
    return scale, h, d, c


def test_subrange_parity_and_additivity_over_the_auxiliary_index(engine, c2):
    scale, h, d, c = c2
    j_full, k_full = engine.build_jk(d, c, NOCC)
    # oracle on the first 48 auxiliary functions, regenerated on the host
    qs = 48
    b = synth.synth_tensor(SEED, N, NAUX, scale, q_begin=0, q_count=qs)
    engine.synth_tensor(N, NAUX, SEED, scale, q_begin=0, q_count=qs)
    j0, k0 = engine.build_jk(d, c, NOCC)
    j_ref, k_ref, _ = oracle.jk_df_fast(b, d, c, NOCC)
    assert np.max(np.abs(j0 - j_ref)) <= 1e-10 and np.max(np.abs(k0 - k_ref)) <= 1e-10
    # the rest of the range, in two uneven shards: partial J/K add up to the full build
    j_sum, k_sum = j0.copy(), k0.copy()
    for q0, q1 in ((qs, 1000), (1000, NAUX)):
        engine.synth_tensor(N, NAUX, SEED, scale, q_begin=q0, q_count=q1 - q0)
        j, k = engine.build_jk(d, c, NOCC)
        j_sum += j
        k_sum += k
    assert np.max(np.abs(j_sum - j_full)) <= 1e-10 and np.max(np.abs(k_sum - k_full)) <= 1e-10
    engine.synth_tensor(N, NAUX, SEED, scale)


def test_linearity_invariance_and_definiteness(engine, c2):
    scale, h, d, c = c2
    j1, k1 = engine.build_jk(d, c, NOCC)
    rng = np.random.default_rng(1)
    d2 = rng.standard_normal((N, N)); d2 = np.asfortranarray(d2 + d2.T)
    j2, _ = engine.build_jk(d2, c, NOCC, want_k=False)
    j12, _ = engine.build_jk(np.asfortranarray(0.3 * d - 1.7 * d2), c, NOCC, want_k=False)
    assert np.max(np.abs(j12 - (0.3 * j1 - 1.7 * j2))) <= 1e-10 * max(1.0, np.max(np.abs(j2)))
    # K depends on C only through C C^T: rotate the occupied orbitals
    u, _ = np.linalg.qr(rng.standard_normal((NOCC, NOCC)))
    _, k_rot = engine.build_jk(d, np.asfortranarray(c @ u), NOCC, want_j=False)
    assert np.max(np.abs(k_rot - k1)) <= 1e-10 * max(1.0, np.max(np.abs(k1)))
    # K = 2 sum X X^T is positive semi-definite; J and K are exactly symmetric
    assert np.linalg.eigvalsh(k1).min() >= -1e-9
    assert np.array_equal(j1, j1.T) and np.array_equal(k1, k1.T)
    # tr(D J) = sum_Q gamma_Q^2 >= 0
    assert np.sum(d * j1) >= 0.0


def test_fock_energy_and_bit_reproducibility(engine, c2):
    scale, h, d, c = c2
    f1 = engine.build_fock_df(h, d, c, NOCC)
    assert engine.last_gamma_fused()            # 3.5 GB tensor, D = 2CC^T: gamma from the half-transform
    e1 = engine.last_energy()
    try:                                        # ... and the general pass over B gives the same Fock matrix
        engine.set_fuse_threshold(2 ** 62)
        f_general = engine.build_fock_df(h, d, c, NOCC)
        assert not engine.last_gamma_fused()
        assert np.max(np.abs(f_general - f1)) <= 1e-11
    finally:
        engine.set_fuse_threshold(1 << 30)
    j, k = engine.build_jk(d, c, NOCC)
    assert np.max(np.abs(f1 - (h + j - 0.5 * k))) <= 1e-10
    assert abs(e1 - 0.5 * np.sum(d * (h + f1))) <= 1e-9 * max(1.0, abs(e1))
    f2 = engine.build_fock_df(h, d, c, NOCC)
    assert np.array_equal(f1, f2) and engine.last_energy() == e1


def test_device_resident_call_equals_host_call(engine, c2):
    torch = pytest.importorskip("torch")
    scale, h, d, c = c2
    f_host = engine.build_fock_df(h, d, c, NOCC, k_scale=0.2)
    d_h = torch.from_numpy(np.ascontiguousarray(h.T)).cuda()
    d_d = torch.from_numpy(np.ascontiguousarray(d.T)).cuda()
    d_c = torch.from_numpy(np.ascontiguousarray(c.T)).cuda()
    d_f = torch.empty_like(d_h)
    engine.build_fock_device(d_h, d_d, d_c, NOCC, d_f, k_scale=0.2)
    assert np.array_equal(d_f.cpu().numpy().T, f_host)
