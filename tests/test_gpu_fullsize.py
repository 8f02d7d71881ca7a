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
        engine.set_fuse_threshold(32 << 20)
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


# ---- the other BASELINE shapes: c4 (the ~200-atom def2-SVP shape north_star states its target on)
# and c5 (two-spin), each on a 16-function auxiliary sub-range synthesised on the device and
# regenerated on the host for the oracle.  c4 exercises two N tiles of the half-transform
# (n_occ = 241 -> 2 x 128), the one-valid-k-sub last chunk, and n = 1450 = 90.6 tiles.
@pytest.mark.parametrize("q_begin", [0, 3001])
def test_c4_shape_subrange_parity(engine, q_begin):
    cfg = synth.CONFIGS["c4"]
    n, n_occ, naux, qs = cfg["n"], cfg["n_occ"], cfg["naux"], 16
    scale = synth.default_scale(n, naux)
    _, h, d, c = synth.synth_problem(SEED + 4, n, n_occ, naux, with_tensor=False)
    b = synth.synth_tensor(SEED + 4, n, naux, scale, q_begin=q_begin, q_count=qs)
    engine.synth_tensor(n, naux, SEED + 4, scale, q_begin=q_begin, q_count=qs)
    j_ref, k_ref, _ = oracle.jk_df_fast(b, d, c, n_occ)
    f_ref = h + j_ref - 0.5 * 0.2 * k_ref                      # B3LYP: k_scale = 0.2
    f = engine.build_fock_df(h, d, c, n_occ, k_scale=0.2)
    assert engine.last_gamma_fused()                            # D = 2CC^T, shard above the fuse threshold
    assert np.max(np.abs(f - f_ref)) <= 1e-10
    assert abs(engine.last_energy() - oracle.electronic_energy(h, f_ref, d)) <= 1e-9
    j, k = engine.build_jk(d, c, n_occ)
    assert np.max(np.abs(j - j_ref)) <= 1e-10 and np.max(np.abs(k - k_ref)) <= 1e-10
    # one stream and two streams give the same bits; so does the general (unfused) Coulomb pass to 1e-11
    engine.set_overlap(False)
    try:
        assert np.array_equal(engine.build_fock_df(h, d, c, n_occ, k_scale=0.2), f)
    finally:
        engine.set_overlap(True)
    engine.set_fuse_threshold(2 ** 62)
    try:
        f_general = engine.build_fock_df(h, d, c, n_occ, k_scale=0.2)
        assert not engine.last_gamma_fused() and np.max(np.abs(f_general - f)) <= 1e-11
    finally:
        engine.set_fuse_threshold(32 << 20)


def test_c5_shape_two_spin_subrange_parity(engine):
    cfg = synth.CONFIGS["c5"]
    n, na, nb, naux, qs = cfg["n"], cfg["n_alpha"], cfg["n_beta"], cfg["naux"], 16
    scale = synth.default_scale(n, naux)
    ca = synth.synth_orbitals(SEED + 5, n, na)
    cb = synth.synth_orbitals(SEED + 6, n, nb)
    h = synth.synth_core_hamiltonian(SEED + 5, n)
    da, db = oracle.build_density_spin(ca, na), oracle.build_density_spin(cb, nb)
    b = synth.synth_tensor(SEED + 5, n, naux, scale, q_begin=100, q_count=qs)
    engine.synth_tensor(n, naux, SEED + 5, scale, q_begin=100, q_count=qs)
    j_ref, ka_ref, _ = oracle.jk_df_fast(b, da + db, ca, na, rhf_factor=1.0)
    _, kb_ref, _ = oracle.jk_df_fast(b, da + db, cb, nb, rhf_factor=1.0)
    j, ka, kb = engine.build_jk_uhf(da + db, ca, na, cb, nb)
    assert max(np.max(np.abs(j - j_ref)), np.max(np.abs(ka - ka_ref)), np.max(np.abs(kb - kb_ref))) <= 1e-10
    fa, fb = engine.build_fock_df_uhf(h, da, db, ca, na, cb, nb)
    assert np.max(np.abs(fa - (h + j_ref - ka_ref))) <= 1e-10 and np.max(np.abs(fb - (h + j_ref - kb_ref))) <= 1e-10
    e = oracle.uhf_electronic_energy(h, fa, fb, da, db)
    assert abs(e - oracle.uhf_electronic_energy(h, h + j_ref - ka_ref, h + j_ref - kb_ref, da, db)) <= 1e-9


def test_c2_full_auxiliary_range_against_the_oracle(engine, c2):
    """All 1800 auxiliary functions of BASELINE configs[1], not a sub-range: the oracle's J and K are
    accumulated over eight 225-function chunks regenerated on the host (J and K are sums over the
    auxiliary index), the engine builds from the full 3.5 GB resident tensor in one go."""
    scale, h, d, c = c2
    engine.synth_tensor(N, NAUX, SEED, scale)
    j, k = engine.build_jk(d, c, NOCC)
    f = engine.build_fock_df(h, d, c, NOCC)
    j_ref, k_ref = np.zeros((N, N)), np.zeros((N, N))
    for q0 in range(0, NAUX, 225):
        b = synth.synth_tensor(SEED, N, NAUX, scale, q_begin=q0, q_count=225)
        jq, kq, _ = oracle.jk_df_fast(b, d, c, NOCC)
        j_ref += jq
        k_ref += kq
    assert np.max(np.abs(j - j_ref)) <= 1e-10 and np.max(np.abs(k - k_ref)) <= 1e-10
    f_ref = h + j_ref - 0.5 * k_ref
    assert np.max(np.abs(f - f_ref)) <= 1e-10
    assert abs(engine.last_energy() - oracle.electronic_energy(h, f_ref, d)) <= 1e-9
