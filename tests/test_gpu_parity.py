"""Parity of the CUDA engine (through the C ABI) with the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states: J/K/F agree to 1e-10
max-abs, energies to 1e-9 Eh.  Every test runs on seeded inputs small enough for
the oracle to finish in seconds; full-size properties live in test_gpu_fullsize.py.
"""
import numpy as np
import pytest

from metalquicha_b200 import B200Error, synth
from metalquicha_b200.engine import SLOT_ATTENUATED
from oracle import df_fock_oracle as oracle

pytestmark = pytest.mark.gpu

TOL = 1e-10
TOL_E = 1e-9


def _maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


# (n, n_occ, naux): the reference's own water cases, the MBE fragment shapes, ragged n
# (not a multiple of the 16-wide tile or of the 128-wide panel), and occupied counts that
# exercise every half-transform tile width (n_occ -> NB 1..8 and two N tiles).
SHAPES = [
    (24, 5, 116), (24, 5, 139), (48, 10, 227), (72, 15, 340),
    (19, 5, 19), (1, 1, 3), (17, 3, 7), (33, 16, 40), (100, 17, 50), (130, 33, 64),
    (145, 48, 70), (200, 50, 96), (129, 65, 33), (160, 81, 40), (136, 97, 24),
    (150, 113, 20), (257, 128, 12), (140, 130, 10), (144, 144, 9),
]


@pytest.mark.parametrize("n,n_occ,naux", SHAPES)
def test_build_fock_df_matches_oracle(engine, n, n_occ, naux):
    b, h, density, coeff = synth.synth_problem(100 + n, n, n_occ, naux)
    engine.set_tensor(b)
    fock = engine.build_fock_df(h, density, coeff, n_occ)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ)
    assert _maxabs(fock, ref) <= TOL
    assert abs(engine.last_energy() - oracle.electronic_energy(h, ref, density)) <= TOL_E
    assert engine.last_launches() > 0


@pytest.mark.parametrize("n,n_occ,naux", [(24, 5, 116), (130, 33, 97), (256, 241, 12), (140, 130, 10), (688, 80, 30), (300, 120, 80),
                                         (160, 81, 40), (223, 50, 33)])
def test_half_transform_schedule_does_not_change_the_bits(engine, monkeypatch, n, n_occ, naux):
    """The half-transform cuts its last round into pieces and skips an all-padding last n8-block, and the
    accumulation deals the live 8-row blocks of a last-panel-row tile (2, 4 or 6 of 8: n = 130, 160, 300
    here) evenly to its warps (csrc/k_kernels.cu); all three are schedules of the same DMMA sequences, so
    F must not move by a bit when any of them is switched off."""
    b, h, density, coeff = synth.synth_problem(900 + n, n, n_occ, naux)
    engine.set_tensor(b)
    fock = engine.build_fock_df(h, density, coeff, n_occ)
    assert _maxabs(fock, oracle.build_fock_df(h, b, density, coeff, n_occ)) <= TOL
    for switch in ("MQCB200_NO_TAIL_SPLIT", "MQCB200_NO_TRIM", "MQCB200_NO_EDGE_TILES"):
        monkeypatch.setenv(switch, "1")
        assert np.array_equal(engine.build_fock_df(h, density, coeff, n_occ), fock), switch
        monkeypatch.delenv(switch)


@pytest.mark.parametrize("n,n_occ,naux", [(130, 33, 97), (160, 81, 40), (300, 120, 80), (688, 80, 30), (100, 17, 50)])
@pytest.mark.parametrize("splits", [None, "8", "3"])
def test_accumulation_split_classes(engine, monkeypatch, n, n_occ, naux, splits):
    """The accumulation's units come in up to three classes with their own split of the auxiliary range
    (off-diagonal tiles, diagonal tiles at 9/16, tiles of a partly filled last panel row at LB/8:
    csrc/common.cuh k_unit_decode, csrc/k_kernels.cu plan_k).  Forced on and off, with the planner's and with
    forced split counts: every schedule must give F within the tolerance (a different, fixed summation order
    each), and each one reproducibly."""
    b, h, density, coeff = synth.synth_problem(300 + n, n, n_occ, naux)
    engine.set_tensor(b)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ)
    if splits:
        monkeypatch.setenv("MQCB200_KSPLITS", splits)
    monkeypatch.setenv("MQCB200_EDGE_SPLITS", "1")
    f_edge = engine.build_fock_df(h, density, coeff, n_occ)
    assert _maxabs(f_edge, ref) <= TOL
    assert np.array_equal(f_edge, engine.build_fock_df(h, density, coeff, n_occ))
    monkeypatch.setenv("MQCB200_NO_EDGE_TILES", "1")      # same plan, the per-warp row-block skip instead of the even deal
    assert np.array_equal(f_edge, engine.build_fock_df(h, density, coeff, n_occ))
    monkeypatch.delenv("MQCB200_NO_EDGE_TILES")
    monkeypatch.delenv("MQCB200_EDGE_SPLITS")
    monkeypatch.setenv("MQCB200_NO_EDGE_SPLITS", "1")
    f_two = engine.build_fock_df(h, density, coeff, n_occ)
    assert _maxabs(f_two, ref) <= TOL and _maxabs(f_two, f_edge) <= 1e-11


@pytest.mark.parametrize("n,n_occ,naux", [(24, 5, 116), (72, 15, 340), (130, 33, 64), (200, 50, 96)])
def test_jk_match_oracle(engine, n, n_occ, naux):
    b, h, density, coeff = synth.synth_problem(7, n, n_occ, naux)
    engine.set_tensor(b)
    j, k = engine.build_jk(density, coeff, n_occ)
    j_ref, k_ref, _ = oracle.jk_df(b, density, coeff, n_occ)
    assert _maxabs(j, j_ref) <= TOL
    assert _maxabs(k, k_ref) <= TOL
    # exact symmetry of the engine's outputs (it mirrors the lower triangle)
    assert np.array_equal(j, j.T) and np.array_equal(k, k.T)
    j_only, none_k = engine.build_jk(density, coeff, n_occ, want_k=False)
    assert none_k is None and np.array_equal(j_only, j)
    none_j, k_only = engine.build_jk(None, coeff, n_occ, want_j=False)
    assert none_j is None and np.array_equal(k_only, k)


def test_scales_and_attenuated_second_pass(engine):
    """k_scale / j_scale and the range-separated second call (rhf.f90:1094-1105)."""
    n, n_occ, naux = 72, 15, 200
    b, h, density, coeff = synth.synth_problem(11, n, n_occ, naux)
    b_lr = synth.synth_tensor(12, n, naux)
    engine.set_tensor(b)
    engine.set_tensor(b_lr, slot=SLOT_ATTENUATED)
    f = engine.build_fock_df(h, density, coeff, n_occ, k_scale=0.2)
    assert _maxabs(f, oracle.build_fock_df(h, b, density, coeff, n_occ, k_scale=0.2)) <= TOL
    k_lr = engine.build_fock_df(np.zeros_like(h), density, coeff, n_occ, k_scale=0.37, j_scale=0.0,
                                slot=SLOT_ATTENUATED)
    assert _maxabs(k_lr, oracle.build_fock_df(np.zeros_like(h), b_lr, density, coeff, n_occ,
                                              k_scale=0.37, j_scale=0.0)) <= TOL
    fock, e = engine.assemble_fock(h, density, coeff, n_occ, k_scale=0.2, rs_k_lr=0.37)
    fock_ref, e_ref = oracle.assemble_fock_df(h, b, density, coeff, n_occ, k_scale=0.2,
                                              bmat_lr=b_lr, rs_k_lr=0.37)
    assert _maxabs(fock, fock_ref) <= TOL and abs(e - e_ref) <= TOL_E
    fock, e = engine.assemble_fock(h, density, coeff, n_occ)
    fock_ref, e_ref = oracle.assemble_fock_df(h, b, density, coeff, n_occ)
    assert _maxabs(fock, fock_ref) <= TOL and abs(e - e_ref) <= TOL_E
    engine.clear_tensor(SLOT_ATTENUATED)


def test_coeff_leading_dimension_and_extra_columns(engine):
    """Only coeff(:, 1:n_occ) is read (rhf.f90:1618); coeff has n_mo >= n_occ columns."""
    n, n_occ, naux = 50, 9, 40
    b, h, density, _ = synth.synth_problem(3, n, n_occ, naux)
    full = synth.synth_orbitals(3, n, n)            # all n "molecular orbitals"
    engine.set_tensor(b)
    ref = oracle.build_fock_df(h, b, density, full, n_occ)
    assert _maxabs(engine.build_fock_df(h, density, full, n_occ), ref) <= TOL
    padded = np.zeros((n + 6, n), order="F")
    padded[:n, :] = full
    view = padded[:n, :]                            # column-major view, ldc = n + 6
    assert _maxabs(engine.build_fock_df(h, density, view, n_occ), ref) <= TOL
    assert _maxabs(engine.build_fock_df(h, density, np.ascontiguousarray(full), n_occ), ref) <= TOL


def test_strided_views_of_h_and_density_are_copied(engine):
    """Only the coefficient matrix carries a leading dimension through the C ABI; a column-major
    VIEW of h or D (big[:n, :n]) must not be handed over as if it were contiguous."""
    n, n_occ, naux = 40, 7, 30
    b, h, density, coeff = synth.synth_problem(23, n, n_occ, naux)
    engine.set_tensor(b)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ)
    big_h = np.asfortranarray(np.full((n + 5, n + 3), 7.0)); big_h[:n, :n] = h
    big_d = np.asfortranarray(np.full((n + 2, n + 9), -3.0)); big_d[:n, :n] = density
    hv, dv = big_h[:n, :n], big_d[:n, :n]
    assert not hv.flags.f_contiguous and not dv.flags.f_contiguous
    assert _maxabs(engine.build_fock_df(hv, dv, coeff, n_occ), ref) <= TOL
    j, k = engine.build_jk(dv, coeff, n_occ)
    j_ref, k_ref, _ = oracle.jk_df(b, density, coeff, n_occ)
    assert _maxabs(j, j_ref) <= TOL and _maxabs(k, k_ref) <= TOL
    # operands of another size than the resident tensor are refused, not read out of bounds
    assert engine.tensor_shape() == (n, naux, 0, naux)
    with pytest.raises(ValueError, match="resident tensor"):
        engine.build_fock_df(np.zeros((n + 1, n + 1)), np.zeros((n + 1, n + 1)), np.zeros((n + 1, 2)), 2)
    with pytest.raises(ValueError, match="resident tensor"):
        engine.build_jk(np.zeros((n - 1, n - 1)), coeff, n_occ)


def test_non_finite_density_never_takes_the_fused_route(engine):
    """A NaN in D whose other elements match 2CC^T must not set the fuse flag (a max() drops
    NaN): the reference propagates it into J and F."""
    n, n_occ, naux = 96, 20, 40
    b, h, density, coeff = synth.synth_problem(29, n, n_occ, naux)
    engine.set_tensor(b)
    engine.set_fuse_threshold(0)
    try:
        engine.build_fock_df(h, density, coeff, n_occ)
        assert engine.last_gamma_fused()
        bad = density.copy(order="F"); bad[3, 5] = np.nan
        f = engine.build_fock_df(h, bad, coeff, n_occ)
        assert not engine.last_gamma_fused() and np.isnan(f).any()
        asym = density.copy(order="F"); asym[7, 2] += 1e-6        # upper != lower: not an orbital product
        f2 = engine.build_fock_df(h, asym, coeff, n_occ)
        assert not engine.last_gamma_fused()
        assert _maxabs(f2, oracle.build_fock_df(h, b, asym, coeff, n_occ)) <= TOL
    finally:
        engine.set_fuse_threshold(32 << 20)


def test_pseudo_orbital_guess_build(engine):
    """atomic_guess_fock: a non-idempotent guess density through its pseudo-orbitals."""
    n, naux = 60, 80
    rng = np.random.default_rng(5)
    a = rng.standard_normal((n, 14))
    occ = rng.uniform(0.1, 2.0, size=14)
    density = np.asfortranarray((a * occ[None, :]) @ a.T)       # PSD, rank 14, not idempotent
    h = synth.synth_core_hamiltonian(5, n)
    b = synth.synth_tensor(5, n, naux)
    engine.set_tensor(b)
    fock = engine.atomic_guess_fock(h, density)
    pseudo, n_modes = oracle.density_pseudo_orbitals(density)
    assert n_modes == 14
    assert _maxabs(fock, oracle.build_fock_df(h, b, density, pseudo, n_modes)) <= 1e-9
    # K through pseudo-orbitals equals the general-density form sum_P B_P D B_P
    _, k = engine.build_jk(density, pseudo, n_modes)
    assert _maxabs(k, oracle.fitted_exchange_general(b, density)) <= 1e-9
    # full-rank guess: n_modes == n (K cost n/n_occ times larger on that one call)
    d_full = np.asfortranarray(density + 0.01 * np.eye(n))
    pseudo, n_modes = oracle.density_pseudo_orbitals(d_full)
    assert n_modes == n
    assert _maxabs(engine.atomic_guess_fock(h, d_full), oracle.build_fock_df(h, b, d_full, pseudo, n)) <= 1e-9


def test_asymmetric_density_is_symmetrised(engine):
    """The reference sums B_P * D over the full square; a packed engine sees (D + D^T)/2
    implicitly -- identical because every B_P is symmetric."""
    n, n_occ, naux = 40, 6, 30
    b, h, density, coeff = synth.synth_problem(21, n, n_occ, naux)
    rng = np.random.default_rng(21)
    d_asym = np.asfortranarray(density + 0.1 * rng.standard_normal((n, n)))
    engine.set_tensor(b)
    j, _ = engine.build_jk(d_asym, coeff, n_occ, want_k=False)
    j_ref, _, _ = oracle.jk_df(b, d_asym, coeff, n_occ)
    assert _maxabs(j, j_ref) <= TOL


def test_two_spin_build(engine):
    """DF-UHF: J from the total density, K per spin without the factor 2."""
    n, na, nb_, naux = 66, 9, 8, 120
    b = synth.synth_tensor(31, n, naux)
    h = synth.synth_core_hamiltonian(31, n)
    ca = synth.synth_orbitals(31, n, na)
    cb = synth.synth_orbitals(32, n, nb_)
    da, db = oracle.build_density_spin(ca, na), oracle.build_density_spin(cb, nb_)
    engine.set_tensor(b)
    j, ka, kb = engine.build_jk_uhf(da + db, ca, na, cb, nb_)
    j_ref, ka_ref, kb_ref = oracle.jk_df_uhf(b, da + db, ca, na, cb, nb_)
    assert _maxabs(j, j_ref) <= TOL and _maxabs(ka, ka_ref) <= TOL and _maxabs(kb, kb_ref) <= TOL
    fa, fb = engine.build_fock_df_uhf(h, da, db, ca, na, cb, nb_, k_scale=0.5)
    fa_ref, fb_ref = oracle.build_fock_df_uhf(h, b, da, db, ca, na, cb, nb_, k_scale=0.5)
    assert _maxabs(fa, fa_ref) <= TOL and _maxabs(fb, fb_ref) <= TOL
    e = oracle.uhf_electronic_energy(h, fa, fb, da, db)
    assert abs(e - oracle.uhf_electronic_energy(h, fa_ref, fb_ref, da, db)) <= TOL_E
    # restricted vs two-spin with C_alpha = C_beta: K_rhf = 2 K_sigma, same J
    _, k_rhf = engine.build_jk(2.0 * da, ca, na)
    assert _maxabs(k_rhf, 2.0 * ka) <= TOL
    # empty beta channel is skipped, not zeroed
    j1, ka1, kb1 = engine.build_jk_uhf(da, ca, na, None, 0)
    assert kb1 is None and _maxabs(ka1, ka_ref) <= TOL
    fa1, fb1 = engine.build_fock_df_uhf(h, da, np.zeros_like(da), ca, na, None, 0)
    fa1_ref, fb1_ref = oracle.build_fock_df_uhf(h, b, da, np.zeros_like(da), ca, na, cb, 0)
    assert _maxabs(fa1, fa1_ref) <= TOL and _maxabs(fb1, fb1_ref) <= TOL


def test_physical_like_tensor_with_dropped_modes(engine):
    """B = three . metric^(-1/2) with two metric modes below the 1e-10 threshold."""
    n, n_occ, naux = 24, 5, 30
    three, metric = synth.synth_physical_like_tensor(1, n, naux, n_null=2)
    b = np.asfortranarray(oracle.whiten(three, metric))
    _, h, density, coeff = synth.synth_problem(1, n, n_occ, naux, with_tensor=False)
    engine.set_tensor(b)
    fock = engine.build_fock_df(h, density, coeff, n_occ)
    scale = max(1.0, float(np.max(np.abs(fock))))
    assert _maxabs(fock, oracle.build_fock_df(h, b, density, coeff, n_occ)) <= TOL * scale


@pytest.mark.parametrize("n,naux,n_null", [(24, 30, 2), (40, 139, 0), (37, 150, 3), (130, 260, 1)])
def test_device_whitening_matches_reference_gemm(engine, n, naux, n_null):
    """mqcb200_set_tensor_from_3c: b = three . metric^(-1/2) formed on the device in the
    packed layout (integrals.F90:981-987) vs pic_gemm(three, half, b) of the oracle."""
    from metalquicha_b200.engine import metric_inverse_sqrt
    three, metric = synth.synth_physical_like_tensor(3, n, naux, n_null=n_null)
    half_ref = oracle.metric_inverse_sqrt(metric)
    half = metric_inverse_sqrt(metric, engine)          # one-sided Jacobi + GEMM on the device
    assert engine.last_metric_kept == naux - n_null
    assert _maxabs(half, half_ref) <= 1e-9 * max(1.0, float(np.max(np.abs(half_ref))))
    assert _maxabs(half, half.T) <= 1e-12 * max(1.0, float(np.max(np.abs(half_ref))))
    b = np.asfortranarray(three @ half)              # the reference's whitening GEMM on these inputs
    n_occ = max(1, n // 5)
    _, h, density, coeff = synth.synth_problem(3, n, n_occ, naux, with_tensor=False)
    engine.set_tensor(b)
    f_host = engine.build_fock_df(h, density, coeff, n_occ)
    engine.set_tensor_from_3c(three, half, n)
    f_dev = engine.build_fock_df(h, density, coeff, n_occ)
    scale = max(1.0, float(np.max(np.abs(f_host))))
    assert _maxabs(f_dev, f_host) <= TOL * scale
    assert _maxabs(f_dev, oracle.build_fock_df(h, b, density, coeff, n_occ)) <= TOL * scale
    half2 = engine.build_df_tensor(three, metric, n)  # metric^(-1/2) AND the GEMM on the device, one call
    assert _maxabs(half2, half_ref) <= 1e-9 * max(1.0, float(np.max(np.abs(half_ref))))
    assert _maxabs(engine.build_fock_df(h, density, coeff, n_occ), f_host) <= 1e-9 * scale
    # the same GEMM fed slab by slab by the caller, slabs in any order, full-tensor and compact blocks
    engine.whiten_begin(n, naux, half)
    edges = list(range(0, n, 16)) + [n]
    blocks = list(zip(edges[:-1], edges[1:]))[::-1]
    three3 = three.reshape(n, n, naux, order="F")               # (mu, nu, P)
    for nu0, nu1 in blocks:
        blk = np.asfortranarray(three3[:, nu0:nu1, :].reshape(n * (nu1 - nu0), naux, order="F"))
        engine.whiten_push(nu0, blk)
    engine.whiten_end()
    assert np.array_equal(engine.build_fock_df(h, density, coeff, n_occ), f_dev)


def test_response_operator_and_general_density_potential(engine):
    """The reference's other two fitted operators on the same resident tensor
    (mqc_libcint_cphf.F90:499-566 and :568-616), SURVEY 8f row 3 -- direct rank-2 kernels."""
    n, n_occ, naux = 58, 9, 70
    b, _, _, coeff_all = synth.synth_problem(41, n, n, naux)           # all n orbitals
    c_occ = np.asfortranarray(coeff_all[:, :n_occ])
    c_vir = coeff_all[:, n_occ:]
    rng = np.random.default_rng(41)
    x = np.asfortranarray(c_vir @ rng.standard_normal((n - n_occ, n_occ)))   # X = C_vir U
    dtilde = np.asfortranarray(x @ c_occ.T + c_occ @ x.T)                    # symmetric, indefinite
    engine.set_tensor(b)
    for ks in (None, 0.25, 0.0):
        g = engine.response_operator_df(x, c_occ, dtilde, k_scale=ks)
        g_ref = oracle.response_operator_df(b, x, c_occ, dtilde, k_scale=ks)
        assert _maxabs(g, g_ref) <= TOL * max(1.0, float(np.max(np.abs(g_ref))))
        assert np.array_equal(g, g.T)
    # the rank-2 form and the general form agree on the same response density
    g_gen = engine.fitted_potential_general(dtilde)
    assert _maxabs(g_gen, oracle.fitted_potential_general(b, dtilde)) <= TOL
    assert _maxabs(g_gen, engine.response_operator_df(x, c_occ, dtilde)) <= TOL
    # an unstructured symmetric density that integrates to zero (MP2-relaxed-like)
    a = rng.standard_normal((n, n)); d = a + a.T; d -= np.trace(d) / n * np.eye(n)
    d = np.asfortranarray(d)
    g = engine.fitted_potential_general(d, k_scale=0.5)
    g_ref = oracle.fitted_potential_general(b, d, k_scale=0.5)
    assert _maxabs(g, g_ref) <= TOL * max(1.0, float(np.max(np.abs(g_ref))))


@pytest.mark.parametrize("n,n_occ,naux", [(58, 9, 70), (130, 33, 40), (97, 16, 30), (200, 81, 12), (72, 15, 64), (150, 129, 9)])
def test_response_operator_keeps_relative_accuracy_for_a_small_trial_vector(engine, monkeypatch, n, n_occ, naux):
    """Near CPHF convergence |X| ~ 1e-8 |C|.  The direct rank-2 form keeps ~16 digits of g there;
    a difference of squares K[X+C] - K[X-C] keeps eight.  Tolerance is
    RELATIVE to |g|.  Shapes: ragged n, n_occ on and off the 16-wide chunk boundary, two N tiles
    of the stacked half-transform (2*ceil16(n_occ) > 128), fragment-sized n (general path forced)."""
    b, _, _, coeff_all = synth.synth_problem(500 + n, n, n, naux)
    c_occ = np.asfortranarray(coeff_all[:, :n_occ])
    rng = np.random.default_rng(n)
    x = np.asfortranarray(1.0e-8 * (coeff_all[:, n_occ:] @ rng.standard_normal((n - n_occ, n_occ))))
    dtilde = np.asfortranarray(x @ c_occ.T + c_occ @ x.T)
    engine.set_tensor(b)
    g = engine.response_operator_df(x, c_occ, dtilde, k_scale=1.0)
    g_ref = oracle.response_operator_df(b, x, c_occ, dtilde, k_scale=1.0)
    scale = float(np.max(np.abs(g_ref)))
    assert 1e-10 < scale < 1e-5
    assert _maxabs(g, g_ref) <= 1e-10 * scale
    assert np.array_equal(g, engine.response_operator_df(x, c_occ, dtilde, k_scale=1.0))      # bit-reproducible
    monkeypatch.setenv("MQCB200_NO_EDGE_TILES", "1")          # the rank-2 mode takes the edge-tile schedule too
    assert np.array_equal(g, engine.response_operator_df(x, c_occ, dtilde, k_scale=1.0))
    monkeypatch.delenv("MQCB200_NO_EDGE_TILES")
    # coefficient matrices with a leading dimension
    pad = np.zeros((n + 3, 2 * n_occ), order="F")
    pad[:n, :n_occ], pad[:n, n_occ:] = x, c_occ
    assert np.array_equal(g, engine.response_operator_df(pad[:n, :n_occ], pad[:n, n_occ:], dtilde, k_scale=1.0))


def test_singular_metric_is_refused():
    from metalquicha_b200.engine import metric_inverse_sqrt
    with pytest.raises(B200Error, match="singular"):
        metric_inverse_sqrt(np.zeros((5, 5)))


@pytest.mark.parametrize("n,n_occ,naux", [(81, 5, 116), (96, 65, 40), (130, 33, 64), (200, 50, 96),
                                          (257, 128, 12), (140, 130, 10), (100, 16, 40)])
def test_coulomb_vector_from_half_transform(engine, n, n_occ, naux):
    """When the density is the orbitals' own (D = 2CC^T) the engine may take gamma_Q from the
    half-transformed tensor instead of a pass over B; forced on here (threshold 0).  Both
    paths must give the oracle's J, and an inconsistent density must fall back by itself.
    (Shapes are above the one-pass fragment kernel's range, n > 80 or n_occ > 64.)"""
    b, h, density, coeff = synth.synth_problem(300 + n, n, n_occ, naux)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ)
    engine.set_tensor(b)
    try:
        engine.set_fuse_threshold(0)
        f_fused = engine.build_fock_df(h, density, coeff, n_occ)
        assert engine.last_gamma_fused()
        assert _maxabs(f_fused, ref) <= TOL
        assert np.array_equal(f_fused, engine.build_fock_df(h, density, coeff, n_occ))    # bit-reproducible
        # a density that is NOT 2CC^T: the device-side check must refuse the shortcut
        d2 = np.asfortranarray(density + 1e-9 * np.eye(n))
        f2 = engine.build_fock_df(h, d2, coeff, n_occ)
        assert not engine.last_gamma_fused()
        assert _maxabs(f2, oracle.build_fock_df(h, b, d2, coeff, n_occ)) <= TOL
        # two-spin: D_total = Ca Ca^T + Cb Cb^T
        nb_ = max(1, n_occ - 1)
        cb = synth.synth_orbitals(900 + n, n, nb_)
        da, db = oracle.build_density_spin(coeff, n_occ), oracle.build_density_spin(cb, nb_)
        j, ka, kb = engine.build_jk_uhf(da + db, coeff, n_occ, cb, nb_)
        assert engine.last_gamma_fused()
        j_ref, ka_ref, kb_ref = oracle.jk_df_uhf(b, da + db, coeff, n_occ, cb, nb_)
        assert _maxabs(j, j_ref) <= TOL and _maxabs(ka, ka_ref) <= TOL and _maxabs(kb, kb_ref) <= TOL
    finally:
        engine.set_fuse_threshold(2 ** 62)
    try:
        f_general = engine.build_fock_df(h, density, coeff, n_occ)
        assert not engine.last_gamma_fused()
        assert _maxabs(f_general, ref) <= TOL and _maxabs(f_general, f_fused) <= 1e-11
    finally:
        engine.set_fuse_threshold(32 << 20)


def test_device_generator_matches_host_generator(engine):
    """mqcb200_synth_tensor == synth.synth_tensor bit for bit (checked through a build)."""
    n, n_occ, naux = 70, 11, 50
    _, h, density, coeff = synth.synth_problem(9, n, n_occ, naux, with_tensor=False)
    scale = synth.default_scale(n, naux)
    b = synth.synth_tensor(1234, n, naux, scale)
    engine.set_tensor(b)
    f_host = engine.build_fock_df(h, density, coeff, n_occ)
    engine.synth_tensor(n, naux, 1234, scale)
    f_dev = engine.build_fock_df(h, density, coeff, n_occ)
    assert np.array_equal(f_host, f_dev)


def test_repeat_builds_are_bit_identical(engine):
    """MBE differences fragment energies, so the reference relies on bit-reproducible
    repeats (backends/cuest/backend/mqc_cuest_context.f90:209-213)."""
    n, n_occ, naux = 150, 40, 130
    b, h, density, coeff = synth.synth_problem(77, n, n_occ, naux)
    engine.set_tensor(b)
    first = engine.build_fock_df(h, density, coeff, n_occ)
    for _ in range(3):
        assert np.array_equal(first, engine.build_fock_df(h, density, coeff, n_occ))


def test_small_workspace_chunks_the_auxiliary_range(engine):
    """A tight scratch limit forces several half-transform/accumulate rounds."""
    n, n_occ, naux = 90, 20, 300
    b, h, density, coeff = synth.synth_problem(8, n, n_occ, naux)
    engine.set_tensor(b)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ)
    engine.set_workspace_limit(1 << 20)       # 1 MiB: a few auxiliary functions per round
    try:
        assert _maxabs(engine.build_fock_df(h, density, coeff, n_occ), ref) <= TOL
    finally:
        engine.set_workspace_limit(4 << 30)
    assert _maxabs(engine.build_fock_df(h, density, coeff, n_occ), ref) <= TOL


def test_n_occ_zero_and_k_scale_zero(engine):
    n, naux = 30, 20
    b, h, density, coeff = synth.synth_problem(4, n, 4, naux)
    engine.set_tensor(b)
    j_ref, _, _ = oracle.jk_df(b, density, coeff, 4)
    assert _maxabs(engine.build_fock_df(h, density, coeff, 0), h + j_ref) <= TOL
    assert _maxabs(engine.build_fock_df(h, density, coeff, 4, k_scale=0.0), h + j_ref) <= TOL
    j, k = engine.build_jk(density, coeff, 0)
    assert _maxabs(j, j_ref) <= TOL and not k.any()


@pytest.mark.parametrize("n,n_occ,naux", [(200, 150, 20), (300, 257, 8), (81, 65, 30), (96, 96, 17), (128, 1, 40)])
def test_wide_occupied_spaces_and_fragment_path_boundary(engine, n, n_occ, naux):
    """Two and three N tiles in the half-transform (n_occ > 128, > 256), shapes just outside
    the one-pass fragment kernel (n = 81, n_occ = 65), n_occ == n and n_occ == 1."""
    b, h, density, coeff = synth.synth_problem(500 + n, n, n_occ, naux)
    engine.set_tensor(b)
    ref = oracle.build_fock_df(h, b, density, coeff, n_occ)
    scale = max(1.0, float(np.max(np.abs(ref))))
    assert _maxabs(engine.build_fock_df(h, density, coeff, n_occ), ref) <= TOL * scale


def test_empty_shard_contributes_nothing(engine):
    """A rank that holds no auxiliary function (more GPUs than naux slabs) returns H."""
    n, n_occ, naux = 30, 4, 5
    b, h, density, coeff = synth.synth_problem(6, n, n_occ, naux)
    engine.set_tensor_shard(b[:, 0:0], n, naux, 3)
    assert engine.tensor_bytes() == 0
    assert np.array_equal(engine.build_fock_df(h, density, coeff, n_occ), h)
    j, k = engine.build_jk(density, coeff, n_occ)
    assert not j.any() and not k.any()


def test_sharded_tensor_partial_sums(engine):
    """A shard holds auxiliary functions [q_begin, q_begin+q_count); partial J/K add up."""
    n, n_occ, naux = 64, 12, 90
    b, h, density, coeff = synth.synth_problem(15, n, n_occ, naux)
    j_ref, k_ref, _ = oracle.jk_df(b, density, coeff, n_occ)
    j_sum, k_sum = np.zeros((n, n)), np.zeros((n, n))
    for rank in range(3):
        q0, qc = synth.shard_range(naux, 3, rank)
        engine.set_tensor_shard(b[:, q0:q0 + qc], n, naux, q0)
        j, k = engine.build_jk(density, coeff, n_occ)
        j_sum += j
        k_sum += k
    assert _maxabs(j_sum, j_ref) <= TOL and _maxabs(k_sum, k_ref) <= TOL


def test_errors_surface_as_status_and_message(engine):
    """Refuse, never abort (reference: result%error%set + return)."""
    from metalquicha_b200 import B200FockEngine
    with B200FockEngine(0) as fresh:
        with pytest.raises(B200Error, match="no fitted tensor"):
            fresh.build_fock_df(np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((4, 2)), 2)
        with pytest.raises(B200Error, match="slot"):
            fresh.synth_tensor(8, 4, 1, 1.0, slot=5)
    n, n_occ, naux = 20, 3, 10
    b, h, density, coeff = synth.synth_problem(2, n, n_occ, naux)
    engine.set_tensor(b)
    with pytest.raises(B200Error, match="negative"):
        engine.build_fock_df(h, density, coeff, -1)
    # the engine is still usable after a refused call
    assert _maxabs(engine.build_fock_df(h, density, coeff, n_occ),
                   oracle.build_fock_df(h, b, density, coeff, n_occ)) <= TOL
