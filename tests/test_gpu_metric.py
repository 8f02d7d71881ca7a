"""metric_inverse_sqrt (mqc_libcint_integrals.F90:992-1038) on the device -- one-sided Jacobi
eigensolver + GEMM -- against the oracle's dsyev restatement, SURVEY 8 row a5."""
import numpy as np
import pytest

from metalquicha_b200 import B200Error
from oracle import df_fock_oracle as oracle

pytestmark = pytest.mark.gpu


def _metric(seed, naux, n_null=0, cond=1e4):
    rng = np.random.default_rng(seed)
    u, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    lam = np.exp(rng.uniform(0.0, np.log(cond), size=naux)) / cond
    lam[:n_null] = 1.0e-13 * (1.0 + rng.uniform(size=n_null))
    m = (u * lam[None, :]) @ u.T
    return np.asfortranarray(0.5 * (m + m.T))


@pytest.mark.parametrize("naux,n_null", [(1, 0), (2, 0), (19, 0), (116, 0), (139, 3), (340, 1), (601, 0)])
def test_device_metric_inverse_sqrt_matches_the_oracle(engine, naux, n_null):
    metric = _metric(naux, naux, n_null)
    half = engine.metric_inverse_sqrt(metric)
    ref = oracle.metric_inverse_sqrt(metric)
    scale = max(1.0, float(np.max(np.abs(ref))))
    assert float(np.max(np.abs(half - ref))) <= 1e-9 * scale
    assert engine.last_metric_kept == naux - n_null
    # half . metric . half is the projector on the kept modes
    proj = half @ metric @ half
    assert abs(np.trace(proj) - (naux - n_null)) <= 1e-8 * naux
    assert np.array_equal(half, engine.metric_inverse_sqrt(metric))          # bit-reproducible
    ms, sweeps = engine.last_metric()
    assert 1 <= sweeps <= 30


def test_singular_and_indefinite_metrics(engine):
    with pytest.raises(B200Error, match="singular"):
        engine.metric_inverse_sqrt(np.zeros((6, 6)))
    # a negative eigenvalue is dropped like a null one (the reference keeps values(i) > 1e-10 only)
    m = _metric(3, 12)
    w, u = np.linalg.eigh(m)
    w[0] = -0.3
    m = np.asfortranarray((u * w[None, :]) @ u.T)
    half = engine.metric_inverse_sqrt(0.5 * (m + m.T))
    ref = oracle.metric_inverse_sqrt(0.5 * (m + m.T))
    assert engine.last_metric_kept == 11
    assert float(np.max(np.abs(half - ref))) <= 1e-9 * max(1.0, float(np.max(np.abs(ref))))


def test_slab_streamed_whitening_refusals(engine):
    """The slab API refuses what it cannot place, with a message, and leaves the handle usable."""
    from metalquicha_b200 import synth
    n, naux = 40, 24
    three, metric = synth.synth_physical_like_tensor(4, n, naux, n_null=0)
    half = engine.metric_inverse_sqrt(metric)
    blk = np.asfortranarray(three.reshape(n, n, naux, order="F")[:, 0:16, :].reshape(n * 16, naux, order="F"))
    with pytest.raises(B200Error, match="no whitening in progress"):
        engine._whiten = (0, n, naux, 0, naux)
        engine.whiten_push(0, blk)
    engine.whiten_begin(n, naux, half)
    with pytest.raises(B200Error, match="already in progress"):
        engine.whiten_begin(n, naux, half)
    # a failed push (nu_begin not on a 16-column boundary) ends the whitening on a lone rank
    with pytest.raises(B200Error, match="16-wide column tiles"):
        engine.whiten_push(8, blk)
    with pytest.raises(B200Error, match="no whitening in progress"):
        engine.whiten_end()
    # ... and the handle still works
    engine.whiten_begin(n, naux, half)
    three3 = three.reshape(n, n, naux, order="F")
    for nu0, nu1 in ((0, 16), (16, 32), (32, 40)):
        engine.whiten_push(nu0, np.asfortranarray(three3[:, nu0:nu1, :].reshape(n * (nu1 - nu0), naux, order="F")))
    engine.whiten_end()
    _, h, d, c = synth.synth_problem(4, n, 6, naux, with_tensor=False)
    b = np.asfortranarray(three @ oracle.metric_inverse_sqrt(metric))
    f = engine.build_fock_df(h, d, c, 6)
    ref = oracle.build_fock_df(h, b, d, c, 6)
    assert float(np.max(np.abs(f - ref))) <= 1e-10 * max(1.0, float(np.max(np.abs(ref))))
