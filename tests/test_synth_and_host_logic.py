"""Host-side pieces that need no GPU: the synthetic generator, the shard partition,
the fragment FIFO (C ABI and store-backed), the largest-first ordering."""
import numpy as np

from metalquicha_b200 import WorkQueue, farm, synth


def test_synth_tensor_is_symmetric_deterministic_and_slab_addressable():
    n, naux = 21, 9
    b = synth.synth_tensor(5, n, naux)
    assert b.shape == (n * n, naux) and b.flags.f_contiguous
    for p in range(naux):
        slab = b[:, p].reshape(n, n, order="F")
        assert np.array_equal(slab, slab.T)
    assert np.array_equal(b, synth.synth_tensor(5, n, naux))
    assert not np.array_equal(b, synth.synth_tensor(6, n, naux))
    # any slab range can be regenerated on its own (needed for tensors beyond host RAM)
    assert np.array_equal(b[:, 3:7], synth.synth_tensor(5, n, naux, q_begin=3, q_count=4))
    x = b / synth.default_scale(n, naux)
    assert 0.05 < np.std(x) < 1.0 and abs(np.mean(x)) < 0.05


def test_synth_orbitals_are_orthonormal():
    c = synth.synth_orbitals(1, 40, 7)
    assert np.max(np.abs(c.T @ c - np.eye(7))) < 1e-12
    _, h, d, c = synth.synth_problem(1, 30, 6, 5, with_tensor=False)
    assert np.allclose(h, h.T) and np.allclose(d, 2 * c @ c.T) and abs(np.trace(d) - 12) < 1e-10


def test_shard_ranges_partition_the_auxiliary_index():
    for naux in (1, 7, 116, 1800, 6800):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                q0, qc = synth.shard_range(naux, world, r)
                covered.extend(range(q0, q0 + qc))
            assert covered == list(range(naux))
            sizes = [synth.shard_range(naux, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_c_abi_queue_has_queue_t_semantics():
    """queue_init_from_list / queue_pop / queue_is_empty (mqc_work_queue.f90:17-50)."""
    q = WorkQueue([42, 7, 7, 100000000000])
    assert not q.is_empty()
    assert [q.pop() for _ in range(4)] == [(42, True), (7, True), (7, True), (100000000000, True)]
    assert q.is_empty() and q.pop() == (-1, False) and q.pop() == (-1, False)
    q.destroy()
    empty = WorkQueue([])
    assert empty.is_empty() and empty.pop() == (-1, False)


def test_c_abi_queue_is_thread_safe():
    import threading
    ids = list(range(5000))
    q = WorkQueue(ids)
    got = [[] for _ in range(8)]

    def worker(k):
        while True:
            i, ok = q.pop()
            if not ok:
                return
            got[k].append(i)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert sorted(sum(got, [])) == ids
    for g in got:
        assert g == sorted(g)          # FIFO order is preserved per consumer


def test_largest_first_ordering_and_local_worker_loop():
    sizes = [3, 9, 1, 9, 6]
    order = farm.sort_fragments_largest_first(sizes)
    assert order == [1, 3, 4, 0, 2]
    q = farm.LocalWorkQueue(order)
    seen = []
    res = farm.worker_loop(q, lambda i: (seen.append(i), sizes[i] * 2)[1])
    assert seen == order and res == {i: 2 * s for i, s in enumerate(sizes)}
    assert q.pop() == (-1, False)
