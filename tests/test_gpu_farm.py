"""Fragmented runs: independent fragments pulled from the FIFO by several host workers
(engine handles) on one GPU.  Every fragment must be processed exactly once and give the
energy the oracle gives, whichever worker took it."""
import threading

import numpy as np
import pytest

from metalquicha_b200 import B200FockEngine, WorkQueue, farm, synth
from oracle import df_fock_oracle as oracle

pytestmark = pytest.mark.gpu

SHAPES = {"monomer": (24, 5, 40), "dimer": (48, 10, 60), "trimer": (72, 15, 80)}


def test_fragment_farm_on_one_gpu():
    kinds = ["trimer"] * 7 + ["dimer"] * 5 + ["monomer"] * 4
    sizes = [SHAPES[k][0] for k in kinds]
    order = farm.sort_fragments_largest_first(sizes)
    assert [sizes[i] for i in order] == sorted(sizes, reverse=True)

    problems, ref = {}, {}
    for idx, kind in enumerate(kinds):
        n, o, q = SHAPES[kind]
        b, h, d, c = synth.synth_problem(1000 + idx, n, o, q)
        problems[idx] = (b, h, d, c, o)
        ref[idx] = oracle.electronic_energy(h, oracle.build_fock_df(h, b, d, c, o), d)

    n_workers = 3
    engines = [B200FockEngine(0) for _ in range(n_workers)]
    results = [dict() for _ in range(n_workers)]
    queue = WorkQueue(order)

    def do_fragment(w, idx):
        b, h, d, c, o = problems[idx]
        engines[w].set_tensor(b)
        engines[w].build_fock_df(h, d, c, o)
        return engines[w].last_energy()

    def worker(w):
        results[w] = farm.worker_loop(queue, lambda idx: do_fragment(w, idx))

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(n_workers)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for e in engines:
        e.close()

    merged = {}
    for r in results:
        assert not (set(r) & set(merged))          # nobody processed a fragment twice
        merged.update(r)
    assert sorted(merged) == list(range(len(kinds)))
    for idx, e in merged.items():
        assert abs(e - ref[idx]) <= 1e-9
    assert queue.is_empty() and queue.pop() == (-1, False)
