"""Fragment dispatch for fragmented (MBE/GMBE) runs: one fragment per GPU at a time.

Semantics follow the reference's work queue and worker loop
(src/fragmentation/common/mqc_work_queue.f90:10-57; worker
src/fragmentation/mbe/mqc_mbe_mpi_fragment_distribution_scheme.F90:1296-1377):
a FIFO of int64 fragment ids, sorted largest-first by the caller
(sort_fragments_by_size, src/fragmentation/common/mqc_frag_utils.f90:121), from which
every worker pulls the next id when it is idle; a drained queue answers
``(-1, False)``.  There is no collective on the data path: results are scalars.

In the reference the coordinator rank owns the queue and answers requests over MPI.
Here (one process per GPU under torch.distributed) the queue head is an atomic
counter in the c10d store, so no rank is spent on coordination; within one process
``metalquicha_b200.WorkQueue`` (the C ABI's thread-safe FIFO) plays the same role for
one-thread-per-GPU hosts.
"""
from __future__ import annotations

import threading
from typing import Callable, Iterable, Sequence


class DistributedWorkQueue:
    """``queue_t`` shared by all ranks of a process group through its store."""

    def __init__(self, ids: Sequence[int], store, name: str = "mqcb200_fragment_queue", batch: int = 1):
        self.ids = [int(i) for i in ids]          # identical on every rank
        self.store = store
        self.key = name + "/head"
        # ``batch`` ids are claimed per round trip to the store (a TCP exchange with rank 0): the
        # reference's coordinator likewise hands a worker its next fragment while it still
        # computes (mqc_mbe_mpi_fragment_distribution_scheme.F90:1296-1377).  FIFO order is kept;
        # a claimed run of ids is worked off locally, in order, by whichever thread pops next.
        self.batch = max(1, int(batch))
        self._claimed = []
        self._lock = threading.Lock()

    def pop(self):
        """``queue_pop``: ``(item_idx, has_item)``; ``(-1, False)`` once drained."""
        with self._lock:
            if not self._claimed:
                head = self.store.add(self.key, self.batch) - self.batch    # atomic fetch-and-add
                self._claimed = self.ids[head:head + self.batch] if head < len(self.ids) else []
            if not self._claimed:
                return -1, False
            return self._claimed.pop(0), True


class LocalWorkQueue:
    """Single-process stand-in with the same interface (serial_fragment_processor)."""

    def __init__(self, ids: Iterable[int]):
        self.ids = [int(i) for i in ids]
        self.head = 0

    def pop(self):
        if self.head >= len(self.ids):
            return -1, False
        self.head += 1
        return self.ids[self.head - 1], True


def sort_fragments_largest_first(sizes: Sequence[int]):
    """Fragment ids ordered by decreasing size (ties keep input order), so the device
    pools are sized once by the first fragment (backends/cuest/CUEST.md:340-346)."""
    return [i for i, _ in sorted(enumerate(sizes), key=lambda t: (-t[1], t[0]))]


def worker_loop(queue, do_fragment_work: Callable[[int], object]):
    """``node_worker``: pull ids until the queue is drained; returns ``{id: result}``."""
    results = {}
    while True:
        idx, has_item = queue.pop()
        if not has_item:
            return results
        results[idx] = do_fragment_work(idx)
