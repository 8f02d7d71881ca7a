"""world_size-2 tests of the N>1 host logic on CPU (gloo): the auxiliary-index
partition + one sum all-reduce of [J;K], the broadcast of the communicator id, and the
store-backed fragment FIFO.  The per-rank build is played by the oracle here (test
stand-in for the CUDA engine, which needs a GPU); the GPU multi-rank path itself is
covered by tests/test_gpu_multirank.py on a multi-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metalquicha_b200 import farm, synth
from metalquicha_b200.sharded import ShardedFockBuilder
from oracle import df_fock_oracle as oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _OracleEngine:
    """Same surface as B200FockEngine for the calls ShardedFockBuilder makes; the
    all-reduce that the real engine does on NCCL is done here on gloo."""

    def __init__(self):
        self.comm = None

    @staticmethod
    def comm_unique_id():
        return bytes(range(128))

    def comm_init(self, world, rank, uid):
        assert uid == bytes(range(128)) and len(uid) == 128
        self.comm = (world, rank)

    def set_tensor_shard(self, b_shard, n, naux_total, q_begin, slot=0):
        self.b, self.n, self.q_begin = b_shard, n, q_begin

    def build_jk(self, density, coeff, n_occ, slot=0):
        assert self.comm is not None, "build before the communicator exists"
        j, k, _ = oracle.jk_df(self.b, density, coeff, n_occ)
        jk = torch.from_numpy(np.stack([j, k]))
        dist.all_reduce(jk)                              # ONE all-reduce of [J;K]
        return jk[0].numpy(), jk[1].numpy()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, n_occ, naux = 26, 5, 31
        b, h, density, coeff = synth.synth_problem(99, n, n_occ, naux)
        builder = ShardedFockBuilder(_OracleEngine(), rank, world)
        builder.set_tensor(b, n, naux)
        q0, qc = builder.my_range(naux)
        assert builder.engine.b.shape == (n * n, qc) and builder.engine.q_begin == q0
        j, k = builder.build_jk(density, coeff, n_occ)
        j_ref, k_ref, _ = oracle.jk_df(b, density, coeff, n_occ)
        assert np.max(np.abs(j - j_ref)) <= 1e-12 and np.max(np.abs(k - k_ref)) <= 1e-12

        # store-backed FIFO: every fragment handed out exactly once, in FIFO order per rank
        sizes = [3, 1, 2, 3, 1, 2, 3, 3, 1, 1, 2]
        order = farm.sort_fragments_largest_first(sizes)
        store = dist.distributed_c10d._get_default_store()
        q = farm.DistributedWorkQueue(order, store, name="t1")
        mine = list(farm.worker_loop(q, lambda i: sizes[i]).keys())
        pos = [order.index(i) for i in mine]
        assert pos == sorted(pos)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            assert sorted(sum(gathered, [])) == sorted(order)
        assert q.pop() == (-1, False)

        # several ids claimed per round trip to the store (what the fragment farm of bench.py does):
        # still every fragment exactly once, and each rank works its claims off in queue order
        q3 = farm.DistributedWorkQueue(order, store, name="t3", batch=3)
        mine3 = list(farm.worker_loop(q3, lambda i: sizes[i]).keys())
        pos3 = [order.index(i) for i in mine3]
        assert pos3 == sorted(pos3)
        for a in range(0, len(pos3) - len(pos3) % 3, 3):           # claims are runs of consecutive queue positions
            assert pos3[a + 1] == pos3[a] + 1 and pos3[a + 2] == pos3[a] + 2
        gathered3 = [None] * world
        dist.all_gather_object(gathered3, mine3)
        if rank == 0:
            assert sorted(sum(gathered3, [])) == sorted(order)
        assert q3.pop() == (-1, False)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
