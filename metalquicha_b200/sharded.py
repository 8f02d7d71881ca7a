"""Whole-molecule builds sharded over GPUs by auxiliary index.

gamma_Q and X_Q depend only on slab Q, and J and K are sums over Q, so each rank
holds a contiguous auxiliary slab of the packed tensor, receives the full D and
C_occ, and the partial [J;K] are summed by ONE all-reduce per build.  That all-reduce
runs inside libmqcb200.so on NCCL (mqcb200_comm_init); this module is the host
plumbing around it: the partition, and carrying NCCL's 128-byte id from rank 0 to the
other ranks over whatever process group the host program already has (the reference
would use its MPI broadcast, src/parallel/mqc_bcast.f90).

The reference never splits one Fock build across ranks, so this has no
reference counterpart to mirror -- only the per-rank call is the reference's
``build_fock_df``.
"""
from __future__ import annotations

import numpy as np

from .synth import shard_range


class ShardedFockBuilder:
    """One instance per rank.  ``engine`` is this rank's B200FockEngine; ``group`` a
    torch.distributed process group (any backend -- it only carries the 128-byte id)."""

    def __init__(self, engine, rank: int, world: int, group=None):
        self.engine = engine
        self.rank = rank
        self.world = world
        self.group = group
        self._comm_ready = False

    # -- partition ---------------------------------------------------------------------
    def my_range(self, naux: int):
        return shard_range(naux, self.world, self.rank)

    # -- communicator --------------------------------------------------------------------
    def init_comm(self) -> None:
        if self.world == 1 or self._comm_ready:
            return
        import torch.distributed as dist
        box = [self.engine.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=self.group)
        self.engine.comm_init(self.world, self.rank, box[0])
        self._comm_ready = True

    # -- tensor ---------------------------------------------------------------------------
    def set_tensor(self, b_full_or_none, n: int, naux: int, slot: int = 0, slab_loader=None) -> None:
        """Give this rank its slab.  Either ``b_full_or_none`` is the whole ``(n*n, naux)``
        tensor (small cases, tests) or ``slab_loader(q_begin, q_count)`` returns the slab."""
        q0, qc = self.my_range(naux)
        if slab_loader is not None:
            slab = slab_loader(q0, qc)
        else:
            slab = np.asfortranarray(b_full_or_none[:, q0:q0 + qc])
        self.engine.set_tensor_shard(slab, n, naux, q0, slot=slot)

    def synth_tensor(self, n: int, naux: int, seed: int, scale: float, slot: int = 0) -> None:
        q0, qc = self.my_range(naux)
        self.engine.synth_tensor(n, naux, seed, scale, q_begin=q0, q_count=qc, slot=slot)

    # -- build (every rank gets the full Fock matrix back) -------------------------------------
    def build_fock_df(self, h, density, coeff, n_occ, k_scale=None, j_scale=None, slot: int = 0):
        self.init_comm()
        return self.engine.build_fock_df(h, density, coeff, n_occ, k_scale=k_scale, j_scale=j_scale, slot=slot)

    def build_jk(self, density, coeff, n_occ, slot: int = 0):
        self.init_comm()
        return self.engine.build_jk(density, coeff, n_occ, slot=slot)
