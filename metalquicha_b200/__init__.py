"""metalquicha_b200 -- B200-native density-fitted J/K Fock-build engine.

A drop-in for ONE hot path of JorgeG94/metalquicha: ``build_fock_df``
(backends/libcint/mqc_libcint_rhf.f90:1576-1646) and the tensor it consumes.
The product is ``libmqcb200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/mqcb200.h``); this package is the host-side mirror of the reference
interface used by the tests and the benchmark.
"""
from .engine import B200Error, B200FockEngine, WorkQueue  # noqa: F401

__all__ = ["B200FockEngine", "B200Error", "WorkQueue"]
