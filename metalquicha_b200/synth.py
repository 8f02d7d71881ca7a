"""Synthetic inputs for the DF-J/K build.

Real three-centre integrals cannot be generated in this image (no integral code,
no basis-set data), so benchmarks and parity tests run on synthetic ``(B, D, C, H)``
of the named shapes.  ``B`` comes from a counter-based generator that exists twice
-- here in NumPy and in ``csrc/synth.cuh`` on the device -- and the two agree bit
for bit, so a tensor too large for host RAM can be generated on the GPU while the
CPU oracle regenerates any slab it wants to check.
"""
from __future__ import annotations

import math

import numpy as np

XNORM = 2.6429156645611437e-05   # must match MQCB200_SYNTH_XNORM in csrc/synth.cuh
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

# (n, n_occ, naux) of the BASELINE.json configurations; aux counts marked "~" in
# SURVEY 8 are estimates (the basis-set bundle is missing from the reference tree).
CONFIGS = {
    "c1": dict(n=24, n_occ=5, naux=116, what="water RHF/cc-pVDZ + cc-pVDZ-JKFIT"),
    "c1_bench": dict(n=24, n_occ=5, naux=139, what="water cc-pVDZ / cc-pVTZ-JKFIT (validation/check_df.f90:55)"),
    "c2": dict(n=688, n_occ=80, naux=1800, what="(H2O)16 RHF/def2-TZVP, def2-universal-JKFIT"),
    "c3": dict(n=72, n_occ=15, naux=340, what="(H2O)3 trimer of the (H2O)64 MBE-3 def2-SVP run"),
    "c3_dimer": dict(n=48, n_occ=10, naux=227, what="(H2O)2 dimer of the MBE-3 run"),
    "c3_monomer": dict(n=24, n_occ=5, naux=113, what="H2O monomer of the MBE-3 run"),
    "c4": dict(n=1450, n_occ=241, naux=6800, what="C60H122 RKS-B3LYP/def2-SVP (k_scale=0.2)"),
    "c5": dict(n=682, n_alpha=80, n_beta=79, naux=1800, what="OH.(H2O)15 UHF/def2-TZVP"),
}


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def default_scale(n: int, naux: int) -> float:
    """Keeps |J|, |K| of order one for orthonormal occupied orbitals."""
    return 1.0 / math.sqrt(float(naux))


def synth_slab_lower(seed: int, q: int, n: int, scale: float) -> np.ndarray:
    """Symmetric n x n slab ``B_q`` of the synthetic tensor (both triangles filled)."""
    mu, nu = np.tril_indices(n)
    key = _mix64(np.uint64(seed) ^ _mix64(np.array([0x51ED270B0B4F3A1D + q], dtype=np.uint64)))[0]
    r = _mix64(key ^ ((mu.astype(np.uint64) << np.uint64(32)) | nu.astype(np.uint64)))
    s = ((r & np.uint64(0xFFFF)) + ((r >> np.uint64(16)) & np.uint64(0xFFFF))
         + ((r >> np.uint64(32)) & np.uint64(0xFFFF)) + (r >> np.uint64(48))).astype(np.int64)
    x = (s - 131070).astype(np.float64) * XNORM
    d = (mu - nu).astype(np.float64) * (8.0 / float(n))
    decay = 1.0 / (1.0 + d * d)
    v = (x * decay) * scale
    out = np.zeros((n, n))
    out[mu, nu] = v
    out[nu, mu] = v
    return out


def synth_tensor(seed: int, n: int, naux: int, scale: float | None = None,
                 q_begin: int = 0, q_count: int | None = None) -> np.ndarray:
    """``b(n*n, q_count)`` column-major, identical to ``mqcb200_synth_tensor``."""
    if scale is None:
        scale = default_scale(n, naux)
    if q_count is None:
        q_count = naux - q_begin
    b = np.empty((n * n, q_count), order="F")
    for p in range(q_count):
        b[:, p] = synth_slab_lower(seed, q_begin + p, n, scale).reshape(n * n, order="F")
    return b


def synth_orbitals(seed: int, n: int, n_cols: int) -> np.ndarray:
    """Orthonormal columns: QR of an n x n_cols standard-normal matrix (Philox)."""
    rng = np.random.Generator(np.random.Philox(seed))
    a = rng.standard_normal((n, max(n_cols, 1)))
    q, _ = np.linalg.qr(a)
    return np.asfortranarray(q[:, :n_cols])


def synth_core_hamiltonian(seed: int, n: int) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed + 7919))
    a = rng.standard_normal((n, n))
    return np.asfortranarray(0.5 * (a + a.T))


def synth_problem(seed: int, n: int, n_occ: int, naux: int, with_tensor: bool = True):
    """``(b, h, density, coeff)`` for a closed-shell build; ``density = 2 C C^T``."""
    coeff = synth_orbitals(seed, n, n_occ)
    density = np.asfortranarray(2.0 * (coeff @ coeff.T))
    h = synth_core_hamiltonian(seed, n)
    b = synth_tensor(seed, n, naux) if with_tensor else None
    return b, h, density, coeff


def synth_physical_like_tensor(seed: int, n: int, naux: int, n_null: int = 2):
    """A tensor built the way ``build_df_tensor`` builds it: ``three . metric^(-1/2)``
    with a metric that has ``n_null`` eigenvalues below the 1e-10 drop threshold.
    Returns ``(three, metric)``; the whitening itself is the oracle's / engine's job."""
    rng = np.random.Generator(np.random.Philox(seed + 104729))
    three = np.empty((n * n, naux), order="F")
    for p in range(naux):
        a = rng.standard_normal((n, n)) / math.sqrt(n)
        three[:, p] = (0.5 * (a + a.T)).reshape(n * n, order="F")
    u, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    lam = np.exp(rng.uniform(-2.0, 2.0, size=naux))
    lam[:n_null] = 1.0e-13
    metric = (u * lam[None, :]) @ u.T
    metric = 0.5 * (metric + metric.T)
    return three, np.asfortranarray(metric)


def shard_range(naux: int, n_ranks: int, rank: int):
    """Contiguous auxiliary slab ``[q_begin, q_begin + q_count)`` of one rank."""
    q_begin = (naux * rank) // n_ranks
    q_end = (naux * (rank + 1)) // n_ranks
    return q_begin, q_end - q_begin
