"""Synthetic inputs for tests and bench.py: references, reads with the 1 % edit model.

Everything is seeded (numpy PCG64) so that the GPU path, the CPU checker and the fixtures see the same bytes.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_reference(n: int, seed: int = 42) -> bytes:
    rng = np.random.Generator(np.random.PCG64(seed))
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)].tobytes()


def sample_reads(ref: bytes, n_reads: int, read_len: int, seed: int = 42, edit_rate: float = 0.01,
                 n_rate: float = 0.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Forward-strand reads drawn uniformly from ``ref`` with per-base edits
    (80 % substitution / 10 % 1-bp insertion / 10 % deletion), optionally sprinkled with 'N'.

    Returns (bytes uint8[total], offsets int64[n_reads+1], true_start int64[n_reads]).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    r = np.frombuffer(ref, dtype=np.uint8)
    starts = rng.integers(0, max(1, len(r) - read_len), size=n_reads)
    out = []
    offs = np.zeros(n_reads + 1, dtype=np.int64)
    for k in range(n_reads):
        seq = r[starts[k]:starts[k] + read_len]
        if edit_rate > 0:
            hit = rng.random(len(seq)) < edit_rate
            if hit.any():
                kind = rng.random(len(seq))
                res = []
                for p in range(len(seq)):
                    if not hit[p]:
                        res.append(seq[p])
                    elif kind[p] < 0.8:
                        res.append(_ACGT[(np.searchsorted(_ACGT, seq[p]) + 1 + rng.integers(0, 3)) % 4])
                    elif kind[p] < 0.9:
                        res.append(seq[p]); res.append(_ACGT[rng.integers(0, 4)])
                    # else: deletion
                seq = np.array(res, dtype=np.uint8)
        if n_rate > 0:
            seq = seq.copy()
            seq[rng.random(len(seq)) < n_rate] = ord("N")
        out.append(seq)
        offs[k + 1] = offs[k] + len(seq)
    return np.concatenate(out) if out else np.zeros(0, np.uint8), offs, starts.astype(np.int64)


def sample_reads_fast(ref: np.ndarray, n_reads: int, read_len: int, seed: int = 42, edit_rate: float = 0.01
                      ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised variant for millions of fixed-length reads: substitutions only at ``edit_rate``
    (an indel changes the length; bench batches keep one length so the workload is reproducible)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    starts = rng.integers(0, len(ref) - read_len, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = ref[idx]
    if edit_rate > 0:
        hit = rng.random(reads.shape) < edit_rate
        code = (reads >> 1) & 3            # A0 C1 T2 G3
        lut = np.frombuffer(b"ACTG", dtype=np.uint8)
        sub = lut[(code + rng.integers(1, 4, size=reads.shape, dtype=np.uint8)) & 3]
        reads = np.where(hit, sub, reads)
    offs = np.arange(n_reads + 1, dtype=np.int64) * read_len
    return np.ascontiguousarray(reads.reshape(-1)), offs, starts.astype(np.int64)
