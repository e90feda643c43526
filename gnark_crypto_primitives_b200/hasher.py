"""Batched host-side mirrors of the reference's stateful hashers (hash.Hash[T], /root/reference/hash/hash.go:9-18):
same method names and the same bookkeeping, but every `Write` carries one column per ROW of a batch and `Sum` runs all
rows on the GPU.

  Poseidon  hash/native/bn254/poseidon/poseidon.go:94-197   (MaxHashInputs = 16, :10-15)
  MiMC7     hash/native/bn254/mimc7/mimc.go:25-54           (maxInputs = 62, :9)
"""
import numpy as np

from ._lib import FMT_CANONICAL
from .engine import _as_elems


class _BatchHasher:
    MAX_INPUTS = 0

    def __init__(self, engine, fmt=FMT_CANONICAL):
        self._engine = engine
        self._fmt = fmt
        self._cols = []      # each (n, 32) uint8
        self._n = None

    def Write(self, *data):
        """Appends the given columns, each (n, 32) uint8 (one element per row).  Like the reference, a call that
        would exceed the input limit is dropped WHOLE and silently (poseidon.go:103-108, mimc.go:33-38)."""
        cols = [_as_elems(d, name="data").reshape(-1, 32) for d in data]
        for c in cols:
            if self._n is None:
                self._n = c.shape[0]
            if c.shape[0] != self._n:
                raise ValueError("every column must have one element per row of the batch")
        if len(self._cols) + len(cols) > self.MAX_INPUTS:
            return
        self._cols.extend(cols)

    def Reset(self):
        self._cols = []

    def WriteSucceeded(self) -> bool:
        return len(self._cols) > 0

    def _rows(self):
        n = self._n or 0
        if not self._cols:
            return np.zeros((n, 0, 32), dtype=np.uint8)
        return np.ascontiguousarray(np.stack(self._cols, axis=1))

    def Sum(self):
        """-> (digests (n, 32), status (n,)).  With nothing written the engine reports the reference's
        "bad inputs provided" (poseidon.go:41-43)."""
        return self._hash(self._rows())

    def SumIsEqual(self, expected):
        """-> (flags (n,), status (n,)): 1 where Sum() equals `expected` (n, 32)."""
        digests, status = self.Sum()
        exp = _as_elems(expected, digests.shape[0], "expected").reshape(-1, 32)
        flags = (digests == exp).all(axis=1).astype(np.uint8)
        flags[status != 0] = 0
        return flags, status

    def AssertSumIsEqual(self, expected):
        """Raises AssertionError where the reference's api.AssertIsEqual(flag, 1) would fail."""
        flags, _ = self.SumIsEqual(expected)
        if not flags.all():
            raise AssertionError(f"AssertSumIsEqual failed for rows {np.flatnonzero(flags == 0)[:8].tolist()}")


class Poseidon(_BatchHasher):
    MAX_INPUTS = 16

    def _hash(self, rows):
        return self._engine.poseidon_hash(rows, fmt=self._fmt)


class MiMC7(_BatchHasher):
    MAX_INPUTS = 62

    def _hash(self, rows):
        return self._engine.mimc7_hash(rows, fmt=self._fmt)
