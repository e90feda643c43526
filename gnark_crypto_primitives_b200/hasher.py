"""Batched host-side mirrors of the reference's stateful hashers (hash.Hash[T], /root/reference/hash/hash.go:9-18):
same method names and the same bookkeeping, but every `Write` carries one column per ROW of a batch and `Sum` runs all
rows on the GPU.

  Poseidon  hash/native/bn254/poseidon/poseidon.go:94-197   (MaxHashInputs = 16, :10-15)
  MiMC7     hash/native/bn254/mimc7/mimc.go:25-54           (maxInputs = 62, :9)

and of the function-form plug `utils.Hasher` (utils/hashers.go:10): PoseidonHasher (:25-27) and Poseidon2Hasher (:35-37).
(utils.MiMCHasher, :15-22, wraps gnark's std/hash/mimc - a different permutation that lives outside the reference tree -
and has no mirror here.)
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


def _columns(data):
    cols = [_as_elems(d, name="data").reshape(-1, 32) for d in data]
    if cols and any(c.shape[0] != cols[0].shape[0] for c in cols):
        raise ValueError("every column must have one element per row of the batch")
    return cols


def PoseidonHasher(engine, *data, fmt=FMT_CANONICAL):
    """utils.PoseidonHasher (utils/hashers.go:25-27) = poseidon.Hash over a batch: one (n, 32) column per argument ->
    (digests (n, 32), status (n,)).  0 or more than 16 columns raise the engine's "bad inputs provided"."""
    cols = _columns(data)
    n = cols[0].shape[0] if cols else 0
    rows = np.ascontiguousarray(np.stack(cols, axis=1)) if cols else np.zeros((n, 0, 32), dtype=np.uint8)
    return engine.poseidon_hash(rows, fmt=fmt)


def Poseidon2Hasher(engine, *data, fmt=FMT_CANONICAL):
    """utils.Poseidon2Hasher (utils/hashers.go:35-37) = HashPoseidon2Gnark over a batch: 2 columns hash an internal
    node (ordered min, max), 3 columns a leaf (key, value, flag); any other count raises "need 2 or 3 limbs"
    (hash/native/bn254/poseidon2/gnark.go:38-40)."""
    cols = _columns(data)
    if len(cols) not in (2, 3):
        raise ValueError(f"poseidon2: need 2 or 3 limbs, got {len(cols)}")
    return engine.poseidon2_hash(np.ascontiguousarray(np.stack(cols, axis=1)), fmt=fmt)
