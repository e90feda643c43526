"""Shared helpers for the parity tests: seeded synthetic inputs in the engine's wire format."""
import random

import numpy as np

from oracle import smt as osmt
from oracle.field import R


def elems(values):
    values = list(values)
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in values), dtype=np.uint8).reshape(
        len(values), 32).copy()


def ints(arr):
    a = np.ascontiguousarray(arr, dtype=np.uint8).reshape(-1, 32)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def rand_fr(rng):
    return rng.randrange(R)


def dense_proof(rng, n_levels, key_bits=None):
    """Primary 'dense' distribution: siblings[0..n-2] non-zero, siblings[n-1] = 0, root = oracle fold."""
    key = rng.getrandbits(key_bits or n_levels)
    value = rand_fr(rng)
    sib = [rng.randrange(1, R) for _ in range(n_levels - 1)] + [0]
    root = osmt.fold_inclusion(sib, key, value)
    return root, sib, key, value


def census_proof(rng, n_levels, lo=20, hi=28):
    """Secondary 'census-like' distribution: L ~ U[lo,hi] leading siblings with ~10% interior zeros, rest 0."""
    L = rng.randint(lo, min(hi, n_levels - 1))
    key = rng.getrandbits(n_levels)
    value = rand_fr(rng)
    sib = [0 if rng.random() < 0.1 else rng.randrange(1, R) for _ in range(L)]
    sib[L - 1] = rng.randrange(1, R)
    sib += [0] * (n_levels - L)
    root = osmt.fold_inclusion(sib, key, value)
    return root, sib, key, value
