"""gnark_crypto_primitives_b200 — B200-native batch engine for the data-parallel core of
vocdoni/gnark-crypto-primitives (Poseidon over BN254 Fr, SMT proof verification, BabyJubJub ElGamal).

The product is the CUDA library behind include/gcp_b200.h; this package is its thin host-side mirror.
Importing it never touches oracle/ and never computes on the CPU.
"""
from ._lib import (COORDS_TE, MSG_U64, HASHER_POSEIDON, HASHER_POSEIDON2, EngineError, FMT_CANONICAL, FMT_MONTGOMERY, STATUS_ASSERTION, STATUS_KEY_RANGE, STATUS_NONCANONICAL,
                   STATUS_MALFORMED, STATUS_NOT_BOOLEAN, STATUS_OFF_CURVE, STATUS_OK, STATUS_ZERO_DENOM)
from .engine import Engine, Group, PinnedBuffer, R, elems_to_ints, ints_to_elems
from . import hasher

__all__ = [
    "Engine", "Group", "PinnedBuffer", "hasher", "EngineError", "R", "ints_to_elems", "elems_to_ints", "FMT_CANONICAL", "FMT_MONTGOMERY", "COORDS_TE", "MSG_U64", "HASHER_POSEIDON", "HASHER_POSEIDON2", "STATUS_OK",
    "STATUS_NONCANONICAL", "STATUS_KEY_RANGE", "STATUS_NOT_BOOLEAN", "STATUS_OFF_CURVE", "STATUS_ZERO_DENOM", "STATUS_ASSERTION", "STATUS_MALFORMED",
]
