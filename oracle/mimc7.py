"""Oracle: iden3-compatible MiMC7 over BN254 Fr (91 rounds, x^7, Miyaguchi-Preneel with field addition).

Follows /root/reference/hash/native/bn254/mimc7/mimc.go: Sum :47-54, encrypt :80-87, pow7 :74-78, Write :33-38
(more than 62 inputs are silently dropped), constants.go:9-25 (constants[0] = 0).  Pinned by the public iden3
go-iden3-crypto vectors Hash([12]) and Hash([12, 45, 78, 41]) in tests/test_oracle_golden.py.
"""
import struct
from functools import lru_cache
from pathlib import Path

from .field import R

MAX_INPUTS = 62   # mimc.go:9
N_ROUNDS = 91     # constants.go:9
BLOB = Path(__file__).resolve().parent.parent / "gnark_crypto_primitives_b200" / "data" / "mimc7_bn254.bin"


@lru_cache(maxsize=None)
def constants():
    raw = BLOB.read_bytes()
    magic, version, n, _ = struct.unpack_from("<4I", raw, 0)
    assert magic == 0x374D494D and version == 1 and n == N_ROUNDS
    return [int.from_bytes(raw[16 + 32 * i:48 + 32 * i], "little") for i in range(n)]


def _pow7(x):
    x2 = x * x % R
    x4 = x2 * x2 % R
    return x * x2 % R * x4 % R


def _encrypt(m, h):
    x = m
    for c in constants():
        x = _pow7((x + h + c) % R)
    return (x + h) % R


def hash(inputs):
    """mimc.go:47-54 after New/Write: more than 62 inputs -> the Write is dropped and Sum returns 0."""
    if len(inputs) > MAX_INPUTS:
        return 0
    h = 0
    for d in inputs:
        h = (h + _encrypt(d % R, h) + d) % R
    return h
