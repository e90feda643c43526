"""BN254 scalar field Fr and the Poseidon constant blob (oracle side).

r literal: /root/reference/hash/emulated/bn254/mimc7/constants.go:18.
gnark's test engine reduces every Add/Mul/Sub mod r;
the oracle does the same with Python integers.
"""
import struct
from functools import lru_cache
from pathlib import Path

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BLOB = Path(__file__).resolve().parent.parent / "gnark_crypto_primitives_b200" / "data" / "poseidon_bn254.bin"
MAGIC = 0x32425350


def inv(a: int) -> int:
    """a^-1 mod r; 0 -> 0 (callers check for zero denominators before dividing)."""
    a %= R
    return pow(a, -1, R) if a else 0


def to_le32(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def from_le32(b: bytes) -> int:
    return int.from_bytes(b, "little")


@lru_cache(maxsize=None)
def poseidon_tables():
    """{t: dict(RP, C, S, M, P)} with M/P as row-major t*t lists indexed [j*t+i] == m[j][i]."""
    raw = BLOB.read_bytes()
    magic, version, n_t, _ = struct.unpack_from("<4I", raw, 0)
    assert magic == MAGIC and version == 1 and n_t == 16
    base = 16 + 40 * n_t
    out = {}

    def elems(off, n):
        return [from_le32(raw[base + 32 * (off + k): base + 32 * (off + k + 1)]) for k in range(n)]

    for i in range(n_t):
        t, rp, oc, nc, os_, ns, om, nm, op, np_ = struct.unpack_from("<10I", raw, 16 + 40 * i)
        out[t] = dict(RP=rp, C=elems(oc, nc), S=elems(os_, ns), M=elems(om, nm), P=elems(op, np_))
    return out
