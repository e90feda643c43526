#!/usr/bin/env python3
"""Build the Poseidon constant blob from the reference's decimal tables.

TEST/BUILD INFRASTRUCTURE.  Runs only in the dev container (needs
/root/reference); the output blob is committed so the GPU box never reads the
reference.

Source of truth: /root/reference/hash/native/bn254/poseidon/constants.go
  strC :48-2289   round constants, one list per t = 2..17, len 8t + RP
  strM :2291-4412 dense MDS matrices, t x t, indexed m[j][i] (poseidon.go:219)
  strS :4414-22731 sparse partial-round rows, len (2t-1)*RP
  strP :22733-24854 pre-sparse matrix, t x t
  getConstant(X, t) = X[t-2]  (constants.go:14-16)

Blob layout (little endian), consumed by oracle/ (Python + C) and by the engine
(gnark_crypto_primitives_b200/csrc/constants.cpp):
  u32 magic 'PSB2' (0x32425350), u32 version=1, u32 n_t=16, u32 reserved
  16 x { u32 t, RP, offC, nC, offS, nS, offM, nM, offP, nP }  (offsets in elements)
  elements: 32 bytes each, canonical (< r), little-endian integer
"""
import hashlib
import re
import struct
import sys
from pathlib import Path

REF = Path("/root/reference/hash/native/bn254/poseidon/constants.go")
REF_SHA256 = "c7f3fe3430227991b3c6db6e86ba4d434545f1d48c49fc24e122f608d8783120"
OUT = Path(__file__).resolve().parent.parent / "gnark_crypto_primitives_b200" / "data" / "poseidon_bn254.bin"
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]  # poseidon.go:119
MAGIC = 0x32425350


def _parse_nested(text: str):
    """Turn a Go `[][]string{...}` / `[][][]string{...}` literal body into nested lists of int."""
    body = text[text.index("{"):]
    body = body.replace("{", "[").replace("}", "]")
    body = re.sub(r'"(\d+)"', r"\1", body)
    body = re.sub(r",\s*\]", "]", body)
    return eval(body, {"__builtins__": {}})  # digits, brackets and commas only


def load_reference_tables():
    src = REF.read_text()
    digest = hashlib.sha256(src.encode()).hexdigest()
    if digest != REF_SHA256:
        raise SystemExit(f"reference constants.go changed: sha256 {digest}")
    tables = {}
    names = ["strC", "strM", "strS", "strP"]
    starts = {n: src.index(f"var {n} = ") for n in names}
    order = sorted(names, key=lambda n: starts[n])
    for i, n in enumerate(order):
        end = starts[order[i + 1]] if i + 1 < len(order) else len(src)
        chunk = src[starts[n]:end]
        assert re.fullmatch(r'[\s\w=\[\]{}",]*', chunk), n
        tables[n[3:]] = _parse_nested(chunk)
    return tables


def build_blob(tables) -> bytes:
    C, M, S, P = tables["C"], tables["M"], tables["S"], tables["P"]
    assert len(C) == len(M) == len(S) == len(P) == 16
    elems = []
    dirs = []
    for idx in range(16):
        t = idx + 2
        rp = N_ROUNDS_P[idx]
        c, s, m, p = C[idx], S[idx], M[idx], P[idx]
        assert len(c) == 8 * t + rp, (t, len(c))
        assert len(s) == (2 * t - 1) * rp, (t, len(s))
        assert len(m) == t and all(len(row) == t for row in m)
        assert len(p) == t and all(len(row) == t for row in p)
        offC = len(elems); elems += c
        offS = len(elems); elems += s
        offM = len(elems); elems += [m[j][i] for j in range(t) for i in range(t)]
        offP = len(elems); elems += [p[j][i] for j in range(t) for i in range(t)]
        dirs.append((t, rp, offC, len(c), offS, len(s), offM, t * t, offP, t * t))
    assert all(0 <= e < R for e in elems)
    out = struct.pack("<4I", MAGIC, 1, 16, 0)
    for d in dirs:
        out += struct.pack("<10I", *d)
    out += b"".join(e.to_bytes(32, "little") for e in elems)
    return out


MIMC_REF = Path("/root/reference/hash/native/bn254/mimc7/constants.go")
MIMC_OUT = OUT.parent / "mimc7_bn254.bin"
MIMC_MAGIC = 0x374D494D  # 'MIM7'


def build_mimc_blob() -> bytes:
    """hash/native/bn254/mimc7/constants.go:9-25: constants[0] = 0, constants[i] = strConstants[i-1], 91 rounds.
    Layout: u32 magic 'MIM7', u32 version=1, u32 n=91, u32 reserved, then 91 x 32-byte little-endian elements."""
    src = MIMC_REF.read_text()
    cs = [int(x) for x in re.findall(r'"(\d+)"', src)]
    assert len(cs) == 90 and all(0 <= c < R for c in cs)
    elems = [0] + cs
    return struct.pack("<4I", MIMC_MAGIC, 1, len(elems), 0) + b"".join(e.to_bytes(32, "little") for e in elems)


P2_OUT = OUT.parent / "poseidon2_bn254_t2.bin"
P2_MAGIC = 0x32534F50  # 'POS2'


def build_poseidon2_blob() -> bytes:
    """Round keys of poseidon2.NewPermutation(2, 6, 50) (hash/native/bn254/poseidon2/native.go:27) as derived by
    oracle/poseidon2.py::round_keys (gnark-crypto's published Keccak chain; PARITY UNPINNED - no vector in the reference).
    Layout: u32 magic 'POS2', u32 version=1, u32 n=62, u32 reserved, then 62 x 32-byte little-endian elements in round
    order (3 full rounds x 2 keys, 50 partial rounds x 1 key, 3 full rounds x 2 keys)."""
    from oracle import poseidon2
    flat = poseidon2.flat_round_keys()
    assert len(flat) == 62 and all(0 <= k < R for k in flat)
    return struct.pack("<4I", P2_MAGIC, 1, len(flat), 0) + b"".join(k.to_bytes(32, "little") for k in flat)


def main():
    blob = build_blob(load_reference_tables())
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_bytes(blob)
    print(f"wrote {OUT} ({len(blob)} bytes, sha256 {hashlib.sha256(blob).hexdigest()})")
    mblob = build_mimc_blob()
    MIMC_OUT.write_bytes(mblob)
    print(f"wrote {MIMC_OUT} ({len(mblob)} bytes, sha256 {hashlib.sha256(mblob).hexdigest()})")
    pblob = build_poseidon2_blob()
    P2_OUT.write_bytes(pblob)
    print(f"wrote {P2_OUT} ({len(pblob)} bytes, sha256 {hashlib.sha256(pblob).hexdigest()})")


if __name__ == "__main__":
    sys.exit(main())
