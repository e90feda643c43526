"""ctypes access to oracle/c/liboracle.so, the multithreaded C restatement (TEST INFRASTRUCTURE ONLY).

Used by tests/ for parity at sizes the Python oracle cannot reach, and by bench.py as the CPU baseline
("port" of the reference's plain-field path).  build() compiles it with gcc when missing.
"""
import ctypes
import os
import subprocess
from ctypes import c_char_p, c_int, c_size_t, c_void_p
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "c" / "liboracle.so"
BLOB = HERE.parent / "gnark_crypto_primitives_b200" / "data" / "poseidon_bn254.bin"
_lib = None


def build(force=False):
    src = HERE / "c" / "oracle.c"
    if force or not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE / "c"), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(SO))
        L.oracle_init.argtypes = [c_char_p]
        L.oracle_poseidon_hash.argtypes = [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]
        L.oracle_poseidon_multihash.argtypes = [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_int]
        L.oracle_smt_verify.argtypes = [c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                        c_int]
        L.oracle_elgamal_encrypt.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]
        L.oracle_elgamal_add.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int]
        L.oracle_elgamal_tally.argtypes = [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int]
        L.oracle_fixed_base_mul.argtypes = [c_void_p, c_size_t, c_void_p]
        L.oracle_keccak_address.argtypes = [c_void_p, c_size_t, c_void_p, c_int]
        rc = L.oracle_init(str(BLOB).encode())
        if rc != 0:
            raise RuntimeError(f"oracle_init failed: {rc}")
        _lib = L
    return _lib


def default_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _p(a):
    return None if a is None else c_void_p(a.ctypes.data)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def poseidon_hash(inputs, threads=1):
    a = _c(inputs)
    n, arity = a.shape[0], a.shape[1]
    out = np.empty((n, 32), np.uint8)
    st = np.empty(n, np.uint8)
    assert lib().oracle_poseidon_hash(_p(a), arity, n, _p(out), _p(st), threads) == 0
    return out, st


def poseidon_multihash(inputs, threads=1):
    a = _c(inputs)
    n, ln = a.shape[0], a.shape[1]
    out = np.empty((n, 32), np.uint8)
    st = np.empty(n, np.uint8)
    assert lib().oracle_poseidon_multihash(_p(a), ln, n, _p(out), _p(st), threads) == 0
    return out, st


def smt_verify(roots, siblings, keys, values, old_keys=None, old_values=None, is_old0=None, fnc=None, enabled=None,
               literal=True, threads=1):
    sib = _c(siblings)
    n, n_levels = sib.shape[0], sib.shape[1]
    r = _c(roots)
    shared = 1 if r.size == 32 and n != 1 else 0
    k, v, ok, ov = _c(keys), _c(values), _c(old_keys), _c(old_values)
    i0, fn, en = _c(is_old0), _c(fnc), _c(enabled)
    flags = np.empty(n, np.uint8)
    st = np.empty(n, np.uint8)
    oroots = np.empty((n, 32), np.uint8)
    assert lib().oracle_smt_verify(n_levels, n, _p(r), shared, _p(sib), _p(ok), _p(ov), _p(i0), _p(k), _p(v), _p(fn),
                                   _p(en), _p(flags), _p(st), _p(oroots), 1 if literal else 0, threads) == 0
    return flags, st, oroots


def elgamal_encrypt(pk, k, m, threads=1):
    pk, k, m = _c(pk), _c(k), _c(m)
    n = k.size // 32
    per_item = 1 if pk.size == n * 64 and n != 1 else 0
    out = np.empty((n, 4, 32), np.uint8)
    st = np.empty(n, np.uint8)
    assert lib().oracle_elgamal_encrypt(_p(pk), per_item, _p(k), _p(m), n, _p(out), _p(st), threads) == 0
    return out, st


def elgamal_add(a, b, threads=1):
    a, b = _c(a), _c(b)
    n = a.size // 128
    out = np.empty((n, 4, 32), np.uint8)
    st = np.empty(n, np.uint8)
    assert lib().oracle_elgamal_add(_p(a), _p(b), n, _p(out), _p(st), threads) == 0
    return out, st


def elgamal_tally(ct, threads=1):
    ct = _c(ct)
    n_ballots, n_fields = ct.shape[0], ct.shape[1]
    out = np.empty((n_fields, 4, 32), np.uint8)
    st = np.empty(n_fields, np.uint8)
    assert lib().oracle_elgamal_tally(_p(ct), n_ballots, n_fields, _p(out), _p(st), threads) == 0
    return out, st


def fixed_base_mul(scalars):
    s = _c(scalars)
    n = s.size // 32
    out = np.empty((n, 2, 32), np.uint8)
    assert lib().oracle_fixed_base_mul(_p(s), n, _p(out)) == 0
    return out


def keccak_address(pub_xy_be, threads=1):
    a = _c(pub_xy_be)
    n = a.size // 64
    out = np.empty((n, 20), np.uint8)
    assert lib().oracle_keccak_address(_p(a), n, _p(out), threads) == 0
    return out
