"""Engine: one context per GPU over the C ABI (include/gcp_b200.h).

Host-side data are numpy arrays of 32-byte little-endian field elements (dtype uint8, last axis 32); the
device-level methods (`*_dev`) take torch CUDA tensors (or raw device pointers) and a stream, and are what
bench.py times for resident-data throughput.  Nothing here computes field arithmetic on the CPU.
"""
import ctypes
from ctypes import c_void_p

import numpy as np

from . import _lib
from ._lib import EngineError, FMT_CANONICAL, FMT_MONTGOMERY

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def ints_to_elems(values) -> np.ndarray:
    """Iterable of Python ints (each < 2^256) -> (n, 32) uint8, little-endian."""
    values = list(values)
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint8).reshape(len(values), 32).copy()


def elems_to_ints(arr) -> list:
    a = np.ascontiguousarray(arr, dtype=np.uint8).reshape(-1, 32)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def _as_elems(arr, n_expected=None, name="array"):
    a = np.ascontiguousarray(arr)
    if a.dtype != np.uint8:
        if a.dtype == np.uint64 and a.shape[-1] == 4:
            a = a.view(np.uint8)
        else:
            raise TypeError(f"{name}: expected uint8[..., 32] (or uint64[..., 4]) field elements, got {a.dtype}")
    if a.shape[-1] != 32:
        raise ValueError(f"{name}: last axis must be 32 bytes")
    if n_expected is not None and a.size != n_expected * 32:
        raise ValueError(f"{name}: expected {n_expected} elements, got {a.size // 32}")
    return a


def _as_msgs(m, n_expected, fmt):
    """The messages of the fused tallies: field elements, or with MSG_U64 in fmt one uint64 each."""
    if not (fmt & _lib.MSG_U64):
        return _as_elems(m, n_expected, "m")
    a = np.ascontiguousarray(m)
    if a.dtype != np.uint64:
        raise TypeError(f"m: MSG_U64 expects uint64 values, got {a.dtype}")
    if a.size != n_expected:
        raise ValueError(f"m: expected {n_expected} values, got {a.size}")
    return a


def _ptr(a):
    if a is None:
        return None
    return c_void_p(a.ctypes.data)


def _dptr(t):
    """torch tensor / int / None -> device pointer."""
    if t is None:
        return None
    if isinstance(t, int):
        return c_void_p(t)
    return c_void_p(t.data_ptr())


def _u8(arr, n, name):
    if arr is None:
        return None
    a = np.ascontiguousarray(arr, dtype=np.uint8)
    if a.size != n:
        raise ValueError(f"{name}: expected {n} bytes")
    return a


def _ballot_batch_call(fn, handle, n_levels, roots, siblings, packed, keys, values, pub_key, k, m, fmt):
    """Shared argument marshalling of gcp_ballot_batch / gcp_group_ballot_batch."""
    kk = _as_elems(k, name="k")
    if kk.ndim != 3:
        raise ValueError("k must have shape (n_voters, n_fields, 32)")
    n, nf = kk.shape[0], kk.shape[1]
    mm = _as_msgs(m, n * nf, fmt)
    pk = _as_elems(pub_key, 2, "pub_key")
    r = _as_elems(roots, name="roots")
    shared = 1 if r.size == 32 and n != 1 else 0
    kv = _as_elems(keys, n, "keys")
    vv = _as_elems(values, n, "values")
    sib = blob = offs = None
    if packed is None:
        sib = _as_elems(siblings, n * int(n_levels), "siblings")
    else:
        lens = np.fromiter((len(b) for b in packed), dtype=np.uint64, count=len(packed))
        offs = np.zeros(len(packed) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        blob = np.frombuffer(b"".join(bytes(b) for b in packed) or b"\0", dtype=np.uint8)
        if len(packed) != n:
            raise ValueError("one packed proof per voter")
    flags = np.empty(n, dtype=np.uint8)
    status = np.empty(n, dtype=np.uint8)
    tally = np.empty((nf, 4, 32), dtype=np.uint8)
    tstatus = np.empty(nf, dtype=np.uint8)
    rc = fn(handle, int(n_levels), n, _ptr(r), shared, _ptr(sib), _ptr(blob), _ptr(offs), _ptr(kv), _ptr(vv), _ptr(pk),
            _ptr(kk), _ptr(mm), nf, _ptr(flags), _ptr(status), _ptr(tally), _ptr(tstatus), fmt)
    return rc, (flags, status, tally, tstatus)


class PinnedBuffer:
    """Page-locked host memory from gcp_host_alloc, viewed as a numpy uint8 array (`.array`); free with close()."""

    def __init__(self, nbytes: int):
        self._lib = _lib.load()
        p = c_void_p()
        rc = self._lib.gcp_host_alloc(int(nbytes), ctypes.byref(p))
        if rc != 0:
            raise EngineError(rc, "gcp_host_alloc failed")
        self._p = p
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * int(nbytes)).from_address(p.value)) if nbytes else \
            np.zeros(0, np.uint8)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            self._lib.gcp_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device: int = 0, constants_path: str = None):
        self._lib = _lib.load()
        h = c_void_p()
        rc = self._lib.gcp_ctx_create(int(device), constants_path.encode() if constants_path else None, ctypes.byref(h))
        if rc != 0:
            msg = self._lib.gcp_last_error(None)
            raise EngineError(rc, msg.decode() if msg else "gcp_ctx_create failed")
        self._h = h
        self.device = int(device)

    # -- lifecycle ----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gcp_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.gcp_last_error(self._h)
            raise EngineError(rc, msg.decode() if msg else "")

    @property
    def launch_count(self) -> int:
        return int(self._lib.gcp_ctx_launch_count(self._h))

    @staticmethod
    def _stream(stream):
        if stream is None:
            return None
        if isinstance(stream, int):
            return c_void_p(stream)
        return c_void_p(stream.cuda_stream)  # torch.cuda.Stream

    def probe_imad_wide(self) -> float:
        """Sustained IMAD.WIDE.U32 multiply-adds per second on this GPU (roofline denominator)."""
        v = ctypes.c_double(0.0)
        self._check(self._lib.gcp_probe_imad_wide(self._h, ctypes.byref(v)))
        return float(v.value)

    # -- Poseidon -----------------------------------------------------------------------------------
    def poseidon_hash(self, inputs, fmt=FMT_CANONICAL):
        """inputs: (n, arity, 32) uint8 -> (digests (n, 32) uint8, status (n,) uint8).  poseidon.go:38-45."""
        a = _as_elems(inputs, name="inputs")
        if a.ndim != 3:
            raise ValueError("inputs must have shape (n, arity, 32)")
        n, arity = a.shape[0], a.shape[1]
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_poseidon_hash(self._h, _ptr(a), arity, n, _ptr(out), _ptr(status), fmt))
        return out, status

    def poseidon_multihash(self, inputs, fmt=FMT_CANONICAL):
        """inputs: (n, len, 32) uint8, 1 <= len <= 4096.  poseidon.go:54-91."""
        a = _as_elems(inputs, name="inputs")
        if a.ndim != 3:
            raise ValueError("inputs must have shape (n, len, 32)")
        n, ln = a.shape[0], a.shape[1]
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_poseidon_multihash(self._h, _ptr(a), ln, n, _ptr(out), _ptr(status), fmt))
        return out, status

    def poseidon_hash_dev(self, d_in, arity, n, d_out, d_status=None, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_poseidon_hash_dev(self._h, _dptr(d_in), arity, n, _dptr(d_out), _dptr(d_status), fmt,
                                                    self._stream(stream)))

    def poseidon_multihash_dev(self, d_in, length, n, d_out, d_status=None, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_poseidon_multihash_dev(self._h, _dptr(d_in), length, n, _dptr(d_out),
                                                         _dptr(d_status), fmt, self._stream(stream)))

    # -- fixed-base tables ----------------------------------------------------------------------------
    def set_fixed_base_window(self, window_bits: int):
        """Window width of the precomputed tables of G and the shared public key (elgamal/mul.go:26-72 has 4-bit windows):
        8..26 bits, or 0 for the automatic choice (20 bits, 24 once a base has served 2^27 multiplications)."""
        self._check(self._lib.gcp_ctx_set_fixed_base_window(self._h, int(window_bits)))

    def fixed_base_window(self, which: int = 0) -> int:
        """Current window width of G's (0) or the cached public key's (1) table."""
        return int(self._lib.gcp_ctx_fixed_base_window(self._h, int(which)))

    # -- SMT ----------------------------------------------------------------------------------------
    def set_smt_hasher(self, hasher: int):
        """The utils.Hasher plug of the tree/smt gadgets (utils/hashers.go:10-37) for every SMT call of this engine:
        _lib.HASHER_POSEIDON (default) or _lib.HASHER_POSEIDON2."""
        self._check(self._lib.gcp_ctx_set_smt_hasher(self._h, int(hasher)))

    @property
    def smt_hasher(self) -> int:
        return int(self._lib.gcp_ctx_smt_hasher(self._h))

    def smt_leaf_hash(self, keys, values, fmt=FMT_CANONICAL):
        """smt.Hash1 (tree/smt/hash.go:10-19): Poseidon(key, values..., 1).  keys: (n, 32); values: (n, n_values, 32) with
        n_values in 0..14 ((n, 32) is one value per leaf).  Returns (hashes (n, 32), status)."""
        k = _as_elems(keys, name="keys").reshape(-1, 32)
        n = k.shape[0]
        v = np.ascontiguousarray(np.asarray(values, dtype=np.uint8))
        if v.ndim == 2 and v.shape == (n, 32):
            v = v.reshape(n, 1, 32)
        if v.ndim != 3 or v.shape[0] != n or v.shape[2] != 32:
            raise ValueError("values must have shape (n, n_values, 32)")
        n_values = v.shape[1]
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_smt_leaf_hash(self._h, _ptr(k), _ptr(v) if n_values else None, n_values, n, _ptr(out),
                                                _ptr(status), fmt))
        return out, status

    def smt_leaf_hash_dev(self, d_keys, d_values, n_values, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_smt_leaf_hash_dev(self._h, _dptr(d_keys), _dptr(d_values), int(n_values), int(n),
                                                    _dptr(d_out), _dptr(d_status), fmt, self._stream(stream)))

    def smt_verify_with_leaf_hash(self, roots, siblings, keys, hash1_new, old_keys=None, hash1_old=None, is_old0=None,
                                  fnc=None, enabled=None, want_roots=False, fmt=FMT_CANONICAL):
        """smt.VerifierWithLeafHashFlag (tree/smt/verifier.go:171-242) with the caller's leaf hashes."""
        return self.smt_verify(roots, siblings, keys, hash1_new, old_keys, hash1_old, is_old0, fnc, enabled, want_roots,
                               fmt, _fn=self._lib.gcp_smt_verify_with_leaf_hash)

    def smt_verify_with_leaf_hash_dev(self, n_levels, n, d_roots, shared_root, d_siblings, d_keys, d_hash1_new, d_flags,
                                      d_status, d_old_keys=None, d_hash1_old=None, d_is_old0=None, d_fnc=None,
                                      d_enabled=None, d_out_roots=None, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_smt_verify_with_leaf_hash_dev(
            self._h, n_levels, n, _dptr(d_roots), int(bool(shared_root)), _dptr(d_siblings), _dptr(d_old_keys),
            _dptr(d_hash1_old), _dptr(d_is_old0), _dptr(d_keys), _dptr(d_hash1_new), _dptr(d_fnc), _dptr(d_enabled),
            _dptr(d_flags), _dptr(d_status), _dptr(d_out_roots), fmt, self._stream(stream)))

    def smt_verify(self, roots, siblings, keys, values, old_keys=None, old_values=None, is_old0=None, fnc=None,
                   enabled=None, want_roots=False, fmt=FMT_CANONICAL, _fn=None):
        """smt.Verifier (tree/smt/verifier.go:102-121) over a batch.

        siblings: (n, n_levels, 32); roots: (n, 32) or (32,)/(1, 32) for one shared root; keys/values: (n, 32).
        Returns (flags, status[, roots]).
        """
        sib = _as_elems(siblings, name="siblings")
        if sib.ndim != 3:
            raise ValueError("siblings must have shape (n, n_levels, 32)")
        n, n_levels = sib.shape[0], sib.shape[1]
        r = _as_elems(roots, name="roots")
        shared = 1 if r.size == 32 and n != 1 else 0
        if not shared:
            r = _as_elems(r, n, "roots")
        k = _as_elems(keys, n, "keys")
        v = _as_elems(values, n, "values")
        ok = _as_elems(old_keys, n, "old_keys") if old_keys is not None else None
        ov = _as_elems(old_values, n, "old_values") if old_values is not None else None
        i0 = _u8(is_old0, n, "is_old0")
        fn = _u8(fnc, n, "fnc")
        en = _u8(enabled, n, "enabled")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        oroots = np.empty((n, 32), dtype=np.uint8) if want_roots else None
        call = _fn or self._lib.gcp_smt_verify
        self._check(call(self._h, n_levels, n, _ptr(r), shared, _ptr(sib), _ptr(ok), _ptr(ov), _ptr(i0), _ptr(k), _ptr(v),
                         _ptr(fn), _ptr(en), _ptr(flags), _ptr(status), _ptr(oroots), fmt))
        return (flags, status, oroots) if want_roots else (flags, status)

    def smt_verify_packed(self, roots, packed, n_levels, keys, values, offsets=None, old_keys=None, old_values=None,
                          is_old0=None, fnc=None, enabled=None, want_roots=False, fmt=FMT_CANONICAL):
        """smt.Verifier fed with arbo's packed proofs (what GenProof returns; the reference's callers run
        arbo.UnpackSiblings and pad to `levels` on the CPU, tree/smt/wrapper_arbo.go:63-76).

        packed: a list of byte strings, or one bytes/uint8 blob together with `offsets` (n + 1 byte offsets).
        Returns (flags, status[, roots]); a string arbo would reject has status 7 (malformed), flag 0.
        """
        if offsets is None:
            lens = np.fromiter((len(b) for b in packed), dtype=np.uint64, count=len(packed))
            offs = np.zeros(len(packed) + 1, dtype=np.uint64)
            np.cumsum(lens, out=offs[1:])
            blob = np.frombuffer(b"".join(bytes(b) for b in packed), dtype=np.uint8)
        else:
            offs = np.ascontiguousarray(offsets, dtype=np.uint64)
            blob = np.frombuffer(packed, dtype=np.uint8) if isinstance(packed, (bytes, bytearray)) else \
                np.ascontiguousarray(packed, dtype=np.uint8)
            if offs.ndim != 1 or offs.size < 1 or int(offs[-1]) > blob.size:
                raise ValueError("offsets must hold n + 1 byte offsets into packed")
        if blob.size == 0:
            blob = np.zeros(1, dtype=np.uint8)
        n = offs.size - 1
        r = _as_elems(roots, name="roots")
        shared = 1 if r.size == 32 and n != 1 else 0
        if not shared:
            r = _as_elems(r, n, "roots")
        k = _as_elems(keys, n, "keys")
        v = _as_elems(values, n, "values")
        ok = _as_elems(old_keys, n, "old_keys") if old_keys is not None else None
        ov = _as_elems(old_values, n, "old_values") if old_values is not None else None
        i0 = _u8(is_old0, n, "is_old0")
        fn = _u8(fnc, n, "fnc")
        en = _u8(enabled, n, "enabled")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        oroots = np.empty((n, 32), dtype=np.uint8) if want_roots else None
        self._check(self._lib.gcp_smt_verify_packed(self._h, int(n_levels), n, _ptr(r), shared, _ptr(blob), _ptr(offs),
                                                    _ptr(ok), _ptr(ov), _ptr(i0), _ptr(k), _ptr(v), _ptr(fn), _ptr(en),
                                                    _ptr(flags), _ptr(status), _ptr(oroots), fmt))
        return (flags, status, oroots) if want_roots else (flags, status)

    def smt_unpack_siblings_dev(self, n_levels, n, d_packed, packed_bytes, d_offsets, d_siblings, d_bad,
                                fmt=FMT_CANONICAL, stream=None):
        """arbo.UnpackSiblings + zero padding on device buffers (feeds smt_verify_dev)."""
        self._check(self._lib.gcp_smt_unpack_siblings_dev(self._h, n_levels, n, _dptr(d_packed), packed_bytes,
                                                          _dptr(d_offsets), _dptr(d_siblings), _dptr(d_bad), fmt,
                                                          self._stream(stream)))

    def smt_verify_inclusion(self, roots, siblings, keys, values, want_roots=False, fmt=FMT_CANONICAL):
        """smt.InclusionVerifier (tree/smt/verifier.go:29-43)."""
        sib = _as_elems(siblings, name="siblings")
        n, n_levels = sib.shape[0], sib.shape[1]
        r = _as_elems(roots, name="roots")
        shared = 1 if r.size == 32 and n != 1 else 0
        k = _as_elems(keys, n, "keys")
        v = _as_elems(values, n, "values")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        oroots = np.empty((n, 32), dtype=np.uint8) if want_roots else None
        self._check(self._lib.gcp_smt_verify_inclusion(self._h, n_levels, n, _ptr(r), shared, _ptr(sib), _ptr(k),
                                                       _ptr(v), _ptr(flags), _ptr(status), _ptr(oroots), fmt))
        return (flags, status, oroots) if want_roots else (flags, status)

    def smt_verify_exclusion(self, roots, siblings, old_keys, old_values, is_old0, keys, want_roots=False,
                             fmt=FMT_CANONICAL):
        """smt.ExclusionVerifier (tree/smt/verifier.go:66-81)."""
        sib = _as_elems(siblings, name="siblings")
        n, n_levels = sib.shape[0], sib.shape[1]
        r = _as_elems(roots, name="roots")
        shared = 1 if r.size == 32 and n != 1 else 0
        ok = _as_elems(old_keys, n, "old_keys")
        ov = _as_elems(old_values, n, "old_values")
        i0 = _u8(is_old0, n, "is_old0")
        k = _as_elems(keys, n, "keys")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        oroots = np.empty((n, 32), dtype=np.uint8) if want_roots else None
        self._check(self._lib.gcp_smt_verify_exclusion(self._h, n_levels, n, _ptr(r), shared, _ptr(sib), _ptr(ok),
                                                       _ptr(ov), _ptr(i0), _ptr(k), _ptr(flags), _ptr(status),
                                                       _ptr(oroots), fmt))
        return (flags, status, oroots) if want_roots else (flags, status)

    def smt_verify_dev(self, n_levels, n, d_roots, shared_root, d_siblings, d_keys, d_values, d_flags, d_status,
                       d_old_keys=None, d_old_values=None, d_is_old0=None, d_fnc=None, d_enabled=None,
                       d_out_roots=None, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_smt_verify_dev(self._h, n_levels, n, _dptr(d_roots), int(bool(shared_root)),
                                                 _dptr(d_siblings), _dptr(d_old_keys), _dptr(d_old_values),
                                                 _dptr(d_is_old0), _dptr(d_keys), _dptr(d_values), _dptr(d_fnc),
                                                 _dptr(d_enabled), _dptr(d_flags), _dptr(d_status),
                                                 _dptr(d_out_roots), fmt, self._stream(stream)))

    # -- ElGamal ------------------------------------------------------------------------------------
    def elgamal_fixed_base_mul(self, scalars, fmt=FMT_CANONICAL):
        """FixedBaseScalarMulBN254 (elgamal/mul.go:76-166): (n, 32) scalars -> ((n, 2, 32) points, status)."""
        s = _as_elems(scalars, name="scalars").reshape(-1, 32)
        n = s.shape[0]
        out = np.empty((n, 2, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_fixed_base_mul(self._h, _ptr(s), n, _ptr(out), _ptr(status), fmt))
        return out, status

    def elgamal_scalar_mul(self, points, scalars, points2=None, scalars2=None, fmt=FMT_CANONICAL):
        """curve.ScalarMul over a batch (call sites elgamal/encrypt.go:55, ciphertext.go:58,147-160): [s]P, or with a
        second base [s]P + [s2]P2 in one pass sharing the doublings.  Returns ((n, 2, 32) points, status)."""
        s = _as_elems(scalars, name="scalars").reshape(-1, 32)
        n = s.shape[0]
        p = _as_elems(points, 2 * n, "points")
        p2 = s2 = None
        if (points2 is None) != (scalars2 is None):
            raise ValueError("points2 and scalars2 go together")
        if points2 is not None:
            p2 = _as_elems(points2, 2 * n, "points2")
            s2 = _as_elems(scalars2, n, "scalars2")
        out = np.empty((n, 2, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_scalar_mul(self._h, _ptr(p), _ptr(s), _ptr(p2), _ptr(s2), n, _ptr(out),
                                                     _ptr(status), fmt))
        return out, status

    def elgamal_scalar_mul_dev(self, d_points, d_scalars, d_points2, d_scalars2, n, d_out, d_status, fmt=FMT_CANONICAL,
                               stream=None):
        self._check(self._lib.gcp_elgamal_scalar_mul_dev(self._h, _dptr(d_points), _dptr(d_scalars), _dptr(d_points2),
                                                         _dptr(d_scalars2), int(n), _dptr(d_out), _dptr(d_status), fmt,
                                                         self._stream(stream)))

    def elgamal_encrypt(self, pub_key, k, m, fmt=FMT_CANONICAL):
        """(*Ciphertext).Encrypt (elgamal/encrypt.go:42-64).  pub_key: (2, 32) shared or (n, 2, 32) per item.
        Returns ((n, 4, 32) ciphertexts in Serialize order, status)."""
        kk = _as_elems(k, name="k").reshape(-1, 32)
        n = kk.shape[0]
        mm = _as_elems(m, n, "m").reshape(-1, 32)
        pk = _as_elems(pub_key, name="pub_key")
        per_item = 0 if pk.size == 64 and (n != 1 or pk.ndim <= 2) else 1
        if per_item and pk.size != n * 64:
            raise ValueError("pub_key must be one point or n points")
        out = np.empty((n, 4, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_encrypt(self._h, _ptr(pk), per_item, _ptr(kk), _ptr(mm), n, _ptr(out),
                                                  _ptr(status), fmt))
        return out, status

    def elgamal_add(self, a, b, fmt=FMT_CANONICAL):
        """(*Ciphertext).Add (elgamal/ciphertext.go:24-32), element-wise: (n, 4, 32) x2 -> (n, 4, 32)."""
        aa = _as_elems(a, name="a").reshape(-1, 4, 32)
        n = aa.shape[0]
        bb = _as_elems(b, n * 4, "b").reshape(-1, 4, 32)
        out = np.empty((n, 4, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_add(self._h, _ptr(aa), _ptr(bb), n, _ptr(out), _ptr(status), fmt))
        return out, status

    def elgamal_neg(self, a, fmt=FMT_CANONICAL):
        """(*Ciphertext).Neg (elgamal/ciphertext.go:37-46)."""
        aa = _as_elems(a, name="a").reshape(-1, 4, 32)
        n = aa.shape[0]
        out = np.empty((n, 4, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_neg(self._h, _ptr(aa), n, _ptr(out), _ptr(status), fmt))
        return out, status

    def elgamal_is_equal(self, a, b):
        """(*Ciphertext).IsEqual (elgamal/ciphertext.go:79-87): (n, 4, 32) x2 -> (flags (n,), status (n,))."""
        aa = _as_elems(a, name="a").reshape(-1, 4, 32)
        n = aa.shape[0]
        bb = _as_elems(b, n * 4, "b").reshape(-1, 4, 32)
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_is_equal(self._h, _ptr(aa), _ptr(bb), n, _ptr(flags), _ptr(status)))
        return flags, status

    def elgamal_select(self, sel, i1, i2):
        """(*Ciphertext).Select (elgamal/ciphertext.go:90-96): out = sel ? i1 : i2 -> ((n, 4, 32), status)."""
        aa = _as_elems(i1, name="i1").reshape(-1, 4, 32)
        n = aa.shape[0]
        bb = _as_elems(i2, n * 4, "i2").reshape(-1, 4, 32)
        ss = _u8(sel, n, "sel")
        out = np.empty((n, 4, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_select(self._h, _ptr(ss), _ptr(aa), _ptr(bb), n, _ptr(out), _ptr(status)))
        return out, status

    def elgamal_tally(self, ct, fmt=FMT_CANONICAL):
        """Fold of Ciphertext.Add over ballots per field: (n_ballots, n_fields, 4, 32) -> ((n_fields, 4, 32), status)."""
        c = _as_elems(ct, name="ct")
        if c.ndim != 4 or c.shape[2] != 4:
            raise ValueError("ct must have shape (n_ballots, n_fields, 4, 32)")
        nb, nf = c.shape[0], c.shape[1]
        out = np.empty((nf, 4, 32), dtype=np.uint8)
        status = np.empty(nf, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_tally(self._h, _ptr(c), nb, nf, _ptr(out), _ptr(status), fmt))
        return out, status

    def elgamal_encrypt_dev(self, d_pub_key, pk_per_item, d_k, d_m, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_elgamal_encrypt_dev(self._h, _dptr(d_pub_key), int(bool(pk_per_item)), _dptr(d_k),
                                                      _dptr(d_m), n, _dptr(d_out), _dptr(d_status), fmt,
                                                      self._stream(stream)))

    def elgamal_fixed_base_mul_dev(self, d_scalars, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_elgamal_fixed_base_mul_dev(self._h, _dptr(d_scalars), n, _dptr(d_out),
                                                             _dptr(d_status), fmt, self._stream(stream)))

    def elgamal_add_dev(self, d_a, d_b, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_elgamal_add_dev(self._h, _dptr(d_a), _dptr(d_b), n, _dptr(d_out), _dptr(d_status), fmt,
                                                  self._stream(stream)))

    def elgamal_tally_dev(self, d_ct, n_ballots, n_fields, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_elgamal_tally_dev(self._h, _dptr(d_ct), n_ballots, n_fields, _dptr(d_out),
                                                    _dptr(d_status), fmt, self._stream(stream)))

    def elgamal_encrypt_tally(self, pub_key, k, m, fmt=FMT_CANONICAL):
        """Fused Encrypt + tally.  k, m: (n_ballots, n_fields, 32) -> ((n_fields, 4, 32), status (n_fields,)).
        With _lib.MSG_U64 or-ed into fmt, m is a uint64 array (n_ballots, n_fields): 40 instead of 64 bytes per encryption."""
        kk = _as_elems(k, name="k")
        if kk.ndim != 3:
            raise ValueError("k must have shape (n_ballots, n_fields, 32)")
        nb, nf = kk.shape[0], kk.shape[1]
        mm = _as_msgs(m, nb * nf, fmt)
        pk = _as_elems(pub_key, 2, "pub_key")
        out = np.empty((nf, 4, 32), dtype=np.uint8)
        status = np.empty(nf, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_encrypt_tally(self._h, _ptr(pk), _ptr(kk), _ptr(mm), nb, nf, _ptr(out),
                                                        _ptr(status), fmt))
        return out, status

    def elgamal_encrypt_tally_dev(self, d_pub_key, d_k, d_m, n_ballots, n_fields, d_out, d_status, fmt=FMT_CANONICAL,
                                  stream=None):
        self._check(self._lib.gcp_elgamal_encrypt_tally_dev(self._h, _dptr(d_pub_key), _dptr(d_k), _dptr(d_m), n_ballots,
                                                            n_fields, _dptr(d_out), _dptr(d_status), fmt,
                                                            self._stream(stream)))

    # -- Ethereum address ---------------------------------------------------------------------------
    def keccak_address(self, pub_xy_be):
        """DeriveAddress (ecc/secp256k1/ecdsa/address.go:14-40): (n, 64) uint8 X_be||Y_be -> (n, 20) uint8."""
        a = np.ascontiguousarray(pub_xy_be, dtype=np.uint8).reshape(-1, 64)
        n = a.shape[0]
        out = np.empty((n, 20), dtype=np.uint8)
        self._check(self._lib.gcp_keccak_address(self._h, _ptr(a), n, _ptr(out)))
        return out

    def keccak_address_dev(self, d_in, n, d_out, stream=None):
        self._check(self._lib.gcp_keccak_address_dev(self._h, _dptr(d_in), n, _dptr(d_out), self._stream(stream)))

    # -- end-to-end ballot batch (config 5) ---------------------------------------------------------
    def ballot_batch(self, n_levels, roots, keys, values, pub_key, k, m, siblings=None, packed=None, fmt=FMT_CANONICAL):
        """End-to-end ballot batch from host buffers (config 5): census proofs dense (`siblings` (n, n_levels, 32)) or
        arbo packed (`packed`: list of byte strings); k, m: (n_voters, n_fields, 32).
        Returns (flags, status, tally (n_fields, 4, 32), tally_status)."""
        rc, out = _ballot_batch_call(self._lib.gcp_ballot_batch, self._h, n_levels, roots, siblings, packed, keys, values,
                                     pub_key, k, m, fmt)
        self._check(rc)
        return out

    def ballot_batch_dev(self, n_levels, n_voters, d_roots, shared_root, d_siblings, d_keys, d_values, d_pub_key, d_k,
                         d_m, n_fields, d_flags, d_status, d_tally, d_tally_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_ballot_batch_dev(self._h, n_levels, n_voters, _dptr(d_roots), int(bool(shared_root)),
                                                   _dptr(d_siblings), _dptr(d_keys), _dptr(d_values), _dptr(d_pub_key),
                                                   _dptr(d_k), _dptr(d_m), n_fields, _dptr(d_flags), _dptr(d_status),
                                                   _dptr(d_tally), _dptr(d_tally_status), fmt, self._stream(stream)))

    # -- SMT processor ------------------------------------------------------------------------------
    def smt_process_with_leaf_hash(self, old_roots, siblings, old_keys, hash1_old, is_old0, new_keys, hash1_new, fnc0,
                                   fnc1, fmt=FMT_CANONICAL):
        """smt.ProcessorWithLeafHash (tree/smt/processor.go:16-72) -> (new_roots (n, 32), status (n,))."""
        return self.smt_process(old_roots, siblings, old_keys, hash1_old, is_old0, new_keys, hash1_new, fnc0, fnc1, fmt,
                                _fn=self._lib.gcp_smt_process_with_leaf_hash)

    def smt_process(self, old_roots, siblings, old_keys, old_values, is_old0, new_keys, new_values, fnc0, fnc1,
                    fmt=FMT_CANONICAL, _fn=None):
        """smt.Processor (tree/smt/processor.go:10-72) over a batch -> (new_roots (n, 32), status (n,))."""
        sib = _as_elems(siblings, name="siblings")
        if sib.ndim != 3:
            raise ValueError("siblings must have shape (n, n_levels, 32)")
        n, n_levels = sib.shape[0], sib.shape[1]
        args = [_as_elems(x, n, nm) for x, nm in ((old_roots, "old_roots"), (old_keys, "old_keys"),
                                                  (old_values, "old_values"), (new_keys, "new_keys"),
                                                  (new_values, "new_values"))]
        i0, f0, f1 = _u8(is_old0, n, "is_old0"), _u8(fnc0, n, "fnc0"), _u8(fnc1, n, "fnc1")
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        call = _fn or self._lib.gcp_smt_process
        self._check(call(self._h, n_levels, n, _ptr(args[0]), _ptr(sib), _ptr(args[1]), _ptr(args[2]), _ptr(i0),
                         _ptr(args[3]), _ptr(args[4]), _ptr(f0), _ptr(f1), _ptr(out), _ptr(status), fmt))
        return out, status

    def smt_process_arbo(self, old_roots, packed, n_levels, old_keys, old_values, is_old0, new_keys, new_values,
                         fnc0, fnc1, fmt=FMT_CANONICAL):
        """smt.Processor fed the way the reference's addOrUpdate builds its Assignment (wrapper_arbo.go:152-172):
        `packed` are GenProof strings taken AFTER the change; the last unpacked sibling is dropped where
        is_old0 == 0 and fnc1 == 0.  -> (new_roots (n, 32), status (n,))."""
        return self.smt_process_packed(old_roots, packed, n_levels, old_keys, old_values, is_old0, new_keys, new_values,
                                       fnc0, fnc1, fmt, _fn=self._lib.gcp_smt_process_arbo)

    def smt_process_packed(self, old_roots, packed, n_levels, old_keys, old_values, is_old0, new_keys, new_values,
                           fnc0, fnc1, fmt=FMT_CANONICAL, _fn=None):
        """smt.Processor over arbo packed proofs (list of byte strings, siblings used as they are: proofs taken BEFORE
        the change) -> (new_roots (n, 32), status (n,))."""
        n = len(packed)
        lens = np.fromiter((len(b) for b in packed), dtype=np.uint64, count=n)
        offs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        blob = np.frombuffer(b"".join(bytes(b) for b in packed) or b"\0", dtype=np.uint8)
        args = [_as_elems(x, n, nm) for x, nm in ((old_roots, "old_roots"), (old_keys, "old_keys"),
                                                  (old_values, "old_values"), (new_keys, "new_keys"),
                                                  (new_values, "new_values"))]
        i0, f0, f1 = _u8(is_old0, n, "is_old0"), _u8(fnc0, n, "fnc0"), _u8(fnc1, n, "fnc1")
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        fn = _fn or self._lib.gcp_smt_process_packed
        self._check(fn(self._h, int(n_levels), n, _ptr(args[0]), _ptr(blob), _ptr(offs), _ptr(args[1]), _ptr(args[2]),
                       _ptr(i0), _ptr(args[3]), _ptr(args[4]), _ptr(f0), _ptr(f1), _ptr(out), _ptr(status), fmt))
        return out, status

    # -- decryption checks, coordinate conversion ----------------------------------------------------
    def elgamal_assert_decrypt(self, ct, priv_keys, msgs, fmt=FMT_CANONICAL):
        """(*Ciphertext).AssertDecrypt (elgamal/ciphertext.go:50-67) -> (flags, status)."""
        c = _as_elems(ct, name="ct").reshape(-1, 4, 32)
        n = c.shape[0]
        p = _as_elems(priv_keys, n, "priv_keys")
        m = _as_elems(msgs, n, "msgs")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_assert_decrypt(self._h, _ptr(c), _ptr(p), _ptr(m), n, _ptr(flags),
                                                         _ptr(status), fmt))
        return flags, status

    def elgamal_verify_decryption_proof(self, pub_keys, ct, msgs, a1, a2, z, fmt=FMT_CANONICAL):
        """DecryptionProof.Verify (elgamal/ciphertext.go:124-168) -> (flags, status)."""
        c = _as_elems(ct, name="ct").reshape(-1, 4, 32)
        n = c.shape[0]
        pk = _as_elems(pub_keys, 2 * n, "pub_keys")
        m = _as_elems(msgs, n, "msgs")
        p1 = _as_elems(a1, 2 * n, "a1")
        p2 = _as_elems(a2, 2 * n, "a2")
        zz = _as_elems(z, n, "z")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_elgamal_verify_decryption_proof(self._h, _ptr(pk), _ptr(c), _ptr(m), _ptr(p1), _ptr(p2),
                                                                  _ptr(zz), n, _ptr(flags), _ptr(status), fmt))
        return flags, status

    def te_to_rte(self, points):
        """format.FromTEtoRTE (ecc/format/twistededwards.go:42-48): (n, 2, 32) -> ((n, 2, 32), status)."""
        return self._te_rte(points, self._lib.gcp_te_to_rte)

    def rte_to_te(self, points):
        """format.FromRTEtoTE (ecc/format/twistededwards.go:29-37)."""
        return self._te_rte(points, self._lib.gcp_rte_to_te)

    def _te_rte(self, points, fn):
        p = _as_elems(points, name="points").reshape(-1, 2, 32)
        n = p.shape[0]
        out = np.empty((n, 2, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(fn(self._h, _ptr(p), n, _ptr(out), _ptr(status)))
        return out, status

    def eddsa_verify(self, pub_keys_te, sig_r_te, sig_s, msgs, fmt=FMT_CANONICAL):
        """eddsa.Verifier.IsValid (ecc/bn254/eddsa/verifier.go:55-88) -> (flags, status)."""
        s = _as_elems(sig_s, name="sig_s").reshape(-1, 32)
        n = s.shape[0]
        a = _as_elems(pub_keys_te, 2 * n, "pub_keys_te")
        r = _as_elems(sig_r_te, 2 * n, "sig_r_te")
        m = _as_elems(msgs, n, "msgs")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_eddsa_verify(self._h, _ptr(a), _ptr(r), _ptr(s), _ptr(m), n, _ptr(flags), _ptr(status),
                                               fmt))
        return flags, status

    def smt_scan_dev(self, n_levels, n, d_siblings, d_lidx, d_info, stream=None):
        """HBM-bound proof-streaming pass: lidx (uint16) and info (uint8) per proof."""
        self._check(self._lib.gcp_smt_scan_dev(self._h, n_levels, n, _dptr(d_siblings), _dptr(d_lidx), _dptr(d_info),
                                               self._stream(stream)))

    # -- MiMC7 --------------------------------------------------------------------------------------
    def mimc7_hash(self, inputs, fmt=FMT_CANONICAL):
        """MiMC7 (hash/native/bn254/mimc7/mimc.go:47-54): (n, len, 32), 1 <= len <= 62 -> ((n, 32), status)."""
        a = _as_elems(inputs, name="inputs")
        if a.ndim != 3:
            raise ValueError("inputs must have shape (n, len, 32)")
        n, ln = a.shape[0], a.shape[1]
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_mimc7_hash(self._h, _ptr(a), ln, n, _ptr(out), _ptr(status), fmt))
        return out, status

    def mimc7_hash_dev(self, d_in, length, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_mimc7_hash_dev(self._h, _dptr(d_in), length, n, _dptr(d_out), _dptr(d_status), fmt,
                                                 self._stream(stream)))

    # -- Poseidon2, width 2 -------------------------------------------------------------------------
    def poseidon2_set_round_keys(self, keys, fmt=FMT_CANONICAL):
        """Installs the 62 round keys of poseidon2.NewPermutation(2, 6, 50) (hash/native/bn254/poseidon2/native.go:27):
        (62, 32) uint8 in round order.  A Go host passes gnark-crypto's own Parameters.RoundKeys."""
        a = _as_elems(keys, name="keys").reshape(-1, 32)
        self._check(self._lib.gcp_poseidon2_set_round_keys(self._h, _ptr(a), a.shape[0], fmt))

    def poseidon2_hash(self, inputs, fmt=FMT_CANONICAL):
        """HashPoseidon2.Hash / HashPoseidon2Gnark (hash/native/bn254/poseidon2/native.go:30-63, gnark.go:18-54):
        (n, len, 32) with len 2 (node, ordered min/max) or 3 (leaf) -> ((n, 32), status)."""
        a = _as_elems(inputs, name="inputs")
        if a.ndim != 3:
            raise ValueError("inputs must have shape (n, len, 32)")
        n, ln = a.shape[0], a.shape[1]
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_poseidon2_hash(self._h, _ptr(a), ln, n, _ptr(out), _ptr(status), fmt))
        return out, status

    def poseidon2_permutation(self, states, fmt=FMT_CANONICAL):
        """perm2.Permutation (native.go:27,55): (n, 2, 32) -> ((n, 2, 32), status)."""
        a = _as_elems(states, name="states")
        if a.ndim != 3 or a.shape[1] != 2:
            raise ValueError("states must have shape (n, 2, 32)")
        n = a.shape[0]
        out = np.empty((n, 2, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_poseidon2_permutation(self._h, _ptr(a), n, _ptr(out), _ptr(status), fmt))
        return out, status

    def poseidon2_hash_dev(self, d_in, length, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_poseidon2_hash_dev(self._h, _dptr(d_in), length, n, _dptr(d_out), _dptr(d_status), fmt,
                                                     self._stream(stream)))

    def poseidon2_permutation_dev(self, d_in, n, d_out, d_status, fmt=FMT_CANONICAL, stream=None):
        self._check(self._lib.gcp_poseidon2_permutation_dev(self._h, _dptr(d_in), n, _dptr(d_out), _dptr(d_status), fmt,
                                                            self._stream(stream)))


class Group:
    """Several GPUs of one box behind one handle (gcp_group_*, include/gcp_b200.h): the single-process form a Go host
    uses.  Batches are sharded by contiguous index range, one host thread per device, no data-path collective; the
    tallies all-gather their partial ciphertexts with NCCL.  Everything without a group form goes through
    `group.engine(i)`-style per-device contexts on the C side (gcp_group_ctx)."""

    def __init__(self, devices, constants_path: str = None):
        self._lib = _lib.load()
        devs = [int(d) for d in devices]
        arr = (ctypes.c_int * len(devs))(*devs)
        h = c_void_p()
        rc = self._lib.gcp_group_create(arr, len(devs), constants_path.encode() if constants_path else None,
                                        ctypes.byref(h))
        if rc != 0:
            msg = self._lib.gcp_group_last_error(None)
            raise EngineError(rc, msg.decode() if msg else "gcp_group_create failed")
        self._h = h
        self.devices = devs

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gcp_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.gcp_group_last_error(self._h)
            raise EngineError(rc, msg.decode() if msg else "")

    @property
    def size(self) -> int:
        return int(self._lib.gcp_group_size(self._h))

    @property
    def uses_nccl(self) -> bool:
        return bool(self._lib.gcp_group_uses_nccl(self._h))

    def launch_counts(self):
        return [int(self._lib.gcp_ctx_launch_count(c_void_p(self._lib.gcp_group_ctx(self._h, i))))
                for i in range(self.size)]

    def set_fixed_base_window(self, window_bits: int):
        """gcp_ctx_set_fixed_base_window on every device's context (gcp_group_ctx)."""
        for i in range(self.size):
            ctx = c_void_p(self._lib.gcp_group_ctx(self._h, i))
            rc = self._lib.gcp_ctx_set_fixed_base_window(ctx, int(window_bits))
            if rc != 0:
                msg = self._lib.gcp_last_error(ctx)
                raise EngineError(rc, msg.decode() if msg else "")

    def poseidon_hash(self, inputs, fmt=FMT_CANONICAL):
        a = _as_elems(inputs, name="inputs")
        if a.ndim != 3:
            raise ValueError("inputs must have shape (n, arity, 32)")
        n, arity = a.shape[0], a.shape[1]
        out = np.empty((n, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_group_poseidon_hash(self._h, _ptr(a), arity, n, _ptr(out), _ptr(status), fmt))
        return out, status

    def smt_verify(self, roots, siblings, keys, values, old_keys=None, old_values=None, is_old0=None, fnc=None,
                   enabled=None, want_roots=False, fmt=FMT_CANONICAL):
        sib = _as_elems(siblings, name="siblings")
        if sib.ndim != 3:
            raise ValueError("siblings must have shape (n, n_levels, 32)")
        n, n_levels = sib.shape[0], sib.shape[1]
        r = _as_elems(roots, name="roots")
        shared = 1 if r.size == 32 and n != 1 else 0
        k, v = _as_elems(keys, n, "keys"), _as_elems(values, n, "values")
        ok = _as_elems(old_keys, n, "old_keys") if old_keys is not None else None
        ov = _as_elems(old_values, n, "old_values") if old_values is not None else None
        i0, fn, en = _u8(is_old0, n, "is_old0"), _u8(fnc, n, "fnc"), _u8(enabled, n, "enabled")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        oroots = np.empty((n, 32), dtype=np.uint8) if want_roots else None
        self._check(self._lib.gcp_group_smt_verify(self._h, n_levels, n, _ptr(r), shared, _ptr(sib), _ptr(ok), _ptr(ov),
                                                   _ptr(i0), _ptr(k), _ptr(v), _ptr(fn), _ptr(en), _ptr(flags),
                                                   _ptr(status), _ptr(oroots), fmt))
        return (flags, status, oroots) if want_roots else (flags, status)

    def smt_verify_packed(self, roots, packed, n_levels, keys, values, want_roots=False, fmt=FMT_CANONICAL):
        lens = np.fromiter((len(b) for b in packed), dtype=np.uint64, count=len(packed))
        offs = np.zeros(len(packed) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        blob = np.frombuffer(b"".join(bytes(b) for b in packed) or b"\0", dtype=np.uint8)
        n = len(packed)
        r = _as_elems(roots, name="roots")
        shared = 1 if r.size == 32 and n != 1 else 0
        k, v = _as_elems(keys, n, "keys"), _as_elems(values, n, "values")
        flags = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        oroots = np.empty((n, 32), dtype=np.uint8) if want_roots else None
        self._check(self._lib.gcp_group_smt_verify_packed(self._h, int(n_levels), n, _ptr(r), shared, _ptr(blob),
                                                          _ptr(offs), None, None, None, _ptr(k), _ptr(v), None, None,
                                                          _ptr(flags), _ptr(status), _ptr(oroots), fmt))
        return (flags, status, oroots) if want_roots else (flags, status)

    def elgamal_encrypt(self, pub_key, k, m, fmt=FMT_CANONICAL):
        kk = _as_elems(k, name="k").reshape(-1, 32)
        n = kk.shape[0]
        mm = _as_elems(m, n, "m").reshape(-1, 32)
        pk = _as_elems(pub_key, name="pub_key")
        per_item = 0 if pk.size == 64 and (n != 1 or pk.ndim <= 2) else 1
        out = np.empty((n, 4, 32), dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self._check(self._lib.gcp_group_elgamal_encrypt(self._h, _ptr(pk), per_item, _ptr(kk), _ptr(mm), n, _ptr(out),
                                                        _ptr(status), fmt))
        return out, status

    def ballot_batch(self, n_levels, roots, keys, values, pub_key, k, m, siblings=None, packed=None, fmt=FMT_CANONICAL):
        rc, out = _ballot_batch_call(self._lib.gcp_group_ballot_batch, self._h, n_levels, roots, siblings, packed, keys,
                                     values, pub_key, k, m, fmt)
        self._check(rc)
        return out

    def elgamal_tally(self, ct, fmt=FMT_CANONICAL):
        c = _as_elems(ct, name="ct")
        if c.ndim != 4 or c.shape[2] != 4:
            raise ValueError("ct must have shape (n_ballots, n_fields, 4, 32)")
        n_ballots, n_fields = c.shape[0], c.shape[1]
        out = np.empty((n_fields, 4, 32), dtype=np.uint8)
        status = np.empty(n_fields, dtype=np.uint8)
        self._check(self._lib.gcp_group_elgamal_tally(self._h, _ptr(c), n_ballots, n_fields, _ptr(out), _ptr(status), fmt))
        return out, status

    def elgamal_encrypt_tally(self, pub_key, k, m, fmt=FMT_CANONICAL):
        kk = _as_elems(k, name="k")
        if kk.ndim != 3:
            raise ValueError("k must have shape (n_ballots, n_fields, 32)")
        n_ballots, n_fields = kk.shape[0], kk.shape[1]
        mm = _as_msgs(m, n_ballots * n_fields, fmt)
        pk = _as_elems(pub_key, 2, "pub_key")
        out = np.empty((n_fields, 4, 32), dtype=np.uint8)
        status = np.empty(n_fields, dtype=np.uint8)
        self._check(self._lib.gcp_group_elgamal_encrypt_tally(self._h, _ptr(pk), _ptr(kk), _ptr(mm), n_ballots, n_fields,
                                                              _ptr(out), _ptr(status), fmt))
        return out, status
