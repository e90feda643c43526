"""GPU parity: the verifier fed with arbo packed proofs (gcp_smt_verify_packed) against the oracle's
unpack -> pad -> literal state machine, and against the dense entry point on the same proofs."""
import random

import numpy as np
import pytest

from oracle import smt as osmt
from oracle.field import R
from tests.util import census_proof, elems, ints

pytestmark = pytest.mark.gpu

FMT_CANONICAL, FMT_MONTGOMERY = 0, 1
RMONT = 1 << 256


def strip(sib):
    """GenProof's sibling list: down to the leaf, i.e. without the zero padding."""
    last = max((i for i, s in enumerate(sib) if s), default=-1)
    return sib[:last + 1]


def test_real_tree_proofs_packed(engine):
    """Fixture shape of tree/test/verifier_bn254_test.go:36-68 (64 levels, 8-byte keys), proofs kept packed."""
    rng = random.Random(640)
    n_levels = 64
    tree = osmt.Tree(n_levels)
    keys = [rng.getrandbits(64) for _ in range(40)]
    for k in keys:
        tree.add(k, rng.getrandbits(100))
    root = tree.root()
    packed, cases = [], []
    for k in keys:
        p = tree.gen_proof(k)
        packed.append(tree.last_packed)
        cases.append(dict(key=k, value=p["old_value"], old_key=k, old_value=p["old_value"], is_old0=0, fnc=0))
    for _ in range(40):
        k = rng.getrandbits(64)
        p = tree.gen_proof(k)
        packed.append(tree.last_packed)
        cases.append(dict(key=k, value=0, old_key=p["old_key"], old_value=p["old_value"], is_old0=p["is_old0"], fnc=1))
    n = len(cases)
    flags, status, roots = engine.smt_verify_packed(
        elems([root]), packed, n_levels, elems(c["key"] for c in cases), elems(c["value"] for c in cases),
        old_keys=elems(c["old_key"] for c in cases), old_values=elems(c["old_value"] for c in cases),
        is_old0=np.array([c["is_old0"] for c in cases], dtype=np.uint8),
        fnc=np.array([c["fnc"] for c in cases], dtype=np.uint8), want_roots=True)
    got_roots = ints(roots)
    for i, c in enumerate(cases):
        sib, st = osmt.assignment_siblings(packed[i], n_levels)
        assert st == 0
        want = osmt.verifier(1, root, sib, c["old_key"], c["old_value"], c["is_old0"], c["key"], c["value"], c["fnc"])
        assert (int(flags[i]), int(status[i]), got_roots[i]) == want, i
    assert flags.all()


@pytest.mark.parametrize("fmt", [FMT_CANONICAL, FMT_MONTGOMERY])
def test_packed_equals_dense_on_synthetic_batches(engine, fmt):
    """Ragged lengths, interior zeros, every alignment of the 32-byte siblings in the blob, corrupted proofs, both
    element formats (siblings are canonical on the wire in either)."""
    rng = random.Random(4242 + fmt)
    n_levels = 160
    n = 1200
    conv = (lambda v: v * RMONT % R) if fmt == FMT_MONTGOMERY else (lambda v: v)
    roots, sibs, keys, vals, packed = [], [], [], [], []
    for i in range(n):
        root, sib, key, value = census_proof(rng, n_levels, lo=1, hi=60)
        if i % 7 == 3:
            j = rng.randrange(len(strip(sib)))
            sib[j] = rng.randrange(1, R)  # corrupted sibling -> flag 0
        if i % 50 == 10:
            sib = [0] * n_levels  # single-leaf tree: no siblings at all
            root = osmt.hash1(key, value)
        roots.append(root), sibs.append(sib), keys.append(key), vals.append(value)
        packed.append(osmt.pack_siblings(strip(sib)))
    pf, ps, pr = engine.smt_verify_packed(elems(map(conv, roots)), packed, n_levels, elems(map(conv, keys)),
                                          elems(map(conv, vals)), want_roots=True, fmt=fmt)
    dense = elems(conv(s) for row in sibs for s in row).reshape(n, n_levels, 32)
    df, ds, dr = engine.smt_verify(elems(map(conv, roots)), dense, elems(map(conv, keys)), elems(map(conv, vals)),
                                   want_roots=True, fmt=fmt)
    assert np.array_equal(pf, df) and np.array_equal(ps, ds) and np.array_equal(pr, dr)
    assert 0 < int(pf.sum()) < n and not ps.any()
    # sample against the oracle
    for i in rng.sample(range(n), 40):
        f, s, r = osmt.inclusion_verifier(roots[i], sibs[i], keys[i], vals[i])
        assert (int(pf[i]), int(ps[i])) == (f, s)
        assert ints(pr[i:i + 1])[0] == conv(r)


def test_malformed_truncated_and_overlong_proofs(engine):
    rng = random.Random(77)
    n_levels = 32
    root, sib, key, value = census_proof(rng, n_levels, lo=5, hi=9)
    good = osmt.pack_siblings(strip(sib))
    long_sib = [rng.randrange(1, R) for _ in range(40)]  # deeper than n_levels: the tail is dropped
    variants = [
        good,
        good[:-1],                                                       # length field mismatch
        good + b"\x00",
        b"\x03\x00\x00",                                                 # shorter than the header
        (6).to_bytes(2, "little") + (9).to_bytes(2, "little") + b"\x00\x00",  # bitmap past the end
        (45).to_bytes(2, "little") + (1).to_bytes(2, "little") + b"\x03" + bytes(40),  # cut sibling
        (37).to_bytes(2, "little") + (1).to_bytes(2, "little") + b"\x07" + (9).to_bytes(32, "little"),
        osmt.pack_siblings([]),
        osmt.pack_siblings(long_sib),
        osmt.pack_siblings([R + 5] + strip(sib)[1:]),                    # non-canonical sibling on the wire
        b"",
    ]
    n = len(variants)
    flags, status, roots = engine.smt_verify_packed(elems([root] * n), variants, n_levels, elems([key] * n),
                                                    elems([value] * n), want_roots=True)
    for i, b in enumerate(variants):
        s_or, st = osmt.assignment_siblings(b, n_levels)
        if st:
            assert (int(flags[i]), int(status[i]), ints(roots[i:i + 1])[0]) == (0, osmt.STATUS_MALFORMED, 0), i
        else:
            want = osmt.inclusion_verifier(root, s_or, key, value)
            assert (int(flags[i]), int(status[i])) == want[:2], i
            if want[1] == 0:
                assert ints(roots[i:i + 1])[0] == want[2], i
    assert int(flags[0]) == 1 and int(status[9]) == osmt.STATUS_NONCANONICAL
    assert [int(s) for s in status[1:6]] == [7, 7, 7, 7, 7] and int(status[10]) == 7


def test_packed_chunked_host_path(engine, monkeypatch):
    """Several chunks on the two streams (chunk size forced small), shared root."""
    rng = random.Random(99)
    n_levels = 64
    tree = osmt.Tree(n_levels)
    keys = [rng.getrandbits(64) for _ in range(300)]
    for k in keys:
        tree.add(k, rng.getrandbits(64))
    root = tree.root()
    packed, vals = [], []
    for k in keys:
        p = tree.gen_proof(k)
        packed.append(tree.last_packed)
        vals.append(p["old_value"] if len(vals) % 11 else p["old_value"] ^ 1)
    monkeypatch.setenv("GCP_B200_SMT_CHUNK", "37")
    flags, status = engine.smt_verify_packed(elems([root]), packed, n_levels, elems(keys), elems(vals))
    assert not status.any()
    assert [int(f) for f in flags] == [1 if i % 11 else 0 for i in range(len(keys))]


def test_unpack_dev_feeds_verify_dev(engine):
    import torch

    rng = random.Random(5)
    n_levels = 160
    n = 512
    roots, keys, vals, packed, sibs = [], [], [], [], []
    for _ in range(n):
        root, sib, key, value = census_proof(rng, n_levels)
        roots.append(root), keys.append(key), vals.append(value), sibs.append(sib)
        packed.append(osmt.pack_siblings(strip(sib)))
    blob = np.frombuffer(b"".join(packed), dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([len(b) for b in packed], out=offs[1:])
    dev = torch.device("cuda:0")
    d_blob = torch.from_numpy(blob.copy()).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64).copy()).to(dev)
    d_sib = torch.empty((n, n_levels, 32), dtype=torch.uint8, device=dev)
    d_bad = torch.empty(n, dtype=torch.uint8, device=dev)
    engine.smt_unpack_siblings_dev(n_levels, n, d_blob, blob.size, d_offs, d_sib, d_bad)
    torch.cuda.synchronize()
    assert not d_bad.cpu().numpy().any()
    want = elems(s for row in sibs for s in row).reshape(n, n_levels, 32)
    assert np.array_equal(d_sib.cpu().numpy(), want)


def test_random_mutations_of_packed_strings(engine):
    """Seeded byte-level fuzz of the wire format: header fields, bitmap bytes and lengths of honest packed proofs are
    mutated at random; every string must get exactly what arbo.UnpackSiblings + the verifier (oracle) give - status 7
    with flag 0 and root 0, or the verifier's result on the expanded row."""
    rng = random.Random(2024)
    n_levels = 24
    roots, variants, keys, values = [], [], [], []
    for i in range(160):
        root, sib, key, value = census_proof(rng, n_levels, lo=1, hi=14)
        b = bytearray(osmt.pack_siblings(strip(sib)))
        mut = rng.randrange(8)
        if mut == 0 and len(b):
            b[rng.randrange(min(len(b), 4))] ^= 1 << rng.randrange(8)          # header bit flip
        elif mut == 1 and len(b) > 4:
            b[4 + rng.randrange(min(len(b) - 4, 3))] ^= 1 << rng.randrange(8)  # bitmap bit flip
        elif mut == 2:
            b = b[:rng.randrange(len(b) + 1)]                                  # truncation
        elif mut == 3:
            b += bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 40)))
        elif mut == 4 and len(b) > 36:
            j = rng.randrange(4, len(b))
            b[j] ^= 0xFF                                                       # sibling byte (may go non-canonical)
        elif mut == 5:
            b = bytearray(rng.getrandbits(8) for _ in range(rng.randrange(0, 80)))
        roots.append(root), variants.append(bytes(b)), keys.append(key), values.append(value)
    n = len(variants)
    flags, status, out_roots = engine.smt_verify_packed(elems(roots), variants, n_levels, elems(keys), elems(values),
                                                        want_roots=True)
    seen = set()
    for i, b in enumerate(variants):
        s_or, bad = osmt.assignment_siblings(b, n_levels)
        if bad:
            want = (0, osmt.STATUS_MALFORMED, 0)
        else:
            want = osmt.inclusion_verifier(roots[i], s_or, keys[i], values[i])
        got = (int(flags[i]), int(status[i]), ints(out_roots[i:i + 1])[0])
        assert got[:2] == want[:2] and (want[1] != 0 or got[2] == want[2]), (i, b.hex())
        seen.add(want[:2])
    assert {(1, 0), (0, 0), (0, 7)} <= seen
