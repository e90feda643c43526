"""CPU: the C restatement (oracle/c) must agree bit-for-bit with the pinned Python oracle."""
import random

import numpy as np

from oracle import cport
from oracle import edwards as ed
from oracle import elgamal as eg
from oracle import keccak
from oracle import poseidon as pos
from oracle import smt
from oracle.field import R
from tests.util import census_proof, dense_proof, elems, ints


def test_poseidon_all_arities():
    rng = random.Random(1)
    for arity in range(1, 17):
        rows = [[rng.randrange(R) for _ in range(arity)] for _ in range(5)]
        rows[0] = [0] * arity
        rows[1] = [R - 1] * arity
        out, st = cport.poseidon_hash(elems([x for r in rows for x in r]).reshape(5, arity, 32), threads=2)
        assert not st.any() and ints(out) == [pos.hash(r) for r in rows]


def test_multihash_and_noncanonical():
    rng = random.Random(2)
    for length in (17, 60, 257):
        rows = [[rng.randrange(R) for _ in range(length)] for _ in range(2)]
        out, st = cport.poseidon_multihash(elems([x for r in rows for x in r]).reshape(2, length, 32))
        assert ints(out) == [pos.multihash(r) for r in rows]
    out, st = cport.poseidon_hash(elems([1, R]).reshape(1, 2, 32))
    assert int(st[0]) == 1


def test_smt_literal_and_early_out_match_python():
    rng = random.Random(3)
    n_levels = 24
    cases = []
    for enabled in (0, 1):
        for fnc in (0, 1):
            for is0 in (0, 1):
                for same in (0, 1):
                    for corrupt in (0, 1, 2):
                        root, sib, key, value = census_proof(rng, n_levels, 3, 12)
                        ok_ = key if same else rng.getrandbits(n_levels)
                        ov_ = value if same else rng.randrange(R)
                        if corrupt == 1:
                            root = (root + 1) % R
                        if corrupt == 2:
                            sib = list(sib)
                            sib[-1] = 9
                        cases.append((enabled, root, sib, ok_, ov_, is0, key, value, fnc))
    root, sib, key, value = dense_proof(rng, n_levels)
    cases += [(1, root, sib, key, value, 0, key | (1 << n_levels), value, 0), (1, root, sib, key, value, 2, key, value, 0),
              (1, root, sib, key, value, 0, key, R, 0), (1, root, sib, key, value, 0, key, value, 0)]
    want = [smt.verifier(*c) for c in cases]
    n = len(cases)
    for literal in (True, False):
        flags, st, roots = cport.smt_verify(
            elems(c[1] for c in cases), elems([s for c in cases for s in c[2]]).reshape(n, n_levels, 32),
            elems(c[6] for c in cases), elems(c[7] for c in cases), old_keys=elems(c[3] for c in cases),
            old_values=elems(c[4] for c in cases), is_old0=np.array([c[5] for c in cases], np.uint8),
            fnc=np.array([c[8] for c in cases], np.uint8), enabled=np.array([c[0] for c in cases], np.uint8),
            literal=literal, threads=4)
        got = list(zip([int(f) for f in flags], [int(s) for s in st], ints(roots)))
        assert got == want


def test_elgamal_matches_python():
    rng = random.Random(4)
    pk = ed.scalar_mul(ed.G, 0xB200)
    ks = [12345, R - 1, 0, 1, rng.randrange(R), rng.randrange(R)]
    ms = [67890, 0, 5, R - 2, rng.randrange(1 << 16), rng.randrange(1 << 16)]
    out, st = cport.elgamal_encrypt(elems(pk), elems(ks), elems(ms), threads=3)
    assert not st.any()
    want = [eg.serialize(eg.encrypt(pk, k, m)) for k, m in zip(ks, ms)]
    assert [ints(o) for o in out] == want
    # per-item keys, one off-curve
    pks = [ed.scalar_mul(ed.G, 7 + i) for i in range(len(ks))]
    pks[2] = (1, 2)
    out2, st2 = cport.elgamal_encrypt(elems([c for p in pks for c in p]), elems(ks), elems(ms))
    assert [int(s) for s in st2] == [0, 0, 4, 0, 0, 0]
    assert ints(out2[1]) == eg.serialize(eg.encrypt(pks[1], ks[1], ms[1]))
    # add / tally
    summed, st3 = cport.elgamal_add(out[:3], out[3:6])
    for i in range(3):
        a = ((want[i][0], want[i][1]), (want[i][2], want[i][3]))
        b = ((want[i + 3][0], want[i + 3][1]), (want[i + 3][2], want[i + 3][3]))
        assert ints(summed[i]) == eg.serialize(eg.ct_add(a, b))
    tal, st4 = cport.elgamal_tally(out.reshape(3, 2, 4, 32))
    cts = [((w[0], w[1]), (w[2], w[3])) for w in want]
    assert ints(tal[0]) == eg.serialize(eg.tally([cts[0], cts[2], cts[4]]))
    assert ints(tal[1]) == eg.serialize(eg.tally([cts[1], cts[3], cts[5]]))
    fb = cport.fixed_base_mul(elems(ks))
    assert [tuple(ints(p)) for p in fb] == [eg.fixed_base_scalar_mul(k) for k in ks]


def test_keccak_address():
    rng = random.Random(5)
    data = bytes(rng.getrandbits(8) for _ in range(64 * 9))
    out = cport.keccak_address(np.frombuffer(data, np.uint8), threads=2)
    for i in range(9):
        assert out[i].tobytes() == keccak.derive_address(data[64 * i:64 * i + 64])
