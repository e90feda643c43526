"""GPU parity: Poseidon through the C ABI vs the oracle (bit-exact)."""
import random

import numpy as np
import pytest

from oracle import poseidon as opos
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu

KAT = {  # public circomlib vectors
    (1,): 18586133768512220936620570745912940619677854269274689475585506675881198879027,
    (1, 2): 7853200120776062878684798364095072458815029376092732009249414926327459813530,
    (1, 2, 3): 6542985608222806190361240322586112750744169038454362455181422643027100751666,
    (1, 2, 3, 4): 18821383157269793795438455681495246036402687001665670618754263018637548127333,
    tuple(range(1, 17)): 9989051620750914585850546081941653841776809718687451684622678807385399211877,
}


def test_known_answers(engine):
    for inp, want in KAT.items():
        out, st = engine.poseidon_hash(elems(inp).reshape(1, len(inp), 32))
        assert st[0] == 0
        assert ints(out)[0] == want, inp
    # reference fixed input, hash/native/bn254/poseidon/poseidon_test.go:39
    out, _ = engine.poseidon_hash(elems([297262668938251460872476410954775437897592223497]).reshape(1, 1, 32))
    assert ints(out)[0] == 21099541378821686330832093407308585959971016892597585818017774528142419287929


@pytest.mark.parametrize("arity", list(range(1, 17)))
def test_hash_matches_oracle_all_arities(engine, arity):
    rng = random.Random(0xB200 + arity)
    n = 64 if arity <= 3 else 16
    rows = [[rng.randrange(R) for _ in range(arity)] for _ in range(n)]
    rows[0] = [0] * arity
    rows[1] = [R - 1] * arity
    rows[2] = list(range(1, arity + 1))
    out, st = engine.poseidon_hash(elems([x for r in rows for x in r]).reshape(n, arity, 32))
    assert not st.any()
    assert ints(out) == [opos.hash(r) for r in rows]


def test_config1_batch_1024_two_inputs(engine):
    """BASELINE config 1: 1024 two-input hashes (the Hash2 shape of the SMT path)."""
    rng = random.Random(0xB200)
    rows = [[rng.randrange(R), rng.randrange(R)] for _ in range(1024)]
    out, st = engine.poseidon_hash(elems([x for r in rows for x in r]).reshape(1024, 2, 32))
    assert not st.any()
    assert ints(out) == [opos.hash(r) for r in rows]


def test_noncanonical_input_sets_status(engine):
    rows = [[1, 2], [R, 5], [3, 2**256 - 1], [7, 8]]
    out, st = engine.poseidon_hash(elems([x for r in rows for x in r]).reshape(4, 2, 32))
    assert list(st) == [0, 1, 1, 0]
    got = ints(out)
    assert got[0] == opos.hash([1, 2]) and got[3] == opos.hash([7, 8])
    assert got[1] == 0 and got[2] == 0


def test_bad_arity_is_an_error(engine):
    import gnark_crypto_primitives_b200 as g

    with pytest.raises(g.EngineError) as e:
        engine.poseidon_hash(np.zeros((2, 17, 32), dtype=np.uint8))
    assert "bad inputs provided" in str(e.value)  # poseidon.go:41-43
    with pytest.raises(g.EngineError):
        engine.poseidon_multihash(np.zeros((1, 4097, 32), dtype=np.uint8))
    out, st = engine.poseidon_hash(np.zeros((0, 2, 32), dtype=np.uint8))
    assert out.shape == (0, 32)


@pytest.mark.parametrize("length", [1, 16, 17, 32, 33, 60, 255, 256, 257, 300])
def test_multihash_matches_oracle(engine, length):
    rng = random.Random(length)
    n = 3
    rows = [[rng.randrange(R) for _ in range(length)] for _ in range(n)]
    rows[0] = list(range(1, length + 1))
    out, st = engine.poseidon_multihash(elems([x for r in rows for x in r]).reshape(n, length, 32))
    assert not st.any()
    assert ints(out) == [opos.multihash(r) for r in rows]


def test_multihash_1_to_60_regression(engine):
    # inputs of hash/emulated/bn254/poseidon/poseidon_test.go:85-88
    out, _ = engine.poseidon_multihash(elems(range(1, 61)).reshape(1, 60, 32))
    assert ints(out)[0] == 10383247944466245790564312669548703973436539043614368627989218605970689057797


def test_multihash_4096(engine):
    rng = random.Random(4096)
    row = [rng.randrange(R) for _ in range(4096)]
    out, st = engine.poseidon_multihash(elems(row).reshape(1, 4096, 32))
    assert ints(out)[0] == opos.multihash(row)


def test_montgomery_format_roundtrip(engine):
    """GCP_FMT_MONTGOMERY takes gnark-crypto fr.Element memory (x * 2^256 mod r) and returns the same form."""
    import gnark_crypto_primitives_b200 as g

    rng = random.Random(5)
    rows = [[rng.randrange(R), rng.randrange(R)] for _ in range(33)]
    mont = [[x * (1 << 256) % R for x in r] for r in rows]
    out, st = engine.poseidon_hash(elems([x for r in mont for x in r]).reshape(33, 2, 32), fmt=g.FMT_MONTGOMERY)
    assert not st.any()
    rinv = pow(1 << 256, -1, R)
    assert [v * rinv % R for v in ints(out)] == [opos.hash(r) for r in rows]


def _edge_values():
    """Field elements with carry-hostile limb patterns: runs of ones, single bits, r - small, limb boundaries."""
    vals = {0, 1, 2, R - 1, R - 2, R >> 1, (R >> 1) + 1}
    for k in range(0, 254):
        vals.add((1 << k) % R)
        vals.add(((1 << k) - 1) % R)
        vals.add((R - (1 << k)) % R)
    for lo in range(0, 256, 32):
        for hi in range(lo + 32, 257, 32):
            v = ((1 << hi) - 1) ^ ((1 << lo) - 1)          # limbs lo/32 .. hi/32 - 1 all ones
            vals.add(v % R)
            vals.add((v >> 3) % R)
    for limb in range(8):
        vals.add((0xFFFFFFFF << (32 * limb)) % R)
        vals.add((0x80000000 << (32 * limb)) % R)
        vals.add((0x00000001 << (32 * limb)) % R)
    return sorted(vals)


def test_edge_value_patterns_against_c_oracle(engine):
    from oracle import cport

    vals = _edge_values()
    assert len(vals) > 800
    # arity 1: every edge value alone; arity 2: consecutive pairs and (x, x) squares of the same value
    a1 = elems(vals).reshape(len(vals), 1, 32)
    out, st = engine.poseidon_hash(a1)
    want, wst = cport.poseidon_hash(a1, threads=8)
    assert not st.any() and (out == want).all()
    pairs = [(vals[i], vals[(i * 7 + 3) % len(vals)]) for i in range(len(vals))] + [(v, v) for v in vals]
    a2 = elems([x for p in pairs for x in p]).reshape(len(pairs), 2, 32)
    out, st = engine.poseidon_hash(a2)
    want, wst = cport.poseidon_hash(a2, threads=8)
    assert not st.any() and (out == want).all()
    for i in (0, 5, 100, len(vals) - 1):
        assert ints(out[i:i + 1])[0] == opos.hash(list(pairs[i]))


@pytest.mark.parametrize("arity", [5, 10, 11, 12, 16])
def test_wide_arities_edge_patterns_and_bulk_against_c_oracle(engine, arity):
    """The generic kernel takes a whole matrix row (up to 17 products) into one Montgomery reduction; t = 11 is the
    widest row on a non-canonical (< 2r) state, t = 12..17 keep the state canonical (poseidon.cuh bound note).
    Carry-hostile inputs (r-1, all-ones limbs, single bits) in every position plus 4096 random rows, all compared."""
    from oracle import cport

    vals = _edge_values()
    rng = random.Random(arity)
    rows = [[vals[(i * 13 + j * 7) % len(vals)] for j in range(arity)] for i in range(len(vals))]
    rows += [[R - 1] * arity, [R - 2] * arity, [0] * arity, [1] * arity, [(1 << 253) - 1] * arity]
    rows += [[rng.randrange(R) for _ in range(arity)] for _ in range(4096)]
    a = elems([x for row in rows for x in row]).reshape(len(rows), arity, 32)
    out, st = engine.poseidon_hash(a)
    want, wst = cport.poseidon_hash(a, threads=cport.default_threads())
    assert not st.any() and not wst.any()
    assert (out == want).all()
    for i in (0, len(vals), len(vals) + 1):
        assert ints(out[i:i + 1])[0] == opos.hash(rows[i])


def test_full_compare_2pow18_hashes(engine):
    """Every one of 2^18 random two-input hashes against the C oracle (not a sample)."""
    from oracle import cport

    rng = np.random.default_rng(20260)
    n = 1 << 18
    a = rng.integers(0, 256, size=(n, 2, 32), dtype=np.uint8)
    a[:, :, 31] &= 0x1F                                     # < 2^253 < r
    out, st = engine.poseidon_hash(a)
    want, wst = cport.poseidon_hash(a, threads=cport.default_threads())
    assert not st.any() and not wst.any()
    assert (out == want).all()
