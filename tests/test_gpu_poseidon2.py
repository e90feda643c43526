"""GPU parity: width-2 Poseidon2 hasher (hash/native/bn254/poseidon2/native.go:30-63, gnark.go:18-54) through the
C ABI vs oracle/poseidon2.py.  The permutation's round keys are un-vendored (parity unpinned, see the oracle's header);
the key-independent properties - min/max ordering, mod-r reduction, the chaining rule, installable keys - are checked
here with BOTH the default keys and a random key set."""
import random
import threading

import numpy as np
import pytest

from oracle import poseidon2 as op2
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu
M = 1 << 256


def _hash(engine, rows, **kw):
    n, ln = len(rows), len(rows[0])
    out, st = engine.poseidon2_hash(elems([x for r in rows for x in r]).reshape(n, ln, 32), **kw)
    return ints(out), st


def test_permutation_matches_oracle(engine):
    rng = random.Random(1)
    states = [[0, 0], [0, 1], [R - 1, R - 1]] + [[rng.randrange(R), rng.randrange(R)] for _ in range(200)]
    out, st = engine.poseidon2_permutation(elems([x for s in states for x in s]).reshape(len(states), 2, 32))
    assert not st.any()
    got = ints(out)
    assert [got[2 * i:2 * i + 2] for i in range(len(states))] == [op2.permutation(s) for s in states]


@pytest.mark.parametrize("length", [2, 3])
def test_hash_matches_oracle(engine, length):
    rng = random.Random(length)
    rows = [[rng.randrange(R) for _ in range(length)] for _ in range(300)]
    rows[0] = [0] * length
    rows[1] = [R - 1] * length
    rows[2] = [5] * length                      # equal limbs: order irrelevant
    rows[3] = ([7, 3] + [1])[:length]           # descending pair: swapped when length == 2
    got, st = _hash(engine, rows)
    assert not st.any()
    assert got == [op2.hash(r) for r in rows]


def test_internal_node_is_order_independent_and_leaf_is_not(engine):
    rng = random.Random(7)
    pairs = [[rng.randrange(R), rng.randrange(R)] for _ in range(64)]
    a, _ = _hash(engine, pairs)
    b, _ = _hash(engine, [[y, x] for x, y in pairs])
    assert a == b                                # native.go:42-44 / gnark.go:24-36
    la, _ = _hash(engine, [[x, y, 1] for x, y in pairs])
    lb, _ = _hash(engine, [[y, x, 1] for x, y in pairs])
    assert all(p != q for p, q in zip(la, lb))   # leaves keep (key, value, flag) order


def test_canonical_inputs_are_reduced_mod_r_like_safe_big_int(engine):
    rows = [[R + 5, 3], [M - 1, R], [2 * R + 1, 4 * R + 9]]
    got, st = _hash(engine, rows)
    assert not st.any()                          # native.go:37-39: SafeBigInt reduces, no error
    assert got == [op2.hash([x % R for x in r]) for r in rows]
    # the order is decided AFTER the reduction: (R + 5, 3) -> (5, 3) -> (3, 5)
    assert got[0] == op2.hash([3, 5])


def test_montgomery_format(engine):
    import gnark_crypto_primitives_b200 as g

    rng = random.Random(11)
    rows = [[rng.randrange(R) for _ in range(2)] for _ in range(40)] + [[9, 2], [2, 9]]
    got, st = _hash(engine, [[x * M % R for x in r] for r in rows], fmt=g.FMT_MONTGOMERY)
    assert not st.any()
    assert [x * pow(M, -1, R) % R for x in got] == [op2.hash(r) for r in rows]   # ordered by VALUE, not by limb image
    bad, st = _hash(engine, [[R, 1]], fmt=g.FMT_MONTGOMERY)                         # fr.Element invariant broken
    assert int(st[0]) == 1 and bad == [0]


def test_arity_and_empty(engine):
    import gnark_crypto_primitives_b200 as g

    for ln in (1, 4):
        with pytest.raises(g.EngineError, match="need 2 or 3 limbs"):
            engine.poseidon2_hash(np.zeros((1, ln, 32), np.uint8))
    out, st = engine.poseidon2_hash(np.zeros((0, 2, 32), np.uint8))
    assert out.shape == (0, 32) and st.shape == (0,)


def test_installable_round_keys():
    """A context takes gnark-crypto's own keys as data: with a random key set the engine follows the oracle run with
    the same keys, and a second context keeps the default ones."""
    import gnark_crypto_primitives_b200 as g

    rng = random.Random(99)
    flat = [rng.randrange(R) for _ in range(op2.N_KEYS)]
    keys = op2.unflatten(flat)
    rows = [[rng.randrange(R) for _ in range(3)] for _ in range(20)]
    a, b = g.Engine(0), g.Engine(0)
    try:
        a.poseidon2_set_round_keys(elems(flat))
        got, st = _hash(a, rows)
        assert not st.any() and got == [op2.hash(r, keys) for r in rows]
        dflt, _ = _hash(b, rows)
        assert dflt == [op2.hash(r) for r in rows]
        a.poseidon2_set_round_keys(elems([k * M % R for k in op2.flat_round_keys()]), fmt=g.FMT_MONTGOMERY)
        back, _ = _hash(a, rows)
        assert back == dflt
        with pytest.raises(g.EngineError, match="62 round keys"):
            a.poseidon2_set_round_keys(elems(flat[:61]))
        with pytest.raises(g.EngineError, match=">= r"):
            a.poseidon2_set_round_keys(elems(flat[:61] + [R]))
    finally:
        a.close()
        b.close()


def test_merkle_path_with_poseidon2_nodes(engine):
    """The shape the reference uses it in (tree/test/poseidon2_test.go): leaf = H(key, value, 1), nodes = H(min, max)."""
    rng = random.Random(5)
    n, depth = 32, 6
    leaves = [[rng.randrange(R), rng.randrange(R), 1] for _ in range(n)]
    acc, _ = _hash(engine, leaves)
    want = [op2.hash(r) for r in leaves]
    for _ in range(depth):
        sib = [rng.randrange(R) for _ in range(n)]
        acc, st = _hash(engine, [[a, s] for a, s in zip(acc, sib)])
        want = [op2.hash([a, s]) for a, s in zip(want, sib)]
        assert not st.any()
    assert acc == want


def test_large_batch_two_streams_and_threads(engine):
    """> 2^20 items (two chunks on the two streams) from two host threads at once: per-call scratch is not shared."""
    n = (1 << 20) + 1000
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(n, 2, 32), dtype=np.uint8)
    a[:, :, 31] &= 0x1F                                    # < 2^253 < r
    ref, st = engine.poseidon2_hash(a)
    assert not st.any()
    idx = [0, 1, (1 << 20) - 1, 1 << 20, n - 1]
    assert [ints(ref[i])[0] for i in idx] == [op2.hash(ints(a[i])) for i in idx]
    res = {}

    def work(tag, arr):
        res[tag] = engine.poseidon2_hash(arr)[0]

    b = a[::-1].copy()
    ts = [threading.Thread(target=work, args=("a", a)), threading.Thread(target=work, args=("b", b))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert np.array_equal(res["a"], ref) and np.array_equal(res["b"], ref[::-1])
