"""GPU: one context shared by several host threads (Go goroutines call in concurrently; the ABI serialises per ctx)."""
import random
import threading

import pytest

from oracle import poseidon as opos
from oracle.field import R
from tests.util import dense_proof, elems, ints

pytestmark = pytest.mark.gpu


def test_concurrent_calls_on_one_context(engine):
    rng = random.Random(77)
    rows = [[rng.randrange(R), rng.randrange(R)] for _ in range(64)]
    want_hash = [opos.hash(r) for r in rows]
    hin = elems([x for r in rows for x in r]).reshape(64, 2, 32)
    items = [dense_proof(rng, 24) for _ in range(40)]
    sib = elems([s for it in items for s in it[1]]).reshape(40, 24, 32)
    roots, keys, vals = elems(it[0] for it in items), elems(it[2] for it in items), elems(it[3] for it in items)
    errors = []

    def worker(kind):
        try:
            for _ in range(10):
                if kind == 0:
                    out, st = engine.poseidon_hash(hin)
                    assert not st.any() and ints(out) == want_hash
                else:
                    flags, st = engine.smt_verify_inclusion(roots, sib, keys, vals)
                    assert flags.all() and not st.any()
        except Exception as exc:  # surfaced below
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i % 2,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]


def test_two_contexts_on_one_device():
    import gnark_crypto_primitives_b200 as g

    with g.Engine(0) as a, g.Engine(0) as b:
        x = elems([1, 2]).reshape(1, 2, 32)
        assert ints(a.poseidon_hash(x)[0]) == ints(b.poseidon_hash(x)[0]) == [opos.hash([1, 2])]
