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


def test_host_entry_points_keep_their_scratch_for_the_whole_call(engine):
    """gcp_mimc7_hash and gcp_smt_process used to drop the context lock between upload, kernel and read-back, so two
    threads on one context could overwrite (or, on a re-allocation, free) each other's device buffers."""
    import numpy as np

    from oracle import mimc7 as omimc
    from oracle import smt as osmt

    rng = random.Random(123)
    batches = []
    for size in (7, 300, 4000):                     # different sizes: the scratch slots are re-allocated across calls
        rows = [[rng.randrange(R) for _ in range(3)] for _ in range(size)]
        batches.append((elems([x for r in rows for x in r]).reshape(size, 3, 32), rows))
    want_m = [[omimc.hash(r) for r in rows[:5]] for _, rows in batches]

    n_levels = 24
    tree = osmt.Tree(n_levels)
    cases = []
    for _ in range(24):
        k, v = rng.getrandbits(n_levels), rng.randrange(R)
        p = tree.gen_proof(k)
        if p["exists"]:
            continue
        cases.append(dict(old_root=tree.root(), siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                          is_old0=p["is_old0"], new_key=k, new_value=v, fnc0=1, fnc1=0))
        tree.add(k, v)
    n = len(cases)
    pargs = (elems(c["old_root"] for c in cases), elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32),
             elems(c["old_key"] for c in cases), elems(c["old_value"] for c in cases),
             np.array([c["is_old0"] for c in cases], np.uint8), elems(c["new_key"] for c in cases),
             elems(c["new_value"] for c in cases), np.array([c["fnc0"] for c in cases], np.uint8),
             np.array([c["fnc1"] for c in cases], np.uint8))
    want_roots = [c["old_root"] for c in cases[1:]] + [tree.root()]
    errors = []

    def worker(kind):
        try:
            for it in range(12):
                if kind < 3:
                    arr, _ = batches[(kind + it) % 3]
                    out, st = engine.mimc7_hash(arr)
                    assert not st.any() and ints(out[:5]) == want_m[(kind + it) % 3]
                else:
                    roots, st = engine.smt_process(*pargs)
                    assert not st.any() and ints(roots) == want_roots
        except Exception as exc:
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i % 4,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]


def test_contexts_share_the_copy_pool_from_several_threads():
    """Three contexts on one device driven by three host threads with PAGEABLE inputs of several MB and pageable outputs:
    every upload goes through the process-wide copy pool (csrc/hostcopy.h) and the page-locked staging ring, every result
    through the page-locked output arena / ring (d2h_copy).  Results must equal the single-threaded ones, call after call."""
    import numpy as np

    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import _lib

    rng = np.random.default_rng(5)
    n = 1 << 16
    hin = rng.integers(0, 256, size=(n, 2, 32), dtype=np.uint8)       # 4 MB: staged (>= 1 MB), out 2 MB: arena
    hin[:, :, 31] &= 0x0F
    big = rng.integers(0, 256, size=(3 * n, 4, 32), dtype=np.uint8)   # 25 MB in, 25 MB out: the output ring
    big[:, :, 31] &= 0x0F
    with g.Engine(0) as ref:
        want_hash, _ = ref.poseidon_hash(hin)
        want_neg, st = ref.elgamal_neg(big)
        assert not st.any()
    assert _lib.load().gcp_copy_threads() >= 1
    errors = []

    def worker(i):
        try:
            with g.Engine(0) as eng:
                for it in range(4):
                    if (i + it) % 2:
                        out, st = eng.poseidon_hash(hin)
                        assert not st.any() and (out == want_hash).all()
                    else:
                        out, st = eng.elgamal_neg(big)
                        assert not st.any() and (out == want_neg).all()
        except Exception as exc:
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]
