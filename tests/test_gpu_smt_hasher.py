"""GPU parity: the tree's hash function as a plug (utils.Hasher, utils/hashers.go:10-37).  With
gcp_ctx_set_smt_hasher(GCP_HASHER_POSEIDON2) the verifier, the processor and Hash1 must give what the literal oracle
gives when its hFn is utils.Poseidon2Hasher (HashPoseidon2Gnark, hash/native/bn254/poseidon2/gnark.go:18-54: node =
ordered (min, max) pair, leaf = (key, value, 1)).  The Poseidon2 permutation itself is parity-unpinned: this
pins the SMT gadgets' use of the plug, not the round keys."""
import contextlib
import random

import numpy as np
import pytest

import gnark_crypto_primitives_b200 as g
from oracle import poseidon2 as op2
from oracle import smt as osmt
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def oracle_hasher_poseidon2():
    """hFn = Poseidon2Hasher inside oracle/smt.py (Hash1 / Hash2 of tree/smt/hash.go call hFn)."""
    h1, h2 = osmt.hash1, osmt.hash2
    osmt.hash1 = lambda key, *values: op2.hash([key, *values, 1])
    osmt.hash2 = lambda l, r: op2.hash([l, r])
    try:
        yield
    finally:
        osmt.hash1, osmt.hash2 = h1, h2


@pytest.fixture()
def engine_p2():
    with g.Engine(0) as eng:
        assert eng.smt_hasher == g.HASHER_POSEIDON
        eng.set_smt_hasher(g.HASHER_POSEIDON2)
        assert eng.smt_hasher == g.HASHER_POSEIDON2
        yield eng


def test_plug_is_per_context_and_checked(engine, engine_p2):
    assert engine.smt_hasher == g.HASHER_POSEIDON                       # the session engine is untouched
    with pytest.raises(g.EngineError):
        engine_p2.set_smt_hasher(7)
    k, v = elems([5, 6]), elems([7, 8]).reshape(2, 1, 32)
    out, st = engine_p2.smt_leaf_hash(k, v)
    assert not st.any() and ints(out) == [op2.hash([5, 7, 1]), op2.hash([6, 8, 1])]
    out0, st0 = engine_p2.smt_leaf_hash(k, np.zeros((2, 0, 32), np.uint8))                    # Hash1(key) = hFn(key, 1): two limbs, ordered
    assert not st0.any() and ints(out0) == [op2.hash([5, 1]), op2.hash([6, 1])]
    with pytest.raises(g.EngineError) as e:
        engine_p2.smt_leaf_hash(k, np.zeros((2, 2, 32), np.uint8))      # four limbs
    assert "need 2 or 3 limbs" in str(e.value)
    out_d, _ = engine.smt_leaf_hash(k, v)
    assert ints(out_d) == [osmt.hash1(5, 7), osmt.hash1(6, 8)] and ints(out_d) != ints(out)


def test_verifier_with_poseidon2_hasher(engine_p2):
    rng = random.Random(3535)
    n_levels = 32
    with oracle_hasher_poseidon2():
        tree = osmt.Tree(n_levels)
        kv = {rng.getrandbits(n_levels): rng.randrange(R) for _ in range(24)}
        for k, v in kv.items():
            tree.add(k, v)
        root = tree.root()
        cases = []
        for k, v in kv.items():
            p = tree.gen_proof(k)
            cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=k, old_value=v, is_old0=0, key=k, value=v, fnc=0))
            cases.append(dict(cases[-1], value=(v + 1) % R, old_value=(v + 1) % R))
        for _ in range(30):
            k = rng.getrandbits(n_levels)
            if k in kv:
                continue
            p = tree.gen_proof(k)
            cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                              is_old0=p["is_old0"], key=k, value=0, fnc=1))
        cases.append(dict(cases[0], enabled=0, root=9))
        want = [osmt.verifier(c["enabled"], c["root"], c["siblings"], c["old_key"], c["old_value"], c["is_old0"], c["key"],
                              c["value"], c["fnc"]) for c in cases]
    n = len(cases)
    sib = elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32)
    flags, status, roots = engine_p2.smt_verify(
        elems(c["root"] for c in cases), sib, elems(c["key"] for c in cases), elems(c["value"] for c in cases),
        old_keys=elems(c["old_key"] for c in cases), old_values=elems(c["old_value"] for c in cases),
        is_old0=np.array([c["is_old0"] for c in cases], np.uint8), fnc=np.array([c["fnc"] for c in cases], np.uint8),
        enabled=np.array([c["enabled"] for c in cases], np.uint8), want_roots=True)
    got = list(zip([int(f) for f in flags], [int(s) for s in status], ints(roots)))
    assert got == want
    assert sum(w[0] for w in want) >= 24 + 20 and any(w[0] == 0 for w in want)
    # the same proofs under the default plug do not verify: the two hashers build different trees
    with g.Engine(0) as eng0:
        f0, s0 = eng0.smt_verify(elems(c["root"] for c in cases[:8]), sib[:8], elems(c["key"] for c in cases[:8]),
                                 elems(c["value"] for c in cases[:8]))
    assert not f0.any() and not s0.any()


@pytest.mark.parametrize("big", [False, True])
def test_processor_with_poseidon2_hasher(engine_p2, big):
    """Inserts / updates / deletes of a growing Poseidon2 tree; `big` tiles the batch past 1024 transitions so that the
    scan / sort / prep / path pipeline runs, the small batch takes the thread-per-transition kernel."""
    rng = random.Random(3636)
    n_levels = 28
    with oracle_hasher_poseidon2():
        tree = osmt.Tree(n_levels)
        cases = []
        present = {}
        for step in range(40):
            if present and step % 4 == 3:
                k = rng.choice(list(present))
                v = rng.randrange(R)
                p = tree.gen_proof(k)
                c = dict(old_root=tree.root(), siblings=p["siblings"], old_key=k, old_value=present[k], is_old0=0, new_key=k,
                         new_value=v, fnc0=0, fnc1=1)
            else:
                k = rng.getrandbits(n_levels)
                while k in present:
                    k = rng.getrandbits(n_levels)
                v = rng.randrange(R)
                p = tree.gen_proof(k)
                c = dict(old_root=tree.root(), siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                         is_old0=p["is_old0"], new_key=k, new_value=v, fnc0=1, fnc1=0)
            tree.add(k, v)
            present[k] = v
            c["new_root"] = tree.root()
            cases.append(c)
            if c["fnc0"] == 1 and step % 3 == 0:
                cases.append(dict(c, old_root=c["new_root"], fnc1=1, new_root=c["old_root"]))     # the delete that undoes it
        cases.append(dict(cases[0], old_root=(cases[0]["old_root"] + 1) % R, new_root=0))
        want = [osmt.processor(c["old_root"], c["siblings"], c["old_key"], c["old_value"], c["is_old0"], c["new_key"],
                               c["new_value"], c["fnc0"], c["fnc1"]) for c in cases]
    assert [w[0] for w in want[:-1]] == [c["new_root"] for c in cases[:-1]] and want[-1] == (0, osmt.STATUS_ASSERTION)
    order = list(range(len(cases)))
    if big:
        order = [rng.randrange(len(cases)) for _ in range(1200)]
    cs = [cases[i] for i in order]
    n = len(cs)
    out, st = engine_p2.smt_process(
        elems(c["old_root"] for c in cs), elems([s for c in cs for s in c["siblings"]]).reshape(n, n_levels, 32),
        elems(c["old_key"] for c in cs), elems(c["old_value"] for c in cs), np.array([c["is_old0"] for c in cs], np.uint8),
        elems(c["new_key"] for c in cs), elems(c["new_value"] for c in cs), np.array([c["fnc0"] for c in cs], np.uint8),
        np.array([c["fnc1"] for c in cs], np.uint8))
    assert list(zip(ints(out), [int(x) for x in st])) == [want[i] for i in order]
