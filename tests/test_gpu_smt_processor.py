"""GPU parity: smt.Processor (insert / update / delete / nop) through the C ABI vs the literal oracle."""
import random

import numpy as np
import pytest

from oracle import smt as osmt
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu


def run(engine, cases, n_levels):
    n = len(cases)
    out, st = engine.smt_process(
        elems(c["old_root"] for c in cases), elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32),
        elems(c["old_key"] for c in cases), elems(c["old_value"] for c in cases),
        np.array([c["is_old0"] for c in cases], np.uint8), elems(c["new_key"] for c in cases),
        elems(c["new_value"] for c in cases), np.array([c["fnc0"] for c in cases], np.uint8),
        np.array([c["fnc1"] for c in cases], np.uint8))
    want = [osmt.processor(c["old_root"], c["siblings"], c["old_key"], c["old_value"], c["is_old0"], c["new_key"],
                           c["new_value"], c["fnc0"], c["fnc1"]) for c in cases]
    return ints(out), [int(s) for s in st], want


@pytest.mark.parametrize("n_levels", [8, 32, 160])
def test_insert_update_delete_on_a_growing_tree(engine, n_levels):
    rng = random.Random(n_levels)
    tree = osmt.Tree(n_levels)
    cases, expect_roots = [], []
    keys = []
    for step in range(40):
        k = rng.getrandbits(n_levels)
        while k in keys:
            k = rng.getrandbits(n_levels)
        v = rng.randrange(R)
        old_root = tree.root()
        p = tree.gen_proof(k)
        tree.add(k, v)
        keys.append(k)
        base = dict(old_root=old_root, siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                    is_old0=p["is_old0"], new_key=k, new_value=v)
        cases.append(dict(base, fnc0=1, fnc1=0))                                   # insert
        expect_roots.append(tree.root())
        cases.append(dict(base, old_root=tree.root(), fnc0=1, fnc1=1))             # delete = mirror of insert
        expect_roots.append(old_root)
        cases.append(dict(base, fnc0=0, fnc1=0))                                   # nop
        expect_roots.append(old_root)
    for k in keys[:12]:                                                            # updates
        old_root = tree.root()
        p = tree.gen_proof(k)
        v2 = rng.randrange(R)
        tree.add(k, v2)
        cases.append(dict(old_root=old_root, siblings=p["siblings"], old_key=k, old_value=p["old_value"], is_old0=0,
                          new_key=k, new_value=v2, fnc0=0, fnc1=1))
        expect_roots.append(tree.root())
    roots, status, want = run(engine, cases, n_levels)
    assert [(r, s) for r, s in zip(roots, status)] == want
    if n_levels >= 32:   # with 8-bit keys some leaves sit at depth n_levels, which LevIns rejects (status 6) by design
        assert status == [0] * len(cases)
        assert roots == expect_roots
    else:
        assert all(r == e for r, e, s in zip(roots, expect_roots, status) if s == 0) and status.count(0) > 100


def test_assertion_failures(engine):
    rng = random.Random(9)
    n_levels = 16
    tree = osmt.Tree(n_levels)
    for _ in range(6):
        tree.add(rng.getrandbits(n_levels), rng.randrange(R))
    k = rng.getrandbits(n_levels)
    p = tree.gen_proof(k)
    assert not p["exists"]
    good = dict(old_root=tree.root(), siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                is_old0=p["is_old0"], new_key=k, new_value=5, fnc0=1, fnc1=0)
    existing = tree.gen_proof(p["old_key"]) if not p["is_old0"] else None
    cases = [
        dict(good),
        dict(good, old_root=(good["old_root"] + 1) % R),                 # old root does not match the proof
        dict(good, is_old0=2),                                           # processor_test.go:60-61
        dict(good, fnc0=2),
        dict(good, new_key=k | (1 << n_levels)),                         # lowBits
        dict(good, siblings=list(p["siblings"][:-1]) + [7]),             # LevIns: siblings[n-1] != 0
        dict(good, new_value=R),
        dict(good, fnc0=0, fnc1=1),                                      # update with a different key
    ]
    if existing is not None:
        ek = p["old_key"]
        cases.append(dict(old_root=tree.root(), siblings=existing["siblings"], old_key=ek, old_value=existing["old_value"],
                          is_old0=0, new_key=ek, new_value=9, fnc0=1, fnc1=0))   # inserting an existing key
    roots, status, want = run(engine, cases, n_levels)
    assert [(r, s) for r, s in zip(roots, status)] == want
    assert status[0] == 0 and status[1] == 6 and status[2] == 3 and status[3] == 3 and status[4] == 2
    assert status[5] == 6 and status[6] == 1 and status[7] == 6
    if existing is not None:
        assert status[8] == 6
    # reference's own valid case (processor_test.go:46-57): everything zero, nop
    z = dict(old_root=0, siblings=[0, 0, 0, 0], old_key=0, old_value=0, is_old0=0, new_key=0, new_value=0, fnc0=0, fnc1=0)
    roots, status, want = run(engine, [z, dict(z, is_old0=2)], 4)
    assert (roots, status) == ([0, 0], [0, 3]) and want == [(0, 0), (0, 3)]


def test_packed_proofs_give_the_dense_results(engine):
    """gcp_smt_process_packed: the proofs as arbo's GenProof returns them (wrapper_arbo.go:166-179 unpacks on the CPU)."""
    rng = random.Random(166)
    n_levels = 64
    tree = osmt.Tree(n_levels)
    cases, packed = [], []
    for step in range(30):
        k = rng.getrandbits(n_levels)
        v = rng.randrange(R)
        old_root = tree.root()
        p = tree.gen_proof(k)
        packed.append(tree.last_packed)
        tree.add(k, v)
        cases.append(dict(old_root=old_root, siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                          is_old0=p["is_old0"], new_key=k, new_value=v, fnc0=1, fnc1=0))
    cases.append(dict(cases[3]))
    packed.append(packed[3][:-2])                                   # truncated string: arbo.UnpackSiblings errors
    n = len(cases)
    dense_roots, dense_status, want = run(engine, cases, n_levels)
    out, st = engine.smt_process_packed(
        elems(c["old_root"] for c in cases), packed, n_levels, elems(c["old_key"] for c in cases),
        elems(c["old_value"] for c in cases), np.array([c["is_old0"] for c in cases], np.uint8),
        elems(c["new_key"] for c in cases), elems(c["new_value"] for c in cases),
        np.array([c["fnc0"] for c in cases], np.uint8), np.array([c["fnc1"] for c in cases], np.uint8))
    got_roots, got_status = ints(out), [int(s) for s in st]
    assert got_roots[:n - 1] == dense_roots[:n - 1] and got_status[:n - 1] == dense_status[:n - 1]
    assert [(r, s) for r, s in zip(got_roots[:n - 1], got_status[:n - 1])] == want[:n - 1]
    assert got_status[-1] == osmt.STATUS_MALFORMED and got_roots[-1] == 0
    assert all(s == 0 for s in got_status[:n - 1])


def test_arbo_post_insert_proofs(engine):
    """gcp_smt_process_arbo: the reference's own flow (WrapperArbo.addOrUpdate, wrapper_arbo.go:119-185) generates the
    proof AFTER the add and drops the last unpacked sibling when isOld0 == 0 and fnc1 == 0.  The transitions must solve
    (status 0) and give the tree's new root; the same strings through gcp_smt_process_packed (siblings as they are) fail
    the old-root assertion for every insert beside an existing leaf."""
    rng = random.Random(170)
    n_levels = 48
    tree = osmt.Tree(n_levels)
    keys = [rng.getrandbits(n_levels) for _ in range(36)]
    keys += [keys[3], keys[17], keys[17]]                            # updates of existing keys
    asg = [osmt.arbo_add_or_update(tree, k, rng.randrange(R)) for k in keys]
    assert all(a["status"] == 0 for a in asg)
    beside = [i for i, a in enumerate(asg) if a["fnc0"] == 1 and a["is_old0"] == 0]
    assert len(beside) >= 10 and any(a["fnc1"] == 1 for a in asg) and any(a["is_old0"] == 1 for a in asg)
    want = [osmt.processor(a["old_root"], a["siblings"], a["old_key"], a["old_value"], a["is_old0"], a["new_key"],
                           a["new_value"], a["fnc0"], a["fnc1"]) for a in asg]
    assert [w[0] for w in want] == [a["new_root"] for a in asg] and all(w[1] == 0 for w in want)
    args = (elems(a["old_root"] for a in asg), [a["packed"] for a in asg], n_levels, elems(a["old_key"] for a in asg),
            elems(a["old_value"] for a in asg), np.array([a["is_old0"] for a in asg], np.uint8),
            elems(a["new_key"] for a in asg), elems(a["new_value"] for a in asg),
            np.array([a["fnc0"] for a in asg], np.uint8), np.array([a["fnc1"] for a in asg], np.uint8))
    out, st = engine.smt_process_arbo(*args)
    assert [int(x) for x in st] == [0] * len(asg) and ints(out) == [a["new_root"] for a in asg]
    out2, st2 = engine.smt_process_packed(*args)
    st2 = [int(x) for x in st2]
    assert all(st2[i] == osmt.STATUS_ASSERTION for i in beside)
    assert all(st2[i] == 0 for i in range(len(asg)) if i not in beside)
    # a string with nothing to drop: the reference panics on siblingsUnpacked[0:-1]
    empty = osmt.pack_siblings([])
    a = asg[beside[0]]
    out3, st3 = engine.smt_process_arbo(elems([a["old_root"]]), [empty], n_levels, elems([a["old_key"]]),
                                        elems([a["old_value"]]), np.array([0], np.uint8), elems([a["new_key"]]),
                                        elems([a["new_value"]]), np.array([1], np.uint8), np.array([0], np.uint8))
    assert int(st3[0]) == osmt.STATUS_MALFORMED and ints(out3) == [0]


def differential_cases(seed, n_levels, steps=60):
    """Transitions of a growing tree, each with a random function code and a random mutation (selectors outside {0,1},
    elements at and past r, keys past 2^n, wrong old roots, swapped or corrupted siblings, siblings[n-1] != 0)."""
    rng = random.Random(4000 + seed)
    tree = osmt.Tree(n_levels)
    cases = []
    for step in range(steps):
        k = rng.getrandbits(n_levels)
        v = rng.randrange(R)
        p = tree.gen_proof(k)
        c = dict(old_root=tree.root(), siblings=list(p["siblings"]), old_key=p["old_key"], old_value=p["old_value"],
                 is_old0=p["is_old0"], new_key=k, new_value=v, fnc0=rng.choice([0, 1, 1]), fnc1=rng.choice([0, 0, 1]))
        if not p["exists"]:
            tree.add(k, v)
        mut = rng.randrange(10)
        if mut == 0:
            c[rng.choice(["fnc0", "fnc1", "is_old0"])] = rng.choice([2, 7, 255])
        elif mut == 1:
            c[rng.choice(["old_root", "old_value", "new_value", "old_key", "new_key"])] = rng.choice([R, R + 3, 2**256 - 1])
        elif mut == 2:
            c[rng.choice(["old_key", "new_key"])] |= 1 << n_levels
        elif mut == 3:
            c["old_root"] = (c["old_root"] + 1) % R
        elif mut == 4:
            c["siblings"][n_levels - 1] = rng.randrange(1, R)
        elif mut == 5:
            j = rng.randrange(n_levels - 1)
            c["siblings"][j] = rng.choice([R, (c["siblings"][j] + 1) % R, 0])
        elif mut == 6:
            c["is_old0"] ^= 1
        cases.append(c)
    return cases


@pytest.mark.parametrize("seed", [1, 2])
def test_random_differential_with_mutations(engine, seed):
    """Seeded differential run against the literal oracle at an odd level count."""
    n_levels = [19, 45][seed - 1]
    cases = differential_cases(seed, n_levels)
    roots, status, want = run(engine, cases, n_levels)
    for i, w in enumerate(want):
        assert (roots[i], status[i]) == w, (i, cases[i])
    assert len({s for _, s in want}) >= 3


def test_pipeline_form_on_a_large_mixed_batch(engine):
    """From 1024 transitions up the processor runs on the verifier's pipeline (scan, sort by path length, prep kernel,
    two-chain path kernel with staged siblings).  A shuffled batch of 1 500 transitions - honest inserts / updates /
    deletes / nops of a growing tree and every kind of mutated one, path lengths from 0 up - must give exactly the
    oracle's (new root, status) per item, i.e. what the one-thread-per-transition kernel gives on the same items."""
    n_levels = 37
    base = differential_cases(7, n_levels, steps=150)
    # honest transitions of a second tree, incl. deletes (an insert read backwards) and inserts below a shared prefix
    rng = random.Random(99)
    tree = osmt.Tree(n_levels)
    for step in range(100):
        k = rng.getrandbits(n_levels) if step % 5 else (rng.getrandbits(6) | (1 << 20))   # clustered keys: deep splits
        v = rng.randrange(R)
        p = tree.gen_proof(k)
        if p["exists"]:
            continue
        old_root = tree.root()
        tree.add(k, v)
        ins = dict(old_root=old_root, siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                   is_old0=p["is_old0"], new_key=k, new_value=v, fnc0=1, fnc1=0)
        base.append(ins)
        base.append(dict(ins, old_root=tree.root(), fnc1=1))                              # delete
    want_base = [osmt.processor(c["old_root"], c["siblings"], c["old_key"], c["old_value"], c["is_old0"], c["new_key"],
                                c["new_value"], c["fnc0"], c["fnc1"]) for c in base]
    assert sum(1 for w in want_base if w[1] == 0) > 150 and len({w[1] for w in want_base}) >= 4
    order = [rng.randrange(len(base)) for _ in range(1500)]
    cases = [base[i] for i in order]
    n = len(cases)
    sib = elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32)
    out, st = engine.smt_process(elems(c["old_root"] for c in cases), sib, elems(c["old_key"] for c in cases),
                                 elems(c["old_value"] for c in cases), np.array([c["is_old0"] for c in cases], np.uint8),
                                 elems(c["new_key"] for c in cases), elems(c["new_value"] for c in cases),
                                 np.array([c["fnc0"] for c in cases], np.uint8), np.array([c["fnc1"] for c in cases], np.uint8))
    got = list(zip(ints(out), [int(x) for x in st]))
    assert got == [want_base[i] for i in order]
    # the same batch in gnark-crypto's Montgomery memory
    import gnark_crypto_primitives_b200 as g
    rm = (1 << 256) % R
    mont = lambda vals: elems((int(v) * rm) % R if int(v) < R else int(v) for v in vals)
    out_m, st_m = engine.smt_process(mont(c["old_root"] for c in cases), mont(s for c in cases for s in c["siblings"]).reshape(n, n_levels, 32),
                                     mont(c["old_key"] for c in cases), mont(c["old_value"] for c in cases),
                                     np.array([c["is_old0"] for c in cases], np.uint8), mont(c["new_key"] for c in cases),
                                     mont(c["new_value"] for c in cases), np.array([c["fnc0"] for c in cases], np.uint8),
                                     np.array([c["fnc1"] for c in cases], np.uint8), fmt=g.FMT_MONTGOMERY)
    ok = [i for i in range(n) if got[i][1] == 0]
    assert [int(st_m[i]) for i in ok] == [0] * len(ok)
    assert [ints(out_m[i:i + 1])[0] for i in ok] == [(got[i][0] * rm) % R for i in ok]
