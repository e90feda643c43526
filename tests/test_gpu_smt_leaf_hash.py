"""GPU parity: the leaf-hash forms of the SMT gadgets (VerifierWithLeafHash[Flag] verifier.go:129-183,
ProcessorWithLeafHash processor.go:16-72) and the variadic Hash1 (hash.go:10-19), against the literal oracle, on
trees whose leaves carry one, two and three values."""
import random

import numpy as np
import pytest

import gnark_crypto_primitives_b200 as g
from oracle import poseidon as opos
from oracle import smt as osmt
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu

R_MONT = (1 << 256) % R


def to_mont(values):
    return elems((int(v) * R_MONT) % R for v in values)


@pytest.mark.parametrize("n_values", [0, 1, 2, 3, 7, 14])
def test_hash1_is_variadic(engine, n_values):
    """Hash1(key, values..., 1): t = n_values + 3, up to the hasher's 16 inputs."""
    rng = random.Random(1000 + n_values)
    n = 37
    keys = [rng.randrange(R) for _ in range(n)]
    vals = [[rng.randrange(R) for _ in range(n_values)] for _ in range(n)]
    keys[0], keys[1] = 0, R - 1
    if n_values:
        vals[1] = [R - 1] * n_values
    want = [opos.hash([k, *v, 1]) for k, v in zip(keys, vals)]
    v_arr = elems([x for v in vals for x in v]).reshape(n, n_values, 32) if n_values else np.zeros((n, 0, 32), np.uint8)
    out, st = engine.smt_leaf_hash(elems(keys), v_arr)
    assert not st.any() and ints(out) == want
    if n_values == 1:
        assert want == [osmt.hash1(k, v[0]) for k, v in zip(keys, vals)]
    # gnark-crypto fr.Element memory in and out
    v_m = to_mont([x for v in vals for x in v]).reshape(n, n_values, 32) if n_values else v_arr
    out_m, st_m = engine.smt_leaf_hash(to_mont(keys), v_m, fmt=g.FMT_MONTGOMERY)
    assert not st_m.any() and ints(out_m) == [(w * R_MONT) % R for w in want]


def test_hash1_arity_and_canonical_errors(engine):
    keys = elems([1, 2])
    with pytest.raises(g.EngineError) as e:
        engine.smt_leaf_hash(keys, np.zeros((2, 15, 32), np.uint8))          # key + 15 values + 1 = 17 inputs
    assert e.value.code == -1 and "bad inputs provided" in str(e.value)
    out, st = engine.smt_leaf_hash(elems([5, 6]), elems([R, 7]).reshape(2, 1, 32))
    assert [int(x) for x in st] == [1, 0] and ints(out) == [0, opos.hash([6, 7, 1])]


def multi_value_tree(rng, n_levels, n_values, n_leaves):
    tree = osmt.Tree(n_levels)
    leaves = {}
    while len(leaves) < n_leaves:
        k = rng.getrandbits(n_levels)
        leaves[k] = tuple(rng.randrange(R) for _ in range(n_values))
    for k, v in leaves.items():
        tree.add(k, v)
    return tree, leaves


@pytest.mark.parametrize("n_values", [2, 3])
def test_verifier_with_leaf_hash_on_multi_value_leaves(engine, n_values):
    """Inclusion and exclusion proofs of a tree with n_values values per leaf: leaf hashes from gcp_smt_leaf_hash, the
    path from gcp_smt_verify_with_leaf_hash; flags, status and recomputed roots against VerifierWithLeafHashFlag."""
    rng = random.Random(2000 + n_values)
    n_levels = 40
    tree, leaves = multi_value_tree(rng, n_levels, n_values, 24)
    root = tree.root()
    cases = []
    for k, v in leaves.items():
        p = tree.gen_proof(k)
        assert p["exists"] and p["old_value"] == v
        cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=k, old_vals=v, is_old0=0, key=k, vals=v,
                          fnc=0))
        wrong = tuple((x + 1) % R for x in v[:1]) + v[1:]
        cases.append(dict(cases[-1], old_vals=wrong, vals=wrong))            # one value changed: flag 0
    for _ in range(30):
        k = rng.getrandbits(n_levels)
        if k in leaves:
            continue
        p = tree.gen_proof(k)
        ov = p["old_value"] if p["is_old0"] == 0 else (0,) * n_values
        cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=p["old_key"], old_vals=ov,
                          is_old0=p["is_old0"], key=k, vals=(0,) * n_values, fnc=1))
    cases.append(dict(cases[0], enabled=0, root=5))
    n = len(cases)
    flat = lambda name: elems([x for c in cases for x in c[name]]).reshape(n, n_values, 32)
    h_old, st_o = engine.smt_leaf_hash(elems(c["old_key"] for c in cases), flat("old_vals"))
    h_new, st_n = engine.smt_leaf_hash(elems(c["key"] for c in cases), flat("vals"))
    assert not st_o.any() and not st_n.any()
    assert ints(h_new) == [osmt.hash1(c["key"], *c["vals"]) for c in cases]
    sib = elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32)
    flags, status, roots = engine.smt_verify_with_leaf_hash(
        elems(c["root"] for c in cases), sib, elems(c["key"] for c in cases), h_new,
        old_keys=elems(c["old_key"] for c in cases), hash1_old=h_old,
        is_old0=np.array([c["is_old0"] for c in cases], np.uint8), fnc=np.array([c["fnc"] for c in cases], np.uint8),
        enabled=np.array([c["enabled"] for c in cases], np.uint8), want_roots=True)
    want = [osmt.verifier_with_leaf_hash_flag(c["enabled"], c["root"], c["siblings"], c["old_key"],
                                              osmt.hash1(c["old_key"], *c["old_vals"]), c["is_old0"], c["key"],
                                              osmt.hash1(c["key"], *c["vals"]), c["fnc"]) for c in cases]
    got = list(zip([int(f) for f in flags], [int(s) for s in status], ints(roots)))
    # the engine reports level[0] = 0 for a disabled proof, as the gadget computes it
    assert got == [(w[0], w[1], w[2]) for w in want]
    assert sum(w[0] for w in want) >= len(leaves) + 10 and any(w[0] == 0 for w in want)


def test_leaf_hash_form_equals_value_form_and_checks_its_inputs(engine):
    """With Hash1(key, value) supplied by the caller the leaf-hash form must agree with smt.Verifier bit for bit, in both
    element formats; a leaf hash >= r is not a field element (status 1); the inclusion shorthand (no old leaf) works."""
    rng = random.Random(77)
    n_levels = 33
    tree = osmt.Tree(n_levels)
    kv = {rng.getrandbits(n_levels): rng.randrange(R) for _ in range(20)}
    for k, v in kv.items():
        tree.add(k, v)
    root = tree.root()
    ks = list(kv)
    sib_l = [tree.gen_proof(k)["siblings"] for k in ks]
    n = len(ks)
    sib = elems([s for row in sib_l for s in row]).reshape(n, n_levels, 32)
    vals = [kv[k] for k in ks]
    vals[3] = (vals[3] + 1) % R
    f0, s0, r0 = engine.smt_verify(elems([root]), sib, elems(ks), elems(vals), want_roots=True)
    h = [osmt.hash1(k, v) for k, v in zip(ks, vals)]
    f1, s1, r1 = engine.smt_verify_with_leaf_hash(elems([root]), sib, elems(ks), elems(h), want_roots=True)
    assert (f0 == f1).all() and (s0 == s1).all() and (r0 == r1).all() and int(f0[3]) == 0 and int(f0.sum()) == n - 1
    sib_m = to_mont([s for row in sib_l for s in row]).reshape(n, n_levels, 32)
    f2, s2, r2 = engine.smt_verify_with_leaf_hash(to_mont([root]), sib_m, to_mont(ks), to_mont(h), want_roots=True,
                                                  fmt=g.FMT_MONTGOMERY)
    assert (f2 == f0).all() and not s2.any() and ints(r2) == [(x * R_MONT) % R for x in ints(r0)]
    bad = list(h)
    bad[5] = R + 1
    f3, s3 = engine.smt_verify_with_leaf_hash(elems([root]), sib, elems(ks), elems(bad))
    assert int(s3[5]) == 1 and int(f3[5]) == 0 and [int(x) for i, x in enumerate(s3) if i != 5] == [0] * (n - 1)
    with pytest.raises(g.EngineError):
        engine.smt_verify_with_leaf_hash(elems([root]), sib, elems(ks), elems(h), old_keys=elems(ks))   # hash1_old missing


@pytest.mark.parametrize("n_values", [1, 3])
def test_processor_with_leaf_hash(engine, n_values):
    """Insert / update / delete / nop on a growing tree of multi-value leaves through gcp_smt_process_with_leaf_hash,
    against ProcessorWithLeafHash and the tree's own roots; with one value it must equal gcp_smt_process."""
    rng = random.Random(3000 + n_values)
    n_levels = 36
    tree = osmt.Tree(n_levels)
    cases = []
    present = {}
    for step in range(48):
        kind = rng.choice(["insert", "insert", "update", "nop"]) if present else "insert"
        if kind == "insert":
            k = rng.getrandbits(n_levels)
            while k in present:
                k = rng.getrandbits(n_levels)
            v = tuple(rng.randrange(R) for _ in range(n_values))
            p = tree.gen_proof(k)
            ov = p["old_value"] if p["is_old0"] == 0 else (0,) * n_values
            c = dict(old_root=tree.root(), siblings=p["siblings"], old_key=p["old_key"], old_vals=ov, is_old0=p["is_old0"],
                     new_key=k, new_vals=v, fnc0=1, fnc1=0)
            tree.add(k, v)
            present[k] = v
            c["want_root"] = tree.root()
        elif kind == "update":
            k = rng.choice(list(present))
            v = tuple(rng.randrange(R) for _ in range(n_values))
            p = tree.gen_proof(k)
            c = dict(old_root=tree.root(), siblings=p["siblings"], old_key=k, old_vals=present[k], is_old0=0, new_key=k,
                     new_vals=v, fnc0=0, fnc1=1)
            tree.add(k, v)
            present[k] = v
            c["want_root"] = tree.root()
        else:
            k = rng.choice(list(present))
            p = tree.gen_proof(k)
            c = dict(old_root=tree.root(), siblings=p["siblings"], old_key=k, old_vals=present[k], is_old0=0, new_key=k,
                     new_vals=present[k], fnc0=0, fnc1=0, want_root=tree.root())
        cases.append(c)
    # deletes: the transition of an insert read backwards (fnc = (1,1) swaps old and new, processor.go:58-61)
    for c in [c for c in cases if c["fnc0"] == 1][:6]:
        cases.append(dict(c, old_root=c["want_root"], fnc1=1, want_root=c["old_root"]))
    cases.append(dict(cases[0], old_root=(cases[0]["old_root"] + 1) % R, want_root=0))          # wrong old root: assertion
    n = len(cases)
    flat = lambda name: elems([x for c in cases for x in c[name]]).reshape(n, n_values, 32)
    h_old, _ = engine.smt_leaf_hash(elems(c["old_key"] for c in cases), flat("old_vals"))
    h_new, _ = engine.smt_leaf_hash(elems(c["new_key"] for c in cases), flat("new_vals"))
    sib = elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32)
    bits = lambda name: np.array([c[name] for c in cases], np.uint8)
    args = (elems(c["old_root"] for c in cases), sib, elems(c["old_key"] for c in cases))
    out, st = engine.smt_process_with_leaf_hash(*args, h_old, bits("is_old0"), elems(c["new_key"] for c in cases), h_new,
                                                bits("fnc0"), bits("fnc1"))
    want = [osmt.processor_with_leaf_hash(c["old_root"], c["siblings"], c["old_key"], osmt.hash1(c["old_key"], *c["old_vals"]),
                                          c["is_old0"], c["new_key"], osmt.hash1(c["new_key"], *c["new_vals"]), c["fnc0"],
                                          c["fnc1"]) for c in cases]
    assert list(zip(ints(out), [int(s) for s in st])) == want
    assert [w[0] for w in want[:-1]] == [c["want_root"] for c in cases[:-1]] and want[-1] == (0, osmt.STATUS_ASSERTION)
    if n_values == 1:
        out1, st1 = engine.smt_process(*args, flat("old_vals").reshape(n, 32), bits("is_old0"),
                                       elems(c["new_key"] for c in cases), flat("new_vals").reshape(n, 32), bits("fnc0"),
                                       bits("fnc1"))
        assert (out1 == out).all() and (st1 == st).all()


def test_device_resident_forms(engine):
    """gcp_smt_leaf_hash_dev -> gcp_smt_verify_with_leaf_hash_dev on device buffers, 2-value leaves, one stream."""
    import torch

    rng = random.Random(9)
    n_levels, n_values = 30, 2
    tree, leaves = multi_value_tree(rng, n_levels, n_values, 64)
    root = tree.root()
    ks = list(leaves)
    n = len(ks)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_keys = dev(elems(ks))
    d_vals = dev(elems([x for k in ks for x in leaves[k]]).reshape(n, n_values, 32))
    d_sib = dev(elems([s for k in ks for s in tree.gen_proof(k)["siblings"]]).reshape(n, n_levels, 32))
    d_root = dev(elems([root]))
    d_h = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_status = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    engine.smt_leaf_hash_dev(d_keys, d_vals, n_values, n, d_h, d_st, stream=st)
    engine.smt_verify_with_leaf_hash_dev(n_levels, n, d_root, True, d_sib, d_keys, d_h, d_flags, d_status, stream=st)
    torch.cuda.synchronize()
    assert bool(d_flags.all()) and not bool(d_status.any()) and not bool(d_st.any())
    assert ints(d_h.cpu().numpy()) == [osmt.hash1(k, *leaves[k]) for k in ks]
