"""GPU parity: SMT verifier through the C ABI vs the literal state-machine oracle (bit-exact flags/status/roots)."""
import random

import numpy as np
import pytest

from oracle import smt as osmt
from oracle.field import R
from tests.util import census_proof, dense_proof, elems, ints

pytestmark = pytest.mark.gpu


def run_general(engine, cases, n_levels):
    """cases: dicts with enabled, root, siblings, old_key, old_value, is_old0, key, value, fnc."""
    n = len(cases)
    sib = elems([s for c in cases for s in c["siblings"]]).reshape(n, n_levels, 32)
    flags, status, roots = engine.smt_verify(
        elems(c["root"] for c in cases), sib, elems(c["key"] for c in cases), elems(c["value"] for c in cases),
        old_keys=elems(c["old_key"] for c in cases), old_values=elems(c["old_value"] for c in cases),
        is_old0=np.array([c["is_old0"] for c in cases], dtype=np.uint8),
        fnc=np.array([c["fnc"] for c in cases], dtype=np.uint8),
        enabled=np.array([c["enabled"] for c in cases], dtype=np.uint8), want_roots=True)
    want = [osmt.verifier(c["enabled"], c["root"], c["siblings"], c["old_key"], c["old_value"], c["is_old0"], c["key"],
                          c["value"], c["fnc"]) for c in cases]
    return flags, status, ints(roots), want


def test_real_tree_inclusion_and_exclusion(engine):
    """Fixture shape of tree/test/verifier_bn254_test.go:36-68: 64 levels, 10 leaves, 8-byte keys."""
    rng = random.Random(64)
    n_levels = 64
    tree = osmt.Tree(n_levels)
    keys = [rng.getrandbits(64) for _ in range(10)]
    for i, k in enumerate(keys):
        tree.add(k, 10 if i == 0 else rng.getrandbits(64))
    root = tree.root()
    cases = []
    for k in keys:
        p = tree.gen_proof(k)
        assert p["exists"]
        cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=k, old_value=p["old_value"], is_old0=0,
                          key=k, value=p["old_value"], fnc=0))
    absent = [rng.getrandbits(64) for _ in range(40)]
    seen_empty = seen_neighbour = False
    for k in absent:
        p = tree.gen_proof(k)
        assert not p["exists"]
        seen_empty |= p["is_old0"] == 1
        seen_neighbour |= p["is_old0"] == 0
        cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=p["old_key"], old_value=p["old_value"],
                          is_old0=p["is_old0"], key=k, value=0, fnc=1))
        # claiming inclusion of an absent key must give flag 0
        cases.append(dict(enabled=1, root=root, siblings=p["siblings"], old_key=k, old_value=5, is_old0=0, key=k,
                          value=5, fnc=0))
    assert seen_neighbour
    flags, status, roots, want = run_general(engine, cases, n_levels)
    for i, w in enumerate(want):
        assert (int(flags[i]), int(status[i]), roots[i]) == w, i
    assert all(int(f) == 1 for f in flags[:10])


def test_inclusion_form_matches(engine):
    rng = random.Random(7)
    n_levels = 64
    tree = osmt.Tree(n_levels)
    keys = [rng.getrandbits(64) for _ in range(33)]
    vals = [rng.randrange(R) for _ in keys]
    for k, v in zip(keys, vals):
        tree.add(k, v)
    root = tree.root()
    proofs = [tree.gen_proof(k) for k in keys]
    sib = elems([s for p in proofs for s in p["siblings"]]).reshape(len(keys), n_levels, 32)
    # shared root form
    flags, status, roots = engine.smt_verify_inclusion(elems([root]), sib, elems(keys), elems(vals), want_roots=True)
    assert flags.all() and not status.any() and set(ints(roots)) == {root}
    # a wrong value / wrong root / wrong sibling flips the flag, not the status
    bad_vals = list(vals)
    bad_vals[3] = (bad_vals[3] + 1) % R
    flags, status = engine.smt_verify_inclusion(elems([root]), sib, elems(keys), elems(bad_vals))
    assert list(flags) == [0 if i == 3 else 1 for i in range(len(keys))] and not status.any()


def test_state_machine_all_selector_combinations(engine):
    """Every boolean (enabled, fnc, isOld0) combination, with old==new keys and not, valid and corrupted."""
    rng = random.Random(11)
    n_levels = 20
    cases = []
    for enabled in (0, 1):
        for fnc in (0, 1):
            for is0 in (0, 1):
                for same_key in (0, 1):
                    for corrupt in (0, 1, 2):
                        root, sib, key, value = census_proof(rng, n_levels, 3, 12)
                        old_key = key if same_key else rng.getrandbits(n_levels)
                        old_value = value if same_key else rng.randrange(R)
                        if corrupt == 1:
                            root = (root + 1) % R
                        if corrupt == 2:
                            sib = list(sib)
                            sib[-1] = 5
                        cases.append(dict(enabled=enabled, root=root, siblings=sib, old_key=old_key,
                                          old_value=old_value, is_old0=is0, key=key, value=value, fnc=fnc))
    flags, status, roots, want = run_general(engine, cases, n_levels)
    for i, w in enumerate(want):
        assert (int(flags[i]), int(status[i]), roots[i]) == w, (i, cases[i])


def test_assertion_failures_become_status(engine):
    rng = random.Random(3)
    n_levels = 16
    root, sib, key, value = dense_proof(rng, n_levels)
    base = dict(enabled=1, root=root, siblings=sib, old_key=key, old_value=value, is_old0=0, key=key, value=value, fnc=0)
    cases = [dict(base),
             dict(base, key=key | (1 << n_levels)),       # lowBits assertion, tree/smt/utils.go:11-13
             dict(base, is_old0=2),                         # cf. tree/smt/processor_test.go:60-61
             dict(base, fnc=3), dict(base, enabled=2),
             dict(base, value=R),                           # non-canonical
             dict(base, siblings=[R + 1] + list(sib[1:])),
             dict(base, root=2**256 - 1)]
    flags, status, roots, want = run_general(engine, cases, n_levels)
    assert [int(s) for s in status] == [0, 2, 3, 3, 3, 1, 1, 1]
    assert [w[1] for w in want] == [0, 2, 3, 3, 3, 1, 1, 1]
    assert int(flags[0]) == 1 and not flags[1:].any()


def test_lowbits_key7_vs_key5(engine):
    """tree/smt/utils_test.go:27-39 restated: with 3 levels the path bits of key 7 are 1,1,1; key 5 differs."""
    n_levels = 3
    sib = [11, 22, 0]
    root7 = osmt.fold_inclusion(sib, 7, 9)
    flags, status = engine.smt_verify_inclusion(elems([root7, root7]), elems(sib + sib).reshape(2, 3, 32),
                                                elems([7, 5]), elems([9, 9]))
    assert list(flags) == [1, 0] and not status.any()
    assert osmt.inclusion_verifier(root7, sib, 5, 9)[0] == 0


@pytest.mark.parametrize("n_levels", [2, 3, 31, 32, 33, 64, 160, 253])
def test_dense_and_corrupted(engine, n_levels):
    """Primary synthetic distribution of SURVEY.md 8d at small N; every 4th proof corrupted."""
    rng = random.Random(n_levels)
    n = 24 if n_levels >= 160 else 48
    items = []
    for i in range(n):
        root, sib, key, value = dense_proof(rng, n_levels, key_bits=min(n_levels, 253))
        if i % 4 == 1:
            kind = rng.randrange(4)
            sib = list(sib)
            if kind == 0:
                root = (root + 1) % R
            elif kind == 1:
                sib[rng.randrange(n_levels - 1)] = rng.randrange(1, R)
            elif kind == 2:
                value = (value + 1) % R
            else:
                sib[-1] = rng.randrange(1, R)
        items.append((root, sib, key, value))
    sib = elems([s for it in items for s in it[1]]).reshape(n, n_levels, 32)
    flags, status, roots = engine.smt_verify_inclusion(elems(it[0] for it in items), sib, elems(it[2] for it in items),
                                                       elems(it[3] for it in items), want_roots=True)
    want = [osmt.inclusion_verifier(it[0], it[1], it[2], it[3]) for it in items]
    got_roots = ints(roots)
    for i, w in enumerate(want):
        assert (int(flags[i]), int(status[i]), got_roots[i]) == w, i
    assert [int(f) for f in flags] == [0 if i % 4 == 1 else 1 for i in range(n)]


def test_census_like(engine):
    rng = random.Random(99)
    n_levels, n = 160, 64
    items = [census_proof(rng, n_levels) for _ in range(n)]
    sib = elems([s for it in items for s in it[1]]).reshape(n, n_levels, 32)
    flags, status, roots = engine.smt_verify_inclusion(elems(it[0] for it in items), sib, elems(it[2] for it in items),
                                                       elems(it[3] for it in items), want_roots=True)
    assert flags.all() and not status.any()
    assert ints(roots) == [it[0] for it in items]


def test_all_zero_siblings_single_leaf_tree(engine):
    """arbo single-leaf tree: root = leaf hash, no siblings (lidx = 0)."""
    n_levels = 64
    key, value = 12345, 67890
    root = osmt.hash1(key, value)
    sib = elems([0] * n_levels).reshape(1, n_levels, 32)
    flags, status, roots = engine.smt_verify_inclusion(elems([root]), sib, elems([key]), elems([value]), want_roots=True)
    assert int(flags[0]) == 1 and int(status[0]) == 0 and ints(roots)[0] == root
    assert osmt.inclusion_verifier(root, [0] * n_levels, key, value) == (1, 0, root)


def test_chunked_host_path(engine, monkeypatch):
    """Force several chunks through the double-buffered host pipeline."""
    monkeypatch.setenv("GCP_B200_SMT_CHUNK", "7")
    rng = random.Random(1234)
    n_levels, n = 24, 50
    items = [dense_proof(rng, n_levels) for _ in range(n)]
    sib = elems([s for it in items for s in it[1]]).reshape(n, n_levels, 32)
    flags, status, roots = engine.smt_verify_inclusion(elems(it[0] for it in items), sib, elems(it[2] for it in items),
                                                       elems(it[3] for it in items), want_roots=True)
    assert flags.all() and not status.any() and ints(roots) == [it[0] for it in items]


def test_bad_arguments(engine):
    import gnark_crypto_primitives_b200 as g

    with pytest.raises(g.EngineError):
        engine.smt_verify_inclusion(np.zeros((1, 32), np.uint8), np.zeros((1, 1, 32), np.uint8),
                                    np.zeros((1, 32), np.uint8), np.zeros((1, 32), np.uint8))
    with pytest.raises(g.EngineError):
        engine.smt_verify_inclusion(np.zeros((1, 32), np.uint8), np.zeros((1, 254, 32), np.uint8),
                                    np.zeros((1, 32), np.uint8), np.zeros((1, 32), np.uint8))


def test_scan_kernel_reports_path_length_and_flags(engine):
    """gcp_smt_scan_dev: lidx / siblings[n-1]==0 / canonical per proof, for ragged level counts."""
    torch = pytest.importorskip("torch")
    rng = random.Random(2024)
    for n_levels in (2, 5, 16, 17, 33, 160, 253):
        n = 67
        rows, want_lidx, want_info = [], [], []
        for i in range(n):
            L = rng.randint(0, n_levels - 1)
            sib = [0 if rng.random() < 0.2 else rng.randrange(1, R) for _ in range(L)]
            if L:
                sib[L - 1] = rng.randrange(1, R)
            sib += [0] * (n_levels - L)
            info = 3
            if i % 5 == 1:
                sib[n_levels - 1] = 9                 # siblings[n-1] != 0
                info &= ~1
            if i % 7 == 2:
                sib[rng.randrange(n_levels)] = R + rng.randrange(3)   # not canonical (r, r+1, r+2)
                info &= ~2
            if sib[n_levels - 1] != 0:
                info &= ~1
            last = max([j for j in range(n_levels - 1) if sib[j] != 0], default=-1)
            rows.append(sib)
            want_lidx.append(last + 1)
            want_info.append(info)
        d = torch.from_numpy(elems([x for r_ in rows for x in r_]).reshape(n, n_levels, 32)).cuda()
        lidx = torch.empty(n, dtype=torch.int16, device="cuda")
        info = torch.empty(n, dtype=torch.uint8, device="cuda")
        engine.smt_scan_dev(n_levels, n, d, lidx, info, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert lidx.cpu().tolist() == want_lidx, n_levels
        assert info.cpu().tolist() == want_info, n_levels


def test_pinned_host_buffers_give_the_same_results(engine):
    """gcp_host_alloc: the caller's arrays in page-locked memory (what bench.py's e2e uses through torch)."""
    import gnark_crypto_primitives_b200 as g

    rng = random.Random(404)
    n_levels, n = 160, 96
    items = [census_proof(rng, n_levels) for _ in range(n)]
    items[7] = (items[7][0] ^ 1,) + items[7][1:]
    sib = elems([s for it in items for s in it[1]]).reshape(n, n_levels, 32)
    roots, keys, vals = elems(it[0] for it in items), elems(it[2] for it in items), elems(it[3] for it in items)
    want = engine.smt_verify_inclusion(roots, sib, keys, vals)
    buf = g.PinnedBuffer(sib.nbytes)
    try:
        pinned = buf.array.reshape(sib.shape)
        pinned[...] = sib
        got = engine.smt_verify_inclusion(roots, pinned, keys, vals)
    finally:
        buf.close()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert int(want[0][7]) == 0 and int(want[0].sum()) == n - 1
    empty = g.PinnedBuffer(0)
    empty.close()


def test_large_pageable_input_takes_the_staged_copy_path(engine, monkeypatch):
    """Host buffers of 64 MB and more in pageable memory are copied through the library's page-locked staging ring by
    several host threads (capi.cu: h2d_copy).  Same results as the driver-staged path, bit for bit."""
    rng = np.random.default_rng(64)
    n_levels, n = 160, 16384                                  # 84 MB of siblings in one chunk
    sib = rng.integers(0, 256, size=(n, n_levels, 32), dtype=np.uint8)
    sib[:, :, 31] &= 0x0F
    sib[:, n_levels - 1, :] = 0
    keys = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keys[:, 20:] = 0                                          # < 2^160
    vals = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    vals[:, 31] &= 0x0F
    zero_roots = np.zeros((n, 32), np.uint8)
    _, st, roots = engine.smt_verify_inclusion(zero_roots, sib, keys, vals, want_roots=True)
    assert not st.any()
    roots[::5, 0] ^= 1                                        # every fifth proof gets a wrong root
    staged = engine.smt_verify_inclusion(roots, sib, keys, vals, want_roots=True)
    monkeypatch.setenv("GCP_B200_NO_STAGING", "1")
    plain = engine.smt_verify_inclusion(roots, sib, keys, vals, want_roots=True)
    for a, b in zip(staged, plain):
        assert np.array_equal(a, b)
    assert [int(f) for f in staged[0][:10]] == [0, 1, 1, 1, 1, 0, 1, 1, 1, 1] and not staged[1].any()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_differential_over_levels_paths_and_selectors(engine, seed):
    """Seeded differential run against the literal gadget oracle: random n_levels in [2, 253] (not multiples of the path
    kernel's 4-level staging chunk), every path length from 0 to n-1 mixed in ONE batch (the length sort and the warp
    grouping see ragged neighbours), interior zero siblings, all selector values incl. non-boolean ones, keys at and past
    2^n, elements at and past r, and correct as well as corrupted roots."""
    rng = random.Random(1000 + seed)
    n_levels = rng.choice([2, 5, 7, 9, 13, 17, 29, 41, 67, 101, 163, 253][seed - 1::3])
    cases = []
    for i in range(140):
        L = rng.choice([0, 1, n_levels - 1, rng.randrange(n_levels), min(n_levels - 1, rng.randrange(8))])
        sib = [0 if rng.random() < 0.15 else rng.randrange(1, R) for _ in range(L)] + [0] * (n_levels - L)
        if L:
            sib[L - 1] = rng.randrange(1, R)
        key = rng.getrandbits(n_levels)
        value = rng.randrange(R)
        fnc, is0, enabled = rng.choice([0, 0, 1]), rng.choice([0, 0, 1]), rng.choice([1, 1, 1, 0])
        same = rng.random() < 0.3
        old_key = key if same else rng.getrandbits(n_levels)
        old_value = value if same else rng.randrange(R)
        # the root the state machine accepts for these selectors: fold from the leaf it injects
        if fnc == 0:
            root = osmt.fold_inclusion(sib, key, value)
        elif is0:
            root = osmt.verifier(1, 0, sib, old_key, old_value, 1, key, value, 1)[2]
        else:
            root = osmt.verifier(1, 0, sib, old_key, old_value, 0, key, value, 1)[2]
        c = dict(enabled=enabled, root=root, siblings=sib, old_key=old_key, old_value=old_value, is_old0=is0, key=key,
                 value=value, fnc=fnc)
        mut = rng.randrange(12)
        if mut == 0:
            c["root"] = (root + 1) % R
        elif mut == 1:
            c["siblings"] = sib[:-1] + [rng.randrange(1, R)]            # siblings[n-1] != 0
        elif mut == 2:
            c["key"] = key | (1 << n_levels)                             # lowBits assertion
        elif mut == 3:
            c[rng.choice(["is_old0", "fnc", "enabled"])] = rng.choice([2, 3, 255])
        elif mut == 4:
            c[rng.choice(["root", "value", "old_value", "old_key"])] = rng.choice([R, R + 1, 2**256 - 1])
        elif mut == 5 and n_levels > 2:
            j = rng.randrange(n_levels - 1)
            c["siblings"] = sib[:j] + [rng.choice([R, 2**256 - 1])] + sib[j + 1:]
        elif mut == 6 and L:
            j = rng.randrange(L)
            c["siblings"] = sib[:j] + [(sib[j] + 1) % R] + sib[j + 1:]
        cases.append(c)
    flags, status, roots, want = run_general(engine, cases, n_levels)
    for i, w in enumerate(want):
        assert (int(flags[i]), int(status[i]), roots[i]) == w, (n_levels, i, cases[i])
    assert len({w[:2] for w in want}) >= 4        # the batch really mixes flag/status outcomes
