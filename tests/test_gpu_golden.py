"""GPU parity against the COMMITTED golden fixtures (tests/golden/vectors.json) through the C ABI: no oracle code
runs here, the expected outputs are the stored ones."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.util import elems, ints

pytestmark = pytest.mark.gpu
DOC = json.loads((Path(__file__).resolve().parent / "golden" / "vectors.json").read_text())
RMONT = 1 << 256
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def I(xs):
    return [int(x) for x in xs]


def test_poseidon_vectors(engine):
    by_arity = {}
    for e in DOC["poseidon"]["kat"] + DOC["poseidon"]["hash"]:
        by_arity.setdefault(len(e["in"]), []).append(e)
    assert sorted(by_arity) == list(range(1, 17))
    for arity, es in by_arity.items():
        a = elems([x for e in es for x in I(e["in"])]).reshape(len(es), arity, 32)
        out, st = engine.poseidon_hash(a)
        assert not st.any() and ints(out) == [int(e["out"]) for e in es], arity
        # gnark-crypto Montgomery memory in and out
        am = elems([x * RMONT % R for e in es for x in I(e["in"])]).reshape(len(es), arity, 32)
        outm, st = engine.poseidon_hash(am, fmt=1)
        assert not st.any() and ints(outm) == [int(e["out"]) * RMONT % R for e in es], arity
    for e in DOC["poseidon"]["multihash"]:
        out, st = engine.poseidon_multihash(elems(I(e["in"])).reshape(1, len(e["in"]), 32))
        assert int(st[0]) == 0 and ints(out)[0] == int(e["out"]), len(e["in"])


def test_smt_verifier_vectors_dense_and_packed(engine):
    smt = DOC["smt"]
    for n_levels in (smt["n_levels"], 3):
        cases = [c for c in smt["verifier"] if c.get("n_levels", smt["n_levels"]) == n_levels]
        n = len(cases)
        args = dict(old_keys=elems(int(c["old_key"]) for c in cases), old_values=elems(int(c["old_value"]) for c in cases),
                    is_old0=np.array([c["is_old0"] for c in cases], np.uint8),
                    fnc=np.array([c["fnc"] for c in cases], np.uint8),
                    enabled=np.array([c["enabled"] for c in cases], np.uint8), want_roots=True)
        roots, keys, vals = (elems(int(c[k]) for c in cases) for k in ("root", "key", "value"))
        sib = elems([int(x) for c in cases for x in c["siblings"]]).reshape(n, n_levels, 32)
        dense = engine.smt_verify(roots, sib, keys, vals, **args)
        packed = engine.smt_verify_packed(roots, [bytes.fromhex(c["packed"]) for c in cases], n_levels, keys, vals, **args)
        for flags, status, level0 in (dense, packed):
            assert [int(f) for f in flags] == [c["flag"] for c in cases]
            assert [int(s) for s in status] == [c["status"] for c in cases]
            got = ints(level0)
            for i, c in enumerate(cases):
                if c["status"] == 0 and c["enabled"] == 1:
                    assert got[i] == int(c["level0"]), i
    assert sum(c["flag"] for c in smt["verifier"]) > 20


def test_smt_processor_vectors(engine):
    smt = DOC["smt"]
    cases, n_levels = smt["processor"], smt["n_levels"]
    n = len(cases)
    common = (elems(int(c["old_key"]) for c in cases), elems(int(c["old_value"]) for c in cases),
              np.array([c["is_old0"] for c in cases], np.uint8), elems(int(c["new_key"]) for c in cases),
              elems(int(c["new_value"]) for c in cases), np.array([c["fnc0"] for c in cases], np.uint8),
              np.array([c["fnc1"] for c in cases], np.uint8))
    roots = elems(int(c["old_root"]) for c in cases)
    sib = elems([int(x) for c in cases for x in c["siblings"]]).reshape(n, n_levels, 32)
    for out, st in (engine.smt_process(roots, sib, *common),
                    engine.smt_process_packed(roots, [bytes.fromhex(c["packed"]) for c in cases], n_levels, *common)):
        assert [int(s) for s in st] == [c["status"] for c in cases]
        assert ints(out) == [int(c["new_root"]) for c in cases]


def test_elgamal_vectors(engine):
    eg = DOC["elgamal"]
    enc = eg["encrypt"]
    n = len(enc)
    ct, st = engine.elgamal_encrypt(elems([int(x) for e in enc for x in e["pk"]]).reshape(n, 2, 32),
                                    elems(int(e["k"]) for e in enc), elems(int(e["m"]) for e in enc))
    assert not st.any()
    for i, e in enumerate(enc):
        assert ints(ct[i]) == I(e["ct"]), i
    shared = [e for e in enc if e["pk"] == enc[-1]["pk"]]
    ct, st = engine.elgamal_encrypt(elems(I(shared[0]["pk"])).reshape(2, 32), elems(int(e["k"]) for e in shared),
                                    elems(int(e["m"]) for e in shared))
    assert not st.any() and all(ints(ct[i]) == I(e["ct"]) for i, e in enumerate(shared))
    fb = eg["fixed_base"]
    pts, st = engine.elgamal_fixed_base_mul(elems(int(e["s"]) for e in fb))
    assert not st.any() and all(ints(pts[i]) == I(e["p"]) for i, e in enumerate(fb))
    add = eg["add"]
    a = elems([int(x) for e in add for x in e["a"]]).reshape(len(add), 4, 32)
    b = elems([int(x) for e in add for x in e["b"]]).reshape(len(add), 4, 32)
    s, st = engine.elgamal_add(a, b)
    assert not st.any() and all(ints(s[i]) == I(e["sum"]) for i, e in enumerate(add))
    ng, st = engine.elgamal_neg(a)
    assert not st.any() and all(ints(ng[i]) == I(e["neg_a"]) for i, e in enumerate(add))
    t = eg["tally"]
    nb, nf = len(t["ballots"]), len(t["ballots"][0])
    ballots = elems([int(x) for row in t["ballots"] for c in row for x in c]).reshape(nb, nf, 4, 32)
    tal, st = engine.elgamal_tally(ballots)
    assert not st.any() and [ints(tal[f]) for f in range(nf)] == [I(c) for c in t["tally"]]
    k = elems([int(x) for row in t["k"] for x in row]).reshape(nb, nf, 32)
    m = elems([int(x) for row in t["m"] for x in row]).reshape(nb, nf, 32)
    tal2, st = engine.elgamal_encrypt_tally(elems(I(t["pk"])).reshape(2, 32), k, m)
    assert not st.any() and np.array_equal(tal, tal2)
    dp = eg["decryption_proof"]
    flags, st = engine.elgamal_verify_decryption_proof(
        elems([int(x) for e in dp for x in e["pk"]]), elems([int(x) for e in dp for x in e["c1"] + e["c2"]]).reshape(len(dp), 4, 32),
        elems(int(e["msg"]) for e in dp), elems([int(x) for e in dp for x in e["a1"]]),
        elems([int(x) for e in dp for x in e["a2"]]), elems(int(e["z"]) for e in dp))
    assert [int(f) for f in flags] == [e["valid"] for e in dp]
    assert [int(s) == 0 for s in st] == [bool(e["on_curve"]) for e in dp]
    ad = eg["assert_decrypt"]
    flags, st = engine.elgamal_assert_decrypt(elems([int(x) for e in ad for x in e["ct"]]).reshape(len(ad), 4, 32),
                                              elems(int(e["priv"]) for e in ad), elems(int(e["msg"]) for e in ad))
    assert not st.any() and [int(f) for f in flags] == [e["ok"] for e in ad]
    tr = eg["te_to_rte"]
    rte, st = engine.te_to_rte(elems([int(x) for e in tr for x in e["te"]]).reshape(len(tr), 2, 32))
    assert not st.any() and all(ints(rte[i]) == I(e["rte"]) for i, e in enumerate(tr))
    back, st = engine.rte_to_te(rte)
    assert all(ints(back[i]) == I(e["te"]) for i, e in enumerate(tr))
    # GCP_COORDS_TE: iden3 coordinates at the boundary
    import gnark_crypto_primitives_b200 as g
    tc = eg["te_coords"]
    nb, nf = len(tc["ballots_te"]), len(tc["ballots_te"][0])
    k = elems([int(x) for row in tc["k"] for x in row])
    m = elems([int(x) for row in tc["m"] for x in row])
    pk_te = elems(I(tc["pk_te"])).reshape(2, 32)
    ct, st = engine.elgamal_encrypt(pk_te, k, m, fmt=g.COORDS_TE)
    want = [I(c) for row in tc["ballots_te"] for c in row]
    assert not st.any() and [ints(c) for c in ct] == want
    tal, st = engine.elgamal_tally(ct.reshape(nb, nf, 4, 32), fmt=g.COORDS_TE)
    assert not st.any() and [ints(t) for t in tal] == [I(c) for c in tc["tally_te"]]
    tal2, st = engine.elgamal_encrypt_tally(pk_te, k.reshape(nb, nf, 32), m.reshape(nb, nf, 32), fmt=g.COORDS_TE)
    assert not st.any() and np.array_equal(tal, tal2)


def test_smt_leaf_hash_and_arbo_vectors(engine):
    doc = DOC["smt_leaf_hash"]
    n_levels = doc["n_levels"]
    for nv in sorted({len(e["values"]) for e in doc["hash1"]}):
        rows = [e for e in doc["hash1"] if len(e["values"]) == nv]
        vals = elems([int(x) for e in rows for x in e["values"]]).reshape(len(rows), nv, 32) if nv else \
            np.zeros((len(rows), 0, 32), np.uint8)
        out, st = engine.smt_leaf_hash(elems(int(e["key"]) for e in rows), vals)
        assert not st.any() and ints(out) == [int(e["hash"]) for e in rows]
    vc = doc["verifier_with_leaf_hash"]
    n = len(vc)
    flags, status, roots = engine.smt_verify_with_leaf_hash(
        elems(int(c["root"]) for c in vc), elems([int(x) for c in vc for x in c["siblings"]]).reshape(n, n_levels, 32),
        elems(int(c["key"]) for c in vc), elems(int(c["hash1_new"]) for c in vc),
        old_keys=elems(int(c["old_key"]) for c in vc), hash1_old=elems(int(c["hash1_old"]) for c in vc),
        is_old0=np.array([c["is_old0"] for c in vc], np.uint8), fnc=np.array([c["fnc"] for c in vc], np.uint8), want_roots=True)
    assert [int(f) for f in flags] == [c["flag"] for c in vc] and [int(x) for x in status] == [c["status"] for c in vc]
    assert ints(roots) == [int(c["level0"]) for c in vc]
    pa = doc["processor_arbo"]
    u8 = lambda name: np.array([c[name] for c in pa], np.uint8)
    out, st = engine.smt_process_arbo(elems(int(c["old_root"]) for c in pa), [bytes.fromhex(c["packed"]) for c in pa], n_levels,
                                      elems(int(c["old_key"]) for c in pa), elems(int(c["old_value"]) for c in pa), u8("is_old0"),
                                      elems(int(c["new_key"]) for c in pa), elems(int(c["new_value"]) for c in pa), u8("fnc0"),
                                      u8("fnc1"))
    assert not st.any() and ints(out) == [int(c["new_root"]) for c in pa]


def test_eddsa_keccak_mimc7_vectors(engine):
    ed = DOC["eddsa"]
    flags, st = engine.eddsa_verify(elems([int(x) for e in ed for x in e["a"]]), elems([int(x) for e in ed for x in e["r"]]),
                                    elems(int(e["s"]) for e in ed), elems(int(e["msg"]) for e in ed))
    assert [int(f) for f in flags] == [e["flag"] for e in ed]
    assert [int(s) == 0 for s in st] == [bool(e["assertions_hold"]) for e in ed]
    kk = DOC["keccak_address"]
    addr = engine.keccak_address(np.frombuffer(b"".join(bytes.fromhex(e["pub_xy_be"]) for e in kk), np.uint8).reshape(-1, 64))
    assert [bytes(a).hex() for a in addr] == [e["address"] for e in kk]
    for e in DOC["mimc7"]:
        out, st = engine.mimc7_hash(elems(I(e["in"])).reshape(1, len(e["in"]), 32))
        assert int(st[0]) == 0 and ints(out)[0] == int(e["out"])


def test_poseidon2_vectors(engine):
    p2 = DOC["poseidon2"]
    for e in p2["hash"]:
        out, st = engine.poseidon2_hash(elems(I(e["in"])).reshape(1, len(e["in"]), 32))
        assert int(st[0]) == 0 and ints(out)[0] == int(e["out"])
    states = elems([int(x) for e in p2["permutation"] for x in e["in"]]).reshape(-1, 2, 32)
    out, st = engine.poseidon2_permutation(states)
    assert not st.any() and ints(out) == [int(x) for e in p2["permutation"] for x in e["out"]]
