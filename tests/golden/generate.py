"""Writes tests/golden/vectors.json: input/output vectors of every entry point of the path.

Outputs come from oracle/*.py, the CPU restatement that tests/test_oracle_golden.py pins to the reference's own
vectors (public circomlib Poseidon vectors, the static decryption-proof KAT elgamal/ciphertext_test.go:286-345,
encrypt_test.go:144-217, ecc/format/twistededwards.go:17, tree/smt/utils_test.go:27-39).  The reference itself is Go
and cannot run here (no toolchain), so it cannot generate vectors; entries marked "source": "reference" carry values
that are literal in the reference tree or public (circomlib / Ethereum), the others are oracle outputs on seeded
inputs and serve as regression pins for the oracle and as GPU parity vectors that need no oracle at test time.

    python tests/golden/generate.py        # rewrites vectors.json (deterministic)
"""
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import eddsa as oeddsa  # noqa: E402
from oracle import edwards as ed  # noqa: E402
from oracle import elgamal as eg  # noqa: E402
from oracle import keccak  # noqa: E402
from oracle import mimc7  # noqa: E402
from oracle import poseidon2  # noqa: E402
from oracle import poseidon as pos  # noqa: E402
from oracle import smt  # noqa: E402
from oracle import smt as osmt  # noqa: E402
from oracle.field import R  # noqa: E402

S = str  # big integers travel as decimal strings


def poseidon_section(rng):
    out = {"kat": [], "hash": [], "multihash": []}
    for inputs, src in (([1], "reference"), ([1, 2], "reference"), ([1, 2, 3], "reference"), ([1, 2, 3, 4], "reference"),
                        (list(range(1, 17)), "reference"),
                        ([297262668938251460872476410954775437897592223497], "oracle")):  # poseidon_test.go:39 input
        out["kat"].append({"in": [S(x) for x in inputs], "out": S(pos.hash(inputs)), "source": src})
    for arity in range(1, 17):
        for _ in range(2):
            xs = [rng.randrange(R) for _ in range(arity)]
            out["hash"].append({"in": [S(x) for x in xs], "out": S(pos.hash(xs))})
    edge = [0, 1, R - 1, R - 2, (1 << 253) - 1, 1 << 128]
    for a in edge:
        for b in (0, R - 1, 12345):
            out["hash"].append({"in": [S(a), S(b)], "out": S(pos.hash([a, b]))})
    for ln in (17, 32, 60, 256, 257):
        xs = list(range(1, ln + 1)) if ln == 60 else [rng.randrange(R) for _ in range(ln)]  # emulated test: 1..60
        out["multihash"].append({"in": [S(x) for x in xs], "out": S(pos.multihash(xs))})
    return out


def smt_section(rng):
    n_levels = 64                                       # tree/test/verifier_bls12377_test.go:23-27 shape
    tree = smt.Tree(n_levels)
    keys = [rng.getrandbits(64) for _ in range(10)]
    vals = [10] + [rng.getrandbits(64) for _ in range(9)]
    for k, v in zip(keys, vals):
        tree.add(k, v)
    root = tree.root()
    cases = []

    def add_case(enabled, rt, sib, ok, ov, i0, k, v, fnc, packed):
        f, st, r = smt.verifier(enabled, rt, sib, ok, ov, i0, k, v, fnc)
        cases.append({"enabled": enabled, "root": S(rt), "siblings": [S(x) for x in sib], "packed": packed.hex(),
                      "old_key": S(ok), "old_value": S(ov), "is_old0": i0, "key": S(k), "value": S(v), "fnc": fnc,
                      "flag": f, "status": st, "level0": S(r)})

    for k, v in zip(keys, vals):
        p = tree.gen_proof(k)
        pk = tree.last_packed
        add_case(1, root, p["siblings"], k, v, 0, k, v, 0, pk)                    # inclusion
        add_case(1, root, p["siblings"], k, v, 0, k, v ^ 1, 0, pk)                # wrong value
        add_case(1, (root + 1) % R, p["siblings"], k, v, 0, k, v, 0, pk)          # wrong root
    for _ in range(12):
        k = rng.getrandbits(64)
        p = tree.gen_proof(k)
        pk = tree.last_packed
        add_case(1, root, p["siblings"], p["old_key"], p["old_value"], p["is_old0"], k, 0, 1, pk)     # exclusion
        add_case(1, root, p["siblings"], k, 5, 0, k, 5, 0, pk)                                        # false inclusion
        add_case(0, root, p["siblings"], k, 5, 0, k, 5, 0, pk)                                        # disabled
    sib3 = [11, 22, 0]                                   # tree/smt/utils_test.go:27-39
    r3 = smt.fold_inclusion(sib3, 7, 9)
    p3 = smt.pack_siblings([11, 22])
    for key in (7, 5, 8):
        f, st, r = smt.inclusion_verifier(r3, sib3, key, 9)
        cases.append({"enabled": 1, "root": S(r3), "siblings": [S(x) for x in sib3], "packed": p3.hex(), "old_key": S(key),
                      "old_value": "9", "is_old0": 0, "key": S(key), "value": "9", "fnc": 0, "flag": f, "status": st,
                      "level0": S(r), "n_levels": 3})
    # processor: insert / update / delete / nop on a growing tree
    t2 = smt.Tree(n_levels)
    proc = []
    pkeys = []
    for _ in range(8):
        k, v = rng.getrandbits(64), rng.randrange(R)
        old_root = t2.root()
        p = t2.gen_proof(k)
        packed = t2.last_packed
        t2.add(k, v)
        pkeys.append(k)
        for fnc0, fnc1, oroot in ((1, 0, old_root), (1, 1, t2.root()), (0, 0, old_root)):
            nr, st = smt.processor(oroot, p["siblings"], p["old_key"], p["old_value"], p["is_old0"], k, v, fnc0, fnc1)
            proc.append({"old_root": S(oroot), "siblings": [S(x) for x in p["siblings"]], "packed": packed.hex(),
                         "old_key": S(p["old_key"]), "old_value": S(p["old_value"]), "is_old0": p["is_old0"],
                         "new_key": S(k), "new_value": S(v), "fnc0": fnc0, "fnc1": fnc1, "new_root": S(nr), "status": st})
    for k in pkeys[:3]:
        old_root = t2.root()
        p = t2.gen_proof(k)
        packed = t2.last_packed
        v2 = rng.randrange(R)
        nr, st = smt.processor(old_root, p["siblings"], k, p["old_value"], 0, k, v2, 0, 1)
        t2.add(k, v2)
        assert nr == t2.root() and st == 0
        proc.append({"old_root": S(old_root), "siblings": [S(x) for x in p["siblings"]], "packed": packed.hex(),
                     "old_key": S(k), "old_value": S(p["old_value"]), "is_old0": 0, "new_key": S(k), "new_value": S(v2),
                     "fnc0": 0, "fnc1": 1, "new_root": S(nr), "status": st})
    return {"n_levels": n_levels, "verifier": cases, "processor": proc}


def pt(p):
    return [S(p[0]), S(p[1])]


def smt_leaf_hash_section():
    """Leaf-hash forms (verifier.go:129-183, processor.go:16, hash.go:10-19) on a tree with three values per leaf, and
    the reference's post-insert processor flow (wrapper_arbo.go:119-185)."""
    rng = random.Random(0xB202)
    n_levels, n_values = 24, 3
    tree = osmt.Tree(n_levels)
    leaves = {}
    while len(leaves) < 9:
        leaves[rng.getrandbits(n_levels)] = tuple(rng.randrange(R) for _ in range(n_values))
    for k, v in leaves.items():
        tree.add(k, v)
    root = tree.root()
    hash1 = [{"key": S(k), "values": [S(x) for x in v], "hash": S(osmt.hash1(k, *v))} for k, v in leaves.items()]
    hash1.append({"key": S(5), "values": [], "hash": S(osmt.hash1(5))})
    cases = []
    for k, v in list(leaves.items())[:5]:
        p = tree.gen_proof(k)
        for vals in (v, v[:2] + ((v[2] + 1) % R,)):
            h = osmt.hash1(k, *vals)
            f, st, lv = osmt.verifier_with_leaf_hash_flag(1, root, p["siblings"], k, h, 0, k, h, 0)
            cases.append({"root": S(root), "siblings": [S(x) for x in p["siblings"]], "old_key": S(k), "hash1_old": S(h),
                          "is_old0": 0, "key": S(k), "hash1_new": S(h), "fnc": 0, "flag": f, "status": st, "level0": S(lv)})
    for _ in range(6):
        k = rng.getrandbits(n_levels)
        if k in leaves:
            continue
        p = tree.gen_proof(k)
        ho = osmt.hash1(p["old_key"], *p["old_value"]) if p["is_old0"] == 0 else 0
        hn = osmt.hash1(k, 0, 0, 0)
        f, st, lv = osmt.verifier_with_leaf_hash_flag(1, root, p["siblings"], p["old_key"], ho, p["is_old0"], k, hn, 1)
        cases.append({"root": S(root), "siblings": [S(x) for x in p["siblings"]], "old_key": S(p["old_key"]), "hash1_old": S(ho),
                      "is_old0": p["is_old0"], "key": S(k), "hash1_new": S(hn), "fnc": 1, "flag": f, "status": st, "level0": S(lv)})
    t2 = osmt.Tree(n_levels)
    arbo = []
    ks = [rng.getrandbits(n_levels) for _ in range(10)]
    ks += [ks[2], ks[7]]
    for k in ks:
        a = osmt.arbo_add_or_update(t2, k, rng.randrange(R))
        arbo.append({"old_root": S(a["old_root"]), "new_root": S(a["new_root"]), "old_key": S(a["old_key"]),
                     "old_value": S(a["old_value"]), "is_old0": a["is_old0"], "new_key": S(a["new_key"]),
                     "new_value": S(a["new_value"]), "fnc0": a["fnc0"], "fnc1": a["fnc1"], "packed": a["packed"].hex(),
                     "siblings": [S(x) for x in a["siblings"]]})
    return {"n_levels": n_levels, "hash1": hash1, "verifier_with_leaf_hash": cases, "processor_arbo": arbo}


def elgamal_section(rng):
    out = {}
    # elgamal/ciphertext_test.go:289-303 (literal in the reference)
    a1 = (9394823613809705110116613460910105025054013892432913335394773002247992354854,
          11024289076895660735250094443495165598068433425499992095815117261086957091439)
    a2 = (19797710400961090194828422488006966273839297906754012108828771044254185248577,
          14922306070502274021207471871631487833716178512064982802994428541540403297523)
    z = 1742022034800951303918649192268907782873437905421353131642789173698540722240
    pk = (11914791603502957547081391328506057813324763482068493183947042790384502567641,
          14401335135320235427678361547570520415347209769899386704796044467443275407252)
    c1 = (3200797265076621797396943577308832679391396371860226890120121432230653785233,
          5210110328792812562066091196399294499414608384227631465547758111507815530790)
    c2 = (14353965765711180631440746432124851641123026187756655584132953629432908500962,
          18899802722931794583798498860596714297548149427767678529077963923612627261516)
    proofs = []
    for msg, A1, zz, src in ((50, a1, z, "reference"), (50, (a1[0], 0), z, "reference"), (51, a1, z, "oracle"),
                             (50, a1, (z + 1) % ed.ORDER, "oracle")):
        ok = ed.is_on_curve(A1) and eg.verify_decryption_proof(pk, (c1, c2), msg, A1, a2, zz)
        proofs.append({"pk": pt(pk), "c1": pt(c1), "c2": pt(c2), "msg": S(msg), "a1": pt(A1), "a2": pt(a2), "z": S(zz),
                       "valid": int(bool(ok)), "on_curve": int(ed.is_on_curve(A1)), "source": src})
    out["decryption_proof"] = proofs
    enc = []
    fixed_pk = (18604149248430057540085528196797394191454458259161233471314599389622530831795,
                1988784568828097512630242539176296837964596457792502130892628909648459248949)   # encrypt_test.go:188
    items = [(ed.G, 12345, 67890)]                                                                  # encrypt_test.go:152-155
    items += [(fixed_pk, k, 0) for k in (855131146298194990003384743709896434741839908245,
                                         5883442530210657871581412827617735506655215369087356134218551734599178232070,
                                         3979028711588105728532079493967382119023185938755564152610807942458151212832)]
    d = 0xB200
    pk2 = ed.scalar_mul(ed.G, d)
    items += [(pk2, rng.randrange(R), rng.randrange(1 << 16)) for _ in range(12)]
    items += [(pk2, 0, 0), (pk2, R - 1, R - 1), (pk2, ed.ORDER, 1)]
    for P, k, m in items:
        ct = eg.encrypt(P, k, m)
        enc.append({"pk": pt(P), "k": S(k), "m": S(m), "ct": [S(x) for x in eg.serialize(ct)]})
    out["encrypt"] = enc
    out["fixed_base"] = [{"s": S(s), "p": pt(eg.fixed_base_scalar_mul(s))} for s in
                         (0, 1, 2, 15, 16, 12345, ed.ORDER - 1, ed.ORDER, R - 1, rng.randrange(R))]
    cts = [eg.encrypt(pk2, rng.randrange(R), rng.randrange(1 << 16)) for _ in range(6)]
    out["add"] = [{"a": [S(x) for x in eg.serialize(cts[i])], "b": [S(x) for x in eg.serialize(cts[i + 1])],
                   "sum": [S(x) for x in eg.serialize(eg.ct_add(cts[i], cts[i + 1]))],
                   "neg_a": [S(x) for x in eg.serialize(eg.ct_neg(cts[i]))]} for i in range(5)]
    nb, nf = 5, 3
    ks = [[rng.randrange(R) for _ in range(nf)] for _ in range(nb)]
    ms = [[rng.randrange(1 << 16) for _ in range(nf)] for _ in range(nb)]
    ballots = [[eg.encrypt(pk2, ks[b][f], ms[b][f]) for f in range(nf)] for b in range(nb)]
    out["tally"] = {"pk": pt(pk2), "k": [[S(x) for x in r] for r in ks], "m": [[S(x) for x in r] for r in ms],
                    "ballots": [[[S(x) for x in eg.serialize(c)] for c in row] for row in ballots],
                    "tally": [[S(x) for x in eg.serialize(eg.tally([ballots[b][f] for b in range(nb)]))] for f in range(nf)]}
    dec = []
    for _ in range(4):
        dd, msg = rng.randrange(1, ed.ORDER), rng.randrange(1000)
        ct = eg.encrypt(ed.scalar_mul(ed.G, dd), rng.randrange(ed.ORDER), msg)
        dec.append({"ct": [S(x) for x in eg.serialize(ct)], "priv": S(dd), "msg": S(msg), "ok": 1})
        dec.append({"ct": [S(x) for x in eg.serialize(ct)], "priv": S(dd), "msg": S(msg + 1), "ok": 0})
    out["assert_decrypt"] = dec
    # ecc/format/twistededwards.go:17 and twistededwards_test.go:71,73
    b8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
          16950150798460657717958625567821834550301663161624707787222815936182638968203)
    tp = (20284931487578954787250358776722960153090567235942462656834196519767860852891,
          21185575020764391300398134415668786804224896114060668011215204645513129497221)
    out["te_to_rte"] = [{"te": pt(p), "rte": pt(ed.te_to_rte(*p))} for p in (b8, tp)]
    # GCP_COORDS_TE: the same gadgets with every point in iden3 coordinates at the boundary (own generator: the sections
    # after this one keep their vectors)
    r2 = random.Random(0xB201)
    te = lambda p: ed.rte_to_te(*p)
    te_ct = lambda c: [S(x) for q in c for x in te(q)]
    pk3 = ed.scalar_mul(ed.G, r2.randrange(1, ed.ORDER))
    nb2, nf2 = 4, 2
    ks2 = [[r2.randrange(R) for _ in range(nf2)] for _ in range(nb2)]
    ms2 = [[r2.randrange(1 << 16) for _ in range(nf2)] for _ in range(nb2)]
    b2 = [[eg.encrypt(pk3, ks2[b][f], ms2[b][f]) for f in range(nf2)] for b in range(nb2)]
    out["te_coords"] = {"pk_te": pt(te(pk3)), "k": [[S(x) for x in r] for r in ks2], "m": [[S(x) for x in r] for r in ms2],
                        "ballots_te": [[te_ct(c) for c in row] for row in b2],
                        "tally_te": [te_ct(eg.tally([b2[b][f] for b in range(nb2)])) for f in range(nf2)]}
    return out


def eddsa_section(rng):
    out = []
    for i in range(6):
        msg = rng.getrandbits(248)
        a, r, s = oeddsa.sign(rng.randrange(1, ed.ORDER), rng.randrange(1, ed.ORDER), msg)
        if i % 3 == 1:
            msg ^= 1
        if i % 3 == 2:
            s = (s + 1) % ed.ORDER
        flag, ok = oeddsa.is_valid(a, r, s, msg)
        out.append({"a": pt(a), "r": pt(r), "s": S(s), "msg": S(msg), "flag": int(flag), "assertions_hold": int(bool(ok))})
    return out


def keccak_section(rng):
    # secp256k1 generator (private key 1): the public Ethereum vector 0x7e5f4552091a69125d5dfcb7b8c2659029395bdf
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    gy = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
    rows = [gx.to_bytes(32, "big") + gy.to_bytes(32, "big")] + [bytes(rng.getrandbits(8) for _ in range(64)) for _ in range(7)]
    return [{"pub_xy_be": r.hex(), "address": keccak.derive_address(r).hex()} for r in rows]


def mimc7_section(rng):
    out = []
    for xs in ([12], [12, 45], [12, 45, 78, 41], [rng.randrange(R) for _ in range(62)]):   # mimc_test.go:34-51 input 12
        out.append({"in": [S(x) for x in xs], "out": S(mimc7.hash(xs))})
    return out


def poseidon2_section(rng):
    """Width-2 Poseidon2 hasher (hash/native/bn254/poseidon2): outputs of oracle/poseidon2.py - PARITY UNPINNED (the round
    keys are restated from gnark-crypto's published derivation, no vector exists in the reference).  The keys are stored
    so that a Go-side check can compare them with poseidon2.NewParameters(2, 6, 50).RoundKeys directly."""
    rows = [[0, 0], [1, 2], [2, 1], [R - 1, 0], [1, 2, 1], [0, 0, 0]]
    rows += [[rng.randrange(R) for _ in range(2 + i % 2)] for i in range(6)]
    states = [[0, 0], [0, 1], [1, 2], [rng.randrange(R), rng.randrange(R)]]
    return {
        "source": "oracle (unpinned)",
        "seed": poseidon2.seed_string(),
        "round_keys": [S(k) for k in poseidon2.flat_round_keys()],
        "permutation": [{"in": [S(x) for x in st], "out": [S(x) for x in poseidon2.permutation(st)]} for st in states],
        "hash": [{"in": [S(x) for x in r], "out": S(poseidon2.hash(r))} for r in rows],
    }


def main():
    rng = random.Random(0xB200)
    doc = {
        "_about": "generated by tests/golden/generate.py from oracle/*.py (seed 0xB200); integers are decimal strings",
        "poseidon": poseidon_section(rng),
        "smt": smt_section(rng),
        "smt_leaf_hash": smt_leaf_hash_section(),
        "elgamal": elgamal_section(rng),
        "eddsa": eddsa_section(rng),
        "keccak_address": keccak_section(rng),
        "mimc7": mimc7_section(rng),
        "poseidon2": poseidon2_section(rng),
    }
    path = Path(__file__).resolve().parent / "vectors.json"
    path.write_text(json.dumps(doc, indent=0, sort_keys=True) + "\n")
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
