"""CPU: tests/golden/vectors.json is what tests/golden/generate.py produces from the pinned oracle (no drift), and its
"reference" entries are the values that are literal in the reference tree or public."""
import importlib.util
import json
import random
from pathlib import Path

GOLDEN = Path(__file__).resolve().parent / "golden"


def load():
    return json.loads((GOLDEN / "vectors.json").read_text())


def test_fixture_file_is_reproducible_from_the_oracle():
    spec = importlib.util.spec_from_file_location("golden_generate", GOLDEN / "generate.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    rng = random.Random(0xB200)
    doc = load()
    assert gen.poseidon_section(rng) == doc["poseidon"]
    assert gen.smt_section(rng) == doc["smt"]
    assert gen.elgamal_section(rng) == doc["elgamal"]
    assert gen.eddsa_section(rng) == doc["eddsa"]
    assert gen.keccak_section(rng) == doc["keccak_address"]
    assert gen.mimc7_section(rng) == doc["mimc7"]
    assert gen.poseidon2_section(rng) == doc["poseidon2"]


def test_reference_entries_carry_the_published_values():
    doc = load()
    kat = {tuple(int(x) for x in e["in"]): int(e["out"]) for e in doc["poseidon"]["kat"]}
    # circomlib test-suite values
    assert kat[(1,)] == 18586133768512220936620570745912940619677854269274689475585506675881198879027
    assert kat[(1, 2)] == 7853200120776062878684798364095072458815029376092732009249414926327459813530
    assert kat[(1, 2, 3)] == 6542985608222806190361240322586112750744169038454362455181422643027100751666
    assert kat[(1, 2, 3, 4)] == 18821383157269793795438455681495246036402687001665670618754263018637548127333
    assert kat[tuple(range(1, 17))] == 9989051620750914585850546081941653841776809718687451684622678807385399211877
    # elgamal/ciphertext_test.go:286-345: the valid assignment verifies, A1.Y = 0 does not
    proofs = doc["elgamal"]["decryption_proof"]
    assert [p["valid"] for p in proofs if p["source"] == "reference"] == [1, 0]
    # ecc/format/twistededwards.go:17: iden3 B8 maps to gnark's base point
    assert doc["elgamal"]["te_to_rte"][0]["rte"] == [
        "9671717474070082183213120605117400219616337014328744928644933853176787189663",
        "16950150798460657717958625567821834550301663161624707787222815936182638968203"]
    # public Ethereum vector: address of the secp256k1 generator (private key 1)
    assert doc["keccak_address"][0]["address"] == "7e5f4552091a69125d5dfcb7b8c2659029395bdf"
    # tree/smt/utils_test.go:27-39: key 7 verifies, key 5 does not, key 8 fails the range assertion
    tiny = [c for c in doc["smt"]["verifier"] if c.get("n_levels") == 3]
    assert [(c["flag"], c["status"]) for c in tiny] == [(1, 0), (0, 0), (0, 2)]
    # go-iden3-crypto mimc7 test vectors (public): Hash([12]) and Hash([12, 45, 78, 41])
    m = {tuple(int(x) for x in e["in"]): int(e["out"]) for e in doc["mimc7"][:3]}
    assert m[(12,)] == 16051049095595290701999129793867590386356047218708919933694064829788708231421
    assert m[(12, 45, 78, 41)] == 18226366069841799622585958305961373004333097209608110160936134895615261821931


def test_c_port_reproduces_the_fixtures():
    """oracle/c (the CPU baseline of bench.py) against the stored vectors: Poseidon, the verifier, Encrypt, tally."""
    import numpy as np

    from oracle import cport
    from tests.util import elems, ints

    doc = load()
    by_arity = {}
    for e in doc["poseidon"]["kat"] + doc["poseidon"]["hash"]:
        by_arity.setdefault(len(e["in"]), []).append(e)
    for arity, es in by_arity.items():
        out, st = cport.poseidon_hash(elems([int(x) for e in es for x in e["in"]]).reshape(len(es), arity, 32))
        assert not st.any() and ints(out) == [int(e["out"]) for e in es]
    for e in doc["poseidon"]["multihash"]:
        out, st = cport.poseidon_multihash(elems([int(x) for x in e["in"]]).reshape(1, len(e["in"]), 32))
        assert ints(out)[0] == int(e["out"])
    smt = doc["smt"]
    cases = [c for c in smt["verifier"] if "n_levels" not in c]
    n = len(cases)
    flags, status, _ = cport.smt_verify(
        elems(int(c["root"]) for c in cases), elems([int(x) for c in cases for x in c["siblings"]]).reshape(n, smt["n_levels"], 32),
        elems(int(c["key"]) for c in cases), elems(int(c["value"]) for c in cases),
        old_keys=elems(int(c["old_key"]) for c in cases), old_values=elems(int(c["old_value"]) for c in cases),
        is_old0=np.array([c["is_old0"] for c in cases], np.uint8), fnc=np.array([c["fnc"] for c in cases], np.uint8),
        enabled=np.array([c["enabled"] for c in cases], np.uint8), literal=True)
    assert [int(f) for f in flags] == [c["flag"] for c in cases] and [int(s) for s in status] == [c["status"] for c in cases]
    enc = [e for e in doc["elgamal"]["encrypt"]]
    for e in enc:
        ct, st = cport.elgamal_encrypt(elems([int(x) for x in e["pk"]]).reshape(2, 32), elems([int(e["k"])]), elems([int(e["m"])]))
        assert int(st[0]) == 0 and ints(ct[0]) == [int(x) for x in e["ct"]]
    t = doc["elgamal"]["tally"]
    nb, nf = len(t["ballots"]), len(t["ballots"][0])
    tal, st = cport.elgamal_tally(elems([int(x) for row in t["ballots"] for c in row for x in c]).reshape(nb, nf, 4, 32))
    assert [ints(tal[f]) for f in range(nf)] == [[int(x) for x in c] for c in t["tally"]]
