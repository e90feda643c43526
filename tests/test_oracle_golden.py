"""CPU: pin the Python oracle to every golden vector / known-answer test the reference holds for the hot path
(SURVEY.md 8c).  Citations are file:line under the reference tree."""
import hashlib
from pathlib import Path

import pytest

from oracle import edwards as ed
from oracle import elgamal as eg
from oracle import keccak
from oracle import poseidon as pos
from oracle import smt
from oracle.field import R, poseidon_tables

BLOB = Path(__file__).resolve().parent.parent / "gnark_crypto_primitives_b200" / "data" / "poseidon_bn254.bin"


def test_constant_blob_is_the_committed_one():
    # oracle/gen_constants.py derives it from hash/native/bn254/poseidon/constants.go (sha256 c7f3fe34...3120)
    assert hashlib.sha256(BLOB.read_bytes()).hexdigest() == \
        "9e5f79761800da6ad9b6c45be037889b6be9ac6e9e92666a193c7416221305d0"
    tabs = poseidon_tables()
    rp = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]  # poseidon.go:119
    for t in range(2, 18):
        assert tabs[t]["RP"] == rp[t - 2]
        assert len(tabs[t]["C"]) == 8 * t + rp[t - 2] and len(tabs[t]["S"]) == (2 * t - 1) * rp[t - 2]
        assert len(tabs[t]["M"]) == t * t == len(tabs[t]["P"])


def test_poseidon_public_circomlib_vectors():
    assert pos.hash([1]) == 18586133768512220936620570745912940619677854269274689475585506675881198879027
    assert pos.hash([1, 2]) == 7853200120776062878684798364095072458815029376092732009249414926327459813530
    assert pos.hash([1, 2, 3]) == 6542985608222806190361240322586112750744169038454362455181422643027100751666
    assert pos.hash([1, 2, 3, 4]) == 18821383157269793795438455681495246036402687001665670618754263018637548127333
    assert pos.hash(list(range(1, 17))) == \
        9989051620750914585850546081941653841776809718687451684622678807385399211877


def test_poseidon_reference_fixed_inputs_regression():
    # hash/native/bn254/poseidon/poseidon_test.go:39 ; hash/emulated/bn254/poseidon/poseidon_test.go:85-88
    assert pos.hash([297262668938251460872476410954775437897592223497]) == \
        21099541378821686330832093407308585959971016892597585818017774528142419287929
    assert pos.multihash(list(range(1, 61))) == \
        10383247944466245790564312669548703973436539043614368627989218605970689057797


def test_poseidon_arity_errors_and_mul_counts():
    with pytest.raises(pos.PoseidonError):
        pos.hash([])                       # poseidon.go:41-43
    with pytest.raises(pos.PoseidonError):
        pos.hash([1] * 17)                 # Write drops > 16 inputs, poseidon.go:103-108
    with pytest.raises(pos.PoseidonError):
        pos.multihash([1] * 4097)          # poseidon.go:57-59
    assert [pos.field_mul_count(t) for t in (2, 3, 4, 13, 17)] == [414, 594, 772, 3328, 4896]


def test_multihash_structure():
    xs = list(range(100, 132))  # 32 inputs = 2 x Hash16 + Hash2 (poseidon_test.go:60-91 shape)
    assert pos.multihash(xs) == pos.hash([pos.hash(xs[:16]), pos.hash(xs[16:])])
    xs = list(range(1, 300))    # 299 inputs -> 19 chunk hashes -> recursive
    chunks = [pos.hash(xs[i:i + 16]) for i in range(0, 299, 16)]
    assert pos.multihash(xs) == pos.hash([pos.hash(chunks[:16]), pos.hash(chunks[16:])])


# ---- curve / ElGamal ---------------------------------------------------------------------------------
A1 = (9394823613809705110116613460910105025054013892432913335394773002247992354854,
      11024289076895660735250094443495165598068433425499992095815117261086957091439)
A2 = (19797710400961090194828422488006966273839297906754012108828771044254185248577,
      14922306070502274021207471871631487833716178512064982802994428541540403297523)
Z = 1742022034800951303918649192268907782873437905421353131642789173698540722240
PK = (11914791603502957547081391328506057813324763482068493183947042790384502567641,
      14401335135320235427678361547570520415347209769899386704796044467443275407252)
C1 = (3200797265076621797396943577308832679391396371860226890120121432230653785233,
      5210110328792812562066091196399294499414608384227631465547758111507815530790)
C2 = (14353965765711180631440746432124851641123026187756655584132953629432908500962,
      18899802722931794583798498860596714297548149427767678529077963923612627261516)


def test_decryption_proof_static_kat():
    """elgamal/ciphertext_test.go:286-345: valid assignment solves, A1.Y = 0 does not."""
    for p in (A1, A2, PK, C1, C2):
        assert ed.is_on_curve(p)
    assert eg.verify_decryption_proof(PK, (C1, C2), 50, A1, A2, Z)
    assert not eg.verify_decryption_proof(PK, (C1, C2), 50, (A1[0], 0), A2, Z)
    assert not eg.verify_decryption_proof(PK, (C1, C2), 51, A1, A2, Z)


def test_curve_parameters():
    assert ed.is_on_curve(ed.G) and ed.is_on_curve(ed.IDENTITY)
    assert ed.scalar_mul(ed.G, ed.ORDER) == ed.IDENTITY
    assert ed.scalar_mul(ed.G, ed.ORDER + 5) == ed.scalar_mul(ed.G, 5)
    # ecc/format/twistededwards.go:17 maps iden3 B8 to gnark's base point
    b8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
          16950150798460657717958625567821834550301663161624707787222815936182638968203)
    assert ed.te_to_rte(*b8) == ed.G and ed.rte_to_te(*ed.G) == b8
    # ecc/format/twistededwards_test.go:71,73 round trip
    pt = (20284931487578954787250358776722960153090567235942462656834196519767860852891,
          21185575020764391300398134415668786804224896114060668011215204645513129497221)
    assert ed.rte_to_te(*ed.te_to_rte(*pt)) == pt


def test_scalar_mul_fast_path_equals_affine_definition():
    for s in (0, 1, 2, 12345, ed.ORDER - 1, R - 1):
        assert ed.scalar_mul(ed.G, s) == ed.scalar_mul_affine(ed.G, s)
        assert eg.fixed_base_scalar_mul(s) == ed.scalar_mul_affine(ed.G, s)


def test_encrypt_compatibility():
    """elgamal/encrypt_test.go:144-159: PubKey=G, k=12345, m=67890; fixed-base == generic; EncryptedZero == Encrypt(0)."""
    c1, c2 = eg.encrypt(ed.G, 12345, 67890)
    assert c1 == ed.scalar_mul(ed.G, 12345)
    assert c2 == ed.add(ed.scalar_mul(ed.G, 67890), ed.scalar_mul(ed.G, 12345))
    assert c1 == (19918712023960437102123786886411468478902094248734671506464346041569881874462,
                  13276557205153692030187527501273228448057533426731746626187331221465573305487)
    assert c2 == (839909438842816078619007291947839299631027001100725795038371896036563634986,
                  2683297034354865619026157779551553408927035959921603566752318144387126273226)
    assert eg.encrypted_zero(ed.G, 12345) == eg.encrypt(ed.G, 12345, 0)


def test_encrypt_with_specific_data():
    """elgamal/encrypt_test.go:180-217: fixed pubkey, three k (two above the subgroup order), m = 0 must solve."""
    pk = (18604149248430057540085528196797394191454458259161233471314599389622530831795,
          1988784568828097512630242539176296837964596457792502130892628909648459248949)
    assert ed.is_on_curve(pk)
    ks = [855131146298194990003384743709896434741839908245,
          5883442530210657871581412827617735506655215369087356134218551734599178232070,
          3979028711588105728532079493967382119023185938755564152610807942458151212832]
    assert sum(k > ed.ORDER for k in ks) == 2
    for k in ks:
        c1, c2 = eg.encrypt(pk, k, 0)
        assert ed.is_on_curve(c1) and ed.is_on_curve(c2)
        assert (c1, c2) == eg.encrypted_zero(pk, k)
        assert c1 == ed.scalar_mul(ed.G, k % ed.ORDER)


def test_homomorphic_tally_closed_form():
    d = 0xB200
    pk = ed.scalar_mul(ed.G, d)
    ks, ms = [R - 3, 17, 2 ** 200 + 5], [3, 5, 65535]
    cts = [eg.encrypt(pk, k, m) for k, m in zip(ks, ms)]
    total = eg.tally(cts)
    assert total == eg.encrypt(pk, sum(ks) % ed.ORDER, sum(ms) % ed.ORDER)
    assert eg.assert_decrypt(total, d, sum(ms))
    assert eg.ct_neg(eg.ct_neg(cts[0])) == cts[0]
    assert eg.ct_add(cts[0], eg.ct_neg(cts[0])) == eg.new_ciphertext()


def test_off_curve_public_key_is_an_assertion():
    with pytest.raises(ed.CurveError):
        eg.encrypt((1, 2), 5, 7)


# ---- SMT ------------------------------------------------------------------------------------------------
def test_lowbits_binds_to_key():
    """tree/smt/utils_test.go:27-39: key 7 has bits 1,1,1 over 3 levels; key 5 does not; key 8 fails the range assertion."""
    sib = [11, 22, 0]
    root = smt.fold_inclusion(sib, 7, 9)
    assert smt.inclusion_verifier(root, sib, 7, 9) == (1, 0, root)
    assert smt.inclusion_verifier(root, sib, 5, 9)[:2] == (0, 0)
    assert smt.inclusion_verifier(root, sib, 8, 9)[:2] == (0, smt.STATUS_KEY_RANGE)


def test_non_boolean_selector_is_rejected():
    """tree/smt/processor_test.go:60-61 (IsOld0 = 2 must be rejected) restated for the verifier's selectors."""
    assert smt.verifier(1, 1, [5, 6, 0], 1, 2, 2, 7, 4, 0)[1] == smt.STATUS_NOT_BOOLEAN
    assert smt.verifier(2, 1, [5, 6, 0], 1, 2, 0, 7, 4, 0)[1] == smt.STATUS_NOT_BOOLEAN


def test_verifier_on_oracle_built_tree():
    import random

    rng = random.Random(20)
    tree = smt.Tree(64)
    keys = [rng.getrandbits(64) for _ in range(10)]       # tree/test/verifier_bls12377_test.go:23-27 shape
    for k in keys:
        tree.add(k, rng.getrandbits(64))
    root = tree.root()
    for k in keys:
        p = tree.gen_proof(k)
        assert smt.inclusion_verifier(root, p["siblings"], k, p["old_value"]) == (1, 0, root)
        assert smt.fold_inclusion(p["siblings"], k, p["old_value"]) == root
        assert smt.inclusion_verifier(root, p["siblings"], k, p["old_value"] ^ 1)[0] == 0
    kinds = set()
    for _ in range(60):
        k = rng.getrandbits(64)
        p = tree.gen_proof(k)
        kinds.add(p["is_old0"])
        assert smt.exclusion_verifier(root, p["siblings"], p["old_key"], p["old_value"], p["is_old0"], k)[:2] == (1, 0)
        # key-reuse guard: pretending the neighbour has the same key must fail (verifier.go:223-230)
        if not p["is_old0"]:
            assert smt.exclusion_verifier(root, p["siblings"], k, p["old_value"], 0, k)[0] == 0
    assert 0 in kinds


def test_gadget_hashes_every_level():
    cnt = [0]
    sib = [3, 0, 4, 0, 0, 0]
    h = smt.hash1(9, 8)
    smt.verifier_with_leaf_hash_flag(1, 1, sib, 9, h, 0, 9, h, 0, count_hashes=cnt)
    assert cnt[0] == len(sib)            # verifier.go:211-220 runs for all n levels
    assert smt.verifier(0, 123, [5, 6, 7], 1, 2, 0, 3, 4, 0) == (1, 0, 0)   # enabled = 0 bypasses everything


# ---- Keccak / address --------------------------------------------------------------------------------------
def test_keccak_public_vectors():
    assert keccak.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak.keccak256(b"hello").hex() == "1c8aff950685c2ed4bc3174f3472287b56d9517b9c948127319a09a7a36deac8"
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    gy = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
    assert keccak.derive_address(gx.to_bytes(32, "big") + gy.to_bytes(32, "big")).hex() == \
        "7e5f4552091a69125d5dfcb7b8c2659029395bdf"


# ---- SMT processor / EdDSA ---------------------------------------------------------------------
def test_processor_against_tree_transitions():
    import random

    rng = random.Random(5)
    n = 32
    tree = smt.Tree(n)
    keys = []
    for _ in range(25):
        k, v = rng.getrandbits(n), rng.randrange(R)
        old_root, p = tree.root(), tree.gen_proof(k)
        tree.add(k, v)
        keys.append(k)
        new_root = tree.root()
        args = (p["siblings"], p["old_key"], p["old_value"], p["is_old0"], k, v)
        assert smt.processor(old_root, *args, 1, 0) == (new_root, 0)          # insert
        assert smt.processor(new_root, *args, 1, 1) == (old_root, 0)          # delete mirrors insert
        assert smt.processor(old_root, *args, 0, 0) == (old_root, 0)          # nop
        assert smt.processor((old_root + 1) % R, *args, 1, 0) == (0, smt.STATUS_ASSERTION)
    for k in keys[:8]:
        old_root, p = tree.root(), tree.gen_proof(k)
        v2 = rng.randrange(R)
        tree.add(k, v2)
        assert smt.processor(old_root, p["siblings"], k, p["old_value"], 0, k, v2, 0, 1) == (tree.root(), 0)   # update
    # processor_test.go:46-71: all-zero nop is valid, IsOld0 = 2 is rejected
    assert smt.processor(0, [0, 0, 0, 0], 0, 0, 0, 0, 0, 0, 0) == (0, 0)
    assert smt.processor(0, [0, 0, 0, 0], 0, 0, 2, 0, 0, 0, 0) == (0, smt.STATUS_NOT_BOOLEAN)


def test_eddsa_oracle_self_consistency():
    from oracle import eddsa

    a, r, s = eddsa.sign(123456789, 987654321, 42)
    assert eddsa.is_valid(a, r, s, 42) == (1, True)
    assert eddsa.is_valid(a, r, s, 43) == (0, True)
    assert eddsa.is_valid(a, r, s + 1, 42) == (0, True)
    assert eddsa.is_valid((1, 2), r, s, 42) == (0, False)
    # rteB8 (ecc/bn254/eddsa/constants.go:11-18) is gnark's base point
    b8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
          16950150798460657717958625567821834550301663161624707787222815936182638968203)
    assert ed.te_to_rte(*b8) == ed.G


def test_mimc7_public_iden3_vectors():
    from oracle import mimc7

    # go-iden3-crypto mimc7 test vectors; the first input is the one hash/native/bn254/mimc7/mimc_test.go:37 uses
    assert mimc7.hash([12]) == 16051049095595290701999129793867590386356047218708919933694064829788708231421
    assert mimc7.hash([12, 45, 78, 41]) == \
        18226366069841799622585958305961373004333097209608110160936134895615261821931
    assert mimc7.hash([1] * 63) == 0          # mimc.go:33-38: more than 62 inputs are dropped
    assert len(mimc7.constants()) == 91 and mimc7.constants()[0] == 0


def test_poseidon2_wrapper_semantics_and_key_blob():
    """hash/native/bn254/poseidon2: the permutation is un-vendored (PARITY UNPINNED, oracle/poseidon2.py header); what the
    reference's own lines fix is the wrapper - arity, mod-r reduction, min/max order, the chaining rule - and that is
    checked here, together with the committed key blob being the oracle's derivation."""
    import struct
    from pathlib import Path

    from oracle import poseidon2 as p2
    from oracle.field import R

    with pytest.raises(ValueError):
        p2.hash([1])                                       # native.go:31-33
    with pytest.raises(ValueError):
        p2.hash([1, 2, 3, 4])
    assert p2.hash([1, 2]) == p2.hash([2, 1])              # native.go:42-44, hints.go:10-19
    assert p2.hash([1, 2, 1]) != p2.hash([2, 1, 1])        # leaves keep their order
    assert p2.hash([R + 1, 2]) == p2.hash([1, 2])          # native.go:37-39 (SafeBigInt)
    # chaining rule native.go:47-61 written out for a leaf
    cv = 0
    for m in (11, 22, 1):
        cv = (p2.permutation([cv, m])[1] + m) % R
    assert p2.hash([11, 22, 1]) == cv
    # the permutation is a bijection built from invertible layers: distinct inputs, distinct outputs; 0 is not fixed
    outs = {tuple(p2.permutation([a, b])) for a in range(4) for b in range(4)}
    assert len(outs) == 16 and (0, 0) not in outs
    keys = p2.round_keys()
    assert [len(r) for r in keys] == [2] * 3 + [1] * 50 + [2] * 3 and len(p2.flat_round_keys()) == 62
    assert p2.unflatten(p2.flat_round_keys()) == keys
    assert p2.seed_string() == "Poseidon2-BN254[t=2,rF=6,rP=50,d=5]"
    blob = (Path(__file__).resolve().parent.parent / "gnark_crypto_primitives_b200" / "data" /
            "poseidon2_bn254_t2.bin").read_bytes()
    assert struct.unpack_from("<4I", blob) == (0x32534F50, 1, 62, 0)
    assert [int.from_bytes(blob[16 + 32 * i:48 + 32 * i], "little") for i in range(62)] == p2.flat_round_keys()
