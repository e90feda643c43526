"""GPU parity: AssertDecrypt, DecryptionProof.Verify (incl. the reference's static KAT) and TE<->RTE."""
import random

import numpy as np
import pytest

from oracle import edwards as ed
from oracle import elgamal as eg
from oracle import poseidon as opos
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu

# elgamal/ciphertext_test.go:289-303
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


def make_proof(rng, d, msg):
    """Honest Chaum-Pedersen proof for ciphertext Encrypt([d]G, k, msg) (prover side, in the oracle)."""
    pk = ed.scalar_mul(ed.G, d)
    k = rng.randrange(ed.ORDER)
    c1, c2 = eg.encrypt(pk, k, msg)
    dpt = ed.add(c2, ed.neg(eg.fixed_base_scalar_mul(msg)))
    r = rng.randrange(ed.ORDER)
    a1 = ed.scalar_mul(ed.G, r)
    a2 = ed.scalar_mul(c1, r)
    e = opos.multihash([pk[0], pk[1], pk[0], pk[1], c1[0], c1[1], dpt[0], dpt[1], a1[0], a1[1], a2[0], a2[1]])
    z = (r + e * d) % ed.ORDER
    return pk, (c1, c2), a1, a2, z


def test_static_kat_and_batch(engine):
    rng = random.Random(50)
    items = [(PK, (C1, C2), 50, A1, A2, Z),                      # valid assignment of the reference's KAT
             (PK, (C1, C2), 50, (A1[0], 0), A2, Z),              # its invalid assignment: A1.Y = 0 (off curve)
             (PK, (C1, C2), 51, A1, A2, Z)]                      # wrong message: equations fail
    for i in range(5):
        msg = rng.randrange(1000)
        pk, ct, a1, a2, z = make_proof(rng, rng.randrange(1, ed.ORDER), msg)
        items.append((pk, ct, msg, a1, a2, z))
        if i % 2 == 0:
            items.append((pk, ct, msg, a1, a2, (z + 1) % ed.ORDER))
    n = len(items)
    flags, status = engine.elgamal_verify_decryption_proof(
        elems([c for it in items for c in it[0]]), elems([x for it in items for x in eg.serialize(it[1])]).reshape(n, 4, 32),
        elems(it[2] for it in items), elems([c for it in items for c in it[3]]), elems([c for it in items for c in it[4]]),
        elems(it[5] for it in items))
    want = [1 if eg.verify_decryption_proof(it[0], it[1], it[2], it[3], it[4], it[5]) else 0 for it in items]
    assert [int(f) for f in flags] == want
    assert want[:3] == [1, 0, 0]
    assert int(status[0]) == 0 and int(status[1]) == 4 and int(status[2]) == 0


def test_assert_decrypt(engine):
    rng = random.Random(51)
    items = []
    for i in range(10):
        d = rng.randrange(1, ed.ORDER)
        pk = ed.scalar_mul(ed.G, d)
        m = rng.randrange(1 << 16)
        ct = eg.encrypt(pk, rng.randrange(R), m)
        if i % 3 == 1:
            items.append((ct, d, m + 1))
        elif i % 3 == 2:
            items.append((ct, (d + 1) % ed.ORDER, m))
        else:
            items.append((ct, d, m))
    items.append((((1, 2), items[0][0][1]), items[0][1], items[0][2]))   # C1 off curve
    n = len(items)
    flags, status = engine.elgamal_assert_decrypt(
        elems([x for it in items for x in eg.serialize(it[0])]).reshape(n, 4, 32), elems(it[1] for it in items),
        elems(it[2] for it in items))
    want = [1 if eg.assert_decrypt(*it) else 0 for it in items]
    assert [int(f) for f in flags] == want and sum(want) == 4
    assert [int(s) for s in status] == [0] * (n - 1) + [4]


def test_te_rte_roundtrip_and_constants(engine):
    rng = random.Random(52)
    b8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
          16950150798460657717958625567821834550301663161624707787222815936182638968203)
    pt = (20284931487578954787250358776722960153090567235942462656834196519767860852891,       # twistededwards_test.go:71,73
          21185575020764391300398134415668786804224896114060668011215204645513129497221)
    pts = [b8, pt] + [(rng.randrange(R), rng.randrange(R)) for _ in range(20)]
    flat = elems([c for p in pts for c in p]).reshape(len(pts), 2, 32)
    rte, st = engine.te_to_rte(flat)
    assert not st.any()
    assert [tuple(ints(p)) for p in rte] == [ed.te_to_rte(*p) for p in pts]
    assert tuple(ints(rte[0])) == ed.G
    back, st = engine.rte_to_te(rte)
    assert (back == flat).all()
    bad, st = engine.te_to_rte(elems([R, 1]).reshape(1, 2, 32))
    assert int(st[0]) == 1


def test_eddsa_poseidon_verifier(engine):
    from oracle import eddsa as oeddsa

    rng = random.Random(53)
    items = []
    for i in range(9):
        msg = rng.getrandbits(248)                                  # 31 random bytes, verifier_test.go:44
        a, r, s = oeddsa.sign(rng.randrange(1, ed.ORDER), rng.randrange(1, ed.ORDER), msg)
        if i % 3 == 1:
            msg ^= 1
        if i % 3 == 2:
            s = (s + 1) % ed.ORDER
        items.append((a, r, s, msg))
    items.append(((1, 2), items[0][1], items[0][2], items[0][3]))   # A off curve -> PointToRTE assertion
    n = len(items)
    flags, status = engine.eddsa_verify(elems([c for it in items for c in it[0]]), elems([c for it in items for c in it[1]]),
                                        elems(it[2] for it in items), elems(it[3] for it in items))
    want = [oeddsa.is_valid(*it) for it in items]
    assert [int(f) for f in flags] == [w[0] for w in want]
    assert [int(s) == 0 for s in status] == [w[1] for w in want]
    assert [w[0] for w in want] == [1, 0, 0, 1, 0, 0, 1, 0, 0, 0] and int(status[-1]) == 4


def test_proof_entry_points_on_edge_scalars_and_each_off_curve_operand(engine):
    """Responses and keys above the subgroup order (the gadgets take Fr elements as integers, r > l), a maximal message,
    every point operand off the curve in turn, non-canonical elements, and low-order commitment points."""
    from oracle import eddsa as oeddsa

    rng = random.Random(77)
    ell = ed.ORDER
    t2 = (0, R - 1)                                                         # order 2, on the curve
    items = []
    for i in range(4):
        msg = [rng.randrange(1000), R - 1, 0, ell][i]
        pk, ct, a1, a2, z = make_proof(rng, rng.randrange(1, ell), msg)
        items.append((pk, ct, msg, a1, a2, z))
        items.append((pk, ct, msg, a1, a2, z + ell))                        # same point [z]G: still valid
        if z + 7 * ell < R:
            items.append((pk, ct, msg, a1, a2, z + 7 * ell))
        items.append((pk, ct, msg, ed.add(a1, t2), a2, z))                  # commitment moved by a low-order point
        items.append((pk, ct, (msg + 1) % R, a1, a2, z))
    pk, ct, a1, a2, z = items[0][0], items[0][1], items[0][3], items[0][4], items[0][5]
    off = (3, 4)
    assert not ed.is_on_curve(off)
    n_valid_shapes = len(items)
    items += [(off, ct, items[0][2], a1, a2, z), (pk, (off, ct[1]), items[0][2], a1, a2, z),
              (pk, (ct[0], off), items[0][2], a1, a2, z), (pk, ct, items[0][2], off, a2, z),
              (pk, ct, items[0][2], a1, off, z)]
    n = len(items)
    flags, status = engine.elgamal_verify_decryption_proof(
        elems([c for it in items for c in it[0]]), elems([x for it in items for x in eg.serialize(it[1])]).reshape(n, 4, 32),
        elems(it[2] for it in items), elems([c for it in items for c in it[3]]), elems([c for it in items for c in it[4]]),
        elems(it[5] for it in items))
    want = [1 if eg.verify_decryption_proof(*it) else 0 for it in items]
    assert [int(f) for f in flags] == want and sum(want) >= 8
    assert [int(s) for s in status] == [0] * n_valid_shapes + [4] * 5
    # non-canonical response / message: status 1, flag 0
    bad = [(pk, ct, items[0][2], a1, a2, R), (pk, ct, R + 5, a1, a2, z)]
    f2, s2 = engine.elgamal_verify_decryption_proof(
        elems([c for it in bad for c in it[0]]), elems([x for it in bad for x in eg.serialize(it[1])]).reshape(2, 4, 32),
        elems(it[2] for it in bad), elems([c for it in bad for c in it[3]]), elems([c for it in bad for c in it[4]]),
        elems(it[5] for it in bad))
    assert [int(x) for x in f2] == [0, 0] and [int(x) for x in s2] == [1, 1]

    # AssertDecrypt with private keys above the order and a key of R - 1
    dec = []
    for i in range(6):
        d = rng.randrange(1, ell)
        m = [rng.randrange(1 << 16), 0, R - 1][i % 3]
        c = eg.encrypt(ed.scalar_mul(ed.G, d), rng.randrange(R), m)
        dec.append((c, d + ell * (i % 4), m))                               # d, d + l, d + 2l, d + 3l: all < r
    dec.append((dec[0][0], R - 1, dec[0][2]))
    dec.append(((t2, dec[0][0][1]), dec[0][1], dec[0][2]))                  # C1 of order 2
    nd = len(dec)
    flags, status = engine.elgamal_assert_decrypt(
        elems([x for it in dec for x in eg.serialize(it[0])]).reshape(nd, 4, 32), elems(it[1] for it in dec),
        elems(it[2] for it in dec))
    want = [1 if eg.assert_decrypt(*it) else 0 for it in dec]
    assert [int(f) for f in flags] == want and want[:6] == [1] * 6 and not status.any()

    # EdDSA: S above the order (types.go:37-49 reduces S mod l before the multiplication), a maximal message
    sigs = []
    for i in range(4):
        msg = [rng.getrandbits(248), R - 1, 0, 1][i]
        a, r, s = oeddsa.sign(rng.randrange(1, ell), rng.randrange(1, ell), msg)
        sigs.append((a, r, s, msg))
        sigs.append((a, r, s + ell, msg))
    ns = len(sigs)
    flags, status = engine.eddsa_verify(elems([c for it in sigs for c in it[0]]), elems([c for it in sigs for c in it[1]]),
                                        elems(it[2] for it in sigs), elems(it[3] for it in sigs))
    want = [oeddsa.is_valid(*it) for it in sigs]
    assert [int(f) for f in flags] == [w[0] for w in want] and not status.any()
    assert [w[0] for w in want] == [1] * ns
