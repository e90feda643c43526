"""GPU parity: GCP_COORDS_TE - curve points in iden3 / circom twisted-Edwards coordinates at the boundary of the ElGamal
entry points.  Every call must equal FromTEtoRTE (ecc/format/twistededwards.go:42-48) on each input
point, the gadget on gnark's reduced curve, FromRTEtoTE (:29-37) on each output point, composed in the oracle."""
import random

import numpy as np
import pytest

import gnark_crypto_primitives_b200 as g
from oracle import edwards as ed
from oracle import elgamal as eg
from oracle.field import R
from tests.test_gpu_proofs import make_proof
from tests.util import dense_proof, elems, ints

pytestmark = pytest.mark.gpu

TE = g.COORDS_TE
R_MONT = (1 << 256) % R


def te(p):
    return ed.rte_to_te(*p)


def ct_te(ct):
    return [c for p in ct for c in te(p)]


def test_conversion_constants():
    """scalingFactor (ecc/format/twistededwards.go:17): iden3 B8 maps to gnark's generator."""
    b8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
          16950150798460657717958625567821834550301663161624707787222815936182638968203)
    assert ed.te_to_rte(*b8) == ed.G and te(ed.G) == b8


def test_encrypt_fixed_base_add_tally_in_te(engine):
    rng = random.Random(2948)
    n, nf = 24, 3
    d = rng.randrange(1, ed.ORDER)
    pk = ed.scalar_mul(ed.G, d)
    ks = [rng.randrange(R) for _ in range(n)]
    ms = [rng.randrange(1 << 16) for _ in range(n)]
    cts = [eg.encrypt(pk, k, m) for k, m in zip(ks, ms)]
    want = [ct_te(c) for c in cts]
    # shared key
    out, st = engine.elgamal_encrypt(elems(te(pk)), elems(ks), elems(ms), fmt=TE)
    assert not st.any() and [ints(o) for o in out] == want
    # the RTE key read as a TE key is some other point: the flag must matter
    out_wrong, st_wrong = engine.elgamal_encrypt(elems(pk), elems(ks), elems(ms), fmt=TE)
    assert st_wrong.all() or [ints(o) for o in out_wrong] != want
    # a key per item
    pks = [ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)) for _ in range(n)]
    pks[5] = (1, 2)                                                            # off the curve in either coordinate system
    out2, st2 = engine.elgamal_encrypt(elems([c for p in pks for c in (te(p) if p != (1, 2) else p)]).reshape(n, 2, 32),
                                       elems(ks), elems(ms), fmt=TE)
    for i in range(n):
        if i == 5:
            assert int(st2[i]) == g.STATUS_OFF_CURVE
        else:
            assert int(st2[i]) == 0 and ints(out2[i]) == ct_te(eg.encrypt(pks[i], ks[i], ms[i]))
    # fixed base: TE only on the way out
    pts, st3 = engine.elgamal_fixed_base_mul(elems(ks), fmt=TE)
    assert not st3.any() and [tuple(ints(p)) for p in pts] == [te(ed.scalar_mul(ed.G, k % ed.ORDER)) for k in ks]
    # element-wise Add and the tally, TE in and TE out, both element formats
    a = elems([x for w in want for x in w]).reshape(n, 4, 32)
    b = np.roll(a, 1, axis=0)
    s, st4 = engine.elgamal_add(a, b, fmt=TE)
    want_sum = [ct_te((ed.add(cts[i][0], cts[i - 1][0]), ed.add(cts[i][1], cts[i - 1][1]))) for i in range(n)]
    assert not st4.any() and [ints(x) for x in s] == want_sum
    tal, st5 = engine.elgamal_tally(a.reshape(n // nf, nf, 4, 32), fmt=TE)
    want_tal = []
    for f in range(nf):
        acc = ((0, 1), (0, 1))
        for bidx in range(n // nf):
            c = cts[bidx * nf + f]
            acc = (ed.add(acc[0], c[0]), ed.add(acc[1], c[1]))
        want_tal.append(ct_te(acc))
    assert not st5.any() and [ints(x) for x in tal] == want_tal
    a_m = elems([(x * R_MONT) % R for w in want for x in w]).reshape(n // nf, nf, 4, 32)
    tal_m, st6 = engine.elgamal_tally(a_m, fmt=TE | g.FMT_MONTGOMERY)
    assert not st6.any() and [ints(x) for x in tal_m] == [[(v * R_MONT) % R for v in w] for w in want_tal]
    # fused encrypt + tally: TE key in, TE tally out
    k3 = elems(ks).reshape(n // nf, nf, 32)
    m3 = elems(ms).reshape(n // nf, nf, 32)
    tal2, st7 = engine.elgamal_encrypt_tally(elems(te(pk)), k3, m3, fmt=TE)
    assert not st7.any() and (tal2 == tal).all()
    # formats outside the two flag bits are rejected, and the SMT entry points do not take the coordinate flag
    with pytest.raises(g.EngineError):
        engine.elgamal_add(a, b, fmt=4)
    with pytest.raises(g.EngineError):
        engine.poseidon_hash(elems([1, 2]).reshape(1, 2, 32), fmt=TE)


def test_scalar_mul_and_proofs_in_te(engine):
    rng = random.Random(2949)
    n = 10
    pts = [ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)) for _ in range(n)]
    pts2 = [ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)) for _ in range(n)]
    s1 = [rng.randrange(R) for _ in range(n)]
    s2 = [rng.randrange(R) for _ in range(n)]
    flat = lambda ps: elems([c for p in ps for c in te(p)]).reshape(len(ps), 2, 32)
    out, st = engine.elgamal_scalar_mul(flat(pts), elems(s1), fmt=TE)
    assert not st.any() and [tuple(ints(o)) for o in out] == [te(ed.scalar_mul(p, s % ed.ORDER)) for p, s in zip(pts, s1)]
    out, st = engine.elgamal_scalar_mul(flat(pts), elems(s1), flat(pts2), elems(s2), fmt=TE)
    want = [te(ed.add(ed.scalar_mul(p, a % ed.ORDER), ed.scalar_mul(q, b % ed.ORDER))) for p, a, q, b in zip(pts, s1, pts2, s2)]
    assert not st.any() and [tuple(ints(o)) for o in out] == want
    # AssertDecrypt
    items = []
    for i in range(n):
        d = rng.randrange(1, ed.ORDER)
        msg = rng.randrange(1000)
        items.append((eg.encrypt(ed.scalar_mul(ed.G, d), rng.randrange(ed.ORDER), msg), d, msg + (1 if i == 3 else 0)))
    f, st = engine.elgamal_assert_decrypt(elems([x for it in items for x in ct_te(it[0])]).reshape(n, 4, 32),
                                          elems(it[1] for it in items), elems(it[2] for it in items), fmt=TE)
    assert not st.any() and [int(x) for x in f] == [0 if i == 3 else 1 for i in range(n)]
    # the same ciphertexts WITHOUT the flag are not on gnark's curve
    f2, st2 = engine.elgamal_assert_decrypt(elems([x for it in items for x in ct_te(it[0])]).reshape(n, 4, 32),
                                            elems(it[1] for it in items), elems(it[2] for it in items))
    assert (st2 == g.STATUS_OFF_CURVE).all() and not f2.any()
    # DecryptionProof.Verify: the Fiat-Shamir hash is over the gadget's (reduced) coordinates
    proofs = []
    for i in range(6):
        msg = rng.randrange(1000)
        pk, ct, a1, a2, z = make_proof(rng, rng.randrange(1, ed.ORDER), msg)
        proofs.append((pk, ct, msg, a1, a2, (z + (1 if i == 4 else 0)) % ed.ORDER))
    pt = lambda k: elems([c for it in proofs for c in te(it[k])]).reshape(len(proofs), 2, 32)
    f, st = engine.elgamal_verify_decryption_proof(
        pt(0), elems([x for it in proofs for x in ct_te(it[1])]).reshape(len(proofs), 4, 32), elems(it[2] for it in proofs),
        pt(3), pt(4), elems(it[5] for it in proofs), fmt=TE)
    assert not st.any() and [int(x) for x in f] == [1, 1, 1, 1, 0, 1]


def test_ballot_batch_in_te(engine):
    """Config 5 with a circom-side election key: key in TE, tally out in TE, the census proofs untouched."""
    rng = random.Random(2950)
    n_levels, nv, nf = 20, 12, 2
    items = [dense_proof(rng, n_levels) for _ in range(nv)]
    items[4] = (items[4][0] ^ 1,) + items[4][1:]
    pk = ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER))
    ks = [rng.randrange(R) for _ in range(nv * nf)]
    ms = [rng.randrange(1 << 16) for _ in range(nv * nf)]
    sib = elems([s for it in items for s in it[1]]).reshape(nv, n_levels, 32)
    args = (n_levels, elems(it[0] for it in items), elems(it[2] for it in items), elems(it[3] for it in items))
    k3, m3 = elems(ks).reshape(nv, nf, 32), elems(ms).reshape(nv, nf, 32)
    f1, s1, t1, ts1 = engine.ballot_batch(*args, elems(pk), k3, m3, siblings=sib)
    f2, s2, t2, ts2 = engine.ballot_batch(*args, elems(te(pk)), k3, m3, siblings=sib, fmt=TE)
    assert (f1 == f2).all() and (s1 == s2).all() and not ts1.any() and not ts2.any() and int(f1.sum()) == nv - 1
    rte = [ints(t) for t in t1]
    assert [ints(t) for t in t2] == [[c for p in ((w[0], w[1]), (w[2], w[3])) for c in te(p)] for w in rte]
