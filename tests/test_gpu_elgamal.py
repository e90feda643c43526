"""GPU parity: ElGamal (fixed-base, encrypt, add, neg, tally) through the C ABI vs the oracle (bit-exact points)."""
import random

import numpy as np
import pytest

from oracle import cport
from oracle import edwards as ed
from oracle import elgamal as eg
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu

PK_D = 0xB200
PK = ed.scalar_mul(ed.G, PK_D)


def ct_ints(arr):
    return [ints(c) for c in arr]


def test_fixed_base_matches_reference_window_semantics(engine):
    rng = random.Random(1)
    ks = [0, 1, 2, 15, 16, 255, 256, 12345, 67890, ed.ORDER - 1, ed.ORDER, ed.ORDER + 1, R - 1, (1 << 253) + 7] + \
        [rng.randrange(R) for _ in range(50)]
    out, st = engine.elgamal_fixed_base_mul(elems(ks))
    assert not st.any()
    assert [tuple(ints(p)) for p in out] == [eg.fixed_base_scalar_mul(k) for k in ks]  # mul.go:76-166 restated
    assert tuple(ints(out[0])) == ed.IDENTITY


def test_encrypt_compatibility_vector(engine):
    """elgamal/encrypt_test.go:144-159: PubKey = G, k = 12345, m = 67890."""
    out, st = engine.elgamal_encrypt(elems(ed.G), elems([12345]), elems([67890]))
    assert not st.any()
    assert ints(out[0]) == [19918712023960437102123786886411468478902094248734671506464346041569881874462,
                            13276557205153692030187527501273228448057533426731746626187331221465573305487,
                            839909438842816078619007291947839299631027001100725795038371896036563634986,
                            2683297034354865619026157779551553408927035959921603566752318144387126273226]
    # EncryptedZero(k) == Encrypt(k, 0)  (encrypt_test.go:125-139)
    z, _ = engine.elgamal_encrypt(elems(ed.G), elems([12345]), elems([0]))
    assert ints(z[0]) == eg.serialize(eg.encrypted_zero(ed.G, 12345))


def test_encrypt_with_specific_data(engine):
    """elgamal/encrypt_test.go:180-217: fixed pubkey, k1..k3 (two above the subgroup order), m = 0."""
    pk = (18604149248430057540085528196797394191454458259161233471314599389622530831795,
          1988784568828097512630242539176296837964596457792502130892628909648459248949)
    ks = [855131146298194990003384743709896434741839908245,
          5883442530210657871581412827617735506655215369087356134218551734599178232070,
          3979028711588105728532079493967382119023185938755564152610807942458151212832]
    out, st = engine.elgamal_encrypt(elems(pk), elems(ks), elems([0, 0, 0]))
    assert not st.any()
    assert ct_ints(out) == [eg.serialize(eg.encrypt(pk, k, 0)) for k in ks]


def test_encrypt_shared_key_batch(engine):
    rng = random.Random(2)
    n = 300
    ks = [rng.randrange(R) for _ in range(n)]
    ms = [rng.randrange(1 << 16) for _ in range(n)]
    ks[:4] = [0, 1, R - 1, ed.ORDER]
    ms[:4] = [0, R - 1, 1, rng.randrange(R)]
    out, st = engine.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
    assert not st.any()
    want, wst = cport.elgamal_encrypt(elems(PK), elems(ks), elems(ms), threads=8)
    assert not wst.any() and (out == want).all()
    for i in (0, 1, 2, 3, 17):
        assert ints(out[i]) == eg.serialize(eg.encrypt(PK, ks[i], ms[i]))


def test_encrypt_per_item_keys_and_failures(engine):
    rng = random.Random(3)
    n = 12
    pks = [ed.scalar_mul(ed.G, 5 + i) for i in range(n)]
    ks = [rng.randrange(R) for _ in range(n)]
    ms = [rng.randrange(1 << 16) for _ in range(n)]
    pks[3] = (1, 2)            # off curve -> AssertIsOnCurve fails (encrypt.go:49)
    ks[5] = R                  # non-canonical scalar
    pks[7] = (R + 1, pks[7][1])
    out, st = engine.elgamal_encrypt(elems([c for p in pks for c in p]).reshape(n, 2, 32), elems(ks), elems(ms))
    assert [int(s) for s in st] == [0, 0, 0, 4, 0, 1, 0, 1, 0, 0, 0, 0]
    for i in range(n):
        if st[i] == 0:
            assert ints(out[i]) == eg.serialize(eg.encrypt(pks[i], ks[i], ms[i])), i
    # shared off-curve key: every item is flagged
    out2, st2 = engine.elgamal_encrypt(elems((1, 2)), elems(ks[:3]), elems(ms[:3]))
    assert [int(s) for s in st2] == [4, 4, 4]
    # and the cache recovers when a good key follows
    out3, st3 = engine.elgamal_encrypt(elems(PK), elems(ks[:3]), elems(ms[:3]))
    assert not st3.any() and ints(out3[0]) == eg.serialize(eg.encrypt(PK, ks[0], ms[0]))


def test_add_neg_elementwise(engine):
    rng = random.Random(4)
    n = 40
    a = [eg.encrypt(PK, rng.randrange(R), rng.randrange(100)) for _ in range(n)]
    b = [eg.encrypt(PK, rng.randrange(R), rng.randrange(100)) for _ in range(n)]
    b[0] = a[0]                                   # doubling through the unified law
    b[1] = eg.ct_neg(a[1])                        # sum = identity
    a[2] = eg.new_ciphertext()                    # identity operand (ciphertext.go:16-19)
    fa = elems([x for c in a for x in eg.serialize(c)]).reshape(n, 4, 32)
    fb = elems([x for c in b for x in eg.serialize(c)]).reshape(n, 4, 32)
    out, st = engine.elgamal_add(fa, fb)
    assert not st.any()
    assert ct_ints(out) == [eg.serialize(eg.ct_add(x, y)) for x, y in zip(a, b)]
    assert ints(out[1]) == [0, 1, 0, 1]
    neg, st = engine.elgamal_neg(fa)
    assert not st.any() and ct_ints(neg) == [eg.serialize(eg.ct_neg(x)) for x in a]
    back, _ = engine.elgamal_neg(neg)
    assert (back == fa).all()


def test_add_does_not_require_curve_points(engine):
    """ciphertext.go:24-32 performs no on-curve check: garbage in, the same deterministic garbage out."""
    rng = random.Random(5)
    pts = [[rng.randrange(R) for _ in range(4)] for _ in range(6)]
    qts = [[rng.randrange(R) for _ in range(4)] for _ in range(6)]
    out, st = engine.elgamal_add(elems([x for p in pts for x in p]).reshape(6, 4, 32),
                                 elems([x for p in qts for x in p]).reshape(6, 4, 32))
    assert not st.any()
    for i in range(6):
        a = ((pts[i][0], pts[i][1]), (pts[i][2], pts[i][3]))
        b = ((qts[i][0], qts[i][1]), (qts[i][2], qts[i][3]))
        assert ints(out[i]) == eg.serialize(eg.ct_add(a, b))


@pytest.mark.parametrize("n_ballots,n_fields", [(1, 1), (7, 3), (64, 8), (333, 8), (50, 64), (0, 4)])
def test_tally_matches_fold(engine, n_ballots, n_fields):
    rng = random.Random(n_ballots * 100 + n_fields)
    n = n_ballots * n_fields
    ks = [rng.randrange(R) for _ in range(n)]
    ms = [rng.randrange(1 << 16) for _ in range(n)]
    if n:
        cts, st = cport.elgamal_encrypt(elems(PK), elems(ks), elems(ms), threads=8)
        cts = cts.reshape(n_ballots, n_fields, 4, 32)
    else:
        cts = np.zeros((0, n_fields, 4, 32), np.uint8)
    out, st = engine.elgamal_tally(cts)
    assert not st.any()
    # closed form: sum Encrypt(pk, k_i, m_i) == Encrypt(pk, sum k_i mod l, sum m_i mod l) for pk in <G>
    for f in range(n_fields):
        ksum = sum(ks[b * n_fields + f] for b in range(n_ballots)) % ed.ORDER
        msum = sum(ms[b * n_fields + f] for b in range(n_ballots)) % ed.ORDER
        assert ints(out[f]) == eg.serialize(eg.encrypt(PK, ksum, msum)), f
    if n:
        want, _ = cport.elgamal_tally(cts)
        assert (out == want).all()


def test_tally_is_shard_invariant(engine):
    """Multi-GPU rule: tally(shards' partials) == tally(all) bit for bit (associativity)."""
    rng = random.Random(77)
    nb, nf = 96, 8
    ks = [rng.randrange(R) for _ in range(nb * nf)]
    ms = [rng.randrange(1 << 16) for _ in range(nb * nf)]
    cts, _ = cport.elgamal_encrypt(elems(PK), elems(ks), elems(ms), threads=8)
    cts = cts.reshape(nb, nf, 4, 32)
    whole, _ = engine.elgamal_tally(cts)
    for shards in (2, 3, 8):
        parts = [engine.elgamal_tally(c)[0] for c in np.array_split(cts, shards)]
        again, _ = engine.elgamal_tally(np.stack(parts))
        assert (again == whole).all()


def test_montgomery_format(engine):
    import gnark_crypto_primitives_b200 as g

    rng = random.Random(6)
    M = 1 << 256
    ks = [rng.randrange(R) for _ in range(5)]
    ms = [rng.randrange(1000) for _ in range(5)]
    to_m = lambda xs: elems([x * M % R for x in xs])
    out, st = engine.elgamal_encrypt(to_m(PK), to_m(ks), to_m(ms), fmt=g.FMT_MONTGOMERY)
    assert not st.any()
    rinv = pow(M, -1, R)
    got = [[v * rinv % R for v in ints(c)] for c in out]
    assert got == [eg.serialize(eg.encrypt(PK, k, m)) for k, m in zip(ks, ms)]
    tal, st = engine.elgamal_tally(out.reshape(5, 1, 4, 32), fmt=g.FMT_MONTGOMERY)
    want = eg.serialize(eg.tally([eg.encrypt(PK, k, m) for k, m in zip(ks, ms)]))
    assert [v * rinv % R for v in ints(tal[0])] == want


def test_fused_encrypt_tally(engine):
    rng = random.Random(8)
    for nb, nf in ((1, 1), (9, 8), (200, 3), (0, 2)):
        ks = [rng.randrange(R) for _ in range(nb * nf)]
        ms = [rng.randrange(1 << 16) for _ in range(nb * nf)]
        k = elems(ks).reshape(nb, nf, 32)
        m = elems(ms).reshape(nb, nf, 32)
        out, st = engine.elgamal_encrypt_tally(elems(PK), k, m)
        assert not st.any()
        for f in range(nf):
            ksum = sum(ks[b * nf + f] for b in range(nb)) % ed.ORDER
            msum = sum(ms[b * nf + f] for b in range(nb)) % ed.ORDER
            assert ints(out[f]) == eg.serialize(eg.encrypt(PK, ksum, msum))
        if nb:
            cts, _ = engine.elgamal_encrypt(elems(PK), k.reshape(-1, 32), m.reshape(-1, 32))
            tal, _ = engine.elgamal_tally(cts.reshape(nb, nf, 4, 32))
            assert (tal == out).all()
    # off-curve key: every field flagged
    out, st = engine.elgamal_encrypt_tally(elems((1, 2)), elems([1, 2]).reshape(1, 2, 32), elems([3, 4]).reshape(1, 2, 32))
    assert [int(s) for s in st] == [4, 4] and not out.any()


def test_device_forms_report_an_off_curve_key(engine):
    """AssertIsOnCurve(pubKey) (encrypt.go:49) on the device-resident forms: gcp_elgamal_encrypt_tally_dev and
    gcp_ballot_batch_dev must give status 4 and no result, not the tally under a table built from the identity."""
    import torch

    from tests.util import dense_proof

    rng = random.Random(49)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nb, nf = 5, 3
    k = dev(elems(rng.randrange(R) for _ in range(nb * nf)))
    m = dev(elems(rng.randrange(1 << 16) for _ in range(nb * nf)))
    out = torch.full((nf, 4, 32), 7, dtype=torch.uint8, device="cuda")
    st = torch.full((nf,), 9, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    for bad_pk in ((1, 2), (R, 1)):                                          # off the curve / not a field element
        engine.elgamal_encrypt_tally_dev(dev(elems(bad_pk)), k, m, nb, nf, out, st, stream=stream)
        torch.cuda.synchronize()
        assert st.tolist() == [4] * nf and not bool(out.any())
    engine.elgamal_encrypt_tally_dev(dev(elems(PK)), k, m, nb, nf, out, st, stream=stream)
    torch.cuda.synchronize()
    assert st.tolist() == [0] * nf and bool(out.any())
    # ballot batch
    n_levels = 24
    items = [dense_proof(rng, n_levels) for _ in range(nb)]
    d = dict(roots=dev(elems(it[0] for it in items)), sib=dev(elems([s for it in items for s in it[1]])),
             keys=dev(elems(it[2] for it in items)), vals=dev(elems(it[3] for it in items)))
    flags = torch.empty(nb, dtype=torch.uint8, device="cuda")
    pst = torch.empty(nb, dtype=torch.uint8, device="cuda")
    engine.ballot_batch_dev(n_levels, nb, d["roots"], False, d["sib"], d["keys"], d["vals"], dev(elems((1, 2))), k, m, nf,
                            flags, pst, out, st, stream=stream)
    torch.cuda.synchronize()
    assert flags.tolist() == [1] * nb and st.tolist() == [4] * nf and not bool(out.any())


def test_field_edge_patterns_through_curve_arithmetic(engine):
    """Ciphertext.Add on carry-hostile coordinates (no on-curve check in the reference) vs the C oracle."""
    from tests.test_gpu_poseidon import _edge_values

    vals = _edge_values()
    n = len(vals)
    a = elems([vals[(i * k + k) % n] for i in range(n) for k in (1, 3, 5, 7)]).reshape(n, 4, 32)
    b = elems([vals[(i * k + 2 * k + 1) % n] for i in range(n) for k in (11, 13, 17, 19)]).reshape(n, 4, 32)
    out, st = engine.elgamal_add(a, b)
    want, wst = cport.elgamal_add(a, b, threads=8)
    assert (st == wst).all()
    ok = st == 0
    assert ok.sum() > n // 2
    assert (out[ok] == want[ok]).all()


def test_is_equal_and_select(engine):
    """(*Ciphertext).IsEqual / Select (elgamal/ciphertext.go:79-96)."""
    rng = random.Random(7996)
    n = 300
    a = np.frombuffer(bytes(rng.getrandbits(8) for _ in range(n * 128)), dtype=np.uint8).reshape(n, 4, 32).copy()
    a[:, :, 31] &= 0x0F                                   # canonical
    b = a.copy()
    for i in range(0, n, 3):                              # every third differs in one byte of one coordinate
        b[i, i % 4, (i * 7) % 31] ^= 1 << (i % 8)
    flags, st = engine.elgamal_is_equal(a, b)
    assert not st.any() and [int(f) for f in flags] == [0 if i % 3 == 0 else 1 for i in range(n)]
    a2 = a.copy()
    a2[5, 2] = np.frombuffer(int(R).to_bytes(32, "little"), np.uint8)   # non-canonical coordinate
    flags, st = engine.elgamal_is_equal(a2, a2)
    assert int(st[5]) == 1 and int(flags[5]) == 0 and int(flags[6]) == 1
    sel = np.array([i % 2 for i in range(n)], np.uint8)
    sel[9] = 2                                            # api.Select asserts a boolean
    out, st = engine.elgamal_select(sel, a, b)
    for i in range(n):
        if i == 9:
            assert int(st[i]) == 3 and not out[i].any()
        else:
            assert int(st[i]) == 0 and np.array_equal(out[i], a[i] if sel[i] else b[i]), i
    out, st = engine.elgamal_select(np.zeros(0, np.uint8), np.zeros((0, 4, 32), np.uint8), np.zeros((0, 4, 32), np.uint8))
    assert out.shape == (0, 4, 32)


def _window_boundary_scalars():
    """Scalars that sit on the recoding boundaries of the engine's signed windows (20-bit fixed-base, 4-bit variable-base):
    digits exactly half a window (the tie), all-ones runs whose carry ripples through every window, single bits at
    window edges, and the largest canonical values."""
    ks = set()
    for w in (4, 20):
        half, full = 1 << (w - 1), 1 << w
        for i in range(0, 254, w):
            for d in (half - 1, half, half + 1, full - 1):
                ks.add((d << i) % R)
            ks.add(((1 << i) - 1) % R)
            ks.add((1 << i) % R)
        ks.add(sum(half << i for i in range(0, 240, w)) % R)             # every digit on the tie
        ks.add(sum((half + 1) << i for i in range(0, 240, w)) % R)
        ks.add(sum((full - 1) << i for i in range(0, 240, 2 * w)) % R)   # alternating full / empty digits
    ks |= {0, 1, R - 1, R - 2, ed.ORDER - 1, ed.ORDER, ed.ORDER + 1, 2 * ed.ORDER, 7 * ed.ORDER + 5, (1 << 253) - 1, 1 << 253}
    return sorted(ks)


def test_window_boundary_scalars_fixed_and_variable_base(engine):
    ks = _window_boundary_scalars()
    n = len(ks)
    assert n > 300
    out, st = engine.elgamal_fixed_base_mul(elems(ks))
    assert not st.any()
    ms = [ks[(7 * i + 3) % n] for i in range(n)]
    # shared key (fixed-base tables for G and PK) and per-item keys (windowed variable-base) against the C oracle
    enc, st = engine.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
    want, wst = cport.elgamal_encrypt(elems(PK), elems(ks), elems(ms), threads=8)
    assert not st.any() and not wst.any() and (enc == want).all()
    assert (enc[:, :2] == out).all()                                     # C1 = [k]G
    pks = elems([c for _ in range(n) for c in PK]).reshape(n, 2, 32)
    enc2, st2 = engine.elgamal_encrypt(pks, elems(ks), elems(ms))
    assert not st2.any() and (enc2 == want).all()
    # and a sample against the literal Python restatement of mul.go:76-166
    for i in list(range(0, n, 37)) + [n - 1]:
        assert tuple(ints(out[i])) == eg.fixed_base_scalar_mul(ks[i]), hex(ks[i])


def test_public_keys_with_a_cofactor_component(engine):
    """Per-item keys outside <G>: the identity, the points of order 2, 4 and 8, and sums of those with subgroup points.
    They pass AssertIsOnCurve (encrypt.go:49), so Encrypt goes on to [k]pk; the engine must return the group-law value
    (oracle: plain double-and-add).  SURVEY 8c lists gnark's hinted ScalarMul on such keys as an edge the reference
    tree does not pin; the mathematical value is what is checked here."""
    t8 = (438929327410846936349781937479275998929723622237267896546408861753155634789,
          4826523245007015323400664741523384119579596407052839571721035538011798951543)
    low = [ed.IDENTITY, (0, R - 1), ed.scalar_mul(t8, 2), t8, ed.scalar_mul(t8, 3), ed.scalar_mul(t8, 7)]
    assert all(ed.is_on_curve(p) for p in low) and ed.scalar_mul(t8, 8) == ed.IDENTITY and ed.scalar_mul(t8, 4) == (0, R - 1)
    rng = random.Random(88)
    pks = low + [ed.add(ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)), t) for t in low[1:]]
    ks = [rng.randrange(R) for _ in pks]
    ks[0], ks[1], ks[2], ks[3] = 5, 7, 6, ed.ORDER          # odd multiple of the order-2 point, multiple of l on order 8
    ms = [rng.randrange(1 << 16) for _ in pks]
    n = len(pks)
    out, st = engine.elgamal_encrypt(elems([c for p in pks for c in p]).reshape(n, 2, 32), elems(ks), elems(ms))
    assert not st.any()
    for i in range(n):
        assert ints(out[i]) == eg.serialize(eg.encrypt(pks[i], ks[i], ms[i])), i
    # the same keys through the shared-key path (fixed-base table built from the key)
    for i in (1, 3, n - 1):
        o, s = engine.elgamal_encrypt(elems(pks[i]), elems(ks[:4]), elems(ms[:4]))
        assert not s.any()
        assert ct_ints(o) == [eg.serialize(eg.encrypt(pks[i], k, m)) for k, m in zip(ks[:4], ms[:4])], i
    # ciphertext addition and tally with low-order points as operands
    cts = [(low[i % 6], low[(i + 1) % 6]) for i in range(12)]
    flat = elems([x for c in cts for x in eg.serialize(c)]).reshape(12, 4, 32)
    add, st = engine.elgamal_add(flat[:6], flat[6:])
    assert not st.any() and ct_ints(add) == [eg.serialize(eg.ct_add(a, b)) for a, b in zip(cts[:6], cts[6:])]
    tal, st = engine.elgamal_tally(flat.reshape(12, 1, 4, 32))
    assert not st.any() and ct_ints(tal) == [eg.serialize(eg.tally(cts))]


def test_scalar_mul_single_and_double_base(engine):
    """curve.ScalarMul on its own (call sites elgamal/encrypt.go:55, ciphertext.go:58,147-160): [s]P and the one-pass
    [s]P + [s2]P2, window-boundary scalars and keys with a cofactor component included, both element formats."""
    rng = random.Random(4242)
    t8 = (438929327410846936349781937479275998929723622237267896546408861753155634789,
          4826523245007015323400664741523384119579596407052839571721035538011798951543)
    ks = _window_boundary_scalars()[::5] + [rng.randrange(R) for _ in range(40)]
    n = len(ks)
    pts = [ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)) for _ in range(n)]
    pts[0], pts[1], pts[2] = ed.IDENTITY, t8, ed.add(pts[2], t8)
    out, st = engine.elgamal_scalar_mul(elems([c for p in pts for c in p]).reshape(n, 2, 32), elems(ks))
    assert not st.any()
    for i in range(n):
        assert tuple(ints(out[i])) == ed.scalar_mul(pts[i], ks[i]), (i, hex(ks[i]))
    ks2 = [ks[(3 * i + 1) % n] for i in range(n)]
    pts2 = [pts[(5 * i + 2) % n] for i in range(n)]
    out2, st = engine.elgamal_scalar_mul(elems([c for p in pts for c in p]).reshape(n, 2, 32), elems(ks),
                                         elems([c for p in pts2 for c in p]).reshape(n, 2, 32), elems(ks2))
    assert not st.any()
    for i in range(n):
        assert tuple(ints(out2[i])) == ed.add(ed.scalar_mul(pts[i], ks[i]), ed.scalar_mul(pts2[i], ks2[i])), i
    # gnark-crypto element memory (Montgomery) in and out
    mont = lambda v: (v << 256) % R  # noqa: E731
    outm, st = engine.elgamal_scalar_mul(elems([mont(c) for p in pts for c in p]).reshape(n, 2, 32), elems([mont(k) for k in ks]),
                                         fmt=1)
    assert not st.any() and ints(outm) == [mont(v) for v in ints(out)]
    # assertions: off-curve point -> status 4, non-canonical scalar -> status 1, outputs zeroed
    bad_p = elems([1, 2, *pts[3]]).reshape(2, 2, 32)
    o, s = engine.elgamal_scalar_mul(bad_p, elems([5, R]))
    assert list(s) == [4, 1] and not o.any()
