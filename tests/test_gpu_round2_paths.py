"""GPU parity for the code paths added at the end of round 2: normalize_kernel with every points-per-inversion ratio
(the launcher only raises it above 32 on batches of millions of points, so the ratio is forced here), the 9-multiply
Ciphertext.Add in both element formats, and the host-buffer Poseidon call on either side of the 256 KB zero-copy limit."""
import os
import random

import numpy as np
import pytest

from oracle import cport
from oracle import edwards as ed
from oracle import elgamal as eg
from oracle import poseidon as opos
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu

PK = ed.scalar_mul(ed.G, 0xB200)
R_MONT = (1 << 256) % R


def _mont(values):
    return elems([(v * R_MONT) % R for v in values])


@pytest.fixture
def norm_per():
    """Sets GCP_B200_NORM_PER (read by the launcher at every call) and restores it."""
    old = os.environ.get("GCP_B200_NORM_PER")

    def setter(v):
        if v is None:
            os.environ.pop("GCP_B200_NORM_PER", None)
        else:
            os.environ["GCP_B200_NORM_PER"] = str(v)

    yield setter
    setter(old)


@pytest.mark.parametrize("per", [1, 2, 7, 32, 33, 88, 128])
def test_normalisation_ratio_does_not_change_results(engine, norm_per, per):
    """Encrypt, Add and fixed-base outputs all leave through normalize_kernel: identical for every ratio, equal to the oracle
    on a sample; ragged sizes so that the last threads hold fewer points than the others."""
    rng = random.Random(900 + per)
    n = 1000 + per * 3 + 1
    ks = [rng.randrange(R) for _ in range(n)]
    ms = [rng.randrange(1 << 16) for _ in range(n)]
    norm_per(None)
    want_ct, want_st = engine.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
    want_fb, _ = engine.elgamal_fixed_base_mul(elems(ks))
    want_add, _ = engine.elgamal_add(want_ct[: n // 2], want_ct[n // 2: 2 * (n // 2)])
    norm_per(per)
    ct, st = engine.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
    fb, st2 = engine.elgamal_fixed_base_mul(elems(ks))
    add, st3 = engine.elgamal_add(ct[: n // 2], ct[n // 2: 2 * (n // 2)])
    assert not st.any() and not st2.any() and not st3.any() and not want_st.any()
    assert (ct == want_ct).all() and (fb == want_fb).all() and (add == want_add).all()
    for i in (0, 1, n // 2 - 1, n - 1):
        assert ints(ct[i]) == eg.serialize(eg.encrypt(PK, ks[i], ms[i]))
    for i in (0, n // 2 - 1):
        a, b = eg.encrypt(PK, ks[i], ms[i]), eg.encrypt(PK, ks[n // 2 + i], ms[n // 2 + i])
        assert ints(add[i]) == eg.serialize(eg.ct_add(a, b))


@pytest.mark.parametrize("per", [1, 128])
def test_zero_denominator_inside_a_long_inversion_batch(engine, norm_per, per):
    """ciphertext.go:24-32 divides by 1 +- d x1 x2 y1 y2: an item whose denominator is 0 gets status 5 and must not disturb
    the other points that share its Fermat inversion."""
    rng = random.Random(77)
    n = 300
    pts = [[rng.randrange(R) for _ in range(4)] for _ in range(n)]
    qts = [[rng.randrange(R) for _ in range(4)] for _ in range(n)]
    # make 1 + d x1 x2 y1 y2 == 0 for the C1 half of item 5: choose y2 = -1 / (d x1 x2 y1)
    x1, y1, x2 = pts[5][0], pts[5][1], qts[5][0]
    qts[5][1] = (-pow(ed.D * x1 * x2 * y1 % R, -1, R)) % R
    norm_per(per)
    out, st = engine.elgamal_add(elems([x for p in pts for x in p]).reshape(n, 4, 32),
                                 elems([x for p in qts for x in p]).reshape(n, 4, 32))
    assert st[5] == 5 and not np.delete(st, 5).any()
    for i in (0, 4, 6, n - 1):
        a = ((pts[i][0], pts[i][1]), (pts[i][2], pts[i][3]))
        b = ((qts[i][0], qts[i][1]), (qts[i][2], qts[i][3]))
        assert ints(out[i]) == eg.serialize(eg.ct_add(a, b))


def test_add_in_both_element_formats_and_coordinates(engine):
    """The 9-multiply addition takes standard-form inputs without converting them (constants 2d R^3 and 2 / R) and
    fr.Element memory as it is: same ciphertexts either way, arbitrary (off-curve) inputs included, identity, doubling,
    inverse pairs, non-canonical operands."""
    import gnark_crypto_primitives_b200 as g

    rng = random.Random(31)
    n = 257
    a = [[rng.randrange(R) for _ in range(4)] for _ in range(n)]
    b = [[rng.randrange(R) for _ in range(4)] for _ in range(n)]
    ca, cb = eg.encrypt(PK, 5, 7), eg.encrypt(PK, R - 3, 11)
    a[0], b[0] = eg.serialize(ca), eg.serialize(cb)
    a[1], b[1] = eg.serialize(ca), eg.serialize(ca)                    # doubling
    a[2], b[2] = eg.serialize(ca), eg.serialize(eg.ct_neg(ca))         # sum = identity
    a[3], b[3] = eg.serialize(eg.new_ciphertext()), eg.serialize(cb)   # identity operand
    a[4], b[4] = [0, 0, 0, 0], [R - 1, R - 1, R - 1, R - 1]
    flat_a = [x for c in a for x in c]
    flat_b = [x for c in b for x in c]
    std, st = engine.elgamal_add(elems(flat_a).reshape(n, 4, 32), elems(flat_b).reshape(n, 4, 32))
    mont, stm = engine.elgamal_add(_mont(flat_a).reshape(n, 4, 32), _mont(flat_b).reshape(n, 4, 32), fmt=g.FMT_MONTGOMERY)
    assert not st.any() and not stm.any()
    want = [eg.serialize(eg.ct_add(((x[0], x[1]), (x[2], x[3])), ((y[0], y[1]), (y[2], y[3])))) for x, y in zip(a, b)]
    assert [ints(c) for c in std] == want
    assert (mont == _mont([v for w in want for v in w]).reshape(n, 4, 32)).all()
    assert ints(std[2]) == [0, 1, 0, 1]
    # a non-canonical coordinate is status 1 in either format and the item's output is zeroed
    bad = elems(flat_a).reshape(n, 4, 32).copy()
    bad[7, 2] = np.frombuffer(R.to_bytes(32, "little"), dtype=np.uint8)
    out, st = engine.elgamal_add(bad, elems(flat_b).reshape(n, 4, 32))
    assert st[7] == 1 and not np.delete(st, 7).any() and not out[7].any()
    assert [ints(c) for c in np.delete(out, 7, axis=0)] == want[:7] + want[8:]


@pytest.mark.parametrize("arity,n", [(2, 4096), (2, 4097), (1, 8192), (1, 8193), (16, 512), (16, 513), (3, 2730), (3, 2731)])
def test_host_hash_on_both_sides_of_the_zero_copy_limit(engine, arity, n):
    """gcp_poseidon_hash: up to 256 KB of inputs run on mapped page-locked memory, one byte more goes through the staged
    copies; same digests, status bytes included."""
    rng = np.random.default_rng(arity * 100003 + n)
    a = rng.integers(0, 256, size=(n, arity, 32), dtype=np.uint8)
    a[:, :, 31] &= 0x1F                                  # < 2^253 < r: canonical
    a[3, 0] = np.frombuffer(R.to_bytes(32, "little"), dtype=np.uint8)   # one non-canonical row
    out, st = engine.poseidon_hash(a)
    want, wst = cport.poseidon_hash(a, threads=8)
    assert (st == wst).all() and st[3] == 1 and int(st.sum()) == 1
    ok = st == 0
    assert (out[ok] == want[ok]).all()
    for i in (0, n - 1):
        assert ints(out[i:i + 1])[0] == opos.hash(ints(a[i]))
    # repeated calls reuse the mapped buffer: a second, smaller call must not see stale rows
    out2, st2 = engine.poseidon_hash(a[5:9])
    assert (out2 == out[5:9]).all() and not st2.any()
