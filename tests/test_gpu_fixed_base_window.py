"""GPU parity: the window width of the fixed-base tables (elgamal/mul.go:26-72 has 4-bit windows; the engine's tables carry
their own width) changes nothing in the results: FixedBaseScalarMulBN254, Encrypt with a shared key, the fused tally and
AssertDecrypt against the oracle for narrow, odd (windows that straddle limb boundaries) and wide tables, the automatic
widening once a base has served enough multiplications, and a key change after it."""
import os
import random

import numpy as np
import pytest

from oracle import edwards as ed
from oracle import elgamal as eg
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu

PK = ed.scalar_mul(ed.G, 0xB200)


def boundary_scalars(bits):
    """Scalars whose signed windows sit on the recoding boundaries of `bits`-bit windows."""
    half, full = 1 << (bits - 1), 1 << bits
    ks = [0, 1, half - 1, half, half + 1, full - 1, full, full + 1, ed.ORDER - 1, ed.ORDER, ed.ORDER + 1, R - 1]
    for w in (1, 2, (254 // bits) - 1, 254 // bits):
        for d in (half - 1, half, half + 1, full - 1):
            ks.append((d << (bits * w)) % R)
            ks.append(((d << (bits * w)) + (half << (bits * (w - 1)))) % R)
    ks.append(sum(half << (bits * w) for w in range(256 // bits)) % R)        # every digit on the boundary
    ks.append(sum((half + 1) << (bits * w) for w in range(256 // bits)) % R)  # every digit negative with a carry
    return ks


@pytest.fixture()
def own_engine():
    import gnark_crypto_primitives_b200 as g

    eng = g.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("bits", [8, 13, 17, 20, 22])
def test_results_do_not_depend_on_the_window_width(own_engine, bits):
    eng = own_engine
    rng = random.Random(bits)
    eng.set_fixed_base_window(bits)
    assert eng.fixed_base_window(0) == bits
    ks = boundary_scalars(bits) + [rng.randrange(R) for _ in range(40)]
    out, st = eng.elgamal_fixed_base_mul(elems(ks))
    assert not st.any()
    assert [tuple(ints(p)) for p in out] == [ed.scalar_mul(ed.G, k) for k in ks]
    ms = [rng.randrange(1 << 16) for _ in ks]
    ms[:3] = [0, R - 1, ed.ORDER]
    ct, st = eng.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
    assert not st.any() and eng.fixed_base_window(1) == bits
    for i in list(range(12)) + [len(ks) - 1, len(ks) - 41, len(ks) - 42]:
        assert ints(ct[i]) == eg.serialize(eg.encrypt(PK, ks[i], ms[i])), i
    # fused encrypt + tally == closed form over the same scalars (Sum Encrypt(k_i, m_i) = Encrypt(Sum k_i, Sum m_i))
    nb = len(ks) // 2
    k2 = elems(ks[:2 * nb]).reshape(nb, 2, 32)
    m2 = elems(ms[:2 * nb]).reshape(nb, 2, 32)
    tally, st = eng.elgamal_encrypt_tally(elems(PK), k2, m2)
    assert not st.any()
    for f in range(2):
        ksum = sum(ks[f:2 * nb:2]) % ed.ORDER
        msum = sum(ms[f:2 * nb:2]) % ed.ORDER
        assert ints(tally[f]) == eg.serialize(eg.encrypt(PK, ksum, msum))
    # iden3 coordinates at the boundary: the key's table is built from the converted point
    from gnark_crypto_primitives_b200 import _lib
    pk_te = ed.rte_to_te(*PK)
    ct_te, st = eng.elgamal_encrypt(elems(pk_te), elems(ks[:6]), elems(ms[:6]), fmt=_lib.COORDS_TE)
    assert not st.any()
    for i in range(6):
        c1, c2 = eg.encrypt(PK, ks[i], ms[i])
        assert ints(ct_te[i]) == list(ed.rte_to_te(*c1)) + list(ed.rte_to_te(*c2))


def test_wide_table_24_bits(own_engine):
    eng = own_engine
    rng = random.Random(24)
    eng.set_fixed_base_window(24)
    assert eng.fixed_base_window(0) == 24
    ks = boundary_scalars(24) + [rng.randrange(R) for _ in range(30)]
    ms = [rng.randrange(1 << 20) for _ in ks]
    out, st = eng.elgamal_fixed_base_mul(elems(ks))
    assert not st.any()
    assert [tuple(ints(p)) for p in out] == [ed.scalar_mul(ed.G, k) for k in ks]
    ct, st = eng.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
    assert not st.any() and eng.fixed_base_window(1) == 24
    for i in (0, 3, 4, 5, 11, 20, len(ks) - 1):
        assert ints(ct[i]) == eg.serialize(eg.encrypt(PK, ks[i], ms[i])), i
    # an off-curve key at this width: status 4, no result (encrypt.go:49)
    ct, st = eng.elgamal_encrypt(elems((1, 2)), elems(ks[:3]), elems(ms[:3]))
    assert (st == 4).all()
    # back to automatic: the wide tables stay in place, results unchanged
    eng.set_fixed_base_window(0)
    ct2, st = eng.elgamal_encrypt(elems(PK), elems(ks[:5]), elems(ms[:5]))
    assert not st.any()
    for i in range(5):
        assert ints(ct2[i]) == eg.serialize(eg.encrypt(PK, ks[i], ms[i]))


def test_tables_widen_once_a_base_has_served_enough(monkeypatch):
    """The automatic policy with its thresholds lowered: G and the key start at 20 bits, widen (here to 22) after 4 096
    multiplications, a new key starts narrow again, and every result on the way equals the oracle's."""
    import gnark_crypto_primitives_b200 as g

    monkeypatch.setenv("GCP_B200_FB_WIDEN_AT", "4096")
    monkeypatch.setenv("GCP_B200_FB_WBITS", "22")
    eng = g.Engine(0)
    try:
        rng = random.Random(7)
        assert eng.fixed_base_window(0) == 20 and eng.fixed_base_window(1) == 20
        n = 3000
        ks = [rng.randrange(R) for _ in range(n)]
        ms = [rng.randrange(1 << 16) for _ in range(n)]
        first, st = eng.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
        assert not st.any() and eng.fixed_base_window(0) == 20 and eng.fixed_base_window(1) == 20
        second, st = eng.elgamal_encrypt(elems(PK), elems(ks), elems(ms))
        assert not st.any() and eng.fixed_base_window(0) == 22 and eng.fixed_base_window(1) == 22
        assert (first == second).all()
        for i in (0, 1, 2, n - 1):
            assert ints(second[i]) == eg.serialize(eg.encrypt(PK, ks[i], ms[i]))
        other = ed.scalar_mul(ed.G, 77)
        third, st = eng.elgamal_encrypt(elems(other), elems(ks[:50]), elems(ms[:50]))
        assert not st.any() and eng.fixed_base_window(0) == 22 and eng.fixed_base_window(1) == 20
        for i in (0, 49):
            assert ints(third[i]) == eg.serialize(eg.encrypt(other, ks[i], ms[i]))
        # the fused tally under the first key again: its table is rebuilt narrow (the key was replaced), then widens
        nb = 2400
        k2 = elems(ks[:nb * 1]).reshape(nb, 1, 32)
        m2 = elems(ms[:nb * 1]).reshape(nb, 1, 32)
        t1, st = eng.elgamal_encrypt_tally(elems(PK), k2, m2)
        assert not st.any() and eng.fixed_base_window(1) == 20
        t2, st = eng.elgamal_encrypt_tally(elems(PK), k2, m2)
        assert not st.any() and eng.fixed_base_window(1) == 22
        assert (t1 == t2).all()
        assert ints(t1[0]) == eg.serialize(eg.encrypt(PK, sum(ks[:nb]) % ed.ORDER, sum(ms[:nb]) % ed.ORDER))
    finally:
        eng.close()


def test_bad_window_is_rejected(own_engine):
    import gnark_crypto_primitives_b200 as g

    for bits in (-1, 1, 7, 27, 64):
        with pytest.raises(g.EngineError):
            own_engine.set_fixed_base_window(bits)
    assert own_engine.fixed_base_window(0) == 20
