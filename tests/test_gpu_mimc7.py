"""GPU parity: MiMC7 through the C ABI vs the oracle and the public iden3 vectors."""
import random

import numpy as np
import pytest

from oracle import mimc7 as omimc
from oracle.field import R
from tests.util import elems, ints

pytestmark = pytest.mark.gpu


def test_public_iden3_vectors(engine):
    out, st = engine.mimc7_hash(elems([12]).reshape(1, 1, 32))          # input of mimc_test.go:37
    assert int(st[0]) == 0
    assert ints(out)[0] == 16051049095595290701999129793867590386356047218708919933694064829788708231421
    out, st = engine.mimc7_hash(elems([12, 45, 78, 41]).reshape(1, 4, 32))
    assert ints(out)[0] == 18226366069841799622585958305961373004333097209608110160936134895615261821931


@pytest.mark.parametrize("length", [1, 2, 5, 62])
def test_matches_oracle(engine, length):
    rng = random.Random(length)
    n = 12
    rows = [[rng.randrange(R) for _ in range(length)] for _ in range(n)]
    rows[0] = [0] * length
    rows[1] = [R - 1] * length
    if length == 62:
        rows[2] = [rows[2][0]] * 62                                     # mimc_test.go:81-113: 62 x the same value
    out, st = engine.mimc7_hash(elems([x for r in rows for x in r]).reshape(n, length, 32))
    assert not st.any()
    assert ints(out) == [omimc.hash(r) for r in rows]


def test_limits_and_formats(engine):
    import gnark_crypto_primitives_b200 as g

    with pytest.raises(g.EngineError):
        engine.mimc7_hash(np.zeros((1, 63, 32), np.uint8))              # > 62 inputs: the reference drops the Write
    out, st = engine.mimc7_hash(elems([5, R]).reshape(1, 2, 32))
    assert int(st[0]) == 1
    M = 1 << 256
    rows = [[3, 4, 5]]
    out, st = engine.mimc7_hash(elems([x * M % R for x in rows[0]]).reshape(1, 3, 32), fmt=g.FMT_MONTGOMERY)
    assert ints(out)[0] * pow(M, -1, R) % R == omimc.hash(rows[0])
