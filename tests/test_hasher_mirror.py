"""The batched mirrors of the reference's stateful hashers (hash/hash.go:9-18): bookkeeping on the CPU, Sum on the GPU."""
import numpy as np
import pytest

from tests.util import elems, ints


class _NoEngine:
    def poseidon_hash(self, rows, fmt=0):
        raise AssertionError("Sum must not be called in the CPU test")


def test_write_drops_a_whole_call_past_the_limit():
    """poseidon.go:103-108: Write beyond 16 inputs in total is discarded as a whole, silently."""
    from gnark_crypto_primitives_b200 import hasher

    h = hasher.Poseidon(_NoEngine())
    assert not h.WriteSucceeded()
    col = elems([1, 2, 3])
    h.Write(*[col] * 10)
    assert h.WriteSucceeded() and len(h._cols) == 10
    h.Write(*[col] * 7)          # 17 > 16: dropped entirely
    assert len(h._cols) == 10
    h.Write(*[col] * 6)
    assert len(h._cols) == 16
    h.Reset()
    assert not h.WriteSucceeded()
    m = hasher.MiMC7(_NoEngine())
    m.Write(*[col] * 63)         # mimc.go:33-38
    assert not m.WriteSucceeded()
    with pytest.raises(ValueError):
        h.Write(col, elems([1, 2]))


@pytest.mark.gpu
def test_sum_and_sum_is_equal(engine):
    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import hasher

    h = hasher.Poseidon(engine)
    h.Write(elems([1, 1]), elems([2, 5]))
    dig, st = h.Sum()
    assert not st.any()
    assert ints(dig)[0] == 7853200120776062878684798364095072458815029376092732009249414926327459813530   # Poseidon(1, 2)
    flags, _ = h.SumIsEqual(np.stack([dig[0], dig[0]]))
    assert [int(f) for f in flags] == [1, 0]
    h.AssertSumIsEqual(dig)
    with pytest.raises(AssertionError):
        h.AssertSumIsEqual(np.stack([dig[0], dig[0]]))
    h.Reset()
    with pytest.raises(g.EngineError) as e:
        h.Sum()                  # nothing written: the reference's Hash error
    assert "bad inputs provided" in str(e.value)
    m = hasher.MiMC7(engine)
    m.Write(elems([12]))
    dig, st = m.Sum()
    assert ints(dig)[0] == 16051049095595290701999129793867590386356047218708919933694064829788708231421


def test_function_form_hashers_shape_their_rows_like_the_reference():
    """utils/hashers.go:25-37: PoseidonHasher / Poseidon2Hasher take the inputs as separate arguments; the mirrors stack
    one column per argument into the (n, arity, 32) rows of the engine call and keep the reference's arity errors."""
    from gnark_crypto_primitives_b200 import hasher

    seen = {}

    class Recorder:
        def poseidon_hash(self, rows, fmt=0):
            seen["poseidon"] = (rows.copy(), fmt)
            return "digests", "status"

        def poseidon2_hash(self, rows, fmt=0):
            seen["poseidon2"] = (rows.copy(), fmt)
            return "digests2", "status2"

    a, b, c = elems([1, 4]), elems([2, 5]), elems([3, 6])
    assert hasher.PoseidonHasher(Recorder(), a, b, c) == ("digests", "status")
    rows, fmt = seen["poseidon"]
    assert rows.shape == (2, 3, 32) and fmt == 0 and ints(rows[0]) == [1, 2, 3] and ints(rows[1]) == [4, 5, 6]
    assert hasher.Poseidon2Hasher(Recorder(), a, b, fmt=1) == ("digests2", "status2")
    rows, fmt = seen["poseidon2"]
    assert rows.shape == (2, 2, 32) and fmt == 1 and ints(rows[1]) == [4, 5]
    hasher.Poseidon2Hasher(Recorder(), a, b, c)
    assert seen["poseidon2"][0].shape == (2, 3, 32)
    for bad in ((a,), (a, b, c, a)):
        with pytest.raises(ValueError, match="need 2 or 3 limbs"):
            hasher.Poseidon2Hasher(Recorder(), *bad)
    with pytest.raises(ValueError):
        hasher.PoseidonHasher(Recorder(), a, elems([1]))

