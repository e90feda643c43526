"""One pass of each gadget built on variable-base scalar multiplication at a moderate size, through the host API
(for the ncu launch list and `ncu --set full -k regex:varbase_window_kernel`).  argv[1] = log2 items (default 17);
argv[2:] = which of: adec dproof eddsa enc (default all)."""
import random
import sys

import numpy as np

sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g  # noqa: E402
from oracle import eddsa as oeddsa  # noqa: E402
from oracle import edwards as ed  # noqa: E402
from oracle import elgamal as eg  # noqa: E402
from tests.test_gpu_proofs import make_proof  # noqa: E402
from tests.util import elems  # noqa: E402

rng = random.Random(1)
eng = g.Engine(0)
N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
which = set(sys.argv[2:]) or {"adec", "dproof", "eddsa", "enc"}
base = 8


def tile(a, reps):
    return np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))


if "enc" in which:
    pks = [ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)) for _ in range(base)]
    ks = [rng.randrange(1 << 253) for _ in range(base)]
    ms = [rng.randrange(1 << 16) for _ in range(base)]
    ct, st = eng.elgamal_encrypt(tile(elems([c for p in pks for c in p]).reshape(base, 2, 32), N // base),
                                 tile(elems(ks), N // base), tile(elems(ms), N // base))
    assert not st.any()
if "adec" in which:
    items = []
    for _ in range(base):
        d = rng.randrange(1, ed.ORDER)
        msg = rng.randrange(1000)
        items.append((eg.encrypt(ed.scalar_mul(ed.G, d), rng.randrange(ed.ORDER), msg), d, msg))
    f, st = eng.elgamal_assert_decrypt(
        tile(elems([x for it in items for x in eg.serialize(it[0])]).reshape(base, 4, 32), N // base),
        tile(elems(it[1] for it in items), N // base), tile(elems(it[2] for it in items), N // base))
    assert f.all() and not st.any()
if "dproof" in which:
    items = []
    for _ in range(base):
        msg = rng.randrange(1000)
        pk, ct, a1, a2, z = make_proof(rng, rng.randrange(1, ed.ORDER), msg)
        items.append((pk, ct, msg, a1, a2, z))
    f, st = eng.elgamal_verify_decryption_proof(
        tile(elems([c for it in items for c in it[0]]).reshape(base, 2, 32), N // base),
        tile(elems([x for it in items for x in eg.serialize(it[1])]).reshape(base, 4, 32), N // base),
        tile(elems(it[2] for it in items), N // base),
        tile(elems([c for it in items for c in it[3]]).reshape(base, 2, 32), N // base),
        tile(elems([c for it in items for c in it[4]]).reshape(base, 2, 32), N // base),
        tile(elems(it[5] for it in items), N // base))
    assert f.all() and not st.any()
if "eddsa" in which:
    items = []
    for _ in range(base):
        msg = rng.getrandbits(248)
        a, r, s = oeddsa.sign(rng.randrange(1, ed.ORDER), rng.randrange(1, ed.ORDER), msg)
        items.append((a, r, s, msg))
    f, st = eng.eddsa_verify(tile(elems([c for it in items for c in it[0]]).reshape(base, 2, 32), N // base),
                             tile(elems([c for it in items for c in it[1]]).reshape(base, 2, 32), N // base),
                             tile(elems(it[2] for it in items), N // base), tile(elems(it[3] for it in items), N // base))
    assert f.all() and not st.any()
print("done")
