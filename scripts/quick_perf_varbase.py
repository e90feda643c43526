"""Quick throughput probe (not the bench) of the kernels built on VARIABLE-base scalar multiplication: Encrypt with a
key per item, AssertDecrypt, DecryptionProof.Verify, EdDSA-Poseidon.  A few honest items from the oracle, tiled."""
import random, sys, time
import numpy as np
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from oracle import edwards as ed, elgamal as eg, eddsa as oeddsa
from tests.util import elems
from tests.test_gpu_proofs import make_proof

rng = random.Random(1)
eng = g.Engine(0)
N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
base = 8

def tile(a, reps):
    return np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))

def timed(name, fn, n):
    fn()
    t0 = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t0
    print(f"{name}: n={n} {dt*1e3:.1f} ms  {n/dt/1e6:.3f} M/s  ok={out}", flush=True)

# per-item-key encrypt
pks, ks, ms = [], [], []
for _ in range(base):
    pks.append(ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER))); ks.append(rng.randrange(1 << 253)); ms.append(rng.randrange(1 << 16))
pk_a = tile(elems([c for p in pks for c in p]).reshape(base, 2, 32), N // base)
k_a, m_a = tile(elems(ks), N // base), tile(elems(ms), N // base)
want = eg.serialize(eg.encrypt(pks[0], ks[0], ms[0]))
def enc():
    ct, st = eng.elgamal_encrypt(pk_a, k_a, m_a)
    return (not st.any()) and bytes(ct[0].reshape(-1)) == b"".join(int(v).to_bytes(32, "little") for v in want)
timed("encrypt per-item key", enc, N)

# assert decrypt
items = []
for _ in range(base):
    d = rng.randrange(1, ed.ORDER); msg = rng.randrange(1000)
    ct = eg.encrypt(ed.scalar_mul(ed.G, d), rng.randrange(ed.ORDER), msg)
    items.append((ct, d, msg))
ct_a = tile(elems([x for it in items for x in eg.serialize(it[0])]).reshape(base, 4, 32), N // base)
d_a, msg_a = tile(elems(it[1] for it in items), N // base), tile(elems(it[2] for it in items), N // base)
def adec():
    f, st = eng.elgamal_assert_decrypt(ct_a, d_a, msg_a)
    return bool(f.all()) and not st.any()
timed("assert_decrypt", adec, N)

# decryption proof
items = []
for _ in range(base):
    msg = rng.randrange(1000)
    pk, ct, a1, a2, z = make_proof(rng, rng.randrange(1, ed.ORDER), msg)
    items.append((pk, ct, msg, a1, a2, z))
args = [tile(elems([c for it in items for c in it[0]]).reshape(base, 2, 32), N // base),
        tile(elems([x for it in items for x in eg.serialize(it[1])]).reshape(base, 4, 32), N // base),
        tile(elems(it[2] for it in items), N // base),
        tile(elems([c for it in items for c in it[3]]).reshape(base, 2, 32), N // base),
        tile(elems([c for it in items for c in it[4]]).reshape(base, 2, 32), N // base),
        tile(elems(it[5] for it in items), N // base)]
def dproof():
    f, st = eng.elgamal_verify_decryption_proof(*args)
    return bool(f.all()) and not st.any()
timed("decryption_proof.verify", dproof, N)

# eddsa
items = []
for _ in range(base):
    msg = rng.getrandbits(248)
    a, r, s = oeddsa.sign(rng.randrange(1, ed.ORDER), rng.randrange(1, ed.ORDER), msg)
    items.append((a, r, s, msg))
eargs = [tile(elems([c for it in items for c in it[0]]).reshape(base, 2, 32), N // base),
         tile(elems([c for it in items for c in it[1]]).reshape(base, 2, 32), N // base),
         tile(elems(it[2] for it in items), N // base), tile(elems(it[3] for it in items), N // base)]
def edd():
    f, st = eng.eddsa_verify(*eargs)
    return bool(f.all()) and not st.any()
timed("eddsa_verify", edd, N)
