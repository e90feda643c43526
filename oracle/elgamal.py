"""Oracle: ElGamal over the a = -1 BN254 twisted-Edwards curve.

Follows /root/reference/elgamal:
  mul.go:26-72   initFixedBaseTable      table[i][j] = [j * 2^(4i)] G, 63x16 + 1x4
  mul.go:76-166  FixedBaseScalarMulBN254 64 windows, window 0 initialises, zero nibbles skipped
  encrypt.go:42-64 Encrypt, :72-94 EncryptedZero
  ciphertext.go:16-19 NewCiphertext, :24-32 Add, :37-46 Neg, :50-67 AssertDecrypt,
  :98-105 Serialize, :124-168 DecryptionProof.Verify, :173-184 hashPointsToScalar
"""
from functools import lru_cache

from . import edwards as ed
from . import poseidon
from .field import R

N_WINDOWS = 64  # mul.go:27


@lru_cache(maxsize=None)
def fixed_base_table():
    table = []
    for i in range(N_WINDOWS):
        entries = 4 if i == N_WINDOWS - 1 else 16
        row = [ed.IDENTITY]
        for j in range(1, entries):
            row.append(ed.scalar_mul(ed.G, j << (4 * i)))
        table.append(row)
    return table


def fixed_base_scalar_mul(scalar):
    """mul.go:76-166.  scalar is an Fr element used as an integer; bits.ToBinary(254) asserts < 2^254."""
    scalar = int(scalar)
    if not 0 <= scalar < R:
        raise ed.CurveError("non-canonical scalar")
    table = fixed_base_table()
    res = None
    for i in range(N_WINDOWS):
        if i < N_WINDOWS - 1:
            nib = (scalar >> (4 * i)) & 0xF
        else:
            nib = (scalar >> (4 * i)) & 0x3
        contrib = table[i][nib]
        if i == 0:
            res = contrib                       # mul.go:129-131
        elif nib != 0:
            res = ed.add(res, contrib)          # mul.go:135-137 / :159-161
    return res


def encrypt(pub_key, k, m):
    """encrypt.go:42-64 -> ((C1x, C1y), (C2x, C2y))."""
    if not ed.is_on_curve(pub_key):
        raise ed.CurveError("public key not on curve")   # encrypt.go:49
    c1 = fixed_base_scalar_mul(k)
    s = ed.scalar_mul(pub_key, k)
    m_point = fixed_base_scalar_mul(m)
    c2 = ed.add(m_point, s)
    return (c1, c2)


def encrypted_zero(pub_key, k):
    """encrypt.go:72-94."""
    if not ed.is_on_curve(pub_key):
        raise ed.CurveError("public key not on curve")
    c1 = fixed_base_scalar_mul(k)
    s = ed.scalar_mul(pub_key, k)
    return (c1, ed.add(ed.IDENTITY, s))


def new_ciphertext():
    return (ed.IDENTITY, ed.IDENTITY)


def ct_add(x, y):
    """ciphertext.go:24-32."""
    return (ed.add(x[0], y[0]), ed.add(x[1], y[1]))


def ct_neg(x):
    """ciphertext.go:37-46."""
    return (ed.neg(x[0]), ed.neg(x[1]))


def serialize(ct):
    """ciphertext.go:98-105: C1.X, C1.Y, C2.X, C2.Y."""
    return [ct[0][0], ct[0][1], ct[1][0], ct[1][1]]


def tally(cts):
    """Left fold of Ciphertext.Add starting from NewCiphertext (the caller's loop in davinci-node)."""
    acc = new_ciphertext()
    for c in cts:
        acc = ct_add(acc, c)
    return acc


def assert_decrypt(ct, priv_key, m):
    """ciphertext.go:50-67; returns True when the assertions hold."""
    if not (ed.is_on_curve(ct[0]) and ed.is_on_curve(ct[1])):
        return False
    s = ed.scalar_mul(ct[0], priv_key)
    mp = ed.add(ct[1], ed.neg(s))
    return mp == fixed_base_scalar_mul(m)


def verify_decryption_proof(pub_key, ct, msg, a1, a2, z):
    """ciphertext.go:124-168 with hFn = poseidon.MultiHash; True iff every assertion holds."""
    for p in (pub_key, ct[0], ct[1], a1, a2):
        if not ed.is_on_curve(p):
            return False
    m_pt = fixed_base_scalar_mul(msg)
    d_pt = ed.add(ct[1], ed.neg(m_pt))
    coords = []
    for p in (pub_key, pub_key, ct[0], d_pt, a1, a2):      # :141, :173-184
        coords += [p[0], p[1]]
    e = poseidon.multihash(coords)
    zg = fixed_base_scalar_mul(z)
    if ed.add(a1, ed.scalar_mul(pub_key, e)) != zg:
        return False
    return ed.add(a2, ed.scalar_mul(d_pt, e)) == ed.scalar_mul(ct[0], z)
