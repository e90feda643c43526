"""Oracle: gnark / gnark-crypto BN254 twisted-Edwards curve (a = -1 "reduced" BabyJubJub).

The point arithmetic lives in un-vendored dependencies
(github.com/consensys/gnark v0.14.1-0.20251203003358-cce547909fed,
std/algebra/native/twistededwards; github.com/consensys/gnark-crypto
v0.19.3-0.20251115174214-022ec58e8c19, ecc/bn254/twistededwards), so this file
restates their published algorithm (complete affine Edwards addition) and is
pinned by the reference's call sites and static vectors:
  elgamal/ciphertext_test.go:286-345 (decryption-proof KAT),
  ecc/format/twistededwards.go:17 (scalingFactor maps iden3 B8 to this G),
  [ORDER]G == identity.
Call sites restated: elgamal/ciphertext.go:24-46, elgamal/encrypt.go:42-94.
"""
from .field import R, inv

A = R - 1
D = 12181644023421730124874158521699555681764249180949974110617291017600649128846
GX = 9671717474070082183213120605117400219616337014328744928644933853176787189663
GY = 16950150798460657717958625567821834550301663161624707787222815936182638968203
G = (GX, GY)
ORDER = 2736030358979909402780800718157159386076813972158567259200215660948447373041
COFACTOR = 8
IDENTITY = (0, 1)

# ecc/format/twistededwards.go:17
SCALING_FACTOR = 6360561867910373094066688120553762416144456282423235903351243436111059670888


class CurveError(ValueError):
    """An assertion of the gadget would have failed (solver error in the reference)."""


def is_on_curve(p):
    """curve.AssertIsOnCurve: a*x^2 + y^2 == 1 + d*x^2*y^2."""
    x, y = p
    x2, y2 = x * x % R, y * y % R
    return (A * x2 + y2) % R == (1 + D * x2 % R * y2) % R


def add(p, q):
    """curve.Add: complete affine addition."""
    x1, y1 = p
    x2, y2 = q
    x1x2 = x1 * x2 % R
    y1y2 = y1 * y2 % R
    k = D * x1x2 % R * y1y2 % R
    dx = (1 + k) % R
    dy = (1 - k) % R
    if dx == 0 or dy == 0:
        raise CurveError("edwards addition: zero denominator")
    x3 = (x1 * y2 + y1 * x2) % R * inv(dx) % R
    y3 = (y1y2 - A * x1x2) % R * inv(dy) % R
    return (x3, y3)


def neg(p):
    return ((-p[0]) % R, p[1])


def double(p):
    return add(p, p)


def scalar_mul_affine(p, s):
    """[s]P by double-and-add on the affine law above (slow; the definition)."""
    acc = IDENTITY
    base = p
    s = int(s)
    while s:
        if s & 1:
            acc = add(acc, base)
        base = add(base, base)
        s >>= 1
    return acc


def _padd(p, q):
    # projective (X:Y:Z) form of the same complete law; no inversions
    x1, y1, z1 = p
    x2, y2, z2 = q
    a_ = z1 * z2 % R
    b_ = a_ * a_ % R
    c_ = x1 * x2 % R
    d_ = y1 * y2 % R
    e_ = D * c_ % R * d_ % R
    f_ = (b_ - e_) % R
    g_ = (b_ + e_) % R
    x3 = a_ * f_ % R * (((x1 + y1) * (x2 + y2) - c_ - d_) % R) % R
    y3 = a_ * g_ % R * ((d_ - A * c_) % R) % R
    return (x3, y3, f_ * g_ % R)


def scalar_mul(p, s):
    """curve.ScalarMul(P, s): [s]P with s an integer in [0, r) -- NOT reduced mod ORDER by the caller.

    Same group element as scalar_mul_affine (the law is exact); computed projectively for speed.
    """
    acc = (0, 1, 1)
    base = (p[0] % R, p[1] % R, 1)
    s = int(s)
    while s:
        if s & 1:
            acc = _padd(acc, base)
        s >>= 1
        if s:
            base = _padd(base, base)
    if acc[2] == 0:
        raise CurveError("scalar_mul: point at infinity (input off-curve)")
    zi = inv(acc[2])
    return (acc[0] * zi % R, acc[1] * zi % R)


def te_to_rte(x, y):
    """ecc/format/twistededwards.go:42-48: x_RTE = x_TE * (-f)."""
    return (x * ((-SCALING_FACTOR) % R) % R, y)


def rte_to_te(x, y):
    """ecc/format/twistededwards.go:29-37: x_TE = x_RTE / (-f)."""
    return (x * inv((-SCALING_FACTOR) % R) % R, y)
