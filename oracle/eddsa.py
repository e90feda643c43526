"""Oracle: iden3/circomlib EdDSA-Poseidon verification as the gadget evaluates it.

Follows /root/reference/ecc/bn254/eddsa/verifier.go:55-88 (IsValid), :40-49 (PointToRTE), constants.go:11-18
(rteB8 = FromTEtoRTE(babyjub.B8), which equals gnark's base point G), types.go:37-49 (S reduced mod the order).
  h = Poseidon(R.x, R.y, A.x, A.y, msg) on the ORIGINAL (TE, a = 168700) coordinates
  A' = RTE(A), R' = RTE(R), both asserted on the a = -1 curve
  flag = ([S] rteB8 == 8 * [h] A' + R')
`sign` is a test-side signer producing signatures that satisfy the circomlib equation S*B8 = R8 + 8*h*A.
"""
from . import edwards as ed
from . import poseidon


def is_valid(a_te, r_te, s, msg):
    """-> (flag, on_curve_ok).  on_curve_ok False = an AssertIsOnCurve of the gadget fails."""
    h = poseidon.hash([r_te[0], r_te[1], a_te[0], a_te[1], msg])          # verifier.go:58-59
    a = ed.te_to_rte(*a_te)                                               # verifier.go:66-67
    r = ed.te_to_rte(*r_te)
    if not (ed.is_on_curve(a) and ed.is_on_curve(r)):
        return 0, False
    left = ed.scalar_mul(ed.G, s)                                         # verifier.go:69
    r1 = ed.scalar_mul(a, h)                                              # verifier.go:71
    for _ in range(3):
        r1 = ed.double(r1)                                                # verifier.go:72-74
    right = ed.add(r1, r)                                                 # verifier.go:76
    return (1 if left == right else 0), True


def sign(secret, nonce, msg):
    """-> (A_te, R_te, S) with A = [secret]B8, R = [nonce]B8, S = nonce + 8*h*secret mod l."""
    a = ed.scalar_mul(ed.G, secret)
    r = ed.scalar_mul(ed.G, nonce)
    a_te, r_te = ed.rte_to_te(*a), ed.rte_to_te(*r)
    h = poseidon.hash([r_te[0], r_te[1], a_te[0], a_te[1], msg])
    s = (nonce + 8 * h * secret) % ed.ORDER
    return a_te, r_te, s
