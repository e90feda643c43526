"""Oracle (test infrastructure): the width-2 Poseidon2 Merkle-Damgard hasher of the reference.

PARITY UNPINNED.  The reference's own code for this hasher is only the wrapper:
  /root/reference/hash/native/bn254/poseidon2/native.go:27      perm2 = poseidon2.NewPermutation(2, 6, 50)
  /root/reference/hash/native/bn254/poseidon2/native.go:30-63   HashPoseidon2.Hash (mod-r reduction, min/max order, chain)
  /root/reference/hash/native/bn254/poseidon2/gnark.go:18-54    HashPoseidon2Gnark (same values in-circuit)
  /root/reference/hash/native/bn254/poseidon2/hints.go:10-19    MinMaxHint
The permutation and its round keys live in an un-vendored dependency (gnark-crypto
v0.19.3-0.20251115174214-022ec58e8c19, ecc/bn254/fr/poseidon2) and the reference tree holds no vector for it
(poseidon2_test.go compares the gadget with the native hasher on random inputs).  What follows restates that
package's published algorithm (Poseidon2, eprint 2023/323, t = 2: external matrix circ(2, 1), internal matrix
[[2, 1], [1, 3]], S-box x^5) and its round-key derivation (a legacy-Keccak-256 chain seeded with the parameter
string).  The wrapper above it (`hash`) is pinned to the reference lines cited.  Because the keys cannot be checked
here, the engine takes them as DATA: `gcp_poseidon2_set_round_keys` lets a Go host install
`poseidon2.NewParameters(2, 6, 50).RoundKeys` verbatim, after which only the (published) round structure is assumed.
"""
from functools import lru_cache

from .field import R
from .keccak import keccak256

WIDTH = 2
RF = 6    # native.go:27
RP = 50   # native.go:27
DEGREE = 5
N_KEYS = (RF // 2) * WIDTH + RP + (RF // 2) * WIDTH   # 62 elements, flattened in round order


def seed_string(width=WIDTH, rf=RF, rp=RP, d=DEGREE) -> str:
    """gnark-crypto Parameters.String()."""
    return f"Poseidon2-BN254[t={width},rF={rf},rP={rp},d={d}]"


@lru_cache(maxsize=None)
def round_keys(width=WIDTH, rf=RF, rp=RP):
    """gnark-crypto Parameters.initRC: rnd = keccak(seed); every key = keccak(previous digest) read big-endian mod r.
    Full rounds carry `width` keys, partial rounds one.  Returns a list of per-round lists."""
    rnd = keccak256(seed_string(width, rf, rp).encode())
    keys = []
    for i in range(rf + rp):
        n = width if (i < rf // 2 or i >= rf // 2 + rp) else 1
        row = []
        for _ in range(n):
            rnd = keccak256(rnd)
            row.append(int.from_bytes(rnd, "big") % R)
        keys.append(row)
    return keys


def flat_round_keys():
    return [k for row in round_keys() for k in row]


def unflatten(flat, width=WIDTH, rf=RF, rp=RP):
    assert len(flat) == rf * width + rp
    out, p = [], 0
    for i in range(rf + rp):
        n = width if (i < rf // 2 or i >= rf // 2 + rp) else 1
        out.append(list(flat[p:p + n]))
        p += n
    return out


def _sbox(x):
    x2 = x * x % R
    return x2 * x2 % R * x % R


def _external(s):
    t = (s[0] + s[1]) % R                      # circ(2, 1)
    return [(t + s[0]) % R, (t + s[1]) % R]


def _internal(s):
    t = (s[0] + s[1]) % R                      # [[2, 1], [1, 3]]
    return [(s[0] + t) % R, (2 * s[1] + t) % R]


def permutation(state, keys=None):
    """poseidon2.Permutation.Permutation for width 2: external matrix first, rf/2 full rounds, rp partial rounds,
    rf/2 full rounds; a round is matmul(sbox(add_round_key(state)))."""
    keys = round_keys() if keys is None else keys
    s = _external([x % R for x in state])
    half = RF // 2
    for i in range(RF + RP):
        if i < half or i >= half + RP:
            s = [_sbox((s[0] + keys[i][0]) % R), _sbox((s[1] + keys[i][1]) % R)]
            s = _external(s)
        else:
            s = [_sbox((s[0] + keys[i][0]) % R), s[1]]
            s = _internal(s)
    return s


def hash(limbs, keys=None):
    """HashPoseidon2.Hash (native.go:30-63) on integers: 2 or 3 limbs, each reduced mod r (:37-39), a 2-limb call is
    ordered (min, max) (:42-44; gnark.go:24-36), then CV <- Permutation([CV, m])[1] + m over the limbs (:47-61)."""
    if len(limbs) not in (2, 3):
        raise ValueError(f"poseidon2: need 2 or 3 limbs, got {len(limbs)}")   # native.go:31-33
    safe = [int(x) % R for x in limbs]
    if len(safe) == 2 and safe[0] > safe[1]:
        safe = [safe[1], safe[0]]
    cv = 0
    for m in safe:
        st = permutation([cv, m], keys)
        cv = (st[1] + m) % R
    return cv
