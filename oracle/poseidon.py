"""Oracle: circomlib-compatible optimized Poseidon over BN254 Fr.

Follows /root/reference/hash/native/bn254/poseidon/poseidon.go:
  Sum :116-183, sigma :199-203, ark :205-211, mix :213-224, mixLast :226-233,
  Hash :38-45, Write :103-108, MultiHash :54-91.
Equivalent plain-field form: hash/emulated/bn254/poseidon/poseidon_test.go:112-273.
"""
from .field import R, poseidon_tables

MAX_HASH_INPUTS = 16        # poseidon.go:14
MAX_MULTIHASH_INPUTS = 4096  # poseidon.go:12
N_ROUNDS_F = 8


class PoseidonError(ValueError):
    pass


def _sigma(x):
    x2 = x * x % R
    x4 = x2 * x2 % R
    return x4 * x % R


def permute_sum(inputs):
    """poseidon.go:116-183 with inputs already validated (1..16 canonical elements)."""
    t = len(inputs) + 1
    tab = poseidon_tables()[t]
    rp, c, s, m, p = tab["RP"], tab["C"], tab["S"], tab["M"], tab["P"]
    state = [0] + [x % R for x in inputs]                      # :128-135 capacity first
    state = [(state[i] + c[i]) % R for i in range(t)]           # :136

    def mix(st, mat):                                           # :213-224, m[j][i]
        return [sum(mat[j * t + i] * st[j] for j in range(t)) % R for i in range(t)]

    half = N_ROUNDS_F // 2
    for r in range(half - 1):                                   # :138-144
        state = [_sigma(x) for x in state]
        state = [(state[i] + c[(r + 1) * t + i]) % R for i in range(t)]
        state = mix(state, m)
    state = [_sigma(x) for x in state]                          # :146-150
    state = [(state[i] + c[half * t + i]) % R for i in range(t)]
    state = mix(state, p)
    for r in range(rp):                                         # :152-166
        state[0] = (_sigma(state[0]) + c[(half + 1) * t + r]) % R
        base = (2 * t - 1) * r
        new0 = sum(s[base + j] * state[j] for j in range(t)) % R
        for k in range(1, t):
            state[k] = (state[k] + state[0] * s[base + t + k - 1]) % R
        state[0] = new0
    for r in range(half - 1):                                   # :168-174
        state = [_sigma(x) for x in state]
        off = (half + 1) * t + rp + r * t
        state = [(state[i] + c[off + i]) % R for i in range(t)]
        state = mix(state, m)
    state = [_sigma(x) for x in state]                          # :176-178
    return sum(m[j * t + 0] * state[j] for j in range(t)) % R  # :180, :226-233


def hash(inputs):
    """poseidon.go:38-45. 0 or >16 inputs -> error "bad inputs provided"."""
    n = len(inputs)
    if n == 0 or n > MAX_HASH_INPUTS:
        raise PoseidonError("bad inputs provided")
    return permute_sum(list(inputs))


def multihash(inputs):
    """poseidon.go:54-91."""
    n = len(inputs)
    if n <= MAX_HASH_INPUTS:
        return hash(inputs)
    if n > MAX_MULTIHASH_INPUTS:
        raise PoseidonError("the maximum number of inputs supported is %d" % MAX_MULTIHASH_INPUTS)
    hashed = [permute_sum(list(inputs[i:i + MAX_HASH_INPUTS])) for i in range(0, n, MAX_HASH_INPUTS)]
    if len(hashed) == 1:
        return hashed[0]
    if len(hashed) <= MAX_HASH_INPUTS:
        return permute_sum(hashed)
    return multihash(hashed)


def field_mul_count(t):
    """Fr multiplications the reference executes per hash."""
    rp = poseidon_tables()[t]["RP"]
    return 24 * t + 3 * rp + 7 * t * t + rp * (2 * t - 1) + t
