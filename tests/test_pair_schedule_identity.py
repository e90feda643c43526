"""CPU: the algebra behind the kernels' partial rounds in pairs (csrc/poseidon.cuh) in plain integers, checked against the
oracle's literal permutation (oracle/poseidon.py, poseidon.go:152-166) for every width t = 2..17.

A partial round is  x = sigma(s0) + c;  s0' = <S[0..t), (x, s1..s_{t-1})>;  s_k += x * S[t+k-1].  The s_k written by round A
are read only by round B's dot row and rank-1 update, and both are linear in them, so
  round A:  x0 = sigma(s0) + cA;  s0 = <S_A, (x0, s)>                              (no rank-1 update)
  round B:  x1 = sigma(s0) + cB;  s0 = <S_B, (x1, s)> + d * x0,   d = sum_k S_B[k] * S_A[t+k-1]
            s_k += S_A[t+k-1] * x0 + S_B[t+k-1] * x1
gives the same state after two rounds.  The test restates that schedule the two ways the kernels run it - the generic
kernel does one ordinary round first when RP is odd; the t = 3 kernel runs the first round as a round B with x0 = 0 - and
the block form of DESIGN 8 (B rounds per block) for good measure."""
import random

import pytest

from oracle import poseidon as opos
from oracle.field import R, poseidon_tables

HALF = opos.N_ROUNDS_F // 2


def _sigma(x):
    return pow(x, 5, R)


def _permute(inputs, partial_rounds):
    """poseidon.go:116-183 with the partial rounds (:152-166) replaced by `partial_rounds(state, tab) -> state`."""
    t = len(inputs) + 1
    tab = poseidon_tables()[t]
    rp, c, m, p = tab["RP"], tab["C"], tab["M"], tab["P"]

    def mix(st, mat):
        return [sum(mat[j * t + i] * st[j] for j in range(t)) % R for i in range(t)]

    state = [(x + c[i]) % R for i, x in enumerate([0] + list(inputs))]
    for r in range(HALF - 1):
        state = mix([(_sigma(x) + c[(r + 1) * t + i]) % R for i, x in enumerate(state)], m)
    state = mix([(_sigma(x) + c[HALF * t + i]) % R for i, x in enumerate(state)], p)
    state = partial_rounds(state, tab)
    for r in range(HALF - 1):
        off = (HALF + 1) * t + rp + r * t
        state = mix([(_sigma(x) + c[off + i]) % R for i, x in enumerate(state)], m)
    state = [_sigma(x) for x in state]
    return sum(m[j * t] * state[j] for j in range(t)) % R


def _ordinary_round(state, tab, r):
    t, c, s = len(state), tab["C"], tab["S"]
    x = (_sigma(state[0]) + c[(HALF + 1) * t + r]) % R
    row = s[(2 * t - 1) * r:(2 * t - 1) * (r + 1)]
    new0 = (row[0] * x + sum(row[j] * state[j] for j in range(1, t))) % R
    return [new0] + [(state[k] + x * row[t + k - 1]) % R for k in range(1, t)]


def _pair_constant(tab, t, ra):
    """d for the pair (ra, ra + 1): what csrc/kernels.cu: poseidon_pair_constants_kernel / pos3_pair_kernel compute."""
    s = tab["S"]
    row_a, row_b = s[(2 * t - 1) * ra:], s[(2 * t - 1) * (ra + 1):]
    return sum(row_b[k] * row_a[t + k - 1] for k in range(1, t)) % R


def _round_b(state, tab, rb, x0, d):
    """round B of a pair; with x0 = 0 it is one ordinary round (the t = 3 kernel's first partial round)."""
    t, c, s = len(state), tab["C"], tab["S"]
    row_a = s[(2 * t - 1) * (rb - 1):(2 * t - 1) * rb] if rb > 0 else [0] * (2 * t - 1)
    row_b = s[(2 * t - 1) * rb:(2 * t - 1) * (rb + 1)]
    x1 = (_sigma(state[0]) + c[(HALF + 1) * t + rb]) % R
    new0 = (row_b[0] * x1 + sum(row_b[j] * state[j] for j in range(1, t)) + d * x0) % R
    return [new0] + [(state[k] + row_a[t + k - 1] * x0 + row_b[t + k - 1] * x1) % R for k in range(1, t)]


def _round_a(state, tab, ra):
    t, c, s = len(state), tab["C"], tab["S"]
    x0 = (_sigma(state[0]) + c[(HALF + 1) * t + ra]) % R
    row = s[(2 * t - 1) * ra:(2 * t - 1) * (ra + 1)]
    return [(row[0] * x0 + sum(row[j] * state[j] for j in range(1, t))) % R] + state[1:], x0


def one_round_at_a_time(state, tab):
    for r in range(tab["RP"]):
        state = _ordinary_round(state, tab, r)
    return state


def pairs_generic(state, tab):
    """poseidon_permute_generic: an odd RP starts with one ordinary round, then pairs."""
    t, rp = len(state), tab["RP"]
    r = 0
    if rp & 1:
        state = _ordinary_round(state, tab, 0)
        r = 1
    while r < rp:
        state, x0 = _round_a(state, tab, r)
        state = _round_b(state, tab, r + 1, x0, _pair_constant(tab, t, r))
        r += 2
    return state


def pairs_first_round_as_b(state, tab):
    """poseidon_permute_const<3, true>: the first partial round is a round B with x0 = 0 (any d), then A at odd, B at even."""
    t, rp = len(state), tab["RP"]
    assert rp & 1
    state = _round_b(state, tab, 0, 0, 12345)
    for q in range(1, rp, 2):
        state, x0 = _round_a(state, tab, q)
        state = _round_b(state, tab, q + 1, x0, _pair_constant(tab, t, q))
    return state


def blocks(size):
    """DESIGN 8 'next': B rounds per block - round i of a block adds sum_{i' < i} D[i][i'] x_i' to its dot row, the
    rank-1 updates of the whole block are applied at its end."""
    def run(state, tab):
        t, rp, c, s = len(state), tab["RP"], tab["C"], tab["S"]
        r = 0
        while r < rp:
            b = min(size, rp - r)
            rows = [s[(2 * t - 1) * (r + i):(2 * t - 1) * (r + i + 1)] for i in range(b)]
            xs = []
            for i in range(b):
                x = (_sigma(state[0]) + c[(HALF + 1) * t + r + i]) % R
                hist = sum(sum(rows[i][k] * rows[j][t + k - 1] for k in range(1, t)) % R * xs[j] for j in range(i))
                state = [(rows[i][0] * x + sum(rows[i][j] * state[j] for j in range(1, t)) + hist) % R] + state[1:]
                xs.append(x)
            state = [state[0]] + [(state[k] + sum(rows[i][t + k - 1] * xs[i] for i in range(b))) % R for k in range(1, t)]
            r += b
        return state
    return run


@pytest.mark.parametrize("t", range(2, 18))
def test_pair_schedule_equals_the_literal_rounds(t):
    rng = random.Random(0xB200 + t)
    rows = [[rng.randrange(R) for _ in range(t - 1)] for _ in range(3)] + [[0] * (t - 1), [R - 1] * (t - 1)]
    for inputs in rows:
        want = opos.hash(inputs)
        assert _permute(inputs, one_round_at_a_time) == want   # the harness itself
        assert _permute(inputs, pairs_generic) == want
        if poseidon_tables()[t]["RP"] & 1:
            assert _permute(inputs, pairs_first_round_as_b) == want
        for size in (3, 4):
            assert _permute(inputs, blocks(size)) == want


def test_multiply_counts_of_the_schedules():
    """units of 64 wide multiplies per two partial rounds: 4t - 2 products + 2t reductions one round at a time,
    4t - 1 products + t + 1 reductions in pairs: t - 2 saved (bench.py: wide_per_hash)."""
    for t in range(2, 18):
        one = 2 * ((2 * t - 1) + t)
        pair = (t + (t + 1) + 2 * (t - 1)) + (1 + 1 + (t - 1))
        assert one - pair == t - 2
    import bench
    assert bench.wide_per_hash(3, 57) == 61384 - (28 * 64 - (11 * 64 - 8 * 64))  # 28 pairs save a unit each; the round-B-shaped first round costs 3 more
