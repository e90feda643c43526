"""GPU: gcp_group_* (several GPUs behind one handle, the single-process form for the Go host).

A group of ONE device runs everywhere (the driver's 1-GPU box); the group of TWO devices needs a 2-GPU box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_group.py -m gpu`), where the tallies go through ncclAllGather and
must equal the single-device results bit for bit."""
import random

import numpy as np
import pytest

import gnark_crypto_primitives_b200 as g
from oracle import edwards as ed
from oracle import elgamal as eg
from oracle import poseidon as opos
from oracle import smt as osmt
from oracle.field import R
from tests.util import census_proof, elems, ints

pytestmark = pytest.mark.gpu


def n_gpus():
    from gnark_crypto_primitives_b200 import _lib

    return _lib.load().gcp_device_count()


def workload(rng, n_proofs=257, n_ballots=131, n_fields=3):
    n_levels = 64
    proofs = [census_proof(rng, n_levels, lo=3, hi=20) for _ in range(n_proofs)]
    for i in range(0, n_proofs, 9):
        proofs[i] = (proofs[i][0] ^ 1,) + proofs[i][1:]
    pk = ed.scalar_mul(ed.G, 0xB200)
    ks = [[rng.randrange(R) for _ in range(n_fields)] for _ in range(n_ballots)]
    ms = [[rng.randrange(1 << 16) for _ in range(n_fields)] for _ in range(n_ballots)]
    return n_levels, proofs, pk, ks, ms


def run_all(grp, n_levels, proofs, pk, ks, ms):
    n = len(proofs)
    res = {}
    rows = [[p[2], p[3]] for p in proofs]
    res["hash"] = grp.poseidon_hash(elems([x for r in rows for x in r]).reshape(n, 2, 32))
    sib = elems([s for p in proofs for s in p[1]]).reshape(n, n_levels, 32)
    res["smt"] = grp.smt_verify(elems(p[0] for p in proofs), sib, elems(p[2] for p in proofs), elems(p[3] for p in proofs),
                                want_roots=True)
    packed = []
    for p in proofs:
        last = max((i for i, s in enumerate(p[1]) if s), default=-1)
        packed.append(osmt.pack_siblings(p[1][:last + 1]))
    res["packed"] = grp.smt_verify_packed(elems(p[0] for p in proofs), packed, n_levels, elems(p[2] for p in proofs),
                                          elems(p[3] for p in proofs), want_roots=True)
    nb, nf = len(ks), len(ks[0])
    k = elems([x for row in ks for x in row]).reshape(nb, nf, 32)
    m = elems([x for row in ms for x in row]).reshape(nb, nf, 32)
    pk_a = elems(pk).reshape(2, 32)
    ct, st = grp.elgamal_encrypt(pk_a, k.reshape(-1, 32), m.reshape(-1, 32))
    res["encrypt"] = (ct, st)
    res["tally"] = grp.elgamal_tally(ct.reshape(nb, nf, 4, 32))
    res["encrypt_tally"] = grp.elgamal_encrypt_tally(pk_a, k, m)
    return res


def check_against_oracle(res, n_levels, proofs, pk, ks, ms):
    n = len(proofs)
    dig, st = res["hash"]
    assert not st.any()
    for i in (0, 1, n // 2, n - 1):
        assert ints(dig[i:i + 1])[0] == opos.hash([proofs[i][2], proofs[i][3]])
    for key in ("smt", "packed"):
        flags, status, roots = res[key]
        assert not status.any()
        assert [int(f) for f in flags] == [0 if i % 9 == 0 else 1 for i in range(n)]
        for i in (0, 1, n - 1):
            assert (int(flags[i]), int(status[i]), ints(roots[i:i + 1])[0]) == osmt.inclusion_verifier(*proofs[i])
    assert np.array_equal(res["smt"][2], res["packed"][2])
    # closed form of the tally: sum Encrypt(pk, k_i, m_i) = Encrypt(pk, sum k_i mod l, sum m_i mod l)
    nf = len(ks[0])
    for name in ("tally", "encrypt_tally"):
        out, st = res[name]
        assert not st.any()
        for f in range(nf):
            want = eg.encrypt(pk, sum(r[f] for r in ks) % ed.ORDER, sum(r[f] for r in ms) % ed.ORDER)
            assert ints(out[f]) == list(eg.serialize(want)), (name, f)
    ct, st = res["encrypt"]
    assert not st.any() and ints(ct[0]) == list(eg.serialize(eg.encrypt(pk, ks[0][0], ms[0][0])))


def test_group_of_one_device_matches_the_oracle():
    rng = random.Random(81)
    w = workload(rng)
    with g.Group([0]) as grp:
        assert grp.size == 1 and not grp.uses_nccl
        res = run_all(grp, *w)
        assert grp.launch_counts()[0] > 0
    check_against_oracle(res, *w)


def test_group_bad_arguments():
    with pytest.raises(g.EngineError):
        g.Group([])
    with pytest.raises(g.EngineError):
        g.Group([0, 0])
    with pytest.raises(g.EngineError):
        g.Group([4096])
    with g.Group([0]) as grp:
        with pytest.raises(g.EngineError) as e:
            grp.poseidon_hash(np.zeros((4, 17, 32), np.uint8))
        assert "bad inputs provided" in str(e.value)
        out, st = grp.poseidon_hash(np.zeros((0, 2, 32), np.uint8))
        assert out.shape == (0, 32)


def test_group_of_two_devices_is_bit_identical_to_one():
    if n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); the one-device group test covers this box")
    rng = random.Random(82)
    w = workload(rng)
    with g.Group([0]) as one:
        base = run_all(one, *w)
    import torch
    torch.cuda.set_device(0)
    with g.Group([0, 1]) as two:
        assert two.size == 2 and two.uses_nccl
        # creating contexts on other devices must not move the caller's current device (torch keeps allocating on 0)
        assert torch.cuda.current_device() == 0
        res = run_all(two, *w)
        assert torch.cuda.current_device() == 0
        assert all(c > 0 for c in two.launch_counts())
        # fewer ballots than devices: one shard is empty
        k1 = elems([5, 6, 7]).reshape(1, 3, 32)
        m1 = elems([1, 2, 3]).reshape(1, 3, 32)
        t_small = two.elgamal_encrypt_tally(elems(w[2]).reshape(2, 32), k1, m1)
    with g.Group([0]) as one:
        t_small_one = one.elgamal_encrypt_tally(elems(w[2]).reshape(2, 32), k1, m1)
    for key in base:
        for a, b in zip(base[key], res[key]):
            assert np.array_equal(a, b), key
    assert np.array_equal(t_small[0], t_small_one[0]) and not t_small[1].any()
    check_against_oracle(res, *w)


def test_group_tally_in_iden3_coordinates_and_status_merge():
    """GCP_COORDS_TE through the group forms (every visible device): key in TE, tally out in TE, equal to the conversion of
    the RTE result; a shard with a non-canonical scalar marks its field in the merged status (the statuses travel with the
    partial ciphertexts through the all-gather) and that field's ciphertext is zeroed."""
    rng = random.Random(83)
    devices = list(range(min(n_gpus(), 2)))
    nb, nf = 37, 3
    pk = ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER))
    ks = [[rng.randrange(R) for _ in range(nf)] for _ in range(nb)]
    ms = [[rng.randrange(1 << 16) for _ in range(nf)] for _ in range(nb)]
    k = elems([x for r in ks for x in r]).reshape(nb, nf, 32)
    m = elems([x for r in ms for x in r]).reshape(nb, nf, 32)
    with g.Group(devices) as grp:
        t_rte, st0 = grp.elgamal_encrypt_tally(elems(pk), k, m)
        t_te, st1 = grp.elgamal_encrypt_tally(elems(ed.rte_to_te(*pk)), k, m, fmt=g.COORDS_TE)
        assert not st0.any() and not st1.any()
        for f in range(nf):
            w = ints(t_rte[f])
            assert ints(t_te[f]) == [c for p in ((w[0], w[1]), (w[2], w[3])) for c in ed.rte_to_te(*p)]
        # the last ballot (it lands in the last device's shard) carries k >= r in field 1
        k_bad = k.copy()
        k_bad[nb - 1, 1] = elems([R + 5])[0]
        t_bad, st_bad = grp.elgamal_encrypt_tally(elems(pk), k_bad, m)
        assert [int(s) for s in st_bad] == [0, g.STATUS_NONCANONICAL, 0]
        assert not t_bad[1].any() and (t_bad[0] == t_rte[0]).all() and (t_bad[2] == t_rte[2]).all()
