"""GPU parity: end-to-end ballot batch (census inclusion proof + ballot encryption + aggregation, config 5 shape)."""
import random

import numpy as np
import pytest

from oracle import edwards as ed
from oracle import elgamal as eg
from oracle import smt as osmt
from oracle.field import R
from tests.util import census_proof, elems, ints

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_ballot_batch_matches_composition_of_oracles(engine):
    rng = random.Random(0xBA11)
    n_levels, n, nf = 160, 96, 8
    pk = ed.scalar_mul(ed.G, 0xB200)
    items = [census_proof(rng, n_levels) for _ in range(n)]
    for i in range(0, n, 7):                       # some voters present a bad proof
        r, s, k, v = items[i]
        items[i] = ((r + 1) % R, s, k, v)
    items[5] = (items[5][0], items[5][1], items[5][2] | (1 << n_levels), items[5][3])   # assertion failure -> status
    ks = [rng.randrange(R) for _ in range(n * nf)]
    ms = [rng.randrange(1 << 16) for _ in range(n * nf)]
    flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    status = torch.empty(n, dtype=torch.uint8, device="cuda")
    tally = torch.empty((nf, 4, 32), dtype=torch.uint8, device="cuda")
    tstatus = torch.empty(nf, dtype=torch.uint8, device="cuda")
    engine.ballot_batch_dev(n_levels, n, dev(elems(it[0] for it in items)), False,
                            dev(elems([s for it in items for s in it[1]])), dev(elems(it[2] for it in items)),
                            dev(elems(it[3] for it in items)), dev(elems(pk)), dev(elems(ks)), dev(elems(ms)), nf, flags,
                            status, tally, tstatus, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    want = [osmt.inclusion_verifier(*it) for it in items]
    assert [int(f) for f in flags.cpu()] == [w[0] for w in want]
    assert [int(s) for s in status.cpu()] == [w[1] for w in want]
    assert int(status[5]) == 2 and int(flags[5]) == 0
    admitted = [i for i, w in enumerate(want) if w[0] == 1 and w[1] == 0]
    assert 0 < len(admitted) < n
    got = tally.cpu().numpy()
    assert not bool(tstatus.any())
    for f in range(nf):
        ksum = sum(ks[i * nf + f] for i in admitted) % ed.ORDER
        msum = sum(ms[i * nf + f] for i in admitted) % ed.ORDER
        assert ints(got[f]) == eg.serialize(eg.encrypt(pk, ksum, msum)), f


def _workload(rng, n_levels, n, nf):
    pk = ed.scalar_mul(ed.G, 0xB200)
    items = [census_proof(rng, n_levels, lo=2, hi=min(28, n_levels - 1)) for _ in range(n)]
    for i in range(0, n, 7):
        r, s, k, v = items[i]
        items[i] = ((r + 1) % R, s, k, v)
    ks = [rng.randrange(R) for _ in range(n * nf)]
    ms = [rng.randrange(1 << 16) for _ in range(n * nf)]
    return pk, items, ks, ms


def _check(out, pk, items, ks, ms, nf, bad=()):
    flags, status, tally, tstatus = out
    n = len(items)
    want = [osmt.inclusion_verifier(*it) for it in items]
    for i in range(n):
        if i in bad:
            assert (int(flags[i]), int(status[i])) == (0, osmt.STATUS_MALFORMED), i
        else:
            assert (int(flags[i]), int(status[i])) == want[i][:2], i
    admitted = [i for i, w in enumerate(want) if w[0] == 1 and w[1] == 0 and i not in bad]
    assert 0 < len(admitted) < n and not tstatus.any()
    for f in range(nf):
        ksum = sum(ks[i * nf + f] for i in admitted) % ed.ORDER
        msum = sum(ms[i * nf + f] for i in admitted) % ed.ORDER
        assert ints(tally[f]) == eg.serialize(eg.encrypt(pk, ksum, msum)), f


@pytest.mark.parametrize("chunk", [None, "23"])
def test_host_ballot_batch_dense_and_packed(engine, monkeypatch, chunk):
    """gcp_ballot_batch from host buffers: dense rows and arbo packed proofs, one chunk and many chunks."""
    if chunk:
        monkeypatch.setenv("GCP_B200_SMT_CHUNK", chunk)
    rng = random.Random(0xBA12)
    n_levels, n, nf = 64, 90, 3
    pk, items, ks, ms = _workload(rng, n_levels, n, nf)
    args = (n_levels, elems(it[0] for it in items), elems(it[2] for it in items), elems(it[3] for it in items), elems(pk),
            elems(ks).reshape(n, nf, 32), elems(ms).reshape(n, nf, 32))
    dense = engine.ballot_batch(*args, siblings=elems([s for it in items for s in it[1]]).reshape(n, n_levels, 32))
    _check(dense, pk, items, ks, ms, nf)
    packed = []
    for it in items:
        last = max((i for i, s in enumerate(it[1]) if s), default=-1)
        packed.append(osmt.pack_siblings(it[1][:last + 1]))
    packed[11] = packed[11][:-1]                                    # arbo.UnpackSiblings would reject this one
    out = engine.ballot_batch(*args, packed=packed)
    _check(out, pk, items, ks, ms, nf, bad={11})


def test_host_ballot_batch_through_a_group(engine):
    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import _lib

    rng = random.Random(0xBA13)
    n_levels, n, nf = 64, 70, 2
    pk, items, ks, ms = _workload(rng, n_levels, n, nf)
    args = (n_levels, elems(it[0] for it in items), elems(it[2] for it in items), elems(it[3] for it in items), elems(pk),
            elems(ks).reshape(n, nf, 32), elems(ms).reshape(n, nf, 32))
    sib = elems([s for it in items for s in it[1]]).reshape(n, n_levels, 32)
    devices = [0, 1] if _lib.load().gcp_device_count() >= 2 else [0]
    with g.Group(devices) as grp:
        out = grp.ballot_batch(*args, siblings=sib)
    _check(out, pk, items, ks, ms, nf)
    one = engine.ballot_batch(*args, siblings=sib)
    for a, b in zip(out, one):
        assert np.array_equal(a, b)


def test_host_ballot_batch_with_no_voters(engine):
    """Empty batch: the tally is NewCiphertext (identity, identity) per field (elgamal/ciphertext.go:16-19)."""
    nf = 2
    pk = ed.scalar_mul(ed.G, 0xB200)
    flags, status, tally, tstatus = engine.ballot_batch(
        64, np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8), elems(pk),
        np.zeros((0, nf, 32), np.uint8), np.zeros((0, nf, 32), np.uint8), siblings=np.zeros((0, 64, 32), np.uint8))
    assert flags.shape == (0,) and status.shape == (0,) and not tstatus.any()
    for f in range(nf):
        assert ints(tally[f]) == [0, 1, 0, 1]
