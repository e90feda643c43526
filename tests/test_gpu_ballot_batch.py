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
