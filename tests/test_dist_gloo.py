"""CPU, world_size 2, gloo: the N > 1 host logic (index sharding + the tally all-gather plumbing).

The GPU engine cannot run here, so the per-device reduction is played by the oracle's C port; what is under test is
gnark_crypto_primitives_b200.dist: shard bounds, byte all-gather in rank order, tally-of-partials == tally-of-all."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnark_crypto_primitives_b200 import dist as gdist
from tests.util import elems

N_BALLOTS, N_FIELDS = 37, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_cts():
    from oracle import cport
    from oracle import edwards as ed
    from oracle.field import R

    rng = random.Random(99)
    pk = ed.scalar_mul(ed.G, 0xB200)
    n = N_BALLOTS * N_FIELDS
    ks = [rng.randrange(R) for _ in range(n)]
    ms = [rng.randrange(1 << 16) for _ in range(n)]
    cts, st = cport.elgamal_encrypt(elems(pk), elems(ks), elems(ms), threads=4)
    assert not st.any()
    return cts.reshape(N_BALLOTS, N_FIELDS, 4, 32)


def _oracle_tally(ct, n_ballots, n_fields):
    from oracle import cport

    out, st = cport.elgamal_tally(ct.numpy().reshape(n_ballots, n_fields, 4, 32))
    assert not st.any()
    return torch.from_numpy(out.copy())


def _worker(rank, world, port, cts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = gdist.shard_bounds(N_BALLOTS, world, rank)
        local = torch.from_numpy(cts[lo:hi].copy())
        total = gdist.sharded_tally(local, N_FIELDS, _oracle_tally)
        gathered = gdist.allgather_partials(torch.full((N_FIELDS, 4, 32), rank, dtype=torch.uint8))
        q.put((rank, total.numpy().tobytes(), [int(gathered[r].max()) for r in range(world)], (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 37, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [gdist.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        gdist.shard_bounds(4, 2, 2)


def test_single_process_path_needs_no_group():
    cts = _make_cts()
    whole = _oracle_tally(torch.from_numpy(cts), N_BALLOTS, N_FIELDS)
    got = gdist.sharded_tally(torch.from_numpy(cts), N_FIELDS, _oracle_tally)
    assert torch.equal(got, whole)


@pytest.mark.timeout(300)
def test_world2_gloo_tally_matches_whole():
    cts = _make_cts()
    whole = _oracle_tally(torch.from_numpy(cts), N_BALLOTS, N_FIELDS).numpy().tobytes()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cts, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, order, span in results:
        assert total == whole, rank                     # identical on every rank, identical to the 1-process tally
        assert order == [0, 1]                          # gathered in rank order
    assert sorted(r[3] for r in results) == [(0, 18), (18, 37)]
