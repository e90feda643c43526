"""GPU parity at BASELINE.json's full sizes, through size-independent properties plus sampled oracle checks.

config 2: 2^20 dense 160-level inclusion proofs on one GPU (flags follow the construction; a sample is re-verified
          by the oracle's literal state machine).
config 3: 2^24 ballots x 8 fields, Encrypt + homomorphic tally: sum_i Encrypt(pk, k_i, m_i) must equal
          Encrypt(pk, sum k_i mod l, sum m_i mod l) — one O(1) oracle computation pins 2^27 encryptions.
config 1 at scale: 2^22 two-input hashes, sampled against the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def rand_limbs(n, gen):
    """(n, 8) int32 limbs, each in [0, 2^31), top limb < 2^28: canonical elements below 2^252."""
    x = torch.randint(0, 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=gen)
    x[:, 7] &= 0x0FFFFFFF
    return x


def as_bytes(t):
    return t.cpu().numpy().view(np.uint8).reshape(t.shape[:-1] + (32,))


def test_config2_smt_2pow20_dense(engine):
    from bench import N_LEVELS, make_batch
    from oracle import cport

    n = 1 << 20
    batch = make_batch(torch, engine, n, seed=0xB200)
    stream = torch.cuda.current_stream()
    roots_out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    engine.smt_verify_dev(N_LEVELS, n, batch["roots"], False, batch["sib"], batch["keys"], batch["vals"], batch["flags"],
                          batch["status"], d_out_roots=roots_out, stream=stream)
    torch.cuda.synchronize()
    assert not bool(batch["status"].any())
    assert bool((batch["flags"] == batch["expect"]).all())
    assert int(batch["flags"].sum()) == n - n // 16
    # idempotence: a second pass gives identical outputs
    flags2 = torch.empty_like(batch["flags"])
    status2 = torch.empty_like(batch["status"])
    roots2 = torch.empty_like(roots_out)
    engine.smt_verify_dev(N_LEVELS, n, batch["roots"], False, batch["sib"], batch["keys"], batch["vals"], flags2, status2,
                          d_out_roots=roots2, stream=stream)
    torch.cuda.synchronize()
    assert torch.equal(flags2, batch["flags"]) and torch.equal(roots2, roots_out)
    # sampled literal re-verification (valid and corrupted proofs alike)
    idx = torch.cat([torch.arange(0, n, n // 192, device="cuda"), torch.arange(0, 64 * 16, 16, device="cuda")])
    f, s, r = cport.smt_verify(as_bytes(batch["roots"][idx]), as_bytes(batch["sib"][idx]), as_bytes(batch["keys"][idx]),
                               as_bytes(batch["vals"][idx]), literal=True, threads=cport.default_threads())
    assert (f == batch["flags"][idx].cpu().numpy()).all() and not s.any()
    assert (r == as_bytes(roots_out[idx])).all()


def test_config3_elgamal_2pow24_ballots_x8_closed_form(engine):
    from oracle import edwards as ed
    from oracle import elgamal as eg
    from tests.util import elems, ints

    n_ballots, n_fields = 1 << 24, 8
    n = n_ballots * n_fields
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xE16A)
    k = rand_limbs(n, gen)                        # ~half of these exceed the subgroup order
    m = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
    m[:, 0] = torch.randint(0, 1 << 16, (n,), dtype=torch.int32, device="cuda", generator=gen)
    pk_int = ed.scalar_mul(ed.G, 0xB200)
    pk = torch.from_numpy(elems(pk_int)).cuda()
    out = torch.empty((n_fields, 4, 32), dtype=torch.uint8, device="cuda")
    status = torch.empty(n_fields, dtype=torch.uint8, device="cuda")
    engine.elgamal_encrypt_tally_dev(pk, k, m, n_ballots, n_fields, out, status, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert not bool(status.any())
    ksum = k.view(n_ballots, n_fields, 8).to(torch.int64).sum(0).cpu().tolist()
    msum = m.view(n_ballots, n_fields, 8)[:, :, 0].to(torch.int64).sum(0).cpu().tolist()
    got = out.cpu().numpy()
    for f in range(n_fields):
        ks = sum(v << (32 * l) for l, v in enumerate(ksum[f])) % ed.ORDER
        assert ints(got[f]) == eg.serialize(eg.encrypt(pk_int, ks, msum[f] % ed.ORDER)), f
    # the materialised path agrees on a 2^20-ballot slice: tally(encrypt(.)) == encrypt_tally(.)
    nb2 = 1 << 20
    n2 = nb2 * n_fields
    ct = torch.empty((n2, 4, 32), dtype=torch.uint8, device="cuda")
    st = torch.empty(n2, dtype=torch.uint8, device="cuda")
    engine.elgamal_encrypt_dev(pk, False, k[:n2], m[:n2], n2, ct, st, stream=torch.cuda.current_stream())
    t1 = torch.empty((n_fields, 4, 32), dtype=torch.uint8, device="cuda")
    s1 = torch.empty(n_fields, dtype=torch.uint8, device="cuda")
    engine.elgamal_tally_dev(ct, nb2, n_fields, t1, s1, stream=torch.cuda.current_stream())
    t2 = torch.empty_like(t1)
    engine.elgamal_encrypt_tally_dev(pk, k[:n2], m[:n2], nb2, n_fields, t2, s1, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert not bool(st.any()) and torch.equal(t1, t2)
    # and a sample of the materialised ciphertexts matches the oracle
    from oracle import cport
    idx = torch.arange(0, n2, n2 // 64, device="cuda")
    want, wst = cport.elgamal_encrypt(elems(pk_int), as_bytes(k[idx]), as_bytes(m[idx]), threads=cport.default_threads())
    assert (want == ct[idx].cpu().numpy()).all()


def test_config1_poseidon_2pow22_sampled(engine):
    from oracle import cport

    n = 1 << 22
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    inp = rand_limbs(2 * n, gen)
    out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    status = torch.empty(n, dtype=torch.uint8, device="cuda")
    engine.poseidon_hash_dev(inp, 2, n, out, status, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert not bool(status.any())
    idx = torch.arange(0, n, n // 1024, device="cuda")
    want, st = cport.poseidon_hash(as_bytes(inp.view(n, 2, 8)[idx]), threads=cport.default_threads())
    assert (want == as_bytes(out[idx])).all()
