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


def test_config4_keccak_2pow24_sampled(engine):
    """BASELINE config 4 at its stated size: 2^24 public keys -> addresses on the device, 8 192 of them (plus the public
    priv = 1 vector) against the oracle; the host-buffer form on a slice gives the same bytes."""
    from oracle import cport

    n = 1 << 24
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xADD2)
    pub = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=gen)
    g_xy = bytes.fromhex("79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"
                         "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8")   # secp256k1 G (priv = 1)
    pub[12345] = torch.tensor(list(g_xy), dtype=torch.uint8, device="cuda")
    addr = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    engine.keccak_address_dev(pub, n, addr, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert bytes(addr[12345].cpu().numpy()).hex() == "7e5f4552091a69125d5dfcb7b8c2659029395bdf"
    idx = torch.arange(0, n, n // 8192, device="cuda")
    want = cport.keccak_address(pub[idx].cpu().numpy(), threads=cport.default_threads())
    assert (want == addr[idx].cpu().numpy()).all()
    m = 1 << 20
    host = engine.keccak_address(pub[:m].cpu().numpy())
    assert (host == addr[:m].cpu().numpy()).all()


def test_config5_ballot_batch_streamed_closed_form(engine):
    """BASELINE config 5 in miniature with the bench's own shape: voters stream through gcp_ballot_batch_dev in chunks
    (census-like 160-level proofs, every 16th wrong; 8 encrypted fields per voter), the chunk tallies are folded.  Flags
    must follow the construction, a sample of proofs the oracle's literal verifier, and the tally the closed form
    sum over ADMITTED voters: Encrypt(pk, sum k, sum m) - a voter with a bad proof must not be counted."""
    from bench import N_LEVELS, make_census_like, rand_elems
    from oracle import cport
    from oracle import edwards as ed
    from oracle import elgamal as eg
    from tests.util import elems, ints

    chunk, n_chunks, nf = 1 << 17, 4, 8
    pk_int = ed.scalar_mul(ed.G, 0xB200)
    pk = torch.from_numpy(elems(pk_int)).cuda()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(55)
    stream = torch.cuda.current_stream()
    parts = torch.empty((n_chunks, nf, 4, 32), dtype=torch.uint8, device="cuda")
    pst = torch.empty((n_chunks, nf), dtype=torch.uint8, device="cuda")
    ksum = torch.zeros((nf, 8), dtype=torch.int64, device="cuda")
    msum = torch.zeros(nf, dtype=torch.int64, device="cuda")
    for c in range(n_chunks):
        cen = make_census_like(torch, engine, chunk, seed=500 + c, host_forms=False)
        k = rand_elems(torch, chunk * nf, gen)
        m = torch.zeros((chunk * nf, 8), dtype=torch.int32, device="cuda")
        m[:, 0] = torch.randint(0, 1 << 16, (chunk * nf,), dtype=torch.int32, device="cuda", generator=gen)
        flags = torch.empty(chunk, dtype=torch.uint8, device="cuda")
        status = torch.empty(chunk, dtype=torch.uint8, device="cuda")
        engine.ballot_batch_dev(N_LEVELS, chunk, cen["roots"], False, cen["sib"], cen["keys"], cen["vals"], pk, k, m, nf, flags,
                                status, parts[c], pst[c], stream=stream)
        torch.cuda.synchronize()
        expect = torch.from_numpy(cen["expect"]).cuda()
        assert torch.equal(flags, expect) and not bool(status.any())
        adm = expect.to(torch.int64).view(chunk, 1, 1)
        ksum += ((k.view(chunk, nf, 8).to(torch.int64) & 0xFFFFFFFF) * adm).sum(0)
        msum += (m.view(chunk, nf, 8)[:, :, 0].to(torch.int64) * adm.view(chunk, 1)).sum(0)
        if c == 0:   # a sample of this chunk's proofs, valid and wrong alike, against the literal verifier
            idx = torch.arange(0, chunk, chunk // 128, device="cuda")
            f, s, _ = cport.smt_verify(as_bytes(cen["roots"][idx]), as_bytes(cen["sib"][idx]), as_bytes(cen["keys"][idx]),
                                       as_bytes(cen["vals"][idx]), literal=True, threads=cport.default_threads())
            assert (f == flags[idx].cpu().numpy()).all() and not s.any()
    out = torch.empty((nf, 4, 32), dtype=torch.uint8, device="cuda")
    ost = torch.empty(nf, dtype=torch.uint8, device="cuda")
    engine.elgamal_tally_dev(parts, n_chunks, nf, out, ost, stream=stream)
    torch.cuda.synchronize()
    assert not bool(pst.any()) and not bool(ost.any())
    got = out.cpu().numpy()
    ks = ksum.cpu().tolist()
    for f in range(nf):
        kf = sum(int(v) << (32 * l) for l, v in enumerate(ks[f])) % ed.ORDER
        assert ints(got[f]) == eg.serialize(eg.encrypt(pk_int, kf, int(msum[f].item()) % ed.ORDER)), f
