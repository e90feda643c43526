"""Throughput of the secondary kernels of the path, one JSON line per measurement (CUDA events on the launching stream
for device-resident forms, wall clock around the blocking host-buffer calls).  Not the bench: bench.py is the judged line;
this is the before/after instrument for kernel work.

  python scripts/perf_suite.py [section ...] [--tag NAME]     sections: process verify add tally varbase keccak fused hashes
"""
import json
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g  # noqa: E402
from bench import rand_elems  # noqa: E402
from gnark_crypto_primitives_b200.engine import _dptr  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
TAG = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")), "")
SECTIONS = set(args) or {"process", "verify", "add", "tally", "varbase", "keccak", "fused", "hashes"}
eng = g.Engine(0)
gen = torch.Generator(device="cuda")
gen.manual_seed(5)
st = torch.cuda.current_stream()
N_LEVELS = 160


def emit(**kw):
    kw["tag"] = TAG
    print(json.dumps(kw), flush=True)


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def smt_inputs(n, L):
    """L leading non-zero siblings of 160 (L = 159: dense; L = 24: census-like length, no interior zeros)."""
    sib = rand_elems(torch, n * N_LEVELS, gen, nonzero=True).view(n, N_LEVELS, 8)
    sib[:, L:, :] = 0
    keys = rand_elems(torch, n, gen)
    keys[:, 5:] = 0
    return sib, keys


if "process" in SECTIONS or "verify" in SECTIONS:
    for logn, L in ((17, 159), (19, 24)):
        n = 1 << logn
        sib, keys = smt_inputs(n, L)
        ov, nv, roots = rand_elems(torch, n, gen), rand_elems(torch, n, gen), rand_elems(torch, n, gen)
        z = torch.zeros(n, dtype=torch.uint8, device="cuda")
        o = torch.ones(n, dtype=torch.uint8, device="cuda")
        out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
        stt = torch.empty(n, dtype=torch.uint8, device="cuda")
        flags = torch.empty(n, dtype=torch.uint8, device="cuda")
        if "verify" in SECTIONS:
            ms = timeit(lambda: eng.smt_verify_dev(N_LEVELS, n, roots, False, sib, keys, nv, flags, stt, stream=st))
            emit(kernel="smt_verify", n=n, path=L, ms=ms, per_s=n / ms * 1e3, hash2_per_s=n * L / ms * 1e3)
        if "process" in SECTIONS:
            def run():
                rc = eng._lib.gcp_smt_process_dev(eng._h, N_LEVELS, n, _dptr(roots), _dptr(sib), _dptr(keys), _dptr(ov),
                                                  _dptr(z), _dptr(keys), _dptr(nv), _dptr(z), _dptr(o), _dptr(out),
                                                  _dptr(stt), 0, eng._stream(st))
                assert rc == 0
            ms = timeit(run)
            emit(kernel="smt_process_update", n=n, path=L, ms=ms, per_s=n / ms * 1e3, hash2_per_s=2 * n * L / ms * 1e3,
                 status6_frac=float((stt == 6).float().mean().item()))
        del sib

if "process" in SECTIONS:
    # ragged path lengths (census-like: L ~ U[20,28], ~10 % interior zeros): updates with a matching old root, so status 0
    from bench import make_census_like
    n = 1 << 18
    c = make_census_like(torch, eng, n, host_forms=False)
    c["vals"][::16, 0] ^= 2                                   # undo the bench's corruption: every proof valid
    z = torch.zeros(n, dtype=torch.uint8, device="cuda")
    o = torch.ones(n, dtype=torch.uint8, device="cuda")
    nv = rand_elems(torch, n, gen)
    out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    stt = torch.empty(n, dtype=torch.uint8, device="cuda")
    flags = torch.empty(n, dtype=torch.uint8, device="cuda")

    def run():
        rc = eng._lib.gcp_smt_process_dev(eng._h, N_LEVELS, n, _dptr(c["roots"]), _dptr(c["sib"]), _dptr(c["keys"]), _dptr(c["vals"]),
                                          _dptr(z), _dptr(c["keys"]), _dptr(nv), _dptr(z), _dptr(o), _dptr(out), _dptr(stt), 0,
                                          eng._stream(st))
        assert rc == 0
    ms = timeit(run)
    ok = float((stt == 0).float().mean().item())
    msv = timeit(lambda: eng.smt_verify_dev(N_LEVELS, n, c["roots"], False, c["sib"], c["keys"], c["vals"], flags, stt, stream=st))
    emit(kernel="smt_process_update_census_like", n=n, mean_path=c["mean_levels"], ms=ms, per_s=n / ms * 1e3, status0_frac=ok,
         verifier_per_s=n / msv * 1e3, ratio_to_verifier=msv / ms)
    del c

if "hashes" in SECTIONS:
    n = 1 << 22
    inp = rand_elems(torch, 2 * n, gen)
    dig = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    stt = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.poseidon_hash_dev(inp, 2, n, dig, stt, stream=st))
    emit(kernel="poseidon_hash2", n=n, ms=ms, per_s=n / ms * 1e3)
    del inp
    for nsmall in (256, 1024, 4096, 8192):
        a = np.random.default_rng(1).integers(0, 256, size=(nsmall, 2, 32), dtype=np.uint8)
        a[:, :, 31] &= 0x1F
        eng.poseidon_hash(a)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            eng.poseidon_hash(a)
            ts.append(time.perf_counter() - t0)
        d_in = torch.from_numpy(a).cuda()
        d_out = torch.empty((nsmall, 32), dtype=torch.uint8, device="cuda")
        d_st = torch.empty(nsmall, dtype=torch.uint8, device="cuda")
        ms = timeit(lambda: eng.poseidon_hash_dev(d_in, 2, nsmall, d_out, d_st, stream=st), iters=20, warm=3)
        import os
        emit(kernel="poseidon_hash2_host_call", n=nsmall, us=float(np.median(ts)) * 1e6, kernel_only_us=ms * 1e3,
             lanes_layout=(os.environ.get("GCP_B200_NO_LANES") is None and nsmall <= 4096))

if "add" in SECTIONS or "tally" in SECTIONS:
    nf = 8
    for lognb in (20,):
        nb = 1 << lognb
        n = nb * nf
        ct = rand_elems(torch, n * 4, gen).reshape(n, 4, 8)
        tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda")
        tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
        if "tally" in SECTIONS:
            for fmt, name in ((g.FMT_CANONICAL, "canonical"), (g.FMT_MONTGOMERY, "montgomery")):
                ms = timeit(lambda: eng.elgamal_tally_dev(ct, nb, nf, tout, tst, fmt=fmt, stream=st), iters=5)
                emit(kernel="tally", fmt=name, n_ballots=nb, n_fields=nf, ms=ms, ct_per_s=n / ms * 1e3,
                     gb_per_s=n * 128 / ms / 1e6)
        if "add" in SECTIONS:
            m = 1 << 22
            a, b = ct[:m], ct[m:2 * m]
            o = torch.empty((m, 4, 8), dtype=torch.int32, device="cuda")
            so = torch.empty(m, dtype=torch.uint8, device="cuda")
            for fmt, name in ((g.FMT_CANONICAL, "canonical"), (g.FMT_MONTGOMERY, "montgomery")):
                ms = timeit(lambda: eng.elgamal_add_dev(a, b, m, o, so, fmt=fmt, stream=st), iters=5)
                emit(kernel="ct_add", fmt=name, n=m, ms=ms, ct_per_s=m / ms * 1e3)
        del ct

if "fused" in SECTIONS:
    from oracle import edwards as ed
    from tests.util import elems
    pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
    nf = 8
    for lognb in (20, 24):
        nb = 1 << lognb
        n = nb * nf
        k = rand_elems(torch, n, gen)
        m = rand_elems(torch, n, gen)
        m[:, 1:] = 0
        m[:, 0] &= 0xFFFF
        tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda")
        tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
        ms = timeit(lambda: eng.elgamal_encrypt_tally_dev(pk, k, m, nb, nf, tout, tst, stream=st), iters=2)
        emit(kernel="encrypt_tally", n_ballots=nb, n_fields=nf, ms=ms, enc_per_s=n / ms * 1e3)
        if lognb == 20:
            ct = torch.empty((n, 4, 8), dtype=torch.int32, device="cuda")
            est = torch.empty(n, dtype=torch.uint8, device="cuda")
            ms = timeit(lambda: eng.elgamal_encrypt_dev(pk, False, k, m, n, ct, est, stream=st), iters=2)
            emit(kernel="encrypt_shared", n=n, ms=ms, enc_per_s=n / ms * 1e3)
            del ct
        del k, m

if "keccak" in SECTIONS:
    n = 1 << 24
    inp = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=gen)
    out = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.keccak_address_dev(inp, n, out, stream=st), iters=5)
    emit(kernel="keccak_address", n=n, ms=ms, per_s=n / ms * 1e3, gb_per_s=n * 84 / ms / 1e6)
    del inp, out

if "varbase" in SECTIONS:
    from oracle import eddsa as oeddsa
    from oracle import edwards as ed
    from oracle import elgamal as eg
    from tests.test_gpu_proofs import make_proof
    from tests.util import elems
    rng = random.Random(1)
    N = 1 << 19
    base = 8

    def tile(a, reps):
        return np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))

    def timed(name, fn):
        fn()
        t0 = time.perf_counter()
        ok = fn()
        dt = time.perf_counter() - t0
        emit(kernel=name, n=N, ms=dt * 1e3, per_s=N / dt, ok=bool(ok), via="host API, pageable numpy")

    pks, ks, ms_ = [], [], []
    for _ in range(base):
        pks.append(ed.scalar_mul(ed.G, rng.randrange(1, ed.ORDER)))
        ks.append(rng.randrange(1 << 253))
        ms_.append(rng.randrange(1 << 16))
    pk_a = tile(elems([c for p in pks for c in p]).reshape(base, 2, 32), N // base)
    k_a, m_a = tile(elems(ks), N // base), tile(elems(ms_), N // base)
    want = eg.serialize(eg.encrypt(pks[0], ks[0], ms_[0]))

    def enc():
        ct, s = eng.elgamal_encrypt(pk_a, k_a, m_a)
        return (not s.any()) and bytes(ct[0].reshape(-1)) == b"".join(int(v).to_bytes(32, "little") for v in want)
    timed("encrypt_per_key", enc)

    items = []
    for _ in range(base):
        d = rng.randrange(1, ed.ORDER)
        msg = rng.randrange(1000)
        c = eg.encrypt(ed.scalar_mul(ed.G, d), rng.randrange(ed.ORDER), msg)
        items.append((c, d, msg))
    ct_a = tile(elems([x for it in items for x in eg.serialize(it[0])]).reshape(base, 4, 32), N // base)
    d_a, msg_a = tile(elems(it[1] for it in items), N // base), tile(elems(it[2] for it in items), N // base)

    def adec():
        f, s = eng.elgamal_assert_decrypt(ct_a, d_a, msg_a)
        return bool(f.all()) and not s.any()
    timed("assert_decrypt", adec)

    items = []
    for _ in range(base):
        msg = rng.randrange(1000)
        pk_, c, a1, a2, z = make_proof(rng, rng.randrange(1, ed.ORDER), msg)
        items.append((pk_, c, msg, a1, a2, z))
    pargs = [tile(elems([c for it in items for c in it[0]]).reshape(base, 2, 32), N // base),
             tile(elems([x for it in items for x in eg.serialize(it[1])]).reshape(base, 4, 32), N // base),
             tile(elems(it[2] for it in items), N // base),
             tile(elems([c for it in items for c in it[3]]).reshape(base, 2, 32), N // base),
             tile(elems([c for it in items for c in it[4]]).reshape(base, 2, 32), N // base),
             tile(elems(it[5] for it in items), N // base)]

    def dproof():
        f, s = eng.elgamal_verify_decryption_proof(*pargs)
        return bool(f.all()) and not s.any()
    timed("decryption_proof_verify", dproof)

    items = []
    for _ in range(base):
        msg = rng.getrandbits(248)
        a, r, s = oeddsa.sign(rng.randrange(1, ed.ORDER), rng.randrange(1, ed.ORDER), msg)
        items.append((a, r, s, msg))
    eargs = [tile(elems([c for it in items for c in it[0]]).reshape(base, 2, 32), N // base),
             tile(elems([c for it in items for c in it[1]]).reshape(base, 2, 32), N // base),
             tile(elems(it[2] for it in items), N // base), tile(elems(it[3] for it in items), N // base)]

    def edd():
        f, s = eng.eddsa_verify(*eargs)
        return bool(f.all()) and not s.any()
    timed("eddsa_verify", edd)

eng.close()
