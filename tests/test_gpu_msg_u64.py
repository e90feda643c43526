"""GPU parity: GCP_MSG_U64 on the fused tallies (messages as little-endian uint64 instead of 32-byte field elements; the
reference's Encrypt takes msg as a frontend.Variable, elgamal/encrypt.go:42 - ballot fields are small integers).  The
compact form must give exactly the tally of the field-element form and of the oracle's closed form, through the host
pipelines (several chunks), the device forms, the Montgomery element format, the ballot batch and the group."""
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

PK = ed.scalar_mul(ed.G, 0xB200)
R_MONT = 1 << 256


def closed_form(pk, ks, ms, nb, nf):
    return [eg.serialize(eg.encrypt(pk, sum(ks[f::nf]) % ed.ORDER, sum(ms[f::nf]) % ed.ORDER)) for f in range(nf)]


def test_encrypt_tally_u64_messages_host(engine):
    from gnark_crypto_primitives_b200 import _lib

    rng = random.Random(0x64)
    nb, nf = 700, 3
    ks = [rng.randrange(R) for _ in range(nb * nf)]
    ms = [rng.randrange(1 << 16) for _ in range(nb * nf)]
    ms[:6] = [0, 1, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, (1 << 63) + 12345]      # the whole uint64 range
    k = elems(ks).reshape(nb, nf, 32)
    m64 = np.array(ms, dtype=np.uint64).reshape(nb, nf)
    got, st = engine.elgamal_encrypt_tally(elems(PK), k, m64, fmt=_lib.MSG_U64)
    assert not st.any()
    want, st2 = engine.elgamal_encrypt_tally(elems(PK), k, elems(ms).reshape(nb, nf, 32))
    assert not st2.any() and (got == want).all()
    assert [ints(got[f]) for f in range(nf)] == closed_form(PK, ks, ms, nb, nf)
    # fr.Element memory for k (Montgomery form), integers for m
    k_mont = elems([(x * R_MONT) % R for x in ks]).reshape(nb, nf, 32)
    got_m, st = engine.elgamal_encrypt_tally(elems([(c * R_MONT) % R for c in PK]), k_mont, m64,
                                             fmt=_lib.FMT_MONTGOMERY | _lib.MSG_U64)
    assert not st.any()
    assert [[(v * pow(R_MONT, -1, R)) % R for v in ints(got_m[f])] for f in range(nf)] == closed_form(PK, ks, ms, nb, nf)
    # a non-canonical k still gives status 1 on its field; a wrong dtype is refused by the mirror
    bad = k.copy()
    bad[3, 1] = elems([R])[0]
    _, st = engine.elgamal_encrypt_tally(elems(PK), bad, m64, fmt=_lib.MSG_U64)
    assert [int(s) for s in st] == [0, 1, 0]
    with pytest.raises(TypeError):
        engine.elgamal_encrypt_tally(elems(PK), k, m64.astype(np.int32), fmt=_lib.MSG_U64)
    # empty batch: the identity ciphertext per field
    z, st = engine.elgamal_encrypt_tally(elems(PK), np.empty((0, nf, 32), np.uint8), np.empty((0, nf), np.uint64), fmt=_lib.MSG_U64)
    assert not st.any() and all(ints(z[f]) == [0, 1, 0, 1] for f in range(nf))


def test_encrypt_tally_u64_messages_in_several_chunks(monkeypatch):
    """The host pipeline with 1 MB chunks: the message slices must advance by 8 bytes per message."""
    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import _lib

    rng = random.Random(0x65)
    nb, nf = 40000, 2                       # 2.56 MB of scalars: three chunks
    k = np.frombuffer(rng.randbytes(nb * nf * 32), dtype=np.uint8).reshape(nb, nf, 32).copy()
    k[:, :, 31] &= 0x0F                     # canonical
    m64 = np.frombuffer(rng.randbytes(nb * nf * 8), dtype=np.uint64).reshape(nb, nf).copy()
    m64 >>= np.uint64(40)
    ms_fr = np.zeros((nb, nf, 32), dtype=np.uint8)
    ms_fr[:, :, :8] = m64.view(np.uint8).reshape(nb, nf, 8)
    import subprocess, sys, json, os, tempfile
    # GCP_B200_ET_CHUNK_MB is read once per process: run the chunked call in a child
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "k.npy"), k)
        np.save(os.path.join(d, "m.npy"), m64)
        code = ("import sys, numpy as np; sys.path.insert(0, '.');"
                "import gnark_crypto_primitives_b200 as g; from gnark_crypto_primitives_b200 import _lib;"
                "from oracle import edwards as ed; from tests.util import elems;"
                f"k = np.load(r'{d}/k.npy'); m = np.load(r'{d}/m.npy'); e = g.Engine(0);"
                "out, st = e.elgamal_encrypt_tally(elems(ed.scalar_mul(ed.G, 0xB200)), k, m, fmt=_lib.MSG_U64);"
                f"assert not st.any(); np.save(r'{d}/out.npy', out)")
        env = dict(os.environ, GCP_B200_ET_CHUNK_MB="1")
        subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
        chunked = np.load(os.path.join(d, "out.npy"))
    eng = g.Engine(0)
    try:
        want, st = eng.elgamal_encrypt_tally(elems(PK), k, ms_fr)
        assert not st.any() and (chunked == want).all()
    finally:
        eng.close()


def test_device_forms_take_u64_messages(engine):
    from gnark_crypto_primitives_b200 import _lib

    rng = random.Random(0x66)
    nb, nf, n_levels = 120, 4, 64
    ks = [rng.randrange(R) for _ in range(nb * nf)]
    ms = [rng.randrange(1 << 40) for _ in range(nb * nf)]
    dk = torch.from_numpy(elems(ks)).cuda()
    dm = torch.from_numpy(np.array(ms, dtype=np.uint64).view(np.int64)).cuda()
    dpk = torch.from_numpy(elems(PK)).cuda()
    out = torch.empty((nf, 4, 32), dtype=torch.uint8, device="cuda")
    st = torch.empty(nf, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    engine.elgamal_encrypt_tally_dev(dpk, dk, dm, nb, nf, out, st, fmt=_lib.MSG_U64, stream=stream)
    torch.cuda.synchronize()
    assert not bool(st.any())
    assert [ints(out.cpu().numpy()[f]) for f in range(nf)] == closed_form(PK, ks, ms, nb, nf)
    # ballot batch: every 5th voter presents a wrong root and is left out of the tally
    items = [census_proof(rng, n_levels, lo=2, hi=20) for _ in range(nb)]
    for i in range(0, nb, 5):
        r, s, k, v = items[i]
        items[i] = ((r + 1) % R, s, k, v)
    flags = torch.empty(nb, dtype=torch.uint8, device="cuda")
    status = torch.empty(nb, dtype=torch.uint8, device="cuda")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    engine.ballot_batch_dev(n_levels, nb, dev(elems(it[0] for it in items)), False,
                            dev(elems([s for it in items for s in it[1]])), dev(elems(it[2] for it in items)),
                            dev(elems(it[3] for it in items)), dpk, dk, dm, nf, flags, status, out, st,
                            fmt=_lib.MSG_U64, stream=stream)
    torch.cuda.synchronize()
    want = [osmt.inclusion_verifier(*it) for it in items]
    assert [int(f) for f in flags.cpu()] == [w[0] for w in want]
    adm = [i for i, w in enumerate(want) if w[0] == 1 and w[1] == 0]
    assert 0 < len(adm) < nb and not bool(st.any())
    got = out.cpu().numpy()
    for f in range(nf):
        ksum = sum(ks[i * nf + f] for i in adm) % ed.ORDER
        msum = sum(ms[i * nf + f] for i in adm) % ed.ORDER
        assert ints(got[f]) == eg.serialize(eg.encrypt(PK, ksum, msum)), f
    # host form of the ballot batch, dense rows
    fl, stt, tal, tst = engine.ballot_batch(n_levels, elems(it[0] for it in items), elems(it[2] for it in items),
                                            elems(it[3] for it in items), elems(PK), elems(ks).reshape(nb, nf, 32),
                                            np.array(ms, dtype=np.uint64).reshape(nb, nf),
                                            siblings=elems([s for it in items for s in it[1]]).reshape(nb, n_levels, 32),
                                            fmt=_lib.MSG_U64)
    assert (fl == flags.cpu().numpy()).all() and (tal == got).all() and not tst.any()


def test_group_takes_u64_messages():
    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import _lib

    rng = random.Random(0x67)
    nb, nf = 501, 2
    ks = [rng.randrange(R) for _ in range(nb * nf)]
    ms = [rng.randrange(1 << 20) for _ in range(nb * nf)]
    devices = list(range(min(2, torch.cuda.device_count())))
    with g.Group(devices) as grp:
        out, st = grp.elgamal_encrypt_tally(elems(PK), elems(ks).reshape(nb, nf, 32),
                                            np.array(ms, dtype=np.uint64).reshape(nb, nf), fmt=_lib.MSG_U64)
    assert not st.any()
    assert [ints(out[f]) for f in range(nf)] == closed_form(PK, ks, ms, nb, nf)
