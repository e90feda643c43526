"""Fused encrypt+tally from page-locked host scalars (config 3 e2e, 2^23 ballots x 8) for GCP_B200_ET_CHUNK_MB."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems
from oracle import edwards as oed
from tests.util import elems

eng = g.Engine(0)
nb, nf = 1 << 23, 8
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
k = rand_elems(torch, nb * nf, gen)
m = torch.zeros((nb * nf, 8), dtype=torch.int32, device="cuda")
m[:, 0] = torch.randint(0, 1 << 16, (nb * nf,), dtype=torch.int32, device="cuda", generator=gen)
hk = torch.empty(k.shape, dtype=k.dtype).pin_memory(); hk.copy_(k)
hm = torch.empty(m.shape, dtype=m.dtype).pin_memory(); hm.copy_(m)
pk = elems(oed.scalar_mul(oed.G, 0xB200))
out = np.empty((nf, 4, 32), np.uint8); st = np.empty(nf, np.uint8)
res = torch.empty((nf, 4, 32), dtype=torch.uint8, device="cuda"); rst = torch.empty(nf, dtype=torch.uint8, device="cuda")
pkd = torch.from_numpy(pk).cuda()
def dev():
    eng.elgamal_encrypt_tally_dev(pkd, k, m, nb, nf, res, rst, stream=torch.cuda.current_stream()); torch.cuda.synchronize()
def host():
    rc = eng._lib.gcp_elgamal_encrypt_tally(eng._h, pk.ctypes.data, hk.data_ptr(), hm.data_ptr(), nb, nf, out.ctypes.data, st.ctypes.data, 0)
    assert rc == 0
def timed(fn, it=3):
    fn(); t0 = time.perf_counter()
    for _ in range(it): fn()
    return (time.perf_counter() - t0) / it
td, th = timed(dev), timed(host)
print(json.dumps({"chunk_mb": os.environ.get("GCP_B200_ET_CHUNK_MB", "64"), "resident_enc_per_s": nb * nf / td, "host_enc_per_s": nb * nf / th,
                  "same": bool((out == res.cpu().numpy()).all())}))
