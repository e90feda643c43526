"""Quick device-resident throughput probe for the ElGamal kernels (not the bench)."""
import sys
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from oracle import edwards as ed
from tests.util import elems

def rand_elems(n, gen, bits=252):
    lo = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=gen)
    hi = torch.randint(0, 2, (n, 8), dtype=torch.int32, device="cuda", generator=gen)
    x = lo | (hi << 31)
    x[:, 7] &= 0x0FFFFFFF
    if bits <= 32:
        x[:, 1:] = 0
        x[:, 0] &= (1 << bits) - 1
    return x

def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(0xB200)
st = torch.cuda.current_stream()
pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
for logn in (18, 20):
    n = 1 << logn
    k = rand_elems(n, gen); m = rand_elems(n, gen, bits=16)
    out = torch.empty((n, 4, 8), dtype=torch.int32, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.elgamal_encrypt_dev(pk, False, k, m, n, out, status, stream=st))
    print(f"encrypt shared-pk n=2^{logn}: {ms:.2f} ms  {n/ms/1e3:.2f} M enc/s  bad={int((status!=0).sum())}", flush=True)
    ms2 = timeit(lambda: eng.elgamal_fixed_base_mul_dev(k, n, out.view(-1)[: n * 16], status, stream=st))
    print(f"fixed-base n=2^{logn}: {ms2:.2f} ms  {n/ms2/1e3:.2f} M/s", flush=True)
nf = 8
for lognb in (17, 20):
    nb = 1 << lognb
    n = nb * nf
    k = rand_elems(min(n, 1 << 20), gen); m = rand_elems(min(n, 1 << 20), gen, bits=16)
    base = torch.empty((min(n, 1 << 20), 4, 8), dtype=torch.int32, device="cuda"); status = torch.empty(min(n, 1 << 20), dtype=torch.uint8, device="cuda")
    eng.elgamal_encrypt_dev(pk, False, k, m, base.shape[0], base, status, stream=st)
    torch.cuda.synchronize()
    ct = base.repeat((n // base.shape[0], 1, 1)).contiguous()
    tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.elgamal_tally_dev(ct, nb, nf, tout, tst, stream=st))
    print(f"tally n_ballots=2^{lognb} x {nf}: {ms:.2f} ms  {n/ms/1e3:.1f} M ct/s  {n*128/ms/1e6:.1f} GB/s  bad={int(tst.sum())}", flush=True)
    a = ct[: 1 << 20].contiguous(); b = ct[1 << 10: (1 << 20) + (1 << 10)].contiguous() if n > (1 << 20) + (1 << 10) else a
    na = a.shape[0]
    o = torch.empty_like(a); s2 = torch.empty(na, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.elgamal_add_dev(a, b, na, o, s2, stream=st))
    print(f"ct add n={na}: {ms:.2f} ms  {na/ms/1e3:.1f} M add/s", flush=True)
# fused encrypt + tally (config 3 shape)
for lognb in (20,):
    nb = 1 << lognb
    k = rand_elems(nb * nf, gen); m = rand_elems(nb * nf, gen, bits=16)
    tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.elgamal_encrypt_tally_dev(pk, k, m, nb, nf, tout, tst, stream=st))
    print(f"encrypt_tally n_ballots=2^{lognb} x {nf}: {ms:.2f} ms  {nb*nf/ms/1e3:.1f} M enc/s  bad={int(tst.sum())}", flush=True)
