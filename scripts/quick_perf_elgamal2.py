import sys, torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from oracle import edwards as ed
from tests.util import elems
from bench import rand_elems
eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(0xB200)
st = torch.cuda.current_stream()
pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for logn in (20, 23):
    n = 1 << logn
    k = rand_elems(torch, n, gen); m = rand_elems(torch, n, gen); m[:, 1:] = 0; m[:, 0] &= 0xFFFF
    out = torch.empty((n, 4, 8), dtype=torch.int32, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.elgamal_encrypt_dev(pk, False, k, m, n, out, status, stream=st))
    print(f"encrypt n=2^{logn}: {ms:.2f} ms  {n/ms/1e3:.1f} M enc/s", flush=True)
    for nf in (8, 1):
        nb = n // nf
        tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
        ms = timeit(lambda: eng.elgamal_encrypt_tally_dev(pk, k, m, nb, nf, tout, tst, stream=st))
        print(f"encrypt_tally {nb} x {nf}: {ms:.2f} ms  {n/ms/1e3:.1f} M enc/s", flush=True)
