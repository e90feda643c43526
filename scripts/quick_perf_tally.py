"""Device-resident throughput of the plain tally (config 3's fold of Ciphertext.Add), both element formats.
The additions are data-independent, so random canonical elements stand in for ciphertexts (Add performs no curve check)."""
import sys
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems

eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
st = torch.cuda.current_stream()


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


nf = 8
for lognb in (16, 20, 22):
    nb = 1 << lognb; n = nb * nf
    ct = rand_elems(torch, n * 4, gen).reshape(n, 4, 8)
    tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
    for fmt, name in ((g.FMT_CANONICAL, "canonical"), (g.FMT_MONTGOMERY, "montgomery")):
        ms = timeit(lambda: eng.elgamal_tally_dev(ct, nb, nf, tout, tst, fmt=fmt, stream=st))
        print(f"tally {name} n_ballots=2^{lognb} x {nf}: {ms:.3f} ms  {n/ms/1e6:.2f} G ct/s  {n*128/ms/1e6:.1f} GB/s  status={tst.tolist()}", flush=True)
