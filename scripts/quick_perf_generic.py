import sys, torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems
eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(3)
st = torch.cuda.current_stream()
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for arity, logn in ((1, 20), (4, 20), (12, 18), (16, 18)):
    n = 1 << logn
    inp = rand_elems(torch, arity * n, gen); out = torch.empty((n, 8), dtype=torch.int32, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.poseidon_hash_dev(inp, arity, n, out, status, stream=st))
    print(f"poseidon arity={arity} (t={arity+1}) n=2^{logn}: {ms:.2f} ms  {n/ms/1e3:.2f} M hash/s", flush=True)
for length, logn in ((60, 16), (256, 14)):
    n = 1 << logn
    inp = rand_elems(torch, length * n, gen); out = torch.empty((n, 8), dtype=torch.int32, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.poseidon_multihash_dev(inp, length, n, out, status, stream=st))
    print(f"multihash len={length} n=2^{logn}: {ms:.2f} ms  {n/ms/1e3:.3f} M multihash/s  ({n*length/ms/1e3:.1f} M inputs/s)", flush=True)
