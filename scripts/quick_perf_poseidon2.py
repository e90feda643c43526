"""Device-resident throughput of the width-2 Poseidon2 hasher and of MiMC7 (CUDA events, 3 launches each)."""
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
n = 1 << 22
status = torch.empty(n, dtype=torch.uint8, device="cuda")
for length in (2, 3):
    inp = rand_elems(torch, length * n, gen); out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: eng.poseidon2_hash_dev(inp, length, n, out, status, stream=st))
    # per permutation: 62 S-boxes = 124 squarings (100 wide multiplies) + 62 multiplies (128), + 2 conversions per limb
    wide = length * (124 * 100 + 62 * 128) + (length + 1) * 128
    print(f"poseidon2 limbs={length} n=2^22: {ms:.2f} ms  {n/ms/1e3:.1f} M hash/s  {n*wide/ms/1e9:.2f} T wide/s", flush=True)
inp = rand_elems(torch, 2 * n, gen); out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
ms = timeit(lambda: eng.poseidon2_permutation_dev(inp, n, out, status, stream=st))
print(f"poseidon2 permutation n=2^22: {ms:.2f} ms  {n/ms/1e3:.1f} M perm/s", flush=True)
n = 1 << 20
inp = rand_elems(torch, 2 * n, gen); out = torch.empty((n, 8), dtype=torch.int32, device="cuda")
ms = timeit(lambda: eng.mimc7_hash_dev(inp, 2, n, out, status[:n], stream=st))
print(f"mimc7 len=2 n=2^20: {ms:.2f} ms  {n/ms/1e3:.1f} M hash/s", flush=True)
