"""Quick device-resident throughput probe (not the bench): Poseidon t=3 batch and dense SMT proofs."""
import sys, time
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g

def rand_elems(n, gen):
    x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=gen)
    x = x ^ (torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=gen) << 1)
    x[:, 7] &= 0x0FFFFFFF   # < 2^252 < r: canonical
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
for logn in (20, 22):
    n = 1 << logn
    inp = rand_elems(2 * n, gen); out = torch.empty((n, 8), dtype=torch.int32, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.poseidon_hash_dev(inp, 2, n, out, status, stream=st))
    print(f"poseidon t=3 n=2^{logn}: {ms:.3f} ms  {n/ms/1e3:.2f} Mhash/s", flush=True)
n_levels = 160
for logn in (16, 18):
    n = 1 << logn
    sib = rand_elems(n * n_levels, gen).view(n, n_levels, 8); sib[:, n_levels - 1, :] = 0
    keys = rand_elems(n, gen); keys[:, 5:] = 0
    vals = rand_elems(n, gen); roots = rand_elems(n, gen)
    flags = torch.empty(n, dtype=torch.uint8, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.smt_verify_dev(n_levels, n, roots, False, sib, keys, vals, flags, status, stream=st), iters=2)
    print(f"smt dense n_levels=160 n=2^{logn}: {ms:.2f} ms  {n/ms:.1f} kproofs/s  ({n*159/ms/1e3:.2f} Mhash/s) status_nonzero={int((status!=0).sum())}", flush=True)
