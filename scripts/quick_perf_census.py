"""Census-like SMT distribution (SURVEY 8d secondary): L ~ U[20,28] leading siblings, ~10% interior zeros, rest 0."""
import sys, torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems

eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
st = torch.cuda.current_stream()
n_levels = 160
for logn in (20,):
    n = 1 << logn
    sib = rand_elems(torch, n * n_levels, gen, nonzero=True).view(n, n_levels, 8)
    L = torch.randint(20, 29, (n,), device="cuda", generator=gen)
    lev = torch.arange(n_levels, device="cuda").view(1, n_levels)
    keep = lev < L.view(n, 1)
    interior_zero = (torch.rand((n, n_levels), device="cuda", generator=gen) < 0.1) & (lev < (L.view(n, 1) - 1))
    sib[~keep | interior_zero] = 0
    keys = rand_elems(torch, n, gen); keys[:, 5:] = 0
    vals = rand_elems(torch, n, gen); roots = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
    flags = torch.empty(n, dtype=torch.uint8, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
    tmp = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    eng.smt_verify_dev(n_levels, n, roots, False, sib, keys, vals, flags, status, d_out_roots=tmp, stream=st)
    torch.cuda.synchronize(); roots.copy_(tmp)
    def run(): eng.smt_verify_dev(n_levels, n, roots, False, sib, keys, vals, flags, status, stream=st)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    hashes = float(L.sum().item())
    print(f"census n=2^{logn}: {ms:.2f} ms  {n/ms/1e3:.2f} M proofs/s  ({hashes/ms/1e3:.1f} M hash2/s + leaf)  {n*5216/ms/1e6:.1f} GB/s  flags_ok={bool(flags.all())} status_clean={not bool(status.any())}")
