import sys, torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems
eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(3)
st = torch.cuda.current_stream()
n_levels, n = 160, 1 << 20
sib = rand_elems(torch, n * n_levels, gen, nonzero=True).view(n, n_levels, 8)
sib[:, n_levels - 1, :] = 0
L = torch.randint(1, 159, (n,), device="cuda", generator=gen)
lev = torch.arange(n_levels, device="cuda").view(1, n_levels)
sib[(lev >= L.view(n, 1)).unsqueeze(-1).expand(-1, -1, 8)] = 0
lidx = torch.empty(n, dtype=torch.int16, device="cuda"); info = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3): eng.smt_scan_dev(n_levels, n, sib, lidx, info, stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.smt_scan_dev(n_levels, n, sib, lidx, info, stream=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ok = bool((lidx.to(torch.int64) == L).all()) and bool((info == 3).all())
print(f"smt_scan n=2^20 x 160 levels: {ms:.3f} ms  {n*n_levels*32/ms/1e6:.0f} GB/s  lidx_ok={ok}")
