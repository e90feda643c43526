"""Quick device-resident throughput probe of smt.Processor (update: old and new path, 2 Hash2 per level)."""
import sys, torch, ctypes
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems
from gnark_crypto_primitives_b200.engine import _dptr
eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
n_levels = 160
for logn, L in ((16, 159), (17, 159), (18, 24)):
    n = 1 << logn
    sib = rand_elems(torch, n * n_levels, gen, nonzero=True).view(n, n_levels, 8)
    sib[:, L:, :] = 0
    keys = rand_elems(torch, n, gen); keys[:, 5:] = 0
    ov, nv, roots = rand_elems(torch, n, gen), rand_elems(torch, n, gen), rand_elems(torch, n, gen)
    z = torch.zeros(n, dtype=torch.uint8, device="cuda"); o = torch.ones(n, dtype=torch.uint8, device="cuda")
    out = torch.empty((n, 8), dtype=torch.int32, device="cuda"); st = torch.empty(n, dtype=torch.uint8, device="cuda")
    def run():
        rc = eng._lib.gcp_smt_process_dev(eng._h, n_levels, n, _dptr(roots), _dptr(sib), _dptr(keys), _dptr(ov), _dptr(z),
                                          _dptr(keys), _dptr(nv), _dptr(z), _dptr(o), _dptr(out), _dptr(st), 0, None)
        assert rc == 0
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    run(); run()
    e1.record(torch.cuda.default_stream()); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"smt_process update n=2^{logn} path={L}: {ms:.2f} ms  {n/ms:.1f} k transitions/s  ({2*n*L/ms/1e3:.1f} M hash2/s)  status6={(st==6).float().mean().item():.2f}", flush=True)
