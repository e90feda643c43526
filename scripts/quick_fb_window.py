"""Fixed-base window width A/B on one box: build time of G's table and fused encrypt + tally throughput (2^24 ballots x 8
fields resident, CUDA events) with both tables forced to 20, 22, 24 and 26 bits, then the automatic policy.
  python scripts/quick_fb_window.py [log2_ballots]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g  # noqa: E402
from bench import rand_elems  # noqa: E402
from oracle import edwards as ed  # noqa: E402
from tests.util import elems  # noqa: E402

lognb = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nb, nf = 1 << lognb, 8
n = nb * nf
eng = g.Engine(0)
gen = torch.Generator(device="cuda")
gen.manual_seed(5)
st = torch.cuda.current_stream()
pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
k = rand_elems(torch, n, gen)
m = rand_elems(torch, n, gen)
m[:, 1:] = 0
m[:, 0] &= 0xFFFF
tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda")
tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
ref = None
for bits in (20, 22, 24, 26, 0):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.set_fixed_base_window(bits)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    eng.elgamal_encrypt_tally_dev(pk, k, m, nb, nf, tout, tst, stream=st)  # builds the key's table at this width
    torch.cuda.synchronize()
    first_ms = (time.perf_counter() - t0) * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        eng.elgamal_encrypt_tally_dev(pk, k, m, nb, nf, tout, tst, stream=st)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    same = True if ref is None else bool((tout == ref).all())
    ref = tout.clone() if ref is None else ref
    print(json.dumps({"forced_bits": bits, "g_table_build_ms": round(build_ms, 2), "first_call_ms": round(first_ms, 2),
                      "ms": round(ms, 3), "enc_per_s": n / ms * 1e3, "width_g": eng.fixed_base_window(0),
                      "width_pk": eng.fixed_base_window(1), "tally_equal_to_20_bit": same, "status_clean": not bool(tst.any()),
                      "mem_gb": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)}), flush=True)
