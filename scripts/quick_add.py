"""Element-wise Ciphertext.Add, device-resident, both element formats: CUDA-event time per call (and the subject of the
ncu launch list / --set full captures of ct_add_kernel and normalize_kernel).
  python scripts/quick_add.py [log2_ciphertexts] [iters]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g  # noqa: E402
from bench import rand_elems  # noqa: E402

logm = int(sys.argv[1]) if len(sys.argv) > 1 else 22
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m = 1 << logm
eng = g.Engine(0)
gen = torch.Generator(device="cuda")
gen.manual_seed(5)
st = torch.cuda.current_stream()
a = rand_elems(torch, m * 4, gen).reshape(m, 4, 8)
b = rand_elems(torch, m * 4, gen).reshape(m, 4, 8)
o = torch.empty((m, 4, 8), dtype=torch.int32, device="cuda")
so = torch.empty(m, dtype=torch.uint8, device="cuda")
for fmt, name in ((g.FMT_CANONICAL, "canonical"), (g.FMT_MONTGOMERY, "montgomery")):
    eng.elgamal_add_dev(a, b, m, o, so, fmt=fmt, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        eng.elgamal_add_dev(a, b, m, o, so, fmt=fmt, stream=st)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"kernel": "ct_add", "fmt": name, "n": m, "ms": ms, "ct_per_s": m / ms * 1e3}), flush=True)
