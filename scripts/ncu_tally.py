"""Two launches of tally_partial_kernel at 2^20 ballots x 8 fields (standard-form, then Montgomery-form inputs) for
`ncu --set full -k regex:tally_partial_kernel`."""
import sys
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems

eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
st = torch.cuda.current_stream()
nf, nb = 8, 1 << 20
ct = rand_elems(torch, nb * nf * 4, gen).reshape(nb * nf, 4, 8)
tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
for fmt in (g.FMT_CANONICAL, g.FMT_MONTGOMERY):
    eng.elgamal_tally_dev(ct, nb, nf, tout, tst, fmt=fmt, stream=st)
    torch.cuda.synchronize()
print("done")
