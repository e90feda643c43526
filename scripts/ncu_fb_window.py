"""One fused encrypt + tally launch per window width (20 and 24 bits) at 2^19 ballots x 8 fields, device-resident, for
`ncu --set full -k regex:encrypt_tally_partial_kernel` (and the table-build kernels with -k regex:fb_table)."""
import sys

import torch

sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g  # noqa: E402
from bench import rand_elems  # noqa: E402
from oracle import edwards as ed  # noqa: E402
from tests.util import elems  # noqa: E402

nb, nf = 1 << 19, 8
eng = g.Engine(0)
gen = torch.Generator(device="cuda")
gen.manual_seed(5)
st = torch.cuda.current_stream()
pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
k = rand_elems(torch, nb * nf, gen)
m = rand_elems(torch, nb * nf, gen)
m[:, 1:] = 0
m[:, 0] &= 0xFFFF
tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda")
tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
for bits in (20, 24):
    eng.set_fixed_base_window(bits)
    eng.elgamal_encrypt_tally_dev(pk, k, m, nb, nf, tout, tst, stream=st)
    torch.cuda.synchronize()
    assert not bool(tst.any())
print("done")
