import sys, torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from oracle import edwards as ed
from tests.util import elems
from bench import rand_elems
eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(0xB200)
st = torch.cuda.current_stream()
pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
n, nf = 1 << 20, 8
k = rand_elems(torch, n, gen); m = rand_elems(torch, n, gen); m[:, 1:] = 0; m[:, 0] &= 0xFFFF
out = torch.empty((n, 4, 8), dtype=torch.int32, device="cuda"); status = torch.empty(n, dtype=torch.uint8, device="cuda")
tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.elgamal_encrypt_dev(pk, False, k, m, n, out, status, stream=st)
    eng.elgamal_encrypt_tally_dev(pk, k, m, n // nf, nf, tout, tst, stream=st)
torch.cuda.synchronize()
print("ok")
