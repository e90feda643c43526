"""Launches each compute-bound kernel of the path twice at a moderate size (for `ncu --set full -k regex:...`)."""
import sys
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import rand_elems
from oracle import edwards as ed
from tests.util import elems

eng = g.Engine(0)
gen = torch.Generator(device="cuda"); gen.manual_seed(11)
st = torch.cuda.current_stream()
pk = torch.from_numpy(elems(ed.scalar_mul(ed.G, 0xB200))).cuda()
n = 1 << 20
inp = rand_elems(torch, 2 * n, gen); out = torch.empty((n, 8), dtype=torch.int32, device="cuda"); stt = torch.empty(n, dtype=torch.uint8, device="cuda")
n16 = 1 << 17
inp16 = rand_elems(torch, 16 * n16, gen)
ne = 1 << 19
k = rand_elems(torch, ne, gen); m = rand_elems(torch, ne, gen); m[:, 1:] = 0; m[:, 0] &= 0xFFFF
ct = torch.empty((ne, 4, 8), dtype=torch.int32, device="cuda"); est = torch.empty(ne, dtype=torch.uint8, device="cuda")
nf = 8
tout = torch.empty((nf, 4, 8), dtype=torch.int32, device="cuda"); tst = torch.empty(nf, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.poseidon_hash_dev(inp, 2, n, out, stt, stream=st)
    eng.poseidon_hash_dev(inp16, 16, n16, out, stt, stream=st)
    eng.elgamal_encrypt_dev(pk, False, k, m, ne, ct, est, stream=st)
    eng.elgamal_tally_dev(ct, ne // nf, nf, tout, tst, stream=st)
    eng.elgamal_encrypt_tally_dev(pk, k, m, ne // nf, nf, tout, tst, stream=st)
    torch.cuda.synchronize()
print("done")
