"""One launch of each secondary kernel at a moderate size, device-resident, for `ncu --set full -k regex:...`:
keccak_address_kernel, mimc7_kernel, poseidon2_hash_kernel, poseidon2_permutation_kernel, smt_unpack_kernel,
smt_process_kernel, ct_add_kernel, tally_partial_kernel, encrypt_tally_partial_kernel."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g  # noqa: E402
from bench import N_LEVELS, make_census_like, rand_elems  # noqa: E402
from gnark_crypto_primitives_b200.engine import _dptr  # noqa: E402

eng = g.Engine(0)
gen = torch.Generator(device="cuda")
gen.manual_seed(21)
st = torch.cuda.current_stream()
which = set(sys.argv[1:]) or {"keccak", "mimc7", "poseidon2", "unpack", "process", "add", "tally", "fused", "lanes"}
u8 = lambda *shape: torch.empty(shape, dtype=torch.uint8, device="cuda")

if "keccak" in which:
    n = 1 << 22
    pub = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=gen)
    eng.keccak_address_dev(pub, n, u8(n, 20), stream=st)
if "lanes" in which:
    n = 1024                                       # BASELINE config 1: the three-lanes-per-hash latency layout
    eng.poseidon_hash_dev(rand_elems(torch, 2 * n, gen), 2, n, u8(n, 32), u8(n), stream=st)
if "mimc7" in which:
    n = 1 << 20
    eng.mimc7_hash_dev(rand_elems(torch, 2 * n, gen), 2, n, u8(n, 32), u8(n), stream=st)
if "poseidon2" in which:
    n = 1 << 21
    x = rand_elems(torch, 2 * n, gen)
    eng.poseidon2_hash_dev(x, 2, n, u8(n, 32), u8(n), stream=st)
    eng.poseidon2_permutation_dev(x, n, u8(n, 2, 32), u8(n), stream=st)
if "unpack" in which or "process" in which:
    n = 1 << 17
    c = make_census_like(torch, eng, n)
    if "unpack" in which:
        blob = torch.from_numpy(c["blob"]).cuda()
        offs = torch.from_numpy(c["offs"].astype(np.int64)).cuda()
        eng.smt_unpack_siblings_dev(N_LEVELS, n, blob, blob.numel(), offs, torch.empty_like(c["sib"]), u8(n), stream=st)
    if "process" in which:
        z, o = torch.zeros(n, dtype=torch.uint8, device="cuda"), torch.ones(n, dtype=torch.uint8, device="cuda")
        rc = eng._lib.gcp_smt_process_dev(eng._h, N_LEVELS, n, _dptr(c["roots"]), _dptr(c["sib"]), _dptr(c["keys"]),
                                          _dptr(c["vals"]), _dptr(z), _dptr(c["keys"]), _dptr(c["vals"]), _dptr(z), _dptr(o),
                                          _dptr(u8(n, 32)), _dptr(u8(n)), 0, eng._stream(st))
        assert rc == 0
if which & {"add", "tally", "fused"}:
    sk = torch.zeros((1, 8), dtype=torch.int32, device="cuda")
    sk[0, 0] = 0xB200
    pk = u8(1, 2, 32)
    eng.elgamal_fixed_base_mul_dev(sk, 1, pk, u8(1), stream=st)
    n = 1 << 20
    k, m = rand_elems(torch, n, gen), rand_elems(torch, n, gen)
    m[:, 1:] = 0
    ct = u8(n, 4, 32)
    eng.elgamal_encrypt_dev(pk, False, k, m, n, ct, u8(n), stream=st)
    if "add" in which:
        eng.elgamal_add_dev(ct, ct.roll(1, 0).contiguous(), n, u8(n, 4, 32), u8(n), stream=st)
    big = ct.repeat(8, 1, 1).contiguous()
    if "tally" in which:
        eng.elgamal_tally_dev(big, 1 << 20, 8, u8(8, 4, 32), u8(8), stream=st)
    if "fused" in which:
        eng.elgamal_encrypt_tally_dev(pk, k, m, n // 8, 8, u8(8, 4, 32), u8(8), stream=st)
torch.cuda.synchronize()
print("done")
