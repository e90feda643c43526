"""Single-process multi-GPU through the C ABI (gcp_group_*): dense SMT proofs and the fused encrypt+tally from pinned
host buffers, for every group size 1..visible GPUs (powers of two).  Not the contract bench (that is bench.py under
torchrun); this is the path a Go host takes.  Prints one JSON line per group size."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from gnark_crypto_primitives_b200 import _lib
from bench import make_batch, make_census_like, rand_elems, N_LEVELS

log2_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 19
visible = _lib.load().gcp_device_count()
sizes = [s for s in (1, 2, 4, 8) if s <= visible]
torch.cuda.set_device(0)
eng0 = g.Engine(0)
n_max = (1 << log2_per_gpu) * sizes[-1]
# one dense batch generated on GPU 0 in pieces, kept in pinned host memory
piece = 1 << min(log2_per_gpu, 18)
host = {k: torch.empty((n_max,) + shape, dtype=torch.int32).pin_memory()
        for k, shape in (("sib", (N_LEVELS, 8)), ("keys", (8,)), ("vals", (8,)), ("roots", (8,)))}
expect = np.empty(n_max, dtype=np.uint8)
for off in range(0, n_max, piece):
    b = make_batch(torch, eng0, piece, seed=0xB200 + off)
    for k in host:
        host[k][off:off + piece].copy_(b[k])
    expect[off:off + piece] = b["expect"].cpu().numpy()
    del b
torch.cuda.synchronize()
hn = {k: v.numpy().view(np.uint8).reshape(v.shape[:-1] + (32,)) for k, v in host.items()}
# ballots: k uniform, m < 2^16, 8 fields
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
nb_per_gpu, nf = 1 << 20, 8
kk = rand_elems(torch, nb_per_gpu * nf * sizes[-1], gen).cpu().numpy().view(np.uint8).reshape(-1, nf, 32)
mm = np.zeros_like(kk); mm[:, :, 0:2] = kk[:, :, 4:6]
sk = np.zeros((1, 32), np.uint8); sk[0, 0] = 0xB2
pk, _ = eng0.elgamal_fixed_base_mul(sk)
# config 5 shape: census-like packed proofs, 8 fields per voter, 2^17 voters per GPU
nv_per_gpu = 1 << 17
cen = make_census_like(torch, eng0, nv_per_gpu * sizes[-1])
eng0.close()
tallies = {}
for s in sizes:
    n = (1 << log2_per_gpu) * s
    with g.Group(list(range(s))) as grp:
        args = (hn["roots"][:n], hn["sib"][:n], hn["keys"][:n], hn["vals"][:n])
        grp.smt_verify(*args)                      # warm: device pools
        t0 = time.perf_counter()
        flags, status = grp.smt_verify(*args)
        dt = time.perf_counter() - t0
        ok = bool((flags == expect[:n]).all()) and not status.any()
        nb = nb_per_gpu * s
        grp.elgamal_encrypt_tally(pk[0], kk[:nb], mm[:nb])
        t0 = time.perf_counter()
        out, st = grp.elgamal_encrypt_tally(pk[0], kk[:nb], mm[:nb])
        dt2 = time.perf_counter() - t0
        tallies[s] = out
        nv = nv_per_gpu * s
        lens = cen["offs"][:nv + 1]
        bargs = (N_LEVELS, cen["hr"][:nv], cen["hk"][:nv], cen["hv"][:nv], pk[0], kk[:nv], mm[:nv])
        packed_list = None
        lib = grp._lib
        import ctypes
        def ballot():
            fl = np.empty(nv, np.uint8); stv = np.empty(nv, np.uint8); tl = np.empty((nf, 4, 32), np.uint8); ts = np.empty(nf, np.uint8)
            rc = lib.gcp_group_ballot_batch(grp._h, N_LEVELS, nv, cen["hr"].ctypes.data, 0, None, cen["blob"].ctypes.data,
                                            cen["offs"].ctypes.data, cen["hk"].ctypes.data, cen["hv"].ctypes.data,
                                            np.ascontiguousarray(pk[0]).ctypes.data, kk.ctypes.data, mm.ctypes.data, nf,
                                            fl.ctypes.data, stv.ctypes.data, tl.ctypes.data, ts.ctypes.data, 0)
            assert rc == 0, lib.gcp_group_last_error(grp._h)
            return fl, stv, tl, ts
        ballot()
        t0 = time.perf_counter()
        fl, stv, tl, ts = ballot()
        dt3 = time.perf_counter() - t0
        ballot_ok = bool((fl == cen["expect"][:nv]).all()) and not stv.any() and not ts.any()
        print(json.dumps({"group_size": s, "voters": nv, "ballot_batch_e2e_voters_per_s": nv / dt3, "ballot_flags_ok": ballot_ok, "uses_nccl": grp.uses_nccl, "proofs": n, "smt_e2e_proofs_per_s": n / dt,
                          "flags_ok": ok, "ballots": nb, "fields": nf, "encrypt_tally_e2e_enc_per_s": nb * nf / dt2,
                          "tally_status_clean": not st.any()}), flush=True)
# the tally of the first nb_per_gpu ballots must not depend on the group size
with g.Group([0]) as one:
    ref, _ = one.elgamal_encrypt_tally(pk[0], kk[:nb_per_gpu * sizes[-1]], mm[:nb_per_gpu * sizes[-1]])
print(json.dumps({"largest_group_tally_equals_single_device": bool(np.array_equal(ref, tallies[sizes[-1]]))}))
