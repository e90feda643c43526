"""Dense 160-level proofs end to end from pageable vs page-locked (gcp_host_alloc) host memory."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import make_batch, N_LEVELS
eng = g.Engine(0)
n = 1 << 19
b = make_batch(torch, eng, n, seed=3)
sib = b["sib"].cpu().numpy().view(np.uint8).reshape(n, N_LEVELS, 32)          # pageable
roots, keys, vals = (b[k].cpu().numpy().view(np.uint8).reshape(n, 32) for k in ("roots", "keys", "vals"))
expect = b["expect"].cpu().numpy()
buf = g.PinnedBuffer(sib.nbytes)
pinned = buf.array.reshape(sib.shape); pinned[...] = sib
for name, arr in (("pageable", sib), ("pinned (gcp_host_alloc)", pinned)):
    eng.smt_verify_inclusion(roots, arr, keys, vals)
    t0 = time.perf_counter()
    f, s = eng.smt_verify_inclusion(roots, arr, keys, vals)
    dt = time.perf_counter() - t0
    print(f"{name}: {n/dt/1e3:.1f} k proofs/s  ok={bool((f == expect).all()) and not s.any()}", flush=True)
buf.close()
