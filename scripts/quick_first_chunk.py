"""Host-buffer SMT verification with different first-chunk sizes (GCP_B200_FIRST_DIV): census-like 2^18 and dense 2^19."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import gnark_crypto_primitives_b200 as g
from bench import N_LEVELS, make_batch, make_census_like

eng = g.Engine(0)
lib, h = eng._lib, eng._h
c = make_census_like(torch, eng, 1 << 18)
n = 1 << 18
of, os_ = np.empty(n, np.uint8), np.empty(n, np.uint8)
def census():
    rc = lib.gcp_smt_verify_inclusion(h, N_LEVELS, n, c["hr"].ctypes.data, 0, c["dense"].ctypes.data, c["hk"].ctypes.data,
                                      c["hv"].ctypes.data, of.ctypes.data, os_.ctypes.data, None, 0)
    assert rc == 0
def timed(fn, iters=4):
    fn(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    return (time.perf_counter() - t0) / iters
dt = timed(census)
res = {"first_div": os.environ.get("GCP_B200_FIRST_DIV", "1"), "census_dense_rows_e2e_per_s": n / dt, "ok": bool((of == c["expect"]).all())}
del c
nd = 1 << 19
b = make_batch(torch, eng, nd, seed=3)
hb = {k: torch.empty(b[k].shape, dtype=b[k].dtype).pin_memory() for k in ("sib", "keys", "vals", "roots")}
for k in hb: hb[k].copy_(b[k])
exp = b["expect"].cpu().numpy(); del b
of2, os2 = np.empty(nd, np.uint8), np.empty(nd, np.uint8)
def dense():
    rc = lib.gcp_smt_verify_inclusion(h, N_LEVELS, nd, hb["roots"].data_ptr(), 0, hb["sib"].data_ptr(), hb["keys"].data_ptr(),
                                      hb["vals"].data_ptr(), of2.ctypes.data, os2.ctypes.data, None, 0)
    assert rc == 0
dt = timed(dense, iters=2)
res["dense_2p19_e2e_per_s"] = nd / dt
res["dense_ok"] = bool((of2 == exp).all())
print(json.dumps(res))
