"""Device time of the fused M^T M v kernels at a named config (events inside the library, back-to-back and L2-cold)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from smoqyelph_b200 import model as mdl, api
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
m = mdl.config(name)
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.cdw_start(m, 0); elph.update_fdm()
n = m.N * m.Ltau
d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
res = {"config": name, "auto": fdm.tuning}
def t(op=2): return round(fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 300), 2), round(fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 31, flush.data_ptr(), flush.numel() * 4), 2)
res["auto_hot_cold_us"] = t()
fdm.set_fast_path(1); res["v2_hot_cold_us"] = t()
for S in (1, 2, 3, 4, 5):
    fdm.set_fast_path(2 + 256 * S)
    if fdm.tuning["path"] != 3: break
    res[f"v3_S{S}_hot_cold_us"] = t()
    res[f"v3_S{S}_M_hot_cold_us"] = t(0)
    res[f"v3_S{S}_native_hot_cold_us"] = t(102)
print(json.dumps(res))

import time
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
for S in (2, 3):
    fdm.set_fast_path(2 + 256 * S)
    for nit in (200, 2000):
        fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=nit)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        it, eps = fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=nit)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"S={S} CG iters {it} -> {dt/nit*1e6:.2f} us/iter")
