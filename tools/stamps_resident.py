"""Per-phase cycle counts of the resident CG kernel (profiling build: fdm_v3.cu compiled with -DSQ_V3_STAMPS, see tools/README.md).
SMOQYELPH_B200_LIB=tools/_stamps/libsmoqyelph_b200_stamps.so SQ_DEBUG_STAMPS=1 python tools/stamps_resident.py [config] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.bench_state(m)[0] if m.name == "cfg4" else m.random_fields(np.random.default_rng(0), smooth=True)
elph.update_fdm()
n = m.N * m.Ltau
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=50)
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=nit)
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / nit * 1e6:.2f} us per iteration (stamped build)")
