"""Fixed cost of one CG solve (conversions, state set-up, cooperative launch, read-back): solves with maxiter = 1."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config("cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.cdw_start(m, 0); elph.update_fdm()
n = m.N * m.Ltau
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
for nit in (1, 2, 101):
    fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=nit)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
        fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=nit)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print(f"maxiter {nit}: {dt*1e6:.1f} us per solve")
