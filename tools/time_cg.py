"""Per-iteration device time of the CG loop at cfg4 (fixed iteration count)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.cdw_start(m, 0); elph.update_fdm()
n = m.N * m.Ltau
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
P = api.KPMPreconditioner(fdm) if len(sys.argv) > 2 and sys.argv[2] == "kpm" else None
for nit in (200, 2000):
    fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=nit)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    it, eps = fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=nit)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"fusion={'off' if os.environ.get('SQ_NO_CG_FUSION') else 'on'} batch={os.environ.get('SQ_CG_BATCH','16')} kpm={P is not None} iters {it} -> {dt/nit*1e6:.2f} us/iter  tuning {fdm.tuning}")
