"""Driver for an ncu launch list of the KPM-preconditioned CG: one warm solve, then one solve of a few iterations.  argv: config [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = mdl.config(name)
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.bench_state(m)[0] if m.name == "cfg4" else bench.cdw_start(m, 0) if (m.Nhol and len(m.lattice_dims) == 2 and m.N == m.lattice_dims[0] * m.lattice_dims[1]) else m.random_fields(np.random.default_rng(0), smooth=True)
elph.update_fdm()
fdm.set_fast_path(2 + 256 * 3)          # no timing-based tuning under a profiler: register path where it applies, else the fast shared-memory kernel
P = api.KPMPreconditioner(fdm)
n = m.N * m.Ltau
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=nit)
torch.cuda.synchronize()
l0 = fdm.launch_count
it, eps = fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=nit)
torch.cuda.synchronize()
print("iters", it, "launches", fdm.launch_count - l0, "orders max", P.orders.max(), "n>1", int((P.orders > 1).sum()))
