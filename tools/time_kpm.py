import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.cdw_start(m, 0); elph.update_fdm()
P = api.KPMPreconditioner(fdm)
print("orders max", P.orders.max(), "sum", 2 * P.orders.sum())
n = m.N * m.Ltau
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
st = torch.cuda.ExternalStream(fdm.stream)
with torch.cuda.stream(st):
    for _ in range(3): P.ldiv_dev(x.data_ptr(), b.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20): P.ldiv_dev(x.data_ptr(), b.data_ptr())
    e1.record(st); e1.synchronize()
print(f"kpm ldiv: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per apply (slow={os.environ.get('SQ_KPM_SLOW')})")
bb = np.asfortranarray(np.random.default_rng(0).standard_normal((m.Ltau, m.N)) + 0j)
for tol in (1e-5, 1e-10):
    for pre in (None, P):
        t0 = time.perf_counter(); xs, it, eps = fdm.ldiv(bb, preconditioner=pre, tol=tol, maxiter=20000, refresh=False); dt = time.perf_counter() - t0
        print(f"tol {tol:g} precond {pre is not None}: iters {it} wall {dt*1e3:.1f} ms -> {dt/max(it,1)*1e6:.1f} us/iter")
