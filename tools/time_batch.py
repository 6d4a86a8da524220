"""update_greens_estimator! (Nrv = 10) with the KPM preconditioner: batched (cg_batch.cu) against one-by-one.  argv: config"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
x, _ = bench.bench_state(m) if m.name == "cfg4" else (m.random_fields(np.random.default_rng(0), smooth=True), "")
elph.x = x; elph.update_fdm()
P = api.KPMPreconditioner(fdm)
for mode in ("batch", "seq", "plain"):
    os.environ.pop("SQ_NO_BATCH_CG", None)
    if mode == "seq": os.environ["SQ_NO_BATCH_CG"] = "1"
    g = api.GreensEstimator(fdm, Nrv=10, seed=3)
    pre = None if mode == "plain" else P
    g.update_greens_estimator(preconditioner=pre, tol=1e-10)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g2 = api.GreensEstimator(fdm, Nrv=10, seed=4)
    it = g2.update_greens_estimator(preconditioner=pre, tol=1e-10)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{m.name} Nrv=10 {mode}: {dt*1e3:.1f} ms, avg iterations {it:.1f} -> {dt/(10*max(it,1))*1e6:.1f} us per system-iteration; stats {fdm.stats['cg_batched_rhs']}")
