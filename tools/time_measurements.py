"""Device time of the measurement pieces at a named config: Nrv estimator solves, G(Δ,0), density correlation."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
m = mdl.config(name)
fdm = api.SymFermionDetMatrix(m, tol=1e-10, maxiter=100000)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.cdw_start(m, 0) if name == "cfg4" else m.random_fields(np.random.default_rng(0), smooth=True)
elph.update_fdm()
g = api.GreensEstimator(fdm, Nrv=10, seed=1)
def timed(fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps, out
res = {"config": name}
dt, it = timed(lambda: g.update_greens_estimator(tol=1e-10))
res["update_greens_estimator_ms"] = round(dt * 1e3, 1); res["avg_cg_iters"] = it
norb = m.N // int(np.prod(m.lattice_dims))
dt, _ = timed(lambda: g.measure_GD0((0, 0)), 3); res["measure_GD0_ms"] = round(dt * 1e3, 2)
dt, _ = timed(lambda: g.measure_density_correlation(0, norb - 1), 2); res["density_correlation_ms"] = round(dt * 1e3, 2)
dt, _ = timed(lambda: g.measure(), 3); res["scalar_measurements_ms"] = round(dt * 1e3, 2)
print(json.dumps(res))
