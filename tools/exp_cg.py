"""CG / trajectory timing experiments on a named config (GPU)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smoqyelph_b200 import model as mdl, api
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
Nt = int(sys.argv[2]) if len(sys.argv) > 2 else 24
m = mdl.config(name)
rng = np.random.default_rng(0)
x = mdl.thermal_fields(m, rng)
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = x; elph.update_fdm()
print(name, "tuning", fdm.tuning, "x rms", x.std())
b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
P = api.KPMPreconditioner(fdm)
print("kpm orders: max", P.orders.max(), "sum(all freq)", 2 * P.orders.sum(), "n>1:", int((P.orders > 1).sum()) * 2)
for tol in (1e-5, 1e-10):
    for pre in (None, P):
        fdm.ldiv(b, preconditioner=pre, tol=tol, maxiter=20000, refresh=False)
        t0 = time.perf_counter()
        xs, it, eps = fdm.ldiv(b, preconditioner=pre, tol=tol, maxiter=20000, refresh=False)
        dt = time.perf_counter() - t0
        print(f"tol {tol:g} precond {pre is not None}: iters {it} eps {eps:.2e} wall {dt*1e3:.1f} ms  -> {dt/max(it,1)*1e6:.1f} us/iter (incl. H2D/D2H)")
pff = api.PFFCalculator(elph)
for pre in (None, P):
    elph.x = x; elph.update_fdm()
    hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=Nt, seed=1)
    for rep in range(3):
        l0 = fdm.launch_count
        t0 = time.perf_counter()
        acc, its = hmc.hmc_update(preconditioner=pre, tol_action=1e-10, tol_force=1e-5, maxiter=10000)
        dt = time.perf_counter() - t0
        print(f"hmc precond {pre is not None}: acc {acc} iters_avg {its:.1f} dH {hmc.info[1]:.3f} wall {dt*1e3:.1f} ms launches {fdm.launch_count - l0}")
