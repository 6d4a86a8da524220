"""Resident-CG iteration time of the 64-site SSH chain against the number of CTAs (Ltau = 40 ... 320, S = 2 / 3 slices per CTA): the iteration
is a chain of L2 round trips, independent of how many CTAs take part."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, time
from smoqyelph_b200 import model as mdl, api
for beta in (2.0, 4.0, 8.0, 16.0):
    for S in (2, 3):
        os.environ["SQ_V3_RESIDENT_SLAB"] = str(S)
        m = mdl.ossh_chain(64, beta)
        fdm = api.FermionDetMatrix(m, sym=True)
        elph = api.ElectronPhononParameters(m, fdm)
        elph.x = m.random_fields(np.random.default_rng(0), smooth=True)
        elph.update_fdm()
        n = m.N * m.Ltau
        b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
        fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=40)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        nit, rep = 40, 50          # short solves: past convergence the residual hits exact zero
        for _ in range(rep): fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=nit)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(rep): fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, tol=1e-300, maxiter=2 * nit)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        st = fdm.stats
        print(f"Ltau {m.Ltau} S {S} ctas {(m.Ltau + S - 1)//S}: {((t2 - t1) - (t1 - t0)) / (rep * nit) * 1e6:.2f} us/iter (difference of 80- and 40-iteration solves) resident={st['cg_resident']}", flush=True)
