"""Small resident-CG solves for compute-sanitizer (racecheck / synccheck / memcheck): uniform square, honeycomb, per-bond square and chain.
compute-sanitizer --tool racecheck python tools/sanitize_resident.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smoqyelph_b200 import model as mdl, api
import dense_ref as dr

for name, m in (("h16x16", mdl.holstein_square(16, 16, 0.35)), ("hc8", mdl.holstein_honeycomb(8, 0.3)), ("bssh16", mdl.bssh_square(16, 16, 0.4)),
                ("ossh64", mdl.ossh_chain(64, 0.5))):
    rng = np.random.default_rng(0)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    fdm = api.FermionDetMatrix(m, sym=True)
    fdm.update(V, t)
    fdm.set_fast_path(2 + 256 * 3)
    b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    st0 = fdm.stats
    x, it, eps = fdm.ldiv(b, tol=1e-8, maxiter=60)
    st = fdm.stats
    print(name, "Ltau", m.Ltau, "iters", it, "eps", eps, "resident", st["cg_resident"] - st0["cg_resident"], flush=True)
