import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smoqyelph_b200 import model as mdl, api
from oracle import oracle as orc
import dense_ref as dr
os.environ["SQ_SLAB"] = "2"; os.environ["SQ_THREADS"] = "64"
for sym in (True, False):
    m = mdl.holstein_ssh_chain(7, 0.25)
    rng = np.random.default_rng(0)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    ref = orc.RefFDM(m, sym=sym); ref.update(V, t)
    fdm = api.FermionDetMatrix(m, sym=sym); fdm.update(V, t)
    v = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    for op in ("mul_M", "mul_Mt", "mul_MtM"):
        g, w = getattr(fdm, op)(v), getattr(ref, op)(v)
        print(sym, op, np.linalg.norm(g - w) / np.linalg.norm(w))
    x, it, eps = fdm.ldiv(v, tol=1e-10)
    print("cg", it, eps)
