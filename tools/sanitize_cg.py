"""Small resident-CG solve for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smoqyelph_b200 import model as mdl, api
import dense_ref as dr
kind = sys.argv[1] if len(sys.argv) > 1 else "square"
m = mdl.holstein_square(16, 16, 0.4) if kind == "square" else mdl.holstein_honeycomb(8, 0.4)
rng = np.random.default_rng(0)
V, t = dr.build_Vt(m, m.random_fields(rng))
fdm = api.FermionDetMatrix(m, sym=True)
fdm.update(V, t)
fdm.set_fast_path(2 + 256 * 2)
b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
x, it, eps = fdm.ldiv(b, tol=1e-8, maxiter=60)
y = fdm.mul_MtM(b)
print("ok", fdm.tuning, it, eps)
