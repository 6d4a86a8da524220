import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smoqyelph_b200 import model as mdl, api
from oracle import oracle as orc
import dense_ref as dr
for m in [mdl.holstein_square(16, 16, 0.5), mdl.config("cfg4")]:
    rng = np.random.default_rng(0)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    ref = orc.RefFDM(m); ref.update(V, t)
    fdm = api.FermionDetMatrix(m); fdm.update(V, t)
    v = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    print(m.name, fdm.tuning)
    for S in (1, 2, 3):
        for T in (32, 64, 128, 256, 512, 1024):
            try: fdm.set_tuning(S, T)
            except Exception as e: print("skip", S, T); continue
            errs = []
            for op in ("mul_M", "mul_Mt", "mul_MtM"):
                g, w = getattr(fdm, op)(v), getattr(ref, op)(v)
                errs.append(np.linalg.norm(g - w) / np.linalg.norm(w))
            print(S, T, ["%.1e" % e for e in errs])
