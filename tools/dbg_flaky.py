import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smoqyelph_b200 import model as mdl, api
from oracle import oracle as orc
import dense_ref as dr
m = mdl.config("cfg4")
rng = np.random.default_rng(0)
V, t = dr.build_Vt(m, m.random_fields(rng))
ref = orc.RefFDM(m); ref.update(V, t)
v = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
want = ref.mul_MtM(v)
nbad = 0
for rep in range(40):
    fdm = api.FermionDetMatrix(m); fdm.update(V, t)
    for (S, T) in [(1, 128), (1, 256), (3, 1024), (2, 512)]:
        fdm.set_tuning(S, T)
        g = fdm.mul_MtM(v)
        err = np.linalg.norm(g - want) / np.linalg.norm(want)
        if err > 1e-12:
            nbad += 1
            per = np.linalg.norm(g - want, axis=1) / np.linalg.norm(want, axis=1)
            bad = np.nonzero(per > 1e-12)[0]
            e, c, s = fdm.coefficients()
            print("BAD rep", rep, S, T, err, "nbad slices", len(bad), bad[:10], "coef err", np.abs(e - ref.expV).max(), np.abs(c - ref.cosh).max())
            g2 = fdm.mul_MtM(v)
            print("   retry err", np.linalg.norm(g2 - want) / np.linalg.norm(want))
    del fdm
print("done, bad =", nbad)
