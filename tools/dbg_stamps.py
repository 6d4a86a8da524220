import sys, os
os.environ["SQ_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
for name, S, T in [("cfg4", 3, 1024), ("cfg4", 3, 512), ("cfg4", 1, 512), ("cfg1", 2, 128), ("cfg3", 1, 256)]:
    m = mdl.config(name)
    rng = np.random.default_rng(0)
    fdm = api.FermionDetMatrix(m, sym=True)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = m.random_fields(rng); elph.update_fdm()
    fdm.set_fast_path(True); fdm.set_tuning(S, T)
    n = m.N * m.Ltau
    d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
    for _ in range(5): fdm.mul_dev(2, d_out.data_ptr(), d_in.data_ptr())
    torch.cuda.synchronize()
    os.environ["SQ_DEBUG_PRINT"] = "1"
    print(name, S, T, file=sys.stderr)
    for _ in range(3): fdm.mul_dev(2, d_out.data_ptr(), d_in.data_ptr())
    del os.environ["SQ_DEBUG_PRINT"]
