import sys, os
os.environ["SQ_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
m = mdl.config("cfg4")
rng = np.random.default_rng(0)
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = m.random_fields(rng); elph.update_fdm()
n = m.N * m.Ltau
d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
for op in (2, 102):
    for S in (1, 3, 5):
        fdm.set_fast_path(2 + 256 * S)
        fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 5)
        os.environ["SQ_DEBUG_PRINT"] = "1"
        print("op", op, "S", S, fdm.tuning, file=sys.stderr)
        fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 1)
        del os.environ["SQ_DEBUG_PRINT"]
