"""Minimal driver for ncu: a few fused M^T M v launches on a named config."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import dense_ref as dr
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
m = mdl.config(name)
rng = np.random.default_rng(0)
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = m.random_fields(rng); elph.update_fdm()
if len(sys.argv) > 3: fdm.set_tuning(int(sys.argv[2]), int(sys.argv[3]))
n = m.N * m.Ltau
d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
for _ in range(10): fdm.mul_dev(2, d_out.data_ptr(), d_in.data_ptr())
torch.cuda.synchronize()
print("ok", fdm.tuning)
