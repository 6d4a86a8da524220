"""Minimal driver for ncu: a few fused M^T M v launches on a named config.  argv: config [op] (op 102 = register path, native order)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
op = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = mdl.holstein_square(*[int(q) for q in name.split(":")[1:3]], float(name.split(":")[3])) if name.startswith("sq:") else mdl.config(name)   # sq:Lx:Ly:beta
rng = np.random.default_rng(0)
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = m.random_fields(rng); elph.update_fdm()
n = m.N * m.Ltau
d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
us = fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 10)
torch.cuda.synchronize()
print("ok", fdm.tuning, us)
