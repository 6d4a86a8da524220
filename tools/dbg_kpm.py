import sys, os
os.environ["SQ_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config("cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm); elph.x = bench.cdw_start(m, 0); elph.update_fdm()
P = api.KPMPreconditioner(fdm)
n = m.N * m.Ltau
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
for _ in range(3): P.ldiv_dev(x.data_ptr(), b.data_ptr())
torch.cuda.synchronize()
os.environ["SQ_DEBUG_PRINT"] = "1"
for _ in range(3): P.ldiv_dev(x.data_ptr(), b.data_ptr())
