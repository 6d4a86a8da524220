import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from smoqyelph_b200 import model as mdl, api, lib
m = mdl.config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
P = api.KPMPreconditioner(fdm, update=False)
L = lib.load()
n = m.N * m.Ltau
st = torch.cuda.ExternalStream(fdm.stream)
v = np.asfortranarray(np.random.default_rng(0).standard_normal((m.Ltau, m.N)) + 0j)
import time
# time through the host API is dominated by copies; use inactive preconditioner ldiv? use fourier on device via repeated host calls is useless -> time ldiv_dev with bounds forcing order 1 everywhere (pure FFT pair)
P.set_bounds(0.999, 1.001)
print("max order", P.orders.max())
b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
with torch.cuda.stream(st):
    for _ in range(3): P.ldiv_dev(x.data_ptr(), b.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): P.ldiv_dev(x.data_ptr(), b.data_ptr())
    e1.record(st); e1.synchronize()
print(f"SB={os.environ.get('SQ_FFT_SB')} T={os.environ.get('SQ_FFT_T')}: forward+inverse FFT pair {e0.elapsed_time(e1)/50*1e3:.1f} us")
