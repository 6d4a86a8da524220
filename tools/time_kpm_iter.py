import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
for name in sys.argv[1:]:
    m = mdl.config(name)
    fdm = api.FermionDetMatrix(m, sym=True)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = bench.bench_state(m)[0] if m.name == "cfg4" else m.random_fields(np.random.default_rng(0), smooth=True)
    elph.update_fdm()
    P = api.KPMPreconditioner(fdm)
    n = m.N * m.Ltau
    b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
    st = torch.cuda.ExternalStream(fdm.stream)
    for mode in ("fused", "unfused", "fused", "unfused"):
        os.environ.pop("SQ_NO_FFT_FUSION", None)
        if mode == "unfused": os.environ["SQ_NO_FFT_FUSION"] = "1"
        fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20): fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=12)
        e1.record(st); e1.synchronize()
        print(f"{name} {mode}: {e0.elapsed_time(e1)/240*1e3:.1f} us per preconditioned iteration")
