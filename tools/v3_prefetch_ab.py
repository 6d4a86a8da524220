"""A/B of the operand staging of the stand-alone register-path matvec (SQ_V3_PRE: 0 plain loads, 2 bulk async copies into shared memory): agreement of the results and time per launch over a size sweep."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
peak, src = bench.measured_peak()
for Lx, Ly, beta in ((32, 32, 80.0), (32, 32, 160.0), (32, 32, 320.0), (32, 64, 320.0)):
    m = mdl.holstein_square(Lx, Ly, beta)
    fdm = api.FermionDetMatrix(m, sym=True)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = m.random_fields(np.random.default_rng(0), smooth=True); elph.update_fdm()
    n = m.N * m.Ltau
    d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
    B = 40 * m.N * m.Ltau + 16 * m.Nh
    row = {"lattice": f"{Lx}x{Ly}", "Ltau": m.Ltau}
    ref = None
    for pre in ("0", "2"):
        os.environ["SQ_V3_PRE"] = pre
        res = {}
        for S in (1, 2, 3, 4, 5, 7):
            fdm.set_fast_path(2 + 256 * S)
            if fdm.tuning["path"] != 3: continue
            d_out.zero_()
            res[S] = round(fdm.time_mul(102, d_out.data_ptr(), d_in.data_ptr(), 100), 2)
            torch.cuda.synchronize()
            if ref is None: ref = d_out.clone()
            err = float((d_out - ref).abs().max() / ref.abs().max())
            assert err < 1e-13, (pre, S, err)
        b = min(res.values())
        row["pre" + pre] = {"us_by_S": res, "best_frac": round(B / b / 1e3 / peak, 3)}
    print(json.dumps(row), flush=True)
    del fdm, elph
