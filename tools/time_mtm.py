"""Device timing of the fused M^T M v kernel across configs and tunings (CUDA events on the library stream)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smoqyelph_b200 import model as mdl, api
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dense_ref as dr


def time_op(fdm, op, d_out, d_in, reps=50, flush=None):
    st = torch.cuda.ExternalStream(fdm.stream)
    with torch.cuda.stream(st):
        for _ in range(5):
            fdm.mul_dev(op, d_out.data_ptr(), d_in.data_ptr())
        ts = []
        for _ in range(reps):
            if flush is not None:
                flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            fdm.mul_dev(op, d_out.data_ptr(), d_in.data_ptr())
            e1.record(st)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts = np.array(ts)
    return float(np.median(ts)), float(ts.min())


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")   # 256 MB > L2
    for name in names:
        m = mdl.config(name)
        rng = np.random.default_rng(0)
        V, t = dr.build_Vt(m, m.random_fields(rng))
        fdm = api.FermionDetMatrix(m, sym=True)
        fdm.update(V, t)
        n = m.N * m.Ltau
        d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda")
        d_out = torch.zeros_like(d_in)
        B = (40 * m.N + 16 * m.Nh) * m.Ltau
        auto = fdm.tuning
        res = {"config": name, "N": m.N, "Nh": m.Nh, "Ltau": m.Ltau, "bytes": B, "auto": auto}
        med, mn = time_op(fdm, 2, d_out, d_in)
        res["auto_us_hot"] = med
        res["auto_GBs_hot"] = B / med / 1e3
        med, mn = time_op(fdm, 2, d_out, d_in, flush=flush)
        res["auto_us_cold"] = med
        res["auto_GBs_cold"] = B / med / 1e3
        sweep = {}
        if auto["path"] != 1:
            for fast in (0, 1):
                fdm.set_fast_path(fast)
                for S in (1, 2, 3, 4, 5, 6, 8):
                    for T in (128, 256, 512, 576, 1024):
                        try:
                            fdm.set_tuning(S, T)
                        except Exception:
                            continue
                        med, mn = time_op(fdm, 2, d_out, d_in, reps=20)
                        sweep[f"v{fast+1}_S{S}_T{T}"] = round(med, 2)
            fdm.set_fast_path(auto["path"] == 2)
            fdm.set_tuning(auto["slab"], auto["threads"])
        res["sweep_us_hot"] = sweep
        for op, nm in ((0, "M"), (1, "Mt")):
            med, mn = time_op(fdm, op, d_out, d_in, reps=20)
            res[f"{nm}_us_hot"] = med
        print(json.dumps(res))


if __name__ == "__main__":
    main()
