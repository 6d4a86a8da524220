"""Register-path (fdm_v3.cu) check: parity vs the oracle and the shared-memory kernels, CG iteration parity, timing."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from smoqyelph_b200 import model as mdl, api
from oracle import oracle as orc
import dense_ref as dr
from time_mtm import time_op


def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    for name, mk in (("h16x16", lambda: mdl.holstein_square(16, 16, 2.0)), ("h32x16", lambda: mdl.holstein_square(32, 16, 1.0)),
                     ("h16x32", lambda: mdl.holstein_square(16, 32, 1.0)), ("cfg4", lambda: mdl.config("cfg4"))):
        m = mk()
        rng = np.random.default_rng(0)
        V, t = dr.build_Vt(m, m.random_fields(rng))
        ref = orc.RefFDM(m, sym=True); ref.update(V, t)
        fdm = api.FermionDetMatrix(m, sym=True); fdm.update(V, t)
        v = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
        res = {"config": name, "auto": fdm.tuning}
        fdm.set_fast_path(1)
        base = {op: getattr(fdm, op)(v) for op in ("mul_M", "mul_Mt", "mul_MtM", "mul_MMt")}
        for S in (1, 2, 3, 4, 5, 7):
            fdm.set_fast_path(2 + 256 * S)
            assert fdm.tuning["path"] == 3, fdm.tuning
            for op in base:
                got = getattr(fdm, op)(v)
                e = rel(got, getattr(ref, op)(v))
                assert e < 1e-12, (name, S, op, e)
                assert np.array_equal(got, base[op]), (name, S, op, "not bit-identical to v2")
        res["parity"] = "ok (bit-identical to fdm_v2, <1e-12 vs oracle)"
        # CG
        b = v / np.linalg.norm(v)
        fdm.set_fast_path(1)
        x2, it2, e2 = fdm.ldiv(b, tol=1e-8, maxiter=5000)
        fdm.set_fast_path(2 + 256 * 3)
        x3, it3, e3 = fdm.ldiv(b, tol=1e-8, maxiter=5000)
        res["cg"] = {"iters_v2": it2, "iters_v3": it3, "rel": rel(x3, x2)}
        assert abs(it2 - it3) <= 1 and rel(x3, x2) < 1e-7
        n = m.N * m.Ltau
        d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
        B = (40 * m.N) * m.Ltau + 16 * m.Nh
        tm = {}
        fdm.set_fast_path(1)
        tm["v2_hot"] = time_op(fdm, 2, d_out, d_in)[0]; tm["v2_cold"] = time_op(fdm, 2, d_out, d_in, flush=flush)[0]
        for S in (1, 2, 3, 4, 5, 7):
            fdm.set_fast_path(2 + 256 * S)
            tm[f"v3_S{S}_hot"] = time_op(fdm, 2, d_out, d_in)[0]
            tm[f"v3_S{S}_cold"] = time_op(fdm, 2, d_out, d_in, flush=flush)[0]
            tm[f"v3_S{S}_M_hot"] = time_op(fdm, 0, d_out, d_in)[0]
        res["us"] = {k: round(x, 2) for k, x in tm.items()}
        res["bytes_uniform"] = B
        # CG iteration time
        fdm.set_fast_path(2 + 256 * 3)
        t0 = time.perf_counter(); _, it, _ = fdm.ldiv(b, tol=1e-30, maxiter=2000); t1 = time.perf_counter()
        res["cg_us_per_iter_v3"] = round((t1 - t0) / it * 1e6, 2)
        fdm.set_fast_path(1)
        t0 = time.perf_counter(); _, it, _ = fdm.ldiv(b, tol=1e-30, maxiter=2000); t1 = time.perf_counter()
        res["cg_us_per_iter_v2"] = round((t1 - t0) / it * 1e6, 2)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
