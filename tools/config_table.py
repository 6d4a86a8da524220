"""One line per named configuration: fused M^T M v time and GB/s against the algorithmic bytes, CG microseconds per iteration
(plain and KPM-preconditioned), EFA-HMC trajectories/s, and which kernel family ran.  JSON lines on stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench

names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
peak, _ = bench.measured_peak() if hasattr(bench, "measured_peak") else (6542.0, "fallback")
for name in names:
    m = mdl.config(name)
    fdm = api.FermionDetMatrix(m, sym=True)
    elph = api.ElectronPhononParameters(m, fdm)
    rng = np.random.default_rng(0)
    elph.x = bench.bench_state(m)[0] if m.name == "cfg4" else bench.cdw_start(m, 0) if (m.Nhol and len(m.lattice_dims) == 2 and m.N == m.lattice_dims[0] * m.lattice_dims[1]) else m.random_fields(rng, smooth=True)
    elph.update_fdm()
    n = m.N * m.Ltau
    b = torch.randn(n, 2, dtype=torch.float64, device="cuda")
    x = torch.zeros_like(b)
    B_generic = (40 * m.N + 16 * m.Nh) * m.Ltau
    row = {"config": name, "N": m.N, "Nh": m.Nh, "Ltau": m.Ltau, "bytes_generic": B_generic}
    row["matvec_us"] = fdm.time_mul(2, x.data_ptr(), b.data_ptr(), 200)
    row["tuning"] = fdm.tuning
    row["matvec_GBps_generic_bytes"] = B_generic / row["matvec_us"] / 1e3
    row["frac_of_peak"] = row["matvec_GBps_generic_bytes"] / peak
    for label, use_kpm in (("cg_us_per_iter", False), ("cg_kpm_us_per_iter", True)):
        P = api.KPMPreconditioner(fdm, update=False) if use_kpm else None
        if P is not None:
            row["kpm_active"] = P.update()[0]
        # solves to the action tolerance; per-iteration time = device time of three solves / their iterations
        fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-10, maxiter=20000)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tot = 0
        for _ in range(3):
            it, _ = fdm.cg_dev(x.data_ptr(), b.data_ptr(), True, preconditioner=P, tol=1e-10, maxiter=20000)
            tot += it
        torch.cuda.synchronize()
        row[label] = (time.perf_counter() - t0) / tot * 1e6
        row[label.replace("us_per_iter", "iters_to_1e-10")] = it
        if P is not None:
            P.close()
    pff = api.PFFCalculator(elph)
    x_start = elph.x.copy()
    for label, use_kpm in (("trajectories_per_s", False), ("trajectories_per_s_kpm", True)):
        elph.x = x_start
        elph.update_fdm()
        P = api.KPMPreconditioner(fdm, update=False) if use_kpm else None
        hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=bench.NT, seed=5)
        kw = dict(preconditioner=P, tol_action=bench.TOL_ACTION, tol_force=bench.TOL_FORCE, maxiter=bench.MAXITER)
        for _ in range(2):                                 # two untimed trajectories move the fields away from the smooth start
            hmc.hmc_update(**kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ntr = 3
        its, accs = [], 0
        for _ in range(ntr):
            acc, iters = hmc.hmc_update(**kw)
            its.append(iters)
            accs += int(acc)
        torch.cuda.synchronize()
        row[label] = ntr / (time.perf_counter() - t0)
        row[label.replace("trajectories_per_s", "avg_cg_iters")] = float(np.mean(its))
        row[label.replace("trajectories_per_s", "accepted")] = accs
        hmc.close()
        if P is not None:
            P.close()
    print(json.dumps(row), flush=True)
    pff.close(); elph.close(); fdm.close()
