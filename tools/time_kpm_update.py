"""Cost of update_preconditioner! (tau means, device Lanczos, host Sturm bisection, hysteresis test) against the CG iterations of a
preconditioned trajectory.  argv: configs"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
for name in (sys.argv[1:] or ["cfg1", "cfg5", "cfg3", "cfg4"]):
    m = mdl.config(name)
    fdm = api.FermionDetMatrix(m, sym=True)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = bench.bench_state(m)[0] if m.name == "cfg4" else m.random_fields(np.random.default_rng(0), smooth=True)
    elph.update_fdm()
    P = api.KPMPreconditioner(fdm)
    for _ in range(3): P.update()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): P.update()
    torch.cuda.synchronize(); t_same = (time.perf_counter() - t0) / 50
    # with an operator refresh in between (tau means recomputed), as inside a trajectory
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): elph.update_fdm()
    torch.cuda.synchronize(); t_ref = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): elph.update_fdm(); P.update()
    torch.cuda.synchronize(); t_both = (time.perf_counter() - t0) / 50
    # expansion rebuild (bounds moved beyond the hysteresis window)
    b = P.update()[1]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for q in range(5): P.set_bounds(b[0] * (1 - 0.01 * q), b[1])
    torch.cuda.synchronize(); t_exp = (time.perf_counter() - t0) / 5
    pff = api.PFFCalculator(elph)
    hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=24, dt=np.pi / (2 * 24), seed=1)
    for _ in range(2): hmc.hmc_update(preconditioner=P)
    st0 = fdm.stats
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nt = 3
    for _ in range(nt): hmc.hmc_update(preconditioner=P)
    torch.cuda.synchronize(); t_traj = (time.perf_counter() - t0) / nt
    st = fdm.stats
    its = (st["cg_iterations"] - st0["cg_iterations"]) / nt
    print(f"{name}: update {t_same*1e6:.0f} us (refresh alone {t_ref*1e6:.0f}, refresh + update {t_both*1e6:.0f}), expansion rebuild {t_exp*1e3:.2f} ms; "
          f"trajectory {t_traj*1e3:.1f} ms with {its:.0f} CG iterations", flush=True)
