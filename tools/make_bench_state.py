"""Relaxed phonon state for bench.py: the staggered CDW start of cfg4 evolved by NTRAJ EFA-PFF-HMC trajectories on the GPU (Philox seed 77,
KPM-preconditioned solves), rounded to float32 so that the committed file is small; both bench arms load the SAME numbers from it.
Usage (on the GPU box): python tools/make_bench_state.py [ntraj]  -> gpurun_out/bench_state_cfg4_f32.npy (copy to bench_data/)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smoqyelph_b200  # noqa: F401,E402
from smoqyelph_b200 import api, model as mdl  # noqa: E402
import bench  # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = mdl.config("cfg4")
fdm = api.SymFermionDetMatrix(m, tol=1e-10, maxiter=10000)
elph = api.ElectronPhononParameters(m, fdm)
pff = api.PFFCalculator(elph)
P = api.KPMPreconditioner(fdm, update=False)
elph.x = bench.cdw_start(m, 1000)
elph.update_fdm()
hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=bench.NT, seed=77)
for k in range(ntraj):
    acc, it = hmc.hmc_update(preconditioner=P, tol_action=bench.TOL_ACTION, tol_force=bench.TOL_FORCE, maxiter=bench.MAXITER)
    print(f"trajectory {k}: accepted={acc} cg iterations per solve={it:.1f} dH={hmc.info[1]:.3f}", flush=True)
x = np.asarray(elph.x, np.float64).astype(np.float32)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/bench_state_cfg4_f32.npy", x)
print("saved", x.shape, "mean |x|", float(np.abs(x).mean()))
