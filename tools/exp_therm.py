"""Which synthetic phonon state gives stable, representative trajectories at cfg4?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smoqyelph_b200 import model as mdl, api
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
amp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
ntraj = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m = mdl.config(name)
rng = np.random.default_rng(0)
Lx = m.lattice_dims[0]
sites = np.arange(m.N)
stag = np.where(((sites % Lx) + (sites // Lx)) % 2 == 0, 1.0, -1.0)
x = np.asfortranarray(amp * stag[:, None] + 0.3 * mdl.thermal_fields(m, rng))
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = x; elph.update_fdm()
pff = api.PFFCalculator(elph)
hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=24, seed=1)
b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
_, it, eps = fdm.ldiv(b, tol=1e-5, maxiter=50000)
print("start: cg iters @1e-5", it)
for k in range(ntraj):
    t0 = time.perf_counter()
    acc, its = hmc.hmc_update(tol_action=1e-10, tol_force=1e-5, maxiter=20000)
    dt = time.perf_counter() - t0
    xx = elph.x
    print(f"traj {k}: acc {acc} iters_avg {its:.0f} dH {hmc.info[1]:.3f} Sf {hmc.info[2]:.1f}->{hmc.info[3]:.1f} Sb {hmc.info[4]:.1f}->{hmc.info[5]:.1f} wall {dt:.2f}s  <stag x> {np.mean(xx.mean(axis=1)*stag):.3f} rms {xx.std():.3f}")
