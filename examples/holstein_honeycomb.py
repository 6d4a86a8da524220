"""The simulation loop of /root/reference/tutorials/holstein_honeycomb.jl:540-700 on the B200 library (Python twin of the
Julia shim): thermalisation and measurement sweeps of reflection, swap and EFA-PFF-HMC updates, with the estimator solves,
the scalar measurements, the time-displaced Green's function and the density / pair / spin / bond / current correlations on the device.
The model DSL, binning and IO of SmoQyDQMC are replaced by a dictionary of running averages.

    python examples/holstein_honeycomb.py [L] [beta] [N_therm] [N_measurements]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import smoqyelph_b200  # noqa: F401
from smoqyelph_b200 import api, model as mdl


def run_simulation(L=3, beta=4.0, N_therm=20, N_measurements=20, Omega=1.0, alpha=1.5, mu=0.0, Nt=8, Nrv=10, tol=1e-10, seed=0,
                   use_preconditioner=False, device=0):
    rng = np.random.default_rng(seed)
    m = mdl.holstein_honeycomb(L, beta, Omega=Omega, alpha=alpha, mu=mu)
    fdm = api.SymFermionDetMatrix(m, tol=tol, maxiter=10000, device=device)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = mdl.thermal_fields(m, rng)
    elph.update_fdm()
    from smoqyelph_b200.parallel import chain_seed
    # one Philox key per component, hashed from the chain seed (consecutive integers would alias streams between chains)
    pff = api.PFFCalculator(elph, seed=chain_seed(seed, 0, 3))
    P = api.KPMPreconditioner(fdm, seed=chain_seed(seed, 0, 4)) if use_preconditioner else None
    hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=Nt, seed=chain_seed(seed, 0, 1))
    g = api.GreensEstimator(fdm, Nrv=Nrv, seed=chain_seed(seed, 0, 2))
    meta = {k: 0.0 for k in ("hmc_acceptance_rate", "reflection_acceptance_rate", "swap_acceptance_rate", "hmc_iters",
                             "reflection_iters", "swap_iters", "measurement_iters")}

    def sweep():
        acc, it = api.reflection_update(elph, pff, rng=rng, preconditioner=P, tol=tol)
        meta["reflection_acceptance_rate"] += acc; meta["reflection_iters"] += it
        acc, it = api.swap_update(elph, pff, rng=rng, preconditioner=P, tol=tol)
        meta["swap_acceptance_rate"] += acc; meta["swap_iters"] += it
        acc, it = hmc.hmc_update(preconditioner=P, tol_action=tol, tol_force=np.sqrt(tol))
        meta["hmc_acceptance_rate"] += acc; meta["hmc_iters"] += it

    t0 = time.perf_counter()
    for _ in range(N_therm):
        sweep()
    norb = m.nphonon
    obs = {"density": 0.0, "double_occ": 0.0, "greens": 0.0, "density_corr": 0.0, "pair_corr": 0.0, "spin_z_corr": 0.0,
           "bond_corr": 0.0, "current_corr": 0.0}
    zero = (0,) * len(m.lattice_dims)
    nn_bond = ((0, norb - 1), zero)                          # the A-B bond inside the unit cell
    t_nn = np.ones((m.Ltau,) + tuple(m.lattice_dims))        # its effective hopping (Holstein: the bare t = 1 on every slice)
    for _ in range(N_measurements):
        sweep()
        meta["measurement_iters"] += g.update_greens_estimator(preconditioner=P, tol=tol)
        s = g.measure()
        obs["density"] += 2 * s["n"].real / N_measurements                 # both spin species
        obs["double_occ"] += s["double_occ"].real / N_measurements
        obs["greens"] += g.measure_GD0((0, 0)) / N_measurements             # G_AA(r, tau)
        obs["density_corr"] += sum(g.measure_density_correlation(a, b) for a in range(norb) for b in range(norb)) / N_measurements
        obs["pair_corr"] += g.measure_pair_correlation(((0, 0), zero), ((0, 0), zero)) / N_measurements     # on-site s-wave, orbital A
        obs["spin_z_corr"] += g.measure_spin_correlation(0, 0) / N_measurements
        obs["bond_corr"] += g.measure_bond_correlation(nn_bond, nn_bond) / N_measurements
        obs["current_corr"] += g.measure_current_correlation(nn_bond, nn_bond, t_nn, t_nn) / N_measurements
    n_sweeps = N_therm + N_measurements
    for k in ("hmc_acceptance_rate", "reflection_acceptance_rate", "swap_acceptance_rate", "hmc_iters", "reflection_iters", "swap_iters"):
        meta[k] /= n_sweeps
    meta["measurement_iters"] /= max(N_measurements, 1)
    meta["runtime_s"] = time.perf_counter() - t0
    meta["tuning"] = fdm.tuning
    return m, obs, meta


if __name__ == "__main__":
    a = sys.argv[1:]
    L, beta = (int(a[0]) if a else 3), (float(a[1]) if len(a) > 1 else 4.0)
    nt, nm = (int(a[2]) if len(a) > 2 else 20), (int(a[3]) if len(a) > 3 else 20)
    m, obs, meta = run_simulation(L, beta, nt, nm)
    G = obs["greens"]
    cdw = obs["density_corr"]
    print(json.dumps({"model": m.name, "N": m.N, "Ltau": m.Ltau, "density": obs["density"], "double_occ": obs["double_occ"],
                      "G_AA(r=0, tau=0)": G[(0,) * (G.ndim - 1) + (0,)].real, "G_AA(r=0, tau=beta/2)": G[(0,) * (G.ndim - 1) + (m.Ltau // 2,)].real,
                      "equal-time density correlation at r=0": cdw[(0,) * (cdw.ndim - 1) + (0,)].real, **meta}))
