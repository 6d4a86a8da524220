"""Statistical parity (north_star): density, double occupancy and phonon energies of a fixed-seed EFA-PFF-HMC run on the
GPU library agree with an independent fixed-seed run of the CPU oracle within their combined error bars.  The two chains
use different random streams (Philox on the device, numpy on the host), so this is a check of the sampled distribution --
forces, accept/reject, Fourier acceleration, measurements -- not of one trajectory (that is test_gpu_pff_hmc.py)."""
import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

NT, NTHERM, NMEAS, NRV, BIN = 6, 60, 240, 6, 12
TOL_A, TOL_F = 1e-8, 1e-5


def phonon_energies(m, x):
    """Per-site phonon potential energy and the imaginary-time kinetic term of the bosonic action, from the fields."""
    pe = 0.5 * np.mean((m.Omega[:, None] ** 2) * x ** 2)
    dx = np.roll(x, -1, axis=1) - x
    ke = 0.5 * np.mean(m.Mass[:, None] * dx ** 2) / m.dtau ** 2
    return pe, ke


def binned(series):
    a = np.asarray(series, float)
    nb = len(a) // BIN
    b = a[:nb * BIN].reshape(nb, BIN).mean(axis=1)
    return b.mean(), b.std(ddof=1) / np.sqrt(nb)


def run_gpu(m, x0):
    from smoqyelph_b200 import api
    fdm = api.SymFermionDetMatrix(m, tol=TOL_A, maxiter=20000)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = x0
    elph.update_fdm()
    pff = api.PFFCalculator(elph)
    hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=NT, seed=2024)
    g = api.GreensEstimator(fdm, Nrv=NRV, seed=7)
    obs, nacc = {"n": [], "d": [], "pe": [], "ke": []}, 0
    for k in range(NTHERM + NMEAS):
        acc, _ = hmc.hmc_update(tol_action=TOL_A, tol_force=TOL_F)
        if k < NTHERM:
            continue
        nacc += int(acc)
        g.update_greens_estimator(tol=TOL_A)
        ms = g.measure()
        pe, ke = phonon_energies(m, elph.x)
        obs["n"].append(ms["n"].real); obs["d"].append(ms["double_occ"].real); obs["pe"].append(pe); obs["ke"].append(ke)
    return obs, nacc / NMEAS


def run_oracle(m, x0):
    rng = np.random.default_rng(99)
    rf = orc.RefFDM(m, sym=True, tol=TOL_A, maxiter=20000)
    re = orc.RefElPh(m)
    re.set_x(x0)
    re.refresh(rf)
    rp = orc.RefPFF(re, rf)
    ra = orc.RefEFA(re)
    V = m.N * m.Ltau
    obs, nacc = {"n": [], "d": [], "pe": [], "ke": []}, 0
    nr = orc.hmc_random_count(m, NT, False)
    for k in range(NTHERM + NMEAS):
        rnd = rng.standard_normal(nr)
        rnd[0], rnd[-1] = rng.random(), rng.random()            # the two uniform draws (time-step jitter, accept test)
        acc, _ = orc.hmc_update(re, rf, rp, ra, None, NT, np.pi / (2 * NT), 0.05, TOL_A, TOL_F, 20000, rnd)
        if k < NTHERM:
            continue
        nacc += int(acc)
        R = rng.standard_normal((V, NRV)) + 1j * rng.standard_normal((V, NRV))
        R = np.asfortranarray(R / np.abs(R))
        GR = np.zeros((V, NRV), np.complex128, order="F")
        orc.greens_update(rf, None, R, GR, TOL_A, 20000)
        pe, ke = phonon_energies(m, np.array(re.x))
        obs["n"].append(orc.measure("n", R, GR).real); obs["d"].append(orc.measure("double_occ", R, GR).real)
        obs["pe"].append(pe); obs["ke"].append(ke)
    return obs, nacc / NMEAS


def test_observables_agree_within_error_bars():
    m = mdl.holstein_honeycomb(2, 1.0, mu=0.3, alpha=1.0)        # away from half filling so that <n> is not trivially 1
    x0 = m.random_fields(np.random.default_rng(5), amplitude=0.3)
    og, acc_g = run_gpu(m, x0)
    oo, acc_o = run_oracle(m, x0)
    assert acc_g > 0.5 and acc_o > 0.5, (acc_g, acc_o)
    assert abs(acc_g - acc_o) < 0.15
    report = {}
    for key in ("n", "d", "pe", "ke"):
        mg, eg = binned(og[key])
        mo, eo = binned(oo[key])
        sigma = np.hypot(eg, eo)
        report[key] = (mg, eg, mo, eo)
        assert abs(mg - mo) < 4.0 * sigma, (key, report[key])
        assert sigma < 0.1 * max(abs(mo), 0.05), (key, "error bar too large for a meaningful comparison", report[key])
    # sanity: away from half filling, attractive (Holstein) interaction => double occupancy above the uncorrelated value n^2
    n, d = report["n"][0], report["d"][0]
    assert abs(n - 0.5) > 0.02 and d > n * n


def test_tutorial_loop_runs_end_to_end():
    """The reference's own tests are `isnothing(run_simulation(...))` smoke runs of the tutorials: the same loop --
    reflection, swap, HMC, estimator solves, scalar and correlation measurements -- on the library, with sanity bounds."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "holstein_honeycomb.py")
    spec = importlib.util.spec_from_file_location("holstein_honeycomb_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    m, obs, meta = mod.run_simulation(L=3, beta=1.0, N_therm=4, N_measurements=4, Nt=4, Nrv=4, tol=1e-8)
    assert 0.5 < obs["density"] < 1.5                      # half filling up to noise (mu = 0, particle-hole symmetric coupling)
    assert 0.0 < obs["double_occ"] < 1.0
    assert meta["hmc_acceptance_rate"] > 0.25 and meta["hmc_iters"] > 0 and meta["measurement_iters"] > 0
    G = obs["greens"]
    assert G.shape == tuple(m.lattice_dims) + (m.Ltau + 1,)
    assert abs(G[0, 0, 0].real + G[0, 0, m.Ltau].real - 1.0) < 1e-10      # G(0, beta) = 1 - G(0, 0)
    assert np.all(np.isfinite(obs["density_corr"])) and np.all(np.isfinite(obs["pair_corr"])) and np.all(np.isfinite(obs["spin_z_corr"]))
