"""GPU parity: tau-FFT, KPM preconditioner and preconditioned CG through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc
import dense_ref as dr

pytestmark = pytest.mark.gpu


def rand_cvec(rng, m):
    return np.asfortranarray((rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2))


def relerr(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def setup(m, sym, seed=0, smooth=True):
    from smoqyelph_b200 import api
    rng = np.random.default_rng(seed)
    x = m.random_fields(rng, smooth=smooth)
    V, t = dr.build_Vt(m, x)
    ref = orc.RefFDM(m, sym=sym)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m, sym=sym)
    fdm.update(V, t)
    return rng, ref, fdm


@pytest.mark.parametrize("Lt", [1, 2, 3, 4, 5, 7, 12, 20, 35, 44, 80, 128, 200, 243, 320, 400])
def test_tau_fft_matches_numpy(Lt):
    """FourierTransformer: every radix path (2, 3, 4, 5, 7, generic 11) and the twist / normalisation."""
    from smoqyelph_b200 import api
    m = mdl.ossh_chain(10, Lt * 0.05)
    assert m.Ltau == Lt
    rng = np.random.default_rng(Lt)
    fdm = api.FermionDetMatrix(m)
    P = api.KPMPreconditioner(fdm, update=False)
    v = rand_cvec(rng, m)
    theta = np.exp(-1j * np.pi * np.arange(Lt) / Lt)[:, None]
    want = np.fft.fft(theta / np.sqrt(Lt) * v, axis=0)
    got = P.fourier(v, True)
    assert np.abs(got - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    back = P.fourier(got, False)
    assert np.abs(back - v).max() < 1e-13 * np.abs(v).max()


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "mixed"])
def test_preconditioner_matches_oracle(name, sym):
    from smoqyelph_b200 import api
    m = mdl.holstein_ssh_chain(7, 1.0) if name == "mixed" else mdl.config(name)
    rng, ref, fdm = setup(m, sym)
    Pr = orc.RefKPM(ref)
    Pg = api.KPMPreconditioner(fdm, update=False)
    start = rng.standard_normal(m.N)
    # Lanczos bounds (same start vector) and the activation decision
    Pr.update(start)
    act, bounds = Pg.update(start)
    assert act == Pr.active
    np.testing.assert_allclose(bounds, Pr.bounds, rtol=1e-9)
    # identical bounds injected => identical orders and (to rounding) coefficients
    Pg.set_bounds(*Pr.bounds)
    np.testing.assert_array_equal(Pg.orders, Pr.orders)
    for l in (0, 1, len(Pr.orders) // 2, len(Pr.orders) - 1):
        np.testing.assert_allclose(Pg.coefs(l), Pr.coefs(l), rtol=1e-10, atol=1e-13)
    v = rand_cvec(rng, m)
    assert relerr(Pg.ldiv(v), Pr.ldiv(v)) < 1e-11


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "mixed"])
def test_preconditioned_cg_matches_oracle(name, sym):
    from smoqyelph_b200 import api
    m = mdl.holstein_ssh_chain(7, 1.0) if name == "mixed" else mdl.config(name)
    rng, ref, fdm = setup(m, sym)
    Pr = orc.RefKPM(ref)
    Pg = api.KPMPreconditioner(fdm, update=False)
    Pr.update(rng.standard_normal(m.N))
    assert Pr.active
    Pg.set_bounds(*Pr.bounds)
    b = rand_cvec(rng, m)
    xr, itr, _ = ref.cg(b, P=Pr, tol=1e-14, maxiter=5000)
    xg, itg, epsg = fdm.ldiv(b, preconditioner=Pg, tol=1e-14, maxiter=5000, refresh=False)
    assert epsg < 1e-14
    assert relerr(xg, xr) < 1e-11
    for tol in (1e-5, 1e-10):
        _, itr, _ = ref.cg(b, P=Pr, tol=tol, maxiter=5000)
        _, itg, epsg = fdm.ldiv(b, preconditioner=Pg, tol=tol, maxiter=5000, refresh=False)
        assert abs(itg - itr) <= 1, (tol, itg, itr)
        _, it0, _ = fdm.ldiv(b, tol=tol, maxiter=5000)
        assert itg < it0                      # the preconditioner pays for itself on tau-smooth fields
    # warm start + preconditioner
    _, itw, _ = fdm.ldiv(b, x0=xg, preconditioner=Pg, tol=1e-10, refresh=False)
    assert itw == 0


REG_LATTICES = {
    "sq16x16": lambda: mdl.holstein_square(16, 16, 2.0),     # one warp per chain
    "sq32x16": lambda: mdl.holstein_square(32, 16, 1.0),     # two warps per chain: boundary rows through shared memory
    "sq16x32": lambda: mdl.holstein_square(16, 32, 1.0),
    "sq32x32": lambda: mdl.holstein_square(32, 32, 1.5),     # cfg4's lattice: four warps per chain
    "sq16x64": lambda: mdl.holstein_square(16, 64, 0.5),
    "sq32x64": lambda: mdl.holstein_square(32, 64, 0.5),     # eight warps per chain
    "hc8": lambda: mdl.holstein_honeycomb(8, 1.0),
    "hc16": lambda: mdl.holstein_honeycomb(16, 0.6),
    "hc24": lambda: mdl.holstein_honeycomb(24, 0.5),         # cfg5's lattice: three warps per chain
    # per-bond coefficients (SSH couplings: every bond its own tau-mean cosh / sinh): square lattices and chains
    "bssh16": lambda: mdl.bssh_square(16, 16, 1.0),          # cfg3's lattice
    "bssh32x16": lambda: mdl.bssh_square(32, 16, 0.5),
    "bssh32": lambda: mdl.bssh_square(32, 32, 0.5),
    "ossh64": lambda: mdl.ossh_chain(64, 2.0),               # cfg2's lattice: one warp per chain
    "ossh256": lambda: mdl.ossh_chain(256, 1.0),             # four warps per chain
}


@pytest.mark.parametrize("name", list(REG_LATTICES))
def test_register_chebyshev_matches_oracle_and_smem_kernel(name, monkeypatch):
    """The register-engine Chebyshev kernel (kpm_reg.cu): same orders / coefficients, P^-1 v against the oracle and against the
    shared-memory kernel it replaces on these lattices, and the preconditioned CG iteration counts."""
    from smoqyelph_b200 import api
    m = REG_LATTICES[name]()
    rng, ref, fdm = setup(m, True)
    Pr = orc.RefKPM(ref)
    Pg = api.KPMPreconditioner(fdm, update=False)
    Pr.update(rng.standard_normal(m.N))
    assert Pr.active
    Pg.set_bounds(*Pr.bounds)
    np.testing.assert_array_equal(Pg.orders, Pr.orders)
    assert Pr.orders.max() > 2
    v = rand_cvec(rng, m)
    want = Pr.ldiv(v)
    st0 = fdm.stats
    got = Pg.ldiv(v)
    assert fdm.stats["kpm_register"] == st0["kpm_register"] + 1, fdm.stats
    assert relerr(got, want) < 1e-11
    monkeypatch.setenv("SQ_KPM_REG", "0")
    smem = Pg.ldiv(v)
    monkeypatch.delenv("SQ_KPM_REG")
    assert fdm.stats["kpm_smem"] == st0["kpm_smem"] + 1
    assert relerr(got, smem) < 1e-12
    b = rand_cvec(rng, m)
    for tol in (1e-5, 1e-10):
        xr, itr, _ = ref.cg(b, P=Pr, tol=tol, maxiter=5000)
        xg, itg, epsg = fdm.ldiv(b, preconditioner=Pg, tol=tol, maxiter=5000, refresh=False)
        assert abs(itg - itr) <= 1, (name, tol, itg, itr)
        assert epsg < tol and relerr(xg, xr) < 100 * tol


def test_preconditioned_cg_iteration_count_cfg4():
    """Full size of the named configuration 4 (32 x 32, Ltau = 400, max order ~ 160): preconditioned CG iteration counts within +-1 of
    the reference recurrence (src/IterativeSolvers/ConjugateGradient.jl:169-249 with src/KPMPreconditioner.jl:355-414)."""
    from smoqyelph_b200 import api
    m = mdl.config("cfg4")
    rng = np.random.default_rng(5)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    ref = orc.RefFDM(m, sym=True, omp=True)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m, sym=True)
    fdm.update(V, t)
    Pr = orc.RefKPM(ref)
    Pg = api.KPMPreconditioner(fdm, update=False)
    Pr.update(rng.standard_normal(m.N))
    Pg.set_bounds(*Pr.bounds)
    np.testing.assert_array_equal(Pg.orders, Pr.orders)
    assert Pr.orders.max() > 100
    b = rand_cvec(rng, m)
    assert relerr(Pg.ldiv(b), Pr.ldiv(b)) < 1e-11
    for tol in (1e-5, 1e-10):
        xr, itr, _ = ref.cg(b, P=Pr, tol=tol, maxiter=5000)
        xg, itg, epsg = fdm.ldiv(b, preconditioner=Pg, tol=tol, maxiter=5000, refresh=False)
        assert abs(itg - itr) <= 1, (tol, itg, itr)
        assert epsg < tol and relerr(xg, xr) < 100 * tol
    assert fdm.stats["kpm_register"] > 0


def test_tau_independent_fields_exact_inverse_full_size():
    """Size-independent property at cfg4: tau-independent fields => P^-1 M^T M = I at high order."""
    from smoqyelph_b200 import api
    m = mdl.config("cfg4")
    rng = np.random.default_rng(3)
    x = np.repeat(0.3 * rng.standard_normal((m.Nph, 1)), m.Ltau, axis=1)
    V, t = dr.build_Vt(m, x)
    fdm = api.FermionDetMatrix(m)
    fdm.update(V, t)
    P = api.KPMPreconditioner(fdm, a1=30.0, a2=30.0)
    v = rand_cvec(rng, m)
    w = P.ldiv(fdm.mul_MtM(v))
    assert np.abs(w - v).max() / np.abs(v).max() < 1e-7
    _, it, eps = fdm.ldiv(v, preconditioner=P, tol=1e-10, refresh=False)
    assert it <= 3


def test_inactive_preconditioner_is_identity():
    from smoqyelph_b200 import api
    m = mdl.ossh_chain(8, 0.5, t=40.0)           # huge hopping => eps_max >> 2 => deactivated (KPMPreconditioner.jl:570)
    rng, ref, fdm = setup(m, True)
    Pg = api.KPMPreconditioner(fdm, update=False)
    act, _ = Pg.update(rng.standard_normal(m.N))
    assert not act
    v = rand_cvec(rng, m)
    assert np.array_equal(Pg.ldiv(v), v)


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "sq16x16", "hc8"])
def test_batched_multi_rhs_solve_matches_one_by_one(name, monkeypatch):
    """Multi-RHS preconditioned CG (cg_batch.cu): every system of the batch reproduces the one-by-one solve -- solution, iteration
    count (exactly: same recurrence per system) -- and the oracle's preconditioned CG (iterations +-1), including warm starts and
    systems that converge at different iterations."""
    from smoqyelph_b200 import api
    m = REG_LATTICES[name]() if name in REG_LATTICES else mdl.config(name)
    rng, ref, fdm = setup(m, True)
    Pr = orc.RefKPM(ref)
    Pg = api.KPMPreconditioner(fdm, update=False)
    Pr.update(rng.standard_normal(m.N))
    assert Pr.active
    Pg.set_bounds(*Pr.bounds)
    V, nrhs = m.N * m.Ltau, 5
    B = np.asfortranarray(rng.standard_normal((V, nrhs)) + 1j * rng.standard_normal((V, nrhs)))
    B[:, 1] *= 1e-3                                  # different scales / difficulty: the systems stop at different iterations
    B[:, 3] = ref.mul_MtM(B[:, 0].reshape(m.Ltau, m.N, order="F")).ravel(order="F")
    for tol in (1e-5, 1e-10):
        st0 = fdm.stats
        X, its, epss = fdm.ldiv_batch(B, preconditioner=Pg, tol=tol, maxiter=5000, refresh=False)
        assert fdm.stats["cg_batched_rhs"] == st0["cg_batched_rhs"] + nrhs
        for j in range(nrhs):
            bj = B[:, j].reshape(m.Ltau, m.N, order="F")
            x1, it1, eps1 = fdm.ldiv(bj, preconditioner=Pg, tol=tol, maxiter=5000, refresh=False)
            assert its[j] == it1, (name, tol, j, its[j], it1)
            assert relerr(X[:, j].reshape(m.Ltau, m.N, order="F"), x1) < 1e-12
            assert abs(epss[j] - eps1) <= 1e-3 * tol
            _, itr, _ = ref.cg(bj, P=Pr, tol=tol, maxiter=5000)
            assert abs(its[j] - itr) <= 1
    # warm start: from the solution nothing is left to do; from a perturbed solution a few iterations
    X2, its2, _ = fdm.ldiv_batch(B, X0=X, preconditioner=Pg, tol=1e-9, maxiter=5000, refresh=False)
    assert np.all(its2 == 0) and np.array_equal(X2, X)
    X3, its3, _ = fdm.ldiv_batch(B, X0=X * (1 + 1e-4), preconditioner=Pg, tol=1e-10, maxiter=5000, refresh=False)
    assert np.all(its3 > 0) and relerr(X3, X) < 1e-8
    # maxiter cut-off and the fall-back without a preconditioner
    _, itc, epsc = fdm.ldiv_batch(B, preconditioner=Pg, tol=1e-14, maxiter=2, refresh=False)
    assert np.all(itc == 2) and np.all(epsc > 1e-14)
    Xp, itp, _ = fdm.ldiv_batch(B[:, :2], tol=1e-8, maxiter=20000)
    for j in range(2):
        x1, it1, _ = fdm.ldiv(B[:, j].reshape(m.Ltau, m.N, order="F"), tol=1e-8, maxiter=20000)
        assert itp[j] == it1 and relerr(Xp[:, j].reshape(m.Ltau, m.N, order="F"), x1) < 1e-12


def test_greens_estimator_batched_update_matches_sequential(monkeypatch):
    """update_greens_estimator! with a preconditioner runs the Nrv solves as one batch: same G R and average iteration count as the
    one-by-one loop (SQ_NO_BATCH_CG=1), warm-start semantics included."""
    from smoqyelph_b200 import api
    m = REG_LATTICES["sq16x16"]()
    rng, ref, fdm = setup(m, True)
    P = api.KPMPreconditioner(fdm, update=False)
    act, _ = P.update(rng.standard_normal(m.N))      # fixes the bounds: later refreshes stay inside the rbuf / 2 hysteresis
    assert act
    Nrv, V = 6, m.N * m.Ltau
    R = rng.standard_normal((V, Nrv)) + 1j * rng.standard_normal((V, Nrv))
    R = np.asfortranarray(R / np.abs(R))
    res = {}
    for mode in ("batch", "seq"):
        if mode == "seq":
            monkeypatch.setenv("SQ_NO_BATCH_CG", "1")
        g = api.GreensEstimator(fdm, Nrv=Nrv, seed=1)
        st0 = fdm.stats
        a1 = g.update_greens_estimator(preconditioner=P, R=R, tol=1e-10, maxiter=5000)
        a2 = g.update_greens_estimator(preconditioner=P, R=R, tol=1e-12, maxiter=5000)      # warm start from the previous G R
        res[mode] = (a1, a2, g.get()[1], fdm.stats["cg_batched_rhs"] - st0["cg_batched_rhs"])
    assert res["batch"][3] == 2 * Nrv and res["seq"][3] == 0
    assert res["batch"][0] == res["seq"][0] and res["batch"][1] == res["seq"][1]
    assert relerr(res["batch"][2], res["seq"][2]) < 1e-12
    GRr = np.zeros((V, Nrv), np.complex128, order="F")
    Pr = orc.RefKPM(ref)
    Pr.update(rng.standard_normal(m.N))
    orc.greens_update(ref, Pr, R, GRr, 1e-12, 5000)
    assert relerr(res["batch"][2], GRr) < 1e-9
