"""Known-answer tests that pin the CPU oracle (SURVEY.md 8c lists them).

The reference ships no golden vectors for the hot path, so the oracle is pinned against an
independent dense-matrix statement of the definitions (tests/dense_ref.py), inverse identities,
numpy's FFT and finite differences.  CPU only.
"""
import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc
import dense_ref as dr


def small_models():
    return [
        mdl.holstein_honeycomb(2, 0.3),
        mdl.ossh_chain(6, 0.4),
        mdl.bssh_square(2, 4, 0.25),
        mdl.holstein_ssh_chain(5, 0.3),
        mdl.holstein_square(4, 2, 0.2, ph_sym=False),
    ]


def rand_cvec(rng, m):
    return np.asfortranarray((rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2))


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("mi", range(5))
def test_dense_M_Mt(mi, sym):
    m = small_models()[mi]
    rng = np.random.default_rng(mi)
    x = m.random_fields(rng)
    V, t = dr.build_Vt(m, x)
    f = orc.RefFDM(m, sym=sym)
    e = orc.RefElPh(m)
    e.set_x(x)
    e.build_Vt()
    np.testing.assert_allclose(e.V, V, rtol=0, atol=1e-15)
    np.testing.assert_allclose(e.t, t, rtol=0, atol=1e-15)
    f.update(V, t)
    M = dr.dense_M(m, dr.propagators(m, V, t, sym))
    v = rand_cvec(rng, m)
    for name, D in [("mul_M", M), ("mul_Mt", M.T), ("mul_MtM", M.T @ M), ("mul_MMt", M @ M.T)]:
        got = dr.flat(getattr(f, name)(v))
        want = D @ dr.flat(v)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13, name
    w = np.linalg.eigvalsh(M.T @ M)
    assert w.min() > 0


@pytest.mark.parametrize("mi", range(5))
def test_checkerboard_inverse_identities(mi):
    m = small_models()[mi]
    rng = np.random.default_rng(10 + mi)
    f = orc.RefFDM(m, sym=True)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    f.update(V, t)
    v = rand_cvec(rng, m)
    for tr in (False, True):
        w = f.chk(f.chk(v, transposed=tr), transposed=tr, inverse=True)
        assert np.abs(w - v).max() < 1e-13
        for c in range(len(m.colors)):
            w = f.chk(f.chk(v, transposed=tr, color=c), transposed=tr, inverse=True, color=c)
            assert np.abs(w - v).max() < 1e-13
    # transposed sweep is the matrix transpose of the forward sweep (real hoppings)
    a, b = rand_cvec(rng, m).real, rand_cvec(rng, m).real
    lhs = np.vdot(a, f.chk(b, transposed=False).real)
    rhs = np.vdot(f.chk(a, transposed=True).real, b)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


@pytest.mark.parametrize("mi", range(5))
def test_lambda_identities(mi):
    m = small_models()[mi]
    rng = np.random.default_rng(20 + mi)
    x = m.random_fields(rng)
    e = orc.RefElPh(m)
    e.set_x(x)
    Lam = e.Lambda()
    lam, Ld = dr.dense_Lambda(m, x)
    np.testing.assert_allclose(Lam, lam, rtol=1e-15)
    v = rand_cvec(rng, m)
    ops = {"mul": Ld, "mulT": Ld.T, "ldiv": np.linalg.inv(Ld), "ldivT": np.linalg.inv(Ld).T}
    for k, D in ops.items():
        got = dr.flat(e.lam_op(k, Lam, v))
        want = D @ dr.flat(v)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-14, k


@pytest.mark.parametrize("sym", [True, False])
def test_cg_against_dense_solve(sym):
    m = mdl.holstein_ssh_chain(6, 0.4)
    rng = np.random.default_rng(3)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    f = orc.RefFDM(m, sym=sym, tol=1e-14, maxiter=2000)
    f.update(V, t)
    M = dr.dense_M(m, dr.propagators(m, V, t, sym))
    b = rand_cvec(rng, m)
    want = np.linalg.solve(M.T @ M, dr.flat(b))
    x, it, eps = f.cg(b)
    assert eps < 1e-14 and it < 2000
    assert np.linalg.norm(dr.flat(x) - want) / np.linalg.norm(want) < 1e-11
    # warm start from the solution converges immediately
    x2, it2, _ = f.cg(b, x0=x, tol=1e-12)
    assert it2 == 0


def test_fourier_matches_numpy_fft():
    for Lt, n in [(20, 3), (12, 2), (35, 2), (16, 1)]:
        m = mdl.ossh_chain(4, Lt * 0.05)
        assert m.Ltau == Lt
        rng = np.random.default_rng(Lt)
        f = orc.RefFDM(m)
        k = orc.RefKPM(f)
        v = rand_cvec(rng, m)
        theta = np.exp(-1j * np.pi * np.arange(Lt) / Lt)[:, None]
        want = np.fft.fft(theta / np.sqrt(Lt) * v, axis=0)
        got = k.fourier(v, True)
        assert np.abs(got - want).max() < 1e-13
        back = k.fourier(got, False)
        assert np.abs(back - v).max() < 1e-13


def test_kpm_pieces_against_numpy():
    m = mdl.holstein_square(4, 4, 1.0)
    rng = np.random.default_rng(5)
    x = m.random_fields(rng, smooth=True)
    V, t = dr.build_Vt(m, x)
    f = orc.RefFDM(m, sym=True)
    f.update(V, t)
    k = orc.RefKPM(f)
    k.refresh_Bbar()
    # B-bar from tau means of exp(-dtau V), cosh, sinh
    cbar = np.cosh(m.dtau / 2 * np.abs(t[m.perm])).mean(axis=1)
    sbar = (np.sign(t[m.perm]) * np.sinh(m.dtau / 2 * np.abs(t[m.perm]))).mean(axis=1)
    G = dr.gamma(m, cbar, sbar)
    Bbar = G @ np.diag(np.exp(-m.dtau * V).mean(axis=1)) @ G.T
    v = rng.standard_normal(m.N) + 1j * rng.standard_normal(m.N)
    assert np.abs(k.bbar_mul(v) - Bbar @ v).max() < 1e-14
    # Lanczos (n = 20 > N = 16 would break down; use the real spectrum as reference for n<=N)
    k2 = orc.RefKPM(f, n=12)
    k2.refresh_Bbar()
    b = k2.lanczos(rng.standard_normal(m.N))
    w = np.linalg.eigvalsh(Bbar)
    assert abs(b[0] - w.min()) < 2e-2 and abs(b[1] - w.max()) < 2e-2
    assert b[0] >= w.min() - 1e-12 and b[1] <= w.max() + 1e-12
    # coefficients: Chebyshev interpolation of f on the bounds reproduces f at the nodes
    k.set_bounds(0.8 * w.min(), 1.2 * w.max())
    emin, emax = k.bounds
    orders = k.orders
    Lt = m.Ltau
    for l in (0, 1, len(orders) - 1):
        phi = 2 * np.pi / Lt * (l + 0.5)
        want_order = max(1, int(np.floor((emax - emin) * (2.0 / min(phi, 2 * np.pi - phi) + 1.0))))
        assert orders[l] == want_order
        c = k.coefs(l).real
        xs = np.cos(np.pi * (np.arange(orders[l]) + 0.5) / orders[l])          # nodes of the order-n interpolant
        bs = 0.5 * (emax - emin) * xs + 0.5 * (emax + emin)
        fx = 1.0 / (bs**2 - 2 * bs * np.cos(phi) + 1)
        approx = np.polynomial.chebyshev.chebval(xs, c)
        if orders[l] > 8:
            assert np.abs(approx - fx).max() / np.abs(fx).max() < 0.2
        # exact formula check against an independent evaluation of the quadrature
        Nq = 2 * orders[l]
        xq = np.cos(np.pi * (np.arange(Nq) + 0.5) / Nq)
        fq = 1.0 / ((0.5 * (emax - emin) * xq + 0.5 * (emax + emin)) ** 2 - 2 * (0.5 * (emax - emin) * xq + 0.5 * (emax + emin)) * np.cos(phi) + 1)
        cw = np.array([(1 if q == 0 else 2) / Nq * np.sum(fq * np.cos(np.pi * q * (np.arange(Nq) + 0.5) / Nq)) for q in range(orders[l])])
        np.testing.assert_allclose(c, cw, rtol=1e-12, atol=1e-14)


def test_tau_independent_preconditioner_is_exact_inverse():
    """KAT (4): tau-independent fields => P^-1 M^T M -> I as the expansion order grows, and
    preconditioned CG converges in O(1) iterations; validates theta / phi_n / fold conventions."""
    m = mdl.holstein_square(4, 4, 1.0)
    rng = np.random.default_rng(7)
    x = np.repeat(0.3 * rng.standard_normal((m.Nph, 1)), m.Ltau, axis=1)
    V, t = dr.build_Vt(m, x)
    f = orc.RefFDM(m, sym=True, tol=1e-12, maxiter=500)
    f.update(V, t)
    k = orc.RefKPM(f, a1=40.0, a2=40.0)
    k.update(rng.standard_normal(m.N))
    assert k.active
    v = rand_cvec(rng, m)
    w = k.ldiv(f.mul_MtM(v))
    assert np.abs(w - v).max() / np.abs(v).max() < 1e-8
    _, it_p, _ = f.cg(v, P=k)
    _, it_0, _ = f.cg(v)
    assert it_p <= 3 and it_0 > 10
    # the default (low order) preconditioner still cuts the iteration count on smooth fields
    x = m.random_fields(rng, smooth=True)
    V, t = dr.build_Vt(m, x)
    f.update(V, t)
    k2 = orc.RefKPM(f)
    k2.update(rng.standard_normal(m.N))
    xs, it_p, _ = f.cg(v, P=k2)
    xu, it_0, _ = f.cg(v)
    assert it_p < it_0
    assert np.abs(xs - xu).max() / np.abs(xu).max() < 1e-9


def _force_fd(m, sym, exact, rng, h=1e-5, nprobe=6):
    x = m.random_fields(rng)
    f = orc.RefFDM(m, sym=sym, tol=1e-14, maxiter=5000)
    e = orc.RefElPh(m)
    e.set_x(x)
    e.refresh(f)
    pff = orc.RefPFF(e, f, exact_holstein=exact)
    pff.sample(rand_cvec(rng, m))
    Phi = pff.Phi.copy()
    F, Sf, it, eps = pff.force(tol=1e-14, maxiter=5000)
    assert abs(Sf - dr.fermionic_action_dense(m, x, Phi, sym)) < 1e-9 * abs(Sf)
    errs = []
    finite = np.nonzero(np.isfinite(m.Mass))[0]
    for _ in range(nprobe):
        p, l = rng.choice(finite), rng.integers(m.Ltau)
        xp, xm = x.copy(), x.copy()
        xp[p, l] += h
        xm[p, l] -= h
        fd = (dr.fermionic_action_dense(m, xp, Phi, sym) - dr.fermionic_action_dense(m, xm, Phi, sym)) / (2 * h)
        errs.append(abs(F[p, l] - fd) / max(1e-3, abs(fd)))
    return max(errs), F


@pytest.mark.parametrize("sym", [True, False])
def test_force_ssh_matches_finite_differences(sym):
    err, _ = _force_fd(mdl.ossh_chain(6, 0.3), sym, False, np.random.default_rng(11))
    assert err < 1e-6
    err, F = _force_fd(mdl.bssh_square(2, 2, 0.2), sym, False, np.random.default_rng(12))
    assert err < 1e-6
    m = mdl.bssh_square(2, 2, 0.2)
    assert np.all(F[~np.isfinite(m.Mass), :] == 0.0)        # frozen modes feel no force
    err, _ = _force_fd(mdl.holstein_ssh_chain(5, 0.3), sym, False, np.random.default_rng(13))
    assert err < 1e-6


def test_force_holstein_reference_quirk_and_exact_variant():
    """the Holstein-only Sym branch applies Gamma^-T where the exact derivative needs
    Gamma^-1.  The oracle ships the reference behaviour; the exact variant matches finite differences."""
    m = mdl.holstein_square(6, 2, 0.2)          # 6-ring: the colour factors do not commute
    err_exact, _ = _force_fd(m, True, True, np.random.default_rng(14))
    assert err_exact < 1e-6
    err_ref, _ = _force_fd(m, True, False, np.random.default_rng(14))
    assert 1e-6 < err_ref < 5e-2
    err_asym, _ = _force_fd(m, False, False, np.random.default_rng(15))
    assert err_asym < 1e-6


def test_efa_exact_flow():
    m = mdl.bssh_square(2, 2, 0.5)
    rng = np.random.default_rng(21)
    e = orc.RefElPh(m)
    x = m.random_fields(rng)
    e.set_x(x)
    a = orc.RefEFA(e)
    R = rng.standard_normal((m.Nph, m.Ltau))
    p, K = a.init_momentum(R)
    fin = np.isfinite(m.Mass)
    assert abs(K - 0.5 * np.sum(R[fin] ** 2)) < 1e-10 * K
    assert abs(a.kinetic(p) - K) < 1e-10 * K
    assert np.all(p[~fin] == 0)
    H0 = e.bosonic_action() + K
    x1, p1 = a.evolve(x, p, 0.37)
    e.set_x(x1)
    H1 = e.bosonic_action() + a.kinetic(p1)
    assert abs(H1 - H0) < 1e-10 * abs(H0)                 # exact flow conserves the free energy
    x2, p2 = a.evolve(x1, -p1, 0.37)                       # and is time reversible
    assert np.abs(x2 - x).max() < 1e-11 and np.abs(p2 + p).max() < 1e-11
    x3, p3 = a.evolve(x, p, 2 * np.pi)                     # unit frequency for every mode (eta = 0)
    assert np.abs(x3 - x).max() < 1e-10 and np.abs(p3 - p).max() < 1e-10


def test_hmc_energy_conservation_improves_with_Nt():
    m = mdl.holstein_honeycomb(2, 0.5)
    dHs = []
    for Nt in (4, 16):
        rng = np.random.default_rng(31)
        f = orc.RefFDM(m, sym=True, tol=1e-12, maxiter=5000)
        e = orc.RefElPh(m)
        e.set_x(m.random_fields(rng))
        e.refresh(f)
        pff = orc.RefPFF(e, f, exact_holstein=True)
        a = orc.RefEFA(e)
        rnd = np.concatenate([[0.5], rng.standard_normal(orc.hmc_random_count(m, Nt, False) - 2), [0.0]])
        x_before = e.x.copy()
        acc, info = orc.hmc_update(e, f, pff, a, None, Nt, np.pi / (2 * Nt), 0.0, 1e-12, 1e-12, 5000, rnd)
        dHs.append(abs(info[1]))
        assert acc and not np.allclose(e.x, x_before)
    assert dHs[1] < dHs[0] / 6            # leapfrog: error ~ dt^2


def test_scalar_measurements_free_fermions():
    """KAT (8): alpha = 0 => <n> from the estimator equals the exact free-fermion density."""
    m = mdl.holstein_square(2, 2, 0.5, alpha=0.0, mu=-0.4)
    rng = np.random.default_rng(41)
    f = orc.RefFDM(m, sym=True, tol=1e-13, maxiter=2000)
    V, t = dr.build_Vt(m, np.zeros((m.Nph, m.Ltau)))
    f.update(V, t)
    Nrv, Vdim = 64, m.N * m.Ltau
    R = rng.standard_normal((Vdim, Nrv)) + 1j * rng.standard_normal((Vdim, Nrv))
    R = np.asfortranarray(R / np.abs(R))
    GR = np.zeros((Vdim, Nrv), np.complex128, order="F")
    orc.greens_update(f, None, R, GR, 1e-13, 2000)
    M = dr.dense_M(m, dr.propagators(m, V, t, True))
    G = np.linalg.inv(M)
    assert np.abs(GR - G @ R).max() < 1e-9
    n_exact = 1 - np.trace(G) / Vdim
    n_est = orc.measure("n", R, GR)
    assert abs(n_est - n_exact) < 0.05
    # exact density from the single-particle propagator: G(tau,tau) = (1 + B^L)^-1
    B = dr.propagators(m, V, t, True)[0]
    n_ed = 1 - np.trace(np.linalg.inv(np.eye(m.N) + np.linalg.matrix_power(B, m.Ltau))) / m.N
    assert abs(n_exact - n_ed) < 1e-10
    d = orc.measure("double_occ", R, GR)
    assert abs(d.real - n_exact.real**2) < 0.05          # non-interacting: <n_up n_dn> = n^2 per spin


def test_dispersive_couplings_action_and_derivative():
    """Dispersive phonon couplings [unvendored SmoQyDQMC arithmetic, restated]: the derivative the kick uses is the gradient of the
    dispersive part of the bosonic action (central differences), the term vanishes for equal displacements, and a frozen partner
    (M = inf) gives the reduced mass of the free one."""
    m = mdl.with_dispersion(mdl.holstein_square(4, 4, 0.5), 0.7, 0.3)
    m0 = mdl.holstein_square(4, 4, 0.5)
    e, e0 = orc.RefElPh(m), orc.RefElPh(m0)
    rng = np.random.default_rng(0)
    x = m.random_fields(rng)

    def extra(xx):
        e.set_x(xx); e0.set_x(xx)
        return e.bosonic_action() - e0.bosonic_action()
    e.set_x(x); e0.set_x(x)
    G = e.potential_derivative() - e0.potential_derivative()
    h = 1e-5
    for k in [(0, 0), (3, 2), (15, 9), (7, 5)]:
        xp, xm = x.copy(), x.copy()
        xp[k] += h; xm[k] -= h
        fd = (extra(xp) - extra(xm)) / (2 * h)
        assert abs(G[k] - fd) < 1e-6 * max(1.0, abs(fd)), (k, G[k], fd)
    assert abs(extra(np.full_like(x, 0.37))) < 1e-12
    assert extra(x) > 0
