"""Device-side correlation measurements against a numpy restatement of the reference:
measure_GΔ0! (/root/reference/src/Measurements/GreensEstimator.jl:177-233) with _aperiodic_copyto! (:656-671),
_translational_average! (:674-705) and add_contraction_to_correlation! (:718-729)."""
import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
import dense_ref as dr

pytestmark = pytest.mark.gpu


def ref_measure_GD0(R, GR, Ltau, norb, dims, a, b):
    """Literal restatement.  R, GR: (V, Nrv) columns in the reference layout (tau fastest, then orbital, then cells)."""
    Nrv = R.shape[1]
    shape = (Ltau, norb) + tuple(dims)
    Rt = np.conj(R).reshape(shape + (Nrv,), order="F")
    G = GR.reshape(shape + (Nrv,), order="F")
    S = np.zeros((Ltau + 1,) + tuple(dims), complex)
    for i in range(Nrv):
        A = np.concatenate([G[:, a, ..., i], -G[:, a, ..., i]], axis=0)          # aperiodic copy to 2 Ltau
        B = np.concatenate([Rt[:, b, ..., i], -Rt[:, b, ..., i]], axis=0)
        c = np.fft.ifftn(np.fft.fftn(A) * np.fft.ifftn(B))                        # FFTW: fft unnormalised, ifft 1/M
        S[:Ltau] += c[:Ltau]
        S[Ltau] += c[0]
    S /= Nrv
    S[Ltau] = -S[Ltau]
    if a == b:
        S[(Ltau,) + (0,) * len(dims)] += 1.0
    return np.moveaxis(S, 0, -1)                                                  # correlation axes: (L..., Ltau + 1)


def direct_translational_average(R, GR, Ltau, norb, dims, a, b, dl, dr_):
    """Independent O(V) check of one displacement: (1/(2 Ltau Nc)) sum over the doubled torus of A(x + D) B(x)."""
    Nrv = R.shape[1]
    shape = (Ltau, norb) + tuple(dims)
    Rt = np.conj(R).reshape(shape + (Nrv,), order="F")
    G = GR.reshape(shape + (Nrv,), order="F")
    tot = 0.0
    for i in range(Nrv):
        A = np.concatenate([G[:, a, ..., i], -G[:, a, ..., i]], axis=0)
        B = np.concatenate([Rt[:, b, ..., i], -Rt[:, b, ..., i]], axis=0)
        Ash = np.roll(A, shift=[-dl] + [-d for d in dr_], axis=tuple(range(A.ndim)))
        tot += np.mean(Ash * B)
    return tot / Nrv


@pytest.mark.parametrize("name", ["honeycomb", "square", "chain"])
def test_measure_GD0_matches_reference_formula(name):
    from smoqyelph_b200 import api
    m = {"honeycomb": lambda: mdl.holstein_honeycomb(3, 1.0, mu=0.2), "square": lambda: mdl.holstein_square(16, 16, 0.5),
         "chain": lambda: mdl.ossh_chain(12, 0.6)}[name]()
    rng = np.random.default_rng(4)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    fdm.update(V, t)
    g = api.GreensEstimator(fdm, Nrv=5, seed=1)
    g.update_greens_estimator(tol=1e-12)
    R, GR = g.get()
    dims = tuple(m.lattice_dims)
    norb = m.N // int(np.prod(dims))
    for a in range(norb):
        for b in range(norb):
            got = g.measure_GD0((a, b))
            want = ref_measure_GD0(R, GR, m.Ltau, norb, dims, a, b)
            assert got.shape == want.shape == dims + (m.Ltau + 1,)
            assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max()), (name, a, b, np.abs(got - want).max())
    # one displacement against the direct sum (checks the restatement itself)
    dl, dsp = 3 % m.Ltau, [1] + [0] * (len(dims) - 1)
    got = g.measure_GD0((0, norb - 1))
    direct = direct_translational_average(R, GR, m.Ltau, norb, dims, 0, norb - 1, dl, dsp)
    assert abs(got[tuple(dsp) + (dl,)] - direct) < 1e-12


def test_GD0_of_free_fermions_is_the_exact_greens_function():
    """alpha = 0: the stochastic estimator averaged over many random vectors reproduces the exact equal-time and
    time-displaced free-fermion Green's function G(r, tau) = (1/N) sum_k e^{ikr} e^{-tau eps_k} / (1 + e^{-beta eps_k})
    of the checkerboard-decomposed propagator to the statistical error of the estimator (~1/sqrt(Nrv V))."""
    from smoqyelph_b200 import api
    m = mdl.holstein_square(16, 16, 0.4, alpha=0.0, mu=-0.3)
    V, t = dr.build_Vt(m, np.zeros((m.Nph, m.Ltau)))
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    fdm.update(V, t)
    g = api.GreensEstimator(fdm, Nrv=40, seed=3)
    g.update_greens_estimator(tol=1e-12)
    G = g.measure_GD0((0, 0)).real
    # exact G(r, tau) through the dense propagator of the same checkerboard operator (tau-independent fields)
    B = dr.propagators(m, V, t, sym=True)[0]
    G0 = np.linalg.inv(np.eye(m.N) + np.linalg.matrix_power(B, m.Ltau))
    Lx, Ly = m.lattice_dims
    idx = np.arange(m.N)
    x, y = idx % Lx, idx // Lx

    def averaged(Gl, dx, dy):                               # (1/N) sum_r G(r + d | r)
        j = ((x + dx) % Lx) + Lx * ((y + dy) % Ly)
        return np.mean(Gl[j, idx])

    err = 10.0 / np.sqrt(40 * m.N)                           # estimator noise per displacement ~ 1 / sqrt(Nrv N Ltau) x O(1)
    for l in (0, 1, m.Ltau // 2, m.Ltau - 1):
        Gl = np.linalg.matrix_power(B, l) @ G0
        for d in ((0, 0), (1, 0), (0, 1), (3, 2)):
            assert abs(G[d[0], d[1], l] - averaged(Gl, *d)) < err, (l, d, G[d[0], d[1], l], averaged(Gl, *d))
    assert abs(G[0, 0, m.Ltau] - (1.0 - averaged(G0, 0, 0))) < err
    assert abs(G[1, 0, m.Ltau] + averaged(G0, 1, 0)) < err


# ---- four-point contractions (src/Measurements/GreensEstimator.jl:236-652) -----------------------------------------------
def _fields(R, GR, Ltau, norb, dims):
    Nrv = R.shape[1]
    shape = (Ltau, norb) + tuple(dims) + (Nrv,)
    return np.conj(R).reshape(shape, order="F"), GR.reshape(shape, order="F")


def _shift(x, r):
    """ShiftedArrays.circshift(x, (0, (-r)...)): result[tau, i] = x[tau, i + r]."""
    return np.roll(x, shift=[-int(q) for q in r], axis=tuple(range(1, 1 + len(r))))


def _tavg(S, A, B, Ltau):
    """_translational_average! (:674-705) on the periodic Ltau x L... torus."""
    c = np.fft.ifftn(np.fft.fftn(A) * np.fft.ifftn(B))
    S[:Ltau] += c[:Ltau]
    S[Ltau] += c[0]


def ref_contraction(kind, R, GR, Ltau, norb, dims, orbitals, r, tD=None, t0=None):
    """tD, t0: optional hopping weights (Ltau, L...) of _measure_CΔ0! (:626-646) and of the weighted delta terms."""
    Rt, G = _fields(R, GR, Ltau, norb, dims)
    wD = 1.0 if tD is None else tD
    w0 = 1.0 if t0 is None else t0
    Nrv, D = R.shape[1], len(dims)
    a, b, c, d = orbitals
    r1, r2, r3, r4 = (np.asarray(q) for q in r)
    S = np.zeros((Ltau + 1,) + tuple(dims), complex)
    for n in range(Nrv - 1):
        for m in range(n + 1, Nrv):
            GRa, Rtb = _shift(G[:, a, ..., n], r1), _shift(Rt[:, b, ..., n], r2)
            GRc, Rtd = _shift(G[:, c, ..., m], r3), _shift(Rt[:, d, ..., m], r4)
            if kind == "GD0_GD0":
                _tavg(S, wD * GRa * GRc, w0 * Rtb * Rtd, Ltau)
            elif kind == "GDD_G00":
                _tavg(S, wD * GRa * Rtb, w0 * GRc * Rtd, Ltau)
            else:
                _tavg(S, wD * Rtb * GRc, w0 * GRa * Rtd, Ltau)
    S /= Nrv * (Nrv - 1) / 2

    def mean_shifted(orb_g, shift, orb_r, tshift=None):
        """sum over rv of sum(circshift(GR, (0, shift...)) .* Rt) / (Nrv * length): circshift by +s: result[i] = GR[i - s];
        with weights: times circshift(tΔ, (0, tshift...)) .* t0 (:318-326, :560-568)."""
        tot = 0.0
        w = 1.0
        if tD is not None:
            w = np.roll(tD, shift=[int(q) for q in tshift], axis=tuple(range(1, 1 + D))) * t0
        for n in range(Nrv):
            g = np.roll(G[:, orb_g, ..., n], shift=[int(q) for q in shift], axis=tuple(range(1, 1 + D)))
            tot += np.sum(w * g * Rt[:, orb_r, ..., n]) / g.size
        return tot / Nrv

    def at(v):
        return tuple(int(q) % L for q, L in zip(v, dims))

    if kind == "GD0_GD0":
        if a == b:
            S[(Ltau,) + at(-r1 + r2)] -= mean_shifted(c, r1 - r2 - r3 + r4, d, r1 - r2)
        if c == d:
            S[(Ltau,) + at(-r3 + r4)] -= mean_shifted(a, -r1 + r2 + r3 - r4, b, r3 - r4)
        if a == b and c == d and at(r2 - r1) == at(r4 - r3):
            S[(Ltau,) + at(r2 - r1)] += 1 if tD is None else np.mean(np.roll(tD, shift=[int(q) for q in r1 - r2], axis=tuple(range(1, 1 + D))) * t0)
    elif kind == "G0D_GD0":
        if a == b:
            S[(0,) + at(r1 - r2)] -= mean_shifted(c, -r1 + r2 - r3 + r4, d, -r1 + r2)
        if c == d:
            S[(Ltau,) + at(r4 - r3)] -= mean_shifted(a, -r1 + r2 - r3 + r4, b, -r4 + r3)
    return np.moveaxis(S, 0, -1)


@pytest.mark.parametrize("name", ["honeycomb", "square"])
def test_four_point_contractions_and_density_correlation(name):
    from smoqyelph_b200 import api
    m = {"honeycomb": lambda: mdl.holstein_honeycomb(3, 0.6, mu=0.2), "square": lambda: mdl.holstein_square(16, 16, 0.3)}[name]()
    rng = np.random.default_rng(8)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    fdm.update(V, t)
    g = api.GreensEstimator(fdm, Nrv=4, seed=5)
    g.update_greens_estimator(tol=1e-12)
    R, GR = g.get()
    dims = tuple(m.lattice_dims)
    norb = m.N // int(np.prod(dims))
    zero = [(0,) * len(dims)] * 4
    shifted = [(1, 0), (0, 2), (2, 1), (0, 0)]
    cases = [((0, 0, 0, 0), zero), ((norb - 1, 0, 0, norb - 1), zero), ((0, 0, norb - 1, norb - 1), shifted), ((0, norb - 1, norb - 1, 0), shifted)]
    for kind in ("GD0_GD0", "GDD_G00", "G0D_GD0"):
        for orbs, r in cases:
            got = g.measure_contraction(kind, orbs, r)
            want = ref_contraction(kind, R, GR, m.Ltau, norb, dims, orbs, r)
            assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max()), (name, kind, orbs, r, np.abs(got - want).max())
    # bond correlation (src/Measurements/Correlations/bond.jl:2-48): 8 contractions with the bond displacements
    b1 = ((norb - 1, 0), (1,) + (0,) * (len(dims) - 1))
    b2 = ((0, norb - 1), (0,) * (len(dims) - 1) + (1,))
    (bb, ba), r1 = b1
    (bd, bc), r2 = b2
    z = (0,) * len(dims)
    rc = lambda kind, orbs, r: ref_contraction(kind, R, GR, m.Ltau, norb, dims, orbs, r)
    want = 4 * (rc("GDD_G00", (ba, bb, bc, bd), (r1, z, r2, z)) + rc("GDD_G00", (ba, bb, bd, bc), (r1, z, z, r2))
                + rc("GDD_G00", (bb, ba, bc, bd), (z, r1, r2, z)) + rc("GDD_G00", (bb, ba, bd, bc), (z, r1, z, r2))) \
        - 2 * (rc("G0D_GD0", (bc, bb, ba, bd), (r2, z, r1, z)) + rc("G0D_GD0", (bd, bb, ba, bc), (z, z, r1, r2))
               + rc("G0D_GD0", (bc, ba, bb, bd), (r2, r1, z, z)) + rc("G0D_GD0", (bd, ba, bb, bc), (z, r1, z, r2)))
    got = g.measure_bond_correlation(b1, b2)
    assert np.abs(got - want).max() < 1e-11 * max(1.0, np.abs(want).max())
    # spin-resolved forms (bond.jl:66-131): factors 1, exchange terms only for equal spins
    gdd = (rc("GDD_G00", (ba, bb, bc, bd), (r1, z, r2, z)) + rc("GDD_G00", (ba, bb, bd, bc), (r1, z, z, r2))
           + rc("GDD_G00", (bb, ba, bc, bd), (z, r1, r2, z)) + rc("GDD_G00", (bb, ba, bd, bc), (z, r1, z, r2)))
    for spins, w in (((+1, -1), gdd), ((-1, -1), gdd - (4 * gdd - want) / 2)):
        got = g.measure_bond_correlation(b1, b2, spins=spins)
        assert np.abs(got - w).max() < 1e-11 * max(1.0, np.abs(w).max()), spins
    # density correlation (src/Measurements/Correlations/density.jl:2-33) assembled from the contractions
    Rt, G = _fields(R, GR, m.Ltau, norb, dims)
    for a, b in ((0, 0), (0, norb - 1)):
        na = 1 - np.sum(G[:, a] * Rt[:, a]) / G[:, a].size
        nb = 1 - np.sum(G[:, b] * Rt[:, b]) / G[:, b].size
        assert abs(g.measure_n_orbital(a) - na) < 1e-13
        want = 4 * (na + nb - 1) + 4 * ref_contraction("GDD_G00", R, GR, m.Ltau, norb, dims, (a, a, b, b), zero) \
            - 2 * ref_contraction("G0D_GD0", R, GR, m.Ltau, norb, dims, (b, a, a, b), zero)
        got = g.measure_density_correlation(a, b)
        assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
        # spin-resolved (density.jl:33-65)
        dd = ref_contraction("GDD_G00", R, GR, m.Ltau, norb, dims, (a, a, b, b), zero)
        ex = ref_contraction("G0D_GD0", R, GR, m.Ltau, norb, dims, (b, a, a, b), zero)
        assert np.abs(g.measure_density_correlation(a, b, spins=(+1, +1)) - ((na + nb - 1) + dd - ex)).max() < 1e-12 * max(1.0, np.abs(dd).max())
        assert np.abs(g.measure_density_correlation(a, b, spins=(+1, -1)) - ((na + nb - 1) + dd)).max() < 1e-12 * max(1.0, np.abs(dd).max())


@pytest.mark.parametrize("name", ["honeycomb", "square"])
def test_hopping_weighted_contractions_and_current_correlation(name):
    """tΔ / t0 arguments of the contractions and measure_current_correlation! (Correlations/current.jl:2-151)."""
    from smoqyelph_b200 import api
    m = {"honeycomb": lambda: mdl.holstein_honeycomb(3, 0.6, mu=0.2), "square": lambda: mdl.holstein_square(16, 16, 0.3)}[name]()
    rng = np.random.default_rng(9)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    fdm.update(V, t)
    g = api.GreensEstimator(fdm, Nrv=3, seed=6)
    g.update_greens_estimator(tol=1e-12)
    R, GR = g.get()
    dims = tuple(m.lattice_dims)
    norb = m.N // int(np.prod(dims))
    t1 = 1.0 + 0.3 * rng.standard_normal((m.Ltau,) + dims)          # space-time dependent hoppings (as in an SSH model)
    t2 = 1.0 + 0.3 * rng.standard_normal((m.Ltau,) + dims)
    zero = [(0,) * len(dims)] * 4
    shifted = [(1, 0), (0, 2), (2, 1), (0, 0)]
    cases = [((0, 0, 0, 0), zero), ((0, 0, norb - 1, norb - 1), shifted), ((0, norb - 1, norb - 1, 0), shifted), ((norb - 1, norb - 1, 0, 0), shifted)]
    for kind in ("GD0_GD0", "GDD_G00", "G0D_GD0"):
        for orbs, r in cases:
            got = g.measure_contraction(kind, orbs, r, tD=t1, t0=t2)
            want = ref_contraction(kind, R, GR, m.Ltau, norb, dims, orbs, r, t1, t2)
            assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max()), (name, kind, orbs, r, np.abs(got - want).max())
    # one-sided weights (tΔ only) where no delta term needs both
    got = g.measure_contraction("GDD_G00", (0, 0, norb - 1, norb - 1), shifted, tD=t1)
    want = ref_contraction("GDD_G00", R, GR, m.Ltau, norb, dims, (0, 0, norb - 1, norb - 1), shifted, t1, None)
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
    # current correlation: literal restatement of the eight terms
    for b1, b2 in ((((norb - 1, 0), (1,) + (0,) * (len(dims) - 1)), ((0, norb - 1), (0,) * (len(dims) - 1) + (1,))),
                   (((0, 0), (1,) + (0,) * (len(dims) - 1)), ((0, 0), (1,) + (0,) * (len(dims) - 1)))):
        (bb, ba), r1 = b1
        (bd, bc), r2 = b2
        z = (0,) * len(dims)
        rc = lambda kind, orbs, r: ref_contraction(kind, R, GR, m.Ltau, norb, dims, orbs, r, t1, t2)
        gdd = rc("GDD_G00", (ba, bb, bd, bc), (r1, z, z, r2)) - rc("GDD_G00", (ba, bb, bc, bd), (r1, z, r2, z)) \
            - rc("GDD_G00", (bb, ba, bd, bc), (z, r1, z, r2)) + rc("GDD_G00", (bb, ba, bc, bd), (z, r1, r2, z))
        g0d = -rc("G0D_GD0", (bb, ba, bc, bd), (z, z, r1, r2)) + rc("G0D_GD0", (bb, ba, bd, bc), (r2, z, r1, z)) \
            + rc("G0D_GD0", (bd, ba, bb, bc), (z, r1, z, r2)) - rc("G0D_GD0", (bc, ba, bb, bd), (r2, r1, z, z))
        for spins, want in ((None, 4 * gdd + 2 * g0d), ((+1, +1), gdd + g0d), ((+1, -1), gdd)):
            got = g.measure_current_correlation(b1, b2, t1, t2, spins=spins)
            assert np.abs(got - want).max() < 1e-11 * max(1.0, np.abs(want).max()), (name, b1, b2, spins)


# ---- local measurements (tight_binding_measurements.jl:43-133, electron_phonon_measurements.jl) ---------------------------
@pytest.mark.parametrize("name", ["honeycomb", "bssh"])
def test_local_measurements(name):
    from smoqyelph_b200 import api
    m = {"honeycomb": lambda: mdl.holstein_honeycomb(3, 0.6, mu=0.2), "bssh": lambda: mdl.bssh_square(4, 4, 0.5)}[name]()
    rng = np.random.default_rng(12)
    x = m.random_fields(rng, smooth=True)
    V, t = dr.build_Vt(m, x)
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    fdm.update(V, t)
    g = api.GreensEstimator(fdm, Nrv=5, seed=2)
    g.update_greens_estimator(tol=1e-12)
    R, GR = g.get()
    Nrv, Lt, N = R.shape[1], m.Ltau, m.N
    Rt = np.conj(R).reshape((Lt, N, Nrv), order="F")
    G = GR.reshape((Lt, N, Nrv), order="F")
    dims = tuple(m.lattice_dims)
    norb = N // int(np.prod(dims))
    ncell = N // norb
    # on-site energy, literal restatement of :43-63
    eps, mu = rng.standard_normal(N), 0.3
    for orb in range(norb):
        e = 0.0
        for u in range(ncell):
            i = orb + norb * u
            e += (eps[i] - mu) * np.sum(1 - G[:, i, :] * Rt[:, i, :]) / (Lt * Nrv)
        e /= ncell
        assert abs(g.measure_onsite_energy(orb, eps, mu) - e) < 1e-12 * max(1.0, abs(e))
    # bare and modulated hopping energy of the first hopping id (:66-133)
    nt = m.neighbor_table[:, :ncell]
    for tt in (m.t0[:ncell] * (1 + 0.1j), t[:ncell, :].astype(complex)):
        h = 0.0
        for mm in range(ncell):
            i, f = nt[0, mm], nt[1, mm]
            tl = np.broadcast_to(tt[mm], (Lt,))
            h += np.sum(tl[:, None] * G[:, i, :] * Rt[:, f, :] + np.conj(tl)[:, None] * G[:, f, :] * Rt[:, i, :])
        h /= Lt * N * Nrv
        assert abs(g.measure_hopping_energy(nt, tt) - h) < 1e-12 * max(1.0, abs(h))
    if m.Nhol:                                                       # Holstein energy (electron_phonon_measurements.jl)
        want = 0.0
        for c in range(ncell):
            i, p = m.hol_site[c], m.hol_phonon[c]
            a1, a2, a3, a4 = m.hol_alpha[:, c]
            for l in range(Lt):
                n_li = np.sum(1 - G[l, i, :] * Rt[l, i, :]) / Nrv
                xv = x[p, l]
                want += (a2 * xv ** 2 + a4 * xv ** 4) * n_li
                want += (a1 * xv + a3 * xv ** 2) * (n_li - 0.5 if m.hol_phsym[c] else n_li)
        want /= ncell * Lt
        assert abs(g.measure_holstein_energy(x, 0) - want) < 1e-12 * max(1.0, abs(want))
    for ssh_id in range(m.Nssh // ncell):                            # SSH energy, literal restatement of :124-186
        want = 0.0
        for u in range(ncell):
            cpl = ssh_id * ncell + u
            s_i, s_f = m.neighbor_table[:, m.ssh_hopping[cpl]]
            p_i, p_f = m.ssh_phonon[:, cpl]
            a1, a2, a3, a4 = m.ssh_alpha[:, cpl]
            for l in range(Lt):
                dx = x[p_f, l] - x[p_i, l]
                c = a1 * dx + a2 * dx ** 2 + a3 * dx ** 3 + a4 * dx ** 4
                hf = -np.sum(G[l, s_i, :] * Rt[l, s_f, :]) / Nrv
                hr = -np.sum(G[l, s_f, :] * Rt[l, s_i, :]) / Nrv
                want += c * hf + np.conj(c) * hr
        want /= ncell * Lt
        assert abs(g.measure_ssh_energy(x, ssh_id) - want) < 1e-12 * max(1.0, abs(want))


def test_make_measurements_twin_and_orbital_double_occupancy():
    """measure_double_occ(greens_estimator, orbital) (scalar_measurements.jl:98-109, divided by the TOTAL V as the reference does) and the
    accumulation order of make_measurements! (make_measurements.jl:19-146) on a dictionary."""
    from smoqyelph_b200 import api
    m = mdl.holstein_honeycomb(3, 0.6, mu=0.2)
    rng = np.random.default_rng(12)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    fdm.update(V, t)
    g = api.GreensEstimator(fdm, Nrv=4, seed=7)
    meas = {}
    zero = (0, 0)
    corr = [("greens", (0, 1)), ("density", (0, 1)), ("bond", (((0, 1), zero), ((0, 1), (1, 0))))]
    it1 = api.make_measurements(meas, fdm, g, mu=0.2, bosonic_action=1.5, tol=1e-12, correlations=corr)
    R, GR = g.get()
    dims = tuple(m.lattice_dims)
    norb = m.N // int(np.prod(dims))
    Rt, G = _fields(R, GR, m.Ltau, norb, dims)
    Nrv, Vtot = R.shape[1], m.N * m.Ltau
    for a in range(norb):
        d = 0.0
        for i in range(Nrv - 1):
            for j in range(i + 1, Nrv):
                d += np.sum((1 - G[:, a, ..., i] * Rt[:, a, ..., i]) * (1 - G[:, a, ..., j] * Rt[:, a, ..., j])) / Vtot
        d /= Nrv * (Nrv - 1) / 2
        assert abs(g.measure_double_occ_orbital(a) - d) < 1e-13
        assert abs(meas["local"]["double_occ"][a] - d) < 1e-13
        n = 1 - np.sum(G[:, a] * Rt[:, a]) / G[:, a].size
        assert abs(meas["local"]["density"][a] - 2 * n) < 1e-13
    s = g.measure()
    assert abs(sum(meas["local"]["double_occ"]) - s["double_occ"]) < 1e-13          # the orbital parts add up to the global value
    assert abs(meas["global"]["density"] - 2 * s["n"]) < 1e-14 and meas["global"]["sgn"] == 1.0 and meas["global"]["action_bosonic"] == 1.5
    key = ("density", 0, 1)
    assert np.abs(meas["correlations"][key] - g.measure_density_correlation(0, 1)).max() < 1e-13
    # a second configuration accumulates
    it2 = api.make_measurements(meas, fdm, g, mu=0.2, bosonic_action=1.5, tol=1e-12, correlations=corr)
    assert meas["global"]["sgn"] == 2.0 and it1 > 0 and it2 > 0
    assert abs(meas["global"]["chemical_potential"] - 0.4) < 1e-15
