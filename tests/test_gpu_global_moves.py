"""reflection_update! / swap_update! / radial_update! through the C ABI against the same steps on the CPU oracle
(/root/reference/src/reflection_update.jl:67-176, swap_update.jl:68-176, radial_update.jl:88-193), identical random numbers."""
import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def oracle_move(re, rf, rp, mutate, log_jac, R, u, tol):
    """The reference sequence on the oracle objects; returns (accepted, iters, dS)."""
    Sf = rp.sample(R)
    Sb = re.bosonic_action()
    x0 = np.array(re.x, copy=True)
    mutate(re.x)
    re.refresh(rf)
    Sf2, iters, _ = rp.action(tol=tol)
    dS = (Sf2 + re.bosonic_action()) - (Sf + Sb)
    P = min(1.0, float(np.exp(-dS + log_jac)))
    acc = u < P
    if not acc:
        re.set_x(x0)
        re.refresh(rf)
    return acc, iters, dS


def setup(name):
    from smoqyelph_b200 import api
    m = {"honeycomb": lambda: mdl.config("cfg1t"), "bssh": lambda: mdl.bssh_square(4, 4, 0.5),
         "square": lambda: mdl.holstein_square(16, 16, 0.5)}[name]()
    rng = np.random.default_rng(11)
    x = m.random_fields(rng, smooth=True)
    fdm = api.SymFermionDetMatrix(m, tol=1e-12, maxiter=20000)
    elph = api.ElectronPhononParameters(m, fdm)
    elph.x = x
    elph.update_fdm()
    pff = api.PFFCalculator(elph)
    rf = orc.RefFDM(m, sym=True)
    re = orc.RefElPh(m)
    re.set_x(x)
    re.refresh(rf)
    rp = orc.RefPFF(re, rf)
    return api, m, rng, elph, pff, re, rf, rp


def rand_R(rng, m):
    return np.asfortranarray((rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2))


@pytest.mark.parametrize("name", ["honeycomb", "bssh", "square"])
def test_reflection_swap_radial_match_oracle(name):
    api, m, rng, elph, pff, re, rf, rp = setup(name)
    finite = np.nonzero(np.isfinite(m.Mass))[0]
    for move in ("reflection", "swap", "radial", "reflection", "swap", "radial"):
        R, u = rand_R(rng, m), float(rng.random())
        if move == "reflection":
            p = int(rng.choice(finite))
            acc, it = api.reflection_update(elph, pff, tol=1e-12, randoms={"mode": p, "R": R, "u": u})
            info = api.reflection_update.last
            def mut(x): x[p, :] *= -1.0
            racc, rit, rdS = oracle_move(re, rf, rp, mut, 0.0, R, u, 1e-12)
        elif move == "swap":
            i, j = (int(a) for a in rng.choice(finite, 2, replace=False))
            acc, it = api.swap_update(elph, pff, tol=1e-12, randoms={"modes": (i, j), "R": R, "u": u})
            info = api.swap_update.last
            def mut(x): x[[i, j], :] = x[[j, i], :]
            racc, rit, rdS = oracle_move(re, rf, rp, mut, 0.0, R, u, 1e-12)
        else:
            g = float(rng.standard_normal())
            acc, it = api.radial_update(elph, pff, tol=1e-12, sigma=0.5, randoms={"gamma_normal": g, "R": R, "u": u})
            info = api.radial_update.last
            gamma, d = info["gamma"], info["d"]
            assert d == len(finite) * m.Ltau
            def mut(x): x[...] = np.exp(gamma) * x
            racc, rit, rdS = oracle_move(re, rf, rp, mut, d * gamma, R, u, 1e-12)
        assert abs(info["dS"] - rdS) < 1e-8 * max(1.0, abs(rdS)), (move, info["dS"], rdS)
        assert acc == racc and abs(it - rit) <= 1, (move, acc, racc, it, rit)
        assert np.abs(elph.x - re.x).max() < 1e-14, move
    # the operator follows the field after every move (accepted or rejected)
    v = rand_R(rng, m)
    assert np.linalg.norm(elph.fdm.mul_MtM(v) - rf.mul_MtM(v)) < 1e-12 * np.linalg.norm(v)


def test_moves_sample_modes_and_respect_frozen_phonons():
    """Random proposals: frozen (M = inf) modes are never touched; phonon_types / phonon_id restrict the proposal."""
    api, m, rng, elph, pff, re, rf, rp = setup("bssh")
    x0 = elph.x
    frozen = ~np.isfinite(m.Mass)
    assert frozen.any()
    for _ in range(4):
        api.reflection_update(elph, pff, rng=rng, tol=1e-8, phonon_types=[0])
        assert api.reflection_update.last["mode"] < m.n_unit_cells
        api.swap_update(elph, pff, rng=rng, tol=1e-8)
        i, j = api.swap_update.last["modes"]
        assert not frozen[i] and not frozen[j] and i != j
        api.radial_update(elph, pff, rng=rng, tol=1e-8, phonon_id=1, sigma=0.3)
    x1 = elph.x
    assert np.array_equal(x1[frozen], x0[frozen])
    assert np.all(x1[2 * m.n_unit_cells:] == 0.0)


def test_detailed_balance_of_rejection():
    """A rejected move restores x and the operator bit for bit."""
    api, m, rng, elph, pff, re, rf, rp = setup("honeycomb")
    x0 = elph.x
    e0 = elph.fdm.coefficients()[0]
    acc, _ = api.radial_update(elph, pff, tol=1e-10, sigma=1.0, randoms={"gamma_normal": 3.0, "R": rand_R(rng, m), "u": 1.0})
    assert not acc
    assert np.array_equal(elph.x, x0) and np.array_equal(elph.fdm.coefficients()[0], e0)
