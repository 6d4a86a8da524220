"""GPU parity: FermionDetMatrix products and CG through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc
import dense_ref as dr

pytestmark = pytest.mark.gpu

RTOL = 1e-12   # north_star: M^T M v within 1e-12 relative error in Float64


def rand_cvec(rng, m):
    return np.asfortranarray((rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2))


def relerr(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


MODELS = {
    "cfg1t": lambda: mdl.config("cfg1t"),
    "cfg1": lambda: mdl.config("cfg1"),
    "cfg2": lambda: mdl.config("cfg2"),
    "cfg3": lambda: mdl.config("cfg3"),
    "cfg4": lambda: mdl.config("cfg4"),
    "cfg5": lambda: mdl.config("cfg5"),
    "mixed_odd": lambda: mdl.holstein_ssh_chain(7, 0.65),     # odd ring: 3 colours, last colour 1 bond, uncovered sites
    "tiny_L1": lambda: mdl.ossh_chain(4, 0.05),               # Ltau = 1
    "tiny_L2": lambda: mdl.ossh_chain(6, 0.10),               # Ltau = 2
}


def setup(name, sym, seed=0):
    from smoqyelph_b200 import api
    m = MODELS[name]()
    rng = np.random.default_rng(seed)
    x = m.random_fields(rng)
    V, t = dr.build_Vt(m, x)
    ref = orc.RefFDM(m, sym=sym)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m, sym=sym)
    fdm.update(V, t)
    return m, rng, ref, fdm


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("name", list(MODELS))
def test_products_match_oracle(name, sym):
    m, rng, ref, fdm = setup(name, sym)
    e, c, s = fdm.coefficients()
    assert np.abs(e - ref.expV).max() <= 4e-16 * np.abs(ref.expV).max()
    if m.Nh:
        assert np.abs(c - ref.cosh).max() <= 4e-16 * np.abs(ref.cosh).max()
        assert np.abs(s - ref.sinh).max() <= 1e-15
    v = rand_cvec(rng, m)
    for op in ("mul_M", "mul_Mt", "mul_MtM", "mul_MMt"):
        got = getattr(fdm, op)(v)
        want = getattr(ref, op)(v)
        assert relerr(got, want) < RTOL, (name, sym, op, relerr(got, want))


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4", "cfg5", "mixed_odd"])
def test_all_tunings_agree(name):
    m, rng, ref, fdm = setup(name, True)
    v = rand_cvec(rng, m)
    want = ref.mul_MtM(v)
    assert fdm.tuning["path"] in (0, 2, 3)
    results = {}
    for fast in (False, True):
        fdm.set_fast_path(fast)
        for slab in (1, 2, 3, 5):
            for threads in (64, 128, 256, 512, 576, 1024):
                try:
                    fdm.set_tuning(slab, threads)
                except Exception:
                    continue
                got = fdm.mul_MtM(v)
                assert relerr(got, want) < RTOL, (fast, slab, threads)
                assert relerr(fdm.mul_M(v), ref.mul_M(v)) < RTOL
                assert relerr(fdm.mul_Mt(v), ref.mul_Mt(v)) < RTOL
                results[(fast, slab, threads)] = got
    # the fast and the generic fused kernels perform the same arithmetic: bit-identical output
    vals = list(results.values())
    assert all(np.array_equal(vals[0], w) for w in vals[1:])


def test_linearity_and_adjointness_full_size():
    """Size-independent properties at BASELINE's full size (cfg4)."""
    m, rng, ref, fdm = setup("cfg4", True)
    a, b = rand_cvec(rng, m), rand_cvec(rng, m)
    al = 0.3 - 1.1j
    lhs = fdm.mul_MtM(a + al * b)
    rhs = fdm.mul_MtM(a) + al * fdm.mul_MtM(b)
    assert relerr(lhs, rhs) < 1e-13
    # <a, M b> == <M^T a, b> for real M
    l = np.vdot(a, fdm.mul_M(b))
    r = np.vdot(fdm.mul_Mt(a), b)
    assert abs(l - r) < 1e-12 * abs(l)
    # M^T M is positive: <a, M^T M a> = |M a|^2
    assert abs(np.vdot(a, fdm.mul_MtM(a)) - np.linalg.norm(fdm.mul_M(a)) ** 2) < 1e-11 * np.linalg.norm(a) ** 2


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("name", ["cfg1t", "cfg2", "mixed_odd"])
def test_cg_matches_oracle(name, sym):
    m, rng, ref, fdm = setup(name, sym)
    b = rand_cvec(rng, m)
    # solutions compared with both solvers converged far below the comparison tolerance (hard part 3)
    xr, itr, epsr = ref.cg(b, tol=1e-14, maxiter=20000)
    xg, itg, epsg = fdm.ldiv(b, tol=1e-14, maxiter=20000)
    assert epsg < 1e-14 and epsr < 1e-14
    assert relerr(xg, xr) < 1e-11
    # iteration counts at production tolerances
    for tol in (1e-5, 1e-10):
        _, itr, _ = ref.cg(b, tol=tol, maxiter=20000)
        _, itg, epsg = fdm.ldiv(b, tol=tol, maxiter=20000)
        assert abs(itg - itr) <= 1, (tol, itg, itr)
        assert epsg < tol
    # warm start: x0 = solution => 0 iterations ; maxiter cut-off returns maxiter
    _, it0, _ = fdm.ldiv(b, x0=xg, tol=1e-10)
    assert it0 == 0
    _, itc, epsc = fdm.ldiv(b, tol=1e-14, maxiter=3)
    assert itc == 3 and epsc > 1e-14


# ---- register path (fdm_v3.cu) and the resident CG solver -----------------------------------------------------
SQUARES = {
    "h16x16": lambda: mdl.holstein_square(16, 16, 2.0),
    "h16x16odd": lambda: mdl.holstein_square(16, 16, 0.35),    # Ltau = 7: the last CTA of the resident solver owns ONE slice (ragged slab)
    "h32x16": lambda: mdl.holstein_square(32, 16, 1.0),
    "h16x32": lambda: mdl.holstein_square(16, 32, 1.0),
    "h32x32": lambda: mdl.holstein_square(32, 32, 1.5),
    "h16x64": lambda: mdl.holstein_square(16, 64, 0.5),
    "h32x64": lambda: mdl.holstein_square(32, 64, 0.35),    # 64 rows: 16 rows per lane, Ltau = 7 (ragged last CTA)
    # honeycomb lattices (the reference's tutorial lattice; 24 x 24 is BASELINE config 5): honeycomb engine of fdm_v3.cu
    "hc8": lambda: mdl.holstein_honeycomb(8, 1.0),
    "hc16": lambda: mdl.holstein_honeycomb(16, 0.6),
    "hc24": lambda: mdl.holstein_honeycomb(24, 0.5),
    # per-bond engines (SSH couplings: every bond of every slice has its own cosh / sinh, held in registers)
    "bssh16": lambda: mdl.bssh_square(16, 16, 2.0),         # cfg3's lattice
    "ossh64": lambda: mdl.ossh_chain(64, 2.0),              # cfg2's lattice
    "ossh128": lambda: mdl.ossh_chain(128, 1.5),
    "ossh256": lambda: mdl.ossh_chain(256, 1.0),
}


def setup_square(name, seed=0):
    from smoqyelph_b200 import api
    m = SQUARES[name]()
    rng = np.random.default_rng(seed)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    ref = orc.RefFDM(m, sym=True)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m, sym=True)
    fdm.update(V, t)
    return m, rng, ref, fdm


@pytest.mark.parametrize("name", list(SQUARES))
def test_register_path_products(name):
    """One warp per slice-part, colour sweeps in registers: same bits as the shared-memory kernels, 1e-12 vs the oracle."""
    m, rng, ref, fdm = setup_square(name)
    v = rand_cvec(rng, m)
    fdm.set_fast_path(1)
    base = {op: getattr(fdm, op)(v) for op in ("mul_M", "mul_Mt", "mul_MtM", "mul_MMt")}
    for S in (1, 2, 3, 4, 7):
        fdm.set_fast_path(2 + 256 * S)
        assert fdm.tuning["path"] == 3, (name, S, fdm.tuning)
        for op, want in base.items():
            got = getattr(fdm, op)(v)
            assert relerr(got, getattr(ref, op)(v)) < RTOL, (name, S, op)
            assert np.array_equal(got, want), (name, S, op, "register path differs from the shared-memory kernel")


def test_register_path_requires_uniform_colours_and_canonical_order():
    """Disordered hoppings on a lattice without a per-bond engine, or a permuted colour order, must fall back to the shared-memory
    kernels (and stay correct); on 16 x 16 the disordered operator runs on the per-bond register engine, bit-identical."""
    from smoqyelph_b200 import api
    rng = np.random.default_rng(3)
    for (lx, ly, want3) in ((32, 32, False), (16, 16, True)):
        m = mdl.holstein_square(lx, ly, 0.5)
        V, t = dr.build_Vt(m, m.random_fields(rng))
        t = t * (1.0 + 0.1 * rng.standard_normal(t.shape[0]))[:, None]          # bond disorder: colours no longer uniform
        ref = orc.RefFDM(m, sym=True)
        ref.update(V, t)
        fdm = api.FermionDetMatrix(m, sym=True)
        fdm.update(V, t)
        v = rand_cvec(rng, m)
        fdm.set_fast_path(1)
        base = fdm.mul_MtM(v)
        fdm.set_fast_path(2)
        assert (fdm.tuning["path"] == 3) == want3, (lx, ly, fdm.tuning)
        got = fdm.mul_MtM(v)
        assert relerr(got, ref.mul_MtM(v)) < RTOL
        assert np.array_equal(got, base)
    m = mdl.holstein_square(16, 16, 0.5)
    v = rand_cvec(rng, m)
    # y colours before x colours: a valid checkerboard, but not the order the register kernels are written for
    nt, col, _ = mdl._square_bonds(16, 16)
    m2 = mdl.holstein_square(16, 16, 0.5)
    m2.finalize((np.asarray(col) + 2) % 4)
    V2, t2 = dr.build_Vt(m2, m2.random_fields(rng))
    ref2 = orc.RefFDM(m2, sym=True)
    ref2.update(V2, t2)
    fdm2 = api.FermionDetMatrix(m2, sym=True)
    fdm2.update(V2, t2)
    fdm2.set_fast_path(2)
    assert fdm2.tuning["path"] != 3
    assert relerr(fdm2.mul_MtM(v), ref2.mul_MtM(v)) < RTOL


@pytest.mark.parametrize("solver", ["resident", "launches", "launches_tma"])
@pytest.mark.parametrize("name", ["h16x16", "h16x16odd", "h32x16", "h32x32", "hc8", "hc24", "bssh16", "ossh64", "ossh256"])
def test_register_path_cg(name, solver, monkeypatch):
    """CG on the register path in native order: the whole-solve resident kernel (one grid-wide sum per iteration) and the
    two-launches-per-iteration loop (its only fallback) against the oracle's CG."""
    m, rng, ref, fdm = setup_square(name)
    b = rand_cvec(rng, m)
    fdm.set_fast_path(2 + 256 * 3)
    if solver == "launches":
        monkeypatch.setenv("SQ_NO_PERSISTENT_CG", "1")
    elif solver == "launches_tma":                  # first matvec of every solve with the operands staged by bulk async copies
        monkeypatch.setenv("SQ_NO_PERSISTENT_CG", "1")
        monkeypatch.setenv("SQ_V3_PRE", "2")
    st0 = fdm.stats
    xr, itr, epsr = ref.cg(b, tol=1e-14, maxiter=20000)
    xg, itg, epsg = fdm.ldiv(b, tol=1e-14, maxiter=20000)
    assert epsg < 1e-14 and epsr < 1e-14
    assert relerr(xg, xr) < 1e-11
    for tol in (1e-5, 1e-10):
        _, itr, _ = ref.cg(b, tol=tol, maxiter=20000)
        xg, itg, epsg = fdm.ldiv(b, tol=tol, maxiter=20000)
        assert abs(itg - itr) <= 1, (name, solver, tol, itg, itr)
        assert epsg < tol
        assert abs(relerr(ref.mul_MtM(xg), b) - epsg) < 1e-3 * tol + 1e-13       # the returned eps is the true residual
    _, it0, _ = fdm.ldiv(b, x0=xg, tol=1e-9)
    assert it0 == 0
    _, itc, epsc = fdm.ldiv(b, tol=1e-14, maxiter=3)
    assert itc == 3 and epsc > 1e-14
    # warm start from a perturbed solution converges to the same answer
    x2, it2, _ = fdm.ldiv(b, x0=xg * (1 + 1e-3), tol=1e-12)
    assert relerr(x2, xr) < 1e-9 and it2 > 0
    # the solver that was asked for is the one that ran (sq_fdm_stats), and no watchdog fired
    st = fdm.stats
    key = "cg_resident" if solver == "resident" else "cg_launch_loop"
    assert st[key] - st0[key] >= 5, (solver, st0, st)
    assert st["watchdog_aborts"] == 0 and st["instabilities"] == 0
    # maxiter = 0: no iteration, the initial residual is reported (not stale partial sums of an earlier solve)
    _, it00, eps00 = fdm.ldiv(b, tol=1e-10, maxiter=0)
    assert it00 == 0 and abs(eps00 - 1.0) < 1e-12


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "cfg5"])
def test_cg_iteration_counts_at_named_sizes(name):
    """CG iteration counts within +-1 of the reference recurrence (src/IterativeSolvers/ConjugateGradient.jl:93-167) at the FULL size of
    the named configurations 3, 4, 5 and at the production tolerances 1e-5 / 1e-10, on tau-smooth synthetic fields.
    All three run the whole-solve resident register kernel (cfg4 / cfg5: uniform engines, cfg3: per-bond engine)."""
    from smoqyelph_b200 import api
    m = mdl.config(name)
    rng = np.random.default_rng(11)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    ref = orc.RefFDM(m, sym=True, omp=True)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m, sym=True)
    fdm.update(V, t)
    b = rand_cvec(rng, m)
    st0 = fdm.stats
    for tol in (1e-5, 1e-10):
        xr, itr, epsr = ref.cg(b, tol=tol, maxiter=20000)
        xg, itg, epsg = fdm.ldiv(b, tol=tol, maxiter=20000)
        assert abs(itg - itr) <= 1, (name, tol, itg, itr)
        assert epsg < tol and epsr < tol
        assert relerr(xg, xr) < 50 * tol                     # both inside the tolerance ball (cond(M^T M) ~ 10 on these fields)
    st = fdm.stats
    assert st["cg_resident"] - st0["cg_resident"] == 2, st       # cfg3: per-bond register engine
    assert st["watchdog_aborts"] == 0


@pytest.mark.parametrize("name", ["h16x16", "h32x32", "h32x64", "hc8", "hc24"])
def test_register_path_tma_staging_is_bit_identical(name, monkeypatch):
    """The stand-alone native-order matvec with its late operands staged in shared memory by bulk async copies (the variant used
    once the vectors stream from HBM) computes exactly what the plain-load kernel computes, for every slab size."""
    import torch
    m, rng, ref, fdm = setup_square(name)
    n = m.N * m.Ltau
    d_in = torch.randn(n, 2, dtype=torch.float64, device="cuda")
    outs = {}
    for S in (1, 2, 3, 5):
        fdm.set_fast_path(2 + 256 * S)
        if fdm.tuning["path"] != 3:
            continue
        for pre in ("0", "2"):
            monkeypatch.setenv("SQ_V3_PRE", pre)
            d_out = torch.zeros_like(d_in)
            fdm.time_mul(102, d_out.data_ptr(), d_in.data_ptr(), 2)
            torch.cuda.synchronize()
            outs[(S, pre)] = d_out
        assert torch.equal(outs[(S, "0")], outs[(S, "2")]), (name, S)
        assert float(outs[(S, "0")].abs().max()) > 0
    assert outs
