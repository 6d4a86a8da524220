"""Reader / checker of the reference fixtures written by tools/dump_reference_fixtures.jl (tests/golden/ref_<tag>/: raw little-endian
column-major arrays + manifest.txt), and a writer of the same format from the CPU oracle (used to test the pipeline itself)."""
import glob
import os

import numpy as np

from smoqyelph_b200 import model as mdl

HERE = os.path.dirname(os.path.abspath(__file__))
DTYPES = {"c16": np.complex128, "f8": np.float64, "i8": np.int64, "u1": np.uint8}


def fixture_dirs():
    return sorted(d for d in glob.glob(os.path.join(HERE, "golden", "ref_*")) if os.path.exists(os.path.join(d, "manifest.txt")))


def load(d):
    out = {}
    for line in open(os.path.join(d, "manifest.txt")):
        p = line.split()
        if not p:
            continue
        if p[1] == "scalar":
            out[p[0]] = float(p[2])
        else:
            shape = tuple(int(q) for q in p[2:])
            a = np.fromfile(os.path.join(d, p[0] + ".bin"), dtype=DTYPES[p[1]])
            out[p[0]] = a.reshape(shape, order="F")
    return out


def model_of(fx, name):
    """Model with Julia's own checkerboard decomposition (the colour order changes M at O(dtau^2): never recomputed here)."""
    N, Nh = int(fx["N"]), fx["nt"].shape[1]
    hol_alpha = fx["hol_alpha"].reshape(-1, 4, order="F").T if fx["hol_alpha"].size else np.zeros((4, 0))
    ssh_alpha = fx["ssh_alpha"].reshape(-1, 4, order="F").T if fx["ssh_alpha"].size else np.zeros((4, 0))
    m = mdl.Model(name, fx["beta"], fx["dtau"], N, (fx["neighbor_table"] - 1).astype(np.int64), fx["t_bare"].copy(),
                  fx["eps_bare"] - fx["mu"], fx["Omega"].copy(), fx["Omega4"].copy(), fx["M"].copy(),
                  hol_phonon=(fx["hol_phonon"] - 1).astype(np.int64), hol_site=(fx["hol_site"] - 1).astype(np.int64),
                  hol_alpha=np.ascontiguousarray(hol_alpha), hol_phsym=fx["hol_phsym"].astype(np.int32),
                  ssh_phonon=(fx["ssh_phonon"].reshape(2, -1, order="F") - 1).astype(np.int64), ssh_hopping=(fx["ssh_hopping"] - 1).astype(np.int64),
                  ssh_alpha=np.ascontiguousarray(ssh_alpha))
    m.perm = (fx["perm"] - 1).astype(np.int64)
    m.nt_chk = np.ascontiguousarray((fx["nt"] - 1).astype(np.int64))
    m.colors = [(int(lo) - 1, int(hi)) for lo, hi in zip(fx["color_lo"], fx["color_hi"])]
    assert m.Ltau == int(fx["Ltau"]) and Nh == m.Nh
    return m


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def check(fx, backend, name="fixture"):
    """backend: 'oracle' (CPU restatement) or 'cuda' (the library through its C ABI).  Compares with the reference outputs:
    M^T M v / M v / M^T v to 1e-12, coefficients to 1e-14, CG solution to solver accuracy, iteration counts +-1, action and force."""
    m = model_of(fx, name)
    V, t, x = np.asfortranarray(fx["V"]), np.asfortranarray(fx["t"]), np.asfortranarray(fx["x"])
    if backend == "oracle":
        from oracle import oracle as orc
        f = orc.RefFDM(m, sym=True, tol=1e-10, maxiter=50000)
        f.update(V, t)
        e = orc.RefElPh(m)
        e.set_x(x)
        pff = orc.RefPFF(e, f)
        coef = (f.expV, f.cosh, f.sinh)
        cg = lambda b, tol: f.cg(b, tol=tol, maxiter=50000)
        sample = lambda R: pff.sample(R)
        phi = lambda: pff.Phi
        force = lambda tol: pff.force(tol=tol, maxiter=50000)
    else:
        from smoqyelph_b200 import api
        f = api.FermionDetMatrix(m, sym=True, tol=1e-10, maxiter=50000)
        f.update(V, t)
        e = api.ElectronPhononParameters(m, f)
        e.x = x
        pff = api.PFFCalculator(e)
        coef = f.coefficients()
        cg = lambda b, tol: f.ldiv(b, tol=tol, maxiter=50000)
        sample = lambda R: pff.sample_pseudofermion_fields(R)
        phi = lambda: pff.fields()[0]
        force = lambda tol: pff.calculate_derivative_fermionic_action(tol=tol, maxiter=50000)
    res = {}
    res["expV"] = relerr(coef[0], fx["expV"])
    res["cosh"] = relerr(coef[1], fx["cosh_t"])
    res["sinh"] = float(np.abs(coef[2] - fx["sinh_t"]).max())
    b = np.asfortranarray(fx["Phi_in"])
    for op, key in (("mul_MtM", "MtM_Phi"), ("mul_M", "M_Phi"), ("mul_Mt", "Mt_Phi")):
        res[op] = relerr(getattr(f, op)(b), fx[key])
    xs, it14, _ = cg(b, 1e-14)
    res["x_cg"] = relerr(xs, fx["x_cg"])
    res["iters"] = {}
    for tol, nm in ((1e-5, "5"), (1e-10, "10")):
        _, it, _ = cg(b, tol)
        res["iters"][nm] = (int(it), int(fx["cg_iters_" + nm]))
    res["Sf_sample"] = abs(sample(np.asfortranarray(fx["R"])) - fx["Sf_sample"]) / abs(fx["Sf_sample"])
    res["Phi"] = relerr(phi(), fx["Phi"])
    F, Sf, _, _ = force(1e-14)
    res["Sf"] = abs(Sf - fx["Sf"]) / abs(fx["Sf"])
    res["dSdx"] = relerr(F, fx["dSdx"])
    return res


def assert_parity(res):
    assert res["expV"] < 1e-14 and res["cosh"] < 1e-14 and res["sinh"] < 1e-14, res
    assert res["mul_MtM"] < 1e-12 and res["mul_M"] < 1e-12 and res["mul_Mt"] < 1e-12, res          # north_star: 1e-12
    assert res["x_cg"] < 1e-10, res                                                               # both converged to 1e-14: cond x 1e-14
    assert all(abs(a - b) <= 1 for a, b in res["iters"].values()), res                            # north_star: +-1
    assert res["Sf_sample"] < 1e-13 and res["Phi"] < 1e-13, res
    assert res["Sf"] < 1e-10 and res["dSdx"] < 1e-9, res                                          # force linear in Psi (solver accuracy)


def write_from_oracle(d, m, seed=0):
    """Same files as tools/dump_reference_fixtures.jl, produced by the CPU oracle: exercises reader + checker without Julia."""
    from oracle import oracle as orc
    import dense_ref as dr
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    x = m.random_fields(rng, smooth=True)
    V, t = dr.build_Vt(m, x)
    f = orc.RefFDM(m, sym=True, tol=1e-10, maxiter=50000)
    f.update(V, t)
    e = orc.RefElPh(m)
    e.set_x(x)
    pff = orc.RefPFF(e, f)
    lines = []

    def arr(name, a, dt):
        a = np.asfortranarray(np.asarray(a, DTYPES[dt]))
        a.ravel(order="F").tofile(os.path.join(d, name + ".bin"))
        lines.append(f"{name} {dt} " + " ".join(str(s) for s in a.shape))

    def sc(name, v):
        lines.append(f"{name} scalar {float(v)!r}")
    sc("beta", m.beta); sc("dtau", m.dtau); sc("Ltau", m.Ltau); sc("N", m.N); sc("sym", 1); sc("mu", 0.0); sc("nphonon", m.nphonon)
    arr("neighbor_table", m.neighbor_table + 1, "i8"); arr("nt", m.nt_chk + 1, "i8"); arr("perm", m.perm + 1, "i8")
    arr("color_lo", [c[0] + 1 for c in m.colors], "i8"); arr("color_hi", [c[1] for c in m.colors], "i8")
    arr("V", V, "f8"); arr("t", t, "f8"); arr("x", x, "f8"); arr("eps_bare", m.V0, "f8"); arr("t_bare", m.t0, "f8")
    arr("expV", f.expV, "f8"); arr("cosh_t", f.cosh, "f8"); arr("sinh_t", f.sinh, "f8")
    arr("Omega", m.Omega, "f8"); arr("Omega4", m.Omega4, "f8"); arr("M", m.Mass, "f8")
    arr("hol_phonon", m.hol_phonon + 1, "i8"); arr("hol_site", m.hol_site + 1, "i8"); arr("hol_alpha", m.hol_alpha.T, "f8")
    arr("hol_phsym", m.hol_phsym, "i8"); arr("ssh_phonon", m.ssh_phonon + 1, "i8"); arr("ssh_hopping", m.ssh_hopping + 1, "i8")
    arr("ssh_alpha", m.ssh_alpha.T, "f8")
    b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    arr("Phi_in", b, "c16"); arr("MtM_Phi", f.mul_MtM(b), "c16"); arr("M_Phi", f.mul_M(b), "c16"); arr("Mt_Phi", f.mul_Mt(b), "c16")
    for tol, nm in ((1e-5, "5"), (1e-10, "10"), (1e-14, "14")):
        xs, it, eps = f.cg(b, tol=tol, maxiter=50000)
        sc("cg_iters_" + nm, it); sc("cg_eps_" + nm, eps)
        if nm == "14":
            arr("x_cg", xs, "c16")
    R = np.asfortranarray((rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2))
    sc("Sf_sample", pff.sample(R)); arr("R", R, "c16"); arr("Phi", pff.Phi, "c16")
    F, Sf, it, _ = pff.force(tol=1e-14, maxiter=50000)
    arr("dSdx", F, "f8"); sc("Sf", Sf); sc("force_iters", it); arr("Lambda", e.Lambda(), "f8")
    open(os.path.join(d, "manifest.txt"), "w").write("\n".join(lines) + "\n")
