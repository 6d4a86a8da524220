"""ctypes front-end of the CPU oracle (oracle/ref_c.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py; never by the product package.  Parity status: unpinned by
upstream (no golden vectors exist); pinned by tests/test_oracle_kat.py.

Array conventions are the reference's (Julia column-major): space-time vectors are numpy arrays of
shape (Ltau, N) in Fortran order (tau fastest); V (N, Ltau), t (Nh, Ltau), x (Nph, Ltau) Fortran.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_i64 = C.c_int64
c_dbl = C.c_double
c_vp = C.c_void_p


def build(force=False):
    """Compile ref_c.c (serial and OpenMP flavours) into oracle/_build/."""
    src = os.path.join(_HERE, "ref_c.c")
    outs = [os.path.join(_HERE, "_build", n) for n in ("libref_c.so", "libref_c_omp.so")]
    if force or any((not os.path.exists(o)) or os.path.getmtime(o) < os.path.getmtime(src) for o in outs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])


def _ptr(a):
    return a.ctypes.data_as(c_vp) if a is not None else None


def lib(omp=False):
    key = bool(omp)
    if key in _LIBS:
        return _LIBS[key]
    build()
    L = C.CDLL(os.path.join(_HERE, "_build", "libref_c_omp.so" if omp else "libref_c.so"))
    sig = {
        "ref_num_threads": (C.c_int, []),
        "ref_set_num_threads": (None, [C.c_int]),
        "ref_fdm_create": (c_vp, [C.c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_dbl, c_i64]),
        "ref_fdm_destroy": (None, [c_vp]),
        "ref_fdm_update": (None, [c_vp, c_vp, c_vp, c_dbl]),
        "ref_fdm_expV": (c_vp, [c_vp]), "ref_fdm_cosh": (c_vp, [c_vp]), "ref_fdm_sinh": (c_vp, [c_vp]),
        "ref_chk_lmul": (None, [c_vp, c_vp, C.c_int, c_i64, c_i64]),
        "ref_chk_ldiv": (None, [c_vp, c_vp, C.c_int, c_i64, c_i64]),
        "ref_mul_M": (None, [c_vp, c_vp, c_vp]), "ref_mul_Mt": (None, [c_vp, c_vp, c_vp]),
        "ref_mul_MtM": (None, [c_vp, c_vp, c_vp]), "ref_mul_MMt": (None, [c_vp, c_vp, c_vp]),
        "ref_kpm_create": (c_vp, [c_vp, c_dbl, c_i64, c_dbl, c_dbl]), "ref_kpm_destroy": (None, [c_vp]),
        "ref_kpm_update": (None, [c_vp, c_vp]), "ref_kpm_set_bounds": (None, [c_vp, c_dbl, c_dbl]),
        "ref_kpm_refresh_Bbar": (None, [c_vp]),
        "ref_kpm_active": (C.c_int, [c_vp]), "ref_kpm_get_bounds": (None, [c_vp, c_vp]),
        "ref_kpm_ncoef": (c_i64, [c_vp]), "ref_kpm_get_orders": (None, [c_vp, c_vp]),
        "ref_kpm_get_coefs": (None, [c_vp, c_i64, c_vp]), "ref_kpm_lanczos": (None, [c_vp, c_vp, c_vp]),
        "ref_kpm_bbar_mul": (None, [c_vp, c_vp, c_vp]),
        "ref_fourier": (None, [c_vp, c_vp, C.c_int]), "ref_kpm_ldiv": (None, [c_vp, c_vp, c_vp]),
        "ref_cg": (c_i64, [c_vp, c_vp, c_vp, C.c_int, c_vp, c_dbl, c_i64, c_vp]),
        "ref_elph_create": (c_vp, [c_i64, c_i64, c_i64, c_i64, c_dbl, c_vp, c_vp, c_vp,
                                   c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                   c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "ref_elph_destroy": (None, [c_vp]), "ref_elph_x": (c_vp, [c_vp]), "ref_elph_V": (c_vp, [c_vp]),
        "ref_elph_t": (c_vp, [c_vp]), "ref_elph_shift_mu": (None, [c_vp, c_dbl]),
        "ref_elph_build_Vt": (None, [c_vp]), "ref_elph_refresh": (None, [c_vp, c_vp]),
        "ref_update_Lambda": (None, [c_vp, c_vp]),
        "ref_mul_Lambda": (None, [c_vp, c_vp, c_vp, c_i64, c_i64]), "ref_ldiv_Lambda": (None, [c_vp, c_vp, c_vp, c_i64, c_i64]),
        "ref_mul_LambdaT": (None, [c_vp, c_vp, c_vp, c_i64, c_i64]), "ref_ldiv_LambdaT": (None, [c_vp, c_vp, c_vp, c_i64, c_i64]),
        "ref_mul_nuRe_dLambda_dx": (None, [c_vp, c_dbl, c_vp, c_vp, c_vp, c_vp]),
        "ref_mul_nuRe_dM_dx": (None, [c_vp, c_dbl, c_vp, c_vp, c_vp, c_vp, C.c_int]),
        "ref_pff_create": (c_vp, [c_i64, c_i64]), "ref_pff_destroy": (None, [c_vp]),
        "ref_pff_Phi": (c_vp, [c_vp]), "ref_pff_Psi": (c_vp, [c_vp]), "ref_pff_Lambda": (c_vp, [c_vp]),
        "ref_pff_set_exact_holstein": (None, [c_vp, C.c_int]),
        "ref_pff_sample": (c_dbl, [c_vp, c_vp, c_vp, c_vp]),
        "ref_pff_action": (c_dbl, [c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_i64, c_vp, c_vp, c_vp]),
        "ref_pff_force": (c_dbl, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_i64, c_vp, c_vp]),
        "ref_efa_create": (c_vp, [c_vp, c_dbl]), "ref_efa_destroy": (None, [c_vp]),
        "ref_efa_init_momentum": (c_dbl, [c_vp, c_vp, c_vp, c_vp]), "ref_efa_kinetic": (c_dbl, [c_vp, c_vp, c_vp]),
        "ref_efa_evolve": (None, [c_vp, c_vp, c_vp, c_vp, c_dbl]),
        "ref_bosonic_action": (c_dbl, [c_vp]), "ref_anharmonic_derivative": (None, [c_vp, c_vp]),
        "ref_dispersive_derivative": (None, [c_vp, c_vp]), "ref_elph_set_dispersion": (None, [c_vp, c_i64, c_vp, c_vp, c_vp]),
        "ref_hmc_update": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_i64, c_vp, c_vp]),
        "ref_greens_update": (c_dbl, [c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_i64]),
        "ref_measure_n": (None, [c_vp, c_vp, c_i64, c_i64, c_vp]),
        "ref_measure_double_occ": (None, [c_vp, c_vp, c_i64, c_i64, c_vp]),
        "ref_measure_Nsqrd": (None, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIBS[key] = L
    return L


def cvec(model, a=None):
    """(Ltau, N) complex128 Fortran array (zeros, or a copy of a)."""
    out = np.zeros((model.Ltau, model.N), np.complex128, order="F")
    if a is not None:
        out[...] = np.asarray(a).reshape(out.shape, order="F")
    return out


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape, order="F")


class RefFDM:
    """FermionDetMatrix of the oracle (src/FermionDetMatrix.jl)."""

    def __init__(self, model, sym=True, tol=1e-6, maxiter=None, omp=False):
        self.L = lib(omp)
        self.model, self.sym = model, bool(sym)
        m = model
        nt = np.ascontiguousarray(m.nt_chk.T.astype(np.int64))       # (Nh, 2) row-major == (2, Nh) column-major
        perm = np.ascontiguousarray(m.perm.astype(np.int64))
        clo = np.array([c[0] for c in m.colors], np.int64)
        chi = np.array([c[1] for c in m.colors], np.int64)
        self.h = self.L.ref_fdm_create(int(sym), m.Ltau, m.N, m.Nh, _ptr(nt), _ptr(perm), len(m.colors), _ptr(clo), _ptr(chi),
                                       tol, maxiter if maxiter is not None else m.N * m.Ltau)
        self.tol, self.maxiter = tol, maxiter if maxiter is not None else m.N * m.Ltau

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_fdm_destroy(self.h)
            self.h = None

    def update(self, V, t):
        V = np.asfortranarray(V, dtype=np.float64)
        t = np.asfortranarray(t, dtype=np.float64)
        self.L.ref_fdm_update(self.h, _ptr(V), _ptr(t), self.model.dtau)

    @property
    def expV(self):
        return _view(self.L.ref_fdm_expV(self.h), (self.model.Ltau, self.model.N), np.float64)

    @property
    def cosh(self):
        return _view(self.L.ref_fdm_cosh(self.h), (self.model.Ltau, self.model.Nh), np.float64)

    @property
    def sinh(self):
        return _view(self.L.ref_fdm_sinh(self.h), (self.model.Ltau, self.model.Nh), np.float64)

    def _mul(self, fn, v):
        v = cvec(self.model, v)
        out = cvec(self.model)
        fn(self.h, _ptr(out), _ptr(v))
        return out

    def mul_M(self, v): return self._mul(self.L.ref_mul_M, v)
    def mul_Mt(self, v): return self._mul(self.L.ref_mul_Mt, v)
    def mul_MtM(self, v): return self._mul(self.L.ref_mul_MtM, v)
    def mul_MMt(self, v): return self._mul(self.L.ref_mul_MMt, v)

    def chk(self, v, transposed=False, inverse=False, color=None):
        v = cvec(self.model, v)
        lo, hi = (0, self.model.Nh) if color is None else self.model.colors[color]
        (self.L.ref_chk_ldiv if inverse else self.L.ref_chk_lmul)(self.h, _ptr(v), int(transposed), lo, hi)
        return v

    def cg(self, b, x0=None, P=None, tol=None, maxiter=None):
        """ldiv!(x, fdm, b; preconditioner) without the preconditioner refresh.  x0=None <=> x === b."""
        b = cvec(self.model, b)
        same = x0 is None
        x = cvec(self.model, b if same else x0)
        eps = c_dbl(0)
        it = self.L.ref_cg(self.h, _ptr(x), _ptr(x if same else b), int(same), P.h if P is not None else None,
                           self.tol if tol is None else tol, self.maxiter if maxiter is None else maxiter, C.byref(eps))
        return x, int(it), eps.value


class RefKPM:
    """KPMPreconditioner of the oracle (src/KPMPreconditioner.jl)."""

    def __init__(self, fdm, rbuf=0.10, n=20, a1=1.0, a2=1.0):
        self.L, self.fdm = fdm.L, fdm
        self.h = self.L.ref_kpm_create(fdm.h, rbuf, n, a1, a2)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_kpm_destroy(self.h)
            self.h = None

    def update(self, start):
        start = np.ascontiguousarray(start, np.float64)
        self.L.ref_kpm_update(self.h, _ptr(start))

    def set_bounds(self, emin, emax): self.L.ref_kpm_set_bounds(self.h, emin, emax)
    def refresh_Bbar(self): self.L.ref_kpm_refresh_Bbar(self.h)

    @property
    def active(self): return bool(self.L.ref_kpm_active(self.h))

    @property
    def bounds(self):
        b = np.zeros(2)
        self.L.ref_kpm_get_bounds(self.h, _ptr(b))
        return b

    @property
    def orders(self):
        o = np.zeros(self.L.ref_kpm_ncoef(self.h), np.int64)
        self.L.ref_kpm_get_orders(self.h, _ptr(o))
        return o

    def coefs(self, l):
        c = np.zeros(int(self.orders[l]), np.complex128)
        self.L.ref_kpm_get_coefs(self.h, l, _ptr(c))
        return c

    def lanczos(self, start):
        b = np.zeros(2)
        start = np.ascontiguousarray(start, np.float64)
        self.L.ref_kpm_lanczos(self.h, _ptr(start), _ptr(b))
        return b

    def bbar_mul(self, v):
        v = np.ascontiguousarray(v, np.complex128)
        out = np.zeros_like(v)
        self.L.ref_kpm_bbar_mul(self.h, _ptr(out), _ptr(v))
        return out

    def fourier(self, v, forward=True):
        v = cvec(self.fdm.model, v)
        self.L.ref_fourier(self.h, _ptr(v), int(forward))
        return v

    def ldiv(self, v):
        v = cvec(self.fdm.model, v)
        out = cvec(self.fdm.model)
        self.L.ref_kpm_ldiv(self.h, _ptr(out), _ptr(v))
        return out


class RefElPh:
    """The ElectronPhononParameters / FermionPathIntegral fields the path reads."""

    def __init__(self, model, omp=False):
        self.L, self.model = lib(omp), model
        m = model
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        i64 = lambda a: np.ascontiguousarray(a, np.int64)
        keep = [f64(m.Omega), f64(m.Omega4), f64(m.Mass), i64(m.hol_phonon), i64(m.hol_site),
                f64(m.hol_alpha[0]), f64(m.hol_alpha[1]), f64(m.hol_alpha[2]), f64(m.hol_alpha[3]),
                np.ascontiguousarray(m.hol_phsym, np.int32),
                i64(m.ssh_phonon.T), i64(m.ssh_hopping),
                f64(m.ssh_alpha[0]), f64(m.ssh_alpha[1]), f64(m.ssh_alpha[2]), f64(m.ssh_alpha[3]), f64(m.V0), f64(m.t0)]
        k = keep
        self.h = self.L.ref_elph_create(m.Ltau, m.N, m.Nh, m.Nph, m.dtau, _ptr(k[0]), _ptr(k[1]), _ptr(k[2]),
                                        m.Nhol, _ptr(k[3]), _ptr(k[4]), _ptr(k[5]), _ptr(k[6]), _ptr(k[7]), _ptr(k[8]), _ptr(k[9]),
                                        m.Nssh, _ptr(k[10]), _ptr(k[11]), _ptr(k[12]), _ptr(k[13]), _ptr(k[14]), _ptr(k[15]),
                                        _ptr(k[16]), _ptr(k[17]))
        if getattr(m, "Ndisp", 0):
            dp, do, do4 = i64(m.disp_phonon.T), f64(m.disp_Omega), f64(m.disp_Omega4)
            self.L.ref_elph_set_dispersion(self.h, m.Ndisp, _ptr(dp), _ptr(do), _ptr(do4))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_elph_destroy(self.h)
            self.h = None

    @property
    def x(self): return _view(self.L.ref_elph_x(self.h), (self.model.Nph, self.model.Ltau), np.float64)
    @property
    def V(self): return _view(self.L.ref_elph_V(self.h), (self.model.N, self.model.Ltau), np.float64)
    @property
    def t(self): return _view(self.L.ref_elph_t(self.h), (self.model.Nh, self.model.Ltau), np.float64)

    def set_x(self, x): self.x[...] = x
    def shift_mu(self, dmu): self.L.ref_elph_shift_mu(self.h, dmu)
    def build_Vt(self): self.L.ref_elph_build_Vt(self.h)
    def refresh(self, fdm): self.L.ref_elph_refresh(self.h, fdm.h)

    def Lambda(self):
        out = np.zeros((self.model.Ltau, self.model.N), order="F")
        self.L.ref_update_Lambda(_ptr(out), self.h)
        return out

    def lam_op(self, which, Lam, v):
        fn = {"mul": self.L.ref_mul_Lambda, "ldiv": self.L.ref_ldiv_Lambda, "mulT": self.L.ref_mul_LambdaT,
              "ldivT": self.L.ref_ldiv_LambdaT}[which]
        v = cvec(self.model, v)
        out = cvec(self.model)
        fn(_ptr(out), _ptr(np.asfortranarray(Lam)), _ptr(v), self.model.Ltau, self.model.N)
        return out

    def dLambda_dx(self, nu, up, u, Lam):
        F = np.zeros((self.model.Nph, self.model.Ltau), order="F")
        up, u = cvec(self.model, up), cvec(self.model, u)
        self.L.ref_mul_nuRe_dLambda_dx(_ptr(F), nu, _ptr(up), _ptr(u), _ptr(np.asfortranarray(Lam)), self.h)
        return F

    def dM_dx(self, nu, u, v, fdm, exact_holstein=False):
        F = np.zeros((self.model.Nph, self.model.Ltau), order="F")
        u, v = cvec(self.model, u), cvec(self.model, v)
        self.L.ref_mul_nuRe_dM_dx(_ptr(F), nu, _ptr(u), _ptr(v), fdm.h, self.h, int(exact_holstein))
        return F

    def bosonic_action(self): return self.L.ref_bosonic_action(self.h)

    def potential_derivative(self):
        """Anharmonic + dispersive parts of dS_b/dx (the terms the leapfrog kick adds to the fermionic force, EFAPFFHMCUpdater.jl:190-193)."""
        F = np.zeros((self.model.Nph, self.model.Ltau), order="F")
        self.L.ref_anharmonic_derivative(_ptr(F), self.h)
        self.L.ref_dispersive_derivative(_ptr(F), self.h)
        return F


class RefPFF:
    """PFFCalculator of the oracle (src/PFFCalculator.jl)."""

    def __init__(self, elph, fdm, exact_holstein=False):
        self.L, self.elph, self.fdm = elph.L, elph, fdm
        self.h = self.L.ref_pff_create(fdm.model.Ltau, fdm.model.N)
        self.L.ref_pff_set_exact_holstein(self.h, int(exact_holstein))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_pff_destroy(self.h)
            self.h = None

    @property
    def Phi(self): return _view(self.L.ref_pff_Phi(self.h), (self.fdm.model.Ltau, self.fdm.model.N), np.complex128)
    @property
    def Psi(self): return _view(self.L.ref_pff_Psi(self.h), (self.fdm.model.Ltau, self.fdm.model.N), np.complex128)

    def sample(self, R):
        R = cvec(self.fdm.model, R)
        return self.L.ref_pff_sample(self.h, self.elph.h, self.fdm.h, _ptr(R))

    def action(self, P=None, lanczos_start=None, tol=1e-10, maxiter=10000):
        it, eps, im = c_i64(0), c_dbl(0), c_dbl(0)
        ls = np.ascontiguousarray(lanczos_start, np.float64) if lanczos_start is not None else None
        Sf = self.L.ref_pff_action(self.h, self.elph.h, self.fdm.h, P.h if P is not None else None, _ptr(ls), tol, maxiter,
                                   C.byref(it), C.byref(eps), C.byref(im))
        return Sf, it.value, eps.value

    def force(self, P=None, lanczos_start=None, tol=1e-5, maxiter=10000):
        m = self.fdm.model
        F = np.zeros((m.Nph, m.Ltau), order="F")
        it, eps = c_i64(0), c_dbl(0)
        ls = np.ascontiguousarray(lanczos_start, np.float64) if lanczos_start is not None else None
        Sf = self.L.ref_pff_force(_ptr(F), self.h, self.elph.h, self.fdm.h, P.h if P is not None else None, _ptr(ls), tol, maxiter,
                                  C.byref(it), C.byref(eps))
        return F, Sf, it.value, eps.value


class RefEFA:
    def __init__(self, elph, eta=0.0):
        self.L, self.elph = elph.L, elph
        self.h = self.L.ref_efa_create(elph.h, eta)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_efa_destroy(self.h)
            self.h = None

    def init_momentum(self, R):
        m = self.elph.model
        p = np.zeros((m.Nph, m.Ltau), order="F")
        R = np.asfortranarray(R, dtype=np.float64)
        K = self.L.ref_efa_init_momentum(self.h, self.elph.h, _ptr(p), _ptr(R))
        return p, K

    def kinetic(self, p):
        p = np.asfortranarray(p, dtype=np.float64)
        return self.L.ref_efa_kinetic(self.h, self.elph.h, _ptr(p))

    def evolve(self, x, p, dt):
        x = np.asfortranarray(x, dtype=np.float64).copy(order="F")
        p = np.asfortranarray(p, dtype=np.float64).copy(order="F")
        self.L.ref_efa_evolve(self.h, self.elph.h, _ptr(x), _ptr(p), dt)
        return x, p


def hmc_random_count(model, Nt, preconditioned):
    return 1 + 2 * model.Ltau * model.N + model.Nph * model.Ltau + ((Nt + 1) * model.N if preconditioned else 0) + 1


def hmc_update(elph, fdm, pff, efa, P, Nt, dt, delta, tol_action, tol_force, maxiter, rnd):
    """hmc_update! (src/EFAPFFHMCUpdater.jl:102-279).  Returns (accepted, info[8])."""
    rnd = np.ascontiguousarray(rnd, np.float64)
    assert rnd.size >= hmc_random_count(elph.model, Nt, P is not None)
    out = np.zeros(8)
    acc = elph.L.ref_hmc_update(elph.h, fdm.h, pff.h, P.h if P is not None else None, efa.h, Nt, dt, delta,
                                tol_action, tol_force, maxiter, _ptr(rnd), _ptr(out))
    return bool(acc), out


def greens_update(fdm, P, R, GR, tol, maxiter):
    """update_greens_estimator! solves.  R, GR: (V, Nrv) Fortran complex; GR updated in place."""
    V, Nrv = R.shape
    R = np.asfortranarray(R, np.complex128)
    assert GR.flags.f_contiguous and GR.dtype == np.complex128
    return fdm.L.ref_greens_update(fdm.h, P.h if P is not None else None, _ptr(R), _ptr(GR), Nrv, tol, maxiter)


def measure(which, R, GR, Ltau=None, omp=False):
    L = lib(omp)
    V, Nrv = R.shape
    R = np.asfortranarray(R, np.complex128)
    GR = np.asfortranarray(GR, np.complex128)
    out = np.zeros(2)
    if which == "n":
        L.ref_measure_n(_ptr(R), _ptr(GR), V, Nrv, _ptr(out))
    elif which == "double_occ":
        L.ref_measure_double_occ(_ptr(R), _ptr(GR), V, Nrv, _ptr(out))
    elif which == "Nsqrd":
        L.ref_measure_Nsqrd(_ptr(R), _ptr(GR), V, Ltau, Nrv, _ptr(out))
    else:
        raise KeyError(which)
    return complex(out[0], out[1])
