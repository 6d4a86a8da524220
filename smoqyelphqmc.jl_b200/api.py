"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the Julia package (docs/src/api.md of the
reference), so the parity tests read like the reference's own code:

    fdm  = SymFermionDetMatrix(model, maxiter=..., tol=...)      src/FermionDetMatrix.jl:66
    elph = ElectronPhononParameters(model, fdm); elph.x = x; elph.update_fdm()
    P    = KPMPreconditioner(fdm, rbuf=0.1, n=20, a1=1.0, a2=1.0)  src/KPMPreconditioner.jl:198
    pff  = PFFCalculator(elph, fdm)                               src/PFFCalculator.jl:30
    hmc  = EFAPFFHMCUpdater(elph, pff, Nt=..., dt=...)            src/EFAPFFHMCUpdater.jl:40
    accepted, iters = hmc.hmc_update(preconditioner=P, tol_action=..., tol_force=..., maxiter=...)

Every method is a thin call into the C ABI (lib.py); vectors are numpy arrays in the reference's
layout, (Ltau, N) complex128 Fortran order.  A failed library call raises `SqError`, which callers
treat like the reference's "numerical instability" exceptions (reject the update).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l
from .lib import SqError, SqNumericalInstability, check, ptr  # noqa: F401

OP_M, OP_MT, OP_MTM, OP_MMT = 0, 1, 2, 3


def _cvec(model, a=None):
    out = np.zeros((model.Ltau, model.N), np.complex128, order="F")
    if a is not None:
        out[...] = np.asarray(a).reshape(out.shape, order="F")
    return out


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def _i64(a, one_based=True):
    return np.ascontiguousarray(np.asarray(a, np.int64) + (1 if one_based else 0))


class FermionDetMatrix:
    """FermionDetMatrix{T,E}: matrix-free M with its CG workspace (src/FermionDetMatrix.jl:19-55)."""

    def __init__(self, model, sym=True, maxiter=None, tol=1e-6, device=0):
        self.L = _l.load()
        self.model, self.sym = model, bool(sym)
        self.tol = tol
        self.maxiter = int(maxiter if maxiter is not None else model.N * model.Ltau)
        nt = _i64(model.nt_chk.T)                       # (Nh, 2) C-order == (2, Nh) column-major
        perm = _i64(model.perm)
        clo = _i64([c[0] for c in model.colors])         # 1-based inclusive lower bound
        chi = _i64([c[1] - 1 for c in model.colors])     # 1-based inclusive upper bound
        h = C.c_void_p()
        check(self.L.sq_fdm_create(C.byref(h), int(sym), model.Ltau, model.N, model.Nh, ptr(nt), ptr(perm), len(model.colors),
                                   ptr(clo), ptr(chi), tol, self.maxiter, device))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.sq_fdm_destroy(self.h)
            self.h = None

    __del__ = close

    # update!(fdm, fpi): V (N, Ltau), t (Nh, Ltau) Fortran
    def update(self, V, t, dtau=None):
        V = np.asfortranarray(V, np.float64)
        t = np.asfortranarray(t, np.float64)
        check(self.L.sq_fdm_update(self.h, ptr(V), ptr(t), self.model.dtau if dtau is None else dtau))

    def _mul(self, op, v):
        v = _cvec(self.model, v)
        out = _cvec(self.model)
        check(self.L.sq_fdm_mul(self.h, op, ptr(out), ptr(v)))
        return out

    def mul_M(self, v): return self._mul(OP_M, v)
    def mul_Mt(self, v): return self._mul(OP_MT, v)
    def mul_MtM(self, v): return self._mul(OP_MTM, v)
    def mul_MMt(self, v): return self._mul(OP_MMT, v)
    mul = mul_MtM                                        # mul!(v', fdm, v) = M^T M v  (:304-315)

    def ldiv(self, b, x0=None, preconditioner=None, tol=None, maxiter=None, lanczos_start=None, refresh=True):
        """ldiv!(x, fdm, b; preconditioner, tol, maxiter) -> (x, iters, eps).  x0=None <=> x === b."""
        b = _cvec(self.model, b)
        zero = x0 is None
        x = _cvec(self.model, None if zero else x0)
        it, eps = C.c_int64(0), C.c_double(0)
        ls = None if lanczos_start is None else _f64(lanczos_start)
        check(self.L.sq_fdm_cg(self.h, ptr(x), ptr(b), int(zero), preconditioner.h if preconditioner is not None else None,
                               int(refresh and preconditioner is not None), ptr(ls), self.tol if tol is None else tol,
                               self.maxiter if maxiter is None else int(maxiter), C.byref(it), C.byref(eps)))
        return x, it.value, eps.value

    def ldiv_batch(self, B, X0=None, preconditioner=None, tol=None, maxiter=None, lanczos_start=None, refresh=True):
        """nrhs systems at once: B (V, nrhs) complex Fortran (the GreensEstimator's layout) -> (X, iters[nrhs], eps[nrhs]).  Same recurrence,
        warm start (X0) and iteration count per system as ldiv; batched over the systems when a preconditioner is active."""
        V = self.model.Ltau * self.model.N
        B = np.asfortranarray(B, np.complex128)
        assert B.shape[0] == V
        nrhs = B.shape[1]
        zero = X0 is None
        X = np.zeros((V, nrhs), np.complex128, order="F") if zero else np.array(X0, np.complex128, order="F", copy=True)
        it, eps = np.zeros(nrhs, np.int64), np.zeros(nrhs)
        ls = None if lanczos_start is None else _f64(lanczos_start)
        check(self.L.sq_fdm_cg_batch(self.h, ptr(X), ptr(B), nrhs, int(zero), preconditioner.h if preconditioner is not None else None,
                                     int(refresh and preconditioner is not None), ptr(ls), self.tol if tol is None else tol,
                                     self.maxiter if maxiter is None else int(maxiter), ptr(it), ptr(eps)))
        return X, it, eps

    def coefficients(self):
        m = self.model
        e = np.zeros((m.Ltau, m.N), order="F")
        c = np.zeros((m.Ltau, m.Nh), order="F")
        s = np.zeros((m.Ltau, m.Nh), order="F")
        check(self.L.sq_fdm_get_coefficients(self.h, ptr(e), ptr(c), ptr(s)))
        return e, c, s

    # device-resident entry points (internal [l][i] layout; raw device addresses, e.g. torch data_ptr())
    def mul_dev(self, op, d_out, d_in): check(self.L.sq_fdm_mul_dev(self.h, op, ptr(d_out), ptr(d_in)))

    def cg_dev(self, d_x, d_b, zero_start=True, preconditioner=None, tol=None, maxiter=None):
        it, eps = C.c_int64(0), C.c_double(0)
        check(self.L.sq_fdm_cg_dev(self.h, ptr(d_x), ptr(d_b), int(zero_start), preconditioner.h if preconditioner is not None else None,
                                   self.tol if tol is None else tol, self.maxiter if maxiter is None else int(maxiter),
                                   C.byref(it), C.byref(eps)))
        return it.value, eps.value

    @property
    def tuning(self):
        s, t, p = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.L.sq_fdm_get_tuning(self.h, C.byref(s), C.byref(t), C.byref(p)))
        return {"slab": s.value, "threads": t.value, "path": p.value}

    def set_tuning(self, slab, threads): check(self.L.sq_fdm_set_tuning(self.h, slab, threads))
    def set_fast_path(self, enable): check(self.L.sq_fdm_set_fast_path(self.h, int(enable)))

    def time_mul(self, op, d_out, d_in, reps=200, d_flush=None, flush_bytes=0):
        """Device time per launch in microseconds (CUDA events inside the library): back to back (L2-hot), or with a
        flush_bytes write to d_flush before every launch (L2-cold, median)."""
        us = C.c_double(0)
        check(self.L.sq_fdm_time_mul(self.h, op, ptr(d_out), ptr(d_in), reps, ptr(d_flush) if d_flush else None, flush_bytes, C.byref(us)))
        return us.value

    # ---- tau-slab partitioning (multi-GPU) ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_l.load().sq_nccl_unique_id(buf))
        return buf.raw

    def init_slab(self, rank, world, unique_id=None):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(self.L.sq_fdm_init_slab(self.h, int(rank), int(world), buf))

    def mailbox_handle(self) -> bytes:
        """This rank's CUDA IPC handle of the mailbox used by the resident multi-GPU CG (call after init_slab)."""
        buf = C.create_string_buffer(64)
        check(self.L.sq_fdm_mailbox_create(self.h, buf))
        return buf.raw

    def mailbox_open(self, handles):
        """handles: the mailbox handles of all ranks in rank order (e.g. from torch.distributed.all_gather_object)."""
        blob = b"".join(handles)
        check(self.L.sq_fdm_mailbox_open(self.h, C.create_string_buffer(blob, len(blob))))

    def set_slab_range(self, lo, hi): check(self.L.sq_fdm_set_slab_range(self.h, int(lo), int(hi)))

    def set_sharded_solve(self, enable=True):
        """One chain over several GPUs: full state on every rank, only the CG solves are tau-slab partitioned (after init_slab)."""
        check(self.L.sq_fdm_set_sharded_solve(self.h, 1 if enable else 0))

    def init_sharded_solve(self, dist):
        """Convenience for torch.distributed programs: communicator, peer-mapped mailboxes and the sharded-solve switch."""
        rank, world = dist.get_rank(), dist.get_world_size()
        ids = [FermionDetMatrix.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.init_slab(rank, world, ids[0])
        handles = [None] * world
        dist.all_gather_object(handles, self.mailbox_handle())
        self.mailbox_open(handles)
        self.set_sharded_solve(True)

    @property
    def slab(self):
        lo, hi, r, w = C.c_int64(0), C.c_int64(0), C.c_int(0), C.c_int(0)
        check(self.L.sq_fdm_get_slab(self.h, C.byref(lo), C.byref(hi), C.byref(r), C.byref(w)))
        return {"lo": lo.value, "hi": hi.value, "rank": r.value, "world": w.value}

    @property
    def stream(self):
        s = C.c_void_p()
        check(self.L.sq_fdm_stream(self.h, C.byref(s)))
        return s.value

    @property
    def launch_count(self): return int(self.L.sq_fdm_launch_count(self.h))

    STAT_NAMES = ("cg_solves", "cg_resident", "cg_persistent_smem", "cg_launch_loop", "cg_preconditioned", "cg_slab_nccl", "cg_slab_resident",
                  "watchdog_aborts", "instabilities", "cg_iterations", "kpm_register", "kpm_smem", "cg_batched_rhs", "cg_slab_preconditioned")

    @property
    def stats(self):
        """Diagnostic counters (sq_fdm_stats): which solver ran how often, watchdog aborts of the resident kernels, instabilities."""
        out = np.zeros(16, np.int64)
        check(self.L.sq_fdm_stats(self.h, ptr(out), 16))
        return dict(zip(self.STAT_NAMES, (int(v) for v in out)))


def SymFermionDetMatrix(model, **kw): return FermionDetMatrix(model, sym=True, **kw)
def AsymFermionDetMatrix(model, **kw): return FermionDetMatrix(model, sym=False, **kw)


class KPMPreconditioner:
    """KPMPreconditioner(fdm; rng, rbuf, n, a1, a2) (src/KPMPreconditioner.jl:198-284)."""

    def __init__(self, fdm, rbuf=0.10, n=20, a1=1.0, a2=1.0, lanczos_start=None, update=True, seed=None):
        self.L, self.fdm = fdm.L, fdm
        h = C.c_void_p()
        check(self.L.sq_kpm_create(C.byref(h), fdm.h, rbuf, n, a1, a2))
        self.h = h
        if seed is not None:            # `rng` of KPMPreconditioner(fdm; rng, ...): keys the library-drawn Lanczos start vectors
            check(self.L.sq_kpm_set_seed(self.h, int(seed) & (2 ** 64 - 1)))
        if update:
            self.update(lanczos_start)

    def close(self):
        if getattr(self, "h", None):
            self.L.sq_kpm_destroy(self.h)
            self.h = None

    __del__ = close

    def update(self, lanczos_start=None):
        """update_preconditioner!(P, fdm, rng) (:554-597) -> (active, bounds)."""
        act, b = C.c_int(0), np.zeros(2)
        ls = None if lanczos_start is None else _f64(lanczos_start)
        check(self.L.sq_kpm_update(self.h, ptr(ls), C.byref(act), ptr(b)))
        return bool(act.value), b

    def set_bounds(self, emin, emax): check(self.L.sq_kpm_set_bounds(self.h, emin, emax))

    @property
    def orders(self):
        n = C.c_int64(0)
        check(self.L.sq_kpm_get_orders(self.h, C.byref(n), None))
        o = np.zeros(n.value, np.int64)
        check(self.L.sq_kpm_get_orders(self.h, C.byref(n), ptr(o)))
        return o

    def coefs(self, l):
        c = np.zeros(int(self.orders[l]), np.complex128)
        check(self.L.sq_kpm_get_coefs(self.h, l, ptr(c)))
        return c

    def ldiv(self, v):
        v = _cvec(self.fdm.model, v)
        out = _cvec(self.fdm.model)
        check(self.L.sq_kpm_ldiv(self.h, ptr(out), ptr(v)))
        return out

    def ldiv_dev(self, d_out, d_in): check(self.L.sq_kpm_ldiv_dev(self.h, ptr(d_out), ptr(d_in)))

    def fourier(self, v, forward=True):
        v = _cvec(self.fdm.model, v)
        check(self.L.sq_kpm_fourier(self.h, ptr(v), int(forward)))
        return v


class ElectronPhononParameters:
    """Device twin of the SmoQyDQMC ElectronPhononParameters / FermionPathIntegral fields the path reads."""

    def __init__(self, model, fdm):
        self.L, self.model, self.fdm = fdm.L, model, fdm
        m = model
        k = [_f64(m.Omega), _f64(m.Omega4), _f64(m.Mass), _i64(m.hol_phonon), _i64(m.hol_site),
             _f64(m.hol_alpha[0]), _f64(m.hol_alpha[1]), _f64(m.hol_alpha[2]), _f64(m.hol_alpha[3]),
             np.ascontiguousarray(m.hol_phsym, np.int32), _i64(m.ssh_phonon.T), _i64(m.ssh_hopping),
             _f64(m.ssh_alpha[0]), _f64(m.ssh_alpha[1]), _f64(m.ssh_alpha[2]), _f64(m.ssh_alpha[3]), _f64(m.V0), _f64(m.t0)]
        h = C.c_void_p()
        check(self.L.sq_elph_create(C.byref(h), fdm.h, m.dtau, m.Nph, ptr(k[0]), ptr(k[1]), ptr(k[2]),
                                    m.Nhol, ptr(k[3]), ptr(k[4]), ptr(k[5]), ptr(k[6]), ptr(k[7]), ptr(k[8]), ptr(k[9]),
                                    m.Nssh, ptr(k[10]), ptr(k[11]), ptr(k[12]), ptr(k[13]), ptr(k[14]), ptr(k[15]),
                                    ptr(k[16]), ptr(k[17])))
        self.h = h
        if getattr(m, "Ndisp", 0):
            dp, do, do4 = _i64(m.disp_phonon.T), _f64(m.disp_Omega), _f64(m.disp_Omega4)
            check(self.L.sq_elph_set_dispersion(self.h, m.Ndisp, ptr(dp), ptr(do), ptr(do4)))

    def close(self):
        if getattr(self, "h", None):
            self.L.sq_elph_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def x(self):
        out = np.zeros((self.model.Nph, self.model.Ltau), order="F")
        check(self.L.sq_elph_get_x(self.h, ptr(out)))
        return out

    @x.setter
    def x(self, val):
        val = np.asfortranarray(val, np.float64)
        assert val.shape == (self.model.Nph, self.model.Ltau)
        check(self.L.sq_elph_set_x(self.h, ptr(val)))

    def update_fdm(self):
        """update!(fpi, elph, x, +1); update!(fdm, fpi) (src/EFAPFFHMCUpdater.jl:152-153)."""
        check(self.L.sq_elph_refresh_fdm(self.h))

    def shift_mu(self, dmu): check(self.L.sq_elph_shift_mu(self.h, dmu))

    def Vt(self):
        m = self.model
        V = np.zeros((m.N, m.Ltau), order="F")
        t = np.zeros((m.Nh, m.Ltau), order="F")
        check(self.L.sq_elph_get_Vt(self.h, ptr(V), ptr(t)))
        return V, t

    def potential_derivative(self):
        """Anharmonic + dispersive parts of dS_b/dx (what the leapfrog kick adds to the fermionic force)."""
        F = np.zeros((self.model.Nph, self.model.Ltau), order="F")
        check(self.L.sq_elph_potential_derivative(self.h, ptr(F)))
        return F

    def bosonic_action(self):
        s = C.c_double(0)
        check(self.L.sq_elph_bosonic_action(self.h, C.byref(s)))
        return s.value

    # x-mutations of the global moves, on the device (0-based phonon indices here, 1-based in the C ABI)
    def scale_x(self, p_first, p_last, factor): check(self.L.sq_elph_scale_x(self.h, int(p_first) + 1, int(p_last) + 1, float(factor)))
    def swap_x(self, p_i, p_j): check(self.L.sq_elph_swap_x(self.h, int(p_i) + 1, int(p_j) + 1))
    def backup_x(self): check(self.L.sq_elph_backup_x(self.h))
    def restore_x(self): check(self.L.sq_elph_restore_x(self.h))


class PFFCalculator:
    """PFFCalculator(elph, fdm) (src/PFFCalculator.jl:30-53)."""

    def __init__(self, elph, fdm=None, exact_holstein=False, seed=None):
        self.L, self.elph, self.fdm = elph.L, elph, elph.fdm
        h = C.c_void_p()
        check(self.L.sq_pff_create(C.byref(h), elph.h))
        self.h = h
        if seed is not None:            # keys the library-drawn pseudofermion noise (global moves); default: derived from the HMC seed
            check(self.L.sq_pff_set_seed(self.h, int(seed) & (2 ** 64 - 1)))
        check(self.L.sq_pff_set_exact_holstein(self.h, int(exact_holstein)))

    def close(self):
        if getattr(self, "h", None):
            self.L.sq_pff_destroy(self.h)
            self.h = None

    __del__ = close

    def sample_pseudofermion_fields(self, R=None):
        Sf = C.c_double(0)
        Rv = None if R is None else _cvec(self.fdm.model, R)
        check(self.L.sq_pff_sample(self.h, ptr(Rv), C.byref(Sf)))
        return Sf.value

    def calculate_fermionic_action(self, preconditioner=None, lanczos_start=None, tol=1e-10, maxiter=10000):
        Sf, it, eps = C.c_double(0), C.c_int64(0), C.c_double(0)
        ls = None if lanczos_start is None else _f64(lanczos_start)
        check(self.L.sq_pff_action(self.h, preconditioner.h if preconditioner is not None else None, ptr(ls), tol, int(maxiter),
                                   C.byref(Sf), C.byref(it), C.byref(eps)))
        return Sf.value, it.value, eps.value

    def calculate_derivative_fermionic_action(self, dSdx=None, preconditioner=None, lanczos_start=None, tol=1e-5, maxiter=10000):
        m = self.fdm.model
        F = np.zeros((m.Nph, m.Ltau), order="F") if dSdx is None else dSdx
        assert F.flags.f_contiguous and F.dtype == np.float64
        Sf, it, eps = C.c_double(0), C.c_int64(0), C.c_double(0)
        ls = None if lanczos_start is None else _f64(lanczos_start)
        check(self.L.sq_pff_force(self.h, ptr(F), preconditioner.h if preconditioner is not None else None, ptr(ls), tol, int(maxiter),
                                  C.byref(Sf), C.byref(it), C.byref(eps)))
        return F, Sf.value, it.value, eps.value

    def fields(self):
        m = self.fdm.model
        Phi, Psi = _cvec(m), _cvec(m)
        Lam = np.zeros((m.Ltau, m.N), order="F")
        check(self.L.sq_pff_get_fields(self.h, ptr(Phi), ptr(Psi), ptr(Lam)))
        return Phi, Psi, Lam

    def set_Phi(self, Phi):
        Phi = _cvec(self.fdm.model, Phi)
        check(self.L.sq_pff_set_Phi(self.h, ptr(Phi)))

    def lambda_op(self, which, v):
        code = {"mul": 0, "ldiv": 1, "mulT": 2, "ldivT": 3}[which]
        v = _cvec(self.fdm.model, v)
        out = _cvec(self.fdm.model)
        check(self.L.sq_pff_lambda_op(self.h, code, ptr(out), ptr(v)))
        return out

    def dM_dx(self, nu, u, v):
        m = self.fdm.model
        F = np.zeros((m.Nph, m.Ltau), order="F")
        check(self.L.sq_pff_dM_dx(self.h, ptr(F), nu, ptr(_cvec(m, u)), ptr(_cvec(m, v))))
        return F

    def dLambda_dx(self, nu, up, u):
        m = self.fdm.model
        F = np.zeros((m.Nph, m.Ltau), order="F")
        check(self.L.sq_pff_dLambda_dx(self.h, ptr(F), nu, ptr(_cvec(m, up)), ptr(_cvec(m, u))))
        return F


class EFAPFFHMCUpdater:
    """EFAPFFHMCUpdater(; electron_phonon_parameters, Nt, dt, eta, delta) (src/EFAPFFHMCUpdater.jl:40-72)."""

    def __init__(self, elph, pff, Nt, dt=None, eta=0.0, delta=0.05, seed=0):
        self.L, self.elph, self.pff = elph.L, elph, pff
        self.Nt = int(Nt)
        self.dt = float(np.pi / (2 * Nt) if dt is None else dt)
        h = C.c_void_p()
        check(self.L.sq_hmc_create(C.byref(h), pff.h, self.Nt, self.dt, eta, delta, seed))
        self.h = h
        self.info = None

    def close(self):
        if getattr(self, "h", None):
            self.L.sq_hmc_destroy(self.h)
            self.h = None

    __del__ = close

    def hmc_update(self, preconditioner=None, tol_action=1e-10, tol_force=1e-5, maxiter=10000, randoms=None):
        """hmc_update! (:102-279) -> (accepted, iters_avg).  A library failure == numerical instability => rejected."""
        acc = C.c_int(0)
        info = np.zeros(8)
        rnd = None if randoms is None else _f64(randoms)
        check(self.L.sq_hmc_update(self.h, preconditioner.h if preconditioner is not None else None, tol_action, tol_force, int(maxiter),
                                   ptr(rnd), 0 if rnd is None else rnd.size, C.byref(acc), ptr(info)))
        self.info = info
        return bool(acc.value), info[0]

    @property
    def last_reject(self):
        """Reason of the last forced rejection (numerical instability inside the trajectory), '' if the last update was stable."""
        return self.L.sq_hmc_last_reject(self.h).decode()

    def set_seed(self, seed): check(self.L.sq_hmc_set_seed(self.h, int(seed) & (2 ** 64 - 1)))

    def init_momentum(self, R):
        m = self.elph.model
        p = np.zeros((m.Nph, m.Ltau), order="F")
        K = C.c_double(0)
        check(self.L.sq_hmc_init_momentum(self.h, ptr(np.asfortranarray(R, np.float64)), ptr(p), C.byref(K)))
        return p, K.value

    def kinetic(self, p):
        K = C.c_double(0)
        check(self.L.sq_hmc_kinetic(self.h, ptr(np.asfortranarray(p, np.float64)), C.byref(K)))
        return K.value

    def evolve(self, x, p, dt):
        x = np.array(x, np.float64, order="F", copy=True)
        p = np.array(p, np.float64, order="F", copy=True)
        check(self.L.sq_hmc_evolve(self.h, ptr(x), ptr(p), dt))
        return x, p


class GreensEstimator:
    """GreensEstimator(fdm, model_geometry; Nrv, ...) solves + scalar measurements
    (src/Measurements/GreensEstimator.jl:63-175, scalar_measurements.jl)."""

    def __init__(self, fdm, Nrv=10, seed=0):
        self.L, self.fdm, self.Nrv = fdm.L, fdm, int(Nrv)
        h = C.c_void_p()
        check(self.L.sq_greens_create(C.byref(h), fdm.h, self.Nrv, seed))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.sq_greens_destroy(self.h)
            self.h = None

    __del__ = close

    def update_greens_estimator(self, preconditioner=None, R=None, tol=1e-10, maxiter=10000):
        avg = C.c_double(0)
        Rv = None if R is None else np.asfortranarray(R, np.complex128)
        check(self.L.sq_greens_update(self.h, preconditioner.h if preconditioner is not None else None, ptr(Rv), tol, int(maxiter), C.byref(avg)))
        return avg.value

    def get(self):
        V = self.fdm.model.N * self.fdm.model.Ltau
        R = np.zeros((V, self.Nrv), np.complex128, order="F")
        GR = np.zeros((V, self.Nrv), np.complex128, order="F")
        check(self.L.sq_greens_get(self.h, ptr(R), ptr(GR)))
        return R, GR

    def set_GR(self, GR): check(self.L.sq_greens_set_GR(self.h, ptr(np.asfortranarray(GR, np.complex128))))

    def measure_GD0(self, orbitals=(0, 0), norb=None, dims=None):
        """measure_GΔ0! (src/Measurements/GreensEstimator.jl:177-233): G_ab(Δr, Δτ) averaged over translations, returned with
        the reference's `correlation` axes (L..., Ltau + 1).  orbitals are 0-based here."""
        m = self.fdm.model
        dims = tuple(m.lattice_dims) if dims is None else tuple(dims)
        ncell = int(np.prod(dims))
        norb = m.N // ncell if norb is None else int(norb)
        d = _i64(dims, one_based=False)
        out = np.zeros((m.Ltau + 1,) + dims, np.complex128, order="F")
        check(self.L.sq_greens_measure_GD0(self.h, norb, len(dims), ptr(d), int(orbitals[0]) + 1, int(orbitals[1]) + 1, ptr(out)))
        return np.moveaxis(out, 0, -1)

    def _geom(self, norb, dims):
        m = self.fdm.model
        dims = tuple(m.lattice_dims) if dims is None else tuple(dims)
        return (m.N // int(np.prod(dims)) if norb is None else int(norb)), dims

    def measure_contraction(self, kind, orbitals, r=None, norb=None, dims=None, tD=None, t0=None):
        """kind: "GD0_GD0" (measure_GΔ0_GΔ0!), "GDD_G00" (measure_GΔΔ_G00!) or "G0D_GD0" (measure_G0Δ_GΔ0!)
        (src/Measurements/GreensEstimator.jl:236-606); orbitals (a, b, c, d) 0-based, r = (r1, r2, r3, r4) displacement tuples
        (default all zero); tD, t0: optional real hopping weights of shape (Ltau, L...) (the reference's tΔ, t0).
        Returns the contraction with the reference's `correlation` axes (L..., Ltau + 1), coef = 1."""
        m = self.fdm.model
        norb, dims = self._geom(norb, dims)
        code = {"GD0_GD0": 0, "GDD_G00": 1, "G0D_GD0": 2}[kind]
        rr = np.zeros((4, len(dims)), np.int64) if r is None else np.ascontiguousarray(r, np.int64).reshape(4, len(dims))
        orb = np.ascontiguousarray(np.asarray(orbitals, np.int32) + 1)
        out = np.zeros((m.Ltau + 1,) + dims, np.complex128, order="F")
        if tD is None and t0 is None:
            check(self.L.sq_greens_measure_contraction(self.h, code, norb, len(dims), ptr(_i64(dims, one_based=False)), ptr(orb), ptr(rr), ptr(out)))
        else:
            w = []
            for t in (tD, t0):
                if t is None:
                    w.append(None)
                    continue
                t = np.asarray(t)
                if np.iscomplexobj(t):
                    raise ValueError("real hoppings only")
                t = np.asfortranarray(t, np.float64)
                assert t.shape == (m.Ltau,) + dims, "hopping weights have the shape (Ltau, L...)"
                w.append(t)
            check(self.L.sq_greens_measure_contraction_weighted(self.h, code, norb, len(dims), ptr(_i64(dims, one_based=False)), ptr(orb), ptr(rr),
                                                                ptr(w[0]) if w[0] is not None else None, ptr(w[1]) if w[1] is not None else None, ptr(out)))
        return np.moveaxis(out, 0, -1)

    def measure_current_correlation(self, bond1, bond2, t1, t2, coef=1.0, spins=None, norb=None, dims=None):
        """measure_current_correlation!(CC, greens_estimator, b', b'', t', t'', [σ', σ''], coef)
        (src/Measurements/Correlations/current.jl:2-151): four hopping-weighted G(Δ,Δ)G(0,0) and four G(0,Δ)G(Δ,0) contractions.
        A bond is ((orbital_1, orbital_2), displacement); t1, t2 = the effective hoppings of the two bonds, shape (Ltau, L...)
        (make_measurements.jl:316-320).  spins = None: the spin-summed form (factors 4 and 2); (σ', σ''): the spin-resolved one."""
        (b, a), r1 = bond1
        (d, c), r2 = bond2
        r1, r2 = tuple(r1), tuple(r2)
        z = (0,) * len(r1)
        mc = lambda kind, orbs, r: self.measure_contraction(kind, orbs, r, norb=norb, dims=dims, tD=t1, t0=t2)
        f1, f2 = (4.0, 2.0) if spins is None else (1.0, 1.0 if spins[0] == spins[1] else 0.0)
        CC = f1 * coef * (mc("GDD_G00", (a, b, d, c), (r1, z, z, r2)) - mc("GDD_G00", (a, b, c, d), (r1, z, r2, z))
                          - mc("GDD_G00", (b, a, d, c), (z, r1, z, r2)) + mc("GDD_G00", (b, a, c, d), (z, r1, r2, z)))
        if f2:
            CC += f2 * coef * (-mc("G0D_GD0", (b, a, c, d), (z, z, r1, r2)) + mc("G0D_GD0", (b, a, d, c), (r2, z, r1, z))
                               + mc("G0D_GD0", (d, a, b, c), (z, r1, z, r2)) - mc("G0D_GD0", (c, a, b, d), (r2, r1, z, z)))
        return CC

    def measure_double_occ_orbital(self, a, norb=None, dims=None):
        """measure_double_occ(greens_estimator, orbital) (scalar_measurements.jl:98-109; normalised by the total N Ltau as there)."""
        norb, dims = self._geom(norb, dims)
        out = np.zeros(2)
        check(self.L.sq_greens_measure_double_occ_orbital(self.h, norb, int(a) + 1, ptr(out)))
        return complex(out[0], out[1])

    def measure_n_orbital(self, a, norb=None, dims=None):
        norb, dims = self._geom(norb, dims)
        out = np.zeros(2)
        check(self.L.sq_greens_measure_n_orbital(self.h, norb, int(a) + 1, ptr(out)))
        return complex(out[0], out[1])

    def measure_density_correlation(self, a, b, coef=1.0, norb=None, dims=None, spins=None):
        """measure_density_correlation!(DD, greens_estimator, a, b, [σ, σ'], coef)  (src/Measurements/Correlations/density.jl:2-65):
        spin-summed  DD = 4 coef (n_a + n_b - 1) + 4 coef G(Δ,Δ)G(0,0)[a,a,b,b] - 2 coef G(0,Δ)G(Δ,0)[b,a,a,b];
        spin-resolved (spins = (σ, σ')): factors 1, and the exchange term only for equal spins."""
        na, nb = self.measure_n_orbital(a, norb, dims), self.measure_n_orbital(b, norb, dims)
        f1, f2 = (4.0, 2.0) if spins is None else (1.0, 1.0 if spins[0] == spins[1] else 0.0)
        DD = f1 * coef * self.measure_contraction("GDD_G00", (a, a, b, b), norb=norb, dims=dims)
        if f2:
            DD -= f2 * coef * self.measure_contraction("G0D_GD0", (b, a, a, b), norb=norb, dims=dims)
        return DD + f1 * coef * (na + nb - 1)

    def measure_pair_correlation(self, bond1, bond2, coef=1.0, norb=None, dims=None):
        """measure_pair_correlation!(PP, greens_estimator, b', b'', coef)  (src/Measurements/Correlations/pair.jl:2-21).
        A bond is ((orbital_1, orbital_2), displacement); b, a = b'.orbitals, d, c = b''.orbitals."""
        (b, a), r1 = bond1
        (d, c), r2 = bond2
        z = (0,) * len(tuple(r1))
        return coef * self.measure_contraction("GD0_GD0", (a, c, b, d), (tuple(r1), tuple(r2), z, z), norb=norb, dims=dims)

    def measure_bond_correlation(self, bond1, bond2, coef=1.0, norb=None, dims=None, spins=None):
        """measure_bond_correlation!(BB, greens_estimator, b', b'', [σ', σ''], coef)  (src/Measurements/Correlations/bond.jl:2-131): four
        G(Δ,Δ)G(0,0) and four G(0,Δ)G(Δ,0) contractions (spin-summed: factors 4 and 2; spin-resolved: 1, exchange terms only for equal
        spins).  A bond is ((orbital_1, orbital_2), displacement)."""
        (b, a), r1 = bond1
        (d, c), r2 = bond2
        r1, r2 = tuple(r1), tuple(r2)
        z = (0,) * len(r1)
        mc = lambda kind, orbs, r: self.measure_contraction(kind, orbs, r, norb=norb, dims=dims)
        f1, f2 = (4.0, 2.0) if spins is None else (1.0, 1.0 if spins[0] == spins[1] else 0.0)
        BB = f1 * coef * (mc("GDD_G00", (a, b, c, d), (r1, z, r2, z)) + mc("GDD_G00", (a, b, d, c), (r1, z, z, r2))
                          + mc("GDD_G00", (b, a, c, d), (z, r1, r2, z)) + mc("GDD_G00", (b, a, d, c), (z, r1, z, r2)))
        if f2:
            BB -= f2 * coef * (mc("G0D_GD0", (c, b, a, d), (r2, z, r1, z)) + mc("G0D_GD0", (d, b, a, c), (z, z, r1, r2))
                               + mc("G0D_GD0", (c, a, b, d), (r2, r1, z, z)) + mc("G0D_GD0", (d, a, b, c), (z, r1, z, r2)))
        return BB

    def measure_spin_correlation(self, a, b, coef=1.0, norb=None, dims=None):
        """measure_spin_correlation!(SzSz, greens_estimator, a, b, coef)  (src/Measurements/Correlations/spin.jl:2-15)."""
        return -0.5 * coef * self.measure_contraction("G0D_GD0", (b, a, a, b), norb=norb, dims=dims)

    # ---- local measurements (tight_binding_measurements.jl, electron_phonon_measurements.jl) ----
    def weighted_density(self, w):
        """sum_{i,l} w[i,l] n(l,i) with n(l,i) = mean_rv (1 - GR Rt); w: (N, Ltau) real."""
        m = self.fdm.model
        w = np.asfortranarray(w, np.float64)
        assert w.shape == (m.N, m.Ltau)
        out = np.zeros(2)
        check(self.L.sq_greens_weighted_density(self.h, ptr(w), ptr(out)))
        return complex(out[0], out[1])

    def weighted_bonds(self, bonds, w):
        """sum_{m,l} w[m,l] <GR[l,i_m] Rt[l,f_m]> + conj(w[m,l]) <GR[l,f_m] Rt[l,i_m]>; bonds (2, nb) 0-based, w (nb, Ltau) complex."""
        m = self.fdm.model
        bonds = np.asarray(bonds)
        w = np.asfortranarray(w, np.complex128)
        assert bonds.shape[0] == 2 and w.shape == (bonds.shape[1], m.Ltau)
        out = np.zeros(2)
        check(self.L.sq_greens_weighted_bonds(self.h, bonds.shape[1], ptr(_i64(bonds.T)), ptr(w), ptr(out)))
        return complex(out[0], out[1])

    def measure_onsite_energy(self, orbital, eps, mu, norb=None, dims=None):
        """measure_onsite_energy (tight_binding_measurements.jl:43-63): eps (N,) on-site energies."""
        m = self.fdm.model
        norb, dims = self._geom(norb, dims)
        w = np.zeros((m.N, m.Ltau))
        sel = np.arange(m.N) % norb == orbital
        w[sel, :] = ((np.asarray(eps, float)[sel] - mu) / (m.Ltau * (m.N // norb)))[:, None]
        return self.weighted_density(w)

    def measure_hopping_energy(self, bonds, t):
        """measure_bare_hopping_energy / measure_hopping_energy (tight_binding_measurements.jl:66-133): bonds (2, nb) of one hopping
        id, t (nb,) bare or (nb, Ltau) modulated amplitudes."""
        m = self.fdm.model
        t = np.asarray(t, np.complex128)
        w = np.broadcast_to(t[:, None] if t.ndim == 1 else t, (np.asarray(bonds).shape[1], m.Ltau)) / (m.Ltau * m.N)
        return self.weighted_bonds(bonds, w)

    def measure_holstein_energy(self, x, holstein_id=0):
        """measure_holstein_energy (electron_phonon_measurements.jl): couplings of one Holstein id (one per unit cell)."""
        m = self.fdm.model
        nc = m.n_unit_cells
        sl = slice(holstein_id * nc, (holstein_id + 1) * nc)
        site, ph = m.hol_site[sl], m.hol_phonon[sl]
        a1, a2, a3, a4 = (m.hol_alpha[k][sl][:, None] for k in range(4))
        xs = np.asarray(x)[ph, :]
        even, odd = a2 * xs ** 2 + a4 * xs ** 4, a1 * xs + a3 * xs ** 2          # the reference uses x^2 with alpha3 (:127-128)
        w = np.zeros((m.N, m.Ltau))
        w[site, :] = (even + odd) / (nc * m.Ltau)
        e = self.weighted_density(w)
        if m.hol_phsym[sl][0]:
            e -= 0.5 * np.sum(odd) / (nc * m.Ltau)
        return e

    def measure_ssh_energy(self, x, ssh_id=0):
        """measure_ssh_energy (electron_phonon_measurements.jl:124-186): the couplings of one SSH id (one per unit cell),
        eps = sum c(dx) h_forward + conj(c) h_reverse with h = -<G R Rt> and c = a1 dx + a2 dx^2 + a3 dx^3 + a4 dx^4, / (N_cells Ltau)."""
        m = self.fdm.model
        nc = m.n_unit_cells
        sl = slice(ssh_id * nc, (ssh_id + 1) * nc)
        hop = m.ssh_hopping[sl]
        bonds = m.neighbor_table[:, hop]
        p_i, p_f = m.ssh_phonon[0, sl], m.ssh_phonon[1, sl]
        a1, a2, a3, a4 = (m.ssh_alpha[k][sl][:, None] for k in range(4))
        xs = np.asarray(x)
        dx = xs[p_f, :] - xs[p_i, :]
        c = a1 * dx + a2 * dx ** 2 + a3 * dx ** 3 + a4 * dx ** 4
        return self.weighted_bonds(bonds, -c.astype(np.complex128) / (nc * m.Ltau))

    def measure(self):
        out = np.zeros((3, 2))
        check(self.L.sq_greens_measure(self.h, ptr(out[0]), ptr(out[1]), ptr(out[2])))
        return {"n": complex(*out[0]), "double_occ": complex(*out[1]), "Nsqrd": complex(*out[2])}


def make_measurements(measurements, fdm, greens, *, mu=0.0, bosonic_action=None, preconditioner=None, tol=1e-10, maxiter=10000,
                      correlations=(), norb=None, dims=None):
    """make_measurements!(measurement_container, fdm, greens_estimator; ...) -> iters  (src/Measurements/make_measurements.jl:19-90) with a
    plain dictionary in place of SmoQyDQMC's container: refreshes the estimator (update_greens_estimator!), then ADDS this
    configuration's values to `measurements` (created on first use):
      "global"       make_global_measurements! (:93-117): sgn, density_up / density_dn / density, double_occ, Nsqrd, chemical_potential,
                     action_bosonic (if `bosonic_action` is given: a number or a callable)
      "local"        the density part of make_local_measurements! (:120-146): density_up / density_dn / density / double_occ per orbital
      "correlations" one array per entry of `correlations`: ("greens", (a, b)), ("density", (a, b)), ("spin_z", (a, b)),
                     ("pair", (bond', bond'')), ("bond", (bond', bond'')), ("current", (bond', bond'', t', t'')) -- the calls of
                     make_correlation_measurements! (:161-394) -- and ("composite", (label, type, id_pairs, coefficients)), the linear
                     combinations of make_composite_correlation_measurements! (:398-913); the caller divides by the number of calls, as
                     the container's processing does."""
    iters = greens.update_greens_estimator(preconditioner=preconditioner, tol=tol, maxiter=maxiter)
    norb, dims = greens._geom(norb, dims)
    G = measurements.setdefault("global", {})
    s = greens.measure()
    add = lambda d, k, v: d.__setitem__(k, d.get(k, 0.0) + v)
    add(G, "sgn", 1.0)
    add(G, "density_up", s["n"]); add(G, "density_dn", s["n"]); add(G, "density", 2 * s["n"])
    add(G, "double_occ", s["double_occ"]); add(G, "Nsqrd", s["Nsqrd"]); add(G, "chemical_potential", mu)
    if bosonic_action is not None:
        add(G, "action_bosonic", bosonic_action() if callable(bosonic_action) else bosonic_action)
    Lm = measurements.setdefault("local", {k: np.zeros(norb, complex) for k in ("density_up", "density_dn", "density", "double_occ")})
    for a in range(norb):
        n = greens.measure_n_orbital(a, norb, dims)
        Lm["density_up"][a] += n; Lm["density_dn"][a] += n; Lm["density"][a] += 2 * n
        Lm["double_occ"][a] += greens.measure_double_occ_orbital(a, norb, dims)
    Cm = measurements.setdefault("correlations", {})

    def one(name, args):
        if name == "greens":
            return greens.measure_GD0(tuple(args), norb=norb, dims=dims)
        if name == "density":
            return greens.measure_density_correlation(args[0], args[1], norb=norb, dims=dims)
        if name == "spin_z":
            return greens.measure_spin_correlation(args[0], args[1], norb=norb, dims=dims)
        if name == "pair":
            return greens.measure_pair_correlation(args[0], args[1], norb=norb, dims=dims)
        if name == "bond":
            return greens.measure_bond_correlation(args[0], args[1], norb=norb, dims=dims)
        if name == "current":
            return greens.measure_current_correlation(args[0], args[1], args[2], args[3], norb=norb, dims=dims)
        raise ValueError("unknown correlation " + str(name))

    for name, args in correlations:
        if name == "composite":
            # make_composite_correlation_measurements! (make_measurements.jl:398-913): a named linear combination
            # sum_k coefficient_k x correlation(id_pair_k) of one correlation type -- args = (label, type, id_pairs, coefficients)
            label, ctype, pairs, coefs = args
            if len(pairs) != len(coefs):
                raise ValueError("composite correlation: one coefficient per id pair")
            val = sum(c * one(ctype, pr) for pr, c in zip(pairs, coefs))
            Cm[("composite", label)] = Cm.get(("composite", label), 0.0) + val
            continue
        key = (name,) + tuple(repr(a) if isinstance(a, np.ndarray) else a for a in args[:2])
        Cm[key] = Cm.get(key, 0.0) + one(name, args)
    return iters


def update_chemical_potential(fdm, greens, elph, mu, mu_new_fn, preconditioner=None, update_greens_estimator=True, tol=1e-10, maxiter=10000):
    """update_chemical_potential! (src/update_chemical_potential.jl:21-73).  The MuTuner scalar logic stays on
    the host: `mu_new_fn(n, Nsqrd) -> mu'` stands in for MuTuner.update!.  Returns (mu', iters)."""
    iters = 0
    if update_greens_estimator:
        iters = greens.update_greens_estimator(preconditioner=preconditioner, tol=tol, maxiter=maxiter)
    meas = greens.measure()
    n = (2 * meas["n"]).real
    Nsqrd = meas["Nsqrd"].real
    mu_new = mu_new_fn(n, Nsqrd)
    elph.shift_mu(mu_new - mu)
    elph.update_fdm()
    return mu_new, iters


# ------------------------------------------------------------------------------------------------------
# global moves: reflection_update!, swap_update!, radial_update!
# ------------------------------------------------------------------------------------------------------
def _sample_phonon_mode(rng, model, phonon_types=None):
    """SmoQyDQMC._sample_phonon_mode [unvendored]: a phonon type among `phonon_types` (all if None), then a unit cell,
    uniformly; modes with infinite mass are frozen and never proposed."""
    types = list(range(model.nphonon)) if phonon_types is None else list(phonon_types)
    ncell = model.n_unit_cells
    for _ in range(10000):
        mode = int(types[rng.integers(len(types))]) * ncell + int(rng.integers(ncell))
        if np.isfinite(model.Mass[mode]):
            return mode
    raise SqError("no phonon mode with finite mass among the requested phonon types")


def _sample_phonon_mode_pair(rng, model, phonon_type_pairs=None):
    """SmoQyDQMC._sample_phonon_mode_pair [unvendored]: a pair of phonon types, then two different unit cells."""
    nph, ncell = model.nphonon, model.n_unit_cells
    pairs = [(a, b) for a in range(nph) for b in range(nph)] if phonon_type_pairs is None else list(phonon_type_pairs)
    for _ in range(10000):
        a, b = pairs[int(rng.integers(len(pairs)))]
        i, j = int(a) * ncell + int(rng.integers(ncell)), int(b) * ncell + int(rng.integers(ncell))
        if i != j and np.isfinite(model.Mass[i]) and np.isfinite(model.Mass[j]):
            return i, j
    raise SqError("no pair of phonon modes with finite mass among the requested phonon type pairs")


def _global_move(elph, pff, mutate, log_jacobian, preconditioner, tol, maxiter, R, u_accept, rng):
    """Common body of the three moves (src/reflection_update.jl:67-176, swap_update.jl:68-176, radial_update.jl:88-193):
    S = Sf + Sb on fresh pseudofermion fields, mutate x, refresh the operator, S' with the same fields, Metropolis test;
    a failed solve (numerical instability) rejects the move.  Returns (accepted, iters, info)."""
    Sf = pff.sample_pseudofermion_fields(R)
    Sb = elph.bosonic_action()
    elph.backup_x()
    mutate()
    elph.update_fdm()
    dS, iters, stable = np.inf, 0, True
    reason = ""
    try:
        Sf2, iters, _ = pff.calculate_fermionic_action(preconditioner=preconditioner, tol=tol, maxiter=maxiter)
        dS = (Sf2 + elph.bosonic_action()) - (Sf + Sb)
        stable = np.isfinite(dS)
    except SqNumericalInstability as err:       # only instabilities reject (reflection_update.jl:111-127); CUDA / argument errors propagate
        stable, reason = False, str(err)
        import warnings
        warnings.warn("Failed to evaluate the fermionic action for the proposed state, update rejected: " + reason)
    P = min(1.0, float(np.exp(-dS + log_jacobian))) if stable else 0.0
    u = float(rng.random()) if u_accept is None else float(u_accept)
    accepted = u < P
    if not accepted:
        elph.restore_x()
        elph.update_fdm()
    return accepted, int(iters), {"dS": float(dS), "P": P, "stable": stable, "reason": reason}


def reflection_update(elph, pff, rng=None, preconditioner=None, tol=None, maxiter=None, phonon_types=None, randoms=None):
    """reflection_update!(elph, pff; fermion_path_integral, fermion_det_matrix, rng, preconditioner, tol, maxiter,
    phonon_types) -> (accepted, iters)  (src/reflection_update.jl:23-177): x_p -> -x_p for one random phonon mode p.
    `randoms` = {"mode": p, "R": ..., "u": ...} replaces the random draws (parity tests)."""
    rng = np.random.default_rng() if rng is None else rng
    rd = randoms or {}
    mode = rd["mode"] if "mode" in rd else _sample_phonon_mode(rng, elph.model, phonon_types)
    acc, iters, info = _global_move(elph, pff, lambda: elph.scale_x(mode, mode, -1.0), 0.0, preconditioner,
                                    elph.fdm.tol if tol is None else tol, elph.fdm.maxiter if maxiter is None else maxiter,
                                    rd.get("R"), rd.get("u"), rng)
    reflection_update.last = dict(info, mode=mode)
    return acc, iters


def swap_update(elph, pff, rng=None, preconditioner=None, tol=None, maxiter=None, phonon_type_pairs=None, randoms=None):
    """swap_update!(...) -> (accepted, iters)  (src/swap_update.jl:22-177): exchange the fields of two random modes.
    `randoms` = {"modes": (i, j), "R": ..., "u": ...}."""
    rng = np.random.default_rng() if rng is None else rng
    rd = randoms or {}
    i, j = rd["modes"] if "modes" in rd else _sample_phonon_mode_pair(rng, elph.model, phonon_type_pairs)
    acc, iters, info = _global_move(elph, pff, lambda: elph.swap_x(i, j), 0.0, preconditioner,
                                    elph.fdm.tol if tol is None else tol, elph.fdm.maxiter if maxiter is None else maxiter,
                                    rd.get("R"), rd.get("u"), rng)
    swap_update.last = dict(info, modes=(i, j))
    return acc, iters


def radial_update(elph, pff, rng=None, preconditioner=None, tol=None, maxiter=None, phonon_id=None, sigma=1.0, randoms=None):
    """radial_update!(...; phonon_id, σ) -> (accepted, iters)  (src/radial_update.jl:23-195): x' -> e^γ x' for all modes (or all
    modes of one phonon type), γ ~ N(0, σ²/d), d = (# finite-mass modes) Lτ, acceptance min(1, exp(-ΔS + d γ)).
    `randoms` = {"gamma_normal": standard normal, "R": ..., "u": ...}."""
    rng = np.random.default_rng() if rng is None else rng
    rd = randoms or {}
    m = elph.model
    ncell = m.n_unit_cells
    first, last = (0, m.Nph - 1) if phonon_id is None else (int(phonon_id) * ncell, (int(phonon_id) + 1) * ncell - 1)
    d = int(np.count_nonzero(np.isfinite(m.Mass[first:last + 1]))) * m.Ltau
    g = rd["gamma_normal"] if "gamma_normal" in rd else rng.standard_normal()
    gamma = float(g) * sigma / np.sqrt(d)
    acc, iters, info = _global_move(elph, pff, lambda: elph.scale_x(first, last, float(np.exp(gamma))), d * gamma, preconditioner,
                                    elph.fdm.tol if tol is None else tol, elph.fdm.maxiter if maxiter is None else maxiter,
                                    rd.get("R"), rd.get("u"), rng)
    radial_update.last = dict(info, gamma=gamma, d=d)
    return acc, iters
