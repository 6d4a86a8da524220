"""Feasibility study for the next solver step, on the CPU oracle's operator: pipelined CG
(Ghysels-Vanroose), which lets the grid-wide sums travel while the next M^T M product runs, takes the SAME number of iterations
as the reference recurrence at the force tolerance (1e-5: 24 of the 25 solves of a trajectory) without any safeguard, and at the
action tolerance (1e-10) once the recurrence residual is replaced by the true one every 100 iterations; without the replacement it
stagnates above the tolerance (measured 1e-9 ... 1e-6 on larger lattices).  Not a product path: a documented property test."""
import numpy as np

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc
import dense_ref as dr


def cg(A, b, tol, maxiter):
    x = np.zeros_like(b); r = b.copy(); p = r.copy()
    rr = np.vdot(r, r).real; nb = np.sqrt(rr)
    for it in range(1, maxiter + 1):
        Ap = A(p)
        alpha = rr / np.vdot(p, Ap).real
        x += alpha * p; r -= alpha * Ap
        rr_new = np.vdot(r, r).real
        if np.sqrt(rr_new) / nb < tol:
            return x, it
        p = r + (rr_new / rr) * p; rr = rr_new
    return x, maxiter


def pipelined_cg(A, b, tol, maxiter, replace_every=0):
    x = np.zeros_like(b); r = b.copy(); w = A(r); nb = np.sqrt(np.vdot(b, b).real)
    z = np.zeros_like(b); s = np.zeros_like(b); p = np.zeros_like(b)
    gamma_old = alpha_old = 1.0
    for it in range(1, maxiter + 1):
        gamma = np.vdot(r, r).real; delta = np.vdot(w, r).real      # one reduction ...
        if np.sqrt(gamma) / nb < tol:
            return x, it - 1
        q = A(w)                                                    # ... that overlaps with this product
        beta = gamma / gamma_old if it > 1 else 0.0
        alpha = gamma / (delta - beta * gamma / alpha_old) if it > 1 else gamma / delta
        z = q + beta * z; s = w + beta * s; p = r + beta * p
        x += alpha * p; r -= alpha * s; w -= alpha * z
        if replace_every and it % replace_every == 0:               # residual replacement
            r = b - A(x); w = A(r); s = A(p); z = A(s)
        gamma_old, alpha_old = gamma, alpha
    return x, maxiter


def test_pipelined_cg_keeps_the_iteration_counts_of_the_reference_recurrence():
    m = mdl.holstein_square(16, 16, 3.0)
    rng = np.random.default_rng(0)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    ref = orc.RefFDM(m, sym=True)
    ref.update(V, t)
    A = ref.mul_MtM
    b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    res = lambda x: np.linalg.norm(A(x) - b) / np.linalg.norm(b)
    x1, i1 = cg(A, b, 1e-5, 5000)
    x2, i2 = pipelined_cg(A, b, 1e-5, 5000)
    assert abs(i1 - i2) <= 1 and res(x2) < 1.1e-5
    x1, i1 = cg(A, b, 1e-10, 5000)
    x2, i2 = pipelined_cg(A, b, 1e-10, 5000, replace_every=100)
    assert abs(i1 - i2) <= 2 and res(x2) < 1.1e-10 and res(x1) < 1.1e-10
