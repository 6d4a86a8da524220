"""Independent dense-matrix statement of the operator, used to pin the oracle (KATs).

Nothing here shares code with oracle/ref_c.c: matrices are assembled from the DEFINITIONS in
SURVEY.md section 10 (B_l = Gamma_l D_l Gamma_l^T, M = I - subdiagonal B_l + antiperiodic corner),
indexed as a dense (Ltau*N) x (Ltau*N) matrix acting on Fortran-flattened (Ltau, N) vectors.
"""
import numpy as np


def bond_factor(N, i, j, c, s):
    F = np.eye(N)
    F[i, i] = c
    F[j, j] = c
    F[i, j] = s
    F[j, i] = s
    return F


def gamma(model, cosh_l, sinh_l):
    """Gamma_l = F_Nh ... F_1 (bond 1 acts first), checkerboard order."""
    N = model.N
    G = np.eye(N)
    for h in range(model.Nh):
        i, j = model.nt_chk[:, h]
        G = bond_factor(N, i, j, cosh_l[h], sinh_l[h]) @ G
    return G


def propagators(model, V, t, sym=True):
    """List of dense B_l from V (N, Ltau), t (Nh, Ltau) in ORIGINAL hopping order."""
    dt = model.dtau
    dtp = dt / 2 if sym else dt
    Bs = []
    for l in range(model.Ltau):
        tl = t[model.perm, l]
        G = gamma(model, np.cosh(dtp * np.abs(tl)), np.sign(tl) * np.sinh(dtp * np.abs(tl)))
        D = np.diag(np.exp(-dt * V[:, l]))
        Bs.append(G @ D @ G.T if sym else D @ G)
    return Bs


def dense_M(model, Bs):
    L, N = model.Ltau, model.N
    M = np.zeros((L, N, L, N))          # M[l, i, l', j]
    for l in range(L):
        M[l, :, l, :] += np.eye(N)
        if l == 0:
            M[0, :, L - 1, :] += Bs[0]
        else:
            M[l, :, l - 1, :] -= Bs[l]
    # Fortran flattening of (Ltau, N): index = l + i*L
    return M.transpose(1, 0, 3, 2).reshape(N * L, N * L)


def flat(v):
    return np.asarray(v).reshape(-1, order="F")


def unflat(model, v):
    return np.asarray(v).reshape((model.Ltau, model.N), order="F")


def build_Vt(model, x):
    """V (N, Ltau), t (Nh, Ltau) from the phonon field (definition in SURVEY.md 10 / ref_c.c)."""
    L = model.Ltau
    V = np.repeat(model.V0[:, None], L, axis=1).astype(float)
    t = np.repeat(model.t0[:, None], L, axis=1).astype(float)
    for c in range(model.Nhol):
        xp = x[model.hol_phonon[c], :]
        a = model.hol_alpha[:, c]
        V[model.hol_site[c], :] += a[0] * xp + a[1] * xp**2 + a[2] * xp**3 + a[3] * xp**4
    for c in range(model.Nssh):
        dx = x[model.ssh_phonon[1, c], :] - x[model.ssh_phonon[0, c], :]
        a = model.ssh_alpha[:, c]
        t[model.ssh_hopping[c], :] -= a[0] * dx + a[1] * dx**2 + a[2] * dx**3 + a[3] * dx**4
    return np.asfortranarray(V), np.asfortranarray(t)


def dense_Lambda(model, x):
    """Lambda as a dense matrix: (Lambda v)[l] = lam[l+1] * v[l+1] (cyclic)."""
    L, N = model.Ltau, model.N
    lam = -np.ones((L, N))
    lam[0, :] = 1.0
    for c in range(model.Nhol):
        if model.hol_phsym[c]:
            xp = x[model.hol_phonon[c], :]
            lam[:, model.hol_site[c]] *= np.exp(model.dtau * (model.hol_alpha[0, c] * xp + model.hol_alpha[2, c] * xp**3) / 2)
    Lm = np.zeros((L, N, L, N))
    for l in range(L):
        lp = (l + 1) % L
        Lm[l, np.arange(N), lp, np.arange(N)] = lam[lp, :]
    return lam, Lm.transpose(1, 0, 3, 2).reshape(N * L, N * L)


def fermionic_action_dense(model, x, Phi, sym=True):
    """S_f = Phi^dagger (A^T A)^-1 Phi, A = M Lambda, by dense linear algebra."""
    V, t = build_Vt(model, x)
    M = dense_M(model, propagators(model, V, t, sym))
    _, Lm = dense_Lambda(model, x)
    A = M @ Lm
    return float(np.real(np.vdot(flat(Phi), np.linalg.solve(A.T @ A, flat(Phi)))))
