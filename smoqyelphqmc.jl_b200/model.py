"""Host-side model tables for the hot path (numpy only, no GPU).

These builders stand in for the SmoQyDQMC model DSL that stays in Julia for the drop-in
(SURVEY.md 8b): they produce exactly the arrays the C ABI consumes -- the neighbour table, a
checkerboard decomposition (permutation + colour ranges, as `checkerboard_decomposition!` returns
them at /root/reference/src/FermionDetMatrix.jl:95-97), bare hoppings / on-site energies, phonon
parameters and the Holstein / SSH coupling maps read at
/root/reference/src/fermion_det_matrix_dervative.jl:208-211,266-269 and
/root/reference/src/holstein_shift_matrix.jl:7-8.

All indices are 0-based in Python; `lib.py` converts to the 1-based tables of the C ABI.
The named configurations of BASELINE.json are provided by `config(name)`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


@dataclass
class Model:
    name: str
    beta: float
    dtau: float
    N: int                       # sites (orbitals)
    neighbor_table: np.ndarray   # (2, Nh) original hopping order
    t0: np.ndarray               # (Nh,) bare hopping amplitudes
    V0: np.ndarray               # (N,) on-site energy minus chemical potential
    # phonons (PhononParameters): type-major over unit cells, as SmoQyDQMC lays them out
    Omega: np.ndarray
    Omega4: np.ndarray
    Mass: np.ndarray
    # Holstein couplings
    hol_phonon: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    hol_site: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    hol_alpha: np.ndarray = field(default_factory=lambda: np.zeros((4, 0)))   # rows: a, a2, a3, a4
    hol_phsym: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    # SSH couplings
    ssh_phonon: np.ndarray = field(default_factory=lambda: np.zeros((2, 0), np.int64))
    ssh_hopping: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    ssh_alpha: np.ndarray = field(default_factory=lambda: np.zeros((4, 0)))
    # dispersive phonon couplings (DispersionParameters): phonon pairs (2, Ndisp), frequencies per coupling
    disp_phonon: np.ndarray = field(default_factory=lambda: np.zeros((2, 0), np.int64))
    disp_Omega: np.ndarray = field(default_factory=lambda: np.zeros(0))
    disp_Omega4: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # checkerboard decomposition (filled by finalize)
    perm: np.ndarray | None = None          # checkerboard index -> original hopping index
    nt_chk: np.ndarray | None = None        # (2, Nh) permuted neighbour table
    colors: list | None = None              # [(lo, hi)) ranges in checkerboard order
    lattice_dims: tuple = ()

    @property
    def Ltau(self) -> int:
        return int(round(self.beta / self.dtau))

    @property
    def Nh(self) -> int:
        return self.neighbor_table.shape[1]

    @property
    def Nph(self) -> int:
        return self.Omega.shape[0]

    @property
    def n_unit_cells(self) -> int:
        return int(np.prod(self.lattice_dims)) if self.lattice_dims else self.N

    @property
    def nphonon(self) -> int:
        """Phonon modes per unit cell (PhononParameters.nphonon); phonons are laid out type-major."""
        return self.Nph // self.n_unit_cells

    @property
    def Nhol(self) -> int:
        return self.hol_phonon.shape[0]

    @property
    def Nssh(self) -> int:
        return self.ssh_hopping.shape[0]

    @property
    def Ndisp(self) -> int:
        return self.disp_Omega.shape[0]

    def finalize(self, bond_colors=None) -> "Model":
        """Checkerboard decomposition: any colouring in which bonds of one colour share no site."""
        nt = self.neighbor_table
        Nh = nt.shape[1]
        if bond_colors is None:
            bond_colors = greedy_edge_colouring(nt, self.N)
        bond_colors = np.asarray(bond_colors)
        perm = np.argsort(bond_colors, kind="stable").astype(np.int64)
        self.perm = perm
        self.nt_chk = np.ascontiguousarray(nt[:, perm])
        sorted_c = bond_colors[perm]
        self.colors = []
        for c in np.unique(sorted_c):
            idx = np.nonzero(sorted_c == c)[0]
            self.colors.append((int(idx[0]), int(idx[-1]) + 1))
        for lo, hi in self.colors:     # validity: no site twice inside a colour
            sites = self.nt_chk[:, lo:hi].ravel()
            assert len(np.unique(sites)) == len(sites), "invalid checkerboard colouring"
        assert Nh == 0 or self.colors[-1][1] == Nh
        return self

    def random_fields(self, rng, amplitude=0.5, smooth=False) -> np.ndarray:
        """Synthetic phonon field x (Nph, Ltau): i.i.d. N(0,1)*amplitude, or a
        tau-smooth variant resembling thermalised fields.  Frozen (M = inf) modes stay 0."""
        L = self.Ltau
        if smooth:
            x = amplitude * rng.standard_normal((self.Nph, 1)) + 0.1 * amplitude * rng.standard_normal((self.Nph, L))
        else:
            x = amplitude * rng.standard_normal((self.Nph, L))
        x[~np.isfinite(self.Mass), :] = 0.0
        return np.asfortranarray(x)


def thermal_fields(model: Model, rng) -> np.ndarray:
    """Phonon field drawn from the FREE-phonon thermal distribution exp(-S_b) (synthetic but physical
    starting point for trajectory benchmarks): every Matsubara mode w of phonon p is Gaussian with
    variance 1 / (dtau M (Omega^2 + 4 sin^2(pi w / L) / dtau^2))."""
    L = model.Ltau
    w = np.arange(L)
    R = rng.standard_normal((model.Nph, L))
    Rt = np.fft.fft(R, axis=1) / np.sqrt(L)
    k = model.dtau * model.Mass[:, None] * (model.Omega[:, None] ** 2 + 4 * np.sin(np.pi * w / L)[None, :] ** 2 / model.dtau ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        xt = np.where(np.isfinite(k), Rt / np.sqrt(k), 0.0)
    x = np.real(np.fft.ifft(xt, axis=1) * np.sqrt(L))
    x[~np.isfinite(model.Mass), :] = 0.0
    return np.asfortranarray(x)


def greedy_edge_colouring(nt: np.ndarray, N: int) -> np.ndarray:
    Nh = nt.shape[1]
    used = [set() for _ in range(N)]
    col = np.zeros(Nh, np.int64)
    for h in range(Nh):
        i, j = int(nt[0, h]), int(nt[1, h])
        c = 0
        while c in used[i] or c in used[j]:
            c += 1
        col[h] = c
        used[i].add(c)
        used[j].add(c)
    return col


# ------------------------------------------------------------------------------------------------
# lattices
# ------------------------------------------------------------------------------------------------
def _chain_bonds(n):
    i = np.arange(n)
    nt = np.stack([i, (i + 1) % n])
    col = (i % 2) if n % 2 == 0 else None
    return nt, col


def _square_bonds(Lx, Ly):
    x, y = np.meshgrid(np.arange(Lx), np.arange(Ly), indexing="ij")
    site = lambda a, b: (a % Lx) + Lx * (b % Ly)
    s = site(x, y).ravel(order="F")
    xs, ys = x.ravel(order="F"), y.ravel(order="F")
    bx = np.stack([s, site(xs + 1, ys)])
    by = np.stack([s, site(xs, ys + 1)])
    nt = np.concatenate([bx, by], axis=1)
    col = None
    if Lx % 2 == 0 and Ly % 2 == 0:
        col = np.concatenate([xs % 2, 2 + ys % 2])
    return nt, col, bx.shape[1]


def _honeycomb_bonds(L1, L2):
    """Two orbitals (A=0, B=1) per cell, site = orb + 2*(c1 + L1*c2).  Three bond types
    A(r)->B(r), A(r)->B(r-a1), A(r)->B(r-a2); each type is a perfect matching => 3 colours."""
    c1, c2 = np.meshgrid(np.arange(L1), np.arange(L2), indexing="ij")
    c1, c2 = c1.ravel(order="F"), c2.ravel(order="F")
    cell = lambda a, b: (a % L1) + L1 * (b % L2)
    A = 2 * cell(c1, c2)
    nts, cols = [], []
    for k, (d1, d2) in enumerate([(0, 0), (-1, 0), (0, -1)]):
        B = 2 * cell(c1 + d1, c2 + d2) + 1
        nts.append(np.stack([A, B]))
        cols.append(np.full(A.shape, k))
    return np.concatenate(nts, axis=1), np.concatenate(cols)


# ------------------------------------------------------------------------------------------------
# model families of the named configurations
# ------------------------------------------------------------------------------------------------
def holstein_square(Lx, Ly, beta, dtau=0.05, Omega=1.0, alpha=1.5, mu=0.0, t=1.0, ph_sym=True, name=None):
    """Config 4 family (tutorial-style Holstein model on a square lattice)."""
    nt, col, _ = _square_bonds(Lx, Ly)
    N = Lx * Ly
    m = Model(name or f"holstein_square_{Lx}x{Ly}_b{beta:g}", beta, dtau, N, nt.astype(np.int64), np.full(nt.shape[1], t),
              np.full(N, -mu), np.full(N, Omega), np.zeros(N), np.ones(N),
              hol_phonon=np.arange(N, dtype=np.int64), hol_site=np.arange(N, dtype=np.int64),
              hol_alpha=np.stack([np.full(N, alpha), np.zeros(N), np.zeros(N), np.zeros(N)]),
              hol_phsym=np.full(N, int(ph_sym), np.int32), lattice_dims=(Lx, Ly))
    return m.finalize(col)


def holstein_honeycomb(L, beta, dtau=0.05, Omega=1.0, alpha=1.5, mu=0.0, t=1.0, ph_sym=True, name=None):
    """Configs 1 / 5: /root/reference/tutorials/holstein_honeycomb.jl:155-430."""
    nt, col = _honeycomb_bonds(L, L)
    N = 2 * L * L
    # phonons / couplings are type-major: orbital A of every cell, then orbital B of every cell
    site_of = np.concatenate([2 * np.arange(L * L), 2 * np.arange(L * L) + 1]).astype(np.int64)
    m = Model(name or f"holstein_honeycomb_{L}x{L}_b{beta:g}", beta, dtau, N, nt.astype(np.int64), np.full(nt.shape[1], t),
              np.full(N, -mu), np.full(N, Omega), np.zeros(N), np.ones(N),
              hol_phonon=np.arange(N, dtype=np.int64), hol_site=site_of,
              hol_alpha=np.stack([np.full(N, alpha), np.zeros(N), np.zeros(N), np.zeros(N)]),
              hol_phsym=np.full(N, int(ph_sym), np.int32), lattice_dims=(L, L))
    return m.finalize(col)


def ossh_chain(n, beta, dtau=0.05, Omega=1.0, alpha=0.5, mu=0.0, t=1.0, name=None):
    """Config 2: /root/reference/examples/ossh_chain.jl:113-178 -- optical SSH, one phonon per site,
    bond i->i+1 couples phonons (i, i+1): t_eff = t - alpha (X_{i+1} - X_i)."""
    nt, col = _chain_bonds(n)
    i = np.arange(n, dtype=np.int64)
    m = Model(name or f"ossh_chain_{n}_b{beta:g}", beta, dtau, n, nt.astype(np.int64), np.full(n, t), np.full(n, -mu),
              np.full(n, Omega), np.zeros(n), np.ones(n),
              ssh_phonon=np.stack([i, (i + 1) % n]), ssh_hopping=i.copy(),
              ssh_alpha=np.stack([np.full(n, alpha), np.zeros(n), np.zeros(n), np.zeros(n)]), lattice_dims=(n,))
    return m.finalize(col)


def bssh_square(Lx, Ly, beta, dtau=0.05, Omega=1.0, alpha=0.5, mu=0.0, t=1.0, name=None):
    """Config 3: /root/reference/examples/bssh_square.jl:173-239 -- bond SSH: per cell an x-bond
    phonon, a y-bond phonon and a frozen (M = inf) phonon; couplings pair (frozen, bond phonon)."""
    nt, col, nbx = _square_bonds(Lx, Ly)
    N = Lx * Ly
    cells = np.arange(N, dtype=np.int64)
    Nph = 3 * N
    Mass = np.concatenate([np.ones(N), np.ones(N), np.full(N, np.inf)])
    ssh_phonon = np.concatenate([np.stack([2 * N + cells, cells]), np.stack([2 * N + cells, N + cells])], axis=1)
    ssh_hopping = np.concatenate([cells, nbx + cells])
    m = Model(name or f"bssh_square_{Lx}x{Ly}_b{beta:g}", beta, dtau, N, nt.astype(np.int64), np.full(nt.shape[1], t),
              np.full(N, -mu), np.full(Nph, Omega), np.zeros(Nph), Mass,
              ssh_phonon=ssh_phonon.astype(np.int64), ssh_hopping=ssh_hopping.astype(np.int64),
              ssh_alpha=np.stack([np.full(2 * N, alpha), np.zeros(2 * N), np.zeros(2 * N), np.zeros(2 * N)]),
              lattice_dims=(Lx, Ly))
    return m.finalize(col)


def holstein_ssh_chain(n, beta, dtau=0.05, name=None):
    """Not a named config: a small model with BOTH coupling kinds, all polynomial orders non-zero,
    a non-ph-symmetric Holstein coupling, disordered hoppings and an odd colour count -- used to
    exercise every branch of the force in the parity tests."""
    nt, _ = _chain_bonds(n)
    i = np.arange(n, dtype=np.int64)
    rng = np.random.default_rng(1234)
    m = Model(name or f"holstein_ssh_chain_{n}_b{beta:g}", beta, dtau, n, nt.astype(np.int64), 1.0 + 0.2 * rng.standard_normal(n),
              0.3 * rng.standard_normal(n), np.full(n, 1.0), np.full(n, 0.4), np.ones(n),
              hol_phonon=i.copy(), hol_site=(i[::-1]).copy(),
              hol_alpha=np.stack([np.full(n, 1.1), np.full(n, 0.2), np.full(n, 0.05), np.full(n, 0.01)]),
              hol_phsym=(i % 2).astype(np.int32),
              ssh_phonon=np.stack([i, (i + 1) % n]), ssh_hopping=i.copy(),
              ssh_alpha=np.stack([np.full(n, 0.4), np.full(n, 0.1), np.full(n, 0.03), np.full(n, 0.01)]), lattice_dims=(n,))
    return m.finalize(None)


def with_dispersion(m: Model, Omega_d=0.5, Omega4_d=0.0) -> Model:
    """Adds nearest-neighbour dispersive phonon couplings (SmoQyDQMC PhononDispersion: a spring between the phonons of the two ends
    of every bond that carries a phonon on each site) to a Holstein-type model: one coupling per bond of the original neighbour table."""
    ph_of_site = {int(s): int(p) for p, s in zip(m.hol_phonon, m.hol_site)}
    pairs = [(ph_of_site[int(i)], ph_of_site[int(j)]) for i, j in m.neighbor_table.T if int(i) in ph_of_site and int(j) in ph_of_site]
    m.disp_phonon = np.ascontiguousarray(np.array(pairs, np.int64).T.reshape(2, -1))
    m.disp_Omega = np.full(len(pairs), float(Omega_d))
    m.disp_Omega4 = np.full(len(pairs), float(Omega4_d))
    m.name += "_disp"
    return m


def config(name: str) -> Model:
    """Named configurations of BASELINE.json / SURVEY.md 8(d)."""
    table = {
        "cfg1t": lambda: holstein_honeycomb(3, 1.0, name="cfg1t"),
        "cfg1": lambda: holstein_honeycomb(3, 4.0, name="cfg1"),
        "cfg2": lambda: ossh_chain(64, 16.0, name="cfg2"),
        "cfg3": lambda: bssh_square(16, 16, 10.0, name="cfg3"),
        "cfg4": lambda: holstein_square(32, 32, 20.0, name="cfg4"),
        "cfg5": lambda: holstein_honeycomb(24, 4.0, name="cfg5"),
        "h16": lambda: holstein_square(16, 16, 2.0, name="h16"),                      # small register-path lattice (multi-GPU tests)
        "hc8": lambda: holstein_honeycomb(8, 2.0, name="hc8"),
        "b40": lambda: holstein_square(32, 32, 40.0, name="b40"),                     # Ltau = 800: more slices than one GPU keeps resident
        "b80": lambda: holstein_square(32, 32, 80.0, name="b80"),                     # Ltau = 1600
        "sweep64": lambda: holstein_square(64, 64, 20.0, name="sweep64"),
        "sweep128": lambda: holstein_square(128, 128, 20.0, name="sweep128"),
    }
    return table[name]()
