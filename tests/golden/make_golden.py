"""Generate the golden vectors in this directory.

The reference ships no golden vectors for the hot path and cannot run here (no Julia, five un-vendored dependencies), so these
vectors come from the INDEPENDENT dense-matrix statement of the definitions in tests/dense_ref.py (dense M from the checkerboard
propagators, numpy.linalg.solve, central finite differences) -- not from the oracle (oracle/ref_c.c) and not from the CUDA
library, both of which are tested AGAINST them.  They freeze the known answers, so a later change of the oracle, the dense
statement or the kernels that moves any number is caught.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz (deterministic: fixed seeds)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import smoqyelph_b200  # noqa: F401,E402
from smoqyelph_b200 import model as mdl  # noqa: E402
import dense_ref as dr  # noqa: E402

CASES = {
    "holstein_honeycomb_2_b0.3": lambda: mdl.holstein_honeycomb(2, 0.3),
    "ossh_chain_6_b0.4": lambda: mdl.ossh_chain(6, 0.4),
    "bssh_square_2x4_b0.25": lambda: mdl.bssh_square(2, 4, 0.25),
    "holstein_ssh_chain_5_b0.3": lambda: mdl.holstein_ssh_chain(5, 0.3),
    "holstein_square_4x4_b0.2": lambda: mdl.holstein_square(4, 4, 0.2),
}


def build(name):
    m = CASES[name]()
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    x = m.random_fields(rng)
    V, t = dr.build_Vt(m, x)
    v = (rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2)
    out = {"x": x, "v": v, "V": V, "t": t}
    for sym in (True, False):
        M = dr.dense_M(m, dr.propagators(m, V, t, sym))
        tag = "sym" if sym else "asym"
        fv = dr.flat(v)
        out[f"M_v_{tag}"] = dr.unflat(m, M @ fv)
        out[f"Mt_v_{tag}"] = dr.unflat(m, M.T @ fv)
        out[f"MtM_v_{tag}"] = dr.unflat(m, M.T @ (M @ fv))
        out[f"solve_MtM_{tag}"] = dr.unflat(m, np.linalg.solve(M.T @ M, fv))
        out[f"Sf_{tag}"] = np.array(dr.fermionic_action_dense(m, x, v, sym))
    # dS_f/dx by central differences of the dense action (Sym), a handful of components
    comps = [(int(p), int(l)) for p, l in zip(rng.integers(0, m.Nph, 6), rng.integers(0, m.Ltau, 6)) if np.isfinite(m.Mass[p])]
    fd = []
    for p, l in comps:
        h = 1e-5
        xp, xm = x.copy(), x.copy()
        xp[p, l] += h
        xm[p, l] -= h
        fd.append((dr.fermionic_action_dense(m, xp, v, True) - dr.fermionic_action_dense(m, xm, v, True)) / (2 * h))
    out["fd_components"] = np.array(comps, np.int64).reshape(-1, 2)
    out["fd_dSdx_sym"] = np.array(fd)
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
