"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, fails loudly
without a GPU (no CPU fallback), and the host-side logic (model tables, multi-GPU plans) is right."""
import os
import subprocess
import sys

import numpy as np
import pytest

from smoqyelph_b200 import model as mdl, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from smoqyelph_b200 import lib
    L = lib.load()
    declared = lib.header_symbols()
    assert len(declared) >= 50
    bound = set(lib.SIGNATURES) | set(lib.SPECIAL)
    assert set(declared) == bound, (set(declared) ^ bound)
    assert lib.MISSING == []
    for name in declared:
        assert hasattr(L, name), name
    assert L.sq_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device every constructor fails with a clear error (skipped on a GPU box)."""
    from smoqyelph_b200 import lib, api
    if lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.SqError, match="no CPU fallback"):
        api.FermionDetMatrix(mdl.config("cfg1t"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "smoqyelphqmc.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("oracle/ref_c.c", "").replace("the oracle", "").replace("CPU oracle", ""), fn


@pytest.mark.parametrize("name,N,Nh,C,L,Nph", [("cfg1t", 18, 27, 3, 20, 18), ("cfg1", 18, 27, 3, 80, 18), ("cfg2", 64, 64, 2, 320, 64),
                                               ("cfg3", 256, 512, 4, 200, 768), ("cfg4", 1024, 2048, 4, 400, 1024),
                                               ("cfg5", 1152, 1728, 3, 80, 1152)])
def test_named_configs_match_survey_table(name, N, Nh, C, L, Nph):
    m = mdl.config(name)
    assert (m.N, m.Nh, len(m.colors), m.Ltau, m.Nph) == (N, Nh, C, L, Nph)
    # checkerboard validity and the permutation being a permutation
    assert sorted(m.perm.tolist()) == list(range(Nh))
    for lo, hi in m.colors:
        s = m.nt_chk[:, lo:hi].ravel()
        assert len(set(s.tolist())) == len(s)
    assert np.array_equal(m.nt_chk, m.neighbor_table[:, m.perm])
    # every site has the lattice coordination number
    deg = np.bincount(m.neighbor_table.ravel(), minlength=N)
    assert deg.min() == deg.max()


def test_frozen_modes_and_thermal_fields():
    m = mdl.config("cfg3")
    rng = np.random.default_rng(0)
    x = mdl.thermal_fields(m, rng)
    assert x.shape == (m.Nph, m.Ltau) and np.all(x[~np.isfinite(m.Mass)] == 0)
    # free-phonon equal-time variance: <x^2> = (1/L) sum_w 1/(dtau M (Om^2 + 4 sin^2(pi w/L)/dtau^2))
    w = np.arange(m.Ltau)
    var = np.mean(1.0 / (m.dtau * (1.0 + 4 * np.sin(np.pi * w / m.Ltau) ** 2 / m.dtau ** 2)))
    got = x[np.isfinite(m.Mass)].var()
    assert abs(got - var) < 0.05 * var


def test_slab_partition_and_halo_plan():
    for Lt, world in [(400, 8), (400, 3), (20, 8), (7, 2)]:
        cover = []
        for r in range(world):
            lo, hi = parallel.slab_range(Lt, world, r)
            cover += list(range(lo, hi))
            plan = parallel.halo_plan(Lt, world, r)
            assert plan["prev"] == (r - 1) % world and plan["next"] == (r + 1) % world
            assert plan["sign_from_prev"] == (1.0 if lo == 0 else -1.0)
            assert plan["sign_from_next"] == (1.0 if hi == Lt else -1.0)
        assert cover == list(range(Lt))
        sizes = [np.diff(parallel.slab_range(Lt, world, r))[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SQ_ROOT"])
import numpy as np
import torch.distributed as dist
import smoqyelph_b200
from smoqyelph_b200 import parallel
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
seeds = {parallel.chain_seed(100, q, c) for q in range(w) for c in range(5)}
assert len(seeds) == 5 * w and parallel.chain_seed(100, r) != parallel.chain_seed(101, r)      # hashed, no collisions between chains / components
mean, err = parallel.merge_chain_statistics([1.0 + r, 10.0 * (r + 1)], dist)
assert np.allclose(mean, [1.5, 15.0]) and np.allclose(err, [0.5, 5.0]), (mean, err)
t = parallel.max_over_ranks(1.0 + r, dist)
assert t == 2.0
# tau-slab halo exchange emulated with gloo send/recv on a toy vector: M v assembled from slabs == serial M v
L, N = 6, 3
rng = np.random.default_rng(0)
v = rng.standard_normal((L, N)); B = rng.standard_normal((L, N))          # diagonal toy propagators
full = v.copy(); full[0] += B[0] * v[L - 1]; full[1:] -= B[1:] * v[:-1]
plan = parallel.halo_plan(L, w, r)
lo, hi = plan["lo"], plan["hi"]
import torch
send = torch.from_numpy(v[hi - 1].copy()); recv = torch.zeros(N, dtype=torch.float64)
reqs = [dist.isend(send, plan["next"]), dist.irecv(recv, plan["prev"])]
for q in reqs: q.wait()
halo = recv.numpy()
mine = v[lo:hi].copy()
prev = np.vstack([halo[None, :], v[lo:hi - 1]])
sg = np.full(hi - lo, -1.0); sg[0] = plan["sign_from_prev"]
mine += sg[:, None] * B[lo:hi] * prev
assert np.allclose(mine, full[lo:hi])
# the all-to-all of the tau-FFT preconditioner (slab.cu, kpm_ldiv_slab) emulated with gloo: every rank transforms its own slices into
# partial sums for all frequencies, the pieces of rank q's frequencies travel to rank q and are added in rank order; a per-frequency
# operation on the own frequencies; the same on the way back.  Result == the serial  ifft(f * fft(v))  on the rank's slices.
Lt, Ns = 10, 4
vv = rng.standard_normal((Lt, Ns)) + 1j * rng.standard_normal((Lt, Ns))
fmul = 1.0 + np.arange(Lt)[:, None] * (0.5 + 0.1j)                               # stands in for the per-frequency Chebyshev stage
want = np.fft.ifft(fmul * np.fft.fft(vv, axis=0), axis=0)
lo, hi = parallel.slab_range(Lt, w, r)
def exchange_sum(partial):
    pieces = [None] * w
    dist.all_gather_object(pieces, partial)                                      # (the library sends only the destination's rows)
    acc = np.zeros((hi - lo, Ns), complex)
    for q in range(w):                                                           # fixed rank order: deterministic sum
        acc += pieces[q][lo:hi]
    return acc
freq = exchange_sum(parallel.partial_dft(vv[lo:hi], lo, Lt))                     # own frequencies, all slices summed
freq = fmul[lo:hi] * freq
back = exchange_sum(parallel.partial_dft(freq, lo, Lt, inverse=True)) / Lt       # own slices, all frequencies summed
assert np.allclose(back, want[lo:hi], atol=1e-12)
sched = [0, Lt - 1, 1, Lt - 2, 2]
share = parallel.frequency_share(sched, Lt, w, r)
allsh = [None] * w
dist.all_gather_object(allsh, share)
assert sorted(sum(allsh, [])) == sorted(sched) and all(lo <= n < hi for n in share)
assert [parallel.rhs_owner(j, w) for j in range(5)] == [0, 1, 0, 1, 0][:5] if w == 2 else True
dist.destroy_process_group()
print("ok", r)
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SQ_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_global_move_sampling_helpers():
    """Host logic of the global moves (no GPU): phonon layout helpers and the mode / pair sampling rules."""
    import numpy as np
    from smoqyelph_b200 import api, model as mdl
    m = mdl.bssh_square(4, 4, 0.5)                       # 3 phonon types per cell, the third frozen (M = inf)
    assert m.n_unit_cells == 16 and m.nphonon == 3
    rng = np.random.default_rng(0)
    modes = [api._sample_phonon_mode(rng, m) for _ in range(200)]
    assert all(np.isfinite(m.Mass[p]) for p in modes) and {p // 16 for p in modes} == {0, 1}
    assert all(api._sample_phonon_mode(rng, m, phonon_types=[1]) // 16 == 1 for _ in range(50))
    pairs = [api._sample_phonon_mode_pair(rng, m, phonon_type_pairs=[(0, 1)]) for _ in range(50)]
    assert all(i // 16 == 0 and j // 16 == 1 for i, j in pairs)
    h = mdl.holstein_honeycomb(3, 1.0)
    assert h.n_unit_cells == 9 and h.nphonon == 2
    import pytest
    with pytest.raises(api.SqError):
        api._sample_phonon_mode(rng, m, phonon_types=[2])    # only frozen modes: nothing to propose


def test_julia_shim_binds_only_declared_symbols_with_matching_arity():
    """julia/SmoQyElPhB200.jl cannot be executed here (no Julia in the image): at least every `ccall` in it must name a function
    the header declares, with as many argument types as the C prototype has parameters."""
    import re
    from smoqyelph_b200 import lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "julia", "SmoQyElPhB200.jl"), encoding="utf-8").read()
    hdr = re.sub(r"/\*.*?\*/", "", open(lib.HEADER).read(), flags=re.S)
    protos = {}
    for mm in re.finditer(r"\b(sq_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = mm.group(2).strip()
        protos[mm.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    calls = list(re.finditer(r"ccall\(\(:(sq_[A-Za-z0-9_]+), LIB\),\s*[A-Za-z{}\.]+,\s*\(", src))
    assert len(calls) > 30
    for mm in calls:
        name = mm.group(1)
        assert name in protos, f"{name} is not declared in the header"
        # the argument-type tuple: balanced parentheses from the match end
        i, depth, start = mm.end(), 1, mm.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[start:i - 1]
        # split on top-level commas only
        parts, d, cur = [], 0, ""
        for ch in tup:
            if ch in "({[":
                d += 1
            elif ch in ")}]":
                d -= 1
            if ch == "," and d == 0:
                parts.append(cur); cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        assert len(parts) == protos[name], f"{name}: {len(parts)} argument types in the shim, {protos[name]} parameters in the header"


def test_clock_sampler_degrades_without_a_gpu():
    """bench.ClockSampler must never raise: NVML first, nvidia-smi second, empty result otherwise."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    out = bench.ClockSampler(0).stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_make_measurements_accumulates_like_the_reference_container():
    """Host logic of api.make_measurements (make_measurements.jl:19-146) on a stub estimator: values are ADDED per call, the per-orbital
    entries are vectors, correlations are keyed by their arguments, and the estimator is refreshed first."""
    from smoqyelph_b200 import api

    class Stub:
        def __init__(self):
            self.calls = []

        def update_greens_estimator(self, preconditioner=None, tol=None, maxiter=None):
            self.calls.append("update")
            return 7.5

        def _geom(self, norb, dims):
            return 2, (3, 3)

        def measure(self):
            self.calls.append("measure")
            return {"n": 0.5 + 0j, "double_occ": 0.25 + 0j, "Nsqrd": 9.0 + 0j}

        def measure_n_orbital(self, a, norb, dims):
            return 0.4 + 0.1 * a

        def measure_double_occ_orbital(self, a, norb, dims):
            return 0.1 + 0.05 * a

        def measure_GD0(self, orbitals, norb=None, dims=None):
            return np.full((3, 3, 5), orbitals[0] + 10.0 * orbitals[1], complex)

        def measure_density_correlation(self, a, b, norb=None, dims=None):
            return np.ones((3, 3, 5), complex)

    g = Stub()
    meas = {}
    corr = [("greens", (0, 1)), ("density", (1, 1)), ("composite", ("tr_greens", "greens", [(0, 0), (1, 1)], [1.0, -2.0]))]
    for k in range(3):
        it = api.make_measurements(meas, None, g, mu=0.3, bosonic_action=lambda: 2.0, correlations=corr)
        assert it == 7.5 and g.calls[2 * k] == "update" and g.calls[2 * k + 1] == "measure"
    G = meas["global"]
    assert G["sgn"] == 3.0 and abs(G["density"] - 3.0) < 1e-14 and abs(G["density_up"] - 1.5) < 1e-14 and abs(G["double_occ"] - 0.75) < 1e-14
    assert abs(G["Nsqrd"] - 27.0) < 1e-14 and abs(G["chemical_potential"] - 0.9) < 1e-14 and abs(G["action_bosonic"] - 6.0) < 1e-14
    L = meas["local"]
    assert np.allclose(L["density"], [3 * 0.8, 3 * 1.0]) and np.allclose(L["double_occ"], [0.3, 0.45]) and L["density_up"].shape == (2,)
    C = meas["correlations"]
    assert np.allclose(C[("greens", 0, 1)], 30.0) and np.allclose(C[("density", 1, 1)], 3.0)
    assert np.allclose(C[("composite", "tr_greens")], 3 * (1.0 * 0.0 - 2.0 * 11.0))        # sum_k coefficient_k x G_(a_k, b_k)
    with pytest.raises(ValueError):
        api.make_measurements(meas, None, g, correlations=[("nonsense", (0, 0))])
