"""tau-slab partitioning: range logic on one GPU, and the NCCL halo path on 2 GPUs when the box has them."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from smoqyelph_b200 import model as mdl
from oracle import oracle as orc
import dense_ref as dr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "cfg1"])
def test_slab_ranges_tile_the_operator(name):
    """Producing the slices slab by slab (any split, including ragged ones) reproduces the full products."""
    from smoqyelph_b200 import api
    m = mdl.config(name)
    rng = np.random.default_rng(1)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    ref = orc.RefFDM(m)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m)
    fdm.update(V, t)
    v = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    L = m.Ltau
    for cuts in ([0, L // 2, L], [0, 1, L // 3, L - 1, L], [0, 7, L]):
        for op in ("mul_MtM", "mul_M", "mul_Mt"):
            full = np.zeros_like(v)
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                fdm.set_slab_range(lo, hi)
                full[lo:hi] = getattr(fdm, op)(v)[lo:hi]
            assert relerr(full, getattr(ref, op)(v)) < 1e-12, (cuts, op)
    fdm.set_slab_range(0, L)


def test_slab_cg_recurrence_matches_oracle(monkeypatch):
    """The slab CG (local updates + all-reduced scalars) is the reference recurrence: world = 1 through the same code."""
    from smoqyelph_b200 import api
    monkeypatch.setenv("SQ_FORCE_SLAB_CG", "1")
    m = mdl.config("cfg2")
    rng = np.random.default_rng(2)
    V, t = dr.build_Vt(m, m.random_fields(rng))
    ref = orc.RefFDM(m)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m)
    fdm.update(V, t)
    fdm.init_slab(0, 1)
    b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    xr, itr, _ = ref.cg(b, tol=1e-13, maxiter=20000)
    xg, itg, eps = fdm.ldiv(b, tol=1e-13, maxiter=20000)
    assert eps < 1e-13 and relerr(xg, xr) < 1e-10
    for tol in (1e-5, 1e-10):
        _, itr, _ = ref.cg(b, tol=tol, maxiter=20000)
        _, itg, _ = fdm.ldiv(b, tol=tol, maxiter=20000)
        assert abs(itg - itr) <= 1
    _, it0, _ = fdm.ldiv(b, x0=xg, tol=1e-10)
    assert it0 == 0


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "h16"])
def test_slab_preconditioned_cg_single_rank(name, monkeypatch):
    """The tau-slab preconditioned solver (zero-padded partial DFTs, piece sums, per-rank share of the Chebyshev schedule) with
    world = 1 through the same code: solution and iteration counts of the reference's preconditioned recurrence."""
    from smoqyelph_b200 import api
    monkeypatch.setenv("SQ_FORCE_SLAB_CG", "1")
    m = mdl.config(name)
    rng = np.random.default_rng(4)
    V, t = dr.build_Vt(m, m.random_fields(rng, smooth=True))
    ref = orc.RefFDM(m)
    ref.update(V, t)
    fdm = api.FermionDetMatrix(m)
    fdm.update(V, t)
    fdm.init_slab(0, 1)
    Pr = orc.RefKPM(ref)
    Pr.update(rng.standard_normal(m.N))
    assert Pr.active
    Pg = api.KPMPreconditioner(fdm, update=False)
    Pg.set_bounds(*Pr.bounds)
    b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
    xr, itr, _ = ref.cg(b, P=Pr, tol=1e-13, maxiter=5000)
    xg, itg, eps = fdm.ldiv(b, preconditioner=Pg, tol=1e-13, maxiter=5000, refresh=False)
    assert eps < 1e-13 and relerr(xg, xr) < 1e-10
    for tol in (1e-5, 1e-10):
        _, itr, _ = ref.cg(b, P=Pr, tol=tol, maxiter=5000)
        _, itg, _ = fdm.ldiv(b, preconditioner=Pg, tol=tol, maxiter=5000, refresh=False)
        assert abs(itg - itr) <= 1
    assert fdm.stats["cg_slab_preconditioned"] == 3
    _, it0, _ = fdm.ldiv(b, x0=xg, preconditioner=Pg, tol=1e-10, refresh=False)
    assert it0 == 0


def test_two_gpu_halo_exchange_and_cg():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    # cfg3 / cfg1: shared-memory kernels + NCCL loop; h16 / hc8: register path + resident multi-GPU solver over the mailboxes
    for name in ("cfg3", "cfg1", "h16", "hc8"):
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                              "--master-port", "29641", os.path.join(ROOT, "tools", "slab_worker.py"), name, "check", "100"],
                             env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
        out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
        assert out["err_mul_MtM"] < 1e-12 and out["err_mul_M"] < 1e-12 and out["err_mul_Mt"] < 1e-12, out
        assert out["err_cg"] < 1e-10, out
        assert abs(out["iters"][0] - out["iters"][1]) <= 1, out
        # preconditioned: frequency-sharded KPM apply with its two all-to-all exchanges
        assert out["err_cg_kpm"] < 1e-11 * 10, out
        assert abs(out["iters_kpm"][0] - out["iters_kpm"][1]) <= 1, out
        assert out["stats"]["cg_slab_preconditioned"] >= 2, out


def test_sharded_solve_single_rank_runs_the_same_trajectory():
    """Sharded-solve mode with world = 1: the operator keeps its full range outside the solves, the solves go through the slab code;
    a trajectory reproduces the plain one (same seeds) up to the solver tolerance."""
    from smoqyelph_b200 import api
    m = mdl.config("cfg1")
    x0 = m.random_fields(np.random.default_rng(3), smooth=True)
    out = []
    for sharded in (False, True):
        fdm = api.FermionDetMatrix(m, sym=True)
        elph = api.ElectronPhononParameters(m, fdm)
        pff = api.PFFCalculator(elph)
        elph.x = x0
        elph.update_fdm()
        if sharded:
            fdm.init_slab(0, 1)
            fdm.set_sharded_solve(True)
            lo, hi = fdm.slab["lo"], fdm.slab["hi"]
            assert (lo, hi) == (0, m.Ltau)
        hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=8, seed=11)
        acc, iters = hmc.hmc_update(tol_action=1e-10, tol_force=1e-8, maxiter=10000)
        out.append((acc, iters, elph.x.copy()))
    assert out[0][0] == out[1][0]
    assert abs(out[0][1] - out[1][1]) <= 1.0
    assert np.abs(out[0][2] - out[1][2]).max() < 1e-6 * max(1.0, np.abs(out[0][2]).max())


def test_two_gpu_sharded_chain():
    """One chain over 2 GPUs: every rank ends with bit-identical fields, and the chain follows the single-GPU one."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    for name, kpm in (("h16", ""), ("cfg1", ""), ("h16", "kpm"), ("cfg1", "kpm")):
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                              "--master-port", "29643", os.path.join(ROOT, "tools", "shard_worker.py"), name, "2", "0", kpm],
                             env=env, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
        out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
        assert out["ranks_bit_identical"], out
        assert out["accept_one_gpu"] == out["accept_sharded"], out
        assert out["max_abs_dx"] < 1e-3 * max(1.0, out["x_scale"]), out
        assert all(abs(a - b) <= 1.0 for a, b in zip(out["avg_iters_one_gpu"], out["avg_iters_sharded"])), out
        if kpm:
            assert out["stats_sharded"]["cg_slab_preconditioned"] > 0, out
        # measurement solves distributed over the ranks: identical G R everywhere, the single-GPU iteration count and density
        gr = out["greens"]
        assert gr["ranks_bit_identical"], out
        assert abs(gr["iters_one_gpu"] - gr["iters_sharded"]) <= 1.0, out
        assert abs(gr["n_one_gpu"] - gr["n_sharded"]) < 1e-6, out
