"""One Markov chain over several GPUs ("sharded solve", run under torchrun, one rank per GPU): every rank keeps the full state, the CG
solves are tau-slab partitioned.  Runs the same EFA-HMC trajectories (same seeds) first on one GPU per rank (replicated, no
communication), then sharded over all ranks, and reports on rank 0: trajectories/s of both, the largest difference between the two
chains, and whether all ranks hold bit-identical fields.  argv: config [trajectories] [warmup]"""
import ctypes as C, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import smoqyelph_b200  # noqa
from smoqyelph_b200 import api, lib, model as mdl
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
ntraj = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nwarm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
use_kpm = len(sys.argv) > 4 and sys.argv[4] == "kpm"
m = mdl.config(name)
L = lib.load()


def chain(sharded):
    """the same chain on every rank: identical start, identical seeds"""
    fdm = api.FermionDetMatrix(m, sym=True, device=local)
    elph = api.ElectronPhononParameters(m, fdm)
    pff = api.PFFCalculator(elph)
    elph.x = bench.cdw_start(m, 1000) if len(m.lattice_dims) == 2 and m.Nhol else m.random_fields(np.random.default_rng(1), smooth=True)
    elph.update_fdm()
    if sharded:
        fdm.init_sharded_solve(dist)
    hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=bench.NT, seed=77)
    P = api.KPMPreconditioner(fdm, update=False) if use_kpm else None
    log = []

    def trajectory():
        acc = C.c_int(0)
        info = np.zeros(8)
        lib.check(L.sq_hmc_update(hmc.h, P.h if P is not None else None, bench.TOL_ACTION, bench.TOL_FORCE, bench.MAXITER, None, 0, C.byref(acc), lib.ptr(info)))
        log.append((bool(acc.value), float(info[0])))
    for _ in range(nwarm):
        trajectory()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(ntraj):
        trajectory()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    x = elph.x.copy()
    # measurement solves: in the sharded chain the Nrv systems are distributed over the ranks (greens.cu)
    g = api.GreensEstimator(fdm, Nrv=5, seed=3)
    torch.cuda.synchronize(); dist.barrier()
    tg = time.perf_counter()
    git = g.update_greens_estimator(preconditioner=P, tol=1e-10)
    torch.cuda.synchronize()
    tg = torch.tensor([time.perf_counter() - tg], dtype=torch.float64)
    dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    GR = g.get()[1].copy()
    greens = {"iters": float(git), "seconds": float(tg.item()), "GR": GR, "n": complex(g.measure()["n"]).real}
    g.close()
    tuning = dict(fdm.tuning, stats=fdm.stats, greens=greens)
    hmc.close(); pff.close(); elph.close()
    if P is not None:
        P.close()
    fdm.close()
    return x, log, ntraj / float(dt.item()), tuning


x1, log1, rate1, tun1 = chain(False)
xs, logs, rates, tuns = chain(True)
digests = [None] * world
dist.all_gather_object(digests, hashlib.sha256(np.ascontiguousarray(xs).tobytes()).hexdigest())
g1, gs = tun1.pop("greens"), tuns.pop("greens")
gd = [None] * world
dist.all_gather_object(gd, hashlib.sha256(np.ascontiguousarray(gs["GR"]).tobytes()).hexdigest())
# (the two chains end in slightly different fields -- solver tolerance -- so G R is compared through the sharded chain's own operator:
#  rank-to-rank identity, the iteration count and the density it yields)
greens_out = {"ranks_bit_identical": len(set(gd)) == 1, "iters_one_gpu": g1["iters"], "iters_sharded": gs["iters"],
              "n_one_gpu": g1["n"], "n_sharded": gs["n"], "seconds_one_gpu": g1["seconds"], "seconds_sharded": gs["seconds"]}
if rank == 0:
    print(json.dumps({"config": name, "world": world, "trajectories": ntraj,
                      "trajectories_per_s_one_gpu": rate1, "trajectories_per_s_sharded": rates, "speedup": rates / rate1,
                      "max_abs_dx": float(np.abs(xs - x1).max()), "x_scale": float(np.abs(x1).max()),
                      "accept_one_gpu": [a for a, _ in log1], "accept_sharded": [a for a, _ in logs],
                      "avg_iters_one_gpu": [i for _, i in log1], "avg_iters_sharded": [i for _, i in logs],
                      "ranks_bit_identical": len(set(digests)) == 1, "kpm": use_kpm, "stats_sharded": tuns["stats"], "greens": greens_out}))
dist.destroy_process_group()
