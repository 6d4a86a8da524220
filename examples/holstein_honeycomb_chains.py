"""The reference's MPI tutorial (/root/reference/tutorials/holstein_honeycomb_mpi.jl: one Markov chain per process, seed + pID, statistics
merged at the end) on the B200 library: one process per GPU under torchrun, no communication during sampling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 examples/holstein_honeycomb_chains.py [L] [beta] [N_therm] [N_measurements]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import torch
import torch.distributed as dist

import smoqyelph_b200  # noqa: F401
from smoqyelph_b200 import parallel
from holstein_honeycomb import run_simulation

if __name__ == "__main__":
    a = sys.argv[1:]
    L, beta = (int(a[0]) if a else 3), (float(a[1]) if len(a) > 1 else 4.0)
    nt, nm = (int(a[2]) if len(a) > 2 else 20), (int(a[3]) if len(a) > 3 else 20)
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    if "RANK" in os.environ:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    m, obs, meta = run_simulation(L, beta, nt, nm, seed=parallel.chain_seed(0, rank), device=local, use_preconditioner=True)
    names = ["density", "double_occ", "hmc_acceptance_rate", "hmc_iters", "runtime_s"]
    mean, err = parallel.merge_chain_statistics([obs["density"], obs["double_occ"], meta["hmc_acceptance_rate"], meta["hmc_iters"],
                                                 meta["runtime_s"]], dist if dist.is_initialized() else None)
    if rank == 0:
        print(json.dumps({"model": m.name, "chains": dist.get_world_size() if dist.is_initialized() else 1,
                          **{n: {"mean": float(mu), "stderr": float(e)} for n, mu, e in zip(names, mean, err)}}))
    if dist.is_initialized():
        dist.destroy_process_group()
