"""Multi-GPU tau-slab worker (run under torchrun, one rank per GPU): parity of the slab matvec / CG against the CPU
oracle, and CG iterations/s.  Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
import smoqyelph_b200  # noqa
from smoqyelph_b200 import api, model as mdl
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
check = (len(sys.argv) > 2 and sys.argv[2] == "check")
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 500
m = mdl.config(name)
rng = np.random.default_rng(0)                       # same fields on every rank
x = bench.cdw_start(m, 0) if m.Nhol and len(m.lattice_dims) == 2 else m.random_fields(rng, smooth=True)
fdm = api.FermionDetMatrix(m, sym=True, device=local)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = x
elph.update_fdm()
ids = [api.FermionDetMatrix.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
fdm.init_slab(rank, world, ids[0])
if os.environ.get("SQ_SLAB_MAILBOX", "1") != "0":           # peer-mapped mailboxes: the resident multi-GPU CG where it applies
    handles = [None] * world
    dist.all_gather_object(handles, fdm.mailbox_handle())
    fdm.mailbox_open(handles)
lo, hi = fdm.slab["lo"], fdm.slab["hi"]
out = {"config": name, "world": world}
b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))

def gather(v):
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, np.ascontiguousarray(v[lo:hi])))
    full = np.zeros_like(v)
    for a, c, d in parts:
        full[a:c] = d
    return full

if check:
    from oracle import oracle as orc
    import dense_ref as dr
    V, t = dr.build_Vt(m, x)
    ref = orc.RefFDM(m, sym=True); ref.update(V, t)
    rel = lambda a, c: float(np.linalg.norm(a - c) / np.linalg.norm(c))
    n = m.N * m.Ltau
    # device-resident matvec with halo exchange: every rank only holds valid data in its own slab of the input
    st = torch.cuda.ExternalStream(fdm.stream)
    bl = np.zeros_like(b); bl[lo:hi] = b[lo:hi]
    d_in = torch.from_numpy(np.ascontiguousarray(bl).view(np.float64).reshape(m.Ltau, m.N, 2)).cuda()     # [l][i] layout
    d_out = torch.zeros_like(d_in)
    for op, nm in ((2, "mul_MtM"), (0, "mul_M"), (1, "mul_Mt")):
        fdm.mul_dev(op, d_out.data_ptr(), d_in.data_ptr())
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.complex128).reshape(m.Ltau, m.N)
        out["err_" + nm] = rel(gather(got), getattr(ref, nm)(b))
    xr, itr, _ = ref.cg(b, tol=1e-13, maxiter=50000)
    xg, itg, epsg = fdm.ldiv(b, tol=1e-13, maxiter=50000)
    out["err_cg"] = rel(gather(xg), xr)
    _, itr5, _ = ref.cg(b, tol=1e-6, maxiter=50000)
    _, itg5, _ = fdm.ldiv(b, tol=1e-6, maxiter=50000)
    out["iters"] = [int(itr5), int(itg5)]
    # preconditioned solve: P^-1 sharded by Matsubara frequency, two all-to-all exchanges per apply (slab.cu, kpm_ldiv_slab)
    Pr = orc.RefKPM(ref)
    Pr.update(np.random.default_rng(5).standard_normal(m.N))
    if Pr.active:
        Pg = api.KPMPreconditioner(fdm, update=False)
        Pg.set_bounds(*Pr.bounds)
        xr, itr, _ = ref.cg(b, P=Pr, tol=1e-13, maxiter=5000)
        xg, itg, epsg = fdm.ldiv(b, preconditioner=Pg, tol=1e-13, maxiter=5000, refresh=False)
        out["err_cg_kpm"] = rel(gather(xg), xr)
        _, itr5, _ = ref.cg(b, P=Pr, tol=1e-6, maxiter=5000)
        _, itg5, _ = fdm.ldiv(b, preconditioner=Pg, tol=1e-6, maxiter=5000, refresh=False)
        out["iters_kpm"] = [int(itr5), int(itg5)]
        out["stats"] = fdm.stats
# throughput: fixed number of CG iterations on device-resident vectors
n = m.N * m.Ltau
d_b = torch.randn(n, 2, dtype=torch.float64, device="cuda"); d_x = torch.zeros_like(d_b)
fdm.cg_dev(d_x.data_ptr(), d_b.data_ptr(), True, tol=1e-300, maxiter=50)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
fdm.cg_dev(d_x.data_ptr(), d_b.data_ptr(), True, tol=1e-300, maxiter=niter)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
out["cg_us_per_iter"] = float(dt.item()) / niter * 1e6
# preconditioned iterations (frequency-sharded apply)
Pt = api.KPMPreconditioner(fdm, update=False)
act, _ = Pt.update(np.random.default_rng(6).standard_normal(m.N))
if act:
    nk = 20
    fdm.cg_dev(d_x.data_ptr(), d_b.data_ptr(), True, preconditioner=Pt, tol=1e-300, maxiter=4)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    fdm.cg_dev(d_x.data_ptr(), d_b.data_ptr(), True, preconditioner=Pt, tol=1e-300, maxiter=nk)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out["cg_kpm_us_per_iter"] = float(dt.item()) / nk * 1e6
out["resident"] = os.environ.get("SQ_SLAB_MAILBOX", "1") != "0" and not os.environ.get("SQ_NO_RESIDENT_CG")
out["slab"] = [lo, hi]
out["tuning"] = fdm.tuning
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
