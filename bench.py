#!/usr/bin/env python
"""bench.py -- EFA-HMC trajectories/s on the synthetic 32x32 Holstein square lattice (beta=20, dtau=0.05),
plus the fused M^T M v kernel against the HBM roofline and the CPU path timed beside it.

    python bench.py --gpus N --steps K --warmup W            (native arm; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the oracle port on host cores)

A "step" is one hmc_update! trajectory (Nt = 24 leapfrog steps = 24 force solves + 1 action solve,
src/EFAPFFHMCUpdater.jl:102-279) of one Markov chain.  `value` keeps the phonon field resident in HBM;
`e2e` pushes x host->device before and reads it back after every trajectory through the C ABI, as the Julia
drop-in does.  N > 1 runs one independent chain per GPU (the MPI tutorial's mode, weak scaling: every chain starts from the
same relaxed configuration and samples with seed + rank; the job time is the slowest chain's, `per_rank` lists every chain's time
and CG iteration count).  Extra keys of the N > 1 line, all under `tau_slab` (strong scaling of ONE chain, not part of `value`):
CG microseconds per iteration on one GPU, tau-slab partitioned with the host-launched NCCL loop and with the resident kernels that
exchange sums and boundary slices through peer-mapped mailboxes, and trajectories/s of one chain whose solves are sharded over all
GPUs (`chain_over_all_gpus`).  Clocks and throttle reasons are sampled during the timed region through NVML from a thread of this
process (`clocks.source`), `nvidia-smi -lms` being the fallback.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NT = 24
TOL_ACTION, TOL_FORCE, MAXITER = 1e-10, 1e-5, 10000        # tutorials/holstein_honeycomb.jl:64-68,591
ITERS_FILE = os.path.join(ROOT, "profiles", "bench_state.json")


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded along the way (NCCL prints its version banner to stdout
    when a communicator is created) must not add to it: point fd 1 at stderr for the run and keep the real stdout aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def cdw_start(m, seed):
    """Synthetic but physical start: the staggered (charge-density-wave) phonon order of the half-filled
    Holstein model at beta = 20, plus free-phonon thermal fluctuations.  Warm-up trajectories relax it."""
    from smoqyelph_b200 import model as mdl
    rng = np.random.default_rng(seed)
    Lx = m.lattice_dims[0]
    s = np.arange(m.N)
    stag = np.where(((s % Lx) + (s // Lx)) % 2 == 0, 1.0, -1.0)
    return np.asfortranarray(1.5 * stag[:, None] + 0.3 * mdl.thermal_fields(m, rng))


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread of this process (no second process
    touching the GPU while a latency-bound resident kernel runs), `nvidia-smi -lms` as the fallback."""
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading
        self.p, self.thread, self.samples, self.stop_flag = None, None, [], threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:                                 # noqa: BLE001
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons_fn(h)

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(reasons_fn(h))))
                    except Exception:                         # noqa: BLE001
                        pass
                    self.stop_flag.wait(0.1)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:                                     # noqa: BLE001
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                bits = 0
                for _, r in self.samples:
                    bits |= r
                out = {"sm_mhz": float(np.median([c for c, _ in self.samples])), "sm_max_mhz": self.max_mhz,
                       "reasons": sorted(nm for bit, nm in self.REASONS.items() if bits & bit), "samples": len(self.samples), "source": "nvml"}
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [q.strip() for q in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(m):
    """One fused M^T M v: read v, write v', read exp(-dtau V), cosh, sinh once each."""
    return (40 * m.N + 16 * m.Nh) * m.Ltau


# ------------------------------------------------------------------------------------------------------
# CPU arm / cpu_baseline: the oracle port, bounded sample, extrapolated with the trajectory's CG iteration count
# ------------------------------------------------------------------------------------------------------
def cpu_sample(m, x, precond, omp, budget_s, iters_per_traj):
    from oracle import oracle as orc
    rng = np.random.default_rng(7)
    if omp:                                                  # all host cores this process may use, whatever OMP_NUM_THREADS the launcher exported
        orc.lib(True).ref_set_num_threads(len(os.sched_getaffinity(0)))
    f = orc.RefFDM(m, sym=True, tol=TOL_FORCE, maxiter=MAXITER, omp=omp)
    e = orc.RefElPh(m, omp=omp)
    e.set_x(x)
    t0 = time.perf_counter()
    e.refresh(f)
    t_refresh = time.perf_counter() - t0
    pff = orc.RefPFF(e, f)
    R = (rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2)
    pff.sample(R)
    P = None
    if precond:
        P = orc.RefKPM(f)
        P.update(rng.standard_normal(m.N))
    # calibrate: 3 CG iterations, then size the sample to the budget
    b = np.asfortranarray(R)
    t0 = time.perf_counter()
    f.cg(b, P=P, tol=1e-300, maxiter=3)
    t3 = (time.perf_counter() - t0) / 3
    n_it = int(max(5, min(400, budget_s * 0.7 / t3)))
    t0 = time.perf_counter()
    f.cg(b, P=P, tol=1e-300, maxiter=n_it)
    t_iter = (time.perf_counter() - t0) / n_it
    # force evaluation with the solve capped at 2 iterations: the non-CG tail (Lambda ops, M, M^T, dM/dx, dLambda/dx)
    t0 = time.perf_counter()
    pff.force(P=P, lanczos_start=rng.standard_normal(m.N) if P is not None else None, tol=1e-300, maxiter=2)
    t_tail = max(0.0, time.perf_counter() - t0 - 2 * t_iter)
    t_traj = iters_per_traj * t_iter + NT * t_tail + (NT + 2) * t_refresh
    threads = f.L.ref_num_threads()
    sample = (f"{n_it} CG iterations of M^T M (precond={'KPM' if precond else 'I'}) + 1 force tail + 1 operator refresh of the C oracle "
              f"port on the bench phonon state; t_iter={t_iter * 1e3:.2f} ms, t_tail={t_tail * 1e3:.1f} ms, t_refresh={t_refresh * 1e3:.1f} ms; "
              f"extrapolated to one trajectory = {iters_per_traj:.0f} CG iterations (count from the GPU arm, parity-tested +-1) "
              f"+ {NT} tails + {NT + 2} refreshes")
    return 1.0 / t_traj, threads, sample, {"t_iter_ms": t_iter * 1e3, "t_tail_ms": t_tail * 1e3, "t_refresh_ms": t_refresh * 1e3, "n_it": n_it}


def stored_iters(precond):
    try:
        d = json.load(open(ITERS_FILE))
        return float(d["precond_on" if precond else "precond_off"]["cg_iters_per_trajectory"])
    except Exception:
        return 20000.0 if not precond else 4000.0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from smoqyelph_b200 import model as mdl
    m = mdl.config(args.config)
    precond = args.precond == "on"
    x = cdw_start(m, 1000)
    iters = stored_iters(precond)
    vals, info, threads, sample = [], None, 1, ""
    budget = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    for step in range(args.warmup + args.steps):
        v, threads, sample, info = cpu_sample(m, x, precond, True, budget, iters)
        if step >= args.warmup:
            vals.append(v)
    val = float(np.mean(vals))
    line = {"impl": "reference", "metric": "efa_hmc_trajectories_per_s", "value": val, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(m, precond, "cpu"),
            "cpu_baseline": {"value": val, "unit": "trajectories/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "detail": info}
    emit(line)


def workload_config(m, precond, where):
    return {"workload": f"{m.name}: Holstein square {m.lattice_dims[0]}x{m.lattice_dims[1]}, beta={m.beta:g}, dtau={m.dtau:g} "
                        f"(N={m.N}, Ltau={m.Ltau}, N*Ltau={m.N * m.Ltau}), Omega=1, alpha=1.5, mu=0, ph_sym_form; "
                        f"EFA-PFF-HMC Nt={NT}, tol_action={TOL_ACTION:g}, tol_force={TOL_FORCE:g}, SymFermionDetMatrix",
            "preconditioner": "KPM (defaults)" if precond else "I",
            "phonon_state": "staggered CDW order 1.5 + 0.3 x free-phonon thermal noise, relaxed by the warm-up trajectories",
            "parallelism": "independent chains, one per GPU" if where == "gpu" else "OpenMP over tau inside each sweep",
            "l2": "trajectory working set is L2-resident by nature (vector 6.5 MB); roofline kernel timed with an L2 flush "
                  "(256 MB write) between launches"}


# ------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import ctypes as C
    from smoqyelph_b200 import api, lib, model as mdl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    m = mdl.config(args.config)
    precond = args.precond == "on"
    fdm = api.SymFermionDetMatrix(m, tol=TOL_ACTION, maxiter=MAXITER, device=local)
    elph = api.ElectronPhononParameters(m, fdm)
    pff = api.PFFCalculator(elph)
    P = api.KPMPreconditioner(fdm, update=False) if precond else None
    elph.x = cdw_start(m, 1000)                              # every chain starts from the same relaxed configuration ...
    elph.update_fdm()
    stream = torch.cuda.ExternalStream(fdm.stream, device=dev)
    L = lib.load()

    def trajectory(h):
        acc = C.c_int(0)
        info = np.zeros(8)
        lib.check(L.sq_hmc_update(h.h, P.h if P is not None else None, TOL_ACTION, TOL_FORCE, MAXITER, None, 0, C.byref(acc), lib.ptr(info)))
        return bool(acc.value), info

    # ---- warm-up (untimed): relaxes the synthetic start, warms the caches and the clocks
    hmc = api.EFAPFFHMCUpdater(elph, pff, Nt=NT, seed=77)
    for _ in range(args.warmup):
        trajectory(hmc)
    x_w = elph.x                                             # state every timed leg starts from

    def fresh_updater():
        elph.x = x_w
        elph.update_fdm()
        return api.EFAPFFHMCUpdater(elph, pff, Nt=NT, seed=4242 + rank)      # ... and samples with its own seed (seed + rank)

    # ---- value: K trajectories, x resident in HBM, timed with CUDA events on the library stream
    h = fresh_updater()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = fdm.launch_count
    e0.record(stream)
    iters_total, accepted = 0.0, 0
    for _ in range(args.steps):
        acc, info = trajectory(h)
        iters_total += info[0] * (NT + 1)
        accepted += int(acc)
    e1.record(stream)
    e1.synchronize()
    launches = fdm.launch_count - l0
    ms_dev = e0.elapsed_time(e1)
    t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_value = float(t.item())
    value = world * args.steps / (ms_value * 1e-3)

    # ---- e2e: same trajectories, x crosses PCIe both ways every step (pinned host buffer), wall clock
    h = fresh_updater()
    nx = m.Nph * m.Ltau
    pin = torch.empty(nx, dtype=torch.float64).pin_memory()
    pin.copy_(torch.from_numpy(np.ascontiguousarray(x_w.ravel(order="F"))))
    pptr = C.c_void_p(pin.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lib.check(L.sq_elph_set_x(elph.h, pptr))             # H2D
        lib.check(L.sq_elph_refresh_fdm(elph.h))
        trajectory(h)
        lib.check(L.sq_elph_get_x(elph.h, pptr))             # D2H (blocking)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    e2e_value = world * args.steps / float(t.item())

    roofline = None
    iters_per_traj = iters_total / args.steps
    per_rank = None
    if dist is not None:
        # the chains differ (seed + rank), so do their CG iteration counts: the job time is the slowest chain's
        pr = torch.tensor([ms_dev / args.steps, iters_per_traj], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(pr) for _ in range(world)]
        dist.all_gather(allr, pr)
        per_rank = {"ms_per_step": [round(float(q[0]), 2) for q in allr], "cg_iters_per_trajectory": [round(float(q[1]), 1) for q in allr]}
    matvecs_per_traj = iters_per_traj + 2 * (NT + 1)
    if rank == 0:
        # ---- roofline, rank 0.  Two kernels matter: the fused M^T M v kernel on its own (what north_star names), timed by
        #      the library with CUDA events on its stream (back to back = L2-hot as inside a solve; and L2-cold with a 256 MB
        #      write between launches), and the kernel the trajectory actually spends its time in -- with the register path
        #      the whole CG solve is ONE resident launch that applies M^T M once per iteration.
        n = m.N * m.Ltau
        d_in = torch.randn(n, 2, dtype=torch.float64, device=dev)
        d_out = torch.zeros_like(d_in)
        flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)          # 256 MB > 126 MB L2
        B = algorithmic_bytes(m)
        tuning = fdm.tuning
        path = tuning["path"]
        op = 102 if path == 3 else api.OP_MTM                # 102: the register-path kernel on native-order vectors (as in CG)
        t_hot = fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 400) * 1e-6
        t_cold = fdm.time_mul(op, d_out.data_ptr(), d_in.data_ptr(), 60, flush.data_ptr(), flush.numel() * 4) * 1e-6
        # CG iteration time: fixed iteration count, device-resident vectors (one launch per solve on the register path)
        nit = 2000
        fdm.cg_dev(d_out.data_ptr(), d_in.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=200)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fdm.cg_dev(d_out.data_ptr(), d_in.data_ptr(), True, preconditioner=P, tol=1e-300, maxiter=nit)
        torch.cuda.synchronize()
        t_iter = (time.perf_counter() - t0) / nit
        peak, peak_src = measured_peak()
        state = {}
        try:
            state = json.load(open(ITERS_FILE))
        except Exception:
            pass
        # Models without SSH coupling have tau-independent hoppings: the fast kernels read (cosh, sinh) once per kernel
        # instead of once per slice, so the bytes they must move are 40 N Ltau + 16 Nh, not the generic figure.
        uniform = (m.Nssh == 0) and path in (2, 3)
        Bk = (40 * m.N * m.Ltau + 16 * m.Nh) if uniform else B
        names = {0: "k_fdm_fused<2>", 2: "k_fdm_fused_v2<2>", 3: "k_fdm_v3<2,0,1> (register path, native order)"}
        resident = (path == 3) and not precond
        step_s = ms_dev * 1e-3 / args.steps
        matvec = {"kernel": names.get(path, "global passes") + " (fused M^T M v)", "us_per_launch_cold_l2": t_cold * 1e6,
                  "us_per_launch_hot_l2": t_hot * 1e6, "achieved_cold_l2": Bk / t_cold / 1e9, "frac_cold_l2": Bk / t_cold / 1e9 / peak,
                  "achieved_hot_l2": Bk / t_hot / 1e9, "frac_hot_l2": Bk / t_hot / 1e9 / peak, "matvecs_per_s_hot_l2": 1.0 / t_hot,
                  "traffic": state.get("mtm_dram_bytes_per_launch"),
                  "note": "back-to-back launches on this system are paced in ~2 us steps (profiles/README.md); a single-wave kernel of "
                          "~6 us cannot be timed below that from the host, which is why the solver is one resident launch"}
        if resident:
            roofline = {"bound": "hbm", "kernel": "k_cg_v3_resident1 (whole CG solve in one launch; M^T M v once per iteration, x r p resident on chip)",
                        "achieved": Bk / t_iter / 1e9, "peak": peak, "unit": "GB/s", "frac": Bk / t_iter / 1e9 / peak,
                        "traffic": state.get("cg_resident_dram_bytes_per_iteration"), "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": Bk * nit, "algorithmic_bytes_per_unit": Bk, "unit_of_work": "one CG iteration = one fused M^T M v",
                        "units_per_launch": nit, "us_per_unit": t_iter * 1e6,
                        "limiter": "latency: one grid-wide sum per iteration (>= 2 us on this two-die part, tools/ubench/grid_sum.cu) + "
                                   "FP64 issue of one wave; HBM traffic per iteration is ~0.1 MB because the working set stays on chip"}
        else:
            roofline = {"bound": "hbm", "kernel": matvec["kernel"], "achieved": Bk / t_cold / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": Bk / t_cold / 1e9 / peak, "traffic": matvec["traffic"], "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": Bk, "limiter": "shared-memory crossbar + barrier latency (DESIGN.md section 4), not HBM"}
        roofline.update({"generic_formula_bytes_per_unit": B,
                         "bytes_note": "tau-uniform hoppings: (cosh, sinh) read once per kernel" if uniform else "generic (40 N + 16 Nh) Ltau",
                         "cg_us_per_iteration": t_iter * 1e6, "cg_iterations_per_s": 1.0 / t_iter,
                         "share_of_step": iters_per_traj * t_iter / step_s, "matvec_kernel": matvec, "tuning": tuning})

    barrier()
    # ---- tau-slab strong scaling of the CG solve (N > 1): the same M^T M system partitioned over the ranks with
    #      NCCL halo exchange + all-reduced dot products, against the single-GPU solve timed on every rank first
    def finish(tau_slab, cpu=None):
        line = {"metric": "efa_hmc_trajectories_per_s", "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(m, precond, "gpu"),
                "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": nx * 8, "d2h_bytes_per_step": nx * 8 + 64},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "cg_iters_per_trajectory": iters_per_traj, "matvecs_per_trajectory": matvecs_per_traj,
                "acceptance": accepted / args.steps, "per_rank": per_rank, "tau_slab": tau_slab}
        emit(line)

    tau_slab = None
    if world > 1:
        # the strong-scaling section is an extra: a watchdog makes sure the headline line is printed even if it stalls
        import threading
        stage = ["start"]

        def bail():
            if rank == 0:
                finish({"error": "tau-slab section exceeded its time limit at stage '%s'" % stage[0]})
            os._exit(0)
        dog = threading.Timer(240.0, bail)
        dog.daemon = True
        dog.start()
        try:
            n = m.N * m.Ltau
            d_b = torch.randn(n, 2, dtype=torch.float64, device=dev)
            d_x = torch.zeros_like(d_b)
            nit = 400

            def timed_cg():
                """us per iteration (max over ranks), or None if any rank failed -- every rank always takes part in the collectives"""
                bad = 0.0
                try:
                    fdm.cg_dev(d_x.data_ptr(), d_b.data_ptr(), True, tol=1e-300, maxiter=40)
                except Exception as e:                            # noqa: BLE001
                    bad = 1.0
                    sys.stderr.write("rank %d: tau-slab CG failed at stage %s: %s\n" % (rank, stage[0], e))
                barrier()
                t0 = time.perf_counter()
                if not bad:
                    try:
                        fdm.cg_dev(d_x.data_ptr(), d_b.data_ptr(), True, tol=1e-300, maxiter=nit)
                        torch.cuda.synchronize()
                    except Exception as e:                        # noqa: BLE001
                        bad = 1.0
                        sys.stderr.write("rank %d: tau-slab CG failed at stage %s: %s\n" % (rank, stage[0], e))
                dt = time.perf_counter() - t0
                tt = torch.tensor([dt, bad], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return None if tt[1].item() > 0 else float(tt[0].item()) / nit * 1e6

            stage[0] = "one GPU"
            us1 = timed_cg()
            ids = [api.FermionDetMatrix.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            fdm.init_slab(rank, world, ids[0])
            stage[0] = "NCCL loop"
            os.environ["SQ_NO_RESIDENT_CG"] = "1"                 # (a) host-launched NCCL loop
            us_nccl = timed_cg()
            del os.environ["SQ_NO_RESIDENT_CG"]
            stage[0] = "mailboxes"
            handles = [None] * world                              # (b) resident kernels + peer-mapped mailboxes (CUDA IPC over NVLink)
            dist.all_gather_object(handles, fdm.mailbox_handle())
            fdm.mailbox_open(handles)
            stage[0] = "resident"
            usN = timed_cg() if us_nccl is not None else None
            best = min([u for u in (usN, us_nccl) if u is not None], default=None)
            tau_slab = {"what": "unpreconditioned CG iterations on M^T M, cfg4, tau-slab partitioned (strong scaling)",
                        "cg_us_per_iter_1gpu": us1, "cg_us_per_iter": usN, "cg_us_per_iter_nccl_loop": us_nccl, "n_gpus": world,
                        "cg_iters_per_s": 1e6 / best if best else None, "speedup_vs_1gpu": us1 / best if best and us1 else None,
                        "comm": "resident kernel per rank; grid-wide sums and boundary slices as device-initiated stores into peer-mapped "
                                "mailboxes (CUDA IPC over NVLink); cg_us_per_iter_nccl_loop = host-launched NCCL send/recv + all-reduces"}
            # (c) whole trajectories of ONE chain over all GPUs: full state on every rank, CG solves partitioned ("sharded solve")
            stage[0] = "sharded chain"
            if usN is not None:
                try:
                    fdm.set_sharded_solve(True)
                    xs = [x_w if rank == 0 else None]
                    dist.broadcast_object_list(xs, src=0)
                    elph.x = xs[0]
                    elph.update_fdm()
                    hs = api.EFAPFFHMCUpdater(elph, pff, Nt=NT, seed=4242)
                    trajectory(hs)                                # untimed: tunes the kernels for the slab range
                    torch.cuda.synchronize()
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(args.steps):
                        trajectory(hs)
                    torch.cuda.synchronize()
                    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    rate = args.steps / float(tt.item())
                    tau_slab["chain_over_all_gpus"] = {"trajectories_per_s": rate, "trajectories_per_s_one_gpu": value / world,
                                                       "speedup_vs_1gpu": rate / (value / world),
                                                       "note": "cfg4 is half a wave of work per GPU at N = 2: one GPU per chain is the faster "
                                                               "choice at this size; Ltau = 800 gives 2.0x on 2 GPUs (profiles/README.md)"}
                except Exception as e:                            # noqa: BLE001
                    tau_slab["chain_over_all_gpus"] = {"error": str(e)}
        except Exception as e:                                # noqa: BLE001  (a rank that fails here must still print / exit cleanly)
            sys.stderr.write("rank %d: tau-slab section failed at stage %s: %s\n" % (rank, stage[0], e))
            tau_slab = {"error": "stage '%s': %s" % (stage[0], e)}
        dog.cancel()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu:
        v, cores, sample, detail = cpu_sample(m, x_w, precond, False, args.cpu_budget, iters_per_traj)
        cpu = {"value": v, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": sample, "detail": detail}

    finish(tau_slab, cpu)
    if dist is not None:
        dist.destroy_process_group()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--precond", default=os.environ.get("SQ_BENCH_PRECOND", "off"), choices=["on", "off"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    import smoqyelph_b200  # noqa: F401
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
