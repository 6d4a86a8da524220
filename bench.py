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


STATE_FILE = os.path.join(ROOT, "bench_data", "bench_state_cfg4_f32.npy")


def cdw_start(m, seed):
    """Synthetic but physical start: the staggered (charge-density-wave) phonon order of the half-filled
    Holstein model at beta = 20, plus free-phonon thermal fluctuations."""
    from smoqyelph_b200 import model as mdl
    rng = np.random.default_rng(seed)
    Lx = m.lattice_dims[0]
    s = np.arange(m.N)
    stag = np.where(((s % Lx) + (s // Lx)) % 2 == 0, 1.0, -1.0)
    return np.asfortranarray(1.5 * stag[:, None] + 0.3 * mdl.thermal_fields(m, rng))


def bench_state(m):
    """The phonon field BOTH arms time: cdw_start relaxed by 20 EFA-PFF-HMC trajectories (tools/make_bench_state.py, run once on a
    B200 and committed as float32), so that the CPU arm -- which cannot afford warm-up trajectories -- sees the same thermalised,
    tau-rough field as the GPU arm.  (On the raw CDW start the field is tau-smooth and a preconditioned solve takes 8 iterations
    instead of ~100: not a representative workload.)  Falls back to the raw start for other configurations."""
    if m.name == "cfg4" and os.path.exists(STATE_FILE):
        x = np.load(STATE_FILE).astype(np.float64)
        if x.shape == (m.Nph, m.Ltau):
            return np.asfortranarray(x), "cdw_start(seed 1000) relaxed by 20 EFA-PFF-HMC trajectories on a B200 (bench_data/bench_state_cfg4_f32.npy, tools/make_bench_state.py)"
    return cdw_start(m, 1000), "staggered CDW order 1.5 + 0.3 x free-phonon thermal noise (seed 1000), NOT relaxed"


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
# CPU arm / cpu_baseline: the oracle port on the box's host cores.  One sample = the pieces of a trajectory timed FOR REAL on the
# bench phonon state: one complete force evaluation (CG to tol_force -- its iteration count is measured, not assumed -- plus the
# Lambda / M / M^T / dM/dx / dLambda/dx tail), one operator refresh, and (first sample only) one complete action evaluation (CG to
# tol_action).  A trajectory is Nt force evaluations + 1 action evaluation + Nt + 2 refreshes, so
#     t_trajectory = Nt t_force + t_action + (Nt + 2) t_refresh.
# The only extrapolation left is "the Nt force solves of a trajectory cost Nt times the first one"; a complete CPU trajectory at cfg4
# takes about a minute per step, which does not fit a run of a few minutes.
# ------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, m, x, precond, omp):
        from oracle import oracle as orc
        self.orc, self.m, self.precond = orc, m, precond
        if omp:                                              # all host cores this process may use, whatever OMP_NUM_THREADS the launcher exported
            orc.lib(True).ref_set_num_threads(len(os.sched_getaffinity(0)))
        rng = np.random.default_rng(7)
        self.rng = rng
        self.f = orc.RefFDM(m, sym=True, tol=TOL_FORCE, maxiter=MAXITER, omp=omp)
        self.e = orc.RefElPh(m, omp=omp)
        self.e.set_x(x)
        t0 = time.perf_counter()
        self.e.refresh(self.f)
        self.t_refresh = time.perf_counter() - t0
        self.pff = orc.RefPFF(self.e, self.f)
        R = (rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N))) / np.sqrt(2)
        self.pff.sample(R)
        self.P = None
        if precond:
            self.P = orc.RefKPM(self.f)
            self.P.update(rng.standard_normal(m.N))
        self.threads = self.f.L.ref_num_threads()
        self.t_action = self.it_action = None
        self.real_seconds = 0.0

    def _start(self):
        return self.rng.standard_normal(self.m.N) if self.P is not None else None

    def sample(self, quick=False):
        """-> trajectories/s from one real force evaluation (quick: capped at 3 CG iterations, for warm-up steps)."""
        t0 = time.perf_counter()
        _, _, it_f, _ = self.pff.force(P=self.P, lanczos_start=self._start(), tol=TOL_FORCE, maxiter=3 if quick else MAXITER)
        t_force = time.perf_counter() - t0
        self.real_seconds += t_force
        if quick:
            return None
        if self.t_action is None:
            t0 = time.perf_counter()
            _, it_a, _ = self.pff.action(P=self.P, lanczos_start=self._start(), tol=TOL_ACTION, maxiter=MAXITER)
            self.t_action, self.it_action = time.perf_counter() - t0, it_a
            self.real_seconds += self.t_action
        t_traj = NT * t_force + self.t_action + (NT + 2) * self.t_refresh
        self.last = {"t_force_s": t_force, "cg_iters_force": int(it_f), "t_action_s": self.t_action, "cg_iters_action": int(self.it_action),
                     "t_refresh_s": self.t_refresh}
        return 1.0 / t_traj

    def describe(self, nsamples):
        d = self.last
        return (f"{nsamples} x [one complete force evaluation: CG on M^T M to {TOL_FORCE:g} (precond={'KPM' if self.precond else 'I'}, "
                f"{d['cg_iters_force']} iterations measured) + Lambda/M/M^T/dM/dx tail, {d['t_force_s']:.2f} s] + one complete action evaluation "
                f"(CG to {TOL_ACTION:g}, {d['cg_iters_action']} iterations, {d['t_action_s']:.2f} s) + operator refresh ({d['t_refresh_s'] * 1e3:.0f} ms) of the C "
                f"oracle port on the bench phonon state, all timed for real ({self.real_seconds:.1f} s of CPU work in this run); "
                f"trajectory = {NT} x force + action + {NT + 2} x refresh")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from smoqyelph_b200 import model as mdl
    m = mdl.config(args.config)
    precond = args.precond != "off"                          # stock configuration of every shipped driver: a KPMPreconditioner (tutorials/holstein_honeycomb.jl:507)
    x0, state_label = bench_state(m)
    arm = CpuArm(m, x0, precond, True)
    t_wall = time.perf_counter()
    vals = []
    for step in range(args.warmup + args.steps):
        if step < args.warmup:
            arm.sample(quick=True)
            continue
        if vals and time.perf_counter() - t_wall > 150.0:    # keep the whole run within a few minutes: reuse the samples taken so far
            break
        vals.append(arm.sample())
    val = float(np.mean(vals))
    line = {"impl": "reference", "metric": "efa_hmc_trajectories_per_s", "value": val, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(m, "KPM (defaults)" if precond else "I", "cpu", state_label),
            "cpu_baseline": {"value": val, "unit": "trajectories/s", "cores": arm.threads, "kind": "port", "sample": arm.describe(len(vals))},
            "e2e": {"value": val, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "real_timed_seconds": arm.real_seconds, "samples": len(vals),
            "note": "ONE Markov chain on the host cores at every --gpus N (the reference has no intra-chain parallelism beyond threads); "
                    "ms_per_step is the modelled time of a whole trajectory, real_timed_seconds what this run actually spent in its samples",
            "detail": arm.last}
    emit(line)


def workload_config(m, precond_label, where, state_label, extra=None):
    cfg = {"workload": f"{m.name}: Holstein square {m.lattice_dims[0]}x{m.lattice_dims[1]}, beta={m.beta:g}, dtau={m.dtau:g} "
                       f"(N={m.N}, Ltau={m.Ltau}, N*Ltau={m.N * m.Ltau}), Omega=1, alpha=1.5, mu=0, ph_sym_form; "
                       f"EFA-PFF-HMC Nt={NT}, tol_action={TOL_ACTION:g}, tol_force={TOL_FORCE:g}, SymFermionDetMatrix",
           "preconditioner": precond_label,
           "phonon_state": state_label + "; both arms start their timed region from this very field (the native arm's warm-up trajectories "
                           "are discarded)",
           "parallelism": {"gpu": "one Markov chain on one GPU", "gpu_strong": "ONE Markov chain, CG solves tau-slab partitioned over all GPUs "
                           "(strong scaling; full state replicated, halos and dot products through peer-mapped mailboxes)",
                           "cpu": "OpenMP over tau inside each sweep"}[where],
           "l2": "trajectory working set is L2-resident by nature (vector 6.5 MB); roofline kernel timed with an L2 flush "
                 "(256 MB write) between launches"}
    if extra:
        cfg.update(extra)
    return cfg


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

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m = mdl.config(args.config)
    fdm = api.SymFermionDetMatrix(m, tol=TOL_ACTION, maxiter=MAXITER, device=local)
    elph = api.ElectronPhononParameters(m, fdm)
    pff = api.PFFCalculator(elph)
    P_kpm = api.KPMPreconditioner(fdm, update=False)
    x0, state_label = bench_state(m)                         # the state both arms time (see config.phonon_state)
    elph.x = x0
    elph.update_fdm()
    stream = torch.cuda.ExternalStream(fdm.stream, device=dev)
    L = lib.load()
    nx = m.Nph * m.Ltau

    def trajectory(h, P):
        acc = C.c_int(0)
        info = np.zeros(8)
        lib.check(L.sq_hmc_update(h.h, P.h if P is not None else None, TOL_ACTION, TOL_FORCE, MAXITER, None, 0, C.byref(acc), lib.ptr(info)))
        return bool(acc.value), info

    def fresh_updater(seed):
        elph.x = x0
        elph.update_fdm()
        return api.EFAPFFHMCUpdater(elph, pff, Nt=NT, seed=seed)

    def timed_chain(P, steps, seed):
        """`steps` trajectories from x0, x resident in HBM, CUDA events on the library stream -> (ms, CG iterations, accepted, launches)."""
        h = fresh_updater(seed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = fdm.launch_count
        e0.record(stream)
        iters, accepted = 0.0, 0
        for _ in range(steps):
            acc, info = trajectory(h, P)
            iters += info[0] * (NT + 1)
            accepted += int(acc)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), iters, accepted, fdm.launch_count - l0

    # ---- warm-up (untimed, discarded): warms caches, clocks and the autotuner for both solver configurations
    hw = fresh_updater(77)
    for k in range(args.warmup):
        trajectory(hw, P_kpm if (k == args.warmup - 1 and args.precond != "off") else None)

    # ---- which preconditioner?  The reference's drivers pass a KPMPreconditioner; the library accepts either, exactly as the reference
    #      (`preconditioner = I` or a KPMPreconditioner).  `auto` times one trajectory of each from the bench state and keeps the faster.
    choice = {}
    if args.precond == "auto":
        for label, P in (("I", None), ("KPM (defaults)", P_kpm)):
            ms, it, _, _ = timed_chain(P, 1, 4242)
            choice[label] = {"trajectories_per_s": 1e3 / max_over_ranks(ms), "cg_iters_per_trajectory": it}
        use_kpm = choice["KPM (defaults)"]["trajectories_per_s"] > choice["I"]["trajectories_per_s"]
    else:
        use_kpm = args.precond == "on"
    P = P_kpm if use_kpm else None
    plabel = "KPM (defaults)" if use_kpm else "I"
    cfg_extra = {"preconditioner_choice": {"mode": args.precond, "timed_one_trajectory_each": choice or None,
                                           "reference_arm_uses": "KPM (defaults), the stock configuration of the shipped drivers"}}

    pin = torch.empty(nx, dtype=torch.float64).pin_memory()
    pin.copy_(torch.from_numpy(np.ascontiguousarray(x0.ravel(order="F"))))
    pptr = C.c_void_p(pin.data_ptr())

    def e2e_chain(P, steps, seed):
        """same trajectories through the C ABI with HOST buffers: x crosses PCIe both ways every step (pinned), wall clock"""
        h = fresh_updater(seed)
        pin.copy_(torch.from_numpy(np.ascontiguousarray(x0.ravel(order="F"))))
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            lib.check(L.sq_elph_set_x(elph.h, pptr))             # H2D
            lib.check(L.sq_elph_refresh_fdm(elph.h))
            trajectory(h, P)
            lib.check(L.sq_elph_get_x(elph.h, pptr))             # D2H (blocking)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    sampler = None
    per_rank = chains = tau_slab = None
    if world == 1:
        # ---- value: K trajectories of one chain on one GPU
        sampler = ClockSampler(local)
        ms_dev, iters_total, accepted, launches = timed_chain(P, args.steps, 4242)
        clocks = sampler.stop()
        ms_value = ms_dev
        value = args.steps / (ms_value * 1e-3)
        e2e_value = args.steps / e2e_chain(P, args.steps, 4242)
        scaling, where = "weak", "gpu"
    else:
        # ---- extra: independent chains, one per GPU (the reference's MPI mode; weak scaling) -- NOT the headline
        ms_c, it_c, _, _ = timed_chain(P, args.steps, 4242 + rank)
        pr = torch.tensor([ms_c / args.steps, it_c / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(pr) for _ in range(world)]
        dist.all_gather(allr, pr)
        chains = {"what": "independent chains, one per GPU, seed + rank (tutorials/holstein_honeycomb_mpi.jl mode; weak scaling)",
                  "trajectories_per_s": world * args.steps / (max_over_ranks(ms_c) * 1e-3),
                  "ms_per_step": [round(float(q[0]), 2) for q in allr], "cg_iters_per_trajectory": [round(float(q[1]), 1) for q in allr]}
        # ---- value (north_star's target curve): ONE chain whose CG solves are tau-slab partitioned over all GPUs (strong scaling).
        #      A watchdog prints the line with an error note if this section stalls.
        import threading
        stage = ["start"]
        state = {}

        def bail():
            if rank == 0:
                emit({"metric": "efa_hmc_trajectories_per_s", "value": None, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                      "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "independent_chains": chains,
                      "error": "tau-slab section exceeded its time limit at stage '%s'" % stage[0]})
            os._exit(0)
        dog = threading.Timer(420.0, bail)
        dog.daemon = True
        dog.start()
        n = m.N * m.Ltau
        g = torch.Generator(device="cpu").manual_seed(99)
        d_b = torch.randn(n, 2, dtype=torch.float64, generator=g).to(dev)
        d_x1 = torch.zeros_like(d_b)
        d_xN = torch.zeros_like(d_b)
        elph.x = x0
        elph.update_fdm()
        stage[0] = "one-GPU solve"
        it1, eps1 = fdm.cg_dev(d_x1.data_ptr(), d_b.data_ptr(), True, tol=1e-10, maxiter=MAXITER)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fdm.cg_dev(d_xN.data_ptr(), d_b.data_ptr(), True, tol=1e-300, maxiter=400)
        torch.cuda.synchronize()
        us1 = max_over_ranks(time.perf_counter() - t0) / 400 * 1e6
        def kpm_iter_us():
            """preconditioned CG, microseconds per iteration (short solves: the recurrence breaks down once the residual is exactly zero)"""
            P_kpm.update(np.random.default_rng(6).standard_normal(m.N))
            fdm.cg_dev(d_xN.data_ptr(), d_b.data_ptr(), True, preconditioner=P_kpm, tol=1e-300, maxiter=4)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                fdm.cg_dev(d_xN.data_ptr(), d_b.data_ptr(), True, preconditioner=P_kpm, tol=1e-300, maxiter=15)
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t0) / 60 * 1e6
        stage[0] = "one-GPU preconditioned solve"
        d_xk1 = torch.zeros_like(d_b)
        P_kpm.update(np.random.default_rng(6).standard_normal(m.N))
        itk1, _ = fdm.cg_dev(d_xk1.data_ptr(), d_b.data_ptr(), True, preconditioner=P_kpm, tol=1e-10, maxiter=MAXITER)
        usk1 = kpm_iter_us()
        stage[0] = "init sharded solve"
        fdm.init_sharded_solve(dist)
        stage[0] = "sharded solve parity"
        itN, epsN = fdm.cg_dev(d_xN.data_ptr(), d_b.data_ptr(), True, tol=1e-10, maxiter=MAXITER)
        torch.cuda.synchronize()
        err = float((torch.linalg.norm(d_xN - d_x1) / torch.linalg.norm(d_x1)).item())
        barrier()
        t0 = time.perf_counter()
        fdm.cg_dev(d_xN.data_ptr(), d_b.data_ptr(), True, tol=1e-300, maxiter=400)
        torch.cuda.synchronize()
        usN = max_over_ranks(time.perf_counter() - t0) / 400 * 1e6
        stage[0] = "sharded preconditioned solve"
        itkN, _ = fdm.cg_dev(d_xN.data_ptr(), d_b.data_ptr(), True, preconditioner=P_kpm, tol=1e-10, maxiter=MAXITER)
        torch.cuda.synchronize()
        errk = float((torch.linalg.norm(d_xN - d_xk1) / torch.linalg.norm(d_xk1)).item())
        uskN = kpm_iter_us()
        st = fdm.stats
        kpm_slab = {"what": "KPM-preconditioned CG, P^-1 sharded by Matsubara frequency: two all-to-all exchanges (grouped ncclSend/ncclRecv) per apply",
                    "parity": {"err": max_over_ranks(errk), "iters_1gpu": int(itk1), "iters_ngpu": int(itkN)},
                    "cg_us_per_iter_1gpu": usk1, "cg_us_per_iter": uskN, "sharded_solves": st["cg_slab_preconditioned"]}
        tau_slab = {"what": "one chain, every unpreconditioned CG solve tau-slab partitioned over all GPUs", "preconditioned": kpm_slab,
                    "parity": {"what": "same right-hand side solved to 1e-10 on one GPU and sharded over all GPUs",
                               "err": max_over_ranks(err), "iters_1gpu": int(it1), "iters_ngpu": int(itN), "eps_1gpu": eps1, "eps_ngpu": epsN},
                    "cg_us_per_iter_1gpu": us1, "cg_us_per_iter": usN, "cg_speedup_vs_1gpu": us1 / usN,
                    "solver": "resident kernels + peer-mapped mailboxes" if st["cg_slab_resident"] else "host-launched NCCL loop",
                    "comm": "grid-wide sums and boundary slices as device-initiated stores into peer-mapped mailboxes (CUDA IPC over NVLink); "
                            "solution gathered with one grouped ncclBroadcast per solve"}
        stage[0] = "sharded chain"
        trajectory(fresh_updater(4242), None)                    # untimed: tunes the kernels for the slab range
        sampler = ClockSampler(local) if rank == 0 else None
        ms_dev, iters_total, accepted, launches = timed_chain(None, args.steps, 4242)     # same seeds on every rank: one chain
        clocks = sampler.stop() if sampler else None
        ms_value = max_over_ranks(ms_dev)
        value = args.steps / (ms_value * 1e-3)                   # ONE chain: total work fixed as N grows
        xs = elph.x
        h = torch.tensor([float(np.abs(xs).sum()), float((xs * np.arange(1, xs.size + 1).reshape(xs.shape, order="F")).sum())], dtype=torch.float64, device=dev)
        hh = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hh, h)
        tau_slab["ranks_bit_identical"] = bool(all(torch.equal(q, hh[0]) for q in hh))
        stage[0] = "sharded e2e"
        e2e_value = args.steps / e2e_chain(None, args.steps, 4242)
        dog.cancel()
        P, plabel, scaling, where = None, "I", "strong", "gpu_strong"
        cfg_extra["preconditioner_choice"]["note"] = "strong-scaling leg: unpreconditioned (the sharded preconditioned solve is reported separately)"

    iters_per_traj = iters_total / args.steps
    matvecs_per_traj = iters_per_traj + 2 * (NT + 1)
    roofline = None
    if rank == 0 and world == 1:
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

        def cg_iter_time(Pq):
            # fixed iteration count, device-resident vectors, CUDA events on the library stream (one launch per solve on the register
            # path).  Preconditioned: short solves, the recurrence breaks down once the residual reaches exact zero.
            nit, reps = (2000, 1) if Pq is None else (15, 20)
            fdm.cg_dev(d_out.data_ptr(), d_in.data_ptr(), True, preconditioner=Pq, tol=1e-300, maxiter=min(nit, 200))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fdm.cg_dev(d_out.data_ptr(), d_in.data_ptr(), True, preconditioner=Pq, tol=1e-300, maxiter=nit)
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1) * 1e-3 / (nit * reps), nit * reps
        t_iter, nit = cg_iter_time(None)
        t_iter_kpm, _ = cg_iter_time(P_kpm)
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
        resident = (path == 3) and fdm.stats["cg_resident"] > 0
        step_s = ms_dev * 1e-3 / args.steps
        matvec = {"kernel": names.get(path, "global passes") + " (fused M^T M v)", "us_per_launch_cold_l2": t_cold * 1e6,
                  "us_per_launch_hot_l2": t_hot * 1e6, "achieved_cold_l2": Bk / t_cold / 1e9, "frac_cold_l2": Bk / t_cold / 1e9 / peak,
                  "achieved_hot_l2": Bk / t_hot / 1e9, "frac_hot_l2": Bk / t_hot / 1e9 / peak, "matvecs_per_s_hot_l2": 1.0 / t_hot,
                  "traffic": state.get("mtm_dram_bytes_per_launch"),
                  "note": "back-to-back launches on this system are paced in ~2 us steps (profiles/README.md); a single-wave kernel of "
                          "~6 us cannot be timed below that from the host, which is why the solver is one resident launch"}
        if path == 3:
            # the same fused kernel on a batch of vectors at the named size (grid dimension z = vector; what a multi-RHS iteration
            # launches): 10 x 16.4 MB leaves the 126 MB L2, so this is the kernel against HBM rather than against launch latency
            batched = {}
            for nb in (10, 20):
                dbi = torch.randn(nb * n, 2, dtype=torch.float64, device=dev)
                dbo = torch.zeros_like(dbi)
                tb = fdm.time_mul(300 + nb, dbo.data_ptr(), dbi.data_ptr(), 30) * 1e-6
                batched[f"n{nb}"] = {"vectors": nb, "us_per_launch": tb * 1e6, "achieved": nb * Bk / tb / 1e9, "frac": nb * Bk / tb / 1e9 / peak}
                del dbi, dbo
            matvec["batched_native_order"] = batched
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
                         "cg_us_per_iteration_kpm": t_iter_kpm * 1e6,
                         "share_of_step": (iters_per_traj * (t_iter_kpm if use_kpm else t_iter)) / step_s, "matvec_kernel": matvec, "tuning": tuning,
                         "solver_stats": fdm.stats})

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu:
        arm = CpuArm(m, x0, True, False)                     # the stock (KPM) configuration on ONE core: the reference is single-threaded
        v = arm.sample()
        cpu = {"value": v, "unit": "trajectories/s", "cores": arm.threads, "kind": "port", "sample": arm.describe(1), "detail": arm.last}

    line = {"metric": "efa_hmc_trajectories_per_s", "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(m, plabel, where, state_label, cfg_extra),
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": nx * 8, "d2h_bytes_per_step": nx * 8 + 64},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "cg_iters_per_trajectory": iters_per_traj, "matvecs_per_trajectory": matvecs_per_traj,
            "acceptance": accepted / args.steps, "independent_chains": chains, "tau_slab": tau_slab}
    emit(line)
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
    ap.add_argument("--precond", default=os.environ.get("SQ_BENCH_PRECOND", "auto"), choices=["auto", "on", "off"],
                    help="native arm: auto = time one trajectory with preconditioner = I and one with a KPMPreconditioner, keep the faster; "
                         "reference arm: KPM (the stock configuration) unless 'off'")
    ap.add_argument("--no-cpu", action="store_true")
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
