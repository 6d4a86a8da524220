"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two modes:
  * independent chains -- the reference's only parallel mode (tutorials/holstein_honeycomb_mpi.jl:60-72):
    rank r runs its own Markov chain with seeds `chain_seed(seed, r, component)`; nothing is communicated during sampling and
    the per-chain statistics are merged at the end (`merge_chain_statistics`).
  * tau-slab partitioning -- rank g owns the contiguous slices [lo, hi) of every [l][i] array; M couples
    slice l to l-1 and M^T to l+1, so a matvec needs one boundary slice from each ring neighbour
    (`slab_range`, `halo_plan`).  The wrap-around link (l = 0 <- L-1) carries the antiperiodic + sign.

Everything here is pure host logic and runs under the gloo backend in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def chain_seed(seed: int, rank: int, component: int = 0) -> int:
    """Seed of the chain run by `rank` (the MPI tutorial seeds its Xoshiro with seed + pID; for the counter-based Philox streams of
    this library consecutive integers would make chain r's Greens stream equal chain r + 1's HMC stream, so the seed is a splitmix64
    hash of (seed, rank, component) -- component 0 the chain's numpy rng, 1 HMC, 2 GreensEstimator, 3 pseudofermion noise, 4 KPM)."""
    m = (1 << 64) - 1
    z = (int(seed) + 0x9E3779B97F4A7C15 * (1 + int(rank) + 1000003 * int(component))) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return (z ^ (z >> 31)) & m


def slab_range(Ltau: int, world: int, rank: int):
    """Contiguous, balanced partition of the Ltau slices: the first Ltau % world ranks get one extra slice."""
    base, extra = divmod(int(Ltau), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def halo_plan(Ltau: int, world: int, rank: int):
    """Ring neighbours and signs for the one-slice halos of M (needs v[lo-1]) and M^T (needs v[hi])."""
    lo, hi = slab_range(Ltau, world, rank)
    prev_rank, next_rank = (rank - 1) % world, (rank + 1) % world
    return {
        "lo": lo, "hi": hi, "prev": prev_rank, "next": next_rank,
        # (M v)[l] = v[l] - B_l v[l-1] for l >= 1, + for l = 0: the halo received from `prev` enters with this sign
        "sign_from_prev": 1.0 if lo == 0 else -1.0,
        # (M^T v)[l] = v[l] - B_{l+1}^T v[l+1] for l < L-1, + for l = L-1
        "sign_from_next": 1.0 if hi == Ltau else -1.0,
    }


def rhs_owner(j: int, world: int) -> int:
    """Rank that solves measurement system j of a sharded chain (greens.cu: column j on rank j mod world)."""
    return int(j) % int(world)


def frequency_share(schedule, Ltau: int, world: int, rank: int):
    """This rank's share of the Chebyshev schedule in tau-slab mode: the scheduled Matsubara frequencies (order > 1, longest recurrence
    first) that fall into the rank's index range -- a rank owns the frequencies with the same indices as its time slices (slab.cu,
    kpm_ldiv_slab)."""
    lo, hi = slab_range(Ltau, world, rank)
    return [int(n) for n in schedule if lo <= int(n) < hi]


def partial_dft(v_own, lo: int, Ltau: int, inverse: bool = False):
    """The contribution of the rows [lo, lo + len(v_own)) to the length-Ltau DFT along axis 0 (all other rows zero): what every rank
    computes before the all-to-all of the tau-FFT preconditioner.  Summed over the ranks it is the full transform (the DFT is linear)."""
    v_own = np.asarray(v_own)
    pad = np.zeros((Ltau,) + v_own.shape[1:], dtype=np.complex128)
    pad[lo:lo + v_own.shape[0]] = v_own
    return (np.fft.ifft(pad, axis=0) * Ltau) if inverse else np.fft.fft(pad, axis=0)


def merge_chain_statistics(values, dist=None):
    """Mean and standard error over chains of per-chain means.  `values`: 1-D array of this rank's per-chain
    observables.  With torch.distributed initialised the chains of all ranks are gathered first."""
    v = np.atleast_1d(np.asarray(values, dtype=np.float64))
    if dist is not None and dist.is_initialized():
        import torch
        t = torch.from_numpy(v.copy())
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        v = torch.stack(out).numpy()          # (world, nobs)
    else:
        v = v[None, :]
    mean = v.mean(axis=0)
    err = v.std(axis=0, ddof=1) / np.sqrt(v.shape[0]) if v.shape[0] > 1 else np.zeros_like(mean)
    return mean, err


def max_over_ranks(seconds: float, dist=None) -> float:
    if dist is not None and dist.is_initialized():
        import torch
        t = torch.tensor([seconds], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(seconds)
