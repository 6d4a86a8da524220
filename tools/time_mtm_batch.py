"""Fused M^T M v of the register path on a batch of n vectors at the named size (what the multi-RHS solver launches per iteration):
time per launch and fraction of the measured HBM peak, library order (op 200 + n) and native order (op 300 + n), L2-cold by construction
once n x 16 MB leaves the 126 MB L2.  argv: config"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoqyelph_b200 import model as mdl, api
import bench
m = mdl.config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
fdm = api.FermionDetMatrix(m, sym=True)
elph = api.ElectronPhononParameters(m, fdm)
elph.x = bench.bench_state(m)[0] if m.name == "cfg4" else m.random_fields(np.random.default_rng(0), smooth=True)
elph.update_fdm()
peak, _ = bench.measured_peak()
V = m.N * m.Ltau
Bu = 40 * m.N * m.Ltau + 16 * m.Nh
for S in (3, 5):
    fdm.set_fast_path(2 + 256 * S)
    for n in (1, 2, 4, 10, 20):
        d_in = torch.randn(n * V, 2, dtype=torch.float64, device="cuda"); d_out = torch.zeros_like(d_in)
        for base, label in ((200, "library order"), (300, "native order")):
            us = fdm.time_mul(base + n, d_out.data_ptr(), d_in.data_ptr(), 50)
            print(f"{m.name} S={S} n={n:2d} {label:14s}: {us:7.1f} us per launch = {us/n:6.2f} us per vector, {n*Bu/us/1e3:7.0f} GB/s = {n*Bu/us/1e3/peak:.3f} of the measured peak", flush=True)
