import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from smoqyelph_b200 import model as mdl
from oracle import oracle as orc
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
omp = len(sys.argv) > 2 and sys.argv[2] == "omp"
m = mdl.config(name)
rng = np.random.default_rng(0)
Lx = m.lattice_dims[0]; sites = np.arange(m.N)
stag = np.where(((sites % Lx) + (sites // Lx)) % 2 == 0, 1.0, -1.0)
x = np.asfortranarray(1.5 * stag[:, None] + 0.3 * mdl.thermal_fields(m, rng))
f = orc.RefFDM(m, sym=True, omp=omp); e = orc.RefElPh(m, omp=omp); e.set_x(x); e.refresh(f)
print("threads", f.L.ref_num_threads())
b = np.asfortranarray(rng.standard_normal((m.Ltau, m.N)) + 1j * rng.standard_normal((m.Ltau, m.N)))
t0 = time.perf_counter(); 
for _ in range(5): f.mul_MtM(b)
print("MtM ms", (time.perf_counter() - t0) / 5 * 1e3)
P = orc.RefKPM(f); P.update(rng.standard_normal(m.N)); print("kpm active", P.active, "orders sum", P.orders.sum() * 2)
t0 = time.perf_counter()
for _ in range(3): P.ldiv(b)
print("kpm ldiv ms", (time.perf_counter() - t0) / 3 * 1e3)
for pre in (None, P):
    t0 = time.perf_counter(); xs, it, eps = f.cg(b, P=pre, tol=1e-5, maxiter=40); dt = time.perf_counter() - t0
    print("cg 40 iters precond", pre is not None, "ms/iter", dt / 40 * 1e3, eps)
