"""One CG and one KPM-PCG solve on config B (development aid; run under ncu for a launch list)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from helpers import engine_holstein_like, oracle_holstein

Ls = int(sys.argv[1]) if len(sys.argv) > 1 else 32
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
om, rng = oracle_holstein("square", Ls, beta, 0.1, mu=-1.0)
em = engine_holstein_like(om)
n = om.Ndim
g = rng.normal(size=n)
b = np.zeros(n)
E.mulMT_(b, em, g)
P = E.SymmetricKPMPreconditioner(em)
info = E.setup_(P, rng.normal(size=2 * om.N))
for rep in range(2):
    x = np.zeros(n)
    t0 = time.perf_counter()
    it, res, flag = E.ldiv_(x, em, b)
    t1 = time.perf_counter()
    print(f"CG  iters={it} wall={1e3*(t1-t0):.2f} ms  {1e6*(t1-t0)/it:.2f} us/iter")
    x = np.zeros(n)
    t0 = time.perf_counter()
    it, res, flag = E.ldiv_(x, em, b, P)
    t1 = time.perf_counter()
    print(f"PCG iters={it} wall={1e3*(t1-t0):.2f} ms  {1e6*(t1-t0)/it:.2f} us/iter")
fa = E.FourierAccelerator(em)
E.update_Q_(fa, em, 0.0, 10.0, 1.0)
dyn = E.RungeKuttaDynamics(em, 1e-3)
for rep in range(3):
    t0 = time.perf_counter()
    it = E.evolve_(em, dyn, fa, P, eta=rng.normal(size=n), g1=rng.normal(size=n), g2=rng.normal(size=n),
                   arnoldi1=rng.normal(size=2 * om.N), arnoldi2=rng.normal(size=2 * om.N))
    print(f"RK step: {1e3*(time.perf_counter()-t0):.2f} ms, pcg iters {it}")
