"""Phase timings of elph_langevin_step (Runge-Kutta, KPM-preconditioned) at config B, through ELPH_TRACE=1 (development aid).
    ELPH_TRACE=1 python scripts/time_langevin.py 2> trace.txt
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

m, rng = workloads.config("B")
fa = E.FourierAccelerator(m)
E.update_Q_(fa, m, 0.0, 10.0, 1.0)
P = E.SymmetricKPMPreconditioner(m)
dyn = E.RungeKuttaDynamics(m, 1e-3)
for step in range(4):
    eta, g1, g2 = rng.normal(size=m.Ndof), rng.normal(size=m.Ndim), rng.normal(size=m.Ndim)
    a1, a2 = rng.normal(size=2 * m.Nsites), rng.normal(size=2 * m.Nsites)
    sys.stderr.write(f"---- step {step}\n")
    sys.stderr.flush()
    t0 = time.perf_counter()
    it = E.evolve_(m, dyn, fa, P, eta=eta, g1=g1, g2=g2, arnoldi1=a1, arnoldi2=a2)
    dt = time.perf_counter() - t0
    sys.stderr.write(f"step {step}: {dt * 1e3:.3f} ms, iters {it}\n")
m.close()
