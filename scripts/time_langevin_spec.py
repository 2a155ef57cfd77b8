"""Langevin RK steps/s at config B through the C ABI with page-locked host noise, speculative KPM set-up on / off (tuning key 25),
several repetitions of the bench's 5-step measurement."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

for spec in (1, 0, 1, 0):
    m, rng = workloads.config("B")
    m._call("elph_set_tuning", 25, spec)
    fa = E.FourierAccelerator(m)
    E.update_Q_(fa, m, 0.0, 10.0, 1.0)
    P = E.SymmetricKPMPreconditioner(m)
    dyn = E.RungeKuttaDynamics(m, 1e-3)
    n, Ns = m.Ndim, m.Nsites
    noise = [dict(eta=rng.normal(size=n), g1=rng.normal(size=n), g2=rng.normal(size=n), arnoldi1=rng.normal(size=2 * Ns),
                  arnoldi2=rng.normal(size=2 * Ns)) for _ in range(6)]
    for nz in noise:
        for key in ("eta", "g1", "g2"):
            m.pin_host(nz[key])
    E.evolve_(m, dyn, fa, P, **noise[5])
    rates = []
    for rep in range(4):
        t0 = time.perf_counter()
        its = [E.evolve_(m, dyn, fa, P, **noise[k]) for k in range(5)]
        rates.append(5 / (time.perf_counter() - t0))
    print("speculate", spec, "steps/s per repetition", [round(r, 1) for r in rates], "iters", its)
    m.close()
