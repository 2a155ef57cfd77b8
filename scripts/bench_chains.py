"""K independent Langevin chains on one GPU (one handle, stream and host thread each): aggregate Runge-Kutta steps/s for the
one-kernel preconditioned solve with 1/K of the SMs per chain (tuning key 20) against the launch-per-phase solve (key 17 = 0).
Development aid behind the choice made in bench.py."""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

nsm = torch.cuda.get_device_properties(0).multi_processor_count
for K, fused, grid in ((1, 1, 0), (2, 1, 74), (4, 1, 36), (8, 1, 18), (8, 0, 0), (4, 0, 0), (12, 0, 0), (8, 1, 0)):
    chains = []
    nst = 5
    for c in range(K):
        m, r = workloads.holstein("square", 32, 20.0, 0.1, mu=-1.0, seed=4321 + c, eps=0.3)
        n = m.Ndim
        f = E.FourierAccelerator(m)
        E.update_Q_(f, m, 0.0, 10.0, 1.0)
        P = E.SymmetricKPMPreconditioner(m)
        m._call("elph_set_tuning", 17, fused)
        m._call("elph_set_tuning", 20, grid)
        d = E.RungeKuttaDynamics(m, 1e-3)
        nz = [dict(eta=r.normal(size=n), g1=r.normal(size=n), g2=r.normal(size=n), arnoldi1=r.normal(size=2 * m.Nsites),
                   arnoldi2=r.normal(size=2 * m.Nsites)) for _ in range(nst + 1)]
        for z in nz:
            for key in ("eta", "g1", "g2"):
                m.pin_host(z[key])
        chains.append((m, f, P, d, nz))

    def run(c, first, steps):
        m, f, P, d, nz = chains[c]
        for k in range(first, first + steps):
            E.evolve_(m, d, f, P, **nz[k])

    for first, steps in ((0, 1), (1, nst)):
        ths = [threading.Thread(target=run, args=(c, first, steps)) for c in range(K)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        dt = time.perf_counter() - t0
    print(f"K = {K:2d} chains, fused = {fused}, CTAs per solve = {grid or nsm}: {K * nst / dt:8.1f} steps/s aggregate, {nst / dt:7.1f} per chain", flush=True)
    for m, f, P, d, nz in chains:
        for z in nz:
            for key in ("eta", "g1", "g2"):
                m.unpin_host(z[key])
        m.close()
