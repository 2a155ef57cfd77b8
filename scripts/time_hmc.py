"""Timing of the HMC pieces on the config-D lattices (development aid): M^T M product, one CG solve with the
different loop strategies, one whole trajectory."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import elphdynamics_b200 as E
from elphdynamics_b200 import hmc as ehmc
from elphdynamics_b200 import workloads

name = sys.argv[1] if len(sys.argv) > 1 else "D_honeycomb"
m, rng = workloads.config(name)
torch.cuda.set_stream(torch.cuda.Stream())
m.set_stream(torch.cuda.current_stream().cuda_stream)
lib, h, n = m._lib, m.handle, m.Ndim
print(name, "N", m.Nsites, "L", m.Ltau, "Nb", m.Nbonds, "groups", list(m.group_sizes))
v = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(v)


def timeit(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print(f"MTM: {timeit(lambda: lib.elph_dev_mulMTM(h, v.data_ptr(), y.data_ptr())):8.2f} us")
print(f"M  : {timeit(lambda: lib.elph_dev_mulM(h, v.data_ptr(), y.data_ptr())):8.2f} us")
b_dev = torch.randn(n, dtype=torch.float64, device="cuda")
x_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
it, eps = C.c_int64(), C.c_double()
for label, keys in (("persistent", {5: 1}), ("graph", {5: 0, 3: 1}), ("launches", {5: 0, 3: 0})):
    for k, val in keys.items():
        lib.elph_set_tuning(h, k, val)
    for rep in range(3):
        x_dev.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lib.elph_dev_cg_solve(h, b_dev.data_ptr(), x_dev.data_ptr(), 0, 0.0, 0, C.byref(it), C.byref(eps))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"CG {label:10s}: {it.value} iters, {dt * 1e3:.3f} ms, {dt / max(it.value, 1) * 1e6:.2f} us/iter, eps {eps.value:.2e}")
lib.elph_set_tuning(h, 5, 1)
lib.elph_set_tuning(h, 3, 1)
from elphdynamics_b200._lib import SolveInfo
for nrhs in (1, 2, 10):
    Bd = torch.randn(nrhs, n, dtype=torch.float64, device="cuda")
    Xd = torch.zeros_like(Bd)
    infos = (SolveInfo * nrhs)()
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lib.elph_dev_solve_batch(h, nrhs, Bd.data_ptr(), Xd.data_ptr(), 0, 1.0, infos)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    its = [infos[k].iters for k in range(nrhs)]
    print(f"solve_batch nrhs={nrhs:2d}: {dt * 1e3:.3f} ms, {dt / nrhs * 1e3:.3f} ms per solve, iters {its}")

fa = E.FourierAccelerator(m)
E.update_M_(fa, m, 0.0, 10.0, 1.0, 0.0)
hm = ehmc.HybridMonteCarlo(m, 0.01, 0.1, 0.0, 10)
draws = [dict(R_v=rng.normal(size=m.Ndof), R_plus=rng.normal(size=m.Ndim), R_minus=rng.normal(size=m.Ndim),
              uniform=float(rng.uniform())) for _ in range(4)]
for d in draws:
    l0 = m.launch_count()
    t0 = time.perf_counter()
    acc, its = ehmc.update_(m, hm, fa, None, **d)
    dt = time.perf_counter() - t0
    print(f"trajectory: {dt * 1e3:.2f} ms, launches {m.launch_count() - l0}, iters(avg) {its}, accepted {acc}")
lib.elph_set_tuning(h, 9, 0)           # inner loop of the multi-timestep integrator step by step (for comparison)
for d in draws[:2]:
    l0 = m.launch_count()
    t0 = time.perf_counter()
    acc, its = ehmc.update_(m, hm, fa, None, **d)
    dt = time.perf_counter() - t0
    print(f"trajectory, unfused inner loop: {dt * 1e3:.2f} ms, launches {m.launch_count() - l0}, iters(avg) {its}, accepted {acc}")
lib.elph_set_tuning(h, 9, 1)
