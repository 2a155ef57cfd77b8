"""Quick device-side timing of the fused M^T M kernel and the CG loop (development aid)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from helpers import engine_holstein_like, oracle_holstein

Ls = int(sys.argv[1]) if len(sys.argv) > 1 else 32
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
om, rng = oracle_holstein("square", Ls, beta, 0.1, mu=-1.0)
em = engine_holstein_like(om)
st = torch.cuda.current_stream()
em.set_stream(st.cuda_stream)
n = om.Ndim
v = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(v)
lib = em._lib


def timeit(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


for chunk in (1, 2, 4, 0):
    lib.elph_set_chunk(em.handle, chunk)
    us = timeit(lambda: lib.elph_dev_mulMTM(em.handle, v.data_ptr(), y.data_ptr()))
    print(f"MTM single  chunk={chunk}: {us:8.2f} us  -> {24*n/us/1e3:8.1f} GB/s algorithmic")
for nrep in (16, 64, 128):
    V = torch.randn(nrep, n, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(V)
    D = torch.rand(nrep, n, dtype=torch.float64, device="cuda") + 0.5
    for chunk in (1, 2, 4, 8):
        lib.elph_set_chunk(em.handle, chunk)
        us = timeit(lambda: lib.elph_dev_mulMTM_replicas(em.handle, nrep, D.data_ptr(), n, V.data_ptr(), Y.data_ptr(), n), iters=50, warm=5)
        print(f"MTM replicas={nrep} chunk={chunk}: {us:8.2f} us -> {24*n*nrep/us/1e3:8.1f} GB/s, {nrep/us*1e6:10.0f} matvecs/s")
lib.elph_set_chunk(em.handle, 0)
# CG
g = rng.normal(size=n)
b = np.zeros(n)
E.mulMT_(b, em, g)
x = np.zeros(n)
t0 = time.time(); it, res, flag = E.ldiv_(x, em, b); t1 = time.time()
print(f"CG (host API): iters={it} resid={res:.3e} flag={flag} wall={1e3*(t1-t0):.2f} ms -> {(t1-t0)/it*1e6:.2f} us/iter")
P = E.SymmetricKPMPreconditioner(em)
info = E.setup_(P, rng.normal(size=2*om.N))
print("kpm", info.active, info.e_min, info.e_max, info.total_order, info.max_order)
x = np.zeros(n)
t0 = time.time(); it, res, flag = E.ldiv_(x, em, b, P); t1 = time.time()
print(f"PCG (host API): iters={it} resid={res:.3e} flag={flag} wall={1e3*(t1-t0):.2f} ms -> {(t1-t0)/it*1e6:.2f} us/iter")
z = np.zeros(n)
t0 = time.time()
for _ in range(20): E.kpm_ldiv_(z, P, b)
print(f"KPM apply (host API incl copies): {(time.time()-t0)/20*1e6:.1f} us")
print("launches", em.launch_count())
