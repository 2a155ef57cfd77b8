"""Wall time of setup!(P) (KPM preconditioner, src/KPMPreconditioners.jl:269-320) at config B: host Arnoldi vs device Arnoldi
(tuning key 19).  Development aid; run under `ncu --metrics gpu__time_duration.sum -k regex:arnoldi` for the kernel time."""
import sys
import time

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

m, rng = workloads.config("B")
P = E.SymmetricKPMPreconditioner(m)
noise = rng.normal(size=2 * m.Nsites)
for dev in (0, 1, 0, 1):
    m._call("elph_set_tuning", 19, dev)
    for _ in range(5):
        E.setup_(P, noise)
    t0 = time.perf_counter()
    for _ in range(50):
        info = E.setup_(P, noise)
    dt = (time.perf_counter() - t0) / 50
    print(f"device Arnoldi = {dev}: {dt * 1e6:8.1f} us per set-up   e_min {info.e_min:.12f} e_max {info.e_max:.12f}")
import ctypes as C
m._call("elph_set_tuning", 19, 1)
m._call("elph_set_tuning", 12, 1)
E.setup_(P, noise)
buf = (C.c_ulonglong * 8)()
m._lib.elph_debug_pipe_prof.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
m._lib.elph_debug_pipe_prof(m.handle, 1, buf)
print("arnoldi kernel cycles: run A: products", buf[0], "Gram-Schmidt", buf[1], "| run A^-1: products", buf[2], "Gram-Schmidt", buf[3])
m.close()
