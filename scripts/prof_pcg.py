"""ncu target: KPM set-up + two KPM-preconditioned solves on config B through elph_dev_cg_solve (development aid).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/prof_pcg.py      (launch list)
    ncu --set full --import-source on --clock-control none -k regex:pcg_fused -c 1 python scripts/prof_pcg.py
"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

m, rng = workloads.config("B")
lib, h, n = m._lib, m.handle, m.Ndim
P = E.SymmetricKPMPreconditioner(m)
E.setup_(P, rng.normal(size=2 * m.Nsites))
g = rng.normal(size=n)
b = np.zeros(n)
E.mulMT_(b, m, g)
b_dev = torch.from_numpy(np.ascontiguousarray(b.reshape(m.Nsites, m.Ltau).T)).reshape(-1).cuda()
x_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
it, eps = C.c_int64(), C.c_double()
for rep in range(2):
    x_dev.zero_()
    l0 = m.launch_count()
    lib.elph_dev_cg_solve(h, b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, 0, C.byref(it), C.byref(eps))
    torch.cuda.synchronize()
    print(f"solve {rep}: {it.value} iterations, eps {eps.value:.3e}, {m.launch_count() - l0} engine launches")
m.close()
