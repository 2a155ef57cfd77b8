"""Unpreconditioned CG on one GPU: single-reduction persistent kernel (csrc/cg_p2p.cu, tuning key 7 = 1) against the
two-reduction persistent kernel (csrc/cg_persistent.cu, key 7 = 0).  Prints one JSON line per lattice.  Development aid
and the source of the K3 numbers in DESIGN.md.

    python scripts/bench_cg1r.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
for (Ls, beta, eps_f) in ((32, 20.0, 0.3), (32, 20.0, 1.0), (32, 5.0, 0.3), (64, 10.0, 0.3), (64, 5.0, 0.3), ("C", 10.0, 0.3)):
    if Ls == "C":
        m, rng = workloads.config("C")          # SSH 32x32, Ltau = 200
        Ls = 32
    else:
        m, rng = workloads.holstein("square", Ls, beta, 0.1, seed=1234, eps=eps_f)
    lib = m._lib
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    b = torch.from_numpy(rng.normal(size=m.Ndim)).cuda()
    out = {"model": type(m).__name__, "lattice": f"{Ls}x{Ls}xL{m.Ltau}", "roughness": eps_f}
    xs = {}
    for key7 in (0, 1):
        lib.elph_set_tuning(m.handle, 7, key7)
        it, ep = C.c_int64(), C.c_double()
        x = torch.zeros_like(b)
        best = 1e9
        for _ in range(3):
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = lib.elph_dev_cg_solve(m.handle, b.data_ptr(), x.data_ptr(), 0, 0.0, 0, C.byref(it), C.byref(ep))
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
            assert st == 0, lib.elph_last_error(m.handle)
        y = torch.zeros_like(b)
        lib.elph_dev_mulMTM(m.handle, x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        res = float(torch.linalg.norm(y - b) / torch.linalg.norm(b))
        xs[key7] = x.clone()
        out["single_reduction" if key7 else "two_reductions"] = {"iters": it.value, "eps": ep.value, "true_residual": res,
                                                                  "seconds": best, "us_per_iter": best / it.value * 1e6}
    out["rel_diff_x"] = float(torch.linalg.norm(xs[0] - xs[1]) / torch.linalg.norm(xs[0]))
    print(json.dumps(out), flush=True)
    m.close()
