"""Per-phase cycle counters of the pipelined CG kernel (tuning key 12), averaged over the CTAs.  Development aid.

    python scripts/prof_cgpipe.py
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
PH = ["wait_coeff", "update", "post", "fetch_rows", "ghost+product", "publish", "stop+x"]
for (Ls, beta, yss) in ((32, 20.0, (1,)), (64, 5.0, (4,)), (64, 10.0, (4,))):
    m, rng = workloads.holstein("square", Ls, beta, 0.1, seed=1234, eps=0.3)
    lib = m._lib
    lib.elph_debug_pipe_prof.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    b = torch.from_numpy(rng.normal(size=m.Ndim)).cuda()
    lib.elph_set_tuning(m.handle, 10, 1)
    lib.elph_set_tuning(m.handle, 12, 1)
    for ys in yss:
        lib.elph_set_tuning(m.handle, 11, ys)
        it, ep = C.c_int64(), C.c_double()
        x = torch.zeros_like(b)
        for _ in range(2):
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = lib.elph_dev_cg_solve(m.handle, b.data_ptr(), x.data_ptr(), 0, 0.0, 0, C.byref(it), C.byref(ep))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        var = C.c_int32()
        lib.elph_get_tuning(m.handle, 100, C.byref(var))
        ys = (var.value // 10) % 10
        ncta = m.Ltau * ys
        buf = np.zeros((ncta + ys, 8), dtype=np.uint64)
        lib.elph_debug_pipe_prof(m.handle, ncta + ys, buf.ctypes.data)
        red = buf[ncta, :6].astype(np.float64) / it.value
        per = buf[:ncta, :7].astype(np.float64) / it.value
        out = {"lattice": f"{Ls}x{Ls}xL{m.Ltau}", "variant": var.value, "iters": it.value, "us_per_iter": dt / it.value * 1e6,
               "cycles_per_iter_mean": {k: round(float(v), 1) for k, v in zip(PH, per.mean(axis=0))},
               "cycles_per_iter_max_cta": {k: round(float(v), 1) for k, v in zip(PH, per.max(axis=0))},
               "cycles_per_iter_min_cta": {k: round(float(v), 1) for k, v in zip(PH, per.min(axis=0))},
               "busy_cycles_percentiles_0_25_50_75_100": [round(float(v), 1) for v in np.percentile(per[:, 1:].sum(axis=1), [0, 25, 50, 75, 100])],
               "reducer_cycles_per_iter": dict(zip(["own_slots_wait", "all_slots_wait", "cross_gpu", "coeff", "push", "stop_rule"], [round(float(v), 1) for v in red])),
               "total_cycles_mean": round(float(per.sum(axis=1).mean()), 1)}
        print(json.dumps(out), flush=True)
    m.close()
