import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads
torch.cuda.set_device(0); torch.cuda.set_stream(torch.cuda.Stream())
cases = [(64, 40.0, 10, 4, 6, 1988)]
for (Ls, beta, variant, ys, spc, expect) in cases:
    for mode in (1, 4, 5, 6, 0, 1):
        bad = []; ts = []
        for trial in range(10):
            m, rng = workloads.holstein("square", Ls, beta, 0.1, seed=1234, eps=0.3)
            lib = m._lib
            m.set_stream(torch.cuda.current_stream().cuda_stream)
            b = torch.from_numpy(rng.normal(size=m.Ndim)).cuda()
            for k, v in ((10, 1), (13, variant), (11, ys), (14, spc), (15, mode)):
                lib.elph_set_tuning(m.handle, k, v)
            for rep in range(2):
                it, ep = C.c_int64(), C.c_double()
                x = torch.zeros_like(b)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                st = lib.elph_dev_cg_solve(m.handle, b.data_ptr(), x.data_ptr(), 0, 0.0, 0, C.byref(it), C.byref(ep))
                torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) / max(it.value, 1) * 1e6)
                if abs(it.value - expect) > 2 or st != 0:
                    bad.append((trial, rep, it.value, st))
            m.close()
        print(f"L{int(beta*10)} {Ls}x{Ls} v{variant} ys{ys} spc{spc} sync_mode {mode}: us/iter {min(ts):.1f} failures {bad}", flush=True)
