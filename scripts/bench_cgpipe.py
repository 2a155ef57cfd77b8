"""Unpreconditioned CG on one GPU: pipelined persistent kernel (csrc/cg_pipe.cu) against the single-reduction
(csrc/cg_p2p.cu) and two-reduction (csrc/cg_persistent.cu) kernels.  One JSON line per lattice.  Development aid and
the source of the K3 numbers in DESIGN.md.

    python scripts/bench_cgpipe.py [quick]
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
quick = "quick" in sys.argv
# (lattice side, beta): 64-wide slabs of 50 / 100 slices = the per-GPU share of config E on 8 / 4 GPUs
cases = [(32, 20.0), ("C", 10.0), (64, 5.0), (64, 10.0), (64, 20.0), (64, 40.0)] if not quick else [(64, 20.0), (64, 40.0)]
for (Ls, beta) in cases:
    if Ls == "C":
        m, rng = workloads.config("C")          # SSH 32x32, Ltau = 200
        Ls = 32
    else:
        m, rng = workloads.holstein("square", Ls, beta, 0.1, seed=1234, eps=0.3)
    lib = m._lib
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    b = torch.from_numpy(rng.normal(size=m.Ndim)).cuda()
    out = {"model": type(m).__name__, "lattice": f"{Ls}x{Ls}xL{m.Ltau}"}
    ref = None
    modes = [("two_reductions", 0, 0, 0, 0), ("single_reduction", 1, 0, 0, 0), ("pipelined", -1, 1, 0, 0)]
    if Ls == 64 and m.Ltau >= 200:
        modes += [(f"pipe_v10_spc{spc}", -1, 1, 4, 10 + 100 * spc) for spc in ((2, 3, 4) if m.Ltau == 200 else (3, 4, 5, 6))]
    if Ls == 32 and type(m).__name__ == "HolsteinModel":
        modes += [("pipe_v7", -1, 1, 1, 7), ("pipe_v1_ys2", -1, 1, 2, 1), ("pipe_v2_ys2", -1, 1, 2, 2), ("pipe_v7_ys2", -1, 1, 2, 7)]
    elif Ls == 64:
        modes += [("pipe_v3_ys4", -1, 1, 4, 3), ("pipe_v8_ys4", -1, 1, 4, 8), ("pipe_v4_ys4", -1, 1, 4, 4), ("pipe_v3_ys8", -1, 1, 8, 3),
                  ("pipe_v8_ys8", -1, 1, 8, 8), ("pipe_v8_ys2", -1, 1, 2, 8), ("pipe_v4_ys8", -1, 1, 8, 4), ("pipe_v9_ys4", -1, 1, 4, 9), ("pipe_v9_ys8", -1, 1, 8, 9),
                  ("pipe_v5_ys2", -1, 1, 2, 5)]
    for name, key7, key10, ys, variant in modes:
        lib.elph_set_tuning(m.handle, 7, key7)
        lib.elph_set_tuning(m.handle, 10, key10)
        lib.elph_set_tuning(m.handle, 11, ys)
        lib.elph_set_tuning(m.handle, 14, variant // 100)
        variant = variant % 100
        lib.elph_set_tuning(m.handle, 13, variant)
        it, ep = C.c_int64(), C.c_double()
        x = torch.zeros_like(b)
        best = 1e9
        ok = True
        for _ in range(3):
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = lib.elph_dev_cg_solve(m.handle, b.data_ptr(), x.data_ptr(), 0, 0.0, 0, C.byref(it), C.byref(ep))
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
            if st != 0:
                out[name] = {"error": lib.elph_last_error(m.handle).decode()}
                ok = False
                break
        if not ok:
            continue
        y = torch.zeros_like(b)
        lib.elph_dev_mulMTM(m.handle, x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        res = float(torch.linalg.norm(y - b) / torch.linalg.norm(b))
        var = C.c_int32()
        lib.elph_get_tuning(m.handle, 100, C.byref(var))
        if variant and var.value // 100 != variant:
            out[name] = "does not fit"
            continue
        if ref is None:
            ref = x.clone()
        spc_ = C.c_int32()
        lib.elph_get_tuning(m.handle, 101, C.byref(spc_))
        out[name] = {"iters": it.value, "true_residual": res, "us_per_iter": round(best / it.value * 1e6, 3),
                     "variant": var.value, "slices_per_cta": spc_.value, "rel_diff_x": float(torch.linalg.norm(x - ref) / torch.linalg.norm(ref))}
    print(json.dumps(out), flush=True)
    m.close()
