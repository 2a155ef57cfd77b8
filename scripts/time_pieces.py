"""Warm device-side timing (CUDA events) of the pieces of a PCG iteration on config B (development aid)."""
import sys

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
torch.cuda.set_stream(torch.cuda.Stream())
em.set_stream(torch.cuda.current_stream().cuda_stream)
lib, h, n = em._lib, em.handle, om.Ndim
P = E.SymmetricKPMPreconditioner(em)
info = E.setup_(P, rng.normal(size=2 * om.N))
fa = E.FourierAccelerator(em)
E.update_Q_(fa, em, 0.0, 10.0, 1.0)
v = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(v)
nu = torch.empty(2 * n, dtype=torch.float64, device="cuda")


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


print("kpm orders: total", info.total_order, "max", info.max_order)
print(f"MTM                : {timeit(lambda: lib.elph_dev_mulMTM(h, v.data_ptr(), y.data_ptr())):8.2f} us")
print(f"KPM apply (2 FFT + poly): {timeit(lambda: lib.elph_dev_kpm_apply(h, v.data_ptr(), y.data_ptr())):8.2f} us")
print(f"fourier_accelerate (2 FFT in one kernel): {timeit(lambda: lib.elph_dev_fourier_accelerate(h, v.data_ptr(), y.data_ptr(), 1.0, 0)):8.2f} us")
import ctypes as C
import time
g = rng.normal(size=n)
bh = np.zeros(n)
E.mulMT_(bh, em, g)
b_dev = torch.from_numpy(np.ascontiguousarray(bh.reshape(om.N, om.L).T)).reshape(-1).cuda()
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
    print(f"CG {label:10s}: {it.value} iters, {dt*1e3:.3f} ms, {dt/it.value*1e6:.2f} us/iter")
lib.elph_set_tuning(h, 5, 1)
lib.elph_set_tuning(h, 3, 1)
for label, usep in (("PCG launches per phase", 1), ("PCG one persistent kernel", 1)):
    lib.elph_set_tuning(h, 17, 1 if "persistent" in label else 0)
    for rep in range(3):
        x_dev.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lib.elph_dev_cg_solve(h, b_dev.data_ptr(), x_dev.data_ptr(), usep, 0.0, 0, C.byref(it), C.byref(eps))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{label}: {it.value} iters, {dt*1e3:.3f} ms, {dt/it.value*1e6:.2f} us/iter")
for py in (2, 4, 8):
    lib.elph_set_tuning(h, 2, py)
    print(f"KPM apply py={py}: {timeit(lambda: lib.elph_dev_kpm_apply(h, v.data_ptr(), y.data_ptr())):8.2f} us")
lib.elph_set_tuning(h, 2, 0)
for fast in (0, 1):
    lib.elph_set_tuning(h, 16, fast)
    print(f"KPM apply tanh-form sweeps={fast}: {timeit(lambda: lib.elph_dev_kpm_apply(h, v.data_ptr(), y.data_ptr())):8.2f} us")
    yy = y.clone()
    if fast == 0:
        y0 = yy
    else:
        print("   relative difference between the two forms:", float((yy - y0).norm() / y0.norm()))

# where the chain kernel spends its cycles (cluster of the longest polynomial)
lib.elph_set_tuning(h, 16, 1)
lib.elph_set_tuning(h, 12, 1)
lib.elph_dev_kpm_apply(h, v.data_ptr(), y.data_ptr())
buf = (C.c_ulonglong * 16)()
lib.elph_debug_pipe_prof.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
lib.elph_debug_pipe_prof(h, 2, buf)
for cta in range(2):
    t = [buf[8 * cta + k] for k in range(7)]
    print("chain CTA", cta, "cycles: load", t[1] - t[0], "poly1", t[2] - t[1], "swap", t[3] - t[2], "poly2", t[4] - t[3], "swap", t[5] - t[4],
          "store", t[6] - t[5], "| per sweep", (t[2] - t[1]) / (info.max_order - 1), (t[4] - t[3]) / (info.max_order - 1))
lib.elph_set_tuning(h, 12, 0)

# phases of the fused PCG kernel (CTA 0..3): cycles summed over the iterations
lib.elph_set_tuning(h, 17, 1)
lib.elph_set_tuning(h, 12, 1)
x_dev.zero_()
lib.elph_dev_cg_solve(h, b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, 0, C.byref(it), C.byref(eps))
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
lib.elph_debug_pipe_prof(h, 4, buf)
names = ["FFT phases", "barriers after FFT phases", "chains", "barrier after chains", "product", "barrier after product"]
for cta in range(4):
    print("fused PCG CTA", cta, it.value, "iterations; cycles per iteration:",
          {nm: round(buf[8 * cta + k] / max(1, it.value)) for k, nm in enumerate(names)},
          "of which fft_smem fwd/inv", round(buf[8 * cta + 6] / max(1, it.value)), round(buf[8 * cta + 7] / max(1, it.value)))
lib.elph_set_tuning(h, 12, 0)
