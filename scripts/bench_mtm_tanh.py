"""Headline kernel with and without the tanh form of the sweeps (tuning key 24): 256 replicas of 32x32xL200, CUDA events over
200 launches after a 0.3 s soak (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elphdynamics_b200 import workloads
m, rng = workloads.config("B")
torch.cuda.set_stream(torch.cuda.Stream()); m.set_stream(torch.cuda.current_stream().cuda_stream)
R, n = 256, m.Ndim
V = torch.randn(R, n, dtype=torch.float64, device="cuda"); Y = torch.empty_like(V); Y2 = torch.empty_like(V)
base = torch.from_numpy(np.ascontiguousarray(m.expnV.reshape(m.Nsites, m.Ltau).T)).reshape(-1).cuda()
D = base.unsqueeze(0).repeat(R, 1) * (1.0 + 0.01 * torch.rand(R, n, dtype=torch.float64, device="cuda"))
for rep in range(2):
    for key in (0, 1):
        m._call("elph_set_tuning", 24, key)
        out = Y if key else Y2
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.3:
            for _ in range(20): m._lib.elph_dev_mulMTM_replicas(m.handle, R, D.data_ptr(), n, V.data_ptr(), out.data_ptr(), n)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): m._lib.elph_dev_mulMTM_replicas(m.handle, R, D.data_ptr(), n, V.data_ptr(), out.data_ptr(), n)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 5.0
        print(f"tanh form = {key}: {us:7.1f} us per launch, {24.0 * n * R / us / 1e3:7.1f} GB/s")
print("relative difference", float((Y - Y2).norm() / Y2.norm()))
m.close()
