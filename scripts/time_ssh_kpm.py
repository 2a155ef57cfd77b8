import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads
m, rng = workloads.config("C")
torch.cuda.set_stream(torch.cuda.Stream()); m.set_stream(torch.cuda.current_stream().cuda_stream)
P = E.SymmetricKPMPreconditioner(m)
info = E.setup_(P, rng.normal(size=2 * m.Nsites))
print("active", info.active, "orders total", info.total_order, "max", info.max_order)
v = torch.randn(m.Ndim, dtype=torch.float64, device="cuda"); y = torch.empty_like(v)
for key1 in (0, 1):
    m._call("elph_set_tuning", 1, key1)
    for _ in range(5): m._lib.elph_dev_kpm_apply(m.handle, v.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): m._lib.elph_dev_kpm_apply(m.handle, v.data_ptr(), y.data_ptr())
    e1.record(); torch.cuda.synchronize()
    print("SSH config C KPM apply, generic kernels =", key1, ":", e0.elapsed_time(e1) / 50 * 1e3, "us")
m.close()
