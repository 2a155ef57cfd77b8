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
m._call("elph_set_tuning", 1, 0)
import ctypes as C, time, numpy as np
g = rng.normal(size=m.Ndim); b = np.zeros(m.Ndim); E.mulMT_(b, m, g)
bd = torch.from_numpy(np.ascontiguousarray(b.reshape(m.Nsites, m.Ltau).T)).reshape(-1).cuda()
xd = torch.zeros(m.Ndim, dtype=torch.float64, device="cuda")
it, eps = C.c_int64(), C.c_double()
for fused in (1, 0):
    m._call("elph_set_tuning", 17, fused)
    for rep in range(3):
        xd.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
        m._lib.elph_dev_cg_solve(m.handle, bd.data_ptr(), xd.data_ptr(), 1, 0.0, 0, C.byref(it), C.byref(eps))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"SSH config C PCG, one kernel = {fused}: {it.value} iterations, {dt / it.value * 1e6:.1f} us per iteration")
m._call("elph_set_tuning", 17, 1)
fa = E.FourierAccelerator(m); E.update_Q_(fa, m, 0.0, 10.0, 0.1)
dyn = E.RungeKuttaDynamics(m, 1e-3)
nz = [dict(eta=rng.normal(size=m.Ndof), g1=rng.normal(size=m.Ndim), g2=rng.normal(size=m.Ndim), arnoldi1=rng.normal(size=2 * m.Nsites),
           arnoldi2=rng.normal(size=2 * m.Nsites)) for _ in range(5)]
E.evolve_(m, dyn, fa, P, **nz[0])
t0 = time.perf_counter()
its = [E.evolve_(m, dyn, fa, P, **z) for z in nz[1:]]
print("SSH config C Langevin RK:", 4 / (time.perf_counter() - t0), "steps/s, PCG iterations", its)
m.close()
