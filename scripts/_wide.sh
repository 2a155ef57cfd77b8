timeout 300 python -m pytest tests/test_gpu_kpm_wide.py -m gpu -x -q 2>&1 | tail -12
for w in 1 0; do
ELPH_KPM_WIDE=$w timeout 200 python - <<'PY'
import os, sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0, "/root/repo")
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads
wide = int(os.environ["ELPH_KPM_WIDE"])
m, rng = workloads.holstein("square", 64, 40.0, 0.1, mu=-1.0, seed=5)
m._call("elph_set_tuning", 26, wide)
m.set_stream(torch.cuda.current_stream().cuda_stream)
P = E.SymmetricKPMPreconditioner(m)
info = E.setup_(P, rng.normal(size=2 * m.Nsites))
v = torch.randn(m.Ndim, dtype=torch.float64, device="cuda"); z = torch.empty_like(v)
lib = m._lib
for _ in range(3): lib.elph_dev_kpm_apply(m.handle, v.data_ptr(), z.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): lib.elph_dev_kpm_apply(m.handle, v.data_ptr(), z.data_ptr())
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
b = torch.randn(m.Ndim, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
it, eps = C.c_int64(), C.c_double()
best = 1e9
for _ in range(2):
    x.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
    lib.elph_dev_cg_solve(m.handle, b.data_ptr(), x.data_ptr(), 1, 0.0, 0, C.byref(it), C.byref(eps)); torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print("wide", wide, "max order", info.max_order, "apply us", round(us, 1), "pcg iters", it.value, "us/iter", round(best * 1e6 / it.value, 1), "solve ms", round(best * 1e3, 2))
PY
done
