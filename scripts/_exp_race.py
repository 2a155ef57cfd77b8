import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads
torch.cuda.set_device(0); torch.cuda.set_stream(torch.cuda.Stream())
mode = int(sys.argv[1])
m, rng = workloads.holstein("square", 64, 1.2, 0.1, seed=1234, eps=0.3)
lib = m._lib
m.set_stream(torch.cuda.current_stream().cuda_stream)
b = torch.from_numpy(rng.normal(size=m.Ndim)).cuda()
for k, v in ((10, 1), (13, 10), (11, 4), (14, 6), (15, mode)):
    lib.elph_set_tuning(m.handle, k, v)
it, ep = C.c_int64(), C.c_double()
x = torch.zeros_like(b)
st = lib.elph_dev_cg_solve(m.handle, b.data_ptr(), x.data_ptr(), 0, 0.0, 12, C.byref(it), C.byref(ep))
torch.cuda.synchronize()
v = C.c_int32(); lib.elph_get_tuning(m.handle, 100, C.byref(v))
print("mode", mode, "iters", it.value, "status", st, "variant", v.value)
m.close()
