"""ncu target: a few KPM applies on config B (development aid)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import elphdynamics_b200 as E
from helpers import engine_holstein_like, oracle_holstein
om, rng = oracle_holstein("square", 32, 20.0, 0.1, mu=-1.0)
em = engine_holstein_like(om)
P = E.SymmetricKPMPreconditioner(em)
E.setup_(P, rng.normal(size=2 * om.N))
v = torch.randn(om.Ndim, dtype=torch.float64, device="cuda"); y = torch.empty_like(v)
for _ in range(12):
    em._lib.elph_dev_kpm_apply(em.handle, v.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
