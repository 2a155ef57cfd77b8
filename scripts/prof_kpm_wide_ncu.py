"""ncu --set full target: the chain of the lowest frequency alone at 64x64xL400, wide (8-CTA cluster) then 2-CTA kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

m, rng = workloads.holstein("square", 64, 40.0, 0.1, mu=-1.0, seed=5)
m.set_stream(torch.cuda.current_stream().cuda_stream)
P = E.SymmetricKPMPreconditioner(m)
E.setup_(P, rng.normal(size=2 * m.Nsites))
L, N = m.Ltau, m.Nsites
nu_in = torch.randn(L, N, dtype=torch.complex128, device="cuda")
nu_out = torch.zeros_like(nu_in)
m._call("elph_kpm_set_omega_subset", 0, 200)
for wide in (1, 0):
    m._call("elph_set_tuning", 26, wide)
    m._call("elph_dev_kpm_chains", nu_in.data_ptr(), nu_out.data_ptr())
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m._call("elph_dev_kpm_chains", nu_in.data_ptr(), nu_out.data_ptr())
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
m.close()
