"""Chain kernels of the KPM apply at 64x64xL400, wide (8-CTA cluster) against 2-CTA, for all frequencies and for the lowest
frequency alone (the critical path): CUDA-event times."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads

m, rng = workloads.holstein("square", 64, 40.0, 0.1, mu=-1.0, seed=5)
m.set_stream(torch.cuda.current_stream().cuda_stream)
P = E.SymmetricKPMPreconditioner(m)
info = E.setup_(P, rng.normal(size=2 * m.Nsites))
L, N = m.Ltau, m.Nsites
nu_in = torch.randn(L, N, dtype=torch.complex128, device="cuda")
nu_out = torch.zeros_like(nu_in)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for first, stride, label in ((0, 1, "all 200 frequencies"), (0, 200, "w = 0 alone"), (1, 200, "w = 1 alone"), (0, 8, "25 frequencies (w = 0, 8, ...)")):
    m._call("elph_kpm_set_omega_subset", first, stride)
    for wide in (1, 0):
        m._call("elph_set_tuning", 26, wide)
        for _ in range(2):
            m._call("elph_dev_kpm_chains", nu_in.data_ptr(), nu_out.data_ptr())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            m._call("elph_dev_kpm_chains", nu_in.data_ptr(), nu_out.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        print(f"{label}: wide={wide} {e0.elapsed_time(e1) * 1e3 / 5:.1f} us  (max order {info.max_order})")
m.close()
