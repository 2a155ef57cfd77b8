"""ncu target: SSH replica batch (elph_dev_mulMTM_replicas_ssh) at config C, 128 replicas, 48 B per lattice point.
    ncu --set full --clock-control none -k regex:ssh_square -s 2 -c 1 python scripts/prof_ssh_replicas.py"""
import sys

import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elphdynamics_b200 import workloads

m, rng = workloads.config("C")
R, n = 128, m.Ndim
L, N, Nph = m.Ltau, m.Nsites, m.Nph
ts = 4 * L * N
X = 0.3 * torch.randn(R, Nph * L, dtype=torch.float64, device="cuda")
T = torch.empty(R * ts, dtype=torch.float64, device="cuda")
V = torch.randn(R, n, dtype=torch.float64, device="cuda")
Y = torch.empty_like(V)
m._call("elph_dev_ssh_replica_tables", R, X.data_ptr(), Nph * L, T.data_ptr(), ts)
for _ in range(5):
    m._call("elph_dev_mulMTM_replicas_ssh", R, T.data_ptr(), ts, V.data_ptr(), Y.data_ptr(), n)
torch.cuda.synchronize()
print("algorithmic bytes per launch", 48 * n * R)
m.close()
