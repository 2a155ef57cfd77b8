"""Launch list of ONE application of the sharded KPM preconditioner with the transposes through the arenas (csrc/kpm_shard.cu),
world = 1 (the ring closes on the GPU itself):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python scripts/prof_kpm_shard.py [Lside] [Ltau]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedKPM, ShardedOperator

Ls = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
DTAU = 0.1


def make_model(Lt):
    m = E.HolsteinModel(E.Lattice(E.UnitCell(2, 1), Ls), Lt * DTAU, DTAU, tol=1e-5, maxiter=10000)
    m.assign_omega(1.0); m.assign_lambda(1.0); m.assign_mu(-1.0)
    m.assign_t(1.0, 0, 0, (1, 0, 0)); m.assign_t(1.0, 0, 0, (0, 1, 0))
    m.initialize_model_()
    return m


m, aux = make_model(L), make_model(L)
rs = np.random.default_rng(99)
N = m.Nsites
m.x = (rs.integers(-1, 2, size=(N, 1)) + 0.7 * rs.normal(size=(N, 1)) + 0.3 * rs.normal(size=(N, L))).reshape(-1)
be = CudaSlabBackend(m, 0, L)
op = ShardedOperator(be, RingComm(0, 1))
op.update_model()
be.kpm_init(aux)
P = ShardedKPM(op, N, L)
assert P.enable_fused(0)
P.setup(rs.normal(size=2 * N))
b, z = be.empty(), be.empty()
b[1:L + 1].normal_()
P.ldiv(z, b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
P.ldiv(z, b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
be.kpm_shard_check()
m.close()
aux.close()
