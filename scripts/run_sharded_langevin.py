"""A Holstein Langevin chain on ONE lattice tau-sharded over the GPUs of a node, KPM-preconditioned, through the public Python
layer (elphdynamics_b200.sharded) -- the multi-GPU counterpart of the loop in src/RunSimulation.jl:25-140.  One rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 \
        scripts/run_sharded_langevin.py [Lside] [Ltau] [nsteps] [rk|euler|heun]

Every rank draws the SAME global noise from a seeded generator and keeps its slab of it (a production driver would draw only its
slab from a counter-based generator).  Prints per step: preconditioned CG iterations, <x>, <x^2> over the whole lattice.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200.sharded import (CudaSlabBackend, RingComm, ShardedKPM, ShardedLangevin, ShardedOperator, slab_bounds)

args = sys.argv[1:]
Ls = int(args[0]) if len(args) > 0 else 32
L = int(args[1]) if len(args) > 1 else 200
nsteps = int(args[2]) if len(args) > 2 else 5
method = args[3] if len(args) > 3 else "rk"
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DTAU, DT = 0.1, 1e-3


def make_model(Lt):
    """examples/holstein_langevin_square.toml:39-76: t = 1, omega = 1, lambda = 1, mu = -1."""
    m = E.HolsteinModel(E.Lattice(E.UnitCell(2, 1), Ls), Lt * DTAU, DTAU, tol=1e-5, maxiter=10000)
    m.assign_omega(1.0); m.assign_lambda(1.0); m.assign_mu(-1.0)
    m.assign_t(1.0, 0, 0, (1, 0, 0)); m.assign_t(1.0, 0, 0, (0, 1, 0))
    m.initialize_model_()
    return m


tau0, lloc = slab_bounds(L, world, rank)
slab, aux = make_model(lloc), make_model(L)           # this rank's slab; the global-lattice handle of the preconditioner
N = slab.Nsites
rng = np.random.default_rng(2024)
x0 = rng.integers(-1, 2, size=(1, N)) + 0.7 * rng.normal(size=(1, N)) + 0.3 * rng.normal(size=(L, N))     # [tau][site]
be = CudaSlabBackend(slab, tau0, L)
comm = RingComm(rank, world)
op = ShardedOperator(be, comm, tol=1e-5, maxiter=10000)
op.enable_p2p()                                       # halo exchange inside the product kernel, peer-memory CG for the fallback
be.make_fft_plan(L)
be.kpm_init(aux)                                      # n = 20, buf = 0.05, c1 = c2 = 1 (examples/...toml:154-168)
P = ShardedKPM(op, N, L)
P.enable_fused(tau0)                                  # transposes of the preconditioner through peer memory
# Fourier acceleration diagonal Q(k) of this rank's site block (src/FourierAcceleration.jl:213-217, mass 1)
s0, nloc = slab_bounds(N, world, rank)
k = np.arange(L)[:, None]
Q = (1.0 + DTAU * 1.0 + 4.0 / DTAU) / (1.0 + DTAU * 1.0 + (2 - 2 * np.cos(2 * np.pi * k / L)) / DTAU) * np.ones((1, nloc))
lang = ShardedLangevin(op, N, L, tau0, torch.from_numpy(np.ascontiguousarray(Q)).cuda(), DT, P=P)
lang.set_x(x0[tau0:tau0 + lloc])


def slab_of(a):
    t = be.empty()
    t[1:lloc + 1] = torch.from_numpy(np.ascontiguousarray(a[tau0:tau0 + lloc])).cuda()
    return t


step = {"rk": lang.evolve_rk, "heun": lang.evolve_heun}.get(method)
t0 = time.perf_counter()
for n in range(nsteps):
    eta, g1, g2 = rng.normal(size=(L, N)), rng.normal(size=(L, N)), rng.normal(size=(L, N))
    a1, a2 = rng.normal(size=2 * N), rng.normal(size=2 * N)
    if method == "euler":
        it = lang.evolve_euler(slab_of(eta), slab_of(g1), a1)
    else:
        it = step(slab_of(eta), slab_of(g1), slab_of(g2), a1, a2)
    own = lang.xh[1:lloc + 1]
    sums = torch.stack([own.sum(), (own * own).sum()])
    comm.allreduce_sum(sums)
    if rank == 0:
        print(json.dumps({"step": n, "pcg_iters": it, "flag": lang.last_flag, "kpm_active": P.active,
                          "x_mean": float(sums[0]) / (L * N), "x2_mean": float(sums[1]) / (L * N)}), flush=True)
if P.fused:
    be.kpm_shard_check()
if rank == 0:
    print(json.dumps({"lattice": f"{Ls}x{Ls}xL{L}", "n_gpus": world, "method": method, "steps": nsteps,
                      "seconds_per_step": (time.perf_counter() - t0) / nsteps}), flush=True)
if world > 1:
    dist.barrier()
slab.close()
aux.close()
if world > 1:
    dist.destroy_process_group()
