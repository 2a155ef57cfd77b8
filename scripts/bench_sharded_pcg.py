"""KPM-preconditioned CG on ONE tau-sharded lattice (ShardedKPM: omega-sharded application of the preconditioner) against the
plain CG of the same sharded lattice.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        scripts/bench_sharded_pcg.py [Lside] [Ltau] [--p2p] [--fused]

Prints one JSON line on rank 0.  Source of the `tau_sharded.pcg` numbers in DESIGN.md / profiles.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elphdynamics_b200 as E
from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedKPM, ShardedOperator, slab_bounds

args = [a for a in sys.argv[1:] if not a.startswith("--")]
Ls = int(args[0]) if len(args) > 0 else 64
Lglob = int(args[1]) if len(args) > 1 else 400
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
torch.cuda.set_stream(torch.cuda.Stream())
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DTAU = 0.1


def make_model(L):
    m = E.HolsteinModel(E.Lattice(E.UnitCell(2, 1), Ls), L * DTAU, DTAU, tol=1e-5, maxiter=10000)
    m.assign_omega(1.0); m.assign_lambda(1.0); m.assign_mu(-1.0)
    m.assign_t(1.0, 0, 0, (1, 0, 0)); m.assign_t(1.0, 0, 0, (0, 1, 0))
    m.initialize_model_()
    return m


tau0, lloc = slab_bounds(Lglob, world, rank)
m = make_model(lloc)
aux = make_model(Lglob)
rs = np.random.default_rng(99)
N = m.Nsites
x0 = rs.integers(-1, 2, size=(N, 1)) + 0.7 * rs.normal(size=(N, 1))
xg = x0 + 0.3 * rs.normal(size=(N, Lglob))
bg = rs.normal(size=(Lglob, N))
noise = rs.normal(size=2 * N)
m.x = np.ascontiguousarray(xg[:, tau0:tau0 + lloc]).reshape(-1)
be = CudaSlabBackend(m, tau0, Lglob)
comm = RingComm(rank, world)
op = ShardedOperator(be, comm, tol=1e-5, maxiter=10000)
if "--p2p" in sys.argv:
    op.enable_p2p()
op.update_model()
be.kpm_init(aux)
P = ShardedKPM(op, N, Lglob)
if "--fused" in sys.argv:
    assert P.enable_fused(tau0)
b = be.empty()
b[1:lloc + 1] = torch.from_numpy(bg[tau0:tau0 + lloc]).cuda()
x = be.empty()
out = {"lattice": f"{Ls}x{Ls}xL{Lglob}", "n_gpus": world, "slab_slices": lloc, "halo": "peer memory" if comm.peer_halo else "nccl",
       "kpm_transposes": "peer memory (kpm_shard.cu)" if P.fused else "nccl all-to-all"}


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps):
    fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_setup = timed(lambda: P.setup(noise), 3)
out["kpm_setup_ms"] = t_setup * 1e3
out["kpm"] = {"active": P.active, "orders_max": int(be.kpm_orders().max()), "orders_total": int(be.kpm_orders().sum()),
              "frequencies_this_rank": len(P.my_w)}
z = be.empty()
out["kpm_apply_us"] = timed(lambda: P.ldiv(z, b), 10) * 1e6
y = be.empty()
out["mtm_us"] = timed(lambda: op.mulMTM(y, b), 20) * 1e6
res = {}


def solve():
    x.zero_()
    res["it"], res["eps"] = op.solve_pcg(x, b, P)


t_pcg = timed(solve, 2)
out["pcg"] = {"iters": res["it"], "eps": res["eps"], "seconds": t_pcg, "us_per_iter": t_pcg / res["it"] * 1e6}
if world == 1:
    # the single-GPU engine on the same lattice, field, right-hand side and Arnoldi vectors (its own kpm apply + CG loop)
    import ctypes as C
    mm = make_model(Lglob)
    mm.x = np.ascontiguousarray(xg).reshape(-1)
    mm.set_stream(torch.cuda.current_stream().cuda_stream)
    E.update_model_(mm)
    Pm = E.SymmetricKPMPreconditioner(mm)
    E.setup_(Pm, noise)
    bd = b[1:lloc + 1].contiguous().reshape(-1)
    xd = torch.zeros_like(bd)
    it3, eps3 = C.c_int64(), C.c_double()

    def one():
        xd.zero_()
        st = mm._lib.elph_dev_cg_solve(mm.handle, bd.data_ptr(), xd.data_ptr(), 1, 0.0, 0, C.byref(it3), C.byref(eps3))
        assert st == 0, mm._lib.elph_last_error(mm.handle)
    t3 = timed(one, 2)
    out["single_gpu_engine_pcg"] = {"iters": it3.value, "eps": eps3.value, "seconds": t3, "us_per_iter": t3 / max(it3.value, 1) * 1e6,
                                    "x_relerr_vs_sharded": float((xd - x[1:lloc + 1].reshape(-1)).norm() / xd.norm())}
    mm.close()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
m.close()
aux.close()
if world > 1:
    dist.destroy_process_group()
