"""Strong scaling of ONE tau-sharded lattice: CG on M^T M with the collectives inside the kernel (csrc/cg_p2p.cu)
against the NCCL-between-launches baseline (ShardedOperator.solve_cg).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_p2p.py [Lside] [Ltau] [--baseline]

Prints one JSON line on rank 0.  Development aid and the source of the `tau_sharded` numbers in DESIGN.md.
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
from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedOperator, slab_bounds

args = [a for a in sys.argv[1:] if not a.startswith("--")]
Ls = int(args[0]) if len(args) > 0 else 64
Lglob = int(args[1]) if len(args) > 1 else 400
baseline = "--baseline" in sys.argv
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
torch.cuda.set_stream(torch.cuda.Stream())
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DTAU = 0.1
tau0, lloc = slab_bounds(Lglob, world, rank)
m = E.HolsteinModel(E.Lattice(E.UnitCell(2, 1), Ls), lloc * DTAU, DTAU, tol=1e-5, maxiter=10000)
m.assign_omega(1.0); m.assign_lambda(1.0); m.assign_mu(-1.0)
m.assign_t(1.0, 0, 0, (1, 0, 0)); m.assign_t(1.0, 0, 0, (0, 1, 0))
m.initialize_model_()
# the same global synthetic field on every world size: generated globally, sliced per rank (host layout site-major)
rs = np.random.default_rng(99)
N = m.Nsites
x0 = rs.integers(-1, 2, size=(N, 1)) + 0.7 * rs.normal(size=(N, 1))
xg = x0 + 0.3 * rs.normal(size=(N, Lglob))
bg = rs.normal(size=(Lglob, N))
m.x = np.ascontiguousarray(xg[:, tau0:tau0 + lloc]).reshape(-1)
be = CudaSlabBackend(m, tau0, Lglob)
comm = RingComm(rank, world)
op = ShardedOperator(be, comm, tol=1e-5, maxiter=10000)
op.update_model()
b = be.empty()
b[1:lloc + 1] = torch.from_numpy(bg[tau0:tau0 + lloc]).cuda()
x = be.empty()
out = {"lattice": f"{Ls}x{Ls}xL{Lglob}", "n_gpus": world, "slab_slices": lloc}
try:
    be.p2p_setup(comm)
    it, eps = be.cg_p2p(x, b)          # warm-up (module load, IPC mappings)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        it, eps = be.cg_p2p(x, b)
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    import ctypes as _C
    var = _C.c_int32()
    m._lib.elph_get_tuning(m.handle, 100, _C.byref(var))
    out["p2p"] = {"iters": it, "eps": eps, "seconds": float(t.item()), "us_per_iter": float(t.item()) / it * 1e6,
                  "kernel": ("pipelined (cg_pipe.cu) variant*100+ys*10+warps = %d" % var.value) if var.value else "single-reduction (cg_p2p.cu)"}
except RuntimeError as e:
    out["p2p"] = {"error": str(e)[:200]}
if baseline:
    x.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it2, eps2 = op.solve_cg(x, b, maxiter=60)      # bounded: the baseline costs ~100 us per iteration
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    out["nccl_launch_baseline"] = {"iters": it2, "us_per_iter": dt2 / it2 * 1e6}
if world == 1:
    # single-GPU engine on the whole lattice for comparison (graph path when the slices are not co-resident)
    import ctypes as C
    lib = m._lib
    it3, eps3 = C.c_int64(), C.c_double()
    mm, _ = E.workloads.holstein("square", Ls, Lglob * DTAU, DTAU, seed=5)
    mm.set_stream(torch.cuda.current_stream().cuda_stream)
    bd = torch.randn(mm.Ndim, dtype=torch.float64, device="cuda")
    xd = torch.zeros_like(bd)
    for _ in range(2):
        xd.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lib.elph_dev_cg_solve(mm.handle, bd.data_ptr(), xd.data_ptr(), 0, 0.0, 0, C.byref(it3), C.byref(eps3))
        torch.cuda.synchronize()
        dt3 = time.perf_counter() - t0
    out["single_gpu_engine"] = {"iters": it3.value, "us_per_iter": dt3 / max(it3.value, 1) * 1e6}
    mm.close()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
m.close()
if world > 1:
    dist.destroy_process_group()
