"""tau-sharded M^T M of ONE lattice: microseconds per product with the halo exchange inside the product kernel (tuning key 22),
as exchange kernel + product kernel, and the open-slab product alone (no exchange).  torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 scripts/bench_halo.py [Lside] [Ltau]
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
from elphdynamics_b200.sharded import MTM_MODE, CudaSlabBackend, RingComm, ShardedOperator, slab_bounds

args = [a for a in sys.argv[1:] if not a.startswith("--")]
Ls = int(args[0]) if len(args) > 0 else 64
Lglob = int(args[1]) if len(args) > 1 else 400
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
torch.cuda.set_stream(torch.cuda.Stream())
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tau0, lloc = slab_bounds(Lglob, world, rank)
m = E.HolsteinModel(E.Lattice(E.UnitCell(2, 1), Ls), lloc * 0.1, 0.1, tol=1e-5, maxiter=10000)
m.assign_omega(1.0); m.assign_lambda(1.0); m.assign_mu(-1.0)
m.assign_t(1.0, 0, 0, (1, 0, 0)); m.assign_t(1.0, 0, 0, (0, 1, 0))
m.initialize_model_()
rs = np.random.default_rng(99)
m.x = np.ascontiguousarray((rs.normal(size=(m.Nsites, 1)) + 0.3 * rs.normal(size=(m.Nsites, Lglob)))[:, tau0:tau0 + lloc]).reshape(-1)
be = CudaSlabBackend(m, tau0, Lglob)
comm = RingComm(rank, world)
op = ShardedOperator(be, comm, tol=1e-5, maxiter=10000)
op.update_model()
assert be.p2p_setup(comm)
v = be.empty(); v.normal_(); y = be.empty()
out = {"lattice": f"{Ls}x{Ls}xL{Lglob}", "n_gpus": world, "slab_slices": lloc}


def timeit(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for label, key in (("halo_inside_product_kernel", 1), ("exchange_kernel_then_product", 0)):
    m._call("elph_set_tuning", 22, key)
    out[label + "_us"] = timeit(lambda: be.matvec_halo(MTM_MODE, v, y))
m._call("elph_set_tuning", 22, 1)
out["product_alone_no_exchange_us"] = timeit(lambda: be.matvec(MTM_MODE, v, y))
for c in (1, 2, 3, 4):
    m._call("elph_set_tuning", 0, c)
    out[f"halo_inside_chunk{c}_us"] = timeit(lambda: be.matvec_halo(MTM_MODE, v, y))
m._call("elph_set_tuning", 0, 0)
if rank == 0:
    print(json.dumps(out), flush=True)
m.close()
if world > 1:
    dist.destroy_process_group()
