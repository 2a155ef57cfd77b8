"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_targets.py
Kept tiny: the sanitizer slows kernels down by one to two orders of magnitude."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import elphdynamics_b200 as E
from elphdynamics_b200 import greens as eg
from elphdynamics_b200 import hmc as ehmc
from elphdynamics_b200 import workloads

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def holstein_square():
    m, rng = workloads.holstein("square", 32, 0.4, 0.1, mu=-0.5)          # register-tile kernels, persistent CG (1R)
    v = rng.normal(size=m.Ndim)
    y = np.zeros(m.Ndim)
    for f in (E.mulM_, E.mulMT_, E.mulMTM_):
        f(y, m, v)
    x = np.zeros(m.Ndim)
    print("holstein 32x32 CG", E.ldiv_(x, m, v))
    m._call("elph_set_tuning", 7, 0)
    print("two-reduction CG", E.ldiv_(np.zeros(m.Ndim), m, v))
    P = E.SymmetricKPMPreconditioner(m)
    fa = E.FourierAccelerator(m)
    E.update_Q_(fa, m, 0.0, 10.0, 1.0)
    it = E.evolve_(m, E.RungeKuttaDynamics(m, 1e-3), fa, P, eta=rng.normal(size=m.Ndof), g1=rng.normal(size=m.Ndim),
                   g2=rng.normal(size=m.Ndim), arnoldi1=rng.normal(size=2 * m.Nsites), arnoldi2=rng.normal(size=2 * m.Nsites))
    print("RK step, PCG iterations", it)
    Gr = eg.EstimateGreensFunction(m, 2)
    eg.update_(Gr, m, None, R=rng.normal(size=(2, m.Ndim)))
    eg.setup_pair_(Gr, 0, 1)
    print("greens", abs(Gr.G_D0).max())
    Gr.close()
    m.close()


def generic_and_hmc():
    m, rng = workloads.holstein("honeycomb", 4, 0.4, 0.1, mu=-0.3, tol=1e-7)   # generic kernels, generic persistent CG
    fa = E.FourierAccelerator(m)
    E.update_Q_(fa, m, 0.0, 10.0, 1.0)
    E.update_M_(fa, m, 0.0, 10.0, 1.0, 0.0)
    h = ehmc.HybridMonteCarlo(m, 0.01, 0.03, 0.0, 2)
    print("hmc", ehmc.update_(m, h, fa, None, R_v=rng.normal(size=m.Ndof), R_plus=rng.normal(size=m.Ndim),
                              R_minus=rng.normal(size=m.Ndim), uniform=0.5))
    upd = ehmc.SwapUpdate(m, 1, 1)
    print("swap", ehmc.special_update_(m, h, upd, None, targets=[(0, 1)], R_plus=[rng.normal(size=m.Ndim)],
                                       R_minus=[rng.normal(size=m.Ndim)], uniforms=[0.5]))
    m.close()


def ssh_square():
    m, rng = workloads.ssh_square(Lside=32, beta=0.2, dtau=0.05)           # SSH register-tile kernels, SSH 1R CG
    v = rng.normal(size=m.Ndim)
    y = np.zeros(m.Ndim)
    E.mulMTM_(y, m, v)
    print("ssh 32x32 CG", E.ldiv_(np.zeros(m.Ndim), m, v))
    d = np.zeros(m.Ndof)
    E.muldMdx_(d, v, m, y)
    m.close()


def sharded_p2p():
    import torch
    from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedOperator
    m, rng = workloads.holstein("square", 32, 0.5, 0.1, mu=-0.5)
    be = CudaSlabBackend(m, 0, m.Ltau)
    op = ShardedOperator(be, RingComm(0, 1))
    op.update_model()
    assert op.enable_p2p()
    b, x = be.empty(), be.empty()
    b[1:m.Ltau + 1].normal_()
    print("peer-memory CG (world 1)", op.solve(x, b))
    m.close()


def sharded_kpm_and_ssh():
    """Round 2: the sharded KPM application through the arenas (kpm_shard.cu), the all-to-all form's column FFTs and chain subset,
    the speculative set-up, and the SSH open-slab kernels (world 1: the ring closes on the GPU itself)."""
    import torch
    from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedKPM, ShardedOperator
    m, rng = workloads.holstein("square", 32, 0.8, 0.1, mu=-0.5)
    aux, _ = workloads.holstein("square", 32, 0.8, 0.1, mu=-0.5)
    be = CudaSlabBackend(m, 0, m.Ltau)
    op = ShardedOperator(be, RingComm(0, 1), tol=1e-6)
    op.update_model()
    be.kpm_init(aux)
    P = ShardedKPM(op, m.Nsites, m.Ltau)
    noise = rng.normal(size=2 * m.Nsites)
    P.setup(noise)
    b, x = be.empty(), be.empty()
    b[1:m.Ltau + 1].normal_()
    print("sharded PCG, all-to-all form", op.ldiv(x, b, P=P))
    assert P.enable_fused(0)
    print("sharded PCG, arena form", op.ldiv(x, b, P=P))
    be.kpm_shard_check()
    m.close()
    aux.close()
    ms, rs = workloads.ssh_square(Lside=32, beta=0.2, dtau=0.05)
    bs = CudaSlabBackend(ms, 0, ms.Ltau)
    ops = ShardedOperator(bs, RingComm(0, 1), tol=1e-6)
    ops.update_model()
    v, y = bs.empty(), bs.empty()
    v[1:ms.Ltau + 1].normal_()
    for f in (ops.mulM, ops.mulMT, ops.mulMTM):
        f(y, v)
    out = bs.empty_field()
    bs.muldMdx(y, v, out, 1.0)
    print("ssh slab", float(y.abs().max()), float(out.abs().max()))
    ms.close()


for name, fn in (("holstein", holstein_square), ("generic", generic_and_hmc), ("ssh", ssh_square), ("p2p", sharded_p2p),
                 ("shardkpm", sharded_kpm_and_ssh)):
    if which in ("all", name):
        fn()
print("sanitize targets done")
