"""tau-sharded (multi-GPU) fermion-matrix products and CG: one process per GPU, torch.distributed for the plumbing.

The space-time lattice is cut along imaginary time: rank r owns a contiguous slab of tau-slices of
every vector and of expnV.  M couples slice tau only to tau-1 and M^T only to tau+1, so one product needs ONE halo
exchange (the ring closure between the last and the first rank carries the antiperiodic sign, which the kernels
apply to GLOBAL slice 0), and CG needs two scalar all-reduces per iteration.  The reference has no counterpart
(single process); the arithmetic is the reference's `mulM!/mulMT!/mulMTM!` (src/HolsteinModels.jl:569-684) and
`solve!` (src/IterativeSolvers.jl:239-314).

Vectors are torch tensors of shape (Lloc + 2, N): row 0 = left halo, rows 1..Lloc = own slices, row Lloc+1 = right
halo, in the engine's slice-major layout.  The local arithmetic is delegated to a *slab backend*; the product backend
is `CudaSlabBackend` (libelph_b200.so through the C ABI).  Tests inject a NumPy backend to exercise this host logic
over gloo on CPU; the package itself has no CPU arithmetic.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

M_MODE, MT_MODE, MTM_MODE = 0, 1, 2


def slab_bounds(L: int, world: int, rank: int):
    """Contiguous near-equal tau-slabs: (tau0, Lloc)."""
    base, rem = divmod(L, world)
    lloc = base + (1 if rank < rem else 0)
    tau0 = rank * base + min(rank, rem)
    return tau0, lloc


class RingComm:
    """Halo exchange on the tau-ring and scalar all-reduce over torch.distributed (nccl or gloo)."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.left = (rank - 1) % world
        self.right = (rank + 1) % world

    def exchange(self, v, lloc: int, lo: bool = True, hi: bool = True):
        """Fill v[0] with the left neighbour's last own slice and v[lloc+1] with the right neighbour's first own slice."""
        import torch.distributed as dist
        if self.world == 1:
            if lo:
                v[0].copy_(v[lloc])
            if hi:
                v[lloc + 1].copy_(v[1])
            return
        ops = []
        # sends first (first own slice -> left, last own slice -> right), then receives in the matching order
        # (from the right: its first slice; from the left: its last slice) so that world == 2 pairs them correctly.
        if hi:
            ops.append(dist.P2POp(dist.isend, v[1], self.left, self.group))
        if lo:
            ops.append(dist.P2POp(dist.isend, v[lloc], self.right, self.group))
        if hi:
            ops.append(dist.P2POp(dist.irecv, v[lloc + 1], self.right, self.group))
        if lo:
            ops.append(dist.P2POp(dist.irecv, v[0], self.left, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def allreduce_sum(self, t):
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class CudaSlabBackend:
    """Local slab arithmetic on the GPU through libelph_b200.so (device-pointer entry points)."""

    def __init__(self, model, tau0: int, Lglob: int):
        import torch
        self.torch = torch
        self.model = model                     # HolsteinModel created with Ltau = Lloc
        self.lib = model._lib
        self.h = model.handle
        self.N, self.lloc = model.Nsites, model.Ltau
        self._check(self.lib.elph_set_shard(self.h, tau0, Lglob))
        model.set_stream(torch.cuda.current_stream().cuda_stream)
        self.scal = torch.zeros(2, dtype=torch.float64, device="cuda")
        p = C.c_void_p()
        self._check(self.lib.elph_dev_ptr_expnV(self.h, C.byref(p)))
        # view of the handle's halo'd expnV allocation as a (Lloc+2, N) tensor for the halo exchange
        self._D_ptr = p.value - self.N * 8

    def _check(self, st):
        if st != 0:
            raise RuntimeError(self.lib.elph_last_error(self.h).decode())

    def empty(self):
        return self.torch.zeros(self.lloc + 2, self.N, dtype=self.torch.float64, device="cuda")

    def own_ptr(self, v):
        return v.data_ptr() + self.N * 8

    def D_tensor(self):
        """The handle's expnV with halos, wrapped as a CUDA tensor (no copy)."""
        torch = self.torch
        n = (self.lloc + 2) * self.N
        iface = {"shape": (n,), "typestr": "<f8", "data": (self._D_ptr, False), "version": 3}

        class _W:
            __cuda_array_interface__ = iface
        return torch.as_tensor(_W(), device="cuda").view(self.lloc + 2, self.N)

    def update_model(self):
        self._check(self.lib.elph_dev_update_model(self.h))

    def matvec(self, mode, v, y):
        self._check(self.lib.elph_dev_shard_matvec(self.h, mode, self.own_ptr(v), self.own_ptr(y)))

    def muldMdx(self, u, v, out, scale=1.0):
        self._check(self.lib.elph_dev_shard_muldMdx(self.h, self.own_ptr(u), self.own_ptr(v), self.own_ptr(out), float(scale)))

    def lincomb(self, out, a, X, b=0.0, Y=None):
        n = self.lloc * self.N
        self._check(self.lib.elph_dev_lincomb(self.h, self.own_ptr(out), float(a), self.own_ptr(X), float(b),
                                              None if Y is None else self.own_ptr(Y), 0.0, None, n))

    def dot(self, a, b):
        self._check(self.lib.elph_dev_dot(self.h, self.own_ptr(a), self.own_ptr(b), self.lloc * self.N, self.scal.data_ptr()))
        return self.scal[:1].clone()


class ShardedOperator:
    """The fermion matrix of one tau-sharded lattice: products and plain CG on M^T M."""

    def __init__(self, backend, comm: RingComm, tol: float = 1e-5, maxiter: int = 10000, kappa_max: float = 1e12):
        self.be, self.comm = backend, comm
        self.lloc = backend.lloc
        self.tol, self.maxiter, self.kappa_max = tol, maxiter, kappa_max
        self.halo_exchanges = 0

    def update_model(self):
        """update_model! on the slab, then refresh the right expnV halo (D(b) is needed to recompute (M v)(b))."""
        self.be.update_model()
        self.comm.exchange(self.be.D_tensor(), self.lloc, lo=False, hi=True)

    def _mul(self, mode, y, v):
        need_lo = mode in (M_MODE, MTM_MODE)
        need_hi = mode in (MT_MODE, MTM_MODE)
        self.comm.exchange(v, self.lloc, lo=need_lo, hi=need_hi)
        self.halo_exchanges += 1
        self.be.matvec(mode, v, y)

    def mulM(self, y, v):
        self._mul(M_MODE, y, v)

    def mulMT(self, y, v):
        self._mul(MT_MODE, y, v)

    def mulMTM(self, y, v):
        self._mul(MTM_MODE, y, v)

    def gdot(self, a, b) -> float:
        return float(self.comm.allreduce_sum(self.be.dot(a, b)).item())

    def solve_cg(self, x, b, tol: float = 0.0, maxiter: int = 0):
        """Plain CG, src/IterativeSolvers.jl:239-314, with the reference stop rule.  Returns (iters, eps)."""
        be = self.be
        tol = tol or self.tol
        maxiter = maxiter or self.maxiter
        r, p, z = be.empty(), be.empty(), be.empty()
        normb = math.sqrt(self.gdot(b, b))
        self.mulMTM(r, x)
        be.lincomb(r, 1.0, b, -1.0, r)
        be.lincomb(p, 1.0, r)
        rdotr = self.gdot(r, r)
        eps0 = math.sqrt(rdotr) / normb
        eps, kmin = eps0, 0.0
        for j in range(1, maxiter + 1):
            self.mulMTM(z, p)
            alpha = rdotr / self.gdot(p, z)
            be.lincomb(x, 1.0, x, alpha, p)
            be.lincomb(r, 1.0, r, -alpha, z)
            nr = self.gdot(r, r)
            eps = math.sqrt(nr) / normb
            with np.errstate(all="ignore"):
                k = float((2.0 * j / np.log(2.0 * eps0 / eps)) ** 2)
            if k > kmin:
                kmin = k
            if eps < tol or kmin > self.kappa_max:
                return j, eps
            beta = nr / rdotr
            rdotr = nr
            be.lincomb(p, 1.0, r, beta, p)
        return maxiter, eps
