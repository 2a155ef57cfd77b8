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
        self.peer_halo = None        # set by ShardedOperator.enable_p2p(): halo exchange through peer memory instead of NCCL

    def exchange(self, v, lloc: int, lo: bool = True, hi: bool = True, rows_of_sites: bool = True):
        """Fill v[0] with the left neighbour's last own slice and v[lloc+1] with the right neighbour's first own slice.
        ``rows_of_sites`` = False for tables whose rows are not Nsites doubles (they do not fit the peer-memory halo rows)."""
        import torch.distributed as dist
        if self.peer_halo is not None and rows_of_sites:   # peer-memory push inside one kernel (csrc/cg_pipe.cu), both directions
            self.peer_halo(v)
            return
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

    def all_to_all(self, recv, send, recv_counts, send_counts):
        """All-to-all of 1-D buffers with per-peer element counts (the tau <-> site transposes of the tau-FFT)."""
        import torch.distributed as dist
        if self.world == 1:
            recv.copy_(send)
            return
        dist.all_to_all_single(recv, send, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                               group=self.group)


class CudaSlabBackend:
    """Local slab arithmetic on the GPU through libelph_b200.so (device-pointer entry points)."""

    def __init__(self, model, tau0: int, Lglob: int):
        import torch
        self.torch = torch
        self.model = model                     # HolsteinModel / SSHModel created with Ltau = Lloc
        self.lib = model._lib
        self.h = model.handle
        self.N, self.lloc = model.Nsites, model.Ltau
        self.is_ssh = getattr(model, "kind", 0) == 1
        self.Nph = model.Nph if self.is_ssh else self.N     # phonon fields per time slice
        self._check(self.lib.elph_set_shard(self.h, tau0, Lglob))
        model.set_stream(torch.cuda.current_stream().cuda_stream)
        self.scal = torch.zeros(2, dtype=torch.float64, device="cuda")
        p = C.c_void_p()
        if self.is_ssh:
            # the table that couples to the neighbour slab is (cosh, sinh)[tau][column]: (Lloc+2) rows of 2*Ncolumns doubles
            self._check(self.lib.elph_dev_ptr_cosh_sinh(self.h, C.byref(p)))
            self.ncs = 2 * model.Nbonds
            self._D_ptr = p.value - self.ncs * 8
        else:
            self._check(self.lib.elph_dev_ptr_expnV(self.h, C.byref(p)))
            # view of the handle's halo'd expnV allocation as a (Lloc+2, N) tensor for the halo exchange
            self.ncs = self.N
            self._D_ptr = p.value - self.N * 8

    def _check(self, st):
        if st != 0:
            raise RuntimeError(self.lib.elph_last_error(self.h).decode())

    def empty(self):
        return self.torch.zeros(self.lloc + 2, self.N, dtype=self.torch.float64, device="cuda")

    def empty_field(self):
        """A halo'd slab of phonon-field shape (Lloc + 2, Nph): forces, noise, the field itself (Nph = Nsites for Holstein)."""
        return self.torch.zeros(self.lloc + 2, self.Nph, dtype=self.torch.float64, device="cuda")

    def own_ptr(self, v):
        return v.data_ptr() + v.shape[1] * 8

    def D_tensor(self):
        """The handle's time-dependent table with halos -- expnV (Holstein, rows of Nsites) or (cosh, sinh) (SSH, rows of
        2*Ncolumns) -- wrapped as a CUDA tensor (no copy)."""
        torch = self.torch
        n = (self.lloc + 2) * self.ncs
        iface = {"shape": (n,), "typestr": "<f8", "data": (self._D_ptr, False), "version": 3}

        class _W:
            __cuda_array_interface__ = iface
        return torch.as_tensor(_W(), device="cuda").view(self.lloc + 2, self.ncs)

    def update_model(self):
        self._check(self.lib.elph_dev_update_model(self.h))

    def matvec(self, mode, v, y):
        self._check(self.lib.elph_dev_shard_matvec(self.h, mode, self.own_ptr(v), self.own_ptr(y)))

    def matvec_halo(self, mode, v, y):
        self._check(self.lib.elph_dev_shard_matvec_halo(self.h, mode, self.own_ptr(v), self.own_ptr(y)))

    def muldMdx(self, u, v, out, scale=1.0):
        self._check(self.lib.elph_dev_shard_muldMdx(self.h, self.own_ptr(u), self.own_ptr(v), self.own_ptr(out), float(scale)))

    def lincomb(self, out, a, X, b=0.0, Y=None):
        n = self.lloc * out.shape[1]              # site vectors (Nsites per slice) and phonon fields (Nph per slice) alike
        self._check(self.lib.elph_dev_lincomb(self.h, self.own_ptr(out), float(a), self.own_ptr(X), float(b),
                                              None if Y is None else self.own_ptr(Y), 0.0, None, n))

    def dot(self, a, b):
        self._check(self.lib.elph_dev_dot(self.h, self.own_ptr(a), self.own_ptr(b), self.lloc * self.N, self.scal.data_ptr()))
        return self.scal[:1].clone()

    # ---- peer-memory CG (csrc/cg_p2p.cu): arenas shared through CUDA IPC, collectives inside the kernel --------------
    def p2p_setup(self, comm: "RingComm") -> bool:
        """Export this rank's arena, all-gather the IPC handles and slab lengths over torch.distributed, open the peers'.
        Returns True when the arenas are open on EVERY rank (halo exchange through peer memory); ``_p2p_ready`` tells
        whether the persistent CG kernels apply as well (all slices of every slab co-resident)."""
        import torch.distributed as dist
        buf = (C.c_ubyte * 64)()
        ok = True
        try:
            self._check(self.lib.elph_shard_p2p_export(self.h, comm.rank, comm.world, buf))
        except RuntimeError:
            ok = False
        mine = (bytes(buf), int(self.lloc), ok)
        if comm.world > 1:
            gathered = [None] * comm.world
            dist.all_gather_object(gathered, mine, group=comm.group)
        else:
            gathered = [mine]
        if all(g[2] for g in gathered):
            handles = (C.c_ubyte * (64 * comm.world)).from_buffer_copy(b"".join(g[0] for g in gathered))
            lengths = (C.c_int64 * comm.world)(*[g[1] for g in gathered])
            try:
                self._check(self.lib.elph_shard_p2p_open(self.h, handles, lengths))
            except RuntimeError:
                ok = False
        else:
            ok = False
        cg_ok = False
        if ok:
            avail = C.c_int32()
            self._check(self.lib.elph_shard_cg_available(self.h, C.byref(avail)))
            cg_ok = bool(avail.value)
        if comm.world > 1:                       # every rank must take the same path
            flags = [None] * comm.world
            dist.all_gather_object(flags, (ok, cg_ok), group=comm.group)
            ok = all(f[0] for f in flags)
            cg_ok = all(f[1] for f in flags)
        self._p2p_open = ok                      # arenas mapped: halo exchange through peer memory
        self._p2p_ready = ok and cg_ok           # ... and every slab is co-resident: the persistent CG kernels apply
        return ok

    def halo_p2p(self, v):
        """Halo slices of a (Lloc+2, N) slab tensor through peer memory: one kernel, no NCCL call (elph_dev_shard_halo)."""
        self._check(self.lib.elph_dev_shard_halo(self.h, self.own_ptr(v)))

    def cg_p2p(self, x, b, tol: float = 0.0, maxiter: int = 0):
        """Whole CG solve (x0 = 0) in one persistent kernel per GPU; returns (iters, eps).  x, b: halo'd slab tensors."""
        it, eps = C.c_int64(), C.c_double()
        self._check(self.lib.elph_dev_shard_cg_p2p(self.h, self.own_ptr(b), self.own_ptr(x), float(tol), int(maxiter),
                                                   C.byref(it), C.byref(eps)))
        return int(it.value), float(eps.value)

    # ---- pieces needed by the sharded Langevin step -------------------------------------------------------------
    def _wrap(self, ptr, shape):
        torch = self.torch
        n = int(np.prod(shape))
        iface = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

        class _W:
            __cuda_array_interface__ = iface
        return torch.as_tensor(_W(), device="cuda").view(*shape)

    def x_tensor(self):
        """The handle's phonon field slab [Lloc][N] wrapped as a CUDA tensor (no copy)."""
        p = C.c_void_p()
        self._check(self.lib.elph_dev_ptr_x(self.h, C.byref(p)))
        return self._wrap(p.value, (self.lloc, self.Nph))

    def dSbdx(self, dS, xh, shifted=True):
        self._check(self.lib.elph_dev_shard_dSbdx(self.h, self.own_ptr(dS), self.own_ptr(xh), 1 if shifted else 0))

    def make_fft_plan(self, Lglob: int):
        """A 1-site handle whose only job is to own the tau-FFT plan of the GLOBAL time extent."""
        import elphdynamics_b200 as E
        aux = E.HolsteinModel(E.Lattice(E.UnitCell(1, 1), 1), Lglob * self.model.dtau, self.model.dtau)
        assert aux.Ltau == Lglob
        aux.initialize_model_()
        aux.set_stream(self.torch.cuda.current_stream().cuda_stream)
        self._fft_aux = aux

    def fa_cols(self, vin, vout, diag, power):
        """fourier_accelerate! on a [L][ncols] block (all time slices of a subset of the sites)."""
        aux = self._fft_aux
        st = aux._lib.elph_dev_fourier_accelerate_cols(aux.handle, vin.data_ptr(), vout.data_ptr(), vin.shape[1],
                                                       diag.data_ptr(), float(power))
        if st != 0:
            raise RuntimeError(aux._lib.elph_last_error(aux.handle).decode())


    # ---- KPM preconditioner of the sharded lattice (ShardedKPM): site-sharded FFT stage, omega-sharded chain stage ------
    def kpm_init(self, aux_model, n: int = 20, buf: float = 0.05, c1: float = 1.0, c2: float = 1.0):
        """``aux_model``: a model of the same kind for the GLOBAL lattice (all sites, global Ltau, same hoppings / couplings) -- it
        owns the FFT plan of the global time extent, the polynomial coefficients and the chain kernels; its own field is never
        used."""
        from .models import SymmetricKPMPreconditioner, update_model_
        assert aux_model.Nsites == self.N
        aux_model.set_stream(self.torch.cuda.current_stream().cuda_stream)
        update_model_(aux_model)               # SSH: expmu of the handle is what the set-up copies (it has no time index)
        self._kpm_aux = aux_model
        self._kpm_P = SymmetricKPMPreconditioner(aux_model, n, buf, c1, c2)
        self.kpm_L = aux_model.Ltau

    def kpm_set_subset(self, first: int, stride: int):
        self._kpm_aux._call("elph_kpm_set_omega_subset", int(first), int(stride))

    def kpm_setup_bar(self, eVbar, noise):
        """setup!(P) (src/KPMPreconditioners.jl:269-321) from the all-reduced tau-mean; returns (active, recomputed)."""
        P = self._kpm_P
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        if noise.size != 2 * self.N:
            raise ValueError("arnoldi_noise must hold 2*Nsites values")
        self._kpm_aux._call("elph_dev_kpm_setup_bar", eVbar.data_ptr(), noise.ctypes.data_as(C.POINTER(C.c_double)),
                            C.byref(P.info))
        return bool(P.info.active), bool(P.info.recomputed)

    def kpm_orders(self):
        return self._kpm_P.orders()

    def kpm_window(self):
        """(lambda_lo, lambda_hi, e_min, e_max) of the last set-up."""
        i = self._kpm_P.info
        return i.lambda_lo, i.lambda_hi, i.e_min, i.e_max

    def tau_to_omega_cols(self, cols):
        """[tau][col] real -> [omega][col] complex (tau_to_omega!, src/TimeFreqFFTs.jl:31-75)."""
        nu = self.torch.empty(cols.shape, dtype=self.torch.complex128, device=cols.device)
        self._kpm_aux._call("elph_dev_tau_to_omega_cols", cols.data_ptr(), nu.data_ptr(), cols.shape[1])
        return nu

    def omega_to_tau_cols(self, nu):
        """[omega][col] complex -> [tau][col] real (omega_to_tau!, :78-122)."""
        out = self.torch.empty(nu.shape, dtype=self.torch.float64, device=nu.device)
        self._kpm_aux._call("elph_dev_omega_to_tau_cols", nu.data_ptr(), out.data_ptr(), nu.shape[1])
        return out

    def kpm_chains(self, nu_in, nu_out):
        """The Chebyshev recurrences of this rank's frequencies on [L][N] complex buffers indexed by the global frequency."""
        self._kpm_aux._call("elph_dev_kpm_chains", nu_in.data_ptr(), nu_out.data_ptr())


    def kpm_shard_setup(self, comm: "RingComm", tau0: int) -> bool:
        """Arenas of the fused application (csrc/kpm_shard.cu): export this rank's, all-gather the IPC handles and the slab
        starts over torch.distributed, open the peers'.  True when every rank succeeded."""
        import torch.distributed as dist
        aux = self._kpm_aux
        buf = (C.c_ubyte * 64)()
        ok = True
        try:
            aux._call("elph_kpm_shard_export", comm.rank, comm.world, int(tau0), int(self.lloc), buf)
        except Exception:
            ok = False
        mine = (bytes(buf), int(tau0), ok)
        if comm.world > 1:
            gathered = [None] * comm.world
            dist.all_gather_object(gathered, mine, group=comm.group)
        else:
            gathered = [mine]
        if all(g[2] for g in gathered):
            handles = (C.c_ubyte * (64 * comm.world)).from_buffer_copy(b"".join(g[0] for g in gathered))
            starts = (C.c_int64 * comm.world)(*[g[1] for g in gathered])
            try:
                aux._call("elph_kpm_shard_open", handles, starts)
            except Exception:
                ok = False
        else:
            ok = False
        if comm.world > 1:                       # every rank must take the same path
            flags = [None] * comm.world
            dist.all_gather_object(flags, ok, group=comm.group)
            ok = all(flags)
        return ok

    def kpm_shard_apply(self, r, z):
        self._kpm_aux._call("elph_dev_kpm_shard_apply", self.own_ptr(r), self.own_ptr(z))

    def kpm_shard_check(self):
        self._kpm_aux._call("elph_kpm_shard_check")


class TauSiteTranspose:
    """The all-to-all pair around every tau-FFT: tau-sharded [Lloc][N] <-> site-sharded [L][Nloc]."""

    def __init__(self, comm: RingComm, N: int, Lglob: int, lloc: int):
        w, r = comm.world, comm.rank
        self.comm, self.N, self.L, self.lloc = comm, N, Lglob, lloc
        self.site_spans = [slab_bounds(N, w, q) for q in range(w)]
        self.tau_spans = [slab_bounds(Lglob, w, q) for q in range(w)]
        self.s0, self.nloc = self.site_spans[r]
        self.to_cols_counts = ([lloc * n for (_, n) in self.site_spans], [lt * self.nloc for (_, lt) in self.tau_spans])

    def to_cols(self, own):
        """own: [Lloc][N] (the own slices of a slab vector) -> [L][Nloc]: all time slices of this rank's site block."""
        import torch
        send = torch.cat([own[:, s:s + n].reshape(-1) for (s, n) in self.site_spans])
        send_counts, recv_counts = self.to_cols_counts
        recv = torch.empty(self.L * self.nloc, dtype=own.dtype, device=own.device)
        self.comm.all_to_all(recv, send, recv_counts, send_counts)
        return recv.view(self.L, self.nloc)             # chunks arrive in rank = tau order: [tau][site_local]

    def to_slab(self, cols, out_own):
        """cols: [L][Nloc] -> out_own [Lloc][N] (written in place)."""
        import torch
        send_counts, recv_counts = self.to_cols_counts
        back = torch.empty(self.lloc * self.N, dtype=cols.dtype, device=cols.device)
        self.comm.all_to_all(back, cols.contiguous().reshape(-1), send_counts, recv_counts)
        off = 0
        for (s, n) in self.site_spans:
            out_own[:, s:s + n] = back[off:off + self.lloc * n].view(self.lloc, n)
            off += self.lloc * n


class ShardedKPM:
    """``SymmetricKPMPreconditioner`` (src/KPMPreconditioners.jl:219-481) of a tau-sharded lattice (Holstein or SSH).

    setup!: the tau-mean of expnV (update_A!, :332-350; SSH: of the per-bond (cosh, sinh) pairs, :355-381) is a local sum over the
    slab's rows of the time-dependent table + one all-reduce; the
    Arnoldi bounds, the hysteresis and the coefficients (:269-321, :781-942) act on Nsites-vectors and are replicated on every
    rank from the same injected start vectors, so every rank holds identical polynomials.
    ldiv! (:426-481) runs on three shardings with an all-to-all between them:
        tau-sharded r  --all-to-all-->  site-sharded: twisted tau-FFT of all slices of Nsites/P sites
                       --all-to-all-->  omega-sharded: the Chebyshev recurrences (:606-679) of the frequencies w = rank, rank + P,
                                        ... (the polynomial order falls monotonically with w, so dealing the frequencies round
                                        robin is the longest-first assignment) on ALL sites
                       --all-to-all-->  site-sharded: mirror frequencies L-1-w = conj (:464-466), inverse FFT
                       --all-to-all-->  tau-sharded z.
    Only the cld(L,2) independent frequencies travel; the receiver rebuilds the mirrors.
    """
    is_identity = False

    def __init__(self, op: "ShardedOperator", N: int, Lglob: int, transpose: TauSiteTranspose | None = None):
        import torch
        self.op, self.be, self.comm = op, op.be, op.comm
        self.N, self.L, self.lloc = N, Lglob, op.lloc
        self.Lo2 = (Lglob + 1) // 2
        self.tr = transpose or TauSiteTranspose(self.comm, N, Lglob, self.lloc)
        w, r = self.comm.world, self.comm.rank
        self.my_w = list(range(r, self.Lo2, w))
        self.nw = [len(range(q, self.Lo2, w)) for q in range(w)]
        self.be.kpm_set_subset(r, w)
        self.active = False
        self.recomputed = False
        dev = self.be.empty().device
        self.nu_in = torch.zeros(Lglob, N, dtype=torch.complex128, device=dev)
        self.nu_out = torch.zeros(Lglob, N, dtype=torch.complex128, device=dev)
        self.applies = 0
        self.fused = False

    def enable_fused(self, tau0: int) -> bool:
        """Use the one-call application with the transposes through peer memory (csrc/kpm_shard.cu) where the backend has it and
        every rank could open every arena; otherwise the all-to-all form stays.  ``tau0``: first global slice of this slab."""
        setup = getattr(self.be, "kpm_shard_setup", None)
        self.fused = bool(setup and setup(self.comm, tau0))
        return self.fused

    def setup(self, arnoldi_noise):
        """setup!(P) with the 2*Nsites Arnoldi start values injected (the same array on every rank)."""
        D = self.be.D_tensor()[1:self.lloc + 1]
        acc = D.sum(dim=0)
        self.comm.allreduce_sum(acc)
        acc /= float(self.L)
        self.active, self.recomputed = self.be.kpm_setup_bar(acc, arnoldi_noise)

    def ldiv(self, z, r):
        """z = P^-1 r on halo'd slab tensors (own slices only)."""
        import torch
        lloc, Lo2, L = self.lloc, self.Lo2, self.L
        self.applies += 1
        if not self.active:                      # identity (:475-478)
            z[1:lloc + 1] = r[1:lloc + 1]
            return
        if self.fused:
            self.be.kpm_shard_apply(r, z)
            return
        tr, comm = self.tr, self.comm
        w, me = comm.world, comm.rank
        nloc = tr.nloc
        cols = tr.to_cols(r[1:lloc + 1])
        nu = self.be.tau_to_omega_cols(cols)                         # [L][nloc]; rows < Lo2 are the independent frequencies
        # site-sharded -> omega-sharded
        send = torch.cat([torch.view_as_real(nu[q:Lo2:w]).reshape(-1) for q in range(w)])
        send_counts = [2 * self.nw[q] * nloc for q in range(w)]
        recv_counts = [2 * self.nw[me] * n for (_, n) in tr.site_spans]
        recv = torch.empty(sum(recv_counts), dtype=torch.float64, device=send.device)
        comm.all_to_all(recv, send, recv_counts, send_counts)
        nmine = self.nw[me]
        off = 0
        for (s, n) in tr.site_spans:
            blk = torch.view_as_complex(recv[off:off + 2 * nmine * n].view(nmine, n, 2))
            self.nu_in[me:Lo2:w, s:s + n] = blk
            off += 2 * nmine * n
        self.be.kpm_chains(self.nu_in, self.nu_out)
        # omega-sharded -> site-sharded (independent frequencies only)
        own_rows = self.nu_out[me:Lo2:w]
        send2 = torch.cat([torch.view_as_real(own_rows[:, s:s + n].contiguous()).reshape(-1) for (s, n) in tr.site_spans])
        recv2 = torch.empty(sum(send_counts), dtype=torch.float64, device=send.device)
        comm.all_to_all(recv2, send2, send_counts, recv_counts)
        nu2 = torch.empty(L, nloc, dtype=torch.complex128, device=send.device)
        off = 0
        blocks = []
        for q in range(w):
            blk = torch.view_as_complex(recv2[off:off + 2 * self.nw[q] * nloc].view(self.nw[q], nloc, 2))
            nu2[q:Lo2:w] = blk
            blocks.append(blk)
            off += 2 * self.nw[q] * nloc
        for q in range(w):                                           # mirrors after ALL direct rows: for odd L the middle
            if self.nw[q]:                                           # frequency ends up conjugated in place, as in :464-466
                idx = torch.arange(q, Lo2, w, device=send.device)
                nu2[L - 1 - idx] = torch.conj(blocks[q])
        zc = self.be.omega_to_tau_cols(nu2)
        tr.to_slab(zc, z[1:lloc + 1])


class ShardedOperator:
    """The fermion matrix of one tau-sharded lattice: products, plain and preconditioned CG on M^T M."""

    def __init__(self, backend, comm: RingComm, tol: float = 1e-5, maxiter: int = 10000, kappa_max: float = 1e12):
        self.be, self.comm = backend, comm
        self.lloc = backend.lloc
        self.tol, self.maxiter, self.kappa_max = tol, maxiter, kappa_max
        self.halo_exchanges = 0

    def update_model(self):
        """update_model! on the slab, then refresh the right expnV halo (D(b) is needed to recompute (M v)(b))."""
        self.be.update_model()
        D = self.be.D_tensor()
        self.comm.exchange(D, self.lloc, lo=False, hi=True, rows_of_sites=(D.shape[1] == getattr(self.be, "N", D.shape[1])))

    def _mul(self, mode, y, v):
        if self.comm.peer_halo is not None and hasattr(self.be, "matvec_halo"):
            self.halo_exchanges += 1
            self.be.matvec_halo(mode, v, y)      # halo push through peer memory + product, one call
            return
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

    def enable_p2p(self) -> bool:
        """Use the peer-memory CG (one persistent kernel per GPU, collectives inside it) for solve() where it applies."""
        setup = getattr(self.be, "p2p_setup", None)
        ok = bool(setup and setup(self.comm))
        if ok and hasattr(self.be, "halo_p2p"):
            self.comm.peer_halo = self.be.halo_p2p
        return ok

    def solve(self, x, b, tol: float = 0.0, maxiter: int = 0):
        """solve!(x, A, b, cg) with x0 = 0 (what every caller on the hot path does, src/LangevinDynamics.jl:355-360):
        the peer-memory CG when enable_p2p() succeeded, else the NCCL-between-launches CG."""
        if getattr(self.be, "_p2p_ready", False):
            return self.be.cg_p2p(x, b, tol or self.tol, maxiter or self.maxiter)
        x.zero_()
        return self.solve_cg(x, b, tol, maxiter)

    def _true_residual(self, x, b):
        res = self.be.empty()
        self.mulMTM(res, x)
        self.be.lincomb(res, 1.0, b, -1.0, res)
        num, den = self.gdot(res, res), self.gdot(b, b)
        return math.sqrt(num) / math.sqrt(den) if den > 0 else float("nan")

    def ldiv(self, x, b, tol: float = 0.0, P=None):
        """``ldiv!(x, model, b)`` without a preconditioner (src/Models.jl:141-186) on the sharded lattice: solve from x0 = 0,
        then the TRUE relative residual |b - A x| / |b| with one more product.  When it exceeds sqrt(tol): ``flag`` 1 if the solve
        ran into maxiter, 2 if the solver reported a convergence that the true residual does not confirm -- and ``x`` is zeroed in
        both cases, as the reference does, so that a failed solve never feeds the force.  Returns ``(iters, residual, flag)``.
        With a preconditioner ``P`` (ShardedKPM): ``ldiv!(x, model, b, P)`` (:74-137) -- preconditioned CG, the same check, and on
        failure the unpreconditioned solve with 10 x maxiter as the fallback."""
        tol = tol or self.tol
        if P is not None and not getattr(P, "is_identity", False):
            x.zero_()
            iters, _ = self.solve_pcg(x, b, P, tol)
            residual = self._true_residual(x, b)
            if residual <= math.sqrt(tol):
                return iters, residual, 0
            iters, _ = self.solve(x, b, tol, maxiter=10 * self.maxiter)
            residual = self._true_residual(x, b)
            flag = 0
            if residual > math.sqrt(tol):
                flag = 1 if iters == self.maxiter else 2      # src/Models.jl:160 compares with solver.maxiter
                x.zero_()
            return iters, residual, flag
        iters, _ = self.solve(x, b, tol)
        residual = self._true_residual(x, b)
        flag = 0
        if residual > math.sqrt(tol):
            flag = 1 if iters == self.maxiter else 2
            x.zero_()
        return iters, residual, flag

    def solve_pcg(self, x, b, P, tol: float = 0.0, maxiter: int = 0):
        """Preconditioned CG, src/IterativeSolvers.jl:153-234, on the sharded lattice: per iteration one product (one halo
        exchange), one preconditioner application (ShardedKPM.ldiv) and two scalar round trips (p.Ap; |r|^2 and r.z together).
        Returns (iters, eps)."""
        be = self.be
        tol = tol or self.tol
        maxiter = maxiter or self.maxiter
        r, p, z = be.empty(), be.empty(), be.empty()
        normb = math.sqrt(self.gdot(b, b))
        self.mulMTM(r, x)
        be.lincomb(r, 1.0, b, -1.0, r)
        P.ldiv(z, r)
        be.lincomb(p, 1.0, z)
        rdotz = self.gdot(r, z)
        eps0 = math.sqrt(self.gdot(r, r)) / normb
        eps, kmin = eps0, 0.0
        import torch
        for j in range(1, maxiter + 1):
            self.mulMTM(z, p)
            alpha = rdotz / self.gdot(p, z)
            be.lincomb(x, 1.0, x, alpha, p)
            be.lincomb(r, 1.0, r, -alpha, z)
            # |r|^2 decides the stop rule and r.z the next direction.  The application of the preconditioner is queued BEFORE |r|^2
            # is read back, so the round trip of the scalars (all-reduce + device-to-host) hides behind it and both travel together:
            # two host synchronisations per iteration instead of three.  Same values, same decisions as :205-227; the one
            # application after the last iteration is computed and dropped.
            rr_t = be.dot(r, r)
            P.ldiv(z, r)
            rz_t = be.dot(r, z)
            both = torch.cat([rr_t.reshape(1), rz_t.reshape(1)])
            self.comm.allreduce_sum(both)
            rr, new_rdotz = (float(v) for v in both.tolist())
            eps = math.sqrt(rr) / normb
            with np.errstate(all="ignore"):
                k = float((2.0 * j / np.log(2.0 * eps0 / eps)) ** 2)
            if k > kmin:
                kmin = k
            if eps < tol or kmin > self.kappa_max:
                return j, eps
            beta = new_rdotz / rdotz
            rdotz = new_rdotz
            be.lincomb(p, 1.0, z, beta, p)
        return maxiter, eps

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


class ShardedLangevin:
    """Langevin updates of a tau-sharded lattice -- Holstein (plain CG, or KPM-preconditioned CG with a ShardedKPM) or SSH (plain
    CG; phonon fields on the bonds, every field its own primary field: src/SSHModels.jl:567-576 is then the identity) --, reference:
    src/LangevinDynamics.jl:81-119 (Euler), :162-225 (Runge-Kutta), :334-384 (forces).

    Collectives per force evaluation: one halo exchange per product (CG iterations + M^T g + the force's v(tau-1)),
    two scalar all-reduces per CG iteration, one x halo for the bosonic gradient; per Fourier acceleration: an
    all-to-all pair (tau-sharded -> site-sharded, local tau-FFT of all slices of Nsites/P sites, and back).
    """

    def __init__(self, op: ShardedOperator, N: int, Lglob: int, tau0: int, Q_site_block, dt: float, P=None):
        """``Q_site_block``: the acceleration diagonal of this rank's site block, [k][site_local] (Lglob x Nloc).
        ``P``: a ShardedKPM (the solves of the force become ``ldiv!(x, model, b, P)``) or None."""
        self.op, self.be, self.comm = op, op.be, op.comm
        self.N, self.L, self.tau0, self.lloc = N, Lglob, tau0, op.lloc
        self.dt = float(dt)
        # the Fourier acceleration acts on phonon fields: Nph columns per time slice (= Nsites for Holstein, the bonds for SSH)
        self.Nph = int(getattr(self.be, "Nph", N))
        self.empty_field = getattr(self.be, "empty_field", self.be.empty)
        self.tr = P.tr if (P is not None and self.Nph == N) else TauSiteTranspose(self.comm, self.Nph, Lglob, self.lloc)
        self.site_spans, self.tau_spans = self.tr.site_spans, self.tr.tau_spans
        self.s0, self.nloc = self.tr.s0, self.tr.nloc
        self.P = P
        self.Q = Q_site_block
        self.xh = self.empty_field()                   # halo'd master copy of the phonon field slab
        self.last_iters, self.last_residual, self.last_flag = 0, 0.0, 0

    # ---- Fourier acceleration through the all-to-all transposes ---------------------------------------------------
    def fourier_accelerate(self, v, power):
        """v: halo'd tensor; returns a new halo'd tensor with Re iFFT(Q^power FFT v) on the own slices."""
        torch = self.be.torch
        cols = self.tr.to_cols(v[1:self.lloc + 1])
        out_cols = torch.empty_like(cols)
        self.be.fa_cols(cols, out_cols, self.Q, power)
        out = self.empty_field()
        self.tr.to_slab(out_cols, out[1:self.lloc + 1])
        return out

    # ---- field / forces -------------------------------------------------------------------------------------------
    def set_x(self, x_slab):
        """x_slab: (Lloc, N) tensor/array in the engine layout."""
        self.xh[1:self.lloc + 1] = self.be.torch.as_tensor(x_slab, dtype=self.xh.dtype).to(self.xh.device)
        self._push_x()

    def _push_x(self):
        self.be.x_tensor().copy_(self.xh[1:self.lloc + 1])
        self.op.update_model()

    def calc_dSdx(self, g, arnoldi_noise=None):
        """dS/dx = -2 g^T (dM/dx) M^-1 g + dSb/dx (shifted), src/LangevinDynamics.jl:334-384; with a preconditioner, setup!(P)
        first (:353) from the injected Arnoldi start values."""
        be, op = self.be, self.op
        b, x, dS = be.empty(), be.empty(), self.empty_field()
        if self.P is not None:
            self.P.setup(arnoldi_noise)
        op.mulMT(b, g)
        iters, self.last_residual, self.last_flag = op.ldiv(x, b, P=self.P)
        self.last_iters = iters
        self.comm.exchange(x, self.lloc, lo=True, hi=False)     # the force needs (M^-1 g)(tau-1)
        be.muldMdx(g, x, dS, -2.0)
        self.comm.exchange(self.xh, self.lloc)                  # bosonic gradient couples x(tau +- 1), periodic
        be.dSbdx(dS, self.xh, True)
        return dS

    def evolve_euler(self, eta, g, arnoldi_noise=None):
        """src/LangevinDynamics.jl:81-119; eta, g: halo'd tensors holding this rank's slab of the injected noise."""
        be = self.be
        self._push_x()
        dS = self.calc_dSdx(g, arnoldi_noise)
        QdS = self.fourier_accelerate(dS, 1.0)
        sqQeta = self.fourier_accelerate(eta, 0.5)
        dx = self.empty_field()
        be.lincomb(dx, math.sqrt(2.0 * self.dt), sqQeta, -self.dt, QdS)
        be.lincomb(self.xh, 1.0, self.xh, 1.0, dx)
        self._push_x()
        return self.last_iters

    def evolve_rk(self, eta, g1, g2, arnoldi_noise1=None, arnoldi_noise2=None):
        """src/LangevinDynamics.jl:162-225."""
        be = self.be
        self._push_x()
        dS1 = self.calc_dSdx(g1, arnoldi_noise1)
        dx = self.empty_field()
        be.lincomb(dx, math.sqrt(2.0 * self.dt), eta, -self.dt, dS1)
        be.lincomb(self.xh, 1.0, self.xh, 1.0, dx)
        self._push_x()
        dS2 = self.calc_dSdx(g2, arnoldi_noise2)
        be.lincomb(self.xh, 1.0, self.xh, -1.0, dx)
        self._push_x()
        be.lincomb(dS1, 0.5, dS2, 0.5, dS1)
        QdS = self.fourier_accelerate(dS1, 1.0)
        sqQeta = self.fourier_accelerate(eta, 0.5)
        be.lincomb(dx, math.sqrt(2.0 * self.dt), sqQeta, -self.dt, QdS)
        be.lincomb(self.xh, 1.0, self.xh, 1.0, dx)
        self._push_x()
        return self.last_iters

    def evolve_heun(self, eta, g1, g2, arnoldi_noise1=None, arnoldi_noise2=None):
        """src/LangevinDynamics.jl:272-324 (Heun): xi = sqrt(Q) eta once, both force evaluations Fourier accelerated; returns
        div(iters1 + iters2, 2) as the reference does (:324)."""
        be = self.be
        s2dt = math.sqrt(2.0 * self.dt)
        xi = self.fourier_accelerate(eta, 0.5)                    # :293
        self._push_x()                                            # :296
        dS1 = self.calc_dSdx(g1, arnoldi_noise1)                  # :298
        it1 = self.last_iters
        dG1 = self.fourier_accelerate(dS1, 1.0)                   # :301
        dx = self.empty_field()
        be.lincomb(dx, s2dt, xi, -self.dt, dG1)                   # :304
        be.lincomb(self.xh, 1.0, self.xh, 1.0, dx)                # :307
        self._push_x()                                            # :308
        dS2 = self.calc_dSdx(g2, arnoldi_noise2)                  # :312
        it2 = self.last_iters
        dG2 = self.fourier_accelerate(dS2, 1.0)                   # :315
        be.lincomb(self.xh, 1.0, self.xh, -1.0, dx)               # :318
        # x'' = x + sqrt(2 dt) xi - dt (dG + dG') / 2             # :321
        be.lincomb(dx, 0.5, dG1, 0.5, dG2)
        be.lincomb(dx, s2dt, xi, -self.dt, dx)
        be.lincomb(self.xh, 1.0, self.xh, 1.0, dx)
        self._push_x()                                            # :322
        return (it1 + it2) // 2

