"""tau-sharded products and CG.

CPU: the host-side logic (slab bounds, ring halo exchange incl. the antiperiodic closure, scalar all-reduces, CG
control flow) runs over gloo with world_size 2 and 3, with a NumPy slab backend built from the oracle.
GPU: the open-slab CUDA kernels are checked (a) on one GPU with several slabs driven in-process and (b) with real
NCCL ranks when at least two GPUs are visible.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import oracle_holstein, relerr  # noqa: E402
from oracle import checkerboard as cb  # noqa: E402
from oracle.solvers import ConjugateGradient, solve_cg  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _OracleKpmMixin:
    """KPM pieces of the NumPy slab backends (oracle/kpm.py per frequency, numpy.fft per column)."""

    def kpm_init(self, aux_model=None, n=20, buf=0.05, c1=1.0, c2=1.0):
        from oracle.kpm import KPMPreconditioner
        P = KPMPreconditioner(self.om, n, buf, c1, c2)
        P.update_A = lambda: None                    # the tau-mean comes from the sharded driver (local sums + all-reduce)
        self._P, self.kpm_L, self._sub = P, self.om.L, (0, 1)

    def kpm_set_subset(self, first, stride):
        self._sub = (first, stride)

    def kpm_setup_bar(self, bar, noise):
        if self.om.kind == "ssh":                    # tau-mean of the (cosh, sinh) pairs, [Ncolumns][2]
            self._P.coshbar[:] = bar.numpy()[0::2]
            self._P.sinhbar[:] = bar.numpy()[1::2]
            self._P.expnVbar[:] = self.om.expmu
        else:
            self._P.expnVbar[:] = bar.numpy()
        self._P.setup(noise)
        return self._P.active, self._P.recomputed

    def kpm_orders(self):
        return self._P.order.copy()

    def kpm_window(self):
        return self._P.lam_lo, self._P.lam_hi, self._P.e_min, self._P.e_max

    def tau_to_omega_cols(self, cols):
        L = self.Lg
        theta = np.exp(-1j * np.pi * np.arange(L) / L)
        return self.torch.from_numpy(np.fft.fft(theta[:, None] * cols.numpy(), axis=0))

    def omega_to_tau_cols(self, nu):
        L = self.Lg
        theta = np.exp(-1j * np.pi * np.arange(L) / L)
        return self.torch.from_numpy(np.real(np.conj(theta)[:, None] * np.fft.ifft(nu.numpy(), axis=0)))

    def kpm_chains(self, nu_in, nu_out):
        L, (first, stride) = self.Lg, self._sub
        a, o = nu_in.numpy(), nu_out.numpy()
        for w in range(first, self._P.Lo2, stride):
            o[w] = self._P.mul_block(w, a[w])
            o[L - 1 - w] = np.conj(o[w])


class OracleSlabBackend(_OracleKpmMixin):
    """NumPy slab arithmetic from the oracle's checkerboard sweeps (TEST ONLY; the package has no CPU backend)."""

    def __init__(self, om, tau0, lloc):
        import torch
        self.torch = torch
        self.om, self.tau0, self.lloc, self.N, self.Lg = om, tau0, lloc, om.N, om.L
        E = om.expnV.reshape(om.N, om.L).T          # [tau][site]
        idx = [(tau0 + t - 1) % om.L for t in range(lloc + 2)]
        self.D = torch.from_numpy(np.ascontiguousarray(E[idx])).clone()
        self.D[0] = 0.0
        self.D[lloc + 1] = 0.0                      # the halo is filled by the exchange, like on the GPU
        X = om.x.reshape(om.N, om.L).T
        self.x = torch.from_numpy(np.ascontiguousarray(X[tau0:tau0 + lloc])).clone()

    def empty(self):
        return self.torch.zeros(self.lloc + 2, self.N, dtype=self.torch.float64)

    def D_tensor(self):
        return self.D

    def x_tensor(self):
        return self.x

    def update_model(self):
        om, L = self.om, self.lloc
        X = self.x.numpy()
        self.D.numpy()[1:L + 1] = np.exp(-om.dtau * (om.lam[None, :] * X + om.lam2[None, :] * X ** 2 + -om.mu[None, :]))

    def dSbdx(self, dS, xh, shifted=True):
        om, L, dt = self.om, self.lloc, self.om.dtau
        x = xh.numpy()
        own, up, dn = x[1:L + 1], x[2:L + 2], x[0:L]
        d = dt * om.omega[None, :] ** 2 * own - (dt * om.lam[None, :] if shifted else 0.0)
        d = d + dt * 4 * om.omega4[None, :] * own ** 3
        d = d - (up + dn - 2.0 * own) / dt
        dS.numpy()[1:L + 1] += d

    def muldMdx(self, u, v, out, scale=1.0):
        om, L, dt = self.om, self.lloc, self.om.dtau
        un, vn, Dn, X = u.numpy(), v.numpy(), self.D.numpy(), self.x.numpy()
        own = np.arange(1, L + 1)
        sign = np.where((self.tau0 + own - 1) % self.Lg == 0, -1.0, 1.0)[:, None]
        d = sign * dt * (om.lam[None, :] + 2 * om.lam2[None, :] * X) * Dn[own] * vn[own - 1]
        y = self._K(un[own], True)
        out.numpy()[1:L + 1] = scale * y * d

    def make_fft_plan(self, Lglob):
        pass

    def fa_cols(self, vin, vout, diag, power):
        a = np.fft.fft(vin.numpy().astype(np.complex128), axis=0) * diag.numpy() ** power
        vout.numpy()[:] = np.real(np.fft.ifft(a, axis=0))

    def _K(self, slices, transpose):
        Y = np.ascontiguousarray(slices.T)           # (N, nsl)
        f = cb.checkerboard_transpose_mul if transpose else cb.checkerboard_mul
        f(Y, self.om.neighbor_table, self.om.cosht, self.om.sinht, self.om.group_offsets)
        return Y.T

    def _w(self, v, D, ts):
        """w(t) = v(t) -/+ K D(t) v(t-1) for slab indices ts (1-based rows of the halo'd arrays)."""
        ts = np.asarray(ts)
        Bv = self._K(D[ts] * v[ts - 1], False)
        sign = np.where((self.tau0 + ts - 1) % self.Lg == 0, 1.0, -1.0)[:, None]
        return v[ts] + sign * Bv

    def matvec(self, mode, v, y):
        vn, Dn, L = v.numpy(), self.D.numpy(), self.lloc
        own = np.arange(1, L + 1)
        if mode == 0:
            out = self._w(vn, Dn, own)
        else:
            w = vn if mode == 1 else np.zeros_like(vn)
            if mode == 2:
                w[1:L + 2] = self._w(vn, Dn, np.arange(1, L + 2))
            u = self._K(w[own + 1], True)
            sign = np.where((self.tau0 + own) % self.Lg == 0, 1.0, -1.0)[:, None]
            out = w[own] + sign * Dn[own + 1] * u
        y.numpy()[1:L + 1] = out

    def lincomb(self, out, a, X, b=0.0, Y=None):
        L = self.lloc
        r = a * X[1:L + 1]
        if Y is not None:
            r = r + b * Y[1:L + 1]
        out[1:L + 1] = r

    def dot(self, a, b):
        L = self.lloc
        return (a[1:L + 1] * b[1:L + 1]).sum().reshape(1)


class OracleSSHSlabBackend(_OracleKpmMixin):
    """NumPy slab arithmetic for the SSH model from the oracle's sweeps with per-slice (cosh, sinh) tables (TEST ONLY): what the
    open-slab SSH kernels compute, for the gloo tests of the host logic (table halo with rows of 2*Ncolumns doubles, field-shaped
    slabs of Nph columns next to site vectors of Nsites columns)."""
    is_ssh = True

    def __init__(self, om, tau0, lloc):
        import torch
        self.torch = torch
        self.om, self.tau0, self.lloc, self.N, self.Lg, self.Nph, self.Nb = om, tau0, lloc, om.N, om.L, om.Nph, om.Nbonds
        self.CS = torch.zeros(lloc + 2, 2 * om.Nbonds, dtype=torch.float64)      # rows of (cosh, sinh) per column, interleaved
        X = om.x.reshape(om.Nph, om.L).T
        self.x = torch.from_numpy(np.ascontiguousarray(X[tau0:tau0 + lloc])).clone()
        self.update_model()

    def empty(self):
        return self.torch.zeros(self.lloc + 2, self.N, dtype=self.torch.float64)

    def empty_field(self):
        return self.torch.zeros(self.lloc + 2, self.Nph, dtype=self.torch.float64)

    def D_tensor(self):
        return self.CS

    def x_tensor(self):
        return self.x

    def update_model(self):
        """src/SSHModels.jl:510-540 on the slab's own rows; the halo rows belong to the exchange."""
        om, L = self.om, self.lloc
        X = self.x.numpy()                                   # (lloc, Nph)
        tp = np.repeat(om.t[None, :], L, axis=0)             # (lloc, bond) in the original bond order
        for ph in range(om.Nph):
            xt = X[:, ph]
            tp[:, om.phonon_to_bond[ph]] = om.t[om.phonon_to_bond[ph]] - (om.alpha[ph] * xt + np.sign(xt) * om.alpha2[ph] * xt ** 2)
        cs = self.CS.numpy()
        cols = om.checkerboard_perm                          # bond -> column
        c = np.zeros((L, om.Nbonds))
        sh = np.zeros((L, om.Nbonds))
        c[:, cols] = np.cosh(om.dtau * tp)
        sh[:, cols] = np.sinh(om.dtau * tp)
        cs[1:L + 1, 0::2] = c
        cs[1:L + 1, 1::2] = sh

    def _K(self, slices, rows, transpose):
        """K(t) or K^T(t) applied to slices[k] with the table of halo'd row rows[k]."""
        cs = self.CS.numpy()
        Y = np.ascontiguousarray(slices.T)                   # (N, nsl)
        c = np.ascontiguousarray(cs[rows, 0::2].T)           # (Nb, nsl)
        sh = np.ascontiguousarray(cs[rows, 1::2].T)
        f = cb.checkerboard_transpose_mul if transpose else cb.checkerboard_mul
        f(Y, self.om.neighbor_table, c, sh, self.om.group_offsets)
        return Y.T

    def _w(self, v, ts):
        ts = np.asarray(ts)
        Bv = self._K(self.om.expmu[None, :] * v[ts - 1], ts, False)
        sign = np.where((self.tau0 + ts - 1) % self.Lg == 0, 1.0, -1.0)[:, None]
        return v[ts] + sign * Bv

    def matvec(self, mode, v, y):
        vn, L = v.numpy(), self.lloc
        own = np.arange(1, L + 1)
        if mode == 0:
            out = self._w(vn, own)
        else:
            w = vn if mode == 1 else np.zeros_like(vn)
            if mode == 2:
                w[1:L + 2] = self._w(vn, np.arange(1, L + 2))
            u = self._K(w[own + 1], own + 1, True)
            sign = np.where((self.tau0 + own) % self.Lg == 0, 1.0, -1.0)[:, None]
            out = w[own] + sign * self.om.expmu[None, :] * u
        y.numpy()[1:L + 1] = out

    def muldMdx(self, u, v, out, scale=1.0):
        """src/SSHModels.jl:707-829 on the slab: bond-sequential recurrence, vectorised over the own slices."""
        om, L, dt = self.om, self.lloc, self.om.dtau
        un, vn, cs, X = u.numpy(), v.numpy(), self.CS.numpy(), self.x.numpy()
        own = np.arange(1, L + 1)
        b = (om.expmu[None, :] * vn[own - 1]).T.copy()       # (N, lloc)
        c = self._K(un[own], own, True).T.copy()
        acc = np.zeros((L, om.Nph))
        flip = ((self.tau0 + own - 1) % self.Lg == 0)
        for n in range(om.Nbonds):
            bond = om.inv_checkerboard_perm[n]
            ph = om.bond_to_phonon[bond]
            i, j = om.neighbor_table[0, n], om.neighbor_table[1, n]
            ch, sh = cs[own, 2 * n], cs[own, 2 * n + 1]
            bi, bj = b[i].copy(), b[j].copy()
            b[i] = ch * bi + sh * bj
            b[j] = ch * bj + sh * bi
            ci, cj = c[i].copy(), c[j].copy()
            c[i] = ch * ci - sh * cj
            c[j] = ch * cj - sh * ci
            if ph >= 0:
                dK = om.alpha[ph] + 2 * om.alpha2[ph] * X[:, ph]
                dm = c[j] * dt * dK * b[i] + (c[i] * dt * dK) * b[j]
                acc[:, ph] += np.where(flip, -dm, dm)
        out.numpy()[1:L + 1] = scale * acc

    def dSbdx(self, dS, xh, shifted=True):
        om, L, dt = self.om, self.lloc, self.om.dtau
        x = xh.numpy()
        own, up, dn = x[1:L + 1], x[2:L + 2], x[0:L]
        d = dt * om.omega[None, :] ** 2 * own + dt * 4 * om.omega4[None, :] * own ** 3 - (up + dn - 2.0 * own) / dt
        dS.numpy()[1:L + 1] += d

    def make_fft_plan(self, Lglob):
        pass

    def fa_cols(self, vin, vout, diag, power):
        a = np.fft.fft(vin.numpy().astype(np.complex128), axis=0) * diag.numpy() ** power
        vout.numpy()[:] = np.real(np.fft.ifft(a, axis=0))

    def lincomb(self, out, a, X, b=0.0, Y=None):
        L = self.lloc
        r = a * X[1:L + 1]
        if Y is not None:
            r = r + b * Y[1:L + 1]
        out[1:L + 1] = r

    def dot(self, a, b):
        L = self.lloc
        return (a[1:L + 1] * b[1:L + 1]).sum().reshape(1)


def _global_reference(om, rng):
    v = rng.normal(size=om.Ndim)
    outs = {}
    for name, fn in (("M", om.mulM), ("MT", om.mulMT), ("MTM", om.mulMTM)):
        y = np.zeros(om.Ndim)
        fn(y, v)
        outs[name] = y.reshape(om.N, om.L).T.copy()
    return v.reshape(om.N, om.L).T.copy(), outs


def _check_rank(op, be, comm_rank, tau0, lloc, V, outs, om, b_glob, x_ref, it_ref):
    import torch
    v = be.empty()
    v[1:lloc + 1] = torch.from_numpy(V[tau0:tau0 + lloc]).to(v.device)
    y = be.empty()
    op.update_model()
    for name, fn in (("M", op.mulM), ("MT", op.mulMT), ("MTM", op.mulMTM)):
        fn(y, v)
        got = y[1:lloc + 1].cpu().numpy()
        assert relerr(got, outs[name][tau0:tau0 + lloc]) <= 1e-12, (name, comm_rank)
    b = be.empty()
    b[1:lloc + 1] = torch.from_numpy(b_glob[tau0:tau0 + lloc]).to(b.device)
    x = be.empty()
    it, eps = op.solve_cg(x, b)
    assert abs(it - it_ref) <= 2, (it, it_ref)
    assert relerr(x[1:lloc + 1].cpu().numpy(), x_ref[tau0:tau0 + lloc]) <= 1e-3   # both stop at eps < 1e-5
    # solve(): the x0 = 0 entry the dynamics use; without enable_p2p() it is the loop above and ignores what x held
    x2 = be.empty()
    x2.fill_(7.0)
    it2, eps2 = op.solve(x2, b)
    assert it2 == it and relerr(x2[1:lloc + 1].cpu().numpy(), x[1:lloc + 1].cpu().numpy()) <= 1e-12
    # ldiv!: the solve + true-residual check + flags of src/Models.jl:141-186 (what the sharded force evaluation calls)
    x3 = be.empty()
    it3, res3, flag3 = op.ldiv(x3, b)
    assert it3 == it and flag3 == 0 and res3 <= 2e-5
    assert relerr(x3[1:lloc + 1].cpu().numpy(), x[1:lloc + 1].cpu().numpy()) <= 1e-12
    keep = op.maxiter
    op.maxiter = 3                                   # cut off: residual > sqrt(tol), flag 1, x zeroed on every rank
    it4, res4, flag4 = op.ldiv(x3, b)
    op.maxiter = keep
    assert it4 == 3 and flag4 == 1 and res4 > 1e-5 ** 0.5 and float(x3[1:lloc + 1].abs().max()) == 0.0


def _problem(seed=7, Ls=4, beta=1.1):
    om, rng = oracle_holstein("square", Ls, beta, 0.1, mu=-0.5, seed=seed)
    V, outs = _global_reference(om, rng)
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    x = np.zeros(om.Ndim)
    it = solve_cg(x, om, b, ConjugateGradient(om.Ndim, tol=1e-5, maxiter=5000))
    to_eng = lambda a: a.reshape(om.N, om.L).T.copy()
    return om, V, outs, to_eng(b), to_eng(x), it


def _cpu_worker(rank, world, port):
    import torch.distributed as dist
    from elphdynamics_b200.sharded import RingComm, ShardedOperator, slab_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        om, V, outs, b, x_ref, it_ref = _problem()
        tau0, lloc = slab_bounds(om.L, world, rank)
        be = OracleSlabBackend(om, tau0, lloc)
        op = ShardedOperator(be, RingComm(rank, world), tol=1e-5, maxiter=5000)
        assert not op.enable_p2p()       # a backend without peer memory keeps the collective-between-launches solver
        _check_rank(op, be, rank, tau0, lloc, V, outs, om, b, x_ref, it_ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_host_logic_over_gloo(world):
    import torch.multiprocessing as mp
    mp.spawn(_cpu_worker, args=(world, _free_port()), nprocs=world, join=True)


# ---------------------------------------------------------------- KPM-preconditioned solve of the sharded lattice
def _pcg_problem(Ls=4, beta=2.1, seed=7, kind="holstein"):
    """Global oracle: setup!(P), one application z = P^-1 r, and ldiv!(x, model, b, P) (src/Models.jl:74-137)."""
    from oracle.kpm import KPMPreconditioner
    from oracle.solvers import ldiv
    if kind == "ssh":
        from helpers_ssh import oracle_ssh
        om, rng = oracle_ssh(Lside=Ls, beta=beta, dtau=0.05, seed=seed)
    else:
        om, rng = oracle_holstein("square", Ls, beta, 0.1, mu=-0.5, seed=seed)
    noise = rng.normal(size=2 * om.N)
    P = KPMPreconditioner(om, n=min(20, om.N))
    P.setup(noise)
    assert P.active
    r = rng.normal(size=om.Ndim)
    z = np.zeros(om.Ndim)
    P.ldiv(z, r)
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    x = np.zeros(om.Ndim)
    cg = ConjugateGradient(om.Ndim, tol=1e-8, maxiter=5000)
    it, res, flag = ldiv(x, om, b, cg, P)
    assert flag == 0
    eng = lambda a: np.ascontiguousarray(a.reshape(om.N, om.L).T)
    return om, noise, P, eng(r), eng(z), eng(b), eng(x), it


def _check_pcg(make_backend, comm, rank, world, device="cpu", Ls=4, beta=2.1, aux=None, p2p=False, fused=False, kind="holstein"):
    import torch
    from elphdynamics_b200.sharded import ShardedKPM, ShardedOperator, slab_bounds
    om, noise, Pref, r, z_ref, b, x_ref, it_ref = _pcg_problem(Ls, beta, kind=kind)
    tau0, lloc = slab_bounds(om.L, world, rank)
    be = make_backend(om, tau0, lloc)
    be.kpm_init(aux(om) if aux else None, n=min(20, om.N))
    op = ShardedOperator(be, comm, tol=1e-8, maxiter=5000)
    if p2p:
        assert op.enable_p2p()
    op.update_model()
    P = ShardedKPM(op, om.N, om.L)
    if fused:
        assert P.enable_fused(tau0)                # transposes through peer memory, one call per application
    else:
        assert not hasattr(be, "kpm_shard_setup") or not P.fused
    P.setup(noise)
    assert P.active and P.recomputed
    assert np.array_equal(np.asarray(be.kpm_orders()), Pref.order)
    # 20 Arnoldi steps amplify last-bit differences of the Gram-Schmidt sums (1e-9 in the bounds): compare the bounds at 1e-6 as
    # the single-GPU tests do, then evaluate the oracle's polynomials on the engine's window for the 1e-11 comparison
    lo, hi, e_min, e_max = be.kpm_window()
    # (SSH at weak coupling: A is close to translation invariant, the spectrum is nearly degenerate and 20 Krylov steps run into
    # the noise floor of the orthogonalisation, so the extreme Ritz values depend on its rounding: measured 1e-4 .. 5e-4 between
    # the engine's two-pass classical Gram-Schmidt and the oracle's modified Gram-Schmidt at 32x32 -- identical to 13 digits between
    # the sharded set-up and the single-GPU engine.  The bounds only enter through a 5 % buffer, a hysteresis and floor() of the
    # orders, which are compared exactly above.)
    btol = 1e-6 if kind == "holstein" else 1e-3
    assert abs(e_min - Pref.e_min) <= btol * Pref.e_min and abs(e_max - Pref.e_max) <= btol * Pref.e_max, (e_min, Pref.e_min, e_max, Pref.e_max)
    assert abs(lo - Pref.lam_lo) <= btol * Pref.lam_lo and abs(hi - Pref.lam_hi) <= btol * Pref.lam_hi
    if (lo, hi) != (Pref.lam_lo, Pref.lam_hi):
        from oracle.kpm import kpm_coefficients
        Pref.lam_lo, Pref.lam_hi = lo, hi
        Pref.lam_avg, Pref.lam_mag = (hi + lo) / 2, (hi - lo) / 2
        Pref.coeff = [kpm_coefficients(int(Pref.order[w]), lo, hi, Pref.phis[w]) for w in range(Pref.Lo2)]
        zz = np.zeros(om.Ndim)
        Pref.ldiv(zz, np.ascontiguousarray(r.T).reshape(-1))
        z_ref = np.ascontiguousarray(zz.reshape(om.N, om.L).T)

    def slab(a):
        t = be.empty()
        t[1:lloc + 1] = torch.from_numpy(a[tau0:tau0 + lloc]).to(device)
        return t
    z = be.empty()
    P.ldiv(z, slab(r))
    assert relerr(z[1:lloc + 1].cpu().numpy(), z_ref[tau0:tau0 + lloc]) <= 1e-11, rank
    x = be.empty()
    x.fill_(5.0)                                   # output only
    it, res, flag = op.ldiv(x, slab(b), P=P)
    assert flag == 0 and abs(it - it_ref) <= 2, (it, it_ref)
    assert res <= 1e-4
    assert relerr(x[1:lloc + 1].cpu().numpy(), x_ref[tau0:tau0 + lloc]) <= 1e-6
    # second set-up on the same field: inside the hysteresis window, polynomials kept (:296-309)
    P.setup(noise)
    assert P.active and not P.recomputed
    # cut-off preconditioned solve: the fallback of ldiv! runs the plain CG with 10 x maxiter and converges
    keep = op.maxiter
    op.maxiter = 3
    it2, res2, flag2 = op.ldiv(x, slab(b), P=P)
    op.maxiter = keep
    assert flag2 in (0, 2) and it2 > 3
    if fused:
        be.kpm_shard_check()                       # no barrier of the fused applications timed out
    return be


def _cpu_pcg_worker(rank, world, port, Ls, beta, kind="holstein"):
    import torch.distributed as dist
    from elphdynamics_b200.sharded import RingComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cls = OracleSSHSlabBackend if kind == "ssh" else OracleSlabBackend
        _check_pcg(lambda om, t0, ll: cls(om, t0, ll), RingComm(rank, world), rank, world, Ls=Ls, beta=beta, kind=kind)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ssh_kpm_pcg_host_logic_over_gloo(world):
    """ShardedKPM on a tau-sharded SSH lattice over gloo: the tau-mean of the (cosh, sinh) table by local sums + all-reduce
    (update_A!, src/KPMPreconditioners.jl:355-381), the omega-sharded application and the preconditioned solve against the oracle."""
    import torch.multiprocessing as mp
    mp.spawn(_cpu_pcg_worker, args=(world, _free_port(), 4, 1.05, "ssh"), nprocs=world, join=True)


@pytest.mark.parametrize("world,Ls,beta", [(2, 4, 2.1), (3, 4, 2.1), (3, 4, 0.4)])
def test_sharded_kpm_pcg_host_logic_over_gloo(world, Ls, beta):
    """ShardedKPM (tau-mean all-reduce, the four all-to-alls of an application, round-robin frequencies, mirror rebuild incl.
    the odd-Ltau middle frequency at beta = 2.1 and fewer frequencies than ranks at beta = 0.4) and the preconditioned CG with
    its ldiv! wrapper over gloo, NumPy slab backend, against the oracle's global preconditioner and solve."""
    import torch.multiprocessing as mp
    mp.spawn(_cpu_pcg_worker, args=(world, _free_port(), Ls, beta), nprocs=world, join=True)


def _langevin_reference(om, method, dt, seed=21, precond=False):
    """Global oracle step (plain CG, or KPM-preconditioned with injected Arnoldi start vectors) + the injected noise (engine
    layout) + Q in [k][site] layout."""
    from oracle import langevin as olang
    from oracle.fourier import FourierAccelerator
    from oracle.kpm import KPMPreconditioner
    rng = np.random.default_rng(seed)
    eta, g1, g2 = rng.normal(size=om.Ndof), rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    an1, an2 = rng.normal(size=2 * om.N), rng.normal(size=2 * om.N)
    _langevin_reference.arnoldi = (an1, an2)
    P = KPMPreconditioner(om, n=min(20, om.N)) if precond else None
    fa = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
    fa.update_Q(0.0, 10.0, 1.0)
    cg = ConjugateGradient(om.Ndim, tol=1e-10, maxiter=20000)
    x0 = om.x.copy()
    if method == "euler":
        it = olang.evolve_euler(om, cg, fa, P, dt, eta, g1, an1)
    elif method == "heun":
        it = olang.evolve_heun(om, cg, fa, P, dt, eta, g1, g2, an1, an2)
    else:
        it = olang.evolve_rk(om, cg, fa, P, dt, eta, g1, g2, an1, an2)
    x1 = om.x.copy()
    om.x[:] = x0
    om.update_model()
    eng = lambda a: np.ascontiguousarray(a.reshape(om.N, om.L).T)
    return eng(eta), eng(g1), eng(g2), eng(fa.Q), eng(x0), eng(x1), it


def _check_langevin(make_backend, comm, rank, world, method, device="cpu", Ls=4, beta=1.1, p2p=False, precond=False, aux=None):
    import torch
    from elphdynamics_b200.sharded import ShardedKPM, ShardedLangevin, ShardedOperator, slab_bounds
    om, rng = oracle_holstein("square", Ls, beta, 0.1, mu=-0.5, seed=7)
    dt = 1e-3
    eta, g1, g2, Q, x0, x1, it_ref = _langevin_reference(om, method, dt, precond=precond)
    an1, an2 = _langevin_reference.arnoldi
    tau0, lloc = slab_bounds(om.L, world, rank)
    s0, nloc = slab_bounds(om.N, world, rank)
    be = make_backend(om, tau0, lloc)
    be.make_fft_plan(om.L)
    op = ShardedOperator(be, comm, tol=1e-10, maxiter=20000)
    if p2p:
        assert op.enable_p2p()      # the solves of the step run in the peer-memory persistent kernel
    Qb = torch.from_numpy(np.ascontiguousarray(Q[:, s0:s0 + nloc])).to(device)
    P = None
    if precond:
        be.kpm_init(aux(om) if aux else None, n=min(20, om.N))
        P = ShardedKPM(op, om.N, om.L)
    lang = ShardedLangevin(op, om.N, om.L, tau0, Qb, dt, P=P)
    lang.set_x(x0[tau0:tau0 + lloc])

    def slab(a):
        t = be.empty()
        t[1:lloc + 1] = torch.from_numpy(a[tau0:tau0 + lloc]).to(device)
        return t
    if method == "euler":
        it = lang.evolve_euler(slab(eta), slab(g1), an1)
    elif method == "heun":
        it = lang.evolve_heun(slab(eta), slab(g1), slab(g2), an1, an2)
    else:
        it = lang.evolve_rk(slab(eta), slab(g1), slab(g2), an1, an2)
    assert abs(it - it_ref) <= 2, (it, it_ref)
    assert lang.last_flag == 0 and (P is None or P.active)
    got = lang.xh[1:lloc + 1].cpu().numpy()
    assert relerr(got - x0[tau0:tau0 + lloc], (x1 - x0)[tau0:tau0 + lloc]) <= 1e-7, (method, rank)


def _cpu_langevin_worker(rank, world, port, method, precond):
    import torch.distributed as dist
    from elphdynamics_b200.sharded import RingComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _check_langevin(lambda om, t0, ll: OracleSlabBackend(om, t0, ll), RingComm(rank, world), rank, world, method,
                        precond=precond, beta=2.0 if precond else 1.1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,method,precond", [(2, "rk", False), (3, "euler", False), (2, "rk", True), (3, "heun", False)])
def test_sharded_langevin_host_logic_over_gloo(world, method, precond):
    """Whole tau-sharded Langevin step (halo exchanges, all-reduces, all-to-all transposes around the tau-FFT; with
    ``precond`` the KPM set-ups and preconditioned solves of ShardedKPM) over gloo, NumPy slab backend, against the oracle's
    global step with identical injected noise."""
    import torch.multiprocessing as mp
    mp.spawn(_cpu_langevin_worker, args=(world, _free_port(), method, precond), nprocs=world, join=True)


def test_slab_bounds_cover_the_time_axis():
    from elphdynamics_b200.sharded import slab_bounds
    for L in (11, 20, 200, 400):
        for world in (1, 2, 3, 4, 8):
            spans = [slab_bounds(L, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(s[1] for s in spans) == L
            for (t0, l0), (t1, _) in zip(spans[:-1], spans[1:]):
                assert t0 + l0 == t1
            assert max(s[1] for s in spans) - min(s[1] for s in spans) <= 1


# ------------------------------------------------------------------------------------------- GPU
def _engine_slab(om, tau0, lloc):
    """Engine model for one slab: Ltau = lloc, field = the slab of the global field."""
    import elphdynamics_b200 as E
    lat = om.lat
    elat = E.Lattice(E.UnitCell(lat.ndim, lat.norbits), lat.L1, lat.L2, lat.L3)
    em = E.HolsteinModel(elat, lloc * om.dtau, om.dtau, tol=om.tol, maxiter=om.maxiter)
    assert em.Ltau == lloc
    em.assign_omega(om.omega); em.assign_mu(om.mu); em.assign_lambda(om.lam); em.assign_lambda2(om.lam2); em.assign_omega4(om.omega4)
    off = 0
    for (o1, o2, d), cnt in zip(om.geom_defs, om.geom.def_counts):
        em.assign_t(om.t[off:off + cnt], o1, o2, d)
        off += cnt
    em.initialize_model_()
    em.x = np.ascontiguousarray(om.x.reshape(om.N, om.L)[:, tau0:tau0 + lloc]).reshape(-1)
    return em


class _InProcessRing:
    """All slabs live in one process on one GPU: the 'exchange' copies between the slabs' tensors."""

    def __init__(self, world):
        self.world = world
        self.registry = {}      # id(tensor of rank r) -> list of the matching tensors of all ranks

    def comm(self, rank):
        ring = self

        class _C:
            def exchange(self, v, lloc, lo=True, hi=True):
                group = ring.registry[id(v)]
                w = ring.world
                if lo:
                    left = group[(rank - 1) % w]
                    v[0].copy_(left[left.shape[0] - 2])
                if hi:
                    v[lloc + 1].copy_(group[(rank + 1) % w][1])

            def allreduce_sum(self, t):
                return t
        return _C()


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,world", [(4, 1), (4, 3), (32, 2), (32, 4)])
def test_open_slab_kernels_single_gpu(Ls, world):
    """Several slabs on ONE GPU, halos copied in-process: checks the open-slab kernels (generic and register/shuffle),
    the global-tau sign and uneven slab lengths against the oracle's global products."""
    import torch
    from elphdynamics_b200.sharded import CudaSlabBackend, ShardedOperator, slab_bounds
    om, V, outs, b, x_ref, it_ref = _problem(Ls=Ls, beta=1.1 if Ls == 4 else 0.9)
    ring = _InProcessRing(world)
    slabs = []
    for r in range(world):
        tau0, lloc = slab_bounds(om.L, world, r)
        em = _engine_slab(om, tau0, lloc)
        be = CudaSlabBackend(em, tau0, om.L)
        slabs.append((em, be, tau0, lloc))
    # expnV halos
    for r, (em, be, tau0, lloc) in enumerate(slabs):
        be.update_model()
    Ds = [be.D_tensor() for (_, be, _, _) in slabs]
    for r, (em, be, tau0, lloc) in enumerate(slabs):
        Ds[r][lloc + 1].copy_(Ds[(r + 1) % world][1])
    vs, ys = [], []
    for (em, be, tau0, lloc) in slabs:
        v = be.empty()
        v[1:lloc + 1] = torch.from_numpy(V[tau0:tau0 + lloc]).cuda()
        vs.append(v)
        ys.append(be.empty())
    for r in range(world):
        ring.registry[id(vs[r])] = vs
    for mode, name in ((0, "M"), (1, "MT"), (2, "MTM")):
        for r, (em, be, tau0, lloc) in enumerate(slabs):
            ring.comm(r).exchange(vs[r], lloc)
            be.matvec(mode, vs[r], ys[r])
            got = ys[r][1:lloc + 1].cpu().numpy()
            assert relerr(got, outs[name][tau0:tau0 + lloc]) <= 1e-12, (name, r)
    # force <dM/dx> = u^T dM/dx v on the slab (left halo of v)
    rng = np.random.default_rng(3)
    u_g = rng.normal(size=om.Ndim)
    d_ref = np.zeros(om.Ndof)
    om.muldMdx(d_ref, u_g, V.T.reshape(-1))
    d_ref = d_ref.reshape(om.N, om.L).T
    U = u_g.reshape(om.N, om.L).T
    for r, (em, be, tau0, lloc) in enumerate(slabs):
        u = be.empty()
        u[1:lloc + 1] = torch.from_numpy(np.ascontiguousarray(U[tau0:tau0 + lloc])).cuda()
        out = be.empty()
        be.muldMdx(u, vs[r], out, 1.0)
        assert relerr(out[1:lloc + 1].cpu().numpy(), d_ref[tau0:tau0 + lloc]) <= 1e-9, r
    if world == 1:
        em, be, tau0, lloc = slabs[0]
        from elphdynamics_b200.sharded import RingComm
        op = ShardedOperator(be, RingComm(0, 1), tol=1e-5, maxiter=5000)
        bb = be.empty()
        bb[1:lloc + 1] = torch.from_numpy(b).cuda()
        x = be.empty()
        it, eps = op.solve_cg(x, bb)
        assert abs(it - it_ref) <= 2 and relerr(x[1:lloc + 1].cpu().numpy(), x_ref) <= 1e-3
    for em, *_ in slabs:
        em.close()


# ------------------------------------------------------------------------------------------- SSH slabs
def _engine_ssh_slab(om, tau0, lloc):
    """Engine SSH model for one slab: Ltau = lloc, field = the slab of the global field."""
    import elphdynamics_b200 as E
    lat = om.lat
    elat = E.Lattice(E.UnitCell(lat.ndim, lat.norbits), lat.L1, lat.L2, lat.L3)
    em = E.SSHModel(elat, lloc * om.dtau, om.dtau, tol=om.tol, maxiter=om.maxiter)
    assert em.Ltau == lloc
    em.assign_mu(om.mu)
    for bd in om.bond_defs:
        em.assign_hopping(bd.t, bd.omega, bd.omega4, bd.alpha, bd.alpha2, bd.o1, bd.o2, bd.d, bd.name)
    em.initialize_model_()
    em.x = np.ascontiguousarray(om.x.reshape(om.Nph, om.L)[:, tau0:tau0 + lloc]).reshape(-1)
    return em


def _ssh_problem(Ls=4, beta=1.0):
    from helpers_ssh import oracle_ssh
    om, rng = oracle_ssh(Lside=Ls, beta=beta, dtau=0.05, seed=11)
    V, outs = _global_reference(om, rng)
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    x = np.zeros(om.Ndim)
    it = solve_cg(x, om, b, ConjugateGradient(om.Ndim, tol=1e-5, maxiter=5000))
    u = rng.normal(size=om.Ndim)
    d = np.zeros(om.Ndof)
    om.muldMdx(d, u, np.ascontiguousarray(V.T).reshape(-1))
    eng = lambda a: np.ascontiguousarray(a.reshape(om.N, om.L).T)
    return om, V, outs, eng(b), eng(x), it, eng(u), np.ascontiguousarray(d.reshape(om.Nph, om.L).T)


def _cuda_ssh_backend(om, tau0, lloc):
    from elphdynamics_b200.sharded import CudaSlabBackend
    return CudaSlabBackend(_engine_ssh_slab(om, tau0, lloc), tau0, om.L)


def _check_ssh_rank(comm, rank, world, Ls=4, beta=1.0, make_backend=_cuda_ssh_backend, device="cuda"):
    """Products, plain CG and the force <dM/dx> of a tau-sharded SSH lattice (open-slab generic kernels with the per-slice
    (cosh, sinh) table halo) against the oracle's global operator (src/SSHModels.jl:581-701, :745-830)."""
    import torch
    from elphdynamics_b200.sharded import ShardedOperator, slab_bounds
    om, V, outs, b, x_ref, it_ref, U, d_ref = _ssh_problem(Ls, beta)
    tau0, lloc = slab_bounds(om.L, world, rank)
    be = make_backend(om, tau0, lloc)
    assert be.is_ssh and be.Nph == om.Nph
    op = ShardedOperator(be, comm, tol=1e-5, maxiter=5000)
    op.update_model()

    def slab(a):
        t = be.empty()
        t[1:lloc + 1] = torch.from_numpy(a[tau0:tau0 + lloc]).to(device)
        return t
    v, y = slab(V), be.empty()
    for name, fn in (("M", op.mulM), ("MT", op.mulMT), ("MTM", op.mulMTM)):
        fn(y, v)
        assert relerr(y[1:lloc + 1].cpu().numpy(), outs[name][tau0:tau0 + lloc]) <= 1e-12, (name, rank)
    x = be.empty()
    it, eps = op.solve_cg(x, slab(b))
    assert abs(it - it_ref) <= 2, (it, it_ref)
    assert relerr(x[1:lloc + 1].cpu().numpy(), x_ref[tau0:tau0 + lloc]) <= 1e-3
    comm.exchange(v, lloc, lo=True, hi=False)
    out = be.empty_field()
    be.muldMdx(slab(U), v, out, 1.0)
    assert relerr(out[1:lloc + 1].cpu().numpy(), d_ref[tau0:tau0 + lloc]) <= 1e-9, rank
    if hasattr(be, "model"):
        be.model.close()


def _check_ssh_langevin(comm, rank, world, method="rk", Ls=4, beta=1.0, make_backend=_cuda_ssh_backend, device="cuda"):
    """One Langevin step of a tau-sharded SSH lattice (fields on the bonds: Nph columns in the Fourier acceleration and the
    bosonic gradient, Nsites columns in the solves) against the oracle's global step with identical injected noise."""
    import torch
    from helpers_ssh import oracle_ssh
    from oracle import langevin as olang
    from oracle.fourier import FourierAccelerator
    from elphdynamics_b200.sharded import ShardedLangevin, ShardedOperator, slab_bounds
    om, rng = oracle_ssh(Lside=Ls, beta=beta, dtau=0.05, seed=11)
    assert np.array_equal(om.primary_field, np.arange(om.Ndof))
    dt = 1e-3
    eta, g1, g2 = rng.normal(size=om.Ndof), rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    fa = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
    fa.update_Q(0.0, 10.0, 0.1)
    cg = ConjugateGradient(om.Ndim, tol=1e-10, maxiter=20000)
    x0 = om.x.copy()
    it_ref = olang.evolve_euler(om, cg, fa, None, dt, eta, g1) if method == "euler" else olang.evolve_rk(om, cg, fa, None, dt, eta, g1, g2)
    x1 = om.x.copy()
    om.x[:] = x0
    om.update_model()
    engf = lambda a: np.ascontiguousarray(a.reshape(om.Nph, om.L).T)     # fields: [tau][phonon]
    engv = lambda a: np.ascontiguousarray(a.reshape(om.N, om.L).T)       # site vectors: [tau][site]
    tau0, lloc = slab_bounds(om.L, world, rank)
    s0, nloc = slab_bounds(om.Nph, world, rank)
    be = make_backend(om, tau0, lloc)
    be.make_fft_plan(om.L)
    op = ShardedOperator(be, comm, tol=1e-10, maxiter=20000)
    Qb = torch.from_numpy(np.ascontiguousarray(engf(fa.Q)[:, s0:s0 + nloc])).to(device)
    lang = ShardedLangevin(op, om.N, om.L, tau0, Qb, dt)
    lang.set_x(engf(x0)[tau0:tau0 + lloc])

    def slab(a, field):
        t = be.empty_field() if field else be.empty()
        t[1:lloc + 1] = torch.from_numpy(a[tau0:tau0 + lloc]).to(device)
        return t
    if method == "euler":
        it = lang.evolve_euler(slab(engf(eta), True), slab(engv(g1), False))
    else:
        it = lang.evolve_rk(slab(engf(eta), True), slab(engv(g1), False), slab(engv(g2), False))
    assert abs(it - it_ref) <= 2, (it, it_ref)
    assert lang.last_flag == 0
    got = lang.xh[1:lloc + 1].cpu().numpy()
    assert relerr(got - engf(x0)[tau0:tau0 + lloc], engf(x1 - x0)[tau0:tau0 + lloc]) <= 1e-7, (method, rank)
    if hasattr(be, "model"):
        be.model.close()


def _cpu_ssh_worker(rank, world, port, what):
    import torch.distributed as dist
    from elphdynamics_b200.sharded import RingComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mk = lambda om, t0, ll: OracleSSHSlabBackend(om, t0, ll)
        if what == "operator":
            _check_ssh_rank(RingComm(rank, world), rank, world, Ls=4, beta=1.0, make_backend=mk, device="cpu")
        else:
            _check_ssh_langevin(RingComm(rank, world), rank, world, what, Ls=4, beta=1.0, make_backend=mk, device="cpu")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,what", [(2, "operator"), (3, "operator"), (2, "rk"), (3, "euler")])
def test_sharded_ssh_host_logic_over_gloo(world, what):
    """tau-sharded SSH lattice over gloo with a NumPy slab backend: the (cosh, sinh) table halo (rows of 2*Ncolumns doubles, refreshed
    after every update_model!), products, CG and force against the oracle's global operator, and a whole Langevin step with
    field-shaped slabs (Nph columns) next to site vectors."""
    import torch.multiprocessing as mp
    mp.spawn(_cpu_ssh_worker, args=(world, _free_port(), what), nprocs=world, join=True)


@pytest.mark.gpu
@pytest.mark.parametrize("method,Ls", [("euler", 4), ("rk", 32)])
def test_sharded_ssh_langevin_single_gpu(method, Ls):
    from elphdynamics_b200.sharded import RingComm
    _check_ssh_langevin(RingComm(0, 1), 0, 1, method, Ls=Ls, beta=1.0 if Ls == 4 else 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,beta", [(4, 1.0), (32, 0.5)])
def test_sharded_ssh_single_gpu(Ls, beta):
    """world = 1: the SSH open-slab path (table halo = the slab's own first row, self-exchanged) against the oracle."""
    from elphdynamics_b200.sharded import RingComm
    _check_ssh_rank(RingComm(0, 1), 0, 1, Ls=Ls, beta=beta)


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,world", [(4, 3), (32, 2)])
def test_open_slab_ssh_kernels_single_gpu(Ls, world):
    """Several SSH slabs on ONE GPU, halos copied in-process: the (cosh, sinh) row of the right neighbour's first slice, the
    global-tau sign and uneven slab lengths against the oracle's global products and force."""
    import torch
    from elphdynamics_b200.sharded import CudaSlabBackend, slab_bounds
    om, V, outs, b, x_ref, it_ref, U, d_ref = _ssh_problem(Ls, 1.0 if Ls == 4 else 0.5)
    ring = _InProcessRing(world)
    slabs = []
    for r in range(world):
        tau0, lloc = slab_bounds(om.L, world, r)
        em = _engine_ssh_slab(om, tau0, lloc)
        be = CudaSlabBackend(em, tau0, om.L)
        be.update_model()
        slabs.append((em, be, tau0, lloc))
    Ds = [be.D_tensor() for (_, be, _, _) in slabs]
    for r, (em, be, tau0, lloc) in enumerate(slabs):
        Ds[r][lloc + 1].copy_(Ds[(r + 1) % world][1])
    vs, ys = [], []
    for (em, be, tau0, lloc) in slabs:
        v = be.empty()
        v[1:lloc + 1] = torch.from_numpy(V[tau0:tau0 + lloc]).cuda()
        vs.append(v)
        ys.append(be.empty())
    for r in range(world):
        ring.registry[id(vs[r])] = vs
    for mode, name in ((0, "M"), (1, "MT"), (2, "MTM")):
        for r, (em, be, tau0, lloc) in enumerate(slabs):
            ring.comm(r).exchange(vs[r], lloc)
            be.matvec(mode, vs[r], ys[r])
            assert relerr(ys[r][1:lloc + 1].cpu().numpy(), outs[name][tau0:tau0 + lloc]) <= 1e-12, (name, r)
    for r, (em, be, tau0, lloc) in enumerate(slabs):
        u = be.empty()
        u[1:lloc + 1] = torch.from_numpy(U[tau0:tau0 + lloc]).cuda()
        out = be.empty_field()
        be.muldMdx(u, vs[r], out, 1.0)
        assert relerr(out[1:lloc + 1].cpu().numpy(), d_ref[tau0:tau0 + lloc]) <= 1e-9, r
    for em, *_ in slabs:
        em.close()


def _gpu_worker(rank, world, port):
    import torch
    import torch.distributed as dist
    from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedOperator, slab_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        om, V, outs, b, x_ref, it_ref = _problem(Ls=32, beta=0.9)
        tau0, lloc = slab_bounds(om.L, world, rank)
        em = _engine_slab(om, tau0, lloc)
        be = CudaSlabBackend(em, tau0, om.L)
        op = ShardedOperator(be, RingComm(rank, world), tol=1e-5, maxiter=5000)
        _check_rank(op, be, rank, tau0, lloc, V, outs, om, b, x_ref, it_ref)
        _check_p2p_cg(be, op.comm, tau0, lloc, b, x_ref, it_ref)
        dist.barrier()
        em.close()
        _check_ssh_rank(RingComm(rank, world), rank, world, Ls=32, beta=0.5)     # SSH slabs: table halo through NCCL
        dist.barrier()
        _check_ssh_langevin(RingComm(rank, world), rank, world, "rk", Ls=32, beta=0.5)
        dist.barrier()
        _check_pcg(_cuda_ssh_backend, RingComm(rank, world), rank, world, device="cuda", Ls=32, beta=1.05, aux=_engine_ssh_global,
                   fused=True, kind="ssh")
        dist.barrier()
        # KPM-preconditioned solve: omega-sharded application through NCCL all-to-alls, products with the halo through NCCL and
        # (second pass) through peer memory inside the product kernel
        for p2p, fused, beta in ((False, False, 2.0), (True, False, 2.0), (True, True, 2.0), (True, True, 2.1), (False, True, 0.5)):
            _check_pcg(_cuda_backend, RingComm(rank, world), rank, world, device="cuda", Ls=32, beta=beta, aux=_engine_global,
                       p2p=p2p, fused=fused)
            dist.barrier()
    finally:
        dist.destroy_process_group()


def _check_p2p_cg(be, comm, tau0, lloc, b_glob, x_ref, it_ref):
    """Peer-memory CG (csrc/cg_p2p.cu): one persistent kernel per GPU, halo + all-reduce inside the kernel."""
    import torch
    assert be.p2p_setup(comm) and be._p2p_ready
    # halo exchange through peer memory (elph_dev_shard_halo) against the torch.distributed exchange, repeated (tag parity)
    for rep in range(3):
        v = be.empty()
        v[1:lloc + 1] = torch.from_numpy(b_glob[tau0:tau0 + lloc] + rep).to(v.device)
        ref = v.clone()
        comm.exchange(ref, lloc)
        be.halo_p2p(v)
        torch.cuda.synchronize()
        assert torch.equal(v, ref)
    # M^T M with the exchange INSIDE the product kernel (mtm_square.cu, HALO; tuning key 22) against exchange + product, the two
    # forms interleaved so that the tag parity of the arena rows is exercised
    from elphdynamics_b200.sharded import MTM_MODE
    for rep, fused in enumerate((1, 0, 1, 1, 0, 1)):
        v = be.empty()
        v[1:lloc + 1] = torch.from_numpy(b_glob[tau0:tau0 + lloc] * (1.0 + 0.1 * rep)).to(v.device)
        ref = v.clone()
        comm.exchange(ref, lloc)
        y_ref = be.empty()
        be.matvec(MTM_MODE, ref, y_ref)
        be.model._call("elph_set_tuning", 22, fused)
        y = be.empty()
        be.matvec_halo(MTM_MODE, v, y)
        torch.cuda.synchronize()
        assert torch.equal(v, ref), (rep, fused)                       # the halo rows of v are filled either way
        assert torch.equal(y[1:lloc + 1], y_ref[1:lloc + 1]), (rep, fused)
    be.model._call("elph_set_tuning", 22, 1)
    b = be.empty()
    b[1:lloc + 1] = torch.from_numpy(b_glob[tau0:tau0 + lloc]).to(b.device)
    x = be.empty()
    it, eps = be.cg_p2p(x, b, 1e-5, 7)        # cut off by maxiter: every rank returns the same count and residual
    assert it == 7 and eps > 1e-5
    for rep in range(3):                      # repeated solves: the barrier sequence numbers carry over
        x = be.empty()
        x.fill_(3.0)                          # output only (x0 = 0 inside)
        it, eps = be.cg_p2p(x, b, 1e-5, 5000)
        assert abs(it - it_ref) <= 2, (it, it_ref)
        assert eps < 1e-5
        assert relerr(x[1:lloc + 1].cpu().numpy(), x_ref[tau0:tau0 + lloc]) <= 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,beta", [(32, 0.9), (32, 4.0), (64, 0.5)])
def test_p2p_cg_single_gpu(Ls, beta):
    """world = 1: the ring closes on the GPU itself (the pushes land in its own halo rows, one mailbox slot)."""
    from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedOperator
    om, V, outs, b, x_ref, it_ref = _problem(Ls=Ls, beta=beta)
    em = _engine_slab(om, 0, om.L)
    be = CudaSlabBackend(em, 0, om.L)
    op = ShardedOperator(be, RingComm(0, 1), tol=1e-5, maxiter=5000)
    op.update_model()
    _check_p2p_cg(be, op.comm, 0, om.L, b, x_ref, it_ref)
    em.close()


def _cuda_backend(om, tau0, lloc):
    from elphdynamics_b200.sharded import CudaSlabBackend
    return CudaSlabBackend(_engine_slab(om, tau0, lloc), tau0, om.L)


def _engine_global(om):
    """The auxiliary engine model of the GLOBAL lattice that owns the FFT plan, the polynomials and the chain kernels."""
    return _engine_slab(om, 0, om.L)


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,beta", [(4, 2.1), (32, 2.0), (32, 3.1)])
def test_sharded_kpm_pcg_single_gpu(Ls, beta):
    """world = 1: ShardedKPM on the CUDA backend (column FFTs, set-up from the supplied tau-mean, chain kernels on the frequency
    subset -- generic at 4x4, register tiles at 32x32; odd Ltau at beta = 2.1 / 3.1) and the preconditioned solve against the
    oracle's global preconditioner and ldiv!."""
    from elphdynamics_b200.sharded import RingComm
    _check_pcg(_cuda_backend, RingComm(0, 1), 0, 1, device="cuda", Ls=Ls, beta=beta, aux=_engine_global)


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,beta", [(4, 2.1), (32, 2.0), (32, 3.1), (64, 1.2)])
def test_sharded_kpm_fused_single_gpu(Ls, beta):
    """world = 1: the application with the transposes through the arenas (csrc/kpm_shard.cu; the ring closes on the GPU itself:
    copy-in, pulling forward FFT, gather, chains, pulling inverse FFT, gather, four barrier kernels) against the oracle."""
    from elphdynamics_b200.sharded import RingComm
    _check_pcg(_cuda_backend, RingComm(0, 1), 0, 1, device="cuda", Ls=Ls, beta=beta, aux=_engine_global, fused=True)


def _engine_ssh_global(om):
    return _engine_ssh_slab(om, 0, om.L)


@pytest.mark.gpu
@pytest.mark.parametrize("Ls,beta,fused", [(4, 1.05, False), (32, 1.0, False), (32, 1.05, True)])
def test_sharded_ssh_kpm_pcg_single_gpu(Ls, beta, fused):
    """world = 1: ShardedKPM on SSH slabs (set-up from the supplied mean of the (cosh, sinh) table, chain kernels with per-bond
    tables -- generic at 4x4, register tiles at 32x32; odd Ltau at beta = 1.05) and the preconditioned solve against the oracle."""
    from elphdynamics_b200.sharded import RingComm
    _check_pcg(_cuda_ssh_backend, RingComm(0, 1), 0, 1, device="cuda", Ls=Ls, beta=beta, aux=_engine_ssh_global, fused=fused,
               kind="ssh")


@pytest.mark.gpu
def test_sharded_langevin_kpm_single_gpu():
    """The sharded Runge-Kutta step with KPM-preconditioned solves (two set-ups from injected Arnoldi vectors), world = 1."""
    from elphdynamics_b200.sharded import RingComm
    _check_langevin(_cuda_backend, RingComm(0, 1), 0, 1, "rk", device="cuda", Ls=32, beta=2.0, precond=True, aux=_engine_global)


@pytest.mark.gpu
def test_sharded_langevin_p2p_single_gpu():
    """The sharded Runge-Kutta step with its solves in the peer-memory CG kernel (world = 1: the ring closes on itself)."""
    from elphdynamics_b200.sharded import RingComm
    _check_langevin(_cuda_backend, RingComm(0, 1), 0, 1, "rk", device="cuda", Ls=32, beta=0.8, p2p=True)


@pytest.mark.gpu
@pytest.mark.parametrize("method,Ls", [("euler", 4), ("rk", 32), ("heun", 32)])
def test_sharded_langevin_single_gpu(method, Ls):
    """world = 1: the whole sharded driver (open-slab kernels, halo self-exchange, FFT plan handle, column FFT, slab
    bosonic gradient) on one GPU against the oracle's global step."""
    from elphdynamics_b200.sharded import RingComm
    _check_langevin(_cuda_backend, RingComm(0, 1), 0, 1, method, device="cuda", Ls=Ls, beta=1.1 if Ls == 4 else 0.8)


def _gpu_langevin_worker(rank, world, port):
    import torch
    import torch.distributed as dist
    from elphdynamics_b200.sharded import RingComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _check_langevin(_cuda_backend, RingComm(rank, world), rank, world, "rk", device="cuda", Ls=32, beta=0.8)
        dist.barrier()
        _check_langevin(_cuda_backend, RingComm(rank, world), rank, world, "rk", device="cuda", Ls=32, beta=0.8, p2p=True)
        dist.barrier()
        _check_langevin(_cuda_backend, RingComm(rank, world), rank, world, "rk", device="cuda", Ls=32, beta=2.0, precond=True,
                        aux=_engine_global)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_langevin_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_gpu_langevin_worker, args=(world, _free_port()), nprocs=world, join=True)


@pytest.mark.gpu
def test_sharded_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_gpu_worker, args=(world, _free_port()), nprocs=world, join=True)
