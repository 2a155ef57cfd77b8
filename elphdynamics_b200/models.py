"""Host-side mirror of the reference's model/operator API (src/Models.jl, src/HolsteinModels.jl,
src/SSHModels.jl) on top of the C ABI of libelph_b200.so.

Names follow the reference with ``!`` written as a trailing underscore:
``mulM_(y, model, v)``, ``mulMT_``, ``mulMTM_``, ``muldMdx_``, ``update_model_``, ``ldiv_``.
(Python normalises identifiers with NFKC, so ``mulMᵀ_`` also resolves to ``mulMT_``.)
Every operator call goes to the GPU library; nothing is computed in Python.
Vectors are float64 numpy arrays in the reference host layout ``index = site*Ltau + tau``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Config, KpmInfo, SolveInfo, check, ptr, out_ptr
from .lattices import Lattice, assemble_checkerboard, calc_neighbor_table

HOLSTEIN, SSH = 0, 1


def _f64(a, n=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"{name}: expected {n} entries, got {a.size}")
    return a


class ConjugateGradient:
    """Parameters of ``ConjugateGradient`` (src/IterativeSolvers.jl:36-57); the work vectors
    r, p, z live on the device inside the model's handle."""

    def __init__(self, N: int, tol: float = 1e-4, maxiter: int = 0, kappa_max: float = 1e12):
        self.tol = float(tol)
        self.maxiter = int(maxiter) if maxiter >= 1 else int(N)
        self.kappa_max = float(kappa_max)
        self.N = int(N)


class AbstractModel:
    """``AbstractModel`` (src/Models.jl:65): owns the engine handle."""
    kind = -1

    def __init__(self):
        self._h = C.c_void_p()
        self._lib = None
        self.mul_by_M = False      # CG solves M^T M x = b (src/HolsteinModels.jl:286-288)
        self.transposed = False

    # ------------------------------------------------------------------ engine plumbing
    def _create(self, cfg: Config, keep):
        self._lib = _lib.load()
        self._keep = keep
        h = C.c_void_p()
        st = self._lib.elph_create(C.byref(cfg), C.byref(h))
        check(st, None)
        self._h = h
        self._keep = None

    def close(self):
        if self._lib is not None and self._h:
            for addr in list(getattr(self, "_pinned", {})):      # page locks taken through this handle
                self._lib.elph_host_unregister(self._h, C.c_void_p(addr))
            self._pinned = {}
            self._lib.elph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("model is not initialised (call initialize_model_)")
        return self._h

    def _call(self, name, *args):
        check(getattr(self._lib, name)(self.handle, *args), self.handle)

    def set_stream(self, stream_ptr: int):
        self._call("elph_set_stream", C.c_void_p(stream_ptr))

    def synchronize(self):
        self._call("elph_synchronize")

    def launch_count(self) -> int:
        return int(self._lib.elph_launch_count(self.handle))

    # ------------------------------------------------------------------ field access
    @property
    def x(self) -> np.ndarray:
        """Phonon fields (copy; authoritative copy lives on the device)."""
        out = np.empty(self.Ndof)
        self._call("elph_get_x", ptr(out))
        return out

    @x.setter
    def x(self, value):
        self._call("elph_set_x", ptr(_f64(value, self.Ndof, "x")))

    def pin_host(self, array: np.ndarray):
        """Page-lock a long-lived host array that is passed to the host-buffer entry points repeatedly (noise vectors,
        right-hand sides): ``elph_host_register``.  Call ``unpin_host`` before the array goes away."""
        assert array.flags["C_CONTIGUOUS"]
        if self._h is None or not self._h:
            raise RuntimeError("model is closed")
        if array.ctypes.data in getattr(self, "_pinned", {}):
            return                       # already page-locked through this handle
        self._call("elph_host_register", C.c_void_p(array.ctypes.data), array.nbytes)
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        self._pinned[array.ctypes.data] = array          # keeps the array alive while it is registered

    def unpin_host(self, array: np.ndarray):
        if self._h is None or not self._h:
            return                       # the handle is gone; the driver released the registration with the context
        self._call("elph_host_unregister", C.c_void_p(array.ctypes.data))
        getattr(self, "_pinned", {}).pop(array.ctypes.data, None)

    def set_mu(self, mu):
        self.mu = _f64(mu, self.Nsites, "mu").copy()
        self._call("elph_set_mu", ptr(self.mu))

    def __len__(self):
        return self.Ndim

    @property
    def shape(self):
        return (self.Ndim, self.Ndim)


class HolsteinModel(AbstractModel):
    """``HolsteinModel(lattice, beta, dtau; ...)`` (src/HolsteinModels.jl:196-313)."""
    kind = HOLSTEIN

    def __init__(self, lattice: Lattice, beta: float, dtau: float, tol: float = 1e-4, maxiter: int = 10000,
                 device: int = -1):
        super().__init__()
        self.lattice = lattice
        self.beta, self.dtau = float(beta), float(dtau)
        self.Ltau = int(round(beta / dtau))
        self.Nsites = lattice.nsites
        self.Nph = self.Nsites
        self.Ndof = self.Nph * self.Ltau
        self.Ndim = self.Ndof
        self.Nbonds = 0
        self.nbonds = 0
        self.device = device
        self.mu = np.zeros(self.Nsites)
        self.omega = np.zeros(self.Nph)
        self.omega4 = np.zeros(self.Nph)
        self.lam = np.zeros(self.Nph)
        self.lam2 = np.zeros(self.Nph)
        self.t = np.zeros(0)
        self._neighbor_table = np.zeros((2, 0), dtype=np.int64)
        self.solver = ConjugateGradient(self.Ndim, tol=tol, maxiter=maxiter)

    # assign_* (src/HolsteinModels.jl:324-444).  Disorder (stddev) draws come from the caller's RNG
    # in the reference; here a per-site/per-bond array may be passed instead of a scalar.
    def _assign(self, arr, value, orbit):
        if getattr(self, "_h", None):
            # the engine copied the parameters at initialize_model_: changing the host copy would silently be ignored
            raise RuntimeError("assign_* after initialize_model_ does not reach the device: use set_mu (the chemical potential is "
                               "the only parameter the reference changes during a run, src/MuFinder.jl) or build a new model")
        v = np.asarray(value, dtype=np.float64)
        if orbit is None:
            arr[:] = v
        else:
            sel = self.lattice.site_to_orbit == orbit
            arr[sel] = v if v.ndim == 0 else v[sel]

    def assign_mu(self, value, orbit=None):
        self._assign(self.mu, value, orbit)

    def assign_lambda(self, value, orbit=None):
        self._assign(self.lam, value, orbit)

    def assign_lambda2(self, value, orbit=None):
        self._assign(self.lam2, value, orbit)

    def assign_omega(self, value, orbit=None):
        self._assign(self.omega, value, orbit)

    def assign_omega4(self, value, orbit=None):
        self._assign(self.omega4, value, orbit)

    def assign_t(self, t, o1: int, o2: int, v):
        """``assign_t!`` (src/HolsteinModels.jl:420-444); ``t`` scalar or one value per new bond."""
        new = calc_neighbor_table(self.lattice, o1, o2, v)
        self.nbonds += 1
        self._neighbor_table = np.concatenate([self._neighbor_table, new], axis=1)
        tv = np.asarray(t, dtype=np.float64)
        self.t = np.concatenate([self.t, np.full(new.shape[1], float(tv)) if tv.ndim == 0 else tv])

    def initialize_model_(self):
        """``initialize_model!`` (src/HolsteinModels.jl:484-517) + engine creation."""
        self.Nbonds = self.t.size
        tabs = assemble_checkerboard(self._neighbor_table)
        self.neighbor_table = tabs.neighbor_table
        self.checkerboard_perm = tabs.checkerboard_perm
        self.inv_checkerboard_perm = tabs.inv_checkerboard_perm
        self.group_sizes = tabs.group_sizes
        self.cosht = np.cosh(self.dtau * self.t)[tabs.inv_checkerboard_perm]
        self.sinht = np.sinh(self.dtau * self.t)[tabs.inv_checkerboard_perm]
        nt = np.ascontiguousarray(self.neighbor_table.T)           # Nbonds (i,j) pairs = Julia (2,Nbonds) column-major
        cfg = Config()
        cfg.model, cfg.index_base, cfg.device = HOLSTEIN, 0, self.device
        cfg.Ltau, cfg.Nsites, cfg.Nbonds, cfg.Nph = self.Ltau, self.Nsites, self.Nbonds, self.Nph
        cfg.dtau = self.dtau
        keep = [nt, self.cosht, self.sinht, self.lam, self.lam2, self.mu, self.omega, self.omega4]
        cfg.neighbor_table = ptr(nt, np.int64)
        cfg.cosht, cfg.sinht = ptr(self.cosht), ptr(self.sinht)
        cfg.lambda_, cfg.lambda2, cfg.mu = ptr(self.lam), ptr(self.lam2), ptr(self.mu)
        cfg.omega, cfg.omega4 = ptr(self.omega), ptr(self.omega4)
        cfg.cg_tol, cfg.cg_maxiter, cfg.cg_kappa_max = self.solver.tol, self.solver.maxiter, self.solver.kappa_max
        self._create(cfg, keep)

    @property
    def expnV(self) -> np.ndarray:
        """exp(-dtau V[x]) as of the last update_model_ (host layout)."""
        out = np.empty(self.Ndim)
        self._call("elph_get_expnV", ptr(out))
        return out


# ---------------------------------------------------------------------------------------------
# operator functions (reference names)
# ---------------------------------------------------------------------------------------------
def update_model_(model: AbstractModel):
    """``update_model!`` (src/HolsteinModels.jl:526-549, src/SSHModels.jl:510-562)."""
    model._call("elph_update_model")


def _vec_io(model, name, *vectors_in, out, n_out=None):
    ins = [_f64(v, None) for v in vectors_in]
    model._call(name, *[ptr(v) for v in ins], out_ptr(out, model.Ndim if n_out is None else n_out, "output"))


def mulM_(y, model, v):
    """``mulM!(y, model, v)`` (src/HolsteinModels.jl:569, src/SSHModels.jl:581)."""
    _vec_io(model, "elph_mulM", _f64(v, model.Ndim, "v"), out=y)


def mulMT_(y, model, v):
    """``mulMᵀ!(y, model, v)`` (src/HolsteinModels.jl:631, src/SSHModels.jl:646)."""
    _vec_io(model, "elph_mulMT", _f64(v, model.Ndim, "v"), out=y)


def mulMTM_(y, model, v):
    """``mulMᵀM!(y, model, v)`` (src/Models.jl:215-224) -- one fused kernel."""
    _vec_io(model, "elph_mulMTM", _f64(v, model.Ndim, "v"), out=y)


def mul_(y, model, v):
    """``mul!(y, model, v)`` dispatch on mul_by_M / transposed (src/Models.jl:192-209)."""
    if model.mul_by_M:
        (mulMT_ if model.transposed else mulM_)(y, model, v)
    else:
        if model.transposed:
            raise NotImplementedError("M M^T is only used by the GMRES/BiCGStab paths (out of scope)")
        mulMTM_(y, model, v)


def muldMdx_(dMdx, u, model, v):
    """``muldMdx!(dMdx, u, model, v)`` (src/HolsteinModels.jl:691, src/SSHModels.jl:707)."""
    model._call("elph_muldMdx", ptr(_f64(u, model.Ndim, "u")), ptr(_f64(v, model.Ndim, "v")), out_ptr(dMdx, model.Ndof, "dMdx"))


def ldiv_(x, model, b, P=None, tol_power: float = 1.0):
    """``ldiv!(x, model, b[, P])`` -> ``(iters, residual_error, flag)`` (src/Models.jl:74-186).
    ``x`` holds the initial guess on entry (callers zero it) and the solution on exit."""
    info = SolveInfo()
    use_p = 0 if (P is None or getattr(P, "is_identity", False)) else 1
    model._call("elph_solve", ptr(_f64(b, model.Ndim, "b")), out_ptr(x, model.Ndim, "x"), use_p, float(tol_power), C.byref(info))
    model.last_solve_info = info
    return info.astuple()


def ldiv_batch_(X, model, B, P=None, tol_power: float = 1.0):
    """``fill!(X, 0); ldiv!(X[:,k], model, B[:,k][, P])`` for every column k in one call -- the measurement solves of
    ``update!(Gr, model, P)`` (src/GreensFunctions.jl:201-234).  ``B``, ``X``: arrays of shape (nrhs, Ndim).
    Returns a list of ``(iters, residual_error, flag)``."""
    B = np.ascontiguousarray(B, dtype=np.float64)
    if B.ndim != 2 or B.shape[1] != model.Ndim:
        raise ValueError(f"B must have shape (nrhs, {model.Ndim})")
    if not (isinstance(X, np.ndarray) and X.dtype == np.float64 and X.flags.c_contiguous and X.shape == B.shape):
        raise ValueError("X must be a C-contiguous float64 array of the same shape as B")
    nrhs = B.shape[0]
    infos = (SolveInfo * nrhs)()
    use_p = 0 if (P is None or getattr(P, "is_identity", False)) else 1
    model._call("elph_solve_batch", nrhs, ptr(B), ptr(X), use_p, float(tol_power), infos)
    return [infos[k].astuple() for k in range(nrhs)]


def update_Gr_(MinvR, model, R, P=None):
    """The solves of ``update!(Gr, model, P)`` (src/GreensFunctions.jl:201-234) for all random vectors at once:
    ``MinvR[k] = (MᵀM)⁻¹ Mᵀ R[k]``.  ``R`` (shape (nv, Ndim)) is drawn by the caller; ``setup_(P)`` is called before."""
    R = np.ascontiguousarray(R, dtype=np.float64)
    if R.ndim != 2 or R.shape[1] != model.Ndim:
        raise ValueError(f"R must have shape (nv, {model.Ndim})")
    if not (isinstance(MinvR, np.ndarray) and MinvR.dtype == np.float64 and MinvR.flags.c_contiguous and MinvR.shape == R.shape):
        raise ValueError("MinvR must be a C-contiguous float64 array of the same shape as R")
    nv = R.shape[0]
    infos = (SolveInfo * nv)()
    use_p = 0 if (P is None or getattr(P, "is_identity", False)) else 1
    model._call("elph_Minv_batch", nv, ptr(R), ptr(MinvR), use_p, infos)
    return [infos[k].astuple() for k in range(nv)]


def solve_(x, model, b, P=None, tol: float = 0.0, maxiter: int = 0):
    """Raw ``solve!(x, A, b, cg[, P])`` (src/IterativeSolvers.jl:153, :239): returns the iteration count."""
    it = C.c_int64()
    eps = C.c_double()
    use_p = 0 if (P is None or getattr(P, "is_identity", False)) else 1
    model._call("elph_cg_solve", ptr(_f64(b, model.Ndim, "b")), out_ptr(x, model.Ndim, "x"), use_p, float(tol), int(maxiter),
                C.byref(it), C.byref(eps))
    model.last_eps = eps.value
    return int(it.value)


class Identity:
    """``LinearAlgebra.I`` as the preconditioner (src/IterativeSolvers.jl:14-17)."""
    is_identity = True


I = Identity()


class SymmetricKPMPreconditioner:
    """``SymmetricKPMPreconditioner(model, n, buf, c1, c2)`` (src/KPMPreconditioners.jl:219-235)."""
    is_identity = False

    def __init__(self, model: AbstractModel, n: int = 20, buf: float = 0.05, c1: float = 1.0, c2: float = 1.0):
        self.model = model
        self.n, self.buf, self.c1, self.c2 = int(n), float(buf), float(c1), float(c2)
        model._call("elph_kpm_configure", self.n, self.buf, self.c1, self.c2)
        self.info = KpmInfo()

    @property
    def active(self):
        return bool(self.info.active)

    def orders(self) -> np.ndarray:
        out = np.zeros((self.model.Ltau + 1) // 2, dtype=np.int64)
        self.model._call("elph_kpm_get_orders", ptr(out, np.int64))
        return out

    def coeff(self, w: int) -> np.ndarray:
        n = int(self.orders()[w])
        out = np.zeros(2 * n)
        self.model._call("elph_kpm_get_coeff", int(w), ptr(out))
        return out[0::2] + 1j * out[1::2]


def setup_(P, arnoldi_noise=None):
    """``setup!(P)`` (src/KPMPreconditioners.jl:259-321).  The 2*Nsites Arnoldi start values the
    reference draws from ``model.rng`` are passed in."""
    if P is None or getattr(P, "is_identity", False):
        return None
    noise = _f64(arnoldi_noise, 2 * P.model.Nsites, "arnoldi_noise")
    P.model._call("elph_kpm_setup", ptr(noise), C.byref(P.info))
    return P.info


def kpm_ldiv_(vout, P, vin):
    """``ldiv!(vout, P, vin)`` (src/KPMPreconditioners.jl:426-481)."""
    if getattr(P, "is_identity", False):
        vout[:] = vin
        return
    P.model._call("elph_kpm_apply", ptr(_f64(vin, P.model.Ndim, "vin")), out_ptr(vout, P.model.Ndim, "vout"))


class SSHModel(AbstractModel):
    """``SSHModel(lattice, beta, dtau; ...)`` (src/SSHModels.jl:79-314): phonons live on bonds and modulate the
    hopping, ``t' = t - (alpha x + sign(x) alpha2 x^2)``."""
    kind = SSH

    def __init__(self, lattice: Lattice, beta: float, dtau: float, tol: float = 1e-4, maxiter: int = 10000, device: int = -1):
        super().__init__()
        self.lattice = lattice
        self.beta, self.dtau = float(beta), float(dtau)
        self.Ltau = int(round(beta / dtau))
        self.Nsites = lattice.nsites
        self.Ndim = self.Nsites * self.Ltau
        self.device = device
        self.mu = np.zeros(self.Nsites)
        self.bond_definitions = []
        self.solver = ConjugateGradient(self.Ndim, tol=tol, maxiter=maxiter)
        self.Nbonds = self.Nph = self.Ndof = 0

    def assign_mu(self, value, orbit=None):
        """``assign_μ!`` (src/SSHModels.jl:332-343)."""
        if getattr(self, "_h", None):
            raise RuntimeError("assign_mu after initialize_model_ does not reach the device: use set_mu")
        v = np.asarray(value, dtype=np.float64)
        sel = slice(None) if orbit is None else (self.lattice.site_to_orbit == orbit)
        self.mu[sel] = v if v.ndim == 0 else v[sel]

    def assign_hopping(self, t, omega, omega4, alpha, alpha2, o1: int, o2: int, v, name: str = ""):
        """``assign_hopping!`` (src/SSHModels.jl:319-327); a phonon lives on the bond iff omega != 0."""
        self.bond_definitions.append(dict(t=float(t), omega=float(omega), omega4=float(omega4), alpha=float(alpha),
                                          alpha2=float(alpha2), o1=o1, o2=o2, v=tuple(v), name=name))

    def initialize_model_(self):
        """``initialize_model!`` (src/SSHModels.jl:348-505) + engine creation."""
        tables, t, om, om4, al, al2, p2b, b2p, names = [], [], [], [], [], [], [], [], []
        ntypes = 0
        for i, bd in enumerate(self.bond_definitions):
            nn = calc_neighbor_table(self.lattice, bd["o1"], bd["o2"], bd["v"])
            n_new = nn.shape[1]
            tables.append(nn)
            t += [bd["t"]] * n_new
            if bd["omega"] != 0.0:
                ntypes += 1
                names.append(bd["name"] if bd["name"] else f"__unnamed_{i}")
                om += [bd["omega"]] * n_new
                om4 += [bd["omega4"]] * n_new
                al += [bd["alpha"]] * n_new
                al2 += [bd["alpha2"]] * n_new
                p2b += list(range(i * n_new, (i + 1) * n_new))
                b2p += list(range((ntypes - 1) * n_new, ntypes * n_new))
            else:
                b2p += [-1] * n_new
        nt = np.concatenate(tables, axis=1) if tables else np.zeros((2, 0), dtype=np.int64)
        tabs = assemble_checkerboard(nt)
        self.neighbor_table = tabs.neighbor_table
        self.checkerboard_perm = np.ascontiguousarray(tabs.checkerboard_perm, dtype=np.int64)
        self.inv_checkerboard_perm = np.ascontiguousarray(tabs.inv_checkerboard_perm, dtype=np.int64)
        self.group_sizes = tabs.group_sizes
        self.t = np.asarray(t, dtype=np.float64)
        self.omega, self.omega4 = np.asarray(om, dtype=np.float64), np.asarray(om4, dtype=np.float64)
        self.alpha, self.alpha2 = np.asarray(al, dtype=np.float64), np.asarray(al2, dtype=np.float64)
        self.phonon_to_bond = np.asarray(p2b, dtype=np.int64)
        self.bond_to_phonon = np.asarray(b2p, dtype=np.int64)
        self.nph = ntypes
        self.Nbonds, self.Nph = self.t.size, self.omega.size
        self.Ndof = self.Nph * self.Ltau
        primary = np.arange(self.Ndof, dtype=np.int64)
        if ntypes > 0:
            per = self.Ndof // ntypes
            pf = primary.reshape(ntypes, per)
            for a in range(ntypes):
                for b in range(a + 1, ntypes):
                    if names[a] == names[b] and pf[b, 0] > a * per:
                        pf[b, :] = np.arange(a * per, (a + 1) * per)
        self.primary_field = primary
        ntp = np.ascontiguousarray(self.neighbor_table.T)
        cfg = Config()
        cfg.model, cfg.index_base, cfg.device = SSH, 0, self.device
        cfg.Ltau, cfg.Nsites, cfg.Nbonds, cfg.Nph = self.Ltau, self.Nsites, self.Nbonds, self.Nph
        cfg.dtau = self.dtau
        cfg.neighbor_table = ptr(ntp, np.int64)
        cfg.mu, cfg.omega, cfg.omega4 = ptr(self.mu), ptr(self.omega), ptr(self.omega4)
        cfg.t, cfg.alpha, cfg.alpha2 = ptr(self.t), ptr(self.alpha), ptr(self.alpha2)
        cfg.checkerboard_perm = ptr(self.checkerboard_perm, np.int64)
        cfg.inv_checkerboard_perm = ptr(self.inv_checkerboard_perm, np.int64)
        cfg.phonon_to_bond = ptr(self.phonon_to_bond, np.int64)
        cfg.bond_to_phonon = ptr(self.bond_to_phonon, np.int64)
        cfg.primary_field = ptr(self.primary_field, np.int64)
        cfg.cg_tol, cfg.cg_maxiter, cfg.cg_kappa_max = self.solver.tol, self.solver.maxiter, self.solver.kappa_max
        self._create(cfg, [ntp])

    def cosh_sinh(self):
        """(cosht, sinht) as (Nbonds, Ltau) arrays = C-order view of the reference's (Ltau, Nbonds) matrices."""
        c = np.empty((self.Nbonds, self.Ltau))
        s = np.empty((self.Nbonds, self.Ltau))
        self._call("elph_get_cosh_sinh", ptr(c), ptr(s))
        return c, s
