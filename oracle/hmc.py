"""Oracle: Hybrid Monte Carlo update.  TEST INFRASTRUCTURE ONLY.

Follows ``src/HMC.jl``:
  * ``HybridMonteCarlo`` :20-279 (Nt = round(tr/dt), dt' = dt/Nb)
  * ``update!`` :310-335, ``standard_update!`` :343-473, ``multitimestep_update!`` :479-638
  * ``refresh_v!`` :648-660, ``refresh_phi!`` :666-692
  * ``calc_H/K/S/Sf`` :698-783, ``calc_dSdx!`` :749-759, ``calc_dSfdx!`` :790-814
  * ``calc_O^-1 Lambda phi!`` :820-915 (two solves with tol^power)
  * Lambda operators (Holstein only; no-ops for SSH) :921-1030

Randomness is INJECTED: ``R_v`` (Ndof), ``R_plus``/``R_minus`` (Ndim), one 2N Arnoldi vector per
``calc_Oinv`` call (Nt+2 of them, in call order) and the Metropolis uniform.
Reference quirk kept: the multi-timestep path drops the iteration count of the first solve
(``iters += iters`` at :515).
"""
from __future__ import annotations

import math

import numpy as np

from .action import calc_dSbdx, calc_Sb
from .solvers import ldiv


class HybridMonteCarlo:
    def __init__(self, model, dt: float, tr: float, alpha: float, Nb: int):
        assert 0.0 <= alpha < 1.0
        self.Ndof, self.Ndim = model.Ndof, model.Ndim
        self.dt, self.tr, self.alpha, self.Nb = float(dt), float(tr), float(alpha), int(Nb)
        self.Nt = int(round(tr / dt))
        self.dtp = dt / Nb
        z = lambda n: np.zeros(n)
        self.x0, self.dSdx, self.v, self.v0 = z(self.Ndof), z(self.Ndof), z(self.Ndof), z(self.Ndof)
        self.Lam = np.ones(self.Ndim)
        self.Rp, self.Rm = z(self.Ndim), z(self.Ndim)
        self.phip, self.phim = z(self.Ndim), z(self.Ndim)
        self.Lphip, self.Lphim = z(self.Ndim), z(self.Ndim)
        self.Op, self.Om = z(self.Ndim), z(self.Ndim)
        self.u, self.y = z(self.Ndim), z(self.Ndof)
        self.H = self.S = self.K = 0.0
        self.iters = 0
        self.accepted = False


# ------------------------------------------------------------------ Lambda operators (:921-1030)
def update_Lam(hmc, model):
    if model.kind != "holstein":
        return
    X = model.x.reshape(model.N, model.L)
    hmc.Lam[:] = np.exp(-model.dtau * (model.lam[:, None] * X + model.lam2[:, None] * X ** 2) / 2).reshape(-1)


def mulLam(out, v, hmc, model):
    if model.kind != "holstein":
        return
    N, L = model.N, model.L
    U, O, Lm = v.reshape(N, L), out.reshape(N, L), hmc.Lam.reshape(N, L)
    u1 = U[:, 0].copy()
    O[:, :L - 1] = -Lm[:, 1:] * U[:, 1:]
    O[:, L - 1] = Lm[:, 0] * u1


def mulLaminv(out, v, hmc, model):
    if model.kind != "holstein":
        return
    N, L = model.N, model.L
    U, O, Lm = v.reshape(N, L), out.reshape(N, L), hmc.Lam.reshape(N, L)
    uL = U[:, L - 1].copy()
    O[:, 1:] = -(1.0 / Lm[:, 1:]) * U[:, :L - 1]
    O[:, 0] = (1.0 / Lm[:, 0]) * uL


def muldLamdx(dLdx, vl, vr, hmc, model):
    if model.kind != "holstein":
        return
    N, L, dt = model.N, model.L, model.dtau
    D, VL, VR = dLdx.reshape(N, L), vl.reshape(N, L), vr.reshape(N, L)
    X, Lm = model.x.reshape(N, L), hmc.Lam.reshape(N, L)
    lam, lam2 = model.lam[:, None], model.lam2[:, None]
    D[:, 0] += VL[:, 0] * (-dt * (model.lam / 2 + model.lam2 * X[:, 0])) * Lm[:, 0] * VR[:, L - 1]
    D[:, 1:] += VL[:, 1:] * (dt * (lam / 2 + lam2 * X[:, 1:])) * Lm[:, 1:] * VR[:, :L - 1]


# ------------------------------------------------------------------ pieces
def refresh_v(hmc, model, fa, R_v):
    R = np.asarray(R_v, dtype=np.float64).copy()
    if model.kind == "ssh":
        R = R[model.primary_field]
    hmc.v[:] = hmc.alpha * hmc.v + math.sqrt(1.0 - hmc.alpha ** 2) * fa.accelerate(R, -0.5, use_mass=True)


def refresh_phi(hmc, model, R_plus, R_minus):
    update_Lam(hmc, model)
    hmc.Rp[:] = R_plus
    hmc.Rm[:] = R_minus
    model.mulMT(hmc.Lphip, hmc.Rp)
    mulLaminv(hmc.phip, hmc.Lphip, hmc, model)
    model.mulMT(hmc.Lphim, hmc.Rm)
    mulLaminv(hmc.phim, hmc.Lphim, hmc, model)
    hmc.S = float(np.dot(hmc.Rp, hmc.Rp) / 2 + np.dot(hmc.Rm, hmc.Rm) / 2)
    hmc.S += calc_Sb(model)
    return hmc.S


def calc_Oinv(hmc, model, cg, P, power, arnoldi_noise=None):
    """calc_O^-1 Lambda phi! (:820-915).  Returns (iters, flag)."""
    tol = cg.tol
    cg.tol = tol ** power
    hmc.iters = 0
    if P is not None and not getattr(P, "is_identity", False):
        P.setup(arnoldi_noise)
    update_Lam(hmc, model)
    mulLam(hmc.Lphip, hmc.phip, hmc, model)
    mulLam(hmc.Lphim, hmc.phim, hmc, model)
    hmc.Op[:] = 0.0
    it, _, flag = ldiv(hmc.Op, model, hmc.Lphip, cg, P)
    hmc.iters += it
    if flag == 0:
        hmc.Om[:] = 0.0
        it, _, flag = ldiv(hmc.Om, model, hmc.Lphim, cg, P)
        hmc.iters += it
    if flag == 0:
        hmc.iters = -(-hmc.iters // 2)
    cg.tol = tol
    return hmc.iters, flag


def calc_Sf(hmc):
    return float(np.dot(hmc.Lphip, hmc.Op) / 2 + np.dot(hmc.Lphim, hmc.Om) / 2)


def calc_K(hmc, model, fa):
    mv = fa.accelerate(hmc.v, 1.0, use_mass=True)
    if model.kind == "holstein":
        hmc.K = float(np.dot(hmc.v, mv) / 2)
    else:
        prim = model.primary_field == np.arange(model.Ndof)
        hmc.K = float(np.sum(hmc.v[prim] * mv[prim] / 2))
    return hmc.K


def calc_H(hmc, model, fa):
    S = calc_Sf(hmc) + calc_Sb(model)
    hmc.S = S
    K = calc_K(hmc, model, fa)
    hmc.H = S + K
    return hmc.H, S, K


def calc_dSfdx(hmc, model):
    """dSdx += fermionic force (:790-814)."""
    dM = np.zeros(model.Ndof)
    for O, phi in ((hmc.Op, hmc.phip), (hmc.Om, hmc.phim)):
        model.mulM(hmc.u, O)
        model.muldMdx(dM, hmc.u, O)
        hmc.dSdx += -dM
    muldLamdx(hmc.dSdx, hmc.phip, hmc.Op, hmc, model)
    muldLamdx(hmc.dSdx, hmc.phim, hmc.Om, hmc, model)


def calc_dSdx(hmc, model):
    calc_dSfdx(hmc, model)
    calc_dSbdx(hmc.dSdx, model, False)


# ------------------------------------------------------------------ updates
def _finish(hmc, model, fa, cg, P, flag, iters, H0, noise_iter, uniform):
    Pacc = 0.0
    H1 = float("nan")
    if flag == 0:
        it, flag = calc_Oinv(hmc, model, cg, P, 2.0, next(noise_iter))
        iters += it
        if flag == 0:
            H1, _, _ = calc_H(hmc, model, fa)
            dH = H1 - H0
            Pacc = min(1.0, math.exp(-dH))
    hmc.H1 = H1
    hmc.Pacc = Pacc
    if uniform < Pacc and flag == 0:
        hmc.accepted = True
    else:
        model.x[:] = hmc.x0
        hmc.v[:] = -hmc.v0
        model.update_model()
        hmc.accepted = False
    return hmc.accepted, float(-(-iters // (hmc.Nt + 2)))


def standard_update(model, hmc, fa, cg, P, R_v, R_plus, R_minus, arnoldi_noises, uniform):
    """src/HMC.jl:343-473."""
    noise_iter = iter(arnoldi_noises if arnoldi_noises is not None else [None] * (hmc.Nt + 2))
    dt = hmc.dt
    model.update_model()
    refresh_v(hmc, model, fa, R_v)
    hmc.x0[:] = model.x
    hmc.v0[:] = hmc.v
    refresh_phi(hmc, model, R_plus, R_minus)
    iters, flag = calc_Oinv(hmc, model, cg, P, 2.0, next(noise_iter))
    H0 = float("nan")
    if flag == 0:
        H0, _, _ = calc_H(hmc, model, fa)
        hmc.dSdx[:] = 0.0
        calc_dSdx(hmc, model)
        Q = fa.accelerate(hmc.dSdx, -1.0, use_mass=True)
        for _ in range(hmc.Nt):
            hmc.v[:] = hmc.v - dt / 2 * Q
            model.x[:] = model.x + dt * hmc.v
            model.update_model()
            it, flag = calc_Oinv(hmc, model, cg, P, 1.0, next(noise_iter))
            iters += it
            if flag > 0:
                break
            hmc.dSdx[:] = 0.0
            calc_dSdx(hmc, model)
            Q = fa.accelerate(hmc.dSdx, -1.0, use_mass=True)
            hmc.v[:] = hmc.v - dt / 2 * Q
    hmc.H0 = H0
    return _finish(hmc, model, fa, cg, P, flag, iters, H0, noise_iter, uniform)


def multitimestep_update(model, hmc, fa, cg, P, R_v, R_plus, R_minus, arnoldi_noises, uniform):
    """src/HMC.jl:479-638."""
    noise_iter = iter(arnoldi_noises if arnoldi_noises is not None else [None] * (hmc.Nt + 2))
    dt, dtp, Nb = hmc.dt, hmc.dtp, hmc.Nb
    iters = 0
    model.update_model()
    refresh_v(hmc, model, fa, R_v)
    hmc.x0[:] = model.x
    hmc.v0[:] = hmc.v
    refresh_phi(hmc, model, R_plus, R_minus)
    _itrs, flag = calc_Oinv(hmc, model, cg, P, 2.0, next(noise_iter))
    iters += iters          # sic (:515): the first solve's count is dropped
    H0 = float("nan")
    if flag == 0:
        H0, _, _ = calc_H(hmc, model, fa)
        hmc.dSdx[:] = 0.0
        calc_dSfdx(hmc, model)
        Qf = fa.accelerate(hmc.dSdx, -1.0, use_mass=True)
        for _ in range(hmc.Nt):
            hmc.v[:] = hmc.v - dt / 2 * Qf
            hmc.dSdx[:] = 0.0
            calc_dSbdx(hmc.dSdx, model, False)
            Qb = fa.accelerate(hmc.dSdx, -1.0, use_mass=True)
            for _tp in range(Nb):
                hmc.v[:] = hmc.v - dtp / 2 * Qb
                model.x[:] = model.x + dtp * hmc.v
                hmc.dSdx[:] = 0.0
                calc_dSbdx(hmc.dSdx, model, False)
                Qb = fa.accelerate(hmc.dSdx, -1.0, use_mass=True)
                hmc.v[:] = hmc.v - dtp / 2 * Qb
            model.update_model()
            it, flag = calc_Oinv(hmc, model, cg, P, 1.0, next(noise_iter))
            iters += it
            if flag > 0:
                break
            hmc.dSdx[:] = 0.0
            calc_dSfdx(hmc, model)
            Qf = fa.accelerate(hmc.dSdx, -1.0, use_mass=True)
            hmc.v[:] = hmc.v - dt / 2 * Qf
    hmc.H0 = H0
    return _finish(hmc, model, fa, cg, P, flag, iters, H0, noise_iter, uniform)


def update(model, hmc, fa, cg, P, R_v, R_plus, R_minus, arnoldi_noises, uniform):
    """``update!`` (:310-335)."""
    if hmc.Ndof == 0:
        return True, 0.0
    if hmc.Nb == 1:
        return standard_update(model, hmc, fa, cg, P, R_v, R_plus, R_minus, arnoldi_noises, uniform)
    return multitimestep_update(model, hmc, fa, cg, P, R_v, R_plus, R_minus, arnoldi_noises, uniform)


def special_update(model, hmc, cg, P, kind, targets, R_plus, R_minus, uniforms, arnoldi_noises=None):
    """``special_update!`` (src/SpecialUpdates.jl:97-160 reflection, :233-290 Holstein swap, :296-366 SSH swap) with
    the random draws injected: ``targets`` are 0-based phonon columns (reflection) or pairs of columns (swap).
    Returns (acceptance ratio, log of (accepted, S0, S1, iters, flag))."""
    L = model.L
    x = model.x.reshape(-1, L)                 # host layout: column = phonon, tau fastest
    log = []
    accepted = 0.0
    model.update_model()                       # :123 / :259 / :311

    def move(t):
        if kind == "reflect":
            x[t] = -x[t]                       # :129
        else:
            i, j = t
            tmp = x[i].copy()                  # swap!(x_i, x_j) :269, :335
            x[i] = x[j]
            x[j] = tmp
        model.update_model()

    for k, t in enumerate(targets):
        S0 = refresh_phi(hmc, model, R_plus[k], R_minus[k])
        move(t)
        iters, flag = calc_Oinv(hmc, model, cg, P, 2.0, None if arnoldi_noises is None else arnoldi_noises[k])
        S1 = calc_Sf(hmc) + calc_Sb(model)     # calc_S, src/HMC.jl:745-754
        Pacc = min(1.0, math.exp(-(S1 - S0)))
        ok = (uniforms[k] < Pacc) and flag == 0
        if ok:
            accepted += 1.0
        else:
            move(t)
        log.append((ok, S0, S1, iters, flag))
    return (accepted / len(targets) if len(targets) else 0.0), log
