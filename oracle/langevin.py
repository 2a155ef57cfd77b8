"""Oracle: Langevin dynamics updates and the fermion force.  TEST INFRASTRUCTURE ONLY.

Follows ``src/LangevinDynamics.jl``:
  * ``evolve!`` Euler :81-119, Runge-Kutta :162-225, Heun :272-324
  * ``calc_dSdx!`` :334-345, ``calc_dSfdx!`` :350-384

Randomness is INJECTED.  In the reference every draw comes from ``model.rng`` in
this order per force evaluation: ``randn!(rng,R)`` (wasted, overwritten at :360),
``randn!(rng,g)``, then inside ``setup!(P)`` the 2N Arnoldi start values
(``src/KPMPreconditioners.jl:859-861,902-904``).  ``eta`` is drawn first through
``randn!(eta,model)`` (SSH remaps it through ``primary_field``,
``src/SSHModels.jl:567-576``).  Callers pass ``eta``, ``g`` and ``arnoldi_noise``
arrays; the wasted draws are simply not needed.
"""
from __future__ import annotations

import math

import numpy as np

from .action import calc_dSbdx
from .solvers import ldiv


def calc_dSfdx(dSfdx, g, Minv_g, model, cg, P, arnoldi_noise=None):
    """src/LangevinDynamics.jl:350-384 with g injected.  Returns (iters, resid, flag)."""
    if P is not None and not getattr(P, "is_identity", False):
        P.setup(arnoldi_noise)
    Minv_g[:] = 0.0
    model.mulMT(model.v2, g)                       # b = M^T g  (into model.v'')
    iters, err, flag = ldiv(Minv_g, model, model.v2, cg, P)
    model.muldMdx(dSfdx, g, Minv_g)
    dSfdx *= -2.0
    return iters, err, flag


def calc_dSdx(dSdx, g, Minv_g, model, cg, P, arnoldi_noise=None):
    """src/LangevinDynamics.jl:334-345."""
    out = calc_dSfdx(dSdx, g, Minv_g, model, cg, P, arnoldi_noise)
    calc_dSbdx(dSdx, model, True)
    return out


def _eta(model, eta):
    eta = np.asarray(eta, dtype=np.float64).copy()
    if model.kind == "ssh":
        eta = eta[model.primary_field]
    return eta


def evolve_euler(model, cg, fa, P, dt, eta, g, arnoldi_noise=None):
    """src/LangevinDynamics.jl:81-119.  Returns iters."""
    model.update_model()
    eta = _eta(model, eta)
    dSdx = np.zeros(model.Ndof)
    Minv = np.zeros(model.Ndim)
    iters, _, _ = calc_dSdx(dSdx, np.asarray(g, dtype=np.float64), Minv, model, cg, P, arnoldi_noise)
    QdSdx = fa.accelerate(dSdx, 1.0)
    sqrtQeta = fa.accelerate(eta, 0.5)
    dx = math.sqrt(2.0 * dt) * sqrtQeta - dt * QdSdx
    model.x += dx
    model.update_model()
    return iters


def evolve_rk(model, cg, fa, P, dt, eta, g1, g2, arnoldi_noise1=None, arnoldi_noise2=None, trace=None):
    """src/LangevinDynamics.jl:162-225.  Returns iters of the SECOND solve (:198)."""
    model.update_model()
    eta = _eta(model, eta)
    dSdx = np.zeros(model.Ndof)
    dSdx2 = np.zeros(model.Ndof)
    Minv = np.zeros(model.Ndim)
    it1, _, _ = calc_dSdx(dSdx, np.asarray(g1, dtype=np.float64), Minv, model, cg, P, arnoldi_noise1)
    if trace is not None:
        trace["dSdx1"] = dSdx.copy()
        trace["iters1"] = it1
    dx = math.sqrt(2 * dt) * eta - dt * dSdx
    model.x[:] = model.x + dx
    model.update_model()
    iters, _, _ = calc_dSdx(dSdx2, np.asarray(g2, dtype=np.float64), Minv, model, cg, P, arnoldi_noise2)
    if trace is not None:
        trace["dSdx2"] = dSdx2.copy()
        trace["iters2"] = iters
    model.x[:] = model.x - dx
    model.update_model()
    dSdx = (dSdx2 + dSdx) / 2.0
    QdSdx = fa.accelerate(dSdx, 1.0)
    sqrtQeta = fa.accelerate(eta, 0.5)
    dx = math.sqrt(2.0 * dt) * sqrtQeta - dt * QdSdx
    model.x[:] = model.x + dx
    model.update_model()
    return iters


def evolve_heun(model, cg, fa, P, dt, eta, g1, g2, arnoldi_noise1=None, arnoldi_noise2=None):
    """src/LangevinDynamics.jl:272-324.  Returns div(iters1+iters2, 2)."""
    eta = _eta(model, eta)
    xi = fa.accelerate(eta, 0.5)
    model.update_model()
    dSdx = np.zeros(model.Ndof)
    dSdx2 = np.zeros(model.Ndof)
    Minv = np.zeros(model.Ndim)
    it1, _, _ = calc_dSdx(dSdx, np.asarray(g1, dtype=np.float64), Minv, model, cg, P, arnoldi_noise1)
    dG = fa.accelerate(dSdx, 1.0)
    dx = math.sqrt(2 * dt) * xi - dt * dG
    model.x[:] = model.x + dx
    model.update_model()
    it2, _, _ = calc_dSdx(dSdx2, np.asarray(g2, dtype=np.float64), Minv, model, cg, P, arnoldi_noise2)
    dG2 = fa.accelerate(dSdx2, 1.0)
    model.x[:] = model.x - dx
    model.x[:] = model.x + math.sqrt(2 * dt) * xi - dt * (dG + dG2) / 2
    model.update_model()
    return (it1 + it2) // 2
