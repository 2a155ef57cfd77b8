"""Oracle: bosonic (phonon) action and its gradient.  TEST INFRASTRUCTURE ONLY.

Follows ``src/PhononAction.jl``:
  * ``calc_Sb`` Holstein :11-66 (dispersion branch :40-60 is dead code in every
    shipped example and references an undefined ``L``; left out), SSH :68-107
  * ``calc_dSbdx!`` Holstein :114-187, SSH :189-233 -- ACCUMULATES into dSbdx.
"""
from __future__ import annotations

import numpy as np


def calc_Sb(model, shifted: bool = False) -> float:
    L, dt = model.L, model.dtau
    if model.kind == "holstein":
        X = model.x.reshape(model.N, L)
        Xm1 = np.roll(X, 1, axis=1)
        w, w4, lam = model.omega[:, None], model.omega4[:, None], model.lam[:, None]
        Sb = np.sum(w ** 2 * X ** 2 / 2 + w4 * X ** 4 - lam * X * float(shifted))
        Sb += np.sum((X - Xm1) ** 2 / dt ** 2 / 2)
        return float(dt * Sb)
    # SSH: only primary phonons contribute (:79-103)
    X = model.x.reshape(model.Nph, L)
    Sb = 0.0
    for i in range(model.Nph):
        f0 = i * L
        if model.primary_field[f0] == f0:
            x = X[i]
            xm1 = np.roll(x, 1)
            Sb += float(np.sum(dt * model.omega[i] ** 2 * x ** 2 / 2 + dt * model.omega4[i] * x ** 4))
            Sb += float(np.sum((x - xm1) ** 2 / dt / 2))
    return Sb


def calc_dSbdx(dSbdx, model, shifted: bool = False):
    """dSbdx += dSb/dx."""
    L, dt = model.L, model.dtau
    nph = model.Nph
    assert dSbdx.size == model.Ndof
    X = model.x.reshape(nph, L)
    D = dSbdx.reshape(nph, L)
    Xp = np.roll(X, -1, axis=1)
    Xm = np.roll(X, 1, axis=1)
    w2 = (dt * model.omega * model.omega)[:, None]
    w4 = (dt * 4 * model.omega4)[:, None]
    if model.kind == "holstein":
        dtlam = (dt * model.lam)[:, None]
        D += w2 * X - dtlam * float(shifted)
        D += w4 * X * X * X
        D -= (Xp + Xm - 2.0 * X) / dt
    else:
        d = w2 * X
        d = d + w4 * X * X * X
        d = d - (Xp + Xm - 2.0 * X) / dt
        D += d
