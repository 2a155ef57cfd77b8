"""Oracle: conjugate gradient and the ``ldiv!`` solve wrappers.  TEST INFRASTRUCTURE ONLY.

Follows:
  * ``src/IterativeSolvers.jl:153-234``  preconditioned CG  ``solve!(x,A,b,cg,P)``
  * ``src/IterativeSolvers.jl:239-314``  plain CG           ``solve!(x,A,b,cg)``
  * ``src/Models.jl:74-137``             ``ldiv!(x,model,b,P)``  (true residual, flags, fallback)
  * ``src/Models.jl:139-186``            ``ldiv!(x,model,b)``

``A`` is any object with ``mul(y, x)``; ``P`` any object with ``ldiv(z, r)``
(the reference duck-types on ``mul!(y,A,x)`` and ``ldiv!(z,P,r)``).
"""
from __future__ import annotations

import math

import numpy as np


class ConjugateGradient:
    """Workspace + defaults, src/IterativeSolvers.jl:36-57."""

    def __init__(self, n: int, tol: float = 1e-4, maxiter: int = 0, kappa_max: float = 1e12):
        self.tol = float(tol)
        self.maxiter = int(maxiter) if maxiter >= 1 else n
        self.kappa_max = float(kappa_max)
        self.N = n
        self.r = np.zeros(n)
        self.p = np.zeros(n)
        self.z = np.zeros(n)
        self.history = []   # eps_j per iteration (debug aid, not in the reference)


def _kappa(j, eps0, eps):
    """(2j / log(2 eps0/eps))^2 with IEEE semantics (log(<=0) -> nan/-inf like Julia
    under @fastmath would not trap either)."""
    with np.errstate(all="ignore"):
        return float((2.0 * j / np.log(2.0 * eps0 / eps)) ** 2)


def solve_pcg(x, A, b, cg: ConjugateGradient, P, maxiter: int = 0, tol: float = 0.0, kappa_max: float = 0.0) -> int:
    """src/IterativeSolvers.jl:153-234."""
    r, p, z = cg.r, cg.p, cg.z
    if maxiter == 0:
        maxiter = cg.maxiter
    if tol == 0.0:
        tol = cg.tol
    if kappa_max == 0.0:
        kappa_max = cg.kappa_max
    cg.history = []
    normb = np.linalg.norm(b)
    A.mul(r, x)
    r[:] = b - r                      # axpby!(1.0,b,-1.0,r)
    P.ldiv(z, r)
    p[:] = z
    rdotz = float(np.dot(r, z))
    eps0 = float(np.linalg.norm(r) / normb)
    kappa_min = 0.0
    for j in range(1, maxiter + 1):
        A.mul(z, p)
        alpha = rdotz / float(np.dot(p, z))
        x += alpha * p
        r -= alpha * z
        eps = float(np.linalg.norm(r) / normb)
        cg.history.append(eps)
        k = _kappa(j, eps0, eps)
        kappa_min = max(kappa_min, k) if not math.isnan(k) else kappa_min
        if eps < tol or kappa_min > kappa_max:
            return j
        P.ldiv(z, r)
        new_rdotz = float(np.dot(r, z))
        beta = new_rdotz / rdotz
        rdotz = new_rdotz
        p[:] = z + beta * p           # axpby!(1.0,z,beta,p)
    return maxiter


def solve_cg(x, A, b, cg: ConjugateGradient, maxiter: int = 0, tol: float = 0.0, kappa_max: float = 0.0) -> int:
    """src/IterativeSolvers.jl:239-314.  (The reference's ``iszero(tol) || kmin > kmax``
    at :252 reads an unassigned variable only when ``tol != 0``; every caller
    passes the default ``tol = 0.0`` so the short-circuit always wins.)"""
    r, p, z = cg.r, cg.p, cg.z
    if maxiter == 0:
        maxiter = cg.maxiter
    if tol == 0.0:
        tol = cg.tol
    if kappa_max == 0.0:
        kappa_max = cg.kappa_max
    cg.history = []
    normb = np.linalg.norm(b)
    A.mul(r, x)
    r[:] = b - r
    p[:] = r
    rdotr = float(np.dot(r, r))
    eps0 = float(np.linalg.norm(r) / normb)
    kappa_min = 0.0
    for j in range(1, maxiter + 1):
        A.mul(z, p)
        alpha = rdotr / float(np.dot(p, z))
        x += alpha * p
        r -= alpha * z
        eps = float(np.linalg.norm(r) / normb)
        cg.history.append(eps)
        k = _kappa(j, eps0, eps)
        kappa_min = max(kappa_min, k) if not math.isnan(k) else kappa_min
        if eps < tol or kappa_min > kappa_max:
            return j
        new_rdotr = float(np.dot(r, r))
        beta = new_rdotr / rdotr
        rdotr = new_rdotr
        p[:] = r + beta * p
    return maxiter


class Identity:
    """``LinearAlgebra.I`` used as a preconditioner, src/IterativeSolvers.jl:14-17."""
    is_identity = True

    def ldiv(self, z, r):
        z[:] = r

    def setup(self, *a, **k):  # ``setup!(op) = nothing``, src/KPMPreconditioners.jl:323-326
        return None


def ldiv_noprecond(x, model, b, cg: ConjugateGradient, maxiter: int = 0):
    """``ldiv!(x, model, b; maxiter)``, src/Models.jl:139-186.
    Returns ``(iters, residual_error, flag)``."""
    if maxiter == 0:
        maxiter = cg.maxiter
    iters = solve_cg(x, model, b, cg, maxiter=maxiter)
    v = model.v3
    model.mul(v, x)
    v[:] = v - b
    residual = float(np.linalg.norm(v) / np.linalg.norm(b))
    if residual > math.sqrt(cg.tol):
        # NOTE: compares against solver.maxiter, not the maxiter argument (src/Models.jl:160)
        flag = 1 if iters == cg.maxiter else 2
        x[:] = 0.0
    else:
        flag = 0
    return iters, residual, flag


def ldiv(x, model, b, cg: ConjugateGradient, P=None, maxiter: int = 0):
    """``ldiv!(x, model, b, P; maxiter)``, src/Models.jl:74-137."""
    if maxiter == 0:
        maxiter = cg.maxiter
    if P is None or getattr(P, "is_identity", False):
        return ldiv_noprecond(x, model, b, cg, maxiter=maxiter)
    iters = solve_pcg(x, model, b, cg, P, maxiter=maxiter)
    v = model.v3
    model.mul(v, x)
    v[:] = v - b
    residual = float(np.linalg.norm(v) / np.linalg.norm(b))
    if residual > math.sqrt(cg.tol):
        flag = 1 if iters == maxiter else 2
        x[:] = 0.0
    else:
        flag = 0
    if flag > 0:
        iters, residual, flag = ldiv_noprecond(x, model, b, cg, maxiter=10 * maxiter)
    return iters, residual, flag
