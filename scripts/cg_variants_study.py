"""CPU study (oracle side only): iteration counts of three algebraically equivalent CG recurrences on A = M^T M with
the stop rule of src/IterativeSolvers.jl:211-219 -- the reference's two-reduction loop, the single-reduction form of
Chronopoulos & Gear (csrc/cg_p2p.cu) and the pipelined form of Ghysels & Vanroose (reduction overlapped with the
product).  Decides which recurrences may run in the persistent kernels under the +-2 iteration criterion.

    python scripts/cg_variants_study.py B E
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

from helpers import oracle_holstein  # noqa: E402
from oracle.cref import CRef  # noqa: E402

CFG = {"A": ("square", 4, 2.0, 0.1), "B": ("square", 32, 20.0, 0.1), "E": ("square", 64, 40.0, 0.1),
       "B16": ("square", 16, 20.0, 0.1)}


def stop(j, eps, eps0, kmin, tol, kmax):
    q = 2.0 * j / np.log(2.0 * eps0 / eps)
    kmin = max(kmin, q * q)
    return (eps < tol or kmin > kmax), kmin


def cg_ref(A, b, tol, maxiter, kmax=1e12):
    x = np.zeros_like(b); r = b.copy(); p = r.copy(); z = np.zeros_like(b)
    nb = np.sqrt(b @ b); rr = r @ r; eps0 = np.sqrt(rr) / nb; kmin = 0.0
    for j in range(1, maxiter + 1):
        A(z, p)
        al = rr / (p @ z)
        x += al * p; r -= al * z
        nr = r @ r
        eps = np.sqrt(nr) / nb
        done, kmin = stop(j, eps, eps0, kmin, tol, kmax)
        if done:
            return j, x
        p *= nr / rr; p += r; rr = nr
    return maxiter, x


def cg_cgear(A, b, tol, maxiter, kmax=1e12):
    x = np.zeros_like(b); r = b.copy(); w = np.zeros_like(b); A(w, r)
    p = np.zeros_like(b); s = np.zeros_like(b)
    nb = np.sqrt(b @ b); gam = r @ r; dl = r @ w; eps0 = np.sqrt(gam) / nb; kmin = 0.0
    al, be = gam / dl, 0.0
    for j in range(1, maxiter + 1):
        p *= be; p += r
        s *= be; s += w
        x += al * p; r -= al * s
        A(w, r)
        gn = r @ r; dn = r @ w
        eps = np.sqrt(gn) / nb
        done, kmin = stop(j, eps, eps0, kmin, tol, kmax)
        if done:
            return j, x
        be = gn / gam; al = gn / (dn - be * gn / al); gam = gn
    return maxiter, x


def cg_pipe(A, b, tol, maxiter, kmax=1e12):
    """Ghysels & Vanroose: gamma = (r,r), delta = (w,r) reduced while q = A w is computed."""
    x = np.zeros_like(b); r = b.copy(); w = np.zeros_like(b); A(w, r)
    p = np.zeros_like(b); s = np.zeros_like(b); z = np.zeros_like(b); q = np.zeros_like(b)
    nb = np.sqrt(b @ b); eps0 = np.sqrt(r @ r) / nb; kmin = 0.0
    gam_old = al_old = 1.0
    for j in range(1, maxiter + 1):
        gam = r @ r; dl = w @ r
        A(q, w)
        if j > 1:
            eps = np.sqrt(gam) / nb      # residual after j-1 iterations
            done, kmin = stop(j - 1, eps, eps0, kmin, tol, kmax)
            if done:
                return j - 1, x
            be = gam / gam_old; al = gam / (dl - be * gam / al_old)
        else:
            be = 0.0; al = gam / dl
        z *= be; z += q
        s *= be; s += w
        p *= be; p += r
        x += al * p; r -= al * s; w -= al * z
        gam_old, al_old = gam, al
    return maxiter, x


def main(names):
    for name in names:
        geom, Ls, beta, dtau = CFG[name]
        om, rng = oracle_holstein(geom, Ls, beta, dtau, mu=-1.0, seed=1234, eps=0.3)
        c = CRef(om, native=True)
        A = c.mulMTM
        for trial in range(2):
            g = rng.normal(size=om.Ndim)
            b = np.zeros(om.Ndim); om.mulMT(b, g) if trial == 0 else b.__setitem__(slice(None), g)
            out = []
            for f in (cg_ref, cg_cgear, cg_pipe):
                t0 = time.time()
                it, x = f(A, b, om.tol, om.maxiter)
                chk = np.zeros_like(b); A(chk, x)
                out.append((f.__name__, it, np.linalg.norm(chk - b) / np.linalg.norm(b), time.time() - t0))
            print(name, "rhs", "M^T g" if trial == 0 else "g", " | ".join(f"{n}: {it} it, true res {tr:.3e} ({dt:.0f}s)" for n, it, tr, dt in out),
                  flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["A", "B"])
