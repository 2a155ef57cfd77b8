"""The oracle's Holstein model and KPM preconditioner with their hot loops in C (oracle/c/elph_ref.c) -- the CPU baseline of the
second headline quantity, Langevin steps/s.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The orchestration is the NumPy oracle's own, unchanged (oracle/langevin.py, oracle/solvers.py, oracle/kpm.py setup!,
oracle/fourier.py): only the loops the reference spends its time in are replaced by their C restatement --
``mulM!/mulMT!/mulMTM!`` (src/HolsteinModels.jl:569-684), ``muldMdx!`` (:691-755), the tau-averaged ``A``, ``A^T``, ``A^-1``
(src/KPMPreconditioners.jl:387-420, 758-778) and the Chebyshev recurrences of the preconditioner blocks (:606-693).  What
the reference itself hands to libraries stays with libraries here as well: the tau-FFTs (FFTW there, pocketfft through
``numpy.fft`` here), the <= 20 x 20 eigenvalue problem of the Arnoldi bounds (LAPACK both) and the coefficient DCT.
``tests/test_oracle_c.py`` checks every replaced piece and a whole Runge-Kutta step against the NumPy oracle.
"""
from __future__ import annotations

import ctypes as C
import types

import numpy as np

from . import cref
from .kpm import KPMPreconditioner

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _bind(lib):
    mp = C.POINTER(cref.RefModel)
    lib.ref_muldMdx.argtypes = [_dp, mp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp]
    lib.ref_mulA.argtypes = [_dp, mp, _dp, _dp, _dp, _dp, C.c_int]
    lib.ref_kpm_blocks.argtypes = [mp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int64, _ip, _ip, _dp, _dp, _dp, C.c_int]
    lib.ref_kpm_blocks.restype = C.c_int
    lib.ref_mulMTM_mt.argtypes = [_dp, mp, _dp, _dp, _dp, C.c_int]
    return lib


def accelerate_model(om, native: bool = True, nthreads: int = 1):
    """Replace the operator loops of an oracle Holstein model by the C restatement, in place; returns the model.
    ``nthreads`` > 1 threads one product along tau (the reference itself is single-threaded, src/ElPhDynamics.jl:74-75)."""
    if om.kind != "holstein":
        raise TypeError("the C restatement covers the Holstein model")
    c = cref.CRef(om, native=native)
    _bind(c.lib)
    c.expnV = om.expnV                      # share the table the oracle's update_model writes (same array object)
    scratch = np.zeros(om.Ndim)
    ref = C.byref(c.m)
    lam, lam2 = np.ascontiguousarray(om.lam, dtype=np.float64), np.ascontiguousarray(om.lam2, dtype=np.float64)

    def mulM(self, y, v):
        c.lib.ref_mulM(_p(y), ref, _p(self.expnV), _p(np.ascontiguousarray(v)))

    def mulMT(self, y, v):
        c.lib.ref_mulMT(_p(y), ref, _p(self.expnV), _p(np.ascontiguousarray(v)))

    def mulMTM(self, y, v):
        if nthreads > 1:
            c.lib.ref_mulMTM_mt(_p(y), ref, _p(self.expnV), _p(np.ascontiguousarray(v)), _p(scratch), nthreads)
        else:
            c.lib.ref_mulMTM(_p(y), ref, _p(self.expnV), _p(np.ascontiguousarray(v)), _p(scratch))

    def mul(self, y, v):
        mulMTM(self, y, v)

    def muldMdx(self, dMdx, u, v):
        c.lib.ref_muldMdx(_p(dMdx), ref, _p(self.expnV), _p(self.x), _p(lam), _p(lam2), self.dtau, _p(np.ascontiguousarray(u)),
                          _p(np.ascontiguousarray(v)), _p(scratch))

    for f in (mulM, mulMT, mulMTM, mul, muldMdx):
        setattr(om, f.__name__, types.MethodType(f, om))
    om._cfast = c                           # keeps the ctypes structures alive
    return om


class FastKPM(KPMPreconditioner):
    """``SymmetricKPMPreconditioner`` with ``A``, ``A^T``, ``A^-1`` and the frequency blocks in C; set-up logic, FFTs and
    coefficients are the parent's."""

    def __init__(self, model, *args, nthreads: int = 1, native: bool = True, **kw):
        super().__init__(model, *args, **kw)
        self._c = getattr(model, "_cfast", None) or cref.CRef(model, native=native)
        _bind(self._c.lib)
        self.nthreads = int(nthreads)

    def _A(self, v, mode):
        out = np.empty(self.N)
        self._c.lib.ref_mulA(_p(out), C.byref(self._c.m), _p(self.expnVbar), _p(self.coshbar), _p(self.sinhbar),
                             _p(np.ascontiguousarray(v, dtype=np.float64)), mode)
        return out

    def mulA(self, v, transposed=False):
        self.checkerboard_count += 1
        return self._A(v, 1 if transposed else 0)

    def ldivA(self, v):
        return self._A(v, 2)

    def ldiv(self, vout, vin):
        if not self.active:
            vout[:] = vin
            return
        N, L = self.N, self.L
        # the transforms stay with pocketfft (numpy.fft), 2.8 us per length-200 column; a plain recursive C FFT written for this
        # file was 7 times slower and was dropped again.  The reference uses planned FFTW transforms here.
        nu = self.fft.tau_to_omega(vin).reshape(N, L)
        a1T = np.ascontiguousarray(nu.T)                          # [omega][site]
        a2T = np.zeros((L, N), dtype=np.complex128)
        order = np.ascontiguousarray(self.order, dtype=np.int64)
        off = np.zeros(self.Lo2, dtype=np.int64)
        off[1:] = np.cumsum(order)[:-1]
        coeff = np.ascontiguousarray(np.concatenate(self.coeff))
        self._c.lib.ref_kpm_blocks(C.byref(self._c.m), _p(self.expnVbar), _p(self.coshbar), _p(self.sinhbar), self.lam_avg, self.lam_mag,
                                   self.Lo2, order.ctypes.data_as(_ip), off.ctypes.data_as(_ip), coeff.view(np.float64).ctypes.data_as(_dp),
                                   a1T.view(np.float64).ctypes.data_as(_dp), a2T.view(np.float64).ctypes.data_as(_dp), self.nthreads)
        a2T[L - self.Lo2:] = np.conj(a2T[:self.Lo2][::-1])          # a2T[L-1-w] = conj(a2T[w]), src/KPMPreconditioners.jl:464-466
        vout[:] = self.fft.omega_to_tau_real(np.ascontiguousarray(a2T.T).reshape(-1))


# ---- CPU baseline of Langevin steps/s (bench.py's cpu_baseline / --impl reference legs) ---------------------------------------
def _chain(args):
    """One independent Markov chain (the reference's own scale-out: one process per chain id, src/ElPhDynamics.jl:90-95)."""
    import time
    seed, nsteps, Ls, beta, dtau = args
    import os
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (here, os.path.join(here, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from helpers import oracle_holstein
    from oracle import langevin as olang
    from oracle.fourier import FourierAccelerator
    from oracle.solvers import ConjugateGradient
    om, _ = oracle_holstein("square", Ls, beta, dtau, mu=-1.0, seed=1234, eps=0.3)
    accelerate_model(om, native=True, nthreads=1)
    P = FastKPM(om, nthreads=1)
    fo = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
    fo.update_Q(0.0, 10.0, 1.0)
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    rng = np.random.default_rng(seed)
    draw = lambda: (rng.normal(size=om.Ndof), rng.normal(size=om.Ndim), rng.normal(size=om.Ndim), rng.normal(size=2 * om.N),
                    rng.normal(size=2 * om.N))
    olang.evolve_rk(om, cg, fo, P, 1e-3, *draw())          # warm-up: first-touch, library load
    its = []
    t0 = time.perf_counter()
    for _ in range(nsteps):
        its.append(int(olang.evolve_rk(om, cg, fo, P, 1e-3, *draw())))
    return time.perf_counter() - t0, its


def langevin_baseline(chains: int = 0, nsteps: int = 2, Ls: int = 32, beta: float = 20.0, dtau: float = 0.1):
    """Runge-Kutta steps/s of the KPM-preconditioned Langevin update at the named lattice: one chain on one thread (the
    reference's setting, src/ElPhDynamics.jl:74-75) and ``chains`` independent chains on as many cores."""
    import multiprocessing as mp
    import os
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    secs1, its1 = _chain((4321, nsteps, Ls, beta, dtau))
    out = {"steps_per_s": nsteps / secs1, "cores": 1, "pcg_iters_second_solve": its1,
           "kind": "port (C restatement of the Julia loops; FFTs through pocketfft, the 20x20 eigenvalue problem through LAPACK)",
           "sample": f"{nsteps} Runge-Kutta steps of {Ls}x{Ls}xL{int(round(beta / dtau))} with KPM-preconditioned CG, fresh injected noise per step"}
    chains = chains or (os.cpu_count() or 1)
    if chains > 1:
        with mp.get_context("fork").Pool(chains) as pool:
            import time
            t0 = time.perf_counter()
            res = pool.map(_chain, [(4321 + c, nsteps, Ls, beta, dtau) for c in range(chains)])
            wall = time.perf_counter() - t0
        slowest = max(r[0] for r in res)
        out["all_cores"] = {"chains": chains, "steps_per_s_aggregate": chains * nsteps / slowest, "cores": chains,
                            "wall_seconds_including_set_up": wall,
                            "note": "independent chains, one single-threaded process each (the reference's own scale-out)"}
    return out


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    print(json.dumps(langevin_baseline(a.chains, a.steps)))
