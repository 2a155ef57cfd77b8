"""ctypes access to the C restatement (oracle/c/elph_ref.c).  TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "c" / "elph_ref.c"
BUILD = HERE / "_build"


class RefModel(C.Structure):
    _fields_ = [("N", C.c_int64), ("L", C.c_int64), ("Nb", C.c_int64), ("nt", C.POINTER(C.c_int64)),
                ("cosht", C.POINTER(C.c_double)), ("sinht", C.POINTER(C.c_double))]


def build(native: bool = False) -> Path:
    """gcc -O3 -ffast-math -fopenmp; ``native`` = -march=native into a separate file (timing on this host)."""
    BUILD.mkdir(exist_ok=True)
    out = BUILD / ("libelph_ref_native.so" if native else "libelph_ref.so")
    if out.exists() and out.stat().st_mtime >= SRC.stat().st_mtime:
        return out
    arch = "native" if native else "x86-64-v3"
    cmd = ["gcc", "-O3", f"-march={arch}", "-ffast-math", "-fopenmp", "-fPIC", "-shared", "-o", str(out), str(SRC)]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def load(native: bool = False) -> C.CDLL:
    try:
        path = build(native)
    except Exception:
        path = BUILD / "libelph_ref.so"      # prebuilt portable copy travels with the snapshot
        if not path.exists():
            raise
    lib = C.CDLL(str(path))
    dp = C.POINTER(C.c_double)
    mp = C.POINTER(RefModel)
    lib.ref_mulM.argtypes = [dp, mp, dp, dp]
    lib.ref_mulMT.argtypes = [dp, mp, dp, dp]
    lib.ref_mulMTM.argtypes = [dp, mp, dp, dp, dp]
    lib.ref_cg.argtypes = [dp, mp, dp, dp, C.c_double, C.c_int64, C.c_double, dp, dp]
    lib.ref_cg.restype = C.c_int64
    lib.ref_cg_mt.argtypes = [dp, mp, dp, dp, C.c_double, C.c_int64, C.c_double, dp, dp, C.c_int]
    lib.ref_cg_mt.restype = C.c_int64
    lib.ref_mulMTM_replicas.argtypes = [mp, C.c_int64, C.c_int64, dp, dp, dp, dp, C.c_int]
    lib.ref_mulMTM_replicas.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CRef:
    """The oracle Holstein model's operator through the C restatement."""

    def __init__(self, om, native: bool = False):
        self.lib = load(native)
        self.N, self.L, self.n = om.N, om.L, om.Ndim
        self._nt = np.ascontiguousarray(om.neighbor_table.T, dtype=np.int64)
        self._c = np.ascontiguousarray(om.cosht)
        self._s = np.ascontiguousarray(om.sinht)
        self.m = RefModel(om.N, om.L, om.Nbonds, self._nt.ctypes.data_as(C.POINTER(C.c_int64)), _p(self._c), _p(self._s))
        self.expnV = np.ascontiguousarray(om.expnV)
        self.scratch = np.zeros(self.n)

    def mulM(self, y, v):
        self.lib.ref_mulM(_p(y), C.byref(self.m), _p(self.expnV), _p(np.ascontiguousarray(v)))

    def mulMT(self, y, v):
        self.lib.ref_mulMT(_p(y), C.byref(self.m), _p(self.expnV), _p(np.ascontiguousarray(v)))

    def mulMTM(self, y, v):
        self.lib.ref_mulMTM(_p(y), C.byref(self.m), _p(self.expnV), _p(np.ascontiguousarray(v)), _p(self.scratch))

    def cg(self, x, b, tol=1e-5, maxiter=10000, kappa_max=1e12):
        work = np.zeros(4 * self.n)
        eps = C.c_double()
        it = self.lib.ref_cg(_p(x), C.byref(self.m), _p(self.expnV), _p(np.ascontiguousarray(b)), tol, maxiter, kappa_max,
                             _p(work), C.byref(eps))
        return int(it), eps.value

    def cg_mt(self, x, b, tol=1e-5, maxiter=10000, kappa_max=1e12, nthreads=0):
        """The same recurrences with the loops of one product threaded along tau (full-size parity tests)."""
        work = np.zeros(4 * self.n)
        eps = C.c_double()
        it = self.lib.ref_cg_mt(_p(x), C.byref(self.m), _p(self.expnV), _p(np.ascontiguousarray(b)), tol, maxiter, kappa_max,
                                _p(work), C.byref(eps), nthreads)
        return int(it), eps.value

    def mulMTM_throughput(self, nrep: int, reps: int, nthreads: int = 0, seed: int = 0):
        """Times nrep independent replicas x reps products; returns (seconds, threads)."""
        import time
        rng = np.random.default_rng(seed)
        v = rng.normal(size=(nrep, self.n))
        y = np.zeros_like(v)
        e = np.tile(self.expnV, (nrep, 1))
        s = np.zeros_like(v)
        self.lib.ref_mulMTM_replicas(C.byref(self.m), nrep, 1, _p(e), _p(v), _p(y), _p(s), nthreads)  # warm-up
        t0 = time.perf_counter()
        used = self.lib.ref_mulMTM_replicas(C.byref(self.m), nrep, reps, _p(e), _p(v), _p(y), _p(s), nthreads)
        return time.perf_counter() - t0, int(used)
