"""The engine's own eigenvalue routine for the Arnoldi Hessenberg matrices (csrc/kpm.cu: complex single-shift QR; replaces
LAPACK eigvals! at src/KPMPreconditioners.jl:891,935) against numpy.linalg.eigvals.  Host code only: no GPU needed."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("n", [1, 2, 3, 8, 20, 33])
def test_hessenberg_eigenvalues_against_lapack(n):
    from elphdynamics_b200 import _lib
    lib = _lib.load()
    f = lib.elph_debug_hessenberg_eigvals
    f.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    f.restype = C.c_int32
    rng = np.random.default_rng(n)
    for trial in range(6):
        H = np.triu(rng.normal(size=(n, n)), -1)
        if trial == 1 and n > 3:
            H[n // 2, n // 2 - 1] = 0.0               # a deflated block
        if trial == 2:
            H = np.triu(H + H.T, -1) + np.diag(np.arange(n, dtype=float))   # nearly real spectrum
        if trial == 3:
            # the shape arnoldi_eigenvalue_bounds! produces: projection of a positive operator with spectrum in (0, 2)
            A = np.diag(rng.uniform(0.1, 1.9, size=n)) + 0.01 * rng.normal(size=(n, n))
            Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
            from scipy.linalg import hessenberg
            H = hessenberg(Q.T @ A @ Q)
        H = np.ascontiguousarray(H)
        re, im = np.zeros(n), np.zeros(n)
        assert f(n, H.ctypes.data, re.ctypes.data, im.ctypes.data) == 0
        got = np.sort_complex(re + 1j * im)
        want = np.sort_complex(np.linalg.eigvals(H))
        scale = max(1.0, np.abs(want).max())
        # match the two spectra greedily (sorting complex conjugate pairs is not stable under rounding)
        left = list(want)
        for g in got:
            k = int(np.argmin([abs(g - w) for w in left]))
            assert abs(g - left[k]) <= 1e-9 * scale, (n, trial, g, left[k])
            left.pop(k)
        assert abs(got.real.max() - want.real.max()) <= 1e-10 * scale
