"""CPU restatement (test infrastructure) of the stochastic Green's-function estimator of the reference:
``EstimateGreensFunction`` / ``update!`` / ``setup!`` / ``convolve!`` / ``antiperiodic_copy!`` / ``periodic_product!``
(src/GreensFunctions.jl:23-188, 201-234, 239-296, 361-414, 420-463) -- SURVEY.md section 8(f) rank 3.

Arrays keep the reference's memory order: a Julia array of dimensions (2L, n, L1, L2, L3) is the NumPy C-ordered array of
shape (L3, L2, L1, n, 2L); the six-dimensional outputs (2L, n, n, L1, L2, L3) are (L3, L2, L1, n[s1], n[s2], 2L).
The FFTs are FFTW's (``plan_fft(a, (1,3,4,5))``: unnormalised forward, ``plan_ifft`` scaled by 1/size), restated with
``numpy.fft``.  Parity unpinned by the reference (it ships no tests); pinned here by a brute-force correlation sum in
tests/test_oracle_invariants.py.
"""
from __future__ import annotations

import math

import numpy as np

from .solvers import ldiv


class EstimateGreensFunction:
    """src/GreensFunctions.jl:23-188."""

    def __init__(self, model, nv: int = 2):
        self.nv = max(2, int(nv))
        lat = model.lat
        self.NL, self.L, self.N = model.Ndim, model.L, model.N
        self.L1, self.L2, self.L3, self.ns = lat.L1, lat.L2, lat.L3, lat.norbits
        self.R = np.zeros((self.nv, self.NL))
        self.MinvR = np.zeros((self.nv, self.NL))
        self.n1, self.n2 = 0, 1
        shape6 = (self.L3, self.L2, self.L1, self.ns, self.ns, 2 * self.L)
        self.G_D0 = np.zeros(shape6, dtype=complex)
        self.G_DD_G_00 = np.zeros(shape6, dtype=complex)
        self.G_D0_G_D0 = np.zeros(shape6, dtype=complex)
        self.G_D0_G_0D = np.zeros(shape6, dtype=complex)

    def _shape5(self):
        return (self.L3, self.L2, self.L1, self.ns, 2 * self.L)


def update(Gr: EstimateGreensFunction, model, cg, P, R, arnoldi_noise=None):
    """``update!(estimator, model, P)`` (:201-234) with the random vectors injected (R: (nv, NL))."""
    if P is not None and not getattr(P, "is_identity", False):
        P.setup(arnoldi_noise)
    infos = []
    for i in range(Gr.nv):
        r1 = np.asarray(R[i], dtype=np.float64)
        x = np.zeros(Gr.NL)
        b = np.zeros(Gr.NL)
        model.mulMT(b, r1)                                  # solve M^T M x = M^T r1  (:222-226)
        infos.append(ldiv(x, model, b, cg, P))
        Gr.R[i] = r1
        Gr.MinvR[i] = x
    return infos


def antiperiodic_copy(x, L):
    """y = [x(1..L), -x(1..L)] per site (:420-433); returns shape (N, 2L)."""
    X = np.asarray(x).reshape(-1, L)
    return np.concatenate([X, -X], axis=1)


def periodic_product(y, x, L):
    """z = [x.y (1..L), x.y (1..L)] per site (:439-457)."""
    Z = np.asarray(y).reshape(-1, L) * np.asarray(x).reshape(-1, L)
    return np.concatenate([Z, Z], axis=1)


def _reflect(B, axes):
    """B[n(k)] with n(k) = mod1(-k+2, len), i.e. 0-based (-k) mod len, along the given axes (:388-391)."""
    for ax in axes:
        B = np.roll(np.flip(B, axis=ax), 1, axis=ax)
    return B


def convolve(a, b, Gr: EstimateGreensFunction):
    """``convolve!`` (:361-414) without the accumulation: returns ab'' of shape (L3, L2, L1, n, n, 2L)."""
    A = np.asarray(a, dtype=complex).reshape(Gr._shape5())
    B = np.asarray(b, dtype=complex).reshape(Gr._shape5())
    axes = (0, 1, 2, 4)                                     # Julia dims (1,3,4,5): omega and the three cell axes
    Af = np.fft.fftn(A, axes=axes)
    Bf = _reflect(np.fft.fftn(B, axes=axes), axes)
    V = 2 * Gr.L * Gr.N / Gr.ns
    abp = Af[:, :, :, None, :, :] * Bf[:, :, :, :, None, :] / V     # [k3,k2,k1,s1,s2,w] = a'[w,s2,k] b'[nw,s1,nk] / V
    return np.fft.ifftn(abp, axes=(0, 1, 2, 5))             # Julia dims (1,4,5,6)


def setup(Gr: EstimateGreensFunction, n1: int, n2: int):
    """``setup!(estimator, n1, n2)`` (:239-296), 0-based vector indices."""
    L = Gr.L
    Gr.n1, Gr.n2 = n1, n2
    m1, r1, m2, r2 = Gr.MinvR[n1], Gr.R[n1], Gr.MinvR[n2], Gr.R[n2]
    s2 = math.sqrt(2.0)
    a = (antiperiodic_copy(m1, L) + antiperiodic_copy(m2, L)) / s2
    b = (antiperiodic_copy(r1, L) + antiperiodic_copy(r2, L)) / s2
    Gr.G_D0 = convolve(a, b, Gr)
    Gr.G_D0_G_D0 = convolve(periodic_product(m1, m2, L), periodic_product(r1, r2, L), Gr)
    Gr.G_DD_G_00 = convolve(periodic_product(m2, r2, L), periodic_product(m1, r1, L), Gr)
    Gr.G_D0_G_0D = convolve(periodic_product(m1, r2, L), periodic_product(m2, r1, L), Gr)
    return Gr.G_D0, Gr.G_D0_G_D0, Gr.G_DD_G_00, Gr.G_D0_G_0D


def measure(G, Gr: EstimateGreensFunction, l1, l2, l3, o1, o2, tau):
    """``measure_G...(estimator, l1, l2, l3, o1, o2, tau)`` (:301-345): G[mod1(tau+1, 2L), o2, o1, l1+1, l2+1, l3+1],
    orbitals 1-based as in the reference, everything else 0-based."""
    return G[l3, l2, l1, o1 - 1, o2 - 1, tau % (2 * Gr.L)]
