"""Oracle: KPM (Chebyshev) preconditioner for M^T M.  TEST INFRASTRUCTURE ONLY.

Follows ``src/KPMPreconditioners.jl``:
  * ``KPMExpansion`` constructor :101-146 (lambda_lo=0, lambda_hi=2, phi_w=2pi/L (w+1/2))
  * ``setup!`` :269-321 (activity test, buffer, hysteresis via ``isapprox(rtol=buf)``)
  * ``update_A!`` :332-350 (Holstein tau-mean of expnV), :355-381 (SSH tau-mean of cosh/sinh)
  * ``mul!``/``ldiv!`` on the expansion (A v, A^-1 v) :387-420
  * preconditioner apply ``ldiv!`` :426-481
  * ``mul!(::SymmetricKPMPreconditioner)`` :606-679, ``mulA'!`` :685-693, ``mulA!`` :758-778
  * ``kpm_coefficients!`` :789-839, ``scalar_invM`` :948-951
  * ``arnoldi_eigenvalue_bounds!`` :845-942

Third-party arithmetic absent from /root/reference: ``FFTW.dct!`` (orthonormal
DCT-II, un-normalised again at :812-813 -> plain cosine sums, restated
explicitly here) and LAPACK ``eigvals!`` on the <=20x20 Hessenberg matrix
(``numpy.linalg.eigvals``, the same ``geev``).

Randomness: the reference draws the 2N Arnoldi start values from ``model.rng``
(:859-861, :902-904).  Here they are INJECTED (``arnoldi_noise``) so that engine
and oracle consume identical numbers.
"""
from __future__ import annotations

import math

import numpy as np

from . import checkerboard as cb
from .fourier import TimeFreqFFT


def isapprox(x, y, rtol):
    """Julia ``isapprox(x, y; rtol)`` with atol = 0."""
    return x == y or abs(x - y) <= rtol * max(abs(x), abs(y))


def scalar_invM(x, phi):
    """src/KPMPreconditioners.jl:948-951."""
    return 1.0 / (1.0 - np.exp(-1j * phi) * x)


def kpm_coefficients(order: int, lam_lo: float, lam_hi: float, phi: float) -> np.ndarray:
    """src/KPMPreconditioners.jl:789-839.  c_0 = S_0/(2M), c_m = 2 S_m/(2M),
    S_m = sum_n f(x_n) cos(pi m (n+1/2)/(2M)), n = 0..2M-1."""
    M = order
    NM = 2 * M
    lam_avg = (lam_hi + lam_lo) / 2
    lam_mag = (lam_hi - lam_lo) / 2
    n = np.arange(NM)
    xn = lam_mag * np.cos(np.pi * (n + 0.5) / NM) + lam_avg
    f = scalar_invM(xn, phi)
    m = np.arange(M)
    C = np.cos(np.pi * np.outer(m, n + 0.5) / NM)          # (M, NM)
    S = C @ f
    c = 2.0 * S / NM
    c[0] = S[0] / NM
    return c.astype(np.complex128)


class KPMPreconditioner:
    """``SymmetricKPMPreconditioner`` + its ``KPMExpansion``."""
    is_identity = False

    def __init__(self, model, n: int = 20, buf: float = 0.05, c1: float = 1.0, c2: float = 1.0):
        self.model = model
        N, L = model.N, model.L
        self.N, self.L = N, L
        self.Lo2 = -(-L // 2)                      # cld(L,2)
        self.fft = TimeFreqFFT(N, L)
        self.buf, self.c1, self.c2 = float(buf), float(c1), float(c2)
        self.lam_lo, self.lam_hi = 0.0, 2.0
        self.lam_avg = (self.lam_hi + self.lam_lo) / 2
        self.lam_mag = (self.lam_hi - self.lam_lo) / 2
        self.phis = 2 * np.pi / L * (np.arange(self.Lo2) + 0.5)
        self.order = np.ones(self.Lo2, dtype=np.int64)
        self.coeff = [np.zeros(1, dtype=np.complex128) for _ in range(self.Lo2)]
        self.expnVbar = np.zeros(N)
        self.coshbar = np.zeros(model.Nbonds)
        self.sinhbar = np.zeros(model.Nbonds)
        if model.kind == "holstein":
            self.coshbar[:] = model.cosht
            self.sinhbar[:] = model.sinht
        elif model.kind == "ssh":
            self.expnVbar[:] = model.expmu
        else:
            raise TypeError(model.kind)
        self.n = min(int(n), N)
        self.active = True
        self.e_min = self.e_max = float("nan")
        self.checkerboard_count = 0
        self.recomputed = False

    # ---------------------------------------------------------------- A, A^-1
    def update_A(self):
        m = self.model
        if m.kind == "holstein":
            # src/KPMPreconditioners.jl:332-350: running sum over tau then / L
            E = m.expnV.reshape(m.N, m.L)
            acc = np.zeros(m.N)
            for tau in range(m.L):
                acc += E[:, tau]
            self.expnVbar[:] = acc / m.L
        else:
            # src/KPMPreconditioners.jl:355-381
            accc = np.zeros(m.Nbonds)
            accs = np.zeros(m.Nbonds)
            for tau in range(m.L):
                accc += m.cosht[:, tau]
                accs += m.sinht[:, tau]
            self.coshbar[:] = accc / m.L
            self.sinhbar[:] = accs / m.L
            self.expnVbar[:] = m.expmu

    def mulA(self, v, transposed=False):
        """A = cb(cbar,sbar) diag(eVbar); A^T = diag(eVbar) cb^T   (:758-778)."""
        m = self.model
        self.checkerboard_count += 1
        if transposed:
            out = v.copy()
            cb.checkerboard_transpose_mul(out, m.neighbor_table, self.coshbar, self.sinhbar, m.group_offsets)
            out *= self.expnVbar
            return out
        out = self.expnVbar * v
        cb.checkerboard_mul(out, m.neighbor_table, self.coshbar, self.sinhbar, m.group_offsets)
        return out

    def ldivA(self, v):
        """A^-1 v = (cb^-1 v) ./ eVbar   (:406-420)."""
        m = self.model
        out = v.copy()
        cb.checkerboard_inverse_mul(out, m.neighbor_table, self.coshbar, self.sinhbar, m.group_offsets)
        out /= self.expnVbar
        return out

    # ----------------------------------------------------------------- Arnoldi
    def _arnoldi(self, start, op):
        """One half of ``arnoldi_eigenvalue_bounds!`` (:845-942)."""
        n, m = self.n, self.N
        Q = np.zeros((m, n + 1))
        h = np.zeros((n + 1, n))
        b = np.asarray(start, dtype=np.float64).copy()
        b /= np.linalg.norm(b)
        Q[:, 0] = b
        l = n
        for k in range(n):
            v = op(b)
            for j in range(k + 1):
                h[j, k] = float(np.dot(Q[:, j], v))
                v = v - h[j, k] * Q[:, j]
            h[k + 1, k] = float(np.linalg.norm(v))
            if h[k + 1, k] > 1e-12:
                b = v / h[k + 1, k]
                Q[:, k + 1] = b
            else:
                l = k + 1
                break
        hh = h[:l, :l]
        if np.all(np.isfinite(hh)):
            return float(np.max(np.real(np.linalg.eigvals(hh)))), hh.copy()
        return float("inf"), hh.copy()

    def arnoldi_eigenvalue_bounds(self, arnoldi_noise):
        noise = np.asarray(arnoldi_noise, dtype=np.float64)
        assert noise.size == 2 * self.N
        e_max, self.h_max = self._arnoldi(noise[:self.N], lambda b: self.mulA(b))
        inv_max, self.h_min = self._arnoldi(noise[self.N:], lambda b: self.ldivA(b))
        e_min = 1.0 / inv_max if math.isfinite(inv_max) else -float("inf")
        return e_min, e_max

    # ------------------------------------------------------------------- setup
    def setup(self, arnoldi_noise):
        """src/KPMPreconditioners.jl:269-321."""
        self.update_A()
        e_min, e_max = self.arnoldi_eigenvalue_bounds(arnoldi_noise)
        self.e_min, self.e_max = e_min, e_max
        self.recomputed = False
        if (0.0 < e_min < 1.0) and (1.0 < e_max) and (e_max - e_min) < 2.0:
            lam_lo = max(0.0, (1 - 2 * self.buf) * e_min)
            lam_hi = (1 + 2 * self.buf) * e_max
            if (not isapprox(lam_lo, self.lam_lo, self.buf)) or (not isapprox(lam_hi, self.lam_hi, self.buf)):
                self.lam_lo, self.lam_hi = lam_lo, lam_hi
                self.lam_avg = (lam_hi + lam_lo) / 2
                self.lam_mag = (lam_hi - lam_lo) / 2
                for w in range(self.Lo2):
                    phi = self.phis[w]
                    order = int(math.floor((lam_hi - lam_lo) * (self.c1 / phi + self.c2)))
                    order = max(1, order)
                    self.order[w] = order
                    self.coeff[w] = kpm_coefficients(order, lam_lo, lam_hi, phi)
                self.recomputed = True
            self.active = True
        else:
            self.active = False

    # ------------------------------------------------------------------- apply
    def _mulAprime(self, v, transposed):
        """src/KPMPreconditioners.jl:685-693."""
        return (1 / self.lam_mag) * self.mulA(v, transposed) - (self.lam_avg / self.lam_mag) * v

    def _poly(self, v, c, transposed, conj):
        """sum_m c_m T_m(A') v by the three-term recurrence (:625-676)."""
        order = len(c)
        cc = np.conj(c) if conj else c
        out = cc[0] * v
        if order > 1:
            u_prev = None
            u_n = v.copy()
            u_next = self._mulAprime(u_n, transposed)
            n = 1
            while True:
                n += 1
                u_prev, u_n = u_n, u_next
                out = out + cc[n - 1] * u_n
                if n == order:
                    break
                u_next = self._mulAprime(u_n, transposed)
                u_next = 2 * u_next - u_prev
        return out

    def mul_block(self, w: int, v):
        """``mul!(v', P::SymmetricKPMPreconditioner, v)`` for frequency ``w`` (0-based), :606-679:
        first M^-T[w,w] (conjugated coefficients, A'^T), then M^-1[w,w]."""
        c = self.coeff[w]
        out = self._poly(v, c, transposed=True, conj=True)
        out = self._poly(out, c, transposed=False, conj=False)
        return out

    def ldiv(self, vout, vin):
        """Apply the preconditioner, src/KPMPreconditioners.jl:426-481."""
        self.checkerboard_count = 0
        if not self.active:
            vout[:] = vin
            return
        N, L = self.N, self.L
        nu = self.fft.tau_to_omega(vin).reshape(N, L)        # a2[tau->omega, i]
        a1T = np.ascontiguousarray(nu.T)                     # (L, N): [omega][site]
        a2T = np.zeros((L, N), dtype=np.complex128)
        for w in range(self.Lo2):
            a2T[w] = self.mul_block(w, a1T[w])
            a2T[L - 1 - w] = np.conj(a2T[w])                 # :464-466 (also for the odd-L middle frequency)
        v1 = np.ascontiguousarray(a2T.T).reshape(-1)
        vout[:] = self.fft.omega_to_tau_real(v1)

    # dense debug: block w of the tau-averaged operator, cf. construct_Bbar :953-991
    def construct_Abar(self):
        A = np.zeros((self.N, self.N))
        for col in range(self.N):
            e = np.zeros(self.N)
            e[col] = 1.0
            A[:, col] = self.mulA(e)
        return A
