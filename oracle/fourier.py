"""Oracle: tau<->omega transforms and Fourier acceleration.  TEST INFRASTRUCTURE ONLY.

Follows:
  * ``src/TimeFreqFFTs.jl:32-45``   Theta[tau] = exp(-i*pi*tau/L) (0-based tau), FFT along tau
  * ``src/TimeFreqFFTs.jl:55-73``   tau_to_omega!   nu[:,i] = FFT(Theta .* v[:,i])
  * ``src/TimeFreqFFTs.jl:92-130``  omega_to_tau!   v = conj(Theta) .* iFFT(nu)  (complex / real-part variants)
  * ``src/FourierAcceleration.jl:91-143``  fourier_accelerate!  (plain FFT, * Q^p or M^p, iFFT, real part)
  * ``src/FourierAcceleration.jl:149-266`` update_Q!/update_M!/element_Qi/element_Mi

Third-party arithmetic: FFTW.jl 1.3.2 / FFTW_jll 3.3.9+8 (``Manifest.toml:68-78``),
absent from /root/reference.  Its published conventions are restated with
``numpy.fft``: forward ``sum_j x_j exp(-2*pi*i*j*k/n)`` unnormalised, inverse
scaled by 1/n -- identical in ``numpy.fft.fft/ifft``.
"""
from __future__ import annotations

import numpy as np


class TimeFreqFFT:
    def __init__(self, N: int, L: int):
        self.N, self.L = N, L
        self.theta = np.exp(-1j * np.pi * np.arange(L) / L)      # src/TimeFreqFFTs.jl:37

    def tau_to_omega(self, vin):
        """Returns complex (N*L,) in the same site-major/tau-fastest layout."""
        u = vin.reshape(self.N, self.L)
        return np.fft.fft(self.theta[None, :] * u, axis=1).reshape(-1)

    def omega_to_tau(self, vin):
        """Complex output variant, src/TimeFreqFFTs.jl:92-110."""
        u = vin.reshape(self.N, self.L)
        return (np.conj(self.theta)[None, :] * np.fft.ifft(u, axis=1)).reshape(-1)

    def omega_to_tau_real(self, vin):
        """Real-part variant, src/TimeFreqFFTs.jl:112-130."""
        return np.real(self.omega_to_tau(vin))


def element_Qi(k, omega, dtau, m, L):
    """src/FourierAcceleration.jl:213-217."""
    return (m ** 2 + dtau * omega * omega + 4.0 / dtau) / (m ** 2 + dtau * omega * omega + (2 - 2 * np.cos(2 * np.pi * k / L)) / dtau)


def element_Mi(k, omega, dtau, m0, c, L):
    """src/FourierAcceleration.jl:260-266."""
    kp = np.minimum(k, L - k)
    m = m0 * np.exp(-(c * kp / L) ** 2)
    return dtau * (m ** 2 + omega ** 2 + (2 - 2 * np.cos(2 * np.pi * kp / L)) / dtau ** 2) / (m ** 2 + omega ** 2)


class FourierAccelerator:
    """src/FourierAcceleration.jl:11-82.  ``Q``/``M`` start at zero and are filled
    by ``update_Q``/``update_M`` for phonons with omega_min < omega < omega_max
    (``initialize_fourieraccelerator``, src/ProcessInputFile.jl:516-535)."""

    def __init__(self, Nph: int, L: int, dtau: float, omega):
        self.N, self.L, self.dtau = Nph, L, float(dtau)
        self.omega = np.asarray(omega, dtype=np.float64)
        self.Q = np.zeros(Nph * L)
        self.M = np.zeros(Nph * L)

    def update_Q(self, omega_min, omega_max, m):
        k = np.arange(self.L)
        Q = self.Q.reshape(self.N, self.L)
        for ph in range(self.N):
            if omega_min < self.omega[ph] < omega_max:
                Q[ph, :] = element_Qi(k, self.omega[ph], self.dtau, m, self.L)

    def update_M(self, omega_min, omega_max, m0, c=0.0):
        k = np.arange(self.L)
        M = self.M.reshape(self.N, self.L)
        for ph in range(self.N):
            if omega_min < self.omega[ph] < omega_max:
                M[ph, :] = element_Mi(k, self.omega[ph], self.dtau, m0, c, self.L)

    def accelerate(self, v, power: float, use_mass: bool = False):
        """Real in -> real out, src/FourierAcceleration.jl:131-137 via :91-115."""
        a = v.reshape(self.N, self.L).astype(np.complex128)
        u = np.fft.fft(a, axis=1)
        D = (self.M if use_mass else self.Q).reshape(self.N, self.L)
        u = u * D ** power
        return np.real(np.fft.ifft(u, axis=1)).reshape(-1)
