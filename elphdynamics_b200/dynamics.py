"""Host-side mirror of the reference's transforms, action and Langevin dynamics API:
src/TimeFreqFFTs.jl, src/FourierAcceleration.jl, src/PhononAction.jl, src/LangevinDynamics.jl.

All arithmetic runs in libelph_b200.so.  Randomness is injected by the caller in the order
the reference draws it from ``model.rng`` (eta, then per force evaluation g and the 2*Nsites
Arnoldi start values; the reference's wasted ``randn!(rng, R)`` draws are the caller's business).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._lib import SolveInfo, out_ptr, ptr
from .models import AbstractModel, _f64

EULER, RUNGE_KUTTA, HEUN = 1, 2, 3


# ------------------------------------------------------------------------------ TimeFreqFFTs
class TimeFreqFFT:
    """``TimeFreqFFT(lattice, L)`` (src/TimeFreqFFTs.jl:9-50): the plan lives in the engine."""

    def __init__(self, model: AbstractModel):
        self.model = model
        self.N, self.L = model.Nsites, model.Ltau


def tau_to_omega_(vout, op: TimeFreqFFT, vin):
    """``τ_to_ω!(vout::complex, op, vin::real)`` (src/TimeFreqFFTs.jl:55-73)."""
    if not (isinstance(vout, np.ndarray) and vout.dtype == np.complex128 and vout.flags["C_CONTIGUOUS"] and vout.size == op.model.Ndim):
        raise ValueError(f"vout: expected a C-contiguous complex128 array of {op.model.Ndim} entries")
    op.model._call("elph_tau_to_omega", ptr(_f64(vin, op.model.Ndim, "vin")), vout.ctypes.data_as(C.POINTER(C.c_double)))


def omega_to_tau_(vout, op: TimeFreqFFT, vin):
    """``ω_to_τ!(vout::real, op, vin::complex)`` (src/TimeFreqFFTs.jl:112-130)."""
    vin = np.ascontiguousarray(vin, dtype=np.complex128)
    if vin.size != op.model.Ndim:
        raise ValueError(f"vin: expected {op.model.Ndim} entries, got {vin.size}")
    op.model._call("elph_omega_to_tau", vin.ctypes.data_as(C.POINTER(C.c_double)), out_ptr(vout, op.model.Ndim, "vout"))


# ------------------------------------------------------------------------------ FourierAcceleration
class FourierAccelerator:
    """``FourierAccelerator(model)`` (src/FourierAcceleration.jl:11-82).  ``Q``/``M`` are built on the
    host exactly like ``update_Q!``/``update_M!`` (:149-266, a once-per-run table) and uploaded."""

    def __init__(self, model: AbstractModel):
        self.model = model
        self.N, self.L = model.Nph, model.Ltau
        self.Q = np.zeros(self.N * self.L)
        self.M = np.zeros(self.N * self.L)

    def _upload(self):
        self.model._call("elph_set_fourier_acceleration", ptr(self.Q), ptr(self.M))


def element_Qi(k, omega, dtau, m, L):
    """src/FourierAcceleration.jl:213-217."""
    return (m ** 2 + dtau * omega * omega + 4.0 / dtau) / (m ** 2 + dtau * omega * omega + (2 - 2 * np.cos(2 * np.pi * k / L)) / dtau)


def element_Mi(k, omega, dtau, m0, c, L):
    """src/FourierAcceleration.jl:260-266."""
    kp = np.minimum(k, L - k)
    m = m0 * np.exp(-(c * kp / L) ** 2)
    return dtau * (m ** 2 + omega ** 2 + (2 - 2 * np.cos(2 * np.pi * kp / L)) / dtau ** 2) / (m ** 2 + omega ** 2)


def update_Q_(fa: FourierAccelerator, model, omega_min, omega_max, m):
    """``update_Q!`` (src/FourierAcceleration.jl:149-155,172-193)."""
    k = np.arange(fa.L)
    Q = fa.Q.reshape(fa.N, fa.L)
    for ph in np.nonzero((model.omega > omega_min) & (model.omega < omega_max))[0]:
        Q[ph] = element_Qi(k, model.omega[ph], model.dtau, m, fa.L)
    fa._upload()


def update_M_(fa: FourierAccelerator, model, omega_min, omega_max, m0, c=0.0):
    """``update_M!`` (src/FourierAcceleration.jl:161-167,223-240)."""
    k = np.arange(fa.L)
    M = fa.M.reshape(fa.N, fa.L)
    for ph in np.nonzero((model.omega > omega_min) & (model.omega < omega_max))[0]:
        M[ph] = element_Mi(k, model.omega[ph], model.dtau, m0, c, fa.L)
    fa._upload()


def fourier_accelerate_(vout, fa: FourierAccelerator, v, power: float, use_mass: bool = False):
    """``fourier_accelerate!(v', fa, v, power; use_mass)`` real -> real (src/FourierAcceleration.jl:131-137)."""
    fa.model._call("elph_fourier_accelerate", ptr(_f64(v, fa.N * fa.L, "v")), out_ptr(vout, fa.N * fa.L, "vout"), float(power), 1 if use_mass else 0)


# ------------------------------------------------------------------------------ PhononAction
def calc_Sb(model, shifted: bool = False) -> float:
    """``calc_Sb(model, shifted)`` (src/PhononAction.jl:11-107)."""
    out = C.c_double()
    model._call("elph_Sb", 1 if shifted else 0, C.byref(out))
    return out.value


def calc_dSbdx_(dSbdx, model, shifted: bool = False):
    """``calc_dSbdx!(dSbdx, model, shifted)`` -- accumulates (src/PhononAction.jl:114-233)."""
    model._call("elph_dSbdx", 1 if shifted else 0, out_ptr(dSbdx, model.Ndof, "dSbdx"))


# ------------------------------------------------------------------------------ LangevinDynamics
class Dynamics:
    method = 0

    def __init__(self, model: AbstractModel, dt: float):
        self.Ndof, self.Ndim, self.dt = model.Ndof, model.Ndim, float(dt)
        self.info1 = SolveInfo()
        self.info2 = SolveInfo()


class EulerDynamics(Dynamics):
    """src/LangevinDynamics.jl:25-79."""
    method = EULER


class RungeKuttaDynamics(Dynamics):
    """src/LangevinDynamics.jl:135-160."""
    method = RUNGE_KUTTA


class HeunsDynamics(Dynamics):
    """src/LangevinDynamics.jl:245-270."""
    method = HEUN


def _use_p(P):
    return 0 if (P is None or getattr(P, "is_identity", False)) else 1


def calc_dSdx_(dSdx, g, Minv_g, model, P=None, arnoldi_noise=None):
    """``calc_dSdx!(dSdx, g, M⁻¹g, model, P)`` (src/LangevinDynamics.jl:334-345) with ``g`` injected.
    Returns the iteration count."""
    info = SolveInfo()
    an = None if arnoldi_noise is None else ptr(_f64(arnoldi_noise, 2 * model.Nsites, "arnoldi_noise"))
    model._call("elph_calc_dSdx", ptr(_f64(g, model.Ndim, "g")), an, _use_p(P), out_ptr(dSdx, model.Ndof, "dSdx"),
                None if Minv_g is None else out_ptr(Minv_g, model.Ndim, "Minv_g"), C.byref(info))
    model.last_solve_info = info
    return int(info.iters)


def evolve_(model, dyn: Dynamics, fa: FourierAccelerator, P=None, *, eta, g1, g2=None, arnoldi1=None, arnoldi2=None) -> int:
    """``evolve!(model, dyn, fa, P)`` (src/LangevinDynamics.jl:81,162,272) with injected noise.
    The phonon field stays on the device; read it back with ``model.x``."""
    it = C.c_int64()
    n2 = 2 * model.Nsites
    a1 = None if arnoldi1 is None else ptr(_f64(arnoldi1, n2, "arnoldi1"))
    a2 = None if arnoldi2 is None else ptr(_f64(arnoldi2, n2, "arnoldi2"))
    model._call("elph_langevin_step", dyn.method, dyn.dt, ptr(_f64(eta, model.Ndof, "eta")), ptr(_f64(g1, model.Ndim, "g1")),
                None if g2 is None else ptr(_f64(g2, model.Ndim, "g2")), a1, a2, _use_p(P), C.byref(it),
                C.byref(dyn.info1), C.byref(dyn.info2))
    return int(it.value)
