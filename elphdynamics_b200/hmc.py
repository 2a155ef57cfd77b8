"""Host-side mirror of ``src/HMC.jl`` on top of libelph_b200.so.

``update_(model, hmc, fa, P, ...)`` runs one whole HMC trajectory on the device (leapfrog or the
multi-timestep integrator); the Metropolis decision uses the injected uniform exactly like
``rand(model.rng) < P`` in the reference.  All noise is injected by the caller.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ptr
from .models import AbstractModel, _f64

_IDS = {"v": 0, "phi_plus": 1, "phi_minus": 2, "Lphi_plus": 3, "Lphi_minus": 4, "O_plus": 5, "O_minus": 6, "Lambda": 7, "dSdx": 8}


class HybridMonteCarlo:
    """``HybridMonteCarlo(model, dt, tr, alpha, Nb)`` (src/HMC.jl:188-233): Nt = round(tr/dt), dt' = dt/Nb."""

    def __init__(self, model: AbstractModel, dt: float, tr: float, alpha: float = 0.0, Nb: int = 1):
        if not 0.0 <= alpha < 1.0:
            raise ValueError("alpha must be in [0, 1)")
        self.model = model
        self.Ndof, self.Ndim = model.Ndof, model.Ndim
        self.dt, self.tr, self.alpha, self.Nb = float(dt), float(tr), float(alpha), int(Nb)
        self.Nt = int(round(tr / dt))
        self.dtp = self.dt / self.Nb
        self.accepted = False
        self.H0 = self.H1 = float("nan")
        self.flag = 0

    def get(self, name: str) -> np.ndarray:
        n = self.Ndof if name in ("v", "dSdx") else self.Ndim
        out = np.empty(n)
        self.model._call("elph_hmc_get", _IDS[name], ptr(out))
        return out

    def set_v(self, v):
        self.model._call("elph_hmc_set_v", ptr(_f64(v, self.Ndof, "v")))


def _use_p(P):
    return 0 if (P is None or getattr(P, "is_identity", False)) else 1


def refresh_v_(hmc: HybridMonteCarlo, model, fa, R):
    """``refresh_v!`` (src/HMC.jl:648-660); ``fa`` must have its mass matrix uploaded (update_M_)."""
    model._call("elph_hmc_refresh_v", hmc.alpha, ptr(_f64(R, model.Ndof, "R")))


def refresh_phi_(hmc: HybridMonteCarlo, model, R_plus, R_minus) -> float:
    """``refresh_ϕ!`` (src/HMC.jl:666-692)."""
    S = C.c_double()
    model._call("elph_hmc_refresh_phi", ptr(_f64(R_plus, model.Ndim, "R_plus")), ptr(_f64(R_minus, model.Ndim, "R_minus")), C.byref(S))
    return S.value


def calc_Oinv_(hmc: HybridMonteCarlo, model, P=None, power: float = 1.0, arnoldi_noise=None):
    """``calc_O⁻¹Λϕ!`` (src/HMC.jl:820-915) -> (iters, flag)."""
    it, fl = C.c_int64(), C.c_int32()
    an = None if arnoldi_noise is None else ptr(_f64(arnoldi_noise, 2 * model.Nsites, "arnoldi_noise"))
    model._call("elph_hmc_calc_Oinv", _use_p(P), an, float(power), C.byref(it), C.byref(fl))
    return int(it.value), int(fl.value)


def calc_H(hmc: HybridMonteCarlo, model, fa):
    """``calc_H`` (src/HMC.jl:698-705) -> (H, S, K)."""
    H, S, K = C.c_double(), C.c_double(), C.c_double()
    model._call("elph_hmc_calc_H", C.byref(H), C.byref(S), C.byref(K))
    return H.value, S.value, K.value


def calc_dSdx_(hmc: HybridMonteCarlo, model, fermion_only: bool = False) -> np.ndarray:
    """``fill!(dSdx,0); calc_dSdx!`` / ``calc_dSfdx!`` (src/HMC.jl:749-814)."""
    out = np.empty(model.Ndof)
    model._call("elph_hmc_calc_dSdx", 1 if fermion_only else 0, ptr(out))
    return out


def update_(model, hmc: HybridMonteCarlo, fa, P=None, *, R_v, R_plus, R_minus, arnoldi_noises=None, uniform: float):
    """``update!(model, hmc, fa, P)`` (src/HMC.jl:310-335) -> ``(accepted, iters)``."""
    acc, fl = C.c_int32(), C.c_int32()
    it, H0, H1 = C.c_double(), C.c_double(), C.c_double()
    an = None
    if _use_p(P):
        an = ptr(_f64(np.concatenate([np.asarray(a, dtype=np.float64) for a in arnoldi_noises]),
                      (hmc.Nt + 2) * 2 * model.Nsites, "arnoldi_noises"))
    model._call("elph_hmc_update", hmc.dt, hmc.Nt, hmc.Nb, hmc.alpha, ptr(_f64(R_v, model.Ndof, "R_v")),
                ptr(_f64(R_plus, model.Ndim, "R_plus")), ptr(_f64(R_minus, model.Ndim, "R_minus")), an, _use_p(P), float(uniform),
                C.byref(acc), C.byref(it), C.byref(H0), C.byref(H1), C.byref(fl))
    hmc.accepted, hmc.H0, hmc.H1, hmc.flag = bool(acc.value), H0.value, H1.value, int(fl.value)
    return hmc.accepted, it.value


class ReflectionUpdate:
    """``ReflectionUpdate(model, freq, nsites)`` (src/SpecialUpdates.jl:40-91): active for Holstein models with phonons."""

    def __init__(self, model, freq: int, nsites: int):
        from .models import HolsteinModel
        self.active = isinstance(model, HolsteinModel)        # :79-84 always on for Holstein, :86-90 off for every other model
        self.freq = int(freq)
        self.nsites = min(model.Nph, int(nsites)) if self.active else 0


class SwapUpdate:
    """``SwapUpdate(model, freq, nbonds)`` (src/SpecialUpdates.jl:169-226)."""

    def __init__(self, model, freq: int, nbonds: int):
        from .models import HolsteinModel, SSHModel
        if isinstance(model, HolsteinModel):                  # :193-206
            self.active = not (model.Nbonds == 0 and nbonds > 0)
        elif isinstance(model, SSHModel):                     # :208-222
            self.active = not (model.Nph == 0 and nbonds > 0)
        else:                                                 # :224-229
            self.active = False
        self.freq = int(freq)
        self.nbonds = min(model.Nbonds, int(nbonds)) if isinstance(model, (HolsteinModel, SSHModel)) else 0


def special_update_(model, hmc: HybridMonteCarlo, upd, P=None, *, targets, R_plus, R_minus, uniforms, arnoldi_noises=None):
    """``special_update!(model, hmc, update, P)`` (src/SpecialUpdates.jl:97-160, 233-366) with every random draw injected:
    ``targets`` = the sampled sites (reflection; 0-based phonon columns) or pairs ``(i, j)`` of phonon columns (swap:
    the two sites of each sampled bond for Holstein, the two sampled phonons for SSH), ``R_plus[k]``, ``R_minus[k]``,
    ``uniforms[k]`` (and ``arnoldi_noises[k]`` with a preconditioner) the draws of proposal k.  One device call per
    proposal.  Returns the acceptance ratio; ``hmc.special_log`` holds ``(accepted, S0, S1, iters, flag)`` per proposal."""
    from .models import HolsteinModel
    hmc.special_log = []
    reflect = isinstance(upd, ReflectionUpdate)
    if not upd.active or (reflect and not isinstance(model, HolsteinModel)):
        return 0.0
    n = upd.nsites if reflect else upd.nbonds      # the reference's bookkeeping: accepted / ru.nsites, accepted / su.nbonds
    if len(targets) != n:
        raise ValueError(f"special_update_: {len(targets)} targets for an update configured for {n}")
    accepted = 0.0
    for k, tgt in enumerate(targets):
        i, j = (int(tgt), 0) if reflect else (int(tgt[0]), int(tgt[1]))
        acc, fl, it = C.c_int32(), C.c_int32(), C.c_int64()
        S0, S1 = C.c_double(), C.c_double()
        an = None
        if _use_p(P):
            an = ptr(_f64(arnoldi_noises[k], 2 * model.Nsites, "arnoldi_noise"))
        model._call("elph_hmc_special_update", 0 if reflect else 1, i, j, ptr(_f64(R_plus[k], model.Ndim, "R_plus")),
                    ptr(_f64(R_minus[k], model.Ndim, "R_minus")), an, _use_p(P), float(uniforms[k]), C.byref(acc), C.byref(S0),
                    C.byref(S1), C.byref(it), C.byref(fl))
        hmc.special_log.append((bool(acc.value), S0.value, S1.value, int(it.value), int(fl.value)))
        accepted += float(acc.value)
    return accepted / n if n else 0.0
