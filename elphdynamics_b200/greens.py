"""Host-side mirror of ``src/GreensFunctions.jl`` on top of libelph_b200.so: the stochastic Green's-function
estimator.  The solves of ``update!`` are one batched call (``elph_Minv_batch``), the convolutions of ``setup!`` run on
the device (csrc/greens.cu); the estimators that index the resulting arrays (``measure_GΔ0`` ...) stay with the driver.

Arrays keep the reference's memory order: Julia ``(2L, n, n, L1, L2, L3)`` is the C-ordered NumPy shape
``(L3, L2, L1, n, n, 2L)``.
"""
from __future__ import annotations

import numpy as np

from ._lib import ptr
from .models import AbstractModel, update_Gr_


class EstimateGreensFunction:
    """``EstimateGreensFunction(model, nv)`` (src/GreensFunctions.jl:23-188)."""

    def __init__(self, model: AbstractModel, nv: int = 2):
        self.model = model
        self.nv = max(2, int(nv))
        lat = model.lattice
        self.NL, self.L, self.N = model.Ndim, model.Ltau, model.Nsites
        self.L1, self.L2, self.L3, self.ns = lat.L1, lat.L2, lat.L3, lat.unit_cell.norbits
        self.R = np.zeros((self.nv, self.NL))
        self.MinvR = np.zeros((self.nv, self.NL))
        self.n1, self.n2 = 0, 1
        shape = (self.L3, self.L2, self.L1, self.ns, self.ns, 2 * self.L)
        self.G_D0 = np.zeros(shape, dtype=np.complex128)
        self.G_D0_G_D0 = np.zeros(shape, dtype=np.complex128)
        self.G_DD_G_00 = np.zeros(shape, dtype=np.complex128)
        self.G_D0_G_0D = np.zeros(shape, dtype=np.complex128)
        # the estimator's arrays live as long as the estimator and cross the bus for every pair: page-lock them once
        self._pinned = []
        for a in (self.R, self.MinvR, self.G_D0, self.G_D0_G_D0, self.G_DD_G_00, self.G_D0_G_0D):
            try:
                model.pin_host(a)
                self._pinned.append(a)
            except RuntimeError:
                break                # pinning is an optimisation only

    def close(self):
        """Release the page locks (call before the model is closed)."""
        for a in self._pinned:
            try:
                self.model.unpin_host(a)
            except Exception:
                pass
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def update_(Gr: EstimateGreensFunction, model, P=None, *, R):
    """``update!(estimator, model, P)`` (:201-234) with the random vectors injected (``R``: (nv, Ndim)); call
    ``setup_(P, noise)`` first when preconditioning, as the reference does at :206.  The vectors and the solutions are
    left on the device for ``setup_pair_``.  Returns the per-vector ``(iters, residual, flag)``."""
    Gr.R[:] = np.asarray(R, dtype=np.float64).reshape(Gr.nv, Gr.NL)
    infos = update_Gr_(Gr.MinvR, model, Gr.R, P)
    model._call("elph_greens_load", Gr.nv, ptr(Gr.R), ptr(Gr.MinvR))
    return infos


def setup_pair_(Gr: EstimateGreensFunction, n1: int, n2: int):
    """``setup!(estimator, n1, n2)`` (:239-296), 0-based vector indices: fills ``G_D0``, ``G_D0_G_D0``, ``G_DD_G_00``,
    ``G_D0_G_0D`` (they are overwritten, like the reference's fill! + accumulate)."""
    Gr.n1, Gr.n2 = int(n1), int(n2)
    outs = (Gr.G_D0, Gr.G_D0_G_D0, Gr.G_DD_G_00, Gr.G_D0_G_0D)
    Gr.model._call("elph_greens_setup", Gr.n1, Gr.n2, Gr.L1, Gr.L2, Gr.L3, Gr.ns, *[ptr(o.view(np.float64)) for o in outs])
    return outs


def measure(G: np.ndarray, Gr: EstimateGreensFunction, l1: int, l2: int, l3: int, o1: int, o2: int, tau: int) -> complex:
    """``measure_GΔ0(estimator, l1, l2, l3, o1, o2, tau)`` and its three siblings (:301-345):
    ``G[mod1(tau+1, 2L), o2, o1, l1+1, l2+1, l3+1]`` (orbitals 1-based as in the reference)."""
    return complex(G[l3, l2, l1, o1 - 1, o2 - 1, tau % (2 * Gr.L)])
