"""A short Langevin Markov chain with every random draw injected, on the oracle and on the engine: the driver loop of
``run_simulation!`` (src/RunSimulation.jl:25-140) reduced to what a parity check needs -- updates with the shipped settings
(Runge-Kutta, Fourier acceleration, KPM-preconditioned CG; examples/holstein_langevin_square.toml) and, every
``meas_freq`` steps, <x>, <x^2> and the equal-time on-site Green's function G(0,0) of the stochastic estimator
(src/GreensFunctions.jl:201-296).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import numpy as np


class Noise:
    """The draws of one chain, in the reference's order per step: eta, then per force evaluation g and the Arnoldi start values
    (src/LangevinDynamics.jl:162-225, :350-384); measurements draw nv random vectors and their own Arnoldi values."""

    def __init__(self, seed, Ndof, Ndim, N):
        self.rng = np.random.default_rng(seed)
        self.Ndof, self.Ndim, self.N = Ndof, Ndim, N

    def step(self):
        r = self.rng
        return dict(eta=r.normal(size=self.Ndof), g1=r.normal(size=self.Ndim), g2=r.normal(size=self.Ndim),
                    arnoldi1=r.normal(size=2 * self.N), arnoldi2=r.normal(size=2 * self.N))

    def measurement(self, nv):
        return self.rng.normal(size=(nv, self.Ndim)), self.rng.normal(size=2 * self.N)


class OracleChain:
    def __init__(self, om, dt, tol=None, model_accel=None):
        from oracle.fourier import FourierAccelerator
        from oracle.kpm import KPMPreconditioner
        from oracle.solvers import ConjugateGradient
        self.om, self.dt = om, dt
        self.cg = ConjugateGradient(om.Ndim, tol=tol or om.tol, maxiter=om.maxiter)
        self.fa = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
        self.fa.update_Q(0.0, 10.0, 1.0)
        self.P = KPMPreconditioner(om)

    @property
    def x(self):
        return self.om.x

    def step(self, nz):
        from oracle import langevin as olang
        return olang.evolve_rk(self.om, self.cg, self.fa, self.P, self.dt, nz["eta"], nz["g1"], nz["g2"], nz["arnoldi1"], nz["arnoldi2"])

    def measure(self, R, arnoldi):
        from oracle import greens as og
        Gr = og.EstimateGreensFunction(self.om, R.shape[0])
        og.update(Gr, self.om, self.cg, self.P, R, arnoldi)
        og.setup(Gr, 0, 1)
        x = self.om.x
        return np.array([x.mean(), (x * x).mean(), og.measure(Gr.G_D0, Gr, 0, 0, 0, 1, 1, 0).real])


class EngineChain:
    def __init__(self, em, dt, tol=None):
        import elphdynamics_b200 as E
        self.E, self.em, self.dt = E, em, dt
        if tol:
            em._call("elph_set_solver", float(tol), 0, 0.0)
        self.fa = E.FourierAccelerator(em)
        E.update_Q_(self.fa, em, 0.0, 10.0, 1.0)
        self.P = E.SymmetricKPMPreconditioner(em)
        self.dyn = E.RungeKuttaDynamics(em, dt)
        self.Gr = None

    @property
    def x(self):
        return self.em.x

    def step(self, nz):
        return self.E.evolve_(self.em, self.dyn, self.fa, self.P, **nz)

    def measure(self, R, arnoldi):
        from elphdynamics_b200 import greens as eg
        if self.Gr is None:
            self.Gr = eg.EstimateGreensFunction(self.em, R.shape[0])
        self.E.setup_(self.P, arnoldi)
        eg.update_(self.Gr, self.em, self.P, R=R)
        eg.setup_pair_(self.Gr, 0, 1)
        x = self.em.x
        return np.array([x.mean(), (x * x).mean(), eg.measure(self.Gr.G_D0, self.Gr, 0, 0, 0, 1, 1, 0).real])

    def close(self):
        if self.Gr is not None:
            self.Gr.close()


def run_chain(chain, noise: Noise, burnin: int, nsteps: int, meas_freq: int, nv: int = 2):
    """Returns the measurement series, shape (nsteps // meas_freq, 3): <x>, <x^2>, G(0,0)."""
    for _ in range(burnin):
        chain.step(noise.step())
    out = []
    for k in range(1, nsteps + 1):
        chain.step(noise.step())
        if k % meas_freq == 0:
            out.append(chain.measure(*noise.measurement(nv)))
    return np.array(out)


def binned_mean_and_error(series, nbins=8):
    """Mean and standard error from bin averages (src/BinningAnalysis / the usual Monte Carlo practice)."""
    n = (series.shape[0] // nbins) * nbins
    bins = series[:n].reshape(nbins, -1, series.shape[1]).mean(axis=1)
    return bins.mean(axis=0), bins.std(axis=0, ddof=1) / np.sqrt(nbins)
