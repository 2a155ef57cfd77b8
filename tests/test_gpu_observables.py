"""Row N1 of the coverage contract (north_star: "observables from fixed-seed runs within statistical error bars"): a Markov
chain of the shipped example (examples/holstein_langevin_square.toml: 4x4, beta = 2, Runge-Kutta Langevin, Fourier
acceleration, KPM-preconditioned CG) on the engine against the oracle's restatement of the same driver loop
(src/RunSimulation.jl:25-140).  (i) identical injected noise: the two trajectories agree step by step; (ii) independent
seeds: <x>, <x^2> and the on-site equal-time Green's function agree within their binned error bars."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr
from helpers_chain import EngineChain, Noise, OracleChain, binned_mean_and_error, run_chain

pytestmark = pytest.mark.gpu


def _pair(seed_field=1234):
    om, _ = oracle_holstein("square", 4, 2.0, 0.1, mu=-1.0, seed=seed_field, eps=0.3, tol=1e-5)
    em = engine_holstein_like(om)
    return om, em


def test_trajectories_agree_step_by_step_with_identical_noise():
    """200 Runge-Kutta steps, dt = 1e-3 as shipped.  The solves run at 1e-10 on both sides so that the comparison is one of the
    update arithmetic and not of two Krylov roundings at the shipped 1e-5."""
    om, em = _pair()
    oc, ec = OracleChain(om, 1e-3, tol=1e-10), EngineChain(em, 1e-3, tol=1e-10)
    try:
        noise = Noise(777, om.Ndof, om.Ndim, om.N)
        worst = 0.0
        for k in range(200):
            nz = noise.step()
            it_o = oc.step(nz)
            it_e = ec.step(nz)
            assert abs(it_e - it_o) <= 2, (k, it_e, it_o)
            err = relerr(ec.x, oc.x)
            worst = max(worst, err)
            assert err <= 1e-7, (k, err)
        print("largest relative deviation of the field over 200 steps:", worst)
    finally:
        ec.close()
        em.close()


def test_observables_agree_within_error_bars():
    """Independent seeds, the shipped solver tolerance, dt = 0.02 (larger than the shipped 1e-3 so that a chain of 1000 steps
    decorrelates; both sides integrate with the same dt, so the step-size bias is common).  3 sigma of the combined binned
    errors: the seeds are fixed, so the outcome is deterministic -- the bound is there to catch a bias, not a fluctuation."""
    om, em = _pair()
    oc, ec = OracleChain(om, 0.02), EngineChain(em, 0.02)
    try:
        so = run_chain(oc, Noise(202, om.Ndof, om.Ndim, om.N), burnin=150, nsteps=640, meas_freq=4)
        se = run_chain(ec, Noise(101, om.Ndof, om.Ndim, om.N), burnin=150, nsteps=640, meas_freq=4)
        mo, eo = binned_mean_and_error(so)
        me, ee = binned_mean_and_error(se)
        z = (me - mo) / np.sqrt(eo ** 2 + ee ** 2)
        print("oracle <x>, <x^2>, G(0,0):", mo, "+-", eo)
        print("engine <x>, <x^2>, G(0,0):", me, "+-", ee)
        print("z-scores:", z)
        assert np.all(np.isfinite(z)) and np.all(np.abs(z) <= 3.0), z
        assert 0.0 < me[2] < 1.0 and 0.0 < mo[2] < 1.0      # G(0,0) = 1 - n_sigma is a probability
    finally:
        ec.close()
        em.close()
