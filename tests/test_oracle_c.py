"""The C restatement (CPU baseline) agrees with the NumPy oracle."""
import numpy as np

from helpers import oracle_holstein, relerr
from oracle.cref import CRef
from oracle.solvers import ConjugateGradient, solve_cg


def test_c_restatement_matches_numpy_oracle():
    for geom, Ls in (("square", 4), ("triangular", 3), ("honeycomb", 3), ("square", 8)):
        om, rng = oracle_holstein(geom, Ls, 1.0, 0.1)
        c = CRef(om)
        v = rng.normal(size=om.Ndim)
        yo, yc = np.zeros(om.Ndim), np.zeros(om.Ndim)
        for fo, fc in ((om.mulM, c.mulM), (om.mulMT, c.mulMT), (om.mulMTM, c.mulMTM)):
            fo(yo, v)
            fc(yc, v)
            assert relerr(yc, yo) < 1e-13
        b = np.zeros(om.Ndim)
        om.mulMT(b, v)
        xo, xc = np.zeros(om.Ndim), np.zeros(om.Ndim)
        it_o = solve_cg(xo, om, b, ConjugateGradient(om.Ndim, tol=1e-6, maxiter=5000))
        it_c, _ = c.cg(xc, b, tol=1e-6, maxiter=5000)
        assert abs(it_o - it_c) <= 2 and relerr(xc, xo) < 1e-4


def test_replica_driver_runs():
    om, _ = oracle_holstein("square", 4, 1.0, 0.1)
    secs, threads = CRef(om).mulMTM_throughput(nrep=4, reps=3, nthreads=2)
    assert secs > 0 and threads >= 1
