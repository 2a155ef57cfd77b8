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


def test_c_langevin_pieces_match_numpy_oracle():
    """oracle/cfast.py: force, tau-averaged A / A^T / A^-1, preconditioner apply and a whole KPM-preconditioned Runge-Kutta
    step through the C loops against the pure NumPy oracle (identical injected noise)."""
    from oracle import cfast
    from oracle import langevin as olang
    from oracle.fourier import FourierAccelerator
    from oracle.kpm import KPMPreconditioner
    for geom, Ls, beta in (("square", 4, 2.0), ("honeycomb", 3, 1.0), ("square", 8, 1.5)):
        om, rng = oracle_holstein(geom, Ls, beta, 0.1, mu=-0.4, lam2=0.03)
        of, _ = oracle_holstein(geom, Ls, beta, 0.1, mu=-0.4, lam2=0.03)
        cfast.accelerate_model(of, native=False)
        u, v = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
        do, dc = np.zeros(om.Ndof), np.zeros(om.Ndof)
        om.muldMdx(do, u, v)
        of.muldMdx(dc, u, v)
        assert relerr(dc, do) < 1e-13
        Po, Pc = KPMPreconditioner(om), cfast.FastKPM(of, native=False)
        noise = rng.normal(size=2 * om.N)
        Po.setup(noise)
        Pc.setup(noise)
        assert Po.active == Pc.active and np.array_equal(Po.order, Pc.order)
        assert abs(Pc.e_min - Po.e_min) <= 1e-6 and abs(Pc.e_max - Po.e_max) <= 1e-6      # 20 Arnoldi steps amplify last-bit differences (the GPU tests use 1e-6 too)
        w = rng.normal(size=om.N)
        for tr in (False, True):
            assert relerr(Pc.mulA(w, tr), Po.mulA(w, tr)) < 1e-14
        assert relerr(Pc.ldivA(w), Po.ldivA(w)) < 1e-13
        zo, zc = np.zeros(om.Ndim), np.zeros(om.Ndim)
        Po.ldiv(zo, v)
        keep = (Pc.lam_lo, Pc.lam_hi, Pc.lam_avg, Pc.lam_mag, Pc.coeff)
        Pc.lam_lo, Pc.lam_hi, Pc.lam_avg, Pc.lam_mag, Pc.coeff = Po.lam_lo, Po.lam_hi, Po.lam_avg, Po.lam_mag, Po.coeff   # same window
        Pc.ldiv(zc, v)
        Pc.lam_lo, Pc.lam_hi, Pc.lam_avg, Pc.lam_mag, Pc.coeff = keep
        assert relerr(zc, zo) < 2e-10      # the recurrences of order ~70 amplify the last-bit differences of the fused C loops
        fo = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
        fo.update_Q(0.0, 10.0, 1.0)
        cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
        eta, g1, g2 = rng.normal(size=om.Ndof), rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
        a1, a2 = rng.normal(size=2 * om.N), rng.normal(size=2 * om.N)
        x0 = om.x.copy()
        it_o = olang.evolve_rk(om, cg, fo, Po, 1e-3, eta, g1, g2, a1, a2)
        it_c = olang.evolve_rk(of, cg, fo, Pc, 1e-3, eta, g1, g2, a1, a2)
        assert abs(it_o - it_c) <= 1
        assert relerr(of.x - x0, om.x - x0) < 1e-6       # two PCG solves at tol 1e-4 .. 1e-5: the step agrees to the solve tolerance
