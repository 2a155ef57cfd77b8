"""Batched solves (the n_v measurement vectors of update!(Gr, ...), src/GreensFunctions.jl:201-234):
elph_solve_batch must return, per right-hand side, what ldiv! returns for it alone."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr
from helpers_ssh import engine_ssh_like, oracle_ssh

pytestmark = pytest.mark.gpu

CASES = [
    ("holstein", dict(geom="square", Lside=4, beta=2.0, dtau=0.1)),            # generic persistent kernel, N = 16
    ("holstein", dict(geom="honeycomb", Lside=6, beta=1.0, dtau=0.1)),         # 3 colours
    ("holstein", dict(geom="triangular", Lside=5, beta=1.0, dtau=0.05)),       # ragged colours
    ("holstein", dict(geom="square", Lside=32, beta=1.2, dtau=0.1)),           # register-tile persistent kernel
    ("ssh", dict(Lside=32, beta=0.4, dtau=0.05, mu=0.1)),                      # SSH tables in shared memory
    ("ssh", dict(Lside=4, beta=1.0, dtau=0.05)),                               # no persistent kernel: sequential fallback
]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: c[0] + "-" + "-".join(str(v) for v in c[1].values()))
def pair(request):
    kind, kw = request.param
    if kind == "holstein":
        om, rng = oracle_holstein(**kw)
        em = engine_holstein_like(om)
    else:
        om, rng = oracle_ssh(**kw)
        em = engine_ssh_like(om)
    yield om, em, rng
    em.close()


@pytest.mark.parametrize("nrhs", [1, 2, 10])
def test_batch_equals_single_solves(pair, nrhs):
    import elphdynamics_b200 as E
    om, em, rng = pair
    B = rng.normal(size=(nrhs, om.Ndim))
    B[-1] *= 1e-3                                   # different norms: every right-hand side has its own stop rule
    X = np.full_like(B, 7.0)                        # output only: the initial guess is zero whatever X holds
    infos = E.ldiv_batch_(X, em, B)
    for k in range(nrhs):
        x1 = np.zeros(om.Ndim)
        it, res, fl = E.ldiv_(x1, em, B[k])
        assert infos[k][2] == fl == 0
        assert abs(infos[k][0] - it) <= 2           # identical algorithm; reduction order may differ between kernels
        # both are CG solutions at relative residual tol: they agree to ~tol * cond, not to rounding
        assert relerr(X[k], x1) <= 1e-3 and infos[k][1] <= np.sqrt(om.tol)
    # against the oracle for one column
    from oracle.solvers import ConjugateGradient, ldiv
    xo = np.zeros(om.Ndim)
    ito, reso, flo = ldiv(xo, om, B[0], ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter))
    assert flo == 0 and abs(infos[0][0] - ito) <= 2
    y = np.zeros(om.Ndim)
    om.mulMTM(y, X[0])
    assert relerr(y, B[0]) <= np.sqrt(om.tol)


def test_batch_with_preconditioner_is_the_loop_of_single_solves():
    import elphdynamics_b200 as E
    om, rng = oracle_holstein("square", 8, 4.0, 0.1, eps=0.3)
    em = engine_holstein_like(om)
    P = E.SymmetricKPMPreconditioner(em)
    E.setup_(P, rng.normal(size=2 * om.N))
    B = rng.normal(size=(3, om.Ndim))
    X = np.zeros_like(B)
    infos = E.ldiv_batch_(X, em, B, P)
    for k in range(3):
        x1 = np.zeros(om.Ndim)
        it, res, fl = E.ldiv_(x1, em, B[k], P)
        assert infos[k] == (it, res, fl)
        assert np.array_equal(X[k], x1)
    em.close()


def test_batch_argument_checks():
    import elphdynamics_b200 as E
    om, rng = oracle_holstein("square", 4, 1.0, 0.1)
    em = engine_holstein_like(om)
    with pytest.raises(ValueError):
        E.ldiv_batch_(np.zeros((2, om.Ndim)), em, np.zeros((2, 3)))
    with pytest.raises(E._lib.ElphError):
        em._call("elph_solve_batch", 0, None, None, 0, 1.0, None)
    em.close()


def test_update_Gr_solves(pair):
    """MinvR[k] = M^-1 R[k] (src/GreensFunctions.jl:201-234): check M MinvR = R and the per-vector oracle solve."""
    import elphdynamics_b200 as E
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng = pair
    R = rng.normal(size=(4, om.Ndim))
    MinvR = np.empty_like(R)
    infos = E.update_Gr_(MinvR, em, R)
    assert all(i[2] == 0 for i in infos)
    y = np.zeros(om.Ndim)
    for k in range(4):
        om.mulM(y, MinvR[k])
        assert relerr(y, R[k]) <= 50 * np.sqrt(om.tol)
    b = np.zeros(om.Ndim)
    om.mulMT(b, R[1])
    xo = np.zeros(om.Ndim)
    ito, _, flo = ldiv(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter))
    assert flo == 0 and abs(infos[1][0] - ito) <= 2 and relerr(MinvR[1], xo) <= 1e-3
