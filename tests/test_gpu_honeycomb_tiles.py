"""Register-tile kernels for the honeycomb lattice 32 cells wide (config D: examples/holstein_hmc_honeycomb.toml scaled to 2048
sites): same results as the generic shared-memory kernels (tuning key 21 = 0) and as the oracle."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(0.8, 0.1), (2.0, 0.1)], ids=["L8", "L20"])
def pair(request):
    beta, dtau = request.param
    om, rng = oracle_holstein("honeycomb", 32, beta, dtau, mu=-0.4, omega=1.0, lam=1.0)
    em = engine_holstein_like(om)
    yield om, em, rng
    em.close()


def test_honeycomb_cg_matches_oracle_and_generic(pair):
    import elphdynamics_b200 as E
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng = pair
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    xo = np.zeros(om.Ndim)
    it_o, res_o, fl_o = ldiv(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter))
    xe = np.zeros(om.Ndim)
    it_e, res_e, fl_e = E.ldiv_(xe, em, b)
    assert fl_o == fl_e == 0 and abs(it_e - it_o) <= 2, (it_e, it_o)
    assert relerr(xe, xo) <= 50 * om.tol
    em._call("elph_set_tuning", 21, 0)
    xg = np.zeros(om.Ndim)
    it_g, _, fl_g = E.ldiv_(xg, em, b)
    em._call("elph_set_tuning", 21, 1)
    assert fl_g == 0 and abs(it_g - it_e) <= 1 and relerr(xg, xe) <= 1e-6


def test_honeycomb_batch_of_solves(pair):
    import elphdynamics_b200 as E
    om, em, rng = pair
    B = rng.normal(size=(3, om.Ndim))
    X = np.zeros_like(B)
    infos = E.ldiv_batch_(X, em, B)
    for k in range(3):
        x1 = np.zeros(om.Ndim)
        it, res, fl = E.ldiv_(x1, em, B[k])
        assert infos[k][2] == fl == 0 and abs(infos[k][0] - it) <= 1
        y = np.zeros(om.Ndim)
        om.mulMTM(y, X[k])
        assert relerr(y, B[k]) <= np.sqrt(om.tol)


def test_honeycomb_products(pair):
    import elphdynamics_b200 as E
    om, em, rng = pair
    v = rng.normal(size=om.Ndim)
    for name in ("mulMTM", "mulM", "mulMT"):
        yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
        getattr(om, name)(yo, v)
        getattr(E, name + "_")(ye, em, v)
        assert relerr(ye, yo) <= 1e-12, name
