"""Green's-function estimator: the device convolutions of setup!(estimator, n1, n2)
(src/GreensFunctions.jl:239-296, 361-414) against the NumPy restatement, and update! + setup! end to end."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr
from oracle import greens as og
from oracle.solvers import ConjugateGradient

pytestmark = pytest.mark.gpu

CASES = [
    ("square", 4, 0.5),        # Ltau = 5: 2L = 10 = 2 x 5
    ("honeycomb", 3, 0.7),     # two orbitals, odd extents, 2L = 14 = 2 x 7
    ("chain", 7, 1.1),         # one lattice axis, prime extent, 2L = 22
    ("triangular", 5, 0.3),    # 2L = 6
    ("square", 32, 0.4),       # 2L = 8, the extents of configs B / C
]


@pytest.mark.parametrize("geom,Ls,beta", CASES)
def test_setup_pair_matches_oracle(geom, Ls, beta):
    import elphdynamics_b200 as E
    from elphdynamics_b200 import greens as eg
    om, rng = oracle_holstein(geom, Ls, beta, 0.1, mu=-0.4)
    em = engine_holstein_like(om)
    nv = 3
    Go, Ge = og.EstimateGreensFunction(om, nv), eg.EstimateGreensFunction(em, nv)
    # arbitrary vectors in place of the solves: the convolutions are linear-algebra identities in R and M^-1 R
    Go.R[:] = rng.normal(size=Go.R.shape)
    Go.MinvR[:] = rng.normal(size=Go.R.shape)
    Ge.R[:], Ge.MinvR[:] = Go.R, Go.MinvR
    em._call("elph_greens_load", nv, E._lib.ptr(Ge.R), E._lib.ptr(Ge.MinvR))
    for n1, n2 in ((0, 1), (0, 2), (1, 2)):
        ref = og.setup(Go, n1, n2)
        got = eg.setup_pair_(Ge, n1, n2)
        for name, r, g in zip(("G_D0", "G_D0_G_D0", "G_DD_G_00", "G_D0_G_0D"), ref, got):
            assert g.shape == r.shape
            assert relerr(g, r) <= 1e-12, (name, n1, n2, relerr(g, r))
    # measure_...: the reference's indexing G[mod1(tau+1, 2L), o2, o1, l1+1, l2+1, l3+1]
    ns = Go.ns
    for (l1, l2, o1, o2, tau) in ((0, 0, 1, 1, 0), (1, 0, ns, 1, 3), (Go.L1 - 1, Go.L2 - 1, 1, ns, 2 * om.L - 1)):
        assert eg.measure(Ge.G_D0, Ge, l1, l2, 0, o1, o2, tau) == og.measure(np.asarray(got[0]), Go, l1, l2, 0, o1, o2, tau)
    em.close()


def test_update_then_setup_end_to_end():
    """update!(Gr, model) (batched solves on the device) followed by setup!(Gr, 1, 2): G[D,0] at D = 0 estimates
    1 - <n>, so its real part lies in (0, 1) up to stochastic noise; the arrays match the oracle's from its own solves."""
    from elphdynamics_b200 import greens as eg
    om, rng = oracle_holstein("square", 4, 1.0, 0.1, mu=-0.2, tol=1e-10)
    em = engine_holstein_like(om)
    nv = 2
    R = rng.normal(size=(nv, om.Ndim))
    Go, Ge = og.EstimateGreensFunction(om, nv), eg.EstimateGreensFunction(em, nv)
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    og.update(Go, om, cg, None, R)
    infos = eg.update_(Ge, em, None, R=R)
    assert all(f == 0 for (_, _, f) in infos)
    assert relerr(Ge.MinvR, Go.MinvR) <= 1e-7
    ref = og.setup(Go, 0, 1)
    got = eg.setup_pair_(Ge, 0, 1)
    for r, g in zip(ref, got):
        assert relerr(g, r) <= 1e-6
    em.close()
