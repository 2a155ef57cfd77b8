"""GPU parity tests (optical SSH model): CUDA path through the C ABI vs the NumPy oracle."""
import numpy as np
import pytest

from helpers import relerr
from helpers_ssh import engine_ssh_like, oracle_ssh

pytestmark = pytest.mark.gpu

CASES = [
    dict(Lside=4, beta=2.0, dtau=0.05),                       # examples/ssh_langevin_square.toml as shipped
    dict(Lside=3, beta=0.5, dtau=0.05, alpha2=0.02),          # odd L: 5 ragged colours; non-linear coupling
    dict(geometry="two_site", beta=1.0, dtau=0.1),            # examples/ssh_hmc_two_site.toml (one bond)
    dict(Lside=4, beta=0.5, dtau=0.05, names=["a", "a"]),     # equivalent fields (primary_field folding)
    dict(Lside=4, beta=0.5, dtau=0.05, mixed=True),           # a bond type without phonons
    dict(Lside=8, beta=1.0, dtau=0.05, mu=0.2),
    # Lx = 32 / 64: served by the register-tile kernel of ssh_square.cu (the cases above use the generic kernel)
    dict(Lside=32, beta=0.4, dtau=0.05, mu=0.1, alpha2=0.01),
    dict(Lside=32, beta=0.25, dtau=0.05, mixed=True),
    dict(Lside=64, beta=0.2, dtau=0.05, mu=-0.1),
]


def test_kernel_family(pair):
    import ctypes as C
    om, em, _ = pair
    sq = C.c_int32()
    em._call("elph_get_kernel_info", C.byref(sq), None)
    assert sq.value == (1 if om.lat.L1 in (32, 64) else 0)


@pytest.fixture(scope="module", params=CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def pair(request):
    om, rng = oracle_ssh(**request.param)
    em = engine_ssh_like(om)
    yield om, em, rng
    em.close()


def test_tables_and_update_model(pair):
    om, em, _ = pair
    assert np.array_equal(em.neighbor_table, om.neighbor_table)
    assert np.array_equal(em.checkerboard_perm, om.checkerboard_perm)
    assert np.array_equal(em.inv_checkerboard_perm, om.inv_checkerboard_perm)
    assert np.array_equal(em.phonon_to_bond, om.phonon_to_bond)
    assert np.array_equal(em.bond_to_phonon, om.bond_to_phonon)
    assert np.array_equal(em.primary_field, om.primary_field)
    c, s = em.cosh_sinh()
    assert relerr(c, om.cosht) <= 1e-14 and relerr(s, om.sinht) <= 1e-13


@pytest.mark.parametrize("chunk", [0, 1, 3])
def test_matvecs(pair, chunk):
    import elphdynamics_b200 as E
    om, em, rng = pair
    em._call("elph_set_chunk", chunk)
    v = rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    for fo, fe in ((om.mulM, E.mulM_), (om.mulMT, E.mulMT_), (om.mulMTM, E.mulMTM_)):
        fo(yo, v)
        fe(ye, em, v)
        assert relerr(ye, yo) <= 1e-12, fe.__name__
    em._call("elph_set_chunk", 0)


def test_muldMdx_and_action(pair):
    import elphdynamics_b200 as E
    from oracle.action import calc_dSbdx, calc_Sb
    om, em, rng = pair
    u, v = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    do, de = np.zeros(om.Ndof), np.zeros(om.Ndof)
    om.muldMdx(do, u, v)
    E.muldMdx_(de, u, em, v)
    assert relerr(de, do) <= 1e-9
    assert abs(E.calc_Sb(em) - calc_Sb(om)) <= 1e-12 * abs(calc_Sb(om))
    base = rng.normal(size=om.Ndof)
    a, b = base.copy(), base.copy()
    calc_dSbdx(a, om, True)
    E.calc_dSbdx_(b, em, True)
    assert relerr(b, a) <= 1e-13


def test_equivalent_field_check(pair):
    """update_model! raises when equivalent fields differ (src/SSHModels.jl:547-559)."""
    import elphdynamics_b200 as E
    from elphdynamics_b200._lib import ElphError
    om, em, rng = pair
    pf = om.primary_field
    nonprimary = np.nonzero(pf != np.arange(om.Ndof))[0]
    if nonprimary.size == 0:
        return
    x = om.x.copy()
    x[nonprimary[0]] += 0.5
    em.x = x
    with pytest.raises(ElphError):
        E.update_model_(em)
    em.x = om.x
    E.update_model_(em)


def test_solves_and_langevin(pair):
    import elphdynamics_b200 as E
    from oracle import langevin as olang
    from oracle.fourier import FourierAccelerator
    from oracle.kpm import KPMPreconditioner
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng = pair
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    xo, xe = np.zeros(om.Ndim), np.zeros(om.Ndim)
    it_o, _, f_o = ldiv(xo, om, b, cg)
    it_e, _, f_e = E.ldiv_(xe, em, b)
    assert f_o == f_e == 0 and abs(it_o - it_e) <= 2 and relerr(xe, xo) <= 50 * om.tol
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    info = E.setup_(Pe, noise)
    assert bool(info.active) == Po.active
    if Po.active:
        assert np.array_equal(Pe.orders(), Po.order)
        xo[:] = 0
        xe[:] = 0
        it_o, _, f_o = ldiv(xo, om, b, cg, Po)
        it_e, _, f_e = E.ldiv_(xe, em, b, Pe)
        assert f_o == f_e == 0 and abs(it_o - it_e) <= 2
    # Langevin RK step with injected noise; mass 0.1 as in examples/ssh_langevin_square.toml:84-87
    x0 = om.x.copy()
    cgt = ConjugateGradient(om.Ndim, tol=1e-10, maxiter=om.maxiter)
    em._call("elph_set_solver", 1e-10, 0, 0.0)
    fo = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
    fo.update_Q(0.0, 10.0, 0.1)
    fe = E.FourierAccelerator(em)
    E.update_Q_(fe, em, 0.0, 10.0, 0.1)
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    eta, g1, g2 = rng.normal(size=om.Ndof), rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    a1, a2 = rng.normal(size=2 * om.N), rng.normal(size=2 * om.N)
    it_o = olang.evolve_rk(om, cgt, fo, Po, 1e-3, eta, g1, g2, a1, a2)
    it_e = E.evolve_(em, E.RungeKuttaDynamics(em, 1e-3), fe, Pe, eta=eta, g1=g1, g2=g2, arnoldi1=a1, arnoldi2=a2)
    assert abs(it_o - it_e) <= 2
    assert relerr(em.x - x0, om.x - x0) <= 1e-8
    em._call("elph_set_solver", om.tol, 0, 0.0)
    om.x[:] = x0
    om.update_model()
    em.x = x0
    E.update_model_(em)
