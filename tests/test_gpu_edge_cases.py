"""GPU edge cases: degenerate lattices and time extents, error behaviour of the C ABI."""
import ctypes as C

import numpy as np
import pytest

from helpers import engine_holstein_like, relerr, synthetic_field
from oracle import lattice as olat
from oracle.holstein import HolsteinModel as OracleHolstein

pytestmark = pytest.mark.gpu


def _pair(lat, bonds, t, beta, dtau, seed=0, **kw):
    om = OracleHolstein(lat, bonds, t, beta, dtau, omega=1.0, lam=1.0, mu=-0.4, **kw)
    rng = np.random.default_rng(seed)
    om.x[:] = synthetic_field(rng, om.N, om.L, beta, 1.0, 1.0, 0.3)
    om.update_model()
    return om, engine_holstein_like(om), rng


@pytest.mark.parametrize("beta,dtau", [(0.1, 0.1), (0.2, 0.1), (0.3, 0.1), (1.1, 0.1), (1.3, 0.1), (2.3, 0.1)],
                         ids=["L1", "L2", "L3", "L11-prime", "L13-prime", "L23-prime"])
def test_short_and_prime_time_extents(beta, dtau):
    """Ltau = 1 (M = I + B), 2, 3 and prime lengths (generic-radix FFT stage, antiperiodic wrap on every slice pair)."""
    import elphdynamics_b200 as E
    from oracle.fourier import TimeFreqFFT
    om, em, rng = _pair(olat.Lattice(2, 1, 4), olat.SQUARE_BONDS, 1.0, beta, dtau)
    v = rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    for fo, fe in ((om.mulM, E.mulM_), (om.mulMT, E.mulMT_), (om.mulMTM, E.mulMTM_)):
        fo(yo, v)
        fe(ye, em, v)
        assert relerr(ye, yo) <= 1e-12, fe.__name__
    nu = np.zeros(om.Ndim, dtype=np.complex128)
    E.tau_to_omega_(nu, E.TimeFreqFFT(em), v)
    assert relerr(nu, TimeFreqFFT(om.N, om.L).tau_to_omega(v)) <= 1e-13
    do, de = np.zeros(om.Ndof), np.zeros(om.Ndof)
    om.muldMdx(do, v, yo)
    E.muldMdx_(de, v, em, yo)
    assert relerr(de, do) <= 1e-9
    em.close()


def test_single_site_no_bonds():
    """examples/holstein_hmc_single_site.toml: N = 1, Nbonds = 0 (empty neighbour table, zero colour groups)."""
    import elphdynamics_b200 as E
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng = _pair(olat.Lattice(1, 1, 1), [], np.zeros(0), 2.0, 0.1)
    assert om.Nbonds == 0 and em.Nbonds == 0
    v = rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    for fo, fe in ((om.mulM, E.mulM_), (om.mulMT, E.mulMT_), (om.mulMTM, E.mulMTM_)):
        fo(yo, v)
        fe(ye, em, v)
        assert relerr(ye, yo) <= 1e-13
    b = np.zeros(om.Ndim)
    om.mulMT(b, v)
    xo, xe = np.zeros(om.Ndim), np.zeros(om.Ndim)
    it_o, _, f_o = ldiv(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter))
    it_e, _, f_e = E.ldiv_(xe, em, b)
    assert f_o == f_e == 0 and abs(it_o - it_e) <= 2 and relerr(xe, xo) <= 1e-4
    em.close()


def test_disordered_hoppings_fall_back_to_the_generic_kernel():
    """Per-bond hoppings (sigma_t != 0 in the reference) on a 32x32 lattice: the register/shuffle kernel requires a
    uniform (cosh, sinh) per colour, so the generic kernel must be selected -- and must agree with the oracle."""
    import elphdynamics_b200 as E
    lat = olat.Lattice(2, 1, 32)
    rng0 = np.random.default_rng(5)
    t = 1.0 + 0.1 * rng0.normal(size=2 * lat.nsites)
    om, em, rng = _pair(lat, olat.SQUARE_BONDS, t, 0.5, 0.1)
    sq = C.c_int32()
    em._call("elph_get_kernel_info", C.byref(sq), None)
    assert sq.value == 0
    v = rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    om.mulMTM(yo, v)
    E.mulMTM_(ye, em, v)
    assert relerr(ye, yo) <= 1e-12
    em.close()


def test_maxiter_flag_and_fallback_semantics():
    """ldiv!: hitting maxiter gives flag 1 and a zeroed x (src/Models.jl:156-166); with a preconditioner the failed
    attempt is retried without it at 10*maxiter (:129-133)."""
    import elphdynamics_b200 as E
    from oracle.kpm import KPMPreconditioner
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng = _pair(olat.Lattice(2, 1, 4), olat.SQUARE_BONDS, 1.0, 2.0, 0.1)
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    em._call("elph_set_solver", 1e-10, 3, 0.0)
    cg = ConjugateGradient(om.Ndim, tol=1e-10, maxiter=3)
    xo, xe = np.ones(om.Ndim) * 0, np.zeros(om.Ndim)
    it_o, r_o, f_o = ldiv(xo, om, b, cg)
    it_e, r_e, f_e = E.ldiv_(xe, em, b)
    assert (it_o, f_o) == (3, 1) and (it_e, f_e) == (3, 1) and not xe.any() and abs(r_e - r_o) <= 1e-9 * r_o
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    E.setup_(Pe, noise)
    xo[:] = 0
    it_o, r_o, f_o = ldiv(xo, om, b, cg, Po)            # PCG fails at 3 iterations -> CG at 30
    it_e, r_e, f_e = E.ldiv_(xe, em, b, Pe)
    assert f_o == f_e and abs(it_o - it_e) <= 2
    assert em.last_solve_info.used_fallback == 1
    em.close()


def test_error_reporting_through_the_abi():
    import elphdynamics_b200 as E
    from elphdynamics_b200._lib import Config, ElphError, check, load, ptr
    lib = load()
    cfg = Config()
    cfg.model, cfg.Ltau, cfg.Nsites, cfg.Nbonds, cfg.Nph, cfg.dtau = 0, 4, 2, 1, 2, 0.1
    h = C.c_void_p()
    st = lib.elph_create(C.byref(cfg), C.byref(h))          # neighbor_table / mu are NULL
    assert st != 0 and b"NULL" in lib.elph_last_error(None)
    nt = np.array([[5, 9]], dtype=np.int64)                 # site index out of range
    mu = np.zeros(2)
    cs = np.ones(1)
    cfg.neighbor_table, cfg.mu, cfg.cosht, cfg.sinht = ptr(nt, np.int64), ptr(mu), ptr(cs), ptr(cs)
    st = lib.elph_create(C.byref(cfg), C.byref(h))
    assert st != 0 and b"out of range" in lib.elph_last_error(None)
    assert lib.elph_mulM(None, None, None) != 0             # null handle is an error, not a crash
    lat = E.Lattice(E.UnitCell(2, 1), 4)
    m = E.HolsteinModel(lat, 1.0, 0.1)
    m.assign_t(1.0, 0, 0, (1, 0, 0))
    m.initialize_model_()
    with pytest.raises(ElphError):                          # preconditioned solve before setup!(P)
        m._call("elph_cg_solve", ptr(np.ones(m.Ndim)), ptr(np.zeros(m.Ndim)), 1, 0.0, 0, None, None)
    with pytest.raises(ElphError):                          # fourier acceleration without Q
        m._call("elph_fourier_accelerate", ptr(np.ones(m.Ndof)), ptr(np.zeros(m.Ndof)), 1.0, 0)
    with pytest.raises(ValueError):
        E.mulM_(np.zeros(m.Ndim), m, np.zeros(3))
    m.close()


def test_workload_builders_match_the_test_builders():
    """bench.py builds its models with elphdynamics_b200.workloads (no oracle); same seeds must give the same field
    and tables as the oracle-side builders the parity tests use."""
    from helpers import oracle_holstein, relerr
    from helpers_ssh import oracle_ssh
    from elphdynamics_b200 import workloads
    om, _ = oracle_holstein("triangular", 5, 1.0, 0.1, seed=7, eps=0.5)
    em, _ = workloads.holstein("triangular", 5, 1.0, 0.1, seed=7, eps=0.5)
    assert np.array_equal(em.x, om.x)
    assert np.array_equal(em.neighbor_table, om.neighbor_table)
    assert relerr(em.expnV, om.expnV.reshape(-1)) < 1e-14
    em.close()
    os_, _ = oracle_ssh(4, 1.0, 0.05, seed=11)
    es, _ = workloads.ssh_square(4, 1.0, 0.05, seed=11)
    assert np.array_equal(es.x, os_.x)
    es.close()


def test_handles_are_independent_across_host_threads():
    """One handle per caller thread at a time (include/elph_b200.h); different handles may run concurrently from different
    host threads (independent Markov chains on one GPU): the results equal the sequential ones bit for bit, for the
    persistent cooperative CG (32-wide lattice) and for a KPM-preconditioned Langevin step."""
    import threading
    import elphdynamics_b200 as E
    from elphdynamics_b200 import workloads

    def make(c):
        m, rng = workloads.holstein("square", 32, 0.8, 0.1, mu=-0.5, seed=77 + c, eps=0.3)
        fa = E.FourierAccelerator(m)
        E.update_Q_(fa, m, 0.0, 10.0, 1.0)
        P = E.SymmetricKPMPreconditioner(m)
        nz = dict(eta=rng.normal(size=m.Ndof), g1=rng.normal(size=m.Ndim), g2=rng.normal(size=m.Ndim),
                  arnoldi1=rng.normal(size=2 * m.Nsites), arnoldi2=rng.normal(size=2 * m.Nsites))
        return m, fa, P, nz, rng.normal(size=m.Ndim)

    def work(ch, out, c):
        m, fa, P, nz, b = ch
        x = np.zeros(m.Ndim)
        it, res, flag = E.ldiv_(x, m, b)                                        # persistent cooperative CG
        it2 = E.evolve_(m, E.RungeKuttaDynamics(m, 1e-3), fa, P, **nz)          # graphs, KPM set-up on host threads
        out[c] = (it, flag, x, it2, m.x)

    K = 4
    results = []
    for concurrent in (False, True):
        chains = [make(c) for c in range(K)]
        out = [None] * K
        if concurrent:
            ths = [threading.Thread(target=work, args=(chains[c], out, c)) for c in range(K)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        else:
            for c in range(K):
                work(chains[c], out, c)
        results.append(out)
        for ch in chains:
            ch[0].close()
    for a, b in zip(*results):
        assert a[0] == b[0] and a[1] == b[1] == 0 and a[3] == b[3]
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[4])
