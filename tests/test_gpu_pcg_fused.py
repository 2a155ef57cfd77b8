"""KPM-preconditioned CG as one persistent kernel (csrc/pcg_fused.cu) against the reference recurrence
(src/IterativeSolvers.jl:153-234 with ldiv!(z, P, r) of src/KPMPreconditioners.jl:426-481) as restated by the oracle, and
against the launch-per-phase form of the same solve (tuning key 17 = 0): iteration counts within +-2, solutions to the
solver tolerance, identical stop-rule bookkeeping (maxiter, initial guess)."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr

pytestmark = pytest.mark.gpu

# (beta, dtau): Ltau = 12 (radix 4, 3), 7 (odd prime: middle frequency is its own mirror), 30 (2, 3, 5), 5, 64
CASES = [(1.2, 0.1), (0.7, 0.1), (3.0, 0.1), (0.5, 0.1), (6.4, 0.1)]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"L{int(round(c[0] / c[1]))}")
def setup(request):
    import elphdynamics_b200 as E
    from oracle.kpm import KPMPreconditioner, kpm_coefficients
    beta, dtau = request.param
    om, rng = oracle_holstein("square", 32, beta, dtau, mu=-0.8, omega=1.0, lam=1.2)
    em = engine_holstein_like(om)
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    info = E.setup_(Pe, noise)
    assert bool(info.active) == Po.active
    if Po.active:
        # identical spectral window on both sides (the 20-step Arnoldi estimates agree to ~1e-8 only)
        Po.lam_lo, Po.lam_hi = info.lambda_lo, info.lambda_hi
        Po.lam_avg, Po.lam_mag = (Po.lam_hi + Po.lam_lo) / 2, (Po.lam_hi - Po.lam_lo) / 2
        Po.coeff = [kpm_coefficients(int(Po.order[w]), Po.lam_lo, Po.lam_hi, Po.phis[w]) for w in range(Po.Lo2)]
    yield om, em, rng, Po, Pe
    em.close()


def _launches(em):
    return em.launch_count()


def test_fused_pcg_matches_oracle_and_unfused(setup):
    import elphdynamics_b200 as E
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng, Po, Pe = setup
    if not Po.active:
        pytest.skip("preconditioner inactive for this field")
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    xo = np.zeros(om.Ndim)
    it_o, res_o, fl_o = ldiv(xo, om, b, cg, Po)
    xe = np.zeros(om.Ndim)
    l0 = _launches(em)
    it_e, res_e, fl_e = E.ldiv_(xe, em, b, Pe)
    n_fused = _launches(em) - l0
    assert fl_o == fl_e == 0
    assert abs(it_e - it_o) <= 2, (it_e, it_o)
    assert relerr(xe, xo) <= 50 * om.tol
    assert res_e <= np.sqrt(om.tol)
    # the launch-per-phase form of the same solve
    em._call("elph_set_tuning", 17, 0)
    x2 = np.zeros(om.Ndim)
    l0 = _launches(em)
    it2, res2, fl2 = E.ldiv_(x2, em, b, Pe)
    n_unfused = _launches(em) - l0
    em._call("elph_set_tuning", 17, 1)
    assert fl2 == 0 and abs(it2 - it_e) <= 1, (it2, it_e)
    assert relerr(x2, xe) <= 1e-6
    # full-length tau-FFTs inside the fused kernel instead of the half-length form used for even Ltau
    em._call("elph_set_tuning", 18, 0)
    x3 = np.zeros(om.Ndim)
    it3, _, fl3 = E.ldiv_(x3, em, b, Pe)
    em._call("elph_set_tuning", 18, 1)
    assert fl3 == 0 and abs(it3 - it_e) <= 1 and relerr(x3, xe) <= 1e-6
    # the solve itself is ONE launch: M^T M + init + memsets before it, the true-residual check after it
    assert n_fused <= 8 and n_unfused >= n_fused + 3 * it2, (n_fused, n_unfused)


def test_fused_pcg_maxiter_and_initial_guess(setup):
    import ctypes as C
    import torch
    om, em, rng, Po, Pe = setup
    if not Po.active:
        pytest.skip("preconditioner inactive for this field")
    lib, h, n = em._lib, em.handle, om.Ndim
    g = rng.normal(size=n)
    b = np.zeros(n)
    om.mulMT(b, g)
    eng = lambda v: np.ascontiguousarray(v.reshape(om.N, om.L).T).reshape(-1)
    b_dev = torch.from_numpy(eng(b)).cuda()
    it, eps = C.c_int64(), C.c_double()
    out = {}
    for fused in (1, 0):
        em._call("elph_set_tuning", 17, fused)
        x_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
        em._call("elph_dev_cg_solve", b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, 3, C.byref(it), C.byref(eps))    # maxiter = 3
        em.synchronize()
        x3 = x_dev.cpu().numpy().copy()
        assert it.value == 3
        eps3 = eps.value
        # continue from that iterate: the stop rule restarts from the new initial residual
        em._call("elph_dev_cg_solve", b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, 0, C.byref(it), C.byref(eps))
        em.synchronize()
        out[fused] = (x3, eps3, it.value, eps.value, x_dev.cpu().numpy().copy())
        assert eps.value < om.tol
    em._call("elph_set_tuning", 17, 1)
    assert relerr(out[1][0], out[0][0]) <= 1e-10           # three iterations: same iterate to rounding
    assert abs(out[1][1] - out[0][1]) <= 1e-9 * out[0][1]
    assert abs(out[1][2] - out[0][2]) <= 1
    assert relerr(out[1][4], out[0][4]) <= 1e-6
    # against the oracle's recurrence, three iterations
    from oracle.solvers import ConjugateGradient, solve_pcg
    xo = np.zeros(n)
    solve_pcg(xo, om, b, ConjugateGradient(n, tol=om.tol, maxiter=om.maxiter), Po, maxiter=3)
    assert relerr(np.ascontiguousarray(out[1][0].reshape(om.L, om.N).T).reshape(-1), xo) <= 1e-9



@pytest.mark.parametrize("geom,Lside", [("square", 32), ("square", 4), ("honeycomb", 6), ("triangular", 5), ("square", 64)])
def test_device_arnoldi_equals_host_arnoldi(geom, Lside):
    """arnoldi_eigenvalue_bounds! (src/KPMPreconditioners.jl:845-942): the two Krylov runs as one kernel (tuning key 19, default)
    against the same loops on the host and against the oracle."""
    import elphdynamics_b200 as E
    from oracle.kpm import KPMPreconditioner
    om, rng = oracle_holstein(geom, Lside, 0.8, 0.1, mu=-0.5)
    em = engine_holstein_like(om)
    Pe = E.SymmetricKPMPreconditioner(em)
    Po = KPMPreconditioner(om)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    em._call("elph_set_tuning", 19, 0)
    ih = E.setup_(Pe, noise)
    em._call("elph_set_tuning", 19, 1)
    idv = E.setup_(Pe, noise)
    for a, b in ((ih.e_min, idv.e_min), (ih.e_max, idv.e_max)):
        assert abs(a - b) <= 1e-9 * abs(a), (a, b)
    assert abs(idv.e_min - Po.e_min) <= 1e-6 * abs(Po.e_min) and abs(idv.e_max - Po.e_max) <= 1e-6 * abs(Po.e_max)
    assert bool(idv.active) == Po.active
    if Po.active:
        assert np.array_equal(Pe.orders(), Po.order)
    em.close()


@pytest.mark.parametrize("beta,dtau", [(0.6, 0.05), (0.35, 0.05)])
def test_ssh_kpm_apply_register_chains(beta, dtau):
    """ldiv!(z, P, r) for the SSH model on a 32x32 lattice: the cluster-split chain kernel with the tau-averaged (cosh, sinh) of
    every bond in shared memory (src/KPMPreconditioners.jl:355-381, 606-679) against the oracle and against the generic
    shared-memory kernel (tuning key 1)."""
    import elphdynamics_b200 as E
    from helpers_ssh import engine_ssh_like, oracle_ssh
    from oracle.kpm import KPMPreconditioner, kpm_coefficients
    from oracle.solvers import ConjugateGradient, ldiv
    om, rng = oracle_ssh(Lside=32, beta=beta, dtau=dtau, mu=0.1)
    em = engine_ssh_like(om)
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    info = E.setup_(Pe, noise)
    assert bool(info.active) == Po.active
    if not Po.active:
        pytest.skip("preconditioner inactive for this field")
    Po.lam_lo, Po.lam_hi = info.lambda_lo, info.lambda_hi
    Po.lam_avg, Po.lam_mag = (Po.lam_hi + Po.lam_lo) / 2, (Po.lam_hi - Po.lam_lo) / 2
    Po.coeff = [kpm_coefficients(int(Po.order[w]), Po.lam_lo, Po.lam_hi, Po.phis[w]) for w in range(Po.Lo2)]
    r = rng.normal(size=om.Ndim)
    zo = np.zeros(om.Ndim)
    Po.ldiv(zo, r)
    l0 = em.launch_count()
    ze = np.zeros(om.Ndim)
    E.kpm_ldiv_(ze, Pe, r)
    assert relerr(ze, zo) <= 1e-11
    em._call("elph_set_tuning", 1, 1)       # generic kernels
    zg = np.zeros(om.Ndim)
    E.kpm_ldiv_(zg, Pe, r)
    em._call("elph_set_tuning", 1, 0)
    assert relerr(zg, zo) <= 1e-11 and relerr(ze, zg) <= 1e-12
    # and the preconditioned solve through it
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    xo, xe = np.zeros(om.Ndim), np.zeros(om.Ndim)
    it_o, _, f_o = ldiv(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter), Po)
    l0 = em.launch_count()
    it_e, _, f_e = E.ldiv_(xe, em, b, Pe)
    n_fused = em.launch_count() - l0
    assert f_o == f_e == 0 and abs(it_o - it_e) <= 2 and relerr(xe, xo) <= 50 * om.tol
    # the SSH solve is one persistent kernel as well (per-slice tables resident in shared memory); launch-per-phase form for comparison
    em._call("elph_set_tuning", 17, 0)
    x2 = np.zeros(om.Ndim)
    l0 = em.launch_count()
    it2, _, f2 = E.ldiv_(x2, em, b, Pe)
    n_unfused = em.launch_count() - l0
    em._call("elph_set_tuning", 17, 1)
    assert f2 == 0 and abs(it2 - it_e) <= 1 and relerr(x2, xe) <= 1e-6
    assert n_fused <= 8 and n_unfused >= n_fused + 3 * it2, (n_fused, n_unfused)
    em.close()
